// parser.cpp -- the `.fray` scene language.
//
// Behavioural mirror of DefaultSceneParser / ParsedBlockImpl in /root/reference/src/scene.cpp:61-753:
// line-oriented blocks `Class [name] {` ... `}`, `//` and `#` comments, `/*` ... `*/` only at line starts,
// randfloat()/randint() macro expansion, properties filled in the fixed class order
// Settings, Camera, Environment, Light, Geometry, Texture, Shader, Node, and shader-less nodes moved to
// Scene::superNodes. Error and warning texts go to stderr like the reference's; parseScene() returns false
// where the reference's main() would exit with -3 (src/main.cpp:503-506).
#include <cctype>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <sys/stat.h>

#include "scene.h"
#include "../csrc/rng.cuh"

namespace fray {

int g_verbose = 0;

static bool fileExists(const std::string& fn) // src/util.cpp:57-66
{
	std::string t = fn;
	if (!t.empty() && t.back() == '/') t.pop_back();
	struct stat st;
	return stat(t.c_str(), &st) == 0;
}

static std::string trim(const std::string& s)
{
	size_t a = 0, b = s.size();
	while (b > 0 && isspace((unsigned char) s[b - 1])) b--;
	while (a < b && isspace((unsigned char) s[a])) a++;
	return s.substr(a, b - a);
}

static std::vector<std::string> tokenize(const std::string& s) // src/util.cpp:69-82
{
	std::vector<std::string> out;
	size_t i = 0, n = s.size();
	while (i < n) {
		while (i < n && isspace((unsigned char) s[i])) i++;
		if (i >= n) break;
		size_t j = i;
		while (j < n && !isspace((unsigned char) s[j])) j++;
		out.push_back(s.substr(i, j - i));
		i = j;
	}
	return out;
}

static SyntaxError syntaxError(int line, const char* fmt, ...)
{
	char buf[160];
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(buf, sizeof(buf), fmt, ap);
	va_end(ap);
	return SyntaxError{ line, buf };
}

static std::string despace(std::string s) // '(', ')' and ',' are separators, src/scene.cpp:181-188, 655-666
{
	for (char& c: s)
		if (c == ',' || c == '(' || c == ')') c = ' ';
	return s;
}

static void get3Doubles(int line, const std::string& expr, double& a, double& b, double& c)
{
	if (3 != sscanf(despace(expr).c_str(), "%lf%lf%lf", &a, &b, &c)) throw syntaxError(line, "Expected three double values");
}

class Parser;

class Block: public ParsedBlock {
public:
	struct Line {
		int line;
		std::string name, value;
		bool recognized;
	};
	std::vector<Line> lines;
	int blockBegin = 0, blockEnd = 0;
	Parser* parser = nullptr;
	SceneElement* element = nullptr;

	Line* find(const char* name)
	{
		for (auto& l: lines)
			if (l.name == name) {
				l.recognized = true;
				return &l;
			}
		return nullptr;
	}

	bool getIntProp(const char* name, int* value, int lo, int hi) override
	{
		Line* l = find(name);
		if (!l) return false;
		int x;
		if (1 != sscanf(l->value.c_str(), "%d", &x)) throw syntaxError(l->line, "Invalid integer");
		if (x < lo || x > hi) throw syntaxError(l->line, "Value outside the allowed bounds (%d .. %d)\n", lo, hi);
		*value = x;
		return true;
	}
	bool getBoolProp(const char* name, bool* value) override
	{
		Line* l = find(name);
		if (!l) return false;
		*value = !(l->value == "off" || l->value == "false" || l->value == "0");
		return true;
	}
	bool getFloatProp(const char* name, float* value, float lo, float hi) override
	{
		Line* l = find(name);
		if (!l) return false;
		float x;
		if (1 != sscanf(l->value.c_str(), "%f", &x)) throw syntaxError(l->line, "Invalid float");
		if (x < lo || x > hi) throw syntaxError(l->line, "Value outside the allowed bounds (%f .. %f)\n", lo, hi);
		*value = x;
		return true;
	}
	bool getDoubleProp(const char* name, double* value, double lo, double hi) override
	{
		Line* l = find(name);
		if (!l) return false;
		double x;
		if (1 != sscanf(l->value.c_str(), "%lf", &x)) throw syntaxError(l->line, "Invalid double");
		if (x < lo || x > hi) throw syntaxError(l->line, "Value outside the allowed bounds (%f .. %f)\n", lo, hi);
		*value = x;
		return true;
	}
	bool getColorProp(const char* name, Color* value, float lo, float hi) override
	{
		Line* l = find(name);
		if (!l) return false;
		Color c;
		if (3 != sscanf(despace(l->value).c_str(), "%f%f%f", &c.r, &c.g, &c.b)) throw syntaxError(l->line, "Invalid color");
		const float comp[3] = { c.r, c.g, c.b };
		const char* names = "RGB";
		for (int i = 0; i < 3; i++)
			if (comp[i] < lo || comp[i] > hi)
				throw syntaxError(l->line, "Color %c value outside the allowed bounds (%f .. %f)\n", names[i], lo, hi);
		*value = c;
		return true;
	}
	bool getVectorProp(const char* name, Vec3* value) override
	{
		Line* l = find(name);
		if (!l) return false;
		Vec3 v;
		if (3 != sscanf(despace(l->value).c_str(), "%lf%lf%lf", &v.x, &v.y, &v.z)) throw syntaxError(l->line, "Invalid vector");
		*value = v;
		return true;
	}
	bool getGeometryProp(const char* name, Geometry** value) override;
	bool getShaderProp(const char* name, Shader** value) override;
	bool getTextureProp(const char* name, Texture** value) override;
	bool getNodeProp(const char* name, Node** value) override;
	bool getStringProp(const char* name, std::string* value) override
	{
		Line* l = find(name);
		if (!l) return false;
		*value = l->value;
		return true;
	}
	bool getFilenameProp(const char* name, std::string* value) override;
	bool getBitmapFileProp(const char* name, Bitmap& bmp) override;
	void getTransformProp(Transform& T) override // applied in file order, src/scene.cpp:297-320
	{
		for (auto& l: lines) {
			double x, y, z;
			if (l.name == "scale") {
				l.recognized = true;
				get3Doubles(l.line, l.value, x, y, z);
				T.scale(x, y, z);
			} else if (l.name == "rotate") {
				l.recognized = true;
				get3Doubles(l.line, l.value, x, y, z);
				T.rotate(x, y, z);
			} else if (l.name == "translate") {
				l.recognized = true;
				get3Doubles(l.line, l.value, x, y, z);
				T.translate(Vec3(x, y, z));
			}
		}
	}
	void requiredProp(const char* name) override
	{
		if (!find(name)) throw syntaxError(blockEnd, "Required property `%s' not defined", name);
	}
	void signalError(const char* msg) override { throw syntaxError(blockEnd, "%s", msg); }
	void signalWarning(const char* msg) override { fprintf(stderr, "Warning (at line %d): %s\n", blockEnd, msg); }
	int getBlockLines() override { return (int) lines.size(); }
	void getBlockLine(int idx, int& srcLine, std::string& head, std::string& tail) override
	{
		lines[idx].recognized = true;
		srcLine = lines[idx].line;
		head = lines[idx].name;
		tail = lines[idx].value;
	}
	SceneParser& getParser() override;
};

class Parser: public SceneParser {
public:
	std::string rootDir;
	Scene* s = nullptr;
	Rng macroRng; // the reference expands the macros with generator 0 (src/scene.cpp:405)

	template <class T> static T* byName(const std::vector<T*>& v, const char* name)
	{
		for (T* e: v)
			if (e->name == name) return e;
		return nullptr;
	}
	Shader* findShaderByName(const char* name) override { return byName(s->shaders, name); }
	Texture* findTextureByName(const char* name) override { return byName(s->textures, name); }
	Geometry* findGeometryByName(const char* name) override { return byName(s->geometries, name); }
	Node* findNodeByName(const char* name) override { return byName(s->nodes, name); }
	bool resolveFullPath(std::string& path) override // src/scene.cpp:710-721
	{
		std::string full = rootDir + path;
		if (!fileExists(full)) return false;
		path = full;
		return true;
	}

	SceneElement* newSceneElement(const std::string& cls); // src/scene.cpp:821-848
	void expandMacros(int srcLine, std::string& line);     // src/scene.cpp:609-653
	bool parse(const char* filename, Scene* scene);
};

bool Block::getGeometryProp(const char* name, Geometry** value)
{
	Line* l = find(name);
	if (!l) return false;
	Geometry* g = parser->findGeometryByName(l->value.c_str());
	if (!g) throw syntaxError(l->line, "Geometry not defined");
	*value = g;
	return true;
}
bool Block::getShaderProp(const char* name, Shader** value)
{
	Line* l = find(name);
	if (!l) return false;
	Shader* sh = parser->findShaderByName(l->value.c_str());
	if (!sh) throw syntaxError(l->line, "Shader not defined");
	*value = sh;
	return true;
}
bool Block::getTextureProp(const char* name, Texture** value)
{
	Line* l = find(name);
	if (!l) return false;
	Texture* t = parser->findTextureByName(l->value.c_str());
	if (!t) throw syntaxError(l->line, "Texture not defined");
	*value = t;
	return true;
}
bool Block::getNodeProp(const char* name, Node** value)
{
	Line* l = find(name);
	if (!l) return false;
	Node* n = parser->findNodeByName(l->value.c_str());
	if (!n) throw syntaxError(l->line, "Node not defined");
	*value = n;
	return true;
}
bool Block::getFilenameProp(const char* name, std::string* value)
{
	Line* l = find(name);
	if (!l) return false;
	*value = l->value;
	if (parser->resolveFullPath(*value)) return true;
	throw FileNotFoundError{ l->line, l->value };
}
bool Block::getBitmapFileProp(const char* name, Bitmap& bmp)
{
	Line* l = find(name);
	if (!l) return false;
	std::string fn = l->value;
	if (!parser->resolveFullPath(fn)) throw FileNotFoundError{ l->line, fn };
	return bmp.loadImage(fn.c_str());
}
SceneParser& Block::getParser() { return *parser; }

SceneElement* Parser::newSceneElement(const std::string& c)
{
	if (c == "GlobalSettings") return &s->settings;
	if (c == "Plane") return new Plane;
	if (c == "Sphere") return new Sphere;
	if (c == "Cube") return new Cube;
	if (c == "CsgPlus") return new CsgOp(FRAY_GEOM_CSG_PLUS);
	if (c == "CsgAnd") return new CsgOp(FRAY_GEOM_CSG_AND);
	if (c == "CsgMinus") return new CsgOp(FRAY_GEOM_CSG_MINUS);
	if (c == "Lambert") return new Lambert;
	if (c == "Phong") return new Phong;
	if (c == "CheckerTexture") return new CheckerTexture;
	if (c == "BitmapTexture") return new BitmapTexture;
	if (c == "Refl") return new Reflection;
	if (c == "Refr") return new Refraction;
	if (c == "Layered") return new Layered;
	if (c == "Fresnel") return new FresnelTexture;
	if (c == "Node") return new Node;
	if (c == "CubemapEnvironment") return new CubemapEnvironment;
	if (c == "Camera") return new Camera;
	if (c == "Mesh") return new Mesh;
	if (c == "BumpTexture") return new BumpTexture;
	if (c == "Const") return new ConstantShader;
	if (c == "PointLight") return new PointLight;
	if (c == "RectLight") return new RectLight;
	return nullptr;
}

void Parser::expandMacros(int srcLine, std::string& line)
{
	// `randfloat(a, b)` / `randint(a, b)` anywhere in the line are overwritten in place by a number and padded
	// with blanks up to the closing parenthesis.
	for (int pass = 0; pass < 2; pass++) {
		const char* kw = pass == 0 ? "randfloat" : "randint";
		size_t p;
		while ((p = line.find(kw)) != std::string::npos) {
			size_t open = line.find('(', p);
			if (open == std::string::npos) throw syntaxError(srcLine, "%s in inexpected format", kw);
			size_t close = line.find(')', open);
			if (close == std::string::npos) throw syntaxError(srcLine, "%s in inexpected format", kw);
			std::string args = line.substr(open + 1, close - open - 1);
			char text[32];
			if (pass == 0) {
				float lo, hi;
				if (2 != sscanf(args.c_str(), "%f,%f", &lo, &hi)) throw syntaxError(srcLine, "bad randfloat format (expected: randfloat(<min>, <max>))");
				if (lo > hi) throw syntaxError(srcLine, "bad randfloat format (min > max)");
				snprintf(text, sizeof(text), "%.5f", macroRng.randfloat() * (hi - lo) + lo);
			} else {
				int lo, hi;
				if (2 != sscanf(args.c_str(), "%d,%d", &lo, &hi)) throw syntaxError(srcLine, "bad randint format (expected: randint(<min>, <max>))");
				if (lo > hi) throw syntaxError(srcLine, "bad randint format (min > max)");
				snprintf(text, sizeof(text), "%d", macroRng.randint(lo, hi));
			}
			size_t span = close - p + 1, len = strlen(text);
			if (len >= span) throw syntaxError(srcLine, "%s expansion does not fit", kw);
			line.replace(p, span, std::string(text) + std::string(span - len, ' '));
		}
	}
}

bool Parser::parse(const char* filename, Scene* scene)
{
	s = scene;
	s->environment = nullptr;
	macroRng.init(42, 0xFFFFFFFFu, 0xFFFFFFFFu, 0);
	std::ifstream in(filename);
	if (!in) {
		fprintf(stderr, "Cannot open scene file `%s'!\n", filename);
		s->lastError = std::string("Cannot open scene file ") + filename;
		return false;
	}
	std::string fn = filename;
	size_t slash = fn.find_last_of("/\\");
	rootDir = slash == std::string::npos ? "" : fn.substr(0, slash + 1);

	std::vector<std::unique_ptr<Block>> blocks;
	Block* cur = nullptr;
	SceneElement* curObj = nullptr;
	bool commentedOut = false;
	int curLine = 0;
	std::string raw;
	auto fail = [&](const char* fmt, auto... args) {
		char buf[256];
		snprintf(buf, sizeof(buf), fmt, args...);
		fprintf(stderr, "%s\n", buf);
		s->lastError = buf;
		return false;
	};
	try {
		while (std::getline(in, raw)) {
			curLine++;
			if (commentedOut) {
				if (raw.size() >= 2 && raw[0] == '*' && raw[1] == '/') commentedOut = false;
				continue;
			}
			size_t c1 = raw.find("//"), c2 = raw.find('#');
			size_t cut = std::min(c1, c2);
			if (cut != std::string::npos) raw.erase(cut);
			std::string line = trim(raw);
			if (line.empty()) continue;
			if (line[0] == '/' && line.size() > 1 && line[1] == '*') {
				commentedOut = true;
				continue;
			}
			expandMacros(curLine, line);
			std::vector<std::string> tok = tokenize(line);
			if (tok.empty()) continue;
			if (!curObj) {
				if (tok.size() == 1) {
					if (tok[0] == "{") return fail("Excess `}' on line %d", curLine);
					return fail("Unexpected token `%s' on line %d", tok[0].c_str(), curLine);
				}
				if (tok.size() > 3) return fail("Unexpected content on line %d!", curLine);
				if (tok.back() != "{") return fail("A object definition should end with a `{' (on line %d)", curLine);
				curObj = newSceneElement(tok[0]);
				if (!curObj) return fail("Unknown object class `%s' on line %d", tok[0].c_str(), curLine);
				curObj->name = tok[1]; // "{" for anonymous blocks, exactly like src/scene.cpp:487
				blocks.emplace_back(new Block);
				cur = blocks.back().get();
				cur->parser = this;
				cur->element = curObj;
				cur->blockBegin = curLine;
				switch (curObj->getElementType()) {
					case ELEM_GEOMETRY: s->geometries.push_back((Geometry*) curObj); break;
					case ELEM_SHADER: s->shaders.push_back((Shader*) curObj); break;
					case ELEM_TEXTURE: s->textures.push_back((Texture*) curObj); break;
					case ELEM_NODE: s->nodes.push_back((Node*) curObj); break;
					case ELEM_ENVIRONMENT: delete s->environment; s->environment = (Environment*) curObj; break;
					case ELEM_CAMERA:
						delete s->camera;
						s->camera = (Camera*) curObj;
						s->camera->owner = s;
						break;
					case ELEM_LIGHT: s->lights.push_back((Light*) curObj); break;
					default: break;
				}
			} else if (tok.size() == 1) {
				if (tok[0] != "}") return fail("Unexpected token in object definition on line %d: `%s'", curLine, tok[0].c_str());
				cur->blockEnd = curLine;
				curObj = nullptr;
				cur = nullptr;
			} else {
				size_t i = tok[0].size();
				while (i < line.size() && isspace((unsigned char) line[i])) i++;
				size_t last = line.size() - 1;
				std::string value = line.substr(i);
				if (i < last && line[i] == '"' && line[last] == '"') value = line.substr(i + 1, last - i - 1);
				cur->lines.push_back(Block::Line{ curLine, tok[0], value, false });
			}
		}
		if (curObj) return fail("Unfinished object definition at EOF!");

		static const ElementType order[] = { ELEM_SETTINGS, ELEM_CAMERA, ELEM_ENVIRONMENT, ELEM_LIGHT, ELEM_GEOMETRY, ELEM_TEXTURE, ELEM_SHADER, ELEM_NODE };
		for (ElementType et: order)
			for (auto& pb: blocks) {
				if (pb->element->getElementType() != et) continue;
				pb->element->fillProperties(*pb);
				for (auto& l: pb->lines)
					if (!l.recognized && g_verbose)
						fprintf(stderr, "%s:%d: Warning: the property `%s' isn't recognized!\n", filename, l.line, l.name.c_str());
			}
	} catch (const SyntaxError& err) {
		return fail("%s:%d: Syntax error on line %d: %s", filename, err.line, err.line, err.msg.c_str());
	} catch (const FileNotFoundError& err) {
		return fail("%s:%d: Required file not found (%s) (required at line %d)", filename, err.line, err.filename.c_str(), err.line);
	}
	if (!s->camera) return fail("%s: the scene defines no Camera", filename);
	// nodes without a shader are not renderable objects (CSG operands etc.), src/scene.cpp:563-568
	for (int i = (int) s->nodes.size() - 1; i >= 0; i--)
		if (!s->nodes[i]->shader) {
			s->superNodes.push_back(s->nodes[i]);
			s->nodes.erase(s->nodes.begin() + i);
		}
	return true;
}

Scene::~Scene()
{
	for (auto* e: geometries) delete e;
	for (auto* e: nodes) delete e;
	for (auto* e: superNodes) delete e;
	for (auto* e: textures) delete e;
	for (auto* e: shaders) delete e;
	for (auto* e: lights) delete e;
	delete environment;
	delete camera;
}

bool Scene::parseScene(const char* sceneFile)
{
	Parser p;
	return p.parse(sceneFile, this);
}

void Scene::beginRender()
{
	for (auto* e: geometries) e->beginRender();
	for (auto* e: textures) e->beginRender();
	for (auto* e: shaders) e->beginRender();
	for (auto* e: superNodes) e->beginRender();
	for (auto* e: nodes) e->beginRender();
	for (auto* e: lights) e->beginRender();
	camera->beginRender();
	settings.beginRender();
	if (environment) environment->beginRender();
}

void Scene::beginFrame()
{
	for (auto* e: geometries) e->beginFrame();
	for (auto* e: textures) e->beginFrame();
	for (auto* e: shaders) e->beginFrame();
	for (auto* e: superNodes) e->beginFrame();
	for (auto* e: nodes) e->beginFrame();
	for (auto* e: lights) e->beginFrame();
	camera->beginFrame();
	settings.beginFrame();
	if (environment) environment->beginFrame();
}

int Scene::samplesPerPixel() const
{
	int spp = settings.wantAA ? 5 : 1; // COUNT_OF(offsets), src/main.cpp:55-61
	if (camera && camera->dof) spp = std::max(spp, camera->numDOFSamples);
	if (settings.gi) spp = std::max(spp, settings.numPaths);
	return spp;
}

} // namespace fray
