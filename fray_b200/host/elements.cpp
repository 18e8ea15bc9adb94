// elements.cpp -- property parsing and per-render / per-frame preparation of the scene element classes.
//
// Property names, defaults, ranges and derived quantities follow the reference:
//   geometry   /root/reference/src/geometry.h:60-127      lights       src/lights.h:49-88, src/lights.cpp:37-46
//   shaders    src/shading.h:139-255, src/shading.cpp:303-355   camera  src/camera.h:56-75, src/camera.cpp:34-57
//   textures   src/shading.h:42-107,227-237               settings     src/scene.cpp:799-814
//   cubemap    src/environment.cpp:31-52
#include <cstdio>
#include <cstring>
#include <sys/stat.h>

#include "scene.h"

namespace fray {

void Plane::fillProperties(ParsedBlock& pb)
{
	pb.getDoubleProp("y", &height);
	pb.getDoubleProp("limit", &limit);
}

void Sphere::fillProperties(ParsedBlock& pb)
{
	pb.getVectorProp("O", &O);
	pb.getDoubleProp("R", &R);
}

void Cube::fillProperties(ParsedBlock& pb)
{
	pb.getVectorProp("O", &O);
	pb.getDoubleProp("halfSide", &halfSide);
}

void CsgOp::fillProperties(ParsedBlock& pb)
{
	pb.requiredProp("left");
	pb.requiredProp("right");
	pb.getGeometryProp("left", &left);
	pb.getGeometryProp("right", &right);
}

void Mesh::fillProperties(ParsedBlock& pb)
{
	std::string fn;
	if (pb.getFilenameProp("file", &fn)) {
		if (!loadFromOBJ(fn.c_str())) pb.signalError("Could not parse OBJ file!");
	} else {
		pb.requiredProp("file");
	}
	pb.getBoolProp("faceted", &faceted);
	pb.getBoolProp("backfaceCulling", &backfaceCulling);
	pb.getBoolProp("useKDTree", &useKD);
}

void CheckerTexture::fillProperties(ParsedBlock& pb)
{
	pb.getColorProp("color1", &color1);
	pb.getColorProp("color2", &color2);
	pb.getDoubleProp("scaling", &scaling);
}

void BitmapTexture::fillProperties(ParsedBlock& pb)
{
	pb.getDoubleProp("scaling", &scaling);
	scaling = 1 / scaling; // stored inverted, src/shading.h:66-67
	if (!pb.getBitmapFileProp("file", bmp)) pb.requiredProp("file");
}

void BumpTexture::fillProperties(ParsedBlock& pb)
{
	pb.getDoubleProp("strength", &bumpIntensity);
	pb.getDoubleProp("scaling", &scaling); // NOT inverted, src/shading.h:91-92
	if (!pb.getBitmapFileProp("file", bumpTex)) pb.requiredProp("file");
}

void BumpTexture::beginRender() { bumpTex.differentiate(); }

void FresnelTexture::fillProperties(ParsedBlock& pb) { pb.getDoubleProp("ior", &ior, 1e-6, 10); }

void Lambert::fillProperties(ParsedBlock& pb)
{
	pb.getColorProp("color", &color);
	pb.getTextureProp("texture", &diffuseTex);
}

void Phong::fillProperties(ParsedBlock& pb)
{
	pb.getColorProp("color", &color);
	pb.getTextureProp("texture", &diffuseTex);
	pb.getDoubleProp("specularExponent", &exponent);
	pb.getDoubleProp("specularMultiplier", &specularMultiplier);
	pb.getColorProp("specularColor", &specularColor);
}

void Reflection::fillProperties(ParsedBlock& pb)
{
	double m = 1;
	pb.getDoubleProp("multiplier", &m);
	mult = Color((float) m, (float) m, (float) m);
	pb.getDoubleProp("glossiness", &glossiness, 0, 1);
	pb.getIntProp("numSamples", &numSamples, 1);
}

void Reflection::beginFrame()
{
	pureReflection = (glossiness == 1.0);
	deflectionScaling = std::pow(10.0, 2 - 4 * glossiness);
}

void Refraction::fillProperties(ParsedBlock& pb)
{
	double m = 1;
	pb.getDoubleProp("multiplier", &m);
	mult = Color((float) m, (float) m, (float) m);
	pb.getDoubleProp("ior", &ior, 1e-6, 10);
}

void Layered::addLayer(Shader* shader, Color opacity, Texture* texture)
{
	if (layers.size() < 32) layers.push_back(Layer{ shader, opacity, texture });
}

// first / last blank-separated token of `s`, removed from it (getFrontToken / getLastToken, src/scene.cpp:668-695)
static bool popFront(std::string& s, std::string& tok)
{
	size_t i = 0, n = s.size();
	while (i < n && isspace((unsigned char) s[i])) i++;
	if (i == n) return false;
	size_t j = i;
	while (j < n && !isspace((unsigned char) s[j])) j++;
	if (j == n) return false; // nothing may follow: the reference treats that as an error
	tok = s.substr(i, j - i);
	s = s.substr(j);
	return true;
}
static bool popBack(std::string& s, std::string& tok)
{
	int i = (int) s.size() - 1;
	while (i >= 0 && isspace((unsigned char) s[i])) i--;
	if (i < 0) return false;
	int j = i;
	while (j >= 0 && !isspace((unsigned char) s[j])) j--;
	if (j < 0) return false;
	tok = s.substr(j + 1, i - j);
	s = s.substr(0, j + 1);
	return true;
}
static std::string stripPunct(const std::string& s) // src/scene.cpp:697-708
{
	std::string r;
	for (char c: s)
		if (!isspace((unsigned char) c) && c != ',') r += c;
	return r;
}

void Layered::fillProperties(ParsedBlock& pb)
{
	// lines of the form: layer <shader> (r, g, b) [<texture> | NULL]     (src/shading.cpp:313-355)
	for (int i = 0; i < pb.getBlockLines(); i++) {
		int srcLine;
		std::string head, value;
		pb.getBlockLine(i, srcLine, head, value);
		if (head != "layer") continue;
		std::string shaderName, textureName;
		bool err = !popFront(value, shaderName);
		if (!err) shaderName = stripPunct(shaderName);
		if (value.empty()) err = true;
		if (!err && value.back() != ')') {
			err = !popBack(value, textureName);
			if (!err) textureName = stripPunct(textureName);
		}
		if (!err && textureName == "NULL") textureName.clear();
		Shader* shader = nullptr;
		Texture* texture = nullptr;
		if (!err) err = !(shader = pb.getParser().findShaderByName(shaderName.c_str()));
		if (!err && !textureName.empty()) err = !(texture = pb.getParser().findTextureByName(textureName.c_str()));
		if (err) throw SyntaxError{ srcLine, "Expected a line like `layer <shader>, <color>[, <texture>]'" };
		double x, y, z;
		std::string nums = value;
		for (char& c: nums)
			if (c == '(' || c == ')' || c == ',') c = ' ';
		if (3 != sscanf(nums.c_str(), "%lf%lf%lf", &x, &y, &z)) throw SyntaxError{ srcLine, "Expected three double values" };
		addLayer(shader, Color((float) x, (float) y, (float) z), texture);
	}
}

void Node::fillProperties(ParsedBlock& pb)
{
	pb.getGeometryProp("geometry", &geometry);
	pb.getShaderProp("shader", &shader);
	pb.getTransformProp(T);
	pb.getTextureProp("bump", &bump);
}

void Light::fillProperties(ParsedBlock& pb)
{
	pb.getColorProp("color", &color);
	pb.getFloatProp("power", &power);
}

void PointLight::fillProperties(ParsedBlock& pb)
{
	Light::fillProperties(pb);
	pb.getVectorProp("pos", &pos);
}

void RectLight::fillProperties(ParsedBlock& pb)
{
	Light::fillProperties(pb);
	pb.getIntProp("xSubd", &xSubd, 1);
	pb.getIntProp("ySubd", &ySubd, 1);
	pb.getTransformProp(T);
}

void RectLight::beginFrame()
{
	center = T.point(Vec3(0, 0, 0));
	Vec3 a = T.point(Vec3(-0.5, 0.0, -0.5));
	Vec3 b = T.point(Vec3(0.5, 0.0, -0.5));
	Vec3 c = T.point(Vec3(0.5, 0.0, 0.5));
	float width = (float) (b - a).length();
	float height = (float) (b - c).length();
	area = width * height; // float product widened to double, src/lights.cpp:43-45
}

void Camera::fillProperties(ParsedBlock& pb)
{
	if (!pb.getVectorProp("position", &pos)) pb.requiredProp("position");
	pb.getDoubleProp("aspectRatio", &aspectRatio, 1e-6);
	pb.getDoubleProp("fov", &fov, 0.0001, 179);
	pb.getDoubleProp("yaw", &yaw);
	pb.getDoubleProp("pitch", &pitch, -90, 90);
	pb.getDoubleProp("roll", &roll);
	pb.getBoolProp("dof", &dof);
	pb.getDoubleProp("fNumber", &fNumber, 0);
	pb.getIntProp("numSamples", &numDOFSamples, 1);
	pb.getDoubleProp("focalPlaneDist", &focalPlaneDist, 0.1);
	pb.getBoolProp("autofocus", &autofocus);
	pb.getDoubleProp("stereoSeparation", &stereoSeparation, 0.0);
	pb.getColorProp("leftMask", &leftMask);
	pb.getColorProp("rightMask", &rightMask);
	apertureSize = 4.5 / fNumber;
}

void Camera::beginFrame()
{
	// the corners of a screen at z = 1 whose half-diagonal (measured to the top-left corner, projected on
	// the z = 1 plane) spans tan(fov/2); then rotate by roll, pitch, yaw.
	Vec3 BC = Vec3(-aspectRatio, 1, 1) - Vec3(0, 0, 1);
	double lenBC = BC.length();
	double lenWanted = std::tan(radians(fov / 2));
	double m = lenWanted / lenBC;
	topLeft = Vec3(-aspectRatio * m, +m, 1);
	topRight = Vec3(+aspectRatio * m, +m, 1);
	bottomLeft = Vec3(-aspectRatio * m, -m, 1);
	w = owner ? owner->settings.frameWidth : 0;
	h = owner ? owner->settings.frameHeight : 0;
	Mat3 rotation = rotZ(radians(roll)) * rotX(radians(pitch)) * rotY(radians(yaw));
	topLeft = topLeft * rotation;
	topRight = topRight * rotation;
	bottomLeft = bottomLeft * rotation;
	frontDir = Vec3(0, 0, 1) * rotation;
	upDir = Vec3(0, 1, 0) * rotation;
	rightDir = Vec3(1, 0, 0) * rotation;
	apertureSize = 1.0 / fNumber;
}

void Camera::move(double rx, double ry) { pos = pos + rx * rightDir + ry * frontDir; }

void Camera::rotate(double rx, double ry)
{
	yaw += rx;
	pitch += ry;
	pitch = std::min(pitch, +90.0);
	pitch = std::max(pitch, -90.0);
}

bool CubemapEnvironment::loadMaps(const char* folder)
{
	static const char* prefixes[2] = { "neg", "pos" };
	static const char* axes[3] = { "x", "y", "z" };
	static const char* suffixes[2] = { ".bmp", ".exr" };
	int n = 0;
	for (int pi = 0; pi < 2; pi++)
		for (int axis = 0; axis < 3; axis++) {
			std::unique_ptr<Bitmap> map(new Bitmap);
			for (int si = 0; si < 2; si++) {
				std::string fn = std::string(folder) + "/" + prefixes[pi] + axes[axis] + suffixes[si];
				struct stat st;
				if (stat(fn.c_str(), &st) == 0 && map->loadImage(fn.c_str())) break;
			}
			if (!map->isOK()) return false;
			maps[n++] = std::move(map);
		}
	loaded = true;
	return true;
}

void CubemapEnvironment::fillProperties(ParsedBlock& pb)
{
	std::string folder;
	if (!pb.getFilenameProp("folder", &folder)) pb.requiredProp("folder");
	if (!loadMaps(folder.c_str())) {
		// the reference only prints this and later crashes on the NULL maps; we refuse the scene instead
		fprintf(stderr, "CubemapEnvironment: Could not load maps from `%s'\n", folder.c_str());
		pb.signalError("CubemapEnvironment: could not load the six cube faces");
	}
}

void GlobalSettings::fillProperties(ParsedBlock& pb)
{
	pb.getIntProp("frameWidth", &frameWidth);
	pb.getIntProp("frameHeight", &frameHeight);
	pb.getColorProp("ambientLight", &ambientLight);
	pb.getIntProp("maxTraceDepth", &maxTraceDepth);
	pb.getBoolProp("dbg", &dbg);
	pb.getBoolProp("wantAA", &wantAA);
	pb.getFloatProp("saturation", &saturation, 0, 1);
	pb.getBoolProp("wantPrepass", &wantPrepass);
	pb.getBoolProp("gi", &gi);
	pb.getIntProp("pathsPerPixel", &numPaths, 1);
	pb.getIntProp("numThreads", &numThreads);
	pb.getBoolProp("interactive", &interactive);
	pb.getBoolProp("fullscreen", &fullscreen);
}

} // namespace fray
