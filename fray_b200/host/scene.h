// scene.h -- host-side scene object model of fray-b200.
//
// Mirrors the reference's API surface so that a `.fray` file means the same thing here:
//   SceneElement life-cycle fillProperties -> beginRender -> beginFrame   /root/reference/src/scene.h:59-141
//   ParsedBlock property getters                                          src/scene.h:143-194, src/scene.cpp:136-357
//   class factory (23 classes)                                            src/scene.cpp:821-848
//   Scene / GlobalSettings                                                src/scene.h:252-299, src/scene.cpp:783-819
// The classes hold PARAMETERS only: intersection, shading and sampling run on the GPU
// (fray_b200/csrc) from the flattened tables produced by flatten() (include/fray_gpu.h).
#pragma once
#include <climits>
#include <memory>
#include <string>
#include <vector>

#include "vec.h"
#include "fray_gpu.h"

namespace fray {

// progress / statistics lines the reference prints to stdout ("Mesh loaded", "KD Tree ... built"); the command line tool
// turns them on, library users get a quiet stdout
extern int g_verbose;

enum ElementType { // src/scene.h:30-39
	ELEM_GEOMETRY, ELEM_SHADER, ELEM_NODE, ELEM_TEXTURE, ELEM_ENVIRONMENT, ELEM_CAMERA, ELEM_SETTINGS, ELEM_LIGHT,
};

struct SyntaxError { // src/scene.h:219-224
	int line;
	std::string msg;
};
struct FileNotFoundError { // src/scene.h:226-231
	int line;
	std::string filename;
};

class Bitmap { // src/bitmap.h:30-60: float RGB image, (0,0) top-left
public:
	int width = -1, height = -1;
	std::vector<Color> data;
	bool isOK() const { return !data.empty(); }
	void generateEmptyImage(int w, int h);
	Color getPixel(int x, int y) const; // black outside, src/bitmap.cpp:67-71
	void setPixel(int x, int y, const Color& c);
	void differentiate(); // src/bitmap.cpp:300-315
	bool loadBMP(const char* fn);  // 8/24/32-bit uncompressed, src/bitmap.cpp:117-195
	bool saveBMP(const char* fn) const; // 24-bit, byte-exact with src/bitmap.cpp:197-236
	bool loadEXR(const char* fn);  // scan-line RGBA half/float; NONE, ZIPS, ZIP, PIZ (own decoder, image.cpp / exr.cpp)
	bool saveEXR(const char* fn) const; // HALF RGBA, A = 1, uncompressed scan lines (src/bitmap.cpp:266-284 semantics)
	bool loadImage(const char* fn); // by extension, src/bitmap.cpp:286-291
	bool saveImage(const char* fn) const;
};

class ParsedBlock;
class SceneParser;
struct Scene;

class SceneElement {
public:
	std::string name;
	virtual ~SceneElement() {}
	virtual ElementType getElementType() const = 0;
	virtual void fillProperties(ParsedBlock&) {}
	virtual void beginRender() {}
	virtual void beginFrame() {}
	virtual void* getInterface(int) { return nullptr; } // src/scene.h:138
};

class Geometry;
class Shader;
class Texture;
struct Node;

class ParsedBlock { // src/scene.h:143-194
public:
	virtual ~ParsedBlock() {}
	virtual bool getIntProp(const char* name, int* value, int minValue = INT_MIN, int maxValue = INT_MAX) = 0;
	virtual bool getBoolProp(const char* name, bool* value) = 0;
	virtual bool getFloatProp(const char* name, float* value, float minValue = -1e17f, float maxValue = 1e17f) = 0;
	virtual bool getDoubleProp(const char* name, double* value, double minValue = -1e120, double maxValue = 1e120) = 0;
	virtual bool getColorProp(const char* name, Color* value, float minComp = -1e17f, float maxComp = 1e17f) = 0;
	virtual bool getVectorProp(const char* name, Vec3* value) = 0;
	virtual bool getGeometryProp(const char* name, Geometry** value) = 0;
	virtual bool getShaderProp(const char* name, Shader** value) = 0;
	virtual bool getTextureProp(const char* name, Texture** value) = 0;
	virtual bool getNodeProp(const char* name, Node** value) = 0;
	virtual bool getStringProp(const char* name, std::string* value) = 0;
	virtual bool getFilenameProp(const char* name, std::string* value) = 0;
	virtual bool getBitmapFileProp(const char* name, Bitmap& bmp) = 0;
	virtual void getTransformProp(Transform& T) = 0;
	virtual void requiredProp(const char* name) = 0;
	virtual void signalError(const char* msg) = 0;
	virtual void signalWarning(const char* msg) = 0;
	virtual int getBlockLines() = 0;
	virtual void getBlockLine(int idx, int& srcLine, std::string& head, std::string& tail) = 0;
	virtual SceneParser& getParser() = 0;
};

class SceneParser { // src/scene.h:196-217
public:
	virtual ~SceneParser() {}
	virtual Shader* findShaderByName(const char* name) = 0;
	virtual Texture* findTextureByName(const char* name) = 0;
	virtual Geometry* findGeometryByName(const char* name) = 0;
	virtual Node* findNodeByName(const char* name) = 0;
	virtual bool resolveFullPath(std::string& path) = 0;
};

// ---- geometries (src/geometry.h:54-154, src/mesh.h:55-100) ----------------------------------------
class Geometry: public SceneElement {
public:
	ElementType getElementType() const override { return ELEM_GEOMETRY; }
	virtual int geomType() const = 0; // FRAY_GEOM_*
};

class Plane: public Geometry {
public:
	double limit = 128, height = 0;
	void fillProperties(ParsedBlock& pb) override;
	int geomType() const override { return FRAY_GEOM_PLANE; }
};

class Sphere: public Geometry {
public:
	Vec3 O;
	double R = 1;
	void fillProperties(ParsedBlock& pb) override;
	int geomType() const override { return FRAY_GEOM_SPHERE; }
};

class Cube: public Geometry {
public:
	Vec3 O;
	double halfSide = 1;
	void fillProperties(ParsedBlock& pb) override;
	int geomType() const override { return FRAY_GEOM_CUBE; }
};

class CsgOp: public Geometry {
public:
	Geometry* left = nullptr;
	Geometry* right = nullptr;
	int op;
	explicit CsgOp(int op): op(op) {}
	void fillProperties(ParsedBlock& pb) override;
	int geomType() const override { return op; }
};

struct Triangle { // src/triangle.h:30-41
	int v[3], n[3], t[3];
	Vec3 gnormal, dNdx, dNdy, AB, AC, ABcrossAC;
};

struct BBox { // src/bbox.h:61-211 (only what the builder needs)
	Vec3 vmin = Vec3(+kInf, +kInf, +kInf), vmax = Vec3(-kInf, -kInf, -kInf);
	void add(const Vec3& p);
	bool inside(const Vec3& p) const;
	bool testIntersect(const Vec3& start, const Vec3& dir, const Vec3& rdir) const;
	bool intersectTriangle(const Vec3& A, const Vec3& B, const Vec3& C) const;
	void split(int axis, double where, BBox& left, BBox& right) const;
};

class Mesh: public Geometry {
public:
	std::vector<Vec3> vertices, normals, uvs;
	std::vector<Triangle> triangles;
	BBox bbox;
	// flattened KD tree (include/fray_gpu.h FrayGpuKdNode); empty = brute force
	std::vector<FrayGpuKdNode> kdNodes;
	std::vector<int32_t> leafRefs;
	int maxTreeDepth = 0, numNodes = 0;
	long long nodeDepthSum = 0;
	bool faceted = false, useKD = true, backfaceCulling = true;

	void fillProperties(ParsedBlock& pb) override;
	bool loadFromOBJ(const char* filename); // src/mesh.cpp:203-258
	void beginRender() override;            // src/mesh.cpp:67-94
	int geomType() const override { return FRAY_GEOM_MESH; }
private:
	void prepareTriangles();                // src/mesh.cpp:273-313
	void buildKD(int nodeIdx, const std::vector<int>& tris, const BBox& box, int depth); // src/mesh.cpp:320-355
};

// ---- textures (src/shading.h:33-107, 227-237) -----------------------------------------------------
class Texture: public SceneElement {
public:
	ElementType getElementType() const override { return ELEM_TEXTURE; }
	virtual int texType() const = 0; // FRAY_TEX_*
};

class CheckerTexture: public Texture {
public:
	Color color1 = Color(0.7f, 0.7f, 0.7f), color2 = Color(0.2f, 0.2f, 0.2f);
	double scaling = 1;
	void fillProperties(ParsedBlock& pb) override;
	int texType() const override { return FRAY_TEX_CHECKER; }
};

class BitmapTexture: public Texture {
public:
	Bitmap bmp;
	double scaling = 1;
	void fillProperties(ParsedBlock& pb) override;
	int texType() const override { return FRAY_TEX_BITMAP; }
};

struct BumpMapperInterface { // src/shading.h:75-81; the GPU reads the parameters, there is nothing to call on the host
	static const int ID = 0x20180522;
	virtual ~BumpMapperInterface() {}
};

class BumpTexture: public Texture, public BumpMapperInterface {
public:
	Bitmap bumpTex;
	double scaling = 1, bumpIntensity = 10.0;
	void fillProperties(ParsedBlock& pb) override;
	void beginRender() override; // differentiate(), src/shading.cpp:387-390
	void* getInterface(int id) override { return id == BumpMapperInterface::ID ? (BumpMapperInterface*) this : nullptr; }
	int texType() const override { return FRAY_TEX_BUMP; }
};

class FresnelTexture: public Texture {
public:
	double ior = 1;
	void fillProperties(ParsedBlock& pb) override;
	int texType() const override { return FRAY_TEX_FRESNEL; }
};

// ---- shaders (src/shading.h:109-255) --------------------------------------------------------------
class Shader: public SceneElement {
public:
	Texture* diffuseTex = nullptr;
	ElementType getElementType() const override { return ELEM_SHADER; }
	virtual int shaderType() const = 0; // FRAY_SHADER_*
};

class ConstantShader: public Shader {
public:
	Color color = Color(1, 0, 0);
	int shaderType() const override { return FRAY_SHADER_CONST; }
};

class Lambert: public Shader {
public:
	Color color = Color(1, 1, 1);
	void fillProperties(ParsedBlock& pb) override;
	int shaderType() const override { return FRAY_SHADER_LAMBERT; }
};

class Phong: public Shader {
public:
	Color color = Color(1, 1, 1);
	double exponent = 10.0, specularMultiplier = 0.25;
	Color specularColor = Color(0.75f, 0.75f, 0.75f);
	void fillProperties(ParsedBlock& pb) override;
	int shaderType() const override { return FRAY_SHADER_PHONG; }
};

class Reflection: public Shader {
public:
	double deflectionScaling = 0, glossiness = 1.0;
	bool pureReflection = false;
	int numSamples = 10;
	Color mult = Color(1, 1, 1);
	void fillProperties(ParsedBlock& pb) override;
	void beginFrame() override; // src/shading.h:197-201
	int shaderType() const override { return FRAY_SHADER_REFL; }
};

class Refraction: public Shader {
public:
	double ior = 1;
	Color mult = Color(1, 1, 1);
	void fillProperties(ParsedBlock& pb) override;
	int shaderType() const override { return FRAY_SHADER_REFR; }
};

class Layered: public Shader {
public:
	struct Layer { Shader* shader; Color opacity; Texture* texture; };
	std::vector<Layer> layers; // at most 32, src/shading.h:244, src/shading.cpp:303-311
	void fillProperties(ParsedBlock& pb) override;
	void addLayer(Shader* shader, Color opacity = Color(1, 1, 1), Texture* texture = nullptr);
	int shaderType() const override { return FRAY_SHADER_LAYERED; }
};

// ---- node, lights, camera, environment, settings --------------------------------------------------
struct Node: public SceneElement { // src/geometry.h:158-177
	Geometry* geometry = nullptr;
	Shader* shader = nullptr;
	Transform T;
	Texture* bump = nullptr;
	ElementType getElementType() const override { return ELEM_NODE; }
	void fillProperties(ParsedBlock& pb) override;
};

class Light: public SceneElement { // src/lights.h:32-54
public:
	Color color = Color(1, 1, 1);
	float power = 1;
	ElementType getElementType() const override { return ELEM_LIGHT; }
	void fillProperties(ParsedBlock& pb) override;
	virtual int lightType() const = 0;
};

class PointLight: public Light { // src/lights.h:57-75
public:
	Vec3 pos;
	void fillProperties(ParsedBlock& pb) override;
	int lightType() const override { return FRAY_LIGHT_POINT; }
};

class RectLight: public Light { // src/lights.h:77-99
public:
	Transform T;
	int xSubd = 1, ySubd = 1; // the reference leaves these uninitialised when omitted
	Vec3 center;
	double area = 0;
	void fillProperties(ParsedBlock& pb) override;
	void beginFrame() override; // src/lights.cpp:37-46
	int lightType() const override { return FRAY_LIGHT_RECT; }
};

class Camera: public SceneElement { // src/camera.h:37-86
public:
	Vec3 topLeft, topRight, bottomLeft, frontDir, upDir, rightDir;
	double w = 0, h = 0, apertureSize = 0;
	Vec3 pos;
	double yaw = 0, pitch = 0, roll = 0, fov = 90.0, aspectRatio = 1.3333, focalPlaneDist = 5.0, fNumber = 2.0;
	bool dof = false, autofocus = true;
	int numDOFSamples = 32;
	double stereoSeparation = 0;
	Color leftMask = Color(1, 0, 0), rightMask = Color(0, 1, 1);
	Scene* owner = nullptr; // for frame size (the reference asks the SDL surface, src/camera.cpp:45-46)
	ElementType getElementType() const override { return ELEM_CAMERA; }
	void fillProperties(ParsedBlock& pb) override;
	void beginFrame() override; // src/camera.cpp:34-57
	void move(double rx, double ry);   // src/camera.cpp:95-98
	void rotate(double rx, double ry); // src/camera.cpp:100-106
};

class Environment: public SceneElement { // src/environment.h:36-48
public:
	bool loaded = false;
	ElementType getElementType() const override { return ELEM_ENVIRONMENT; }
};

class CubemapEnvironment: public Environment { // src/environment.h:51-77
public:
	std::unique_ptr<Bitmap> maps[6]; // negx,negy,negz,posx,posy,posz
	bool loadMaps(const char* folder); // src/environment.cpp:31-52
	void fillProperties(ParsedBlock& pb) override;
};

struct GlobalSettings: public SceneElement { // src/scene.h:252-278, defaults src/scene.cpp:783-797
	int frameWidth = 800, frameHeight = 600;
	Color ambientLight;
	bool wantAA = true, gi = false;
	int maxTraceDepth = 4;
	bool dbg = false;
	float saturation = 1;
	bool wantPrepass = true;
	int numPaths = 10;
	int numThreads = 0;
	bool interactive = false, fullscreen = false;
	ElementType getElementType() const override { return ELEM_SETTINGS; }
	void fillProperties(ParsedBlock& pb) override;
};

struct Scene { // src/scene.h:280-299
	std::vector<Geometry*> geometries;
	std::vector<Shader*> shaders;
	std::vector<Node*> nodes;
	std::vector<Node*> superNodes;
	std::vector<Texture*> textures;
	std::vector<Light*> lights;
	Environment* environment = nullptr;
	Camera* camera = nullptr;
	GlobalSettings settings;
	std::string lastError;

	Scene() {}
	~Scene();
	Scene(const Scene&) = delete;
	Scene& operator=(const Scene&) = delete;

	bool parseScene(const char* sceneFile); // src/scene.cpp:751-755
	void beginRender();                     // src/scene.cpp:757-768
	void beginFrame();                      // src/scene.cpp:770-781
	int samplesPerPixel() const;            // src/main.cpp:395-400
};

// ---- flattening (the drop-in boundary, include/fray_gpu.h) ----------------------------------------
struct FlatScene {
	FrayGpuScene view; // pointers into the vectors below
	std::vector<FrayGpuNode> nodes;
	std::vector<FrayGpuGeometry> geometries;
	std::vector<FrayGpuMesh> meshes;
	std::vector<FrayGpuShader> shaders;
	std::vector<FrayGpuLayer> layers;
	std::vector<FrayGpuTexture> textures;
	std::vector<FrayGpuBitmap> bitmaps;
	std::vector<FrayGpuLight> lights;
	std::vector<double> vertices, normals, uvs;
	std::vector<int32_t> tri_v, tri_n, tri_t;
	std::vector<double> tri_gnormal, tri_dndx, tri_dndy, tri_ab, tri_ac, tri_abxac;
	std::vector<FrayGpuKdNode> kd_nodes;
	std::vector<int32_t> leaf_refs;
	std::vector<float> texels;
	void rebind(); // refresh `view` after the vectors changed
};

// Scene must have had beginRender() and beginFrame() called. Returns false and sets scene.lastError on failure.
bool flatten(Scene& scene, FlatScene& out);
void flattenCamera(const Scene& scene, FrayGpuCamera& out);

} // namespace fray
