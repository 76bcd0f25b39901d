// Material / Triangle / Mesh -- host-side scene types with the API of the reference's mesh.h
// (CG_Project/mesh.h:10-125 Material, :131-167 Triangle, :172-201 Mesh), written from scratch.
//
// Behaviour kept on purpose:
//   * Material::cleanup() clears the has_* flags and the name only -- VALUES SURVIVE, so a material that
//     omits a key inherits the previous material's value (flag clear)                     (mesh.h:43-53)
//   * is_valid() == Kd || Ka || Ks || Tr set                                              (mesh.h:55-56)
// Pinned (the reference leaves these indeterminate): a value that was never set by any
// earlier statement is Kd=Ka=Ks=(0,0,0), Ns=0, Ni=1, Tr=1, illum=0.
#pragma once
#include <map>
#include <string>
#include <vector>
#include "Vertex.h"

class Material {
public:
    // which statements of the MTL block were seen; the bit values are the RT_HAS_* flags of include/rt_b200.h
    enum Field : unsigned { kKd = 1u, kKa = 2u, kKs = 4u, kNs = 8u, kNi = 16u, kTr = 32u, kIllum = 64u };

    Material() : seen_(0u), scalar_{0.f, 1.f, 1.f}, illum_(0), name_("empty") {}   // Ns = 0, Ni = 1, Tr = 1 (pinned)

    // Forget which fields were set (and the name); the VALUES stay -- see the note above.
    void cleanup() { seen_ = 0u; name_ = "empty"; }
    bool is_valid() const { return (seen_ & (kKd | kKa | kKs | kTr)) != 0u; }
    unsigned seen() const { return seen_; }

    bool has_Kd() const { return has(kKd); }
    bool has_Ka() const { return has(kKa); }
    bool has_Ks() const { return has(kKs); }
    bool has_Ns() const { return has(kNs); }
    bool has_Ni() const { return has(kNi); }
    bool has_illum() const { return has(kIllum); }
    bool has_Tr() const { return has(kTr); }

    void set_Kd(float r, float g, float b) { colour_[0] = Vec3Df(r, g, b); seen_ |= kKd; }
    void set_Ka(float r, float g, float b) { colour_[1] = Vec3Df(r, g, b); seen_ |= kKa; }
    void set_Ks(float r, float g, float b) { colour_[2] = Vec3Df(r, g, b); seen_ |= kKs; }
    void set_Ns(float v) { scalar_[0] = v; seen_ |= kNs; }
    void set_Ni(float v) { scalar_[1] = v; seen_ |= kNi; }
    void set_Tr(float v) { scalar_[2] = v; seen_ |= kTr; }
    void set_illum(int v) { illum_ = v; seen_ |= kIllum; }
    void set_textureName(const std::string& s) { texture_ = s; }
    void set_name(const std::string& s) { name_ = s; }

    const Vec3Df& Kd() const { return colour_[0]; }   // diffuse
    const Vec3Df& Ka() const { return colour_[1]; }   // ambient
    const Vec3Df& Ks() const { return colour_[2]; }   // specular (also the reflection weight, flag or not)
    float Ns() const { return scalar_[0]; }           // shininess
    float Ni() const { return scalar_[1]; }           // index of refraction
    float Tr() const { return scalar_[2]; }           // "d" / "Tr": 1 = opaque
    int illum() const { return illum_; }
    const std::string& textureName() const { return texture_; }
    const std::string& name() const { return name_; }

private:
    bool has(unsigned f) const { return (seen_ & f) != 0u; }
    unsigned seen_;
    Vec3Df colour_[3];   // Kd, Ka, Ks
    float scalar_[3];    // Ns, Ni, Tr
    int illum_;
    std::string name_, texture_;
};

// Vertex ids v[3] and texture-coordinate ids t[3] of one face.
class Triangle {
public:
    unsigned int v[3];
    unsigned int t[3];
    Triangle() : v{0, 0, 0}, t{0, 0, 0} {}
    Triangle(unsigned int v0, unsigned int t0, unsigned int v1, unsigned int t1, unsigned int v2, unsigned int t2)
        : v{v0, v1, v2}, t{t0, t1, t2} {}
    virtual ~Triangle() {}
};

class Mesh {
public:
    Mesh() {}
    Mesh(const std::vector<Vertex>& v, const std::vector<Triangle>& t) : vertices(v), triangles(t) {}

    // OBJ (+ referenced MTL) loader; same grammar as the reference (mesh.cpp:95-331). Returns false when
    // the OBJ cannot be opened (the reference crashes in fclose(NULL) there -- pinned to an error).
    bool loadMesh(const char* filename, bool randomizeTriangulation);
    bool loadMtl(const char* filename, std::map<std::string, unsigned int>& materialIndex);
    void computeVertexNormals();
    // The GL preview (mesh.cpp:53-90) is out of scope (no GL in a headless build); kept as no-ops so
    // skeleton code that calls them still links.
    void draw() {}
    void drawSmooth() {}

    std::vector<Vertex> vertices;
    std::vector<Vec3Df> texcoords;
    std::vector<Triangle> triangles;
    std::vector<unsigned int> triangleMaterials;  // one material index per triangle
    std::vector<Material> materials;              // [0] is the built-in default material

    bool verbose = false;  // print the reference's progress lines ("Load material file ...")
};
