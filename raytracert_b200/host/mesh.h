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
    Material() : Ns_(0.f), Ni_(1.f), illum_(0), Tr_(1.f) { cleanup(); }

    void cleanup() {
        Kd_set_ = Ka_set_ = Ks_set_ = Ns_set_ = Ni_set_ = Tr_set_ = illum_set_ = false;
        name_ = "empty";
    }
    bool is_valid() const { return Kd_set_ || Ka_set_ || Ks_set_ || Tr_set_; }

    bool has_Kd() const { return Kd_set_; }
    bool has_Ka() const { return Ka_set_; }
    bool has_Ks() const { return Ks_set_; }
    bool has_Ns() const { return Ns_set_; }
    bool has_Ni() const { return Ni_set_; }
    bool has_illum() const { return illum_set_; }
    bool has_Tr() const { return Tr_set_; }

    void set_Kd(float r, float g, float b) { Kd_ = Vec3Df(r, g, b); Kd_set_ = true; }
    void set_Ka(float r, float g, float b) { Ka_ = Vec3Df(r, g, b); Ka_set_ = true; }
    void set_Ks(float r, float g, float b) { Ks_ = Vec3Df(r, g, b); Ks_set_ = true; }
    void set_Ns(float v) { Ns_ = v; Ns_set_ = true; }
    void set_Ni(float v) { Ni_ = v; Ni_set_ = true; }
    void set_illum(int v) { illum_ = v; illum_set_ = true; }
    void set_Tr(float v) { Tr_ = v; Tr_set_ = true; }
    void set_textureName(const std::string& s) { textureName_ = s; }
    void set_name(const std::string& s) { name_ = s; }

    const Vec3Df& Kd() const { return Kd_; }
    const Vec3Df& Ka() const { return Ka_; }
    const Vec3Df& Ks() const { return Ks_; }
    float Ns() const { return Ns_; }
    float Ni() const { return Ni_; }
    int illum() const { return illum_; }
    float Tr() const { return Tr_; }
    const std::string& textureName() const { return textureName_; }
    const std::string& name() const { return name_; }

private:
    Vec3Df Kd_, Ka_, Ks_;
    float Ns_, Ni_;
    int illum_;
    float Tr_;
    bool Kd_set_, Ka_set_, Ks_set_, Ns_set_, Ni_set_, illum_set_, Tr_set_;
    std::string name_, textureName_;
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
