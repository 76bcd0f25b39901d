// Mesh (+ face normals) -> rt_scene: the SoA float4 buffers librt_b200 uploads to HBM.
// Replaces nothing in the reference one-to-one: it is the bridge between the reference's AoS containers
// (mesh.h:184-200, raytracing.cpp:33) and the C ABI (include/rt_b200.h).
#pragma once
#include <vector>
#include "mesh.h"
#include "../../include/rt_b200.h"

struct FlatScene {
    std::vector<float> v0, v1, v2, normal;   // 4 floats per triangle
    std::vector<uint32_t> tri_material;
    std::vector<rt_material> materials;
    std::vector<rt_sphere> spheres;
    rt_scene view() const {
        rt_scene s;
        s.n_triangles = (uint32_t)tri_material.size();
        s.v0 = v0.data(); s.v1 = v1.data(); s.v2 = v2.data(); s.normal = normal.data();
        s.tri_material = tri_material.data();
        s.n_materials = (uint32_t)materials.size();
        s.materials = materials.data();
        s.n_spheres = (uint32_t)spheres.size();
        s.spheres = spheres.empty() ? nullptr : spheres.data();
        return s;
    }
};

// Per-triangle unit face normal normalize(cross(p1-p0, p2-p0)) -- the arithmetic of calculateNormals()
// (raytracing.cpp:78-86): float cross product, Vec3Df::normalize (sqrt, reciprocal, three multiplies);
// a degenerate triangle keeps the zero vector. Appends, like the reference's push_back.
inline void append_face_normals(const Mesh& mesh, std::vector<Vec3Df>& out) {
    for (size_t i = 0; i < mesh.triangles.size(); ++i) {
        const Vec3Df& p0 = mesh.vertices[mesh.triangles[i].v[0]].p;
        Vec3Df n = Vec3Df::crossProduct(mesh.vertices[mesh.triangles[i].v[1]].p - p0, mesh.vertices[mesh.triangles[i].v[2]].p - p0);
        n.normalize();
        out.push_back(n);
    }
}

inline rt_material flatten_material(const Material& m) {
    rt_material r;
    for (int c = 0; c < 3; ++c) { r.Kd[c] = m.Kd()[c]; r.Ka[c] = m.Ka()[c]; r.Ks[c] = m.Ks()[c]; }
    r.Ns = m.Ns(); r.Ni = m.Ni(); r.Tr = m.Tr();
    static_assert(Material::kKd == RT_HAS_KD && Material::kKa == RT_HAS_KA && Material::kKs == RT_HAS_KS && Material::kNs == RT_HAS_NS &&
                  Material::kNi == RT_HAS_NI && Material::kTr == RT_HAS_TR, "Material::Field must match the RT_HAS_* bits");
    r.flags = m.seen() & (RT_HAS_KD | RT_HAS_KA | RT_HAS_KS | RT_HAS_NS | RT_HAS_NI | RT_HAS_TR);
    r.pad[0] = r.pad[1] = r.pad[2] = 0;
    return r;
}

// Returns false if a triangle refers to a vertex or material that does not exist (the reference would
// read out of bounds).
inline bool flatten_mesh(const Mesh& mesh, const std::vector<Vec3Df>& face_normals, FlatScene& out) {
    const size_t n = mesh.triangles.size();
    if (face_normals.size() != n || mesh.triangleMaterials.size() < n) return false;
    out.v0.assign(4 * n, 0.f); out.v1.assign(4 * n, 0.f); out.v2.assign(4 * n, 0.f); out.normal.assign(4 * n, 0.f);
    out.tri_material.resize(n);
    for (size_t i = 0; i < n; ++i) {
        const Triangle& t = mesh.triangles[i];
        float* dst[3] = {&out.v0[4 * i], &out.v1[4 * i], &out.v2[4 * i]};
        for (int k = 0; k < 3; ++k) {
            if (t.v[k] >= mesh.vertices.size()) return false;
            const Vec3Df& p = mesh.vertices[t.v[k]].p;
            dst[k][0] = p[0]; dst[k][1] = p[1]; dst[k][2] = p[2];
        }
        for (int c = 0; c < 3; ++c) out.normal[4 * i + c] = face_normals[i][c];
        if (mesh.triangleMaterials[i] >= mesh.materials.size()) return false;
        out.tri_material[i] = mesh.triangleMaterials[i];
    }
    out.materials.clear();
    for (size_t i = 0; i < mesh.materials.size(); ++i) out.materials.push_back(flatten_material(mesh.materials[i]));
    return true;
}
