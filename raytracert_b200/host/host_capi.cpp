// librt_host.so -- C entry points over the C++ host side (loader, flatten, camera, PPM) so that the
// Python tests/bench can drive exactly the code the C++ drop-in (main.cpp) runs. No CUDA in here.
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>
#include "flatten.h"
#include "glu_math.h"
#include "image.h"
#include "mesh.h"

namespace {
struct HostScene {
    Mesh mesh;
    std::vector<Vec3Df> normals;
    FlatScene flat;
    rt_scene view;
};
}

extern "C" {

// OBJ -> Mesh -> face normals -> flat SoA scene. NULL when the file cannot be opened or is inconsistent.
void* rth_load_obj(const char* path) {
    HostScene* h = new HostScene();
    if (!h->mesh.loadMesh(path, true)) { delete h; return nullptr; }
    for (size_t i = 0; i < h->mesh.triangles.size(); ++i)
        for (int k = 0; k < 3; ++k)
            if (h->mesh.triangles[i].v[k] >= h->mesh.vertices.size()) { delete h; return nullptr; }
    h->mesh.computeVertexNormals();
    append_face_normals(h->mesh, h->normals);
    if (!flatten_mesh(h->mesh, h->normals, h->flat)) { delete h; return nullptr; }
    h->view = h->flat.view();
    return h;
}

void rth_free(void* handle) { delete static_cast<HostScene*>(handle); }

void rth_counts(void* handle, int* nv, int* nt, int* nm) {
    HostScene* h = static_cast<HostScene*>(handle);
    *nv = (int)h->mesh.vertices.size(); *nt = (int)h->mesh.triangles.size(); *nm = (int)h->mesh.materials.size();
}

void rth_get_vertices(void* handle, float* out) {
    HostScene* h = static_cast<HostScene*>(handle);
    for (size_t i = 0; i < h->mesh.vertices.size(); ++i)
        for (int c = 0; c < 3; ++c) out[3 * i + c] = h->mesh.vertices[i].p[c];
}

void rth_get_triangles(void* handle, uint32_t* idx, uint32_t* mat) {
    HostScene* h = static_cast<HostScene*>(handle);
    for (size_t i = 0; i < h->mesh.triangles.size(); ++i) {
        for (int c = 0; c < 3; ++c) idx[3 * i + c] = h->mesh.triangles[i].v[c];
        mat[i] = h->mesh.triangleMaterials[i];
    }
}

void rth_get_normals(void* handle, float* out) {
    HostScene* h = static_cast<HostScene*>(handle);
    for (size_t i = 0; i < h->normals.size(); ++i)
        for (int c = 0; c < 3; ++c) out[3 * i + c] = h->normals[i][c];
}

// 16 floats, same layout as the oracle harness: Kd Ns | Ka Ni | Ks Tr | flags 0 0 0.
void rth_get_material(void* handle, int i, float* out, char* name, int name_cap) {
    HostScene* h = static_cast<HostScene*>(handle);
    const rt_material m = flatten_material(h->mesh.materials[i]);
    for (int c = 0; c < 3; ++c) { out[c] = m.Kd[c]; out[4 + c] = m.Ka[c]; out[8 + c] = m.Ks[c]; }
    out[3] = m.Ns; out[7] = m.Ni; out[11] = m.Tr; out[12] = (float)m.flags; out[13] = out[14] = out[15] = 0.f;
    if (name && name_cap > 0) { strncpy(name, h->mesh.materials[i].name().c_str(), name_cap - 1); name[name_cap - 1] = 0; }
}

// Face normals for a flat indexed mesh, with calculateNormals()'s arithmetic (raytracing.cpp:78-86).
void rth_face_normals(int nv, const float* verts, int nt, const uint32_t* idx, float* out) {
    Mesh m;
    for (int i = 0; i < nv; ++i) m.vertices.push_back(Vertex(Vec3Df(verts[3 * i], verts[3 * i + 1], verts[3 * i + 2])));
    for (int i = 0; i < nt; ++i) m.triangles.push_back(Triangle(idx[3 * i], 0, idx[3 * i + 1], 0, idx[3 * i + 2], 0));
    std::vector<Vec3Df> n;
    append_face_normals(m, n);
    for (int i = 0; i < nt; ++i)
        for (int c = 0; c < 3; ++c) out[3 * i + c] = n[i][c];
}

const rt_scene* rth_scene(void* handle) { return &static_cast<HostScene*>(handle)->view; }

// The four produceRay() calls of main.cpp:355-358 for a W x H viewport: corner c at window
// (x_I, y_new = H - y_I), unprojected at depth 0 (origin) and 1 (dest), cast to float.
int rth_corner_rays(const double* modelview, const double* projection, int W, int H, float* out24) {
    glu::Mat4 mv, pr;
    memcpy(mv.m, modelview, sizeof(mv.m));
    memcpy(pr.m, projection, sizeof(pr.m));
    const int viewport[4] = {0, 0, W, H};
    const int cx[4] = {0, 0, W - 1, W - 1};
    const int cy[4] = {0, H - 1, 0, H - 1};
    for (int c = 0; c < 4; ++c) {
        const int y_new = viewport[3] - cy[c];
        for (int depth = 0; depth < 2; ++depth) {
            double p[3];
            if (!glu::unproject(cx[c], y_new, depth, mv, pr, viewport, p)) return -1;
            for (int k = 0; k < 3; ++k) out24[c * 6 + depth * 3 + k] = float(p[k]);
        }
    }
    return 0;
}

// modelview = T(0,0,-4) (main.cpp:217-219), projection = gluPerspective(50, (float)w/h, 1, 10) (main.cpp:294).
void rth_default_camera(int W, int H, double* modelview, double* projection, float* eye) {
    glu::Mat4 mv = glu::translate(0, 0, -4);
    glu::Mat4 pr = glu::perspective(50, (float)W / H, 1, 10);
    memcpy(modelview, mv.m, sizeof(mv.m));
    memcpy(projection, pr.m, sizeof(pr.m));
    double e[3] = {0, 0, 0}; glu::camera_position(mv, e);
    for (int k = 0; k < 3; ++k) eye[k] = float(e[k]);
}

void rth_lookat_camera(const double* eye_in, const double* center, const double* up, int W, int H, double* modelview, double* projection, float* eye) {
    glu::Mat4 mv = glu::look_at(eye_in, center, up);
    glu::Mat4 pr = glu::perspective(50, (float)W / H, 1, 10);
    memcpy(modelview, mv.m, sizeof(mv.m));
    memcpy(projection, pr.m, sizeof(pr.m));
    double e[3] = {0, 0, 0}; glu::camera_position(mv, e);
    for (int k = 0; k < 3; ++k) eye[k] = float(e[k]);
}

int rth_write_ppm(const char* path, const float* rgb, int W, int H) {
    Image img(W, H);
    memcpy(img._image.data(), rgb, sizeof(float) * 3 * (size_t)W * H);
    return img.writeImage(path) ? 0 : -1;
}

}  // extern "C"
