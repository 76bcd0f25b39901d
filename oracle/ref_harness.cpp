// oracle/_ref harness -- TEST INFRASTRUCTURE ONLY.
//
// Links the UNMODIFIED reference translation units /root/reference/CG_Project/{raytracing,mesh}.cpp
// (compiled where they lie, behind oracle/shim/) and plays the part of the reference's main.cpp:
// it defines the globals main.cpp owns (main.cpp:17-18,130,137-141), restates the frame loop of the
// 'r' key handler (main.cpp:347-395) and exposes everything over a small C ABI so tests/ and
// bench.py's reference arm can drive the real reference code.  Nothing in the product path may link
// or load this file.
//
// What is reference code and what is harness code:
//   * performRayTracing / trace / intersectMesh / shade / init / calculateNormals / Mesh::loadMesh /
//     Mesh::loadMtl ... : the reference, untouched.
//   * ref_render's y/x/subx/suby loop: restatement of main.cpp:360-393 (operation for operation,
//     float for float); OpenMP over rows is added here, the reference code underneath is unchanged
//     (8 threads give a bit-identical image).
//   * UB pins: ref_pin_material_scalars() gives Tr/Ni a defined value where the MTL never
//     set one (the reference leaves them uninitialised, mesh.h:119-122).
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <vector>
#include <string>
#include "raytracing.h"  // the reference's own header (pulls mesh.h, Vec3D.h)

// ---- globals that main.cpp would define (main.cpp:17-18,130,137-141) -------------------------
Vec3Df MyCameraPosition;
std::vector<Vec3Df> MyLightPositions;
Mesh MyMesh;
unsigned int WindowSize_X = 500;
unsigned int WindowSize_Y = 500;
unsigned int RayTracingResolutionX = 500;
unsigned int RayTracingResolutionY = 500;

// produceRay lives in main.cpp (needs GL state); only the debug 'd' key uses it from raytracing.cpp.
void produceRay(int, int, Vec3Df& origin, Vec3Df& dest) { origin = Vec3Df(0, 0, 0); dest = Vec3Df(0, 0, -1); }

// ---- external-linkage knobs inside raytracing.cpp (raytracing.cpp:15-37) ------------------------
extern bool Ambient, Diffuse, Reflection, Shadows, Specular, Refraction, DebugMode;
extern int max_lvl;
extern std::vector<Vec3Df> normals;
int intersectMesh(Vec3Df origin, Vec3Df dest, Vec3Df* intersectOut);  // raytracing.cpp:161

extern "C" {

// Loads an OBJ through the reference's own init() (raytracing.cpp:42-73). Returns #triangles.
int ref_load_obj(const char* path) {
    normals.clear();
    MyMesh.triangleMaterials.clear();  // loadMesh never clears it (mesh.cpp:97-99)
    MyLightPositions.clear();
    FILE* f = fopen(path, "r");  // pinned UB (iii): missing OBJ is an error, not fclose(NULL)
    if (!f) return -1;
    fclose(f);
    std::vector<char> buf(path, path + strlen(path) + 1);
    init(buf.data());
    return (int)MyMesh.triangles.size();
}

void ref_counts(int* nv, int* nt, int* nm) {
    *nv = (int)MyMesh.vertices.size();
    *nt = (int)MyMesh.triangles.size();
    *nm = (int)MyMesh.materials.size();
}

void ref_get_vertices(float* out) {
    for (size_t i = 0; i < MyMesh.vertices.size(); ++i)
        for (int c = 0; c < 3; ++c) out[3 * i + c] = MyMesh.vertices[i].p[c];
}

void ref_get_triangles(uint32_t* idx, uint32_t* mat) {
    for (size_t i = 0; i < MyMesh.triangles.size(); ++i) {
        for (int c = 0; c < 3; ++c) idx[3 * i + c] = MyMesh.triangles[i].v[c];
        mat[i] = MyMesh.triangleMaterials[i];
    }
}

void ref_get_normals(float* out) {
    for (size_t i = 0; i < normals.size(); ++i)
        for (int c = 0; c < 3; ++c) out[3 * i + c] = normals[i][c];
}

// 16 floats: Kd xyz Ns | Ka xyz Ni | Ks xyz Tr | flags(bitmask as float) 0 0 0.
// flags: 1 Kd, 2 Ka, 4 Ks, 8 Ns, 16 Ni, 32 Tr.  Values whose flag is clear may be indeterminate.
void ref_get_material(int i, float* out, char* name, int name_cap) {
    Material& m = MyMesh.materials[i];
    for (int c = 0; c < 3; ++c) { out[c] = m.Kd()[c]; out[4 + c] = m.Ka()[c]; out[8 + c] = m.Ks()[c]; }
    out[3] = m.Ns(); out[7] = m.Ni(); out[11] = m.Tr();
    int flags = (m.has_Kd() ? 1 : 0) | (m.has_Ka() ? 2 : 0) | (m.has_Ks() ? 4 : 0) | (m.has_Ns() ? 8 : 0) |
                (m.has_Ni() ? 16 : 0) | (m.has_Tr() ? 32 : 0);
    out[12] = (float)flags; out[13] = out[14] = out[15] = 0.f;
    if (name && name_cap > 0) { strncpy(name, m.name().c_str(), name_cap - 1); name[name_cap - 1] = 0; }
}

// Bypass the loader: fill MyMesh from flat arrays (so the GPU box needs no OBJ of the reference's).
// materials: n x 16 floats in ref_get_material's layout; flagged fields go through the public setters.
void ref_set_scene(int nv, const float* verts, int nt, const uint32_t* idx, const uint32_t* mat, int nm,
                   const float* mats) {
    normals.clear();
    MyLightPositions.clear();
    MyMesh.vertices.clear();
    MyMesh.triangles.clear();
    MyMesh.triangleMaterials.clear();
    MyMesh.materials.clear();
    for (int i = 0; i < nv; ++i) MyMesh.vertices.push_back(Vertex(Vec3Df(verts[3 * i], verts[3 * i + 1], verts[3 * i + 2])));
    for (int i = 0; i < nt; ++i) {
        MyMesh.triangles.push_back(Triangle(idx[3 * i], 0, idx[3 * i + 1], 0, idx[3 * i + 2], 0));
        MyMesh.triangleMaterials.push_back(mat[i]);
    }
    for (int i = 0; i < nm; ++i) {
        const float* m = mats + 16 * i;
        int flags = (int)m[12];
        Material M;
        // deterministic contents first (a flag-less field keeps the value but not the flag)
        M.set_Kd(m[0], m[1], m[2]); M.set_Ka(m[4], m[5], m[6]); M.set_Ks(m[8], m[9], m[10]);
        M.set_Ns(m[3]); M.set_Ni(m[7]); M.set_Tr(m[11]);
        Material V = M;   // values now defined
        V.cleanup();      // cleanup() clears flags only (mesh.h:43-53) -- exactly the leak semantics
        if (flags & 1) V.set_Kd(m[0], m[1], m[2]);
        if (flags & 2) V.set_Ka(m[4], m[5], m[6]);
        if (flags & 4) V.set_Ks(m[8], m[9], m[10]);
        if (flags & 8) V.set_Ns(m[3]);
        if (flags & 16) V.set_Ni(m[7]);
        if (flags & 32) V.set_Tr(m[11]);
        MyMesh.materials.push_back(V);
    }
    calculateNormals();  // raytracing.cpp:78-86
}

// Pin (i) of SURVEY 8c: Tr/Ni of material i get a defined value (set_Tr/set_Ni, mesh.h:78-85).
void ref_set_material_tr(int i, float v) { MyMesh.materials[i].set_Tr(v); }
void ref_set_material_ni(int i, float v) { MyMesh.materials[i].set_Ni(v); }

void ref_set_camera(const float* eye) { MyCameraPosition = Vec3Df(eye[0], eye[1], eye[2]); }

void ref_set_lights(int n, const float* xyz) {
    MyLightPositions.clear();
    for (int i = 0; i < n; ++i) MyLightPositions.push_back(Vec3Df(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]));
}

void ref_set_toggles(int ambient, int diffuse, int specular, int reflection, int shadows, int refraction) {
    Ambient = ambient; Diffuse = diffuse; Specular = specular; Reflection = reflection; Shadows = shadows;
    Refraction = refraction; DebugMode = false;
}

void ref_set_max_lvl(int l) { max_lvl = l; }

// corners: o00 d00 o01 d01 o10 d10 o11 d11 (3 floats each), i.e. produceRay at
// (0,0), (0,H-1), (W-1,0), (W-1,H-1) as in main.cpp:355-358.
// Renders the pixel lattice y = y0, y0+ystep, ... < H  x  x = x0, x0+xstep, ... < W (the full frame for
// 0,1,0,1).  rgb: 3*W*H floats (clamped like RGBValue, main.cpp:24-42), pixels not rendered are left
// untouched.  sample_rgb (optional): 3 floats per sample, sample index ((y*W+x)*pfX+subx)*pfY+suby.
// sample_prim (optional): primary primitive id per sample.  OpenMP runs over the lattice's pixels; each
// pixel is computed by exactly the reference's per-pixel code, so the result does not depend on threads.
void ref_render(const float* corners, int W, int H, int pfX, int pfY, int y0, int ystep, int x0, int xstep, float* rgb,
                float* sample_rgb, int32_t* sample_prim, int nthreads) {
    Vec3Df origin00(corners[0], corners[1], corners[2]), dest00(corners[3], corners[4], corners[5]);
    Vec3Df origin01(corners[6], corners[7], corners[8]), dest01(corners[9], corners[10], corners[11]);
    Vec3Df origin10(corners[12], corners[13], corners[14]), dest10(corners[15], corners[16], corners[17]);
    Vec3Df origin11(corners[18], corners[19], corners[20]), dest11(corners[21], corners[22], corners[23]);
    unsigned int pixelfactorX_ = pfX, pixelfactorY_ = pfY;
    unsigned int WindowSize_X_ = W, WindowSize_Y_ = H;
    float divX = (WindowSize_X_ * pixelfactorX_ - 1);  // main.cpp:360
    float divY = (WindowSize_Y_ * pixelfactorY_ - 1);  // main.cpp:361
    int raysPerPixel = (pixelfactorX_ * pixelfactorY_); // main.cpp:362
    if (ystep < 1) ystep = 1;
    if (xstep < 1) xstep = 1;
    const long nrows = (H - y0 + ystep - 1) / ystep, ncols = (W - x0 + xstep - 1) / xstep;
    if (nrows <= 0 || ncols <= 0) return;
    (void)nthreads;
#pragma omp parallel for schedule(dynamic, 16) num_threads(nthreads)
    for (long item = 0; item < nrows * ncols; ++item) {
        unsigned int y = y0 + (unsigned int)(item / ncols) * ystep;
        unsigned int x = x0 + (unsigned int)(item % ncols) * xstep;
        Vec3Df rgbv = Vec3Df(0, 0, 0);
        for (int subx = 0; subx < (int)pixelfactorX_; subx++) {
            for (int suby = 0; suby < (int)pixelfactorY_; suby++) {
                float xscale = 1.0f - (float(x) * pixelfactorX_ + subx) / divX;  // main.cpp:380
                float yscale = 1.0f - (float(y) * pixelfactorY_ + suby) / divY;  // main.cpp:381
                Vec3Df origin = yscale * (xscale * origin00 + (1 - xscale) * origin10) +
                                (1 - yscale) * (xscale * origin01 + (1 - xscale) * origin11);
                Vec3Df dest = yscale * (xscale * dest00 + (1 - xscale) * dest10) +
                              (1 - yscale) * (xscale * dest01 + (1 - xscale) * dest11);
                Vec3Df c = performRayTracing(origin, dest);  // main.cpp:388
                rgbv += c;
                size_t s = (((size_t)y * W + x) * pfX + subx) * pfY + suby;
                if (sample_rgb) { sample_rgb[3 * s] = c[0]; sample_rgb[3 * s + 1] = c[1]; sample_rgb[3 * s + 2] = c[2]; }
                if (sample_prim) { Vec3Df tmp; sample_prim[s] = intersectMesh(origin, dest, &tmp); }
            }
        }
        rgbv = rgbv / raysPerPixel;  // main.cpp:391
        float ch[3] = {rgbv[0], rgbv[1], rgbv[2]};
        for (int c = 0; c < 3; ++c) {  // RGBValue ctor, main.cpp:29-41
            if (ch[c] > 1) ch[c] = 1.0;
            if (ch[c] < 0) ch[c] = 0.0;
            rgb[3 * ((size_t)W * y + x) + c] = ch[c];
        }
    }
}

// Batch of single rays through performRayTracing (raytracing.cpp:410) + primary id (raytracing.cpp:161).
void ref_trace(int n, const float* origins, const float* dests, float* rgb, int32_t* prim, float* hit) {
    for (int i = 0; i < n; ++i) {
        Vec3Df o(origins[3 * i], origins[3 * i + 1], origins[3 * i + 2]);
        Vec3Df d(dests[3 * i], dests[3 * i + 1], dests[3 * i + 2]);
        Vec3Df c = performRayTracing(o, d);
        rgb[3 * i] = c[0]; rgb[3 * i + 1] = c[1]; rgb[3 * i + 2] = c[2];
        if (prim || hit) {
            Vec3Df I;
            int id = intersectMesh(o, d, &I);
            if (prim) prim[i] = id;
            if (hit) { hit[3 * i] = I[0]; hit[3 * i + 1] = I[1]; hit[3 * i + 2] = I[2]; }
        }
    }
}

// The reference's quantiser (main.cpp:116-117): truncation of v*255.0f.
void ref_quantise(const float* rgb, int n, unsigned char* out) {
    for (int i = 0; i < n; ++i) out[i] = (unsigned char)(rgb[i] * 255.0f);
}

}  // extern "C"
