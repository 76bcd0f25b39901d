// raytracing.cpp -- the plug-in half of the TI1805 contract (reference: CG_Project/raytracing.cpp), with
// every computation moved behind the C ABI of librt_b200:
//   init()               raytracing.cpp:42-73    -> loadMesh + calculateNormals + rt_init + rt_upload_scene
//   performRayTracing()  raytracing.cpp:410-416  -> rt_trace (batch of one)
//   trace()              raytracing.cpp:381-406  -> rt_trace with the remaining recursion budget
//   renderFrame()        main.cpp:347-395        -> rt_render + rt_download_framebuffer
//   yourKeyboardFunc()   raytracing.cpp:453-553  -> same keys, same toggles (GL wireframe / ray debugger are UI, out of scope)
// A failing library call prints rt_last_error() and the functions return black / false: there is no CPU fallback.
#include "raytracing.h"

#include <cstdio>
#include <cstring>
#include <iostream>

#include "flatten.h"
#include "../../include/rt_b200.h"

// raytracing.cpp:15-29
bool Ambient = true, Diffuse = true, Reflection = true, Shadows = true, Specular = true, Refraction = true;
bool DebugMode = false;
bool WireFrame = false;
#define pixelfactor 3
unsigned int pixelfactorX = pixelfactor;
unsigned int pixelfactorY = pixelfactor;
int max_lvl = 10;
std::vector<Vec3Df> normals;
std::vector<Sphere> MySpheres;
int RtGpuCount = 1;

static bool g_rt_ready = false;
bool RtFailed = false;  // set when any library call failed (the app exits non-zero)

static bool rt_ok(int rc, const char* what) {
    if (rc == RT_OK) return true;
    printf("%s failed (%d): %s\n", what, rc, rt_last_error());
    RtFailed = true;
    return false;
}

// false: the frame cannot be described to the library (more lights than RT_MAX_LIGHTS; the reference's 'L' key has no
// limit, main.cpp:334-336) -- an error, never a silently different image.
static bool fill_params(rt_params& p) {
    memset(&p, 0, sizeof(p));
    p.width = WindowSize_X; p.height = WindowSize_Y;
    p.pixelfactor_x = pixelfactorX; p.pixelfactor_y = pixelfactorY;
    p.max_lvl = max_lvl;
    p.features = (Ambient ? RT_AMBIENT : 0) | (Diffuse ? RT_DIFFUSE : 0) | (Specular ? RT_SPECULAR : 0) |
                 (Reflection ? RT_REFLECTION : 0) | (Shadows ? RT_SHADOWS : 0) | (Refraction ? RT_REFRACTION : 0);
    for (int k = 0; k < 3; ++k) p.camera[k] = MyCameraPosition[k];
    p.n_lights = (uint32_t)MyLightPositions.size();
    if (p.n_lights > RT_MAX_LIGHTS) {
        printf("error: %u lights, librt_b200 renders at most %d (RT_MAX_LIGHTS)\n", p.n_lights, RT_MAX_LIGHTS);
        RtFailed = true;
        return false;
    }
    for (uint32_t i = 0; i < p.n_lights; ++i)
        for (int k = 0; k < 3; ++k) p.lights[i][k] = MyLightPositions[i][k];
    return true;
}

void calculateNormals() {  // raytracing.cpp:78-86
    append_face_normals(MyMesh, normals);
}

bool uploadScene() {
    FlatScene flat;
    if (!flatten_mesh(MyMesh, normals, flat)) {
        printf("scene is inconsistent (vertex / material index out of range)\n");
        return false;
    }
    // spheres carry their own Material: append it to the table
    for (size_t i = 0; i < MySpheres.size(); ++i) {
        rt_sphere s;
        memset(&s, 0, sizeof(s));
        for (int k = 0; k < 3; ++k) s.center[k] = MySpheres[i].getCenter()[k];
        s.radius = MySpheres[i].getRadius();
        s.material = (uint32_t)flat.materials.size();
        flat.materials.push_back(flatten_material(MySpheres[i].getMaterial()));
        flat.spheres.push_back(s);
    }
    if (!g_rt_ready) {
        if (!rt_ok(rt_init(RtGpuCount), "rt_init")) return false;
        g_rt_ready = true;
    }
    rt_scene view = flat.view();
    return rt_ok(rt_upload_scene(&view), "rt_upload_scene");
}

void init(char* fileName) {  // raytracing.cpp:42-73
    normals.clear();
    MyMesh.triangleMaterials.clear();
    if (!MyMesh.loadMesh(fileName, true)) {
        printf("cannot load %s\n", fileName);
        return;
    }
    MyMesh.computeVertexNormals();
    calculateNormals();
    // one light at the position the camera starts from (raytracing.cpp:72)
    MyLightPositions.push_back(MyCameraPosition);
    uploadScene();
}

Material getMaterial(int index) {  // raytracing.cpp:373-376
    return MyMesh.materials[MyMesh.triangleMaterials[index]];
}

bool performRayTracingBatch(int n, const float* origins, const float* dests, float* rgb, int* prim_id) {
    rt_params p;
    if (!fill_params(p)) return false;
    return rt_ok(rt_trace(&p, n, origins, dests, rgb, prim_id, nullptr), "rt_trace");
}

Vec3Df trace(const Vec3Df& origin, const Vec3Df& dest, int lvl) {  // raytracing.cpp:381-406
    // a ray that starts at level `lvl` may still spawn max_lvl - lvl continuation rays
    rt_params p;
    if (!fill_params(p)) return Vec3Df(0, 0, 0);
    p.max_lvl = max_lvl - lvl;
    if (p.max_lvl < 0) { p.max_lvl = 0; p.features &= ~(RT_REFLECTION | RT_REFRACTION); }
    float rgb[3] = {0, 0, 0};
    if (!rt_ok(rt_trace(&p, 1, origin.pointer(), dest.pointer(), rgb, nullptr, nullptr), "rt_trace")) return Vec3Df(0, 0, 0);
    return Vec3Df(rgb[0], rgb[1], rgb[2]);
}

Vec3Df performRayTracing(const Vec3Df& origin, const Vec3Df& dest) {  // raytracing.cpp:410-416
    return trace(origin, dest, 0);
}

bool renderFrame(const Vec3Df origin[4], const Vec3Df dest[4], float* rgb) {
    rt_params p;
    if (!fill_params(p)) return false;
    // corner order of main.cpp:355-358: (0,0) (0,H-1) (W-1,0) (W-1,H-1)
    for (int c = 0; c < 4; ++c)
        for (int k = 0; k < 3; ++k) { p.corners[c * 6 + k] = origin[c][k]; p.corners[c * 6 + 3 + k] = dest[c][k]; }
    if (!rt_ok(rt_render(&p), "rt_render")) return false;
    return rt_ok(rt_download_framebuffer(rgb, nullptr), "rt_download_framebuffer");
}

void yourDebugDraw() {}  // GL ray debugger: UI, out of scope

void yourKeyboardFunc(char key, int x, int y) {  // raytracing.cpp:453-553
    switch (key) {
        case '1': Ambient = !Ambient; break;
        case '2': Diffuse = !Diffuse; break;
        case '3': Specular = !Specular; break;
        case '4': Reflection = !Reflection; break;
        case '5': Shadows = !Shadows; break;
        case '6': Refraction = !Refraction; break;
        case '+': pixelfactorX++; pixelfactorY++; break;
        case '-':
            pixelfactorX--; pixelfactorY--;
            if (pixelfactorX < 1) pixelfactorX = 1;
            if (pixelfactorY < 1) pixelfactorY = 1;
            break;
        case '0':
            DebugMode = !DebugMode;
            std::cout << "Debug Mode:\n 0 to enable / disable debug mode\n d to shoot a ray trace.\n";
            break;
        case 'd':  // shoot one ray at the mouse position
            if (DebugMode) {
                Vec3Df origin, dest;
                produceRay(x, y, origin, dest);
                Vec3Df pixelcolor = trace(origin, dest, 0);
                std::cout << "Ray trace color = " << pixelcolor << std::endl;
            }
            return;
        case 'w': WireFrame = !WireFrame; break;  // GL preview only; kept as a flag
        default: break;
    }
    std::cout << std::endl << "------SETTINGS------" << std::endl
              << "Ammbient " << (Ambient ? "ON" : "OFF") << std::endl
              << "Diffuse " << (Diffuse ? "ON" : "OFF") << std::endl
              << "Specular " << (Specular ? "ON" : "OFF") << std::endl
              << "Reflection " << (Reflection ? "ON" : "OFF") << std::endl
              << "Shadow " << (Shadows ? "ON" : "OFF") << std::endl
              << "Refraction " << (Refraction ? "ON" : "OFF") << std::endl
              << "pixelfactorX = " << pixelfactorX << std::endl
              << "pixelfactorY = " << pixelfactorY << std::endl
              << "DebugMode " << (DebugMode ? "ON" : "OFF") << std::endl
              << "--------------------" << std::endl;
}
