// main.cpp -- headless skeleton of the TI1805 ray tracer: the reference's main.cpp (CG_Project/main.cpp)
// without GLUT.  It owns the same globals (main.cpp:17-18,130,137-141), keeps produceRay() (main.cpp:300-320,
// gluUnProject at depth 0 and 1, here through glu_math.h), the keyboard() dispatch ('L', 'l', 'r', then
// yourKeyboardFunc -- main.cpp:328-418) and the Image/PPM writer (image.h), but the 'r' handler hands the
// frame to the GPU: four produceRay() corner calls -> renderFrame() (rt_render + rt_download_framebuffer)
// -> Image -> writeImage("result.ppm").
//
//   rt_main [scene.obj] [--size WxH] [--pf N] [--lvl N] [--eye x,y,z --center x,y,z] [--light x,y,z]...
//           [--sphere cx,cy,cz,r,material_index]... [--gpus N] [--cull] [--out result.ppm] [--keys STRING]
//
// --keys replays key presses in order (default "r"); e.g. --keys "5r" renders with shadows toggled off,
// "Lr" adds a light at the camera first.  Without --eye the camera is the reference's start-up pose:
// modelview = T(0,0,-4), eye (0,0,4) (main.cpp:217-222).
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <string>

#include "glu_math.h"
#include "image.h"
#include "raytracing.h"
#include "../../include/rt_b200.h"

// ---- globals the skeleton owns (main.cpp:17-18,130,137-141) --------------------------------------
Vec3Df MyCameraPosition;
std::vector<Vec3Df> MyLightPositions;
Mesh MyMesh;
unsigned int WindowSize_X = 500;
unsigned int WindowSize_Y = 500;
unsigned int RayTracingResolutionX = 500;
unsigned int RayTracingResolutionY = 500;

// ---- the GL state produceRay() reads (glGetDoublev / glGetIntegerv in the reference) -----------------
static glu::Mat4 g_modelview = glu::translate(0, 0, -4);
static glu::Mat4 g_projection;
static int g_viewport[4] = {0, 0, 500, 500};
static std::string g_out = "result.ppm";

static Vec3Df getCameraPosition() {  // traqueboule.h:209-219: inverse(modelview) * origin
    double e[3] = {0, 0, 0};
    glu::camera_position(g_modelview, e);
    return Vec3Df(float(e[0]), float(e[1]), float(e[2]));
}

void produceRay(int x_I, int y_I, Vec3Df* origin, Vec3Df* dest) {  // main.cpp:300-320
    int y_new = g_viewport[3] - y_I;
    double p[3];
    glu::unproject(x_I, y_new, 0, g_modelview, g_projection, g_viewport, p);
    origin->p[0] = float(p[0]); origin->p[1] = float(p[1]); origin->p[2] = float(p[2]);
    glu::unproject(x_I, y_new, 1, g_modelview, g_projection, g_viewport, p);
    dest->p[0] = float(p[0]); dest->p[1] = float(p[1]); dest->p[2] = float(p[2]);
}
void produceRay(int x_I, int y_I, Vec3Df& origin, Vec3Df& dest) { produceRay(x_I, y_I, &origin, &dest); }  // main.cpp:322-325

void keyboard(unsigned char key, int x, int y) {  // main.cpp:328-418
    printf("key %d pressed at %d,%d\n", key, x, y);
    fflush(stdout);
    switch (key) {
        case 'L': MyLightPositions.push_back(getCameraPosition()); break;
        case 'l': if (!MyLightPositions.empty()) MyLightPositions[MyLightPositions.size() - 1] = getCameraPosition(); break;
        case 'r': {
            std::cout << "Raytracing" << std::endl;
            Image result(WindowSize_X, WindowSize_Y);
            Vec3Df origin[4], dest[4];
            produceRay(0, 0, &origin[0], &dest[0]);
            produceRay(0, WindowSize_Y - 1, &origin[1], &dest[1]);
            produceRay(WindowSize_X - 1, 0, &origin[2], &dest[2]);
            produceRay(WindowSize_X - 1, WindowSize_Y - 1, &origin[3], &dest[3]);
            auto t0 = std::chrono::steady_clock::now();
            if (!renderFrame(origin, dest, result._image.data())) { printf("render failed: no image written\n"); break; }
            double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
            rt_stats st;
            if (rt_get_stats(&st) == RT_OK) {
                double rays = double(st.primary_rays + st.shadow_rays + st.bounce_rays);
                printf("%ux%u, %ux%u rays/pixel, %u triangles, %d GPU(s): %.2f ms (device %.2f ms), %.0f rays "
                       "(%llu primary, %llu shadow, %llu bounce), %.1f Mrays/s\n",
                       WindowSize_X, WindowSize_Y, pixelfactorX, pixelfactorY, st.n_triangles, (int)st.n_gpus, ms, st.ms_total, rays,
                       (unsigned long long)st.primary_rays, (unsigned long long)st.shadow_rays, (unsigned long long)st.bounce_rays,
                       rays / (st.ms_total > 0 ? st.ms_total : ms) / 1e3);
            }
            result.writeImage(g_out.c_str());
            break;
        }
        case 27: exit(0);
    }
    yourKeyboardFunc(key, x, y);
}

static bool parse_floats(const char* s, float* out, int n) {
    for (int i = 0; i < n; ++i) {
        char* end = nullptr;
        out[i] = strtof(s, &end);
        if (end == s) return false;
        s = (*end == ',') ? end + 1 : end;
    }
    return true;
}

int main(int argc, char** argv) {
    std::string scene = "cube.obj", keys = "r";
    float eye[3], center[3] = {0, 0, 0};
    bool have_eye = false, cull = false;
    std::vector<Vec3Df> lights;
    struct SphereArg { float v[5]; };
    std::vector<SphereArg> spheres;
    for (int i = 1; i < argc; ++i) {
        std::string a = argv[i];
        auto next = [&](const char* what) -> const char* {
            if (i + 1 >= argc) { printf("%s needs a value\n", what); exit(2); }
            return argv[++i];
        };
        if (a == "--size") { if (sscanf(next("--size"), "%ux%u", &WindowSize_X, &WindowSize_Y) != 2) { printf("--size WxH\n"); return 2; } }
        else if (a == "--pf") { pixelfactorX = pixelfactorY = (unsigned)atoi(next("--pf")); }
        else if (a == "--lvl") { max_lvl = atoi(next("--lvl")); }
        else if (a == "--eye") { if (!parse_floats(next("--eye"), eye, 3)) return 2; have_eye = true; }
        else if (a == "--center") { if (!parse_floats(next("--center"), center, 3)) return 2; }
        else if (a == "--light") { float l[3]; if (!parse_floats(next("--light"), l, 3)) return 2; lights.push_back(Vec3Df(l[0], l[1], l[2])); }
        else if (a == "--sphere") { SphereArg s; if (!parse_floats(next("--sphere"), s.v, 5)) return 2; spheres.push_back(s); }
        else if (a == "--gpus") { RtGpuCount = atoi(next("--gpus")); }
        else if (a == "--cull") { cull = true; }
        else if (a == "--out") { g_out = next("--out"); }
        else if (a == "--keys") { keys = next("--keys"); }
        else if (a[0] == '-') { printf("unknown option %s\n", a.c_str()); return 2; }
        else scene = a;
    }
    if (pixelfactorX < 1) pixelfactorX = pixelfactorY = 1;
    RayTracingResolutionX = WindowSize_X; RayTracingResolutionY = WindowSize_Y;

    // GL state of main(): modelview (main.cpp:217-219) and reshape()'s projection (main.cpp:288-296)
    g_viewport[2] = (int)WindowSize_X; g_viewport[3] = (int)WindowSize_Y;
    g_projection = glu::perspective(50, (float)WindowSize_X / WindowSize_Y, 1, 10);
    if (have_eye) {
        const double e[3] = {eye[0], eye[1], eye[2]}, c[3] = {center[0], center[1], center[2]}, up[3] = {0, 1, 0};
        g_modelview = glu::look_at(e, c, up);
    }
    MyCameraPosition = getCameraPosition();  // main.cpp:222

    std::vector<char> name(scene.begin(), scene.end());
    name.push_back(0);
    // options persist across rt_init, and tile culling wants to be set BEFORE the upload (spatially sorted tiles)
    if (cull && rt_set_option(RT_OPT_TILE_CULLING, 1) != RT_OK) { printf("%s\n", rt_last_error()); return 1; }  // same image, fewer tests
    init(name.data());  // main.cpp:258
    if (MyMesh.triangles.empty() && spheres.empty()) { printf("no geometry loaded from %s\n", scene.c_str()); return 1; }
    if (!lights.empty()) MyLightPositions = lights;  // replaces the start-up light at the eye
    if (!spheres.empty()) {
        for (const SphereArg& s : spheres) {
            size_t m = (size_t)s.v[4];
            if (m >= MyMesh.materials.size()) { printf("--sphere: material %zu does not exist\n", m); return 2; }
            MySpheres.push_back(Sphere(Vec3Df(s.v[0], s.v[1], s.v[2]), s.v[3], MyMesh.materials[m]));
        }
        if (!uploadScene()) return 1;
    }
    for (char k : keys) keyboard((unsigned char)k, 0, 0);
    rt_shutdown();
    return RtFailed ? 1 : 0;
}
