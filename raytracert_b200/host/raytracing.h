// The TI1805 skeleton <-> plug-in contract, same names and meaning as the reference's raytracing.h
// (CG_Project/raytracing.h:8-41).  main.cpp owns the globals; raytracing.cpp implements the functions --
// here by driving librt_b200 (include/rt_b200.h) instead of tracing on the CPU.  There is no CPU path.
#pragma once
#include <vector>
#include "Sphere.h"
#include "mesh.h"

extern Mesh MyMesh;                            // main mesh
extern std::vector<Vec3Df> MyLightPositions;
extern Vec3Df MyCameraPosition;                // current camera eye
extern unsigned int WindowSize_X;              // image width
extern unsigned int WindowSize_Y;              // image height
extern unsigned int RayTracingResolutionX;
extern unsigned int RayTracingResolutionY;
extern unsigned int pixelfactorX;              // sub-samples per pixel in x
extern unsigned int pixelfactorY;              // sub-samples per pixel in y

// raytracing.cpp:15-29 -- the plug-in's own knobs (keys '1'..'6'), now fields of rt_params
extern bool Ambient, Diffuse, Specular, Reflection, Shadows, Refraction, DebugMode;
extern int max_lvl;
extern std::vector<Vec3Df> normals;            // face normals (raytracing.cpp:33)
extern std::vector<Sphere> MySpheres;          // analytic spheres added to the scene (extension)
extern bool RtFailed;                         // a librt_b200 call failed (there is no CPU fallback)
extern int RtGpuCount;                         // GPUs rt_init() is asked for (default 1)

// Loads the mesh, computes the face normals and uploads the scene to the GPU(s) (rt_init + rt_upload_scene).
void init(char* fileName);
// Re-flatten MyMesh/normals/MySpheres and upload again (after the host changed the scene).
bool uploadScene();

// Defined by the skeleton (main.cpp): window pixel -> ray origin (near plane) and destination (far plane).
void produceRay(int x_I, int y_I, Vec3Df& origin, Vec3Df& dest);

Material getMaterial(int index);

// One ray through the GPU path (rt_trace, batch of 1).  `lvl` is the recursion level the ray starts at.
Vec3Df trace(const Vec3Df& origin, const Vec3Df& dest, int lvl);
Vec3Df performRayTracing(const Vec3Df& origin, const Vec3Df& dest);
// Batch form of performRayTracing: n rays, 3 floats each; rgb out (3*n). Returns false on a library error.
bool performRayTracingBatch(int n, const float* origins, const float* dests, float* rgb, int* prim_id);

// The body of the 'r' key handler (main.cpp:347-395 of the reference): one rt_render + download.
// rgb: 3*WindowSize_X*WindowSize_Y floats, already clamped like RGBValue.
bool renderFrame(const Vec3Df origin[4], const Vec3Df dest[4], float* rgb);

void yourDebugDraw();
void yourKeyboardFunc(char key, int x, int y);
void calculateNormals();
