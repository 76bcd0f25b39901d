/* rt_oracle.c -- plain-C restatement of the reference's render hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
 * this; the product path (raytracert_b200/csrc, raytracert_b200/host) never does.
 *
 * Pinned: bit-exact (float RGB per sample, per pixel, primary primitive ids, hit points) against
 * oracle/_ref (the unmodified reference sources compiled here) by tests/test_oracle.py on every shipped
 * scene, and against the committed fixtures under tests/golden/ that oracle/_ref generated.  The
 * reference itself ships no golden vectors.
 *
 * Every function names the reference lines it follows (paths under /root/reference/CG_Project).
 * Arithmetic is IEEE binary32, one rounding per operation, in the reference's evaluation order; build
 * with -O2 -ffp-contract=off (oracle/Makefile).  Double appears only where the reference's expression
 * is double: sqrt() of a float inside Vec3D::getLength (Vec3D.h:138-140) and `root >= 0.0` etc.
 *
 * Pinned UB: materials arrive with every scalar defined (the caller resolves the
 * uninitialised Tr/Ni of the reference to 1); a triangle without a known material uses material 0.
 *
 * Extension WITHOUT a reference ("parity unpinned"): analytic spheres.  Sphere.h in the
 * reference is orphaned and does not compile; the semantics implemented here are this repo's own
 * definition (orc_set_spheres) and are what the CUDA path is checked against.
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct { float x, y, z; } v3;

/* ---- Vec3D.h --------------------------------------------------------------------------------- */
static inline v3 V(float x, float y, float z) { v3 r = {x, y, z}; return r; }
static inline v3 vadd(v3 a, v3 b) { return V(a.x + b.x, a.y + b.y, a.z + b.z); }        /* Vec3D.h:24-26 */
static inline v3 vsub(v3 a, v3 b) { return V(a.x - b.x, a.y - b.y, a.z - b.z); }        /* Vec3D.h:28-30 */
static inline v3 vneg(v3 a) { return V(-a.x, -a.y, -a.z); }                              /* Vec3D.h:32-34 */
static inline v3 vscale(v3 a, float s) { return V(a.x * s, a.y * s, a.z * s); }          /* Vec3D.h:12-18 */
static inline v3 vmul(v3 a, v3 b) { return V(a.x * b.x, a.y * b.y, a.z * b.z); }         /* Vec3D.h:20-22 */
static inline v3 vdiv(v3 a, float s) { return V(a.x / s, a.y / s, a.z / s); }            /* Vec3D.h:36-38 */
static inline float vdot(v3 a, v3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }       /* Vec3D.h:192-194 */
static inline v3 vcross(v3 a, v3 b) {                                                    /* Vec3D.h:185-191 */
    return V(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
static inline float vlength(v3 a) { return (float)sqrt((double)vdot(a, a)); }            /* Vec3D.h:135-140 */
static inline void vnormalize(v3* a) {                                                   /* Vec3D.h:142-151 */
    float len = vlength(*a);
    if (len == 0.0f) return;
    float rez = 1.0f / len;
    a->x *= rez; a->y *= rez; a->z *= rez;
}
static inline float vdistance(v3 a, v3 b) { return vlength(vsub(a, b)); }                /* Vec3D.h:199-202 */
static inline float fmax_std(float a, float b) { return (a < b) ? b : a; }               /* std::max */

/* ---- scene (mesh.h:172-201 + raytracing.cpp:33) -------------------------------------------------- */
typedef struct {
    v3 Kd, Ka, Ks;
    float Ns, Ni, Tr;
    int has_Kd, has_Ka, has_Ks, has_Ns, has_Ni, has_Tr;
} material_t;

typedef struct { v3 center; float radius; uint32_t material; } sphere_t;

static int g_nv, g_nt, g_nm, g_ns;
static v3* g_verts;
static uint32_t* g_idx;       /* 3 per triangle */
static uint32_t* g_trimat;
static v3* g_normals;         /* raytracing.cpp:33 */
static material_t* g_mats;
static sphere_t* g_spheres;
static v3 g_camera;           /* MyCameraPosition */
static v3 g_lights[64];
static int g_nlights;
static int Ambient = 1, Diffuse = 1, Reflection = 1, Shadows = 1, Specular = 1, Refraction = 1; /* raytracing.cpp:15-20 */
static int max_lvl = 10;                                                                          /* raytracing.cpp:29 */
static uint64_t g_counts[3];  /* intersectMesh calls: primary, shadow, bounce */

enum { RAY_PRIMARY = 0, RAY_SHADOW = 1, RAY_BOUNCE = 2 };

/* calculateNormals, raytracing.cpp:78-86 */
static void calculate_normals(void) {
    for (int i = 0; i < g_nt; i++) {
        v3 edge01 = vsub(g_verts[g_idx[3 * i + 1]], g_verts[g_idx[3 * i]]);
        v3 edge02 = vsub(g_verts[g_idx[3 * i + 2]], g_verts[g_idx[3 * i]]);
        v3 n = vcross(edge01, edge02);
        vnormalize(&n);
        g_normals[i] = n;
    }
}

/* rayIntersectTriangle, raytracing.cpp:99-154 */
static int ray_intersect_triangle(v3 R0, v3 R1, v3 T0, v3 T1, v3 T2, v3* out) {
    const float SMALL_NUM = 0.00001f;
    v3 u = vsub(T1, T0);
    v3 v = vsub(T2, T0);
    v3 n = vcross(u, v);
    if (n.x == 0 && n.y == 0 && n.z == 0) return 0;       /* :109, isNullVector :92-94 */
    v3 dir = vsub(R1, R0);
    v3 w0 = vsub(R0, T0);
    float b = vdot(n, dir);
    float a = -vdot(n, w0);
    if (fabsf(b) < SMALL_NUM) return 0;                     /* :115 */
    float r = a / b;
    if (r < 0) return 0;                                    /* :125 (no upper bound) */
    v3 I = vadd(R0, vscale(dir, r));                        /* :130  R[0] + r * dir */
    float uu = vdot(u, u), uv = vdot(u, v), vv = vdot(v, v);
    v3 w = vsub(I, T0);
    float wu = vdot(w, u), wv = vdot(w, v);
    float D = uv * uv - uu * vv;
    float s = (uv * wv - vv * wu) / D;
    if (s < 0 || s > 1) return 0;                           /* :145 (NaN passes) */
    float t = (uv * wu - uu * wv) / D;
    if (t < 0 || (s + t) > 1) return 0;                     /* :149 */
    *out = I;
    return 1;
}

/* Sphere extension (no reference; see header): nearest root t >= 1e-4 of |O + t*d - C| = R with
 * d = normalize(dest - origin), all in float; hit point O + t*d; returns the distance |hit - O|. */
static int ray_intersect_sphere(v3 R0, v3 R1, const sphere_t* sp, v3* out) {
    v3 d = vsub(R1, R0);
    vnormalize(&d);
    v3 oc = vsub(R0, sp->center);
    float bq = vdot(oc, d);
    float cq = vdot(oc, oc) - sp->radius * sp->radius;
    float disc = bq * bq - cq;
    if (disc < 0) return 0;
    float sq = (float)sqrt((double)disc);
    float t = -bq - sq;
    if (!(t > 1e-4f)) t = -bq + sq;
    if (!(t > 1e-4f)) return 0;
    *out = vadd(R0, vscale(d, t));
    return 1;
}

/* intersectMesh, raytracing.cpp:161-192.  Returns the primitive index or -1; spheres (extension) are
 * numbered after the triangles: id = n_triangles + sphere index. */
static int intersect_mesh(v3 origin, v3 dest, v3* out, int kind) {
    v3 intersect = V(0, 0, 0);
    int index = -1;
    float dist = FLT_MAX;
#pragma omp atomic
    g_counts[kind]++;
    for (int i = 0; i < g_nt; i++) {
        v3 tmp;
        if (ray_intersect_triangle(origin, dest, g_verts[g_idx[3 * i]], g_verts[g_idx[3 * i + 1]], g_verts[g_idx[3 * i + 2]], &tmp)) {
            float tempDist = vdistance(origin, tmp);        /* :182 */
            if (tempDist < dist) { dist = tempDist; index = i; intersect = tmp; }
        }
    }
    for (int i = 0; i < g_ns; i++) {
        v3 tmp;
        if (ray_intersect_sphere(origin, dest, &g_spheres[i], &tmp)) {
            float tempDist = vdistance(origin, tmp);
            if (tempDist < dist) { dist = tempDist; index = g_nt + i; intersect = tmp; }
        }
    }
    *out = intersect;
    return index;
}

/* getMaterial, raytracing.cpp:373-376 (+ sphere extension) */
static material_t get_material(int index) {
    uint32_t m = (index < g_nt) ? g_trimat[index] : g_spheres[index - g_nt].material;
    return g_mats[m];
}

static v3 trace(v3 origin, v3 dest, int lvl, int kind);

/* diffuseOnly, raytracing.cpp:197-205 (normal is normalised IN PLACE: it is a reference parameter) */
static v3 diffuse_only(v3* normal, const material_t* m, v3 lightpos) {
    v3 c = V(0, 0, 0);
    vnormalize(normal);
    vnormalize(&lightpos);  /* the light POSITION used as a direction (:200-202) */
    c = vadd(c, vscale(m->Kd, fmax_std(vdot(*normal, lightpos), 0.0f)));
    return c;
}

/* blinnPhongSpecularOnly, raytracing.cpp:210-232 */
static v3 blinn_phong_specular_only(v3 P, v3* normal, const material_t* m, v3 lightpos) {
    v3 c = V(0, 0, 0);
    v3 Vv = vsub(g_camera, P);     /* global camera eye, also for bounced rays (:212) */
    vnormalize(normal);
    vnormalize(&Vv);
    v3 L = vsub(lightpos, P);
    vnormalize(&L);
    v3 H = vadd(Vv, L);
    vnormalize(&H);
    float spec = fmax_std(vdot(H, *normal), 0.0f);
    spec = powf(spec, m->Ns);      /* std::pow(float,float) (:226) */
    c = vadd(c, vscale(m->Ks, spec));
    return c;
}

/* isShadow, raytracing.cpp:241-261 */
static int is_shadow(v3 P, v3 light) {
    if (Shadows) {
        v3 tmp;
        P = vadd(P, V(0.1f, 0.1f, 0.1f));                  /* :246 */
        int index = intersect_mesh(P, light, &tmp, RAY_SHADOW);
        if (index == -1) return 0;
        material_t m = get_material(index);
        if (m.has_Tr && m.Tr < 1.0) return 0;               /* :254 */
        return 1;
    }
    return 0;
}

/* addOffset, raytracing.cpp:266-271 */
static void add_offset(v3* point, const v3* towards) {
    v3 d = vsub(*towards, *point);
    vnormalize(&d);
    d = vscale(d, (float)0.01);                             /* vector *= 0.01 with T = float */
    *point = vadd(*point, d);
}

/* reflection, raytracing.cpp:277-285 */
static v3 reflection(v3 ray, v3 P, v3* normal, int lvl) {
    vnormalize(&ray);
    v3 R = vsub(ray, vscale(*normal, 2 * vdot(*normal, ray)));   /* ray - (2*dot*normal) */
    v3 point = P;
    v3 dest = vadd(P, R);
    add_offset(&point, &dest);
    return trace(point, dest, lvl, RAY_BOUNCE);
}

/* refraction, raytracing.cpp:290-330 */
static v3 refraction(v3 ray, v3 P, v3* normal, const material_t* m, int lvl) {
    float ni = m->Ni;
    vnormalize(&ray);
    float check = vdot(ray, *normal);
    if (check < 0) {
        float angle = acosf(check);
        if (angle <= 2 && angle > 0)                        /* :298 grazing hack */
            return vmul(m->Ks, reflection(ray, P, normal, lvl + 1));
        float nr = 1 / ni;
        float dn = vdot(*normal, ray);
        float root = 1 - powf(nr, 2) * (1 - powf(dn, 2));   /* :302 */
        if (root >= 0.0) {
            root = (float)sqrt((double)root);
            v3 T = vsub(vscale(vsub(ray, vscale(*normal, vdot(*normal, ray))), nr), vscale(*normal, root)); /* :307 */
            v3 point = P;
            v3 dest = vadd(P, T);
            add_offset(&point, &dest);
            return vscale(trace(point, dest, lvl + 1, RAY_BOUNCE), 1 - m->Tr);                              /* :311 */
        }
    } else {
        float nr = ni;
        v3 nn = vneg(*normal);
        float root = 1 - powf(nr, 2) * (1 - powf(vdot(nn, ray), 2));                                        /* :316 */
        if (root >= 0.0) {
            root = (float)sqrt((double)root);
            v3 T = vsub(vscale(vsub(ray, vscale(nn, vdot(nn, ray))), nr), vscale(nn, root));                /* :321 */
            v3 point = P;
            v3 dest = vadd(point, T);
            add_offset(&point, &dest);
            return vscale(trace(point, dest, lvl + 1, RAY_BOUNCE), 1 - m->Tr);                              /* :325 */
        }
    }
    return V(0, 0, 0);
}

/* shade, raytracing.cpp:335-368 */
static v3 shade(v3 ray, v3 P, v3* normal, const material_t* m, int lvl) {
    v3 c = V(0, 0, 0);
    if (Ambient && m->has_Ka) c = vadd(c, m->Ka);
    for (int i = 0; i < g_nlights; i++) {
        v3 L = g_lights[i];
        if (!is_shadow(P, L)) {
            if (Diffuse && m->has_Kd) c = vadd(c, vscale(diffuse_only(normal, m, L), m->Tr));
            if (Specular && m->has_Ks && m->has_Ns) c = vadd(c, vscale(blinn_phong_specular_only(P, normal, m, L), m->Tr));
        }
    }
    if (Refraction && (m->Tr < 1) && lvl < max_lvl)
        c = vadd(c, refraction(ray, P, normal, m, lvl + 1));
    else if (Reflection && lvl < max_lvl)
        c = vadd(c, vmul(m->Ks, reflection(ray, P, normal, lvl + 1)));   /* traced even when Ks == 0 */
    return c;
}

/* trace, raytracing.cpp:381-406 */
static v3 trace(v3 origin, v3 dest, int lvl, int kind) {
    v3 I;
    int index = intersect_mesh(origin, dest, &I, kind);
    if (index == -1) return V(0, 0, 0);
    v3 ray = vsub(dest, origin);
    v3 normal;
    if (index < g_nt) normal = g_normals[index];            /* never flipped toward the ray (:394) */
    else { normal = vsub(I, g_spheres[index - g_nt].center); vnormalize(&normal); }  /* Sphere.h:33-37 */
    material_t m = get_material(index);
    return shade(ray, I, &normal, &m, lvl);
}

/* performRayTracing, raytracing.cpp:410-416 */
static v3 perform_ray_tracing(v3 origin, v3 dest) { return trace(origin, dest, 0, RAY_PRIMARY); }

/* ---- C ABI (same shape as oracle/ref_harness.cpp so tests can swap the two) --------------------- */

void orc_set_scene(int nv, const float* verts, int nt, const uint32_t* idx, const uint32_t* mat, int nm, const float* mats) {
    free(g_verts); free(g_idx); free(g_trimat); free(g_normals); free(g_mats);
    g_nv = nv; g_nt = nt; g_nm = nm; g_ns = 0;
    g_verts = (v3*)malloc(sizeof(v3) * (nv > 0 ? nv : 1));
    g_idx = (uint32_t*)malloc(sizeof(uint32_t) * 3 * (nt > 0 ? nt : 1));
    g_trimat = (uint32_t*)malloc(sizeof(uint32_t) * (nt > 0 ? nt : 1));
    g_normals = (v3*)malloc(sizeof(v3) * (nt > 0 ? nt : 1));
    g_mats = (material_t*)malloc(sizeof(material_t) * (nm > 0 ? nm : 1));
    for (int i = 0; i < nv; i++) g_verts[i] = V(verts[3 * i], verts[3 * i + 1], verts[3 * i + 2]);
    memcpy(g_idx, idx, sizeof(uint32_t) * 3 * nt);
    for (int i = 0; i < nt; i++) g_trimat[i] = (mat[i] < (uint32_t)nm) ? mat[i] : 0;
    for (int i = 0; i < nm; i++) {
        const float* m = mats + 16 * i;
        int flags = (int)m[12];
        material_t M;
        M.Kd = V(m[0], m[1], m[2]); M.Ns = m[3];
        M.Ka = V(m[4], m[5], m[6]); M.Ni = m[7];
        M.Ks = V(m[8], m[9], m[10]); M.Tr = m[11];
        M.has_Kd = flags & 1; M.has_Ka = flags & 2; M.has_Ks = flags & 4; M.has_Ns = flags & 8; M.has_Ni = flags & 16; M.has_Tr = flags & 32;
        g_mats[i] = M;
    }
    calculate_normals();
    g_nlights = 0;
}

/* rows: cx cy cz radius material(as float) */
void orc_set_spheres(int n, const float* rows) {
    free(g_spheres);
    g_spheres = (sphere_t*)malloc(sizeof(sphere_t) * (n > 0 ? n : 1));
    g_ns = n;
    for (int i = 0; i < n; i++) {
        g_spheres[i].center = V(rows[5 * i], rows[5 * i + 1], rows[5 * i + 2]);
        g_spheres[i].radius = rows[5 * i + 3];
        g_spheres[i].material = (uint32_t)rows[5 * i + 4];
    }
}

void orc_get_normals(float* out) { for (int i = 0; i < g_nt; i++) { out[3 * i] = g_normals[i].x; out[3 * i + 1] = g_normals[i].y; out[3 * i + 2] = g_normals[i].z; } }
void orc_set_camera(const float* eye) { g_camera = V(eye[0], eye[1], eye[2]); }
void orc_set_lights(int n, const float* xyz) { g_nlights = n > 64 ? 64 : n; for (int i = 0; i < g_nlights; i++) g_lights[i] = V(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]); }
void orc_set_toggles(int ambient, int diffuse, int specular, int reflection_, int shadows, int refraction_) {
    Ambient = ambient; Diffuse = diffuse; Specular = specular; Reflection = reflection_; Shadows = shadows; Refraction = refraction_;
}
void orc_set_max_lvl(int l) { max_lvl = l; }
void orc_reset_counts(void) { g_counts[0] = g_counts[1] = g_counts[2] = 0; }
void orc_get_counts(uint64_t* out) { out[0] = g_counts[0]; out[1] = g_counts[1]; out[2] = g_counts[2]; }

/* Frame loop of the 'r' handler, main.cpp:347-395 (corner order main.cpp:355-358; scales :380-381;
 * bilinear rays :383-386; sample sum subx-outer/suby-inner :377-390; average :391; clamp :29-41). */
void orc_render(const float* c, int W, int H, int pfX, int pfY, int y0, int ystep, int x0, int xstep, float* rgb, float* sample_rgb, int32_t* sample_prim, int nthreads) {
    v3 origin00 = V(c[0], c[1], c[2]), dest00 = V(c[3], c[4], c[5]);
    v3 origin01 = V(c[6], c[7], c[8]), dest01 = V(c[9], c[10], c[11]);
    v3 origin10 = V(c[12], c[13], c[14]), dest10 = V(c[15], c[16], c[17]);
    v3 origin11 = V(c[18], c[19], c[20]), dest11 = V(c[21], c[22], c[23]);
    unsigned int pixelfactorX = pfX, pixelfactorY = pfY, WindowSize_X = W, WindowSize_Y = H;
    float divX = (WindowSize_X * pixelfactorX - 1);
    float divY = (WindowSize_Y * pixelfactorY - 1);
    int raysPerPixel = (pixelfactorX * pixelfactorY);
    if (ystep < 1) ystep = 1;
    if (xstep < 1) xstep = 1;
    /* pixel lattice y = y0 + i*ystep, x = x0 + j*xstep (the full frame for 0,1,0,1); threads share the pixel list */
    const long nrows = (H - y0 + ystep - 1) / ystep, ncols = (W - x0 + xstep - 1) / xstep;
    if (nrows <= 0 || ncols <= 0) return;
    (void)nthreads;
#pragma omp parallel for schedule(dynamic, 16) num_threads(nthreads)
    for (long item = 0; item < nrows * ncols; item++) {
        unsigned int y = y0 + (unsigned int)(item / ncols) * ystep;
        unsigned int x = x0 + (unsigned int)(item % ncols) * xstep;
        v3 acc = V(0, 0, 0);
        for (int subx = 0; subx < (int)pixelfactorX; subx++) {
            for (int suby = 0; suby < (int)pixelfactorY; suby++) {
                float xscale = 1.0f - ((float)x * pixelfactorX + subx) / divX;
                float yscale = 1.0f - ((float)y * pixelfactorY + suby) / divY;
                v3 origin = vadd(vscale(vadd(vscale(origin00, xscale), vscale(origin10, 1 - xscale)), yscale),
                                 vscale(vadd(vscale(origin01, xscale), vscale(origin11, 1 - xscale)), 1 - yscale));
                v3 dest = vadd(vscale(vadd(vscale(dest00, xscale), vscale(dest10, 1 - xscale)), yscale),
                               vscale(vadd(vscale(dest01, xscale), vscale(dest11, 1 - xscale)), 1 - yscale));
                v3 col = perform_ray_tracing(origin, dest);
                acc = vadd(acc, col);
                size_t s = (((size_t)y * W + x) * pfX + subx) * pfY + suby;
                if (sample_rgb) { sample_rgb[3 * s] = col.x; sample_rgb[3 * s + 1] = col.y; sample_rgb[3 * s + 2] = col.z; }
                if (sample_prim) {
                    v3 tmp;
                    sample_prim[s] = intersect_mesh(origin, dest, &tmp, RAY_PRIMARY);
#pragma omp atomic
                    g_counts[RAY_PRIMARY]--;   /* the id query is not one of the reference's rays */
                }
            }
        }
        acc = vdiv(acc, raysPerPixel);
        float ch[3] = {acc.x, acc.y, acc.z};
        for (int k = 0; k < 3; k++) {
            if (ch[k] > 1) ch[k] = 1.0f;
            if (ch[k] < 0) ch[k] = 0.0f;
            rgb[3 * ((size_t)W * y + x) + k] = ch[k];
        }
    }
}

void orc_trace(int n, const float* origins, const float* dests, float* rgb, int32_t* prim, float* hit) {
    for (int i = 0; i < n; i++) {
        v3 o = V(origins[3 * i], origins[3 * i + 1], origins[3 * i + 2]);
        v3 d = V(dests[3 * i], dests[3 * i + 1], dests[3 * i + 2]);
        v3 col = perform_ray_tracing(o, d);
        rgb[3 * i] = col.x; rgb[3 * i + 1] = col.y; rgb[3 * i + 2] = col.z;
        if (prim || hit) {
            v3 I;
            int id = intersect_mesh(o, d, &I, RAY_PRIMARY);
            g_counts[RAY_PRIMARY]--;
            if (prim) prim[i] = id;
            if (hit) { hit[3 * i] = I.x; hit[3 * i + 1] = I.y; hit[3 * i + 2] = I.z; }
        }
    }
}

/* One (ray, triangle) pair: rayIntersectTriangle + the distance intersectMesh compares (raytracing.cpp:179-183).
 * Returns 1 on a hit and writes Vec3Df::distance(origin, I).  Used by the filter soundness tests. */
int orc_ray_triangle(const float* R0, const float* R1, const float* T0, const float* T1, const float* T2, float* dist) {
    v3 I;
    v3 o = V(R0[0], R0[1], R0[2]);
    if (!ray_intersect_triangle(o, V(R1[0], R1[1], R1[2]), V(T0[0], T0[1], T0[2]), V(T1[0], T1[1], T1[2]), V(T2[0], T2[1], T2[2]), &I)) return 0;
    *dist = vdistance(o, I);
    return 1;
}

/* Image::writeImage quantiser, main.cpp:116-117 */
void orc_quantise(const float* rgb, int n, unsigned char* out) {
    for (int i = 0; i < n; i++) out[i] = (unsigned char)(rgb[i] * 255.0f);
}
