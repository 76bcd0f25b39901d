// Headless stand-ins for the GL matrix stack + GLU calls the reference's camera path uses:
// gluPerspective (main.cpp:294), glTranslatef (main.cpp:219), gluUnProject (main.cpp:312,316).
// Those live in the system's OpenGL/GLU (un-vendored, no version pinned -- SURVEY 8c); this is a
// restatement of the published SGI/Mesa algorithm in double precision: column-major 4x4 matrices,
// P*M product, general inverse by cofactors, NDC -> object space, divide by w.
// The 24 corner floats produced here are an INPUT to both the oracle and the GPU path, so this file
// cannot cause a parity mismatch; it only has to be a sane camera.
#pragma once
#include <cmath>
#include <cstring>

namespace glu {

struct Mat4 { double m[16]; };  // column-major, like glGetDoublev

inline Mat4 identity() { Mat4 r; for (int i = 0; i < 16; ++i) r.m[i] = (i % 5 == 0) ? 1.0 : 0.0; return r; }

// r = a * b (apply b first), column-major.
inline Mat4 mul(const Mat4& a, const Mat4& b) {
    Mat4 r;
    for (int c = 0; c < 4; ++c)
        for (int row = 0; row < 4; ++row) {
            double s = 0.0;
            for (int k = 0; k < 4; ++k) s += a.m[k * 4 + row] * b.m[c * 4 + k];
            r.m[c * 4 + row] = s;
        }
    return r;
}

inline Mat4 translate(double x, double y, double z) { Mat4 r = identity(); r.m[12] = x; r.m[13] = y; r.m[14] = z; return r; }

inline Mat4 perspective(double fovy_deg, double aspect, double z_near, double z_far) {
    Mat4 r; memset(r.m, 0, sizeof(r.m));
    const double half = fovy_deg / 2.0 * M_PI / 180.0;
    const double cot = std::cos(half) / std::sin(half);
    r.m[0] = cot / aspect;
    r.m[5] = cot;
    r.m[10] = -(z_far + z_near) / (z_far - z_near);
    r.m[11] = -1.0;
    r.m[14] = -2.0 * z_near * z_far / (z_far - z_near);
    return r;
}

inline Mat4 look_at(const double eye[3], const double center[3], const double up[3]) {
    double f[3] = {center[0] - eye[0], center[1] - eye[1], center[2] - eye[2]};
    double fl = std::sqrt(f[0] * f[0] + f[1] * f[1] + f[2] * f[2]);
    for (int i = 0; i < 3; ++i) f[i] /= fl;
    double s[3] = {f[1] * up[2] - f[2] * up[1], f[2] * up[0] - f[0] * up[2], f[0] * up[1] - f[1] * up[0]};
    double sl = std::sqrt(s[0] * s[0] + s[1] * s[1] + s[2] * s[2]);
    for (int i = 0; i < 3; ++i) s[i] /= sl;
    double u[3] = {s[1] * f[2] - s[2] * f[1], s[2] * f[0] - s[0] * f[2], s[0] * f[1] - s[1] * f[0]};
    Mat4 r = identity();
    r.m[0] = s[0]; r.m[4] = s[1]; r.m[8] = s[2];
    r.m[1] = u[0]; r.m[5] = u[1]; r.m[9] = u[2];
    r.m[2] = -f[0]; r.m[6] = -f[1]; r.m[10] = -f[2];
    return mul(r, translate(-eye[0], -eye[1], -eye[2]));
}

// General 4x4 inverse (cofactor expansion). Returns false for a singular matrix.
inline bool invert(const Mat4& a, Mat4& out) {
    const double* m = a.m;
    double inv[16];
    inv[0] = m[5] * m[10] * m[15] - m[5] * m[11] * m[14] - m[9] * m[6] * m[15] + m[9] * m[7] * m[14] + m[13] * m[6] * m[11] - m[13] * m[7] * m[10];
    inv[4] = -m[4] * m[10] * m[15] + m[4] * m[11] * m[14] + m[8] * m[6] * m[15] - m[8] * m[7] * m[14] - m[12] * m[6] * m[11] + m[12] * m[7] * m[10];
    inv[8] = m[4] * m[9] * m[15] - m[4] * m[11] * m[13] - m[8] * m[5] * m[15] + m[8] * m[7] * m[13] + m[12] * m[5] * m[11] - m[12] * m[7] * m[9];
    inv[12] = -m[4] * m[9] * m[14] + m[4] * m[10] * m[13] + m[8] * m[5] * m[14] - m[8] * m[6] * m[13] - m[12] * m[5] * m[10] + m[12] * m[6] * m[9];
    inv[1] = -m[1] * m[10] * m[15] + m[1] * m[11] * m[14] + m[9] * m[2] * m[15] - m[9] * m[3] * m[14] - m[13] * m[2] * m[11] + m[13] * m[3] * m[10];
    inv[5] = m[0] * m[10] * m[15] - m[0] * m[11] * m[14] - m[8] * m[2] * m[15] + m[8] * m[3] * m[14] + m[12] * m[2] * m[11] - m[12] * m[3] * m[10];
    inv[9] = -m[0] * m[9] * m[15] + m[0] * m[11] * m[13] + m[8] * m[1] * m[15] - m[8] * m[3] * m[13] - m[12] * m[1] * m[11] + m[12] * m[3] * m[9];
    inv[13] = m[0] * m[9] * m[14] - m[0] * m[10] * m[13] - m[8] * m[1] * m[14] + m[8] * m[2] * m[13] + m[12] * m[1] * m[10] - m[12] * m[2] * m[9];
    inv[2] = m[1] * m[6] * m[15] - m[1] * m[7] * m[14] - m[5] * m[2] * m[15] + m[5] * m[3] * m[14] + m[13] * m[2] * m[7] - m[13] * m[3] * m[6];
    inv[6] = -m[0] * m[6] * m[15] + m[0] * m[7] * m[14] + m[4] * m[2] * m[15] - m[4] * m[3] * m[14] - m[12] * m[2] * m[7] + m[12] * m[3] * m[6];
    inv[10] = m[0] * m[5] * m[15] - m[0] * m[7] * m[13] - m[4] * m[1] * m[15] + m[4] * m[3] * m[13] + m[12] * m[1] * m[7] - m[12] * m[3] * m[5];
    inv[14] = -m[0] * m[5] * m[14] + m[0] * m[6] * m[13] + m[4] * m[1] * m[14] - m[4] * m[2] * m[13] - m[12] * m[1] * m[6] + m[12] * m[2] * m[5];
    inv[3] = -m[1] * m[6] * m[11] + m[1] * m[7] * m[10] + m[5] * m[2] * m[11] - m[5] * m[3] * m[10] - m[9] * m[2] * m[7] + m[9] * m[3] * m[6];
    inv[7] = m[0] * m[6] * m[11] - m[0] * m[7] * m[10] - m[4] * m[2] * m[11] + m[4] * m[3] * m[10] + m[8] * m[2] * m[7] - m[8] * m[3] * m[6];
    inv[11] = -m[0] * m[5] * m[11] + m[0] * m[7] * m[9] + m[4] * m[1] * m[11] - m[4] * m[3] * m[9] - m[8] * m[1] * m[7] + m[8] * m[3] * m[5];
    inv[15] = m[0] * m[5] * m[10] - m[0] * m[6] * m[9] - m[4] * m[1] * m[10] + m[4] * m[2] * m[9] + m[8] * m[1] * m[6] - m[8] * m[2] * m[5];
    double det = m[0] * inv[0] + m[1] * inv[4] + m[2] * inv[8] + m[3] * inv[12];
    if (det == 0.0) return false;
    det = 1.0 / det;
    for (int i = 0; i < 16; ++i) out.m[i] = inv[i] * det;
    return true;
}

// gluUnProject: window (winx, winy, winz) -> object space.
inline bool unproject(double winx, double winy, double winz, const Mat4& model, const Mat4& proj, const int viewport[4], double out[3]) {
    Mat4 inv;
    if (!invert(mul(proj, model), inv)) return false;
    double in[4] = {(winx - viewport[0]) / viewport[2] * 2.0 - 1.0, (winy - viewport[1]) / viewport[3] * 2.0 - 1.0, winz * 2.0 - 1.0, 1.0};
    double r[4];
    for (int row = 0; row < 4; ++row) r[row] = inv.m[row] * in[0] + inv.m[4 + row] * in[1] + inv.m[8 + row] * in[2] + inv.m[12 + row] * in[3];
    if (r[3] == 0.0) return false;
    out[0] = r[0] / r[3]; out[1] = r[1] / r[3]; out[2] = r[2] / r[3];
    return true;
}

// Camera position the way traqueboule.h:209-219 gets it: inverse(modelview) * (0,0,0,1).
inline bool camera_position(const Mat4& model, double out[3]) {
    Mat4 inv;
    if (!invert(model, inv)) return false;
    out[0] = inv.m[12] / inv.m[15]; out[1] = inv.m[13] / inv.m[15]; out[2] = inv.m[14] / inv.m[15];
    return true;
}

}  // namespace glu
