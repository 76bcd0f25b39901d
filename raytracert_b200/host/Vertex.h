// Vertex {position p, normal n} -- same surface as the reference's Vertex.h:9-23.
// Only p is used by the ray tracer; n is the (preview-only) averaged vertex normal.
#pragma once
#include "Vec3D.h"

class Vertex {
public:
    Vec3Df p;
    Vec3Df n;
    Vertex() {}
    Vertex(const Vec3Df& pos) : p(pos) {}
    Vertex(const Vec3Df& pos, const Vec3Df& nrm) : p(pos), n(nrm) {}
    virtual ~Vertex() {}
};
