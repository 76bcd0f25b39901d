// Sphere -- analytic sphere primitive with the surface of the reference's Sphere.h (CG_Project/Sphere.h:14-55:
// Sphere(center, radius, material), getNormalAt(intersection), findIntersection(origin, destination)).
//
// The reference's file is orphaned (included nowhere, not in code.pro, does not compile: missing ';',
// uses Material without mesh.h) and its arithmetic is not a usable definition (no sqrt of the discriminant,
// center.center where ray.ray is meant) -- SURVEY 8a-S, "parity unpinned".  The semantics here are this
// repo's own and are the SAME in oracle/rt_oracle.c (ray_intersect_sphere) and on the GPU
// (csrc/rt_common.cuh exact_ray_sphere): nearest root t > 1e-4 of |O + t*d - C| = r with
// d = normalize(destination - origin), float arithmetic in this exact order; miss => 0.
#pragma once
#include <cmath>
#include "Vec3D.h"
#include "mesh.h"

class Sphere {
public:
    Sphere() : c_(0, 0, 0), r_(1) {}
    Sphere(Vec3Df _center, float _radius, Material _material) : c_(_center), r_(_radius), m_(_material) {}
    virtual ~Sphere() {}

    const Vec3Df& getCenter() const { return c_; }
    float getRadius() const { return r_; }
    const Material& getMaterial() const { return m_; }

    // Unit outward normal at a surface point.
    virtual Vec3Df getNormalAt(Vec3Df& intersection) {
        Vec3Df n = intersection - c_;
        n.normalize();
        return n;
    }

    // Distance along the unit direction origin -> destination to the nearest intersection, 0 on a miss
    // (float arithmetic in this exact order: it is the definition the oracle and the GPU share).
    virtual float findIntersection(Vec3Df& origin, Vec3Df& destination) {
        Vec3Df d = destination - origin;
        d.normalize();
        const Vec3Df oc = origin - c_;
        const float half_b = Vec3Df::dotProduct(oc, d);
        const float c = Vec3Df::dotProduct(oc, oc) - r_ * r_;
        const float disc = half_b * half_b - c;
        if (disc < 0) return 0;
        const float root = (float)std::sqrt((double)disc);
        float t = -half_b - root;
        if (!(t > 1e-4f)) t = -half_b + root;
        return (t > 1e-4f) ? t : 0;
    }

private:
    Vec3Df c_;
    float r_;
    Material m_;
};
