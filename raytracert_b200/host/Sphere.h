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
    Vec3Df center;
    float radius;
    Material material;

public:
    Sphere() : center(0, 0, 0), radius(1) {}
    Sphere(Vec3Df _center, float _radius, Material _material) : center(_center), radius(_radius), material(_material) {}
    virtual ~Sphere() {}

    const Vec3Df& getCenter() const { return center; }
    float getRadius() const { return radius; }
    const Material& getMaterial() const { return material; }

    virtual Vec3Df getNormalAt(Vec3Df& intersection) {
        Vec3Df result = intersection - center;
        result.normalize();
        return result;
    }

    // Distance along the unit direction origin -> destination to the nearest intersection, 0 on a miss.
    virtual float findIntersection(Vec3Df& origin, Vec3Df& destination) {
        Vec3Df d = destination - origin;
        d.normalize();
        Vec3Df oc = origin - center;
        float bq = Vec3Df::dotProduct(oc, d);
        float cq = Vec3Df::dotProduct(oc, oc) - radius * radius;
        float disc = bq * bq - cq;
        if (disc < 0) return 0;
        float sq = (float)std::sqrt((double)disc);
        float t = -bq - sq;
        if (!(t > 1e-4f)) t = -bq + sq;
        if (!(t > 1e-4f)) return 0;
        return t;
    }
};
