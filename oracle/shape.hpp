// ORACLE — TEST INFRASTRUCTURE ONLY (see geometry.hpp).
// Restatement of src/shape.rs, src/primitive.rs, src/intersection.rs.
#pragma once
#include "geometry.hpp"
#include "sampling.hpp"

namespace orc {

struct ShapeIntersection {  // intersection.rs:9-14 (+ the Moeller-Trumbore u,v kept for the S3 parity check)
    V3 location, normal;
    double uv[2];
    double bary_u, bary_v;
};

enum ShapeKind { SPHERE = 0, TRIANGLE = 1, DISK = 2 };

struct Shape {  // shape.rs:24-47
    int kind;
    // Sphere / Disk
    Transformation object_to_world, world_to_object;
    double radius, inner_radius;
    // Triangle
    V3 v0, e1, e2, n0, n01, n02;
    double uv0[2], uv01[2], uv02[2];

    static Shape new_sphere(V3 origin, double radius) {  // shape.rs:56-71
        Shape s{};
        s.kind = SPHERE;
        s.radius = radius;
        s.object_to_world = translate(origin.x, origin.y, origin.z);
        s.world_to_object = translate(-origin.x, -origin.y, -origin.z);
        return s;
    }
    static bool new_triangle(V3 v0, V3 v1, V3 v2, Shape& out) {  // shape.rs:72-97
        V3 e1 = v1 - v0, e2 = v2 - v0;
        V3 n0 = cross(e2, e1);
        double mag = magnitude(n0);
        if (mag == 0.0) return false;
        n0 = n0 / mag;
        out = Shape{};
        out.kind = TRIANGLE;
        out.v0 = v0; out.e1 = e1; out.e2 = e2; out.n0 = n0;
        out.n01 = {0, 0, 0}; out.n02 = {0, 0, 0};
        out.uv0[0] = 0; out.uv0[1] = 0; out.uv01[0] = 1; out.uv01[1] = 0; out.uv02[0] = 1; out.uv02[1] = 1;
        return true;
    }
    static Shape new_disk(V3 origin, double rotate_x_deg, double rotate_y_deg, double radius, double inner_radius) {  // shape.rs:133-153
        Shape s{};
        s.kind = DISK;
        s.object_to_world = tmul(tmul(translate(origin.x, origin.y, origin.z), rotate_x(to_radians(rotate_x_deg))),
                                 rotate_y(to_radians(rotate_y_deg)));
        s.world_to_object = s.object_to_world.inverse();
        s.radius = radius;
        s.inner_radius = inner_radius;
        return s;
    }

    ShapeIntersection to_world(const ShapeIntersection& si) const {  // transformation.rs:488-496
        ShapeIntersection r = si;
        r.location = xf_point(object_to_world, si.location);
        r.normal = xf_normal(object_to_world, si.normal);
        return r;
    }

    // shape.rs:157-311.  Updates ray.max_distance on a hit.
    bool intersect(Ray& ray, ShapeIntersection& out) const {
        switch (kind) {
            case SPHERE: {
                Ray obj_ray = xf_ray(world_to_object, ray);
                V3 oc{obj_ray.origin.x, obj_ray.origin.y, obj_ray.origin.z};
                double a = magnitude_squared(obj_ray.direction);
                double b = 2.0 * dot(oc, obj_ray.direction);
                double c = magnitude_squared(oc) - radius * radius;  // radius.powf(2.0): LLVM folds pow(x,2) to x*x
                double discriminant = b * b - 4.0 * a * c;
                if (discriminant < 0.0) return false;
                double discriminant_sqrt = std::sqrt(discriminant);
                double inv_2_a = 1.0 / (2.0 * a);
                double distance = (-b - discriminant_sqrt) * inv_2_a;
                for (int root = 0; root < 2; ++root) {
                    if (root == 1) distance = (-b + discriminant_sqrt) * inv_2_a;
                    if (obj_ray.update_max_distance(distance)) {
                        V3 location = obj_ray.at(distance);
                        ray.update_max_distance(distance);
                        double phi = std::atan2(location.y, location.x);
                        if (phi < 0.0) phi += PI * 2.0;
                        double u = phi / (PI * 2.0);
                        double theta = std::acos(location.z / radius);
                        double v = theta * FRAC_1_PI;
                        ShapeIntersection si{location, V3{location.x, location.y, location.z} / radius, {u, v}, u, v};
                        out = to_world(si);
                        return true;
                    }
                }
                return false;
            }
            case TRIANGLE: {  // shape.rs:214-262 (Moeller-Trumbore)
                V3 P = cross(ray.direction, e2);
                double denominator = dot(P, e1);
                if (denominator > -EPSILON && denominator < EPSILON) return false;
                V3 T = ray.origin - v0;
                double u = dot(P, T) / denominator;
                if (u < 0.0 || u > 1.0) return false;
                V3 Q = cross(T, e1);
                double v = dot(Q, ray.direction) / denominator;
                if (v < 0.0 || u + v > 1.0) return false;
                double distance = dot(cross(T, e1), e2) / denominator;
                if (ray.update_max_distance(distance)) {
                    out.location = ray.at(distance);
                    out.normal = normalized(n0 + n01 * u + n02 * v);
                    out.uv[0] = uv0[0] + uv01[0] * u + uv02[0] * v;
                    out.uv[1] = uv0[1] + uv01[1] * u + uv02[1] * v;
                    out.bary_u = u;
                    out.bary_v = v;
                    return true;
                }
                return false;
            }
            default: {  // DISK shape.rs:263-309
                Ray obj_ray = xf_ray(world_to_object, ray);
                if (obj_ray.direction.z == 0.0) return false;
                double t = -obj_ray.origin.z / obj_ray.direction.z;
                if (!obj_ray.contains_distance(t)) return false;
                V3 location{obj_ray.origin.x + obj_ray.direction.x * t, obj_ray.origin.y + obj_ray.direction.y * t, 0.0};
                double distance_squared = location.x * location.x + location.y * location.y;  // powf(2.0)
                if (distance_squared < inner_radius * inner_radius || distance_squared > radius * radius) return false;
                double theta = std::atan2(location.y, location.x);
                if (theta < 0.0) theta += PI * 2.0;
                double u = theta / (PI * 2.0);
                double v = std::sqrt(distance_squared) / radius;
                if (ray.update_max_distance(t)) {
                    ShapeIntersection si{location, V3{0, 0, 1}, {u, v}, u, v};
                    out = to_world(si);
                    return true;
                }
                return false;
            }
        }
    }

    // shape.rs:314-400
    bool intersects(const Ray& ray) const {
        switch (kind) {
            case SPHERE: {
                Ray obj_ray = xf_ray(world_to_object, ray);
                V3 oc{obj_ray.origin.x, obj_ray.origin.y, obj_ray.origin.z};
                double a = magnitude_squared(obj_ray.direction);
                double b = 2.0 * dot(oc, obj_ray.direction);
                double c = magnitude_squared(oc) - radius * radius;
                double discriminant = b * b - 4.0 * a * c;
                if (discriminant < 0.0) return false;
                double discriminant_sqrt = std::sqrt(discriminant);
                double inv_2_a = 1.0 / (2.0 * a);
                double distance = (-b - discriminant_sqrt) * inv_2_a;
                if (obj_ray.contains_distance(distance)) return true;
                distance = (-b + discriminant_sqrt) * inv_2_a;
                return obj_ray.contains_distance(distance);
            }
            case TRIANGLE: {
                V3 P = cross(ray.direction, e2);
                double denominator = dot(P, e1);
                if (denominator > -EPSILON && denominator < EPSILON) return false;
                V3 T = ray.origin - v0;
                double u = dot(P, T) / denominator;
                if (u < 0.0 || u > 1.0) return false;
                V3 Q = cross(T, e1);
                double v = dot(Q, ray.direction) / denominator;
                if (v < 0.0 || u + v > 1.0) return false;
                double distance = dot(cross(T, e1), e2) / denominator;
                return ray.contains_distance(distance);
            }
            default: {
                Ray obj_ray = xf_ray(world_to_object, ray);
                if (obj_ray.direction.z == 0.0) return false;
                double t = -obj_ray.origin.z / obj_ray.direction.z;
                if (!obj_ray.contains_distance(t)) return false;
                V3 location{obj_ray.origin.x + obj_ray.direction.x * t, obj_ray.origin.y + obj_ray.direction.y * t, 0.0};
                double distance_squared = location.x * location.x + location.y * location.y;
                if (distance_squared < inner_radius * inner_radius || distance_squared > radius * radius) return false;
                return ray.contains_distance(t);
            }
        }
    }

    Bounds bounds() const {  // shape.rs:402-438
        switch (kind) {
            case SPHERE:
                return xf_bounds(object_to_world, Bounds::make({-radius, -radius, -radius}, {radius, radius, radius}));
            case TRIANGLE: {
                V3 v1 = v0 + e1, v2 = v0 + e2;
                return Bounds::make({rmin(v1.x, rmin(v2.x, v0.x)), rmin(v1.y, rmin(v2.y, v0.y)), rmin(v1.z, rmin(v2.z, v0.z))},
                                    {rmax(v1.x, rmax(v2.x, v0.x)), rmax(v1.y, rmax(v2.y, v0.y)), rmax(v1.z, rmax(v2.z, v0.z))});
            }
            default:
                return xf_bounds(object_to_world, Bounds::make({-radius, -radius, 0.0}, {radius, radius, 0.0}));
        }
    }

    V3 sample(double su, double sv) const {  // shape.rs:445-470
        switch (kind) {
            case SPHERE: {
                V3 p = V3{0, 0, 0} + sample_sphere(su, sv) * radius;
                return xf_point(object_to_world, p);
            }
            case TRIANGLE: {
                double b1, b2;
                sample_triangle(su, sv, b1, b2);
                return v0 + e1 * b1 + e2 * b2;
            }
            default: {
                double x, y;
                sample_disk(su, sv, x, y);
                return xf_point(object_to_world, V3{x * radius, y * radius, 0.0});
            }
        }
    }

    double area() const {  // shape.rs:504-514
        switch (kind) {
            case SPHERE: return PI * (radius * radius);
            case TRIANGLE: return magnitude(cross(e1, e2)) / 2.0;
            default: return PI * (radius * radius - inner_radius * inner_radius);
        }
    }

    // shape.rs:487-502; `location`/`normal` are those of the receiver intersection.
    double pdf_from(V3 location, V3 normal, V3 w_i) const {
        Ray ray = Ray::make(location, w_i);
        ShapeIntersection si;
        if (intersect(ray, si)) {
            double distance_squared = magnitude_squared(si.location - location);
            double cos_theta = std::fabs(dot(w_i, normal));
            return distance_squared / (cos_theta * area());
        }
        return 0.0;
    }
};

}  // namespace orc
