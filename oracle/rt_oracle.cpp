// rt_oracle.cpp — TEST INFRASTRUCTURE ONLY.  A double-precision CPU restatement of the
// reference's per-pixel hot path, driven by the same flattened scene description
// (include/rt_b200.h) the CUDA library consumes.  Only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline leg may load this; the product never does.
//
// Pinned against the real reference: tests/test_oracle_port.py checks this file against
// the golden fixtures in tests/golden/ that oracle/ref_driver.cpp produced from the
// UNMODIFIED headers of /root/reference (primary hits bit-close, material/texture known
// answers, converged-image statistics).
//
// Every function cites the reference lines it restates.  What is deliberately different:
//   * random numbers come from counter-based Philox with the device's dimension
//     assignment (philox_ref.h) instead of rand() (rtweekend.h:26-29), so that this checker
//     and the GPU trace the same sample sequence; the reference itself is not
//     reproducible;
//   * the closest surface hit is found first and media are evaluated against it
//     afterwards (the reference interleaves them in BVH order, constant_medium.h:20-53 —
//     same distribution), with density * multiplicity;
//   * translate/rotate_y chains arrive pre-composed as one rigid transform per primitive
//     (hittable.h:46-58, 101-139 applied once instead of per wrapper).
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <thread>
#include <vector>

#include "philox_ref.h"
#include "rt_b200.h"

namespace oracle {

const double kInf = std::numeric_limits<double>::infinity();
const double kPi = 3.1415926535897932385;  // rtweekend.h:15

struct V {
    double x = 0, y = 0, z = 0;
    double operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
};
inline V mk(const double* p) { return V{p[0], p[1], p[2]}; }
inline V operator+(V a, V b) { return V{a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V operator-(V a, V b) { return V{a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V operator-(V a) { return V{-a.x, -a.y, -a.z}; }
inline V operator*(V a, V b) { return V{a.x * b.x, a.y * b.y, a.z * b.z}; }
inline V operator*(double s, V a) { return V{s * a.x, s * a.y, s * a.z}; }
inline V operator/(V a, double s) { return (1 / s) * a; }  // vec3.h:96-98
inline double dot(V a, V b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline V cross(V a, V b) { return V{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline double length_squared(V a) { return a.x * a.x + a.y * a.y + a.z * a.z; }  // vec3.h:46-48
inline double length(V a) { return std::sqrt(length_squared(a)); }
inline V unit_vector(V a) { return a / length(a); }  // vec3.h:104-106

struct Ray {
    V o, d;
    double tm = 0;
    V at(double t) const { return o + t * d; }  // ray.h:22-24
};

struct HitRec {  // hittable.h:11-27 plus the primitive id the reference does not carry
    V p, normal;
    int material = -1;
    double t = 0;
    bool front_face = false;
    double u = 0, v = 0;
    int prim = -1;
    void set_face_normal(const Ray& r, V outward) {  // hittable.h:22-25
        front_face = dot(r.d, outward) < 0;
        normal = front_face ? outward : -outward;
    }
};

struct Xf {
    bool identity = true;
    double r[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, t[3] = {0, 0, 0};
    V to_obj_point(V p) const {
        if (identity) return p;
        V q{p.x - t[0], p.y - t[1], p.z - t[2]};
        return V{r[0] * q.x + r[3] * q.y + r[6] * q.z, r[1] * q.x + r[4] * q.y + r[7] * q.z, r[2] * q.x + r[5] * q.y + r[8] * q.z};
    }
    V to_obj_dir(V d) const {
        if (identity) return d;
        return V{r[0] * d.x + r[3] * d.y + r[6] * d.z, r[1] * d.x + r[4] * d.y + r[7] * d.z, r[2] * d.x + r[5] * d.y + r[8] * d.z};
    }
    V to_world_point(V p) const {
        if (identity) return p;
        return V{r[0] * p.x + r[1] * p.y + r[2] * p.z + t[0], r[3] * p.x + r[4] * p.y + r[5] * p.z + t[1],
                 r[6] * p.x + r[7] * p.y + r[8] * p.z + t[2]};
    }
    V to_world_dir(V d) const {
        if (identity) return d;
        return V{r[0] * d.x + r[1] * d.y + r[2] * d.z, r[3] * d.x + r[4] * d.y + r[5] * d.z, r[6] * d.x + r[7] * d.y + r[8] * d.z};
    }
};

struct Box {
    double lo[3] = {kInf, kInf, kInf}, hi[3] = {-kInf, -kInf, -kInf};
    void grow(V p) {
        for (int k = 0; k < 3; k++) { lo[k] = std::min(lo[k], p[k]); hi[k] = std::max(hi[k], p[k]); }
    }
    void grow(const Box& b) {
        for (int k = 0; k < 3; k++) { lo[k] = std::min(lo[k], b.lo[k]); hi[k] = std::max(hi[k], b.hi[k]); }
    }
    // aabb.h:61-85 (slab test, early out when the interval closes)
    bool hit(const Ray& r, double tmin, double tmax) const {
        for (int axis = 0; axis < 3; axis++) {
            const double adinv = 1.0 / r.d[axis];
            double t0 = (lo[axis] - r.o[axis]) * adinv, t1 = (hi[axis] - r.o[axis]) * adinv;
            if (t0 < t1) {
                if (t0 > tmin) tmin = t0;
                if (t1 < tmax) tmax = t1;
            } else {
                if (t1 > tmin) tmin = t1;
                if (t0 < tmax) tmax = t0;
            }
            if (tmax <= tmin) return false;
        }
        return true;
    }
};

struct Prim {
    int type = 0, index = 0, id = -1;
    Xf xf;
    Box box;  // world space, padded like aabb.h:98-105
};

struct Scene {
    const rt_scene_desc* d = nullptr;
    std::vector<Prim> world, boundary;
    // median-split BVH over `world` (bvh.h:13-45: longest axis, sort by box minimum, halve)
    struct Node { Box box; int left = -1, right = -1, first = 0, count = 0; };
    std::vector<Node> nodes;
    std::vector<int> order;
    // camera (Camera.txt:136-175)
    V center, pixel00, du, dv, disk_u, disk_v, background;
    double defocus_angle = 0;
};

Xf make_xf(const rt_scene_desc* d, int idx) {
    Xf x;
    if (idx >= 0) {
        x.identity = false;
        std::memcpy(x.r, d->xforms[idx].r, sizeof x.r);
        std::memcpy(x.t, d->xforms[idx].t, sizeof x.t);
    }
    return x;
}

Prim make_prim(const rt_scene_desc* d, const rt_prim_ref& ref, int id) {
    Prim p;
    p.type = ref.type;
    p.index = ref.index;
    p.id = id;
    auto corner = [&](V q) { p.box.grow(p.xf.to_world_point(q)); };
    if (ref.type == RT_PRIM_SPHERE) {
        const rt_sphere& s = d->spheres[ref.index];
        p.xf = make_xf(d, s.xform);
        for (int m = 0; m < 2; m++) {  // sphere.h:12-28: both ends of the motion
            V c = mk(s.center0) + (double)m * mk(s.center_vec);
            for (int i = 0; i < 8; i++)
                corner(V{c.x + ((i & 1) ? s.radius : -s.radius), c.y + ((i & 2) ? s.radius : -s.radius), c.z + ((i & 4) ? s.radius : -s.radius)});
        }
    } else if (ref.type == RT_PRIM_QUAD) {
        const rt_quad& q = d->quads[ref.index];
        p.xf = make_xf(d, q.xform);
        V Q = mk(q.Q), u = mk(q.u), v = mk(q.v);
        corner(Q); corner(Q + u); corner(Q + v); corner(Q + u + v);  // quad.h:22-27
    } else {
        const rt_triangle& t = d->triangles[ref.index];
        p.xf = make_xf(d, t.xform);
        corner(mk(t.p0)); corner(mk(t.p1)); corner(mk(t.p2));  // triangle.h:45-60
    }
    for (int k = 0; k < 3; k++) {  // generous padding: the box is only an accelerator here
        double pad = 1e-4 + 1e-9 * std::max(std::fabs(p.box.lo[k]), std::fabs(p.box.hi[k]));
        p.box.lo[k] -= pad;
        p.box.hi[k] += pad;
    }
    return p;
}

// sphere.h:32-58 + get_sphere_uv :67-73
bool hit_sphere(const rt_sphere& s, const Ray& r, double tmin, double tmax, HitRec& rec) {
    V current_center = mk(s.center0) + r.tm * mk(s.center_vec);
    V oc = current_center - r.o;
    double a = length_squared(r.d);
    double h = dot(r.d, oc);
    double radius = std::fmax(0, s.radius);
    double c = length_squared(oc) - radius * radius;
    double discriminant = h * h - a * c;
    if (discriminant < 0) return false;
    double sqrtd = std::sqrt(discriminant);
    double root = (h - sqrtd) / a;
    if (!(tmin < root && root < tmax)) {  // interval::surrounds
        root = (h + sqrtd) / a;
        if (!(tmin < root && root < tmax)) return false;
    }
    rec.t = root;
    rec.p = r.at(rec.t);
    V outward = (rec.p - current_center) / radius;
    rec.set_face_normal(r, outward);
    double theta = std::acos(-outward.y);
    double phi = std::atan2(-outward.z, outward.x) + kPi;
    rec.u = phi / (2 * kPi);
    rec.v = theta / kPi;
    rec.material = s.material;
    return true;
}

// quad.h:10-21 (constructor quantities) + :29-73
bool hit_quad(const rt_quad& q, const Ray& r, double tmin, double tmax, HitRec& rec) {
    V Q = mk(q.Q), u = mk(q.u), v = mk(q.v);
    V n = cross(u, v);
    V normal = unit_vector(n);
    double D = dot(normal, Q);
    V w = n / dot(n, n);
    double denom = dot(normal, r.d);
    if (std::fabs(denom) < 1e-8) return false;
    double t = (D - dot(normal, r.o)) / denom;
    if (!(tmin <= t && t <= tmax)) return false;  // interval::contains
    V intersection = r.at(t);
    V planar = intersection - Q;
    double alpha = dot(w, cross(planar, v));
    double beta = dot(w, cross(u, planar));
    if (!(0 <= alpha && alpha <= 1) || !(0 <= beta && beta <= 1)) return false;  // is_interior
    rec.u = alpha;
    rec.v = beta;
    rec.t = t;
    rec.p = intersection;
    rec.material = q.material;
    rec.set_face_normal(r, normal);
    return true;
}

// triangle.h:65-113, including its float det / invDet / alpha / beta / gamma
bool hit_triangle(const rt_triangle& tr, const Ray& r, double tmin, double tmax, HitRec& rec) {
    V p0 = mk(tr.p0), p1 = mk(tr.p1), p2 = mk(tr.p2);
    V v0v1 = p1 - p0, v0v2 = p2 - p0;
    V pvec = cross(r.d, v0v2);
    float det = (float)dot(v0v1, pvec);
    if (std::fabs(det) < 1e-8) return false;
    float invDet = 1.0f / det;
    V tvec = r.o - p0;
    double u = dot(tvec, pvec) * invDet;
    if (u < 0.0f || u > 1.0f) return false;
    V qvec = cross(tvec, v0v1);
    double v = dot(r.d, qvec) * invDet;
    if (v < 0.0f || u + v > 1.0f) return false;
    double t = dot(v0v2, qvec) * invDet;
    if (t < tmin || t > tmax) return false;
    float alpha = (float)(1 - u - v), beta = (float)u, gamma = (float)v;
    if (!(0 <= (double)alpha && (double)alpha <= 1) || !(0 <= (double)beta && (double)beta <= 1)) return false;  // is_interior
    // float arithmetic as in triangle.h:103-104 (float * float + ... then widened)
    rec.u = alpha * tr.uv0[0] + beta * tr.uv1[0] + gamma * tr.uv2[0];
    rec.v = alpha * tr.uv0[1] + beta * tr.uv1[1] + gamma * tr.uv2[1];
    rec.t = t;
    rec.p = r.at(t);
    rec.material = tr.material;
    V normal = unit_vector(cross(p1 - p0, p2 - p0));  // triangle.h:21-22
    rec.set_face_normal(r, normal);
    return true;
}

// one leaf under its (pre-composed) translate/rotate_y chain: hittable.h:46-58, 101-139
bool hit_prim(const rt_scene_desc* d, const Prim& p, const Ray& r, double tmin, double tmax, HitRec& rec) {
    Ray ro = r;
    if (!p.xf.identity) {
        ro.o = p.xf.to_obj_point(r.o);
        ro.d = p.xf.to_obj_dir(r.d);
    }
    bool ok;
    if (p.type == RT_PRIM_SPHERE) ok = hit_sphere(d->spheres[p.index], ro, tmin, tmax, rec);
    else if (p.type == RT_PRIM_QUAD) ok = hit_quad(d->quads[p.index], ro, tmin, tmax, rec);
    else ok = hit_triangle(d->triangles[p.index], ro, tmin, tmax, rec);
    if (!ok) return false;
    if (!p.xf.identity) {
        rec.p = p.xf.to_world_point(rec.p);
        rec.normal = p.xf.to_world_dir(rec.normal);
    }
    rec.prim = p.id;
    return true;
}

int build_bvh(Scene& sc, int start, int end) {
    Scene::Node node;
    for (int i = start; i < end; i++) node.box.grow(sc.world[sc.order[i]].box);
    int me = (int)sc.nodes.size();
    sc.nodes.push_back(node);
    int span = end - start;
    if (span <= 2) {
        sc.nodes[me].first = start;
        sc.nodes[me].count = span;
        return me;
    }
    double ext[3] = {node.box.hi[0] - node.box.lo[0], node.box.hi[1] - node.box.lo[1], node.box.hi[2] - node.box.lo[2]};
    int axis = ext[0] > ext[1] ? (ext[0] > ext[2] ? 0 : 2) : (ext[1] > ext[2] ? 1 : 2);  // aabb.h:87-93
    std::sort(sc.order.begin() + start, sc.order.begin() + end,
              [&](int a, int b) { return sc.world[a].box.lo[axis] < sc.world[b].box.lo[axis]; });
    int mid = start + span / 2;
    int l = build_bvh(sc, start, mid);
    int r = build_bvh(sc, mid, end);
    sc.nodes[me].left = l;
    sc.nodes[me].right = r;
    return me;
}

// bvh.h:64-72 + hittable_list.h:22-35: closest hit with a shrinking t_max
bool world_hit(const Scene& sc, int node, const Ray& r, double tmin, double tmax, HitRec& rec) {
    const Scene::Node& n = sc.nodes[node];
    if (!n.box.hit(r, tmin, tmax)) return false;
    if (n.left < 0) {
        bool any = false;
        double closest = tmax;
        HitRec tmp;
        for (int i = 0; i < n.count; i++) {
            if (hit_prim(sc.d, sc.world[sc.order[n.first + i]], r, tmin, closest, tmp)) {
                any = true;
                closest = tmp.t;
                rec = tmp;
            }
        }
        return any;
    }
    bool hl = world_hit(sc, n.left, r, tmin, tmax, rec);
    bool hr = world_hit(sc, n.right, r, tmin, hl ? rec.t : tmax, rec);
    return hl || hr;
}

bool surface_hit(const Scene& sc, const Ray& r, double tmin, double tmax, HitRec& rec) {
    if (sc.world.empty()) return false;
    return world_hit(sc, 0, r, tmin, tmax, rec);
}

// hittable_list.h:22-35 over a medium's boundary primitives
bool boundary_hit(const Scene& sc, const rt_medium& m, const Ray& r, double tmin, double tmax, HitRec& rec) {
    bool any = false;
    double closest = tmax;
    HitRec tmp;
    for (int i = 0; i < m.boundary_count; i++) {
        if (hit_prim(sc.d, sc.boundary[m.boundary_first + i], r, tmin, closest, tmp)) {
            any = true;
            closest = tmp.t;
            rec = tmp;
        }
    }
    return any;
}

// ---- perlin.h:14-50, 72-89 -------------------------------------------------------
double perlin_noise(const rt_perlin& pn, V p) {
    double u = p.x - std::floor(p.x), v = p.y - std::floor(p.y), w = p.z - std::floor(p.z);
    int i = int(std::floor(p.x)), j = int(std::floor(p.y)), k = int(std::floor(p.z));
    double uu = u * u * (3 - 2 * u), vv = v * v * (3 - 2 * v), ww = w * w * (3 - 2 * w);
    double accum = 0.0;
    for (int di = 0; di < 2; di++)
        for (int dj = 0; dj < 2; dj++)
            for (int dk = 0; dk < 2; dk++) {
                int idx = pn.perm_x[(i + di) & 255] ^ pn.perm_y[(j + dj) & 255] ^ pn.perm_z[(k + dk) & 255];
                V c = mk(pn.randvec[idx]);
                V weight_v{u - di, v - dj, w - dk};
                accum += (di * uu + (1 - di) * (1 - uu)) * (dj * vv + (1 - dj) * (1 - vv)) * (dk * ww + (1 - dk) * (1 - ww)) * dot(c, weight_v);
            }
    return accum;
}
double perlin_turb(const rt_perlin& pn, V p, int depth) {
    double accum = 0.0, weight = 1.0;
    V temp_p = p;
    for (int i = 0; i < depth; i++) {
        accum += weight * perlin_noise(pn, temp_p);
        weight *= 0.5;
        temp_p = 2.0 * temp_p;
    }
    return std::fabs(accum);
}

// ---- texture.h ---------------------------------------------------------------------
V texture_value(const rt_scene_desc* d, int tex, double u, double v, V p) {
    const rt_texture& t = d->textures[tex];
    switch (t.type) {
        case RT_TEX_SOLID: return mk(t.color);  // :26-28
        case RT_TEX_CHECKER: {                  // :42-50
            int xi = int(std::floor(t.scale * p.x)), yi = int(std::floor(t.scale * p.y)), zi = int(std::floor(t.scale * p.z));
            bool even = (xi + yi + zi) % 2 == 0;
            return texture_value(d, even ? t.even : t.odd, u, v, p);
        }
        case RT_TEX_CHECKER_TRIANGLE: {  // :66-76
            v = 1.0 - v;
            int ui = int(std::round(t.scale * u * 10)), vi = int(std::round(t.scale * v * 10));
            bool even = (ui + vi) % 2 == 0;
            return texture_value(d, even ? t.even : t.odd, u, v, p);
        }
        case RT_TEX_IMAGE: {  // :90-104 + rtw_stb_image.h:71-97
            if (t.image < 0) return V{0, 1, 1};
            const rt_image& im = d->images[t.image];
            u = u < 0 ? 0 : (u > 1 ? 1 : u);
            v = 1.0 - (v < 0 ? 0 : (v > 1 ? 1 : v));
            int i = int(u * im.width), j = int(v * im.height);
            i = i < 0 ? 0 : (i < im.width ? i : im.width - 1);
            j = j < 0 ? 0 : (j < im.height ? j : im.height - 1);
            const uint8_t* px = im.rgb + 3 * ((size_t)j * im.width + i);
            double s = 1.0 / 255.0;
            return V{s * px[0], s * px[1], s * px[2]};
        }
        default: {  // noise :115
            double val = 1 + std::sin(t.scale * p.z + 10 * perlin_turb(d->perlins[t.perlin], p, 7));
            return V{.5 * val, .5 * val, .5 * val};
        }
    }
}

// vec3.h:107-115: the loop always exits on its first pass
V random_unit_vector(double ux, double uy, double uz) {
    V p{-1 + 2 * ux, -1 + 2 * uy, -1 + 2 * uz};  // random_double(-1,1) = min + (max-min)*u
    double lensq = length_squared(p);
    return p / std::sqrt(lensq);
}
V reflect(V v, V n) { return v - 2 * dot(v, n) * n; }  // vec3.h:125-127
V refract(V uv, V n, double etai_over_etat) {          // vec3.h:128-133
    double cos_theta = std::fmin(dot(-uv, n), 1.0);
    V perp = etai_over_etat * (uv + cos_theta * n);
    V parallel = -std::sqrt(std::fabs(1.0 - length_squared(perp))) * n;
    return perp + parallel;
}
bool near_zero(V v) { return std::fabs(v.x) < 1e-8 && std::fabs(v.y) < 1e-8 && std::fabs(v.z) < 1e-8; }

V emitted(const rt_scene_desc* d, const rt_material& m, const HitRec& rec) {  // material.h:14,99-101,111-113
    if (m.type == RT_MAT_DIFFUSE_LIGHT || m.type == RT_MAT_EMISSIVE_LIGHT) return texture_value(d, m.texture, rec.u, rec.v, rec.p);
    return V{0, 0, 0};
}

bool scatter(const rt_scene_desc* d, const rt_material& m, const Ray& in, const HitRec& rec, U4 u, V& att, Ray& out) {
    out.o = rec.p;
    out.tm = in.tm;
    switch (m.type) {
        case RT_MAT_LAMBERTIAN: {  // material.h:29-38
            V dir = rec.normal + random_unit_vector(u.x, u.y, u.z);
            if (near_zero(dir)) dir = rec.normal;
            out.d = dir;
            att = texture_value(d, m.texture, rec.u, rec.v, rec.p);
            return true;
        }
        case RT_MAT_METAL: {  // material.h:82-88
            V refl = reflect(in.d, rec.normal);
            refl = unit_vector(refl) + m.param * random_unit_vector(u.x, u.y, u.z);
            out.d = refl;
            att = mk(m.albedo);
            return dot(out.d, rec.normal) > 0;
        }
        case RT_MAT_DIELECTRIC: {  // material.h:47-74
            att = V{1, 1, 1};
            double ri = rec.front_face ? (1.0 / m.param) : m.param;
            V unit = unit_vector(in.d);
            double cos_theta = std::fmin(dot(-unit, rec.normal), 1.0);
            double sin_theta = std::sqrt(1.0 - cos_theta * cos_theta);
            bool cannot_refract = ri * sin_theta > 1.0;
            double r0 = (1 - ri) / (1 + ri);
            r0 = r0 * r0;
            double reflectance = r0 + (1 - r0) * std::pow((1 - cos_theta), 5);
            out.d = (cannot_refract || reflectance > u.w) ? reflect(unit, rec.normal) : refract(unit, rec.normal, ri);
            return true;
        }
        case RT_MAT_ISOTROPIC: {  // material.h:129-134
            out.d = random_unit_vector(u.x, u.y, u.z);
            att = texture_value(d, m.texture, rec.u, rec.v, rec.p);
            return true;
        }
        case RT_MAT_SPECULAR: {  // material.h:145-167
            V unit = unit_vector(in.d);
            V refl = reflect(unit, rec.normal);
            V diffuse = random_unit_vector(u.x, u.y, u.z);
            if (!(dot(diffuse, rec.normal) > 0.0)) diffuse = -diffuse;  // vec3.h:116-124
            double factor = std::pow(1.0 - dot(refl, unit), m.param);
            V dir = factor * refl + (1.0 - factor) * diffuse;
            if (near_zero(dir)) dir = rec.normal;
            out.d = dir;
            att = mk(m.albedo);
            return true;
        }
        default: return false;
    }
}

V get_lighting(const rt_scene_desc* d, V p, V normal) {  // Camera.txt:240-272
    V result{0, 0, 0};
    for (int i = 0; i < d->n_lights; i++) {
        const rt_point_light& l = d->lights[i];
        V light_dir = mk(l.position) - p;
        double distance_squared = length_squared(light_dir);
        light_dir = unit_vector(light_dir);
        double diffuse = std::max(dot(normal, light_dir), 0.0);
        double radius_effect = l.size * 0.1;
        if (distance_squared <= l.size * l.size) {
            result = result + diffuse * mk(l.intensity);
        } else {
            double attenuation = 1.0 / (distance_squared + radius_effect);
            result = result + diffuse * (attenuation * mk(l.intensity));
        }
    }
    return result;
}

// world.hit for the full scene: surfaces, then constant_medium.h:20-53 per medium
bool scene_hit(const Scene& sc, const Ray& r, const Rng& rng, uint32_t bounce, HitRec& rec) {
    const double tmin = 0.001;
    bool any = surface_hit(sc, r, tmin, kInf, rec);
    double closest = any ? rec.t : kInf;
    const rt_scene_desc* d = sc.d;
    U4 u4{0, 0, 0, 0};
    for (int m = 0; m < d->n_media; m++) {
        if ((m & 3) == 0) u4 = rng.draw(bounce, RS_MEDIUM + (m >> 2));
        double u = (m & 3) == 0 ? u4.x : ((m & 3) == 1 ? u4.y : ((m & 3) == 2 ? u4.z : u4.w));
        const rt_medium& md = d->media[m];
        HitRec rec1, rec2;
        if (!boundary_hit(sc, md, r, -kInf, kInf, rec1)) continue;
        if (!boundary_hit(sc, md, r, rec1.t + 0.0001, kInf, rec2)) continue;
        if (rec1.t < tmin) rec1.t = tmin;
        if (rec2.t > closest) rec2.t = closest;
        if (rec1.t >= rec2.t) continue;
        if (rec1.t < 0) rec1.t = 0;
        double ray_length = length(r.d);
        double distance_inside = (rec2.t - rec1.t) * ray_length;
        double neg_inv_density = -1 / (md.density * md.multiplicity);
        double hit_distance = neg_inv_density * std::log(u);
        if (hit_distance > distance_inside) continue;
        rec.t = rec1.t + hit_distance / ray_length;
        rec.p = r.at(rec.t);
        Xf x = make_xf(d, md.xform);
        rec.normal = x.to_world_dir(V{1, 0, 0});
        rec.front_face = true;
        rec.material = md.material;
        rec.u = rec.v = 0;
        rec.prim = -1;
        closest = rec.t;
        any = true;
    }
    return any;
}

// Camera.txt:203-238
V ray_color(const Scene& sc, const Ray& r, int depth, const Rng& rng, uint32_t bounce) {
    if (depth <= 0) return V{0, 0, 0};
    HitRec rec;
    if (!scene_hit(sc, r, rng, bounce, rec)) return sc.background;
    const rt_material& m = sc.d->materials[rec.material];
    V color_from_emission = emitted(sc.d, m, rec);
    Ray scattered;
    V attenuation;
    U4 u{0, 0, 0, 0};
    if (m.type != RT_MAT_DIFFUSE_LIGHT && m.type != RT_MAT_EMISSIVE_LIGHT) u = rng.draw(bounce, RS_SCATTER);
    if (!scatter(sc.d, m, r, rec, u, attenuation, scattered)) return color_from_emission;
    V lighting = attenuation * get_lighting(sc.d, rec.p, rec.normal);
    V color_from_scatter = attenuation * ray_color(sc, scattered, depth - 1, rng, bounce + 1);
    return color_from_emission + lighting + color_from_scatter;
}

void init_camera(Scene& sc, int width, int height) {  // Camera.txt:136-175
    const rt_camera& c = sc.d->camera;
    sc.center = mk(c.lookfrom);
    double theta = c.vfov * kPi / 180.0;
    double h = std::tan(theta / 2);
    double viewport_height = 2 * h * c.focus_dist;
    double viewport_width = viewport_height * (double(width) / height);
    V w = unit_vector(mk(c.lookfrom) - mk(c.lookat));
    V u = unit_vector(cross(mk(c.vup), w));
    V v = cross(w, u);
    V viewport_u = viewport_width * u;
    V viewport_v = viewport_height * (-v);
    sc.du = viewport_u / width;
    sc.dv = viewport_v / height;
    V upper_left = sc.center - (c.focus_dist * w) - viewport_u / 2 - viewport_v / 2;
    sc.pixel00 = upper_left + 0.5 * (sc.du + sc.dv);
    double defocus_radius = c.focus_dist * std::tan((c.defocus_angle / 2) * kPi / 180.0);
    sc.disk_u = defocus_radius * u;
    sc.disk_v = defocus_radius * v;
    sc.background = mk(c.background);
    sc.defocus_angle = c.defocus_angle;
}

Ray get_ray(const Scene& sc, int i, int j, const Rng& rng) {  // Camera.txt:177-200
    U4 u = rng.draw(0, RS_CAMERA);
    V pixel_sample = sc.pixel00 + ((i + (u.x - 0.5)) * sc.du) + ((j + (u.y - 0.5)) * sc.dv);
    V origin = sc.center;
    if (sc.defocus_angle > 0) {
        double px = 0, py = 0;
        for (uint32_t attempt = 0;; attempt++) {  // random_in_unit_disk, vec3.h:135-142
            U4 c = rng.draw(0, RS_DEFOCUS + (attempt < 14 ? attempt : 14));
            px = -1 + 2 * c.x; py = -1 + 2 * c.y;
            if (px * px + py * py < 1) break;
            px = -1 + 2 * c.z; py = -1 + 2 * c.w;
            if (px * px + py * py < 1) break;
            if (attempt >= 14) { px = py = 0; break; }
        }
        origin = sc.center + (px * sc.disk_u) + (py * sc.disk_v);
    }
    Ray r;
    r.o = origin;
    r.d = pixel_sample - origin;
    r.tm = u.z;
    return r;
}

bool init_scene(Scene& sc, const rt_scene_desc* d) {
    if (!d || d->struct_size != sizeof(rt_scene_desc) || d->abi_version != RT_B200_ABI_VERSION) return false;
    sc.d = d;
    for (int i = 0; i < d->n_world; i++) sc.world.push_back(make_prim(d, d->world[i], i));
    for (int i = 0; i < d->n_boundary_refs; i++) sc.boundary.push_back(make_prim(d, d->boundary_refs[i], -1));
    sc.order.resize(sc.world.size());
    for (size_t i = 0; i < sc.order.size(); i++) sc.order[i] = (int)i;
    if (!sc.world.empty()) build_bvh(sc, 0, (int)sc.world.size());
    return true;
}

}  // namespace oracle

using namespace oracle;

// rows [0, n) over all host threads (the checker itself has no shared mutable state)
template <class F>
static void parallel_rows(int n, F body) {
    std::atomic<int> next{0};
    int threads = std::max(1, (int)std::thread::hardware_concurrency());
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; t++)
        pool.emplace_back([&]() {
            for (;;) {
                int j = next.fetch_add(1);
                if (j >= n) break;
                body(j);
            }
        });
    for (auto& th : pool) th.join();
}

extern "C" {

// Mean radiance (and single-sample variance) per pixel, samples [spp_begin, spp_begin+spp).
int oracle_render(const rt_scene_desc* d, int width, int height, int spp, int spp_begin, int depth, uint64_t seed, int threads,
                  double* rgb, double* var) {
    Scene sc;
    if (!init_scene(sc, d) || width <= 0 || height <= 0 || spp <= 0) return 1;
    init_camera(sc, width, height);
    if (threads <= 0) threads = (int)std::thread::hardware_concurrency();
    std::atomic<int> next_row{0};
    auto worker = [&]() {
        for (;;) {
            int j = next_row.fetch_add(1);
            if (j >= height) break;
            for (int i = 0; i < width; i++) {
                V acc, acc2;
                Rng rng;
                rng.pixel = (uint32_t)(j * width + i);
                rng.k0 = (uint32_t)(seed & 0xffffffffu);
                rng.k1 = (uint32_t)(seed >> 32);
                for (int s = 0; s < spp; s++) {
                    rng.sample = (uint32_t)(spp_begin + s);
                    Ray r = get_ray(sc, i, j, rng);
                    V c = ray_color(sc, r, depth, rng, 0);
                    if (!(std::isfinite(c.x) && std::isfinite(c.y) && std::isfinite(c.z))) continue;
                    acc = acc + c;
                    acc2 = acc2 + c * c;
                }
                size_t px = ((size_t)j * width + i) * 3;
                for (int k = 0; k < 3; k++) {
                    double mean = acc[k] / spp;
                    rgb[px + k] = mean;
                    if (var) var[px + k] = std::max(0.0, acc2[k] / spp - mean * mean);
                }
            }
        }
    };
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; t++) pool.emplace_back(worker);
    for (auto& t : pool) t.join();
    return 0;
}

// Pixel-centre primary hits (media skipped), same definition as rt_render_aov.
int oracle_primary(const rt_scene_desc* d, int width, int height, int32_t* ids, double* t, double* normal, double* point, double* uv) {
    Scene sc;
    if (!init_scene(sc, d)) return 1;
    init_camera(sc, width, height);
    parallel_rows(height, [&](int j) {
        for (int i = 0; i < width; i++) {
            Ray r;
            r.o = sc.center;
            r.d = sc.pixel00 + ((double)i * sc.du) + ((double)j * sc.dv) - sc.center;
            r.tm = 0;
            HitRec rec;
            bool hit = surface_hit(sc, r, 0.001, kInf, rec);
            size_t px = (size_t)j * width + i;
            ids[px] = hit ? rec.prim : -1;
            t[px] = hit ? rec.t : 0;
            for (int k = 0; k < 3; k++) {
                normal[3 * px + k] = hit ? rec.normal[k] : 0;
                point[3 * px + k] = hit ? rec.p[k] : 0;
            }
            uv[2 * px] = hit ? rec.u : 0;
            uv[2 * px + 1] = hit ? rec.v : 0;
        }
    });
    return 0;
}

// Closest hit of arbitrary rays: rays = n x 9 (o, d, time, tmin, tmax)
int oracle_hit(const rt_scene_desc* d, int n, const double* rays, int32_t* ids, double* t, double* normal, double* uv) {
    Scene sc;
    if (!init_scene(sc, d)) return 1;
    const int chunk = 4096;
    parallel_rows((n + chunk - 1) / chunk, [&](int c) {
        for (int i = c * chunk; i < std::min(n, (c + 1) * chunk); i++) {
            const double* q = rays + 9 * (size_t)i;
            Ray r;
            r.o = mk(q); r.d = mk(q + 3); r.tm = q[6];
            HitRec rec;
            bool hit = surface_hit(sc, r, q[7], q[8], rec);
            ids[i] = hit ? rec.prim : -1;
            t[i] = hit ? rec.t : 0;
            for (int k = 0; k < 3; k++) normal[3 * i + k] = hit ? rec.normal[k] : 0;
            uv[2 * i] = hit ? rec.u : 0;
            uv[2 * i + 1] = hit ? rec.v : 0;
        }
    });
    return 0;
}

// material::emitted + material::scatter with caller-supplied uniforms.
// in: n x 16 (o[3] d[3] time p[3] normal[3] front u v), uniforms n x 4,
// out: n x 16 (scattered, att[3], o[3], d[3], emitted[3], time, pad[2])
int oracle_scatter(const rt_scene_desc* d, int material, int n, const double* in, const double* uniforms, double* out) {
    if (!d || material < 0 || material >= d->n_materials) return 1;
    const rt_material& m = d->materials[material];
    for (int i = 0; i < n; i++) {
        const double* q = in + 16 * (size_t)i;
        Ray r;
        r.o = mk(q); r.d = mk(q + 3); r.tm = q[6];
        HitRec rec;
        rec.p = mk(q + 7); rec.normal = mk(q + 10); rec.front_face = q[13] != 0; rec.u = q[14]; rec.v = q[15];
        rec.material = material;
        U4 u{uniforms[4 * i], uniforms[4 * i + 1], uniforms[4 * i + 2], uniforms[4 * i + 3]};
        V em = emitted(d, m, rec);
        V att;
        Ray s;
        bool ok = scatter(d, m, r, rec, u, att, s);
        double* o = out + 16 * (size_t)i;
        o[0] = ok;
        o[1] = att.x; o[2] = att.y; o[3] = att.z;
        o[4] = s.o.x; o[5] = s.o.y; o[6] = s.o.z;
        o[7] = s.d.x; o[8] = s.d.y; o[9] = s.d.z;
        o[10] = em.x; o[11] = em.y; o[12] = em.z;
        o[13] = s.tm; o[14] = o[15] = 0;
    }
    return 0;
}

int oracle_texture(const rt_scene_desc* d, int texture, int n, const double* uvp, double* rgb) {
    if (!d || texture < 0 || texture >= d->n_textures) return 1;
    for (int i = 0; i < n; i++) {
        V c = texture_value(d, texture, uvp[5 * i], uvp[5 * i + 1], mk(uvp + 5 * i + 2));
        rgb[3 * i] = c.x; rgb[3 * i + 1] = c.y; rgb[3 * i + 2] = c.z;
    }
    return 0;
}

}  // extern "C"
