// ref_driver.cpp — TEST INFRASTRUCTURE ONLY (oracle/): drives the UNMODIFIED reference
// CPU renderer, compiled from the headers where they lie under /root/reference, to
// produce the golden data the parity tests pin against and to time the CPU baseline.
// Nothing in the product links, loads or executes this.
//
// The reference keeps all geometry and camera state private and its hit_record has no
// primitive id (hittable.h:11-27), so this one translation unit is compiled with
// `private` redefined to `public` (after every standard header has been seen) purely
// to READ the reference objects: walking bvh_node::left/right, translate::offset,
// rotate_y::sin_theta/cos_theta, sphere::center/radius, quad::Q/u/v, triangle::p0/p1/p2,
// and to call camera::initialize / get_ray / ray_color (Camera.txt:136-238) from a
// loop that writes float radiance instead of the reference's RGB8 PNG.  No reference
// function is re-implemented here.
//
// rand(): the reference draws every random number from libc rand() (rtweekend.h:26-29),
// which serialises all threads on a glibc lock.  This binary defines its own
// rand()/srand() on top of the thread-local generator of host_rng.h (same [0,2^31)
// range).  A "scripted" mode feeds chosen values for the known-answer vectors.
//
// Modes (see usage()):
//   render   float radiance image from the reference's get_ray/ray_color, dynamic rows
//   time     wall-clock the reference's own camera::render (its threading, its PNG)
//   primary  per-pixel primary-hit dump (leaf index, t, normal, p, u, v) + leaf table
//   kat      known-answer vectors for material::scatter/emitted + texture::value
//   scene    summary (object count) — used to compare scene construction

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <future>
#include <iostream>
#include <limits>
#include <map>
#include <memory>
#include <mutex>
#include <regex>
#include <sstream>
#include <string>
#include <thread>
#include <vector>
#include <cuda_runtime.h>

#include "raytracingoneweekendapplication_b200/host/host_rng.h"

// ---- rand() interposer ------------------------------------------------------------
static thread_local const int* g_script = nullptr;
static thread_local int g_script_len = 0, g_script_pos = 0;
static std::atomic<uint64_t> g_thread_counter{1};
static uint64_t g_seed = 1;
static thread_local bool g_thread_seeded = false;

extern "C" int rand(void) noexcept {
    if (g_script) {
        int v = g_script[g_script_pos % g_script_len];
        g_script_pos++;
        return v;
    }
    if (!g_thread_seeded) {  // worker threads get their own stream
        g_thread_seeded = true;
        rtb200::host_srand(g_seed, g_thread_counter.fetch_add(1));
    }
    return rtb200::host_rand31();
}
extern "C" void srand(unsigned s) noexcept {
    g_seed = s;
    g_thread_seeded = true;
    rtb200::host_srand(s);
}

// ---- the reference, untouched -----------------------------------------------------
#define private public
#include "rtweekend.h"

#include "bvh.h"
#include "hittable_list.h"
#include "hittable.h"
#include "sphere.h"
#include "triangle.h"
#include "camera.h"  // oracle/_ref/camera.h := Camera.txt minus two `const` (Makefile)
#include "quad.h"
#include "point_light.h"
#include "texture.h"
#include "constant_medium.h"
#include "mesh.h"
#define STB_IMAGE_WRITE_IMPLEMENTATION
#include "stb_image_write.h"
#undef private

// scenes.h calls rtb200::host_srand for seeding; route it through srand() above so the
// main thread is marked seeded.
#include "scenes.h"

// ---- object-graph walk ------------------------------------------------------------
struct Step { int kind; double a, b, c; };  // kind 0: translate(a,b,c); 1: rotate_y(sin=a, cos=b)
struct Leaf {
    shared_ptr<hittable> obj;
    std::vector<Step> chain;  // outermost first
    int type;                 // 0 sphere, 1 quad, 2 triangle
};
struct MediumInfo { double density; int multiplicity; };

static void walk(const shared_ptr<hittable>& h, std::vector<Step>& chain, std::vector<Leaf>& leaves,
                 std::vector<MediumInfo>& media, int mult) {
    if (auto l = std::dynamic_pointer_cast<hittable_list>(h)) {
        for (auto& o : l->objects) walk(o, chain, leaves, media, mult);
    } else if (auto b = std::dynamic_pointer_cast<bvh_node>(h)) {
        if (b->left == b->right) {
            walk(b->left, chain, leaves, media, mult * 2);
        } else {
            walk(b->left, chain, leaves, media, mult);
            walk(b->right, chain, leaves, media, mult);
        }
    } else if (auto t = std::dynamic_pointer_cast<translate>(h)) {
        chain.push_back({0, t->offset.x(), t->offset.y(), t->offset.z()});
        walk(t->object, chain, leaves, media, mult);
        chain.pop_back();
    } else if (auto r = std::dynamic_pointer_cast<rotate_y>(h)) {
        chain.push_back({1, r->sin_theta, r->cos_theta, 0});
        walk(r->object, chain, leaves, media, mult);
        chain.pop_back();
    } else if (auto m = std::dynamic_pointer_cast<constant_medium>(h)) {
        media.push_back({-1.0 / m->neg_inv_density, mult});
    } else if (std::dynamic_pointer_cast<sphere>(h)) {
        leaves.push_back({h, chain, 0});
    } else if (std::dynamic_pointer_cast<quad>(h)) {
        leaves.push_back({h, chain, 1});
    } else if (std::dynamic_pointer_cast<triangle>(h)) {
        leaves.push_back({h, chain, 2});
    } else {
        std::fprintf(stderr, "walk: unknown hittable\n");
    }
}

// ray into the leaf's object space, by the same arithmetic as hittable.h:46-48,101-117
static ray to_object(const ray& r, const std::vector<Step>& chain) {
    ray cur = r;
    for (const Step& s : chain) {
        if (s.kind == 0) {
            cur = ray(cur.origin() - vec3(s.a, s.b, s.c), cur.direction(), cur.time());
        } else {
            auto o = point3(s.b * cur.origin().x() - s.a * cur.origin().z(), cur.origin().y(),
                            s.a * cur.origin().x() + s.b * cur.origin().z());
            auto d = vec3(s.b * cur.direction().x() - s.a * cur.direction().z(), cur.direction().y(),
                          s.a * cur.direction().x() + s.b * cur.direction().z());
            cur = ray(o, d, cur.time());
        }
    }
    return cur;
}
static vec3 point_to_world(vec3 p, const std::vector<Step>& chain) {
    for (int i = (int)chain.size() - 1; i >= 0; i--) {
        const Step& s = chain[i];
        if (s.kind == 0) p = p + vec3(s.a, s.b, s.c);
        else p = vec3(s.b * p.x() + s.a * p.z(), p.y(), -s.a * p.x() + s.b * p.z());
    }
    return p;
}
static vec3 dir_to_world(vec3 d, const std::vector<Step>& chain) {
    for (int i = (int)chain.size() - 1; i >= 0; i--) {
        const Step& s = chain[i];
        if (s.kind == 1) d = vec3(s.b * d.x() + s.a * d.z(), d.y(), -s.a * d.x() + s.b * d.z());
    }
    return d;
}
static shared_ptr<material> leaf_material(const Leaf& lf) {
    if (lf.type == 0) return std::static_pointer_cast<sphere>(lf.obj)->mat;
    if (lf.type == 1) return std::static_pointer_cast<quad>(lf.obj)->mat;
    return std::static_pointer_cast<triangle>(lf.obj)->mat;
}

// 10 doubles per leaf: type, then world-space geometry
//   sphere: c0(3) cvec(3) radius 0 0 ; quad: Q(3) u(3) v(3) ; triangle: p0 p1 p2
static void leaf_key(const Leaf& lf, double* out) {
    out[0] = lf.type;
    if (lf.type == 0) {
        auto s = std::static_pointer_cast<sphere>(lf.obj);
        vec3 c = point_to_world(s->center.origin(), lf.chain), cv = dir_to_world(s->center.direction(), lf.chain);
        for (int k = 0; k < 3; k++) { out[1 + k] = c[k]; out[4 + k] = cv[k]; }
        out[7] = s->radius; out[8] = out[9] = 0;
    } else if (lf.type == 1) {
        auto q = std::static_pointer_cast<quad>(lf.obj);
        vec3 Q = point_to_world(q->Q, lf.chain), u = dir_to_world(q->u, lf.chain), v = dir_to_world(q->v, lf.chain);
        for (int k = 0; k < 3; k++) { out[1 + k] = Q[k]; out[4 + k] = u[k]; out[7 + k] = v[k]; }
    } else {
        auto t = std::static_pointer_cast<triangle>(lf.obj);
        vec3 a = point_to_world(t->p0, lf.chain), b = point_to_world(t->p1, lf.chain), c = point_to_world(t->p2, lf.chain);
        for (int k = 0; k < 3; k++) { out[1 + k] = a[k]; out[4 + k] = b[k]; out[7 + k] = c[k]; }
    }
}

template <class T>
static void dump(const std::string& path, const std::vector<T>& v) {
    std::ofstream f(path, std::ios::binary);
    f.write((const char*)v.data(), (std::streamsize)(v.size() * sizeof(T)));
}

struct Scene {
    hittable_list world;
    camera cam;
    std::vector<point_light> lights;
    scene_config cfg;
};

static bool make_scene(Scene& sc, const std::string& name, unsigned seed, const std::string& assets) {
    sc.cfg.asset_dir = assets;
    srand(seed);
    return build_scene(name, seed, sc.world, sc.cam, sc.lights, sc.cfg);
}

// Finds which leaf produced `rec` for ray r: the leaf whose own hit() over the same
// interval returns exactly rec.t (same arithmetic path => bitwise equal).
static int identify(const std::vector<Leaf>& leaves, const ray& r, const hit_record& rec) {
    int found = -1;
    for (size_t i = 0; i < leaves.size(); i++) {
        if (leaf_material(leaves[i]) != rec.mat) continue;
        hit_record tmp;
        ray ro = to_object(r, leaves[i].chain);
        if (leaves[i].obj->hit(ro, interval(0.001, infinity), tmp) && tmp.t == rec.t) {
            found = (int)i;  // the LAST match wins a tie, like hittable_list.h:27-33
        }
    }
    return found;
}

static int usage() {
    std::fprintf(stderr,
                 "ref_driver render  <scene> <seed> <assets> <W> <H> <spp> <depth> <out.f32> [threads]\n"
                 "ref_driver time    <scene> <seed> <assets> <W> <H> <spp> <depth> [stock_rand]\n"
                 "ref_driver primary <scene> <seed> <assets> <W> <H> <out_prefix>\n"
                 "ref_driver kat     <scene> <seed> <assets> <n> <out_prefix>\n"
                 "ref_driver scene   <scene> <seed> <assets>\n");
    return 2;
}

int main(int argc, char** argv) {
    if (argc < 5) return usage();
    std::string mode = argv[1], name = argv[2];
    unsigned seed = (unsigned)std::atoi(argv[3]);
    std::string assets = argv[4];
    Scene sc;
    if (!make_scene(sc, name, seed, assets)) {
        std::fprintf(stderr, "unknown scene %s\n", name.c_str());
        return 2;
    }
    std::vector<Leaf> leaves;
    std::vector<MediumInfo> media;
    {
        std::vector<Step> chain;
        for (auto& o : sc.world.objects) walk(o, chain, leaves, media, 1);
    }

    if (mode == "scene") {
        std::printf("{\"scene\": \"%s\", \"leaves\": %zu, \"media\": %zu", name.c_str(), leaves.size(), media.size());
        std::printf(", \"medium_multiplicity\": [");
        for (size_t i = 0; i < media.size(); i++) std::printf("%s%d", i ? ", " : "", media[i].multiplicity);
        std::printf("], \"medium_density\": [");
        for (size_t i = 0; i < media.size(); i++) std::printf("%s%.17g", i ? ", " : "", media[i].density);
        std::printf("], \"width\": %d, \"height\": %d, \"spp\": %d, \"depth\": %d}\n", sc.cfg.width, sc.cfg.height,
                    sc.cfg.spp, sc.cfg.depth);
        return 0;
    }

    if (mode == "render" || mode == "time") {
        if (argc < 9) return usage();
        int W = std::atoi(argv[5]), H = std::atoi(argv[6]), spp = std::atoi(argv[7]), depth = std::atoi(argv[8]);
        sc.cam.image_width = W;
        sc.cam.aspect_ratio = double(W) / double(H);
        sc.cam.samples_per_pixel = spp;
        sc.cam.max_depth = depth;
        if (mode == "time") {
            // the reference's own render(): its std::async row bands, its PNG write
            sc.cam.image_name = "/tmp/rtb200_ref_time.png";
            auto t0 = std::chrono::steady_clock::now();
            sc.cam.render(sc.world, sc.lights);
            auto t1 = std::chrono::steady_clock::now();
            double s = std::chrono::duration<double>(t1 - t0).count();
            double samples = double(W) * sc.cam.image_height * spp;
            std::printf("{\"scene\": \"%s\", \"width\": %d, \"height\": %d, \"spp\": %d, \"depth\": %d, \"seconds\": %.6f, "
                        "\"msamples_per_s\": %.6f, \"threads\": %u}\n",
                        name.c_str(), W, sc.cam.image_height, spp, depth, s, samples / s * 1e-6,
                        std::thread::hardware_concurrency());
            return 0;
        }
        if (argc < 10) return usage();
        int threads = argc > 10 ? std::atoi(argv[10]) : (int)std::thread::hardware_concurrency();
        sc.cam.initialize();
        if (sc.cam.image_height != H) std::fprintf(stderr, "note: reference height %d != requested %d\n", sc.cam.image_height, H);
        H = sc.cam.image_height;
        std::vector<float> img((size_t)W * H * 3), var((size_t)W * H * 3);
        std::atomic<int> next_row{0};
        auto t0 = std::chrono::steady_clock::now();
        auto worker = [&]() {
            for (;;) {
                int j = next_row.fetch_add(1);
                if (j >= H) break;
                for (int i = 0; i < W; i++) {
                    color acc(0, 0, 0), acc2(0, 0, 0);
                    for (int s = 0; s < spp; s++) {
                        ray r = sc.cam.get_ray(i, j);
                        color c = sc.cam.ray_color(r, depth, sc.world, sc.lights);
                        acc += c;
                        acc2 += c * c;
                    }
                    for (int k = 0; k < 3; k++) {
                        double mean = acc[k] / spp;
                        img[((size_t)j * W + i) * 3 + k] = (float)mean;
                        // variance of ONE sample (for the Monte Carlo error bars of the parity tests)
                        var[((size_t)j * W + i) * 3 + k] = (float)std::max(0.0, acc2[k] / spp - mean * mean);
                    }
                }
            }
        };
        std::vector<std::thread> pool;
        for (int t = 0; t < threads; t++) pool.emplace_back(worker);
        for (auto& t : pool) t.join();
        double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        dump(argv[9], img);
        dump(std::string(argv[9]) + ".var", var);
        std::printf("{\"scene\": \"%s\", \"width\": %d, \"height\": %d, \"spp\": %d, \"seconds\": %.3f, \"threads\": %d}\n",
                    name.c_str(), W, H, spp, s, threads);
        return 0;
    }

    if (mode == "primary") {
        if (argc < 8) return usage();
        int W = std::atoi(argv[5]), H = std::atoi(argv[6]);
        std::string prefix = argv[7];
        sc.cam.image_width = W;
        sc.cam.aspect_ratio = double(W) / double(H);
        sc.cam.initialize();
        H = sc.cam.image_height;
        std::vector<int32_t> ids((size_t)W * H, -1);
        std::vector<double> rec_out((size_t)W * H * 9, 0.0);
        std::atomic<int> next_row{0};
        auto worker = [&]() {
            for (;;) {
                int j = next_row.fetch_add(1);
                if (j >= H) break;
                for (int i = 0; i < W; i++) {
                    // the pixel-centre ray: get_ray (Camera.txt:177-191) with zero offset,
                    // no defocus, time 0
                    auto pixel = sc.cam.pixel00_loc + (double(i) * sc.cam.pixel_delta_u) + (double(j) * sc.cam.pixel_delta_v);
                    ray r(sc.cam.center, pixel - sc.cam.center, 0.0);
                    hit_record rec;
                    // media draw random numbers; primary-hit parity is for surfaces only, so
                    // make every medium miss: log(rand()=0) = -inf => hit_distance = +inf
                    static const int zero = 0;
                    g_script = &zero; g_script_len = 1; g_script_pos = 0;
                    bool hit = sc.world.hit(r, interval(0.001, infinity), rec);
                    g_script = nullptr;
                    size_t px = (size_t)j * W + i;
                    if (!hit) continue;
                    ids[px] = identify(leaves, r, rec);
                    double* o = &rec_out[px * 9];
                    o[0] = rec.t;
                    for (int k = 0; k < 3; k++) { o[1 + k] = rec.normal[k]; o[4 + k] = rec.p[k]; }
                    o[7] = rec.u; o[8] = rec.v;
                }
            }
        };
        std::vector<std::thread> pool;
        for (unsigned t = 0; t < std::thread::hardware_concurrency(); t++) pool.emplace_back(worker);
        for (auto& t : pool) t.join();
        std::vector<double> keys(leaves.size() * 10);
        for (size_t i = 0; i < leaves.size(); i++) leaf_key(leaves[i], &keys[i * 10]);
        dump(prefix + ".ids.i32", ids);
        dump(prefix + ".hit.f64", rec_out);
        dump(prefix + ".leaves.f64", keys);
        std::printf("{\"scene\": \"%s\", \"width\": %d, \"height\": %d, \"leaves\": %zu}\n", name.c_str(), W, H, leaves.size());
        return 0;
    }

    if (mode == "kat") {
        if (argc < 7) return usage();
        int n = std::atoi(argv[5]);
        std::string prefix = argv[6];
        sc.cam.initialize();
        const int W = sc.cam.image_width, H = sc.cam.image_height;
        // record layout (doubles), 40 per case:
        //  0 leaf  1..3 o  4..6 d  7 time  8 t  9..11 p  12..14 normal  15 front  16 u  17 v
        //  18..21 uniforms (x,y,z of the unit vector's cube point in [0,1), dielectric draw)
        //  22 scattered  23..25 attenuation  26..28 scattered.o  29..31 scattered.d  32 scattered.time
        //  33..35 emitted  36..39 pad
        std::vector<double> out;
        rtb200::host_rng_state saved = rtb200::host_rng();
        rtb200::host_srand(seed * 7919u + 17u);
        int made = 0, attempts = 0;
        while (made < n && attempts < n * 50) {
            attempts++;
            double fi = random_double(0, W), fj = random_double(0, H), tm = random_double();
            auto pixel = sc.cam.pixel00_loc + (fi * sc.cam.pixel_delta_u) + (fj * sc.cam.pixel_delta_v);
            ray r(sc.cam.center, pixel - sc.cam.center, tm);
            // 24-bit uniforms so that the FP32 device path can be fed the identical values
            int m[4];
            for (int k = 0; k < 4; k++) m[k] = rand() >> 7;
            hit_record rec;
            static const int zero = 0;
            g_script = &zero; g_script_len = 1; g_script_pos = 0;
            bool hit = sc.world.hit(r, interval(0.001, infinity), rec);
            g_script = nullptr;
            if (!hit) continue;
            int leaf = identify(leaves, r, rec);
            if (leaf < 0) continue;
            // second-bounce records too: scatter once with the reference, then re-hit, so that
            // inside-dielectric (front_face=false) records occur
            bool is_diel = std::dynamic_pointer_cast<dielectric>(rec.mat) != nullptr;
            // vec3::random(min,max) (vec3.h:54-56) is `vec3(rd(), rd(), rd())`: g++ evaluates the
            // constructor arguments right to left, so the FIRST draw lands in z.  Script the
            // draws so that component x gets m[0], y m[1], z m[2].
            int script_uv[3] = {m[2] << 7, m[1] << 7, m[0] << 7};
            int script_d[1] = {m[3] << 7};
            if (is_diel) { g_script = script_d; g_script_len = 1; }
            else { g_script = script_uv; g_script_len = 3; }
            g_script_pos = 0;
            ray scattered;
            color att(0, 0, 0);
            color em = rec.mat->emitted(rec.u, rec.v, rec.p);
            bool sc_ok = rec.mat->scatter(r, rec, att, scattered);
            g_script = nullptr;
            double c[40] = {0};
            c[0] = leaf;
            for (int k = 0; k < 3; k++) { c[1 + k] = r.origin()[k]; c[4 + k] = r.direction()[k]; c[9 + k] = rec.p[k]; c[12 + k] = rec.normal[k]; }
            c[7] = r.time(); c[8] = rec.t; c[15] = rec.front_face; c[16] = rec.u; c[17] = rec.v;
            for (int k = 0; k < 4; k++) c[18 + k] = m[k] / 16777216.0;
            c[22] = sc_ok;
            if (sc_ok) {
                for (int k = 0; k < 3; k++) { c[23 + k] = att[k]; c[26 + k] = scattered.origin()[k]; c[29 + k] = scattered.direction()[k]; }
                c[32] = scattered.time();
            }
            for (int k = 0; k < 3; k++) c[33 + k] = em[k];
            out.insert(out.end(), c, c + 40);
            made++;
        }
        rtb200::host_rng() = saved;
        std::vector<double> keys(leaves.size() * 10);
        for (size_t i = 0; i < leaves.size(); i++) leaf_key(leaves[i], &keys[i * 10]);
        dump(prefix + ".kat.f64", out);
        dump(prefix + ".leaves.f64", keys);
        std::printf("{\"scene\": \"%s\", \"cases\": %d, \"leaves\": %zu}\n", name.c_str(), made, leaves.size());
        return 0;
    }
    return usage();
}
