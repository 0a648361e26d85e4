// rt_wide.cuh — device traversal of the W-wide quantised BVH of bvh_wide.h (W = 8 or 4):
// bvh.h:64-72 + aabb.h:61-85 + hittable_list.h:22-35 as a resumable state machine with the same
// interface as the binary `Trav` of rt_device.cuh, so that render_kernel_v2 runs either.
//
// One node step = five (W = 8) or three (W = 4) 128-bit loads, W slab tests on boxes decoded
// from one byte per plane, and one (base, mask) stack entry for all the children that were hit.
//   * decode: the byte q goes into the mantissa of 1.0f with one PRMT (as_float(0x3F800000 | q << 8)
//     = 1 + q * 2^-15), and ONE FFMA per plane gives the slab distance:
//     t = (1 + q 2^-15) * a + b with a = 2^15 * step / d, b = (p - o) / d - a   (exact value
//     (p + q * step - o) / d; the rounding of the FFMA is ~2^-8 of a grid step, and the builder
//     quantises outwards with 1/16 of a step to spare);
//   * order: children sit in octant slots, a ray of direction signs o takes the pending child with
//     the smallest (slot ^ o) -- front to back without sorting distances;
//   * stack: one 8-byte entry per node with children left over, in SHARED memory laid out
//     [entry][thread]: whatever depth each lane is at, the 32 lanes of a warp touch 32 different
//     banks (local memory would put every lane's entry in its own 32-byte sector: ncu measured
//     1.4-1.9 of 32 bytes used per sector for the binary kernel's stack).
#pragma once
#include "rt_device.cuh"

namespace rtdev {

template <int W>
struct WideFmt;
template <>
struct WideFmt<8> {
    static constexpr int kVec = 5;  // uint4 per node
};
template <>
struct WideFmt<4> {
    static constexpr int kVec = 3;
};

// the slots a ray of octant o prefers at each level of the choice "smallest (slot ^ o)": byte 0 for
// slot bit 2, byte 1 for slot bit 1, byte 2 for slot bit 0
__device__ __forceinline__ uint32_t octant_prefs(uint32_t o) {
    return ((o & 4u) ? 0xF0u : 0x0Fu) | ((o & 2u) ? 0xCC00u : 0x3300u) | ((o & 1u) ? 0xAA0000u : 0x550000u);
}

struct RayConstW {
    float idx, idy, idz, inv_a;
    uint32_t oct;    // bit k: direction component k is negative
    uint32_t prefs;  // octant_prefs(oct)
    __device__ __forceinline__ void set(const Ray& ray) {
        auto safe_inv = [](float d) { return __frcp_rn(fabsf(d) > 1e-20f ? d : copysignf(1e-20f, d)); };
        idx = safe_inv(ray.d.x); idy = safe_inv(ray.d.y); idz = safe_inv(ray.d.z);
        inv_a = __frcp_rn(dot(ray.d, ray.d));
        oct = (ray.d.x < 0.0f ? 1u : 0u) | (ray.d.y < 0.0f ? 2u : 0u) | (ray.d.z < 0.0f ? 4u : 0u);
        prefs = octant_prefs(oct);
    }
};

// as_float(0x3F800000 | byte k of w << 8).  `one` = 0x3F800000 held in a register: PRMT takes ONE
// immediate, and it has to be the selector (with the constant as the immediate the compiler
// materialises a selector register per use: ~40 extra moves per node step)
template <int K>
__device__ __forceinline__ float unit_plus_byte(uint32_t w, uint32_t one) {
    return __uint_as_float(__byte_perm(w, one, 0x7604u | (K << 4)));
}
// (the value travels in the kernel's parameter block, DevScene::f32_one_bits: ptxas folds anything it can see)

// the pending child (as a one-hot mask) a ray visits first: the one with the smallest (slot ^ octant);
// m != 0, m < 256, prefs = octant_prefs(octant)
template <int W>
__device__ __forceinline__ uint32_t select_child(uint32_t m, uint32_t prefs) {
    uint32_t t;
    if (W == 8) {
        t = m & prefs;
        if (t) m = t;
    }
    t = m & (prefs >> 8);
    if (t) m = t;
    t = m & (prefs >> 16);
    if (t) m = t;
    return m;
}

// Slab tests of all W children of `node`; returns the hit mask in slot space.
template <int W>
__device__ __forceinline__ uint32_t wide_node_hits(const uint4* __restrict__ nodes, uint32_t node, const Ray& ray, const RayConstW& rc,
                                                   float tmin, float tmax, uint32_t one, uint32_t& imask, uint32_t& child_base,
                                                   uint32_t& node_oct) {
    const uint4* n = nodes + (size_t)WideFmt<W>::kVec * node;
    const bool nx = rc.idx < 0.0f, ny = rc.idy < 0.0f, nz = rc.idz < 0.0f;
    uint32_t hits = 0;
    if (W == 8) {
        const uint4 n0 = __ldg(n), n1 = __ldg(n + 1), n2 = __ldg(n + 2), n3 = __ldg(n + 3), n4 = __ldg(n + 4);
        const float ax = __uint_as_float((n0.w & 0xffu) << 23) * rc.idx;
        const float ay = __uint_as_float((n0.w & 0xff00u) << 15) * rc.idy;
        const float az = __uint_as_float((n0.w & 0xff0000u) << 7) * rc.idz;
        const float bx = fmaf(__uint_as_float(n0.x) - ray.o.x, rc.idx, -ax);
        const float by = fmaf(__uint_as_float(n0.y) - ray.o.y, rc.idy, -ay);
        const float bz = fmaf(__uint_as_float(n0.z) - ray.o.z, rc.idz, -az);
        imask = n0.w >> 24;
        child_base = n1.x;
        node_oct = rc.oct;
        // near / far plane words of each axis: {children 0..3, children 4..7}
        const uint32_t nx0 = nx ? n3.x : n1.z, nx1 = nx ? n3.y : n1.w, fx0 = nx ? n1.z : n3.x, fx1 = nx ? n1.w : n3.y;
        const uint32_t ny0 = ny ? n3.z : n2.x, ny1 = ny ? n3.w : n2.y, fy0 = ny ? n2.x : n3.z, fy1 = ny ? n2.y : n3.w;
        const uint32_t nz0 = nz ? n4.x : n2.z, nz1 = nz ? n4.y : n2.w, fz0 = nz ? n2.z : n4.x, fz1 = nz ? n2.w : n4.y;
#define RT_WIDE_CHILD(J, NX, NY, NZ, FX, FY, FZ, K)                                                                            \
    {                                                                                                                          \
        const float tn = fmaxf(fmaxf(fmaf(unit_plus_byte<K>(NX, one), ax, bx), fmaf(unit_plus_byte<K>(NY, one), ay, by)),                \
                               fmaxf(fmaf(unit_plus_byte<K>(NZ, one), az, bz), tmin));                                              \
        const float tf = fminf(fminf(fmaf(unit_plus_byte<K>(FX, one), ax, bx), fmaf(unit_plus_byte<K>(FY, one), ay, by)),                \
                               fminf(fmaf(unit_plus_byte<K>(FZ, one), az, bz), tmax));                                              \
        if (tn <= tf) hits |= 1u << (J);                                                                                       \
    }
        RT_WIDE_CHILD(0, nx0, ny0, nz0, fx0, fy0, fz0, 0)
        RT_WIDE_CHILD(1, nx0, ny0, nz0, fx0, fy0, fz0, 1)
        RT_WIDE_CHILD(2, nx0, ny0, nz0, fx0, fy0, fz0, 2)
        RT_WIDE_CHILD(3, nx0, ny0, nz0, fx0, fy0, fz0, 3)
        RT_WIDE_CHILD(4, nx1, ny1, nz1, fx1, fy1, fz1, 0)
        RT_WIDE_CHILD(5, nx1, ny1, nz1, fx1, fy1, fz1, 1)
        RT_WIDE_CHILD(6, nx1, ny1, nz1, fx1, fy1, fz1, 2)
        RT_WIDE_CHILD(7, nx1, ny1, nz1, fx1, fy1, fz1, 3)
    } else {
        const uint4 n0 = __ldg(n), n1 = __ldg(n + 1), n2 = __ldg(n + 2);
        // n1 = {child_base, qlo_x, qlo_y, qlo_z}, n2 = {qhi_x, qhi_y, qhi_z, axes}
        const float ax = __uint_as_float((n0.w & 0xffu) << 23) * rc.idx;
        const float ay = __uint_as_float((n0.w & 0xff00u) << 15) * rc.idy;
        const float az = __uint_as_float((n0.w & 0xff0000u) << 7) * rc.idz;
        const float bx = fmaf(__uint_as_float(n0.x) - ray.o.x, rc.idx, -ax);
        const float by = fmaf(__uint_as_float(n0.y) - ray.o.y, rc.idy, -ay);
        const float bz = fmaf(__uint_as_float(n0.z) - ray.o.z, rc.idz, -az);
        imask = n0.w >> 24;
        child_base = n1.x;
        node_oct = ((rc.oct >> (n2.w & 3u)) & 1u) | (((rc.oct >> ((n2.w >> 2) & 3u)) & 1u) << 1);
        const uint32_t nx0 = nx ? n2.x : n1.y, fx0 = nx ? n1.y : n2.x;
        const uint32_t ny0 = ny ? n2.y : n1.z, fy0 = ny ? n1.z : n2.y;
        const uint32_t nz0 = nz ? n2.z : n1.w, fz0 = nz ? n1.w : n2.z;
        RT_WIDE_CHILD(0, nx0, ny0, nz0, fx0, fy0, fz0, 0)
        RT_WIDE_CHILD(1, nx0, ny0, nz0, fx0, fy0, fz0, 1)
        RT_WIDE_CHILD(2, nx0, ny0, nz0, fx0, fy0, fz0, 2)
        RT_WIDE_CHILD(3, nx0, ny0, nz0, fx0, fy0, fz0, 3)
#undef RT_WIDE_CHILD
    }
    return hits;
}

// Stack views: shared memory [entry][thread] for the render kernel, a local array for the
// one-ray-per-thread kernels (AOV, probes, shadow rays).
template <int STRIDE>
struct SmemStack {
    uint32_t addr;  // shared-window address of this thread's column
    __device__ __forceinline__ static SmemStack make(const uint2* block_base) {
        return SmemStack{(uint32_t)__cvta_generic_to_shared(block_base) + threadIdx.x * (uint32_t)sizeof(uint2)};
    }
    __device__ __forceinline__ uint2 load(int i) const {
        uint2 v;
        asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr + (uint32_t)i * (STRIDE * (uint32_t)sizeof(uint2))));
        return v;
    }
    __device__ __forceinline__ void store(int i, uint2 v) const {
        asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(addr + (uint32_t)i * (STRIDE * (uint32_t)sizeof(uint2))), "r"(v.x), "r"(v.y) : "memory");
    }
};
struct LocalStack {
    uint2* base;
    __device__ __forceinline__ uint2 load(int i) const { return base[i]; }
    __device__ __forceinline__ void store(int i, uint2 v) const { base[i] = v; }
};
constexpr int WIDE_LOCAL_STACK = 32;  // levels a wide tree may have for the local-stack kernels (checked at upload)

template <int W>
struct TravW {
    Hit hit;
    uint32_t nbase;  // first interior child of the node whose hit interior children are pending
    uint32_t nmask;  // bits 0..7 pending interior children (slot space), 8..15 that node's imask, 16..18 its octant
    uint32_t tgrp;   // node << 8 | pending leaf children (slot space); for W = 4 bits 4..5 hold the node's octant
    int sp;

    __device__ __forceinline__ void init(const DevScene& S, float tmax) {
        hit.t = tmax;
        hit.prim = PRIM_NONE;
        hit.u = hit.v = 0.0f;
        sp = 0;
        tgrp = 0;
        nbase = 0;
        nmask = S.n_world > 0 ? 0x0101u : 0u;  // "slot 0 of a parent whose only interior child is node 0"
    }
    __device__ __forceinline__ void clear() {
        sp = 0;
        tgrp = nbase = nmask = 0;
    }
    static constexpr uint32_t kLeafBits = W == 8 ? 0xffu : 0x0fu;
    __device__ __forceinline__ bool done() const { return (((nmask & 0xffu) | (tgrp & kLeafBits) | (uint32_t)sp) == 0u); }
    __device__ __forceinline__ bool wants_node() const { return (tgrp & kLeafBits) == 0u && ((nmask & 0xffu) | (uint32_t)sp) != 0u; }
    __device__ __forceinline__ bool at_leaf() const { return (tgrp & kLeafBits) != 0u; }

    template <bool STATS, class Stack>
    __device__ __forceinline__ void node_step(const DevScene& S, const Ray& ray, const RayConstW& rc, float tmin, const Stack& stack, Stats* st) {
        if ((nmask & 0xffu) == 0u) {
            sp--;
            const uint2 e = stack.load(sp);
            nbase = e.x;
            nmask = e.y;
        }
        const uint32_t bit = select_child<W>(nmask & 0xffu, W == 8 ? rc.prefs : octant_prefs(nmask >> 16));
        nmask &= ~bit;
        if (nmask & 0xffu) {
            stack.store(sp, make_uint2(nbase, nmask));
            sp++;
        }
        const uint32_t node = nbase + (uint32_t)__popc((nmask >> 8) & (bit - 1u));
        if (STATS) { st->node_visits++; st->box_tests += W; }
        uint32_t imask, child_base, noct;
        const uint32_t hits = wide_node_hits<W>(reinterpret_cast<const uint4*>(S.wnodes), node, ray, rc, tmin, hit.t, S.f32_one_bits, imask, child_base, noct);
        if (STATS && hits == 0u) st->empty_steps++;
        nbase = child_base;
        nmask = (hits & imask) | (imask << 8) | (W == 8 ? 0u : (noct << 16));
        tgrp = (node << 8) | (hits & ~imask) | (W == 8 ? 0u : (noct << 4));
    }

    // every pending leaf child of this lane's node: hittable_list.h:22-35 over the leaf's primitives
    template <bool STATS, bool LITE, bool MSPH = true, class Stack>
    __device__ __forceinline__ void leaf_step(const DevScene& S, const Ray& ray, const RayConstW& rc, float tmin, uint32_t origin_prim,
                                              const Stack&, Stats* st) {
        const int32_t* refs = S.wrefs + (size_t)W * (tgrp >> 8);
        uint32_t m = tgrp & kLeafBits;
        const uint32_t prefs = W == 8 ? rc.prefs : octant_prefs((tgrp >> 4) & 3u);
        tgrp = 0;
        while (m) {
            const uint32_t bit = select_child<W>(m, prefs);
            m &= ~bit;
            const uint32_t v = ~(uint32_t)__ldg(refs + (31 - __clz((int)bit)));
            const uint32_t type = v >> 28, cnt = ((v >> 25) & 7u) + 1u, first = v & 0x1ffffffu;
            for (uint32_t i = 0; i < cnt; i++) hit_prim<STATS, LITE>(S, type, first + i, ray, rc.inv_a, tmin, origin_prim, hit, st);
        }
    }
};

// one ray, start to end (AOV, probes, shadow rays)
template <int W, bool STATS>
__device__ __forceinline__ void traverse_wide(const DevScene& S, const Ray& ray, float tmin, float tmax, uint32_t origin_prim, Hit& hit,
                                              Stats* st) {
    RayConstW rc;
    rc.set(ray);
    if (STATS) st->rays++;
    TravW<W> tr;
    uint2 mem[WIDE_LOCAL_STACK];
    LocalStack stack{mem};
    tr.init(S, tmax);
    while (!tr.done()) {
        if (tr.wants_node()) tr.template node_step<STATS>(S, ray, rc, tmin, stack, st);
        else tr.template leaf_step<STATS, false>(S, ray, rc, tmin, origin_prim, stack, st);
    }
    hit = tr.hit;
}

}  // namespace rtdev
