// bvh_wide.h — W-wide (4 or 8) BVH with 8-bit quantised child boxes, collapsed from the
// binned-SAH BVH2 of bvh_build.h.  Replaces bvh.h:13-45 (build) and feeds the device
// traversal of rt_wide.cuh, which replaces bvh.h:64-72 + aabb.h:61-85.
//
// Why: north_star asks for "a linear array of 32-byte nodes (or a compressed BVH4/8, chosen by
// measurement)".  A wide node costs about the same ALU work per ray as the binary one, but
//   * a ray makes ~1/3 of the dependent node fetches (latency chains),
//   * a node of eight children is 80 bytes = five 128-bit loads instead of eight times 32 bytes,
//   * the traversal stack holds ONE (base, mask) entry per node instead of one per child.
//
// Node layout (the compressed-wide-BVH idea of Ylitie, Karras & Laine 2017, restated for
// this kernel; W = 8 -> 80 bytes, W = 4 -> 48 bytes), all words little-endian uint32:
//   w0..w2  origin p of the quantisation grid (float): the low corner of the node's box
//   w3      byte 0..2: biased exponents ex, ey, ez such that the FP32 number with exponent
//           byte e and zero mantissa is 2^15 * (grid step of that axis); byte 3: imask,
//           bit s set = child slot s is an interior node
//   w4      index of the first interior child; the others follow in slot order
//   w5..    W = 8: {spare, qlo_x[0..3], qlo_x[4..7]} {qlo_y, qlo_z} {qhi_x, qhi_y} {qhi_z, spare, spare}
//           W = 4: {qlo_x, qlo_y, qlo_z} {qhi_x, qhi_y, qhi_z, spare}
//           one byte per child and plane: box = p + q * step; an empty slot has lo = 255, hi = 0
// Leaf children do not live in the node: refs[W * node + slot] holds the BVH2 leaf link
// (type, count, first: see bvh_build.h) of slot `slot`, 0 for interior and empty slots.
//
// Child slots are assigned by octant (slot bit k set = the child lies towards +axis k of the
// node's centre), so that a ray with direction signs o visits the children in ascending
// (slot ^ o) order without sorting.
#pragma once
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "bvh_build.h"

namespace rtwide {

template <int W>
struct Layout;
template <>
struct Layout<8> {
    static constexpr int kWords = 20;  // 80 B
    // word index of the quantised plane `plane` (0..5 = lo x,y,z, hi x,y,z), half h (children 4h..4h+3)
    static constexpr int q_word(int plane, int h) { return 6 + 2 * plane + h; }
};
template <>
struct Layout<4> {
    static constexpr int kWords = 12;  // 48 B
    static constexpr int q_word(int plane, int) { return plane < 3 ? 5 + plane : 8 + (plane - 3); }
};

struct Built {
    int width = 0;
    std::vector<uint32_t> words;  // kWords per node, node 0 is the root
    std::vector<int32_t> refs;    // W per node
    uint32_t n_nodes = 0, depth = 0, leaf_children = 0, inner_children = 0;
};

inline float box_area(const float* lo, const float* hi) {
    float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
    if (dx < 0 || dy < 0 || dz < 0) return 0.0f;
    return 2.0f * (dx * dy + dy * dz + dz * dx);
}

struct Child {
    int32_t link;
    float lo[3], hi[3];
};

// Relative costs of the collapse decision: one wide node step against one primitive test.
struct CollapseTuning {
    float node_cost = 4.0f;   // a wide node step (~230 instructions) in units of one primitive test
    float grid_steps = 1.2f;  // how much the 8-bit grid widens a child box per axis, in grid steps (outward rounding + margin)
    bool scale_aware = true;
};

// The children of a wide node = the two children of BVH2 node `n`, with interior children replaced
// by THEIR two children (largest surface area first) until W children are reached.  A replacement
// puts two smaller boxes on the grid of node `n` (step = extent / 250): when they are small against
// that grid (a 0.2-unit sphere next to the 2000-unit ground sphere) the outward rounding inflates
// them so much that every ray "hits" them, so an interior child is only opened when the node step
// it saves is worth more than the extra visits the coarser boxes cause:
//     area(c) * node_cost  >=  sum over its two children x of (area_on_grid(x) - area(x)) * cost(x)
template <int W>
inline int gather_children(const rtbvh::Result& R, int32_t n, Child* out, const CollapseTuning& tune = CollapseTuning()) {
    auto from_node = [&](int32_t node, Child& l, Child& r) {
        const rtbvh::Node& nd = R.nodes[(size_t)node];
        l.link = nd.llink; r.link = nd.rlink;
        for (int k = 0; k < 3; k++) { l.lo[k] = nd.lmin[k]; l.hi[k] = nd.lmax[k]; r.lo[k] = nd.rmin[k]; r.hi[k] = nd.rmax[k]; }
    };
    int cnt = 2;
    from_node(n, out[0], out[1]);
    float step[3];
    for (int k = 0; k < 3; k++) {
        const float lo = std::min(out[0].lo[k], out[1].lo[k]), hi = std::max(out[0].hi[k], out[1].hi[k]);
        step[k] = tune.grid_steps * std::max(hi - lo, 0.0f) / 250.0f;
    }
    auto on_grid_area = [&](const Child& c) {
        float lo[3], hi[3];
        for (int k = 0; k < 3; k++) { lo[k] = c.lo[k] - 0.5f * step[k]; hi[k] = c.hi[k] + 0.5f * step[k]; }
        return box_area(lo, hi);
    };
    auto visit_cost = [&](const Child& c) {
        if (c.link >= 0) return tune.node_cost;
        return (float)(((~(uint32_t)c.link >> 25) & 7u) + 1u);  // primitives in the leaf
    };
    bool closed[W];
    for (int i = 0; i < W; i++) closed[i] = false;
    while (cnt < W) {
        int best = -1;
        float best_area = -1.0f;
        for (int i = 0; i < cnt; i++)
            if (out[i].link >= 0 && !closed[i]) {
                float a = box_area(out[i].lo, out[i].hi);
                if (a > best_area) { best_area = a; best = i; }
            }
        if (best < 0) break;
        Child l, r;
        from_node(out[best].link, l, r);
        if (tune.scale_aware) {
            const float extra = (on_grid_area(l) - box_area(l.lo, l.hi)) * visit_cost(l) + (on_grid_area(r) - box_area(r.lo, r.hi)) * visit_cost(r);
            if (best_area * tune.node_cost < extra) { closed[best] = true; continue; }
        }
        out[best] = l;
        closed[best] = false;
        out[cnt] = r;
        closed[cnt] = false;
        cnt++;
    }
    return cnt;
}

// Greedy assignment of children to octant slots: repeatedly the (child, slot) pair with the
// largest projection of the child's centre offset on the slot's diagonal.
// W = 8: slot bit k <-> axis k.  W = 4: slot bit b <-> axis axes[b] (the two axes along which the
// children's centres spread most; stored in the node).
template <int W>
inline void assign_slots(const Child* c, int cnt, const float* centre, const int* axes, int* slot_of) {
    float cost[W][8];
    for (int i = 0; i < cnt; i++)
        for (int s = 0; s < W; s++) {
            float v = 0;
            for (int b = 0; b < (W == 8 ? 3 : 2); b++) {
                const int k = W == 8 ? b : axes[b];
                float d = 0.5f * (c[i].lo[k] + c[i].hi[k]) - centre[k];
                v += ((s >> b) & 1) ? d : -d;
            }
            cost[i][s] = v;
        }
    bool child_done[W] = {false}, slot_used[W] = {false};
    for (int i = 0; i < W; i++) { child_done[i] = false; slot_used[i] = false; }
    for (int round = 0; round < cnt; round++) {
        int bi = -1, bs = -1;
        float bv = -FLT_MAX;
        for (int i = 0; i < cnt; i++) {
            if (child_done[i]) continue;
            for (int s = 0; s < W; s++) {
                if (slot_used[s]) continue;
                if (cost[i][s] > bv) { bv = cost[i][s]; bi = i; bs = s; }
            }
        }
        child_done[bi] = true;
        slot_used[bs] = true;
        slot_of[bi] = bs;
    }
}

// `world_lo/hi`: bounds of all primitives (used when the BVH2 is a single leaf)
template <int W>
inline void build_wide(const rtbvh::Result& R, bool has_prims, const float* world_lo, const float* world_hi, Built& out,
                       const CollapseTuning& tune = CollapseTuning()) {
    using L = Layout<W>;
    out = Built();
    out.width = W;
    struct Pending { int32_t bvh2; uint32_t depth; };
    std::vector<Pending> todo;  // todo[i] = what wide node i is built from
    auto new_node = [&](int32_t bvh2, uint32_t depth) {
        todo.push_back(Pending{bvh2, depth});
        out.words.resize(out.words.size() + L::kWords, 0u);
        out.refs.resize(out.refs.size() + W, 0);
    };
    new_node(R.root, 1);
    for (size_t ni = 0; ni < todo.size(); ni++) {
        const Pending pd = todo[ni];
        out.depth = std::max(out.depth, pd.depth);
        Child ch[W];
        int cnt = 0;
        if (!has_prims) {
            cnt = 0;
        } else if (pd.bvh2 < 0) {  // the whole world is one BVH2 leaf: a root with one leaf child
            ch[0].link = pd.bvh2;
            for (int k = 0; k < 3; k++) { ch[0].lo[k] = world_lo[k]; ch[0].hi[k] = world_hi[k]; }
            cnt = 1;
        } else {
            cnt = gather_children<W>(R, pd.bvh2, ch, tune);
        }
        float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
        for (int i = 0; i < cnt; i++)
            for (int k = 0; k < 3; k++) { lo[k] = std::min(lo[k], ch[i].lo[k]); hi[k] = std::max(hi[k], ch[i].hi[k]); }
        if (cnt == 0) lo[0] = lo[1] = lo[2] = hi[0] = hi[1] = hi[2] = 0.0f;
        float centre[3];
        for (int k = 0; k < 3; k++) centre[k] = 0.5f * (lo[k] + hi[k]);
        int slot_of[W];
        int axes[2] = {0, 1};
        if (W == 4) {  // the two axes with the widest spread of child centres
            float spread[3];
            for (int k = 0; k < 3; k++) {
                float a = FLT_MAX, b = -FLT_MAX;
                for (int i = 0; i < cnt; i++) { float m = 0.5f * (ch[i].lo[k] + ch[i].hi[k]); a = std::min(a, m); b = std::max(b, m); }
                spread[k] = cnt ? b - a : 0.0f;
            }
            int worst = 0;
            for (int k = 1; k < 3; k++) if (spread[k] < spread[worst]) worst = k;
            axes[0] = worst == 0 ? 1 : 0;
            axes[1] = worst == 2 ? 1 : 2;
        }
        assign_slots<W>(ch, cnt, centre, axes, slot_of);
        // grid: step 2^e with (hi - lo) / 2^e <= 250, so that the outward margins below stay inside [0, 255]
        int e[3];
        double step[3];
        for (int k = 0; k < 3; k++) {
            double ext = (double)hi[k] - (double)lo[k];
            int ek = -100;
            if (ext > 0) {
                ek = (int)std::ceil(std::log2(ext / 250.0));
                while (ext / std::ldexp(1.0, ek) > 250.0) ek++;
                ek = std::max(ek, -100);
            }
            e[k] = ek;
            step[k] = std::ldexp(1.0, ek);
        }
        uint32_t* w = out.words.data() + ni * L::kWords;
        uint8_t q[6][W];
        for (int s = 0; s < W; s++) {
            for (int pl = 0; pl < 3; pl++) q[pl][s] = 255;
            for (int pl = 3; pl < 6; pl++) q[pl][s] = 0;
        }
        uint32_t imask = 0;
        int32_t* refs = out.refs.data() + ni * W;
        Child by_slot[W];
        bool used[W];
        for (int s = 0; s < W; s++) used[s] = false;
        for (int i = 0; i < cnt; i++) { by_slot[slot_of[i]] = ch[i]; used[slot_of[i]] = true; }
        const uint32_t child_base = (uint32_t)todo.size();
        for (int s = 0; s < W; s++) {
            if (!used[s]) continue;
            const Child& c = by_slot[s];
            for (int k = 0; k < 3; k++) {
                // quantise outwards with 1/16 of a step to spare (the device evaluates p + q * step - o
                // with one FP32 rounding of ~2^-8 of a step), then verify in FP32
                double a = ((double)c.lo[k] - (double)lo[k]) / step[k] - 0.0625;
                double b = ((double)c.hi[k] - (double)lo[k]) / step[k] + 0.0625;
                int qa = (int)std::floor(a), qb = (int)std::ceil(b);
                qa = std::max(0, std::min(255, qa));
                qb = std::max(0, std::min(255, qb));
                while (qa > 0 && (float)((double)lo[k] + qa * step[k]) > c.lo[k]) qa--;
                while (qb < 255 && (float)((double)lo[k] + qb * step[k]) < c.hi[k]) qb++;
                q[k][s] = (uint8_t)qa;
                q[3 + k][s] = (uint8_t)qb;
            }
            if (c.link >= 0) {
                imask |= 1u << s;
                new_node(c.link, pd.depth + 1);
                w = out.words.data() + ni * L::kWords;  // new_node may have moved the arrays
                refs = out.refs.data() + ni * W;
                out.inner_children++;
            } else {
                refs[s] = c.link;
                out.leaf_children++;
            }
        }
        float pf[3] = {lo[0], lo[1], lo[2]};
        std::memcpy(&w[0], &pf[0], 4);
        std::memcpy(&w[1], &pf[1], 4);
        std::memcpy(&w[2], &pf[2], 4);
        w[3] = (uint32_t)(e[0] + 15 + 127) | ((uint32_t)(e[1] + 15 + 127) << 8) | ((uint32_t)(e[2] + 15 + 127) << 16) | (imask << 24);
        w[4] = child_base;
        if (W == 4) w[11] = (uint32_t)axes[0] | ((uint32_t)axes[1] << 2);
        for (int pl = 0; pl < 6; pl++)
            for (int h = 0; h < W / 4; h++) {
                uint32_t v = 0;
                for (int b = 0; b < 4; b++) v |= (uint32_t)q[pl][4 * h + b] << (8 * b);
                w[L::q_word(pl, h)] = v;
            }
    }
    out.n_nodes = (uint32_t)todo.size();
}

// ---------------------------------------------------------------------------------------------
// Host replay of the device traversal (rt_wide.cuh): same arithmetic, same selection rule, for the
// CPU tests (tests/bvh_check.cpp) -- returns the leaf links whose quantised boxes the ray enters.
// ---------------------------------------------------------------------------------------------
struct HostRay {
    float o[3], d[3], tmin, tmax;
};

inline float as_float(uint32_t u) {
    float f;
    std::memcpy(&f, &u, 4);
    return f;
}

// which slot of the pending set `m` a ray of octant `oct` visits first: the one with the smallest (slot ^ oct)
template <int W>
inline int select_slot(uint32_t m, uint32_t oct) {
    int best = -1;
    uint32_t best_key = 99;
    for (int s = 0; s < W; s++)
        if ((m >> s) & 1u) {
            uint32_t key = (uint32_t)s ^ (oct & (uint32_t)(W - 1));
            if (key < best_key) { best_key = key; best = s; }
        }
    return best;
}

template <int W>
inline uint32_t node_hits_host(const uint32_t* w, const HostRay& r, const float* id, float tmax) {
    using L = Layout<W>;
    uint32_t hits = 0;
    float a[3], b[3];
    for (int k = 0; k < 3; k++) {
        float s = as_float(((w[3] >> (8 * k)) & 0xffu) << 23);
        a[k] = s * id[k];
        b[k] = std::fmaf(as_float(w[k]) - r.o[k], id[k], -a[k]);
    }
    for (int s = 0; s < W; s++) {
        float tn = r.tmin, tf = tmax;
        for (int k = 0; k < 3; k++) {
            uint32_t qlo = (w[L::q_word(k, s >> 2)] >> (8 * (s & 3))) & 0xffu, qhi = (w[L::q_word(3 + k, s >> 2)] >> (8 * (s & 3))) & 0xffu;
            uint32_t qn = id[k] < 0 ? qhi : qlo, qf = id[k] < 0 ? qlo : qhi;
            float fn = as_float(0x3F800000u | (qn << 8)), ff = as_float(0x3F800000u | (qf << 8));
            tn = std::max(tn, std::fmaf(fn, a[k], b[k]));
            tf = std::min(tf, std::fmaf(ff, a[k], b[k]));
        }
        if (tn <= tf) hits |= 1u << s;
    }
    return hits;
}

inline float safe_inv_wide(float d) { return 1.0f / (std::fabs(d) > 1e-20f ? d : std::copysign(1e-20f, d)); }

// the octant a node's children are ordered by: W = 8 the ray's, W = 4 its two bits on the node's axes
template <int W>
inline uint32_t node_octant(const uint32_t* w, uint32_t oct) {
    if (W == 8) return oct;
    const uint32_t ax = w[11];
    return ((oct >> (ax & 3u)) & 1u) | (((oct >> ((ax >> 2) & 3u)) & 1u) << 1);
}

// visits every leaf child whose box the ray enters (no primitives here, so tmax never shrinks)
template <int W, class F>
inline void traverse_host(const Built& B, const HostRay& r, F&& visit_leaf, uint64_t* node_steps = nullptr) {
    using L = Layout<W>;
    if (B.n_nodes == 0) return;
    float id[3];
    for (int k = 0; k < 3; k++) id[k] = safe_inv_wide(r.d[k]);
    const uint32_t oct = (r.d[0] < 0 ? 1u : 0u) | (r.d[1] < 0 ? 2u : 0u) | (r.d[2] < 0 ? 4u : 0u);
    struct Entry { uint32_t base, mask; };
    std::vector<Entry> stack;
    uint32_t nbase = 0, nmask = 0x0101;  // bits 0..7 pending interior hits, 8..15 imask, 16..18 octant of that node
    while (true) {
        if ((nmask & 0xffu) == 0) {
            if (stack.empty()) break;
            nbase = stack.back().base;
            nmask = stack.back().mask;
            stack.pop_back();
        }
        const int slot = select_slot<W>(nmask & 0xffu, nmask >> 16);
        nmask &= ~(1u << slot);
        if (nmask & 0xffu) stack.push_back(Entry{nbase, nmask});
        const uint32_t imask_parent = (nmask >> 8) & 0xffu;
        const uint32_t node = nbase + (uint32_t)__builtin_popcount(imask_parent & ((1u << slot) - 1u));
        if (node_steps) ++*node_steps;
        const uint32_t* w = B.words.data() + (size_t)node * L::kWords;
        const uint32_t hits = node_hits_host<W>(w, r, id, r.tmax);
        const uint32_t imask = w[3] >> 24;
        uint32_t leaf = hits & ~imask;
        const uint32_t noct = node_octant<W>(w, oct);
        while (leaf) {
            const int s = select_slot<W>(leaf, noct);
            leaf &= ~(1u << s);
            visit_leaf(B.refs[(size_t)node * W + s]);
        }
        nbase = w[4];
        nmask = (hits & imask) | (imask << 8) | (noct << 16);
    }
}

}  // namespace rtwide
