// Host-only consistency check of csrc/bvh_build.h (compiled and run by tests/test_bvh_build.py).
#include "bvh_build.h"
#include "bvh_wide.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>

using namespace rtbvh;

static int fail(const char* what) {
    std::printf("FAIL %s\n", what);
    return 1;
}

static bool inside(const float* lo, const float* hi, const Box& b) {
    for (int k = 0; k < 3; k++)
        if (b.lo[k] < lo[k] || b.hi[k] > hi[k]) return false;
    return true;
}

struct Checker {
    const std::vector<Prim>& P;
    const Result& R;
    std::vector<std::vector<uint32_t>> typed;  // typed[t][i] = input index of the i-th primitive of type t in leaf order
    std::vector<int> seen;
    size_t leaves = 0;
    bool ok = true;

    // returns the bounds of the subtree behind `link`
    Box walk(int32_t link, int depth) {
        Box b;
        if (depth > 64) { ok = false; return b; }
        if (link >= 0) {
            if ((size_t)link >= R.nodes.size()) { ok = false; return b; }
            const Node& n = R.nodes[link];
            Box l = walk(n.llink, depth + 1), r = walk(n.rlink, depth + 1);
            if (!inside(n.lmin, n.lmax, l) || !inside(n.rmin, n.rmax, r)) ok = false;
            b.grow(l);
            b.grow(r);
            return b;
        }
        uint32_t v = ~(uint32_t)link;
        uint32_t type = v >> 28, count = ((v >> 25) & 7u) + 1u, first = v & 0x1ffffffu;
        leaves++;
        if (type > 3 || first + count > typed[type].size()) { ok = false; return b; }
        for (uint32_t i = 0; i < count; i++) {
            uint32_t id = typed[type][first + i];
            if (P[id].type != type) ok = false;
            seen[id]++;
            b.grow(P[id].box);
        }
        return b;
    }
};

// The wide (quantised) BVH collapsed from R: every BVH2 leaf appears exactly once, and the host
// replay of the device traversal enters every leaf whose primitives' boxes a ray hits.
template <int W>
static int check_wide(const std::vector<Prim>& P, const Result& R, std::mt19937& g) {
    const int n = (int)P.size();
    float wlo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, whi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (const Prim& p : P)
        for (int k = 0; k < 3; k++) { wlo[k] = std::min(wlo[k], p.box.lo[k]); whi[k] = std::max(whi[k], p.box.hi[k]); }
    rtwide::Built B;
    rtwide::build_wide<W>(R, n > 0, wlo, whi, B);
    if (B.n_nodes * (size_t)rtwide::Layout<W>::kWords != B.words.size() || B.n_nodes * (size_t)W != B.refs.size()) return fail("wide: array sizes");
    // leaf links of the BVH2
    std::vector<int32_t> leaves;
    std::vector<int32_t> stack{R.root};
    if (n > 0)
        while (!stack.empty()) {
            int32_t l = stack.back();
            stack.pop_back();
            if (l < 0) { leaves.push_back(l); continue; }
            stack.push_back(R.nodes[l].llink);
            stack.push_back(R.nodes[l].rlink);
        }
    std::vector<int32_t> got;
    for (int32_t r : B.refs) if (r != 0) got.push_back(r);
    std::sort(leaves.begin(), leaves.end());
    std::sort(got.begin(), got.end());
    if (leaves != got) return fail("wide: leaf links are not a permutation of the BVH2's");
    if (B.leaf_children != leaves.size()) return fail("wide: leaf count");
    // typed[t][i] = input primitive behind the i-th record of type t
    std::vector<std::vector<uint32_t>> typed(4);
    for (uint32_t id : R.order) typed[P[id].type].push_back(id);
    std::uniform_real_distribution<float> U(-60, 60), D(-1, 1);
    uint64_t steps = 0, rays = 0;
    for (int it = 0; it < 4000 && n > 0; it++) {
        rtwide::HostRay r;
        for (int k = 0; k < 3; k++) { r.o[k] = U(g); r.d[k] = D(g); }
        if (it % 7 == 0) r.d[it % 3] = 0.0f;          // axis-parallel rays
        if (it % 11 == 0) for (int k = 0; k < 3; k++) r.o[k] = P[it % n].centroid[k];  // origins inside boxes
        r.tmin = 0.001f;
        r.tmax = it % 5 == 0 ? 30.0f : INFINITY;
        if (r.d[0] == 0 && r.d[1] == 0 && r.d[2] == 0) continue;
        std::vector<char> entered((size_t)n, 0);
        rtwide::traverse_host<W>(B, r, [&](int32_t link) {
            uint32_t v = ~(uint32_t)link;
            uint32_t type = v >> 28, count = ((v >> 25) & 7u) + 1u, first = v & 0x1ffffffu;
            for (uint32_t i = 0; i < count; i++) entered[typed[type][first + i]] = 1;
        }, &steps);
        rays++;
        for (int i = 0; i < n; i++) {
            // exact slab test in double against the primitive's own box
            double tn = r.tmin, tf = r.tmax;
            bool miss = false;
            for (int k = 0; k < 3 && !miss; k++) {
                if (r.d[k] == 0) { if (r.o[k] < P[i].box.lo[k] || r.o[k] > P[i].box.hi[k]) miss = true; continue; }
                double a = ((double)P[i].box.lo[k] - r.o[k]) / r.d[k], b = ((double)P[i].box.hi[k] - r.o[k]) / r.d[k];
                tn = std::max(tn, std::min(a, b));
                tf = std::min(tf, std::max(a, b));
            }
            if (!miss && tn <= tf && !entered[i]) return fail("wide: the traversal skipped a primitive whose box the ray hits");
        }
    }
    std::printf("wide%d nodes=%u depth=%u leaf_children=%u steps/ray=%.2f\n", W, B.n_nodes, B.depth, B.leaf_children, rays ? (double)steps / rays : 0.0);
    return 0;
}

int main(int argc, char** argv) {
    const int n = argc > 1 ? std::atoi(argv[1]) : 10000;
    const int threads = argc > 2 ? std::atoi(argv[2]) : 0;
    std::mt19937 g(n * 7 + 1);
    std::uniform_real_distribution<float> U(-50, 50), S(0.01f, 3.0f);
    std::vector<Prim> P(n);
    for (int i = 0; i < n; i++) {
        Prim& p = P[i];
        float c[3] = {U(g), U(g), U(g)};
        if (i % 17 == 0) c[0] = c[1] = c[2] = 1.0f;  // coincident centroids
        float s = S(g);
        for (int k = 0; k < 3; k++) { p.box.lo[k] = c[k] - s; p.box.hi[k] = c[k] + s; p.centroid[k] = c[k]; }
        p.type = (uint32_t)(g() % 4);
        p.index = (uint32_t)i;
        p.cost = 1.0f + 0.1f * p.type;
    }
    Result R;
    build_bvh(P, R, Tuning(), threads);
    if (n == 0) {
        if (check_wide<8>(P, R, g) || check_wide<4>(P, R, g)) return 1;
        return R.nodes.size() == 1 ? (std::printf("OK empty\n"), 0) : fail("empty");
    }
    if (R.order.size() != (size_t)n) return fail("order size");
    Checker ck{P, R, std::vector<std::vector<uint32_t>>(4), std::vector<int>(n, 0)};
    for (uint32_t id : R.order) ck.typed[P[id].type].push_back(id);
    ck.walk(R.root, 0);
    if (!ck.ok) return fail("structure");
    for (int i = 0; i < n; i++)
        if (ck.seen[i] != 1) return fail("every primitive exactly once");
    if (ck.leaves != R.leaves) return fail("leaf count");
    // determinism: one thread and many threads give the same tree
    Result R1;
    build_bvh(P, R1, Tuning(), 1);
    if (R1.nodes.size() != R.nodes.size() || R1.order != R.order) return fail("parallel build differs from sequential");
    for (size_t i = 0; i < R.nodes.size(); i++)
        if (std::memcmp(&R.nodes[i], &R1.nodes[i], sizeof(Node)) != 0) return fail("node bytes differ");
    if (int rc = check_wide<8>(P, R, g)) return rc;
    if (int rc = check_wide<4>(P, R, g)) return rc;
    std::printf("OK n=%d nodes=%zu leaves=%u depth=%u\n", n, R.nodes.size(), R.leaves, R.depth);
    return 0;
}
