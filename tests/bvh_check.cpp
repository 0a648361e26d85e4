// Host-only consistency check of csrc/bvh_build.h (compiled and run by tests/test_bvh_build.py).
#include "bvh_build.h"

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
    if (n == 0) return R.nodes.size() == 1 ? (std::printf("OK empty\n"), 0) : fail("empty");
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
    std::printf("OK n=%d nodes=%zu leaves=%u depth=%u\n", n, R.nodes.size(), R.leaves, R.depth);
    return 0;
}
