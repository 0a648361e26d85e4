// bvh_build.h — host-side binned-SAH BVH2 builder for the flattened scene.
//
// Replaces the reference's build (bvh.h:13-45: sort by bbox.min on the longest axis,
// split at the median, one object per leaf, the single object of an odd span stored as
// both children).  Closest-hit results do not depend on the tree, so the device is free
// to use a better one: 16-bin surface-area heuristic on centroid bounds, leaves of up to
// four primitives OF ONE TYPE (so a leaf's primitives are contiguous in one typed array
// and the type test is per leaf, not per primitive), children stored so that one 64-byte
// node holds both child boxes.
#pragma once
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <vector>

namespace rtbvh {

struct Box {
    float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX};
    float hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    void grow(const Box& b) {
        for (int k = 0; k < 3; k++) { lo[k] = std::min(lo[k], b.lo[k]); hi[k] = std::max(hi[k], b.hi[k]); }
    }
    void grow(const float* p) {
        for (int k = 0; k < 3; k++) { lo[k] = std::min(lo[k], p[k]); hi[k] = std::max(hi[k], p[k]); }
    }
    float area() const {
        float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
        if (dx < 0 || dy < 0 || dz < 0) return 0.0f;
        return 2.0f * (dx * dy + dy * dz + dz * dx);
    }
};

struct Prim {
    Box box;
    float centroid[3];
    uint32_t type;   // device primitive type (PT_*)
    uint32_t index;  // index in the typed source array
    float cost;      // relative intersection cost
};

struct Node {  // 64 B, the device layout
    float lmin[3]; int32_t llink;
    float lmax[3]; int32_t rlink;
    float rmin[3]; float pad0;
    float rmax[3]; float pad1;
};

struct Result {
    std::vector<Node> nodes;
    int32_t root = 0;
    std::vector<uint32_t> order;  // primitive indices (into the input vector) in leaf order
    uint32_t depth = 0, leaves = 0;
};

constexpr int kBins = 16;
struct Tuning {
    int max_leaf = 4;       // primitives per leaf (<= 8, the link encoding has 3 count bits)
    float trav_cost = 1.0f; // cost of one node visit relative to a sphere test
};

class Builder {
  public:
    Builder(std::vector<Prim>& prims, Result& out, Tuning t = Tuning()) : P(prims), R(out), tune(t) {}

    void run() {
        ids.resize(P.size());
        for (size_t i = 0; i < P.size(); i++) ids[i] = (uint32_t)i;
        type_cursor[0] = type_cursor[1] = type_cursor[2] = type_cursor[3] = 0;
        if (P.empty()) {
            // one node whose two child boxes are empty: every ray misses
            Node n{};
            for (int k = 0; k < 3; k++) { n.lmin[k] = n.rmin[k] = FLT_MAX; n.lmax[k] = n.rmax[k] = -FLT_MAX; }
            n.llink = n.rlink = 0;
            R.nodes.push_back(n);
            R.root = 0;
            return;
        }
        Box b;
        R.root = build(0, (uint32_t)P.size(), b, 1);
    }

  private:
    std::vector<Prim>& P;
    Result& R;
    Tuning tune;
    std::vector<uint32_t> ids;
    uint32_t type_cursor[4];

    bool single_type(uint32_t a, uint32_t b) const {
        for (uint32_t i = a + 1; i < b; i++)
            if (P[ids[i]].type != P[ids[a]].type) return false;
        return true;
    }

    int32_t make_leaf(uint32_t a, uint32_t b) {
        uint32_t type = P[ids[a]].type, count = b - a;
        uint32_t first = type_cursor[type];
        type_cursor[type] += count;
        for (uint32_t i = a; i < b; i++) R.order.push_back(ids[i]);
        R.leaves++;
        uint32_t v = (type << 28) | ((count - 1) << 25) | first;
        return (int32_t)~v;
    }

    // returns the link of the subtree over ids[a,b) and its bounds
    int32_t build(uint32_t a, uint32_t b, Box& bounds, uint32_t depth) {
        R.depth = std::max(R.depth, depth);
        Box cb;
        bounds = Box();
        float total_cost = 0;
        for (uint32_t i = a; i < b; i++) {
            bounds.grow(P[ids[i]].box);
            cb.grow(P[ids[i]].centroid);
            total_cost += P[ids[i]].cost;
        }
        const uint32_t n = b - a;
        if (n == 1) return make_leaf(a, b);

        int best_axis = -1, best_bin = -1;
        float best_cost = FLT_MAX;
        const float parent_area = std::max(bounds.area(), 1e-30f);
        for (int axis = 0; axis < 3; axis++) {
            float ext = cb.hi[axis] - cb.lo[axis];
            if (!(ext > 0)) continue;
            Box bin_box[kBins];
            float bin_cost[kBins] = {0};
            const float scale = kBins / ext;
            for (uint32_t i = a; i < b; i++) {
                const Prim& p = P[ids[i]];
                int bi = std::min(kBins - 1, std::max(0, (int)((p.centroid[axis] - cb.lo[axis]) * scale)));
                bin_box[bi].grow(p.box);
                bin_cost[bi] += p.cost;
            }
            float right_area[kBins], right_cost[kBins];
            Box acc;
            float c = 0;
            for (int i = kBins - 1; i > 0; i--) {
                acc.grow(bin_box[i]);
                c += bin_cost[i];
                right_area[i] = acc.area();
                right_cost[i] = c;
            }
            acc = Box();
            c = 0;
            for (int i = 0; i < kBins - 1; i++) {
                acc.grow(bin_box[i]);
                c += bin_cost[i];
                if (c == 0 || right_cost[i + 1] == 0) continue;
                float cost = tune.trav_cost + (acc.area() * c + right_area[i + 1] * right_cost[i + 1]) / parent_area;
                if (cost < best_cost) { best_cost = cost; best_axis = axis; best_bin = i; }
            }
        }
        const bool can_leaf = n <= (uint32_t)tune.max_leaf && single_type(a, b);
        if (can_leaf && (best_axis < 0 || total_cost <= best_cost)) return make_leaf(a, b);

        uint32_t mid;
        if (best_axis >= 0) {
            const float ext = cb.hi[best_axis] - cb.lo[best_axis];
            const float scale = kBins / ext;
            const float lo = cb.lo[best_axis];
            auto it = std::partition(ids.begin() + a, ids.begin() + b, [&](uint32_t id) {
                int bi = std::min(kBins - 1, std::max(0, (int)((P[id].centroid[best_axis] - lo) * scale)));
                return bi <= best_bin;
            });
            mid = (uint32_t)(it - ids.begin());
        } else {
            mid = a;  // all centroids coincide
        }
        if (mid == a || mid == b) {
            // degenerate (coincident centroids, or mixed types that binning cannot separate):
            // group by type, then halve
            std::stable_sort(ids.begin() + a, ids.begin() + b, [&](uint32_t x, uint32_t y) { return P[x].type < P[y].type; });
            mid = a + n / 2;
            for (uint32_t i = a + 1; i < b; i++)
                if (P[ids[i]].type != P[ids[i - 1]].type) { mid = i; break; }
        }
        // children are built before the parent is appended; links are patched afterwards
        Box lb, rb;
        int32_t me = (int32_t)R.nodes.size();
        R.nodes.push_back(Node{});
        int32_t l = build(a, mid, lb, depth + 1);
        int32_t r = build(mid, b, rb, depth + 1);
        Node& nd = R.nodes[me];
        const float pad = 1e-6f;
        for (int k = 0; k < 3; k++) {
            // conservative outward rounding so that FP32 slab tests never cull a true hit
            float e = pad * std::max({1.0f, std::fabs(lb.lo[k]), std::fabs(lb.hi[k])});
            nd.lmin[k] = lb.lo[k] - e; nd.lmax[k] = lb.hi[k] + e;
            e = pad * std::max({1.0f, std::fabs(rb.lo[k]), std::fabs(rb.hi[k])});
            nd.rmin[k] = rb.lo[k] - e; nd.rmax[k] = rb.hi[k] + e;
        }
        nd.llink = l;
        nd.rlink = r;
        nd.pad0 = nd.pad1 = 0;
        return me;
    }
};

}  // namespace rtbvh
