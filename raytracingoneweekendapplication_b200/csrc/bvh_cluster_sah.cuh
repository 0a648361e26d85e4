// bvh_cluster_sah.cuh — SAH rebuild of the small subtrees of the device-built radix tree.
//
// tools/tree_quality.py: the radix tree's Morton middle splits cost 26-50 % more node visits per
// ray than the host's SAH tree, and rebuilding only the TOP levels with SAH changes little — the
// loss sits in the lower levels.  So every subtree ("cluster") of 3..kClusterMax primitives
// is rebuilt here, one WARP per cluster, with the host builder's algorithm (8-bin SAH on the three
// centroid axes, leaves of <= kMaxLeaf primitives of one type where the SAH prefers them):
//
//   * the cluster's primitives are a contiguous range [a, b] of the sorted order and its old
//     subtree owns count - 1 node slots (collected by a breadth-first walk before anything is
//     overwritten: in Karras' numbering they are the indices a .. b minus ONE that belongs to an
//     ancestor); the rebuilt subtree re-uses those slots, keeps its root in the old root's slot
//     (the parent's link stays valid) and permutes the sorted order only INSIDE [a, b] — typed
//     arrays stay contiguous per leaf;
//   * everything is warp-synchronous: boxes in shared memory (SoA), bins filled with shared-memory
//     atomics, the 21 candidate splits evaluated by 21 lanes, partition by ballot/popc;
//   * a leaf of several primitives is a slot flagged `as_leaf` (what nodes_kernel and
//     clusters_kernel already understand), a single primitive is the child link ~position.
//
// Included by bvh_device.cuh (inside namespace rtlbvh).
#pragma once

constexpr int kClusterMax = 128;  // primitives per rebuilt cluster: 4 per lane
constexpr int kSahBins = 8;
constexpr int kClusterWarps = 4;  // clusters per thread block

// roots of the clusters: 3 <= count <= kClusterMax and the parent has more
__global__ void __launch_bounds__(256) cluster_roots_kernel(Scratch W, int* roots, unsigned* n_roots) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= W.n - 1) return;
    const unsigned count = __float_as_uint(W.nbox_lo[i].w);
    if (count < 3u || count > (unsigned)kClusterMax) return;
    const int p = W.parent_int[i];
    if (p >= 0 && __float_as_uint(W.nbox_lo[p].w) <= (unsigned)kClusterMax) return;
    roots[atomicAdd(n_roots, 1u)] = i;
}

struct ClusterSmem {
    float lo[3][kClusterMax], hi[3][kClusterMax];
    unsigned val[kClusterMax];               // primitive index (input order) of local element e
    unsigned char type[kClusterMax];
    unsigned char perm[kClusterMax], perm2[kClusterMax];  // current order of the local elements
    unsigned bin_lo[3][kSahBins][3], bin_hi[3][kSahBins][3];  // ordered-uint floats
    float bin_cost[3][kSahBins];
    unsigned bin_cnt[3][kSahBins];
    unsigned stack[kClusterMax];             // lo | hi << 8 | depth << 16, slot in stack_slot
    int stack_slot[kClusterMax];
    int slots[kClusterMax];                  // node slots of the old subtree, root first
};

__device__ __forceinline__ float warp_min(float v) {
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__global__ void __launch_bounds__(32 * kClusterWarps) cluster_sah_kernel(Scratch W, const int* roots, const unsigned* n_roots,
                                                                         unsigned* tallest) {
    __shared__ ClusterSmem smem[kClusterWarps];
    const unsigned lane = threadIdx.x & 31u;
    const unsigned c_idx = blockIdx.x * kClusterWarps + (threadIdx.x >> 5);
    if (c_idx >= *n_roots) return;  // whole warps leave together
    ClusterSmem& S = smem[threadIdx.x >> 5];
    const int root = roots[c_idx];
    const int a = W.range_first[root];
    const int count = (int)__float_as_uint(W.nbox_lo[root].w);
    // ---- load the cluster ----------------------------------------------------------------------
    for (int e = lane; e < count; e += 32) {
        const unsigned s = W.val_out[a + e];
        const float4 lo = W.box_lo[s], hi = W.box_hi[s];
        S.lo[0][e] = lo.x; S.lo[1][e] = lo.y; S.lo[2][e] = lo.z;
        S.hi[0][e] = hi.x; S.hi[1][e] = hi.y; S.hi[2][e] = hi.z;
        S.val[e] = s;
        S.type[e] = (unsigned char)__float_as_uint(lo.w);
        S.perm[e] = (unsigned char)e;
    }
    // ---- the node slots the old subtree owns (breadth-first over the old links) -------------------
    int n_slots = 1;
    if (lane == 0) S.slots[0] = root;
    __syncwarp();
    for (int head = 0; head < n_slots;) {
        const int end = min(n_slots, head + 32);
        const int i = head + (int)lane;
        int kids[2] = {-1, -1};
        if (i < end) {
            const int node = S.slots[i];
            kids[0] = W.left[node];
            kids[1] = W.right[node];
        }
        for (int side = 0; side < 2; side++) {
            const bool push = kids[side] >= 0;
            const unsigned m = __ballot_sync(0xffffffffu, push);
            if (push) S.slots[n_slots + __popc(m & ((1u << lane) - 1u))] = kids[side];
            n_slots += __popc(m);
        }
        head = end;
        __syncwarp();
    }
    int next_slot = 1;  // slots[0] is the root's
    auto take_slot = [&]() { return S.slots[next_slot++]; };
    int sp = 0;
    if (lane == 0) { S.stack[0] = 0u | ((unsigned)count << 8) | (0u << 16); S.stack_slot[0] = root; }
    sp = 1;
    unsigned max_depth = 0;
    __syncwarp();

    while (sp > 0) {
        sp--;
        const unsigned entry = S.stack[sp];
        const int slot = S.stack_slot[sp];
        const int lo = (int)(entry & 0xffu), hi = (int)((entry >> 8) & 0xffu);
        const unsigned depth = entry >> 16;
        const int n = hi - lo;
        max_depth = max(max_depth, depth);
        __syncwarp();
        // ---- bounds, centroid bounds, total cost, type mask of the range -------------------------
        float blo[3] = {3.4e38f, 3.4e38f, 3.4e38f}, bhi[3] = {-3.4e38f, -3.4e38f, -3.4e38f};
        float clo[3] = {3.4e38f, 3.4e38f, 3.4e38f}, chi[3] = {-3.4e38f, -3.4e38f, -3.4e38f};
        float total_cost = 0.0f;
        unsigned mask = 0u;
        for (int i = lo + (int)lane; i < hi; i += 32) {
            const int e = S.perm[i];
            for (int k = 0; k < 3; k++) {
                const float l = S.lo[k][e], h = S.hi[k][e], c = 0.5f * (l + h);
                blo[k] = fminf(blo[k], l); bhi[k] = fmaxf(bhi[k], h);
                clo[k] = fminf(clo[k], c); chi[k] = fmaxf(chi[k], c);
            }
            total_cost += prim_cost(S.type[e]);
            mask |= 1u << S.type[e];
        }
        for (int k = 0; k < 3; k++) {
            blo[k] = warp_min(blo[k]); bhi[k] = warp_max(bhi[k]);
            clo[k] = warp_min(clo[k]); chi[k] = warp_max(chi[k]);
        }
        total_cost = warp_sum(total_cost);
        mask = __reduce_or_sync(0xffffffffu, mask);
        const float parent_area = fmaxf(box_area(blo, bhi), 1e-30f);
        // ---- two primitives (a third of all ranges): the only split is one and one, no bins needed ----
        if (n == 2) {
            const int e0 = S.perm[lo], e1 = S.perm[lo + 1];
            float l0[3], h0[3], l1[3], h1[3];
            for (int k = 0; k < 3; k++) { l0[k] = S.lo[k][e0]; h0[k] = S.hi[k][e0]; l1[k] = S.lo[k][e1]; h1[k] = S.hi[k][e1]; }
            const float split_cost = kClusterTravCost + (box_area(l0, h0) * prim_cost(S.type[e0]) + box_area(l1, h1) * prim_cost(S.type[e1])) / parent_area;
            const bool pair_leaf = kMaxLeaf >= 2 && __popc(mask) == 1 && total_cost <= split_cost;
            if (lane == 0) {
                W.nbox_lo[slot] = make_float4(blo[0], blo[1], blo[2], __uint_as_float(2u));
                W.nbox_hi[slot] = make_float4(bhi[0], bhi[1], bhi[2], __uint_as_float(mask | (pair_leaf ? 0x80u : 0u)));
                W.range_first[slot] = a + lo;
                W.parent_leaf[a + lo] = slot;
                W.parent_leaf[a + lo + 1] = slot;
                W.left[slot] = ~(a + lo);
                W.right[slot] = ~(a + lo + 1);
            }
            continue;
        }
        // ---- bins ----------------------------------------------------------------------------------
        for (int j = lane; j < 3 * kSahBins; j += 32) {
            const int ax = j / kSahBins, b = j % kSahBins;
            for (int k = 0; k < 3; k++) { S.bin_lo[ax][b][k] = 0xffffffffu; S.bin_hi[ax][b][k] = 0u; }
            S.bin_cost[ax][b] = 0.0f;
            S.bin_cnt[ax][b] = 0u;
        }
        __syncwarp();
        float scale[3];
        for (int k = 0; k < 3; k++) scale[k] = chi[k] > clo[k] ? (float)kSahBins / (chi[k] - clo[k]) : 0.0f;
        for (int i = lo + (int)lane; i < hi; i += 32) {
            const int e = S.perm[i];
            for (int ax = 0; ax < 3; ax++) {
                if (scale[ax] == 0.0f) continue;
                const float c = 0.5f * (S.lo[ax][e] + S.hi[ax][e]);
                const int b = min(kSahBins - 1, max(0, (int)((c - clo[ax]) * scale[ax])));
                for (int k = 0; k < 3; k++) {
                    atomicMin(&S.bin_lo[ax][b][k], ordered(S.lo[k][e]));
                    atomicMax(&S.bin_hi[ax][b][k], ordered(S.hi[k][e]));
                }
                atomicAdd(&S.bin_cost[ax][b], prim_cost(S.type[e]));
                atomicAdd(&S.bin_cnt[ax][b], 1u);
            }
        }
        __syncwarp();
        // ---- 21 candidate splits, one per lane -----------------------------------------------------
        float my_cost = 3.4e38f;
        if (lane < 3 * (kSahBins - 1)) {
            const int ax = lane / (kSahBins - 1), split = lane % (kSahBins - 1);  // left = bins 0..split
            if (scale[ax] != 0.0f) {
                float llo[3] = {3.4e38f, 3.4e38f, 3.4e38f}, lhi[3] = {-3.4e38f, -3.4e38f, -3.4e38f};
                float rlo[3] = {3.4e38f, 3.4e38f, 3.4e38f}, rhi[3] = {-3.4e38f, -3.4e38f, -3.4e38f};
                float lc = 0.0f, rc = 0.0f;
                unsigned ln = 0u, rn = 0u;
                for (int b = 0; b < kSahBins; b++) {
                    if (S.bin_cnt[ax][b] == 0u) continue;
                    const bool left = b <= split;
                    for (int k = 0; k < 3; k++) {
                        const float l = unordered(S.bin_lo[ax][b][k]), h = unordered(S.bin_hi[ax][b][k]);
                        if (left) { llo[k] = fminf(llo[k], l); lhi[k] = fmaxf(lhi[k], h); }
                        else { rlo[k] = fminf(rlo[k], l); rhi[k] = fmaxf(rhi[k], h); }
                    }
                    if (left) { lc += S.bin_cost[ax][b]; ln += S.bin_cnt[ax][b]; }
                    else { rc += S.bin_cost[ax][b]; rn += S.bin_cnt[ax][b]; }
                }
                if (ln > 0u && rn > 0u) my_cost = kClusterTravCost + (box_area(llo, lhi) * lc + box_area(rlo, rhi) * rc) / parent_area;
            }
        }
        float best_cost = warp_min(my_cost);
        const unsigned winners = __ballot_sync(0xffffffffu, my_cost == best_cost && my_cost < 3.0e38f);
        const bool have_split = winners != 0u;
        const int best_lane = have_split ? __ffs(winners) - 1 : 0;
        const int best_axis = best_lane / (kSahBins - 1), best_split = best_lane % (kSahBins - 1);
        // ---- leaf? ----------------------------------------------------------------------------------
        const bool can_leaf = n <= kMaxLeaf && __popc(mask) == 1;
        const bool make_leaf = can_leaf && (!have_split || total_cost <= best_cost);
        const unsigned height_bits = 0u;  // heights inside a cluster are not used; the root gets the cluster's height below
        if (lane == 0) {
            W.nbox_lo[slot] = make_float4(blo[0], blo[1], blo[2], __uint_as_float((unsigned)n));
            W.nbox_hi[slot] = make_float4(bhi[0], bhi[1], bhi[2], __uint_as_float(mask | (make_leaf ? 0x80u : 0u) | (height_bits << 8)));
            W.range_first[slot] = a + lo;
        }
        if (make_leaf) {
            for (int i = lo + (int)lane; i < hi; i += 32) W.parent_leaf[a + i] = slot;
            continue;
        }
        // ---- partition perm[lo, hi) ---------------------------------------------------------------
        int mid;
        if (have_split) {
            // stable: left elements keep their order, then the right ones
            int n_left = 0;
            for (int base = lo; base < hi; base += 32) {
                const int i = base + (int)lane;
                bool left = false;
                if (i < hi) {
                    const int e = S.perm[i];
                    const float c = 0.5f * (S.lo[best_axis][e] + S.hi[best_axis][e]);
                    const int b = min(kSahBins - 1, max(0, (int)((c - clo[best_axis]) * scale[best_axis])));
                    left = b <= best_split;
                }
                n_left += __popc(__ballot_sync(0xffffffffu, left));
            }
            int wl = lo, wr = lo + n_left;
            for (int base = lo; base < hi; base += 32) {
                const int i = base + (int)lane;
                bool left = false, valid = i < hi;
                int e = 0;
                if (valid) {
                    e = S.perm[i];
                    const float c = 0.5f * (S.lo[best_axis][e] + S.hi[best_axis][e]);
                    const int b = min(kSahBins - 1, max(0, (int)((c - clo[best_axis]) * scale[best_axis])));
                    left = b <= best_split;
                }
                const unsigned lm = __ballot_sync(0xffffffffu, valid && left), rm = __ballot_sync(0xffffffffu, valid && !left);
                const unsigned below = (1u << lane) - 1u;
                if (valid && left) S.perm2[wl + __popc(lm & below)] = (unsigned char)e;
                if (valid && !left) S.perm2[wr + __popc(rm & below)] = (unsigned char)e;
                wl += __popc(lm);
                wr += __popc(rm);
            }
            __syncwarp();
            for (int i = lo + (int)lane; i < hi; i += 32) S.perm[i] = S.perm2[i];
            mid = lo + n_left;
        } else {
            // coincident centroids (or one bin): group by type if the range is mixed, else halve it
            if (__popc(mask) > 1) {
                const unsigned first_type = (unsigned)__ffs(mask) - 1u;
                int n_left = 0;
                for (int base = lo; base < hi; base += 32) {
                    const int i = base + (int)lane;
                    n_left += __popc(__ballot_sync(0xffffffffu, i < hi && S.type[S.perm[min(i, hi - 1)]] == first_type));
                }
                int wl = lo, wr = lo + n_left;
                for (int base = lo; base < hi; base += 32) {
                    const int i = base + (int)lane;
                    const bool valid = i < hi;
                    const int e = valid ? S.perm[i] : 0;
                    const bool left = valid && S.type[e] == first_type;
                    const unsigned lm = __ballot_sync(0xffffffffu, left), rm = __ballot_sync(0xffffffffu, valid && !left);
                    const unsigned below = (1u << lane) - 1u;
                    if (left) S.perm2[wl + __popc(lm & below)] = (unsigned char)e;
                    if (valid && !left) S.perm2[wr + __popc(rm & below)] = (unsigned char)e;
                    wl += __popc(lm);
                    wr += __popc(rm);
                }
                __syncwarp();
                for (int i = lo + (int)lane; i < hi; i += 32) S.perm[i] = S.perm2[i];
                mid = lo + n_left;
            } else {
                mid = lo + n / 2;
            }
        }
        __syncwarp();
        // ---- children --------------------------------------------------------------------------------
        int child[2];
        const int c_lo[2] = {lo, mid}, c_hi[2] = {mid, hi};
        for (int side = 0; side < 2; side++) {
            if (c_hi[side] - c_lo[side] == 1) {
                child[side] = ~(a + c_lo[side]);
                if (lane == 0) W.parent_leaf[a + c_lo[side]] = slot;
            } else {
                const int s2 = take_slot();
                child[side] = s2;
                if (lane == 0) {
                    W.parent_int[s2] = slot;
                    S.stack[sp] = (unsigned)c_lo[side] | ((unsigned)c_hi[side] << 8) | ((depth + 1u) << 16);
                    S.stack_slot[sp] = s2;
                }
                sp++;
            }
        }
        if (lane == 0) { W.left[slot] = child[0]; W.right[slot] = child[1]; }
        __syncwarp();
    }
    // ---- the new order of the cluster's primitives, and its height ------------------------------------
    __syncwarp();
    for (int e = lane; e < count; e += 32) W.val_out[a + e] = S.val[S.perm[e]];
    if (lane == 0) {
        const float4 h = W.nbox_hi[root];
        W.nbox_hi[root] = make_float4(h.x, h.y, h.z, __uint_as_float((__float_as_uint(h.w) & 0xffu) | ((max_depth + 1u) << 8)));
        atomicMax(tallest, max_depth + 1u);
    }
}
