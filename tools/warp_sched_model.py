"""Monte-Carlo model of the megakernel's warp scheduling, fed with the traversal-shape histograms the
STATS kernel measures (tools/trav_hist.py).  It replays the lane state machines of render_kernel_v2
for one warp and prices every warp-level iteration with the instruction counts of the ncu profile
(node step 70, leaf phase 273, shade phase 1190 warp-instructions), so that scheduling policies can be
compared BEFORE they are written in CUDA.  The model is calibrated when policy "v2" reproduces the
measured lane occupancy (12.0 working / 25.2 holding a ray per node step, 14.5 per leaf visit).

Usage: warp_sched_model.py gpurun_out/trav_hist.jsonl [scene]
"""
import json
import random
import sys

C_NODE, C_LEAF, C_SHADE, C_SWAP = 70.0, 273.0, 1190.0, 50.0


class RayGen:
    def __init__(self, r, rng):
        self.rng = rng
        self.first, self.between, self.tail, self.nleaf = r["first"], r["between"], r["tail"], r["leaf_visits"]
        self.tail_nz = [0] + self.tail[1:]

    def draw(self, hist):
        return self.rng.choices(range(len(hist)), weights=hist)[0]

    def ray(self):
        """Segments of node steps; a leaf visit follows every segment but the last."""
        n = self.draw(self.nleaf)
        if n == 0:
            return [self.draw(self.tail_nz)]
        segs = [self.draw(self.first)] + [self.draw(self.between) for _ in range(n - 1)]
        # rays with leaves: P(tail = 0) so that the overall tail histogram is matched
        p_leafless = self.nleaf[0] / sum(self.nleaf)
        p0 = (self.tail[0] / sum(self.tail)) / (1.0 - p_leafless)
        segs.append(0 if self.rng.random() < p0 else self.draw(self.tail_nz))
        return segs


class Ctx:
    __slots__ = ("segs", "i", "left", "done")

    def __init__(self, segs):
        self.segs, self.i, self.left, self.done = segs, 0, segs[0], False
        self.settle()

    def settle(self):  # a ray whose last segment is exhausted is finished
        if self.left == 0 and self.i == len(self.segs) - 1:
            self.done = True

    def descending(self):
        return not self.done and self.left > 0

    def at_leaf(self):
        return not self.done and self.left == 0

    def step(self):
        self.left -= 1
        self.settle()

    def leaf(self):
        self.i += 1
        self.left = self.segs[self.i]
        self.settle()


def simulate(gen, policy, n_rays=60000, threshold=4, leaf_t=32, contexts=1):
    cost = dict(node=0.0, leaf=0.0, shade=0.0, swap=0.0)
    occ = dict(node_iters=0, node_work=0, node_hold=0, leaf_iters=0, leaf_work=0, shade_iters=0, shade_work=0)
    cur = [None] * 32
    park = [[None] * 32 for _ in range(contexts - 1)]
    issued = finished = 0

    def fresh():
        nonlocal issued
        issued += 1
        return Ctx(gen.ray())

    while finished < n_rays:
        # ---- shade rounds: finished (or empty) contexts are shaded / regenerated -------------
        for rnd in range(contexts):
            work = 0
            for l in range(32):
                if rnd > 0:
                    p = park[rnd - 1][l]
                    if p is None or p.done:       # the parked context needs service: swap it in
                        park[rnd - 1][l], cur[l] = cur[l], p
                        serviced = True
                    else:
                        continue
                c = cur[l]
                if c is None or c.done:
                    if c is not None:
                        finished += 1
                    cur[l] = fresh()
                    work += 1
                if rnd > 0:
                    # keep an in-flight traversal in the registers: swap back if the parked one is in flight
                    p = park[rnd - 1][l]
                    if p is not None and not p.done and (p.i > 0 or p.left < p.segs[0]):
                        park[rnd - 1][l], cur[l] = cur[l], p
            if work:
                cost["shade"] += C_SHADE
                occ["shade_iters"] += 1
                occ["shade_work"] += work
                if rnd > 0:
                    cost["swap"] += 2 * C_SWAP
        # ---- traversal phase -----------------------------------------------------------------------
        while True:
            while True:
                d = [l for l in range(32) if cur[l].descending()]
                if not d:
                    break
                if policy != "v2" and sum(cur[l].at_leaf() for l in range(32)) >= leaf_t:
                    break
                cost["node"] += C_NODE
                occ["node_iters"] += 1
                occ["node_work"] += len(d)
                occ["node_hold"] += sum(not cur[l].done for l in range(32))
                for l in d:
                    cur[l].step()
            lv = [l for l in range(32) if cur[l].at_leaf()]
            if lv:
                cost["leaf"] += C_LEAF
                occ["leaf_iters"] += 1
                occ["leaf_work"] += len(lv)
                for l in lv:
                    cur[l].leaf()
            # finished lanes take their parked ray, if it is waiting for traversal
            swapped = False
            for k in range(contexts - 1):
                for l in range(32):
                    p = park[k][l]
                    if cur[l].done and p is not None and not p.done:
                        park[k][l], cur[l] = cur[l], p
                        swapped = True
            if swapped:
                cost["swap"] += C_SWAP
            active = sum(not cur[l].done for l in range(32))
            if active < threshold:
                break
    total = sum(cost.values())
    return {
        "policy": policy, "contexts": contexts, "threshold": threshold, "leaf_t": leaf_t,
        "warp_instr_per_ray": round(total / finished, 1), **{k: round(v / finished, 1) for k, v in cost.items()},
        "node_work": round(occ["node_work"] / max(1, occ["node_iters"]), 2), "node_hold": round(occ["node_hold"] / max(1, occ["node_iters"]), 2),
        "leaf_work": round(occ["leaf_work"] / max(1, occ["leaf_iters"]), 2), "shade_work": round(occ["shade_work"] / max(1, occ["shade_iters"]), 2),
        "node_iters_per_leaf_phase": round(occ["node_iters"] / max(1, occ["leaf_iters"]), 2),
    }


def main():
    path = sys.argv[1]
    want = sys.argv[2] if len(sys.argv) > 2 else "final"
    rec = [json.loads(l) for l in open(path)]
    r = [x for x in rec if x["scene"] == want][0]
    gen = RayGen(r, random.Random(1))
    print(json.dumps(simulate(gen, "v2")))
    for thr in (8, 12):
        print(json.dumps(simulate(gen, "v2", threshold=thr)))
    for ctxs in (2, 3):
        for thr in (4, 8, 12):
            for lt in (8, 12, 16, 20, 32):
                print(json.dumps(simulate(gen, "multi", threshold=thr, leaf_t=lt, contexts=ctxs)))


if __name__ == "__main__":
    main()
