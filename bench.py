#!/usr/bin/env python
"""bench.py — headline benchmark of the path-tracing hot path (BASELINE.json).

Workload: configs[4], the RTOW book-2 final scene at 3840x2160, depth 50.  A "step" is one
progressive pass of --spp-per-step samples per pixel over the whole frame (default 64, so the
default 16 steps are exactly the 1024-spp BASELINE frame).  Metric: Msamples/s (whole job).

  python bench.py [--gpus N] [--steps K] [--warmup W]           this repo's CUDA path
  python bench.py --impl reference [...]                        the reference CPU renderer

N > 1 is launched by torchrun, one process per GPU; the frame is tile-sharded (tile t -> rank
t mod N), every rank renders into a full-frame int64 fixed-point buffer and ONE NCCL int64 SUM
reduce to rank 0 (inside the timed region) finishes the frame: strong scaling, bit-identical
image at any N.

The JSON line carries, besides the contract keys: `roofline` (FP32 intersection roof, the
binding one per SURVEY 8d; the FP32 peak is measured here with an FMA microbenchmark because
MEASURED_PEAKS.json only holds HBM and bf16 numbers), `roofline_mem` (BVH-node + primitive
bytes against the measured HBM copy bandwidth), `cpu_baseline` (the unmodified reference
renderer timed on this box's host cores on a bounded sample) and `e2e` (host scene description
in, host RGB8 frame out, through rt_upload_scene / rt_render / rt_download every step).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SCENE_DEFAULT = "final"
WORKLOADS = {"final": "C5 RTOW book-2 final scene (BASELINE configs[4])", "book1": "C1 RTOW book-1 final scene (configs[0])",
             "cornell": "C2 Cornell box (configs[1])", "cornell_smoke": "C3 Cornell box with smoke (configs[2])",
             "mesh": "C4 triangle-mesh scene (configs[3])"}
FLOPS = {"box": 24, "sphere": 30, "quad": 48, "tri": 44, "boundary": 40, "shade": 50}
BYTES = {"node": 64, "sphere": 16, "quad": 48, "tri": 48, "boundary": 32, "shade": 16}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=16)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--scene", default=SCENE_DEFAULT)
    ap.add_argument("--width", type=int, default=0)
    ap.add_argument("--height", type=int, default=0)
    ap.add_argument("--spp-per-step", type=int, default=64)
    ap.add_argument("--depth", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--out-png", default="")
    return ap.parse_args()


class ClockSampler:
    """nvidia-smi clocks and throttle reasons DURING the timed region."""

    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) < 7:
                continue
            try:
                sm.append(float(r[0])); smax.append(float(r[1])); power.append(float(r[2]))
            except ValueError:
                continue
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------------------
# the reference CPU renderer (oracle/_ref when it was built from /root/reference, else the port)
# ---------------------------------------------------------------------------------------
def cpu_reference_run(scene: str, width: int, height: int, spp: int, depth: int) -> dict:
    """One timed run of the reference CPU implementation on (width x height x spp).  Returns
    {value (Msamples/s), seconds, cores, kind, sample}."""
    from raytracingoneweekendapplication_b200.assets import ensure_assets

    assets = ensure_assets()
    cores = os.cpu_count() or 1
    driver = os.path.join(ROOT, "oracle", "_ref", "ref_driver")
    sample = f"{scene} {width}x{height} x {spp} spp, depth {depth}"
    if os.path.exists(driver):
        # the reference's own camera::render (its std::async row bands over hardware_concurrency()
        # threads, Camera.txt:59-100), wall-clocked externally (its own print is wrong);
        # rand() is the thread-local interposer, g++ -O3.
        out = subprocess.check_output([driver, "time", scene, "1", assets, str(width), str(height), str(spp), str(depth)],
                                      stderr=subprocess.DEVNULL).decode()
        info = json.loads([l for l in out.splitlines() if l.startswith("{")][-1])
        return {"value": info["msamples_per_s"], "unit": "Msamples/s", "seconds": info["seconds"], "cores": info["threads"],
                "kind": "reference", "sample": sample + f" ({info['width']}x{info['height']} rendered)"}
    from oracle import port
    from raytracingoneweekendapplication_b200 import capi

    sc = capi.Scene(scene)
    t0 = time.time()
    port.render(sc, width, height, spp, depth=depth, seed=1, threads=cores)
    dt = time.time() - t0
    return {"value": width * height * spp / dt * 1e-6, "unit": "Msamples/s", "seconds": dt, "cores": cores, "kind": "port", "sample": sample}


def bounded_cpu_baseline(scene: str, width: int, height: int, depth: int, budget_s: float = 15.0) -> dict:
    """~10-30 s of CPU work on the same workload: one probe at a quarter-size frame, then the
    frame size that fits the budget at 1 spp (cost is linear in pixels x spp, Camera.txt:65-73)."""
    probe = cpu_reference_run(scene, max(64, width // 8), max(36, height // 8), 1, depth)
    rate = max(probe["value"], 1e-6) * 1e6
    want = rate * budget_s
    scale = min(1.0, (want / (width * height)) ** 0.5)
    w, h = max(64, int(width * scale) // 16 * 16), max(36, int(height * scale) // 8 * 8)
    spp = max(1, int(want / (w * h)))
    return cpu_reference_run(scene, w, h, min(spp, 64), depth)


def run_reference(args):
    """--impl reference: rank 0 times the reference CPU renderer; other ranks exit 0."""
    if int(os.environ.get("RANK", "0")) != 0:
        return 0
    from raytracingoneweekendapplication_b200 import capi

    sc = capi.Scene(args.scene)
    width, height = args.width or sc.width, args.height or sc.height
    depth = args.depth or sc.depth
    # a bounded sample per step: the full frame would take minutes per step on a CPU
    probe = cpu_reference_run(args.scene, max(64, width // 8), max(36, height // 8), 1, depth)
    rate = max(probe["value"], 1e-6) * 1e6
    budget = 120.0 / max(1, args.steps + args.warmup)           # whole run within a few minutes
    scale = min(1.0, (rate * budget / (width * height)) ** 0.5)
    w, h = max(64, int(width * scale) // 16 * 16), max(36, int(height * scale) // 8 * 8)
    runs = [cpu_reference_run(args.scene, w, h, 1, depth) for _ in range(args.warmup + args.steps)][args.warmup:]
    secs = sum(r["seconds"] for r in runs)
    value = sum(r["value"] * r["seconds"] for r in runs) / secs if secs > 0 else 0.0
    line = {
        "impl": "reference", "metric": "path-tracing throughput", "value": value, "unit": "Msamples/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * secs / max(1, len(runs)), "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{WORKLOADS.get(args.scene, args.scene)} {width}x{height}, depth {depth}",
                   "scene": args.scene, "step": runs[0]["sample"] if runs else ""},
        "cpu_baseline": {"value": value, "unit": "Msamples/s", "cores": runs[0]["cores"] if runs else 0,
                         "kind": runs[0]["kind"] if runs else "reference", "sample": runs[0]["sample"] if runs else ""},
        "e2e": {"value": value, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ---------------------------------------------------------------------------------------
# this repo's CUDA path
# ---------------------------------------------------------------------------------------
def scene_bytes(d) -> int:
    import ctypes as C

    from raytracingoneweekendapplication_b200 import capi

    n = C.sizeof(capi.rt_scene_desc)
    n += (d.n_world + d.n_boundary_refs) * C.sizeof(capi.rt_prim_ref)
    n += d.n_spheres * C.sizeof(capi.rt_sphere) + d.n_quads * C.sizeof(capi.rt_quad) + d.n_triangles * C.sizeof(capi.rt_triangle)
    n += d.n_media * C.sizeof(capi.rt_medium) + d.n_xforms * C.sizeof(capi.rt_xform)
    n += d.n_materials * C.sizeof(capi.rt_material) + d.n_textures * C.sizeof(capi.rt_texture)
    n += d.n_perlins * C.sizeof(capi.rt_perlin) + d.n_lights * C.sizeof(capi.rt_point_light)
    for i in range(d.n_images):
        n += d.images[i].width * d.images[i].height * 3
    return n


def run_b200(args):
    import torch
    import torch.distributed as dist

    from raytracingoneweekendapplication_b200 import capi, sharding

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch N>1 with torchrun (python -m torch.distributed.run --nproc-per-node N bench.py --gpus N ...)")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    sc = capi.Scene(args.scene)
    width, height = args.width or sc.width, args.height or sc.height
    depth = args.depth or sc.depth
    spp_step = args.spp_per_step
    ctx = capi.Context(local_rank)
    ctx.upload(sc)
    accum = torch.zeros(width * height * 4, dtype=torch.int64, device="cuda")
    ctx.bind_accum(accum.data_ptr(), accum.numel() * 8, width, height)
    # a non-default torch stream: its handle is what rt_render launches on (handle 0, the legacy
    # default stream, would read as "NULL = the context's own stream" in the C ABI), and the
    # torch.cuda.Events below are recorded on the same stream
    tstream = torch.cuda.Stream()
    torch.cuda.set_stream(tstream)
    stream = tstream.cuda_stream
    assert stream != 0
    plan = sharding.plan(width, height, spp_step, rank, world, sharding.RT_SHARD_AUTO)
    shard = dict(shard_rank=rank, shard_count=world, shard_mode=plan.mode)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step(i, first, stats=False):
        ctx.render(width, height, spp_step, max_depth=depth, seed=1, spp_begin=i * spp_step, accumulate=not first, stats=stats,
                   stream=stream, blocking=False, **shard)

    # ---- warm-up (untimed) ---------------------------------------------------------------
    for i in range(args.warmup):
        step(i, first=(i == 0))
    if world > 1:
        sharding.reduce_frame(accum.clone())
    barrier()
    ctx.sync()

    # ---- timed region: exactly K steps + the one frame reduce ----------------------------------
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev_k = torch.cuda.Event(enable_timing=True)
    ev0.record()
    launches = 0
    for i in range(args.steps):
        step(i, first=(i == 0))
        launches += 1
    ev_k.record()
    if world > 1:
        sharding.reduce_frame(accum)
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    ms_kernels = ev0.elapsed_time(ev_k)
    clocks = sampler.stop() if rank == 0 else None
    ctx.sync()
    t = torch.tensor([ms, ms_kernels], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_kernels = float(t[0]), float(t[1])
    total_samples = width * height * spp_step * args.steps
    value = total_samples / (ms * 1e-3) * 1e-6

    if rank == 0 and args.out_png:
        from raytracingoneweekendapplication_b200.host_png import write_png

        write_png(args.out_png, ctx.download(spp_step * args.steps, linear=False, rgb8=True))

    # ---- per-launch algorithmic work (one counted pass of the same step, untimed) ---------------
    ctx.render(width, height, spp_step, max_depth=depth, seed=1, stats=True, **shard)   # blocking: fills counters
    st = ctx.stats()
    flops = (FLOPS["box"] * st["box_tests"] + FLOPS["sphere"] * st["sphere_tests"] + FLOPS["quad"] * st["quad_tests"] +
             FLOPS["tri"] * st["triangle_tests"] + FLOPS["boundary"] * st["boundary_tests"] + FLOPS["shade"] * st["rays"])
    nbytes = (BYTES["node"] * st["node_visits"] + BYTES["sphere"] * st["sphere_tests"] + BYTES["quad"] * st["quad_tests"] +
              BYTES["tri"] * st["triangle_tests"] + BYTES["boundary"] * st["boundary_tests"] + BYTES["shade"] * st["rays"] +
              32 * plan.pixels(width, height))
    launch_ms = ms_kernels / max(1, args.steps)      # one render_kernel launch per step
    counters = torch.tensor([flops, nbytes, st["rays"]], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(counters, op=dist.ReduceOp.SUM)
    fp32_peak = ctx.measure_fp32_peak()
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    # DRAM traffic of one launch, from the committed ncu capture of this very command/config
    traffic = None
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "r1_traffic.json")))
        if tr.get("workload") == f"{args.scene} {width}x{height} {spp_step}spp" and world == 1:
            traffic = tr["traffic_bytes_per_launch"]
    except (OSError, ValueError, KeyError):
        pass
    ach_tflops = (flops / (launch_ms * 1e-3)) * 1e-12
    ach_gbs = (nbytes / (launch_ms * 1e-3)) * 1e-9

    # ---- e2e: host scene description in, host RGB8 frame out, every step ------------------------
    e2e = None
    if not args.no_e2e:
        ctx.bind_accum(None, 0, 0, 0)
        import numpy as np

        frame8 = np.empty((height, width, 3), dtype=np.uint8)
        n_e2e = max(2, min(args.steps, 8))
        # one untimed pass first: the library allocates its frame and output buffers on first use
        ctx.upload(sc)
        ctx.render(width, height, spp_step, max_depth=depth, seed=1, **shard)
        if rank == 0:
            ctx.lib.rt_download(ctx._h, spp_step, None, frame8.ctypes.data)
        barrier()
        t0 = time.perf_counter()
        t_up = t_rd = t_dl = 0.0
        for i in range(n_e2e):
            ta = time.perf_counter()
            ctx.upload(sc)                                                   # H2D: the flattened scene (+ host BVH build)
            tb = time.perf_counter()
            ctx.render(width, height, spp_step, max_depth=depth, seed=1, spp_begin=i * spp_step, **shard)
            if world > 1:
                ptr, nb = ctx.accum_buffer()
                sharding.reduce_frame(_as_tensor(ptr, nb))                   # the library-owned buffer, over NCCL
                torch.cuda.synchronize()
            tc = time.perf_counter()
            if rank == 0:
                ctx.lib.rt_download(ctx._h, spp_step, None, frame8.ctypes.data)   # D2H: the RGB8 frame
            td = time.perf_counter()
            t_up += tb - ta; t_rd += tc - tb; t_dl += td - tc
        barrier()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e = {"value": width * height * spp_step * n_e2e / float(tt[0]) * 1e-6, "unit": "Msamples/s",
               "h2d_bytes_per_step": scene_bytes(sc.desc), "d2h_bytes_per_step": width * height * 3, "steps": n_e2e,
               "call": "rt_upload_scene + rt_render + rt_download(rgb8) per step (what camera::render does), host buffers",
               "ms_upload": 1e3 * t_up / n_e2e, "ms_render": 1e3 * t_rd / n_e2e, "ms_download": 1e3 * t_dl / n_e2e}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            cpu = bounded_cpu_baseline(args.scene, width, height, depth)
        except Exception as exc:  # the baseline is a reported number, never a reason to lose the bench line
            cpu = {"value": None, "unit": "Msamples/s", "cores": os.cpu_count(), "kind": "unavailable", "sample": repr(exc)}

    if rank == 0:
        line = {
            "metric": "path-tracing throughput", "value": value, "unit": "Msamples/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {
                "workload": f"{WORKLOADS.get(args.scene, args.scene)} {width}x{height}, depth {depth}, {spp_step * args.steps} spp "
                            f"({spp_step} spp per step)",
                "scene": args.scene, "width": width, "height": height, "spp_per_step": spp_step, "depth": depth,
                "sharding": {1: "tiles 16x16 interleaved", 2: "samples interleaved"}[plan.mode] if world > 1 else "none",
                "l2": "frame accumulation buffer %d MB > 126 MB L2 is re-read every step; the scene (~1 MB) is "
                      "cache-resident by design" % (width * height * 32 // 2 ** 20),
            },
            "mrays_per_s": float(counters[2]) / (launch_ms * 1e-3) * 1e-6,
            "rays_per_sample": float(counters[2]) / (width * height * spp_step),
            "roofline": {"bound": "fp32", "achieved": ach_tflops, "peak": fp32_peak, "unit": "TFLOP/s",
                         "frac": ach_tflops / fp32_peak if fp32_peak else None, "traffic": traffic,
                         "peak_source": "measured here: FP32 FMA microbenchmark (rt_measure_fp32_peak); MEASURED_PEAKS.json has no FP32 figure",
                         "algorithmic": "24/box + 30/sphere + 48/quad + 44/triangle + 40/boundary test + 50/ray shading (SURVEY 8d), "
                                        "counted on the device for this launch", "launch_ms": launch_ms, "per": "rank 0 launch"},
            "roofline_mem": {"bound": "hbm", "achieved": ach_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": ach_gbs / hbm_peak,
                             "traffic": traffic, "algorithmic_bytes": nbytes, "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650",
                             "algorithmic": "64 B/node visit + 16/48/48 B per sphere/quad/triangle test + 16 B/ray shading fetch + "
                                            "32 B/pixel accumulation", "note": "BVH and primitives are L1/L2 resident: this is a "
                                            "cache-bandwidth figure reported against the HBM roof"},
            "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
            "device_stats": {k: st[k] for k in ("regs_per_thread", "blocks", "threads_per_block", "bvh_nodes", "bvh_depth", "nonfinite_samples",
                                                "local_bytes_per_thread")},
            "ms_render_only": ms_kernels / args.steps,
            # active lanes per warp-level iteration of each phase of the megakernel (counted pass)
            "lane_occupancy": {
                "node_step_working": st["desc_lanes"] / max(1, st["desc_iters"]),
                "node_step_holding_a_ray": st["desc_trav_lanes"] / max(1, st["desc_iters"]),
                "leaf_visit": st["leaf_lanes"] / max(1, st["leaf_iters"]),
                "shade_phase": st["shade_lanes"] / max(1, st["shade_iters"]),
                "node_steps_per_ray": st["node_visits"] / max(1, st["rays"]),
                "leaf_visits_per_ray": st["leaf_lanes"] / max(1, st["rays"]),
                "shade_phases": st["shade_iters"], "node_iters": st["desc_iters"], "leaf_iters": st["leaf_iters"],
            },
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def _as_tensor(ptr: int, nbytes: int):
    """A torch int64 view of a raw device pointer (the library-owned accumulation buffer)."""
    import torch

    class _Holder:
        pass

    h = _Holder()
    h.__cuda_array_interface__ = {"shape": (nbytes // 8,), "typestr": "<i8", "data": (ptr, False), "version": 3}
    return torch.as_tensor(h, device="cuda")


def main():
    # stdout carries exactly ONE JSON line: libraries that chat on fd 1 (NCCL prints its version
    # there) are sent to stderr for the duration of the run, the line goes to the real stdout
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real_stdout, "w", buffering=1)
    args = parse()
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
