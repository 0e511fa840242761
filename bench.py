#!/usr/bin/env python
"""bench.py — headline benchmark of the B200-native segmentation hot path.

Headline (BASELINE.json `metric`, first part; workload C1/C4 = configs[0] at full resolution / configs[3]):
  segmented frames/s on 640 x 480 = 307 200-point Kinect-shaped tabletop frames (table plane + sphere, cylinder, cone),
  every frame a distinct seed of the C4 stream with random object poses. A frame = depthAcquisition + clustersAcquisition
  (obj_segmentation.cpp:251-316, ransac_segmentation.cpp:223-343): k = 50 normals, the supports loop (plane RANSAC,
  index maps, points on the plane), Euclidean clustering, per cluster normals + four RANSAC primitive fits with
  refinement, the selection rule, TrackedShapes out.
  step    = one batch of --frames-per-step frames per GPU through the frame stream (pitt_segment_*_batched).
  value   = frames/s with the clouds already staged in HBM (pitt_segment_clouds_batched).
  e2e     = the same batch from pinned HOST buffers through pitt_segment_frames_batched: the H2D copy of every frame and
            the D2H of its result are inside the timed region.
  N > 1   = weak scaling: every rank (one process per GPU) segments its own block of the stream, no collective on the data
            path; the barrier + max-over-ranks timing of the contract stay.
  parity  = outside the timed region one frame per rank (and the first frame's raw faithful path on rank 0) is compared
            field by field with the CPU oracle: `parity_checked`.

Second part of the metric ("RANSAC hypothesis-point evals/s vs peak", configs[1] = C2): the `ransac` and `roofline` keys
(1 M points x 5000 replayed hypotheses, plane_tc_kernel timed alone by CUDA events recorded around its launch).

--workload c2 | c3 | c5 make those configurations the printed metric instead (C5 / C2: hypothesis split through
pitt_sac_segment_split with an NCCL all-gather, strong scaling for C5; C3: 64 clusters size-balanced over the ranks),
each with an oracle spot check on every rank.

`--impl reference` times the CPU path (the oracle = the PCL restatement; the reference itself cannot be built here, see
DESIGN.md) on the host cores on a bounded sample of the same workload.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# the frame stream drives many CUDA streams per GPU; with the default 8 hardware queues streams that share
# a queue serialise falsely (measured 244 -> 362 frames/s at 4 contexts). Must be set before CUDA initialises.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
# rank 0 prints exactly one JSON line on stdout: keep NCCL's "NCCL version ..." banner (NCCL_DEBUG=VERSION) off it
if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
    os.environ["NCCL_DEBUG"] = "WARN"
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # the banner is printed at every level >= VERSION: send it to stderr

import numpy as np

FRAME_W, FRAME_H = 640, 480
METRIC = "segmented frames/s (307k-pt cloud)"
WORKLOAD_C1 = ("C1/C4: 640x480 = 307200-point Kinect-shaped tabletop frames (table plane + sphere, cylinder, cone; distinct seeds, "
               "random poses), the whole frame path (normals, supports loop, clustering, 4 primitive fits per cluster, selection)")
WORKLOAD_C2 = ("C2: table-plane RANSAC only, 1M-point synthetic cloud x 5000 replayed mt19937(12345) hypotheses per GPU, thr 0.02, "
               "ALL_H, refine + select")
UNIT = "frames/s"
N_POINTS = 1_000_000  # C2
N_HYP = 5000
FLOP_PER_EVAL = 6  # plane: 3 mul + 3 add, unfused


def _measured_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


MEASURED = _measured_peaks()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    def __init__(self, index):
        self.rows = []
        self.proc = None
        self.index = index

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", os.environ.get("PITT_BENCH_CLOCK_MS", "100")],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def wait_first(self, timeout=5.0):
        """nvidia-smi needs ~0.1 s to start: block until it has delivered a row, then forget the idle rows"""
        t0 = time.perf_counter()
        while not self.rows and time.perf_counter() - t0 < timeout and self.proc:
            time.sleep(0.01)
        self.rows.clear()

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1])); power.append(float(r[2]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------ workloads
def frame_seeds(rank, world, per_gpu):
    """C4: one stream of world * per_gpu frames with seeds 0, 1, 2, ...; rank r owns a contiguous block"""
    from pitt_object_table_segmentation_b200 import sharding
    lo, hi = sharding.block_range(rank, world, per_gpu * world)
    return list(range(lo, hi))


def make_frames(seeds, raw=False):
    """distinct full-resolution frames (random object poses), generated on a few host threads (numpy releases the GIL)"""
    from concurrent.futures import ThreadPoolExecutor
    from pitt_object_table_segmentation_b200 import scenes
    gen = (lambda s: scenes.raw_camera_frame(seed=s, width=FRAME_W, height=FRAME_H, random_poses=True)) if raw else \
          (lambda s: scenes.tabletop_frame(seed=s, width=FRAME_W, height=FRAME_H, random_poses=True))
    with ThreadPoolExecutor(min(8, os.cpu_count() or 1)) as ex:
        return list(ex.map(gen, seeds))


def faithful_prefilter():
    import pitt_object_table_segmentation_b200 as pkg
    from pitt_object_table_segmentation_b200 import scenes
    pf = pkg.default_prefilter_params()
    c2w, _ = scenes.camera_pose()
    for i, v in enumerate(c2w.ravel()):
        pf.transform[i] = float(v)
    return pf


def frames_equal(got, want):
    """every field of the frame response, bit for bit (same comparison as tests/test_gpu_c1_full_res.py)"""
    def eq(a, b):
        return np.array_equal(np.asarray(a, np.float32).view(np.uint32), np.asarray(b, np.float32).view(np.uint32))
    if (got["n_supports"], got["n_clusters"]) != (want["n_supports"], want["n_clusters"]):
        return False
    if got["support_sizes"] != want["support_sizes"] or got["on_support_sizes"] != want["on_support_sizes"]:
        return False
    if not eq(got["support_coefficients"], want["support_coefficients"]) or len(got["shapes"]) != len(want["shapes"]):
        return False
    for g, w in zip(got["shapes"], want["shapes"]):
        if (g["tag"], g["n_points"], g["inliers"], g["object_id"]) != (w["tag"], w["n_points"], w["inliers"], w["object_id"]):
            return False
        if not (eq(g["coefficients"], w["coefficients"]) and eq(g["pc_centroid"], w["pc_centroid"]) and
                eq(g["est_centroid"], w["est_centroid"])):
            return False
    return True


def cpu_frames_baseline(frames, threads=None, rounds=1, single=True):
    """The oracle's frame path (the reference's algorithm on its PCL restatement) on the host: frame-parallel on all cores,
    `rounds` frames per thread; plus (single) one frame on one core, which is what the reference's blocking service chain uses."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import orc_binding as O
    fp = O.default_frame_params()
    threads = threads or (os.cpu_count() or 1)

    def one(f):
        return O.segment_frame(f, fp)["n_clusters"]

    single_s = None
    if single:
        t0 = time.perf_counter()
        one(frames[0])
        single_s = time.perf_counter() - t0
    work = [frames[i % len(frames)] for i in range(threads * rounds)]
    t0 = time.perf_counter()
    with ThreadPoolExecutor(threads) as ex:  # ctypes releases the GIL inside the oracle
        list(ex.map(one, work))
    dt = time.perf_counter() - t0
    out = {"value": len(work) / dt, "unit": UNIT, "cores": threads, "kind": "port",
           "sample": f"{len(work)} full-resolution frames ({min(len(work), len(frames))} distinct) on {threads} threads, oracle = PCL "
                     f"restatement (-O2), {dt:.1f} s wall"}
    if single:
        out["frames_per_s_1core"] = 1.0 / single_s
        out["sample"] += f"; one frame on one core: {single_s:.2f} s"
    return out


def cpu_c2_baseline(xyz, samples, seconds_target=12.0, threads=None):
    """The oracle scoring a bounded hypothesis sample of the C2 cloud on the host (countWithinDistance)."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import orc_binding as O
    threads = threads or (os.cpu_count() or 1)
    p = O.default_support_sac_params()
    t0 = time.perf_counter()
    O.sac_score(xyz, None, p, samples[:2])
    per_h = (time.perf_counter() - t0) / 2
    h_total = int(max(threads, min(len(samples), seconds_target / per_h * threads)))
    h_total -= h_total % threads
    chunks = np.array_split(samples[:h_total], threads)
    t0 = time.perf_counter()
    with ThreadPoolExecutor(threads) as ex:
        res = list(ex.map(lambda s: O.sac_score(xyz, None, p, s)[0], chunks))
    dt = time.perf_counter() - t0
    return {"value": float(h_total) * xyz.shape[0] / dt, "unit": "evals/s", "cores": threads, "kind": "port",
            "sample": f"{h_total} of {len(samples)} hypotheses x {xyz.shape[0]} points, oracle countWithinDistance, {dt:.1f} s wall",
            "checksum": int(sum(int(r.sum()) for r in res))}


# ------------------------------------------------------------------------------------------------ reference arm
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "liborc.so"], stdout=subprocess.DEVNULL)
    threads = os.cpu_count() or 1
    if args.workload in ("c2", "c5"):
        from oracle import orc_binding as O
        from pitt_object_table_segmentation_b200 import _abi as A, scenes
        n = N_POINTS if args.workload == "c2" else min(args.c5_points, 4_000_000)  # bounded sample of the C5 scene
        xyz = scenes.plane_outlier_cloud(n, seed=12345 if args.workload == "c2" else 5)
        samples = O.pcl_sample_stream(xyz, A.MODEL_PLANE, N_HYP)
        per_step = max(2.0, min(20.0, 120.0 / max(1, args.steps + args.warmup)))
        vals, last = [], None
        for i in range(args.warmup + args.steps):
            last = cpu_c2_baseline(xyz, samples, seconds_target=per_step, threads=threads)
            if i >= args.warmup:
                vals.append(last["value"])
        value = float(np.mean(vals))
        full = float(N_POINTS) * N_HYP if args.workload == "c2" else float(args.c5_points) * args.c5_hypotheses
        line = {"impl": "reference", "metric": "RANSAC hypothesis-point evals/s", "value": value, "unit": "evals/s",
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": full / value * 1e3,
                "higher_is_better": True, "scaling": "weak" if args.workload == "c2" else "strong", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": {"workload": WORKLOAD_C2 if args.workload == "c2" else
                                       f"C5: {args.c5_points}-point scene resident on every GPU, {args.c5_hypotheses} plane hypotheses split "
                                       "over the ranks (pitt_sac_segment_split, NCCL all-gather of counts, arg-max, refine + select)",
                           "cpu_sample": "bounded hypothesis sample per step, ms_per_step extrapolated to the full job"},
                "cpu_baseline": {"value": value, "unit": "evals/s", "cores": threads, "kind": "port", "sample": last["sample"]},
                "e2e": {"value": value, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
        print(json.dumps(line))
        return
    if args.workload == "c3":
        from concurrent.futures import ThreadPoolExecutor
        from oracle import orc_binding as O
        from pitt_object_table_segmentation_b200 import _abi as A, scenes
        sizes = np.linspace(5000, 50000, 64).astype(int)
        pick = [0, 21, 42, 63][: max(1, min(4, threads))]  # bounded sample: 4 of the 64 clusters, one per thread

        def one(i):
            kind, model = (("cylinder", A.MODEL_CYLINDER), ("cone", A.MODEL_CONE))[i % 2]
            cxyz, _ = scenes.primitive_cluster(kind, int(sizes[i]), 100 + i)
            t0 = time.perf_counter()
            nrm = O.estimate_normals(cxyz, 50)
            pj = O.default_sac_params(model)
            pj.max_iterations = 10000
            O.sac_segment(cxyz, nrm, pj)
            return time.perf_counter() - t0, int(sizes[i])

        vals = []
        for i in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            with ThreadPoolExecutor(len(pick)) as ex:
                r = list(ex.map(one, pick))
            dt = time.perf_counter() - t0
            if i >= args.warmup:
                vals.append(len(pick) / dt)
            if time.perf_counter() - t0 > 60 and i >= args.warmup:
                break
        value = float(np.mean(vals))
        line = {"impl": "reference", "metric": "C3 clusters/s (normals + cylinder/cone RANSAC, 10k iterations)", "value": value,
                "unit": "clusters/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 64 / value * 1e3,
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": "C3: 64 clusters of 5k-50k points (bounded sample: 4 clusters on 4 threads, extrapolated)"},
                "cpu_baseline": {"value": value, "unit": "clusters/s", "cores": len(pick), "kind": "port",
                                 "sample": f"clusters {pick} (sizes {[int(sizes[i]) for i in pick]}), one per thread"},
                "e2e": {"value": value, "unit": "clusters/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
        print(json.dumps(line))
        return
    # headline: frames/s of the oracle's frame path, frame-parallel on all host threads; one step = one frame per thread
    frames = make_frames(list(range(min(threads, 64))))
    vals, last = [], None
    t_all = time.perf_counter()
    steps_done = 0
    for i in range(args.warmup + args.steps):
        last = cpu_frames_baseline(frames, threads=threads, rounds=1, single=False)
        if i >= args.warmup:
            vals.append(last["value"])
            steps_done += 1
        if time.perf_counter() - t_all > 150 and steps_done >= 3:  # bounded: the whole run ends within a few minutes
            break
    value = float(np.mean(vals))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": args.frames_per_step / value * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD_C1, "frames_per_gpu_per_step": args.frames_per_step,
                       "cpu_sample": f"bounded sample of {threads} frames per step, one per host thread; ms_per_step extrapolated to "
                                     f"the {args.frames_per_step}-frame step", "steps_measured": steps_done},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": last["sample"]},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------ GPU arms
class Env:
    """rank / device / distributed plumbing shared by the arms"""

    def __init__(self):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, v):
        t = self.torch.tensor([float(v)], dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(self, v):
        t = self.torch.tensor([float(v)], dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t.item())

    def min_over_ranks(self, v):
        t = self.torch.tensor([float(v)], dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MIN)
        return float(t.item())

    def timed_steps(self, fn, steps, warmup, before_each=None):
        """W untimed steps, then K steps each bracketed by CUDA events on torch's current stream (the step function returns
        only when its device work is complete), barrier + synchronize on both sides, max over ranks"""
        torch = self.torch
        for _ in range(warmup):
            fn()
        self.barrier()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        t_wall = time.perf_counter()
        for i in range(steps):
            if before_each:
                before_each()
            ev[i][0].record()
            fn()
            ev[i][1].record()
        self.barrier()
        wall = time.perf_counter() - t_wall
        ms = sum(a.elapsed_time(b) for a, b in ev)
        return self.max_over_ranks(ms) / steps, wall


def arm_frames(env, args, pkg):
    """headline: the frame stream on full-resolution frames, resident and end to end, + parity check + CPU baseline"""
    torch = env.torch
    from pitt_object_table_segmentation_b200 import _results as R
    per_gpu = args.frames_per_step
    # 16 contexts saturate one B200 when every host thread has a core (spinning waits); with fewer cores than threads the waits
    # sleep and a few more contexts cover the wake-up latency (measured with 2 GPUs on 8 cores: 16 -> 3131, 24 -> 3202 frames/s)
    cores = len(os.sched_getaffinity(0))
    n_ctx = args.frame_contexts if args.frame_contexts > 0 else (16 if env.world * 16 <= cores else 24)
    seeds = frame_seeds(env.rank, env.world, per_gpu)
    t0 = time.perf_counter()
    frames_np = make_frames(seeds)
    pinned = [torch.from_numpy(f).pin_memory() for f in frames_np]
    frames = [p.numpy() for p in pinned]
    gen_s = time.perf_counter() - t0
    fctxs = [pkg.Context(env.local_rank, seed=12345) for _ in range(n_ctx)]
    oversubscribed = env.world * n_ctx > len(os.sched_getaffinity(0))  # more host threads than cores: waits sleep instead of spinning
    for c in fctxs:
        c.set_workers(args.frame_workers)
        c.set_blocking_sync(oversubscribed)
    stager = pkg.Context(env.local_rank, seed=12345)
    clouds = [stager.stage_host_ptr(p.data_ptr(), 16, int(p.shape[0])) for p in pinned]  # resident arm: staged once
    bufs = [R.FrameBuffers(64) for _ in frames]
    out = {}

    def step_resident():
        out["res"] = pkg.segment_clouds_batched(fctxs, clouds, bufs=bufs)

    def step_e2e():
        out["res_e2e"] = pkg.segment_frames_batched(fctxs, frames)

    def launches():
        return sum(c.kernel_launches for c in fctxs)

    clocks = ClockSampler(env.local_rank)
    clocks.start()
    clocks.wait_first()
    # warm-up steps run inside timed_steps; count the launches of the timed region only
    for _ in range(args.warmup):
        step_resident()
    l0 = launches()
    cpu0 = time.process_time()  # user + system time of every thread of this process
    ms_step, wall = env.timed_steps(step_resident, args.steps, 0)
    cpu_s = time.process_time() - cpu0
    n_launch = launches() - l0
    clk = clocks.stop()
    clk["note"] = "sampled every 100 ms over the timed region of the headline (resident) arm"
    e2e_steps = max(3, min(args.steps, 10))
    e2e_ms, _ = env.timed_steps(step_e2e, e2e_steps, 2)
    value = per_gpu * env.world / (ms_step * 1e-3)
    e2e_value = per_gpu * env.world / (e2e_ms * 1e-3)

    # ---- parity, outside the timed region: this rank's first frame (resident and end to end) against the oracle
    from oracle import orc_binding as O
    want = O.segment_frame(frames_np[0], O.default_frame_params())
    ok_res = frames_equal(out["res"][0], want)
    ok_e2e = frames_equal(out["res_e2e"][0], want)
    same = all(frames_equal(a, b) for a, b in zip(out["res"], out["res_e2e"]))
    all_ok = env.min_over_ranks(1.0 if (ok_res and ok_e2e and same) else 0.0) == 1.0
    parity = {"ok": bool(all_ok), "frames_vs_oracle_per_rank": 1, "ranks": env.world,
              "resident_equals_e2e_on_all_frames": bool(env.min_over_ranks(1.0 if same else 0.0) == 1.0),
              "rank0_seed": seeds[0], "rank0_shapes": [s["tag_name"] for s in out["res"][0]["shapes"]],
              "rank0_counts": {"supports": out["res"][0]["n_supports"], "clusters": out["res"][0]["n_clusters"],
                               "support_size": out["res"][0]["support_sizes"][:1]},
              "note": "every field of the frame response bit for bit against the CPU oracle (outside the timed region); "
                      "tests/test_gpu_c1_full_res.py holds the full-size parity suite"}
    shapes_hist = {}
    for r in out["res"]:
        for s in r["shapes"]:
            shapes_hist[s["tag_name"]] = shapes_hist.get(s["tag_name"], 0) + 1

    info = {
        "value": value, "ms_per_step": ms_step, "wall_s": wall, "launches": int(n_launch), "clocks": clk,
        "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": int(sum(f.nbytes for f in frames)),
                "d2h_bytes_per_step": int(sum(C.sizeof(b.res) + r["n_clusters"] * C.sizeof(b.shapes[0]) for b, r in zip(bufs, out["res_e2e"])))},
        "parity": parity, "contexts_per_gpu": n_ctx, "workers_per_context": args.frame_workers,
        "blocking_sync": bool(oversubscribed), "host_cores": len(os.sched_getaffinity(0)),
        "host_cpu_ms_per_frame": 1e3 * cpu_s / float(max(1, args.steps * per_gpu)), "frames_per_gpu_per_step": per_gpu,
        "distinct_frames": per_gpu * env.world, "points_per_frame": int(frames[0].shape[0]),
        "shape_tags_of_this_rank": shapes_hist, "frame_generation_s": gen_s,
        "launches_per_frame": n_launch / float(max(1, args.steps * per_gpu)),
    }
    # ---- the dominant kernel group of the frame path, timed alone: whole-cloud normals (grid build + knn_collect_kernel +
    # knn_finish_kernel) on distinct resident frames (629 MB of clouds in rotation: nothing is warm in L2), CUDA events
    # recorded by the library on the context's stream
    nctx = stager  # the context that owns the resident clouds (the normals are attached to them from its pool)
    nsub = clouds[: min(len(clouds), 64)]
    for cl in nsub:  # first round: the normal arrays are attached to the clouds (allocation), the arena grows
        nctx.estimate_normals_device(cl, 50)
    nms = []
    for cl in nsub:  # second round, timed: 64 x 4.9 MB of clouds + 157 MB of keys per call went through L2 since this cloud's turn
        nctx.estimate_normals_device(cl, 50)
        nms.append(nctx.last_device_ms)
    normals_ms = float(np.median(nms))
    n_pts = int(frames[0].shape[0])
    alg_bytes = n_pts * 32  # 16 B per point in, 16 B per normal out
    info["normals"] = {
        "ms": normals_ms, "points": n_pts, "k": 50,
        "frame_stream_ms_per_frame_per_gpu": (1e3 * env.world / value) if value > 0 else None,
        "queries_per_s": n_pts / (normals_ms * 1e-3),
        "note": "whole-cloud NormalEstimation (pc_manager.cpp:68-78) alone on one stream (latency-bound there: about as long as a "
                "whole frame costs inside the 16-context stream, where other frames' kernels fill its stalls)",
        "roofline": {"bound": "hbm", "kernel": "knn_collect_kernel + knn_finish_kernel (+ 9 grid-build launches)",
                     "achieved": alg_bytes / (normals_ms * 1e-3) / 1e9, "peak": MEASURED.get("hbm_gbs"), "unit": "GB/s",
                     "frac": ((alg_bytes / (normals_ms * 1e-3) / 1e9) / MEASURED["hbm_gbs"]) if MEASURED.get("hbm_gbs") else None,
                     "traffic": 306.0e6, "algorithmic_bytes_per_launch": alg_bytes,
                     "traffic_note": "dram__bytes_read+write of the two kernels from profiles/r02_knn_ncu.md (the 157 MB key "
                                     "array is written by the first kernel and read by the second)",
                     "note": "an exact 50-nearest-neighbour search is not a bandwidth problem: 15 000 executed thread "
                             "instructions per query (about 480 candidate distances, twice), issue- and latency-bound; "
                             "the HBM fraction is reported because SURVEY 8d files the kernel under HBM/L2"}}
    # ---- single-frame latency: one context, the fits of a frame fanned out to 4 helper streams
    lctx = pkg.Context(env.local_rank, seed=12345)
    lctx.set_workers(0)
    lat, lat_dev = [], []
    for rnd in range(2):  # first round: the context's arena and pools grow to the sizes these frames need
        for i in range(8):
            t1 = time.perf_counter()
            cl = lctx.stage_host_ptr(pinned[i % len(pinned)].data_ptr(), 16, int(pinned[i % len(pinned)].shape[0]))
            fr1 = lctx.segment_frame(cl)
            cl.release()
            if rnd == 1:
                lat.append((time.perf_counter() - t1) * 1e3)
                lat_dev.append(float(fr1["device_ms"]))
    info["frame_latency_ms"] = float(np.median(lat))
    info["frame_latency_device_ms"] = float(np.median(lat_dev))
    info["frame_latency_note"] = ("one frame at a time on one context: stage (H2D from pinned memory) + segment_frame + results, wall "
                                  "clock; device_ms = the segment_frame part between two CUDA events")
    lctx.close()

    # ---- faithful variant (BASELINE configs[0] as the reference runs it): the raw camera-frame message through fromROSMsg + 1 cm
    # VoxelGrid + deep filter + transform on the device, then the frame path
    if args.faithful:
        pf = faithful_prefilter()
        raws_np = make_frames(seeds[: min(len(seeds), 64)], raw=True)
        raw_pinned = [torch.from_numpy(f).pin_memory() for f in raws_np]
        raws = [p.numpy() for p in raw_pinned]
        for _ in range(max(3, args.warmup)):  # the arenas / pools of every context grow to their steady size during the first calls
            pkg.segment_frames_batched(fctxs, raws, prefilter=pf)
        torch.cuda.synchronize()
        env.barrier()
        t0 = time.perf_counter()
        reps = 4
        for _ in range(reps):
            res_f = pkg.segment_frames_batched(fctxs, raws, prefilter=pf)
        torch.cuda.synchronize()
        dtf = env.max_over_ranks(time.perf_counter() - t0)
        fa = {"frames_per_s": float(len(raws) * reps * env.world / dtf), "frames": len(raws) * reps * env.world,
              "points_per_message": int(raws[0].shape[0]),
              "note": "raw camera-frame PointCloud2 payload in (pinned host), VoxelGrid 0.01 + deep filter 3.0 + transform "
                      "on the device, then normals/supports/clusters/primitive fits (obj_segmentation.cpp:233-316 order)"}
        if env.rank == 0:
            world_cloud, _ = O.prefilter(raws_np[0], pf)
            want_f = O.segment_frame(world_cloud, O.default_frame_params())
            fa["parity_ok"] = bool(frames_equal(res_f[0], want_f))
            if not args.no_cpu_baseline:
                t0 = time.perf_counter()
                wc, _ = O.prefilter(raws_np[1 % len(raws_np)], pf)
                O.segment_frame(wc, O.default_frame_params())
                fa["cpu_frames_per_s_1core"] = 1.0 / (time.perf_counter() - t0)
        info["faithful"] = fa
    if env.rank == 0 and not args.no_cpu_baseline:
        info["cpu_baseline"] = cpu_frames_baseline(frames_np[: min(64, len(frames_np))], rounds=2, single=True)
    for cl in clouds:
        cl.release()
    stager.close()
    for c in fctxs:
        c.close()
    return info


def arm_c2(env, args, pkg, split):
    """BASELINE configs[1]: table-plane RANSAC, 1 M points x 5000 replayed hypotheses. split = True: the hypothesis set of an
    N x 5000 stream goes through pitt_sac_segment_split (weak scaling, NCCL all-gather of the counts)."""
    torch, dist = env.torch, env.dist
    from pitt_object_table_segmentation_b200 import _abi as A, scenes
    from pitt_object_table_segmentation_b200.api import torch_allgather_callback
    world = env.world if split else 1
    rank = env.rank if split else 0
    dev = env.dev
    stream = torch.cuda.current_stream()
    ctx = pkg.Context(env.local_rank, seed=12345, stream=stream.cuda_stream)
    xyz = scenes.plane_outlier_cloud(N_POINTS, seed=12345)
    n = xyz.shape[0]
    host_pinned = torch.from_numpy(xyz).pin_memory()
    cloud = ctx.stage_host_ptr(host_pinned.data_ptr(), 16, n)
    H_all = N_HYP * world
    samples_all = np.ascontiguousarray(ctx.pcl_sample_stream(cloud, A.MODEL_PLANE, H_all))
    samples = np.ascontiguousarray(samples_all[rank * N_HYP:(rank + 1) * N_HYP])
    p = pkg.default_support_sac_params()
    p.sampler, p.stop, p.max_iterations = A.SAMPLER_REPLAY, A.STOP_ALL_H, H_all
    p.replay_samples = samples_all.ctypes.data_as(A.i32p)
    p.replay_count = H_all
    p1 = pkg.default_support_sac_params()  # this rank's slice alone (kernel timing, single-GPU step)
    p1.sampler, p1.stop, p1.max_iterations = A.SAMPLER_REPLAY, A.STOP_ALL_H, N_HYP
    p1.replay_samples = samples.ctypes.data_as(A.i32p)
    p1.replay_count = N_HYP
    l2_flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2
    d_samples = torch.from_numpy(samples).to(dev)
    d_counts = torch.zeros(N_HYP, dtype=torch.int32, device=dev)
    gather = torch_allgather_callback() if world > 1 else None
    last = {}

    def step_resident():
        if world == 1:
            last["r"] = ctx.sac_segment_count_only(cloud, p1)
        else:
            last["r"] = ctx.sac_segment_split(cloud, p, rank, world, gather, want_inliers=False)

    inl_pinned = torch.empty(n, dtype=torch.int32).pin_memory()
    inl_host = inl_pinned.numpy()
    n_inl, n_co = C.c_int(0), C.c_int(0)
    co = np.zeros(8, np.float32)
    d2h = [0]

    def step_e2e():
        # the call a service callback makes: the request's cloud (pinned host memory) in, inliers + coefficients out
        if world == 1:
            st = ctx.lib.pitt_sac_segment_host(ctx.handle, C.c_void_p(host_pinned.data_ptr()), 16, n, C.byref(p1),
                                               inl_host.ctypes.data_as(A.i32p), n, C.byref(n_inl), co.ctypes.data_as(A.f32p),
                                               C.byref(n_co), None)
            assert st == 0, st
        else:
            cl = ctx.stage_host_ptr(host_pinned.data_ptr(), 16, n)
            st = ctx.lib.pitt_sac_segment_split(ctx.handle, cl.handle, C.byref(p), rank, world, gather, None,
                                                inl_host.ctypes.data_as(A.i32p), n, C.byref(n_inl), co.ctypes.data_as(A.f32p),
                                                C.byref(n_co), None)
            assert st == 0, st
            cl.release()
        d2h[0] = n_inl.value * 4 + 8 * 4 + 16 * 4

    flush = l2_flush.zero_
    l0 = ctx.kernel_launches
    steps = args.steps if args.workload == "c2" else max(5, min(args.steps, 20))
    ms_step, wall = env.timed_steps(step_resident, steps, max(3, args.warmup), before_each=flush)
    launches = (ctx.kernel_launches - l0) * steps // (steps + max(3, args.warmup))
    e2e_ms, _ = env.timed_steps(step_e2e, max(3, min(steps, 10)), 2, before_each=flush)
    evals_step = float(n) * N_HYP * world

    # roofline of the dominant kernel: the library brackets the plane_tc_kernel launch alone with two CUDA events
    def kernel_only():
        ctx.sac_score_device(cloud, p1, d_samples.data_ptr(), N_HYP, d_counts.data_ptr())

    k_ms, _ = env.timed_steps(kernel_only, max(5, min(steps, 20)), 3, before_each=flush)
    ctx.lib.pitt_debug_plane_tc_time_kernel(ctx.handle, 1)
    kk = []
    for _ in range(3 + max(5, min(steps, 20))):
        l2_flush.zero_()
        kernel_only()
        kk.append(float(ctx.lib.pitt_debug_plane_tc_kernel_ms(ctx.handle)))
    ctx.lib.pitt_debug_plane_tc_time_kernel(ctx.handle, 0)
    kk = [v for v in kk[3:] if v > 0]
    kern_ms = float(np.mean(kk)) if kk else k_ms
    kern_ms = env.max_over_ranks(kern_ms) if split else kern_ms
    peak_unfused = ctx.fp32_peak(1)
    peak_ffma = ctx.fp32_peak(0)
    achieved = float(n) * N_HYP * FLOP_PER_EVAL / (kern_ms * 1e-3) / 1e12
    traffic = None
    for name in ("r02_plane_tc_traffic.json", "r01_plane_tc_traffic.json"):
        tpath = os.path.join(ROOT, "profiles", name)
        if os.path.exists(tpath):
            try:
                traffic = json.load(open(tpath)).get("dram_bytes_per_launch")
                break
            except Exception:
                traffic = None
    evals_s = float(n) * N_HYP / (kern_ms * 1e-3)
    bf16_peak = MEASURED.get("bf16_tflops", 1659.2)
    floor_evals_s = 148 * 4 * 32 / 3.125 * 1.965e9
    alg_bytes = n * 16 + N_HYP * (64 + 4)
    roofline = {
        "bound": "fp32", "kernel": "plane_tc_kernel", "workload": "C2: 1M points x 5000 plane hypotheses (RANSAC scoring)",
        "achieved": achieved, "peak": peak_ffma, "unit": "TFLOP/s", "frac": achieved / peak_ffma, "traffic": traffic,
        "peak_source": "measured live: pitt_fp32_peak(kind=0) = FFMA issue rate with immediate operands (MEASURED_PEAKS.json has "
                       "no CUDA-core figure)",
        "frac_of_unfused_peak": achieved / peak_unfused, "peak_unfused_tflops": peak_unfused,
        "fma_pipe_ops_per_eval": 3, "issue_slots_per_eval": 2, "frac_of_epilogue_floor": evals_s / floor_evals_s,
        "tensor_tflops": evals_s * 64.0 / 1e12, "frac_tensor_of_measured_bf16": evals_s * 64.0 / 1e12 / bf16_peak,
        "note": "dot products on tcgen05 (2 chained 128x256x16 BF16 MMAs per tile on exact 3-piece splits, FP32 accumulators in "
                "TMEM), CUDA cores run the saturating-count epilogue; `achieved` counts the ALGORITHMIC 6 flop per evaluation",
        "kernel_ms": kern_ms, "kernel_ms_note": "plane_tc_kernel alone: CUDA events recorded by the library around that launch on "
                                                "the context's stream, mean over the timed calls, L2 flushed before each",
        "score_call_ms": k_ms, "algorithmic_flop_per_launch": float(n) * N_HYP * 6, "algorithmic_bytes_per_launch": alg_bytes,
        "as_hbm": {"bound": "hbm", "achieved": alg_bytes / (kern_ms * 1e-3) / 1e9, "peak": MEASURED.get("hbm_gbs", 6523.3),
                   "unit": "GB/s", "frac": alg_bytes / (kern_ms * 1e-3) / 1e9 / MEASURED.get("hbm_gbs", 6523.3)},
    }
    out = {"metric": "RANSAC hypothesis-point evals/s", "unit": "evals/s", "value": evals_step / (ms_step * 1e-3),
           "ms_per_step": ms_step, "launches_per_step": int(launches),
           "e2e": {"value": evals_step / (e2e_ms * 1e-3), "unit": "evals/s", "ms_per_step": e2e_ms,
                   "h2d_bytes_per_step": n * 16 + H_all * 12, "d2h_bytes_per_step": d2h[0]},
           "roofline": roofline, "n_inliers": int(last["r"]["n_inliers"]), "winner": int(last["r"]["info"].best_hypothesis),
           "config": {"workload": WORKLOAD_C2,
                      "points": n, "hypotheses_per_gpu": N_HYP, "l2": "flushed between timed iterations (256 MB write)",
                      "parallelism": "hypothesis split through pitt_sac_segment_split + NCCL all-gather of counts" if world > 1
                                     else "single GPU"}}
    # oracle spot check on this rank: a few of its hypotheses' counts + the joint winner's refined model
    from oracle import orc_binding as O
    po = O.default_support_sac_params()
    counts = d_counts.cpu().numpy()
    pick = np.array([0, N_HYP // 3, N_HYP - 1])
    c_cpu = O.sac_score(xyz, None, po, samples[pick])[0]
    ok = bool(np.array_equal(counts[pick], c_cpu))
    out["parity_checked"] = {"ok": bool(env.min_over_ranks(1.0 if ok else 0.0) == 1.0) if split else ok,
                             "what": "3 hypotheses' inlier counts per rank against the oracle's countWithinDistance"}
    if env.rank == 0 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_c2_baseline(xyz, samples, seconds_target=10.0)
    cloud.release()
    ctx.close()
    return out


def arm_primitives(env, args, pkg, peak_ffma):
    """BASELINE configs[2] (C3) in one line per model: a 50 000-point cluster x 10 000 hypotheses of the PCL sample stream,
    estimate + score on the device, 20 calls back to back between two CUDA events (per-GPU figure)"""
    torch = env.torch
    from pitt_object_table_segmentation_b200 import _abi as A, scenes
    ctx = pkg.Context(env.local_rank, seed=12345, stream=torch.cuda.current_stream().cuda_stream)
    if peak_ffma is None:
        peak_ffma = ctx.fp32_peak(0)
    primitives = {"points": 50000, "hypotheses": 10000, "peak_ffma_tflops": peak_ffma,
                  "note": "per GPU; algorithmic flop per evaluation from SURVEY 8d (sphere 10, cylinder 69, cone 91) against the live FFMA peak"}
    for kind, model, flop in (("sphere", A.MODEL_SPHERE, 10), ("cylinder", A.MODEL_CYLINDER, 69), ("cone", A.MODEL_CONE, 91)):
        pxyz, _ = scenes.primitive_cluster(kind, 50000, 5)
        pcloud = ctx.stage(pxyz)
        ctx.estimate_normals(pcloud, 50)
        pp = pkg.default_sac_params(model)
        ps = torch.from_numpy(ctx.pcl_sample_stream(pcloud, model, 10000)).to(env.dev)
        pc = torch.zeros(10000, dtype=torch.int32, device=env.dev)
        for _ in range(5):
            ctx.sac_score_device(pcloud, pp, ps.data_ptr(), 10000, pc.data_ptr())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            ctx.sac_score_device(pcloud, pp, ps.data_ptr(), 10000, pc.data_ptr())
        e1.record()
        torch.cuda.synchronize()
        pms = e0.elapsed_time(e1) / 20
        ev = 50000.0 * 10000.0 / (pms * 1e-3)
        primitives[kind] = {"ms": pms, "evals_per_s": ev, "algorithmic_tflops": ev * flop / 1e12,
                            "frac_of_ffma_peak": ev * flop / 1e12 / peak_ffma, "best_count": int(pc.max().item())}
        pcloud.release()
    ctx.close()
    return primitives


def arm_c3(env, args, pkg):
    """The whole C3 job as the reference would run it: 64 clusters of 5 000 .. 50 000 points (32 cylinders, 32 cones), per cluster
    k = 50 normals + SACSegmentationFromNormals with setMaxIterations(10000) and PCL's adaptive stop + LM refinement; clusters
    resident on the device, results on the host. Clusters are independent units: size-balanced over the ranks and, on
    each GPU, over 8 contexts."""
    from concurrent.futures import ThreadPoolExecutor
    from pitt_object_table_segmentation_b200 import _abi as A, scenes, sharding
    torch = env.torch
    sizes = np.linspace(5000, 50000, 64).astype(int)
    owner = sharding.greedy_balance([int(v) for v in sizes], env.world)
    mine = [i for i in range(64) if owner[i] == env.rank]
    n_c3 = 8
    cctx = [pkg.Context(env.local_rank, seed=12345) for _ in range(n_c3)]
    slot = sharding.greedy_balance([int(sizes[i]) for i in mine], n_c3)
    jobs = [[] for _ in range(n_c3)]
    host = {}
    for j, i in enumerate(mine):
        kind, model = (("cylinder", A.MODEL_CYLINDER), ("cone", A.MODEL_CONE))[i % 2]
        cxyz, _ = scenes.primitive_cluster(kind, int(sizes[i]), 100 + i)
        pj = pkg.default_sac_params(model)
        pj.max_iterations = 10000
        jobs[slot[j]].append((i, cctx[slot[j]].stage(cxyz), pj))
        host[i] = cxyz
    results = {}

    def c3_worker(k):
        tot = 0
        for i, cl, pj in jobs[k]:
            cctx[k].estimate_normals_device(cl, 50)
            r = cctx[k].sac_segment_count_only(cl, pj)
            results[i] = r
            tot += r["n_inliers"]
        return tot

    with ThreadPoolExecutor(n_c3) as pool:
        sum(pool.map(c3_worker, range(n_c3)))
        torch.cuda.synchronize()
        times = []
        for _ in range(3):
            env.barrier()
            t0 = time.perf_counter()
            inl_mine = sum(pool.map(c3_worker, range(n_c3)))
            torch.cuda.synchronize()
            times.append(env.max_over_ranks(time.perf_counter() - t0))
    c3_s = float(np.median(times))
    inl_total = int(env.sum_over_ranks(inl_mine))
    # oracle spot check: the smallest cluster of this rank, the whole segment() (coefficients bit for bit, inlier count)
    ok = True
    if mine:
        from oracle import orc_binding as O
        i = min(mine, key=lambda q: sizes[q])
        model = (A.MODEL_CYLINDER, A.MODEL_CONE)[i % 2]
        nrm = O.estimate_normals(host[i], 50)
        po = O.default_sac_params(model)
        po.max_iterations = 10000
        w = O.sac_segment(host[i], nrm, po)
        g = results[i]
        ok = bool(len(w["inliers"]) == g["n_inliers"] and
                  np.array_equal(np.asarray(w["coeffs"], np.float32).view(np.uint32), np.asarray(g["coeffs"], np.float32).view(np.uint32)))
    all_ok = env.min_over_ranks(1.0 if ok else 0.0) == 1.0
    for k in range(n_c3):
        for _, cl, _ in jobs[k]:
            cl.release()
        cctx[k].close()
    return {"clusters": 64, "points_total": int(sizes.sum()), "max_iterations": 10000, "seconds": c3_s, "clusters_per_s": 64 / c3_s,
            "inliers_total": inl_total, "contexts_per_gpu": n_c3, "parity_checked": {"ok": bool(all_ok), "what": "smallest cluster of "
            "every rank: normals + segment() against the oracle (coefficients bit for bit, inlier count)"},
            "note": "wall clock, median of 3, max over ranks; clusters size-balanced over ranks and contexts, normals + segment() each"}


def arm_c5(env, args, pkg):
    """BASELINE configs[4] as written: one giant scene resident on every GPU, the plane hypotheses split over the ranks through
    pitt_sac_segment_split (NCCL all-gather of the counts, earliest arg-max, refine + select everywhere). Strong scaling."""
    torch = env.torch
    from pitt_object_table_segmentation_b200 import _abi as A, scenes
    from pitt_object_table_segmentation_b200.api import torch_allgather_callback
    n, H = args.c5_points, args.c5_hypotheses
    ctx = pkg.Context(env.local_rank, seed=12345, stream=torch.cuda.current_stream().cuda_stream)
    t0 = time.perf_counter()
    xyz = scenes.plane_outlier_cloud(n, seed=5)  # every rank generates the same scene (no broadcast needed)
    gen_s = time.perf_counter() - t0
    cloud = ctx.stage(xyz)
    rng = np.random.default_rng(11)
    s0 = rng.integers(0, n, H)
    s1 = (s0 + 1 + rng.integers(0, n - 2, H)) % n
    s2 = (s1 + 1 + rng.integers(0, n - 3, H)) % n
    samples_all = np.ascontiguousarray(np.stack([s0, s1, s2], 1).astype(np.int32))
    p = pkg.default_support_sac_params()
    p.sampler, p.stop, p.max_iterations = A.SAMPLER_REPLAY, A.STOP_ALL_H, H
    p.replay_samples = samples_all.ctypes.data_as(A.i32p)
    p.replay_count = H
    gather = torch_allgather_callback() if env.world > 1 else None
    last = {}

    def step():
        last["r"] = ctx.sac_segment_split(cloud, p, env.rank, env.world, gather, want_inliers=False)

    steps = max(2, min(args.steps, 5))
    ms_step, _ = env.timed_steps(step, steps, 1)
    r = last["r"]
    # oracle spot check: this rank's first hypotheses + the joint winner's count
    from oracle import orc_binding as O
    po = O.default_support_sac_params()
    H_loc = (H + env.world - 1) // env.world
    mine = samples_all[env.rank * H_loc: env.rank * H_loc + 2]
    win = int(r["info"].best_hypothesis)
    chk = np.concatenate([mine, samples_all[win:win + 1]]) if win >= 0 else mine
    ctx.lib.pitt_debug_plane_mode(0)
    c_gpu = ctx.sac_score(cloud, p, chk)[0]
    c_cpu = O.sac_score(xyz, None, po, chk)[0]
    ok = bool(np.array_equal(c_gpu, c_cpu) and (win < 0 or int(c_cpu[-1]) == int(r["info"].best_count)))
    same_winner = env.min_over_ranks(win) == env.max_over_ranks(win)
    all_ok = env.min_over_ranks(1.0 if ok else 0.0) == 1.0 and same_winner
    cloud.release()
    ctx.close()
    return {"metric": "RANSAC hypothesis-point evals/s", "unit": "evals/s", "value": float(n) * H / (ms_step * 1e-3), "ms_per_step": ms_step,
            "points": n, "hypotheses": H, "winner": win, "winner_count": int(r["info"].best_count), "final_inliers": int(r["n_inliers"]),
            "scene_generation_s": gen_s,
            "parity_checked": {"ok": bool(all_ok), "what": "per rank: 2 hypotheses of its slice + the joint winner against the oracle's "
                               "countWithinDistance on the full scene; all ranks name the same winner"}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c1", choices=["c1", "c2", "c3", "c5"],
                    help="c1 (default): the headline, full-resolution frames/s with the C2 roofline and the C3 figures as extra keys; "
                         "c2 / c3 / c5: that BASELINE configuration alone as the printed metric")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-primitives", action="store_true", help="skip the C3-shaped sphere/cylinder/cone scoring figures and the C3 job")
    ap.add_argument("--no-ransac", action="store_true", help="skip the C2 section (evals/s + roofline of plane_tc_kernel)")
    ap.add_argument("--no-faithful", dest="faithful", action="store_false", help="skip the raw-message (VoxelGrid first) frame variant")
    ap.add_argument("--frames-per-step", type=int, default=128, help="frames per GPU in one step (C4: 1024 frames over 8 GPUs)")
    ap.add_argument("--frame-contexts", type=int, default=0,
                    help="host threads / CUDA streams per GPU for the frame stream (0 = 16: measured 979 / 1082 / 920 frames/s with 8 / 16 / 32)")
    ap.add_argument("--frame-workers", type=int, default=0,
                    help="helper streams per context for the primitive fits of a frame (latency knob; 0 is best for throughput)")
    ap.add_argument("--c5-points", type=int, default=50_000_000)
    ap.add_argument("--c5-hypotheses", type=int, default=100_000)
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    args.warmup = max(3, args.warmup)

    import pitt_object_table_segmentation_b200 as pkg
    env = Env()
    line = None
    common = {"n_gpus": env.world, "steps": args.steps, "warmup": args.warmup, "higher_is_better": True, "vs_baseline": None,
              "dtype": "f32", "data": "synthetic"}
    if args.workload == "c2":
        r = arm_c2(env, args, pkg, split=True)
        line = dict(common, metric=r["metric"], value=r["value"], unit=r["unit"], ms_per_step=r["ms_per_step"], scaling="weak",
                    config=r["config"], e2e=r["e2e"], gpu_launches=r["launches_per_step"] * args.steps, roofline=r["roofline"],
                    parity_checked=r["parity_checked"], cpu_baseline=r.get("cpu_baseline"))
    elif args.workload == "c3":
        r = arm_c3(env, args, pkg)
        line = dict(common, metric="C3 clusters/s (normals + cylinder/cone RANSAC, 10k iterations)", value=r["clusters_per_s"],
                    unit="clusters/s", ms_per_step=r["seconds"] * 1e3, scaling="strong",
                    config={"workload": "C3: 64 clusters of 5k-50k points, normals + SACSegmentationFromNormals (10 000 iterations, "
                                        "adaptive stop, LM refinement), clusters size-balanced over the ranks"},
                    c3_job=r, parity_checked=r["parity_checked"])
    elif args.workload == "c5":
        r = arm_c5(env, args, pkg)
        line = dict(common, metric=r["metric"], value=r["value"], unit=r["unit"], ms_per_step=r["ms_per_step"], scaling="strong",
                    config={"workload": f"C5: {r['points']}-point scene resident on every GPU, {r['hypotheses']} plane hypotheses split "
                                        "over the ranks (pitt_sac_segment_split, NCCL all-gather of counts, arg-max, refine + select)"},
                    c5=r, parity_checked=r["parity_checked"])
    else:
        fr = arm_frames(env, args, pkg)
        ransac = None
        if not args.no_ransac:
            if env.rank == 0:
                class _Solo:  # the C2 section runs on rank 0's GPU alone (per-GPU kernel figure)
                    pass
                solo = Env.__new__(Env)
                solo.__dict__.update(env.__dict__)
                solo.world = 1
                solo.barrier = env.torch.cuda.synchronize
                solo.max_over_ranks = lambda v: float(v)
                solo.min_over_ranks = lambda v: float(v)
                solo.sum_over_ranks = lambda v: float(v)
                ransac = arm_c2(solo, args, pkg, split=False)
            env.barrier()
        primitives = None
        if not args.no_primitives:
            if env.rank == 0:
                primitives = arm_primitives(env, args, pkg, ransac["roofline"]["peak"] if ransac else None)
            env.barrier()
            c3 = arm_c3(env, args, pkg)
            if primitives is not None:
                primitives["c3_job"] = c3
        if env.rank == 0:
            line = dict(common, metric=METRIC, value=fr["value"], unit=UNIT, ms_per_step=fr["ms_per_step"], scaling="weak",
                        config={"workload": WORKLOAD_C1,
                                "frames_per_gpu_per_step": fr["frames_per_gpu_per_step"], "points_per_frame": fr["points_per_frame"],
                                "contexts_per_gpu": fr["contexts_per_gpu"],
                                "l2": "inputs larger than L2: one step reads %d MB of clouds per GPU" % (fr["e2e"]["h2d_bytes_per_step"] >> 20),
                                "parallelism": "frame-sharded, one process per GPU, no data-path collective" if env.world > 1 else "single GPU"},
                        clocks=fr["clocks"], e2e=fr["e2e"], gpu_launches=fr["launches"], parity_checked=fr["parity"],
                        frames={k: v for k, v in fr.items() if k not in ("clocks", "e2e", "parity", "cpu_baseline")},
                        wall_s_timed_region=fr["wall_s"])
            # the frame path's own dominant kernel group (whole-cloud normals) beside the RANSAC scoring roofline
            if "normals" in fr:
                line["roofline_frame"] = dict(fr["normals"]["roofline"], kernel_ms=fr["normals"]["ms"])
            if ransac:
                line["roofline"] = ransac["roofline"]
                line["ransac"] = {k: v for k, v in ransac.items() if k != "roofline"}
            if primitives:
                line["primitives"] = primitives
            if "cpu_baseline" in fr:
                line["cpu_baseline"] = fr["cpu_baseline"]
    if env.rank == 0 and line is not None:
        print(json.dumps(line))
    if env.world > 1:
        env.dist.barrier()
        env.dist.destroy_process_group()


if __name__ == "__main__":
    main()
