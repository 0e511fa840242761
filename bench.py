#!/usr/bin/env python
"""bench.py — headline benchmark of the B200-native segmentation hot path.

Workload (BASELINE.json configs[1], "C2"): table-plane RANSAC on a 1 000 000-point synthetic cloud
(70 % plane z=0 with sigma 2 mm, 30 % uniform outliers, seed 12345), 5000 hypotheses replayed from
PCL's mt19937(12345) sample stream, threshold 0.02, stop rule ALL_H. A step = one seg.segment():
model estimation, scoring of all H x N hypothesis-point pairs, earliest arg-max, least-squares
refinement of the winner and final inlier selection.

metric  = RANSAC hypothesis.point evaluations per second (whole job, all GPUs).
value   = cloud resident in HBM, the step runs through pitt_sac_segment (count-only output).
e2e     = same call with HOST buffers: pitt_stage_cloud from pinned memory (H2D of the cloud inside
          the timed region) + pitt_sac_segment + the inlier list copied back to the host.
N > 1   = weak scaling of the hypothesis split (config 5 pattern): every rank holds the cloud and
          scores its own 5000 hypotheses of a N x 5000 stream, the counts are all-gathered with NCCL,
          every rank takes the earliest arg-max, refines the winner and selects its final inliers
          (pitt_sac_finish_device), i.e. the same work per rank as the N = 1 step.

`--impl reference` times the CPU path (the oracle = the PCL restatement; the reference itself
cannot be built here) on the host cores on a bounded sample of the same workload.
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

N_POINTS = 1_000_000
N_HYP = 5000
FLOP_PER_EVAL = 6  # 3 mul + 3 add, unfused
METRIC = "RANSAC hypothesis-point evals/s"
UNIT = "evals/s"


def make_workload(n_ranks):
    from pitt_object_table_segmentation_b200 import scenes
    xyz = scenes.plane_outlier_cloud(N_POINTS, seed=12345)
    return xyz


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
                                          "--format=csv,noheader,nounits", "-lms", "100"],
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


def cpu_baseline(xyz, samples, seconds_target=12.0, threads=None):
    """The oracle (PCL restatement) scoring a bounded hypothesis sample of the same cloud on the host."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import orc_binding as O
    threads = threads or (os.cpu_count() or 1)
    p = O.default_support_sac_params()
    # calibrate on 2 hypotheses per thread
    t0 = time.perf_counter()
    O.sac_score(xyz, None, p, samples[:2])
    per_h = (time.perf_counter() - t0) / 2
    h_total = int(max(threads, min(len(samples), seconds_target / per_h * threads)))
    h_total -= h_total % threads
    chunks = np.array_split(samples[:h_total], threads)
    t0 = time.perf_counter()
    with ThreadPoolExecutor(threads) as ex:  # ctypes releases the GIL inside the oracle
        res = list(ex.map(lambda s: O.sac_score(xyz, None, p, s)[0], chunks))
    dt = time.perf_counter() - t0
    evals = float(h_total) * xyz.shape[0]
    return {"value": evals / dt, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"{h_total} of {N_HYP} hypotheses x {xyz.shape[0]} points, oracle countWithinDistance, "
                      f"{dt:.1f} s wall", "checksum": int(sum(int(r.sum()) for r in res))}


def cpu_frames_baseline(raws, pf, frames_per_thread=2):
    """The oracle's pre-path + frame path (the reference's algorithm) on the host: one frame on one core (what
    the reference's blocking service chain uses) and frame-parallel on all cores."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import orc_binding as O
    fp = O.default_frame_params()

    def one(raw):
        world_cloud, _ = O.prefilter(raw, pf)
        return O.segment_frame(world_cloud, fp)["n_clusters"]

    one(raws[0])
    t0 = time.perf_counter()
    for r in raws[:2]:
        one(r)
    single = 2 / (time.perf_counter() - t0)
    threads = os.cpu_count() or 1
    work = [raws[i % len(raws)] for i in range(threads * frames_per_thread)]
    t0 = time.perf_counter()
    with ThreadPoolExecutor(threads) as ex:  # ctypes releases the GIL inside the oracle
        list(ex.map(one, work))
    allc = len(work) / (time.perf_counter() - t0)
    return {"frames_per_s_1core": single, "frames_per_s_all_cores": allc, "cores": threads, "kind": "port",
            "sample": f"2 frames on 1 core; {len(work)} frames on {threads} threads (oracle = PCL restatement, -O2)"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "liborc.so"], stdout=subprocess.DEVNULL)
    from oracle import orc_binding as O
    from pitt_object_table_segmentation_b200 import _abi as A
    xyz = make_workload(1)
    samples = O.pcl_sample_stream(xyz, A.MODEL_PLANE, N_HYP)
    threads = os.cpu_count() or 1
    per_step_target = max(2.0, min(20.0, 120.0 / max(1, args.steps + args.warmup)))
    vals = []
    last = None
    for i in range(args.warmup + args.steps):
        r = cpu_baseline(xyz, samples, seconds_target=per_step_target, threads=threads)
        if i >= args.warmup:
            vals.append(r["value"])
        last = r
    value = float(np.mean(vals))
    evals_per_step = N_POINTS * N_HYP
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": evals_per_step / value * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "C2: table-plane RANSAC, 1M-point cloud x 5000 hypotheses (bounded sample per step, "
                               "ms_per_step extrapolated to the full 5e9 evaluations)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": last["sample"]},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-primitives", action="store_true", help="skip the C3-shaped sphere/cylinder/cone scoring figures")
    ap.add_argument("--frames", type=int, default=256, help="frames per GPU for the secondary frames/s metric (0 = skip)")
    ap.add_argument("--frame-contexts", type=int, default=0,
                    help="host threads / CUDA streams per GPU for the frame stream (0 = 32 on one GPU, 16 per GPU otherwise: "
                         "measured 592 / 654 / 683 / 682 frames/s with 16 / 24 / 32 / 48 contexts on one B200)")
    ap.add_argument("--frame-workers", type=int, default=0,
                    help="helper streams per context for the primitive fits of a frame (latency knob; 0 is best for throughput)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    args.warmup = max(3, args.warmup)

    import torch
    import torch.distributed as dist
    import pitt_object_table_segmentation_b200 as pkg
    from pitt_object_table_segmentation_b200 import _abi as A, sharding

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    stream = torch.cuda.current_stream()
    ctx = pkg.Context(local_rank, seed=12345, stream=stream.cuda_stream)

    xyz = make_workload(world)
    n = xyz.shape[0]
    host_pinned = torch.from_numpy(xyz).pin_memory()
    cloud = ctx.stage_host_ptr(host_pinned.data_ptr(), 16, n)
    samples_all = ctx.pcl_sample_stream(cloud, A.MODEL_PLANE, N_HYP * world)
    samples = np.ascontiguousarray(samples_all[rank * N_HYP:(rank + 1) * N_HYP])

    p = pkg.default_support_sac_params()
    p.sampler, p.stop, p.max_iterations = A.SAMPLER_REPLAY, A.STOP_ALL_H, N_HYP
    p.replay_samples = samples.ctypes.data_as(A.i32p)
    p.replay_count = N_HYP

    l2_flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2
    d_samples = torch.from_numpy(samples).to(dev)
    d_samples_all = torch.from_numpy(np.ascontiguousarray(samples_all)).to(dev)
    d_counts = torch.zeros(N_HYP, dtype=torch.int32, device=dev)
    d_all = torch.zeros(N_HYP * world, dtype=torch.int32, device=dev)
    d_best = torch.zeros(2, dtype=torch.int32, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_resident():
        if world == 1:
            return ctx.sac_segment_count_only(cloud, p)
        # hypothesis split: local scoring, NCCL all-gather of the counts, earliest arg-max everywhere
        ctx.sac_score_device(cloud, p, d_samples.data_ptr(), N_HYP, d_counts.data_ptr())
        dist.all_gather_into_tensor(d_all, d_counts)
        ctx.argmax_counts_device(d_all.data_ptr(), N_HYP * world, d_best.data_ptr())
        # every rank holds the cloud: refine the winner and select its final inliers (count only), like N = 1
        return ctx.sac_finish_device(cloud, p, d_samples_all.data_ptr(), N_HYP * world, d_best.data_ptr())

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        t_wall = time.perf_counter()
        l0 = ctx.kernel_launches
        for i in range(steps):
            l2_flush.zero_()  # flush L2 between timed iterations (outside the event pair)
            ev[i][0].record()
            fn()
            ev[i][1].record()
        barrier()
        wall = time.perf_counter() - t_wall
        ms = sum(a.elapsed_time(b) for a, b in ev)
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        timed.launches = ctx.kernel_launches - l0
        return float(t.item()) / steps, wall

    clocks = ClockSampler(local_rank)
    clocks.start()
    clocks.wait_first()
    ms_step, wall = timed(step_resident, args.steps, args.warmup)
    # the timed region lasts tens of milliseconds and nvidia-smi samples every 100 ms: keep the very same loop running
    # (untimed) until a few samples have been taken under the same load
    t_ext = time.perf_counter()
    while len(clocks.rows) < 4 and time.perf_counter() - t_ext < 1.5:
        for _ in range(20):
            step_resident()
        torch.cuda.synchronize()
    clk = clocks.stop()
    clk["note"] = "sampled every 100 ms over the timed region and an untimed continuation of the same step loop"
    launches = timed.launches
    evals_step = float(n) * N_HYP * world
    value = evals_step / (ms_step * 1e-3)

    # ---- end-to-end arm: host buffers in, host buffers out, through the C ABI
    inl_pinned = torch.empty(n, dtype=torch.int32).pin_memory()  # the response's index vector (pinned, like the request's cloud)
    inl_host = inl_pinned.numpy()
    n_inl, n_co = C.c_int(0), C.c_int(0)
    co = np.zeros(8, np.float32)
    handle = C.c_void_p()
    d2h_bytes = [0]

    def step_e2e():
        # the call a service callback makes: the cloud of the request (pinned host memory) in, inliers + coefficients out;
        # the library overlaps the host -> device copy with the scoring (pitt_sac_segment_host)
        st = ctx.lib.pitt_sac_segment_host(ctx.handle, C.c_void_p(host_pinned.data_ptr()), 16, n, C.byref(p),
                                           inl_host.ctypes.data_as(A.i32p), n, C.byref(n_inl), co.ctypes.data_as(A.f32p),
                                           C.byref(n_co), None)
        assert st == 0, st
        d2h_bytes[0] = n_inl.value * 4 + 8 * 4 + 16 * 4
        if world > 1:
            dist.all_gather_into_tensor(d_all, d_counts)

    e2e_steps = max(3, min(args.steps, 10))
    e2e_ms, _ = timed(step_e2e, e2e_steps, 2)
    e2e_value = evals_step / (e2e_ms * 1e-3)

    # ---- roofline of the dominant kernel (plane_score_kernel): estimate + score bracketed by events
    def kernel_only():
        ctx.sac_score_device(cloud, p, d_samples.data_ptr(), N_HYP, d_counts.data_ptr())

    k_ms, _ = timed(kernel_only, max(5, min(args.steps, 20)), 3)
    # the dominant kernel alone: the library brackets the plane_tc_kernel launch with two CUDA events on its stream
    ctx.lib.pitt_debug_plane_tc_kernel_ms.restype = C.c_double
    ctx.lib.pitt_debug_plane_tc_kernel_ms.argtypes = [C.c_void_p]
    ctx.lib.pitt_debug_plane_tc_time_kernel(1)
    kk = []
    for _ in range(3 + max(5, min(args.steps, 20))):
        l2_flush.zero_()
        kernel_only()
        kk.append(float(ctx.lib.pitt_debug_plane_tc_kernel_ms(ctx.handle)))
    ctx.lib.pitt_debug_plane_tc_time_kernel(0)
    kk = [v for v in kk[3:] if v > 0]
    kern_ms = float(np.mean(kk)) if kk else k_ms
    if world > 1:
        tk = torch.tensor([kern_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(tk, op=dist.ReduceOp.MAX)
        kern_ms = float(tk.item())
    peak_unfused = ctx.fp32_peak(1)
    peak_ffma = ctx.fp32_peak(0)
    achieved = float(n) * N_HYP * FLOP_PER_EVAL / (kern_ms * 1e-3) / 1e12
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "r01_plane_tc_traffic.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    # The dominant kernel is plane_tc_kernel (csrc/plane_tc.cu): the dot products run on the tensor cores (tcgen05.mma
    # kind::f16 on exact 3-piece BF16 splits, two chained 128x256x16 MMAs = 32 BF16 MAC slots per evaluation), the CUDA
    # cores execute 3 FMA-pipe operations per evaluation (saturating subtract, sum u, sum u^2; 2 issue slots because the two
    # sums are packed FADD2 / FFMA2) and that pipe is what bounds the kernel. The schema's roofline is
    # therefore quoted against the FP32 FFMA peak measured live (MEASURED_PEAKS.json has no CUDA-core figure): `achieved`
    # counts the ALGORITHMIC 6 flop per evaluation of SURVEY 8d; the fraction of the kernel's own instruction-mix floor
    # (tools/ldtm_probe.cu: 3.125 cycles per evaluation per sub-partition once the TMEM loads are included) and the
    # tensor-pipe fraction are reported beside it.
    evals_s = float(n) * N_HYP / (kern_ms * 1e-3)
    fma_ops_peak = peak_ffma * 1e12 / 2.0           # FMA-pipe operations per second (one FFMA = 2 flop)
    bf16_peak = MEASURED.get("bf16_tflops", 1659.2)
    sm_clock_hz = 1.965e9
    floor_evals_s = 148 * 4 * 32 / 3.125 * sm_clock_hz  # measured floor of FADD.SAT + FADD2/2 + FFMA2/2 + tcgen05.ld per evaluation
    roofline = {
        "bound": "fp32", "kernel": "plane_tc_kernel", "achieved": achieved, "peak": peak_ffma,
        "unit": "TFLOP/s", "frac": achieved / peak_ffma, "traffic": traffic,
        "peak_source": "measured live: pitt_fp32_peak(kind=0) = FFMA issue rate with immediate operands",
        "frac_of_unfused_peak": achieved / peak_unfused, "peak_unfused_tflops": peak_unfused,
        "fma_pipe_ops_per_eval": 3, "frac_fma_pipe_ops": evals_s * 3.0 / fma_ops_peak,
        "issue_slots_per_eval": 2, "frac_of_epilogue_floor": evals_s / floor_evals_s,
        "epilogue_floor_note": "tools/ldtm_probe.cu on B200: 2.0 cycles/eval/sub-partition for the arithmetic alone, 3.125 with "
                               "the TMEM->register loads (they do not overlap with arithmetic on the same sub-partition)",
        "tensor_bf16_mac_slots_per_eval": 32, "tensor_tflops": evals_s * 64.0 / 1e12,
        "frac_tensor_of_measured_bf16": evals_s * 64.0 / 1e12 / bf16_peak,
        "note": "dot products on tcgen05 (2 chained 128x256x16 BF16 MMAs per tile on exact 3-piece splits, FP32 accumulators "
                "in TMEM), CUDA cores run the saturating-count epilogue; FFMA filter kernel (previous dominant kernel) = 1.09 ms on "
                "the same job",
        "kernel_ms": kern_ms, "kernel_ms_note": "plane_tc_kernel alone: CUDA events recorded by the library around that launch, "
                                                "mean over the timed calls, L2 flushed before each",
        "score_call_ms": k_ms, "score_call_note": "estimate + set-up kernels + plane_tc_kernel (pitt_sac_score_device)",
        "algorithmic_flop_per_launch": float(n) * N_HYP * 6,
        "algorithmic_bytes_per_launch": n * 16 + N_HYP * (64 + 4),
        # the same launch against the two ceilings of the schema (MEASURED_PEAKS.json), for completeness: it is neither
        "as_hbm": {"bound": "hbm", "achieved": (n * 16 + N_HYP * (64 + 4)) / (kern_ms * 1e-3) / 1e9,
                   "peak": MEASURED.get("hbm_gbs", 6523.3), "unit": "GB/s",
                   "frac": (n * 16 + N_HYP * (64 + 4)) / (kern_ms * 1e-3) / 1e9 / MEASURED.get("hbm_gbs", 6523.3),
                   "note": "every byte of the cloud is read once per launch (traffic == algorithmic bytes); 16 MB against 3e10 flop"},
        "as_tensor": {"bound": "tensor", "achieved": evals_s * 64.0 / 1e12, "peak": bf16_peak, "unit": "TFLOP/s",
                      "frac": evals_s * 64.0 / 1e12 / bf16_peak,
                      "note": "executed BF16 MMA flop (32 MAC slots per evaluation), not algorithmic flop; ncu: tensor pipe 26 % active"},
    }

    # ---- secondary metric of BASELINE.json: segmented frames/s on 307k-point Kinect-shaped frames (C1/C4)
    frames_info = None
    if args.frames > 0:
        from pitt_object_table_segmentation_b200 import scenes
        n_ctx = args.frame_contexts if args.frame_contexts > 0 else (32 if world == 1 else 16)
        fctxs = [pkg.Context(local_rank, seed=12345) for _ in range(n_ctx)]
        # more host threads than cores on the box (8 ranks x 16 contexts): waits sleep instead of spinning
        oversubscribed = world * n_ctx > (os.cpu_count() or 1)
        for c in fctxs:
            c.set_workers(args.frame_workers)
            c.set_blocking_sync(oversubscribed)
        lo, hi = sharding.block_range(rank, world, args.frames * world)  # weak scaling: args.frames per GPU
        uniq = [torch.from_numpy(scenes.tabletop_frame(seed=lo + i, random_poses=True)).pin_memory() for i in range(min(4, hi - lo))]
        frames = [uniq[i % len(uniq)].numpy() for i in range(hi - lo)]
        pkg.segment_frames_batched(fctxs, frames[: 2 * n_ctx])  # warm-up (arenas, pools, clocks)
        barrier()
        t0 = time.perf_counter()
        res = pkg.segment_frames_batched(fctxs, frames)
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        frames_info = {"frames_per_s": float(len(frames) * world / dt.item()), "frames": len(frames) * world,
                       "points_per_frame": int(frames[0].shape[0]), "contexts_per_gpu": n_ctx,
                       "workers_per_context": args.frame_workers, "blocking_sync": bool(oversubscribed),
                       "host_cores": os.cpu_count(),
                       "shapes_first_frame": [s["tag_name"] for s in res[0]["shapes"]],
                       "note": "full-res 640x480 frame: normals k=50, supports loop, clustering, 4 primitive fits per "
                               "cluster, selection; host buffers in (pinned), results out; wall clock, max over ranks"}
        for c in fctxs:
            c.close()
        # faithful variant (BASELINE configs[0]): the raw 307 200-point camera-frame message (NaN returns, far
        # background) through fromROSMsg + 1 cm VoxelGrid + deep filter + transform on the device, then the frame path
        pf = pkg.default_prefilter_params()
        c2w, _ = scenes.camera_pose()
        for i, v in enumerate(c2w.ravel()):
            pf.transform[i] = float(v)
        raw_uniq = [torch.from_numpy(scenes.raw_camera_frame(seed=lo + i, random_poses=True)).pin_memory() for i in range(min(4, hi - lo))]
        raws = [raw_uniq[i % len(raw_uniq)].numpy() for i in range(hi - lo)]
        fctxs = [pkg.Context(local_rank, seed=12345) for _ in range(n_ctx)]
        # more host threads than cores on the box (8 ranks x 16 contexts): waits sleep instead of spinning
        oversubscribed = world * n_ctx > (os.cpu_count() or 1)
        for c in fctxs:
            c.set_workers(args.frame_workers)
            c.set_blocking_sync(oversubscribed)
        pkg.segment_frames_batched(fctxs, raws[: 2 * n_ctx], prefilter=pf)
        barrier()
        t0 = time.perf_counter()
        res_f = pkg.segment_frames_batched(fctxs, raws, prefilter=pf)
        torch.cuda.synchronize()
        dtf = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dtf, op=dist.ReduceOp.MAX)
        for c in fctxs:
            c.close()
        frames_info["faithful"] = {
            "frames_per_s": float(len(raws) * world / dtf.item()), "frames": len(raws) * world,
            "points_per_message": int(raws[0].shape[0]), "shapes_first_frame": [s["tag_name"] for s in res_f[0]["shapes"]],
            "note": "raw camera-frame PointCloud2 payload in (pinned host), VoxelGrid 0.01 + deep filter 3.0 + transform "
                    "on the device, then normals/supports/clusters/primitive fits (obj_segmentation.cpp:233-316 order)"}
        if rank == 0 and not args.no_cpu_baseline:
            frames_info["faithful"]["cpu_baseline"] = cpu_frames_baseline(raws[:4], pf)
        # single-frame latency: one context, the fits of a frame fanned out to 4 helper streams
        lctx = pkg.Context(local_rank, seed=12345)
        lctx.set_workers(4)
        lat = []
        for i in range(8):
            t1 = time.perf_counter()
            cl = lctx.stage_host_ptr(uniq[i % len(uniq)].data_ptr(), 16, int(uniq[i % len(uniq)].shape[0]))
            lctx.segment_frame(cl)
            cl.release()
            lat.append((time.perf_counter() - t1) * 1e3)
        frames_info["frame_latency_ms"] = float(np.median(lat[2:]))
        frames_info["frame_latency_note"] = "one frame at a time: stage (H2D) + segment_frame + results, 1 context + 4 helper streams"
        lctx.close()

    # ---- BASELINE.json configs[2] (C3) in one line per model: a 50 000-point cluster x 10 000 hypotheses of the PCL sample
    # stream, estimate + score on the device, 20 calls back to back between two CUDA events (rank 0's GPU; per-GPU figure)
    primitives = None
    if rank == 0 and not args.no_primitives:
        from pitt_object_table_segmentation_b200 import scenes
        primitives = {"points": 50000, "hypotheses": 10000, "note": "per GPU; algorithmic flop per evaluation from SURVEY 8d "
                      "(sphere 10, cylinder 69, cone 91) against the live FFMA peak"}
        for kind, model, flop in (("sphere", A.MODEL_SPHERE, 10), ("cylinder", A.MODEL_CYLINDER, 69), ("cone", A.MODEL_CONE, 91)):
            pxyz, _ = scenes.primitive_cluster(kind, 50000, 5)
            pcloud = ctx.stage(pxyz)
            ctx.estimate_normals(pcloud, 50)
            pp = pkg.default_sac_params(model)
            ps = torch.from_numpy(ctx.pcl_sample_stream(pcloud, model, 10000)).to(dev)
            pc = torch.zeros(10000, dtype=torch.int32, device=dev)
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
    # ---- the whole C3 job as the reference would run it: 64 clusters of 5 000 .. 50 000 points (32 cylinders, 32 cones), per
    # cluster k = 50 normals + SACSegmentationFromNormals with setMaxIterations(10000) and PCL's adaptive stop + LM refinement;
    # clusters resident on the device, results (coefficients, inlier count) on the host. Clusters are independent units
    # (SURVEY 8e): size-balanced over the ranks (greedy_balance), and on each GPU over 8 contexts = 8 host threads + streams.
    if not args.no_primitives:
        from concurrent.futures import ThreadPoolExecutor
        from pitt_object_table_segmentation_b200 import scenes
        sizes = np.linspace(5000, 50000, 64).astype(int)
        owner = sharding.greedy_balance([int(v) for v in sizes], world)
        mine = [i for i in range(64) if owner[i] == rank]
        n_c3 = 8
        cctx = [pkg.Context(local_rank, seed=12345) for _ in range(n_c3)]
        slot = sharding.greedy_balance([int(sizes[i]) for i in mine], n_c3)
        jobs = [[] for _ in range(n_c3)]
        for j, i in enumerate(mine):
            kind, model = (("cylinder", A.MODEL_CYLINDER), ("cone", A.MODEL_CONE))[i % 2]
            cxyz, _ = scenes.primitive_cluster(kind, int(sizes[i]), 100 + i)
            pj = pkg.default_sac_params(model)
            pj.max_iterations = 10000
            jobs[slot[j]].append((cctx[slot[j]].stage(cxyz), pj))

        def c3_worker(k):
            tot = 0
            for cl, pj in jobs[k]:
                cctx[k].estimate_normals_device(cl, 50)
                tot += cctx[k].sac_segment_count_only(cl, pj)["n_inliers"]
            return tot

        def c3_pass(pool):
            return sum(pool.map(c3_worker, range(n_c3)))

        with ThreadPoolExecutor(n_c3) as pool:
            c3_pass(pool)
            torch.cuda.synchronize()
            barrier()
            t0 = time.perf_counter()
            inl_mine = c3_pass(pool)
            torch.cuda.synchronize()
            c3_t = torch.tensor([time.perf_counter() - t0, float(inl_mine)], dtype=torch.float64, device=dev)
        if world > 1:
            c3_max = c3_t.clone()
            dist.all_reduce(c3_max, op=dist.ReduceOp.MAX)
            dist.all_reduce(c3_t, op=dist.ReduceOp.SUM)
            c3_s, inl_total = float(c3_max[0].item()), int(c3_t[1].item())
        else:
            c3_s, inl_total = float(c3_t[0].item()), int(c3_t[1].item())
        if primitives is not None:
            primitives["c3_job"] = {"clusters": 64, "points_total": int(sizes.sum()), "max_iterations": 10000, "seconds": c3_s,
                                    "clusters_per_s": 64 / c3_s, "inliers_total": inl_total, "contexts_per_gpu": n_c3,
                                    "note": "wall clock, max over ranks; clusters size-balanced over ranks and contexts, "
                                            "normals + segment() each"}
        for k in range(n_c3):
            for cl, _ in jobs[k]:
                cl.release()
            cctx[k].close()

    line = None
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "C2: table-plane RANSAC only, 1M-point synthetic cloud x 5000 replayed "
                                   "mt19937(12345) hypotheses per GPU, thr 0.02, ALL_H, refine + select",
                       "points": n, "hypotheses_per_gpu": N_HYP, "l2": "flushed between timed iterations (256 MB write)",
                       "parallelism": "hypothesis split + NCCL all-gather of counts" if world > 1 else "single GPU"},
            "clocks": clk,
            "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": e2e_ms, "h2d_bytes_per_step": n * 16 + N_HYP * 12,
                    "d2h_bytes_per_step": d2h_bytes[0]},
            "gpu_launches": int(launches),
            "roofline": roofline,
            "frames": frames_info,
            "primitives": primitives,
            "wall_s_timed_region": wall,
        }
        if not args.no_cpu_baseline:
            from oracle import orc_binding as O
            line["cpu_baseline"] = cpu_baseline(xyz, samples, seconds_target=12.0)
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
