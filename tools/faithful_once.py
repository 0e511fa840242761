"""Probe: one raw camera-frame message through pitt_prefilter_cloud + pitt_segment_frame on one context (launch list under ncu,
phase times with PITT_TRACE=1); then the batched stream at 16 contexts."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import pitt_object_table_segmentation_b200 as pkg
from pitt_object_table_segmentation_b200 import scenes
import bench

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
ctx = pkg.Context(0, seed=12345)
pf = bench.faithful_prefilter()
raw = bench.make_frames([12345], raw=True)[0]
pin = torch.from_numpy(raw).pin_memory().numpy()
for r in range(reps):
    t0 = time.perf_counter()
    cloud, info = ctx.prefilter(pin, pf)
    t1 = time.perf_counter()
    l0 = ctx.kernel_launches
    fr = ctx.segment_frame(cloud)
    t2 = time.perf_counter()
    print("prefilter ms %.3f (n %d -> %d) frame ms %.3f device ms %.3f launches %d clusters %d" % (
        (t1 - t0) * 1e3, info["n_input"], cloud.n, (t2 - t1) * 1e3, fr["device_ms"], ctx.kernel_launches - l0, fr["n_clusters"]), flush=True)
    cloud.release()
if len(sys.argv) > 2:
    n_ctx = int(sys.argv[2])
    ctxs = [pkg.Context(0, seed=12345) for _ in range(n_ctx)]
    raws = [torch.from_numpy(f).pin_memory().numpy() for f in bench.make_frames(list(range(32)), raw=True)]
    for rep in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        pkg.segment_frames_batched(ctxs, raws, prefilter=pf)
        dt = time.perf_counter() - t0
        print("batched %d contexts: %.1f frames/s" % (n_ctx, len(raws) / dt), flush=True)
