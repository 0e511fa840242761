"""Sphere scoring throughput: packed kernel (default on large jobs) against the generic score_kernel (pitt_debug_score_mode(1))."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import pitt_object_table_segmentation_b200 as pkg
from pitt_object_table_segmentation_b200 import _abi as A, scenes

ctx = pkg.Context(0, seed=1, stream=torch.cuda.current_stream().cuda_stream)
peak = ctx.fp32_peak(0)
for n in (50000, 5000, 20000, 500000):
    xyz, _ = scenes.primitive_cluster("sphere", n, 5)
    cloud = ctx.stage(xyz)
    p = pkg.default_sac_params(A.MODEL_SPHERE)
    H = 10000
    samples = ctx.pcl_sample_stream(cloud, A.MODEL_SPHERE, H)
    d_s = torch.from_numpy(samples).cuda()
    d_c = torch.zeros(H, dtype=torch.int32, device="cuda")
    res = {}
    for mode in (0, 1, 2, 3):
        ctx.lib.pitt_debug_score_mode(mode)
        for _ in range(10):
            ctx.sac_score_device(cloud, p, d_s.data_ptr(), H, d_c.data_ptr())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            ctx.sac_score_device(cloud, p, d_s.data_ptr(), H, d_c.data_ptr())
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        res[mode] = d_c.cpu().numpy().copy()
        ev = float(n) * H / (ms * 1e-3)
        print(f"n {n:7d} mode {mode}: {ms:.4f} ms  {ev / 1e12:.3f} Tevals/s  {ev * 10 / 1e12 / peak:.3f} of FFMA peak (10 flop/eval)", flush=True)
    ctx.lib.pitt_debug_score_mode(0)
    assert np.array_equal(res[0], res[1])
    cloud.release()
