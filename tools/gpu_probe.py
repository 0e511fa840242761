"""One-off probe run on the GPU box: FP32 pipe peaks and a first timing of plane scoring (C2)."""
import ctypes as C
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import pitt_object_table_segmentation_b200 as pkg
from pitt_object_table_segmentation_b200 import _abi as A, scenes

out = {}
ctx = pkg.Context(0, stream=torch.cuda.current_stream().cuda_stream)
for kind, name in ((0, "ffma"), (1, "fmul_fadd"), (2, "fmul2_ffma2")):
    out["fp32_peak_" + name] = ctx.fp32_peak(kind)
n, H = 1_000_000, 5000
xyz = scenes.plane_outlier_cloud(n, seed=12345)
cloud = ctx.stage(xyz)
rng = np.random.default_rng(0)
samples = rng.integers(0, n, (H, 3)).astype(np.int32)
d_samples = torch.from_numpy(samples).cuda()
d_counts = torch.zeros(H, dtype=torch.int32, device="cuda")
p = pkg.default_support_sac_params()
for force in (0, 1):
    ctx.lib.pitt_debug_force_generic_plane(force)
    for _ in range(3):
        ctx.sac_score_device(cloud, p, d_samples.data_ptr(), H, d_counts.data_ptr())
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    reps = 10
    for _ in range(reps):
        ctx.sac_score_device(cloud, p, d_samples.data_ptr(), H, d_counts.data_ptr())
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    out["plane_score_ms_" + ("generic" if force else "packed")] = ms
    out["plane_evals_per_s_" + ("generic" if force else "packed")] = n * H / (ms * 1e-3)
    out["counts_sum_" + ("generic" if force else "packed")] = int(d_counts.sum().item())
ctx.lib.pitt_debug_force_generic_plane(0)
print(json.dumps(out, indent=1))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/probe.json", "w"), indent=1)
