"""C3-shaped scoring throughput: 50 000-point cylinder / cone / sphere cluster x 10 000 hypotheses, ALL_H.
Two-tier kernel (default) against the generic score_kernel (pitt_debug_score_mode(1)); CUDA events inside the library."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import pitt_object_table_segmentation_b200 as pkg
from pitt_object_table_segmentation_b200 import _abi as A, scenes

FLOP = {"sphere": 10, "cylinder": 69, "cone": 91}
import torch

ctx = pkg.Context(0, seed=1, stream=torch.cuda.current_stream().cuda_stream)
peak = ctx.fp32_peak(0)
out = {"fp32_ffma_peak_tflops": peak}
for kind, model in (("cylinder", A.MODEL_CYLINDER), ("cone", A.MODEL_CONE), ("sphere", A.MODEL_SPHERE)):
    for n in (50000, 5000):
        xyz, _ = scenes.primitive_cluster(kind, n, 5)
        cloud = ctx.stage(xyz)
        ctx.estimate_normals(cloud, 50)
        p = pkg.default_sac_params(model)
        H = 10000
        for sampler in ("pcl", "random"):
            if sampler == "pcl":
                samples = ctx.pcl_sample_stream(cloud, model, H)
            else:
                samples = np.random.default_rng(0).integers(0, n, (H, A.SAMPLE_SIZE[model])).astype(np.int32)
            for mode in (0, 1):
                if kind == "sphere" and mode == 1:
                    continue
                ctx.lib.pitt_debug_score_mode(mode)
                counts, _, valid = ctx.sac_score(cloud, p, samples)
                d_s = torch.from_numpy(samples).cuda()
                d_c = torch.zeros(H, dtype=torch.int32, device="cuda")
                for _ in range(10):  # warm-up: clocks ramp up only under continuous load
                    ctx.sac_score_device(cloud, p, d_s.data_ptr(), H, d_c.data_ptr())
                ms = []
                for _ in range(3):  # estimate + score on device-resident samples, 20 calls back to back between two events
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    for _ in range(20):
                        ctx.sac_score_device(cloud, p, d_s.data_ptr(), H, d_c.data_ptr())
                    e1.record()
                    torch.cuda.synchronize()
                    ms.append(e0.elapsed_time(e1) / 20)
                ctx.lib.pitt_debug_score_mode(0)
                assert np.array_equal(d_c.cpu().numpy(), counts)
                t = float(np.median(ms))
                ev = n * H / (t * 1e-3)
                key = f"{kind}_n{n}_{sampler}_{'two_tier' if mode == 0 else 'generic'}"
                out[key] = {"device_ms": t, "gevals_per_s": ev / 1e9, "algorithmic_tflops": ev * FLOP[kind] / 1e12,
                            "frac_of_ffma_peak": ev * FLOP[kind] / 1e12 / peak, "valid_hypotheses": int(valid.sum()),
                            "mean_inlier_fraction": float(counts[valid.astype(bool)].mean() / n) if valid.any() else 0.0}
                print(key, json.dumps(out[key]))
        cloud.release()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/c3_probe.json", "w"), indent=1)
