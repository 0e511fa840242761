"""Probe: stage timings of the frame pipeline (C1) and primitive scoring (C3 shape) on one B200."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import pitt_object_table_segmentation_b200 as pkg
from pitt_object_table_segmentation_b200 import _abi as A, scenes

out = {}
ctx = pkg.Context(0)


def timeit(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    t = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        t.append((time.perf_counter() - t0) * 1e3)
    return float(np.median(t))


full = scenes.tabletop_frame(seed=12345)
faithful = scenes.voxel_downsample(full, 0.01)
for name, xyz in (("full_307k", full), ("voxel_%d" % len(faithful), faithful)):
    cloud = ctx.stage(xyz)
    out[name + "_stage_ms"] = timeit(lambda: ctx.stage(xyz).release())
    out[name + "_normals_ms"] = timeit(lambda: ctx.estimate_normals(cloud, 50))
    out[name + "_supports_ms"] = timeit(lambda: ctx.find_supports(cloud))
    out[name + "_frame_ms"] = timeit(lambda: ctx.segment_frame(cloud), reps=3, warm=1)
    fr = ctx.segment_frame(cloud)
    out[name + "_frame_device_ms"] = fr["device_ms"]
    out[name + "_shapes"] = [(s["tag_name"], s["n_points"], s["inliers"]) for s in fr["shapes"]]
    sup = ctx.find_supports(cloud)
    on = sup["supports"][0]["on_support_cloud"]
    con = ctx.stage(on)
    out[name + "_on_support"] = len(on)
    out[name + "_cluster_ms"] = timeit(lambda: ctx.cluster_service(con))
    cl = ctx.cluster_service(con)
    c0 = np.ascontiguousarray(on[cl[0]["inliers"]])
    cc = ctx.stage(c0)
    out[name + "_cluster0_n"] = len(c0)
    out[name + "_cluster0_normals_ms"] = timeit(lambda: ctx.estimate_normals(cc, 50))
    for m in range(4):
        p = pkg.default_sac_params(m)
        out[name + "_cluster0_prim_%s_ms" % A.MODEL_NAMES[m]] = timeit(lambda: ctx.primitive_service(cc, p))
# C3-shaped scoring: 50k point cylinder / cone, 10k hypotheses, ALL_H
for kind, model in (("cylinder", A.MODEL_CYLINDER), ("cone", A.MODEL_CONE), ("sphere", A.MODEL_SPHERE)):
    xyz, _ = scenes.primitive_cluster(kind, 50000, 5)
    cloud = ctx.stage(xyz)
    ctx.estimate_normals(cloud, 50)
    rng = np.random.default_rng(0)
    S = A.SAMPLE_SIZE[model]
    samples = rng.integers(0, 50000, (10000, S)).astype(np.int32)
    p = pkg.default_sac_params(model)
    ms = timeit(lambda: ctx.sac_score(cloud, p, samples))
    out["c3_%s_score_ms" % kind] = ms
    out["c3_%s_device_ms" % kind] = ctx.last_device_ms
    out["c3_%s_evals_per_s" % kind] = 50000 * 10000 / (ctx.last_device_ms * 1e-3)
    p.stop, p.max_iterations, p.sampler = A.STOP_ALL_H, 10000, A.SAMPLER_PHILOX
    t0 = time.perf_counter(); r = ctx.sac_segment(cloud, p); out["c3_%s_segment_allh_ms" % kind] = (time.perf_counter() - t0) * 1e3
    out["c3_%s_segment_inliers" % kind] = len(r["inliers"])
print(json.dumps(out, indent=1))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/probe2.json", "w"), indent=1)
