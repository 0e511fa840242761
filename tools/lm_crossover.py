"""Levenberg-Marquardt refinement: one CTA against the cluster of 8 CTAs, by problem size (rows = inliers)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pitt_object_table_segmentation_b200 as pkg
from pitt_object_table_segmentation_b200 import _abi as A, scenes
ctx = pkg.Context(0)
for kind, model in (("cylinder", A.MODEL_CYLINDER), ("cone", A.MODEL_CONE), ("sphere", A.MODEL_SPHERE)):
    for n in (300, 1000, 2000, 4000, 8000, 16000, 50000):
        xyz, _ = scenes.primitive_cluster(kind, n, 100)
        cl = ctx.stage(xyz)
        ctx.estimate_normals_device(cl, 50)
        p = pkg.default_sac_params(model)
        p.optimize = 0
        base = ctx.sac_segment(cl, p)
        p.optimize = 1
        out = []
        for mode, rows in (("cta", 1 << 30), ("cluster", 1)):
            ctx.lib.pitt_debug_lm_cluster_min(rows)
            ts = []
            for _ in range(6):
                ctx.synchronize(); t0 = time.perf_counter()
                ref, info = ctx.sac_refine(cl, p, base["coeffs"], base["inliers"])
                ts.append((time.perf_counter() - t0) * 1e3)
            out.append((mode, float(np.median(ts[2:])), info.lm_nfev))
        ctx.lib.pitt_debug_lm_cluster_min(4096)
        print(kind, n, "inliers", len(base["inliers"]), " ".join("%s %.3f ms (nfev %d)" % o for o in out))
        cl.release()
