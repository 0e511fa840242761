"""Where one C3 cluster (50 000-point cylinder, setMaxIterations(10000), adaptive stop) spends its time."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pitt_object_table_segmentation_b200 as pkg
from pitt_object_table_segmentation_b200 import _abi as A, scenes
ctx = pkg.Context(0)
for kind, model, n in (("cylinder", A.MODEL_CYLINDER, 50000), ("cone", A.MODEL_CONE, 50000), ("cylinder", A.MODEL_CYLINDER, 5000)):
    xyz, _ = scenes.primitive_cluster(kind, n, 100)
    cl = ctx.stage(xyz)
    p = pkg.default_sac_params(model)
    p.max_iterations = 10000
    for rep in range(3):
        t0 = time.perf_counter(); ctx.estimate_normals_device(cl, 50); ctx.synchronize(); t1 = time.perf_counter()
        r = ctx.sac_segment_count_only(cl, p); t2 = time.perf_counter()
        p2 = p.copy(); p2.optimize = 0
        r2 = ctx.sac_segment_count_only(cl, p2); t3 = time.perf_counter()
    print(kind, n, "normals %.2f ms, segment %.2f ms (without refinement %.2f ms), iterations %d, hypotheses %d, lm nfev %d, inliers %d"
          % ((t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, r["info"].iterations, r["info"].hypotheses, r["info"].lm_nfev, r["n_inliers"]))
