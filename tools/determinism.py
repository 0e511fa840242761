"""Stress test: the same few frames through the frame stream many times on several contexts; every repetition of a frame must give
the same response."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pitt_object_table_segmentation_b200 as pkg
from pitt_object_table_segmentation_b200 import scenes

nctx, reps = int(sys.argv[1]) if len(sys.argv) > 1 else 8, int(sys.argv[2]) if len(sys.argv) > 2 else 16
uniq = [scenes.tabletop_frame(seed=s, random_poses=True) for s in range(4)]
frames = [uniq[i % 4] for i in range(4 * reps)]
ctxs = [pkg.Context(0, seed=12345) for _ in range(nctx)]
res = pkg.segment_frames_batched(ctxs, frames)
bad = 0
for i, r in enumerate(res):
    ref = res[i % 4]
    same = (r["n_clusters"] == ref["n_clusters"] and r["support_sizes"] == ref["support_sizes"] and r["on_support_sizes"] == ref["on_support_sizes"]
            and len(r["shapes"]) == len(ref["shapes"]) and
            all(a["inliers"] == b["inliers"] and a["n_points"] == b["n_points"] and np.array_equal(a["coefficients"], b["coefficients"]) and
                np.array_equal(a["pc_centroid"], b["pc_centroid"]) and np.array_equal(a["est_centroid"], b["est_centroid"])
                for a, b in zip(r["shapes"], ref["shapes"])))
    if not same:
        bad += 1
        if bad <= 3:
            print("frame", i, "differs:", [(a["n_points"], a["inliers"]) for a in r["shapes"]], "vs", [(a["n_points"], a["inliers"]) for a in ref["shapes"]],
                  r["support_sizes"], ref["support_sizes"], r["on_support_sizes"], ref["on_support_sizes"])
print(os.environ.get("PITT_DEBUG_CC_HOST"), "frames", len(res), "mismatching", bad)
