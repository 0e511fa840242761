"""Resident vs end-to-end arms of the bench on the same distinct frames: report the frames whose responses differ."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pitt_object_table_segmentation_b200 as pkg
from pitt_object_table_segmentation_b200 import scenes
from concurrent.futures import ThreadPoolExecutor

n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
with ThreadPoolExecutor(8) as ex:
    frames = list(ex.map(lambda s: scenes.tabletop_frame(seed=s, random_poses=True), range(n)))
ctxs = [pkg.Context(0, seed=12345) for _ in range(16)]
if os.environ.get('FRAME_MODE'):
    pkg.load_library().pitt_debug_frame_mode(int(os.environ['FRAME_MODE']))
stager = pkg.Context(0, seed=12345)
clouds = [stager.stage(f) for f in frames]
passes = int(sys.argv[2]) if len(sys.argv) > 2 else 3
runs = []
for p in range(passes):
    runs.append(pkg.segment_clouds_batched(ctxs, clouds) if p % 2 == 0 else pkg.segment_frames_batched(ctxs, frames))
def key(r):
    return (r["n_supports"], r["n_clusters"], tuple(r["support_sizes"]), tuple(r["on_support_sizes"]), r["support_coefficients"].tobytes(),
            tuple((a["tag"], a["n_points"], a["inliers"], a["coefficients"].tobytes(), a["pc_centroid"].tobytes(), a["est_centroid"].tobytes()) for a in r["shapes"]))
bad = 0
for i in range(n):
    ks = [key(r[i]) for r in runs]
    if len(set(ks)) != 1:
        bad += 1
        print("frame", i, "differs in passes", [p for p in range(len(runs)) if ks[p] != max(set(ks), key=ks.count)])
        import ctypes as C
        for p_, r in enumerate(runs):
            pass
        seen = set()
        for r in runs:
            line = str((r[i]["debug_words"][:7], r[i]["support_sizes"], r[i]["on_support_sizes"], [(a["tag_name"], a["n_points"], a["inliers"], a["coefficients"].tolist()) for a in r[i]["shapes"]]))
            if line not in seen:
                seen.add(line)
                print("   ", line)
print("frames", n, "mismatching", bad)
