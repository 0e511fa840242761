"""Device time of the normal estimator on the full-resolution frame, the 1 cm voxel-grid frame and a uniform sheet, with the
number of queries that needed the general ring search. PITT_KNN_CAVG / PITT_KNN_NEED override the tuning knobs."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pitt_object_table_segmentation_b200 as pkg
from pitt_object_table_segmentation_b200 import scenes

ctx = pkg.Context(0)
full = scenes.tabletop_frame(seed=12345)
rng = np.random.default_rng(0)
sheet = np.ones((300_000, 4), np.float32)
sheet[:, :2] = rng.uniform(-1, 1, (300_000, 2))
sheet[:, 2] = rng.normal(0, 0.002, 300_000)
vol = np.ones((200_000, 4), np.float32)
vol[:, :3] = rng.uniform(-1, 1, (200_000, 3))
for name, xyz in (("full", full), ("voxel", scenes.voxel_downsample(full, 0.01)), ("sheet", sheet), ("volume", vol)):
    cloud = ctx.stage(xyz)
    for _ in range(3):
        ctx.estimate_normals_device(cloud, 50)
    t = []
    for _ in range(10):
        ctx.estimate_normals_device(cloud, 50)
        t.append(ctx.last_device_ms)
    st = (C.c_int64 * 16)()
    ctx.lib.pitt_debug_knn_stats(ctx.handle, 1, None)
    ctx.estimate_normals_device(cloud, 50)
    ctx.lib.pitt_debug_knn_stats(ctx.handle, 0, st)
    print(os.environ.get("PITT_KNN_CAVG"), os.environ.get("PITT_KNN_NEED"), name, len(xyz), "normals device ms %.4f" % float(np.median(t)),
          "finite", st[0], "ring", st[1], "attempts/level", list(st[2:6]), "cand/attempt %.0f" % (st[6] / max(1, sum(st[2:6]))),
          "over64 first/later", st[7], st[8], "under-k", st[9], "max cand", st[10], flush=True)
    cloud.release()
