import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pitt_object_table_segmentation_b200 as pkg
from pitt_object_table_segmentation_b200 import scenes
ctx = pkg.Context(0)
for name, xyz in (("full", scenes.tabletop_frame(seed=12345)), ("voxel", None)):
    if xyz is None:
        xyz = scenes.voxel_downsample(scenes.tabletop_frame(seed=12345), 0.01)
    cloud = ctx.stage(xyz)
    for _ in range(3):
        ctx.estimate_normals(cloud, 50)
    t = []
    for _ in range(10):
        ctx.estimate_normals(cloud, 50)
        t.append(ctx.last_device_ms)
    print(os.environ.get("PITT_KNN_CELL_DIV"), name, len(xyz), "normals device ms", float(np.median(t)))
