import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pitt_object_table_segmentation_b200 as pkg
from pitt_object_table_segmentation_b200 import scenes
ctx = pkg.Context(0)
name = sys.argv[2] if len(sys.argv) > 2 else "full"
if name == "full":
    xyz = scenes.tabletop_frame(seed=12345)
elif name == "voxel":
    xyz = scenes.voxel_downsample(scenes.tabletop_frame(seed=12345), 0.01)
else:
    rng = np.random.default_rng(0)
    xyz = np.ones((300_000, 4), np.float32)
    xyz[:, :2] = rng.uniform(-1, 1, (300_000, 2))
    xyz[:, 2] = rng.normal(0, 0.002, 300_000)
cloud = ctx.stage(xyz)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    ctx.estimate_normals_device(cloud, 50)
print("ms", ctx.last_device_ms)
