import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pitt_object_table_segmentation_b200 as pkg
from pitt_object_table_segmentation_b200 import scenes
ctx = pkg.Context(0)
ctx.set_workers(int(os.environ.get('PITT_WORKERS', '0')))
xyz = scenes.tabletop_frame(seed=12345)
cloud = ctx.stage(xyz)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    t0 = time.perf_counter(); fr = ctx.segment_frame(cloud); dt = time.perf_counter() - t0
l0 = ctx.kernel_launches
fr = ctx.segment_frame(cloud)
print("launches of the last frame", ctx.kernel_launches - l0)
print("frame ms", dt * 1e3, "device ms", fr["device_ms"], "launches", ctx.kernel_launches)
