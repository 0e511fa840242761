import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pitt_object_table_segmentation_b200 as pkg
from pitt_object_table_segmentation_b200 import scenes
ctx = pkg.Context(0)
ctx.set_workers(0)
xyz = scenes.tabletop_frame(seed=12345)
cloud = ctx.stage(xyz)
for i in range(3):
    sys.stderr.write(f"---- frame {i}\n")
    ctx.segment_frame(cloud)
