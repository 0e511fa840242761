"""Sustained loop of one plane scoring mode while nvidia-smi samples clocks/power (tensor path clock investigation)."""
import os, subprocess, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pitt_object_table_segmentation_b200 as pkg
from pitt_object_table_segmentation_b200 import scenes
from tc_check import rand_samples

mode, var, secs = int(sys.argv[1]), int(sys.argv[2]), float(sys.argv[3])
ctx = pkg.Context(0, seed=1, stream=torch.cuda.current_stream().cuda_stream)
n, H = 1_000_000, 5000
xyz = scenes.plane_outlier_cloud(n, seed=12345)
cloud = ctx.stage(xyz)
p = pkg.default_support_sac_params()
d_s = torch.from_numpy(rand_samples(n, H, 7)).cuda()
d_c = torch.zeros(H, dtype=torch.int32, device="cuda")
ctx.lib.pitt_debug_plane_tc_variant(var)
ctx.lib.pitt_debug_plane_mode(mode)
for _ in range(3):
    ctx.sac_score_device(cloud, p, d_s.data_ptr(), H, d_c.data_ptr())
torch.cuda.synchronize()
smi = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.sw_power_cap,clocks_event_reasons.hw_slowdown,clocks_event_reasons.sw_thermal_slowdown",
                        "--format=csv,noheader", "-lms", "100"], stdout=subprocess.PIPE, text=True)
t0 = time.time(); k = 0
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
while time.time() - t0 < secs:
    for _ in range(20):
        ctx.sac_score_device(cloud, p, d_s.data_ptr(), H, d_c.data_ptr())
    k += 20
    torch.cuda.synchronize()
e1.record(); torch.cuda.synchronize()
smi.terminate()
lines = smi.stdout.read().strip().splitlines()
print(f"mode {mode} variant {var}: {e0.elapsed_time(e1) / k:.3f} ms per call sustained over {k} calls")
for l in lines[2::4][:8]:
    print("   ", l)
