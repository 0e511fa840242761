"""Where the end-to-end C2 step (host cloud in, inlier list out) spends its time."""
import ctypes as C
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import pitt_object_table_segmentation_b200 as pkg
from pitt_object_table_segmentation_b200 import _abi as A, scenes

n, H = 1_000_000, 5000
ctx = pkg.Context(0, seed=12345, stream=torch.cuda.current_stream().cuda_stream)
xyz = scenes.plane_outlier_cloud(n, seed=12345)
host = torch.from_numpy(xyz).pin_memory()
cloud = ctx.stage_host_ptr(host.data_ptr(), 16, n)
samples = ctx.pcl_sample_stream(cloud, A.MODEL_PLANE, H)
p = pkg.default_support_sac_params()
p.stop, p.max_iterations, p.sampler = A.STOP_ALL_H, H, A.SAMPLER_REPLAY
keep = np.ascontiguousarray(samples)
p.replay_samples = keep.ctypes.data_as(A.i32p)
p.replay_count = H
inl = np.empty(n, np.int32)
n_inl, n_co = C.c_int(0), C.c_int(0)
co = np.zeros(8, np.float32)
handle = C.c_void_p()


def med(fn, reps=15):
    ts = []
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        ts.append((time.perf_counter() - t0) * 1e3)
    return float(np.median(ts[3:]))


def stage():
    ctx.lib.pitt_stage_cloud(ctx.handle, C.c_void_p(host.data_ptr()), 16, n, C.byref(handle))


def seg():
    ctx.lib.pitt_sac_segment(ctx.handle, handle, C.byref(p), inl.ctypes.data_as(A.i32p), n, C.byref(n_inl),
                             co.ctypes.data_as(A.f32p), C.byref(n_co), None)


def seg_count_only():
    ctx.lib.pitt_sac_segment(ctx.handle, handle, C.byref(p), None, 0, C.byref(n_inl), co.ctypes.data_as(A.f32p), C.byref(n_co), None)


def rel():
    ctx.lib.pitt_release_cloud(ctx.handle, handle)


def both():
    stage(); seg(); rel()


stage()
print("segment (resident cloud, inliers to host):", med(seg), "ms; inliers", n_inl.value)
print("segment count only:", med(seg_count_only), "ms")
rel()
print("stage + release:", med(lambda: (stage(), rel())), "ms")
print("stage + segment + release:", med(both), "ms")
def fused():
    ctx.lib.pitt_sac_segment_host(ctx.handle, C.c_void_p(host.data_ptr()), 16, n, C.byref(p), inl.ctypes.data_as(A.i32p), n, C.byref(n_inl), co.ctypes.data_as(A.f32p), C.byref(n_co), None)
print("pitt_sac_segment_host (chunked copy overlapped):", med(fused), "ms; inliers", n_inl.value)
d = torch.empty(n * 4, dtype=torch.float32, device="cuda")
print("plain torch H2D 16 MB:", med(lambda: d.copy_(host.view(-1), non_blocking=True)), "ms")
hb = torch.empty(n_inl.value, dtype=torch.int32).pin_memory()
db = torch.zeros(n_inl.value, dtype=torch.int32, device="cuda")
print("plain torch D2H of the inlier list:", med(lambda: hb.copy_(db, non_blocking=True)), "ms")
