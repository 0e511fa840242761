import os, sys, ctypes as C
sys.path.insert(0, '/root/repo')
import numpy as np, torch
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
p.replay_samples = keep.ctypes.data_as(A.i32p); p.replay_count = H
inl = torch.empty(n, dtype=torch.int32).pin_memory()
n_inl, n_co = C.c_int(0), C.c_int(0)
co = np.zeros(8, np.float32)
import time
ts = []
for i in range(30):
    torch.cuda.synchronize()
    if os.environ.get("PITT_TRACE") == "2": print("---- call", i, file=sys.stderr)
    t0 = time.perf_counter()
    ctx.lib.pitt_sac_segment_host(ctx.handle, C.c_void_p(host.data_ptr()), 16, n, C.byref(p), C.cast(inl.data_ptr(), A.i32p), n, C.byref(n_inl), co.ctypes.data_as(A.f32p), C.byref(n_co), None)
    ts.append((time.perf_counter() - t0) * 1e3)
print("chunks", os.environ.get("PITT_STREAM_CHUNKS", "4"), "median fused call %.3f ms" % float(np.median(ts[5:])), "min %.3f" % min(ts[5:]))
