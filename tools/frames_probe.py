"""Probe: frames/s of pitt_segment_frames_batched vs the number of contexts (host threads + streams) on one GPU."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import pitt_object_table_segmentation_b200 as pkg
from pitt_object_table_segmentation_b200 import scenes

print("host cores", os.cpu_count())
uniq = [torch.from_numpy(scenes.tabletop_frame(seed=i, random_poses=True)).pin_memory() for i in range(4)]
out = {}
sub = int(os.environ.get("PROBE_SUBSAMPLE", "1"))
workers = int(os.environ.get("PROBE_WORKERS", "4"))
if sub > 1:
    uniq = [torch.from_numpy(np.ascontiguousarray(u.numpy()[::sub])).pin_memory() for u in uniq]
print("points per frame", uniq[0].shape[0], "workers", workers, "CUDA_DEVICE_MAX_CONNECTIONS", os.environ.get("CUDA_DEVICE_MAX_CONNECTIONS"))
for n_ctx in [int(a) for a in (sys.argv[1:] or ["1", "2", "4", "8", "16"])]:
    ctxs = [pkg.Context(0, seed=12345) for _ in range(n_ctx)]
    for c in ctxs:
        c.set_workers(workers)
    frames = [uniq[i % 4].numpy() for i in range(max(32, 8 * n_ctx))]
    pkg.segment_frames_batched(ctxs, frames[: 2 * n_ctx])
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    res = pkg.segment_frames_batched(ctxs, frames)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    out[n_ctx] = len(frames) / dt
    print(f"contexts {n_ctx:3d}: {len(frames) / dt:8.1f} frames/s  ({dt / len(frames) * 1e3 * n_ctx:.2f} ms per frame per context)", flush=True)
    for c in ctxs:
        c.close()
json.dump(out, open("gpurun_out/frames_probe.json", "w"))
