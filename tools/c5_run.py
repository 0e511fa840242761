"""BASELINE.json configs[4] (C5) as written: one 50 M-point scene, 100 000 plane hypotheses split over the ranks
(torchrun, one rank per GPU), NCCL all-gather of the counts, earliest arg-max, refinement + final inliers on every rank.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/c5_run.py [points] [hypotheses]
Prints one JSON line (rank 0). Checks: a slice of every rank's counts against the exact CUDA-core kernel, and that all ranks
agree on the winner."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
import numpy as np
import torch
import torch.distributed as dist

import pitt_object_table_segmentation_b200 as pkg
from pitt_object_table_segmentation_b200 import _abi as A, scenes

n = int(sys.argv[1]) if len(sys.argv) > 1 else 50_000_000
H_total = int(sys.argv[2]) if len(sys.argv) > 2 else 100_000
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
ctx = pkg.Context(local, seed=12345, stream=torch.cuda.current_stream().cuda_stream)
H = H_total // world
t0 = time.perf_counter()
xyz = scenes.plane_outlier_cloud(n, seed=5)  # every rank generates the same scene (no broadcast needed)
t_gen = time.perf_counter() - t0
cloud = ctx.stage(xyz)
rng = np.random.default_rng(11)
samples_all = rng.integers(0, n, (H * world, 3)).astype(np.int32)
samples_all[:, 1] = (samples_all[:, 0] + 1 + rng.integers(0, n - 2, H * world)) % n
samples_all[:, 2] = (samples_all[:, 1] + 1 + rng.integers(0, n - 3, H * world)) % n
mine = np.ascontiguousarray(samples_all[rank * H:(rank + 1) * H])
p = pkg.default_support_sac_params()
p.sampler, p.stop, p.max_iterations = A.SAMPLER_REPLAY, A.STOP_ALL_H, H
p.replay_samples = mine.ctypes.data_as(A.i32p)
p.replay_count = H
d_s = torch.from_numpy(mine).to(dev)
d_all_s = torch.from_numpy(samples_all).to(dev)
d_c = torch.zeros(H, dtype=torch.int32, device=dev)
d_all = torch.zeros(H * world, dtype=torch.int32, device=dev)
d_best = torch.zeros(2, dtype=torch.int32, device=dev)


def step():
    ctx.sac_score_device(cloud, p, d_s.data_ptr(), H, d_c.data_ptr())
    if world > 1:
        dist.all_gather_into_tensor(d_all, d_c)
    else:
        d_all.copy_(d_c)
    ctx.argmax_counts_device(d_all.data_ptr(), H * world, d_best.data_ptr())
    return ctx.sac_finish_device(cloud, p, d_all_s.data_ptr(), H * world, d_best.data_ptr())


step()
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
ms = []
for _ in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    res = step()
    e1.record()
    torch.cuda.synchronize()
    ms.append(e0.elapsed_time(e1))
t = torch.tensor([float(np.median(ms))], dtype=torch.float64, device=dev)
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
# checks
counts = d_c.cpu().numpy()
pick = np.arange(0, H, max(1, H // 64))[:64]
ctx.lib.pitt_debug_plane_mode(1)
exact, _, _ = ctx.sac_score(cloud, p, mine[pick])
ctx.lib.pitt_debug_plane_mode(0)
ok_counts = bool(np.array_equal(counts[pick], exact))
best = d_best.cpu().numpy().astype(np.int64)
allc = d_all.cpu().numpy()
ok_winner = bool(int(best[0]) == int(np.argmax(allc)) and int(best[1]) == int(allc.max()))
flags = torch.tensor([1.0 if (ok_counts and ok_winner) else 0.0, float(best[0])], dtype=torch.float64, device=dev)
if world > 1:
    mn = flags.clone(); dist.all_reduce(mn, op=dist.ReduceOp.MIN)
    mx = flags.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    all_ok = bool(mn[0].item() == 1.0 and mn[1].item() == mx[1].item())
else:
    all_ok = ok_counts and ok_winner
if rank == 0:
    step_ms = float(t.item())
    print(json.dumps({"config": "C5", "points": n, "hypotheses": H * world, "n_gpus": world, "ms_per_step": step_ms,
                      "evals_per_s": float(n) * H * world / (step_ms * 1e-3), "winner": int(best[0]), "winner_count": int(best[1]),
                      "final_inliers": int(res["n_inliers"]) if isinstance(res, dict) and "n_inliers" in res else None,
                      "checks_ok": all_ok, "scene_generation_s": t_gen}))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
