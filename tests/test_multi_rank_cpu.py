"""world_size-2 gloo test (CPU) of the N>1 host logic: hypothesis split + all-gather of counts +
earliest arg-max gives the single-process winner; frame/cluster sharding covers every unit once."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from pitt_object_table_segmentation_b200 import sharding  # noqa: E402


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import orc_binding as O
    from pitt_object_table_segmentation_b200 import _abi as A, scenes
    per_rank = 96
    xyz = scenes.plane_outlier_cloud(6000, seed=77)  # every rank holds the whole cloud
    stream = O.pcl_sample_stream(xyz, A.MODEL_PLANE, per_rank * world)
    lo, hi = sharding.hypothesis_slice(rank, world, per_rank)
    p = O.default_support_sac_params()
    counts, _, _ = O.sac_score(xyz, None, p, stream[lo:hi])
    local = torch.from_numpy(counts.astype(np.int32))
    gathered = [torch.zeros(per_rank, dtype=torch.int32) for _ in range(world)]
    dist.all_gather(gathered, local)
    all_counts = torch.cat(gathered).numpy()
    best, best_count = sharding.earliest_argmax(all_counts)
    np.save(os.path.join(out_dir, f"r{rank}.npy"), np.array([best, best_count], np.int64))
    if rank == 0:
        full, _, _ = O.sac_score(xyz, None, p, stream)
        np.save(os.path.join(out_dir, "full.npy"), full)
        np.save(os.path.join(out_dir, "gathered.npy"), all_counts)
    # frame sharding: every frame owned exactly once
    owned = torch.zeros(37, dtype=torch.int32)
    flo, fhi = sharding.block_range(rank, world, 37)
    owned[flo:fhi] = 1
    dist.all_reduce(owned)
    assert bool((owned == 1).all())
    dist.barrier()
    dist.destroy_process_group()


def test_hypothesis_split_two_ranks(tmp_path, built):
    world = 2
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    full = np.load(tmp_path / "full.npy")
    gathered = np.load(tmp_path / "gathered.npy")
    assert np.array_equal(full, gathered)
    want = sharding.earliest_argmax(full)
    for r in range(world):
        got = np.load(tmp_path / f"r{r}.npy")
        assert (int(got[0]), int(got[1])) == want


def test_sharding_helpers():
    for n, w in ((10, 3), (1024, 8), (5, 8), (0, 2)):
        cover = []
        for r in range(w):
            lo, hi = sharding.block_range(r, w, n)
            cover += list(range(lo, hi))
        assert cover == list(range(n))
    assert sharding.earliest_argmax([3, 9, 9, 1]) == (1, 9)
    sizes = [50000, 5000, 42000, 7000, 30000, 12000]
    owner = sharding.greedy_balance(sizes, 2)
    loads = [sum(s for s, o in zip(sizes, owner) if o == r) for r in range(2)]
    assert abs(loads[0] - loads[1]) <= max(sizes) and sorted(set(owner)) == [0, 1]
