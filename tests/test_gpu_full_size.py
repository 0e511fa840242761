"""GPU parity at the full sizes of BASELINE.json's configs, through size-independent properties and oracle spot
checks (the oracle cannot score 5e9 evaluations in seconds, a random sample of hypotheses can be)."""
import numpy as np
import pytest

import pitt_object_table_segmentation_b200 as pkg
from pitt_object_table_segmentation_b200 import _abi as A, scenes

pytestmark = pytest.mark.gpu


def _device_counts(ctx, cloud, p, samples):
    import torch
    dev = torch.device("cuda", 0)
    d_s = torch.from_numpy(np.ascontiguousarray(samples)).to(dev)
    d_c = torch.zeros(len(samples), dtype=torch.int32, device=dev)
    torch.cuda.synchronize()
    ctx.sac_score_device(cloud, p, d_s.data_ptr(), len(samples), d_c.data_ptr())
    ctx.synchronize()
    return d_c.cpu().numpy()


def test_c2_full_size_plane(ctx, oracle):
    """config 2: 1 M points x 5000 replayed hypotheses. (i) the tensor-core kernel, the FFMA filter kernel and the exact packed kernel agree on
    all 5000 counts, (ii) 40 random hypotheses + the winner agree with the oracle, (iii) the winner's refined model
    and final inlier set equal the oracle's refine + select of the same winner."""
    xyz = scenes.plane_outlier_cloud(1_000_000, seed=12345)
    cloud = ctx.stage(xyz)
    H = 5000
    samples = ctx.pcl_sample_stream(cloud, A.MODEL_PLANE, H)
    p = pkg.default_support_sac_params()
    ctx.lib.pitt_debug_plane_mode(2)
    c_filter = _device_counts(ctx, cloud, p, samples)
    ctx.lib.pitt_debug_plane_mode(3)
    c_tensor = _device_counts(ctx, cloud, p, samples)
    ctx.lib.pitt_debug_plane_mode(1)
    c_exact = _device_counts(ctx, cloud, p, samples)
    ctx.lib.pitt_debug_plane_mode(0)
    c_auto = _device_counts(ctx, cloud, p, samples)
    assert np.array_equal(c_filter, c_exact) and np.array_equal(c_tensor, c_exact) and np.array_equal(c_auto, c_exact)
    rng = np.random.default_rng(0)
    pick = np.unique(np.concatenate([rng.integers(0, H, 40), [int(np.argmax(c_exact))]]))
    c_cpu, co_cpu, _ = oracle.sac_score(xyz, None, p, samples[pick])
    assert np.array_equal(c_exact[pick], c_cpu)
    # whole segment() with ALL_H on the device vs the oracle fed with the device's winner
    p.sampler, p.stop, p.max_iterations = A.SAMPLER_REPLAY, A.STOP_ALL_H, H
    p.replay_samples = samples.ctypes.data_as(A.i32p)
    p.replay_count = H
    got = ctx.sac_segment(cloud, p)
    win = got["info"].best_hypothesis
    assert win == int(np.argmax(c_exact)) and got["info"].best_count == int(c_exact.max())  # earliest arg-max
    w_co = oracle.sac_score(xyz, None, p, samples[win:win + 1])[1][0, :4]
    inl0 = oracle.sac_select(xyz, None, p, w_co)
    ref, _ = oracle.sac_refine(xyz, None, p, w_co, inl0)
    assert np.array_equal(got["coeffs"].view(np.uint32), ref[:4].view(np.uint32))
    assert np.array_equal(got["inliers"], oracle.sac_select(xyz, None, p, ref[:4]))
    assert len(got["inliers"]) > 690_000


@pytest.mark.parametrize("kind,model", [("cylinder", A.MODEL_CYLINDER), ("cone", A.MODEL_CONE)])
def test_c3_large_cluster_scoring(ctx, oracle, kind, model):
    """config 3 shape: a 50 000-point cluster, 10 000 hypotheses; a random sample of them against the oracle"""
    xyz, _ = scenes.primitive_cluster(kind, 50_000, 11)
    cloud = ctx.stage(xyz)
    nrm = ctx.estimate_normals(cloud, 50)
    rng = np.random.default_rng(1)
    samples = rng.integers(0, 50_000, (10_000, A.SAMPLE_SIZE[model])).astype(np.int32)
    p = pkg.default_sac_params(model)
    counts = _device_counts(ctx, cloud, p, samples)
    pick = rng.integers(0, 10_000, 24)
    c_cpu = oracle.sac_score(xyz, nrm, p, samples[pick])[0]
    assert np.array_equal(counts[pick], c_cpu)
    assert counts.max() > 25_000


def test_c5_size_plane_slice(ctx, oracle):
    """config 5 shape on one GPU: the 50 M-point cloud (0.8 GB) with one rank's slice of the hypothesis stream;
    the tensor-core kernel, the filter and the exact kernel agree everywhere and 6 hypotheses agree with the oracle"""
    n = 50_000_000
    xyz = scenes.plane_outlier_cloud(n, seed=5)
    cloud = ctx.stage(xyz)
    rng = np.random.default_rng(2)
    H = 1024
    samples = rng.integers(0, n, (H, 3)).astype(np.int32)
    p = pkg.default_support_sac_params()
    ctx.lib.pitt_debug_plane_mode(2)
    c_filter = _device_counts(ctx, cloud, p, samples)
    ctx.lib.pitt_debug_plane_mode(3)
    c_tensor = _device_counts(ctx, cloud, p, samples)
    ctx.lib.pitt_debug_plane_mode(1)
    c_exact = _device_counts(ctx, cloud, p, samples)
    ctx.lib.pitt_debug_plane_mode(0)
    assert np.array_equal(c_filter, c_exact) and np.array_equal(c_tensor, c_exact)
    pick = rng.integers(0, H, 6)
    c_cpu = oracle.sac_score(xyz, None, p, samples[pick])[0]
    assert np.array_equal(c_exact[pick], c_cpu)
    cloud.release()


def test_segment_host_chunks_a_big_cloud_by_itself(ctx):
    """20 M points (320 MB): pitt_sac_segment_host copies in 2 chunks on its second stream and scores the first while the
    second travels; result identical to the staged sequence"""
    n = 20_000_000
    xyz = scenes.plane_outlier_cloud(n, seed=77)
    rng = np.random.default_rng(3)
    H = 300
    samples = rng.integers(0, n, (H, 3)).astype(np.int32)
    p = pkg.default_support_sac_params()
    p.stop, p.max_iterations, p.sampler = A.STOP_ALL_H, H, A.SAMPLER_REPLAY
    keep = np.ascontiguousarray(samples)
    p.replay_samples = keep.ctypes.data_as(A.i32p)
    p.replay_count = H
    got = ctx.sac_segment_host(xyz, p)
    cloud = ctx.stage(xyz)
    staged = ctx.sac_segment(cloud, p)
    cloud.release()
    assert np.array_equal(got["inliers"], staged["inliers"])
    assert np.array_equal(got["coeffs"].view(np.uint32), staged["coeffs"].view(np.uint32))
    assert got["info"].best_hypothesis == staged["info"].best_hypothesis and got["info"].best_count == staged["info"].best_count
    assert len(got["inliers"]) > n // 2
