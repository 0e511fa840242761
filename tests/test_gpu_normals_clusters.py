"""GPU parity: k-NN + normal estimation (pc_manager.cpp:68-78) and Euclidean clustering
(cluster_segmentation_srv.cpp:57-69) through the C ABI vs the CPU oracle."""
import numpy as np
import pytest

import pitt_object_table_segmentation_b200 as pkg
from pitt_object_table_segmentation_b200 import scenes

pytestmark = pytest.mark.gpu


def _frame(seed, w=160, h=120):
    return scenes.tabletop_frame(seed=seed, width=w, height=h)


@pytest.mark.parametrize("k", [1, 7, 50, 64])
def test_knn_lists_identical(ctx, oracle, k):
    xyz = _frame(3)
    cloud = ctx.stage(xyz)
    gi, gs = ctx.knn(cloud, k)
    ci, cs = oracle.knn(xyz, k)
    assert np.array_equal(gi, ci)
    assert np.array_equal(gs, cs)


def test_knn_with_exact_ties_and_duplicates(ctx, oracle):
    # integer lattice: many exactly equal distances -> (distance, index) order decides
    g = np.stack(np.meshgrid(np.arange(12), np.arange(12), np.arange(3), indexing="ij"), -1).reshape(-1, 3)
    xyz = np.ones((len(g) + 5, 4), np.float32)
    xyz[: len(g), :3] = g * 0.01
    xyz[len(g):, :3] = xyz[:5, :3]  # duplicates
    cloud = ctx.stage(xyz)
    gi, gs = ctx.knn(cloud, 20)
    ci, cs = oracle.knn(xyz, 20)
    assert np.array_equal(gi, ci) and np.array_equal(gs, cs)


@pytest.mark.parametrize("jitter", [0.0, 1e-9])
def test_knn_grid_path_with_ties_and_near_ties(ctx, oracle, jitter):
    # Large enough for the multi-level grid path (> 4096 points). A lattice gives many EXACTLY equal distances per query
    # (index order decides); with a 1e-9 jitter the same distances differ by a few ulps: both cases land in the
    # "ambiguous" groups of the 32-bit sort (upper 26 bits of d2 equal), which are ordered by their true 64-bit keys.
    g = np.stack(np.meshgrid(np.arange(96), np.arange(96), np.arange(2), indexing="ij"), -1).reshape(-1, 3)
    rng = np.random.default_rng(11)
    xyz = np.ones((len(g), 4), np.float32)
    xyz[:, :3] = (g * 0.01 + rng.uniform(-jitter, jitter, g.shape)).astype(np.float32)
    cloud = ctx.stage(xyz)
    for k in (20, 50):
        gi, gs = ctx.knn(cloud, k)
        ci, cs = oracle.knn(xyz, k)
        assert np.array_equal(gs, cs)
        assert np.array_equal(gi, ci)
    got = ctx.estimate_normals(cloud, 50, (0.0, 0.0, 1.0))
    want = oracle.estimate_normals(xyz, 50, (0.0, 0.0, 1.0))
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


def test_knn_small_clouds(ctx, oracle):
    for n in (1, 2, 5, 49):
        xyz = _frame(5)[:n]
        cloud = ctx.stage(xyz)
        gi, gs = ctx.knn(cloud, 50)
        ci, cs = oracle.knn(xyz, 50)
        assert np.array_equal(gi, ci) and np.array_equal(gs, cs)


@pytest.mark.parametrize("seed,w,h", [(1, 160, 120), (2, 320, 240)])
def test_normals_bit_exact(ctx, oracle, seed, w, h):
    xyz = _frame(seed, w, h)
    cloud = ctx.stage(xyz)
    got = ctx.estimate_normals(cloud, 50, (0.0, 0.0, 0.0))
    want = oracle.estimate_normals(xyz, 50, (0.0, 0.0, 0.0))
    # tolerance of the north star is 1e-5 relative; the implementation is bit-exact
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    # sanity: table normals are vertical, unit length
    nn = np.linalg.norm(got[:, :3], axis=1)
    assert np.allclose(nn, 1.0, atol=1e-5)


def test_normals_viewpoint_flip_and_nan_points(ctx, oracle):
    xyz = _frame(7).copy()
    xyz[100, 0] = np.nan
    xyz[200, 2] = np.inf
    cloud = ctx.stage(xyz)
    for vp in ((0.0, 0.0, 5.0), (0.0, 0.0, -5.0)):
        got = ctx.estimate_normals(cloud, 30, vp)
        want = oracle.estimate_normals(xyz, 30, vp)
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
        assert np.isnan(got[100]).all() and np.isnan(got[200]).all()
    assert (ctx.estimate_normals(cloud, 30, (0, 0, 5.0))[:50, 2] > 0).all()
    assert (ctx.estimate_normals(cloud, 30, (0, 0, -5.0))[:50, 2] < 0).all()


def test_normals_fewer_points_than_k(ctx, oracle):
    xyz = _frame(9)[:31]
    cloud = ctx.stage(xyz)
    got = ctx.estimate_normals(cloud, 50)
    want = oracle.estimate_normals(xyz, 50)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


def _objects_only(seed):
    xyz = scenes.tabletop_frame(seed=seed, width=320, height=240)
    return np.ascontiguousarray(xyz[xyz[:, 2] > 0.012])


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_clusters_identical_labels(ctx, oracle, seed):
    xyz = _objects_only(seed)
    cloud = ctx.stage(xyz)
    n = len(xyz)
    mn, mx = int(round(n * 0.01)), int(round(n * 0.99))
    gl, gn = ctx.euclidean_clusters(cloud, 0.03, mn, mx)
    cl, cn = oracle.euclidean_clusters(xyz, 0.03, mn, mx)
    assert gn == cn == 3
    assert np.array_equal(gl, cl)  # identical including the PCL ordering (not just up to permutation)


def test_clusters_size_filter_and_noise(ctx, oracle):
    rng = np.random.default_rng(0)
    blobs = [rng.normal(c, 0.004, (m, 3)) for c, m in (((0, 0, 0), 400), ((0.2, 0, 0), 400), ((0, 0.3, 0), 90), ((0.5, 0.5, 0), 7))]
    noise = rng.uniform(-1, 1, (60, 3)) + 3.0
    xyz = np.ones((sum(len(b) for b in blobs) + 60, 4), np.float32)
    xyz[:, :3] = np.concatenate(blobs + [noise])
    xyz = xyz[rng.permutation(len(xyz))]
    cloud = ctx.stage(xyz)
    for mn, mx in ((1, 10**9), (10, 10**9), (50, 399), (401, 10**9)):
        gl, gn = ctx.euclidean_clusters(cloud, 0.03, mn, mx)
        cl, cn = oracle.euclidean_clusters(xyz, 0.03, mn, mx)
        assert gn == cn
        assert np.array_equal(gl, cl)


def test_clusters_chain_at_threshold(ctx, oracle):
    # points spaced exactly at / just under the tolerance: strict '<' decides connectivity
    tol = 0.03
    xs = np.concatenate([np.arange(20) * np.float32(0.0299), 1.0 + np.arange(20) * np.float32(0.03)])
    xyz = np.ones((40, 4), np.float32)
    xyz[:, 0] = xs
    xyz[:, 1:3] = 0
    cloud = ctx.stage(xyz)
    gl, gn = ctx.euclidean_clusters(cloud, tol, 1, 100)
    cl, cn = oracle.euclidean_clusters(xyz, tol, 1, 100)
    assert gn == cn and np.array_equal(gl, cl)
