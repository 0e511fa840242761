"""GPU parity of the tensor-core plane scoring path (csrc/plane_tc.cu: tcgen05.mma kind::tf32 on exact 3-piece
TF32 splits, saturating-count epilogue, exact re-evaluation of undecided segments) against the CPU oracle's
countWithinDistance (SampleConsensusModelPlane): per-hypothesis counts must be bit-identical."""
import ctypes as C

import numpy as np
import pytest

import pitt_object_table_segmentation_b200 as pkg
from pitt_object_table_segmentation_b200 import _abi as A, scenes

pytestmark = pytest.mark.gpu

U = 2.0 ** -24
TC_ACC_ULPS = 8.0  # csrc/plane_tc.cu


def _score_tc(ctx, cloud, p, samples):
    """counts through the tensor path, (segments scored, segments re-evaluated exactly)"""
    out = (C.c_uint64 * 2)()
    ctx.lib.pitt_debug_plane_tc_stats(1, None)
    ctx.lib.pitt_debug_plane_mode(3)
    try:
        counts = ctx.sac_score(cloud, p, samples)[0]
        ctx.lib.pitt_debug_plane_tc_stats(0, out)
    finally:
        ctx.lib.pitt_debug_plane_tc_stats(0, None)
        ctx.lib.pitt_debug_plane_mode(0)
    return counts, int(out[0]), int(out[1])


@pytest.mark.parametrize("n,H", [(512 * 40, 1100), (512 * 40 + 1, 300), (512 * 40 + 511, 257), (700, 256), (129, 400),
                                 (5, 256), (512 * 9 + 130, 2561), (30000, 5200)])
def test_tc_counts_ragged_sizes(ctx, oracle, n, H):
    """ragged last tile (masked columns), H not a multiple of 128, more than one super-block of hypothesis blocks"""
    xyz = scenes.plane_outlier_cloud(n, seed=71)
    cloud = ctx.stage(xyz)
    p = pkg.default_support_sac_params()
    samples = np.random.default_rng(5).integers(0, n, (H, 3)).astype(np.int32)
    c_gpu, seg, redo = _score_tc(ctx, cloud, p, samples)
    c_cpu = oracle.sac_score(xyz, None, p, samples)[0]
    assert np.array_equal(c_gpu, c_cpu)
    assert seg > 0  # the tensor kernel did the work


def test_tc_points_on_the_threshold(ctx, oracle):
    """adversarial cloud: most points sit within a few ulps of |distance| == threshold of hypothesis 0, so its
    segments cannot be decided from the tensor result and must be re-evaluated in the exact operation order"""
    rng = np.random.default_rng(9)
    n = 512 * 24
    thr = np.float32(0.02)
    xyz = np.ones((n, 4), np.float32)
    xyz[:, 0] = rng.uniform(-1, 1, n)
    xyz[:, 1] = rng.uniform(-1, 1, n)
    sign = rng.choice([-1.0, 1.0], n).astype(np.float32)
    z = (sign * thr).astype(np.float32)
    for k in range(4):
        m = rng.random(n) < 0.4
        up = rng.random(n) < 0.5
        z = np.where(m & up, np.nextafter(z, np.float32(1)), np.where(m & ~up, np.nextafter(z, np.float32(-1)), z)).astype(np.float32)
    xyz[:, 2] = z
    xyz[:3] = [[0, 0, 0, 1], [1, 0, 0, 1], [0, 1, 0, 1]]  # the plane z = 0 as hypothesis 0
    cloud = ctx.stage(xyz)
    p = pkg.default_support_sac_params()
    samples = np.vstack([np.array([[0, 1, 2]], np.int32), rng.integers(0, n, (1023, 3)).astype(np.int32)])
    c_gpu, seg, redo = _score_tc(ctx, cloud, p, samples)
    c_cpu = oracle.sac_score(xyz, None, p, samples)[0]
    assert np.array_equal(c_gpu, c_cpu)
    assert redo >= n // 128  # at least every 128-point segment of hypothesis 0


@pytest.mark.parametrize("scale,thr", [(1.0, 0.02), (40.0, 0.02), (1e-3, 1e-4), (1.0, 3.0), (1e4, 0.05), (1.0, 1e-6)])
def test_tc_scales(ctx, oracle, scale, thr):
    """sigma and the decision window follow the cloud extent and the threshold; out-of-range cases fall back"""
    n = 512 * 30
    xyz = scenes.plane_outlier_cloud(n, seed=81)
    xyz[:, :3] *= np.float32(scale)
    cloud = ctx.stage(xyz)
    p = pkg.default_support_sac_params()
    p.distance_threshold = thr
    samples = np.random.default_rng(6).integers(0, n, (600, 3)).astype(np.int32)
    c_gpu, seg, redo = _score_tc(ctx, cloud, p, samples)
    c_cpu = oracle.sac_score(xyz, None, p, samples)[0]
    assert np.array_equal(c_gpu, c_cpu)
    if seg and thr > 1e-5:
        assert redo < 0.25 * seg  # the window stays narrow on ordinary clouds


def test_tc_far_from_origin_cancellation(ctx, oracle):
    """|d| ~ 60 m against a 2 cm threshold: the dot product cancels 12 bits, the split must stay exact"""
    n = 512 * 20
    xyz = scenes.plane_outlier_cloud(n, seed=83)
    xyz[:, :3] += np.float32(37.0)
    cloud = ctx.stage(xyz)
    p = pkg.default_support_sac_params()
    samples = np.random.default_rng(8).integers(0, n, (512, 3)).astype(np.int32)
    c_gpu, seg, redo = _score_tc(ctx, cloud, p, samples)
    c_cpu = oracle.sac_score(xyz, None, p, samples)[0]
    assert np.array_equal(c_gpu, c_cpu)


def test_tc_non_finite_points_fall_back(ctx, oracle):
    n = 512 * 20
    xyz = scenes.plane_outlier_cloud(n, seed=91)
    xyz[100, 0] = np.nan
    xyz[2000, 2] = np.inf
    xyz[7000, :3] = np.nan
    cloud = ctx.stage(xyz)
    p = pkg.default_support_sac_params()
    samples = np.random.default_rng(7).integers(0, n, (512, 3)).astype(np.int32)
    samples[5] = [100, 3, 4]  # a hypothesis built on a NaN point
    c_gpu, seg, redo = _score_tc(ctx, cloud, p, samples)
    c_cpu = oracle.sac_score(xyz, None, p, samples)[0]
    assert np.array_equal(c_gpu, c_cpu)
    assert seg == 0  # exact kernel took over


def test_tc_degenerate_hypotheses_score_zero(ctx, oracle):
    """collinear / duplicate samples give invalid hypotheses: all-NaN records, certain outliers on the tensor path"""
    n = 512 * 10
    xyz = scenes.plane_outlier_cloud(n, seed=31)
    xyz[10] = xyz[11]
    xyz[12, :3] = 2 * xyz[11, :3] - xyz[13, :3]
    cloud = ctx.stage(xyz)
    p = pkg.default_support_sac_params()
    samples = np.random.default_rng(3).integers(0, n, (300, 3)).astype(np.int32)
    samples[:4] = [[10, 11, 40], [13, 11, 12], [1, 2, 3], [11, 10, 10]]
    c_gpu, seg, redo = _score_tc(ctx, cloud, p, samples)
    c_cpu, _, v_cpu = oracle.sac_score(xyz, None, p, samples)
    assert np.array_equal(c_gpu, c_cpu)
    assert (c_gpu[v_cpu == 0] == 0).all()


@pytest.mark.parametrize("shift", [0.0, 37.0])
def test_tc_accumulator_error_is_inside_the_bound(ctx, shift):
    """numerics probe: the raw TMEM accumulators of hypothesis block 0 x points 0..255 against the real dot product;
    the error the decision window is derived from (TC_ACC_ULPS u m) must hold with a factor 2 to spare"""
    n, H = 20000, 512
    xyz = scenes.plane_outlier_cloud(n, seed=1)
    xyz[:, :3] += np.float32(shift)
    cloud = ctx.stage(xyz)
    p = pkg.default_support_sac_params()
    samples = np.random.default_rng(2).integers(0, n, (H, 3)).astype(np.int32)
    ctx.lib.pitt_debug_plane_tc_dump(1, None)
    ctx.lib.pitt_debug_plane_mode(3)
    try:
        counts, co, valid = ctx.sac_score(cloud, p, samples)
    finally:
        ctx.lib.pitt_debug_plane_mode(0)
        buf = np.zeros(128 * 256 + 2, np.float32)
        got = ctx.lib.pitt_debug_plane_tc_dump(0, buf.ctypes.data_as(C.POINTER(C.c_float)))
    assert got == buf.size
    acc = buf[:-2].reshape(128, 256).astype(np.float64)
    sigma = float(buf[-2])
    assert sigma >= 1.0 and np.log2(sigma) == int(np.log2(sigma))
    a = co[:128, :4].astype(np.float64)
    pts = xyz[:256, :3].astype(np.float64)
    real = a[:, :3] @ pts.T + a[:, 3:4]
    m = np.abs(a[:, :3]) @ np.abs(pts.T) + np.abs(a[:, 3:4])
    ok = valid[:128].astype(bool)
    err = (np.abs(acc / sigma - real) / (U * m))[ok]
    assert err.max() < TC_ACC_ULPS / 2, err.max()


def test_tc_is_the_default_for_large_jobs(ctx, oracle):
    """mode 0 routes jobs of >= 2^27 evaluations through the tensor kernel"""
    n, H = 300_000, 512
    xyz = scenes.plane_outlier_cloud(n, seed=12)
    cloud = ctx.stage(xyz)
    p = pkg.default_support_sac_params()
    samples = np.random.default_rng(4).integers(0, n, (H, 3)).astype(np.int32)
    out = (C.c_uint64 * 2)()
    ctx.lib.pitt_debug_plane_tc_stats(1, None)
    try:
        c_auto = ctx.sac_score(cloud, p, samples)[0]
        ctx.lib.pitt_debug_plane_tc_stats(0, out)
    finally:
        ctx.lib.pitt_debug_plane_tc_stats(0, None)
    assert int(out[0]) > 0
    pick = np.random.default_rng(5).integers(0, H, 48)
    assert np.array_equal(c_auto[pick], oracle.sac_score(xyz, None, p, samples[pick])[0])
