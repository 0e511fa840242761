"""GPU parity: plane RANSAC (supports_segmentation_srv.cpp:89-111, plane_segmentation_srv.cpp:52-67)
through the C ABI vs the CPU oracle, bit-exact."""
import ctypes as C

import numpy as np
import pytest

import pitt_object_table_segmentation_b200 as pkg
from pitt_object_table_segmentation_b200 import _abi as A, scenes

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,seed", [(3, 1), (17, 2), (1000, 3), (20000, 4), (100003, 5)])
def test_segment_supports_params_bit_exact(ctx, oracle, n, seed):
    xyz = scenes.plane_outlier_cloud(n, seed=seed)
    cloud = ctx.stage(xyz)
    got = ctx.sac_segment(cloud, pkg.default_support_sac_params())
    want = oracle.sac_segment(xyz, None, oracle.default_support_sac_params())
    assert np.array_equal(got["inliers"], want["inliers"])
    assert np.array_equal(got["coeffs"], want["coeffs"])
    gi, wi = got["info"], want["info"]
    assert (gi.iterations, gi.skipped, gi.best_hypothesis, gi.best_count, gi.n_inliers_model) == \
           (wi.iterations, wi.skipped, wi.best_hypothesis, wi.best_count, wi.n_inliers_model)
    assert np.array_equal(np.array(gi.model_coeffs[:4]), np.array(wi.model_coeffs[:4]))


@pytest.mark.parametrize("optimize", [0, 1])
def test_segment_primitive_plane_params(ctx, oracle, optimize):
    xyz = scenes.plane_outlier_cloud(30000, seed=11, plane_frac=0.4, sigma=0.003)
    cloud = ctx.stage(xyz)
    p = pkg.default_sac_params(A.MODEL_PLANE)
    p.optimize = optimize
    got = ctx.sac_segment(cloud, p)
    want = oracle.sac_segment(xyz, None, p)
    assert np.array_equal(got["inliers"], want["inliers"])
    assert np.array_equal(got["coeffs"], want["coeffs"])
    assert got["info"].iterations == want["info"].iterations


def test_score_counts_all_hypotheses(ctx, oracle):
    """per-hypothesis counts, coefficients and validity of a replayed mt19937 sample stream"""
    xyz = scenes.plane_outlier_cloud(40000, seed=21)
    cloud = ctx.stage(xyz)
    p = pkg.default_support_sac_params()
    samples = oracle.pcl_sample_stream(xyz, A.MODEL_PLANE, 700)
    assert np.array_equal(ctx.pcl_sample_stream(cloud, A.MODEL_PLANE, 700), samples)
    # (generic kernel) / (FFMA filter + exact re-evaluation) / (exact packed kernel)
    for force_generic, plane_mode in ((1, 0), (0, 2), (0, 1), (0, 0)):
        ctx.lib.pitt_debug_force_generic_plane(force_generic)
        ctx.lib.pitt_debug_plane_mode(plane_mode)
        c_gpu, co_gpu, v_gpu = ctx.sac_score(cloud, p, samples)
        ctx.lib.pitt_debug_force_generic_plane(0)
        ctx.lib.pitt_debug_plane_mode(0)
        c_cpu, co_cpu, v_cpu = oracle.sac_score(xyz, None, p, samples)
        assert np.array_equal(v_gpu, v_cpu)
        assert np.array_equal(co_gpu, co_cpu)
        assert np.array_equal(c_gpu, c_cpu)


def test_degenerate_samples_are_skipped(ctx, oracle):
    xyz = scenes.plane_outlier_cloud(5000, seed=31)
    xyz[10] = xyz[11]  # duplicate point -> collinear triple when both are drawn
    xyz[12, :3] = 2 * xyz[11, :3] - xyz[13, :3]
    cloud = ctx.stage(xyz)
    p = pkg.default_support_sac_params()
    samples = np.array([[10, 11, 40], [13, 11, 12], [1, 2, 3], [11, 10, 10]], np.int32)
    c_gpu, co_gpu, v_gpu = ctx.sac_score(cloud, p, samples)
    c_cpu, co_cpu, v_cpu = oracle.sac_score(xyz, None, p, samples)
    assert np.array_equal(v_gpu, v_cpu) and np.array_equal(c_gpu, c_cpu)
    np.testing.assert_array_equal(co_gpu, co_cpu)


def test_all_h_stop_rule_and_replay(ctx, oracle):
    xyz = scenes.plane_outlier_cloud(60000, seed=41)
    cloud = ctx.stage(xyz)
    samples = oracle.pcl_sample_stream(xyz, A.MODEL_PLANE, 1500)
    p = pkg.default_support_sac_params()
    p.sampler, p.stop, p.max_iterations = A.SAMPLER_REPLAY, A.STOP_ALL_H, 1500
    p.replay_samples = samples.ctypes.data_as(A.i32p)
    p.replay_count = 1500
    got = ctx.sac_segment(cloud, p)
    want = oracle.sac_segment(xyz, None, p)
    assert got["info"].best_hypothesis == want["info"].best_hypothesis
    assert got["info"].best_count == want["info"].best_count
    assert np.array_equal(got["inliers"], want["inliers"])
    assert np.array_equal(got["coeffs"], want["coeffs"])


def test_select_and_refine_entry_points(ctx, oracle):
    xyz = scenes.plane_outlier_cloud(25000, seed=51)
    cloud = ctx.stage(xyz)
    p = pkg.default_support_sac_params()
    co = np.array([0.01, -0.02, 0.9997, 0.003], np.float32)
    inl_g = ctx.sac_select(cloud, p, co)
    inl_c = oracle.sac_select(xyz, None, p, co)
    assert np.array_equal(inl_g, inl_c) and len(inl_g) > 1000
    ref_g, _ = ctx.sac_refine(cloud, p, co, inl_g)
    ref_c, _ = oracle.sac_refine(xyz, None, p, co, inl_c)
    assert np.array_equal(ref_g, ref_c)


def test_empty_and_tiny_clouds(ctx, oracle):
    p = pkg.default_support_sac_params()
    for n in (0, 1, 2):
        xyz = scenes.plane_outlier_cloud(max(n, 1), seed=3)[:n]
        cloud = ctx.stage(xyz)
        got = ctx.sac_segment(cloud, p)
        assert len(got["inliers"]) == 0 and len(got["coeffs"]) == 0


def test_philox_sampler_finds_the_plane(ctx):
    xyz = scenes.plane_outlier_cloud(50000, seed=61)
    cloud = ctx.stage(xyz)
    p = pkg.default_support_sac_params()
    p.sampler, p.stop, p.max_iterations = A.SAMPLER_PHILOX, A.STOP_ALL_H, 512
    got = ctx.sac_segment(cloud, p)
    assert abs(abs(got["coeffs"][2]) - 1.0) < 1e-3 and len(got["inliers"]) > 0.68 * 50000
    again = ctx.sac_segment(cloud, p)
    assert np.array_equal(got["inliers"], again["inliers"])  # counter-based: reproducible


def _filter_stats(ctx):
    out = (C.c_uint64 * 2)()
    ctx.lib.pitt_debug_plane_filter_stats(1, out)
    return int(out[0]), int(out[1])


def _score_with_stats(ctx, cloud, p, samples):
    ctx.lib.pitt_debug_plane_filter_stats(1, None)
    ctx.lib.pitt_debug_plane_mode(2)
    try:
        counts = ctx.sac_score(cloud, p, samples)[0]
        pairs, redo = _filter_stats(ctx)
    finally:
        ctx.lib.pitt_debug_plane_filter_stats(0, None)
        ctx.lib.pitt_debug_plane_mode(0)
    return counts, pairs, redo


@pytest.mark.parametrize("n", [512 * 40, 512 * 40 + 1, 512 * 40 + 511, 700])
def test_filter_kernel_ragged_tiles(ctx, oracle, n):
    """full tiles take the FFMA filter, the ragged last tile the exact warp path"""
    xyz = scenes.plane_outlier_cloud(n, seed=71)
    cloud = ctx.stage(xyz)
    p = pkg.default_support_sac_params()
    samples = np.random.default_rng(5).integers(0, n, (1100, 3)).astype(np.int32)
    c_gpu, pairs, redo = _score_with_stats(ctx, cloud, p, samples)
    c_cpu = oracle.sac_score(xyz, None, p, samples)[0]
    assert np.array_equal(c_gpu, c_cpu)
    assert pairs > 0  # the filter kernel did the work


def test_filter_kernel_points_on_the_threshold(ctx, oracle):
    """adversarial cloud: most points sit within a few ulps of |distance| == threshold of the hypotheses, so
    nearly every (hypothesis, tile) pair is uncertain and must be re-evaluated in the exact operation order"""
    rng = np.random.default_rng(9)
    n = 512 * 24
    thr = np.float32(0.02)
    xyz = np.ones((n, 4), np.float32)
    xyz[:, 0] = rng.uniform(-1, 1, n)
    xyz[:, 1] = rng.uniform(-1, 1, n)
    # heights at +-thr +- {0, 1, 2, 3} ulps, plus a few exact multiples
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
    c_gpu, pairs, redo = _score_with_stats(ctx, cloud, p, samples)
    c_cpu = oracle.sac_score(xyz, None, p, samples)[0]
    assert np.array_equal(c_gpu, c_cpu)
    assert redo >= n // 512  # at least every tile of hypothesis 0
    ctx.lib.pitt_debug_plane_mode(1)
    try:
        c_exact = ctx.sac_score(cloud, p, samples)[0]
    finally:
        ctx.lib.pitt_debug_plane_mode(0)
    assert np.array_equal(c_exact, c_cpu)


@pytest.mark.parametrize("scale,thr", [(1.0, 0.02), (40.0, 0.02), (1e-3, 1e-4), (1.0, 3.0), (1e4, 0.05)])
def test_filter_kernel_scales(ctx, oracle, scale, thr):
    """the filter's scale and band follow the cloud extent and the threshold; out-of-range cases fall back"""
    n = 512 * 30
    xyz = scenes.plane_outlier_cloud(n, seed=81)
    xyz[:, :3] *= np.float32(scale)
    cloud = ctx.stage(xyz)
    p = pkg.default_support_sac_params()
    p.distance_threshold = thr
    samples = np.random.default_rng(6).integers(0, n, (600, 3)).astype(np.int32)
    c_gpu, pairs, redo = _score_with_stats(ctx, cloud, p, samples)
    c_cpu = oracle.sac_score(xyz, None, p, samples)[0]
    assert np.array_equal(c_gpu, c_cpu)
    if pairs:
        assert redo < 0.25 * pairs  # the band stays narrow on ordinary clouds


def test_filter_kernel_non_finite_points_fall_back(ctx, oracle):
    n = 512 * 20
    xyz = scenes.plane_outlier_cloud(n, seed=91)
    xyz[100, 0] = np.nan
    xyz[2000, 2] = np.inf
    xyz[7000, :3] = np.nan
    cloud = ctx.stage(xyz)
    p = pkg.default_support_sac_params()
    samples = np.random.default_rng(7).integers(0, n, (512, 3)).astype(np.int32)
    samples[5] = [100, 3, 4]  # a hypothesis built on a NaN point
    c_gpu, pairs, redo = _score_with_stats(ctx, cloud, p, samples)
    c_cpu = oracle.sac_score(xyz, None, p, samples)[0]
    assert np.array_equal(c_gpu, c_cpu)
    assert pairs == 0  # exact kernel took over


def test_hypothesis_split_finish_matches_single_segment(ctx, oracle):
    """config 5 pattern on one GPU: two 'ranks' score halves of the sample stream, the counts are
    concatenated (the all-gather), earliest arg-max, pitt_sac_finish_device == one seg.segment() over all"""
    import torch
    xyz = scenes.plane_outlier_cloud(50000, seed=101)
    cloud = ctx.stage(xyz)
    H = 1200
    samples = oracle.pcl_sample_stream(xyz, A.MODEL_PLANE, 2 * H)
    p = pkg.default_support_sac_params()
    p.sampler, p.stop, p.max_iterations = A.SAMPLER_REPLAY, A.STOP_ALL_H, 2 * H
    p.replay_samples = samples.ctypes.data_as(A.i32p)
    p.replay_count = 2 * H
    want = oracle.sac_segment(xyz, None, p)
    dev = torch.device("cuda", 0)
    d_all_samples = torch.from_numpy(samples).to(dev)
    d_counts = torch.zeros(2 * H, dtype=torch.int32, device=dev)
    for r in range(2):
        part = d_all_samples[r * H:(r + 1) * H].contiguous()
        torch.cuda.synchronize()  # torch's stream and the ctx stream are independent
        ctx.sac_score_device(cloud, p, part.data_ptr(), H, d_counts[r * H:].data_ptr())
        ctx.synchronize()
    d_best = torch.zeros(2, dtype=torch.int32, device=dev)
    torch.cuda.synchronize()
    ctx.argmax_counts_device(d_counts.data_ptr(), 2 * H, d_best.data_ptr())
    got = ctx.sac_finish_device(cloud, p, d_all_samples.data_ptr(), 2 * H, d_best.data_ptr(), want_inliers=True)
    assert got["info"].best_hypothesis == want["info"].best_hypothesis
    assert got["info"].best_count == want["info"].best_count
    assert np.array_equal(got["inliers"], want["inliers"])
    assert np.array_equal(got["coeffs"], want["coeffs"])


def _same_result(a, b):
    assert np.array_equal(a["inliers"], b["inliers"])
    assert np.array_equal(a["coeffs"].view(np.uint32), b["coeffs"].view(np.uint32))
    ai, bi = a["info"], b["info"]
    assert (ai.iterations, ai.skipped, ai.best_hypothesis, ai.best_count, ai.n_inliers_model, ai.hypotheses) == \
           (bi.iterations, bi.skipped, bi.best_hypothesis, bi.best_count, bi.n_inliers_model, bi.hypotheses)


@pytest.fixture
def chunked(ctx, request):
    """force the chunked copy of pitt_sac_segment_host (by default only clouds of >= 16 M points are chunked)"""
    k = getattr(request, "param", 4)
    ctx.lib.pitt_debug_stream_chunks(k)
    yield k
    ctx.lib.pitt_debug_stream_chunks(0)


@pytest.mark.parametrize("chunked", [0, 1, 3, 4], indirect=True)
@pytest.mark.parametrize("n", [1 << 18, (1 << 18) + 777, 300000])
def test_segment_host_streams_the_cloud_and_matches_pcl(ctx, oracle, n, chunked):
    """pitt_sac_segment_host (fromROSMsg + seg.segment() fused; the H2D copy in chunks under the scoring) against the staged
    call and the oracle, with the reference's own parameters (mt19937 sampler, adaptive stop, 10 iterations)"""
    xyz = scenes.plane_outlier_cloud(n, seed=31)
    p = pkg.default_support_sac_params()
    got = ctx.sac_segment_host(xyz, p)
    cloud = ctx.stage(xyz)
    staged = ctx.sac_segment(cloud, p)
    _same_result(got, staged)
    want = oracle.sac_segment(xyz, None, oracle.default_support_sac_params())
    assert np.array_equal(got["inliers"], want["inliers"])
    assert np.array_equal(got["coeffs"], want["coeffs"])
    assert len(got["inliers"]) > n // 2


@pytest.mark.parametrize("chunked", [0, 4], indirect=True)
@pytest.mark.parametrize("plane_mode,H", [(0, 700), (3, 700), (1, 300), (0, 40)])
def test_segment_host_all_hypotheses_replay(ctx, oracle, plane_mode, H, chunked):
    """ALL_H over a replayed sample table: exact, tensor-core (forced per chunk) and generic scoring kernels on chunks"""
    n = 280000
    xyz = scenes.plane_outlier_cloud(n, seed=32, plane_frac=0.5)
    samples = oracle.pcl_sample_stream(xyz, A.MODEL_PLANE, H)
    p = pkg.default_support_sac_params()
    p.stop, p.max_iterations, p.sampler = A.STOP_ALL_H, H, A.SAMPLER_REPLAY
    keep = np.ascontiguousarray(samples)
    p.replay_samples = keep.ctypes.data_as(A.i32p)
    p.replay_count = H
    ctx.lib.pitt_debug_plane_mode(plane_mode)
    try:
        got = ctx.sac_segment_host(xyz, p)
    finally:
        ctx.lib.pitt_debug_plane_mode(0)
    cloud = ctx.stage(xyz)
    staged = ctx.sac_segment(cloud, p)
    _same_result(got, staged)
    c_cpu, _, _ = oracle.sac_score(xyz, None, p, samples)
    assert got["info"].best_hypothesis == int(np.argmax(c_cpu)) and got["info"].best_count == int(c_cpu.max())


@pytest.mark.parametrize("chunked", [0, 4], indirect=True)
def test_segment_host_on_an_organised_frame(ctx, oracle, chunked):
    """a Kinect-ordered frame: the four chunks are image stripes with different extents (the tensor path derives its scale
    per chunk from the chunk and the sample points)"""
    xyz = scenes.tabletop_frame(seed=3)
    p = pkg.default_support_sac_params()
    p.stop, p.max_iterations, p.sampler = A.STOP_ALL_H, 600, A.SAMPLER_PCL_MT19937
    ctx.lib.pitt_debug_plane_mode(3)
    try:
        got = ctx.sac_segment_host(xyz, p)
    finally:
        ctx.lib.pitt_debug_plane_mode(0)
    cloud = ctx.stage(xyz)
    ctx.lib.pitt_debug_plane_mode(1)
    try:
        staged = ctx.sac_segment(cloud, p)
    finally:
        ctx.lib.pitt_debug_plane_mode(0)
    _same_result(got, staged)


@pytest.mark.parametrize("n,cols", [(5000, 4), (270000, 3), (0, 4), (2, 4)])
def test_segment_host_falls_back_to_the_staged_sequence(ctx, oracle, n, cols):
    """small clouds, point_step 12 and empty inputs take stage + segment + release"""
    xyz = scenes.plane_outlier_cloud(max(n, 1), seed=33)[:n]
    p = pkg.default_support_sac_params()
    got = ctx.sac_segment_host(xyz[:, :cols], p)
    want = oracle.sac_segment(xyz, None, oracle.default_support_sac_params())
    assert np.array_equal(got["inliers"], want["inliers"])
    assert np.array_equal(got["coeffs"], want["coeffs"])


@pytest.mark.parametrize("spoil", ["far_points", "nan_points", "inf_points"])
def test_segment_host_single_launch_bounds_check(ctx, oracle, spoil):
    """single-launch streaming derives the tensor path's scale from the sample points (doubled) and verifies afterwards that
    no point of the cloud exceeded it: points far outside, NaN and inf coordinates that are not among the samples must send
    the job to the exact kernel and change nothing in the result"""
    n, H = 300000, 700
    xyz = scenes.plane_outlier_cloud(n, seed=35)
    rng = np.random.default_rng(4)
    samples = rng.integers(0, n - 100, (H, 3)).astype(np.int32)  # the last 100 points are never sampled
    samples[:, 1] = (samples[:, 0] + 1 + rng.integers(0, n - 200, H)) % (n - 100)
    samples[:, 2] = (samples[:, 1] + 7) % (n - 100)
    if spoil == "far_points":
        xyz[-50:, :3] *= np.float32(500.0)
    elif spoil == "nan_points":
        xyz[-50:, 1] = np.nan
    else:
        xyz[-50:, 2] = np.inf
    p = pkg.default_support_sac_params()
    p.stop, p.max_iterations, p.sampler = A.STOP_ALL_H, H, A.SAMPLER_REPLAY
    keep = np.ascontiguousarray(samples)
    p.replay_samples = keep.ctypes.data_as(A.i32p)
    p.replay_count = H
    got = ctx.sac_segment_host(xyz, p)
    cloud = ctx.stage(xyz)
    ctx.lib.pitt_debug_plane_mode(1)
    try:
        staged = ctx.sac_segment(cloud, p)
    finally:
        ctx.lib.pitt_debug_plane_mode(0)
    _same_result(got, staged)
    c_cpu, _, _ = oracle.sac_score(xyz, None, p, samples[:40])
    c_gpu, _, _ = ctx.sac_score(cloud, p, samples[:40])
    assert np.array_equal(c_cpu, c_gpu)


def test_philox_device_known_answers_and_sample_sets(ctx):
    """the PHILOX sampler's generator on the device against the Random123 known-answer vectors (Philox-4x32-10), and the sample
    sets it draws against the Python restatement of the derivation (csrc/sac.cu::philox_samples_kernel)"""
    import ctypes as C
    from pitt_object_table_segmentation_b200 import philox
    u32x4, u32x2 = C.c_uint32 * 4, C.c_uint32 * 2
    for ctr, key, want in philox.KAT:
        out = u32x4()
        assert ctx.lib.pitt_debug_philox(ctx.handle, u32x4(*ctr), u32x2(*key), out) == 0
        assert tuple(out) == want
    for S, n, stream in ((3, 1000, 1), (4, 37, 2), (2, 5, 7)):
        H = 500
        got = np.zeros((H, S), np.int32)
        assert ctx.lib.pitt_debug_philox_samples(ctx.handle, H, S, n, stream, got.ctypes.data_as(A.i32p)) == 0
        assert np.array_equal(got, philox.sample_sets(H, S, n, seed=12345, stream_id=stream))
