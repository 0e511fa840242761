"""GPU parity: plane RANSAC (supports_segmentation_srv.cpp:89-111, plane_segmentation_srv.cpp:52-67)
through the C ABI vs the CPU oracle, bit-exact."""
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
    for force_generic in (0, 1):
        ctx.lib.pitt_debug_force_generic_plane(force_generic)
        c_gpu, co_gpu, v_gpu = ctx.sac_score(cloud, p, samples)
        ctx.lib.pitt_debug_force_generic_plane(0)
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
