"""GPU parity of the pre-path: fromROSMsg + VoxelGrid (pc_manager.cpp:55-67) + deep
filter (deep_filter_srv.cpp:27-58) + transformPointCloud (obj_segmentation.cpp:248) vs the CPU oracle, bit-exact."""
import numpy as np
import pytest

import pitt_object_table_segmentation_b200 as pkg
from pitt_object_table_segmentation_b200 import scenes

pytestmark = pytest.mark.gpu


def _params(leaf=0.01, deep=True, thr=-1.0, transform=True):
    p = pkg.default_prefilter_params()
    for a in range(3):
        p.leaf[a] = leaf
    p.apply_deep_filter = 1 if deep else 0
    p.deep_threshold = thr
    p.apply_transform = 1 if transform else 0
    c2w, _ = scenes.camera_pose()
    for i, v in enumerate(c2w.ravel()):
        p.transform[i] = float(v)
    return p


@pytest.mark.parametrize("w,h,step", [(320, 240, 16), (200, 150, 32), (640, 480, 16), (64, 48, 12)])
def test_prefilter_matches_oracle(ctx, oracle, w, h, step):
    raw = scenes.raw_camera_frame(seed=5, width=w, height=h, point_step=step)
    p = _params()
    cloud, info = ctx.prefilter(raw, p)
    got = ctx.get_points(cloud)
    want, winfo = oracle.prefilter(raw, p)
    assert info == winfo
    assert info["n_voxel"] < info["n_input"] and info["n_further"] > 0 and info["n_closer"] == len(want)
    assert got.shape == want.shape
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


@pytest.mark.parametrize("leaf,deep,thr,transform", [(0.0, True, 2.0, True), (0.02, False, -1.0, False), (0.005, True, 0.5, False),
                                                     (0.3, True, -1.0, True), (1e-5, True, -1.0, True)])
def test_prefilter_stage_switches_and_limits(ctx, oracle, leaf, deep, thr, transform):
    """each stage can be skipped; a huge leaf puts thousands of points in one voxel; a leaf that is too small
    for the extent makes PCL pass the input through (voxel_overflow)"""
    raw = scenes.raw_camera_frame(seed=9, width=160, height=120)
    p = _params(leaf, deep, thr, transform)
    cloud, info = ctx.prefilter(raw, p)
    got = ctx.get_points(cloud)
    want, winfo = oracle.prefilter(raw, p)
    assert info == winfo
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    if leaf == 1e-5:
        assert info["voxel_overflow"] == 1


def test_prefilter_empty_and_all_nan(ctx, oracle):
    p = _params()
    for raw in (np.zeros((0, 4), np.float32), np.full((100, 4), np.nan, np.float32)):
        cloud, info = ctx.prefilter(raw, p)
        want, winfo = oracle.prefilter(raw, p)
        assert cloud.n == 0 and len(want) == 0 and info == winfo


def test_raw_frame_stream_matches_manual_pipeline(ctx, oracle):
    """pitt_segment_raw_frames_batched = prefilter + segment_frame per frame; equals the oracle run on the
    oracle's pre-filtered cloud (the faithful C1 variant: 1 cm VoxelGrid first)"""
    raws = [scenes.raw_camera_frame(seed=s, width=320, height=240, random_poses=True) for s in range(3)]
    p = _params()
    ctxs = [pkg.Context(0, seed=12345) for _ in range(2)]
    try:
        got = pkg.segment_frames_batched(ctxs, raws, prefilter=p)
    finally:
        for c in ctxs:
            c.close()
    for raw, g in zip(raws, got):
        world, _ = oracle.prefilter(raw, p)
        want = oracle.segment_frame(world, oracle.default_frame_params())
        assert (g["n_supports"], g["n_clusters"]) == (want["n_supports"], want["n_clusters"])
        assert g["support_sizes"] == want["support_sizes"]
        for a, b in zip(g["shapes"], want["shapes"]):
            assert (a["tag"], a["n_points"], a["inliers"]) == (b["tag"], b["n_points"], b["inliers"])
            assert np.array_equal(np.asarray(a["coefficients"], np.float32).view(np.uint32),
                                  np.asarray(b["coefficients"], np.float32).view(np.uint32))
