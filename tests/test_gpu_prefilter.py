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


def _arm_params(rng, n_boxes=4, dense=0, rotate=True, translate=True, centre=(0.0, 0.0, 1.0)):
    p = pkg.default_arm_filter_params()
    p.n_boxes = n_boxes
    p.input_is_dense = dense
    for k in range(4):
        for a in range(3):
            p.box[k].translation[a] = float(centre[a] + rng.uniform(-0.3, 0.3)) if translate else 0.0
            p.box[k].rotation_rpy[a] = float(rng.uniform(-3.0, 3.0)) if rotate else 0.0
    return p


@pytest.mark.parametrize("n_boxes,dense,rotate,translate", [(4, 0, True, True), (4, 1, True, True), (1, 0, False, True),
                                                            (2, 0, True, False), (3, 1, False, False), (0, 0, True, True)])
def test_arm_filter_matches_oracle(ctx, oracle, n_boxes, dense, rotate, translate):
    """chained negative CropBoxes of the arm filter service (arm_filter_srv.cpp:66-103, 134-141)"""
    rng = np.random.default_rng(100 + n_boxes * 8 + dense * 4 + rotate * 2 + translate)
    raw = scenes.raw_camera_frame(seed=9, width=320, height=240, point_step=16)  # camera frame, NaN returns included
    xyz = np.ones((len(raw), 4), np.float32)
    xyz[:, :3] = raw[:, :3]
    xyz[::97, 0] = np.inf  # a few infinite coordinates as well
    p = _arm_params(rng, n_boxes, dense, rotate, translate)
    # make the boxes big enough to bite into the scene
    for k in range(4):
        for a in range(3):
            p.box[k].min_pt[a] *= 3.0
            p.box[k].max_pt[a] *= 3.0
    cloud = ctx.stage(xyz)
    out, removed = ctx.arm_filter(cloud, p)
    got = ctx.get_points(out)
    want, wremoved = oracle.arm_filter(xyz, p)
    assert removed == wremoved
    assert got.shape == want.shape
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    if n_boxes == 0:
        assert len(got) == len(xyz)
    else:
        assert 0 < len(got) < len(xyz) and sum(removed) == len(xyz) - len(got)


def test_arm_filter_points_on_the_box_faces(ctx, oracle):
    """points exactly on a face are inside (the comparisons are strict), one ulp further they are outside"""
    rng = np.random.default_rng(7)
    p = pkg.default_arm_filter_params()
    p.n_boxes = 1
    p.input_is_dense = 1
    lo = np.array(p.box[0].min_pt[:], np.float32)
    hi = np.array(p.box[0].max_pt[:], np.float32)
    pts = rng.uniform(lo, hi, (6000, 3)).astype(np.float32)
    face = rng.integers(0, 6, len(pts))
    for f in range(6):
        sel = face == f
        a = f % 3
        v = lo[a] if f < 3 else hi[a]
        pts[sel, a] = v
    nudged = pts.copy()
    for f in range(6):
        sel = (face == f) & (rng.random(len(pts)) < 0.5)
        a = f % 3
        nudged[sel, a] = np.nextafter(pts[sel, a], np.float32(-np.inf if f < 3 else np.inf))
    xyz = np.ones((len(pts), 4), np.float32)
    xyz[:, :3] = nudged
    cloud = ctx.stage(xyz)
    out, removed = ctx.arm_filter(cloud, p)
    got = ctx.get_points(out)
    want, wremoved = oracle.arm_filter(xyz, p)
    assert removed == wremoved and np.array_equal(got.view(np.uint32), want.view(np.uint32))
    assert 0.3 * len(pts) < len(got) < 0.7 * len(pts)


def test_arm_filter_empty_cloud(ctx, oracle):
    p = pkg.default_arm_filter_params()
    cloud = ctx.stage(np.zeros((0, 4), np.float32))
    out, removed = ctx.arm_filter(cloud, p)
    assert out.n == 0 and removed == [0, 0, 0, 0]
