"""GPU parity of the service-shaped entry points (findSupports, clusterize, ransac*Detection,
selection rule, whole frame) through the C ABI vs the CPU oracle."""
import numpy as np
import pytest

import pitt_object_table_segmentation_b200 as pkg
from pitt_object_table_segmentation_b200 import _abi as A, scenes

pytestmark = pytest.mark.gpu


def _eq_f(a, b):
    return np.array_equal(np.asarray(a, np.float32).view(np.uint32), np.asarray(b, np.float32).view(np.uint32))


def _check_supports(got, want):
    assert got["loop_trips"] == want["loop_trips"]
    assert got["used"] == want["used"]
    assert len(got["supports"]) == len(want["supports"])
    for g, w in zip(got["supports"], want["supports"]):
        assert _eq_f(g["coefficients"], w["coefficients"])
        assert np.array_equal(g["inliers"], w["inliers"])
        assert _eq_f(g["support_cloud"], w["support_cloud"])
        assert _eq_f(g["on_support_cloud"], w["on_support_cloud"])


@pytest.mark.parametrize("seed,w,h", [(1, 160, 120), (12345, 320, 240)])
def test_find_supports_tabletop(ctx, oracle, seed, w, h):
    xyz = scenes.tabletop_frame(seed=seed, width=w, height=h)
    nrm = oracle.estimate_normals(xyz, 50)
    cloud = ctx.stage(xyz, normals=nrm)
    got = ctx.find_supports(cloud)
    want = oracle.find_supports(xyz, nrm, oracle.default_support_params())
    assert len(want["supports"]) == 1
    _check_supports(got, want)


def _two_level_scene(seed, tilt_wall=True):
    """table + a raised shelf (second horizontal support) + a vertical wall (non horizontal plane)"""
    r = np.random.default_rng(seed)
    table = np.stack([r.uniform(-0.6, 0.6, 9000), r.uniform(-0.4, 0.4, 9000), r.normal(0, 0.002, 9000)], 1)
    shelf = np.stack([r.uniform(-0.3, 0.3, 3000), r.uniform(0.45, 0.8, 3000), 0.3 + r.normal(0, 0.002, 3000)], 1)
    wall = np.stack([r.uniform(-0.6, 0.6, 5000), 0.9 + r.normal(0, 0.002, 5000), r.uniform(0, 0.8, 5000)], 1)
    box = np.stack([r.uniform(-0.05, 0.05, 800), r.uniform(-0.05, 0.05, 800), r.uniform(0.03, 0.12, 800)], 1)
    ball = r.normal(size=(600, 3))
    ball = ball / np.linalg.norm(ball, axis=1, keepdims=True) * 0.05 + np.array([0.0, 0.6, 0.36])
    pts = np.concatenate([table, shelf, wall, box, ball])
    pts = pts[r.permutation(len(pts))]
    xyz = np.ones((len(pts), 4), np.float32)
    xyz[:, :3] = pts
    return xyz


@pytest.mark.parametrize("seed", [3, 4])
def test_find_supports_multi_level_with_wall(ctx, oracle, seed):
    """several loop trips, a non-horizontal plane (idx map level -1) and two supports"""
    xyz = _two_level_scene(seed)
    cloud = ctx.stage(xyz)
    got = ctx.find_supports(cloud)
    want = oracle.find_supports(xyz, None, oracle.default_support_params())
    assert want["loop_trips"] >= 3 and len(want["supports"]) >= 2
    _check_supports(got, want)


def test_find_supports_custom_params_and_sorted_input(ctx, oracle):
    xyz = _two_level_scene(9)
    xyz = xyz[np.argsort(xyz[:, 0], kind="stable")]  # ascending x: every point raises the running max (C.7)
    cloud = ctx.stage(xyz)
    p = pkg.default_support_params()
    p.ransac_max_iteration_threshold = 30
    p.ransac_distance_point_in_shape_threshold = 0.01
    p.horizontal_axis_len = 3
    p.horizontal_axis[0], p.horizontal_axis[1], p.horizontal_axis[2] = 0.0, 0.0, 1.0
    p.support_edge_remove_offset_len = 3
    p.support_edge_remove_offset[0], p.support_edge_remove_offset[1], p.support_edge_remove_offset[2] = 0.05, 0.03, 0.01
    got = ctx.find_supports(cloud, p)
    want = oracle.find_supports(xyz, None, p)
    _check_supports(got, want)


def test_find_supports_no_plane(ctx, oracle):
    r = np.random.default_rng(0)
    xyz = np.ones((2000, 4), np.float32)
    xyz[:, :3] = r.uniform(-1, 1, (2000, 3))
    cloud = ctx.stage(xyz)
    got = ctx.find_supports(cloud)
    want = oracle.find_supports(xyz, None, oracle.default_support_params())
    _check_supports(got, want)


def _objects(seed, w=320, h=240):
    xyz = scenes.tabletop_frame(seed=seed, width=w, height=h)
    return np.ascontiguousarray(xyz[xyz[:, 2] > 0.012])


def test_cluster_service(ctx, oracle):
    xyz = _objects(5)
    cloud = ctx.stage(xyz)
    got = ctx.cluster_service(cloud)
    want = oracle.cluster_service(xyz, oracle.default_cluster_params())
    assert len(got) == len(want) == 3
    for g, w in zip(got, want):
        assert np.array_equal(g["inliers"], w["inliers"])
        assert _eq_f(g["centroid"], w["centroid"])
        # the n+1 divisor quirk
        true_mean = xyz[g["inliers"], :3].mean(0)
        assert np.allclose(g["centroid"], true_mean * len(g["inliers"]) / (len(g["inliers"]) + 1), atol=1e-5)


def test_cluster_service_small_input_is_skipped(ctx, oracle):
    xyz = _objects(5)[:29]
    cloud = ctx.stage(xyz)
    assert ctx.cluster_service(cloud) == [] == oracle.cluster_service(xyz, oracle.default_cluster_params())


@pytest.mark.parametrize("model,kind", [(A.MODEL_PLANE, "plane"), (A.MODEL_SPHERE, "sphere"),
                                        (A.MODEL_CYLINDER, "cylinder"), (A.MODEL_CONE, "cone"),
                                        (A.MODEL_CYLINDER, "sphere"), (A.MODEL_CONE, "plane")])
def test_primitive_service(ctx, oracle, model, kind):
    xyz, _ = scenes.primitive_cluster(kind, 3000, 40 + model)
    nrm = oracle.estimate_normals(xyz, 50)
    cloud = ctx.stage(xyz, normals=nrm)
    p = pkg.default_sac_params(model)
    got = ctx.primitive_service(cloud, p)
    want = oracle.primitive_service(xyz, nrm, p)
    assert np.array_equal(got["inliers"], want["inliers"])
    assert 0 not in got["inliers"]  # inlierToVectorMsg drops index 0
    assert _eq_f(got["coefficients"], want["coefficients"])
    assert got["centroid_valid"] == want["centroid_valid"]
    if got["centroid_valid"]:
        assert _eq_f(got["centroid"], want["centroid"])
    if model in (A.MODEL_CYLINDER, A.MODEL_CONE):
        assert len(got["coefficients"]) == 8  # height appended


def test_primitive_service_no_model_found(ctx, oracle):
    # 2 points: no cylinder can be sampled... 1 point cloud -> getSamples fails
    xyz = np.ones((1, 4), np.float32)
    nrm = np.zeros((1, 4), np.float32)
    cloud = ctx.stage(xyz, normals=nrm)
    p = pkg.default_sac_params(A.MODEL_CYLINDER)
    got = ctx.primitive_service(cloud, p)
    want = oracle.primitive_service(xyz, nrm, p)
    assert len(got["inliers"]) == 0 and _eq_f(got["coefficients"], want["coefficients"])
    assert list(got["coefficients"]) == [-1.0]


def test_selection_rule_matches(oracle):
    rng = np.random.default_rng(1)
    cases = [(0, 0, 0, 0), (10, 0, 0, 0), (0, 10, 0, 0), (0, 0, 10, 0), (0, 0, 0, 10), (5, 5, 5, 5),
             (100, 90, 100, 90), (0, 0, 100, 90), (0, 0, 100, 89), (0, 0, 1000, 900), (3, 7, 7, 6)]
    cases += [tuple(int(v) for v in rng.integers(0, 50, 4)) for _ in range(300)]
    for pl, sp, cy, co in cases:
        assert pkg.select_primitive(pl, sp, cy, co) == oracle.select_primitive(pl, sp, cy, co), (pl, sp, cy, co)
    assert pkg.select_primitive(0, 0, 100, 90) == A.TAG_CONE      # cone over cylinder priority 0.9
    assert pkg.select_primitive(0, 0, 100, 89) == A.TAG_CYLINDER


@pytest.mark.parametrize("seed,w,h", [(12345, 320, 240), (7, 240, 180)])
def test_segment_frame(ctx, oracle, seed, w, h):
    xyz = scenes.tabletop_frame(seed=seed, width=w, height=h)
    cloud = ctx.stage(xyz)
    got = ctx.segment_frame(cloud)
    want = oracle.segment_frame(xyz, oracle.default_frame_params())
    assert (got["n_supports"], got["n_clusters"]) == (want["n_supports"], want["n_clusters"])
    assert got["support_sizes"] == want["support_sizes"] and got["on_support_sizes"] == want["on_support_sizes"]
    assert _eq_f(got["support_coefficients"], want["support_coefficients"])
    assert len(got["shapes"]) == len(want["shapes"]) == 3
    for g, wv in zip(got["shapes"], want["shapes"]):
        assert (g["tag"], g["n_points"], g["inliers"], g["object_id"]) == (wv["tag"], wv["n_points"], wv["inliers"], wv["object_id"])
        assert _eq_f(g["coefficients"], wv["coefficients"])
        assert _eq_f(g["pc_centroid"], wv["pc_centroid"])
        assert _eq_f(g["est_centroid"], wv["est_centroid"])
    if w >= 320:  # enough points per object for the reference's selection rule to name all three
        assert sorted(s["tag_name"] for s in got["shapes"]) == ["cone", "cylinder", "sphere"]


def test_segment_frames_batched_matches_single(ctx, oracle):
    frames = [scenes.tabletop_frame(seed=s, width=200, height=150, random_poses=True) for s in range(6)]
    ctxs = [pkg.Context(0, seed=12345) for _ in range(3)]
    got = pkg.segment_frames_batched(ctxs, frames)
    for f, g in zip(frames, got):
        single = ctx.segment_frame(ctx.stage(f))
        assert g["n_clusters"] == single["n_clusters"] and g["support_sizes"] == single["support_sizes"]
        for a, b in zip(g["shapes"], single["shapes"]):
            assert (a["tag"], a["inliers"]) == (b["tag"], b["inliers"])
            assert _eq_f(a["coefficients"], b["coefficients"])
    want = oracle.segment_frame(frames[0], oracle.default_frame_params())
    assert [s["tag"] for s in got[0]["shapes"]] == [s["tag"] for s in want["shapes"]]
    for c in ctxs:
        c.close()


def test_segment_frames_batched_ragged_empty_and_short_streams(oracle):
    """the double-buffered frame stream (next frame's copy under the current frame's kernels) on a ragged stream: frames of
    different sizes, an empty frame, a frame below the 30-point guard of depthAcquisition (obj_segmentation.cpp:251), more
    contexts than frames, and no frames at all"""
    big = scenes.tabletop_frame(seed=21, width=200, height=150, random_poses=True)
    small = scenes.tabletop_frame(seed=22, width=120, height=90, random_poses=True)
    frames = [big, np.zeros((0, 4), np.float32), small, big[:25].copy(), small[::2].copy(), big]
    fp = oracle.default_frame_params()
    want = [oracle.segment_frame(f, fp) if len(f) else None for f in frames]
    for n_ctx in (1, 2, 8):
        ctxs = [pkg.Context(0, seed=12345) for _ in range(n_ctx)]
        try:
            for _ in range(2):  # second pass: buffers come back from the contexts' pools
                got = pkg.segment_frames_batched(ctxs, frames)
                assert len(got) == len(frames)
                for f, g, w in zip(frames, got, want):
                    if len(f) <= 30:
                        assert g["n_supports"] == 0 and g["n_clusters"] == 0 and g["shapes"] == []
                        continue
                    assert (g["n_supports"], g["n_clusters"], g["support_sizes"], g["on_support_sizes"]) == \
                           (w["n_supports"], w["n_clusters"], w["support_sizes"], w["on_support_sizes"])
                    assert _eq_f(g["support_coefficients"], w["support_coefficients"])
                    for a, b in zip(g["shapes"], w["shapes"]):
                        assert (a["tag"], a["n_points"], a["inliers"]) == (b["tag"], b["n_points"], b["inliers"])
                        assert _eq_f(a["coefficients"], b["coefficients"]) and _eq_f(a["pc_centroid"], b["pc_centroid"])
            assert len(pkg.segment_frames_batched(ctxs, [])) == 0
        finally:
            for c in ctxs:
                c.close()


@pytest.mark.parametrize("workers", [0, 1, 3, 12])
def test_segment_frame_is_independent_of_worker_count(oracle, workers):
    """the primitive fits of a frame fan out to helper streams/threads; results must not depend on that"""
    xyz = scenes.tabletop_frame(seed=99, width=260, height=200, random_poses=True)
    c = pkg.Context(0, seed=12345)
    try:
        c.set_workers(workers)
        got = c.segment_frame(c.stage(xyz))
        again = c.segment_frame(c.stage(xyz))  # helpers are reused across frames
    finally:
        c.close()
    want = oracle.segment_frame(xyz, oracle.default_frame_params())
    for res in (got, again):
        assert (res["n_supports"], res["n_clusters"]) == (want["n_supports"], want["n_clusters"])
        assert len(res["shapes"]) == len(want["shapes"])
        for g, wv in zip(res["shapes"], want["shapes"]):
            assert (g["tag"], g["n_points"], g["inliers"], g["object_id"]) == (wv["tag"], wv["n_points"], wv["inliers"], wv["object_id"])
            assert _eq_f(g["coefficients"], wv["coefficients"])
            assert _eq_f(g["est_centroid"], wv["est_centroid"])


@pytest.mark.parametrize("seed,w,h", [(21, 320, 240), (22, 400, 300)])
def test_segment_frame_fit_paths_agree(ctx, oracle, seed, w, h):
    """the primitive fits of a frame run batched per model (every launch serves all clusters, device-side PCL stop rule, one
    result copy); one asynchronous chain per fit (mode 2) and the round-1 path with one synchronous seg.segment() per fit
    (mode 1) must give the same TrackedShapes, and all equal the oracle's"""
    xyz = scenes.tabletop_frame(seed=seed, width=w, height=h, random_poses=True)
    cloud = ctx.stage(xyz)
    results = [ctx.segment_frame(cloud)]
    for mode in (1, 2):
        ctx.lib.pitt_debug_frame_mode(mode)
        try:
            results.append(ctx.segment_frame(cloud))
        finally:
            ctx.lib.pitt_debug_frame_mode(0)
    want = oracle.segment_frame(xyz, oracle.default_frame_params())
    for res in results:
        assert (res["n_supports"], res["n_clusters"]) == (want["n_supports"], want["n_clusters"])
        assert len(res["shapes"]) == len(want["shapes"]) == 3
        for g, wv in zip(res["shapes"], want["shapes"]):
            assert (g["tag"], g["n_points"], g["inliers"], g["object_id"]) == (wv["tag"], wv["n_points"], wv["inliers"], wv["object_id"])
            assert _eq_f(g["coefficients"], wv["coefficients"])
            assert _eq_f(g["pc_centroid"], wv["pc_centroid"])
            assert _eq_f(g["est_centroid"], wv["est_centroid"])
