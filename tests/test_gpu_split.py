"""GPU tests of the multi-GPU boundary: pitt_sac_segment_split (hypothesis split behind the C ABI) and contexts on
two devices inside one process."""
import ctypes as C

import numpy as np
import pytest

import pitt_object_table_segmentation_b200 as pkg
from pitt_object_table_segmentation_b200 import _abi as A, scenes
from pitt_object_table_segmentation_b200.api import ALLGATHER_FN, _DevArray

pytestmark = pytest.mark.gpu


def _all_h_params(samples):
    p = pkg.default_support_sac_params()
    p.sampler, p.stop, p.max_iterations = A.SAMPLER_REPLAY, A.STOP_ALL_H, len(samples)
    p.replay_samples = samples.ctypes.data_as(A.i32p)
    p.replay_count = len(samples)
    return p


def test_split_world_1_equals_segment(ctx, oracle):
    xyz = scenes.plane_outlier_cloud(60_000, seed=21)
    cloud = ctx.stage(xyz)
    samples = np.ascontiguousarray(ctx.pcl_sample_stream(cloud, A.MODEL_PLANE, 700))
    p = _all_h_params(samples)
    want = ctx.sac_segment(cloud, p)
    got = ctx.sac_segment_split(cloud, p, 0, 1, None)
    assert np.array_equal(got["inliers"], want["inliers"])
    assert np.array_equal(got["coeffs"].view(np.uint32), want["coeffs"].view(np.uint32))
    assert got["info"].best_hypothesis == want["info"].best_hypothesis and got["info"].best_count == want["info"].best_count
    ref = oracle.sac_segment(xyz, None, p)
    assert np.array_equal(got["inliers"], ref["inliers"])


@pytest.mark.parametrize("world,H", [(2, 700), (3, 1000), (4, 257)])
def test_split_emulated_ranks_agree_with_single_gpu(ctx, world, H):
    """`world` ranks emulated one after the other on one GPU: a first pass captures every rank's slice of counts (what the
    all-gather would carry), a second pass hands every rank the gathered buffer. Every rank must return the single-GPU result;
    H not divisible by world exercises the padding of the last slice."""
    import torch
    xyz = scenes.plane_outlier_cloud(50_000, seed=33)
    cloud = ctx.stage(xyz)
    samples = np.ascontiguousarray(ctx.pcl_sample_stream(cloud, A.MODEL_PLANE, H))
    samples[5] = samples[5][[0, 0, 1]]  # a degenerate triple: computeModelCoefficients fails, PCL skips it, it must never win
    p = _all_h_params(samples)
    want = ctx.sac_segment(cloud, p)
    H_loc = (H + world - 1) // world
    sent = {}

    def make_cb(rank, gathered):
        def cb(user, d_send, d_recv, count, stream):
            assert count == H_loc
            # the collective must be enqueued on (or ordered after) the context's stream: the counts are produced there
            with torch.cuda.stream(torch.cuda.ExternalStream(stream)):
                send = torch.as_tensor(_DevArray(d_send, count), device="cuda")
                recv = torch.as_tensor(_DevArray(d_recv, count * world), device="cuda")
                sent[rank] = send.clone()
                if gathered is not None:
                    recv.copy_(gathered)
                else:
                    recv.fill_(-1)
            return 0
        return ALLGATHER_FN(cb)

    for r in range(world):  # pass A: capture
        ctx.sac_segment_split(cloud, p, r, world, make_cb(r, None), want_inliers=False)
    gathered = torch.cat([sent[r] for r in range(world)])
    counts = ctx.sac_score(cloud, p, samples)[0]
    g = gathered.cpu().numpy()
    assert np.array_equal(g[:H][g[:H] >= 0], counts[g[:H] >= 0]) and g[5] == -1 and np.all(g[H:] == -1)
    for r in range(world):  # pass B: every rank finishes from the gathered counts
        got = ctx.sac_segment_split(cloud, p, r, world, make_cb(r, gathered))
        assert np.array_equal(got["inliers"], want["inliers"])
        assert np.array_equal(got["coeffs"].view(np.uint32), want["coeffs"].view(np.uint32))
        assert got["info"].best_hypothesis == want["info"].best_hypothesis and got["info"].best_count == want["info"].best_count


def test_split_rejects_the_adaptive_stop(ctx):
    xyz = scenes.plane_outlier_cloud(5_000, seed=1)
    cloud = ctx.stage(xyz)
    p = pkg.default_support_sac_params()
    with pytest.raises(pkg.PittError):
        ctx.sac_segment_split(cloud, p, 0, 1, None)


def test_contexts_on_two_devices_in_one_process(oracle):
    """pitt_segment_frames_batched with contexts on device 0 and device 1 (the opt-in to > 48 KB of dynamic shared memory is per
    device: plane_tc_kernel, lm_kernel, knn kernels). Skipped on a single-GPU box."""
    lib = pkg.load_library()
    if lib.pitt_device_count() < 2:
        pytest.skip("needs two GPUs in one process")
    frames = [scenes.tabletop_frame(seed=s, width=320, height=240, random_poses=True) for s in range(4)]
    ctxs = [pkg.Context(0, seed=12345), pkg.Context(1, seed=12345)]
    try:
        got = pkg.segment_frames_batched(ctxs, frames)
        # the tensor path on both devices (>= 2^27 evaluations), same process
        xyz = scenes.plane_outlier_cloud(70_000, seed=7)
        res = []
        for c in ctxs:
            cloud = c.stage(xyz)
            samples = np.ascontiguousarray(c.pcl_sample_stream(cloud, A.MODEL_PLANE, 2048))
            res.append(c.sac_score(cloud, pkg.default_support_sac_params(), samples)[0])
        assert np.array_equal(res[0], res[1])
    finally:
        for c in ctxs:
            c.close()
    fp = oracle.default_frame_params()
    for f, g in zip(frames, got):
        w = oracle.segment_frame(f, fp)
        assert [s["tag"] for s in g["shapes"]] == [s["tag"] for s in w["shapes"]]
        assert [s["inliers"] for s in g["shapes"]] == [s["inliers"] for s in w["shapes"]]
