"""GPU parity: sphere / cylinder / cone RANSAC (sphere_segmentation_srv.cpp:58-73, cylinder…:111-126,
cone…:112-127) incl. Levenberg-Marquardt refinement, through the C ABI vs the CPU oracle."""
import numpy as np
import pytest

import pitt_object_table_segmentation_b200 as pkg
from pitt_object_table_segmentation_b200 import _abi as A, scenes

pytestmark = pytest.mark.gpu

KINDS = {A.MODEL_PLANE: "plane", A.MODEL_SPHERE: "sphere", A.MODEL_CYLINDER: "cylinder", A.MODEL_CONE: "cone"}


def _cluster(kind, n, seed, oracle):
    xyz, truth = scenes.primitive_cluster(kind, n, seed)
    nrm = oracle.estimate_normals(xyz, 50, (0.0, 0.0, 0.0))
    return xyz, nrm, truth


@pytest.mark.parametrize("model", [A.MODEL_SPHERE, A.MODEL_CYLINDER, A.MODEL_CONE])
def test_score_counts_bit_exact(ctx, oracle, model):
    xyz, nrm, _ = _cluster(KINDS[model], 6000, 100 + model, oracle)
    cloud = ctx.stage(xyz, normals=nrm)
    p = pkg.default_sac_params(model)
    samples = oracle.pcl_sample_stream(xyz, model, 400)
    c_gpu, co_gpu, v_gpu = ctx.sac_score(cloud, p, samples)
    c_cpu, co_cpu, v_cpu = oracle.sac_score(xyz, nrm, p, samples)
    assert np.array_equal(v_gpu, v_cpu)
    assert np.array_equal(co_gpu.view(np.uint32), co_cpu.view(np.uint32))
    assert np.array_equal(c_gpu, c_cpu)
    assert c_gpu.max() > 1000  # the right model is among the hypotheses


@pytest.mark.parametrize("model", [A.MODEL_CYLINDER, A.MODEL_CONE])
def test_fast_path_equals_exact_path_on_wrong_shapes(ctx, oracle, model):
    # scoring a cylinder on a cone cloud (and vice versa) puts many points near the threshold
    other = "cone" if model == A.MODEL_CYLINDER else "cylinder"
    xyz, nrm, _ = _cluster(other, 5000, 7, oracle)
    cloud = ctx.stage(xyz, normals=nrm)
    p = pkg.default_sac_params(model)
    p.normal_distance_weight = 0.1  # widen the FP32 band
    samples = oracle.pcl_sample_stream(xyz, model, 300)
    c_gpu, _, _ = ctx.sac_score(cloud, p, samples)
    c_cpu, _, _ = oracle.sac_score(xyz, nrm, p, samples)
    assert np.array_equal(c_gpu, c_cpu)


@pytest.fixture
def lm_path(ctx, request):
    """Levenberg-Marquardt on one CTA (small problems) or on a cluster of 8 CTAs (>= 4096 rows by default): force either"""
    ctx.lib.pitt_debug_lm_cluster_min(1 if request.param == "cluster" else 1 << 30)
    yield request.param
    ctx.lib.pitt_debug_lm_cluster_min(4096)


@pytest.mark.parametrize("lm_path", ["cta", "cluster"], indirect=True)
@pytest.mark.parametrize("n", [4000, 700, 23000])
@pytest.mark.parametrize("model", [A.MODEL_SPHERE, A.MODEL_CYLINDER, A.MODEL_CONE])
def test_lm_refine_bit_exact(ctx, oracle, model, n, lm_path):
    xyz, nrm, _ = _cluster(KINDS[model], n, 200 + model, oracle)
    cloud = ctx.stage(xyz, normals=nrm)
    p = pkg.default_sac_params(model)
    p.optimize = 0
    base = oracle.sac_segment(xyz, nrm, p)
    assert len(base["inliers"]) > n // 8
    ref_c, info_c = oracle.sac_refine(xyz, nrm, p, base["coeffs"], base["inliers"])
    ref_g, info_g = ctx.sac_refine(cloud, p, base["coeffs"], base["inliers"])
    assert (info_g.lm_info, info_g.lm_nfev) == (info_c.lm_info, info_c.lm_nfev)
    assert np.array_equal(ref_g.view(np.uint32), ref_c.view(np.uint32)), (ref_g, ref_c)


@pytest.mark.parametrize("model,n,seed", [(A.MODEL_SPHERE, 5000, 1), (A.MODEL_CYLINDER, 5000, 2), (A.MODEL_CONE, 5000, 3),
                                          (A.MODEL_CYLINDER, 20000, 4), (A.MODEL_CONE, 12000, 5), (A.MODEL_SPHERE, 800, 6)])
def test_segment_bit_exact(ctx, oracle, model, n, seed):
    xyz, nrm, _ = _cluster(KINDS[model], n, seed, oracle)
    cloud = ctx.stage(xyz, normals=nrm)
    p = pkg.default_sac_params(model)
    got = ctx.sac_segment(cloud, p)
    want = oracle.sac_segment(xyz, nrm, p)
    gi, wi = got["info"], want["info"]
    assert (gi.iterations, gi.skipped, gi.best_hypothesis, gi.best_count, gi.n_inliers_model) == \
           (wi.iterations, wi.skipped, wi.best_hypothesis, wi.best_count, wi.n_inliers_model)
    assert (gi.lm_info, gi.lm_nfev) == (wi.lm_info, wi.lm_nfev)
    # north star: coefficients within 1e-5 relative, inlier sets bit-exact; we get both bit-exact
    np.testing.assert_allclose(got["coeffs"], want["coeffs"], rtol=1e-5, atol=1e-7)
    assert np.array_equal(got["coeffs"].view(np.uint32), want["coeffs"].view(np.uint32))
    assert np.array_equal(got["inliers"], want["inliers"])


@pytest.mark.parametrize("lm_path", ["cluster"], indirect=True)
def test_wrong_model_on_each_shape_cluster_lm(ctx, oracle, lm_path):
    """degenerate fits (rank-deficient Jacobians, early exits) through the cluster path of the LM kernel"""
    test_wrong_model_on_each_shape(ctx, oracle)


def test_wrong_model_on_each_shape(ctx, oracle):
    # every model on every shape (what ransac_segmentation.cpp does per cluster)
    for kind in ("sphere", "cylinder", "cone", "plane"):
        xyz, nrm, _ = _cluster(kind, 3000, 77, oracle)
        cloud = ctx.stage(xyz, normals=nrm)
        for model in (A.MODEL_PLANE, A.MODEL_SPHERE, A.MODEL_CYLINDER, A.MODEL_CONE):
            p = pkg.default_sac_params(model)
            got = ctx.sac_segment(cloud, p)
            want = oracle.sac_segment(xyz, nrm, p)
            assert np.array_equal(got["inliers"], want["inliers"]), (kind, model)
            assert np.array_equal(got["coeffs"].view(np.uint32), want["coeffs"].view(np.uint32)), (kind, model)


def test_cylinder_needs_normals(ctx):
    xyz, _ = scenes.primitive_cluster("cylinder", 500, 1)
    cloud = ctx.stage(xyz)
    with pytest.raises(pkg.PittError):
        ctx.sac_segment(cloud, pkg.default_sac_params(A.MODEL_CYLINDER))


def _nudge(v, steps, rng):
    """move each float32 of v by a random number of ulps in [-steps, steps]"""
    out = v.astype(np.float32).copy()
    k = rng.integers(-steps, steps + 1, out.shape)
    for _ in range(steps):
        up = k > 0
        dn = k < 0
        out = np.where(up, np.nextafter(out, np.float32(np.inf)), np.where(dn, np.nextafter(out, np.float32(-np.inf)), out)).astype(np.float32)
        k = k - np.sign(k)
    return out


@pytest.mark.parametrize("thr,radius", [(0.007, 0.06), (0.007, 0.004), (1e-4, 0.5), (0.02, 3.0)])
def test_sphere_interval_predicate_on_the_threshold(ctx, oracle, thr, radius):
    """the sphere predicate is evaluated as an interval test on the squared distance; points whose distance to
    the centre sits within a few ulps of r - thr, r + thr (and of r itself) must classify exactly like PCL"""
    rng = np.random.default_rng(17)
    n = 8192
    c = np.array([0.3, -0.2, 0.9], np.float32)
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    which = rng.integers(0, 3, n)
    rad = np.where(which == 0, radius - thr, np.where(which == 1, radius + thr, radius))
    rad = np.maximum(rad, 0.0)
    xyz = np.ones((n, 4), np.float32)
    xyz[:, :3] = _nudge((c + d * rad[:, None]).astype(np.float32), 3, rng)
    # four exact points of the sphere first: hypothesis 0 = the true sphere (centre c up to rounding)
    base = np.array([[1, 0, 0], [0, 1, 0], [0, 0, 1], [-1, 0, 0]], np.float32) * np.float32(radius) + c
    xyz[:4, :3] = base
    cloud = ctx.stage(xyz)
    p = pkg.default_sac_params(A.MODEL_SPHERE)
    p.distance_threshold = thr
    p.radius_min, p.radius_max = 0.0, 100.0
    samples = np.vstack([np.array([[0, 1, 2, 3]], np.int32), rng.integers(0, n, (255, 4)).astype(np.int32)])
    c_gpu, co_gpu, v_gpu = ctx.sac_score(cloud, p, samples)
    c_cpu, co_cpu, v_cpu = oracle.sac_score(xyz, None, p, samples)
    assert np.array_equal(v_gpu, v_cpu)
    assert np.array_equal(c_gpu, c_cpu)
    inl_g = ctx.sac_select(cloud, p, co_cpu[0, :4])
    inl_c = oracle.sac_select(xyz, None, p, co_cpu[0, :4])
    assert np.array_equal(inl_g, inl_c)
    if radius > thr:
        assert 0 < len(inl_c) < n  # the construction really straddles the threshold


@pytest.mark.parametrize("model,w", [(A.MODEL_CYLINDER, 0.001), (A.MODEL_CONE, 0.0006), (A.MODEL_CYLINDER, 0.0), (A.MODEL_CONE, 0.3),
                                     (A.MODEL_CYLINDER, 1.0)])
def test_cylinder_cone_prefilter_near_threshold(ctx, oracle, model, w):
    """shell of points at axis distances spread tightly around r +- thr/(1-w): the certain-outlier pre-filter
    (euclidean term only) must never change a count"""
    rng = np.random.default_rng(23)
    kind = KINDS[model]
    xyz, nrm, truth = _cluster(kind, 4000, 31, oracle)
    # radial jitter concentrated at the decision boundary of the true model
    p = pkg.default_sac_params(model)
    p.normal_distance_weight = w
    thr = p.distance_threshold
    centre = xyz[:, :3].mean(axis=0)
    scale = 1.0 + rng.choice([-1.0, 1.0], len(xyz))[:, None] * (thr / max(1.0 - w, 0.05)) * rng.uniform(0.98, 1.02, (len(xyz), 1)) / 0.05
    xyz2 = xyz.copy()
    xyz2[:, :3] = (centre + (xyz[:, :3] - centre) * scale).astype(np.float32)
    xyz2[:2000] = xyz[:2000]  # keep half of the true surface so that good hypotheses exist
    nrm2 = oracle.estimate_normals(xyz2, 50, (0.0, 0.0, 0.0))
    cloud = ctx.stage(xyz2, normals=nrm2)
    samples = np.vstack([oracle.pcl_sample_stream(xyz2, model, 300),
                         rng.integers(0, 2000, (300, A.SAMPLE_SIZE[model])).astype(np.int32)])
    c_gpu, _, v_gpu = ctx.sac_score(cloud, p, samples)
    c_cpu, _, v_cpu = oracle.sac_score(xyz2, nrm2, p, samples)
    assert np.array_equal(v_gpu, v_cpu)
    assert np.array_equal(c_gpu, c_cpu)


def _score_modes(ctx, cloud, p, samples):
    """counts from the two-tier kernel (default) and from the generic kernel (test hook)"""
    c2, _, v2 = ctx.sac_score(cloud, p, samples)
    ctx.lib.pitt_debug_score_mode(1)
    try:
        c1, _, v1 = ctx.sac_score(cloud, p, samples)
    finally:
        ctx.lib.pitt_debug_score_mode(0)
    return c2, v2, c1, v1


@pytest.mark.parametrize("model,w,normals", [(A.MODEL_CYLINDER, 0.001, "random"), (A.MODEL_CONE, 0.0006, "random"),
                                             (A.MODEL_CYLINDER, 0.001, "broken"), (A.MODEL_CONE, 0.0006, "broken"),
                                             (A.MODEL_CYLINDER, 0.003, "random"), (A.MODEL_CONE, 0.002, "estimated")])
def test_certain_inlier_shortcut_near_its_boundary(ctx, oracle, model, w, normals):
    """the two-tier kernel accepts an evaluation without looking at the normal when the euclidean term alone keeps the
    score below the threshold for ANY angle (d_euclid <= D_in = (thr - band - w pi/2) / (1 - w)). Points are spread
    tightly around that boundary and given random / zero / NaN / huge normals: counts must equal PCL's."""
    rng = np.random.default_rng(29)
    kind = KINDS[model]
    xyz, nrm, truth = _cluster(kind, 4000, 41, oracle)
    p = pkg.default_sac_params(model)
    p.normal_distance_weight = w
    thr = p.distance_threshold
    band = 2e-3 * w + 2e-6
    d_in = (thr - band - w * np.pi / 2) / (1.0 - w)
    assert d_in > 0
    centre = xyz[:, :3].mean(axis=0)
    # radial jitter (relative to the cluster centre, which is on the axis for these scenes only approximately: the
    # spread factor makes up for it) concentrated at +- d_in
    scale = 1.0 + rng.choice([-1.0, 1.0], len(xyz))[:, None] * d_in * rng.uniform(0.9, 1.1, (len(xyz), 1)) / 0.05
    xyz2 = xyz.copy()
    xyz2[:, :3] = (centre + (xyz[:, :3] - centre) * scale).astype(np.float32)
    xyz2[:1500] = xyz[:1500]
    if normals == "estimated":
        nrm2 = oracle.estimate_normals(xyz2, 50, (0.0, 0.0, 0.0))
    else:
        nrm2 = nrm.copy()
        v = rng.normal(size=(len(xyz2), 3)).astype(np.float32)
        v /= np.linalg.norm(v, axis=1, keepdims=True)
        nrm2[1500:, :3] = v[1500:]  # the angle term is as large as it gets for half of the cloud
        if normals == "broken":
            nrm2[1500:1700, :3] = 0.0
            nrm2[1700:1800, 0] = np.nan
            nrm2[1800:1900, :3] *= np.float32(1e20)
            nrm2[1900:2000, :3] *= np.float32(1e-20)
            nrm2[2000:2050, 1] = np.inf
    cloud = ctx.stage(xyz2, normals=nrm2)
    samples = np.vstack([oracle.pcl_sample_stream(xyz2, model, 200),
                         rng.integers(0, 1500, (400, A.SAMPLE_SIZE[model])).astype(np.int32)])
    c2, v2, c1, v1 = _score_modes(ctx, cloud, p, samples)
    c_cpu, _, v_cpu = oracle.sac_score(xyz2, nrm2, p, samples)
    assert np.array_equal(v2, v_cpu) and np.array_equal(v1, v_cpu)
    assert np.array_equal(c1, c_cpu)
    assert np.array_equal(c2, c_cpu)
    assert c_cpu.max() > 500


@pytest.mark.parametrize("model", [A.MODEL_CYLINDER, A.MODEL_CONE])
@pytest.mark.parametrize("n,offset", [(50000, 0.0), (3000, 0.0), (20000, 1500.0), (4097, 0.0)])
def test_two_tier_scoring_equals_generic(ctx, oracle, model, n, offset):
    """C3-shaped job (big cluster, thousands of hypotheses): the two-tier kernel against the generic one; a slice against
    the oracle. offset = 1500 m moves the cloud beyond the |coordinate| <= 1000 guard of the certain-inlier shortcut."""
    rng = np.random.default_rng(5)
    xyz, _ = scenes.primitive_cluster(KINDS[model], n, 9)
    xyz[:, :3] += np.float32(offset)
    cloud = ctx.stage(xyz)
    vp = (float(offset), float(offset), float(offset))
    nrm = ctx.estimate_normals(cloud, 50, vp)
    p = pkg.default_sac_params(model)
    S = A.SAMPLE_SIZE[model]
    samples = rng.integers(0, n, (2000, S)).astype(np.int32)
    c2, v2, c1, v1 = _score_modes(ctx, cloud, p, samples)
    assert np.array_equal(v2, v1)
    assert np.array_equal(c2, c1)
    if offset == 0.0:
        assert c2.max() > n // 4
    c_cpu, _, v_cpu = oracle.sac_score(xyz, nrm, p, samples[:48])
    assert np.array_equal(c2[:48], c_cpu)


@pytest.mark.parametrize("model,n,outliers", [(A.MODEL_CYLINDER, 30000, 0.0), (A.MODEL_CONE, 30000, 0.0), (A.MODEL_CONE, 4000, 0.8),
                                               (A.MODEL_CYLINDER, 4000, 0.93), (A.MODEL_SPHERE, 6000, 0.85)])
def test_c3_segment_with_10000_iterations(ctx, oracle, model, n, outliers):
    """config 3 as the reference would run it: setMaxIterations(10000) with PCL's adaptive stop. The device scores the sample
    stream in batches (256, then 4096) and stops where PCL stops; clean clusters end after a few hypotheses, clusters
    drowned in outliers need several batches. Everything equal to the oracle, iteration counts included."""
    rng = np.random.default_rng(77)
    xyz, _ = scenes.primitive_cluster(KINDS[model], n, 21)
    if outliers > 0:
        m = int(n * outliers)
        lo, hi = xyz[:, :3].min(0) - 0.05, xyz[:, :3].max(0) + 0.05
        xyz[rng.choice(n, m, replace=False), :3] = rng.uniform(lo, hi, (m, 3)).astype(np.float32)
    nrm = oracle.estimate_normals(xyz, 50, (0.0, 0.0, 0.0))
    cloud = ctx.stage(xyz, normals=nrm)
    p = pkg.default_sac_params(model)
    p.max_iterations = 10000
    got = ctx.sac_segment(cloud, p)
    want = oracle.sac_segment(xyz, nrm, p)
    gi, wi = got["info"], want["info"]
    assert (gi.iterations, gi.skipped, gi.best_hypothesis, gi.best_count, gi.n_inliers_model) == \
           (wi.iterations, wi.skipped, wi.best_hypothesis, wi.best_count, wi.n_inliers_model)
    assert (gi.lm_info, gi.lm_nfev) == (wi.lm_info, wi.lm_nfev)
    assert np.array_equal(got["coeffs"].view(np.uint32), want["coeffs"].view(np.uint32))
    assert np.array_equal(got["inliers"], want["inliers"])
    if outliers > 0:
        assert gi.iterations > 200 and (model == A.MODEL_CYLINDER or gi.iterations + gi.skipped > 256)  # beyond the first batch
    else:
        assert gi.iterations < 256 and gi.hypotheses <= 256  # ... and here inside it: 256 hypotheses scored, not 10 001


@pytest.mark.parametrize("model", [A.MODEL_PLANE, A.MODEL_SPHERE, A.MODEL_CYLINDER, A.MODEL_CONE])
@pytest.mark.parametrize("n", [5, 31, 1023, 1024, 1025, 5000, 16384])
def test_fused_select_of_small_clouds(ctx, oracle, model, n):
    """selectWithinDistance on clouds of <= 16384 points is one single-CTA launch (record preparation + predicate + ordered
    compaction); same ascending index list as the four-launch path and the oracle"""
    xyz, _ = scenes.primitive_cluster(KINDS[model], max(n, 64), 300 + model)
    xyz = np.ascontiguousarray(xyz[:n])
    nrm = oracle.estimate_normals(xyz, min(50, n), (0.0, 0.0, 0.0)) if n >= 3 else np.zeros((n, 4), np.float32)
    cloud = ctx.stage(xyz, normals=nrm)
    p = pkg.default_sac_params(model)
    p0 = p.copy()
    p0.optimize = 0
    if n >= 1000:
        co = oracle.sac_segment(xyz, nrm, p0)["coeffs"]  # a good model of this very cloud
    else:
        big, _ = scenes.primitive_cluster(KINDS[model], 4000, 300 + model)
        bn = oracle.estimate_normals(big, 50, (0.0, 0.0, 0.0))
        co = oracle.sac_segment(big, bn, p0)["coeffs"]  # some model of the same kind
    got = ctx.sac_select(cloud, p, co)
    ctx.lib.pitt_debug_select_no_fuse(1)
    try:
        unfused = ctx.sac_select(cloud, p, co)
    finally:
        ctx.lib.pitt_debug_select_no_fuse(0)
    want = oracle.sac_select(xyz, nrm, p, co)
    assert np.array_equal(got, unfused)
    assert np.array_equal(got, want)
    if n >= 1000:
        assert len(want) > n // 8


@pytest.mark.parametrize("n,H,mode", [(100001, 1003, 0), (200001, 512, 0), (33000, 4097, 0), (40000, 1003, 2), (40000, 1003, 3)])
def test_sphere_packed_kernel_equals_generic(ctx, oracle, n, H, mode):
    """large sphere jobs (n x H >= 1e8, H >= 512) take sphere_score_kernel (lanes = hypotheses, packed f32x2): counts must
    equal the generic kernel's and the oracle's for every hypothesis; ragged n and H, NaN / infinite points,
    half of the points moved onto the two surfaces |d - r| = thr of the true sphere"""
    rng = np.random.default_rng(n + H)
    xyz, truth = scenes.primitive_cluster("sphere", n, 3)
    p = pkg.default_sac_params(A.MODEL_SPHERE)
    thr = p.distance_threshold
    P = xyz[:, :3].astype(np.float64)
    sol = np.linalg.lstsq(np.c_[2.0 * P, np.ones(n)], (P * P).sum(axis=1), rcond=None)[0]  # algebraic sphere fit
    centre = sol[:3]
    v = P - centre
    r = np.linalg.norm(v, axis=1, keepdims=True)
    r0 = float(np.median(r))
    assert abs(r0 - truth["radius"]) < 1e-3
    k = n // 2
    target = r0 + rng.choice([-1.0, 1.0], (k, 1)) * thr * (1.0 + rng.integers(-4, 5, (k, 1)) * 6e-8)
    xyz[:k, :3] = (centre + v[:k] / r[:k] * target).astype(np.float32)
    xyz[k] = (np.nan, 0.0, 0.0, 1.0)
    xyz[k + 1] = (np.inf, 0.1, 0.2, 1.0)
    xyz[k + 2] = (1e20, -1e20, 3.0, 1.0)
    cloud = ctx.stage(xyz)
    samples = oracle.pcl_sample_stream(xyz, A.MODEL_SPHERE, H)
    samples[5] = samples[5][0]            # degenerate sample: invalid hypothesis
    samples[H - 1] = (k, 1, 2, 3)         # through the NaN point
    ctx.lib.pitt_debug_score_mode(mode)  # 0: automatic (packed kernel at this size), 2 / 3: packed kernel, 512- / 128-point tiles
    try:
        c2, _, v2 = ctx.sac_score(cloud, p, samples)
        ctx.lib.pitt_debug_score_mode(1)  # generic kernel
        c1, _, v1 = ctx.sac_score(cloud, p, samples)
    finally:
        ctx.lib.pitt_debug_score_mode(0)
    assert np.array_equal(v2, v1)
    assert np.array_equal(c2, c1)
    c_cpu, _, v_cpu = oracle.sac_score(xyz, None, p, samples)
    assert np.array_equal(c2, c_cpu) and np.array_equal(v2, v_cpu)
    assert c2[5] == 0 and c2[H - 1] == 0
    assert c2.max() > n // 4
    cloud.release()
