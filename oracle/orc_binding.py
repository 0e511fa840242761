"""ctypes binding of the CPU oracle (oracle/liborc.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs as the checker / reported baseline. The product package never imports it.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from pitt_object_table_segmentation_b200 import _abi as A

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    so = os.path.join(_HERE, "liborc.so")
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".cpp", ".h"))]
    srcs.append(os.path.join(_HERE, "..", "include", "pitt_b200.h"))
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "liborc.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.orc_mt19937_nth.restype = C.c_uint32
    return _LIB


def _f4(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    assert a.ndim == 2 and a.shape[1] == 4, a.shape
    return a


def _fp(a):
    return a.ctypes.data_as(A.f32p) if a is not None else None


def _ip(a):
    return a.ctypes.data_as(A.i32p) if a is not None else None


def default_sac_params(model):
    p = A.SacParams()
    lib().orc_default_sac_params(int(model), C.byref(p))
    return p


def default_support_sac_params():
    p = A.SacParams()
    lib().orc_default_support_sac_params(C.byref(p))
    return p


def mt19937_nth(seed, nth):
    return int(lib().orc_mt19937_nth(C.c_uint32(seed), int(nth)))


def sac_segment(xyz4, nrm4, params):
    xyz4 = _f4(xyz4)
    nrm4 = _f4(nrm4) if nrm4 is not None else None
    n = xyz4.shape[0]
    inl = np.empty(max(n, 1), np.int32)
    n_inl, n_co = C.c_int(0), C.c_int(0)
    co = np.zeros(8, np.float32)
    info = A.SacInfo()
    keep = _hold_replay(params)
    st = lib().orc_sac_segment(_fp(xyz4), _fp(nrm4), n, C.byref(params), _ip(inl), n, C.byref(n_inl), _fp(co),
                               C.byref(n_co), C.byref(info))
    del keep
    assert st == 0, st
    return {"inliers": inl[: n_inl.value].copy(), "coeffs": co[: n_co.value].copy(), "info": info}


def _hold_replay(params):
    return params.replay_samples  # the caller keeps the numpy array alive; nothing to do


def sac_score(xyz4, nrm4, params, samples):
    xyz4 = _f4(xyz4)
    nrm4 = _f4(nrm4) if nrm4 is not None else None
    samples = np.ascontiguousarray(samples, np.int32)
    H = samples.shape[0]
    counts = np.zeros(H, np.int32)
    co = np.zeros((H, 8), np.float32)
    valid = np.zeros(H, np.uint8)
    st = lib().orc_sac_score(_fp(xyz4), _fp(nrm4), xyz4.shape[0], C.byref(params), _ip(samples), H, _ip(counts),
                             _fp(co), valid.ctypes.data_as(C.POINTER(C.c_uint8)))
    assert st == 0, st
    return counts, co, valid


def sac_select(xyz4, nrm4, params, coeffs):
    xyz4 = _f4(xyz4)
    nrm4 = _f4(nrm4) if nrm4 is not None else None
    n = xyz4.shape[0]
    co = np.zeros(8, np.float32)
    co[: len(coeffs)] = coeffs
    inl = np.empty(max(n, 1), np.int32)
    n_inl = C.c_int(0)
    st = lib().orc_sac_select(_fp(xyz4), _fp(nrm4), n, C.byref(params), _fp(co), _ip(inl), n, C.byref(n_inl))
    assert st == 0, st
    return inl[: n_inl.value].copy()


def sac_refine(xyz4, nrm4, params, coeffs, inliers):
    xyz4 = _f4(xyz4)
    nrm4 = _f4(nrm4) if nrm4 is not None else None
    co = np.zeros(8, np.float32)
    co[: len(coeffs)] = coeffs
    inliers = np.ascontiguousarray(inliers, np.int32)
    out = np.zeros(8, np.float32)
    info = A.SacInfo()
    st = lib().orc_sac_refine(_fp(xyz4), _fp(nrm4), xyz4.shape[0], C.byref(params), _fp(co), _ip(inliers),
                              len(inliers), _fp(out), C.byref(info))
    assert st == 0, st
    return out[: A.N_COEFFS[params.model]].copy(), info


def pcl_sample_stream(xyz4, model, count):
    xyz4 = _f4(xyz4)
    S = A.SAMPLE_SIZE[model]
    out = np.zeros((count, S), np.int32)
    st = lib().orc_pcl_sample_stream(_fp(xyz4), xyz4.shape[0], int(model), int(count), _ip(out))
    assert st == 0, st
    return out


def knn(xyz4, k):
    xyz4 = _f4(xyz4)
    n = xyz4.shape[0]
    idx = np.zeros((n, k), np.int32)
    sq = np.zeros((n, k), np.float32)
    st = lib().orc_knn(_fp(xyz4), n, int(k), _ip(idx), _fp(sq))
    assert st == 0, st
    return idx, sq


def estimate_normals(xyz4, k=50, viewpoint=(0.0, 0.0, 0.0)):
    xyz4 = _f4(xyz4)
    n = xyz4.shape[0]
    out = np.zeros((n, 4), np.float32)
    vp = (C.c_float * 3)(*viewpoint)
    st = lib().orc_estimate_normals(_fp(xyz4), n, int(k), vp, _fp(out))
    assert st == 0, st
    return out


def euclidean_clusters(xyz4, tolerance, min_size, max_size):
    xyz4 = _f4(xyz4)
    n = xyz4.shape[0]
    labels = np.full(n, -1, np.int32)
    nc = C.c_int(0)
    st = lib().orc_euclidean_clusters(_fp(xyz4), n, C.c_double(tolerance), int(min_size), int(max_size), _ip(labels),
                                      C.byref(nc))
    assert st == 0, st
    return labels, nc.value


# ---------------------------------------------------------------- service-shaped entry points
from pitt_object_table_segmentation_b200 import _results as R  # noqa: E402


def default_support_params():
    p = A.SupportParams()
    p.min_iterative_cloud_percentual_size = -1.0
    p.min_iterative_plane_percentual_size = -1.0
    p.variance_threshold_for_horizontal = -1.0
    p.ransac_distance_point_in_shape_threshold = -1.0
    p.ransac_model_normal_distance_weigth = -1.0
    p.ransac_max_iteration_threshold = -1
    p.horizontal_axis_len = 1
    p.support_edge_remove_offset_len = 1
    p.normals_k = 50
    return p


def default_cluster_params():
    p = A.ClusterParams()
    p.tolerance, p.min_rate, p.max_rate, p.min_input_size = 0.03, 0.01, 0.99, 30
    return p


def default_frame_params():
    p = A.FrameParams()
    p.support = default_support_params()
    p.cluster = default_cluster_params()
    p.plane = default_sac_params(A.MODEL_PLANE)
    p.sphere = default_sac_params(A.MODEL_SPHERE)
    p.cylinder = default_sac_params(A.MODEL_CYLINDER)
    p.cone = default_sac_params(A.MODEL_CONE)
    p.normals_k, p.min_points, p.cone_over_cylinder_priority = 50, 30, 0.9
    return p


def find_supports(xyz4, nrm4, params, supports_cap=4):
    xyz4 = _f4(xyz4)
    nrm4 = _f4(nrm4) if nrm4 is not None else None
    b = R.SupportBuffers(xyz4.shape[0], supports_cap)
    st = lib().orc_find_supports(_fp(xyz4), _fp(nrm4), xyz4.shape[0], C.byref(params), C.byref(b.res))
    assert st == 0, st
    return b.to_python()


def cluster_service(xyz4, params):
    xyz4 = _f4(xyz4)
    b = R.ClusterBuffers(xyz4.shape[0])
    st = lib().orc_cluster_service(_fp(xyz4), xyz4.shape[0], C.byref(params), C.byref(b.res))
    assert st == 0, st
    return b.to_python()


def primitive_service(xyz4, nrm4, params):
    xyz4 = _f4(xyz4)
    nrm4 = _f4(nrm4) if nrm4 is not None else None
    b = R.PrimitiveBuffers(xyz4.shape[0])
    st = lib().orc_primitive_service(_fp(xyz4), _fp(nrm4), xyz4.shape[0], C.byref(params), C.byref(b.res))
    assert st == 0, st
    return b.to_python()


def select_primitive(plane, sphere, cylinder, cone, prio=0.9):
    return int(lib().orc_select_primitive(C.c_int64(plane), C.c_int64(sphere), C.c_int64(cylinder), C.c_int64(cone),
                                          C.c_float(prio)))


def segment_frame(xyz4, params, shapes_cap=64):
    xyz4 = _f4(xyz4)
    b = R.FrameBuffers(shapes_cap)
    st = lib().orc_segment_frame(_fp(xyz4), xyz4.shape[0], C.byref(params), C.byref(b.res))
    assert st == 0, st
    return b.to_python()


def prefilter(raw, params):
    """orc_prefilter: raw = (n, point_step/4) float32 payload; returns (world cloud n x 4, info dict)"""
    raw = np.ascontiguousarray(raw, np.float32)
    n = raw.shape[0]
    out = np.zeros((max(n, 1), 4), np.float32)
    n_out = C.c_int(0)
    info = A.PrefilterInfo()
    st = lib().orc_prefilter(C.c_void_p(raw.ctypes.data), raw.shape[1] * 4, n, C.byref(params), _fp(out), max(n, 1),
                             C.byref(n_out), C.byref(info))
    assert st == 0, st
    return out[: n_out.value].copy(), {k: getattr(info, k) for k, _ in A.PrefilterInfo._fields_}


def arm_filter(xyz4, params):
    """orc_arm_filter: chained negative CropBoxes; returns (kept cloud n x 4, removed per box)"""
    xyz4 = np.ascontiguousarray(xyz4, np.float32)
    n = xyz4.shape[0]
    out = np.zeros((max(n, 1), 4), np.float32)
    n_out = C.c_int(0)
    removed = (C.c_int * 4)()
    st = lib().orc_arm_filter(_fp(xyz4), n, C.byref(params), _fp(out), max(n, 1), C.byref(n_out), removed)
    assert st == 0, st
    return out[: n_out.value].copy(), list(removed)
