"""Python host mirror of the C ABI (include/pitt_b200.h) — thin ctypes calls into libpitt_b200.so.

There is no CPU fallback: `Context()` raises when the CUDA library is missing or no B200 is visible.
The classes below only marshal numpy arrays; every number is produced by the CUDA kernels in csrc/.
"""
import ctypes as C
import os

import numpy as np

from . import _abi as A

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpitt_b200.so")
_LIB = None

# every symbol declared in include/pitt_b200.h
EXPORTED_SYMBOLS = [
    "pitt_create", "pitt_create_on_stream", "pitt_destroy", "pitt_last_error", "pitt_version", "pitt_device_count",
    "pitt_synchronize", "pitt_set_blocking_sync", "pitt_set_workers", "pitt_default_prefilter_params", "pitt_prefilter_cloud",
    "pitt_prefilter_staged", "pitt_default_arm_filter_params", "pitt_arm_filter", "pitt_get_points", "pitt_segment_raw_frames_batched", "pitt_default_sac_params", "pitt_default_support_sac_params", "pitt_default_support_params",
    "pitt_default_cluster_params", "pitt_default_frame_params", "pitt_stage_cloud", "pitt_stage_cloud_device",
    "pitt_set_normals", "pitt_cloud_size", "pitt_cloud_has_normals", "pitt_cloud_device_points",
    "pitt_cloud_device_normals", "pitt_release_cloud", "pitt_estimate_normals", "pitt_get_normals", "pitt_knn",
    "pitt_sac_segment", "pitt_sac_segment_host", "pitt_sac_score", "pitt_sac_score_device", "pitt_argmax_counts_device", "pitt_sac_finish_device", "pitt_sac_select", "pitt_sac_refine",
    "pitt_pcl_sample_stream", "pitt_euclidean_clusters", "pitt_find_supports", "pitt_cluster_service",
    "pitt_primitive_service", "pitt_select_primitive", "pitt_segment_frame", "pitt_segment_frames_batched", "pitt_fp32_peak", "pitt_last_device_ms",
    "pitt_kernel_launches", "pitt_segment_clouds_batched", "pitt_sac_segment_split",
]
# include/pitt_b200_debug.h: test / measurement hooks, not part of the drop-in boundary
DEBUG_SYMBOLS = [
    "pitt_debug_plane_mode", "pitt_debug_force_generic_plane", "pitt_debug_score_mode", "pitt_debug_select_no_fuse",
    "pitt_debug_lm_cluster_min", "pitt_debug_stream_chunks", "pitt_debug_plane_filter_stats", "pitt_debug_plane_tc_stats",
    "pitt_debug_plane_tc_cta_cycles", "pitt_debug_plane_tc_dump", "pitt_debug_plane_tc_acc_ulps", "pitt_debug_plane_tc_variant",
    "pitt_debug_plane_tc_nwq", "pitt_debug_plane_tc_time_kernel", "pitt_debug_plane_tc_kernel_ms", "pitt_debug_knn_stats", "pitt_debug_knn_timeline", "pitt_debug_frame_mode", "pitt_debug_philox", "pitt_debug_philox_samples",
]


# pitt_allgather_fn of include/pitt_b200.h: (user, d_send, d_recv, count_per_rank, cuda_stream) -> 0 on success
ALLGATHER_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p)


class _DevArray:
    """a raw device pointer as a __cuda_array_interface__ object (int32), so that torch can wrap it without a copy"""

    def __init__(self, ptr, count):
        self.__cuda_array_interface__ = {"shape": (int(count),), "typestr": "<i4", "data": (int(ptr), False), "version": 2}


def torch_allgather_callback(group=None):
    """pitt_allgather_fn on top of torch.distributed (NCCL under torchrun, gloo in the CPU tests is not applicable: the
    buffers are device memory). The collective is ordered on the context's stream (passed to the callback)."""
    import torch
    import torch.distributed as dist

    def _cb(user, d_send, d_recv, count, stream):
        try:
            world = dist.get_world_size(group)
            # the counts are produced on the context's stream: make it torch's current stream for the collective
            # (ctypes hands a NULL void* over as None: a context created on the legacy default stream)
            ts = torch.cuda.ExternalStream(int(stream)) if stream else torch.cuda.default_stream()
            with torch.cuda.stream(ts):
                send = torch.as_tensor(_DevArray(d_send, count), device="cuda")
                recv = torch.as_tensor(_DevArray(d_recv, count * world), device="cuda")
                dist.all_gather_into_tensor(recv, send, group=group)
            return 0
        except Exception as e:  # noqa: BLE001 - reported through the status code
            import sys
            print(f"[pitt] all-gather callback failed: {e}", file=sys.stderr)
            return 1

    return ALLGATHER_FN(_cb)


class PittError(RuntimeError):
    pass


def load_library():
    """dlopen libpitt_b200.so (built in-tree by __graft_entry__.build()). Raises if it is missing."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise PittError(f"{LIB_PATH} not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                        "(there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    vp = C.c_void_p
    lib.pitt_create.restype = vp
    lib.pitt_create.argtypes = [C.c_int, C.c_uint64]
    lib.pitt_create_on_stream.restype = vp
    lib.pitt_create_on_stream.argtypes = [C.c_int, C.c_uint64, vp]
    lib.pitt_destroy.argtypes = [vp]
    lib.pitt_last_error.restype = C.c_char_p
    lib.pitt_last_error.argtypes = [vp]
    lib.pitt_version.restype = C.c_char_p
    lib.pitt_synchronize.argtypes = [vp]
    lib.pitt_set_workers.argtypes = [vp, C.c_int]
    lib.pitt_set_blocking_sync.argtypes = [vp, C.c_int]
    lib.pitt_prefilter_cloud.argtypes = [vp, vp, C.c_int, C.c_int, C.POINTER(A.PrefilterParams), C.POINTER(vp),
                                         C.POINTER(A.PrefilterInfo)]
    lib.pitt_prefilter_staged.argtypes = [vp, vp, C.POINTER(A.PrefilterParams), C.POINTER(vp), C.POINTER(A.PrefilterInfo)]
    lib.pitt_get_points.argtypes = [vp, vp, A.f32p]
    lib.pitt_default_arm_filter_params.argtypes = [C.POINTER(A.ArmFilterParams)]
    lib.pitt_arm_filter.argtypes = [vp, vp, C.POINTER(A.ArmFilterParams), C.POINTER(vp), A.i32p]
    lib.pitt_stage_cloud.argtypes = [vp, vp, C.c_int, C.c_int, C.POINTER(vp)]
    lib.pitt_stage_cloud_device.argtypes = [vp, vp, C.c_int, C.POINTER(vp)]
    lib.pitt_set_normals.argtypes = [vp, vp, vp, C.c_int]
    lib.pitt_cloud_size.argtypes = [vp]
    lib.pitt_cloud_has_normals.argtypes = [vp]
    lib.pitt_cloud_device_points.restype = vp
    lib.pitt_cloud_device_points.argtypes = [vp]
    lib.pitt_cloud_device_normals.restype = vp
    lib.pitt_cloud_device_normals.argtypes = [vp]
    lib.pitt_release_cloud.argtypes = [vp, vp]
    lib.pitt_estimate_normals.argtypes = [vp, vp, C.c_int, A.f32p]
    lib.pitt_get_normals.argtypes = [vp, vp, A.f32p]
    lib.pitt_knn.argtypes = [vp, vp, C.c_int, A.i32p, A.f32p]
    lib.pitt_sac_segment.argtypes = [vp, vp, C.POINTER(A.SacParams), A.i32p, C.c_int, C.POINTER(C.c_int), A.f32p,
                                     C.POINTER(C.c_int), C.POINTER(A.SacInfo)]
    lib.pitt_sac_segment_host.argtypes = [vp, vp, C.c_int, C.c_int, C.POINTER(A.SacParams), A.i32p, C.c_int, C.POINTER(C.c_int),
                                          A.f32p, C.POINTER(C.c_int), C.POINTER(A.SacInfo)]
    lib.pitt_sac_score.argtypes = [vp, vp, C.POINTER(A.SacParams), A.i32p, C.c_int, A.i32p, A.f32p,
                                   C.POINTER(C.c_uint8)]
    lib.pitt_sac_score_device.argtypes = [vp, vp, C.POINTER(A.SacParams), vp, C.c_int, vp]
    lib.pitt_argmax_counts_device.argtypes = [vp, vp, C.c_int, vp]
    lib.pitt_sac_finish_device.argtypes = [vp, vp, C.POINTER(A.SacParams), vp, C.c_int, vp, A.i32p, C.c_int,
                                           C.POINTER(C.c_int), A.f32p, C.POINTER(C.c_int), C.POINTER(A.SacInfo)]
    lib.pitt_sac_select.argtypes = [vp, vp, C.POINTER(A.SacParams), A.f32p, A.i32p, C.c_int, C.POINTER(C.c_int)]
    lib.pitt_sac_refine.argtypes = [vp, vp, C.POINTER(A.SacParams), A.f32p, A.i32p, C.c_int, A.f32p,
                                    C.POINTER(A.SacInfo)]
    lib.pitt_pcl_sample_stream.argtypes = [vp, vp, C.c_int, C.c_int, A.i32p]
    lib.pitt_euclidean_clusters.argtypes = [vp, vp, C.c_double, C.c_int, C.c_int, A.i32p, C.POINTER(C.c_int)]
    lib.pitt_find_supports.argtypes = [vp, vp, C.POINTER(A.SupportParams), C.POINTER(A.SupportResult)]
    lib.pitt_cluster_service.argtypes = [vp, vp, C.POINTER(A.ClusterParams), C.POINTER(A.ClustersResult)]
    lib.pitt_primitive_service.argtypes = [vp, vp, C.POINTER(A.SacParams), C.POINTER(A.PrimitiveResult)]
    lib.pitt_select_primitive.argtypes = [C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_float]
    lib.pitt_segment_frame.argtypes = [vp, vp, C.POINTER(A.FrameParams), C.POINTER(A.FrameResult)]
    lib.pitt_segment_frames_batched.argtypes = [C.POINTER(vp), C.c_int, C.POINTER(vp), A.i32p, C.c_int, C.c_int,
                                                C.POINTER(A.FrameParams), C.POINTER(A.FrameResult)]
    lib.pitt_sac_segment_split.argtypes = [vp, vp, C.POINTER(A.SacParams), C.c_int, C.c_int, ALLGATHER_FN, vp, A.i32p, C.c_int,
                                           C.POINTER(C.c_int), A.f32p, C.POINTER(C.c_int), C.POINTER(A.SacInfo)]
    lib.pitt_segment_clouds_batched.argtypes = [C.POINTER(vp), C.c_int, C.POINTER(vp), C.c_int, C.POINTER(A.FrameParams),
                                                C.POINTER(A.FrameResult)]
    lib.pitt_debug_plane_tc_time_kernel.argtypes = [vp, C.c_int]
    lib.pitt_debug_plane_tc_time_kernel.restype = None
    lib.pitt_debug_knn_stats.argtypes = [vp, C.c_int, C.POINTER(C.c_int64)]
    lib.pitt_debug_knn_timeline.argtypes = [vp, C.c_int, C.POINTER(C.c_uint64), C.c_int]
    u32p = C.POINTER(C.c_uint32)
    lib.pitt_debug_philox.argtypes = [vp, u32p, u32p, u32p]
    lib.pitt_debug_philox_samples.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_uint32, A.i32p]
    lib.pitt_debug_plane_tc_kernel_ms.argtypes = [vp]
    lib.pitt_debug_plane_tc_kernel_ms.restype = C.c_double
    lib.pitt_fp32_peak.argtypes = [vp, C.c_int, C.POINTER(C.c_double)]
    lib.pitt_last_device_ms.restype = C.c_double
    lib.pitt_last_device_ms.argtypes = [vp]
    lib.pitt_kernel_launches.restype = C.c_int64
    lib.pitt_kernel_launches.argtypes = [vp]
    _LIB = lib
    return lib


def default_sac_params(model):
    p = A.SacParams()
    load_library().pitt_default_sac_params(int(model), C.byref(p))
    return p


def default_support_sac_params():
    p = A.SacParams()
    load_library().pitt_default_support_sac_params(C.byref(p))
    return p


def default_support_params():
    p = A.SupportParams()
    load_library().pitt_default_support_params(C.byref(p))
    return p


def default_cluster_params():
    p = A.ClusterParams()
    load_library().pitt_default_cluster_params(C.byref(p))
    return p


def default_frame_params():
    p = A.FrameParams()
    load_library().pitt_default_frame_params(C.byref(p))
    return p


def select_primitive(plane, sphere, cylinder, cone, priority=0.9):
    return int(load_library().pitt_select_primitive(int(plane), int(sphere), int(cylinder), int(cone),
                                                    C.c_float(priority)))


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


class Cloud:
    """A cloud staged in HBM (pitt_cloud). Created through Context.stage()."""

    def __init__(self, ctx, handle, n):
        self.ctx, self.handle, self.n = ctx, handle, n

    def release(self):
        if self.handle:
            self.ctx.lib.pitt_release_cloud(self.ctx.handle, self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass

    def set_normals(self, nrm):
        nrm = _f32(nrm)
        assert nrm.shape == (self.n, 4) or nrm.shape == (self.n, 8)
        self.ctx._check(self.ctx.lib.pitt_set_normals(self.ctx.handle, self.handle, nrm.ctypes.data, nrm.shape[1] * 4))
        return self

    @property
    def device_points(self):
        return self.ctx.lib.pitt_cloud_device_points(self.handle)


class Context:
    """One CUDA stream on one B200 (pitt_ctx)."""

    def __init__(self, device=0, seed=12345, stream=None):
        self.lib = load_library()
        if self.lib.pitt_device_count() <= 0:
            raise PittError("no CUDA device visible: libpitt_b200 has no CPU fallback")
        if stream is None:
            self.handle = self.lib.pitt_create(int(device), int(seed))
        else:
            self.handle = self.lib.pitt_create_on_stream(int(device), int(seed), C.c_void_p(int(stream)))
        if not self.handle:
            raise PittError("pitt_create failed (no usable CUDA device)")

    def set_workers(self, n_workers):
        """helper streams/threads for the independent primitive fits of a frame (0 = all on this ctx's stream)"""
        self._check(self.lib.pitt_set_workers(self.handle, int(n_workers)))

    def close(self):
        if getattr(self, "handle", None):
            self.lib.pitt_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, status, allow=()):
        if status != A.PITT_OK and status not in allow:
            raise PittError(f"status {status}: {self.lib.pitt_last_error(self.handle).decode()}")
        return status

    @property
    def last_device_ms(self):
        return float(self.lib.pitt_last_device_ms(self.handle))

    @property
    def kernel_launches(self):
        return int(self.lib.pitt_kernel_launches(self.handle))

    def synchronize(self):
        self._check(self.lib.pitt_synchronize(self.handle))

    # ---- staging
    def stage(self, xyz, normals=None):
        """xyz: (n,3) or (n,4) float32 host array (numpy) -> Cloud in HBM."""
        xyz = _f32(xyz)
        assert xyz.ndim == 2 and xyz.shape[1] in (3, 4), xyz.shape
        h = C.c_void_p()
        ptr = xyz.ctypes.data if xyz.shape[0] else None
        self._check(self.lib.pitt_stage_cloud(self.handle, ptr, xyz.shape[1] * 4, xyz.shape[0], C.byref(h)))
        c = Cloud(self, h, xyz.shape[0])
        if normals is not None:
            c.set_normals(normals)
        return c

    def stage_host_ptr(self, ptr, stride, n):
        h = C.c_void_p()
        self._check(self.lib.pitt_stage_cloud(self.handle, C.c_void_p(ptr), stride, n, C.byref(h)))
        return Cloud(self, h, n)

    def prefilter(self, raw, params=None):
        """pitt_prefilter_cloud: raw = (n, point_step/4) float32 host array with x,y,z in columns 0..2
        (a PointCloud2 payload). Returns (Cloud in the world frame, info dict)."""
        raw = np.ascontiguousarray(raw, np.float32)
        params = params if params is not None else default_prefilter_params()
        h = C.c_void_p()
        info = A.PrefilterInfo()
        self._check(self.lib.pitt_prefilter_cloud(self.handle, C.c_void_p(raw.ctypes.data), raw.shape[1] * 4, raw.shape[0],
                                                  C.byref(params), C.byref(h), C.byref(info)))
        n = self.lib.pitt_cloud_size(h)
        return Cloud(self, h, n), {k: getattr(info, k) for k, _ in A.PrefilterInfo._fields_}

    def set_blocking_sync(self, enable=True):
        self._check(self.lib.pitt_set_blocking_sync(self.handle, 1 if enable else 0))

    def arm_filter(self, cloud, params):
        """pitt_arm_filter: chained negative CropBoxes (arm_filter_srv.cpp:66-103). Returns (Cloud, removed per box)."""
        h = C.c_void_p()
        removed = (C.c_int32 * 4)()
        self._check(self.lib.pitt_arm_filter(self.handle, cloud.handle, C.byref(params), C.byref(h), removed))
        return Cloud(self, h, self.lib.pitt_cloud_size(h)), list(removed)

    def get_points(self, cloud):
        out = np.zeros((cloud.n, 4), np.float32)
        self._check(self.lib.pitt_get_points(self.handle, cloud.handle, out.ctypes.data_as(A.f32p)))
        return out

    def stage_device(self, d_ptr, n):
        h = C.c_void_p()
        self._check(self.lib.pitt_stage_cloud_device(self.handle, C.c_void_p(d_ptr), n, C.byref(h)))
        return Cloud(self, h, n)

    # ---- normals / kNN
    def estimate_normals(self, cloud, k=50, viewpoint=(0.0, 0.0, 0.0)):
        vp = (C.c_float * 3)(*viewpoint)
        self._check(self.lib.pitt_estimate_normals(self.handle, cloud.handle, int(k), vp))
        out = np.zeros((cloud.n, 4), np.float32)
        self._check(self.lib.pitt_get_normals(self.handle, cloud.handle, out.ctypes.data_as(A.f32p)))
        return out

    def estimate_normals_device(self, cloud, k=50, viewpoint=(0.0, 0.0, 0.0)):
        """pitt_estimate_normals without copying the normals back (they stay attached to the cloud)"""
        vp = (C.c_float * 3)(*viewpoint)
        self._check(self.lib.pitt_estimate_normals(self.handle, cloud.handle, int(k), vp))

    def knn(self, cloud, k):
        idx = np.zeros((cloud.n, k), np.int32)
        sq = np.zeros((cloud.n, k), np.float32)
        self._check(self.lib.pitt_knn(self.handle, cloud.handle, int(k), idx.ctypes.data_as(A.i32p),
                                      sq.ctypes.data_as(A.f32p)))
        return idx, sq

    # ---- sample consensus
    def sac_segment_count_only(self, cloud, params):
        """seg.segment() without copying the inlier list back (count + coefficients only)."""
        n_inl, n_co = C.c_int(0), C.c_int(0)
        co = np.zeros(8, np.float32)
        info = A.SacInfo()
        self._check(self.lib.pitt_sac_segment(self.handle, cloud.handle, C.byref(params), None, 0, C.byref(n_inl),
                                              co.ctypes.data_as(A.f32p), C.byref(n_co), C.byref(info)))
        return {"n_inliers": n_inl.value, "coeffs": co[: n_co.value].copy(), "info": info}

    def argmax_counts_device(self, d_counts, H, d_best):
        self._check(self.lib.pitt_argmax_counts_device(self.handle, C.c_void_p(d_counts), int(H), C.c_void_p(d_best)))

    def sac_finish_device(self, cloud, params, d_samples_all, H_all, d_best, want_inliers=False):
        """winner (device {index, count}) of a hypothesis split -> refined model + final inliers"""
        n_inl, n_co = C.c_int(0), C.c_int(0)
        co = np.zeros(8, np.float32)
        info = A.SacInfo()
        inl = np.empty(max(cloud.n, 1), np.int32) if want_inliers else None
        self._check(self.lib.pitt_sac_finish_device(
            self.handle, cloud.handle, C.byref(params), C.c_void_p(d_samples_all), int(H_all), C.c_void_p(d_best),
            inl.ctypes.data_as(A.i32p) if want_inliers else None, cloud.n if want_inliers else 0, C.byref(n_inl),
            co.ctypes.data_as(A.f32p), C.byref(n_co), C.byref(info)))
        return {"n_inliers": n_inl.value, "coeffs": co[: n_co.value].copy(), "info": info,
                "inliers": inl[: n_inl.value].copy() if want_inliers else None}

    def sac_segment(self, cloud, params):
        n = cloud.n
        inl = np.empty(max(n, 1), np.int32)
        n_inl, n_co = C.c_int(0), C.c_int(0)
        co = np.zeros(8, np.float32)
        info = A.SacInfo()
        self._check(self.lib.pitt_sac_segment(self.handle, cloud.handle, C.byref(params), inl.ctypes.data_as(A.i32p), n,
                                              C.byref(n_inl), co.ctypes.data_as(A.f32p), C.byref(n_co), C.byref(info)))
        return {"inliers": inl[: n_inl.value].copy(), "coeffs": co[: n_co.value].copy(), "info": info}

    def sac_segment_split(self, cloud, params, rank, world, allgather=None, want_inliers=True):
        """pitt_sac_segment_split: one cloud, the hypothesis set split over `world` ranks; `allgather` = an ALLGATHER_FN
        (torch_allgather_callback() under torchrun), None with world == 1"""
        n = cloud.n
        inl = np.empty(max(n, 1), np.int32) if want_inliers else None
        n_inl, n_co = C.c_int(0), C.c_int(0)
        co = np.zeros(8, np.float32)
        info = A.SacInfo()
        cb = allgather if allgather is not None else C.cast(None, ALLGATHER_FN)
        self._check(self.lib.pitt_sac_segment_split(self.handle, cloud.handle, C.byref(params), int(rank), int(world), cb, None,
                                                    inl.ctypes.data_as(A.i32p) if want_inliers else None, n if want_inliers else 0,
                                                    C.byref(n_inl), co.ctypes.data_as(A.f32p), C.byref(n_co), C.byref(info)))
        return {"n_inliers": n_inl.value, "inliers": inl[: n_inl.value].copy() if want_inliers else None,
                "coeffs": co[: n_co.value].copy(), "info": info}

    def sac_segment_host(self, xyz, params, host_ptr=None):
        """fromROSMsg + seg.segment() on a host cloud (n x 4 float32, or n x 3) in one call; host_ptr = address of a pinned
        copy of the same array (the H2D then runs at full PCIe speed and overlaps the scoring)."""
        xyz = np.ascontiguousarray(xyz, np.float32)
        n, stride = xyz.shape[0], xyz.shape[1] * 4
        inl = np.empty(max(n, 1), np.int32)
        n_inl, n_co = C.c_int(0), C.c_int(0)
        co = np.zeros(8, np.float32)
        info = A.SacInfo()
        ptr = C.c_void_p(host_ptr) if host_ptr is not None else xyz.ctypes.data_as(C.c_void_p)
        self._check(self.lib.pitt_sac_segment_host(self.handle, ptr, stride, n, C.byref(params), inl.ctypes.data_as(A.i32p), n,
                                                   C.byref(n_inl), co.ctypes.data_as(A.f32p), C.byref(n_co), C.byref(info)))
        return {"inliers": inl[: n_inl.value].copy(), "coeffs": co[: n_co.value].copy(), "info": info}

    def sac_score(self, cloud, params, samples):
        samples = np.ascontiguousarray(samples, np.int32)
        H = samples.shape[0]
        counts = np.zeros(H, np.int32)
        co = np.zeros((H, 8), np.float32)
        valid = np.zeros(H, np.uint8)
        self._check(self.lib.pitt_sac_score(self.handle, cloud.handle, C.byref(params), samples.ctypes.data_as(A.i32p),
                                            H, counts.ctypes.data_as(A.i32p), co.ctypes.data_as(A.f32p),
                                            valid.ctypes.data_as(C.POINTER(C.c_uint8))))
        return counts, co, valid

    def sac_score_device(self, cloud, params, d_samples, H, d_counts):
        self._check(self.lib.pitt_sac_score_device(self.handle, cloud.handle, C.byref(params), C.c_void_p(d_samples),
                                                   int(H), C.c_void_p(d_counts)))

    def sac_select(self, cloud, params, coeffs):
        co = np.zeros(8, np.float32)
        co[: len(coeffs)] = coeffs
        inl = np.empty(max(cloud.n, 1), np.int32)
        n_inl = C.c_int(0)
        self._check(self.lib.pitt_sac_select(self.handle, cloud.handle, C.byref(params), co.ctypes.data_as(A.f32p),
                                             inl.ctypes.data_as(A.i32p), cloud.n, C.byref(n_inl)))
        return inl[: n_inl.value].copy()

    def sac_refine(self, cloud, params, coeffs, inliers):
        co = np.zeros(8, np.float32)
        co[: len(coeffs)] = coeffs
        inliers = np.ascontiguousarray(inliers, np.int32)
        out = np.zeros(8, np.float32)
        info = A.SacInfo()
        self._check(self.lib.pitt_sac_refine(self.handle, cloud.handle, C.byref(params), co.ctypes.data_as(A.f32p),
                                             inliers.ctypes.data_as(A.i32p), len(inliers),
                                             out.ctypes.data_as(A.f32p), C.byref(info)))
        return out[: A.N_COEFFS[params.model]].copy(), info

    def pcl_sample_stream(self, cloud, model, count):
        out = np.zeros((count, A.SAMPLE_SIZE[model]), np.int32)
        self._check(self.lib.pitt_pcl_sample_stream(self.handle, cloud.handle, int(model), int(count),
                                                    out.ctypes.data_as(A.i32p)))
        return out

    # ---- clustering
    def euclidean_clusters(self, cloud, tolerance, min_size, max_size):
        labels = np.full(cloud.n, -1, np.int32)
        nc = C.c_int(0)
        self._check(self.lib.pitt_euclidean_clusters(self.handle, cloud.handle, float(tolerance), int(min_size),
                                                     int(max_size), labels.ctypes.data_as(A.i32p), C.byref(nc)))
        return labels, nc.value

    # ---- measurement
    def fp32_peak(self, kind):
        t = C.c_double(0)
        self._check(self.lib.pitt_fp32_peak(self.handle, int(kind), C.byref(t)))
        return t.value


# ---------------------------------------------------------------- service-shaped entry points
from . import _results as R  # noqa: E402


def _find_supports(self, cloud, params=None, supports_cap=4):
    """findSupports (supports_segmentation_srv.cpp:241-361) on a staged cloud."""
    params = params if params is not None else default_support_params()
    b = R.SupportBuffers(cloud.n, supports_cap)
    self._check(self.lib.pitt_find_supports(self.handle, cloud.handle, C.byref(params), C.byref(b.res)))
    return b.to_python()


def _cluster_service(self, cloud, params=None):
    """clusterize (cluster_segmentation_srv.cpp:38-108)."""
    params = params if params is not None else default_cluster_params()
    b = R.ClusterBuffers(cloud.n)
    self._check(self.lib.pitt_cluster_service(self.handle, cloud.handle, C.byref(params), C.byref(b.res)))
    return b.to_python()


def _primitive_service(self, cloud, params):
    """ransac{Plane,Sphere,Cylinder,Cone}Detection (…_segmentation_srv.cpp)."""
    b = R.PrimitiveBuffers(cloud.n)
    self._check(self.lib.pitt_primitive_service(self.handle, cloud.handle, C.byref(params), C.byref(b.res)))
    return b.to_python()


def _segment_frame(self, cloud, params=None, shapes_cap=64):
    """depthAcquisition + clustersAcquisition from the world-frame cloud on."""
    params = params if params is not None else default_frame_params()
    b = R.FrameBuffers(shapes_cap)
    self._check(self.lib.pitt_segment_frame(self.handle, cloud.handle, C.byref(params), C.byref(b.res)))
    return b.to_python()


Context.find_supports = _find_supports
Context.cluster_service = _cluster_service
Context.primitive_service = _primitive_service
Context.segment_frame = _segment_frame


def default_arm_filter_params():
    p = A.ArmFilterParams()
    load_library().pitt_default_arm_filter_params(C.byref(p))
    return p


def default_prefilter_params():
    p = A.PrefilterParams()
    load_library().pitt_default_prefilter_params(C.byref(p))
    return p


def segment_clouds_batched(contexts, clouds, params=None, shapes_cap=64, bufs=None):
    """pitt_segment_clouds_batched: the frame stream over clouds that are already staged in HBM. `bufs` (a list of
    FrameBuffers, one per cloud) can be passed to keep allocations out of a timed region."""
    params = params if params is not None else default_frame_params()
    lib = load_library()
    n = len(clouds)
    bufs = bufs if bufs is not None else [R.FrameBuffers(shapes_cap) for _ in range(n)]
    res = (A.FrameResult * n)()
    for i, b in enumerate(bufs):
        res[i] = b.res
    handles = (C.c_void_p * n)(*[c.handle for c in clouds])
    ctxs = (C.c_void_p * len(contexts))(*[c.handle for c in contexts])
    st = lib.pitt_segment_clouds_batched(ctxs, len(contexts), handles, n, C.byref(params), res)
    if st != A.PITT_OK:
        raise PittError(f"pitt_segment_clouds_batched status {st}")
    for i, b in enumerate(bufs):
        b.res = res[i]
    return R.FrameResults(bufs)


def segment_frames_batched(contexts, frames, params=None, shapes_cap=64, prefilter=None):
    """pitt_segment_frames_batched: `frames` is a list of (n,4) float32 host arrays (ideally pinned);
    one host thread per context drives its stream. Returns the per-frame result dicts."""
    params = params if params is not None else default_frame_params()
    lib = load_library()
    n = len(frames)
    frames = [np.ascontiguousarray(f, np.float32) for f in frames]
    bufs = [R.FrameBuffers(shapes_cap) for _ in range(n)]
    res = (A.FrameResult * n)()
    for i, b in enumerate(bufs):
        res[i] = b.res
    ptrs = (C.c_void_p * n)(*[f.ctypes.data for f in frames])
    counts = np.array([f.shape[0] for f in frames], np.int32)
    ctxs = (C.c_void_p * len(contexts))(*[c.handle for c in contexts])
    stride = int(frames[0].shape[1]) * 4 if n else 16
    st = lib.pitt_segment_raw_frames_batched(ctxs, len(contexts), ptrs, counts.ctypes.data_as(A.i32p), stride, n,
                                             C.byref(prefilter) if prefilter is not None else None, C.byref(params), res)
    if st != A.PITT_OK:
        raise PittError(f"pitt_segment_frames_batched status {st}")
    for i, b in enumerate(bufs):
        b.res = res[i]
    return R.FrameResults(bufs)
