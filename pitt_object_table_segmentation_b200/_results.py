"""Allocation of the caller-owned result buffers of the service-shaped C ABI calls and their
conversion to plain Python/numpy values. Pure marshalling, shared by the product binding (api.py)
and the oracle binding (oracle/orc_binding.py) because both speak the structs of pitt_b200.h."""
import ctypes as C

import numpy as np

from . import _abi as A


class SupportBuffers:
    def __init__(self, n0, supports_cap=4):
        self.n0 = n0
        self.sup = (A.Support * supports_cap)()
        self.maps = np.zeros(max(1, supports_cap * n0), np.int32)
        self.points = np.zeros((max(1, supports_cap * 2 * n0), 4), np.float32)
        self.res = A.SupportResult()
        self.res.supports = C.cast(self.sup, C.POINTER(A.Support))
        self.res.supports_cap = supports_cap
        self.res.maps = self.maps.ctypes.data_as(A.i32p)
        self.res.maps_cap = self.maps.size
        self.res.points = self.points.ctypes.data_as(A.f32p)
        self.res.points_cap = self.points.shape[0]

    def to_python(self):
        r = self.res
        out = {"loop_trips": r.loop_trips, "supports": [],
               "used": {"min_cloud": r.used_min_iterative_cloud_percentual_size,
                        "min_plane": r.used_min_iterative_plane_percentual_size,
                        "max_var": r.used_max_variance_threshold_for_horizontal,
                        "min_var": r.used_min_variance_threshold_for_horizontal,
                        "max_iter": r.used_ransac_max_iteration_threshold,
                        "thr": r.used_ransac_distance_point_in_shape_threshold,
                        "w": r.used_ransac_model_normal_distance_weigth,
                        "axis": list(r.used_horizontal_axis), "offset": list(r.used_support_edge_remove_offset)}}
        for i in range(r.n_supports):
            s = self.sup[i]
            out["supports"].append({
                "coefficients": np.array([s.a, s.b, s.c, s.d], np.float32),
                "inliers": self.maps[s.map_offset: s.map_offset + s.n_map].copy(),
                "support_cloud": self.points[s.support_offset: s.support_offset + s.n_support].copy(),
                "on_support_cloud": self.points[s.on_support_offset: s.on_support_offset + s.n_on_support].copy(),
            })
        return out


class ClusterBuffers:
    def __init__(self, n, clusters_cap=256):
        self.cl = (A.Cluster * clusters_cap)()
        self.idx = np.zeros(max(1, n), np.int32)
        self.res = A.ClustersResult()
        self.res.clusters = C.cast(self.cl, C.POINTER(A.Cluster))
        self.res.clusters_cap = clusters_cap
        self.res.indices = self.idx.ctypes.data_as(A.i32p)
        self.res.indices_cap = self.idx.size

    def to_python(self):
        out = []
        for i in range(self.res.n_clusters):
            c = self.cl[i]
            out.append({"inliers": self.idx[c.offset: c.offset + c.n].copy(),
                        "centroid": np.array([c.x_centroid, c.y_centroid, c.z_centroid], np.float32)})
        return out


class PrimitiveBuffers:
    def __init__(self, n):
        self.inl = np.zeros(max(1, n), np.int32)
        self.res = A.PrimitiveResult()
        self.res.inliers = self.inl.ctypes.data_as(A.i32p)
        self.res.inliers_cap = self.inl.size

    def to_python(self):
        r = self.res
        return {"inliers": self.inl[: r.n_inliers].copy(),
                "coefficients": np.array(r.coefficients[: r.n_coefficients], np.float32),
                "centroid": np.array([r.x_centroid, r.y_centroid, r.z_centroid], np.float32),
                "centroid_valid": bool(r.centroid_valid), "info": r.info}


class FrameBuffers:
    def __init__(self, shapes_cap=64):
        self.shapes = (A.TrackedShape * shapes_cap)()
        self.res = A.FrameResult()
        self.res.shapes = C.cast(self.shapes, C.POINTER(A.TrackedShape))
        self.res.shapes_cap = shapes_cap

    def to_python(self):
        r = self.res
        shapes = []
        for i in range(min(r.n_shapes, r.shapes_cap)):
            t = self.shapes[i]
            shapes.append({"object_id": t.object_id, "tag": t.shape_tag, "tag_name": A.TAG_NAMES[t.shape_tag],
                           "pc_centroid": np.array([t.x_pc_centroid, t.y_pc_centroid, t.z_pc_centroid], np.float32),
                           "est_centroid": np.array([t.x_est_centroid, t.y_est_centroid, t.z_est_centroid], np.float32),
                           "coefficients": np.array(t.coefficients[: t.n_coefficients], np.float32),
                           "n_points": t.n_points,
                           "inliers": (t.inl_plane, t.inl_sphere, t.inl_cylinder, t.inl_cone)})
        ns = min(r.n_supports, 8)
        return {"n_supports": r.n_supports, "n_clusters": r.n_clusters, "shapes": shapes,
                "support_coefficients": np.array(r.support_coefficients[: 4 * ns], np.float32).reshape(ns, 4),
                "support_sizes": list(r.support_sizes[:ns]), "on_support_sizes": list(r.on_support_sizes[:ns]),
                "device_ms": r.device_ms,
                "debug_words": [int(np.float32(v).view(np.uint32)) for v in r.support_coefficients[16:24]]}


class FrameResults:
    """The responses of a batched call: a read-only sequence over the caller-owned result structs the C ABI has filled. The
    dict view of a frame (FrameBuffers.to_python) is built when the frame is first looked at, not for all frames up front."""

    def __init__(self, bufs):
        self._bufs = list(bufs)
        self._views = [None] * len(self._bufs)

    def __len__(self):
        return len(self._bufs)

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[j] for j in range(*i.indices(len(self)))]
        if i < 0:
            i += len(self)
        if self._views[i] is None:
            self._views[i] = self._bufs[i].to_python()
        return self._views[i]

    def __iter__(self):
        return (self[i] for i in range(len(self)))
