"""ctypes mirrors of the POD structs in include/pitt_b200.h (keep in the same order as the header)."""
import ctypes as C

PITT_OK, PITT_ERR_INVALID, PITT_ERR_CUDA, PITT_ERR_CAPACITY, PITT_ERR_STATE = 0, 1, 2, 3, 4
MODEL_PLANE, MODEL_SPHERE, MODEL_CYLINDER, MODEL_CONE = 0, 1, 2, 3
SAMPLER_PCL_MT19937, SAMPLER_PHILOX, SAMPLER_REPLAY = 0, 1, 2
STOP_PCL_ADAPTIVE, STOP_ALL_H = 0, 1
TAG_UNKNOWN, TAG_PLANE, TAG_SPHERE, TAG_CONE, TAG_CYLINDER = 0, 1, 2, 3, 4
TAG_NAMES = {0: "unknown", 1: "plane", 2: "sphere", 3: "cone", 4: "cylinder"}
MODEL_NAMES = {0: "plane", 1: "sphere", 2: "cylinder", 3: "cone"}
SAMPLE_SIZE = {0: 3, 1: 4, 2: 2, 3: 3}
N_COEFFS = {0: 4, 1: 4, 2: 7, 3: 7}

i32, i64, f32, f64 = C.c_int32, C.c_int64, C.c_float, C.c_double
i32p, f32p = C.POINTER(C.c_int32), C.POINTER(C.c_float)


class SacParams(C.Structure):
    _fields_ = [
        ("model", i32), ("max_iterations", i32),
        ("distance_threshold", f64), ("probability", f64), ("normal_distance_weight", f64),
        ("radius_min", f64), ("radius_max", f64), ("min_angle", f64), ("max_angle", f64),
        ("eps_angle", f64), ("axis", f32 * 3),
        ("optimize", i32), ("sampler", i32), ("stop", i32),
        ("replay_samples", i32p), ("replay_count", i32), ("reserved", i32),
    ]

    def copy(self):
        other = SacParams()
        C.memmove(C.byref(other), C.byref(self), C.sizeof(SacParams))
        return other


class SacInfo(C.Structure):
    _fields_ = [
        ("iterations", i32), ("skipped", i32), ("hypotheses", i32), ("best_hypothesis", i32),
        ("best_count", i32), ("n_inliers_model", i32), ("lm_info", i32), ("lm_nfev", i32),
        ("model_coeffs", f32 * 8), ("device_ms", f64),
    ]


class SupportParams(C.Structure):
    _fields_ = [
        ("min_iterative_cloud_percentual_size", f32), ("min_iterative_plane_percentual_size", f32),
        ("variance_threshold_for_horizontal", f32), ("ransac_distance_point_in_shape_threshold", f32),
        ("ransac_model_normal_distance_weigth", f32), ("ransac_max_iteration_threshold", i32),
        ("horizontal_axis_len", i32), ("horizontal_axis", f32 * 3),
        ("support_edge_remove_offset_len", i32), ("support_edge_remove_offset", f32 * 3),
        ("normals_k", i32), ("compute_discarded_normals", i32),
    ]


class CropBox(C.Structure):
    _fields_ = [("min_pt", f32 * 3), ("max_pt", f32 * 3), ("translation", f32 * 3), ("rotation_rpy", f32 * 3)]


class ArmFilterParams(C.Structure):
    _fields_ = [("n_boxes", i32), ("input_is_dense", i32), ("box", CropBox * 4)]


class Support(C.Structure):
    _fields_ = [
        ("n_map", i32), ("n_support", i32), ("n_on_support", i32),
        ("a", f32), ("b", f32), ("c", f32), ("d", f32),
        ("map_offset", i64), ("support_offset", i64), ("on_support_offset", i64),
    ]


class SupportResult(C.Structure):
    _fields_ = [
        ("n_supports", i32), ("loop_trips", i32),
        ("used_min_iterative_cloud_percentual_size", f32), ("used_min_iterative_plane_percentual_size", f32),
        ("used_max_variance_threshold_for_horizontal", f32), ("used_min_variance_threshold_for_horizontal", f32),
        ("used_ransac_max_iteration_threshold", i32),
        ("used_ransac_distance_point_in_shape_threshold", f32), ("used_ransac_model_normal_distance_weigth", f32),
        ("used_horizontal_axis", f32 * 3), ("used_support_edge_remove_offset", f32 * 3),
        ("supports", C.POINTER(Support)), ("supports_cap", i32),
        ("maps", i32p), ("maps_cap", i64), ("maps_used", i64),
        ("points", f32p), ("points_cap", i64), ("points_used", i64),
    ]


class ClusterParams(C.Structure):
    _fields_ = [("tolerance", f64), ("min_rate", f64), ("max_rate", f64), ("min_input_size", i32), ("reserved", i32)]


class Cluster(C.Structure):
    _fields_ = [("n", i32), ("offset", i32), ("x_centroid", f32), ("y_centroid", f32), ("z_centroid", f32)]


class ClustersResult(C.Structure):
    _fields_ = [
        ("n_clusters", i32), ("clusters", C.POINTER(Cluster)), ("clusters_cap", i32),
        ("indices", i32p), ("indices_cap", i32), ("indices_used", i32),
    ]


class PrimitiveResult(C.Structure):
    _fields_ = [
        ("n_inliers", i32), ("n_coefficients", i32), ("coefficients", f32 * 8),
        ("x_centroid", f32), ("y_centroid", f32), ("z_centroid", f32), ("centroid_valid", i32),
        ("inliers", i32p), ("inliers_cap", i32), ("info", SacInfo),
    ]


class TrackedShape(C.Structure):
    _fields_ = [
        ("object_id", i32), ("shape_tag", i32),
        ("x_pc_centroid", f32), ("y_pc_centroid", f32), ("z_pc_centroid", f32),
        ("x_est_centroid", f32), ("y_est_centroid", f32), ("z_est_centroid", f32),
        ("n_coefficients", i32), ("coefficients", f32 * 8), ("n_points", i32),
        ("inl_plane", i32), ("inl_sphere", i32), ("inl_cylinder", i32), ("inl_cone", i32),
    ]


class FrameParams(C.Structure):
    _fields_ = [
        ("support", SupportParams), ("cluster", ClusterParams),
        ("plane", SacParams), ("sphere", SacParams), ("cylinder", SacParams), ("cone", SacParams),
        ("normals_k", i32), ("min_points", i32), ("viewpoint", f32 * 3),
        ("cone_over_cylinder_priority", f32),
    ]


class PrefilterParams(C.Structure):
    _fields_ = [("leaf", f32 * 3), ("apply_deep_filter", i32), ("deep_threshold", f32), ("apply_transform", i32),
                ("transform", f32 * 16)]


class PrefilterInfo(C.Structure):
    _fields_ = [("n_input", i32), ("n_voxel", i32), ("n_closer", i32), ("n_further", i32), ("voxel_overflow", i32),
                ("used_deep_threshold", f32)]


class FrameResult(C.Structure):
    _fields_ = [
        ("n_supports", i32), ("n_clusters", i32), ("n_shapes", i32),
        ("shapes", C.POINTER(TrackedShape)), ("shapes_cap", i32),
        ("support_coefficients", f32 * 32), ("support_sizes", i32 * 8), ("on_support_sizes", i32 * 8),
        ("device_ms", f64),
    ]
