"""Seeded synthetic inputs for the segmentation hot path.

All generators are deterministic functions of their seed (numpy Philox, a counter-based PRNG) and
return float32 clouds of shape (n, 4) = {x, y, z, 1.0}, the pcl::PointXYZ layout.
"""
import numpy as np


def _rng(seed):
    return np.random.Generator(np.random.Philox(key=int(seed)))


def _pack(xyz):
    out = np.ones((xyz.shape[0], 4), np.float32)
    out[:, :3] = xyz.astype(np.float32)
    return out


def plane_outlier_cloud(n=1_000_000, seed=12345, plane_frac=0.7, sigma=0.002):
    """Config C2/C5: plane z=0 over 2x2 m with Gaussian noise + uniform outliers in 2x2x1 m, shuffled."""
    r = _rng(seed)
    n_pl = int(round(n * plane_frac))
    pl = np.empty((n_pl, 3))
    pl[:, 0:2] = r.uniform(-1.0, 1.0, (n_pl, 2))
    pl[:, 2] = r.normal(0.0, sigma, n_pl)
    out = np.empty((n - n_pl, 3))
    out[:, 0:2] = r.uniform(-1.0, 1.0, (n - n_pl, 2))
    out[:, 2] = r.uniform(0.0, 1.0, n - n_pl)
    xyz = np.concatenate([pl, out], 0)
    xyz = xyz[r.permutation(n)]
    return _pack(xyz)


def _ray_sphere(o, d, c, rad):
    oc = o - c
    b = d @ oc
    cc = oc @ oc - rad * rad
    disc = b * b - cc
    t = np.where(disc > 0, -b - np.sqrt(np.maximum(disc, 0)), np.inf)
    return np.where(t > 1e-6, t, np.inf)


def _ray_cyl_z(o, d, c, rad, h):
    """finite cylinder, axis +z from c (base centre) to c+h, with top cap"""
    ox, oy = o[0] - c[0], o[1] - c[1]
    a = d[:, 0] ** 2 + d[:, 1] ** 2
    b = ox * d[:, 0] + oy * d[:, 1]
    cc = ox * ox + oy * oy - rad * rad
    disc = b * b - a * cc
    with np.errstate(divide="ignore", invalid="ignore"):
        t = (-b - np.sqrt(np.maximum(disc, 0))) / a
    z = o[2] + t * d[:, 2]
    t = np.where((disc > 0) & (t > 1e-6) & (z >= c[2]) & (z <= c[2] + h), t, np.inf)
    # top cap
    with np.errstate(divide="ignore", invalid="ignore"):
        tc = (c[2] + h - o[2]) / d[:, 2]
    px, py = o[0] + tc * d[:, 0] - c[0], o[1] + tc * d[:, 1] - c[1]
    tc = np.where((tc > 1e-6) & (px * px + py * py <= rad * rad), tc, np.inf)
    return np.minimum(t, tc)


def _ray_cone_z(o, d, apex, half_angle, h):
    """cone with apex up at `apex`, opening downwards along -z, height h"""
    k = np.tan(half_angle) ** 2
    ox, oy, oz = o[0] - apex[0], o[1] - apex[1], o[2] - apex[2]
    a = d[:, 0] ** 2 + d[:, 1] ** 2 - k * d[:, 2] ** 2
    b = ox * d[:, 0] + oy * d[:, 1] - k * oz * d[:, 2]
    cc = ox * ox + oy * oy - k * oz * oz
    disc = b * b - a * cc
    sq = np.sqrt(np.maximum(disc, 0))
    best = np.full(d.shape[0], np.inf)
    with np.errstate(divide="ignore", invalid="ignore"):
        for t in ((-b - sq) / a, (-b + sq) / a):
            z = oz + t * d[:, 2]
            ok = (disc > 0) & (t > 1e-6) & (z <= 0) & (z >= -h)
            best = np.where(ok & (t < best), t, best)
    return best


DEFAULT_OBJECTS = (
    ("sphere", (-0.25, 0.05), 0.06),
    ("cylinder", (0.0, -0.02), 0.04, 0.15),
    ("cone", (0.25, 0.05), np.deg2rad(20.0), 0.15),
)


def tabletop_frame(seed=12345, width=640, height=480, objects=DEFAULT_OBJECTS, noise=0.0015, jitter=2e-7,
                   random_poses=False):
    """Config C1/C4: Kinect-shaped frame of a table (plane z=0, world z up) with a sphere, a cylinder
    and a cone standing on it. Pinhole fx=fy=525, cx=319.5, cy=239.5 (scaled with the resolution),
    camera 0.9 m above the table pitched 40 deg down. Returns the cloud in the WORLD frame, which is
    what obj_segmentation hands to the hot path after pcl::transformPointCloud
    (obj_segmentation.cpp:248). Row-major pixel order, every ray hits (no NaN)."""
    r = _rng(seed)
    sx = width / 640.0
    fx = fy = 525.0 * sx
    cx, cy = (width - 1) / 2.0, (height - 1) / 2.0
    u, v = np.meshgrid(np.arange(width), np.arange(height))
    dc = np.stack([(u.ravel() - cx) / fx, (v.ravel() - cy) / fy, np.ones(width * height)], 1)
    dc /= np.linalg.norm(dc, axis=1, keepdims=True)
    pitch = np.deg2rad(40.0)
    # camera axes in world: x_c -> +x, y_c (image down) and z_c (forward) pitched down
    fwd = np.array([0.0, np.cos(pitch), -np.sin(pitch)])
    right = np.array([1.0, 0.0, 0.0])
    down = np.cross(fwd, right)
    Rm = np.stack([right, down, fwd], 1)  # world = Rm @ cam
    d = dc @ Rm.T
    o = np.array([0.0, -0.9 / np.tan(pitch) * 0.75, 0.9])
    objs = list(objects)
    if random_poses:
        offs = r.uniform(-0.04, 0.04, (len(objs), 2))
        objs = [(ob[0], (ob[1][0] + offs[i, 0], ob[1][1] + offs[i, 1])) + tuple(ob[2:]) for i, ob in enumerate(objs)]
    with np.errstate(divide="ignore", invalid="ignore"):
        t = np.where(d[:, 2] < 0, -o[2] / d[:, 2], np.inf)
    for ob in objs:
        kind, (px, py) = ob[0], ob[1]
        if kind == "sphere":
            t = np.minimum(t, _ray_sphere(o, d, np.array([px, py, ob[2]]), ob[2]))
        elif kind == "cylinder":
            t = np.minimum(t, _ray_cyl_z(o, d, np.array([px, py, 0.0]), ob[2], ob[3]))
        elif kind == "cone":
            t = np.minimum(t, _ray_cone_z(o, d, np.array([px, py, ob[3]]), ob[2], ob[3]))
    t = np.where(np.isfinite(t), t, 5.0)
    t = t + r.normal(0.0, noise, t.shape)  # range noise along the ray
    xyz = o[None, :] + t[:, None] * d
    xyz = xyz + r.uniform(-jitter, jitter, xyz.shape)  # sub-micron jitter: breaks exact kNN ties
    return _pack(xyz)


def camera_pose():
    """(4x4 float32 camera->world matrix, 4x4 world->camera) of the synthetic Kinect of tabletop_frame"""
    pitch = np.deg2rad(40.0)
    fwd = np.array([0.0, np.cos(pitch), -np.sin(pitch)])
    right = np.array([1.0, 0.0, 0.0])
    down = np.cross(fwd, right)
    Rm = np.stack([right, down, fwd], 1)
    o = np.array([0.0, -0.9 / np.tan(pitch) * 0.75, 0.9])
    M = np.eye(4)
    M[:3, :3] = Rm
    M[:3, 3] = o
    return M.astype(np.float32), np.linalg.inv(M).astype(np.float32)


def raw_camera_frame(seed=12345, width=640, height=480, point_step=16, nan_fraction=0.02, far_fraction=0.01, **kw):
    """What the sensor publishes (the input of depthAcquisition, obj_segmentation.cpp:233): the tabletop
    frame in the CAMERA frame as a PointCloud2-like payload of `point_step` bytes per point (x,y,z float32
    at 0,4,8; the rest is colour/padding), with some NaN returns and some points beyond the deep threshold."""
    world = tabletop_frame(seed=seed, width=width, height=height, **kw)
    _, w2c = camera_pose()
    cam = (world[:, :3].astype(np.float64) @ w2c[:3, :3].astype(np.float64).T + w2c[:3, 3].astype(np.float64)).astype(np.float32)
    r = _rng(seed + 977)
    n = cam.shape[0]
    far = r.random(n) < far_fraction
    cam[far, 2] += np.float32(3.5)  # background returns
    bad = r.random(n) < nan_fraction
    cam[bad] = np.nan
    raw = np.zeros((n, point_step // 4), np.float32)
    raw[:, :3] = cam
    if point_step >= 16:
        raw[:, 3] = 1.0
    return raw


def voxel_downsample(xyz4, leaf=0.01):
    """pcl::VoxelGrid (centroid per occupied voxel, output ordered by voxel index) — host helper used
    to build the *faithful* C1 variant (PCManager::downSampling, pc_manager.cpp:55-67)."""
    p = xyz4[:, :3].astype(np.float64)
    mn = np.floor(p.min(0) / leaf).astype(np.int64)
    ijk = np.floor(p / leaf).astype(np.int64) - mn
    dims = ijk.max(0) + 1
    key = ijk[:, 0] + dims[0] * (ijk[:, 1] + dims[1] * ijk[:, 2])
    order = np.argsort(key, kind="stable")
    key_s = key[order]
    uniq, start, cnt = np.unique(key_s, return_index=True, return_counts=True)
    sums = np.add.reduceat(xyz4[order, :3].astype(np.float32), start, axis=0)
    return _pack(sums / cnt[:, None].astype(np.float32))


def primitive_cluster(kind, n, seed, sigma=0.001):
    """Config C3: one object cluster seen from one side (half the surface), n points."""
    r = _rng(seed)
    if kind == "cylinder":
        rad, h = r.uniform(0.02, 0.06), r.uniform(0.1, 0.3)
        th = r.uniform(-0.5 * np.pi, 0.5 * np.pi, n)
        z = r.uniform(0, h, n)
        rr = rad + r.normal(0, sigma, n)
        xyz = np.stack([rr * np.cos(th), rr * np.sin(th), z], 1)
        truth = dict(kind=kind, radius=rad, height=h)
    elif kind == "cone":
        ha, h = np.deg2rad(r.uniform(15.0, 30.0)), r.uniform(0.1, 0.2)
        th = r.uniform(-0.5 * np.pi, 0.5 * np.pi, n)
        s = np.sqrt(r.uniform(0.01, 1.0, n)) * h  # distance below the apex, area-uniform
        rr = s * np.tan(ha) + r.normal(0, sigma, n)
        xyz = np.stack([rr * np.cos(th), rr * np.sin(th), h - s], 1)
        truth = dict(kind=kind, half_angle=ha, height=h)
    elif kind == "sphere":
        rad = r.uniform(0.03, 0.08)
        v = r.normal(size=(n, 3))
        v[:, 0] = np.abs(v[:, 0])
        v /= np.linalg.norm(v, axis=1, keepdims=True)
        xyz = v * (rad + r.normal(0, sigma, n))[:, None] + np.array([0, 0, rad])
        truth = dict(kind=kind, radius=rad)
    else:  # plane patch (box face)
        xyz = np.stack([r.uniform(-0.05, 0.05, n), r.normal(0, sigma, n), r.uniform(0, 0.12, n)], 1)
        truth = dict(kind="plane")
    # random rigid pose (small tilt) + offset so that nothing is axis aligned
    ax = r.normal(size=3)
    ax /= np.linalg.norm(ax)
    ang = r.uniform(0, 0.3)
    K = np.array([[0, -ax[2], ax[1]], [ax[2], 0, -ax[0]], [-ax[1], ax[0], 0]])
    Rm = np.eye(3) + np.sin(ang) * K + (1 - np.cos(ang)) * K @ K
    xyz = xyz @ Rm.T + r.uniform(-0.3, 0.3, 3) + np.array([0.0, 0.8, 0.2])
    truth["R"] = Rm
    return _pack(xyz), truth
