"""Regenerates tests/golden/*.npz from the CPU oracle (run from the repo root: python tests/golden/make_golden.py).

The reference has no golden vectors and cannot be built here, so these fixtures are the
ORACLE's outputs on small seeded inputs ("parity unpinned" against the reference itself). They pin the
oracle against silent drift and give the GPU tests a second, file-based comparison target.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import orc_binding as O  # noqa: E402
from pitt_object_table_segmentation_b200 import _abi as A, scenes  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def info_vec(i):
    return np.array([i.iterations, i.skipped, i.hypotheses, i.best_hypothesis, i.best_count, i.n_inliers_model,
                     i.lm_info, i.lm_nfev], np.int32)


def main():
    # 1. table-plane RANSAC (config 2 shape, small)
    xyz = scenes.plane_outlier_cloud(4000, seed=12345)
    p = O.default_support_sac_params()
    samples = O.pcl_sample_stream(xyz, A.MODEL_PLANE, 64)
    counts, co, valid = O.sac_score(xyz, None, p, samples)
    seg = O.sac_segment(xyz, None, p)
    np.savez_compressed(os.path.join(OUT, "plane_c2_small.npz"), xyz=xyz, samples=samples, counts=counts, coeffs=co,
                        valid=valid, seg_inliers=seg["inliers"], seg_coeffs=seg["coeffs"], seg_info=info_vec(seg["info"]))
    # 2. primitives with normals
    for kind, model in (("sphere", A.MODEL_SPHERE), ("cylinder", A.MODEL_CYLINDER), ("cone", A.MODEL_CONE)):
        xyz, _ = scenes.primitive_cluster(kind, 1500, 12345)
        nrm = O.estimate_normals(xyz, 50)
        pm = O.default_sac_params(model)
        samples = O.pcl_sample_stream(xyz, model, 48)
        counts, co, valid = O.sac_score(xyz, nrm, pm, samples)
        seg = O.sac_segment(xyz, nrm, pm)
        srv = O.primitive_service(xyz, nrm, pm)
        np.savez_compressed(os.path.join(OUT, f"primitive_{kind}.npz"), xyz=xyz, normals=nrm, samples=samples,
                            counts=counts, coeffs=co, valid=valid, seg_inliers=seg["inliers"], seg_coeffs=seg["coeffs"],
                            seg_info=info_vec(seg["info"]), srv_inliers=srv["inliers"], srv_coeffs=srv["coefficients"],
                            srv_centroid=srv["centroid"])
    # 3. tabletop frame (config 1 shape, 160x120)
    xyz = scenes.tabletop_frame(seed=12345, width=160, height=120)
    nrm = O.estimate_normals(xyz, 50)
    knn_idx, _ = O.knn(xyz[:2000], 8)
    sup = O.find_supports(xyz, nrm, O.default_support_params())
    on = sup["supports"][0]["on_support_cloud"]
    labels, nc = O.euclidean_clusters(on, 0.03, int(round(len(on) * 0.01)), int(round(len(on) * 0.99)))
    fr = O.segment_frame(xyz, O.default_frame_params())
    np.savez_compressed(os.path.join(OUT, "tabletop_160x120.npz"), xyz=xyz, normals=nrm, knn_idx_first2000_k8=knn_idx,
                        support_coeffs=sup["supports"][0]["coefficients"], support_map=sup["supports"][0]["inliers"],
                        on_support=on, cluster_labels=labels, n_clusters=np.int32(nc),
                        shape_tags=np.array([s["tag"] for s in fr["shapes"]], np.int32),
                        shape_coeffs=np.array([np.pad(s["coefficients"], (0, 8 - len(s["coefficients"]))) for s in fr["shapes"]], np.float32),
                        shape_inliers=np.array([s["inliers"] for s in fr["shapes"]], np.int32),
                        shape_est_centroid=np.array([s["est_centroid"] for s in fr["shapes"]], np.float32))
    print("golden fixtures written to", OUT)


if __name__ == "__main__":
    main()
