"""CTA timeline of knn_collect_kernel on the 307 200-point frame (pitt_debug_knn_timeline): where the tail comes from."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pitt_object_table_segmentation_b200 as pkg
from pitt_object_table_segmentation_b200 import scenes

ctx = pkg.Context(0)
lib = pkg.load_library()
xyz = scenes.tabletop_frame(seed=int(sys.argv[1]) if len(sys.argv) > 1 else 12345)
cloud = ctx.stage(xyz)
for _ in range(3):
    ctx.estimate_normals_device(cloud, 50)
print("normals ms (not recording)", ctx.last_device_ms)
lib.pitt_debug_knn_timeline(ctx.handle, 1, None, 0)
for _ in range(2):
    ctx.estimate_normals_device(cloud, 50)
buf = np.zeros(6 * 8192, np.uint64)
n = lib.pitt_debug_knn_timeline(ctx.handle, 0, buf.ctypes.data_as(C.POINTER(C.c_uint64)), 8192)
t = buf[: 6 * n].reshape(n, 6).astype(np.int64)
t0 = t[:, 0].min()
start, end, sm = (t[:, 0] - t0) * 1e-3, (t[:, 1] - t0) * 1e-3, t[:, 2]
dur = end - start
print(f"CTAs {n}  span {end.max():.1f} us  CTA duration us: mean {dur.mean():.1f} p50 {np.median(dur):.1f} p90 {np.percentile(dur, 90):.1f} "
      f"p99 {np.percentile(dur, 99):.1f} max {dur.max():.1f}")
nsm = int(sm.max()) + 1
last_end = np.array([end[sm == s].max() if np.any(sm == s) else 0 for s in range(nsm)])
print(f"SMs {nsm}: last CTA of an SM ends at us: min {last_end.min():.1f} p50 {np.median(last_end):.1f} max {last_end.max():.1f}")
order = np.argsort(start)
print("start time of the CTA launched last: %.1f us" % start.max())
for lo in range(0, 100, 10):
    sel = (start >= np.percentile(start, lo)) & (start <= np.percentile(start, min(100, lo + 10)))
    print(f"  CTAs started in [{np.percentile(start, lo):6.1f}, {np.percentile(start, min(100, lo + 10)):6.1f}] us: mean duration {dur[sel].mean():6.1f}")
# resident CTAs over time
ts = np.linspace(0, end.max(), 21)
print("resident CTAs at", " ".join(f"{x:.0f}us:{int(np.sum((start <= x) & (end > x)))}" for x in ts))
slow = np.argsort(-dur)[:12]
print("slowest CTAs (block, start, duration, candidates, passes, handovers):", [(int(b), round(float(start[b]), 1), round(float(dur[b]), 1), int(t[b, 3]), int(t[b, 4]), int(t[b, 5])) for b in slow])
print("all CTAs: mean candidates %.0f passes %.1f handovers %.2f; corr(duration, candidates) %.3f corr(duration, passes) %.3f" % (t[:, 3].mean(), t[:, 4].mean(), t[:, 5].mean(), np.corrcoef(dur, t[:, 3])[0, 1], np.corrcoef(dur, t[:, 4])[0, 1]))
early = start < 1.0
A = np.stack([t[early, 3], t[early, 4], t[early, 5], np.ones(early.sum())], 1).astype(float)
coef = np.linalg.lstsq(A, dur[early], rcond=None)[0]
print("first-wave CTAs: duration ~ %.4f us/candidate + %.3f us/pass + %.2f us/handover + %.1f" % tuple(coef))
blk = np.arange(n)
for lo in range(0, n, max(1, n // 12)):
    hi = min(n, lo + max(1, n // 12))
    print(f"  blocks [{lo},{hi}): mean duration {dur[lo:hi].mean():6.1f} start {start[lo:hi].mean():6.1f}")
