"""Development probe for the tensor-core plane path (csrc/plane_tc.cu): numerics of the TMEM accumulators against
the real dot product, count parity against the exact kernel, timing against the FFMA filter kernel.
usage: python tools/tc_check.py numerics|parity|time [n] [H]"""
import ctypes as C
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pitt_object_table_segmentation_b200 as pkg
from pitt_object_table_segmentation_b200 import _abi as A, scenes


def rand_samples(n, H, seed):
    rng = np.random.default_rng(seed)
    s = rng.integers(0, n, size=(H, 3), dtype=np.int64)
    s[:, 1] = (s[:, 0] + 1 + rng.integers(0, n - 2, size=H)) % n
    bad = (s[:, 2] == s[:, 0]) | (s[:, 2] == s[:, 1])
    s[bad, 2] = (s[bad, 0] + 2) % n
    return s.astype(np.int32)


def score(ctx, cloud, p, samples, mode):
    ctx.lib.pitt_debug_plane_mode(mode)
    try:
        return ctx.sac_score(cloud, p, samples)
    finally:
        ctx.lib.pitt_debug_plane_mode(0)


def numerics(ctx, n=20000, H=512, seeds=(1, 2, 3)):
    u = 2.0 ** -24
    worst = 0.0
    for seed in seeds:
        xyz = scenes.plane_outlier_cloud(n, seed=seed)
        if seed == 3:  # far from the origin: heavy cancellation between a x + b y + c z and d
            xyz[:, :3] += np.float32(37.0)
        cloud = ctx.stage(xyz)
        p = pkg.default_support_sac_params()
        samples = rand_samples(n, H, seed)
        ctx.lib.pitt_debug_plane_tc_dump(1, None)
        counts, co, valid = score(ctx, cloud, p, samples, 3)
        buf = np.zeros(128 * 256 + 2, np.float32)
        got = ctx.lib.pitt_debug_plane_tc_dump(0, buf.ctypes.data_as(C.POINTER(C.c_float)))
        assert got == buf.size, got
        acc = buf[:-2].reshape(128, 256).astype(np.float64)
        sigma, Cc = float(buf[-2]), float(buf[-1])
        a = co[:128, :4].astype(np.float64)
        pts = xyz[:256, :3].astype(np.float64)
        real = a[:, :3] @ pts.T + a[:, 3:4]
        m = np.abs(a[:, :3]) @ np.abs(pts.T) + np.abs(a[:, 3:4])
        ok = valid[:128].astype(bool)
        err = np.abs(acc / sigma - real) / (u * m)
        err = err[ok]
        print(f"seed {seed}: sigma={sigma:g} C={Cc:g} max|s~/sigma - s_real|/(u m) = {err.max():.3f}  mean = {err.mean():.3f}  "
              f"p99.9 = {np.quantile(err, 0.999):.3f}  (valid hyps {ok.sum()})")
        worst = max(worst, err.max())
        cloud.release()
    print("worst accumulation error:", worst, "u m")


def parity(ctx, n, H, seed=5):
    xyz = scenes.plane_outlier_cloud(n, seed=seed)
    cloud = ctx.stage(xyz)
    p = pkg.default_support_sac_params()
    samples = rand_samples(n, H, seed)
    out = (C.c_uint64 * 2)()
    ctx.lib.pitt_debug_plane_tc_stats(1, None)
    c3, _, _ = score(ctx, cloud, p, samples, 3)
    ctx.lib.pitt_debug_plane_tc_stats(0, out)
    c1, _, _ = score(ctx, cloud, p, samples, 1)
    bad = np.nonzero(c3 != c1)[0]
    print(f"n={n} H={H}: mismatches {bad.size} / {H}; segments {out[0]}, re-evaluated {out[1]} "
          f"({100.0 * out[1] / max(out[0], 1):.3f} %)")
    if bad.size:
        print("  first:", bad[:8], c3[bad[:8]], c1[bad[:8]])
    cloud.release()
    return bad.size == 0


def timing(ctx, n, H):
    import torch
    ctx.close()
    ctx = pkg.Context(0, seed=1, stream=torch.cuda.current_stream().cuda_stream)
    xyz = scenes.plane_outlier_cloud(n, seed=12345)
    cloud = ctx.stage(xyz)
    p = pkg.default_support_sac_params()
    samples = rand_samples(n, H, 7)
    d_s = torch.from_numpy(samples).cuda()
    d_c = torch.zeros(H, dtype=torch.int32, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    res = {}
    variants = [int(v) for v in os.environ.get("TC_VARIANTS", "0").split(",")]
    for mode, var in [(2, 0)] + [(3, v) for v in variants]:
        ctx.lib.pitt_debug_plane_tc_variant(var)
        ctx.lib.pitt_debug_plane_mode(mode)
        ts = []
        for it in range(10):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ctx.sac_score_device(cloud, p, d_s.data_ptr(), H, d_c.data_ptr())
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ctx.lib.pitt_debug_plane_mode(0)
        if var == 0:
            res[mode] = (min(ts[2:]), d_c.cpu().numpy().copy())
        print(f"mode {mode} variant {var}: {min(ts[2:]):.3f} ms per call (whole pitt_sac_score_device) -> "
              f"{n * H / min(ts[2:]) / 1e9:.2f} Gevals/ms ... {n * H / (min(ts[2:]) * 1e-3) / 1e12:.2f} Tevals/s")
    print("counts equal:", np.array_equal(res[2][1], res[3][1]))


if __name__ == "__main__":
    pkg.load_library().pitt_debug_plane_tc_nwq(int(os.environ.get("TC_NWQ", "4")))
    what = sys.argv[1]
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
    H = int(sys.argv[3]) if len(sys.argv) > 3 else 512
    ctx = pkg.Context(0)
    if what == "numerics":
        numerics(ctx)
    elif what == "parity":
        ok = parity(ctx, n, H)
        sys.exit(0 if ok else 1)
    elif what == "time":
        timing(ctx, n, H)
    ctx.close()


def cta_cycles(n, H):
    """per-CTA cycle counts of one tensor-path launch (load balance)"""
    ctx = pkg.Context(0)
    xyz = scenes.plane_outlier_cloud(n, seed=12345)
    cloud = ctx.stage(xyz)
    p = pkg.default_support_sac_params()
    samples = rand_samples(n, H, 7)
    for var in [int(v) for v in os.environ.get("TC_VARIANTS", "8,9,12").split(",")]:  # 8 = DBG kernel with full work
        ctx.lib.pitt_debug_plane_tc_variant(var)
        score(ctx, cloud, p, samples, 3)
        ctx.lib.pitt_debug_plane_tc_stats(1, None)
        score(ctx, cloud, p, samples, 3)
        ctx.lib.pitt_debug_plane_tc_stats(0, None)
        buf = (C.c_uint64 * 176)()
        ctx.lib.pitt_debug_plane_tc_cta_cycles(buf)
        v = np.array(list(buf), dtype=np.uint64)[:148]
        cyc = (v & np.uint64((1 << 48) - 1)).astype(np.int64)
        sm = (v >> np.uint64(48)).astype(np.int64)
        order = np.argsort(cyc)
        print(f"variant {var}: cycles min {cyc.min()} median {int(np.median(cyc))} max {cyc.max()}; distinct SMs {len(set(sm.tolist()))}")
        print("  slowest CTAs (block, sm, cycles):", [(int(b), int(sm[b]), int(cyc[b])) for b in order[-6:]])
        print("  fastest CTAs (block, sm, cycles):", [(int(b), int(sm[b]), int(cyc[b])) for b in order[:4]])
        mm = [int(x) for x in list(buf)[160:166]]
        print(f"  MMA thread of CTA 0: tiles {mm[5]}; per tile cycles: empty-wait {mm[0] / max(mm[5], 1):.0f}, MMA issue {mm[1] / max(mm[5], 1):.0f}, "
              f"commit {mm[2] / max(mm[5], 1):.0f}; per hb A-wait {mm[3] * 4 / max(mm[5], 1):.0f}; B-wait total {mm[4]}")
        ee = [int(x) for x in list(buf)[166:173]]
        print(f"  epilogue warp 0 of CTA 0: tiles {ee[3]}; per tile cycles: waiting for full {ee[0] / max(ee[3], 1):.0f}, "
              f"arrive -> next full seen {ee[1] / max(ee[3], 1):.0f}, full seen -> arrive (TMEM loads + first half of the math) {ee[2] / max(ee[3], 1):.0f}; "
              f"TMEM loads {ee[4] / max(ee[3], 1):.0f}, math {ee[5] / max(ee[3], 1):.0f}, tail {ee[6] / max(ee[3], 1):.0f}")
        print("  blocks 0..29 mean", cyc[:30].mean(), " blocks 30..147 mean", cyc[30:].mean())
    ctx.lib.pitt_debug_plane_tc_variant(0)


if __name__ == "__main__" and sys.argv[1] == "cta":
    cta_cycles(int(sys.argv[2]), int(sys.argv[3]))
