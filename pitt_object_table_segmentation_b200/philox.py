"""Philox-4x32-10 (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3", SC'11) and the minimal-sample derivation of the
PITT_SAMPLER_PHILOX sampler, restated in Python: the checker of csrc/sac.cu::philox_samples_kernel (tests only)."""
import numpy as np

M0, M1 = 0xD2511F53, 0xCD9E8D57
W0, W1 = 0x9E3779B9, 0xBB67AE85

# Random123 known-answer vectors (kat_vectors, philox4x32 10 rounds): (counter, key, expected)
KAT = [
    ((0, 0, 0, 0), (0, 0), (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)),
    ((0xFFFFFFFF,) * 4, (0xFFFFFFFF,) * 2, (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)),
    ((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0), (0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1)),
]


def philox4x32_10(counter, key):
    c, k = [int(v) for v in counter], [int(v) for v in key]
    for _ in range(10):
        p0, p1 = M0 * c[0], M1 * c[2]
        c = [(p1 >> 32) ^ c[1] ^ k[0], p1 & 0xFFFFFFFF, (p0 >> 32) ^ c[3] ^ k[1], p0 & 0xFFFFFFFF]
        k = [(k[0] + W0) & 0xFFFFFFFF, (k[1] + W1) & 0xFFFFFFFF]
    return tuple(c)


def sample_sets(H, S, n, seed, stream_id):
    """hypothesis h: counter (h, stream_id, 0x9E3779B9, 0xBB67AE85), key = the context seed; draw i is uniform in [0, n - i) from
    word i (multiply-shift) and skips over the earlier picks in ascending order (sampling without replacement)"""
    out = np.zeros((H, S), np.int32)
    key = (seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    for h in range(H):
        rnd = philox4x32_10((h, stream_id, 0x9E3779B9, 0xBB67AE85), key)
        s = []
        for i in range(S):
            v = (rnd[i] * (n - i)) >> 32
            for t in sorted(s):
                if v >= t:
                    v += 1
            s.append(v)
        out[h] = s
    return out
