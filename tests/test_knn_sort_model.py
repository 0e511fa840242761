"""CPU model of the k-NN finishing sort (csrc/knn.cu, knn_sort_body): the candidates of a query are sorted on ONE 32-bit
word each (bits of d2 with the low 6 bits replaced by the slot number) through the generated 64-input network
(tools/gen_sort64.py -> csrc/knn_sort64.inc), then adjacent words that agree in their upper 26 bits are put in order by
their true 64-bit (d2, index) keys, bubble passes until nothing moves. The model checks the claim the kernel rests on: the
result is the order of the 64-bit keys exactly, for any number of candidates, exact ties, few-ulp near ties and duplicates.
(The GPU side of the same claim: tests/test_gpu_normals_clusters.py::test_knn_grid_path_with_ties_and_near_ties.)"""
import os
import random
import re
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
from gen_sort64 import network  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _sort_model(net, c, d2_bits, idx):
    true = [(d2_bits[s] << 32) | idx[s] for s in range(c)]
    key = [((d2_bits[s] & ~63) | s) if s < c else (0x80000000 | (s << 6)) for s in range(64)]
    for i, j in net:
        if key[j] < key[i]:
            key[i], key[j] = key[j], key[i]
    if min(key[s] ^ key[s + 1] for s in range(63)) < 64:
        again = True
        while again:
            again = False
            for s in range(63):
                if (key[s] ^ key[s + 1]) < 64:
                    if true[key[s + 1] & 63] < true[key[s] & 63]:
                        key[s], key[s + 1] = key[s + 1], key[s]
                        again = True
    assert all(k >= 0x80000000 for k in key[c:])  # the pads stay behind every real candidate
    return [true[k & 63] for k in key[:c]], true


def test_generated_network_is_the_one_the_kernel_includes():
    inc = open(os.path.join(ROOT, "pitt_object_table_segmentation_b200", "csrc", "knn_sort64.inc")).read()
    pairs = [(int(a), int(b)) for a, b in re.findall(r"CE\((\d+),(\d+)\)", inc)]
    assert pairs == network(64) and len(pairs) == 543


def test_32_bit_sort_with_true_key_fix_up_equals_the_64_bit_order():
    net = network(64)
    rng = random.Random(7)
    for trial in range(4000):
        c = rng.randint(1, 64)
        base = rng.randint(0, 0x7F000000)
        mode = trial % 4
        if mode == 0:    # unrelated distances
            d = [rng.randint(0, 0x7F7FFFFF) for _ in range(c)]
        elif mode == 1:  # all within a few hundred ulps of each other
            d = [base + rng.randint(0, 200) for _ in range(c)]
        elif mode == 2:  # a handful of values: long runs of exact ties, some one ulp apart, some across a 64-ulp boundary
            vals = [base + 64 * rng.randint(0, 3) + rng.choice([0, 0, 1, 63]) for _ in range(4)]
            d = [rng.choice(vals) for _ in range(c)]
        else:            # duplicates of one point (d2 = 0 for the query itself and its copies)
            d = [rng.choice([0, base])] * c
        idx = rng.sample(range(1 << 22), c)
        got, true = _sort_model(net, c, d, idx)
        assert got == sorted(true)
