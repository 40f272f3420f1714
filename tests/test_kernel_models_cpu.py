"""CPU checks of the two device algorithms whose correctness is an argument (oracle/kernel_models.py):
the warp-batched Fisher-Yates of K5 against numpy's legacy RandomState (what the reference's
trim_hotspot_vertices evaluates, randomwalk.py:256-260), and the one-sided mirror search of K0 against
a set-based symmetry check.  The GPU tests then compare the kernels with the same references."""
import numpy as np
import pytest

from oracle import clib, kernel_models
from tests.helpers import SYMMETRIC, graph_flags


@pytest.mark.parametrize("n,seed", [(1, 3), (2, 1), (3, 5), (5, 0), (33, 7), (100, 1), (624, 9), (625, 2), (1249, 11),
                                    (5000, 42), (20011, 123456789), (7001, 2 ** 32 - 1)])
def test_warp_batched_shuffle_is_numpy_legacy_permutation(n, seed):
    stats = {}
    got = kernel_models.warp_shuffle_permutation(n, seed, stats)
    assert got == np.random.RandomState(seed).permutation(n).tolist()
    if n >= 20000:
        assert stats["parallel"] > 5 * stats.get("replayed", 0)      # clashes are the exception


def test_fixed_point_rejection_needs_few_rounds():
    stats = {}
    kernel_models.warp_shuffle_permutation(30011, 5, stats)
    assert stats["rounds"] <= 33


def _symmetric_graph(rng, n, m):
    a, b = rng.integers(0, n, m), rng.integers(0, n, m)
    hub = np.full(n // 3, 2)
    a, b = np.concatenate([a, hub, [4]]), np.concatenate([b, rng.permutation(n)[: n // 3], [4]])
    lo, hi = np.minimum(a, b), np.maximum(a, b)
    _, idx = np.unique(lo.astype(np.int64) << 32 | hi, return_index=True)
    lo, hi = lo[idx], hi[idx]
    w = rng.uniform(0.5, 2.0, len(lo))
    loop = lo == hi
    return np.concatenate([lo, hi[~loop]]), np.concatenate([hi, lo[~loop]]), np.concatenate([w, w[~loop]])


def test_one_sided_mirror_search_equals_the_set_based_check():
    rng = np.random.default_rng(3)
    n = 120
    src, dst, w = _symmetric_graph(rng, n, 700)

    def both(s, d, x):
        row_ptr, col, ws, _ = clib.csr_from_arcs(s, d, x, n)
        want = bool(graph_flags(row_ptr, col, ws) & SYMMETRIC)
        assert kernel_models.one_sided_symmetric(row_ptr, col, ws) == want
        return want

    assert both(src, dst, w)
    for i in rng.integers(0, len(src), 40).tolist():
        if src[i] == dst[i]:
            continue
        keep = np.ones(len(src), dtype=bool)
        keep[i] = False
        assert not both(src[keep], dst[keep], w[keep])               # a mirror is missing
        w2 = w.copy()
        w2[i] = np.nextafter(w2[i], 3.0)
        assert not both(src, dst, w2)                                # a mirror weighs an ulp more
    present = set(zip(src.tolist(), dst.tolist()))
    free = [(u, v) for u in range(15) for v in range(15, 30) if (u, v) not in present][:6]
    for u, v in free:
        assert not both(np.append(src, u), np.append(dst, v), np.append(w, 1.0))
        assert both(np.append(src, [u, v]), np.append(dst, [v, u]), np.append(w, [1.5, 1.5]))
    assert both(np.append(src, 9), np.append(dst, 9), np.append(w, 1.0)) or (9, 9) in present
    # directed random graphs: almost surely asymmetric, and the two checks agree on every one
    for _ in range(20):
        s, d = rng.integers(0, 30, 80), rng.integers(0, 30, 80)
        _, idx = np.unique(s.astype(np.int64) << 32 | d, return_index=True)
        both(s[idx], d[idx], np.ones(len(idx)))
