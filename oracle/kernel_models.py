"""TEST INFRASTRUCTURE ONLY -- host models of two device algorithms whose correctness rests on an
argument rather than on a line-by-line restatement of the reference, so that the argument itself is
checked on the CPU (``tests/test_kernel_models_cpu.py``) against the reference's own behaviour:

* ``warp_shuffle_permutation``: the batching of ``csrc/trim.cu`` (K5) -- numpy's legacy
  ``RandomState(seed).permutation(n)``, which pandas' ``DataFrame.sample(n, random_state=seed)`` evaluates
  inside the reference's ``trim_hotspot_vertices`` (randomwalk.py:256-260), run 32 draws at a time:
  chunked MT19937 refill, fixed-point rejection, parallel swaps with an in-order replay on clashes.
* ``one_sided_symmetric``: the mirror check of ``csrc/csr_build.cu`` (K0) -- search from one side of every
  mirrored pair only and balance the counts.

Nothing under node2vec_b200/ imports this module.
"""
import numpy as np

_N, _M = 624, 397


def _mt_seed(seed):
    key = np.zeros(_N, dtype=np.uint64)
    s = seed & 0xFFFFFFFF
    for pos in range(_N):
        key[pos] = s
        s = (1812433253 * (s ^ (s >> 30)) + pos + 1) & 0xFFFFFFFF
    return [int(x) for x in key]


def _refill_in_chunks(key):
    """genrand's refill in ascending 32-word chunks, every chunk reading before it writes."""
    key = list(key)
    for c in range(0, _N, 32):
        new = {}
        for i in range(c, min(c + 32, _N)):
            y = (key[i] & 0x80000000) | (key[0 if i + 1 == _N else i + 1] & 0x7FFFFFFF)
            m = i + _M if i < _N - _M else i + _M - _N
            new[i] = key[m] ^ (y >> 1) ^ (0x9908B0DF if y & 1 else 0)
        for i, v in new.items():
            key[i] = v
    return key


def _temper(y):
    y ^= y >> 11
    y ^= (y << 7) & 0x9D2C5680
    y ^= (y << 15) & 0xEFC60000
    y ^= y >> 18
    return y & 0xFFFFFFFF


def warp_shuffle_permutation(n, seed, stats=None):
    """RandomState(seed).permutation(n) the way one warp of trim_sample_kernel computes it."""
    stats = {} if stats is None else stats
    arr = list(range(n))
    key, pos, i_top = _mt_seed(seed), _N, n - 1
    while i_top >= 1:
        if pos == _N:
            key, pos = _refill_in_chunks(key), 0
        m = min(32, _N - pos)
        d = [_temper(key[pos + lane]) if lane < m else 0 for lane in range(32)]
        pos += m
        acc, rounds = (1 << m) - 1, 0
        while True:                                   # fixed point of the sequential rejection loop
            rounds += 1
            a, j, ik = [False] * 32, [0] * 32, [0] * 32
            for lane in range(32):
                ii = i_top - bin(acc & ((1 << lane) - 1)).count("1")
                if lane < m and ii >= 1:
                    mask = ii
                    for sh in (1, 2, 4, 8, 16):
                        mask |= mask >> sh
                    j[lane], ik[lane] = d[lane] & mask, ii
                    a[lane] = j[lane] <= ii
            now = sum(1 << lane for lane in range(32) if a[lane])
            if now == acc:
                break
            acc = now
        stats["rounds"] = max(stats.get("rounds", 0), rounds)
        cnt = bin(acc).count("1")
        i_low = i_top - cnt + 1
        lanes = [lane for lane in range(32) if a[lane]]
        js = [j[lane] for lane in lanes]
        clash = len(set(js)) != len(js) or any(j[lane] >= i_low and j[lane] != ik[lane] for lane in lanes)
        if clash:                                     # replay in order (lane 0 on the device)
            for lane in lanes:
                arr[j[lane]], arr[ik[lane]] = arr[ik[lane]], arr[j[lane]]
            stats["replayed"] = stats.get("replayed", 0) + 1
        else:                                         # all loads, then all stores
            x = {lane: arr[j[lane]] for lane in lanes}
            y = {lane: arr[ik[lane]] for lane in lanes}
            for lane in lanes:
                arr[j[lane]], arr[ik[lane]] = y[lane], x[lane]
            stats["parallel"] = stats.get("parallel", 0) + 1
        i_top -= cnt
    return arr


def one_sided_symmetric(row_ptr, col, w):
    """SYMMETRIC flag of a SIMPLE sorted CSR as check_symmetric decides it: the arc whose head is the
    smaller end by (degree, id) searches the head's row for its mirror; balance = #searchers - #others."""
    deg = np.diff(row_ptr)
    bad, balance = False, 0
    for a in range(len(deg)):
        for i in range(row_ptr[a], row_ptr[a + 1]):
            b = int(col[i])
            if a == b:
                continue
            if not (deg[b] < deg[a] or (deg[b] == deg[a] and b < a)):
                balance -= 1
                continue
            balance += 1
            lo, hi = int(row_ptr[b]), int(row_ptr[b + 1])
            k = lo + int(np.searchsorted(col[lo:hi], a))
            if k >= hi or col[k] != a or w[k] != w[i]:
                bad = True
    return not bad and balance == 0
