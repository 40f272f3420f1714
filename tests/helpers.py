import json
import os

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


def unhex(xs):
    return [float.fromhex(x) for x in xs]


# ---- independent numpy construction of the device records (for the host replay) -------
import numpy as np  # noqa: E402

UNIT, SYMMETRIC, SIMPLE = 1, 2, 4


def pack_arcs(row_ptr, col, alias, probs):
    """thr / dst / alias_dst exactly as include/n2v_b200.h documents n2v_arc_t."""
    probs = np.asarray(probs, dtype=np.float64)
    scaled = np.ceil(probs * 4294967296.0)
    thr = np.where(probs >= 1.0, 4294967295.0, np.minimum(scaled, 4294967295.0)).astype(np.uint32)
    deg = np.diff(row_ptr)
    base_of_arc = np.repeat(row_ptr[:-1], deg)
    alias_dst = np.where(probs >= 1.0, col, col[base_of_arc + alias]).astype(np.int32)
    return thr, col.astype(np.int32), alias_dst


def graph_flags(row_ptr, col, w):
    deg = np.diff(row_ptr)
    src = np.repeat(np.arange(len(deg)), deg)
    flags = 0
    if np.all(w == 1.0):
        flags |= UNIT
    key = src.astype(np.int64) << 32 | col.astype(np.int64)
    simple = len(np.unique(key)) == len(key)
    if simple:
        flags |= SIMPLE
        fwd = {(int(a), int(b)): float(x) for a, b, x in zip(src, col, w)}
        if all(fwd.get((b, a)) == x for (a, b), x in fwd.items()):
            flags |= SYMMETRIC
    return flags


def chi_square_ok(counts, probs, alpha=1e-4):
    """Pearson chi-square of observed counts against a law; pools cells with expectation < 5."""
    from scipy import stats
    counts = np.asarray(counts, dtype=np.float64)
    probs = np.asarray(probs, dtype=np.float64)
    n = counts.sum()
    exp = probs * n
    order = np.argsort(exp)
    counts, exp = counts[order], exp[order]
    small = exp < 5
    if small.any() and (~small).any():
        counts = np.concatenate([[counts[small].sum()], counts[~small]])
        exp = np.concatenate([[exp[small].sum()], exp[~small]])
    if len(exp) < 2:
        return True, 1.0
    chi2 = ((counts - exp) ** 2 / np.maximum(exp, 1e-300)).sum()
    pval = stats.chi2.sf(chi2, len(exp) - 1)
    return pval > alpha, pval
