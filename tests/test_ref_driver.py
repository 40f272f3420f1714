"""The bench's reference arm (oracle/ref_driver.py) against the reference's own entry point: the driver
chains the UNMODIFIED transformer functions with pandas merges; `tests/golden/fugue_verbatim.json` holds
what the reference's `node2vec.fugue.random_walk` itself returned on the same graphs (run verbatim on the
Fugue stand-in by tests/golden/make_golden.py).  Same seed => the same walks, row for row."""
import numpy as np
import pytest

from oracle import clib, ref_driver
from tests.helpers import load_golden, unhex


def test_ref_driver_reproduces_the_reference_entry_point():
    rw, root = ref_driver.load_reference()
    if rw is None:
        pytest.skip("no copy of the reference here (neither /root/reference nor baseline/_ref)")
    fx = load_golden("fugue_verbatim.json")
    done = 0
    for case in fx["walks"]:
        if case["walk_seed"] is not None:
            continue                                   # the driver walks plain start-vertex lists
        src, dst, w = np.asarray(case["src"]), np.asarray(case["dst"]), np.asarray(unhex(case["weight"]))
        n = int(max(src.max(), dst.max())) + 1
        row_ptr, col, ws, _ = clib.csr_from_arcs(src, dst, w, n)
        prm = case["params_after"]
        adj = ref_driver.AdjacencyRows(rw, row_ptr, col, ws)
        starts = np.flatnonzero(np.diff(row_ptr) > 0).tolist()
        paths, steps, hot = ref_driver.walk(rw, adj, starts, prm["num_walks"], prm["walk_length"], prm["return_param"],
                                            prm["inout_param"], seed=case["random_seed"])
        want = case["native"]["walk"]
        assert sorted(paths) == sorted(want), case["graph"]
        assert steps > 0 and hot > 0
        done += 1
    assert done >= 2


def test_bench_arms_share_one_config_dict():
    """Both arms of bench.py describe the workload with the same dict (the driver compares them)."""
    import bench
    for name, w in bench.WORKLOADS.items():
        a = bench.workload_config(name, w, 8)
        b = bench.workload_config(name, dict(w), 8)
        assert a == b and a["workload"] == name and set(a) >= {"graph", "p", "q", "num_walks", "walk_length", "dim", "sgns", "l2"}
    stats = {"steps": 100, "trials": 181, "probes": 113}
    r = bench.walk_roofline(stats, 790e6, 58.0, "hbm", "rmat20", "n")
    assert abs(r["bytes_per_step"] - (16 + 1.81 * (12 + 4 * 113 / 181) + 4)) < 1e-9      # SURVEY 8d formula
    assert abs(r["sectors_per_step"] - 2.94) < 1e-9 and r["traffic"] and 0 < r["frac"] < 1
    assert bench.sgns_bytes_per_pair(128, 5) == 7168.0
