"""Materialised distance matrix at n = 20 000 for the two HBM-bound feature counts of the BASELINE configs (C1: F = 19 in
groups [6,6,6,1]; C2: F = 4 in groups [1,1,1,1]): 8 n^2 bytes written / device time, TMA kernel and the general tile kernel.

    python tools/bench_distmat.py [n]     prints one JSON object"""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from __graft_entry__ import load_package  # noqa: E402


def main():
    m = load_package()
    m.check(m.lib().mcbb_init(0))
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
    dev = torch.device("cuda:0")
    out = {"n": n, "hbm_copy_peak_gbs": 6551.7}
    D = torch.empty((n, n), dtype=torch.float64, device=dev)
    cases_ = [("C1_F19", [6, 6, 6, 1], [1.0, 0.5, 0.5, 1.0]), ("C2_F4", [1, 1, 1, 1], [1.0, 0.75, 0.5, 1.0])]
    for f in filter(None, os.environ.get("FS", "").split(",")):  # FS=8,12,32: one group of f features (scan over F)
        cases_.append(("F%s" % f, [int(f)], [1.0]))
    for name, gl, gw in cases_:
        F = sum(gl)
        X = torch.randn((F, n), dtype=torch.float64, device=dev)
        glen, gwt = np.asarray(gl, dtype=np.int32), np.asarray(gw)
        variants = (("tma", {}), ("tma_128rows_1cta", {b"dist_tma_r": 128}), ("tma_128rows_512thr", {b"dist_tma_rpt": 4}),
                    ("tiles", {b"dist_tma": 0}))
        for kernel, hooks in variants:
            for key in (b"dist_tma", b"dist_tma_r", b"dist_tma_rpt"):
                m.lib().mcbb_test_hook_set(key, hooks.get(key, -2 ** 63))
            best = None
            for _ in range(4):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                m.check(m.lib().mcbb_distance_matrix_dev(C.c_void_p(X.data_ptr()), n, len(gl), glen.ctypes.data_as(m.abi.c_int32_p),
                                                         gwt.ctypes.data_as(m.abi.c_double_p), C.c_void_p(D.data_ptr()),
                                                         C.c_void_p(torch.cuda.current_stream().cuda_stream)))
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1)
                best = ms if best is None else min(best, ms)
            gbs = 8.0 * n * n / (best * 1e-3) / 1e9
            out["%s_%s" % (name, kernel)] = {"ms": best, "gbs_written": gbs, "frac_of_hbm_copy": gbs / 6551.7,
                                             "fp64_adds_per_s": n * (n - 1) / 2 * 2 * F / (best * 1e-3)}
        for key in (b"dist_tma", b"dist_tma_r", b"dist_tma_rpt"):
            m.lib().mcbb_test_hook_set(key, -2 ** 63)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
