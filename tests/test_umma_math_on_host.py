"""CPU test (-m "not gpu") of the arithmetic the tcgen05 ensemble integrator's exactness rests on, compiled for the host
from the PRODUCT source (csrc/umma_math.cuh via tests/host_emul/umma_math_emul.cu):
  * the branch-free sincos against libm,
  * u = round(x 2^50) + 2^51 bit for bit,
  * the coupling sum  sum_j A_ij s_j  rebuilt from the radix-256 digit sums (what the INT8 tensor cores accumulate):
    equal to the correctly rounded value of the exact integer sum — for ANY summation order, i.e. more than an FP64
    loop can promise."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = os.path.join(HERE, "host_emul", "umma_math_emul.cu")
OUT = os.path.join(HERE, "host_emul", "_build", "libumma_math_emul.so")


@pytest.fixture(scope="module")
def um():
    deps = [SRC] + [os.path.join(ROOT, "mcbb.jl_b200", "csrc", f) for f in ("umma_math.cuh", "common.cuh")]
    if not os.path.exists(OUT) or any(os.path.getmtime(d) > os.path.getmtime(OUT) for d in deps):
        os.makedirs(os.path.dirname(OUT), exist_ok=True)
        cmd = ["nvcc", "-std=c++17", "-O2", "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-Xcompiler",
               "-fPIC,-ffp-contract=off", "-I" + os.path.join(ROOT, "include"), "-o", OUT, SRC]
        r = subprocess.run(cmd, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-2000:]
    L = C.CDLL(OUT)
    dp, up, ip = C.POINTER(C.c_double), C.POINTER(C.c_uint32), C.POINTER(C.c_int32)
    L.emul_um_sincos.argtypes = [dp, C.c_int64, dp, dp]
    L.emul_um_fixed52.argtypes = [dp, C.c_int64, up, up]
    L.emul_um_recombine.argtypes = [up, ip, C.c_int64, dp]
    return L, dp, up, ip


def _sincos(um, x):
    L, dp, _, _ = um
    x = np.ascontiguousarray(x, dtype=np.float64)
    s, c = np.empty_like(x), np.empty_like(x)
    L.emul_um_sincos(x.ctypes.data_as(dp), len(x), s.ctypes.data_as(dp), c.ctypes.data_as(dp))
    return s, c


def _fixed52(um, x):
    L, dp, up, _ = um
    x = np.ascontiguousarray(x, dtype=np.float64)
    lo, hi = np.empty(len(x), np.uint32), np.empty(len(x), np.uint32)
    L.emul_um_fixed52(x.ctypes.data_as(dp), len(x), lo.ctypes.data_as(up), hi.ctypes.data_as(up))
    return lo, hi


def test_branch_free_sincos_accuracy(um):
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.uniform(-10, 10, 200000), rng.uniform(-2000, 2000, 200000), rng.uniform(-1e5, 1e5, 100000),
                        np.array([0.0, -0.0, np.pi / 4, -np.pi / 4, np.pi / 2, np.pi, 1e-300, 1.0, -1.0]),
                        np.arange(-64, 65) * (np.pi / 4)])
    s, c = _sincos(um, x)
    xl = x.astype(np.longdouble)
    es, ec = np.abs(s - np.sin(xl)).astype(float), np.abs(c - np.cos(xl)).astype(float)
    small = np.abs(x) <= 2000  # phases of a run of the bench workload stay well inside
    assert es[small].max() < 2.5e-16 and ec[small].max() < 2.5e-16, (es[small].max(), ec[small].max())
    assert es.max() < 1e-14 and ec.max() < 1e-14  # two-term Cody-Waite: the error grows with |x| / (pi/2)
    assert np.all(np.abs(s) <= 1.0) and np.all(np.abs(c) <= 1.0)  # the fixed-point split needs [-1, 1]
    assert np.abs(s * s + c * c - 1.0).max() < 1e-15
    s0, c0 = _sincos(um, np.array([0.0]))
    assert s0[0] == 0.0 and c0[0] == 1.0


def test_fixed_point_split_is_exact(um):
    rng = np.random.default_rng(1)
    x = np.concatenate([rng.uniform(-1, 1, 100000), [1.0, -1.0, 0.0, -0.0, 2.0 ** -50, -2.0 ** -50, 2.0 ** -51, 3 * 2.0 ** -51,
                                                     1 - 2.0 ** -53, -1 + 2.0 ** -53, 2.0 ** -60]])
    lo, hi = _fixed52(um, x)
    u = (hi.astype(np.uint64) << np.uint64(32)) | lo.astype(np.uint64)
    want = np.array([int(np.rint(v * 2.0 ** 50)) + 2 ** 51 for v in x], dtype=np.uint64)  # x 2^50 is exact; rint = RNE
    assert np.array_equal(u, want)
    assert u.max() <= 2 ** 51 + 2 ** 50 and u.min() >= 2 ** 50 and np.all(hi < 2 ** 20)  # 7 digits, d_7 = 0


def test_digit_sums_recombine_to_the_exact_coupling_sum(um):
    """One row of P = A sin(Y) the way the tensor cores see it: every s_j becomes seven radix-256 digits, the MMA adds
    digit k over the neighbours (exact in int32), and um_recombine joins the seven sums and removes the offsets."""
    L, dp, up, ip = um
    rng = np.random.default_rng(2)
    n_rows = 4000
    v = np.zeros((n_rows, 7), dtype=np.uint32)
    degh = np.zeros(n_rows, dtype=np.int32)
    want = np.zeros(n_rows)
    fp64_loop_differs = 0
    for r in range(n_rows):
        deg = int(rng.integers(0, 129)) if r > 3 else (0, 1, 128, 128)[r]
        sj = rng.uniform(-1, 1, deg) if r != 3 else np.ones(deg)
        lo, hi = _fixed52(um, sj)
        u = (hi.astype(np.uint64) << np.uint64(32)) | lo.astype(np.uint64)
        digits = np.array([(u >> np.uint64(8 * k)) & np.uint64(255) for k in range(7)], dtype=np.uint64)  # [7, deg]
        v[r] = digits.sum(axis=1).astype(np.uint32)
        assert v[r].max() < 2 ** 15
        degh[r] = deg << 19
        exact = sum(int(q) - 2 ** 51 for q in u)  # integer sum of round(s_j 2^50)
        want[r] = float(exact)  # correctly rounded, once
        loop = 0.0
        for q in u:
            loop += float(int(q) - 2 ** 51)
        fp64_loop_differs += loop != want[r]
    out = np.empty(n_rows)
    L.emul_um_recombine(v.ctypes.data_as(up), degh.ctypes.data_as(ip), n_rows, out.ctypes.data_as(dp))
    assert np.array_equal(out, want)
    assert out[0] == 0.0 and out[3] == 128 * 2.0 ** 50
    assert fp64_loop_differs > 0  # a plain FP64 accumulation of the same terms is NOT always correctly rounded
