"""ctypes wrapper over oracle/liblrr_oracle.so (the C restatement).  TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

from . import linreg_oracle as O

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    so = os.path.join(_HERE, "liblrr_oracle.so")
    src = os.path.join(_HERE, "linreg_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        # -march=native is dropped: the .so is built here and travels to a different host CPU.
        subprocess.check_call(
            ["gcc", "-O3", "-mavx2", "-mfma", "-fopenmp", "-fPIC", "-std=gnu11", "-shared", "-o", so, src, "-lm"]
        )
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "liblrr_oracle.so")
        if not os.path.exists(so):
            build()
        L = ctypes.CDLL(so)
        dp = ctypes.POINTER(ctypes.c_double)
        L.lrr_oracle_bed.argtypes = [
            ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int32,
            dp, dp, dp, dp, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32,
            dp, dp, dp, dp, dp, dp,
        ]
        L.lrr_oracle_bed.restype = None
        L.lrr_oracle_two_sided_p.argtypes = [ctypes.c_double, ctypes.c_double]
        L.lrr_oracle_two_sided_p.restype = ctypes.c_double
        L.lrr_oracle_max_threads.restype = ctypes.c_int
        _LIB = L
    return _LIB


def _dp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def prepare(ys, cov):
    """Driver prologue (RU:88-128, LR:47-78) once; reuse across calls of `run_prepared`."""
    y, c, idx = O.complete_samples(ys, cov)
    n, K, d, Qt, Qty, yyp = O.prologue(y, c)
    P = y.shape[1]
    return {
        "n": n, "K": K, "P": P, "d": d,
        "y": np.ascontiguousarray(y),
        "Qt": np.ascontiguousarray(Qt if K > 0 else np.zeros((1, n))),
        "Qty": np.ascontiguousarray(Qty if K > 0 else np.zeros((1, P))),
        "yyp": np.ascontiguousarray(yyp),
        "idx": np.ascontiguousarray(idx, dtype=np.int32),
    }


def run_prepared(bed_rows, prep, block_size=16, n_threads=0):
    """The per-partition loop (LR:95-193) over PLINK-coded rows uint8 [M, stride]."""
    bed_rows = np.ascontiguousarray(bed_rows, dtype=np.uint8)
    M, stride = bed_rows.shape
    n, K, P = prep["n"], prep["K"], prep["P"]
    out = {
        "n": np.full(M, n, dtype=np.int32),
        "sum_x": np.empty(M),
        "y_transpose_x": np.empty((M, P)),
        "beta": np.empty((M, P)),
        "standard_error": np.empty((M, P)),
        "t_stat": np.empty((M, P)),
        "p_value": np.empty((M, P)),
    }
    lib().lrr_oracle_bed(
        bed_rows.ctypes.data, M, stride, prep["idx"].ctypes.data, n, _dp(prep["Qt"]), _dp(prep["y"]), _dp(prep["Qty"]),
        _dp(prep["yyp"]), K, P, block_size, n_threads,
        _dp(out["sum_x"]), _dp(out["y_transpose_x"]), _dp(out["beta"]), _dp(out["standard_error"]),
        _dp(out["t_stat"]), _dp(out["p_value"]),
    )
    out["_d"] = prep["d"]
    return out


def linreg_group_bed(bed_rows, n_samples, ys, cov, block_size=16, n_threads=0):
    """LR:46-195 over PLINK-coded rows `bed_rows` uint8 [M, stride]; ys [N,P], cov [N,K] (NaN = missing)."""
    bed_rows = np.ascontiguousarray(bed_rows, dtype=np.uint8)
    assert bed_rows.shape[1] >= (n_samples + 3) // 4
    return run_prepared(bed_rows, prepare(ys, cov), block_size, n_threads)


def max_threads():
    return int(lib().lrr_oracle_max_threads())
