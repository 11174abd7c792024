"""ctypes binding of include/lrr_b200.h.  There is no CPU fallback: if the CUDA library is missing this raises."""
from __future__ import annotations

import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# LRR_B200_LIB overrides the library file (A/B builds of the same sources during kernel tuning)
LIB_PATH = os.environ.get("LRR_B200_LIB") or os.path.join(HERE, "liblrr_b200.so")

KERNEL_AUTO, KERNEL_FP64, KERNEL_TC, KERNEL_TC4 = 0, 1, 2, 3
KERNELS = {"auto": KERNEL_AUTO, "fp64": KERNEL_FP64, "tc": KERNEL_TC, "tc4": KERNEL_TC4}

c_i32p = ctypes.POINTER(ctypes.c_int32)
c_f64p = ctypes.POINTER(ctypes.c_double)


class GroupOut(ctypes.Structure):
    """lrr_group_out"""

    _fields_ = [
        ("n", ctypes.c_void_p),
        ("n_missing", ctypes.c_void_p),
        ("sum_x", ctypes.c_void_p),
        ("y_transpose_x", ctypes.c_void_p),
        ("beta", ctypes.c_void_p),
        ("standard_error", ctypes.c_void_p),
        ("t_stat", ctypes.c_void_p),
        ("p_value", ctypes.c_void_p),
        ("log10_p", ctypes.c_void_p),
    ]


class ScoreOut(ctypes.Structure):
    """lrr_score_out"""

    _fields_ = [("chi_sq_stat", ctypes.c_void_p), ("p_value", ctypes.c_void_p), ("n_missing", ctypes.c_void_p)]


class LogitOut(ctypes.Structure):
    """lrr_logit_out"""

    _fields_ = [(k, ctypes.c_void_p) for k in ("beta", "standard_error", "z_stat", "chi_sq_stat", "p_value",
                                               "n_iterations", "converged", "exploded")]


# name -> (restype, argtypes); must list every symbol include/lrr_b200.h declares (tests check this)
SIGNATURES = {
    "lrr_version": (ctypes.c_char_p, []),
    "lrr_create": (ctypes.c_int, [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int]),
    "lrr_destroy": (None, [ctypes.c_void_p]),
    "lrr_last_error": (ctypes.c_char_p, [ctypes.c_void_p]),
    "lrr_packed_stride": (ctypes.c_int64, [ctypes.c_int64]),
    "lrr_pack_bed": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64,
                                    ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p]),
    "lrr_pack_dosage_i8": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64,
                                          ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p]),
    "lrr_unpack_dosage_i8": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64,
                                            ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p]),
    "lrr_unpack_bed": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64,
                                      ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p]),
    "lrr_bn_fill": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int64,
                                   ctypes.c_int64, ctypes.c_int64, ctypes.c_uint64, ctypes.c_void_p, ctypes.c_int64,
                                   ctypes.c_void_p, ctypes.c_void_p]),
    "lrr_clear_groups": (ctypes.c_int, [ctypes.c_void_p]),
    "lrr_add_group": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32,
                                     ctypes.c_int32, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                     ctypes.c_void_p, ctypes.c_void_p]),
    "lrr_add_group_weighted": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32, ctypes.c_int32,
                                              ctypes.c_int32, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                              ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "lrr_num_groups": (ctypes.c_int, [ctypes.c_void_p]),
    "lrr_reserve": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64]),
    "lrr_run": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64,
                               ctypes.c_int64, ctypes.POINTER(GroupOut), ctypes.c_int32, ctypes.c_int32,
                               ctypes.c_void_p]),
    "lrr_run_dense": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64,
                                     ctypes.POINTER(GroupOut), ctypes.c_int32, ctypes.c_void_p]),
    "lrr_run_dense_u16": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64,
                                         ctypes.c_double, ctypes.POINTER(GroupOut), ctypes.c_int32, ctypes.c_void_p]),
    "lrr_set_logit_model": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32, ctypes.c_int32, ctypes.c_void_p,
                                           ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                           ctypes.c_void_p, ctypes.c_double]),
    "lrr_run_logit": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64,
                                     ctypes.c_int32, ctypes.c_int32, ctypes.c_double, ctypes.POINTER(LogitOut),
                                     ctypes.c_void_p]),
    "lrr_run_logit_dense": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64,
                                           ctypes.c_int32, ctypes.c_int32, ctypes.c_double, ctypes.POINTER(LogitOut),
                                           ctypes.c_void_p]),
    "lrr_qchisqtail1": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p]),
    "lrr_at_times": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64,
                                    ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32, ctypes.c_void_p,
                                    ctypes.c_void_p]),
    "lrr_set_guard": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int]),
    "lrr_last_recomputed": (ctypes.c_int64, [ctypes.c_void_p]),
    "lrr_launch_count": (ctypes.c_int64, [ctypes.c_void_p]),
    "lrr_last_kernel": (ctypes.c_int, [ctypes.c_void_p]),
    "lrr_set_timing": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int]),
    "lrr_last_sweep_ms": (ctypes.c_float, [ctypes.c_void_p]),
    "lrr_last_sweep_shape": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_int64)]),
    "lrr_stream_begin": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_void_p), ctypes.c_void_p, ctypes.c_int64,
                                        ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, ctypes.c_int32]),
    "lrr_stream_run": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.POINTER(GroupOut), ctypes.c_int32,
                                      ctypes.c_int32]),
    "lrr_stream_end": (None, [ctypes.c_void_p, ctypes.c_void_p]),
    "lrr_trim": (ctypes.c_int, [ctypes.c_void_p]),
    "lrr_last_stream_h2d_ms": (ctypes.c_float, [ctypes.c_void_p]),
    "lrr_last_stream_timeline": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]),
    "lrr_set_score_model": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32, ctypes.c_int32,
                                           ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                           ctypes.c_void_p, ctypes.c_void_p]),
    "lrr_run_score": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64,
                                     ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p]),
    "lrr_run_score_dense": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64,
                                           ctypes.c_void_p, ctypes.c_void_p]),
    "lrr_student_t_two_sided": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_double,
                                               ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
}

_lib = None


class LrrError(RuntimeError):
    """Non-zero return from the C ABI (message = lrr_last_error)."""

    def __init__(self, code, message):
        super().__init__(f"lrr_b200 error {code}: {message}")
        self.code = code
        self.message = message


def load():
    """Load liblrr_b200.so.  Fails loudly when the CUDA extension has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build the sm_100a CUDA library first "
            "(python -m hail_b200.build, or __graft_entry__.build()). hail_b200 has no CPU fallback."
        )
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class Context:
    """One lrr_ctx bound to one CUDA device (not thread-safe; one per device, like the reference's
    thread-local BLAS handles, hail/hail/src/is/hail/linalg/BLAS.scala:13-35)."""

    def __init__(self, device: int = 0):
        self.lib = load()
        h = ctypes.c_void_p()
        rc = self.lib.lrr_create(ctypes.byref(h), int(device))
        if rc != 0:
            raise LrrError(rc, self.lib.lrr_last_error(None).decode())
        self.handle = h
        self.device = int(device)

    def check(self, rc):
        if rc != 0:
            raise LrrError(rc, self.lib.lrr_last_error(self.handle).decode())

    def close(self):
        if getattr(self, "handle", None):
            self.lib.lrr_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def last_recomputed(self) -> int:
        """Rows of the last lrr_run that the tolerance guard sent to the float64 recompute (synchronises)."""
        return int(self.lib.lrr_last_recomputed(self.handle))

    @property
    def launch_count(self) -> int:
        return int(self.lib.lrr_launch_count(self.handle))

    @property
    def last_sweep_shape(self):
        """(sweep launches, MMA columns over them, columns of two-plane-capable launches, digit columns in use) of the last lrr_run."""
        out = (ctypes.c_int64 * 4)()
        self.check(self.lib.lrr_last_sweep_shape(self.handle, out))
        return tuple(int(v) for v in out)

    @property
    def last_kernel(self) -> str:
        k = int(self.lib.lrr_last_kernel(self.handle))
        return {v: n for n, v in KERNELS.items()}.get(k, str(k))


_contexts = {}


def context(device: int = 0) -> Context:
    ctx = _contexts.get(device)
    if ctx is None or ctx.handle is None:
        ctx = _contexts[device] = Context(device)
    return ctx
