"""ctypes/numpy binding of oracle/liboracle_join.so and (when present) oracle/_ref/shared.so. Checker only."""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
LIB = HERE / "liboracle_join.so"
REF = HERE / "_ref" / "shared.so"

_i32, _i64, _u32, _u64, _vp = C.c_int32, C.c_int64, C.c_uint32, C.c_uint64, C.c_void_p


def build_oracle(force: bool = False) -> Path:
    """make -C oracle: the C restatement, and _ref/shared.so when /root/reference is present."""
    if force or not LIB.exists() or LIB.stat().st_mtime < (HERE / "oracle_join.c").stat().st_mtime:
        subprocess.run(["make", "-C", str(HERE), "-s", "all"], check=True)
    return LIB


def _p(a: np.ndarray | None):
    return None if a is None else a.ctypes.data_as(_vp)


def _memref(a: np.ndarray):
    return [a.ctypes.data_as(_vp), a.ctypes.data_as(_vp), 0, a.size, 1]


def load_reference_check():
    """The reference's own check() (shared_stuff/shared.cpp:129-172) from oracle/_ref/shared.so, or None."""
    if not REF.exists():
        return None
    lib = C.CDLL(str(REF))
    lib.check.restype = _i32
    lib.check.argtypes = [_vp, _vp, _i64, _i64, _i64] * 4

    def check(R, S, outR, outS) -> int:
        arrs = [np.ascontiguousarray(x, dtype=np.int32) for x in (R, S, outR, outS)]
        args = []
        for a in arrs:
            args += _memref(a)
        return int(lib.check(*args))
    return check


class Oracle:
    def __init__(self):
        build_oracle()
        lib = C.CDLL(str(LIB))
        sig = {
            "oracle_check_i32": (_i32, [_vp, _i64, _vp, _i64, _vp, _vp, _i64]),
            "oracle_check_i64": (_i32, [_vp, _i64, _vp, _i64, _vp, _vp, _i64]),
            "oracle_check_memref": (_i32, [_vp, _vp, _i64, _i64, _i64] * 4),
            "oracle_v1_join_i32": (_i64, [_vp, _i64, _vp, _i64, _i32, _vp, _vp, _i64, C.c_int]),
            "oracle_v1_join_i64": (_i64, [_vp, _i64, _vp, _i64, _i32, _vp, _vp, _i64, C.c_int]),
            "oracle_nested_join_i32": (_i64, [_vp, _i64, _vp, _i64, _vp, _vp, _i64]),
            "oracle_nested_join_rows_i32": (_i64, [_vp, _i64, _i32, _vp, _i64, _i32, _vp, _i64]),
            "oracle_select_i32": (_i64, [_vp, _i64, _i32, _i32, _vp, _vp, _i64]),
            "oracle_select_i64": (_i64, [_vp, _i64, _i32, _i64, _vp, _vp, _i64]),
            "oracle_select_f32": (_i64, [_vp, _i64, _i32, C.c_float, _vp, _vp, _i64]),
            "oracle_select_f64": (_i64, [_vp, _i64, _i32, C.c_double, _vp, _vp, _i64]),
            "oracle_init_index": (None, [_vp, _i64]),
            "oracle_pair_digest": (None, [_vp, _vp, _i64, _vp, _vp]),
            "oracle_gen_i32": (None, [_vp, _i64, C.c_int, _u64, _i64, _u64, _u32, _u64]),
            "oracle_gen_i64": (None, [_vp, _i64, C.c_int, _u64, _i64, _u64, _u32, _u64]),
            "oracle_perm": (_u64, [_u64, _u64, _u64]),
            "oracle_perm_inv": (_u64, [_u64, _u64, _u64]),
            "oracle_max_threads": (C.c_int, []),
        }
        for name, (res, args) in sig.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        self.lib = lib

    # ---- check(): shared.cpp:129-172 ----
    def check(self, R, S, outR, outS) -> int:
        kd = np.int64 if np.asarray(R).dtype == np.int64 else np.int32
        R = np.ascontiguousarray(R, dtype=kd); S = np.ascontiguousarray(S, dtype=kd)
        outR = np.ascontiguousarray(outR, dtype=np.int32); outS = np.ascontiguousarray(outS, dtype=np.int32)
        assert outR.size == outS.size
        fn = self.lib.oracle_check_i64 if kd == np.int64 else self.lib.oracle_check_i32
        return int(fn(_p(R), R.size, _p(S), S.size, _p(outR), _p(outS), outR.size))

    # ---- join_v1 as loops: join_v1.mlir:180-521 ----
    def join(self, R, S, H: int | None = None, threads: int = 1):
        """(outR, outS) of the chained-table restatement. H defaults to max(1, nR // 2) buckets."""
        kd = np.int64 if np.asarray(R).dtype == np.int64 else np.int32
        R = np.ascontiguousarray(R, dtype=kd); S = np.ascontiguousarray(S, dtype=kd)
        if H is None:
            H = max(1, min(R.size // 2, 2**31 - 1))
        fn = self.lib.oracle_v1_join_i64 if kd == np.int64 else self.lib.oracle_v1_join_i32
        n = int(fn(_p(R), R.size, _p(S), S.size, H, None, None, 0, threads))
        if n < 0:
            raise RuntimeError(f"oracle join failed: {n}")
        outR = np.empty(n, dtype=np.int32); outS = np.empty(n, dtype=np.int32)
        if n:
            got = int(fn(_p(R), R.size, _p(S), S.size, H, _p(outR), _p(outS), n, threads))
            assert got == n
        return outR, outS

    def join_timed(self, R, S, H: int | None = None, threads: int = 0):
        """One full init+build+count+probe pass into preallocated buffers; returns (n_out, seconds)."""
        import time
        kd = np.int64 if np.asarray(R).dtype == np.int64 else np.int32
        R = np.ascontiguousarray(R, dtype=kd); S = np.ascontiguousarray(S, dtype=kd)
        if H is None:
            H = max(1, min(R.size // 2, 2**31 - 1))
        fn = self.lib.oracle_v1_join_i64 if kd == np.int64 else self.lib.oracle_v1_join_i32
        cap = int(fn(_p(R), R.size, _p(S), S.size, H, None, None, 0, threads))
        outR = np.empty(max(cap, 1), dtype=np.int32); outS = np.empty(max(cap, 1), dtype=np.int32)
        t0 = time.perf_counter()
        n = int(fn(_p(R), R.size, _p(S), S.size, H, _p(outR), _p(outS), cap, threads))
        return n, time.perf_counter() - t0

    def nested_join(self, R, S):
        R = np.ascontiguousarray(R, dtype=np.int32); S = np.ascontiguousarray(S, dtype=np.int32)
        n = int(self.lib.oracle_nested_join_i32(_p(R), R.size, _p(S), S.size, None, None, 0))
        outR = np.empty(n, dtype=np.int32); outS = np.empty(n, dtype=np.int32)
        self.lib.oracle_nested_join_i32(_p(R), R.size, _p(S), S.size, _p(outR), _p(outS), n)
        return outR, outS

    def nested_join_rows(self, tx, ty):
        """nested-loop.mlir:78-188 with row materialisation (:165-187): (n, x_cols + y_cols - 1) int32 array, x-major order."""
        tx = np.ascontiguousarray(tx, dtype=np.int32); ty = np.ascontiguousarray(ty, dtype=np.int32)
        args = (_p(tx), tx.shape[0], tx.shape[1], _p(ty), ty.shape[0], ty.shape[1])
        n = int(self.lib.oracle_nested_join_rows_i32(*args, None, 0))
        out = np.empty((n, tx.shape[1] + ty.shape[1] - 1), dtype=np.int32)
        self.lib.oracle_nested_join_rows_i32(*args, _p(out), n)
        return out

    def select(self, col, op: int, c):
        """Experiments/selection.mlir:34-155: (values, row ids) of the rows with `value OP c`, input order."""
        col = np.ascontiguousarray(col)
        fn = {np.dtype(np.int32): self.lib.oracle_select_i32, np.dtype(np.int64): self.lib.oracle_select_i64,
              np.dtype(np.float32): self.lib.oracle_select_f32, np.dtype(np.float64): self.lib.oracle_select_f64}[col.dtype]
        n = int(fn(_p(col), col.size, op, c, None, None, 0))
        vals = np.empty(n, dtype=col.dtype); rows = np.empty(n, dtype=np.int32)
        fn(_p(col), col.size, op, c, _p(vals), _p(rows), n)
        return vals, rows

    def pair_digest(self, outR, outS) -> tuple[int, int]:
        outR = np.ascontiguousarray(outR, dtype=np.int32); outS = np.ascontiguousarray(outS, dtype=np.int32)
        s, x = _u64(0), _u64(0)
        self.lib.oracle_pair_digest(_p(outR), _p(outS), outR.size, C.byref(s), C.byref(x))
        return int(s.value), int(x.value)

    def generate(self, n: int, key_bytes: int, kind: int, seed: int, lo: int = 0, domain: int = 1, p16: int = 0, key_mul: int = 0):
        out = np.empty(n, dtype=np.int32 if key_bytes == 4 else np.int64)
        fn = self.lib.oracle_gen_i32 if key_bytes == 4 else self.lib.oracle_gen_i64
        fn(_p(out), n, kind, seed, lo, domain, p16, key_mul)
        return out

    def perm(self, i: int, n: int, seed: int) -> int:
        return int(self.lib.oracle_perm(i, n, seed))

    def perm_inv(self, y: int, n: int, seed: int) -> int:
        return int(self.lib.oracle_perm_inv(y, n, seed))

    def max_threads(self) -> int:
        return int(self.lib.oracle_max_threads())


def sorted_pairs(outR, outS) -> np.ndarray:
    """Canonical form of a result: pairs sorted lexicographically (shared.cpp:168-169), as an (n, 2) int32 array."""
    p = np.stack([np.asarray(outR, dtype=np.int32), np.asarray(outS, dtype=np.int32)], axis=1)
    if p.shape[0]:
        p = p[np.lexsort((p[:, 1], p[:, 0]))]
    return p
