"""ctypes binding of oracle/_ref/libloradsref{32,64}.so -- TEST INFRASTRUCTURE.

The library is the UNTOUCHED reference (compiled from /root/reference/src_semi by
oracle/Makefile) plus oracle/ref_harness.c.  Only tests/, the golden-vector
generator, __graft_entry__.smoke() and bench.py's reference / cpu_baseline legs may
import this module; the product path never does.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int64)


def available(bits: int = 32) -> bool:
    return os.path.exists(os.path.join(REF_DIR, f"libloradsref{bits}.so"))


def cli_path(bits: int = 32) -> str:
    return os.path.join(REF_DIR, f"lorads_ref{bits}")


def _load(bits: int):
    lib = C.CDLL(os.path.join(REF_DIR, f"libloradsref{bits}.so"))
    lib.refh_open.restype = C.c_void_p
    lib.refh_open.argtypes = [C.c_char_p, C.c_double]
    lib.refh_info.restype = C.c_int64
    lib.refh_info.argtypes = [C.c_void_p, C.c_int, C.c_int]
    lib.refh_dinfo.restype = C.c_double
    lib.refh_dinfo.argtypes = [C.c_void_p, C.c_int]
    lib.refh_reader_csc.argtypes = [C.c_void_p, C.c_int, _ip, _ip, _dp]
    lib.refh_pattern.argtypes = [C.c_void_p, C.c_int, _ip, _ip]
    lib.refh_factor_ptr.restype = _dp
    lib.refh_factor_ptr.argtypes = [C.c_void_p, C.c_char, C.c_int]
    lib.refh_vec_ptr.restype = _dp
    lib.refh_vec_ptr.argtypes = [C.c_void_p, C.c_char]
    lib.refh_blinsys_ptr.restype = _dp
    lib.refh_blinsys_ptr.argtypes = [C.c_void_p, C.c_int]
    lib.refh_auv.argtypes = [C.c_void_p, C.c_int, C.c_char, C.c_char, _dp]
    lib.refh_obj_auv.restype = C.c_double
    lib.refh_obj_auv.argtypes = [C.c_void_p, C.c_int, C.c_char, C.c_char]
    lib.refh_wsum_mulrk.argtypes = [C.c_void_p, C.c_int, _dp, C.c_int, C.c_char, _dp]
    lib.refh_alm_cal_grad.restype = C.c_double
    lib.refh_alm_cal_grad.argtypes = [C.c_void_p, C.c_double]
    lib.refh_cg_matvec.argtypes = [C.c_void_p, C.c_int, C.c_char, _dp, _dp]
    lib.refh_update_sdp_var_one.restype = C.c_int64
    lib.refh_update_sdp_var_one.argtypes = [C.c_void_p, C.c_int, C.c_char, C.c_char, C.c_double, C.c_double, C.c_int64]
    lib.refh_alm_prepare.restype = C.c_double
    lib.refh_alm_prepare.argtypes = [C.c_void_p, C.c_double]
    lib.refh_alm_inner_iter.restype = C.c_int64
    lib.refh_alm_inner_iter.argtypes = [C.c_void_p, C.c_double, C.c_int64, _dp]
    lib.refh_time_alm_inner_iters.restype = C.c_double
    lib.refh_time_alm_inner_iters.argtypes = [C.c_void_p, C.c_double, C.c_int64, _dp]
    lib.refh_solve.argtypes = [C.c_void_p, C.c_double, C.c_int, _dp]
    lib.refh_lp_ptr.restype = _dp
    lib.refh_lp_ptr.argtypes = [C.c_void_p, C.c_char]
    lib.refh_admm_init_constr.argtypes = [C.c_void_p]
    lib.refh_admm_update_var.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_int64]
    return lib


def _d(a):
    return a.ctypes.data_as(_dp)


class RefSolver:
    """The reference solver state after main.c's init path, with kernel-level entry points."""

    def __init__(self, dat_s_path: str, bits: int = 32, times_log_rank: float = 0.0):
        self.lib = _load(bits)
        self.ctx = self.lib.refh_open(dat_s_path.encode(), float(times_log_rank))
        if not self.ctx:
            raise RuntimeError(f"reference reader failed on {dat_s_path}")
        self.m = self.info(0)
        self.n_cones = self.info(1)

    def info(self, what: int, cone: int = 0) -> int:
        return int(self.lib.refh_info(self.ctx, what, cone))

    def dinfo(self, what: int) -> float:
        return float(self.lib.refh_dinfo(self.ctx, what))

    def dim(self, cone=0):
        return self.info(2, cone)

    def rank(self, cone=0):
        return self.info(3, cone)

    def reader_csc(self, cone=0):
        nnz = self.info(7, cone)
        beg = np.zeros(self.m + 2, np.int64)
        idx = np.zeros(nnz, np.int64)
        elem = np.zeros(nnz, np.float64)
        self.lib.refh_reader_csc(self.ctx, cone, beg.ctypes.data_as(_ip), idx.ctypes.data_as(_ip), _d(elem))
        return beg, idx, elem

    def pattern(self, cone=0):
        k = self.info(4, cone)
        rows = np.zeros(k, np.int64)
        cols = np.zeros(k, np.int64)
        self.lib.refh_pattern(self.ctx, cone, rows.ctypes.data_as(_ip), cols.ctypes.data_as(_ip))
        return rows, cols

    def factor(self, which: str, cone=0) -> np.ndarray:
        """Live numpy view (n x r, Fortran order) of R / U / V / G(rad) / M(2temp)."""
        n, r = self.dim(cone), self.rank(cone)
        p = self.lib.refh_factor_ptr(self.ctx, which.encode(), cone)
        return np.ctypeslib.as_array(p, shape=(r, n)).T

    def vec(self, which: str) -> np.ndarray:
        """Live view of l(ambda) / s (constrValSum) / b / m (M1temp) / q (ARDSum) / Q (ADDSum) / v (constrVio)."""
        p = self.lib.refh_vec_ptr(self.ctx, which.encode())
        return np.ctypeslib.as_array(p, shape=(self.m,))

    @property
    def n_lp(self) -> int:
        return self.info(9)

    def lp(self, which: str) -> np.ndarray:
        """Live view of the LP vectors rLp ('R') / uLp ('U') / vLp ('V') / gradLp ('G')."""
        p = self.lib.refh_lp_ptr(self.ctx, which.encode())
        return np.ctypeslib.as_array(p, shape=(self.n_lp,))

    def admm_init_constr(self):
        """constrVal / constrValSum from (U, V) and (uLp, vLp): head of LORADSADMMOptimize, lorads_admm.c:47-48."""
        self.lib.refh_admm_init_constr(self.ctx)

    def admm_update_var(self, rho: float, tol: float, maxit: int):
        """One Gauss-Seidel sweep over every block: LORADSUpdateSDPVar / LORADSUpdateSDPLPVar."""
        self.lib.refh_admm_update_var(self.ctx, rho, tol, maxit)

    def blinsys(self, cone=0) -> np.ndarray:
        n, r = self.dim(cone), self.rank(cone)
        p = self.lib.refh_blinsys_ptr(self.ctx, cone)
        return np.ctypeslib.as_array(p, shape=(r, n)).T

    def auv(self, u: str, v: str, cone=0) -> np.ndarray:
        out = np.zeros(self.m)
        self.lib.refh_auv(self.ctx, cone, u.encode(), v.encode(), _d(out))
        return out

    def obj_auv(self, u: str, v: str, cone=0) -> float:
        return float(self.lib.refh_obj_auv(self.ctx, cone, u.encode(), v.encode()))

    def wsum_mulrk(self, w: np.ndarray, add_c: bool, x: str, cone=0) -> np.ndarray:
        n, r = self.dim(cone), self.rank(cone)
        w = np.ascontiguousarray(w, dtype=np.float64)
        out = np.zeros((r, n))
        self.lib.refh_wsum_mulrk(self.ctx, cone, _d(w), int(add_c), x.encode(), _d(out))
        return out.T

    def alm_cal_grad(self, rho: float) -> float:
        return float(self.lib.refh_alm_cal_grad(self.ctx, rho))

    def cg_matvec(self, x: np.ndarray, no_update: str, cone=0) -> np.ndarray:
        """x: n x r (any order) -> res n x r (Fortran-order view)."""
        n, r = self.dim(cone), self.rank(cone)
        xin = np.ascontiguousarray(np.asarray(x, dtype=np.float64).T)   # (r, n) C-order == col-major n x r
        res = np.zeros((r, n))
        self.lib.refh_cg_matvec(self.ctx, cone, no_update.encode(), _d(xin), _d(res))
        return res.T

    def update_sdp_var_one(self, upd: str, noupd: str, rho: float, tol: float, maxit: int, cone=0) -> int:
        return int(self.lib.refh_update_sdp_var_one(self.ctx, cone, upd.encode(), noupd.encode(), rho, tol, maxit))

    def alm_prepare(self, rho: float) -> float:
        return float(self.lib.refh_alm_prepare(self.ctx, rho))

    def alm_inner_iter(self, rho: float, counter: int):
        out = np.zeros(8)
        root = int(self.lib.refh_alm_inner_iter(self.ctx, rho, counter, _d(out)))
        return root, dict(tau=out[0], lag_norm_sq=out[1], pinf=out[2], p1=out[3], p2=out[4])

    def time_alm_inner_iters(self, rho: float, iters: int) -> float:
        out = np.zeros(8)
        return float(self.lib.refh_time_alm_inner_iters(self.ctx, rho, iters, _d(out)))

    def solve(self, time_limit: float = 0.0, reopt_level: int = 2) -> dict:
        out = np.zeros(16)
        self.lib.refh_solve(self.ctx, time_limit, reopt_level, _d(out))
        keys = ["pobj", "dobj", "pinf", "gap", "dinf", "alm_inner", "admm_iter", "cg_iter", "seconds", "status"]
        return dict(zip(keys, out[:10].tolist()))
