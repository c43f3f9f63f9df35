"""ctypes binding of oracle/liblorads_oracle.so (the plain-C restatement) -- TEST INFRASTRUCTURE.

Same method names as oracle/ref.py:RefSolver so that a test can run one check against either checker.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "liblorads_oracle.so")
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int64)


def build(force: bool = False) -> str:
    src = os.path.join(HERE, "lorads_oracle.c")
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-std=gnu11", "-fPIC", "-shared", "-o", LIB, src, "-lm"])
    return LIB


def _load():
    lib = C.CDLL(build())
    lib.orc_create.restype = C.c_void_p
    lib.orc_create.argtypes = [C.c_int64, C.c_int64, _ip, _dp]
    lib.orc_set_cone.argtypes = [C.c_void_p, C.c_int64, _ip, _ip, _dp]
    lib.orc_determine_rank.argtypes = [C.c_void_p, C.c_double]
    lib.orc_init_vars.argtypes = [C.c_void_p, C.c_int64]
    lib.orc_info.restype = C.c_int64
    lib.orc_info.argtypes = [C.c_void_p, C.c_int, C.c_int64]
    lib.orc_dinfo.restype = C.c_double
    lib.orc_dinfo.argtypes = [C.c_void_p, C.c_int]
    lib.orc_pattern.argtypes = [C.c_void_p, C.c_int64, _ip, _ip]
    lib.orc_factor_ptr.restype = _dp
    lib.orc_factor_ptr.argtypes = [C.c_void_p, C.c_char, C.c_int64]
    lib.orc_vec_ptr.restype = _dp
    lib.orc_vec_ptr.argtypes = [C.c_void_p, C.c_char]
    lib.orc_auv.argtypes = [C.c_void_p, C.c_int64, C.c_char, C.c_char, _dp]
    lib.orc_obj_auv.restype = C.c_double
    lib.orc_obj_auv.argtypes = [C.c_void_p, C.c_int64, C.c_char, C.c_char]
    lib.orc_wsum_mulrk.argtypes = [C.c_void_p, C.c_int64, _dp, C.c_int, C.c_char, _dp]
    lib.orc_alm_cal_grad.restype = C.c_double
    lib.orc_alm_cal_grad.argtypes = [C.c_void_p, C.c_double]
    lib.orc_alm_prepare.restype = C.c_double
    lib.orc_alm_prepare.argtypes = [C.c_void_p, C.c_double]
    lib.orc_cg_matvec.argtypes = [C.c_void_p, C.c_int64, C.c_char, _dp, _dp]
    lib.orc_update_sdp_var_one.restype = C.c_int64
    lib.orc_update_sdp_var_one.argtypes = [C.c_void_p, C.c_int64, C.c_char, C.c_char, C.c_double, C.c_double, C.c_int64]
    lib.orc_alm_inner_iter.restype = C.c_int64
    lib.orc_alm_inner_iter.argtypes = [C.c_void_p, C.c_double, C.c_int64, _dp]
    return lib


def _d(a):
    return a.ctypes.data_as(_dp)


def _i(a):
    return a.ctypes.data_as(_ip)


class OracleSolver:
    def __init__(self, inst, times_log_rank: float = 2.0, lbfgs_len: int = 2):
        self.lib = _load()
        dims = np.asarray(inst.blk_dims, dtype=np.int64)
        b = np.ascontiguousarray(inst.b, dtype=np.float64)
        self.m = inst.m
        self.ctx = self.lib.orc_create(inst.m, len(inst.cones), _i(dims), _d(b))
        for k, cone in enumerate(inst.cones):
            beg = np.ascontiguousarray(cone.beg, dtype=np.int64)
            idx = np.ascontiguousarray(cone.idx, dtype=np.int64)
            elem = np.ascontiguousarray(cone.elem, dtype=np.float64)
            self.lib.orc_set_cone(self.ctx, k, _i(beg), _i(idx), _d(elem))
        self.lib.orc_determine_rank(self.ctx, float(times_log_rank))
        self.lib.orc_init_vars(self.ctx, lbfgs_len)
        self.n_cones = len(inst.cones)

    def info(self, what, cone=0):
        return int(self.lib.orc_info(self.ctx, what, cone))

    def dinfo(self, what):
        return float(self.lib.orc_dinfo(self.ctx, what))

    def dim(self, cone=0):
        return self.info(2, cone)

    def rank(self, cone=0):
        return self.info(3, cone)

    def pattern(self, cone=0):
        k = self.info(4, cone)
        rows = np.zeros(k, np.int64)
        cols = np.zeros(k, np.int64)
        self.lib.orc_pattern(self.ctx, cone, _i(rows), _i(cols))
        return rows, cols

    def factor(self, which, cone=0):
        n, r = self.dim(cone), self.rank(cone)
        p = self.lib.orc_factor_ptr(self.ctx, which.encode(), cone)
        return np.ctypeslib.as_array(p, shape=(r, n)).T

    def vec(self, which):
        p = self.lib.orc_vec_ptr(self.ctx, which.encode())
        return np.ctypeslib.as_array(p, shape=(self.m,))

    def auv(self, u, v, cone=0):
        out = np.zeros(self.m)
        self.lib.orc_auv(self.ctx, cone, u.encode(), v.encode(), _d(out))
        return out

    def obj_auv(self, u, v, cone=0):
        return float(self.lib.orc_obj_auv(self.ctx, cone, u.encode(), v.encode()))

    def wsum_mulrk(self, w, add_c, x, cone=0):
        n, r = self.dim(cone), self.rank(cone)
        w = np.ascontiguousarray(w, dtype=np.float64)
        out = np.zeros((r, n))
        self.lib.orc_wsum_mulrk(self.ctx, cone, _d(w), int(add_c), x.encode(), _d(out))
        return out.T

    def alm_cal_grad(self, rho):
        return float(self.lib.orc_alm_cal_grad(self.ctx, rho))

    def alm_prepare(self, rho):
        return float(self.lib.orc_alm_prepare(self.ctx, rho))

    def cg_matvec(self, x, no_update, cone=0):
        xin = np.ascontiguousarray(np.asarray(x, dtype=np.float64).T)
        res = np.zeros_like(xin)
        self.lib.orc_cg_matvec(self.ctx, cone, no_update.encode(), _d(xin), _d(res))
        return res.T

    def update_sdp_var_one(self, upd, noupd, rho, tol, maxit, cone=0):
        return int(self.lib.orc_update_sdp_var_one(self.ctx, cone, upd.encode(), noupd.encode(), rho, tol, maxit))

    def alm_inner_iter(self, rho, counter):
        out = np.zeros(8)
        root = int(self.lib.orc_alm_inner_iter(self.ctx, rho, counter, _d(out)))
        return root, dict(tau=out[0], lag_norm_sq=out[1], pinf=out[2], p1=out[3], p2=out[4])
