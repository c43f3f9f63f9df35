"""ctypes binding of liblorads_b200.so (include/lorads_b200.h).

Python is the *host side* here only because the tests and the benchmark are Python; the API mirrors the
reference's own function names for the path (see INTEGRATION.md for the C drop-in).  There is no CPU
fallback: if the CUDA library is missing or no GPU is usable, construction raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .sdpa import Instance

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "liblorads_b200.so")

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int64)


class Params(C.Structure):
    """lb2_params == lorads_params (reference src_semi/lorads.h:82-105)."""

    _fields_ = [
        ("initRho", C.c_double), ("rhoMax", C.c_double), ("rhoCellingALM", C.c_double), ("rhoCellingADMM", C.c_double),
        ("maxALMIter", C.c_int64), ("maxADMMIter", C.c_int64), ("timesLogRank", C.c_double), ("rhoFreq", C.c_int64),
        ("rhoFactor", C.c_double), ("ALMRhoFactor", C.c_double), ("phase1Tol", C.c_double), ("phase2Tol", C.c_double),
        ("timeSecLimit", C.c_double), ("heuristicFactor", C.c_double), ("lbfgsListLength", C.c_int64),
        ("endTauTol", C.c_double), ("endALMSubTol", C.c_double), ("l2Rescaling", C.c_int), ("reoptLevel", C.c_int64),
        ("dyrankLevel", C.c_int64), ("highAccMode", C.c_int), ("verbose", C.c_int),
    ]


class Result(C.Structure):
    _fields_ = [
        ("pObj", C.c_double), ("dObj", C.c_double), ("pInfeasL1", C.c_double), ("dInfeasL1", C.c_double),
        ("pdGap", C.c_double), ("pInfeasInf", C.c_double), ("dInfeasInf", C.c_double),
        ("almOuterIter", C.c_int64), ("almInnerIter", C.c_int64), ("admmIter", C.c_int64), ("cgIter", C.c_int64),
        ("almRho", C.c_double), ("admmRho", C.c_double), ("solveSeconds", C.c_double), ("almSeconds", C.c_double),
        ("admmSeconds", C.c_double), ("status", C.c_int), ("finalRank0", C.c_int64), ("kernelLaunches", C.c_int64),
    ]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


EXPORTS = [
    "lb2_last_error", "lb2_version", "lb2_default_params", "lb2_create", "lb2_set_cone_data", "lb2_preprocess",
    "lb2_determine_rank", "lb2_init_vars", "lb2_destroy", "lb2_comm_unique_id", "lb2_comm_init", "lb2_info", "lb2_dinfo",
    "lb2_get_pattern", "lb2_set_factor", "lb2_get_factor", "lb2_set_vec", "lb2_get_vec", "lb2_auv", "lb2_wsum_mulrk",
    "lb2_alm_cal_grad", "lb2_cg_matvec", "lb2_update_sdp_var_one", "lb2_alm_prepare", "lb2_alm_inner_iter",
    "lb2_time_alm_inner_iters", "lb2_alm_run_host", "lb2_bench_kernel", "lb2_alm_optimize", "lb2_alm_to_admm", "lb2_admm_optimize", "lb2_dual_infeasibility",
    "lb2_solve", "lb2_get_solution", "lb2_reopt", "lb2_average_uv", "lb2_copy_r_to_v", "lb2_get_state", "lb2_set_state", "lb2_host_presolve", "lb2_host_line_search", "lb2_host_rank_rule",
    "lb2_set_lp_data", "lb2_get_lp_vec", "lb2_set_lp_vec", "lb2_admm_init_constr", "lb2_admm_update_var",
    "lb2_layout_build", "lb2_layout_info", "lb2_layout_get", "lb2_layout_free",
    "lb2_read_sdpa", "lb2_sdpa_save", "lb2_sdpa_info", "lb2_sdpa_get", "lb2_sdpa_free", "lb2_sdpa_last_error",
]

_lib = None


def load_library():
    """Loads the CUDA library; raises (never falls back) when it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -m lorads_b200.build` (nvcc, sm_100a). "
            "There is no CPU fallback for the LoRADS kernel layer.")
    lib = C.CDLL(LIB_PATH)
    lib.lb2_last_error.restype = C.c_char_p
    lib.lb2_version.restype = C.c_char_p
    lib.lb2_default_params.argtypes = [C.POINTER(Params)]
    lib.lb2_create.argtypes = [C.POINTER(C.c_void_p), C.c_int64, C.c_int64, _ip, _dp, C.c_int]
    lib.lb2_set_cone_data.argtypes = [C.c_void_p, C.c_int64, _ip, _ip, _dp]
    lib.lb2_set_lp_data.argtypes = [C.c_void_p, C.c_int64, _ip, _ip, _dp]
    lib.lb2_get_lp_vec.argtypes = [C.c_void_p, C.c_char, _dp]
    lib.lb2_set_lp_vec.argtypes = [C.c_void_p, C.c_char, _dp]
    lib.lb2_admm_init_constr.argtypes = [C.c_void_p]
    lib.lb2_admm_update_var.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_int64]
    lib.lb2_preprocess.argtypes = [C.c_void_p]
    lib.lb2_determine_rank.argtypes = [C.c_void_p, C.c_double]
    lib.lb2_init_vars.argtypes = [C.c_void_p, C.c_int64, C.c_double]
    lib.lb2_destroy.argtypes = [C.c_void_p]
    lib.lb2_destroy.restype = None
    lib.lb2_comm_unique_id.argtypes = [C.c_void_p]
    lib.lb2_comm_init.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
    lib.lb2_info.restype = C.c_int64
    lib.lb2_info.argtypes = [C.c_void_p, C.c_int, C.c_int64]
    lib.lb2_dinfo.restype = C.c_double
    lib.lb2_dinfo.argtypes = [C.c_void_p, C.c_int]
    lib.lb2_get_pattern.argtypes = [C.c_void_p, C.c_int64, _ip, _ip]
    lib.lb2_set_factor.argtypes = [C.c_void_p, C.c_char, C.c_int64, _dp]
    lib.lb2_get_factor.argtypes = [C.c_void_p, C.c_char, C.c_int64, _dp]
    lib.lb2_set_vec.argtypes = [C.c_void_p, C.c_char, _dp]
    lib.lb2_get_vec.argtypes = [C.c_void_p, C.c_char, _dp]
    lib.lb2_auv.argtypes = [C.c_void_p, C.c_int64, C.c_char, C.c_char, _dp, _dp]
    lib.lb2_wsum_mulrk.argtypes = [C.c_void_p, C.c_int64, _dp, C.c_int, C.c_char, _dp]
    lib.lb2_alm_cal_grad.argtypes = [C.c_void_p, C.c_double, _dp]
    lib.lb2_cg_matvec.argtypes = [C.c_void_p, C.c_int64, C.c_char, _dp, _dp]
    lib.lb2_update_sdp_var_one.argtypes = [C.c_void_p, C.c_int64, C.c_char, C.c_char, C.c_double, C.c_double, C.c_int64, _ip]
    lib.lb2_alm_prepare.argtypes = [C.c_void_p, C.c_double, _dp]
    lib.lb2_alm_inner_iter.argtypes = [C.c_void_p, C.c_double, C.c_int64, _dp, _ip]
    lib.lb2_time_alm_inner_iters.argtypes = [C.c_void_p, C.c_double, C.c_int64, _dp, _dp]
    lib.lb2_alm_run_host.argtypes = [C.c_void_p, _dp, _dp, C.c_double, C.c_int64, _dp, _dp]
    lib.lb2_bench_kernel.argtypes = [C.c_void_p, C.c_int, C.c_int64, C.c_int, _dp]
    lib.lb2_alm_optimize.argtypes = [C.c_void_p, C.POINTER(Params), C.c_double]
    lib.lb2_alm_to_admm.argtypes = [C.c_void_p, C.POINTER(Params)]
    lib.lb2_admm_optimize.argtypes = [C.c_void_p, C.POINTER(Params), C.c_int64, C.c_double]
    lib.lb2_dual_infeasibility.argtypes = [C.c_void_p]
    lib.lb2_solve.argtypes = [C.c_void_p, C.POINTER(Params), C.POINTER(Result)]
    lib.lb2_get_solution.argtypes = [C.c_void_p, C.c_int64, _dp, _dp]
    lib.lb2_host_presolve.argtypes = [C.c_int64, C.c_int64, _ip, _ip, _dp, _ip, _ip, _ip]
    lib.lb2_host_line_search.restype = C.c_int64
    lib.lb2_host_line_search.argtypes = [C.c_double, _dp, C.c_double, C.c_double, _dp]
    lib.lb2_host_rank_rule.restype = C.c_int64
    lib.lb2_host_rank_rule.argtypes = [C.c_int64, C.c_int64, C.c_int64, C.c_double, _ip]
    lib.lb2_read_sdpa.argtypes = [C.c_char_p, C.POINTER(C.c_void_p)]
    lib.lb2_sdpa_info.restype = C.c_int64
    lib.lb2_sdpa_info.argtypes = [C.c_void_p, C.c_int, C.c_int64]
    lib.lb2_sdpa_get.argtypes = [C.c_void_p, C.c_int64, _ip, _ip, _dp, _dp]
    lib.lb2_sdpa_free.argtypes = [C.c_void_p]
    lib.lb2_sdpa_save.argtypes = [C.c_void_p, C.c_char_p]
    lib.lb2_sdpa_free.restype = None
    lib.lb2_sdpa_last_error.restype = C.c_char_p
    _lib = lib
    return lib


def default_params(**overrides) -> Params:
    p = Params()
    load_library().lb2_default_params(C.byref(p))
    for k, v in overrides.items():
        setattr(p, k, v)
    return p


def _d(a):
    return a.ctypes.data_as(_dp)


def _i(a):
    return a.ctypes.data_as(_ip)


class Lb2Error(RuntimeError):
    pass


def host_presolve(cone, m: int):
    """Host-only pre-solve of one cone: returns (info dict, pattern rows, pattern cols)."""
    lib = load_library()
    info = np.zeros(10, np.int64)
    beg = np.ascontiguousarray(cone.beg, dtype=np.int64)
    idx = np.ascontiguousarray(cone.idx, dtype=np.int64)
    elem = np.ascontiguousarray(cone.elem, dtype=np.float64)
    rc = lib.lb2_host_presolve(cone.n, m, _i(beg), _i(idx), _d(elem), _i(info), None, None)
    if rc != 0:
        raise Lb2Error(lib.lb2_last_error().decode())
    rows = np.zeros(int(info[0]) if not info[1] else 0, np.int64)
    cols = np.zeros_like(rows)
    if not info[1]:
        lib.lb2_host_presolve(cone.n, m, _i(beg), _i(idx), _d(elem), _i(info), _i(rows), _i(cols))
    keys = ["psize", "dense_path", "dense_cone", "n_act", "nnzA", "nnzC", "n_nonzero_coeff", "n_split_rows", "rank_one_objective"]
    return dict(zip(keys, info.tolist())), rows, cols


def host_layout(cone, m: int) -> dict:
    """The device layouts of one cone as numpy arrays, built on the host only (item lists of A and [A;C], the
    constraint table transposed by pattern position, C on the pattern, the symmetric adjacency)."""
    lib = load_library()
    lib.lb2_layout_build.argtypes = [C.c_int64, C.c_int64, _ip, _ip, _dp, C.POINTER(C.c_void_p)]
    lib.lb2_layout_info.argtypes = [C.c_void_p, C.c_int]
    lib.lb2_layout_info.restype = C.c_int64
    lib.lb2_layout_get.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    lib.lb2_layout_free.argtypes = [C.c_void_p]
    h = C.c_void_p()
    beg = np.ascontiguousarray(cone.beg, dtype=np.int64)
    idx = np.ascontiguousarray(cone.idx, dtype=np.int64)
    elem = np.ascontiguousarray(cone.elem, dtype=np.float64)
    if lib.lb2_layout_build(cone.n, m, _i(beg), _i(idx), _d(elem), C.byref(h)) != 0:
        raise Lb2Error(lib.lb2_last_error().decode())
    try:
        info = [int(lib.lb2_layout_info(h, k)) for k in range(12)]
        n, psize, dense, n_act, nnzA, itA, itAC, nadj, tileA, tileAC, nT, nsplit = info
        out = dict(n=n, psize=psize, dense_path=bool(dense), n_act=n_act, nnzA=nnzA, tile_A=tileA, tile_AC=tileAC, n_split_AC=nsplit)

        def geti(which, count):
            a = np.zeros(max(count, 1), np.int32)
            lib.lb2_layout_get(h, which, a.ctypes.data_as(C.c_void_p))
            return a[:count]

        def getd(which, count):
            a = np.zeros(max(count, 1), np.float64)
            lib.lb2_layout_get(h, which, a.ctypes.data_as(C.c_void_p))
            return a[:count]
        npat = 0 if dense else psize
        out.update(act_idx=geti(0, n_act), P_row=geti(1, npat), P_col=geti(2, npat),
                   A_ptr=geti(3, n_act + 1), A_irow=geti(4, itA), A_icol=geti(5, itA), A_coef=getd(20, itA),
                   AC_ptr=geti(6, n_act + 2), AC_irow=geti(7, itAC), AC_icol=geti(8, itAC), AC_coef=getd(21, itAC),
                   T_con=geti(10, nT), T_val=getd(22, nT), C_onP=getd(23, psize), c_rank1=float(getd(24, 1)[0]))
        if not dense:
            out.update(T_ptr=geti(9, psize + 1), adj_ptr=geti(11, n + 1), adj_col=geti(12, nadj), adj_pos=geti(13, nadj))
            # vertex-centric class split (layout.hpp VcLayout)
            vc_on, nu, nl, n_single, nnz_res, nd = (int(lib.lb2_layout_info(h, k)) for k in range(30, 36))
            out.update(vc_on=bool(vc_on), vc_n_single=n_single, vc_nnz_res=nnz_res,
                       vc_order=geti(30, n), vc_order_l=geti(31, n),
                       vc_u_ptr=geti(32, n + 1), vc_u_mid=geti(33, n), vc_u_col=geti(34, nu), vc_u_tag=geti(35, nu),
                       vc_u_val=getd(36, nu), vc_d_con=geti(37, nd), vc_d_coef=getd(38, nd),
                       vc_l_ptr=geti(39, n + 1), vc_l_row=geti(40, nl), vc_l_col=geti(50, nl), vc_l_con=geti(41, nl), vc_l_coef=getd(42, nl),
                       vc_Tr_ptr=geti(43, psize + 1), vc_Tr_con=geti(44, nnz_res), vc_Tr_val=getd(45, nnz_res),
                       vc_res_ptr=geti(46, n_act + 1), vc_res_irow=geti(47, nnz_res), vc_res_icol=geti(48, nnz_res),
                       vc_res_coef=getd(49, nnz_res))
        return out
    finally:
        lib.lb2_layout_free(h)


def read_sdpa(path: str, save_binary: str | None = None) -> Instance:
    """Fast .dat-s ingest through the library (same arrays as the reference's LReadSDPA); `path` may also be a binary
    instance cache written by an earlier call with `save_binary`."""
    from .sdpa import Cone
    lib = load_library()
    h = C.c_void_p()
    if lib.lb2_read_sdpa(path.encode(), C.byref(h)) != 0:
        raise Lb2Error(lib.lb2_sdpa_last_error().decode())
    try:
        if save_binary is not None and lib.lb2_sdpa_save(h, save_binary.encode()) != 0:
            raise Lb2Error(lib.lb2_sdpa_last_error().decode())
        m = int(lib.lb2_sdpa_info(h, 0, 0))
        nblk = int(lib.lb2_sdpa_info(h, 1, 0))
        b = np.zeros(m)
        lib.lb2_sdpa_get(h, -1, None, None, None, _d(b))
        cones = []
        for k in range(nblk):
            n, nnz = int(lib.lb2_sdpa_info(h, 2, k)), int(lib.lb2_sdpa_info(h, 3, k))
            beg, idx, elem = np.zeros(m + 2, np.int64), np.zeros(nnz, np.int64), np.zeros(nnz)
            lib.lb2_sdpa_get(h, k, _i(beg), _i(idx), _d(elem), None)
            cones.append(Cone(n=n, beg=beg, idx=idx, elem=elem))
        lp = None
        n_lp = int(lib.lb2_sdpa_info(h, 4, 0))
        if n_lp > 0:
            from .sdpa import LpBlock
            nnz = int(lib.lb2_sdpa_info(h, 3, nblk))
            beg, idx, elem = np.zeros(m + 2, np.int64), np.zeros(nnz, np.int64), np.zeros(nnz)
            lib.lb2_sdpa_get(h, nblk, _i(beg), _i(idx), _d(elem), None)
            lp = LpBlock(n=n_lp, beg=beg, idx=idx, elem=elem)
        return Instance(m=m, b=b, cones=cones, name=path, lp=lp)
    finally:
        lib.lb2_sdpa_free(h)


def host_line_search(rho: float, sums, p1: float, p2: float, tau0: float = 0.0):
    lib = load_library()
    sums = np.ascontiguousarray(sums, dtype=np.float64)
    tau = C.c_double(tau0)
    root = lib.lb2_host_line_search(rho, _d(sums), p1, p2, C.byref(tau))
    return int(root), tau.value


def host_rank_rule(n: int, n_nonzero_coeff: int, n_cones: int, times_log_rank: float = 2.0):
    lib = load_library()
    cap = C.c_int64(0)
    r = lib.lb2_host_rank_rule(n, n_nonzero_coeff, n_cones, times_log_rank, C.byref(cap))
    return int(r), int(cap.value)


class Solver:
    """Device-resident LoRADS solver state built from reader-format arrays (main.c:266-304 sequence)."""

    def __init__(self, inst: Instance, device: int = 0, times_log_rank: float = 2.0, lbfgs_len: int = 2,
                 init_rho: float = 0.0, comm=None):
        self.lib = load_library()
        self.h = C.c_void_p()
        self.m = inst.m
        dims = np.asarray(inst.blk_dims, dtype=np.int64)
        b = np.ascontiguousarray(inst.b, dtype=np.float64)
        self._ck(self.lib.lb2_create(C.byref(self.h), inst.m, len(inst.cones), _i(dims), _d(b), device))
        for k, cone in enumerate(inst.cones):
            beg = np.ascontiguousarray(cone.beg, dtype=np.int64)
            idx = np.ascontiguousarray(cone.idx, dtype=np.int64)
            elem = np.ascontiguousarray(cone.elem, dtype=np.float64)
            self._ck(self.lib.lb2_set_cone_data(self.h, k, _i(beg), _i(idx), _d(elem)))
        self.n_lp = 0
        if getattr(inst, "lp", None) is not None:
            lp = inst.lp
            self.n_lp = int(lp.n)
            self._ck(self.lib.lb2_set_lp_data(self.h, lp.n, _i(np.ascontiguousarray(lp.beg, dtype=np.int64)),
                                              _i(np.ascontiguousarray(lp.idx, dtype=np.int64)),
                                              _d(np.ascontiguousarray(lp.elem, dtype=np.float64))))
        self._ck(self.lib.lb2_preprocess(self.h))
        self._ck(self.lib.lb2_determine_rank(self.h, float(times_log_rank)))
        if comm is not None:
            uid, rank, world = comm
            self._ck(self.lib.lb2_comm_init(self.h, uid, rank, world))
        self._ck(self.lib.lb2_init_vars(self.h, lbfgs_len, float(init_rho)))
        self.n_cones = len(inst.cones)

    def _ck(self, rc):
        if rc != 0:
            raise Lb2Error(f"lorads_b200 error {rc}: {self.lib.lb2_last_error().decode()}")

    def close(self):
        if getattr(self, "h", None) and self.h:
            self.lib.lb2_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- queries
    def info(self, what: int, cone: int = 0) -> int:
        return int(self.lib.lb2_info(self.h, what, cone))

    def dinfo(self, what: int) -> float:
        return float(self.lib.lb2_dinfo(self.h, what))

    def dim(self, cone=0):
        return self.info(2, cone)

    def rank(self, cone=0):
        return self.info(3, cone)

    @property
    def launches(self) -> int:
        return self.info(11)

    def pattern(self, cone=0):
        k = self.info(4, cone)
        rows = np.zeros(k, np.int64)
        cols = np.zeros(k, np.int64)
        self._ck(self.lib.lb2_get_pattern(self.h, cone, _i(rows), _i(cols)))
        return rows, cols

    # ---- state transfer
    def set_factor(self, which: str, x: np.ndarray, cone=0):
        """x: n x r array (any memory order); sent as the reference's column-major matElem."""
        cm = np.ascontiguousarray(np.asarray(x, dtype=np.float64).T)
        assert cm.shape == (self.rank(cone), self.dim(cone)), cm.shape
        self._ck(self.lib.lb2_set_factor(self.h, which.encode(), cone, _d(cm)))

    def get_factor(self, which: str, cone=0) -> np.ndarray:
        cm = np.zeros((self.rank(cone), self.dim(cone)))
        self._ck(self.lib.lb2_get_factor(self.h, which.encode(), cone, _d(cm)))
        return cm.T

    def set_vec(self, which: str, v: np.ndarray):
        v = np.ascontiguousarray(v, dtype=np.float64)
        assert v.shape == (self.m,)
        self._ck(self.lib.lb2_set_vec(self.h, which.encode(), _d(v)))

    def get_vec(self, which: str) -> np.ndarray:
        v = np.zeros(self.m)
        self._ck(self.lib.lb2_get_vec(self.h, which.encode(), _d(v)))
        return v

    def get_lp(self, which: str) -> np.ndarray:
        """LP vectors: 'R' rLp, 'U' uLp, 'V' vLp, 'G' gradLp, 'x' product held in constrValLP, 'c' objective."""
        v = np.zeros(self.n_lp)
        self._ck(self.lib.lb2_get_lp_vec(self.h, which.encode(), _d(v)))
        return v

    def set_lp(self, which: str, v: np.ndarray):
        v = np.ascontiguousarray(v, dtype=np.float64)
        assert v.shape == (self.n_lp,)
        self._ck(self.lib.lb2_set_lp_vec(self.h, which.encode(), _d(v)))

    # ---- hot-path operators (names follow oracle/ref.py, i.e. the reference functions)
    def auv(self, u: str, v: str, cone=0, with_obj=False):
        out = np.zeros(self.m)
        obj = C.c_double(0.0)
        self._ck(self.lib.lb2_auv(self.h, cone, u.encode(), v.encode(), _d(out), C.byref(obj) if with_obj else None))
        return (out, obj.value) if with_obj else out

    def obj_auv(self, u: str, v: str, cone=0) -> float:
        return self.auv(u, v, cone, with_obj=True)[1]

    def wsum_mulrk(self, w: np.ndarray, add_c: bool, x: str, cone=0) -> np.ndarray:
        w = np.ascontiguousarray(w, dtype=np.float64)
        out = np.zeros((self.rank(cone), self.dim(cone)))
        self._ck(self.lib.lb2_wsum_mulrk(self.h, cone, _d(w), int(add_c), x.encode(), _d(out)))
        return out.T

    def alm_cal_grad(self, rho: float) -> float:
        lag = C.c_double(0.0)
        self._ck(self.lib.lb2_alm_cal_grad(self.h, rho, C.byref(lag)))
        return lag.value

    def cg_matvec(self, x: np.ndarray, no_update: str, cone=0) -> np.ndarray:
        xin = np.ascontiguousarray(np.asarray(x, dtype=np.float64).T)
        res = np.zeros_like(xin)
        self._ck(self.lib.lb2_cg_matvec(self.h, cone, no_update.encode(), _d(xin), _d(res)))
        return res.T

    def update_sdp_var_one(self, upd: str, noupd: str, rho: float, tol: float, maxit: int, cone=0) -> int:
        it = C.c_int64(0)
        self._ck(self.lib.lb2_update_sdp_var_one(self.h, cone, upd.encode(), noupd.encode(), rho, tol, maxit, C.byref(it)))
        return it.value

    def admm_init_constr(self):
        self._ck(self.lib.lb2_admm_init_constr(self.h))

    def admm_update_var(self, rho: float, tol: float, maxit: int):
        self._ck(self.lib.lb2_admm_update_var(self.h, rho, tol, maxit))

    def alm_prepare(self, rho: float) -> float:
        lag = C.c_double(0.0)
        self._ck(self.lib.lb2_alm_prepare(self.h, rho, C.byref(lag)))
        return lag.value

    def alm_inner_iter(self, rho: float, counter: int):
        out = np.zeros(8)
        root = C.c_int64(0)
        self._ck(self.lib.lb2_alm_inner_iter(self.h, rho, counter, _d(out), C.byref(root)))
        return root.value, dict(tau=out[0], lag_norm_sq=out[1], pinf=out[2], p1=out[3], p2=out[4])

    def time_alm_inner_iters(self, rho: float, iters: int):
        """Returns (seconds by CUDA events, iterations actually done)."""
        out = np.zeros(8)
        sec = C.c_double(0.0)
        self._ck(self.lib.lb2_time_alm_inner_iters(self.h, rho, iters, _d(out), C.byref(sec)))
        return sec.value, int(out[5])

    def alm_run_host(self, R_in: np.ndarray, lam_in: np.ndarray, rho: float, iters: int, R_out: np.ndarray):
        """R_in / R_out: flat column-major concatenation over cones (host, ideally pinned)."""
        out = np.zeros(8)
        self._ck(self.lib.lb2_alm_run_host(self.h, _d(R_in), _d(lam_in), rho, iters, _d(R_out), _d(out)))
        return int(out[5]), out

    def dual_infeasibility(self) -> float:
        """DIMACS dual infeasibility sum_cones |min(lambda_min(C - A^*(lambda)), 0)| / (1 + |C|_1) at the current
        multipliers (calculate_dual_infeasibility_solver, lorads_solver.c:1007-1037)."""
        self._ck(self.lib.lb2_dual_infeasibility(self.h))
        return self.dinfo(11)

    def bench_kernel(self, which: int, reps: int, flush_l2: bool = False) -> float:
        ms = C.c_double(0.0)
        self._ck(self.lib.lb2_bench_kernel(self.h, which, reps, int(flush_l2), C.byref(ms)))
        return ms.value

    # ---- phases
    def solve(self, params: Params | None = None) -> dict:
        p = params if params is not None else default_params()
        res = Result()
        self._ck(self.lib.lb2_solve(self.h, C.byref(p), C.byref(res)))
        return res.as_dict()

    def solution(self, cone=0):
        R = np.zeros((self.rank(cone), self.dim(cone)))
        lam = np.zeros(self.m)
        self._ck(self.lib.lb2_get_solution(self.h, cone, _d(R), _d(lam)))
        return R.T, lam
