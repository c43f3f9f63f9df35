"""oracle/lp_restate.py -- TEST INFRASTRUCTURE, not product code.

numpy restatement of the reference's LP-cone operators (the LP block of an SDPA file: x_j = u_j v_j >= 0, columns
a_j of the m x nLp constraint matrix).  Pinned on the golden vectors the compiled reference produced with its own
LP function set (tests/golden/lp_*.npz, tests/test_oracle_lp_golden.py); only tests/ may import it.

Reference functions restated (src_semi/...):
  lp_cone_proc / lp_cone_presolve          data/lorads_lp_conic.c:14-113      -> LpOracle.__init__
  lp_cone_AUV + LORADSInitConstrValSumLP   lorads_lp_conic.c:164-167, lorads_alg/lorads_alg_common.c:144-158 -> constr
  lp_cone_objAUV                           lorads_lp_conic.c:186-195          -> obj
  ALMSetGradLP / ALMCalGradLP              lorads_alg/lorads_alm.c:56-99      -> grad
  LORADSUpdateLPVarOne + LORADSUpdateSDPLPVar  lorads_alg/lorads_admm.c:595-628, lorads_alg_common.c:225-249 -> sweep
  calculate_dual_infeasibility_solver (LP part)  data/lorads_solver.c:1015-1023  -> dual_infeasibility
  LORADSNrm1Obj / Nrm2Obj / NrmInfObj (LP part)  lorads_solver.c:149-183, lorads_alg_common.c:300-316 -> obj_norm_terms
"""
import numpy as np


class LpOracle:
    def __init__(self, lp, m: int):
        self.n, self.m = int(lp.n), int(m)
        beg, idx, elem = np.asarray(lp.beg), np.asarray(lp.idx), np.asarray(lp.elem)
        self.c = np.zeros(self.n)
        self.c[idx[beg[0]:beg[1]]] = elem[beg[0]:beg[1]]
        self.cols = [[] for _ in range(self.n)]             # column j: [(row, value)], rows ascending
        for i in range(self.m):
            for k in range(beg[i + 1], beg[i + 2]):
                self.cols[idx[k]].append((i, elem[k]))
        self.nrm2sq = np.array([np.sqrt(sum(a * a for _, a in col)) ** 2 for col in self.cols])

    def constr(self, u, v):
        """sum_j constrValLP[j] = sum_j a_j (u_j v_j)"""
        out = np.zeros(self.m)
        for j, col in enumerate(self.cols):
            uv = u[j] * v[j]
            for i, a in col:
                out[i] += a * uv
        return out

    def obj(self, u, v):
        return float(np.dot(self.c, u * v))

    def grad(self, w, r):
        """gradLp_j = 2 (c_j + a_j^T w) r_j with w = -lambda - rho b + rho constrValSum"""
        g = np.zeros(self.n)
        for j, col in enumerate(self.cols):
            s = self.c[j]
            for i, a in col:
                s += a * w[i]
            g[j] = 2 * s * r[j]
        return g

    def sweep(self, rho, b, lam, cvs, x, u, v):
        """In-place Gauss-Seidel sweep over the LP columns; x[j] is the product held in constrValLP[j]."""
        for j, col in enumerate(self.cols):
            for upd, noupd in ((u, v), (v, u)):
                w = self.c[j]
                for i, a in col:
                    w += a * (((cvs[i] - b[i]) - a * x[j]) * rho - lam[i])
                M2 = w * noupd[j]
                M2 = M2 - rho * noupd[j]
                upd[j] = (-1.0 * M2 / rho) / (1 + self.nrm2sq[j] * noupd[j] * noupd[j])
                xn = u[j] * v[j]
                for i, a in col:
                    cvs[i] = (cvs[i] - a * x[j]) + a * xn
                x[j] = xn

    def dual_infeasibility(self, lam):
        t = 0.0
        for j, col in enumerate(self.cols):
            s = self.c[j]
            for i, a in col:
                s += a * (-lam[i])
            t += abs(min(s, 0.0))
        return t

    def obj_norm_terms(self):
        """(|c|_1, its contribution to the squared 2-norm -- the reference squares the 1-norm --, inf-norm term of
        the -DUNDER_BLAS build: the entry AFTER the first largest one, clamped)"""
        l1 = float(np.abs(self.c).sum())
        arg = int(np.argmax(np.abs(self.c)))
        return l1, l1 * l1, float(abs(self.c[min(arg + 1, self.n - 1)]))
