"""TEST INFRASTRUCTURE ONLY -- ctypes binding of the CPU oracle (oracle/liboracle_phys.so).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
PARITY UNPINNED (see towr_oracle.hpp): the reference has no golden vectors for this path.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

STAGES = {"1.1": 0, "1.2": 1, "2.1": 2, "2.2": 3, "3": 4, "4": 5}


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "liboracle_phys.so")
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".cpp", ".hpp"))]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "liboracle_phys.so")
        if not os.path.exists(so) or os.path.exists("/usr/bin/make") or True:
            try:
                so = build()
            except Exception:
                if not os.path.exists(so):
                    raise
        L = C.CDLL(so)
        dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
        L.chdo_create.restype = C.c_void_p
        L.chdo_create.argtypes = [C.c_int, C.c_int, C.c_double, dp, dp, C.c_double, C.c_double, C.c_double, C.c_double,
                                  dp, dp, dp, dp, dp, dp, ip, ip, dp, dp]
        for name, res, args in [
            ("chdo_destroy", None, [C.c_void_p]),
            ("chdo_set_stage", None, [C.c_void_p, C.c_int]),
            ("chdo_n", C.c_int, [C.c_void_p]),
            ("chdo_m", C.c_int, [C.c_void_p]),
            ("chdo_total_time", C.c_double, [C.c_void_p]),
            ("chdo_get_x", None, [C.c_void_p, dp]),
            ("chdo_set_x", None, [C.c_void_p, dp]),
            ("chdo_var_set_sizes", None, [C.c_void_p, ip]),
            ("chdo_var_bounds", None, [C.c_void_p, dp, dp]),
            ("chdo_num_constraint_sets", C.c_int, [C.c_void_p]),
            ("chdo_constraint_set_rows", C.c_int, [C.c_void_p, C.c_int]),
            ("chdo_constraint_set_name", C.c_char_p, [C.c_void_p, C.c_int]),
            ("chdo_con_bounds", None, [C.c_void_p, dp, dp]),
            ("chdo_cost", C.c_double, [C.c_void_p]),
            ("chdo_num_costs", C.c_int, [C.c_void_p]),
            ("chdo_cost_term", C.c_double, [C.c_void_p, C.c_int]),
            ("chdo_grad", None, [C.c_void_p, dp]),
            ("chdo_cons", None, [C.c_void_p, dp]),
            ("chdo_jac", C.c_int, [C.c_void_p, ip, ip, dp]),
            ("chdo_cost_hessian", C.c_int, [C.c_void_p, ip, ip, dp]),
            ("chdo_lag_hessian", C.c_int, [C.c_void_p, dp, ip, ip, dp]),
            ("chdo_row_times", None, [C.c_void_p, dp]),
            ("chdo_forget", None, [C.c_void_p]),
            ("chdo_dur_blocks", C.c_int, [C.c_void_p, ip, ip]),
            ("chdo_var_times", None, [C.c_void_p, dp, dp]),
            ("chdo_solve_stage", C.c_int, [C.c_void_p, C.c_int, C.c_int, dp, C.c_int]),
            ("chdo_sample", C.c_int, [C.c_void_p, dp]),
            ("chdo_spline_point", None, [C.c_void_p, C.c_int, C.c_double, dp]),
            ("chdo_spline_num_polys", C.c_int, [C.c_void_p, C.c_int]),
            ("chdo_spline_poly_durations", None, [C.c_void_p, C.c_int, dp]),
            ("chdo_spline_jac", None, [C.c_void_p, C.c_int, C.c_double, C.c_int, dp]),
            ("chdo_euler", None, [C.c_void_p, C.c_double, dp, dp, dp]),
        ]:
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _LIB = L
    return _LIB


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


class OracleProblem:
    """One sequence's NLP as the reference would assemble it (ifopt stacking order)."""

    def __init__(self, p, weights=(0.4, 1.7, 0.3, 0.1, 0.1)):
        L = lib()
        f64 = lambda a: np.ascontiguousarray(a, dtype=np.float64)
        self._keep = [f64(p.hip_left), f64(p.hip_right), f64(p.inertia), f64(p.base_lin), f64(p.base_ang), f64(p.ee_pos),
                      f64(p.floor_normal), f64(p.floor_point),
                      np.ascontiguousarray(p.ee_start_contact, dtype=np.int32),
                      np.ascontiguousarray([len(d) for d in p.ee_durations], dtype=np.int32),
                      f64(np.concatenate(p.ee_durations)), f64(weights)]
        k = self._keep
        self.n_ee = p.n_ee
        self.F = p.n_frames
        self.n_dur = sum(len(d) - 1 for d in p.ee_durations)
        self.h = L.chdo_create(p.n_frames, p.n_ee, p.dt, _dp(k[0]), _dp(k[1]), p.max_leg_length, p.max_heel_length,
                               p.heel_dist, p.body_mass, _dp(k[2]), _dp(k[3]), _dp(k[4]), _dp(k[5]), _dp(k[6]), _dp(k[7]),
                               _ip(k[8]), _ip(k[9]), _dp(k[10]), _dp(k[11]))
        self.L = L

    def __del__(self):
        try:
            self.L.chdo_forget(self.h)
            self.L.chdo_destroy(self.h)
        except Exception:
            pass

    def set_stage(self, st):
        self.L.chdo_set_stage(self.h, STAGES.get(st, st))

    @property
    def n(self):
        return self.L.chdo_n(self.h)

    @property
    def m(self):
        return self.L.chdo_m(self.h)

    @property
    def total_time(self):
        return self.L.chdo_total_time(self.h)

    def get_x(self):
        x = np.zeros(self.n)
        self.L.chdo_get_x(self.h, _dp(x))
        return x

    def set_x(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        assert x.shape == (self.n,)
        self.L.chdo_set_x(self.h, _dp(x))

    def var_set_sizes(self):
        out = np.zeros(2 + 3 * self.n_ee, dtype=np.int32)
        self.L.chdo_var_set_sizes(self.h, _ip(out))
        return out

    def var_bounds(self):
        lo, hi = np.zeros(self.n), np.zeros(self.n)
        self.L.chdo_var_bounds(self.h, _dp(lo), _dp(hi))
        return lo, hi

    def constraint_sets(self):
        return [(self.L.chdo_constraint_set_name(self.h, i).decode(), self.L.chdo_constraint_set_rows(self.h, i))
                for i in range(self.L.chdo_num_constraint_sets(self.h))]

    def con_bounds(self):
        lo, hi = np.zeros(self.m), np.zeros(self.m)
        self.L.chdo_con_bounds(self.h, _dp(lo), _dp(hi))
        return lo, hi

    def cost(self):
        return self.L.chdo_cost(self.h)

    def cost_terms(self):
        return np.array([self.L.chdo_cost_term(self.h, i) for i in range(self.L.chdo_num_costs(self.h))])

    def grad(self):
        g = np.zeros(self.n)
        self.L.chdo_grad(self.h, _dp(g))
        return g

    def cons(self):
        g = np.zeros(self.m)
        self.L.chdo_cons(self.h, _dp(g))
        return g

    def jac(self):
        """scipy CSR (m x n)"""
        import scipy.sparse as sp
        nnz = self.L.chdo_jac(self.h, None, None, None)
        ri, ci, v = np.zeros(nnz, np.int32), np.zeros(nnz, np.int32), np.zeros(nnz)
        self.L.chdo_jac(self.h, _ip(ri), _ip(ci), _dp(v))
        return sp.csr_matrix((v, (ri, ci)), shape=(self.m, self.n))

    def cost_hessian(self):
        import scipy.sparse as sp
        nnz = self.L.chdo_cost_hessian(self.h, None, None, None)
        ri, ci, v = np.zeros(nnz, np.int32), np.zeros(nnz, np.int32), np.zeros(nnz)
        self.L.chdo_cost_hessian(self.h, _ip(ri), _ip(ci), _dp(v))
        return sp.csr_matrix((v, (ri, ci)), shape=(self.n, self.n))

    def lag_hessian(self, y):
        import scipy.sparse as sp
        y = np.ascontiguousarray(y, dtype=np.float64)
        assert y.shape == (self.m,)
        nnz = self.L.chdo_lag_hessian(self.h, _dp(y), None, None, None)
        ri, ci, v = np.zeros(nnz, np.int32), np.zeros(nnz, np.int32), np.zeros(nnz)
        self.L.chdo_lag_hessian(self.h, _dp(y), _ip(ri), _ip(ci), _dp(v))
        return sp.csr_matrix((v, (ri, ci)), shape=(self.n, self.n))

    def row_times(self):
        t = np.zeros(self.m)
        self.L.chdo_row_times(self.h, _dp(t))
        return t

    def var_times(self):
        t0, t1 = np.zeros(self.n), np.zeros(self.n)
        self.L.chdo_var_times(self.h, _dp(t0), _dp(t1))
        return t0, t1

    def solve_stage(self, stage, max_iter=None, verbose=False):
        """chd-ipm on the CPU (oracle/ipm_oracle.cpp).  Returns a dict like PhysBatch.solve_stage."""
        st = STAGES.get(stage, stage)
        if max_iter is None:
            max_iter = {0: 7000, 1: 7000, 2: 7000, 3: 2500, 4: 2000, 5: 7000}[st]  # phys_optim.cpp:571,640,652,706,743
        stats = np.zeros(12)
        status = self.L.chdo_solve_stage(self.h, st, int(max_iter), _dp(stats), int(verbose))
        return dict(status=status, iters=int(stats[8]), f=stats[0], E0=stats[1], viol=stats[2], dual=stats[3],
                    compl=stats[4], mu=stats[5], delta_w=stats[6], ls_fail=int(stats[7]), Na=int(stats[9]),
                    nb=int(stats[10]), w=int(stats[11]))

    def solve(self):
        """Staged schedule of phys_optim.cpp:554-749: 1.1, 1.2 | 2.1, 2.2 | 3 (| 4 only if stage 3 failed, :713-749)
        with the three SaveSolution snapshots.  `stages` lists the solved stages in order; `stage_ids` names them."""
        out = {}
        res, ids = [], []
        for st in ("1.1", "1.2"):
            res.append(self.solve_stage(st)), ids.append(st)
        out["no_dynamics"] = self.sample()
        for st in ("2.1", "2.2"):
            res.append(self.solve_stage(st)), ids.append(st)
        out["dynamics"] = self.sample()
        dyn_ok = res[3]["status"] == 0
        if self.n_dur <= 96:                                 # CHD_MAX_DUR of the product (csrc/chd_core.h): beyond it stage 3 is not attempted
            res.append(self.solve_stage("3")), ids.append("3")
            dur_ok = res[-1]["status"] == 0                  # :709
        else:
            dur_ok = False
        if not dur_ok:
            res.append(self.solve_stage("4")), ids.append("4")
            dur_ok = res[-1]["status"] == 0                  # :746
        out["durations"] = self.sample()
        out["stages"] = res
        out["stage_ids"] = ids
        out["success"] = (dyn_ok, dur_ok)
        return out

    def sample(self):
        nf = self.L.chdo_sample(self.h, None)
        out = np.zeros((nf, 6 + 7 * self.n_ee))
        got = self.L.chdo_sample(self.h, _dp(out))
        return out[:got]

    def spline_point(self, sid, t):
        out = np.zeros(9)
        self.L.chdo_spline_point(self.h, sid, float(t), _dp(out))
        return out.reshape(3, 3)

    def spline_poly_durations(self, sid):
        out = np.zeros(self.L.chdo_spline_num_polys(self.h, sid))
        self.L.chdo_spline_poly_durations(self.h, sid, _dp(out))
        return out

    def spline_jac(self, sid, t, deriv, rows):
        out = np.zeros((3, rows))
        self.L.chdo_spline_jac(self.h, sid, float(t), deriv, _dp(out))
        return out

    def euler(self, t):
        R, w, wd = np.zeros(9), np.zeros(3), np.zeros(3)
        self.L.chdo_euler(self.h, float(t), _dp(R), _dp(w), _dp(wd))
        return R.reshape(3, 3), w, wd
