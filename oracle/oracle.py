"""ctypes front-end of the CPU oracle (oracle/liboracle.so).

*** TEST INFRASTRUCTURE ONLY. ***  Imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs -- never by the product package.  See oracle/oracle.hpp.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("oracle_capi.cpp", "oracle.hpp", "oracle_grad.hpp")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "liboracle.so"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(so):
            build()
        L = C.CDLL(so)
        L.orc_model_create.restype = C.c_void_p
        L.orc_dir_derivative.restype = C.c_double
        L.orc_birth_death.restype = C.c_double
        L.orc_digamma.restype = C.c_double
        _LIB = L
    return _LIB


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def _i32(x):
    return np.ascontiguousarray(x, dtype=np.int32)


def _f64(x):
    return np.ascontiguousarray(x, dtype=np.float64)


class Oracle:
    """CPU oracle for one model.  `md` is any object with the ModelDesc attributes (duck-typed so
    that oracle/ has no import of the product package)."""

    def __init__(self, md):
        L = lib()
        self.N = len(md.parent)
        self.S = 5 + 2 * self.N
        self.K = self.N - 2
        self.calibrations_available = len(md.cal_node) > 0
        parent, c0, c1 = _i32(md.parent), _i32(md.child0), _i32(md.child1)
        mu, prec = _f64(md.mean), _f64(md.precision).reshape(-1)
        if md.likelihood == 2:
            mu, prec = np.zeros(1), np.zeros(1)
        sp_row = _i32(getattr(md, "sparse_row", np.zeros(0, np.int32)))
        sp_col = _i32(getattr(md, "sparse_col", np.zeros(0, np.int32)))
        sp_val = _f64(getattr(md, "sparse_val", np.zeros(0)))
        if md.likelihood == 3:
            prec = np.zeros(1)
        arrs = [_i32(md.cal_node), _f64(md.cal_lo), _f64(md.cal_lo_p), _f64(md.cal_hi), _f64(md.cal_hi_p),
                _i32(md.con_young), _i32(md.con_old), _f64(md.con_p),
                _i32(md.brace_off), _i32(md.brace_node), _f64(md.brace_sd)]
        self._keep = [parent, c0, c1, mu, prec, sp_row, sp_col, sp_val] + arrs
        D, I = C.c_double, C.c_int
        self.h = C.c_void_p(L.orc_model_create(
            I(self.N), _p(parent, I), _p(c0, I), _p(c1, I), I(self.K), _p(mu, D), _p(prec, D),
            D(md.logdet_sigma), I(md.clock_model), I(md.likelihood), D(md.ht),
            I(len(md.cal_node)), _p(arrs[0], I), _p(arrs[1], D), _p(arrs[2], D), _p(arrs[3], D), _p(arrs[4], D),
            I(len(md.con_young)), _p(arrs[5], I), _p(arrs[6], I), _p(arrs[7], D),
            I(len(md.brace_sd)), _p(arrs[8], I), _p(arrs[9], I), _p(arrs[10], D),
            I(len(sp_val)), _p(sp_row, I), _p(sp_col, I), _p(sp_val, D)))
        if not self.h:
            raise ValueError("oracle: model rejected (root not bifurcating?)")
        self.mask = self.get_mask(self.calibrations_available)

    def __del__(self):
        try:
            if getattr(self, "h", None):
                lib().orc_model_destroy(self.h)
        except Exception:
            pass

    def branch_index(self) -> np.ndarray:
        out = np.empty(self.N, np.int32)
        lib().orc_branch_index(self.h, _p(out, C.c_int))
        return out

    def get_mask(self, calibrations_available: bool) -> np.ndarray:
        out = np.empty(self.S, np.uint8)
        lib().orc_mask(self.h, C.c_int(int(calibrations_available)), _p(out, C.c_uint8))
        return out

    def to_vector(self, x: np.ndarray) -> np.ndarray:
        x = _f64(x)
        th = np.empty(self.S)
        D = lib().orc_to_vector(self.h, _p(self.mask, C.c_uint8), _p(x, C.c_double), _p(th, C.c_double))
        return th[:D].copy()

    def from_vector(self, x: np.ndarray, theta: np.ndarray) -> np.ndarray:
        x, theta = _f64(x), _f64(theta)
        out = np.empty(self.S)
        lib().orc_from_vector(self.h, _p(self.mask, C.c_uint8), _p(x, C.c_double), _p(theta, C.c_double),
                              C.c_int(len(theta)), _p(out, C.c_double))
        return out

    def eval(self, states: np.ndarray, nthreads: int = 1, generic: bool = False):
        """-> (out [B,7] = lnA, lnB, lnC, lnPrior, lnLik, lnJac, lnPost; status [B])"""
        X = _f64(states).reshape(-1, self.S)
        B = X.shape[0]
        out = np.empty((B, 7))
        st = np.empty(B, np.int32)
        if generic:
            lib().orc_eval_generic(self.h, C.c_int(B), _p(X, C.c_double), _p(out, C.c_double), _p(st, C.c_int))
        else:
            lib().orc_eval(self.h, C.c_int(B), _p(X, C.c_double), _p(out, C.c_double), _p(st, C.c_int),
                           C.c_int(nthreads))
        return out, st

    def eval_grad(self, states: np.ndarray, nthreads: int = 1):
        """value + analytic gradient (CPU port) -> (out [B,7], grad [B,S], status [B])"""
        X = _f64(states).reshape(-1, self.S)
        B = X.shape[0]
        out = np.empty((B, 7))
        grad = np.empty((B, self.S))
        st = np.empty(B, np.int32)
        lib().orc_eval_grad(self.h, C.c_int(B), _p(X, C.c_double), _p(self.mask, C.c_uint8), _p(out, C.c_double),
                            _p(grad, C.c_double), _p(st, C.c_int), C.c_int(nthreads))
        return out, grad, st

    def grad_dual(self, state: np.ndarray) -> np.ndarray:
        """gradient ground truth: forward-mode duals through the restated reference code"""
        x = _f64(state)
        g = np.empty(self.S)
        lib().orc_grad_dual(self.h, _p(x, C.c_double), _p(self.mask, C.c_uint8), _p(g, C.c_double))
        return g

    def dir_derivative(self, state: np.ndarray, direction: np.ndarray):
        x, d = _f64(state), _f64(direction)
        val = C.c_double()
        dd = lib().orc_dir_derivative(self.h, _p(x, C.c_double), _p(d, C.c_double), C.byref(val))
        return float(dd), float(val.value)


def birth_death(child0, child1, br, la, mu, rho, condition_on_mrca: bool) -> float:
    c0, c1, t = _i32(child0), _i32(child1), _f64(br)
    return float(lib().orc_birth_death(C.c_int(len(c0)), _p(c0, C.c_int), _p(c1, C.c_int), _p(t, C.c_double),
                                       C.c_double(la), C.c_double(mu), C.c_double(rho),
                                       C.c_int(int(condition_on_mrca))))


def compute_de(la, mu, rho, dt, e0, nearcrit=False):
    out = np.empty(2)
    lib().orc_compute_de(C.c_double(la), C.c_double(mu), C.c_double(rho), C.c_double(dt), C.c_double(e0),
                         C.c_int(int(nearcrit)), _p(out, C.c_double))
    return float(out[0]), float(out[1])


def digamma(x: float) -> float:
    return float(lib().orc_digamma(C.c_double(x)))


def max_threads() -> int:
    return int(lib().orc_max_threads())
