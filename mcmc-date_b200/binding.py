"""ctypes binding of the C ABI (include/mcmcdate_b200.h) -- the same entry points a Haskell
`foreign import ccall` shim would bind.

No CPU fallback: if libmcmcdate_b200.so is missing or no CUDA device is present, construction of an
`Evaluator` raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import model as _m

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MCD_LIB_PATH") or os.path.join(_HERE, "libmcmcdate_b200.so")  # override: experiments only
_LIB = None

EXPORTS = [
    "mcd_create", "mcd_destroy", "mcd_last_error", "mcd_state_len", "mcd_dim", "mcd_branch_index", "mcd_mask",
    "mcd_hmc_dim", "mcd_to_vector", "mcd_from_vector", "mcd_eval", "mcd_eval_grad", "mcd_eval_grad_theta", "mcd_eval_grad_theta_async", "mcd_wait", "mcd_eval_async", "mcd_eval_grad_async", "mcd_leapfrog", "mcd_nuts",
    "mcd_chains_set", "mcd_chains_get", "mcd_chains_nuts", "mcd_mh_step", "mcd_mh_cycle", "mcd_mc3_configure", "mcd_mc3_swap", "mcd_mc3_slots",
    "mcd_chains_out_device", "mcd_chains_stats_device", "mcd_mh_set_incremental", "mcd_mh_get_incremental",
    "mcd_eval_device", "mcd_set_contraction", "mcd_get_contraction",
    "mcd_eval_grad_device", "mcd_kernel_launches", "mcd_synchronize", "mcd_version", "mcd_set_kernel_timing",
    "mcd_kernel_times", "mcd_comm_unique_id", "mcd_comm_init", "mcd_allgather_stats", "mcd_comm_destroy",
]


class MhProposalC(C.Structure):
    """struct mcd_mh_proposal"""
    _fields_ = [("kind", C.c_int32), ("node", C.c_int32), ("param", C.c_double), ("tune", C.c_double),
                ("use_root_jacobian", C.c_int32), ("repeat", C.c_int32)]


# proposal kinds (include/mcmcdate_b200.h)
(MH_SLIDE_NODE, MH_SCALE_SUBTREE, MH_PULLEY, MH_SLIDE_BRACE, MH_SCALE_BRANCH, MH_SCALE_RATE_SUBTREE, MH_SCALE_NORM_TREE_CONTRA_M,
 MH_SCALE_NORM_TREE_CONTRA_H, MH_SCALE_VAR_TREE, MH_SCALE_VAR_TREE_AUTO, MH_SLIDE_NODE_CONTRA, MH_SCALE_SUBTREE_CONTRA,
 MH_SLIDE_BRACE_CONTRA, MH_SLIDE_ROOT_CONTRA, MH_SCALE_RATES_TREE_CONTRA, MH_SCALE_SCALAR, MH_SCALE_H_M_CONTRA) = range(17)


class ModelDescC(C.Structure):
    """struct mcd_model_desc"""
    _fields_ = [
        ("n_nodes", C.c_int32), ("parent", C.POINTER(C.c_int32)),
        ("clock_model", C.c_int32), ("likelihood", C.c_int32),
        ("mean", C.POINTER(C.c_double)), ("precision", C.POINTER(C.c_double)),
        ("logdet_sigma", C.c_double), ("ht", C.c_double),
        ("n_cal", C.c_int32), ("cal_node", C.POINTER(C.c_int32)),
        ("cal_lo", C.POINTER(C.c_double)), ("cal_lo_p", C.POINTER(C.c_double)),
        ("cal_hi", C.POINTER(C.c_double)), ("cal_hi_p", C.POINTER(C.c_double)),
        ("n_con", C.c_int32), ("con_young", C.POINTER(C.c_int32)), ("con_old", C.POINTER(C.c_int32)),
        ("con_p", C.POINTER(C.c_double)),
        ("n_brace", C.c_int32), ("brace_off", C.POINTER(C.c_int32)), ("brace_node", C.POINTER(C.c_int32)),
        ("brace_sd", C.POINTER(C.c_double)),
        ("device", C.c_int32), ("max_batch", C.c_int32),
        ("precision_chol", C.POINTER(C.c_double)),
        ("n_sparse", C.c_int32), ("sparse_row", C.POINTER(C.c_int32)), ("sparse_col", C.POINTER(C.c_int32)),
        ("sparse_val", C.POINTER(C.c_double)),
    ]


def load_library():
    """dlopen the in-tree shared library and declare prototypes.  Raises if it is not built."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} not built -- run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    vp, i32, dp, ip, u8p = C.c_void_p, C.c_int32, C.POINTER(C.c_double), C.POINTER(C.c_int32), C.POINTER(C.c_uint8)
    L.mcd_create.argtypes = [C.POINTER(ModelDescC), C.POINTER(vp)]
    L.mcd_destroy.argtypes = [vp]
    L.mcd_destroy.restype = None
    L.mcd_last_error.argtypes = [vp]
    L.mcd_last_error.restype = C.c_char_p
    L.mcd_version.restype = C.c_char_p
    for f in ("mcd_state_len", "mcd_dim", "mcd_hmc_dim", "mcd_synchronize"):
        getattr(L, f).argtypes = [vp]
    L.mcd_branch_index.argtypes = [vp, ip]
    L.mcd_mask.argtypes = [vp, u8p]
    L.mcd_to_vector.argtypes = [vp, dp, dp]
    L.mcd_from_vector.argtypes = [vp, dp, dp, dp]
    L.mcd_eval.argtypes = [vp, i32, dp, dp, ip]
    L.mcd_eval_grad.argtypes = [vp, i32, dp, dp, dp, ip]
    L.mcd_eval_grad_theta.argtypes = [vp, i32, dp, dp, dp, dp, ip]
    L.mcd_eval_grad_theta_async.argtypes = [vp, i32, dp, dp, dp, dp, ip]
    L.mcd_eval_grad_theta_async.restype = C.c_int64
    L.mcd_wait.argtypes = [vp, C.c_int64]
    L.mcd_eval_async.argtypes = [vp, i32, dp, dp, ip]
    L.mcd_eval_async.restype = C.c_int64
    L.mcd_eval_grad_async.argtypes = [vp, i32, dp, dp, dp, ip]
    L.mcd_eval_grad_async.restype = C.c_int64
    L.mcd_leapfrog.argtypes = [vp, i32, i32, dp, dp, dp, dp, dp, dp, dp, dp, dp, ip]
    L.mcd_nuts.argtypes = [vp, i32, dp, dp, dp, dp, dp, i32, C.c_uint64, C.c_uint32, dp, dp, dp, ip, ip]
    L.mcd_chains_set.argtypes = [vp, i32, dp]
    L.mcd_chains_get.argtypes = [vp, i32, dp, dp, ip]
    L.mcd_chains_nuts.argtypes = [vp, dp, dp, i32, C.c_uint64, C.c_uint32, dp, ip, ip]
    L.mcd_mh_step.argtypes = [vp, i32, i32, C.c_double, C.c_double, i32, C.c_uint64, C.c_uint32, ip]
    u64p = C.POINTER(C.c_uint64)
    L.mcd_mh_cycle.argtypes = [vp, i32, C.POINTER(MhProposalC), i32, C.c_uint64, C.c_uint32, u64p, u64p, C.POINTER(C.c_uint32)]
    L.mcd_mc3_configure.argtypes = [vp, i32, i32, i32, dp, dp]
    L.mcd_mc3_swap.argtypes = [vp, i32, C.c_uint64, C.c_uint32, vp, ip]
    L.mcd_mc3_slots.argtypes = [vp, ip]
    L.mcd_chains_out_device.argtypes = [vp]
    L.mcd_chains_out_device.restype = vp
    L.mcd_chains_stats_device.argtypes = [vp, vp]
    L.mcd_mh_set_incremental.argtypes = [vp, i32, i32]
    L.mcd_mh_get_incremental.argtypes = [vp]
    L.mcd_eval_device.argtypes = [vp, i32, vp, vp, vp, vp]
    L.mcd_eval_grad_device.argtypes = [vp, i32, vp, vp, vp, vp, vp]
    L.mcd_set_contraction.argtypes = [vp, i32]
    L.mcd_get_contraction.argtypes = [vp]
    L.mcd_kernel_launches.argtypes = [vp]
    L.mcd_kernel_launches.restype = C.c_int64
    L.mcd_set_kernel_timing.argtypes = [vp, C.c_int]
    L.mcd_kernel_times.argtypes = [vp, dp, C.POINTER(C.c_int64)]
    L.mcd_comm_unique_id.argtypes = [vp]
    L.mcd_comm_init.argtypes = [vp, i32, i32, C.c_char_p]
    L.mcd_allgather_stats.argtypes = [vp, vp]
    L.mcd_comm_destroy.argtypes = [vp]
    _LIB = L
    return L


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


class Evaluator:
    """One model on one GPU (an `mcd_handle`)."""

    def __init__(self, md: _m.ModelDesc, device: int = 0, max_batch: int = 0, supply_cholesky: bool = True):
        L = load_library()
        self._L = L
        self.md = md
        prec = np.ascontiguousarray(md.precision, dtype=np.float64).reshape(-1)
        self._keep = [prec]
        d = ModelDescC()
        d.n_nodes = md.n_nodes
        d.parent = _ip(md.parent)
        d.clock_model, d.likelihood = md.clock_model, md.likelihood
        d.mean, d.precision = _dp(md.mean), _dp(prec)
        d.logdet_sigma, d.ht = md.logdet_sigma, md.ht
        d.n_cal, d.cal_node = md.n_cal, _ip(md.cal_node)
        d.cal_lo, d.cal_lo_p, d.cal_hi, d.cal_hi_p = _dp(md.cal_lo), _dp(md.cal_lo_p), _dp(md.cal_hi), _dp(md.cal_hi_p)
        d.n_con, d.con_young, d.con_old, d.con_p = md.n_con, _ip(md.con_young), _ip(md.con_old), _dp(md.con_p)
        d.n_brace, d.brace_off, d.brace_node, d.brace_sd = md.n_brace, _ip(md.brace_off), _ip(md.brace_node), _dp(md.brace_sd)
        d.device, d.max_batch = device, max_batch
        d.precision_chol = None
        d.n_sparse = len(md.sparse_val)
        d.sparse_row, d.sparse_col, d.sparse_val = _ip(md.sparse_row), _ip(md.sparse_col), _dp(md.sparse_val)
        if md.likelihood == _m.LIK_FULL and supply_cholesky:
            try:  # numpy's LAPACK Cholesky is much faster than the library's fallback loop
                chol = np.ascontiguousarray(np.linalg.cholesky(md.precision))
                self._keep.append(chol)
                d.precision_chol = _dp(chol)
            except np.linalg.LinAlgError:
                pass
        h = C.c_void_p()
        rc = L.mcd_create(C.byref(d), C.byref(h))
        if rc != 0:
            raise RuntimeError("mcd_create failed: " + L.mcd_last_error(None).decode())
        self.h = h
        self.S = L.mcd_state_len(h)
        self.K = L.mcd_dim(h)
        self.D = L.mcd_hmc_dim(h)
        self.device = device

    def close(self):
        if getattr(self, "h", None):
            self._L.mcd_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise RuntimeError("mcmcdate_b200: " + self._L.mcd_last_error(self.h).decode())

    # model queries -------------------------------------------------------------------------
    def branch_index(self) -> np.ndarray:
        out = np.empty(self.md.n_nodes, np.int32)
        self._check(self._L.mcd_branch_index(self.h, _ip(out)))
        return out

    def mask(self) -> np.ndarray:
        out = np.empty(self.S, np.uint8)
        self._check(self._L.mcd_mask(self.h, out.ctypes.data_as(C.POINTER(C.c_uint8))))
        return out

    def to_vector(self, state: np.ndarray) -> np.ndarray:
        x = np.ascontiguousarray(state, dtype=np.float64)
        th = np.empty(self.D)
        self._check(self._L.mcd_to_vector(self.h, _dp(x), _dp(th)))
        return th

    def from_vector(self, base_state: np.ndarray, theta: np.ndarray) -> np.ndarray:
        x = np.ascontiguousarray(base_state, dtype=np.float64)
        th = np.ascontiguousarray(theta, dtype=np.float64)
        out = np.empty(self.S)
        self._check(self._L.mcd_from_vector(self.h, _dp(x), _dp(th), _dp(out)))
        return out

    # host-buffer evaluation ----------------------------------------------------------------
    def eval(self, states: np.ndarray, out: np.ndarray | None = None, status: np.ndarray | None = None):
        X = np.ascontiguousarray(states, dtype=np.float64).reshape(-1, self.S)
        B = X.shape[0]
        out = np.empty((B, _m.OUT_COLS)) if out is None else out
        status = np.empty(B, np.int32) if status is None else status
        self._check(self._L.mcd_eval(self.h, B, _dp(X), _dp(out), _ip(status)))
        return out, status

    def eval_grad(self, states: np.ndarray, out: np.ndarray | None = None, grad: np.ndarray | None = None,
                  status: np.ndarray | None = None):
        X = np.ascontiguousarray(states, dtype=np.float64).reshape(-1, self.S)
        B = X.shape[0]
        out = np.empty((B, _m.OUT_COLS)) if out is None else out
        grad = np.empty((B, self.S)) if grad is None else grad
        status = np.empty(B, np.int32) if status is None else status
        self._check(self._L.mcd_eval_grad(self.h, B, _dp(X), _dp(out), _dp(grad), _ip(status)))
        return out, grad, status

    def eval_grad_theta(self, theta: np.ndarray, base_state: np.ndarray, out=None, grad_theta=None, status=None):
        """HMC form: theta [B, D] (masked, reversed order) + one shared base state -> out, d/dtheta, status"""
        T = np.ascontiguousarray(theta, dtype=np.float64).reshape(-1, self.D)
        base = np.ascontiguousarray(base_state, dtype=np.float64)
        B = T.shape[0]
        out = np.empty((B, _m.OUT_COLS)) if out is None else out
        grad_theta = np.empty((B, self.D)) if grad_theta is None else grad_theta
        status = np.empty(B, np.int32) if status is None else status
        self._check(self._L.mcd_eval_grad_theta(self.h, B, _dp(T), _dp(base), _dp(out), _dp(grad_theta), _ip(status)))
        return out, grad_theta, status

    def leapfrog(self, theta0, momentum0, base_state, inv_mass, step_size, n_steps: int):
        """n_steps leapfrog steps, resident on the device -> (theta, momentum, out, energy[B,2], status)"""
        T = np.ascontiguousarray(theta0, dtype=np.float64).reshape(-1, self.D)
        P = np.ascontiguousarray(momentum0, dtype=np.float64).reshape(-1, self.D)
        B = T.shape[0]
        base = np.ascontiguousarray(base_state, dtype=np.float64)
        im = np.ascontiguousarray(inv_mass, dtype=np.float64)
        eps = np.ascontiguousarray(np.broadcast_to(np.asarray(step_size, dtype=np.float64), (B,)))
        th, pm = np.empty_like(T), np.empty_like(P)
        out, en, st = np.empty((B, _m.OUT_COLS)), np.empty((B, 2)), np.empty(B, np.int32)
        self._check(self._L.mcd_leapfrog(self.h, B, int(n_steps), _dp(T), _dp(P), _dp(base), _dp(im), _dp(eps), _dp(th),
                                         _dp(pm), _dp(out), _dp(en), _ip(st)))
        return th, pm, out, en, st

    def nuts(self, theta0, base_state, inv_mass, step_size, max_depth: int = 10, seed: int = 0, iteration: int = 0,
             momentum0=None):
        """one NUTS transition per chain on the device -> (theta, out, accept_stat, info[B,4], status);
        info = (depth, leapfrog steps, diverged, valid points)"""
        T = np.ascontiguousarray(theta0, dtype=np.float64).reshape(-1, self.D)
        B = T.shape[0]
        base = np.ascontiguousarray(base_state, dtype=np.float64)
        im = np.ascontiguousarray(inv_mass, dtype=np.float64)
        eps = np.ascontiguousarray(np.broadcast_to(np.asarray(step_size, dtype=np.float64), (B,)))
        mom = None if momentum0 is None else np.ascontiguousarray(momentum0, dtype=np.float64).reshape(B, self.D)
        th = np.empty_like(T)
        out, acc = np.empty((B, _m.OUT_COLS)), np.empty(B)
        info, st = np.empty((B, 4), np.int32), np.empty(B, np.int32)
        self._check(self._L.mcd_nuts(self.h, B, _dp(T), _dp(base), _dp(im), _dp(eps), _dp(mom) if mom is not None else None,
                                     int(max_depth), int(seed), int(iteration), _dp(th), _dp(out), _dp(acc), _ip(info), _ip(st)))
        return th, out, acc, info, st

    # ---- chains resident in HBM + Metropolis-Hastings moves
    def chains_set(self, states):
        X = np.ascontiguousarray(states, dtype=np.float64).reshape(-1, self.S)
        self._n_resident = X.shape[0]
        self._check(self._L.mcd_chains_set(self.h, X.shape[0], _dp(X)))

    def n_resident(self) -> int:
        return int(getattr(self, "_n_resident", 0))

    def chains_get(self):
        B = self._n_resident
        X, out, st = np.empty((B, self.S)), np.empty((B, _m.OUT_COLS)), np.empty(B, np.int32)
        self._check(self._L.mcd_chains_get(self.h, B, _dp(X), _dp(out), _ip(st)))
        return X, out, st

    def chains_nuts(self, inv_mass, step_size, max_depth: int = 10, seed: int = 0, iteration: int = 0):
        """one NUTS transition of every resident chain -> (accept_stat[B], info[B, 4], status[B])"""
        B = self._n_resident
        im = np.ascontiguousarray(inv_mass, dtype=np.float64)
        eps = np.ascontiguousarray(np.broadcast_to(np.asarray(step_size, dtype=np.float64), (B,)))
        acc, info, st = np.empty(B), np.empty((B, 4), np.int32), np.empty(B, np.int32)
        self._check(self._L.mcd_chains_nuts(self.h, _dp(im), _dp(eps), int(max_depth), int(seed), int(iteration), _dp(acc),
                                            _ip(info), _ip(st)))
        return acc, info, st

    def mh_step(self, kind: int, node: int, sd: float, tune: float = 1.0, use_root_jacobian: bool = False, seed: int = 0,
                iteration: int = 0, want_accepted: bool = True):
        """one proposal (MH_* kind; sd = standard deviation or gamma shape) on every resident chain -> accepted flags (1 / 0 / -1)"""
        acc = np.empty(self._n_resident, np.int32) if want_accepted else None
        self._check(self._L.mcd_mh_step(self.h, int(kind), int(node), float(sd), float(tune), int(use_root_jacobian), int(seed),
                                        int(iteration), _ip(acc) if acc is not None else None))
        return acc

    def mh_cycle(self, proposals, n_iterations: int = 1, seed: int = 0, iteration0: int = 0):
        """sweeps over a list of (kind, node, param, tune, use_root_jacobian, repeat) tuples, enqueued back to back
        -> (accepted[n_props], invalid[n_props], next unused iteration value)"""
        if len(proposals) > 4096:      # the C entry point takes 4096 list entries per call
            acc, inv, k = [], [], int(iteration0)
            assert n_iterations == 1, "long cycles: one sweep per call"
            for c0 in range(0, len(proposals), 4096):
                a, i, k = self.mh_cycle(proposals[c0:c0 + 4096], 1, seed, k)
                acc.append(a)
                inv.append(i)
            return np.concatenate(acc), np.concatenate(inv), k
        n = len(proposals)
        arr = (MhProposalC * n)()
        for i, (kind, node, param, tune, jac, rep) in enumerate(proposals):
            arr[i] = MhProposalC(int(kind), int(node), float(param), float(tune), int(jac), int(rep))
        acc, inv = np.zeros(n, np.uint64), np.zeros(n, np.uint64)
        nxt = C.c_uint32(0)
        u64p = C.POINTER(C.c_uint64)
        self._check(self._L.mcd_mh_cycle(self.h, n, arr, int(n_iterations), int(seed), int(iteration0), acc.ctypes.data_as(u64p),
                                         inv.ctypes.data_as(u64p), C.byref(nxt)))
        return acc, inv, int(nxt.value)

    def mc3_configure(self, n_global: int, chain_offset: int, chains_per_group: int, ladder_prior=None, ladder_lik=None):
        """heated chains: groups of chains_per_group temperatures (0: cold chains again)"""
        lp = np.ascontiguousarray(ladder_prior if ladder_prior is not None else [1.0], dtype=np.float64)
        ll = np.ascontiguousarray(ladder_lik if ladder_lik is not None else [1.0], dtype=np.float64)
        self._mc3 = (int(n_global), int(chains_per_group))
        self._check(self._L.mcd_mc3_configure(self.h, int(n_global), int(chain_offset), int(chains_per_group), _dp(lp), _dp(ll)))

    def mc3_swap(self, pair: int = -1, seed: int = 0, iteration: int = 0, d_stats_global: int = 0, want_accepted: bool = True):
        """one swap attempt per group between the slots (pair, pair + 1) -> accepted flag per group"""
        n_global, Cg = getattr(self, "_mc3", (0, 0))
        acc = np.empty(max(1, n_global // max(1, Cg)), np.int32) if want_accepted else None
        self._check(self._L.mcd_mc3_swap(self.h, int(pair), int(seed), int(iteration), d_stats_global or None,
                                         _ip(acc) if acc is not None else None))
        return acc

    def mc3_slots(self):
        sl = np.empty(max(1, getattr(self, "_mc3", (0, 0))[0]), np.int32)
        self._check(self._L.mcd_mc3_slots(self.h, _ip(sl)))
        return sl

    def mh_set_incremental(self, on: bool, refresh_every: int = 0):
        """incremental evaluation of small moves (call before chains_set)"""
        self._check(self._L.mcd_mh_set_incremental(self.h, int(on), int(refresh_every)))

    def mh_incremental_active(self) -> bool:
        return self._L.mcd_mh_get_incremental(self.h) == 1

    # ---- MC3 swap statistics across GPUs (NCCL inside the library)
    @staticmethod
    def comm_unique_id() -> bytes:
        """rank 0: the 128-byte NCCL id the other ranks need for comm_init"""
        L = load_library()
        buf = C.create_string_buffer(128)
        if L.mcd_comm_unique_id(buf) != 0:
            raise RuntimeError("mcd_comm_unique_id: " + (L.mcd_last_error(None) or b"").decode())
        return buf.raw

    def comm_init(self, world: int, rank: int, unique_id: bytes):
        self._check(self._L.mcd_comm_init(self.h, int(world), int(rank), C.c_char_p(bytes(unique_id))))

    def allgather_stats(self, d_stats_global: int):
        """(ln prior, ln lik) of the resident chains of every rank -> device buffer [world * n_resident][2]"""
        self._check(self._L.mcd_allgather_stats(self.h, d_stats_global))

    def comm_destroy(self):
        self._check(self._L.mcd_comm_destroy(self.h))

    def chains_stats_device(self, d_stats: int):
        """(ln prior, ln lik) of the resident chains -> device buffer [n][2] (send buffer of the MC3 all-gather)"""
        self._check(self._L.mcd_chains_stats_device(self.h, d_stats))

    def chains_out_device(self) -> int:
        return int(self._L.mcd_chains_out_device(self.h) or 0)

    def nuts_ptr(self, B: int, theta0: int, base: int, inv_mass: int, eps: int, momentum0: int, max_depth: int, seed: int,
                 iteration: int, theta_out: int, out: int, accept_stat: int, info: int, status: int):
        """mcd_nuts on caller-owned (e.g. pinned) host buffers given as raw addresses; momentum0 = 0: drawn on the device"""
        dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int32)
        c = lambda a: C.cast(a, dp)
        self._check(self._L.mcd_nuts(self.h, B, c(theta0), c(base), c(inv_mass), c(eps), c(momentum0) if momentum0 else None,
                                     int(max_depth), int(seed), int(iteration), c(theta_out), c(out), c(accept_stat),
                                     C.cast(info, ip), C.cast(status, ip)))

    def leapfrog_ptr(self, B: int, n_steps: int, theta0: int, mom0: int, base: int, inv_mass: int, eps: int, theta_out: int,
                     mom_out: int, out: int, energy: int, status: int):
        """mcd_leapfrog on caller-owned (e.g. pinned) host buffers given as raw addresses"""
        dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int32)
        c = lambda a: C.cast(a, dp)
        self._check(self._L.mcd_leapfrog(self.h, B, int(n_steps), c(theta0), c(mom0), c(base), c(inv_mass), c(eps), c(theta_out),
                                         c(mom_out), c(out), c(energy), C.cast(status, ip)))

    def eval_grad_theta_ptr(self, B: int, theta_ptr: int, base_ptr: int, out_ptr: int, gtheta_ptr: int, status_ptr: int):
        dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int32)
        self._check(self._L.mcd_eval_grad_theta(self.h, B, C.cast(theta_ptr, dp), C.cast(base_ptr, dp), C.cast(out_ptr, dp),
                                                C.cast(gtheta_ptr, dp), C.cast(status_ptr, ip)))

    # raw-pointer entry points (host or device addresses as ints) ---------------------------
    def eval_grad_theta_async_ptr(self, B: int, theta_ptr: int, base_ptr: int, out_ptr: int, gtheta_ptr: int, status_ptr: int) -> int:
        """asynchronous form (pinned host buffers): returns a ticket; results are valid after wait(ticket)"""
        dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int32)
        t = self._L.mcd_eval_grad_theta_async(self.h, B, C.cast(theta_ptr, dp), C.cast(base_ptr, dp), C.cast(out_ptr, dp),
                                              C.cast(gtheta_ptr, dp), C.cast(status_ptr, ip))
        if t < 0:
            self._check(-1)
        return int(t)

    def eval_grad_async_ptr(self, B: int, states_ptr: int, out_ptr: int, grad_ptr: int, status_ptr: int) -> int:
        dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int32)
        t = self._L.mcd_eval_grad_async(self.h, B, C.cast(states_ptr, dp), C.cast(out_ptr, dp), C.cast(grad_ptr, dp), C.cast(status_ptr, ip))
        if t < 0:
            self._check(-1)
        return int(t)

    def eval_async_ptr(self, B: int, states_ptr: int, out_ptr: int, status_ptr: int) -> int:
        dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int32)
        t = self._L.mcd_eval_async(self.h, B, C.cast(states_ptr, dp), C.cast(out_ptr, dp), C.cast(status_ptr, ip))
        if t < 0:
            self._check(-1)
        return int(t)

    def wait(self, ticket: int):
        self._check(self._L.mcd_wait(self.h, int(ticket)))

    def eval_grad_ptr(self, B: int, states_ptr: int, out_ptr: int, grad_ptr: int, status_ptr: int):
        dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int32)
        self._check(self._L.mcd_eval_grad(self.h, B, C.cast(states_ptr, dp), C.cast(out_ptr, dp), C.cast(grad_ptr, dp),
                                          C.cast(status_ptr, ip)))

    def eval_ptr(self, B: int, states_ptr: int, out_ptr: int, status_ptr: int):
        dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int32)
        self._check(self._L.mcd_eval(self.h, B, C.cast(states_ptr, dp), C.cast(out_ptr, dp), C.cast(status_ptr, ip)))

    def eval_grad_device(self, B: int, d_states: int, d_out: int, d_grad: int, d_status: int, stream: int = 0):
        self._check(self._L.mcd_eval_grad_device(self.h, B, d_states, d_out, d_grad, d_status, stream))

    def eval_device(self, B: int, d_states: int, d_out: int, d_status: int, stream: int = 0):
        self._check(self._L.mcd_eval_device(self.h, B, d_states, d_out, d_status, stream))

    def kernel_launches(self) -> int:
        return int(self._L.mcd_kernel_launches(self.h))

    def set_contraction(self, mode):
        """'dmma' (FP64 tensor instructions) or 'i8s6' / 'i8s7' (INT8 tensor cores, n base-256 digit planes)"""
        modes = {"dmma": 0, "i8s6": 6, "i8s7": 7}
        self._check(self._L.mcd_set_contraction(self.h, modes[mode] if isinstance(mode, str) else int(mode)))

    def get_contraction(self) -> int:
        return int(self._L.mcd_get_contraction(self.h))

    def set_kernel_timing(self, on: bool):
        self._check(self._L.mcd_set_kernel_timing(self.h, int(on)))

    def kernel_times(self):
        """-> (ms[3] summed over calls: residual, contraction, posterior; number of calls)"""
        ms = np.zeros(3)
        n = C.c_int64()
        self._check(self._L.mcd_kernel_times(self.h, _dp(ms), C.byref(n)))
        return ms, int(n.value)

    def synchronize(self):
        self._check(self._L.mcd_synchronize(self.h))
