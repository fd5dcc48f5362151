"""Model description consumed by the C ABI (include/mcmcdate_b200.h: mcd_model_desc).

One `ModelDesc` gathers what the reference closes over when it builds its prior/likelihood
functions in `getMcmcProps` (app/Main.hs:370-457): the mean tree's topology, the likelihood data
(`LikelihoodData`, app/Probability.hs:210-235), the relaxed-clock model, the mean root height `ht`
(app/Main.hs:394) and the calibration / constraint / brace tables.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from . import tree as _tree

# RelaxedMolecularClockModel (app/Probability.hs:88-93)
UNCORRELATED_GAMMA, UNCORRELATED_LOGNORMAL, UNCORRELATED_WHITENOISE, AUTOCORRELATED_LOGNORMAL = 0, 1, 2, 3
CLOCK_NAMES = {"ug": 0, "ul": 1, "uw": 2, "al": 3}
# LikelihoodData constructors (app/Probability.hs:210-235)
LIK_FULL, LIK_UNIVARIATE, LIK_NONE, LIK_SPARSE = 0, 1, 2, 3

# per-chain status bits returned by the evaluator (include/mcmcdate_b200.h)
ST_REF_ERROR, ST_ZERO, ST_NAN, ST_NEARCRIT, ST_LEAF_HEIGHT, ST_FP64_FALLBACK = 1, 2, 4, 8, 16, 32

# columns of the per-chain output row
OUT_LNA, OUT_LNB, OUT_LNC, OUT_LNPRIOR, OUT_LNLIK, OUT_LNJAC, OUT_LNPOST = range(7)
OUT_COLS = 8  # 7 values + 1 pad (rows are 64-byte aligned)


def _f64(x):
    return np.ascontiguousarray(x, dtype=np.float64)


def _i32(x):
    return np.ascontiguousarray(x, dtype=np.int32)


@dataclass
class ModelDesc:
    parent: np.ndarray                      # [N] int32, pre-order, root -1
    mean: np.ndarray                        # [K]
    precision: np.ndarray                   # [K,K] (LIK_FULL) or [K] variances (LIK_UNIVARIATE)
    logdet_sigma: float
    clock_model: int = UNCORRELATED_LOGNORMAL
    likelihood: int = LIK_FULL
    ht: float = 1.0
    cal_node: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int32))
    cal_lo: np.ndarray = field(default_factory=lambda: np.zeros(0))      # <= 0: no lower bound
    cal_lo_p: np.ndarray = field(default_factory=lambda: np.zeros(0))
    cal_hi: np.ndarray = field(default_factory=lambda: np.zeros(0))      # +inf: no upper bound
    cal_hi_p: np.ndarray = field(default_factory=lambda: np.zeros(0))
    con_young: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int32))
    con_old: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int32))
    con_p: np.ndarray = field(default_factory=lambda: np.zeros(0))
    brace_off: np.ndarray = field(default_factory=lambda: np.zeros(1, np.int32))
    brace_node: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int32))
    brace_sd: np.ndarray = field(default_factory=lambda: np.zeros(0))
    # LIK_SPARSE: association list ((i, j), v) of the sparse precision (SparseS, app/Main.hs:75-81)
    sparse_row: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int32))
    sparse_col: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int32))
    sparse_val: np.ndarray = field(default_factory=lambda: np.zeros(0))

    def __post_init__(self):
        self.parent = _i32(self.parent)
        self.mean = _f64(self.mean)
        self.precision = _f64(self.precision)
        for k in ("cal_node", "con_young", "con_old", "brace_off", "brace_node", "sparse_row", "sparse_col"):
            setattr(self, k, _i32(getattr(self, k)))
        for k in ("cal_lo", "cal_lo_p", "cal_hi", "cal_hi_p", "con_p", "brace_sd", "sparse_val"):
            setattr(self, k, _f64(getattr(self, k)))
        self.child0, self.child1 = _tree.children_from_parent(self.parent)

    # sizes -----------------------------------------------------------------------------------
    @property
    def n_nodes(self) -> int:
        return len(self.parent)

    @property
    def n_leaves(self) -> int:
        return (self.n_nodes + 1) // 2

    @property
    def dim(self) -> int:  # K
        return self.n_nodes - 2

    @property
    def state_len(self) -> int:  # S = 5 + 2N, canonical order (app/State.hs:70-100)
        return 5 + 2 * self.n_nodes

    @property
    def n_cal(self) -> int:
        return len(self.cal_node)

    @property
    def n_con(self) -> int:
        return len(self.con_young)

    @property
    def n_brace(self) -> int:
        return len(self.brace_sd)

    @property
    def calibrations_available(self) -> bool:
        return self.n_cal > 0


# canonical state layout helpers: [lambda, mu, H, h[N], m, v, r[N]]
def state_slices(n_nodes: int):
    N = n_nodes
    return {
        "lambda": 0, "mu": 1, "H": 2, "h": slice(3, 3 + N), "m": 3 + N, "v": 4 + N,
        "r": slice(5 + N, 5 + 2 * N),
    }
