"""Shared helpers for the tests."""
import os

import numpy as np

from mcmc_date_b200 import model

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FIXTURES = ["06-leaves-constant-rate", "12-leaves-variable-rate", "24-leaves-braces", "mtcdnapri-7-leaves",
            "06-leaves-pinned-node", "10-leaves-autocorrelated-rate", "25-leaves-bastien"]
TOL = 1e-10  # north_star: 1e-10 relative error in FP64


def load_fixture(name, clock=1, likelihood=None):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    lik = int(z["likelihood"]) if likelihood is None else likelihood
    prec = z["precision"]
    logdet = float(z["logdet_sigma"])
    if lik == model.LIK_UNIVARIATE:
        var = 1.0 / np.diag(prec)
        prec, logdet = var, float(np.sum(np.log(var)))
    md = model.ModelDesc(
        parent=z["parent"], mean=z["mean"], precision=prec, logdet_sigma=logdet, clock_model=clock,
        likelihood=lik, ht=float(z["ht"]), cal_node=z["cal_node"], cal_lo=z["cal_lo"], cal_lo_p=z["cal_lo_p"],
        cal_hi=z["cal_hi"], cal_hi_p=z["cal_hi_p"], con_young=z["con_young"], con_old=z["con_old"], con_p=z["con_p"],
        brace_off=z["brace_off"], brace_node=z["brace_node"], brace_sd=z["brace_sd"])
    return md, z


def relerr(a, b):
    """|a-b| / max(1,|b|), with matching infinities / NaNs counting as exact"""
    a, b = np.asarray(a, float), np.asarray(b, float)
    same = (a == b) | (np.isnan(a) & np.isnan(b))
    with np.errstate(invalid="ignore"):
        e = np.abs(a - b) / np.maximum(1.0, np.abs(b))
    return np.where(same, 0.0, np.where(np.isfinite(e), e, np.inf))


def grad_relerr(g, ref):
    """gradient error relative to max(1, largest reference component of that chain)"""
    sc = np.maximum(1.0, np.nanmax(np.abs(ref), axis=-1, keepdims=True))
    return np.abs(g - ref) / sc


def grad_relerr_scalar_block(g, ref, n_nodes):
    """the scalar block (lambda, mu, H, m, v) of the gradient on ITS OWN scale: |g - ref| / max(1, largest scalar-block reference
    component of that chain).  The rate gradients are orders of magnitude larger, so the chain-wide normwise measure of
    grad_relerr says little about these five entries."""
    idx = np.array([0, 1, 2, 3 + n_nodes, 4 + n_nodes])
    gs, rs = np.asarray(g)[..., idx], np.asarray(ref)[..., idx]
    sc = np.maximum(1.0, np.nanmax(np.abs(rs), axis=-1, keepdims=True))
    return np.abs(gs - rs) / sc


def grad_relerr_componentwise(g, ref, floor):
    """|g - ref| / max(floor, |ref|) for every component"""
    return np.abs(np.asarray(g) - np.asarray(ref)) / np.maximum(floor, np.abs(ref))
