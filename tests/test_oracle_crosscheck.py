"""Oracle (C++) vs the independent mpmath restatement (oracle/mp_oracle.py) and vs its own golden
outputs; analytic CPU-port gradient vs forward-mode duals and vs 50-digit finite differences."""
import mpmath as mp
import numpy as np
import pytest

from mcmc_date_b200 import model, synth
from oracle import mp_oracle as MP
from oracle import oracle as O
from util import FIXTURES, TOL, grad_relerr, load_fixture, relerr


@pytest.mark.parametrize("name", FIXTURES)
@pytest.mark.parametrize("clock", [0, 1, 2, 3])
def test_oracle_matches_mpmath(name, clock):
    md, z = load_fixture(name, clock)
    orc = O.Oracle(md)
    X = z["states"]
    out, st = orc.eval(X[1:4])
    for b in range(3):
        p = MP.ln_posterior_parts(md, X[1 + b])
        ref = [p["A"], p["B"], p["C"], p["prior"], p["lik"], p["jac"], p["post"]]
        for j in range(7):
            err = abs((mp.mpf(out[b, j]) - ref[j]) / max(1, abs(ref[j])))
            assert err < 1e-12, (name, clock, b, j, float(err))
    assert (st == 0).all()


@pytest.mark.parametrize("lik", [model.LIK_UNIVARIATE, model.LIK_NONE])
def test_oracle_matches_mpmath_other_likelihoods(lik):
    md, z = load_fixture("12-leaves-variable-rate", 1, likelihood=lik)
    orc = O.Oracle(md)
    out, st = orc.eval(z["states"][2:3])
    p = MP.ln_posterior_parts(md, z["states"][2])
    assert abs((mp.mpf(out[0, 4]) - p["lik"]) / max(1, abs(p["lik"]))) < 1e-12
    assert abs((mp.mpf(out[0, 6]) - p["post"]) / max(1, abs(p["post"]))) < 1e-12


@pytest.mark.parametrize("name", FIXTURES)
@pytest.mark.parametrize("clock", [0, 1, 2, 3])
def test_oracle_reproduces_golden(name, clock):
    """the committed golden vectors are what the oracle computes today (guards oracle edits)"""
    md, z = load_fixture(name, clock)
    orc = O.Oracle(md)
    out, grad, st = orc.eval_grad(z["states"])
    assert np.array_equal(st, z[f"status_{clock}"])
    assert relerr(out, z[f"out_{clock}"]).max() < 1e-13
    nv = int(z["n_valid"])
    assert grad_relerr(grad[:nv], z[f"grad_{clock}"][:nv]).max() < 1e-13


@pytest.mark.parametrize("name", FIXTURES)
@pytest.mark.parametrize("clock", [0, 1, 2, 3])
def test_port_gradient_matches_duals(name, clock):
    """hand-derived gradient (what the GPU evaluates) == AD of the restated reference code"""
    md, z = load_fixture(name, clock)
    orc = O.Oracle(md)
    X = z["states"][:4]
    _, grad, _ = orc.eval_grad(X)
    for b in range(4):  # state 0 (initWith) is exactly critical (la == mu == 1)
        gd = orc.grad_dual(X[b])
        assert grad_relerr(grad[b], gd).max() < TOL
        assert grad_relerr(z[f"graddual_{clock}"][b], gd).max() < 1e-13


def test_gradient_matches_high_precision_finite_differences():
    md, z = load_fixture("12-leaves-variable-rate", 2)
    orc = O.Oracle(md)
    x = z["states"][3]
    gd = orc.grad_dual(x)
    free = np.nonzero(orc.mask)[0]
    for j in list(free[:6]) + list(free[-4:]) + [int(free[len(free) // 2])]:
        fd = float(MP.grad_fd(md, x, int(j)))
        assert abs(fd - gd[j]) <= 1e-9 * max(1.0, abs(fd)), (j, fd, gd[j])


def test_sparse_likelihood_oracle():
    """Sparse likelihood: oracle vs mpmath, port gradient vs duals, and sparse == full on the same matrix"""
    md, h = synth.synthetic_model(12, seed=3, likelihood=model.LIK_SPARSE, n_cal=2, clock_model=2)
    X = synth.synthetic_states(md, h, 3)
    orc = O.Oracle(md)
    out, grad, st = orc.eval_grad(X)
    p = MP.ln_posterior_parts(md, X[0])
    assert abs((mp.mpf(out[0, 4]) - p["lik"]) / max(1, abs(p["lik"]))) < 1e-12
    assert grad_relerr(grad[0], orc.grad_dual(X[0])).max() < TOL
    dense = np.zeros((md.dim, md.dim))
    np.add.at(dense, (md.sparse_row, md.sparse_col), md.sparse_val)
    md_full = model.ModelDesc(parent=md.parent, mean=md.mean, precision=dense, logdet_sigma=md.logdet_sigma,
                              clock_model=2, ht=md.ht, cal_node=md.cal_node, cal_lo=md.cal_lo, cal_lo_p=md.cal_lo_p,
                              cal_hi=md.cal_hi, cal_hi_p=md.cal_hi_p)
    of, gf, _ = O.Oracle(md_full).eval_grad(X)
    assert relerr(out, of).max() < 1e-13 and grad_relerr(grad, gf).max() < 1e-12


def test_generic_and_double_likelihood_agree():
    """reduceVMV (generic HMC target) and the BLAS-like Double path are the same number"""
    md, z = load_fixture("24-leaves-braces", 1)
    orc = O.Oracle(md)
    a, _ = orc.eval(z["states"][:8])
    b, _ = orc.eval(z["states"][:8], generic=True)
    assert relerr(a, b).max() < 1e-13


def test_near_critical_value_and_gradient():
    """|la - mu| < 1e-6 switches the reference to first-order formulas (BirthDeath.hs:90-126) that
    differ from the exact density by O(|la - mu|) (~1e-6 absolute).  The port (and the GPU) run the
    literal near-critical recursion and its reverse-mode gradient there, so both still match the
    dual-number truth to 1e-10."""
    md, z = load_fixture("24-leaves-braces", 2)
    orc = O.Oracle(md)
    for d in (3e-7, -9e-7, 0.0):
        x = z["states"][5].copy()
        x[1] = x[0] + d
        out, grad, st = orc.eval_grad(x[None])
        assert st[0] & model.ST_NEARCRIT
        gd = orc.grad_dual(x)
        assert grad_relerr(grad[0], gd).max() < TOL
    # the switch is a (tiny) discontinuity of the reference itself
    x = z["states"][5].copy()
    x[1] = x[0] + 0.99e-6
    a, _ = orc.eval(x[None])
    x[1] = x[0] + 1.01e-6
    b, sb = orc.eval(x[None])
    assert not (sb[0] & model.ST_NEARCRIT)
    assert 1e-8 < abs(a[0, 1] - b[0, 1]) < 1e-4


def test_large_tree_directional_derivative():
    """1000-leaf tree: analytic gradient . u == one dual-number pass along u"""
    md, h = synth.synthetic_model(1000, seed=11, n_cal=16, n_con=8, n_brace=4)
    X = synth.synthetic_states(md, h, 2)
    orc = O.Oracle(md)
    out, grad, st = orc.eval_grad(X)
    rng = np.random.default_rng(3)
    for b in range(2):
        u = rng.normal(size=md.state_len) * orc.mask
        dd, val = orc.dir_derivative(X[b], u)
        assert val == pytest.approx(out[b, 6], rel=1e-12)
        assert abs(dd - grad[b] @ u) <= TOL * max(1.0, abs(dd))
