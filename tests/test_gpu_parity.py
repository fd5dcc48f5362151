"""GPU parity tests proper: the CUDA path, called through the C ABI, against the oracle on the same
seeded inputs, against the committed golden vectors, and -- at BASELINE.json's full size -- through
size-independent properties.  Tolerance: 1e-10 relative (north_star) for floating point, exact for
index maps and status words."""
import ctypes as C

import numpy as np
import pytest

from mcmc_date_b200 import binding, model, synth
from oracle import oracle as O
from util import FIXTURES, TOL, grad_relerr, grad_relerr_scalar_block, load_fixture, relerr

pytestmark = pytest.mark.gpu


def _finite_rows(out):
    return np.isfinite(out[:, 6])


@pytest.mark.parametrize("name", FIXTURES)
@pytest.mark.parametrize("clock", [0, 1, 2, 3])
def test_fixture_parity_value_gradient_status(name, clock):
    """reference datasets (configs[0..3]): valid states, the reference's initial state, edge states"""
    md, z = load_fixture(name, clock)
    X = z["states"]
    ev = binding.Evaluator(md)
    out, grad, st = ev.eval_grad(X)
    # committed golden vectors
    assert np.array_equal(st, z[f"status_{clock}"]), (st, z[f"status_{clock}"])
    assert relerr(out[:, :7], z[f"out_{clock}"]).max() < TOL
    fin = _finite_rows(z[f"out_{clock}"])
    assert fin.sum() >= int(z["n_valid"])
    assert grad_relerr(grad[fin], z[f"grad_{clock}"][fin]).max() < TOL
    # dual-number gradient truth of the first four states
    assert grad_relerr(grad[:4], z[f"graddual_{clock}"]).max() < TOL
    # live oracle
    orc = O.Oracle(md)
    oo, og, ost = orc.eval_grad(X)
    assert np.array_equal(st, ost)
    assert relerr(out[:, :7], oo).max() < TOL
    assert grad_relerr(grad[fin], og[fin]).max() < TOL
    assert grad_relerr_scalar_block(grad[fin], og[fin], md.n_nodes).max() < TOL    # (lambda, mu, H, m, v) on their own scale
    # masked entries are exactly zero
    assert (grad[:, orc.mask == 0] == 0).all()
    # value-only entry point returns the same numbers
    out2, st2 = ev.eval(X)
    # (separate template instantiations: FMA contraction may differ in the last bits)
    # and the value-only path evaluates the quadratic form through the Cholesky factor, |L^T dx|^2)
    assert relerr(out2[:, :7], out[:, :7]).max() < TOL and np.array_equal(st2, st)
    assert relerr(out2[:, :7], z[f"out_{clock}"]).max() < TOL
    ev.close()


@pytest.mark.parametrize("name", FIXTURES)
def test_index_maps_bit_exact(name):
    md, z = load_fixture(name)
    ev = binding.Evaluator(md)
    orc = O.Oracle(md)
    assert np.array_equal(ev.branch_index(), z["branch_index"])
    assert np.array_equal(ev.branch_index(), orc.branch_index())
    assert np.array_equal(ev.mask(), z["mask"])
    x = z["states"][2]
    th = ev.to_vector(x)
    assert np.array_equal(th, orc.to_vector(x))
    assert np.array_equal(ev.from_vector(x, th * 2), orc.from_vector(x, th * 2))
    assert ev.D == int(z["mask"].sum())
    ev.close()


@pytest.mark.parametrize("lik", [model.LIK_UNIVARIATE, model.LIK_NONE])
@pytest.mark.parametrize("clock", [1, 2])
def test_other_likelihood_kinds(lik, clock):
    md, z = load_fixture("24-leaves-braces", clock, likelihood=lik)
    X = z["states"][: int(z["n_valid"])]
    ev = binding.Evaluator(md)
    out, grad, st = ev.eval_grad(X)
    oo, og, ost = O.Oracle(md).eval_grad(X)
    assert np.array_equal(st, ost)
    assert relerr(out[:, :7], oo).max() < TOL and grad_relerr(grad, og).max() < TOL
    ev.close()


@pytest.mark.parametrize("n_leaves,B,clock", [(2, 5, 1), (3, 9, 0), (60, 300, 2), (200, 130, 3), (49, 257, 1)])
def test_synthetic_trees_ragged_batches(n_leaves, B, clock):
    """random topologies incl. the smallest trees (K = 1, 3), batches that are not multiples of the tile"""
    md, h = synth.synthetic_model(n_leaves, seed=100 + n_leaves, clock_model=clock, n_cal=min(3, n_leaves - 1),
                                  n_con=2 if n_leaves > 10 else 0, n_brace=1 if n_leaves > 10 else 0)
    X = synth.synthetic_states(md, h, B)
    ev = binding.Evaluator(md)
    out, grad, st = ev.eval_grad(X)
    oo, og, ost = O.Oracle(md).eval_grad(X, nthreads=4)
    assert np.array_equal(st, ost)
    assert relerr(out[:, :7], oo).max() < TOL and grad_relerr(grad, og).max() < TOL
    ev.close()


@pytest.mark.parametrize("n_leaves,clock", [(150, 1), (150, 2), (40, 3), (1000, 1)])
def test_near_critical_birth_death_on_all_kernel_paths(n_leaves, clock):
    """|lambda - mu| < 1e-6: the literal near-critical D/E recursion (value and reverse-mode gradient) in the
    large-tree kernel (CTA per chain) and in the fused small-tree kernel"""
    md, h = synth.synthetic_model(n_leaves, seed=9 + n_leaves, clock_model=clock, n_cal=3, n_con=2, n_brace=1)
    X = synth.synthetic_states(md, h, 40)
    X[1, 1] = X[1, 0] + 2e-7
    X[2, 1] = X[2, 0]
    X[3, 1] = X[3, 0] - 9.9e-7
    X[4, 1] = X[4, 0] + 1.01e-6          # just outside: closed form
    ev = binding.Evaluator(md)
    out, grad, st = ev.eval_grad(X)
    oo, og, ost = O.Oracle(md).eval_grad(X, nthreads=4)
    assert np.array_equal(st, ost) and (st[1:4] == model.ST_NEARCRIT).all() and st[4] == 0
    assert relerr(out[:, :7], oo).max() < TOL and grad_relerr(grad, og).max() < TOL
    o2, s2 = ev.eval(X)
    assert relerr(o2[:, :7], oo).max() < TOL
    ev.close()


def test_thousand_leaf_tree_against_oracle():
    """configs[4] shape at a batch the oracle finishes in seconds"""
    md, h = synth.synthetic_model(1000, seed=synth.BASE_SEED + 4, n_cal=16, n_con=8, n_brace=4)
    X = synth.synthetic_states(md, h, 300)
    ev = binding.Evaluator(md)
    out, grad, st = ev.eval_grad(X)
    orc = O.Oracle(md)
    oo, og, ost = orc.eval_grad(X, nthreads=8)
    assert np.array_equal(st, ost) and (st == 0).all()
    assert relerr(out[:, :7], oo).max() < TOL
    assert grad_relerr(grad, og).max() < TOL
    assert grad_relerr_scalar_block(grad, og, md.n_nodes).max() < TOL             # (lambda, mu, H, m, v) on their own scale
    # gradient truth at this size: one dual-number pass along a random direction
    u = np.random.default_rng(5).normal(size=md.state_len) * orc.mask
    dd, _ = orc.dir_derivative(X[0], u)
    assert abs(grad[0] @ u - dd) <= TOL * max(1.0, abs(dd))
    ev.close()


@pytest.fixture(scope="module")
def full_size():
    """BASELINE.json's full size: 1000-leaf tree, 8192 chains"""
    md, h = synth.synthetic_model(1000, seed=synth.BASE_SEED + 4, n_cal=16, n_con=8, n_brace=4)
    X = synth.synthetic_states(md, h, 8192, seed=synth.BASE_SEED + 5)
    ev = binding.Evaluator(md, max_batch=8192)
    out, grad, st = ev.eval_grad(X)
    yield md, X, ev, out, grad, st
    ev.close()


def test_full_size_deterministic_and_chunk_invariant(full_size):
    md, X, ev, out, grad, st = full_size
    assert np.isfinite(out[:, :7]).all() and (st == 0).all()
    out2, grad2, st2 = ev.eval_grad(X)
    assert np.array_equal(out, out2) and np.array_equal(grad, grad2)            # bit-wise reproducible
    # a chain's result does not depend on which batch / tile / chunk it travels in
    sub = slice(1000, 1300)
    o3, g3, _ = ev.eval_grad(X[sub])
    assert np.array_equal(o3, out[sub]) and np.array_equal(g3, grad[sub])
    perm = np.random.default_rng(0).permutation(len(X))[:2048]
    o4, g4, _ = ev.eval_grad(X[perm])
    assert np.array_equal(o4, out[perm]) and np.array_equal(g4, grad[perm])
    o5, s5 = ev.eval(X)            # value-only: triangular (Cholesky) contraction
    assert relerr(o5[:, :7], out[:, :7]).max() < 1e-11


def test_full_size_spot_check_against_oracle(full_size):
    md, X, ev, out, grad, st = full_size
    idx = np.random.default_rng(1).choice(len(X), size=24, replace=False)
    oo, og, ost = O.Oracle(md).eval_grad(X[idx], nthreads=8)
    assert relerr(out[idx, :7], oo).max() < TOL
    assert grad_relerr(grad[idx], og).max() < TOL


def test_full_size_gradient_is_the_derivative_of_the_value(full_size):
    """central differences of the GPU's own ln posterior along random directions"""
    md, X, ev, out, grad, st = full_size
    mask = ev.mask().astype(float)
    rng = np.random.default_rng(2)
    idx = rng.choice(len(X), size=64, replace=False)
    U = rng.normal(size=(64, md.state_len)) * mask
    U /= np.linalg.norm(U, axis=1, keepdims=True)
    eps = 1e-6
    Xp, Xm = X[idx] + eps * U * np.abs(X[idx]), X[idx] - eps * U * np.abs(X[idx])
    op, _ = ev.eval(Xp)
    om, _ = ev.eval(Xm)
    fd = (op[:, 6] - om[:, 6]) / (2 * eps)
    an = np.einsum("ij,ij->i", grad[idx], U * np.abs(X[idx]))
    assert np.abs(fd - an).max() <= 2e-5 * np.maximum(1.0, np.abs(an)).max()


def test_full_size_quadratic_form_scaling(full_size):
    """lnL is a quadratic form in the distances: scaling m by c scales d by c, so
    lnL(c) - const = -1/2 (c d - mu)^T P (c d - mu) must be an exact parabola in c"""
    md, X, ev, out, grad, st = full_size
    N = md.n_nodes
    Xs = np.repeat(X[:8], 3, axis=0)
    c = np.tile([0.5, 1.0, 1.5], 8)
    Xs[:, 3 + N] *= c
    o, _ = ev.eval(Xs)
    L = o[:, 4].reshape(8, 3)
    second = L[:, 0] - 2 * L[:, 1] + L[:, 2]          # = -(0.5)^2 d^T P d  (constant second difference)
    Xt = np.repeat(X[:8], 3, axis=0)
    c2 = np.tile([1.0, 1.5, 2.0], 8)
    Xt[:, 3 + N] *= c2
    o2, _ = ev.eval(Xt)
    L2 = o2[:, 4].reshape(8, 3)
    second2 = L2[:, 0] - 2 * L2[:, 1] + L2[:, 2]
    assert np.abs(second - second2).max() <= 1e-9 * np.abs(second).max()


def test_device_entry_points_match_host_entry_points(full_size):
    import torch
    md, X, ev, out, grad, st = full_size
    B = 2048
    dev = torch.device("cuda", 0)
    d_x = torch.from_numpy(X[:B]).to(dev)
    d_out = torch.empty((B, model.OUT_COLS), dtype=torch.float64, device=dev)
    d_grad = torch.empty((B, md.state_len), dtype=torch.float64, device=dev)
    d_st = torch.empty(B, dtype=torch.int32, device=dev)
    n0 = ev.kernel_launches()
    ev.eval_grad_device(B, d_x.data_ptr(), d_out.data_ptr(), d_grad.data_ptr(), d_st.data_ptr(),
                        torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert ev.kernel_launches() - n0 == 5          # residual split, contraction, FP64 fall-back (flagged chains only), posterior, scalar assembly
    assert np.array_equal(d_out.cpu().numpy(), out[:B]) and np.array_equal(d_grad.cpu().numpy(), grad[:B])
    assert np.array_equal(d_st.cpu().numpy(), st[:B])


def test_theta_packed_entry_point(full_size):
    """mcd_eval_grad_theta == fromVectorWith -> mcd_eval_grad -> toVector (app/Hamiltonian.hs:49-60)"""
    md, X, ev, out, grad, st = full_size
    B = 1500
    mask = ev.mask().astype(bool)
    theta = X[:B][:, mask][:, ::-1].copy()               # toVector: free entries, reversed
    assert np.array_equal(theta[3], ev.to_vector(X[3]))
    o, gt, s = ev.eval_grad_theta(theta, X[0])            # fixed entries are identical across chains
    assert np.array_equal(o, out[:B]) and np.array_equal(s, st[:B])
    assert np.array_equal(gt, grad[:B][:, mask][:, ::-1])
    # small tree, ragged batch, against the oracle's dual-number gradient
    md2, z = load_fixture("24-leaves-braces", 1)
    ev2 = binding.Evaluator(md2)
    orc = O.Oracle(md2)
    Xs = z["states"][1:4]
    th = np.array([orc.to_vector(x) for x in Xs])
    o2, g2, s2 = ev2.eval_grad_theta(th, Xs[0])
    for b in range(3):
        gd = orc.grad_dual(Xs[b])
        assert grad_relerr(g2[b], orc.to_vector(gd)).max() < TOL
    ev2.close()


@pytest.mark.parametrize("n_leaves,B", [(12, 100), (40, 64), (300, 257), (1000, 130)])
def test_sparse_precision_likelihood(n_leaves, B):
    """LikelihoodData `Sparse` (app/Probability.hs:178-184): association-list precision as the reference
    stores it; small trees run densified in the fused kernel, large ones through the CSR contraction"""
    md, h = synth.synthetic_model(n_leaves, seed=500 + n_leaves, likelihood=model.LIK_SPARSE, n_cal=3, n_con=2,
                                  n_brace=1, clock_model=1 + (n_leaves % 3))
    assert len(md.sparse_val) < 20 * md.dim + 400
    X = synth.synthetic_states(md, h, B)
    ev = binding.Evaluator(md)
    out, grad, st = ev.eval_grad(X)
    oo, og, ost = O.Oracle(md).eval_grad(X, nthreads=4)
    assert np.array_equal(st, ost)
    assert relerr(out[:, :7], oo).max() < TOL and grad_relerr(grad, og).max() < TOL
    o2, s2 = ev.eval(X)
    assert relerr(o2[:, :7], oo).max() < TOL
    ev.close()
    # an asymmetric association list (only the upper triangle listed twice as heavy) is the same quadratic form
    up = md.sparse_row <= md.sparse_col
    md.sparse_val = np.where(md.sparse_row == md.sparse_col, md.sparse_val, 2.0 * md.sparse_val)[up]
    md.sparse_row, md.sparse_col = md.sparse_row[up], md.sparse_col[up]
    ev = binding.Evaluator(md)
    o3, g3, s3 = ev.eval_grad(X)
    assert relerr(o3[:, :7], oo).max() < TOL and grad_relerr(g3, og).max() < TOL
    ev.close()


def test_value_only_cholesky_path():
    """mcd_eval uses quad = |L^T dx|^2 (triangular contraction, half the flops): with the factor supplied
    by the caller, factorised by the library, and -- for an indefinite matrix -- the symmetric fallback"""
    md, h = synth.synthetic_model(300, seed=77, n_cal=4, n_con=2, n_brace=1)
    X = synth.synthetic_states(md, h, 200)
    oo, ost = O.Oracle(md).eval(X, nthreads=4)
    for supply in (True, False):
        ev = binding.Evaluator(md, supply_cholesky=supply)
        ev.set_contraction("dmma")       # the triangular path belongs to the FP64 DMMA contraction
        out, st = ev.eval(X)
        assert relerr(out[:, :7], oo).max() < TOL and np.array_equal(st, ost)
        ev.close()
    # indefinite "precision": Cholesky fails -> symmetric product; the density formula is still evaluated
    P = md.precision.copy()
    P[5, 5] = -abs(P[5, 5])
    md2 = model.ModelDesc(parent=md.parent, mean=md.mean, precision=P, logdet_sigma=md.logdet_sigma, ht=md.ht,
                          cal_node=md.cal_node, cal_lo=md.cal_lo, cal_lo_p=md.cal_lo_p, cal_hi=md.cal_hi,
                          cal_hi_p=md.cal_hi_p)
    ev = binding.Evaluator(md2)
    ev.set_contraction("dmma")
    out, st = ev.eval(X[:50])
    o2, s2 = O.Oracle(md2).eval(X[:50])
    assert relerr(out[:, :7], o2).max() < TOL
    ev.set_contraction("i8s7")
    out, st = ev.eval(X[:50])
    assert relerr(out[:, :7], o2).max() < TOL
    ev.close()


@pytest.mark.parametrize("n_leaves,B", [(60, 70), (300, 257), (1000, 200)])
def test_contraction_pipes_agree_with_oracle(n_leaves, B):
    """the FP64 DMMA contraction and the INT8 tensor-core (Ozaki-split) contraction with 7 and 6 base-256 digit
    planes all meet the 1e-10 bar here (6 planes is the documented coarse mode: ~2^-8 times the error).  The default
    of a new handle is i8s7."""
    md, h = synth.synthetic_model(n_leaves, seed=1234 + n_leaves, n_cal=4, n_con=2, n_brace=1)
    X = synth.synthetic_states(md, h, B)
    oo, og, ost = O.Oracle(md).eval_grad(X, nthreads=8)
    ev = binding.Evaluator(md)
    assert ev.get_contraction() == 7
    errs = {}
    for mode, tol in (("i8s7", TOL), ("dmma", TOL), ("i8s6", TOL)):
        ev.set_contraction(mode)
        out, grad, st = ev.eval_grad(X)
        assert np.array_equal(st, ost)
        errs[mode] = (relerr(out[:, :7], oo).max(), grad_relerr(grad, og).max())
        assert errs[mode][0] < tol and errs[mode][1] < tol, (mode, errs[mode])
        o2, s2 = ev.eval(X)              # value-only entry point on the same pipe
        assert relerr(o2[:, :7], oo).max() < tol
    # seven planes are of FP64-GEMM quality: within a small factor of the DMMA path's own rounding error
    assert errs["i8s7"][0] < max(50 * errs["dmma"][0], 1e-13) and errs["i8s7"][1] < max(50 * errs["dmma"][1], 1e-13)
    ev.close()


def _check_against_oracle(md, X, n_oracle=None, expect_fallback=None):
    """value / gradient / status parity of the default (INT8) path against the oracle, with the scalar gradient block on its
    own scale; returns (status words of the CUDA path, worst errors)"""
    ev = binding.Evaluator(md)
    assert ev.get_contraction() == 7
    out, grad, st = ev.eval_grad(X)
    n = len(X) if n_oracle is None else n_oracle
    orc = O.Oracle(md)
    oo, og, ost = orc.eval_grad(X[:n], nthreads=4)
    assert np.array_equal(st[:n] & ~model.ST_FP64_FALLBACK, ost)
    ev_val = relerr(out[:n, :7], oo).max()
    eg = grad_relerr(grad[:n], og).max()
    es = grad_relerr_scalar_block(grad[:n], og, md.n_nodes).max()
    assert ev_val < TOL and eg < TOL and es < TOL, (ev_val, eg, es)
    if expect_fallback is not None:
        assert np.array_equal((st & model.ST_FP64_FALLBACK) != 0, expect_fallback), np.nonzero((st & model.ST_FP64_FALLBACK) != 0)[0]
    # the value-only entry point (triangular Cholesky-factor contraction, its own FP64 fall-back for flagged chains)
    outv, stv = ev.eval(X)
    assert relerr(outv[:, :7], out[:, :7]).max() < TOL and np.array_equal(stv, st)
    # the FP64 tensor-instruction contraction on the same inputs: same answer
    ev.set_contraction("dmma")
    out2, grad2, st2 = ev.eval_grad(X)
    assert relerr(out2[:, :7], out[:, :7]).max() < TOL and grad_relerr(grad2, grad).max() < TOL
    ev.close()
    return st, (ev_val, eg, es)


def test_int8_contraction_with_badly_scaled_precision():
    """Sigma^-1 = D P D with the diagonal D spread over ten orders of magnitude (branches of very different length and
    variance): the power-of-two equilibration built at mcd_create makes the digit planes those of P itself"""
    md, h = synth.synthetic_model(200, seed=41, n_cal=3, n_con=2, n_brace=1)
    K = md.dim
    rng = np.random.default_rng(5)
    dvec = 10.0 ** rng.uniform(-5.0, 5.0, K)
    P = md.precision * dvec[:, None] * dvec[None, :]
    P = 0.5 * (P + P.T)
    md.precision = P
    md.logdet_sigma = md.logdet_sigma - 2.0 * float(np.sum(np.log(dvec)))
    X = synth.synthetic_states(md, h, 300, seed=77)
    _check_against_oracle(md, X, n_oracle=64)


def test_int8_contraction_with_the_inverse_of_a_sample_covariance():
    """what `prepare` really produces (app/Main.hs:214-230): the LU inverse of a sample covariance of correlated branch lengths --
    ill-conditioned, rows with a wide dynamic range, symmetric only up to rounding"""
    md, h = synth.synthetic_model(150, seed=43, n_cal=2)
    K = md.dim
    rng = np.random.default_rng(9)
    n_s = 3 * K
    A = rng.normal(size=(K, 12)) * (0.3 * md.mean[:, None])           # a few strong common factors (rate variation shared by clades)
    Z = rng.normal(size=(n_s, 12)) @ A.T + rng.normal(size=(n_s, K)) * (0.02 * md.mean[None, :] + 1e-5) + md.mean[None, :]
    mean = Z.mean(axis=0)
    cov = np.cov(Z, rowvar=False, ddof=1)
    sign, logdet = np.linalg.slogdet(cov)
    assert sign == 1.0
    prec = np.linalg.inv(cov)                                           # not symmetrised, like invlndet
    assert np.linalg.cond(cov) > 1e4 and np.abs(prec - prec.T).max() > 0.0
    md.mean, md.precision, md.logdet_sigma = mean, prec, float(logdet)
    X = synth.synthetic_states(md, h, 300, seed=78)
    _check_against_oracle(md, X, n_oracle=64)


def test_int8_contraction_falls_back_to_fp64_for_chains_with_outlier_residuals():
    """digits are relative to a chain's largest (standardised) residual: chains where one residual dwarfs the others by more
    than 512 x the mean are recomputed in plain FP64 and say so in their status word; results stay within the bar either way"""
    md, h = synth.synthetic_model(200, seed=47, n_cal=3)
    N = md.n_nodes
    B = 260
    X = synth.synthetic_states(md, h, B, seed=79)
    X[:, 5 + N + 1:] = 1.0 + 0.01 * (X[:, 5 + N + 1:] - 1.0)          # residuals of ordinary size
    expect = np.zeros(B, bool)
    for b, factor in ((3, 1e4), (130, 1e7), (259, 3e3)):                # one branch rate off by orders of magnitude
        leaf = int(np.nonzero(md.child0 < 0)[0][5 + b % 7])
        X[b, 5 + N + leaf] *= factor
        expect[b] = True
    st, errs = _check_against_oracle(md, X, expect_fallback=expect)
    assert ((st & model.ST_FP64_FALLBACK) != 0).sum() == 3


def test_int8_contraction_is_bit_reproducible_and_tiling_invariant():
    """integer accumulation: any batch split and any chain position gives bit-identical results"""
    md, h = synth.synthetic_model(300, seed=4321, n_cal=4, n_con=2, n_brace=1)
    X = synth.synthetic_states(md, h, 300)
    ev = binding.Evaluator(md)
    out, grad, st = ev.eval_grad(X)
    out2, grad2, _ = ev.eval_grad(X)
    assert np.array_equal(out, out2) and np.array_equal(grad, grad2)
    perm = np.random.default_rng(0).permutation(300)
    o3, g3, _ = ev.eval_grad(X[perm][:77])
    assert np.array_equal(o3, out[perm][:77]) and np.array_equal(g3, grad[perm][:77])
    ev.close()


def test_contraction_kernel_variants_are_bit_identical():
    """the three INT8 contraction kernels -- plane-granular pipeline (default), stage-granular pipeline of round 1 (MCD_OZ_V1) and
    CTA pairs with cta_group::2 MMAs (MCD_OZ_PAIR) -- multiply the same integers: outputs and gradients agree bit for bit, also for a
    batch that is an odd number of 128-chain tiles (the pair kernel pads to 256) and on the value-only triangular path"""
    import hashlib
    import os
    import subprocess
    import sys
    code = (
        "import sys, hashlib, numpy as np\n"
        "sys.path.insert(0, %r)\n"
        "from mcmc_date_b200 import binding, synth\n"
        "md, h = synth.synthetic_model(300, seed=31, n_cal=3, n_con=2, n_brace=1)\n"
        "ev = binding.Evaluator(md)\n"
        "m = hashlib.sha256()\n"
        "for B in (640, 1100, 130):\n"
        "    X = synth.synthetic_states(md, h, B, seed=500 + B)\n"
        "    out, grad, st = ev.eval_grad(X)\n"
        "    out2, st2 = ev.eval(X)\n"
        "    for a in (out, grad, st, out2, st2): m.update(np.ascontiguousarray(a).tobytes())\n"
        "print(m.hexdigest())\n" % os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    digests = {}
    for name, env in (("default", {}), ("v1", {"MCD_OZ_V1": "1"}), ("pair", {"MCD_OZ_PAIR": "1"})):
        e = dict(os.environ, **env)
        for k in ("MCD_OZ_V1", "MCD_OZ_PAIR"):
            if k not in env:
                e.pop(k, None)
        r = subprocess.run([sys.executable, "-c", code], env=e, capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, (name, r.stderr[-500:])
        digests[name] = r.stdout.strip().splitlines()[-1]
    assert digests["default"] == digests["v1"] == digests["pair"], digests


def test_int8_contraction_non_finite_states():
    """NaN / inf in a state poison that chain only (NaN propagates like in the FP64 product), on both pipes"""
    md, h = synth.synthetic_model(300, seed=99, n_cal=4, n_con=2, n_brace=1)
    X = synth.synthetic_states(md, h, 140)
    X[3, 3 + 7] = np.nan                 # a height
    X[5, 5 + md.n_nodes + 9] = np.inf    # a rate
    X[130, 2] = np.nan                   # H
    orc = O.Oracle(md)
    oo, og, ost = orc.eval_grad(X, nthreads=4)
    ev = binding.Evaluator(md)
    for mode in ("i8s7", "dmma"):
        ev.set_contraction(mode)
        out, grad, st = ev.eval_grad(X)
        bad = np.array([3, 5, 130])
        good = np.setdiff1d(np.arange(140), bad)
        assert np.array_equal(st[good], ost[good]) and relerr(out[good, :7], oo[good]).max() < TOL
        assert grad_relerr(grad[good], og[good]).max() < TOL
        assert not np.isfinite(out[bad, 6]).any() and np.array_equal(np.isnan(out[bad, 6]), np.isnan(oo[bad, 6]))
        assert np.array_equal(st[bad], ost[bad])
    ev.close()


def _oracle_leapfrog(orc, x0, theta, mom, inv_mass, eps, L):
    """half kick, drift, ..., half kick with the ORACLE's gradient (host loop, float64)"""
    theta, mom = theta.copy(), mom.copy()

    def grad(th):
        X = np.array([orc.from_vector(x0, t) for t in th])
        out, g, st = orc.eval_grad(X)
        return out, np.array([orc.to_vector(gi) for gi in g]), st

    out, g, st = grad(theta)
    h0 = -out[:, 6] + 0.5 * (mom * mom * inv_mass).sum(axis=1)
    stat = st.copy()
    for l in range(L):
        mom = mom + 0.5 * eps[:, None] * g
        theta = theta + eps[:, None] * inv_mass * mom
        out, g, st = grad(theta)
        stat |= st
        mom = mom + 0.5 * eps[:, None] * g
    h1 = -out[:, 6] + 0.5 * (mom * mom * inv_mass).sum(axis=1)
    return theta, mom, out, np.stack([h0, h1], axis=1), stat


@pytest.mark.parametrize("n_leaves,B,L", [(24, 33, 6), (300, 20, 4)])
def test_device_resident_leapfrog(n_leaves, B, L):
    """mcd_leapfrog == the same integrator driven step by step from the host with the oracle's gradient"""
    md, h = synth.synthetic_model(n_leaves, seed=31 + n_leaves, n_cal=3, n_con=2, n_brace=0)
    X = synth.synthetic_states(md, h, B)
    ev = binding.Evaluator(md)
    orc = O.Oracle(md)
    theta = np.array([orc.to_vector(x) for x in X])
    rng = np.random.default_rng(4)
    _, g0, _ = orc.eval_grad(X)
    gth = np.array([orc.to_vector(gi) for gi in g0])
    inv_mass = 1.0 / np.maximum(1.0, np.abs(gth).max(axis=0)) ** 2       # crude per-parameter scale
    mom = rng.normal(size=theta.shape) / np.sqrt(inv_mass)
    eps = np.full(B, 0.02) * rng.uniform(0.5, 1.0, B)
    th, pm, out, en, st = ev.leapfrog(theta, mom, X[0], inv_mass, eps, L)
    rt, rp, ro, re, rs = _oracle_leapfrog(orc, X[0], theta, mom, inv_mass, eps, L)
    assert np.array_equal(st, rs)
    ok = rs == 0
    assert ok.sum() >= B // 2
    assert (np.abs(th - rt)[ok] <= 1e-9 * np.maximum(1.0, np.abs(rt[ok]))).all()
    assert (np.abs(pm - rp)[ok] <= 1e-9 * np.maximum(1.0, np.abs(rp[ok]).max(axis=1, keepdims=True))).all()
    assert relerr(out[ok, :7], ro[ok]).max() < 1e-9 and relerr(en[ok], re[ok]).max() < 1e-9
    # the integrator is symplectic and time-reversible: flip the momentum, integrate back, land on the start
    th2, pm2, _, en2, _ = ev.leapfrog(th[ok], -pm[ok], X[0], inv_mass, eps[ok], L)
    assert (np.abs(th2 - theta[ok]) <= 1e-8 * np.maximum(1.0, np.abs(theta[ok]))).all()
    assert relerr(en2[:, 1], en[ok, 0]).max() < 1e-9
    ev.close()


@pytest.mark.parametrize("n_leaves,B,eps0,max_depth", [(24, 24, 0.02, 6), (300, 12, 0.004, 5)])
def test_device_nuts_matches_host_restatement(n_leaves, B, eps0, max_depth):
    """mcd_nuts == the same Algorithm-3 tree building driven chain by chain from the host with the oracle's
    gradient and bit-identical Philox uniforms (tests/nuts_ref.py): tree depths, leapfrog counts, valid-point counts
    and the chosen points agree"""
    import nuts_ref
    md, h = synth.synthetic_model(n_leaves, seed=61 + n_leaves, n_cal=3, n_con=2, n_brace=0)
    X = synth.synthetic_states(md, h, B)
    ev = binding.Evaluator(md)
    orc = O.Oracle(md)
    theta = np.array([orc.to_vector(x) for x in X])
    _, g0, _ = orc.eval_grad(X)
    gth = np.array([orc.to_vector(gi) for gi in g0])
    inv_mass = 1.0 / np.maximum(1.0, np.abs(gth).max(axis=0)) ** 2
    rng = np.random.default_rng(8)
    mom = rng.normal(size=theta.shape) / np.sqrt(inv_mass)
    eps = eps0 * np.exp(rng.uniform(np.log(0.3), np.log(60.0), B))   # short trees, deep trees and divergences
    seed, iteration = 0x1234567890ABCDEF, 7
    th, out, acc, info, st = ev.nuts(theta, X[0], inv_mass, eps, max_depth=max_depth, seed=seed, iteration=iteration,
                                     momentum0=mom)
    depths = set()
    for b in range(B):
        rt, ro, ra, rinfo, rst = nuts_ref.nuts_chain(orc, X[0], theta[b], mom[b], inv_mass, eps[b], max_depth, seed,
                                                     iteration, b)
        assert tuple(info[b]) == tuple(rinfo), (b, info[b], rinfo)
        assert st[b] == rst
        assert (np.abs(th[b] - rt) <= 1e-8 * np.maximum(1.0, np.abs(rt))).all()
        assert relerr(out[b, :7], ro).max() < 1e-8 and abs(acc[b] - ra) < 1e-8
        depths.add(int(info[b, 0]))
    assert len(depths) >= 3          # the batch really mixes trees of different depths
    # same seed / iteration: bit-identical; another iteration: other slice variables and directions
    th2, out2, acc2, info2, _ = ev.nuts(theta, X[0], inv_mass, eps, max_depth=max_depth, seed=seed, iteration=iteration,
                                        momentum0=mom)
    assert np.array_equal(th, th2) and np.array_equal(info, info2)
    ev.close()


def test_device_nuts_draws_its_own_momenta():
    """momentum0 = NULL: momenta ~ N(0, M) from Philox on the device; reproducible, and the chain moves"""
    md, h = synth.synthetic_model(24, seed=85, n_cal=3, n_con=2, n_brace=0)
    X = synth.synthetic_states(md, h, 512)
    ev = binding.Evaluator(md)
    orc = O.Oracle(md)
    theta = np.array([orc.to_vector(x) for x in X])
    inv_mass = np.full(ev.D, 1e-4)
    a = ev.nuts(theta, X[0], inv_mass, 0.01, max_depth=5, seed=99, iteration=3)
    b = ev.nuts(theta, X[0], inv_mass, 0.01, max_depth=5, seed=99, iteration=3)
    c = ev.nuts(theta, X[0], inv_mass, 0.01, max_depth=5, seed=99, iteration=4)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[3], b[3])
    assert not np.array_equal(a[0], c[0])
    moved = (np.abs(a[0] - theta).max(axis=1) > 0)
    assert moved.mean() > 0.5 and (a[3][:, 1] >= 1).all()
    assert np.isfinite(a[1][:, 6]).all() and (a[2] >= 0).all() and (a[2] <= 1).all()
    # the returned ln-posterior parts are those of the returned point
    o2, _, _ = ev.eval_grad(np.array([orc.from_vector(X[0], t) for t in a[0]]))
    assert relerr(a[1][:, :7], o2[:, :7]).max() < 1e-9
    ev.close()


def _mh_setup(n_leaves, B, clock=1, n_brace=1, seed_off=0):
    import mh_ref
    md, h = synth.synthetic_model(n_leaves, seed=301 + n_leaves + seed_off, clock_model=clock, n_cal=3, n_con=2, n_brace=n_brace)
    X = synth.synthetic_states(md, h, B)
    ev = binding.Evaluator(md)
    orc = O.Oracle(md)
    parent = [int(p) for p in md.parent]
    braces = [[int(x) for x in md.brace_node[md.brace_off[b]:md.brace_off[b + 1]]] for b in range(md.n_brace)]
    ev.chains_set(X)
    Xr = X.copy()
    out_r, st_r = orc.eval(Xr)
    out_r = np.concatenate([out_r, np.zeros((B, 1))], axis=1) if out_r.shape[1] == 7 else out_r
    return mh_ref, md, X, ev, orc, parent, braces, Xr, out_r, st_r


def _mh_compare(ev, Xr, out_r, st_r):
    Xd, out_d, st_d = ev.chains_get()
    assert (np.abs(Xd - Xr) <= 1e-11 * np.maximum(1.0, np.abs(Xr))).all()
    assert relerr(out_d[:, :7], out_r[:, :7]).max() < 1e-9 and np.array_equal(st_d, st_r)
    return Xd, out_d, st_d


@pytest.mark.parametrize("n_leaves,B", [(24, 64), (300, 40)])
def test_device_resident_mh_proposals_match_host_restatement(n_leaves, B):
    """chains resident in HBM: slide-node and scale-sub-tree proposals (first-party code of the reference), value-only
    evaluation, accept / restore -- step by step against tests/mh_ref.py (oracle values, same Philox uniforms)"""
    mh_ref, md, X, ev, orc, parent, braces, Xr, out_r, st_r = _mh_setup(n_leaves, B, n_brace=0)
    child, size, inner, inner_list = mh_ref.topology(parent)
    Xd, out_d, st_d = ev.chains_get()
    assert np.array_equal(Xd, X) and relerr(out_d[:, :7], out_r[:, :7]).max() < TOL and np.array_equal(st_d, st_r)
    root_child = 1 if child[1] else child[0][1]          # an inner child of the root: lifted with jacobianRootBranch
    deep = max(inner_list, key=lambda i: (size[i] > 3, -size[i]))
    steps = [(mh_ref.SLIDE_NODE, root_child, 0.05, True), (mh_ref.SCALE_SUBTREE, root_child, 0.05, True),
             (mh_ref.SLIDE_NODE, -1, 0.02, False), (mh_ref.SCALE_SUBTREE, -1, 0.02, False), (mh_ref.SLIDE_NODE, deep, 0.5, False),
             (mh_ref.SCALE_SUBTREE, deep, 0.01, False), (mh_ref.SLIDE_NODE, -1, 0.01, False), (mh_ref.SLIDE_NODE, -1, 0.01, False)]
    n_acc = n_rej = 0
    for it, (kind, node, sd, jac) in enumerate(steps):
        acc_d = ev.mh_step(kind, node, sd, tune=1.3, use_root_jacobian=jac, seed=4242, iteration=it)
        acc_r = mh_ref.mh_step(orc, parent, Xr, out_r, st_r, kind, node, sd, 1.3, jac, 4242, it)
        assert np.array_equal(acc_d, acc_r), (it, np.nonzero(acc_d != acc_r))
        n_acc += int((acc_r == 1).sum())
        n_rej += int((acc_r == 0).sum())
    assert n_acc > B and n_rej > B // 4                    # both branches of the accept kernel were exercised
    Xd, out_d, st_d = _mh_compare(ev, Xr, out_r, st_r)
    # the resident ln-posterior parts are those of the resident states
    o2, s2 = ev.eval(Xd)
    assert relerr(out_d[:, :7], o2[:, :7]).max() < 1e-12
    with pytest.raises(RuntimeError):
        ev.mh_step(mh_ref.SLIDE_NODE, 0, 0.1)              # the root does not slide
    ev.close()


@pytest.mark.parametrize("n_leaves,B,clock", [(24, 48, 1), (24, 48, 3), (200, 24, 0)])
def test_every_proposal_of_the_reference_cycle(n_leaves, B, clock):
    """all 17 proposal kinds of app/Definitions.hs:145-279 (time tree, rate tree, contrary, braces, hyper-parameters),
    each restated literally in tests/mh_ref.py: proposed state, Hastings factor, Jacobian, acceptance -- step by step"""
    mh_ref, md, X, ev, orc, parent, braces, Xr, out_r, st_r = _mh_setup(n_leaves, B, clock=clock, n_brace=2, seed_off=7)
    child, size, inner, inner_list = mh_ref.topology(parent)
    assert child[1] and child[child[0][1]], "both root children must be inner nodes for the pulley (regenerate the tree)"
    R = mh_ref
    big = max(inner_list, key=lambda i: size[i])
    # (kind, node, param, tune, lifted with jacobianRootBranch)   -- parameters as in app/Definitions.hs
    pt = 1.0 if n_leaves < 100 else 0.02                   # the pulley moves every node: small steps on large trees
    steps = [(R.PULLEY, 0, 0.01, pt, True), (R.SLIDE_NODE, 1, 0.01, 2.0, True), (R.SCALE_SUBTREE, big, 0.01, pt, False),
             (R.SLIDE_BRACE, 0, 0.01, 0.01, False), (R.SLIDE_BRACE, -1, 0.01, 0.02, False),
             (R.SCALE_BRANCH, 1, 100.0, 20.0, True), (R.SCALE_BRANCH, -1, 100.0, 20.0, False),
             (R.SCALE_RATE_SUBTREE, big, 100.0, 2.0, False), (R.SCALE_RATE_SUBTREE, -1, 100.0, 5.0, False),
             (R.SCALE_NORM_TREE_CONTRA_M, 0, 100.0, 1.0, True), (R.SCALE_NORM_TREE_CONTRA_H, 0, 100.0, 1.0, True),
             (R.SCALE_VAR_TREE, 0, 100.0, 1.0, True), (R.SCALE_VAR_TREE_AUTO, 0, 100.0, 1.0, True),
             (R.SLIDE_NODE_CONTRA, 1, 0.1, 0.3, True), (R.SLIDE_NODE_CONTRA, -1, 0.1, 0.1, False),
             (R.SCALE_SUBTREE_CONTRA, big, 0.1, 0.1, False), (R.SCALE_SUBTREE_CONTRA, -1, 0.1, 0.05, False),
             (R.SLIDE_BRACE_CONTRA, 1, 0.1, 0.002, False), (R.SLIDE_ROOT_CONTRA, 0, 10.0, 0.5, True),
             (R.SCALE_RATES_TREE_CONTRA, 0, 0.1, 0.1, True),
             (R.SCALE_SCALAR, 0, 10.0, 1.0, False), (R.SCALE_SCALAR, 1, 10.0, 1.0, False), (R.SCALE_SCALAR, 2, 3000.0, 1.0, False),
             (R.SCALE_SCALAR, 3, 10.0, 1.0, False), (R.SCALE_SCALAR, 4, 10.0, 0.5, False), (R.SCALE_H_M_CONTRA, 0, 10.0, 1.0, False),
             (R.SCALE_VAR_TREE, 0, 2.0, 4.0, True),          # shape 0.5 < 1: boosted gamma draw, many NaN rates
             (R.PULLEY, 0, 0.01, 3.0 * pt, True), (R.SLIDE_NODE_CONTRA, -1, 0.1, 0.1, False)]
    seen_acc, seen_rej = set(), set()
    for it, (kind, node, param, tune, jac) in enumerate(steps):
        acc_d = ev.mh_step(kind, node, param, tune=tune, use_root_jacobian=jac, seed=777, iteration=100 + it)
        acc_r, lrs = R.mh_step(orc, parent, Xr, out_r, st_r, kind, node, param, tune, jac, 777, 100 + it, braces=braces,
                               return_lr=True)
        assert np.array_equal(acc_d, acc_r), (it, kind, np.nonzero(acc_d != acc_r), lrs[acc_d != acc_r])
        _mh_compare(ev, Xr, out_r, st_r)
        if (acc_r == 1).any():
            seen_acc.add(kind)
        if (acc_r == 0).any():
            seen_rej.add(kind)
    assert seen_acc == set(range(17)) and len(seen_rej) >= 15, (sorted(seen_acc), sorted(seen_rej))
    with pytest.raises(RuntimeError):
        ev.mh_step(R.SCALE_SCALAR, 7, 10.0)
    with pytest.raises(RuntimeError):
        ev.mh_step(R.SLIDE_BRACE, 5, 0.01)
    ev.close()


@pytest.mark.parametrize("n_leaves,B", [(60, 96), (24, 96), (24, 400)])
def test_mh_cycle_equals_single_steps_and_counts(n_leaves, B):
    """mcd_mh_cycle = the same proposals enqueued back to back: identical chains, acceptance counts per list entry.  24 leaves: the
    whole call is ONE launch (mh_small_cycle_kernel; one warp per CTA for 96 chains, eight for 400) against one launch per step"""
    mh_ref, md, X, ev, orc, parent, braces, Xr, out_r, st_r = _mh_setup(n_leaves, B, n_brace=1)
    R = mh_ref
    props = [(R.SLIDE_NODE, -1, 0.01, 1.0, 0, 3), (R.SCALE_BRANCH, -1, 100.0, 10.0, 0, 2), (R.SCALE_SCALAR, 3, 10.0, 1.0, 0, 1),
             (R.SLIDE_BRACE, 0, 0.01, 0.01, 0, 1), (R.SCALE_NORM_TREE_CONTRA_M, 0, 100.0, 1.0, 1, 1)]
    acc, inv, nxt = ev.mh_cycle(props, n_iterations=2, seed=5, iteration0=10)
    assert nxt == 10 + 2 * 8
    Xc, out_c, st_c = ev.chains_get()
    ev.chains_set(X)
    it, acc2 = 10, np.zeros(len(props), np.int64)
    for sweep in range(2):
        for i, (kind, node, param, tune, jac, rep) in enumerate(props):
            for _ in range(rep):
                a = ev.mh_step(kind, node, param, tune=tune, use_root_jacobian=bool(jac), seed=5, iteration=it)
                acc2[i] += int((a == 1).sum())
                it += 1
    Xs, out_s, st_s = ev.chains_get()
    assert np.array_equal(Xc, Xs) and np.array_equal(out_c, out_s) and np.array_equal(st_c, st_s)
    assert np.array_equal(acc.astype(np.int64), acc2) and (inv == 0).all() and (acc2 > 0).all()
    ev.close()


@pytest.mark.parametrize("n_leaves", [24, 150])
def test_heated_chains_and_mc3_swaps(n_leaves):
    """MC3 (app/Main.hs:476-479): heated acceptance ratio and slot swaps against the host restatement; stepping-stone heats
    (likelihood only, app/Main.hs:511-543).  24 leaves: single-launch small-tree step; 150 leaves: fused incremental step."""
    C, G = 4, 12
    mh_ref, md, X, ev, orc, parent, braces, Xr, out_r, st_r = _mh_setup(n_leaves, C * G, n_brace=0)
    R = mh_ref
    ladder = np.array([1.0, 0.7, 0.4, 0.1])
    ev.mc3_configure(C * G, 0, C, ladder, ladder)
    slot = np.arange(C * G) % C
    cos = np.arange(C * G)
    n_sw = 0
    for it in range(6):
        kind, node, param = [(R.SLIDE_NODE, -1, 0.05), (R.SCALE_BRANCH, -1, 20.0), (R.SCALE_SUBTREE, -1, 0.05)][it % 3]
        beta = ladder[slot]
        acc_d = ev.mh_step(kind, node, param, tune=1.0, seed=9, iteration=it)
        acc_r = R.mh_step(orc, parent, Xr, out_r, st_r, kind, node, param, 1.0, False, 9, it, beta_prior=beta, beta_lik=beta)
        assert np.array_equal(acc_d, acc_r)
        pair = -1 if it % 2 else it // 2 % (C - 1)
        sw_d = ev.mc3_swap(pair, seed=9, iteration=1000 + it)
        sw_r = R.mc3_swap(out_r[:, 3:5], slot, cos, ladder, ladder, C, pair, 9, 1000 + it)
        assert np.array_equal(sw_d, sw_r) and np.array_equal(ev.mc3_slots(), slot)
        n_sw += int(sw_r.sum())
    assert 0 < n_sw < 6 * G
    _mh_compare(ev, Xr, out_r, st_r)
    # every group still holds every temperature exactly once
    assert (np.sort(slot.reshape(G, C), axis=1) == np.arange(C)).all()
    # stepping stone: the likelihood alone is heated
    betas = np.linspace(0.0, 1.0, C * G)
    ev.mc3_configure(C * G, 0, C * G, np.ones(C * G), betas)
    for it in range(3):
        acc_d = ev.mh_step(R.SLIDE_NODE, -1, 0.05, seed=10, iteration=it)
        acc_r = R.mh_step(orc, parent, Xr, out_r, st_r, R.SLIDE_NODE, -1, 0.05, 1.0, False, 10, it, beta_lik=betas)
        assert np.array_equal(acc_d, acc_r)
    _mh_compare(ev, Xr, out_r, st_r)
    ev.mc3_configure(0, 0, 0)                                # cold again
    acc_d = ev.mh_step(R.SLIDE_NODE, -1, 0.05, seed=11, iteration=0)
    acc_r = R.mh_step(orc, parent, Xr, out_r, st_r, R.SLIDE_NODE, -1, 0.05, 1.0, False, 11, 0)
    assert np.array_equal(acc_d, acc_r)
    ev.close()


def test_mh_and_mc3_do_not_depend_on_the_split_over_handles():
    """the multi-GPU layout on one GPU: one handle holding all chains vs two handles holding half each (global chain offsets,
    all-gathered swap statistics): bit-identical chains and slot tables -- what tools/mc3_bench.py relies on over NCCL"""
    import torch
    import mh_ref as R
    C, G = 8, 6
    n = C * G
    md, h = synth.synthetic_model(24, seed=77, clock_model=3, n_cal=3, n_con=2, n_brace=0)
    X = synth.synthetic_states(md, h, n)
    ladder = 1.0 / (1.0 + 0.3 * np.arange(C))
    whole = binding.Evaluator(md)
    halves = [binding.Evaluator(md), binding.Evaluator(md)]
    whole.chains_set(X)
    whole.mc3_configure(n, 0, C, ladder, ladder)
    for r, ev in enumerate(halves):
        ev.chains_set(X[r * n // 2:(r + 1) * n // 2])
        ev.mc3_configure(n, r * n // 2, C, ladder, ladder)
    props = [(R.SLIDE_NODE, -1, 0.05, 1.0, 0, 2), (R.SCALE_BRANCH, -1, 50.0, 1.0, 0, 2), (R.SLIDE_NODE_CONTRA, -1, 0.1, 0.3, 0, 1),
             (R.SCALE_SCALAR, 4, 10.0, 1.0, 0, 1)]
    gathered = torch.empty((n, 2), dtype=torch.float64, device="cuda")
    k = 0
    for it in range(4):
        _, _, k2 = whole.mh_cycle(props, 1, seed=3, iteration0=k)
        for ev in halves:
            ev.mh_cycle(props, 1, seed=3, iteration0=k)
        k = k2
        for r, ev in enumerate(halves):                       # the "all-gather"
            ev.chains_stats_device(gathered[r * n // 2:(r + 1) * n // 2].data_ptr())
        torch.cuda.synchronize()
        a = whole.mc3_swap(-1, seed=4, iteration=it)
        bs = [ev.mc3_swap(-1, seed=4, iteration=it, d_stats_global=gathered.data_ptr()) for ev in halves]
        assert np.array_equal(a, bs[0]) and np.array_equal(a, bs[1])
    Xw, ow, sw = whole.chains_get()
    parts = [ev.chains_get() for ev in halves]
    assert np.array_equal(Xw, np.concatenate([p[0] for p in parts])) and np.array_equal(ow, np.concatenate([p[1] for p in parts]))
    sl = whole.mc3_slots()
    assert np.array_equal(sl, halves[0].mc3_slots()) and np.array_equal(sl, halves[1].mc3_slots())
    assert (sl != np.arange(n) % C).any() and (np.abs(Xw - X).max(axis=1) > 0).all()
    with pytest.raises(RuntimeError):
        halves[1].mc3_swap(-1, seed=4, iteration=99)          # groups span handles: the gathered table is required
    for ev in [whole] + halves:
        ev.close()


@pytest.mark.parametrize("clock", [0, 1, 2, 3])
def test_incremental_evaluation_of_small_moves(clock):
    """small moves scored from the cached y = Sigma^-1 dx (mh_delta_kernel) vs the same run with full evaluations, and vs
    the oracle-driven host restatement: same decisions, same chains; values within accumulated rounding.  Includes chains in
    the near-critical birth-death regime, a refresh every 5 steps, and proposals that fall back to the full evaluation."""
    import mh_ref as R
    n_leaves, B = 300, 48
    md, h = synth.synthetic_model(n_leaves, seed=911 + clock, clock_model=clock, n_cal=4, n_con=3, n_brace=2)
    X = synth.synthetic_states(md, h, B)
    X[::7, 1] = X[::7, 0] + 3e-7                                # |lambda - mu| < 1e-6: literal D/E recursion
    parent = [int(p) for p in md.parent]
    braces = [[int(x) for x in md.brace_node[md.brace_off[b]:md.brace_off[b + 1]]] for b in range(md.n_brace)]
    child, size, inner, inner_list = R.topology(parent)
    small = [i for i in inner_list if 4 <= size[i] <= 32]
    big = max(inner_list, key=lambda i: size[i])
    # sub trees too large for the per-chain path but not hanging off the root: rank-limited contraction over their k-blocks
    mids = sorted([i for i in inner_list if size[i] > 64 and parent[i] != 0], key=lambda i: size[i])
    assert len(mids) >= 2
    mid, mid2 = mids[0], mids[-1]
    orc = O.Oracle(md)
    inc, full = binding.Evaluator(md), binding.Evaluator(md)
    inc.mh_set_incremental(True, refresh_every=5)
    full.mh_set_incremental(False)
    inc.chains_set(X)
    full.chains_set(X)
    assert inc.mh_incremental_active() and not full.mh_incremental_active()
    Xr = X.copy()
    out_r, st_r = orc.eval(Xr)
    assert (st_r[::7] & 8).all()
    steps = [(R.SLIDE_NODE, -1, 0.01, 1.0, False), (R.SLIDE_NODE, 1, 0.01, 1.0, True), (R.SCALE_BRANCH, -1, 100.0, 10.0, False),
             (R.SLIDE_NODE_CONTRA, -1, 0.1, 0.1, False), (R.SLIDE_BRACE, 0, 0.01, 0.01, False), (R.SLIDE_BRACE_CONTRA, 1, 0.1, 0.002, False),
             (R.SCALE_SUBTREE, small[0], 0.01, 1.0, False), (R.SCALE_SUBTREE_CONTRA, small[-1], 0.1, 0.1, False),
             (R.SCALE_RATE_SUBTREE, small[len(small) // 2], 100.0, 5.0, False),
             (R.SCALE_SCALAR, 0, 10.0, 0.02, False),            # full path; keeps / moves chains in and out of near-criticality
             (R.SCALE_SUBTREE, big, 0.01, 0.02, False),         # hangs off the root: evaluated from scratch
             (R.SCALE_SUBTREE, mid, 0.01, 0.1, False), (R.SCALE_SUBTREE_CONTRA, mid2, 0.1, 0.02, False),
             (R.SCALE_RATE_SUBTREE, mid, 100.0, 3.0, False), (R.SCALE_SUBTREE_CONTRA, mid, 0.1, 0.03, False),
             (R.SCALE_RATE_SUBTREE, mid2, 100.0, 3.0, False), (R.SCALE_SUBTREE, mid2, 0.01, 0.05, False),
             (R.SLIDE_NODE, -1, 0.01, 1.0, False), (R.SCALE_NORM_TREE_CONTRA_M, 0, 100.0, 1.0, True),
             (R.SLIDE_NODE_CONTRA, 1, 0.1, 0.3, True), (R.SCALE_BRANCH, 1, 100.0, 10.0, True), (R.SLIDE_NODE, -1, 0.01, 3.0, False),
             (R.SLIDE_BRACE, -1, 0.01, 0.02, False), (R.SCALE_BRANCH, -1, 100.0, 30.0, False), (R.SLIDE_NODE, -1, 0.01, 0.5, False)]
    n_acc = n_rej = 0
    for it, (kind, node, param, tune, jac) in enumerate(steps):
        a_i = inc.mh_step(kind, node, param, tune=tune, use_root_jacobian=jac, seed=31, iteration=it)
        a_f = full.mh_step(kind, node, param, tune=tune, use_root_jacobian=jac, seed=31, iteration=it)
        a_r = R.mh_step(orc, parent, Xr, out_r, st_r, kind, node, param, tune, jac, 31, it, braces=braces)
        assert np.array_equal(a_i, a_f) and np.array_equal(a_i, a_r), (it, kind)
        Xi, oi, si = inc.chains_get()
        Xf, of, sf = full.chains_get()
        assert np.array_equal(Xi, Xf) and np.array_equal(si, sf)
        assert relerr(oi[:, :7], of[:, :7]).max() < 1e-11, (it, kind, relerr(oi[:, :7], of[:, :7]).max())
        n_acc += int((a_r == 1).sum())
        n_rej += int((a_r == 0).sum())
    assert n_acc > 4 * B and n_rej > 2 * B
    _mh_compare(inc, Xr, out_r, st_r)
    # the cached values are those of the resident states
    Xi, oi, si = inc.chains_get()
    o2, s2 = inc.eval(Xi)
    assert relerr(oi[:, :7], o2[:, :7]).max() < 1e-11 and np.array_equal(si, s2)
    # an invalid chain at upload: that set of chains runs with full evaluations
    Xbad = X.copy()
    Xbad[3, 3 + inner_list[5]] = 2.0
    inc.chains_set(Xbad)
    assert not inc.mh_incremental_active()
    inc.chains_set(X)
    assert inc.mh_incremental_active()
    inc.close()
    full.close()


def test_nuts_on_resident_chains_equals_the_host_buffer_call():
    """mcd_chains_nuts = mcd_nuts on the chains' own HMC vectors, written back to their state rows (the Hamiltonian proposal
    of the reference's cycle, app/Definitions.hs:276-278), followed by ordinary proposals on the moved chains"""
    import mh_ref as R
    md, h = synth.synthetic_model(150, seed=77, n_cal=3, n_con=2, n_brace=1)
    B = 40
    X = synth.synthetic_states(md, h, B)
    X[:, 2] = X[0, 2]                                           # with calibrations H is free; nothing else is shared
    ev = binding.Evaluator(md)
    orc = O.Oracle(md)
    theta = np.array([orc.to_vector(x) for x in X])
    inv_mass = np.full(ev.D, 1e-4)
    th1, out1, acc1, info1, st1 = ev.nuts(theta, X[0], inv_mass, 0.01, max_depth=4, seed=5, iteration=2)
    ev.chains_set(X)
    acc2, info2, st2 = ev.chains_nuts(inv_mass, 0.01, max_depth=4, seed=5, iteration=2)
    assert np.array_equal(info1, info2) and np.array_equal(acc1, acc2) and np.array_equal(st1, st2)
    Xd, out_d, st_d = ev.chains_get()
    Xexp = np.array([orc.from_vector(X[b], th1[b]) for b in range(B)])
    assert np.array_equal(Xd, Xexp) and (np.abs(Xd - X).max(axis=1) > 0).mean() > 0.5
    assert relerr(out_d[:, :7], out1[:, :7]).max() < 1e-10
    # the cached contraction results follow: incremental moves after the transition agree with the oracle-driven restatement
    parent = [int(p) for p in md.parent]
    Xr = Xd.copy()
    out_r, st_r = orc.eval(Xr)
    for it in range(3):
        a_d = ev.mh_step(R.SLIDE_NODE, -1, 0.01, seed=8, iteration=it)
        a_r = R.mh_step(orc, parent, Xr, out_r, st_r, R.SLIDE_NODE, -1, 0.01, 1.0, False, 8, it)
        assert np.array_equal(a_d, a_r)
    _mh_compare(ev, Xr, out_r, st_r)
    ev.close()


@pytest.mark.parametrize("clock", [0, 1])
def test_mh_sampler_draws_from_the_prior(clock):
    """End-to-end statistical check of proposals + Hastings factors + Jacobians + acceptance: without data (likelihood NoData)
    the chains must converge to the prior, whose hyper-parameter marginals are known in closed form -- m ~ Exponential(ht),
    v ~ Gamma(3/2, 1/6) (every branch-rate density is normalised given v), E[r_i] = 1.  The chains start far away
    (v = 1, m = 4 / ht, rates 2).  Only proposals whose stated Jacobian is exact are used (see the two reference quirks
    pinned in tests/test_host_logic.py), none lifted with the root-branch Jacobian, so the target is the prior itself."""
    import mh_ref as R
    B = 4096
    md, h = synth.synthetic_model(12, seed=5 + clock, clock_model=clock, likelihood=model.LIK_NONE, n_cal=0)
    X = synth.synthetic_states(md, h, B)
    N = md.n_nodes
    X[:, 3 + N] = 4.0 / md.ht
    X[:, 4 + N] = 1.0
    X[:, 5 + N + 1:5 + 2 * N] = 2.0
    ev = binding.Evaluator(md)
    ev.chains_set(X)
    props = [(R.SCALE_SCALAR, 0, 10.0, 3.0, 0, 1), (R.SCALE_SCALAR, 1, 10.0, 3.0, 0, 1), (R.SCALE_SCALAR, 3, 10.0, 5.0, 0, 2),
             (R.SCALE_SCALAR, 4, 10.0, 5.0, 0, 2), (R.SCALE_BRANCH, -1, 10.0, 5.0, 0, 2 * N), (R.SLIDE_NODE, -1, 0.1, 1.0, 0, 8),
             (R.SCALE_SUBTREE, -1, 0.1, 1.0, 0, 4), (R.SCALE_RATE_SUBTREE, -1, 20.0, 2.0, 0, 4),
             (R.SLIDE_NODE_CONTRA, -1, 0.1, 1.0, 0, 4), (R.SCALE_NORM_TREE_CONTRA_M, 0, 50.0, 1.0, 0, 1),
             (R.SCALE_VAR_TREE_AUTO, 0, 50.0, 1.0, 0, 1)]
    k = 0
    _, _, k = ev.mh_cycle(props, 150, seed=21, iteration0=k)       # burn-in
    ms, vs, rs = [], [], []
    for _ in range(6):                                            # thinned samples (chains are independent of each other)
        _, _, k = ev.mh_cycle(props, 25, seed=21, iteration0=k)
        Xd, out, st = ev.chains_get()
        assert np.isfinite(out[:, 6]).all() and (st == 0).all()
        ms.append(Xd[:, 3 + N]); vs.append(Xd[:, 4 + N]); rs.append(Xd[:, 5 + N + 1:5 + 2 * N].mean(axis=1))
    m_all, v_all, r_all = np.concatenate(ms), np.concatenate(vs), np.concatenate(rs)
    # tolerances: 6 standard errors of a mean over >= 4096 independent chains (thinned repeats only help)
    se = lambda sd: 6.0 * sd / np.sqrt(B)
    assert abs(m_all.mean() - 1.0 / md.ht) < se(1.0 / md.ht), (m_all.mean(), 1.0 / md.ht)
    assert abs(v_all.mean() - 0.25) < se(np.sqrt(1.5) / 6.0), v_all.mean()
    assert abs(v_all.var() - 1.5 / 36.0) < 0.15 * 1.5 / 36.0, v_all.var()
    assert abs(r_all.mean() - 1.0) < 0.02, r_all.mean()
    # Kolmogorov distance of v to Gamma(3/2, 1/6) and of m to Exponential(ht) on the last sample
    from scipy import stats
    assert stats.kstest(vs[-1], stats.gamma(1.5, scale=1.0 / 6.0).cdf).statistic < 0.04
    assert stats.kstest(ms[-1], stats.expon(scale=1.0 / md.ht).cdf).statistic < 0.04
    ev.close()


def test_mh_edge_cases():
    """smallest trees, a leaf child of the root, argument checks that mirror the reference's proposal constructors"""
    import mh_ref as R
    # 3 leaves: ((a,b),c) -- one inner node below the root, the root's other child is a leaf
    parent = np.array([-1, 0, 1, 1, 0], np.int32)
    md = model.ModelDesc(parent=parent, mean=np.array([0.9, 0.3, 0.35]), precision=np.diag([30.0, 90.0, 80.0]) + 1.0,
                         logdet_sigma=-11.0, clock_model=model.UNCORRELATED_GAMMA)
    B = 64
    rng = np.random.default_rng(1)
    X = np.zeros((B, md.state_len))
    X[:, 0], X[:, 1], X[:, 2] = 1.2, 0.7, 1.0
    X[:, 3] = 1.0
    X[:, 4] = rng.uniform(0.2, 0.8, B)
    X[:, 3 + 5], X[:, 4 + 5] = 1.0, 0.3
    X[:, 5 + 5 + 1:] = rng.lognormal(0, 0.2, (B, 4))
    ev = binding.Evaluator(md)
    orc = O.Oracle(md)
    ev.chains_set(X)
    Xr = X.copy()
    out_r, st_r = orc.eval(Xr)
    par = [int(p) for p in parent]
    for it, (kind, node, param, jac) in enumerate([(R.SLIDE_NODE, 1, 0.2, True), (R.SCALE_SUBTREE, -1, 0.2, True),
                                                    (R.SLIDE_NODE_CONTRA, 1, 0.2, True), (R.SCALE_BRANCH, 4, 20.0, True),
                                                    (R.SCALE_RATES_TREE_CONTRA, 0, 0.2, True), (R.SCALE_VAR_TREE, 0, 20.0, True),
                                                    (R.SCALE_RATE_SUBTREE, 1, 20.0, True), (R.SCALE_SUBTREE_CONTRA, 1, 0.2, True)]):
        a_d = ev.mh_step(kind, node, param, use_root_jacobian=jac, seed=2, iteration=it)
        a_r = R.mh_step(orc, par, Xr, out_r, st_r, kind, node, param, 1.0, jac, 2, it)
        assert np.array_equal(a_d, a_r), (it, kind)
    _mh_compare(ev, Xr, out_r, st_r)
    for bad in [(R.PULLEY, 0, 0.1),               # pulleyUltrametric: a sub tree of the root is a leaf
                (R.SLIDE_NODE, 4, 0.1),           # path leads to a leaf
                (R.SLIDE_NODE, 0, 0.1),           # the root
                (R.SCALE_BRANCH, 0, 10.0),        # the stem
                (R.SLIDE_BRACE, 0, 0.1),          # no braces in this model
                (R.SLIDE_NODE, 1, -0.1),          # standard deviation
                (99, 0, 0.1)]:
        with pytest.raises(RuntimeError):
            ev.mh_step(*bad)
    with pytest.raises(RuntimeError):
        ev.mc3_swap(0)                                          # no ladder configured
    with pytest.raises(RuntimeError):
        ev.mc3_configure(B, 0, 7, np.ones(7), np.ones(7))       # 64 chains do not split into groups of 7
    ev.close()
    # a cherry: 2 leaves, no inner node below the root
    md2 = model.ModelDesc(parent=np.array([-1, 0, 0], np.int32), mean=np.array([1.0]), precision=np.array([[4.0]]), logdet_sigma=-1.4)
    ev2 = binding.Evaluator(md2)
    x2 = np.array([[1.0, 0.5, 1.0, 1.0, 0.0, 0.0, 1.0, 0.2, 0.0, 1.1, 0.9]])
    ev2.chains_set(x2)
    with pytest.raises(RuntimeError):
        ev2.mh_step(R.SLIDE_NODE, -1, 0.1)                      # nothing to slide
    with pytest.raises(RuntimeError):
        ev2.mh_step(R.SCALE_RATES_TREE_CONTRA, 0, 0.1)          # "no internal nodes to scale"
    a = ev2.mh_step(R.SCALE_BRANCH, -1, 20.0, seed=1, iteration=0)
    assert a[0] in (0, 1)
    ev2.close()


def test_full_size_mh_cycle_is_consistent_with_fresh_evaluations():
    """BASELINE's largest shape (1000 leaves, K = 1997): a slice of the reference's proposal cycle -- small moves (fused
    incremental step), large sub-tree moves (k-block-range contraction), global moves (from scratch) -- then the resident
    ln-posterior parts, statuses and cached contraction results must be those of a fresh evaluation of the resident states,
    and a second run from the same start must reproduce the chains bit for bit."""
    from mcmc_date_b200 import mh_cycle
    md, h = synth.synthetic_model(1000, seed=synth.BASE_SEED + 4, n_cal=16, n_con=8, n_brace=4)
    B = 1024
    X = synth.synthetic_states(md, h, B)
    props = mh_cycle.reference_cycle(md)
    rng = np.random.default_rng(0)
    pick = sorted(rng.choice(len(props), size=260, replace=False))
    sub = [props[i][:5] + (1,) for i in pick] + [p[:5] + (1,) for p in props if p[0] in (binding.MH_SLIDE_BRACE, binding.MH_PULLEY,
                                                                                       binding.MH_SCALE_VAR_TREE)]
    ev = binding.Evaluator(md, max_batch=B)
    ev.mh_set_incremental(True, refresh_every=100)
    runs = []
    for _ in range(2):
        ev.chains_set(X)
        assert ev.mh_incremental_active()
        acc, inv, _ = ev.mh_cycle(sub, 1, seed=77, iteration0=5)
        runs.append(ev.chains_get())
        assert (inv == 0).all() and (acc > 0).sum() > 0.9 * len(sub)
    (X1, o1, s1), (X2, o2, s2) = runs
    assert np.array_equal(X1, X2) and np.array_equal(o1, o2) and np.array_equal(s1, s2)
    assert (np.abs(X1 - X).max(axis=1) > 0).all()
    of, sf = ev.eval(X1)
    assert relerr(o1[:, :7], of[:, :7]).max() < 1e-10 and np.array_equal(s1, sf)
    # the cached y still describes the resident states: one more small move per chain, scored incrementally, then checked again
    ev.mh_step(binding.MH_SLIDE_NODE, -1, 0.001, seed=78, iteration=0)
    X3, o3, s3 = ev.chains_get()
    of3, sf3 = ev.eval(X3)
    assert relerr(o3[:, :7], of3[:, :7]).max() < 1e-10 and np.array_equal(s3, sf3)
    ev.close()


def test_asynchronous_theta_calls_overlap_safely():
    """mcd_eval_grad_theta_async / mcd_wait: several calls in flight on different host buffers (and different inputs) give
    exactly what the synchronous call gives; tickets older than the ring and a changed base state are handled"""
    import torch
    md, h = synth.synthetic_model(300, seed=12, n_cal=3, n_con=2, n_brace=1)
    B = 700
    ev = binding.Evaluator(md)
    mask = ev.mask().astype(bool)
    D = ev.D
    Xs = [synth.synthetic_states(md, h, B, seed=100 + i) for i in range(3)]
    for X in Xs:
        X[:, 2] = Xs[0][0, 2]
    thetas = [torch.from_numpy(np.ascontiguousarray(X[:, mask][:, ::-1])).pin_memory() for X in Xs]
    base = torch.from_numpy(Xs[0][0].copy()).pin_memory()
    ref = [ev.eval_grad_theta(t.numpy(), base.numpy()) for t in thetas]
    outs = [(torch.empty((B, model.OUT_COLS), dtype=torch.float64).pin_memory(), torch.empty((B, D), dtype=torch.float64).pin_memory(),
             torch.empty(B, dtype=torch.int32).pin_memory()) for _ in range(3)]
    tickets = []
    for rep in range(4):                                   # 12 calls: the ticket ring (8) wraps
        for i in range(3):
            o, g, s_ = outs[i]
            tickets.append(ev.eval_grad_theta_async_ptr(B, thetas[i].data_ptr(), base.data_ptr(), o.data_ptr(), g.data_ptr(), s_.data_ptr()))
    assert tickets == list(range(tickets[0], tickets[0] + 12))
    for t in (tickets[0], tickets[-1], tickets[5]):
        ev.wait(t)
    ev.synchronize()
    for i in range(3):
        o, g, s_ = outs[i]
        assert np.array_equal(o.numpy(), ref[i][0]) and np.array_equal(g.numpy(), ref[i][1]) and np.array_equal(s_.numpy(), ref[i][2])
    # a different base state (other fixed entries) is picked up
    base2 = base.clone()
    base2[5 + md.n_nodes] = 0.5                             # the rate stem is fixed: it must come from the base state
    t = ev.eval_grad_theta_async_ptr(B, thetas[0].data_ptr(), base2.data_ptr(), outs[0][0].data_ptr(), outs[0][1].data_ptr(), outs[0][2].data_ptr())
    ev.wait(t)
    ref2 = ev.eval_grad_theta(thetas[0].numpy(), base2.numpy())
    assert np.array_equal(outs[0][0].numpy(), ref2[0]) and np.array_equal(outs[0][1].numpy(), ref2[1])
    with pytest.raises(RuntimeError):
        ev.wait(10 ** 6)
    # full-state forms
    hx = [torch.from_numpy(X).pin_memory() for X in Xs[:2]]
    S = md.state_len
    fo = [(torch.empty((B, model.OUT_COLS), dtype=torch.float64).pin_memory(), torch.empty((B, S), dtype=torch.float64).pin_memory(),
           torch.empty(B, dtype=torch.int32).pin_memory()) for _ in range(2)]
    tk = [ev.eval_grad_async_ptr(B, hx[i].data_ptr(), fo[i][0].data_ptr(), fo[i][1].data_ptr(), fo[i][2].data_ptr()) for i in range(2)]
    for t in tk:
        ev.wait(t)
    for i in range(2):
        o, g, s_ = ev.eval_grad(Xs[i])
        assert np.array_equal(fo[i][0].numpy(), o) and np.array_equal(fo[i][1].numpy(), g) and np.array_equal(fo[i][2].numpy(), s_)
    t = ev.eval_async_ptr(B, hx[0].data_ptr(), fo[0][0].data_ptr(), fo[0][2].data_ptr())
    ev.wait(t)
    o, s_ = ev.eval(Xs[0])
    assert np.array_equal(fo[0][0].numpy(), o) and np.array_equal(fo[0][2].numpy(), s_)
    ev.close()


def test_mixed_entry_points_are_ordered_without_explicit_waits():
    """The pipelined host-buffer calls leave work in flight on four streams over the handle's shared work buffers; device-pointer
    calls run on the caller's streams and the resident-chain calls on the library's.  The library orders them against each
    other itself: mixing them WITHOUT mcd_wait / mcd_synchronize in between gives what the serial sequence gives."""
    import torch
    md, h = synth.synthetic_model(300, seed=12, n_cal=3, n_con=2, n_brace=1)
    B = 1400
    ev = binding.Evaluator(md)
    mask = ev.mask().astype(bool)
    D, S = ev.D, md.state_len
    XA, XB, XC = (synth.synthetic_states(md, h, B, seed=300 + i) for i in range(3))
    for X in (XB, XC):
        X[:, 2] = XA[0, 2]
    theta = torch.from_numpy(np.ascontiguousarray(XA[:, mask][:, ::-1])).pin_memory()
    base = torch.from_numpy(XA[0].copy()).pin_memory()
    # serial reference
    refA = ev.eval_grad_theta(theta.numpy(), base.numpy())
    refB = ev.eval_grad(XB)
    ev.chains_set(XC)
    accC = ev.mh_step(binding.MH_SCALE_BRANCH, -1, 50.0, seed=9, iteration=1)
    refC = ev.chains_get()
    ev.synchronize()
    dev = torch.device("cuda", 0)
    dXB = torch.from_numpy(XB).to(dev)
    for rep in range(3):
        oA = (torch.empty((B, model.OUT_COLS), dtype=torch.float64).pin_memory(), torch.empty((B, D), dtype=torch.float64).pin_memory(),
              torch.empty(B, dtype=torch.int32).pin_memory())
        d_out = torch.zeros((B, model.OUT_COLS), dtype=torch.float64, device=dev)
        d_grad = torch.zeros((B, S), dtype=torch.float64, device=dev)
        d_st = torch.zeros(B, dtype=torch.int32, device=dev)
        s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        torch.cuda.synchronize()
        t = ev.eval_grad_theta_async_ptr(B, theta.data_ptr(), base.data_ptr(), oA[0].data_ptr(), oA[1].data_ptr(), oA[2].data_ptr())
        # no wait: a device-pointer call on one user stream, the resident-chain calls on the library's stream, then another
        # device-pointer call on a second user stream
        ev.eval_grad_device(B, dXB.data_ptr(), d_out.data_ptr(), d_grad.data_ptr(), d_st.data_ptr(), s1.cuda_stream)
        ev.chains_set(XC)
        acc = ev.mh_step(binding.MH_SCALE_BRANCH, -1, 50.0, seed=9, iteration=1)
        d_out2 = torch.zeros_like(d_out)
        d_grad2 = torch.zeros_like(d_grad)
        d_st2 = torch.zeros_like(d_st)
        ev.eval_grad_device(B, dXB.data_ptr(), d_out2.data_ptr(), d_grad2.data_ptr(), d_st2.data_ptr(), s2.cuda_stream)
        t2 = ev.eval_grad_theta_async_ptr(B, theta.data_ptr(), base.data_ptr(), oA[0].data_ptr(), oA[1].data_ptr(), oA[2].data_ptr())
        Xc, oc, sc = ev.chains_get()
        ev.wait(t)
        ev.wait(t2)
        torch.cuda.synchronize()
        assert np.array_equal(oA[0].numpy(), refA[0]) and np.array_equal(oA[1].numpy(), refA[1]) and np.array_equal(oA[2].numpy(), refA[2])
        for o_, g_, s_ in ((d_out, d_grad, d_st), (d_out2, d_grad2, d_st2)):
            assert np.array_equal(o_.cpu().numpy(), refB[0]) and np.array_equal(g_.cpu().numpy(), refB[1]) and np.array_equal(s_.cpu().numpy(), refB[2])
        assert np.array_equal(acc, accC) and np.array_equal(Xc, refC[0]) and np.array_equal(oc, refC[1]) and np.array_equal(sc, refC[2])
    ev.close()


def test_library_collective_on_a_single_rank():
    """mcd_comm_* / mcd_allgather_stats with a one-rank communicator: NCCL is found at run time, the gathered table equals the local
    (ln prior, ln likelihood) columns (tools/mc3_bench.py runs the same calls on 2..8 GPUs and compares with torch.distributed)"""
    import torch
    md, h = synth.synthetic_model(24, seed=24, n_cal=3, n_con=2, n_brace=1)
    B = 64
    X = synth.synthetic_states(md, h, B)
    ev = binding.Evaluator(md)
    ev.chains_set(X)
    with pytest.raises(RuntimeError):
        ev.allgather_stats(0)
    uid = binding.Evaluator.comm_unique_id()
    assert len(uid) == 128 and any(uid)
    ev.comm_init(1, 0, uid)
    dev = torch.device("cuda", 0)
    g = torch.full((B, 2), float("nan"), dtype=torch.float64, device=dev)
    with pytest.raises(RuntimeError):
        ev.allgather_stats(0)                       # null buffer
    ev.allgather_stats(g.data_ptr())
    _, out, _ = ev.chains_get()
    assert np.array_equal(g.cpu().numpy(), out[:, 3:5])
    ev.comm_destroy()
    with pytest.raises(RuntimeError):
        ev.allgather_stats(g.data_ptr())            # no communicator any more
    ev.close()


def test_error_behaviour():
    md, z = load_fixture("12-leaves-variable-rate")
    ev = binding.Evaluator(md)
    L = binding.load_library()
    # empty batch is fine; null buffers are an error with a message, not a crash
    assert L.mcd_eval(ev.h, 0, None, None, None) == 0
    assert L.mcd_eval(ev.h, 4, None, None, None) != 0 and b"null" in L.mcd_last_error(ev.h)
    ev.close()
    # model validation mirrors the reference's load-time `error`s
    def expect_fail(**kw):
        base = dict(parent=md.parent, mean=md.mean, precision=md.precision, logdet_sigma=md.logdet_sigma, ht=md.ht)
        base.update(kw)
        with pytest.raises((RuntimeError, ValueError)):
            binding.Evaluator(model.ModelDesc(**base))
    expect_fail(ht=0.0)                                                   # exponential: rate <= 0
    expect_fail(brace_off=[0, 2], brace_node=[2, 5], brace_sd=[0.0])      # braceSoftF: sd <= 0
    expect_fail(con_young=[2], con_old=[16], con_p=[1.5])                 # probabilityMass
    p = md.precision.copy()
    p[0, 1] += 1.0
    expect_fail(precision=p)                                              # not symmetric
    expect_fail(parent=np.array([-1, 0, 1, 1, 1, 0, 0], np.int32), mean=np.zeros(5), precision=np.eye(5))  # multifurcating
    expect_fail(parent=np.array([-1, 0, 1, 2, 2], np.int32), mean=np.zeros(3), precision=np.eye(3))       # unary root
