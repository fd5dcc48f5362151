"""Host-side logic (no GPU): index maps bit-exact against the literal restatement, mask / vector
packing, Newick + prepare + loaders, and that the C-ABI library loads and exports every symbol
include/mcmcdate_b200.h declares."""
import ctypes
import os
import re

import numpy as np
import pytest

from mcmc_date_b200 import binding, model, prepare, sharding, synth, tree
from oracle import oracle as O
from util import FIXTURES, load_fixture

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _model_for(parent):
    n = len(parent)
    return model.ModelDesc(parent=parent, mean=np.zeros(n - 2), precision=np.zeros(0), logdet_sigma=0.0,
                           likelihood=model.LIK_NONE)


@pytest.mark.parametrize("seed", range(40))
def test_branch_index_closed_form_is_bit_exact(seed):
    """closed form of getBranches + sumFirstTwo == literal list construction"""
    rng = np.random.default_rng(seed)
    n = int(rng.integers(2, 60))
    parent = tree.random_topology(n, rng)
    md = _model_for(parent)
    lit = O.Oracle(md).branch_index()
    assert np.array_equal(tree.branch_index(parent), lit)
    assert sorted(set(lit[1:])) == list(range(len(parent) - 2))   # onto 0..K-1
    assert (lit[1:] == 0).sum() == 2                              # exactly the two root branches merge


def test_branch_index_edge_cases():
    # two leaves: K = 1
    p = np.array([-1, 0, 0], np.int32)
    assert tree.branch_index(p).tolist() == [-1, 0, 0]
    # left root child is a leaf: range 2 <= i <= s_l is empty
    p = np.array([-1, 0, 0, 2, 2], np.int32)
    assert tree.branch_index(p).tolist() == [-1, 0, 0, 1, 2]
    assert np.array_equal(O.Oracle(_model_for(p)).branch_index(), tree.branch_index(p))
    with pytest.raises(ValueError):
        tree.branch_index(np.array([-1, 0, 1, 1], np.int32))  # root not bifurcating


def test_newick_roundtrip_and_preorder():
    t = tree.parse_newick("((f:0.3,e:0.26):0.19,((d:0.5,c:0.01)x:0.54,(b:0.3,a:0.26):0.37):0)[c];")
    parent, c0, c1, names, lens = tree.flatten_preorder(t)
    assert parent.tolist() == [-1, 0, 1, 1, 0, 4, 5, 5, 4, 8, 8]
    assert names[2] == "f" and names[5] == "x" and lens[4] == 0.0
    assert c0.tolist()[:2] == [1, 2] and c1.tolist()[0] == 4
    h = tree.node_heights_from_lengths(parent, np.array([0, 1, 1, 1, 1, 0.5, 0.5, 0.5, 0.5, 0.5, 0.5]))
    assert h[0] == 1.0


@pytest.mark.parametrize("name", FIXTURES)
def test_mask_and_vector_packing(name):
    """getMask / toVector / fromVectorWith (app/Hamiltonian.hs:33-60): free = everything except root
    height, leaf heights, rate stem, and H unless calibrations exist; theta is in REVERSED order."""
    md, z = load_fixture(name)
    orc = O.Oracle(md)
    N = md.n_nodes
    mask = orc.mask
    assert np.array_equal(mask, z["mask"])
    n = md.n_leaves
    D = 2 + (1 if md.calibrations_available else 0) + (n - 2) + 2 + (N - 1)
    assert mask.sum() == D
    assert mask[3] == 0 and mask[5 + N] == 0 and mask[0] == mask[1] == mask[3 + N] == mask[4 + N] == 1
    x = z["states"][1]
    th = orc.to_vector(x)
    assert np.array_equal(th, x[mask.astype(bool)][::-1])
    th2 = th + 1.0
    y = orc.from_vector(x, th2)
    assert np.array_equal(y[mask.astype(bool)], x[mask.astype(bool)] + 1.0)
    assert np.array_equal(y[~mask.astype(bool)], x[~mask.astype(bool)])


def test_prepare_matches_numpy_and_is_invertible():
    path = "/root/reference/tests/12-leaves-variable-rate/data/test.treelist"
    if not os.path.exists(path):
        pytest.skip("reference checkout not present (GPU box)")
    pr = prepare.prepare_from_treelist(open(path).read())
    md, z = load_fixture("12-leaves-variable-rate")
    assert np.array_equal(pr["parent"], md.parent)
    assert np.allclose(pr["mean"], md.mean, rtol=0, atol=0)
    assert np.allclose(pr["precision"] @ pr["cov"], np.eye(21), atol=1e-8)
    assert np.allclose(pr["precision"], md.precision, rtol=1e-13, atol=0)


def test_data_and_meantree_formats_roundtrip(tmp_path):
    """<name>.data (aeson LikelihoodDataStore JSON) and <name>.meantree as `prepare` writes them
    (app/Main.hs:75-99,286-307) feed the same model the tree list gave"""
    md, z = load_fixture("12-leaves-variable-rate")
    names = [str(x) for x in z["leaf_names"]]
    lengths = np.abs(np.random.default_rng(0).normal(0.2, 0.05, md.n_nodes))
    lengths[0] = 0.0
    f = tmp_path / "a.data"
    prepare.write_data_file(str(f), "FullS", mean=md.mean, precision=md.precision, logdet_sigma=md.logdet_sigma)
    obj = __import__("json").load(open(f))
    assert obj["tag"] == "FullS" and len(obj["contents"]) == 3 and len(obj["contents"][1]) == md.dim
    mt = prepare.mean_tree_newick(md.parent, names, lengths)
    md2, x0 = prepare.model_from_data_file(str(f), mt)
    assert np.array_equal(md2.parent, md.parent) and np.array_equal(md2.precision, md.precision)
    assert md2.logdet_sigma == md.logdet_sigma and np.array_equal(md2.mean, md.mean)
    assert x0[3] == 1.0 and len(x0) == md.state_len
    # inner nodes carry their pre-order index as label (assignIndices)
    p2, _, _, nm2, ln2 = tree.flatten_preorder(tree.parse_newick(mt))
    assert nm2[0] == "0" and nm2[1] == "1" and np.allclose(ln2, lengths)
    for kind, kw in (("UnivariateS", dict(mean=md.mean, variances=1.0 / np.diag(md.precision))),
                     ("NoLikelihoodS", {}),
                     ("SparseS", dict(mean=md.mean, sparse=([0, 1, 1], [0, 1, 0], [2.0, 3.0, 0.5]), logdet_sigma=1.5))):
        prepare.write_data_file(str(f), kind, **kw)
        d = prepare.read_data_file(str(f))
        assert d["likelihood"] == {"UnivariateS": 1, "NoLikelihoodS": 2, "SparseS": 3}[kind]
    assert d["sparse"][2].tolist() == [2.0, 3.0, 0.5]


def test_loaders_pin_survey_integer_fixtures():
    """node indices derived in SURVEY.md 8c"""
    _, z12 = load_fixture("12-leaves-variable-rate")
    assert z12["parent"].tolist() == [-1, 0, 1, 2, 2, 1, 5, 5, 7, 8, 8, 7, 0, 12, 13, 14, 14, 16, 16, 13, 12, 20, 20]
    assert z12["cal_node"].tolist() == [0, 14, 2] and z12["con_young"].tolist() == [2] and z12["con_old"].tolist() == [16]
    assert float(z12["ht"]) == 1050.0
    _, z24 = load_fixture("24-leaves-braces")
    assert len(z24["parent"]) == 47 and np.nonzero(z24["parent"] == 0)[0].tolist() == [1, 24]
    assert z24["cal_node"].tolist() == [0, 25, 12]
    assert z24["con_young"].tolist() == [24, 26] and z24["con_old"].tolist() == [9, 19]
    assert z24["brace_node"].tolist() == [6, 36] and float(z24["brace_sd"][0]) == 1e-4
    _, z6 = load_fixture("06-leaves-constant-rate")
    assert z6["parent"].tolist() == [-1, 0, 1, 1, 0, 4, 5, 5, 4, 8, 8] and z6["cal_node"].tolist() == [0]
    _, z7 = load_fixture("mtcdnapri-7-leaves")
    assert len(z7["parent"]) == 13 and float(z7["ht"]) == 50.0


def test_initial_state_is_valid():
    md, z = load_fixture("24-leaves-braces")
    x0 = z["states"][0]
    N = md.n_nodes
    h = x0[3:3 + N]
    assert h[0] == 1.0 and (h[md.child0 < 0] == 0).all()
    assert (h[md.parent[1:]] - h[1:] > 0).all()
    out, st = O.Oracle(md).eval(x0[None])
    assert np.isfinite(out[0, 6]) and st[0] == model.ST_NEARCRIT  # initWith has lambda == mu == 1


def test_synthetic_generators_are_seeded_and_valid():
    md, h = synth.synthetic_model(50, seed=3, n_cal=4, n_con=3, n_brace=2)
    md2, h2 = synth.synthetic_model(50, seed=3, n_cal=4, n_con=3, n_brace=2)
    assert np.array_equal(md.parent, md2.parent) and np.array_equal(md.precision, md2.precision)
    assert np.array_equal(md.precision, md.precision.T)
    assert np.linalg.eigvalsh(md.precision).min() > 0
    X = synth.synthetic_states(md, h, 32)
    N = md.n_nodes
    t = X[:, 3 + md.parent[1:]] - X[:, 3 + np.arange(1, N)]
    assert (t > 0).all()
    out, st = O.Oracle(md).eval(X)
    assert np.isfinite(out).all() and (st == 0).all()


def test_shard_ranges_partition_chains():
    for B, W in [(8192, 8), (64, 8), (8192, 1), (12, 4)]:
        spans = [sharding.shard_range(B, W, r) for r in range(W)]
        assert spans[0][0] == 0 and spans[-1][1] == B
        assert all(spans[i][1] == spans[i + 1][0] for i in range(W - 1))
        assert len({b - a for a, b in spans}) == 1          # the all-gather and the MC3 slot tables need equal shards
    for B, W in [(10, 4), (7, 3)]:
        with pytest.raises(ValueError):
            sharding.shard_range(B, W, 0)


def test_library_loads_and_exports_every_declared_symbol():
    """no compute calls: dlopen + symbol lookup only"""
    header = open(os.path.join(ROOT, "include", "mcmcdate_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(mcd_[a-z0-9_]+)\s*\(", header)))
    assert "mcd_eval_grad" in declared and "mcd_create" in declared
    assert os.path.exists(binding.LIB_PATH), "libmcmcdate_b200.so not built"
    lib = ctypes.CDLL(binding.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert sorted(binding.EXPORTS) == declared
    L = binding.load_library()
    assert b"sm_100a" in L.mcd_version()


def test_create_fails_loudly_without_gpu_or_with_bad_model():
    """no CPU fallback: mcd_create reports an error instead of evaluating on the host"""
    import torch
    md, _ = load_fixture("06-leaves-constant-rate")
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="no CUDA device|CUDA"):
            binding.Evaluator(md)
    bad = model.ModelDesc(parent=np.array([-1, 0, 1, 1], np.int32), mean=np.zeros(2), precision=np.eye(2),
                          logdet_sigma=0.0)
    with pytest.raises(RuntimeError):
        binding.Evaluator(bad)


def test_philox_known_answers():
    """the counter-based generator behind mcd_nuts: Random123's published Philox4x32-10 vectors"""
    import nuts_ref
    assert nuts_ref.philox4x32_10((0, 0, 0, 0), (0, 0)) == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    assert nuts_ref.philox4x32_10((0xFFFFFFFF,) * 4, (0xFFFFFFFF,) * 2) == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    assert nuts_ref.philox4x32_10((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0)) == [
        0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]
    u = [nuts_ref.uniform(5, 1, 2, d) for d in range(2000)]
    assert 0 < min(u) and max(u) < 1 and abs(sum(u) / len(u) - 0.5) < 0.03


def test_constraint_validation_and_pruning():
    """loadConstraints (Constraint.hs:275-374): bogus constraints are errors, a vacuous one is dropped, conflicts are
    errors, duplicates and redundant constraints are pruned"""
    #            0
    #       1          6
    #    2     5(e)  7(f)  8
    #  3(a) 4(b)         9(c) 10(d)      pre-order; internal: 0, 1, 2, 6, 8
    t = tree.parse_newick("(((a:1,b:1):1,e:2):1,(f:1,(c:1,d:1):1):1);")
    parent, c0, c1, names, lens = tree.flatten_preorder(t)
    assert parent.tolist() == [-1, 0, 1, 2, 2, 1, 0, 6, 6, 8, 8]
    hdr = "Name,YoungerLeafA,YoungerLeafB,OlderLeafA,OlderLeafB,ProbabilityMass\n"
    load = lambda body, **kw: prepare.load_constraints(hdr + body, parent, names, **kw)
    # plain: (a,b) younger than (c,d)
    c = load("c1,a,b,c,d,0.025\n")
    assert c["young"].tolist() == [2] and c["old"].tolist() == [8] and c["p"].tolist() == [0.025]
    # both nodes equal / younger node is an ancestor of the older one: errors
    with pytest.raises(ValueError, match="both nodes are equal"):
        load("bad,a,b,a,b,0.025\n")
    with pytest.raises(ValueError, match="younger node is direct ancestor"):
        load("bad,a,e,a,b,0.025\n")
    # older node is an ancestor of the younger one: vacuous -> dropped with a warning (or an error on request)
    msgs = []
    c = load("vac,a,b,a,e,0.025\nc1,a,b,c,d,0.025\n", log=msgs.append)
    assert c["names"] == ["c1"] and any("Dropping constraint" in m for m in msgs)
    with pytest.raises(ValueError, match="old node is direct ancestor"):
        load("vac,a,b,a,e,0.025\n", on_problem="error")
    # duplicates: the later one goes
    c = load("c1,a,b,c,d,0.025\nc1again,b,a,d,c,0.1\n")
    assert c["names"] == ["c1"]
    # redundancy: given (a,e) < (c,d), the constraint (a,b) < (f,c) [descendant of the young, ancestor of the old] is implied
    c = load("strong,a,e,c,d,0.025\nweak,a,b,f,c,0.025\n")
    assert c["names"] == ["strong"] and c["young"].tolist() == [1] and c["old"].tolist() == [8]
    c = load("weak,a,b,f,c,0.025\nstrong,a,e,c,d,0.025\n")          # order of the file does not matter
    assert c["names"] == ["strong"]
    # conflict: (a,b) < (c,d) and (f,c) < (a,b): the second one's young node is an ancestor of the first one's old node
    with pytest.raises(ValueError, match="conflicting"):
        load("c1,a,b,c,d,0.025\nc2,f,c,a,b,0.025\n")
    # unrelated constraints all survive, order kept
    c = load("c1,a,b,c,d,0.025\nc5,c,d,a,e,0.05\n")   # (a,b) < (c,d) < (a,e): a chain, neither implied by the other
    assert c["names"] == ["c1", "c5"]
    with pytest.raises(ValueError, match="probabilityMass"):
        load("c1,a,b,c,d,1.5\n")


# ---------------------------------------------------------------------------------------------- MH proposals (host side)
def _small_mh_model():
    import mh_ref
    md, h = synth.synthetic_model(9, seed=5, n_cal=2, n_con=0, n_brace=1)
    X = synth.synthetic_states(md, h, 1)
    parent = [int(p) for p in md.parent]
    braces = [[int(x) for x in md.brace_node[md.brace_off[b]:md.brace_off[b + 1]]] for b in range(md.n_brace)]
    return mh_ref, md, X[0], parent, braces


@pytest.mark.parametrize("kind", range(17))
def test_proposal_jacobians_are_the_determinants_of_the_moves(kind):
    """The reference has no tests for its proposals.  Property check of the restatement (tests/mh_ref.py), which the CUDA
    kernels are compared with: for every proposal kind the Jacobian factor the reference states equals |det| of the move as a
    map (touched coordinates, sampled value) -> (new coordinates, value that samples the reverse move), by central differences."""
    R, md, x, parent, braces = _small_mh_model()
    topo = R.topology(parent)
    child, size, inner, inner_list = topo
    N = len(parent)
    node = {R.SCALE_BRANCH: 3, R.SLIDE_BRACE: 0, R.SLIDE_BRACE_CONTRA: 0, R.SCALE_SCALAR: 4}.get(kind, inner_list[1])
    mult = kind in R.MULT_KINDS
    shift = kind in (R.SLIDE_BRACE, R.SLIDE_BRACE_CONTRA, R.PULLEY)
    OH, OM, OV = 3, 3 + N, 4 + N
    # the coordinate whose new value is sampled directly, its current value, and a sampled value inside the bounds
    if kind in (R.SLIDE_NODE, R.SLIDE_NODE_CONTRA, R.SCALE_SUBTREE, R.SCALE_SUBTREE_CONTRA):
        own, cur = OH + node, x[OH + node]
        forced = cur * 1.01
    elif kind == R.SLIDE_ROOT_CONTRA:
        own, cur = 2, x[2]
        forced = cur * 1.02
    elif kind == R.SCALE_RATES_TREE_CONTRA:
        c = max(child[0], key=lambda i: x[OH + i])
        own, cur = OH + c, x[OH + c]
        forced = cur * 0.99
    else:
        own, cur, forced = None, None, (1.07 if mult else 1e-4)

    def move(z):
        """z = (touched coordinates [own first], sampled value) -> (new coordinates [own first], reverse sample)"""
        xx = x.copy()
        xx[idx] = z[:-1]
        R.FORCED = z[-1]
        try:
            y, lqj, _ = R.propose(xx, parent, topo, braces, kind, node, 0.5, 1.0, 1, 0, 0)
        finally:
            R.FORCED = None
        assert y is not None
        rev = 1.0 / z[-1] if mult else (-z[-1] if shift else xx[own])
        return np.concatenate([y[idx], [rev]]), y

    # touched coordinates of this move
    R.FORCED = forced
    try:
        y0, lqj0, _ = R.propose(x, parent, topo, braces, kind, node, 0.5, 1.0, 1, 0, 0)
        lq_only = None
    finally:
        R.FORCED = None
    assert y0 is not None
    idx = [int(i) for i in np.nonzero(y0 != x)[0]]
    if own is not None:
        idx = [own] + [i for i in idx if i != own]
    z0 = np.concatenate([x[idx], [forced]])
    n = len(z0)
    J = np.zeros((n, n))
    for j in range(n):
        e = 1e-6 * max(1.0, abs(z0[j]))
        zp, zm = z0.copy(), z0.copy()
        zp[j] += e
        zm[j] -= e
        J[:, j] = (move(zp)[0] - move(zm)[0]) / (2 * e)
    if own is not None:
        # the sampled value IS the new value of `own`: the move is (own, rest, new) -> (new, rest', own)
        assert abs(move(z0)[0][0] - forced) < 1e-15
    _, logdet = np.linalg.slogdet(J)
    # ln(q |J|) minus the Hastings factor q = the Jacobian the reference states
    R.FORCED = forced
    try:
        if mult:
            kk, th = 0.5 / 1.0, 1.0 / 0.5
            lnq = ((kk - 1.0) * np.log(1.0 / forced) - (1.0 / forced) / th) - ((kk - 1.0) * np.log(forced) - forced / th)
        else:
            # recompute q alone from the truncated normal the move used: ln(q |J|) - ln|J| must be antisymmetric; take it
            # from the restatement by proposing with a Jacobian-free twin where one exists, else from the bounds
            lnq = None
    finally:
        R.FORCED = None
    if kind == R.SCALE_VAR_TREE:
        # Reference quirk, preserved: scaleVarianceAndTree states n ln(u - u/n + 1/n) (Unconstrained.hs:308-350), the product
        # of the DIAGONAL of the move's Jacobian matrix (every rate also moves with the sample mean); the determinant is
        # u^(n-1).  Parity is with the reference, so the restatement and the kernel keep the stated factor.
        stated = lqj0 - lnq
        assert abs(stated - np.sum(np.log(np.abs(np.diag(J))))) < 1e-6
        assert abs(logdet - (N - 2) * np.log(forced)) < 1e-6 and abs(stated - logdet) > 1e-4
    elif lnq is not None:
        assert abs((lqj0 - lnq) - logdet) < 1e-6, (kind, lqj0 - lnq, logdet)
    else:
        # truncated-normal moves: the reverse move y -> x (sampling the old value) has ln(q |J|) = -(forward), and its |J|
        # is the inverse determinant; so forward + reverse = 0 and forward - reverse = 2 (ln q + ln |J|)
        R.FORCED = -forced if shift else cur
        try:
            xb, lqj_rev, _ = R.propose(y0, parent, topo, braces, kind, node, 0.5, 1.0, 1, 0, 0)
        finally:
            R.FORCED = None
        assert xb is not None and np.abs(xb - x).max() < 1e-12          # the reverse move exists and undoes the move
        assert abs(lqj0 + lqj_rev) < 1e-9                               # detailed balance of q |J|
        # |J| alone: q = z(x) / z(x') depends on the two sampled values and the bounds only; compare determinants
        # through a second, independent route: ln|J| = ln(q |J|) - ln q with ln q from truncated_normal_sample
        # evaluated on the same bounds (Jacobian-free kinds have ln|J| = 0)
        R.FORCED = forced
        try:
            R.propose(x, parent, topo, braces, kind, node, 0.5, 1.0, 1, 0, 0)
        finally:
            R.FORCED = None
        if kind == R.SLIDE_ROOT_CONTRA:
            # Second reference quirk, preserved: slideRootContrarily uses -n ln u with n = nInnerNodes INCLUDING the root
            # (Contrary.hs:173-189, 246-267) although n - 1 relative heights are divided by u: stated = determinant - ln u.
            assert abs(R.LAST[1] - (logdet - np.log(forced / cur))) < 1e-6
        else:
            assert abs(R.LAST[1] - logdet) < 1e-6, (kind, R.LAST[1], logdet)  # the stated Jacobian is the determinant


def test_reference_cycle_mirrors_definitions():
    """app/Definitions.hs:125-279 on the 24-leaves-braces data set: which proposals exist, their weights and lifts"""
    from mcmc_date_b200 import binding as B, mh_cycle
    md, z = load_fixture("24-leaves-braces")
    cyc = mh_cycle.reference_cycle(md)
    N = md.n_nodes
    w = int(np.floor(np.log(N) / np.log(1.3)))
    assert mh_cycle.weight_n_branches(N) == w == 14
    kinds = [c[0] for c in cyc]
    n_inner = (N - 1) // 2 - 1                                   # inner nodes below the root
    assert kinds.count(B.MH_SLIDE_NODE) == n_inner and kinds.count(B.MH_SCALE_SUBTREE) == n_inner
    assert kinds.count(B.MH_SLIDE_NODE_CONTRA) == n_inner and kinds.count(B.MH_SCALE_SUBTREE_CONTRA) == n_inner
    assert kinds.count(B.MH_SCALE_BRANCH) == N - 1 and kinds.count(B.MH_SCALE_RATE_SUBTREE) == n_inner
    assert kinds.count(B.MH_SLIDE_BRACE) == md.n_brace == 1 and kinds.count(B.MH_SLIDE_BRACE_CONTRA) == 1
    assert kinds.count(B.MH_SCALE_SCALAR) == 5                   # lambda, mu, m, v and (calibrations) H
    assert kinds.count(B.MH_PULLEY) == 1 and kinds.count(B.MH_SLIDE_ROOT_CONTRA) == 1
    by = {}
    for c in cyc:
        by.setdefault(c[0], []).append(c)
    assert all(c[5] == 5 for c in by[B.MH_SLIDE_NODE]) and all(c[2] == 0.01 for c in by[B.MH_SLIDE_NODE])
    assert all(3 <= c[5] <= 8 for c in by[B.MH_SCALE_SUBTREE]) and max(c[5] for c in by[B.MH_SCALE_SUBTREE]) == 8
    assert all(c[5] == w for c in by[B.MH_SCALE_SCALAR]) and by[B.MH_PULLEY][0][5] == 6
    # [R] proposals (children of the root and the global ones) are lifted with jacobianRootBranch, [O] ones are not
    parent = np.asarray(md.parent)
    for c in by[B.MH_SLIDE_NODE] + by[B.MH_SCALE_BRANCH] + by[B.MH_SLIDE_NODE_CONTRA]:
        assert bool(c[4]) == (parent[c[1]] == 0)
    assert all(c[4] == 1 for c in by[B.MH_SCALE_NORM_TREE_CONTRA_M] + by[B.MH_SCALE_VAR_TREE] + by[B.MH_SLIDE_ROOT_CONTRA])
    assert all(c[4] == 0 for c in by[B.MH_SCALE_SCALAR] + by[B.MH_SCALE_H_M_CONTRA] + by[B.MH_SLIDE_BRACE])
    # without calibrations the time-height proposals disappear
    cyc2 = mh_cycle.reference_cycle(md, calibrations_available=False)
    assert len(cyc) - len(cyc2) == 4 and B.MH_SLIDE_ROOT_CONTRA not in [c[0] for c in cyc2]


def test_mc3_swap_restatement_keeps_permutations():
    import mh_ref as R
    rng = np.random.default_rng(3)
    C, G = 5, 7
    stats = rng.normal(size=(C * G, 2)) * 5
    slot, cos = np.arange(C * G) % C, np.arange(C * G)
    ladder = 1.0 / (1.0 + 0.5 * np.arange(C))
    tot = 0
    for it in range(30):
        tot += int(R.mc3_swap(stats, slot, cos, ladder, ladder, C, -1, 11, it).sum())
        assert (np.sort(slot.reshape(G, C), axis=1) == np.arange(C)).all()
        assert all(slot[cos[g * C + p]] == p for g in range(G) for p in range(C))
    assert 0 < tot < 30 * G


@pytest.mark.parametrize("kind", range(17))
def test_proposals_touch_the_state_entries_their_lenses_name(kind):
    """Which entries of the state a proposal may change is fixed by the lens it is lifted through in app/Definitions.hs:145-279
    (e.g. `ratesTimeTreeL = tripleLens timeBirthRate rateMean timeTree`, :236-237, for scaleRatesAndTreeContrarily -- the
    PFunction's local name `mu` is the MEAN RATE, lib/Mcmc/Tree/Proposal/Contrary.hs:435-436, not the death rate).  The host
    restatement the CUDA kernels are compared with step by step must change exactly entries inside that set, and must change the
    scalars the lens names."""
    R, md, x, parent, braces = _small_mh_model()
    topo = R.topology(parent)
    child, size, inner, inner_list = topo
    N = len(parent)
    LA, MU, H, OH, OM, OV, OR = 0, 1, 2, 3, 3 + N, 4 + N, 5 + N
    heights = set(range(OH + 1, OH + N))       # the root's relative height is never touched
    rates = set(range(OR + 1, OR + N))         # the stem of the rate tree is never touched
    table = {   # kind: (lens of app/Definitions.hs, entries it may change, scalars it must change)
        R.SLIDE_NODE: ("timeTree", heights, set()), R.SCALE_SUBTREE: ("timeTree", heights, set()),
        R.PULLEY: ("timeTree", heights, set()), R.SLIDE_BRACE: ("timeTree", heights, set()),
        R.SCALE_BRANCH: ("rateTree", rates, set()), R.SCALE_RATE_SUBTREE: ("rateTree", rates, set()),
        R.SCALE_NORM_TREE_CONTRA_M: ("rateMeanRateTreeL", rates | {OM}, {OM}),
        R.SCALE_NORM_TREE_CONTRA_H: ("timeHeightRateTreeL", rates | {H}, {H}),
        R.SCALE_VAR_TREE: ("rateVarianceRateTreeL", rates | {OV}, {OV}),
        R.SCALE_VAR_TREE_AUTO: ("rateMeanVarianceTreeL", rates | {OV}, {OV}),          # the mean is used but unchanged
        R.SLIDE_NODE_CONTRA: ("timeRateTreesL", heights | rates, set()),
        R.SCALE_SUBTREE_CONTRA: ("timeRateTreesL", heights | rates, set()),
        R.SLIDE_BRACE_CONTRA: ("timeRateTreesL", heights | rates, set()),
        R.SLIDE_ROOT_CONTRA: ("heightTimeRateTreesLens", heights | rates | {H}, {H}),
        R.SCALE_RATES_TREE_CONTRA: ("ratesTimeTreeL", heights | {LA, OM}, {LA, OM}),
        R.SCALE_H_M_CONTRA: ("timeHeightRateMeanL", {H, OM}, {H, OM}),
    }
    cases = [(kind, inner_list[1])] if kind != R.SCALE_SCALAR else [(kind, j) for j in range(5)]
    for kd, node in cases:
        if kd == R.SCALE_SCALAR:
            off = [LA, MU, H, OM, OV][node]
            may, must = {off}, {off}
        else:
            _, may, must = table[kd]
            node = {R.SCALE_BRANCH: 3, R.SLIDE_BRACE: 0, R.SLIDE_BRACE_CONTRA: 0}.get(kd, node)
        y, lqj, _ = R.propose(x.copy(), parent, topo, braces, kd, node, 0.5 if kd not in R.MULT_KINDS else 20.0, 1.0, 5, 0, 3)
        assert y is not None
        changed = set(np.nonzero(y != x)[0].tolist())
        assert changed and changed <= may, (kd, sorted(changed - may))
        assert must <= changed, (kd, sorted(must - changed))


def _leafdist(nd, acc=None, path=0.0):
    """leaf -> distance from the root of a nested-dict tree"""
    acc = {} if acc is None else acc
    if not nd["children"]:
        acc[nd["name"]] = path
    for c in nd["children"]:
        _leafdist(c, acc, path + (c["length"] or 0.0))
    return acc


def _pairwise(nd):
    """all leaf-to-leaf path lengths (an invariant of the unrooted tree)"""
    out = {}

    def rec(n):
        if not n["children"]:
            return {n["name"]: 0.0}
        sides = []
        for c in n["children"]:
            sides.append({k: v + (c["length"] or 0.0) for k, v in rec(c).items()})
        for i in range(len(sides)):
            for j in range(i + 1, len(sides)):
                for a, da in sides[i].items():
                    for b, db in sides[j].items():
                        out[frozenset((a, b))] = da + db
        merged = {}
        for s_ in sides:
            merged.update(s_)
        return merged

    rec(nd)
    return out


def test_outgroup_rerooting():
    """prepare.outgroup (elynx `outgroup`, app/Main.hs:179-180): the root moves to the branch that separates the outgroup, the
    unrooted tree (all leaf-to-leaf distances) is unchanged, the split branch is halved, and the sub-tree order follows the
    documented `descend` rule"""
    t = tree.parse_newick("((a:1,b:2)x:1,(c:3,(d:4,e:5)y:2)z:1)r;")
    same = prepare.outgroup({"a", "b"}, t)
    assert same is t                                              # already rooted there: unchanged (`roots t` starts with t)
    assert prepare.outgroup({"c", "d", "e"}, t) is t              # either side of the bipartition
    r = prepare.outgroup({"d", "e"}, t)
    assert [prepare._leaf_set(c) for c in r["children"]] == [frozenset("abc"), frozenset("de")]
    assert r["children"][0]["length"] == 1.0 and r["children"][1]["length"] == 1.0       # y's branch of 2 split in halves
    assert r["name"] == "r" and r["children"][0]["name"] == "z" and r["children"][1]["name"] == "y"
    # upside-down part: [former up-going part with the two root branches joined (1 + 1), then the former sibling c]
    up = r["children"][0]["children"]
    assert up[0]["name"] == "x" and up[0]["length"] == 2.0 and up[1]["name"] == "c" and up[1]["length"] == 3.0
    pw0, pw1 = _pairwise(t), _pairwise(r)
    assert pw0.keys() == pw1.keys() and all(abs(pw0[k] - pw1[k]) < 1e-12 for k in pw0)
    # a single leaf as outgroup, two levels down
    r2 = prepare.outgroup({"e"}, t)
    assert [sorted(prepare._leaf_set(c)) for c in r2["children"]] == [["a", "b", "c", "d"], ["e"]]
    pw2 = _pairwise(r2)
    assert all(abs(pw0[k] - pw2[k]) < 1e-12 for k in pw0)
    assert r2["children"][1]["length"] == 2.5 and r2["children"][0]["children"][1]["name"] == "d"
    # multifurcating (unrooted) input: a bifurcating root is introduced on the leftmost branch first
    u = tree.parse_newick("(a:1,b:2,(c:3,d:4):5);")
    ru = prepare.outgroup({"c", "d"}, u)
    assert [sorted(prepare._leaf_set(c)) for c in ru["children"]] in ([["a", "b"], ["c", "d"]], [["c", "d"], ["a", "b"]])
    pu0, pu1 = _pairwise(u), _pairwise(ru)
    assert all(abs(pu0[k] - pu1[k]) < 1e-12 for k in pu0)
    parent, c0, c1, names, lens = tree.flatten_preorder(ru)      # bifurcating everywhere now
    assert len(parent) == 7
    with pytest.raises(ValueError):
        prepare.outgroup({"a", "c"}, t)                           # not monophyletic on either side
    with pytest.raises(ValueError):
        prepare.outgroup(set(), t)


def test_prepare_reroots_the_tree_list_and_checks_topologies():
    rng = np.random.default_rng(3)
    L = rng.uniform(0.5, 2.0, size=(24, 8))
    fmt = "((a:{0},b:{1}):{2},(c:{3},(d:{4},e:{5}):{6}):{7});"
    trees_ = [fmt.format(*row) for row in L]
    rooted = "((d:1,e:1):1,((a:1,b:1):1,c:1):1);"              # same topology, rooted on the branch above (d, e)
    pr = prepare.prepare_from_treelist("\n".join(trees_), rooted_tree_text=rooted)
    names = pr["names"]
    assert [names[i] for i in range(len(names)) if names[i]] == ["a", "b", "c", "d", "e"]
    assert np.nonzero(pr["parent"] == 0)[0].tolist() == [1, 6]          # ((a,b),c) first, (d,e) second
    kept = L[len(L) // 6:]
    assert len(pr["mean"]) == 7 and pr["mean"][0] == pytest.approx(kept[:, 6].mean())    # the split branch: both halves summed
    assert pr["mean"][1] == pytest.approx((kept[:, 2] + kept[:, 7]).mean())             # the two old root branches joined
    assert np.allclose(pr["precision"] @ pr["cov"], np.eye(7), atol=1e-8)
    with pytest.raises(ValueError, match="single topology"):
        prepare.prepare_from_treelist("\n".join(trees_), rooted_tree_text="((d:1,e:1):1,((a:1,c:1):1,b:1):1);")
    with pytest.raises(ValueError, match="duplicate leaves"):
        prepare.prepare_from_treelist("((a:1,a:2):1,(c:3,(d:4,e:5):2):1);\n" * 6)
    with pytest.raises(ValueError, match="equal sub tree orders"):
        prepare.prepare_from_treelist("\n".join(trees_[:12] + ["((b:2,a:1):1,(c:3,(d:4,e:5):2):1);"] * 6))


def test_calibrations_from_an_mcmctree_labelled_tree():
    """loadCalibrationsFromTree (lib/Mcmc/Tree/Prior/Node/CalibrationFromTree.hs): quoted labels, L / U / B bounds, default
    probability mass 0.01, leftmost / rightmost leaf naming, pre-order of the labelled tree"""
    mean_tree = tree.parse_newick("((((h:1,(c:1,b:1):1):1,g:1):1,(o:1,s:1):1):1,x:1);")
    parent, c0, c1, names, _ = tree.flatten_preorder(mean_tree)
    text = "((((h, (c, b)) 'B(6,8,2.5e-2,2.5e-2)', g) 'L(9)', (o, s)) 'B(12,16)', x) 'U(100,2.5e-2)';"
    cal = prepare.load_calibrations_from_tree(text, parent, names)
    assert cal["names"] == ["h-x", "h-s", "h-g", "h-b"] and cal["node"].tolist() == [0, 1, 2, 3]
    assert cal["lo"].tolist() == [0.0, 12.0, 9.0, 6.0] and cal["hi"].tolist() == [100.0, 16.0, np.inf, 8.0]
    assert cal["lo_p"].tolist() == [0.5, 0.01, 0.01, 0.025] and cal["hi_p"].tolist() == [0.025, 0.01, 0.5, 0.025]
    assert prepare.mean_root_height(cal) == 50.0                      # getMeanRootHeight: b / 2 without a lower bound
    # attoparsec's `double` needs a leading digit: MCMCtree's 'B(.06,.08)' is not a calibration for McmcDate
    assert prepare._parse_mcmctree_label("B(.06,.08)") is None and prepare._parse_mcmctree_label("U(1.0)") == (None, None, 1.0, 0.01)
    with pytest.raises(ValueError, match="no calibrations"):
        prepare.load_calibrations_from_tree("((h,(c,b)),x);", parent, names)
    with pytest.raises(ValueError, match="Lower boundary larger equal"):
        prepare.load_calibrations_from_tree("((((h,(c,b))'B(8,6)',g),(o,s)),x);", parent, names)
    # the reference's own file for the 7-taxon set is what the fixture was built from
    _, z7 = load_fixture("mtcdnapri-7-leaves")
    assert z7["cal_node"].tolist() == [0, 1, 3] and z7["cal_hi"].tolist() == [100.0, 16.0, 8.0] and z7["cal_lo"].tolist() == [0.0, 12.0, 6.0]
    assert z7["cal_hi_p"].tolist() == [0.025] * 3


def test_auto_tuner_and_cycle_order():
    from mcmc_date_b200 import mh_cycle
    md, _ = load_fixture("mtcdnapri-7-leaves")
    cyc = mh_cycle.reference_cycle(md)
    assert sum(e[5] for e in cyc) == 253                                  # proposal steps per iteration of the 7-taxon cycle
    assert [mh_cycle.optimal_rate(d) for d in (1, 2, 3, 4, 5, 50)] == pytest.approx([0.44, 0.3885, 0.337, 0.2855, 0.234, 0.234])
    dims = {e[:2]: mh_cycle.proposal_dimension(md, e) for e in cyc}
    from mcmc_date_b200 import binding as B
    assert dims[(B.MH_SCALE_RATES_TREE_CONTRA, 0)] == 7 and dims[(B.MH_SLIDE_ROOT_CONTRA, 0)] == 9 and dims[(B.MH_PULLEY, 0)] == 5 \
        if (B.MH_PULLEY, 0) in dims else True
    assert dims[(B.MH_SCALE_NORM_TREE_CONTRA_M, 0)] == 13 and dims[(B.MH_SCALE_H_M_CONTRA, 0)] == 2
    lst, idx = mh_cycle.random_order(cyc, np.random.default_rng(0))
    assert len(lst) == 253 and np.bincount(idx, minlength=len(cyc)).tolist() == [e[5] for e in cyc] and all(e[5] == 1 for e in lst)
    acc = np.array([0.9 * e[5] for e in cyc])
    tuned = mh_cycle.auto_tune(md, cyc, acc, np.array([float(e[5]) for e in cyc]))
    assert all(t[3] > c[3] for t, c in zip(tuned, cyc))                   # rates above every target: larger steps
    assert tuned[0][3] == pytest.approx(np.exp(2 * (0.9 - 0.44)))


def test_prepare_on_every_reference_data_set_with_its_rooted_tree():
    """all six analysis.conf files of the reference's tests/: the tree list re-rooted at the outgroup of the configured rooted tree
    passes the topology checks and gives what the un-rooted call gives (the lists are already rooted there); the MCMCtree-labelled
    calibration tree of 06-leaves-constant-rate (the reference's own use of `calibrations="data/calibrations.tree"`) loads to the
    same table as its CSV twin"""
    base = "/root/reference/tests"
    if not os.path.isdir(base):
        pytest.skip("reference checkout not present (GPU box)")
    seen = 0
    for d in sorted(os.listdir(base)):
        conf = open(os.path.join(base, d, "analysis.conf")).read()
        get = lambda k: re.search(rf'^{k}="([^"]+)"', conf, re.M).group(1)
        rooted = open(os.path.join(base, d, get("rooted_tree"))).read()
        tl = open(os.path.join(base, d, get("trees"))).read()
        a = prepare.prepare_from_treelist(tl, rooted_tree_text=rooted)
        b = prepare.prepare_from_treelist(tl)
        assert np.array_equal(a["parent"], b["parent"]) and np.array_equal(a["mean"], b["mean"]) and a["names"] == b["names"]
        seen += 1
    assert seen == 6
    d = os.path.join(base, "06-leaves-constant-rate", "data")
    pr = prepare.prepare_from_treelist(open(os.path.join(d, "test.treelist")).read(), rooted_tree_text=open(os.path.join(d, "time.tree")).read())
    c_csv = prepare.load_calibrations(open(os.path.join(d, "calibrations.csv")).read(), pr["parent"], pr["names"])
    c_tree = prepare.load_calibrations_from_tree(open(os.path.join(d, "calibrations.tree")).read(), pr["parent"], pr["names"])
    for k in ("node", "lo", "lo_p", "hi", "hi_p"):
        assert np.array_equal(c_csv[k], c_tree[k]), k
    assert c_tree["names"] == ["a-f"] and prepare.mean_root_height(c_tree) == 1.0
