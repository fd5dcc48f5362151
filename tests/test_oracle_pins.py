"""Pin the oracle against the reference's own known answers: the
RevBayes-checked doc-comment values of lib/Mcmc/Tree/Prior/BirthDeath.hs:249-271 -- the only
golden numbers the reference holds for this path."""
import math

import numpy as np
import pytest

from mcmc_date_b200 import tree
from oracle import oracle as O

T = "(((a:1.0,b:1.0):1.0,c:2.0):1.0,d:3.0):0.0;"


def _tree():
    parent, c0, c1, names, lens = tree.flatten_preorder(tree.parse_newick(T))
    return c0, c1, lens


@pytest.mark.parametrize("mu,expect", [
    (0.0, -10.09861228866811), (0.01, -10.07675364864067), (0.05, -9.993307032921498),
    (0.1, -9.898174270006024), (0.2, -9.73975910235509), (0.5, -9.54137886890279)])
def test_birth_death_death_rates(mu, expect):  # BirthDeath.hs:264-265
    c0, c1, lens = _tree()
    got = math.log(1 / 3) + O.birth_death(c0, c1, lens, 1.0, mu, 1.0, True)
    assert got == pytest.approx(expect, abs=5e-15)


@pytest.mark.parametrize("rho,expect", [(1.0, -10.09861228866811), (0.9, -9.809211822253452), (0.8, -9.498032504556043)])
def test_birth_death_sampling(rho, expect):  # BirthDeath.hs:267-268
    c0, c1, lens = _tree()
    got = math.log(1 / 3) + O.birth_death(c0, c1, lens, 1.0, 0.0, rho, True)
    assert got == pytest.approx(expect, abs=5e-15)


def test_birth_death_general():  # BirthDeath.hs:270-271
    c0, c1, lens = _tree()
    got = math.log(1 / 3) + O.birth_death(c0, c1, lens, 0.2, 0.5, 0.8, True)
    assert got == pytest.approx(-9.700151607658995, abs=5e-15)


def test_birth_death_single_branch():  # BirthDeath.hs:252-254 (value printed in linear domain)
    got = math.exp(O.birth_death([-1], [-1], [1.0], 1.2, 3.2, 1.0, False))
    assert got == pytest.approx(5.8669248906043234e-2, rel=1e-14)


def test_birth_death_stem_tree():  # BirthDeath.hs:256-258: zero stem -> probability 0 with the origin condition
    parent, c0, c1, names, lens = tree.flatten_preorder(tree.parse_newick("(a:0.4,(b:0.2,c:0.2):0.2):0.0;"))
    assert O.birth_death(c0, c1, lens, 1.2, 3.2, 1.0, False) == -math.inf
    got = math.exp(O.birth_death(c0, c1, lens, 1.2, 3.2, 1.0, True))
    # MRCA conditioning = product of the two root subtrees, each with its stem (BirthDeath.hs:173-175)
    a = O.birth_death([-1], [-1], [0.4], 1.2, 3.2, 1.0, False)
    p2, c02, c12, _, l2 = tree.flatten_preorder(tree.parse_newick("(b:0.2,c:0.2):0.2;"))
    b = O.birth_death(c02, c12, l2, 1.2, 3.2, 1.0, False)
    assert got == pytest.approx(math.exp(a + b), rel=1e-14)


def test_compute_de_limits():
    # computeDENearCritical (BirthDeath.hs:90-114) is a FIRST-ORDER approximation of computeDE in
    # d = la - mu: the two differ by O(|d|) (this is why the CUDA path runs the literal near-critical
    # recursion for |d| < 1e-6 instead of the exact closed form)
    prev = None
    for d in (1e-3, 1e-4, 1e-5):
        D, E = O.compute_de(1.0 + d, 1.0, 1.0, 0.7, 0.2)
        Dn, En = O.compute_de(1.0 + d, 1.0, 1.0, 0.7, 0.2, nearcrit=True)
        rel = abs(D - Dn) / D
        assert rel < d and abs(E - En) < d
        if prev is not None:
            assert rel == pytest.approx(prev / 10, rel=0.05)   # linear in d
        prev = rel
    # E(0) = e0 for rho = 1, D(0) = 1
    D0, E0 = O.compute_de(1.3, 0.4, 1.0, 0.0, 0.25)
    assert D0 == pytest.approx(1.0, rel=1e-14) and E0 == pytest.approx(0.25, rel=1e-14)


def test_digamma_against_mpmath():
    import mpmath as mp
    for x in [1e-3, 0.05, 0.3, 1.0, 1.5, 2.0, 7.7, 10.0, 33.3, 1e3, 1e6]:
        assert O.digamma(x) == pytest.approx(float(mp.digamma(x)), rel=2e-14, abs=2e-14)
