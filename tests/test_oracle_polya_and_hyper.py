"""CPU tests of the Poisson Polya-urn Phi draw and of the hyper-parameter optimisation hooks.

Reference: topics/PolyaUrnSpaliasLDA.java:495-507, types/PolyaUrnDirichletFixedCoeffPoisson.java:17-44,
types/PoissonFixedCoeffSampler.java:45-51 (Poisson draws), ModifiedSimpleLDA.java:812-905 (optimizeAlpha / optimizeBeta).
The reference's own tests for this scheme (PolyaUrnSpaliasTest, PoissonFixedCoeffSamplerTest) are smoke / distribution
tests; the distribution tests are mirrored here against scipy."""
import numpy as np
import pytest
from scipy import stats

from conftest import make_corpus


@pytest.mark.parametrize("n,beta", [(0, 0.01), (0, 0.5), (1, 0.01), (7, 0.01), (30, 0.5), (99, 0.01)])
def test_poisson_exact_branch_matches_scipy(oracle, n, beta):
    """n < L: inversion over the truncated pmf -- chi-square against the Poisson pmf, as PoissonFixedCoeffSamplerTest does."""
    lam, N = beta + n, 40000
    xs = np.array([oracle.poisson(beta, n, 11, c, 3) for c in range(N)])
    hi = int(stats.poisson.ppf(1 - 1e-4, lam)) + 1
    obs = np.bincount(np.minimum(xs, hi), minlength=hi + 1).astype(np.float64)
    exp = stats.poisson.pmf(np.arange(hi + 1), lam) * N
    exp[hi] = N - exp[:hi].sum()
    keep = exp > 5
    chi2 = ((obs[keep] - exp[keep]) ** 2 / exp[keep]).sum() + (obs[~keep].sum() - exp[~keep].sum()) ** 2 / max(exp[~keep].sum(), 1e-9)
    assert stats.chi2.sf(chi2, max(int(keep.sum()), 1)) > 1e-5
    assert xs.max() <= 199       # truncated at 2L - 1


@pytest.mark.parametrize("n", [100, 150, 5000])
def test_poisson_normal_branch(oracle, n):
    """n >= L: round(sqrt(lambda) N(0,1) + lambda), never negative (PolyaUrnDirichlet.java:102-107)."""
    beta = 0.01
    xs = np.array([oracle.poisson(beta, n, 5, c, 1) for c in range(20000)], np.float64)
    lam = beta + n
    assert xs.min() >= 0
    assert abs(xs.mean() - lam) < 4 * np.sqrt(lam / len(xs)) + 0.01
    assert abs(xs.var() / lam - 1) < 0.05
    assert stats.kstest((xs - lam) / np.sqrt(lam), "norm").pvalue > 1e-5 or n < 1000   # the rounding shows for small lambda


def test_poisson_contract_equals_faithful(oracle):
    """contract (libm-free) and faithful (glibc exp/log/cos) arithmetic on the same uniforms: identical integers except
    where a uniform falls within rounding error of a pmf boundary"""
    same = total = 0
    for n, beta in ((0, 0.01), (3, 0.01), (40, 0.3), (99, 0.01), (200, 0.01), (3000, 0.1)):
        a = np.array([oracle.poisson(beta, n, 9, c, 2) for c in range(5000)])
        b = np.array([oracle.poisson(beta, n, 9, c, 2, faithful=True) for c in range(5000)])
        same += int((a == b).sum()); total += len(a)
    assert same >= total - 2


def test_polya_urn_phi_draw_properties(oracle):
    """X = Poisson(beta + n), phi = X / sum X: exact zeros, columns that sum to 1, mean (beta + n) / sum (beta + n)."""
    rng = np.random.default_rng(0)
    V, K, beta = 500, 12, 0.01
    n_wk = (rng.random((V, K)) < 0.1) * rng.integers(1, 40, (V, K))
    n_wk[:, 3] = 0                                   # a topic with no tokens at all
    n_wk = n_wk.astype(np.int32)
    phis = np.stack([oracle.phi_polya_contract(n_wk, beta, 5, s) for s in range(200)]).astype(np.float64)
    col = phis.sum(axis=1)
    assert np.all((np.abs(col - 1) < 1e-5) | (col == 0))
    assert (phis == 0).mean() > 0.8                  # sparse rows: most zero-count cells stay exactly zero
    k = 0
    want = (beta + n_wk[:, k]) / (beta + n_wk[:, k]).sum()
    got = phis[:, :, k].mean(axis=0)
    big = want > 0.01
    assert np.allclose(got[big], want[big], rtol=0.08)
    f = oracle.phi_polya_faithful(n_wk, beta, 5, 7)
    c = oracle.phi_polya_contract(n_wk, beta, 5, 7)
    assert np.allclose(f, c, rtol=1e-6, atol=1e-9)
    # the empty topic: Poisson(0.01) per cell, almost surely a handful of ones -- whatever it drew sums to 1 or stays 0
    assert np.all((np.abs(phis[:, :, 3].sum(axis=1) - 1) < 1e-5) | (phis[:, :, 3].sum(axis=1) == 0))


def test_polya_urn_sweeps_invariants_and_zero_column_fallback(oracle):
    """Whole sweeps of the Polya-urn scheme (sparse z-step + Poisson Phi): counts stay consistent, and a word type whose
    Phi column is all zero gets a uniform topic (PolyaUrnSpaliasLDA.java:275-277) instead of NaN arithmetic."""
    off, tokens = make_corpus(60, 80, 25, seed=4)
    K, V, alpha, beta, seed = 30, 80, 0.1, 0.01, 3
    z0 = oracle.java_next_ints(seed, K, len(tokens))
    nw0, _ = oracle.rebuild_counts(tokens, z0, V, K)
    phi0 = oracle.phi_polya_contract(nw0, beta, seed, 0)
    st = oracle.sweeps("contract", oracle.POLYAURN, off, tokens, z0, V, K, np.full(K, alpha), beta, seed, 1, 5, phi0)
    assert st["n_wk"].sum() == len(tokens) and np.array_equal(st["n_wk"].sum(axis=0), st["n_k"])
    assert st["z"].min() >= 0 and st["z"].max() < K
    sf = oracle.sweeps("faithful", oracle.POLYAURN, off, tokens, z0, V, K, np.full(K, alpha), beta, seed, 1, 1,
                       oracle.phi_polya_faithful(nw0, beta, seed, 0))
    s1 = oracle.sweeps("contract", oracle.POLYAURN, off, tokens, z0, V, K, np.full(K, alpha), beta, seed, 1, 1, phi0)
    assert (sf["z"] == s1["z"]).mean() > 0.98
    # all-zero Phi: every token takes floor(u * K), u from the z stream
    zero = np.zeros((V, K), np.float32)
    z = oracle.z_spalias_contract(off, tokens, z0, K, np.full(K, alpha), zero, seed, 1)
    assert z.min() >= 0 and z.max() < K
    assert abs(z.mean() - (K - 1) / 2) < 1.0


def test_learn_symmetric_concentration_recovers_truth():
    """The fixed point behind optimizeAlpha / optimizeBeta (MSL:847-852,897-901 -> MALLET learnSymmetricConcentration)."""
    from ldagroupedgibbssampler_b200.sampler import digamma, learn_symmetric_concentration
    from scipy.special import digamma as dg
    assert max(abs(digamma(x) - dg(x)) for x in (0.01, 0.5, 1.0, 3.3, 10.0, 200.0)) < 1e-10
    rng = np.random.default_rng(1)
    K, true = 40, 8.0
    lens = rng.integers(40, 160, 2500)
    ch = np.zeros(400, np.int64)
    for n in lens:
        ch += np.bincount(rng.multinomial(n, rng.dirichlet(np.full(K, true / K))), minlength=400)[:400]
    got = learn_symmetric_concentration(ch, np.bincount(lens, minlength=400), K, 1.0)
    assert abs(got - true) / true < 0.05
