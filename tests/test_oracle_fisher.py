"""The oracle's Fisher tail against two independent implementations (scipy, mpmath)."""
import mpmath as mp
import numpy as np
import pytest
from scipy import stats


def mp_fisher_greater(a, b, c, d):
    mp.mp.dps = 60
    white, black, draws = a + c, b + d, a + b
    tot = mp.binomial(white + black, draws)
    s = mp.mpf(0)
    for x in range(a, min(white, draws) + 1):
        s += mp.binomial(white, x) * mp.binomial(black, draws - x) / tot
    return s


TABLES = [
    (21, 2886, 1, 2906), (21, 2886, 0, 2907), (28, 2501, 1, 2528), (26, 2914, 1, 2939), (1, 9, 1, 9),
    (5, 5, 5, 5), (3, 997, 1, 999), (100, 5900, 2, 5998), (1000, 999000, 167, 999833), (350, 999650, 167, 999833),
    (10000, 990000, 167, 999833), (2, 48, 1, 49), (60, 5940, 1, 5999), (12, 24988, 5, 24995),
]


@pytest.mark.parametrize("t", TABLES)
def test_oracle_vs_mpmath(oracle, t):
    want = mp_fisher_greater(*t)
    got = oracle.fisher(*t)
    if want < mp.mpf("1e-300"):
        assert got < 1e-290
    else:
        assert abs(got - want) / want < 1e-10, (t, got, want)


def test_oracle_vs_scipy_random(oracle):
    rng = np.random.default_rng(7)
    for _ in range(300):
        n = int(rng.integers(10, 50000))
        k = int(rng.integers(1, min(n, 400)))
        e = int(rng.integers(0, min(n, 40)))
        want = stats.fisher_exact([[k, n - k], [e, n - e]], alternative="greater")[1]
        got = oracle.fisher(k, n - k, e, n - e)
        assert got == pytest.approx(want, rel=1e-7, abs=1e-300)


def test_doc_sensitivity_bound(oracle):
    """SURVEY App. C-8 (screenshot juliet_hiv-phasing.png): G99G GGG->GGT, 0.72 % of 2907 reads,
    is the weakest call shown; with the restated table/Bonferroni defaults it must be called."""
    p1 = oracle.fisher(21, 2886, 1, 2906)
    p0 = oracle.fisher(21, 2886, 0, 2907)
    assert p1 == pytest.approx(5.31e-6, rel=5e-3)
    assert p0 == pytest.approx(4.60e-7, rel=5e-3)
    assert p1 * 947 < 0.01          # PR 99 + RT 560 + IN 288 codon positions
    assert oracle.fisher(28, 2501, 1, 2528) == pytest.approx(5.2e-8, rel=2e-2)   # K65R
    assert oracle.fisher(26, 2914, 1, 2939) == pytest.approx(2.0e-7, rel=3e-2)   # T215Y
