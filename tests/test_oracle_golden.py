"""Pins the CPU oracle against the reference's own golden vectors."""
import pytest

import golden_runner as R
from golden import reference_vectors as G


@pytest.mark.parametrize("case", G.ENCODE_CASES, ids=[c["id"] for c in G.ENCODE_CASES])
def test_encode_golden(oracle, case):
    R.run_encode_case(oracle, case)


@pytest.mark.parametrize("case", G.DECODE_CASES, ids=[c[0] for c in G.DECODE_CASES])
def test_decode_golden(oracle, case):
    R.run_decode_case(oracle, case)


def test_seal(oracle):
    R.run_seal_case(oracle)


def test_lengths(oracle):
    R.run_length_cases(oracle)


def test_misc(oracle):
    R.run_misc_cases(oracle)


def test_guided_quantile_equals_the_table_search(oracle):
    """quantize.rs:580-779 restated with the reference's control flow (orc_qgauss_quantile_guided, used by the
    single-thread lazy-model CPU baseline) returns the unique bin the table holds, whatever the starting guess."""
    import numpy as np
    rng = np.random.default_rng(11)
    for lo, hi, mean, std in [(-50, 50, 3.2, 9.6), (-50, 50, -70.0, 0.3), (-5, 5, 100.0, 1e-3), (0, 1, 0.4, 2.0),
                              (-1000, 1000, 12.5, 300.0), (-64, 64, 0.0, 1e4)]:
        cdf = oracle.qgauss_cdf(lo, hi, mean, std)
        qs = np.concatenate([rng.integers(0, 1 << 24, size=400), [0, 1, (1 << 24) - 1], cdf[:-1], cdf[1:] - 1])
        for q in qs:
            q = int(q)
            s = int(np.searchsorted(cdf, q, side="right")) - 1
            assert oracle.qgauss_quantile_guided(lo, hi, mean, std, q) == (lo + s, int(cdf[s]), int(cdf[s + 1] - cdf[s]))
    syms = np.clip(np.rint(rng.normal(3.2, 9.6, size=5000)), -50, 50).astype(np.int32)
    cdf = oracle.qgauss_cdf(-50, 50, 3.2, 9.6)
    words = oracle.ans_encode_qgauss_lazy(syms, -50, 50, 3.2, 9.6)
    assert np.array_equal(words, oracle.ans_encode_iid(syms, cdf, -50))
    assert np.array_equal(oracle.ans_decode_qgauss_lazy(words, syms.size, -50, 50, 3.2, 9.6), syms)
