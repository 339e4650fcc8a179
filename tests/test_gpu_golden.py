"""The reference's golden vectors replayed against the CUDA path:
`constriction_b200.stream` = the reference's Python API surface on top of the C ABI kernels."""
import pytest

import golden_runner as R
from golden import reference_vectors as G

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def api():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import constriction_b200.stream as S
    return S


@pytest.mark.parametrize("case", G.ENCODE_CASES, ids=[c["id"] for c in G.ENCODE_CASES])
def test_encode_golden(api, case):
    R.run_encode_case(api, case)


@pytest.mark.parametrize("case", G.DECODE_CASES, ids=[c[0] for c in G.DECODE_CASES])
def test_decode_golden(api, case):
    R.run_decode_case(api, case)


def test_seal(api):
    R.run_seal_case(api)


def test_lengths(api):
    R.run_length_cases(api)


def test_misc(api):
    R.run_misc_cases(api)


def test_mixed_models_in_one_stream(api, oracle):
    """The reference's `compress_many` pattern (stack.rs:1293-1454, queue.rs:1175-1330): one coder, alternating
    blocks of symbols under different model kinds -- i.i.d. categorical, i.i.d. Gaussian, per-symbol Gaussian
    parameters (the table-free kernels), per-symbol categorical rows, uniform -- then decoded back; every
    intermediate `pos()` and the final words must equal the oracle's."""
    import numpy as np
    O = oracle
    rng = np.random.default_rng(2024)
    blocks = []
    for rep in range(3):
        pmf = rng.dirichlet(np.ones(17)).astype(np.float64)
        blocks.append(("cat", rng.choice(17, size=300, p=pmf).astype(np.int32), pmf))
        blocks.append(("gauss", np.clip(np.rint(rng.normal(2.0, 7.0, 257)), -30, 30).astype(np.int32), (2.0, 7.0)))
        m, s = rng.normal(0, 10, 200), np.exp(rng.uniform(-1, 3, 200))
        blocks.append(("gauss_params", np.clip(np.rint(rng.normal(m, s)), -30, 30).astype(np.int32), (m, s)))
        rows = rng.dirichlet(np.ones(9), size=64).astype(np.float32)
        blocks.append(("cat_rows", np.array([rng.choice(9, p=r / r.sum()) for r in rows], dtype=np.int32), rows))
        blocks.append(("uniform", rng.integers(0, 11, size=33).astype(np.int32), 11))

    def models(mod):
        out = []
        for kind, syms, par in blocks:
            if kind == "cat":
                out.append((mod.Categorical(par, perfect=False), ()))
            elif kind == "gauss":
                out.append((mod.QuantizedGaussian(-30, 30, *par), ()))
            elif kind == "gauss_params":
                out.append((mod.QuantizedGaussian(-30, 30), par))
            elif kind == "cat_rows":
                out.append((mod.Categorical(perfect=False), (par,)))
            else:
                out.append((mod.Uniform(par), ()))
        return out

    ours, theirs = models(api.model), models(O)
    # stack: encode the blocks last to first, decode first to last
    a, oa = api.stack.AnsCoder(), O.AnsCoder()
    for (kind, syms, _), (m1, p1), (m2, p2) in reversed(list(zip(blocks, ours, theirs))):
        a.encode_reverse(syms, m1, *p1)
        oa.encode_reverse(syms, m2, *p2)
        assert a.pos() == oa.pos(), kind
    assert np.array_equal(a.get_compressed(), oa.get_compressed())
    for (kind, syms, _), (m1, p1) in zip(blocks, ours):
        got = a.decode(m1, *p1) if p1 else a.decode(m1, syms.size)
        assert np.array_equal(got, syms), kind
    assert a.is_empty()
    # queue
    r, orr = api.queue.RangeEncoder(), O.RangeEncoder()
    for (kind, syms, _), (m1, p1), (m2, p2) in zip(blocks, ours, theirs):
        r.encode(syms, m1, *p1)
        orr.encode(syms, m2, *p2)
        assert r.pos() == orr.pos(), kind
    assert np.array_equal(r.get_compressed(), orr.get_compressed())
    d = api.queue.RangeDecoder(r.get_compressed())
    for (kind, syms, _), (m1, p1) in zip(blocks, ours):
        got = d.decode(m1, *p1) if p1 else d.decode(m1, syms.size)
        assert np.array_equal(got, syms), kind
    assert d.maybe_exhausted()
