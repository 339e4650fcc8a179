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
