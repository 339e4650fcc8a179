"""The reference's own pytest files (tests/python/test_constriction.py, test_docexamples{,_f32}.py, test_lazy_f{32,64}.py)
against the CUDA path.  The reference tree does not exist on the GPU box, so the files travel as DATA: every call
they make on coder / model objects and every value they get back was recorded while they ran (and passed) on the
oracle (tests/golden/make_reference_trace.py, 120 tests, 736 calls); here the same calls are made on
`constriction_b200.stream` and every result -- compressed words, decoded symbols, positions, raised exceptions --
must be identical.  (ChainCoder / Huffman tests are out of scope and were not recorded.)"""
import pytest

import trace_replay as T

pytestmark = pytest.mark.gpu
TRACES = T.load()


@pytest.fixture(scope="module")
def api():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import constriction_b200.stream as S
    return S


@pytest.mark.parametrize("trace", TRACES, ids=[t["test"] for t in TRACES])
def test_reference_test_file_replayed_on_the_cuda_path(api, trace):
    T.replay(api, trace)
