"""The reference's OWN pytest files (tests/python/*.py, 3,320 lines), executed in place against the oracle.

Runs only where the reference tree is mounted (this container; the GPU boxes do not have it, and the CUDA path is
compared with the oracle there instead): every `test_*` function of every file under /root/reference/tests/python
is called with `constriction` resolving to a shim package whose `stream.{stack,queue,model}` modules are the
oracle's restatement of the Python API (oracle/refapi.py).  All their `assert`s -- every golden `uint32` array the
reference holds for this path -- thereby pin the oracle.  Tests of components that are out of scope for this project
(ChainCoder, Huffman / symbol codes) are expected to fail with AttributeError and are reported as xfail."""
import importlib.util
import os
import sys
import types

import pytest

REF_TESTS = "/root/reference/tests/python"
OUT_OF_SCOPE = ("chain", "huffman", "symbol")  # constriction.stream.chain.ChainCoder, constriction.symbol.*


def _shim(api):
    root = types.ModuleType("constriction")
    stream = types.ModuleType("constriction.stream")
    stack, queue, model = (types.ModuleType(f"constriction.stream.{n}") for n in ("stack", "queue", "model"))
    stack.AnsCoder = api.AnsCoder
    queue.RangeEncoder, queue.RangeDecoder = api.RangeEncoder, api.RangeDecoder
    for name in ("QuantizedGaussian", "QuantizedLaplace", "QuantizedCauchy", "Binomial", "Bernoulli", "Categorical", "Uniform",
                 "CustomModel", "ScipyModel"):
        setattr(model, name, getattr(api, name))
    stream.stack, stream.queue, stream.model = stack, queue, model
    root.stream = stream
    return {"constriction": root, "constriction.stream": stream, "constriction.stream.stack": stack,
            "constriction.stream.queue": queue, "constriction.stream.model": model}


def _collect():
    if not os.path.isdir(REF_TESTS):
        return []
    out = []
    for fn in sorted(os.listdir(REF_TESTS)):
        if not (fn.startswith("test_") and fn.endswith(".py")):
            continue
        with open(os.path.join(REF_TESTS, fn)) as f:
            for line in f:
                if line.startswith("def test_"):
                    out.append((fn, line[4:line.index("(")]))
    return out


CASES = _collect()


@pytest.mark.skipif(not CASES, reason="the reference tree is not mounted here")
@pytest.mark.parametrize("fn,name", CASES, ids=[f"{f[:-3]}::{n}" for f, n in CASES])
def test_reference_test_passes_on_the_oracle(oracle, fn, name, monkeypatch):
    pytest.importorskip("scipy")
    for k, v in _shim(oracle).items():
        monkeypatch.setitem(sys.modules, k, v)
    spec = importlib.util.spec_from_file_location(f"_ref_{fn[:-3]}", os.path.join(REF_TESTS, fn))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    try:
        getattr(mod, name)()
    except AttributeError as exc:
        if any(word in str(exc) for word in OUT_OF_SCOPE) or any(word in name for word in OUT_OF_SCOPE):
            pytest.xfail(f"out of scope component: {exc}")
        raise


def test_recorded_trace_replays_on_the_oracle(oracle):
    """The committed fixture (tests/golden/reference_trace.json.gz) is self-consistent: replayed on the oracle it
    recorded from, every call returns what was written down.  (The GPU suite replays it on the CUDA path.)"""
    import trace_replay as T
    traces = T.load()
    assert len(traces) >= 100
    for trace in traces:
        T.replay(oracle, trace)
