#!/usr/bin/env python
"""Generates tests/golden/reference_trace.json.gz (run in the container that has /root/reference mounted).

Every test function of the reference's own pytest files (/root/reference/tests/python/test_*.py) is executed with
`constriction` resolving to a RECORDING shim around the oracle's restatement of the Python API (oracle/refapi.py): each
constructor and method call on a coder or model object is written down with its exact arguments (arrays as raw bytes)
and its result.  The reference's asserts run against the oracle while recording, so only traces of tests that PASS are
kept.  tests/test_gpu_reference_trace.py replays the traces against the CUDA path (constriction_b200.stream) on the GPU
box, where the reference tree does not exist, and compares every result -- the reference's test-suite, ported as data.

Python callbacks of CustomModel / ScipyModel are recorded as memo tables {(x, *params): cdf value}: the quantiser only
ever evaluates the CDF at the half-integers of the support, in both implementations.
"""
import base64
import gzip
import importlib.util
import json
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
REF_TESTS = "/root/reference/tests/python"
CLASSES = ["AnsCoder", "RangeEncoder", "RangeDecoder", "QuantizedGaussian", "QuantizedLaplace", "QuantizedCauchy", "Binomial",
           "Bernoulli", "Categorical", "Uniform", "CustomModel", "ScipyModel"]
WHERE = {"AnsCoder": "stack", "RangeEncoder": "queue", "RangeDecoder": "queue"}


class Recorder:
    def __init__(self, api):
        self.api, self.events, self.n = api, [], 0
        self.api_types = tuple(getattr(api, c) for c in CLASSES)

    def new_id(self):
        self.n += 1
        return self.n

    def enc(self, v):
        if isinstance(v, Proxy):
            return {"obj": v._id}
        if isinstance(v, np.ndarray):
            a = np.ascontiguousarray(v)
            return {"nd": a.dtype.str, "shape": list(a.shape), "b64": base64.b64encode(a.tobytes()).decode()}
        if isinstance(v, (np.integer,)):
            return {"int": int(v)}
        if isinstance(v, (np.floating,)):
            return {"f64": float(v).hex()}
        if isinstance(v, bool) or v is None or isinstance(v, int) or isinstance(v, str):
            return v
        if isinstance(v, float):
            return {"f64": v.hex()}
        if isinstance(v, (list, tuple)):
            return {"seq": [self.enc(x) for x in v]}
        raise TypeError(f"cannot record {type(v)}")

    def wrap_result(self, r):
        if isinstance(r, self.api_types):
            p = Proxy(self, r, self.new_id())
            return p, {"newobj": p._id}
        return r, self.enc(r)


class Proxy:
    def __init__(self, rec, target, oid):
        object.__setattr__(self, "_rec", rec)
        object.__setattr__(self, "_t", target)
        object.__setattr__(self, "_id", oid)

    def __getattr__(self, name):
        rec, target, oid = self._rec, self._t, self._id
        fn = getattr(target, name)

        def call(*args, **kwargs):
            ev = {"op": "call", "obj": oid, "name": name, "args": [rec.enc(a) for a in args], "kwargs": {k: rec.enc(v) for k, v in kwargs.items()}}
            raw = [a._t if isinstance(a, Proxy) else a for a in args]
            rkw = {k: (v._t if isinstance(v, Proxy) else v) for k, v in kwargs.items()}
            try:
                r = fn(*raw, **rkw)
            except Exception as exc:
                ev["error"] = type(exc).__name__
                rec.events.append(ev)
                raise
            out, ev["result"] = rec.wrap_result(r)
            rec.events.append(ev)
            return out
        return call


def factory(rec, cls):
    def make(*args, **kwargs):
        oid = rec.new_id()
        ev = {"op": "new", "cls": cls, "id": oid}
        if cls in ("CustomModel", "ScipyModel"):
            memo = {}
            if cls == "ScipyModel":
                dist, lo, hi = args
                cdf = dist.cdf
            else:
                cdf, _ppf, lo, hi = args

            def memo_cdf(x, *params):
                v = float(cdf(x, *params))
                memo[" ".join(float(t).hex() for t in (x,) + tuple(params))] = v.hex()
                return v
            ev.update(cls="CustomModel", lo=int(lo), hi=int(hi), memo=memo)
            target = rec.api.CustomModel(memo_cdf, None, lo, hi)
        else:
            ev.update(args=[rec.enc(a) for a in args], kwargs={k: rec.enc(v) for k, v in kwargs.items()})
            raw = [a._t if isinstance(a, Proxy) else a for a in args]
            try:
                target = getattr(rec.api, cls)(*raw, **{k: (v._t if isinstance(v, Proxy) else v) for k, v in kwargs.items()})
            except Exception as exc:
                ev["error"] = type(exc).__name__
                rec.events.append(ev)
                raise
        rec.events.append(ev)
        return Proxy(rec, target, oid)
    return make


def shim(rec):
    root = types.ModuleType("constriction")
    stream = types.ModuleType("constriction.stream")
    mods = {n: types.ModuleType(f"constriction.stream.{n}") for n in ("stack", "queue", "model")}
    for cls in CLASSES:
        setattr(mods[WHERE.get(cls, "model")], cls, factory(rec, cls))
    stream.stack, stream.queue, stream.model = mods["stack"], mods["queue"], mods["model"]
    root.stream = stream
    out = {"constriction": root, "constriction.stream": stream}
    out.update({f"constriction.stream.{n}": m for n, m in mods.items()})
    return out


def main():
    from oracle import refapi as O
    O.lib()
    traces = []
    for fn in sorted(os.listdir(REF_TESTS)):
        if not (fn.startswith("test_") and fn.endswith(".py")):
            continue
        names = [line[4:line.index("(")] for line in open(os.path.join(REF_TESTS, fn)) if line.startswith("def test_")]
        for name in names:
            rec = Recorder(O)
            saved = {k: sys.modules.get(k) for k in shim(rec)}
            sys.modules.update(shim(rec))
            try:
                spec = importlib.util.spec_from_file_location(f"_ref_{fn[:-3]}", os.path.join(REF_TESTS, fn))
                mod = importlib.util.module_from_spec(spec)
                spec.loader.exec_module(mod)
                getattr(mod, name)()
                traces.append({"test": f"{fn[:-3]}::{name}", "events": rec.events})
            except AttributeError as exc:
                print(f"skipped {fn}::{name}: {exc}")  # ChainCoder / symbol codes: out of scope
            finally:
                for k, v in saved.items():
                    if v is None:
                        sys.modules.pop(k, None)
                    else:
                        sys.modules[k] = v
    path = os.path.join(ROOT, "tests", "golden", "reference_trace.json.gz")
    with gzip.open(path, "wt", compresslevel=9) as f:
        json.dump({"source": "bamler-lab/constriction v0.5.0 tests/python/*.py run on oracle/refapi.py", "traces": traces}, f)
    n_ev = sum(len(t["events"]) for t in traces)
    print(f"{len(traces)} tests, {n_ev} calls -> {path} ({os.path.getsize(path)} bytes)")


if __name__ == "__main__":
    main()
