"""Replays tests/golden/reference_trace.json.gz -- the reference's own pytest files recorded call by call
(tests/golden/make_reference_trace.py) -- against an implementation of the reference's Python API."""
import base64
import gzip
import json
import os

import numpy as np

PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_trace.json.gz")


def load():
    with gzip.open(PATH, "rt") as f:
        return json.load(f)["traces"]


def dec(v, objs):
    if isinstance(v, dict):
        if "obj" in v:
            return objs[v["obj"]]
        if "nd" in v:
            return np.frombuffer(base64.b64decode(v["b64"]), dtype=np.dtype(v["nd"])).reshape(v["shape"]).copy()
        if "int" in v:
            return v["int"]
        if "f64" in v:
            return float.fromhex(v["f64"])
        if "seq" in v:
            return tuple(dec(x, objs) for x in v["seq"])
    return v


def same(got, want):
    if isinstance(want, tuple):
        return isinstance(got, (tuple, list)) and len(got) == len(want) and all(same(g, w) for g, w in zip(got, want))
    if isinstance(want, np.ndarray):
        return isinstance(got, np.ndarray) and got.shape == want.shape and got.dtype == want.dtype and np.array_equal(got, want)
    if isinstance(want, float):
        return float(got) == want
    if want is None:
        return got is None
    return got == want


def replay(api, trace):
    objs = {}
    for i, ev in enumerate(trace["events"]):
        where = f"{trace['test']} call {i}: {ev.get('cls') or ev.get('name')}"
        try:
            if ev["op"] == "new":
                if ev["cls"] == "CustomModel" and "memo" in ev:
                    memo = {k: float.fromhex(v) for k, v in ev["memo"].items()}

                    def cdf(x, *params, _memo=memo):
                        return _memo[" ".join(float(t).hex() for t in (x,) + tuple(params))]
                    objs[ev["id"]] = api.CustomModel(cdf, None, ev["lo"], ev["hi"])
                else:
                    args = [dec(a, objs) for a in ev["args"]]
                    kwargs = {k: dec(v, objs) for k, v in ev["kwargs"].items()}
                    objs[ev["id"]] = getattr(api, ev["cls"])(*args, **kwargs)
                assert "error" not in ev, f"{where}: expected {ev.get('error')}"
                continue
            args = [dec(a, objs) for a in ev["args"]]
            kwargs = {k: dec(v, objs) for k, v in ev["kwargs"].items()}
            got = getattr(objs[ev["obj"]], ev["name"])(*args, **kwargs)
        except Exception as exc:  # noqa: BLE001
            assert ev.get("error") == type(exc).__name__, f"{where}: raised {type(exc).__name__}: {exc}"
            continue
        assert "error" not in ev, f"{where}: expected {ev['error']}"
        res = ev["result"]
        if isinstance(res, dict) and "newobj" in res:
            objs[res["newobj"]] = got
        else:
            want = dec(res, objs)
            assert same(got, want), f"{where}: got {got!r}, the reference test saw {want!r}"
