"""Per-call cost of the Python mirror (one coder object, like the reference's Python API) against the CPU oracle:
where is the break-even message size?  python tools/mirror_latency.py"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import constriction_b200.stream as S
from oracle import refapi as O

rng = np.random.default_rng(0)
print(f"{'symbols':>9s} {'mirror enc+dec ms':>18s} {'oracle (table) ms':>18s} {'oracle (lazy erf) ms':>21s}")
for n in (10, 100, 1000, 10_000, 100_000, 1_000_000, 10_000_000):
    syms = np.clip(np.rint(rng.normal(3.2, 9.6, size=n)), -50, 50).astype(np.int32)
    ms_model, os_model = S.QuantizedGaussian(-50, 50, 3.2, 9.6), O.QuantizedGaussian(-50, 50, 3.2, 9.6)
    def run(api, model):
        c = api.AnsCoder(); c.encode_reverse(syms, model); w = c.get_compressed()
        out = api.AnsCoder(w).decode(model, n); assert np.array_equal(out, syms)
    run(S, ms_model)
    reps = 5 if n <= 100_000 else 2
    t0 = time.perf_counter()
    for _ in range(reps): run(S, ms_model)
    t_m = (time.perf_counter() - t0) / reps
    cdf = O.qgauss_cdf(-50, 50, 3.2, 9.6)
    t0 = time.perf_counter()
    for _ in range(reps):
        w = O.ans_encode_iid(syms, cdf, -50); O.ans_decode_iid(w, n, cdf, -50)
    t_t = (time.perf_counter() - t0) / reps
    if n <= 1_000_000:
        t0 = time.perf_counter()
        w = O.ans_encode_qgauss_lazy(syms, -50, 50, 3.2, 9.6); O.ans_decode_qgauss_lazy(w, n, -50, 50, 3.2, 9.6)
        t_l = time.perf_counter() - t0
    else:
        t_l = float("nan")
    print(f"{n:9d} {t_m * 1e3:18.3f} {t_t * 1e3:18.3f} {t_l * 1e3:21.3f}", flush=True)
