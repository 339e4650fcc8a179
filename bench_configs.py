#!/usr/bin/env python
"""bench_configs.py -- the other BASELINE.json configs, one JSON line each.

`bench.py` is the contract benchmark (configs[1]).  This script times the remaining GPU configs on one
B200 with the same hygiene (CUDA events, >= 3 warm-ups, device-resident inputs) and checks each result
(round trip and/or oracle spot checks):

  config 3  1e8 symbols, per-symbol Categorical models out of a pool of 1e6 256-bin CDFs (1.03 GB, streamed
            from HBM / L2), ANS decode only
  config 4  RangeEncoder, 122,070-symbol streams (the per-GPU shard of "1e9 symbols in 8192 streams on 8 GPUs"
            is 1024 streams; also run with all 8192 streams on one GPU)
  config 5  learned-image-compression shape: int32[64,192,32,32] latents, 192 per-channel QuantizedGaussian
            models, one ANS stream per (image, channel)
  config 6  (not in BASELINE.json) the same latents with one (mean, std) pair per latent,
            evaluated on the device by the table-free Gaussian kernels

    python bench_configs.py [--configs 3,4,5] [--scale 1.0]
"""
from __future__ import annotations

import argparse
import json
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def timed(fn, warmup=3, steps=5):
    import torch
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    ev[0].record()
    for i in range(steps):
        fn()
        ev[i + 1].record()
    torch.cuda.synchronize()
    return ev[0].elapsed_time(ev[-1]) / steps  # ms


def config3(scale):
    """ANS decode with a per-symbol model index into a 1e6 x 256 CDF pool."""
    import torch
    from constriction_b200 import batch as B
    from oracle import refapi as O
    M, A = int(1_000_000 * scale), 256
    n, k = int(100_000_000 * scale), 148 * 1024
    g = torch.Generator(device="cuda")
    g.manual_seed(3)
    # Dirichlet(0.5) rows = normalised Gamma(0.5) draws, float32 (the pmf dtype decides the rounding)
    gam = torch._standard_gamma(torch.full((M, A), 0.5, device="cuda"), generator=g).clamp_min_(1e-30)
    pmf = (gam / gam.sum(dim=1, keepdim=True)).to(torch.float32)
    del gam
    model = B.ModelTable.categorical(pmf)
    idx = torch.randint(0, M, (n,), device="cuda", dtype=torch.int32, generator=g)
    # ANS decoding is surjective: decoding uniformly random words draws every symbol from its model, which is
    # exactly "symbols sampled per row".  ~7 bits/symbol for these rows -> 0.25 words per symbol is ample.
    words_per_stream = int(math.ceil(n / k * 0.25)) + 2
    words = torch.randint(-2**31, 2**31 - 1, (k * words_per_stream,), device="cuda", dtype=torch.int32, generator=g)
    words[words_per_stream - 1::words_per_stream] |= 1  # a valid ANS stream never ends in a zero word
    offsets = (torch.arange(k + 1, device="cuda", dtype=torch.int64) * words_per_stream)
    comp = B.Compressed(words, offsets, k, n, "ans")
    bc = B.BatchCoder()
    out = torch.empty(n, dtype=torch.int32, device="cuda")

    ms = timed(lambda: bc.ans_decode(comp, model, model_index=idx, index_mode=1, out=out))
    bc.check()
    # parity: spot-check streams against the oracle's per-symbol-model decode
    cdfs_dev = None
    ok = True
    cdf_rows = {}
    for s in (0, 77, k - 1):
        sl = slice(s, n, k)
        ids = idx[sl].cpu().numpy().astype(np.uint32)
        rows = np.unique(ids)
        sub = B.ModelTable.categorical(pmf[torch.from_numpy(rows.astype(np.int64)).cuda()]).cdf()
        remap = np.searchsorted(rows, ids).astype(np.uint32)
        w = words[s * words_per_stream:(s + 1) * words_per_stream].cpu().numpy().view(np.uint32)
        want = O.ans_decode_indexed(w, remap, sub, 0)
        ok &= bool(np.array_equal(out[sl].cpu().numpy(), want))
    # every decoded symbol reads its model's CDF row by binary search: 8 probes + the 2 interval ends
    return {"config": 3, "workload": f"{n} symbols, {M} x {A} categorical pool ({M * (A + 1) * 4 / 1e9:.2f} GB), "
                                     f"per-symbol model index, ANS decode, {k} streams",
            "ms": ms, "Msymbols_per_s": n / ms / 1e3, "parity_spot_check": ok,
            "bytes_per_symbol_if_rows_streamed": 4 * (A + 1) + 8 + 1, "note": "binary search touches <= 9 sectors"}


def config4(scale, streams):
    """Range coder, long contiguous streams."""
    import torch
    from constriction_b200 import batch as B
    from oracle import refapi as O
    per = int(122_070 * scale)
    k, n = streams, streams * per
    g = torch.Generator(device="cuda")
    g.manual_seed(4)
    syms = torch.clamp(torch.round(torch.randn(n, device="cuda", generator=g) * 9.6 + 3.2), -50, 50).to(torch.int32)
    off = torch.arange(k + 1, device="cuda", dtype=torch.int64) * per
    model = B.ModelTable.quantized_gaussian(-50, 50, [3.2], [9.6])
    bc = B.BatchCoder()
    state = {}

    def enc():
        state["c"] = bc.range_encode(syms, model, sym_offsets=off, out=state.get("c"))

    ms_enc = timed(enc)
    comp = state["c"]
    out = torch.empty_like(syms)
    ms_dec = timed(lambda: bc.range_decode(comp, model, out=out))
    bc.check()
    ok = bool(torch.equal(out, syms))
    cdf = model.cdf()[0]
    for s in (0, k - 1):
        want = O.range_encode_iid(syms[s * per:(s + 1) * per].cpu().numpy(), cdf, -50)
        ok &= bool(np.array_equal(comp.stream_words(s), want))
    # the same streams with checkpoints every 1024 symbols: identical words, every chunk decoded on its own lane
    cstate = {}

    def enc_ck():
        cstate["c"] = bc.range_encode(syms, model, sym_offsets=off, checkpoint_every=1024)

    ms_enc_ck = timed(enc_ck)
    ccomp = cstate["c"]
    ok &= bool(torch.equal(ccomp.words[:ccomp.total_words()], comp.words[:comp.total_words()]))
    out.zero_()
    ms_dec_ck = timed(lambda: bc.range_decode(ccomp, model, out=out))
    bc.check()
    ok &= bool(torch.equal(out, syms))
    return {"config": 4, "workload": f"{k} RangeEncoder streams x {per} symbols (contiguous), QG(-50,50,3.2,9.6)",
            "ms_encode": ms_enc, "ms_decode": ms_dec, "Msymbols_per_s_encode": n / ms_enc / 1e3,
            "Msymbols_per_s_decode": n / ms_dec / 1e3, "Msymbols_per_s_round_trip": n / (ms_enc + ms_dec) / 1e3,
            "checkpoints_every_1024": {"ms_encode": ms_enc_ck, "ms_decode": ms_dec_ck,
                                       "Msymbols_per_s_decode": n / ms_dec_ck / 1e3,
                                       "extra_bytes_per_symbol": 32.0 / 1024},
            "parity": ok, "note": "one dependent chain per stream: latency-bound, not HBM-bound; with checkpoints the "
                                  "decode runs one lane per 1024-symbol chunk"}


def config2_range(scale):
    """configs[1] workload with the range coder instead of ANS (interleaved deal, shared model)."""
    import torch
    from constriction_b200 import batch as B
    from oracle import refapi as O
    n, k = int(100_000_000 * scale), 148 * 1024
    g = torch.Generator(device="cuda")
    g.manual_seed(2)
    syms = torch.clamp(torch.round(torch.randn(n, device="cuda", generator=g) * 9.6 + 3.2), -50, 50).to(torch.int32)
    model = B.ModelTable.quantized_gaussian(-50, 50, [3.2], [9.6])
    bc = B.BatchCoder()
    state = {}

    def enc():
        state["c"] = bc.range_encode(syms, model, n_streams=k, out=state.get("c"))

    ms_enc = timed(enc)
    comp = state["c"]
    out = torch.empty_like(syms)
    ms_dec = timed(lambda: bc.range_decode(comp, model, out=out))
    bc.check()
    ok = bool(torch.equal(out, syms))
    cdf = model.cdf()[0]
    for s in (0, 1, k - 1):
        want = O.range_encode_iid(syms[s::k].cpu().numpy(), cdf, -50)
        ok &= bool(np.array_equal(comp.stream_words(s), want))
    return {"config": "2-range", "workload": f"{n} i.i.d. symbols, QG(-50,50,3.2,9.6), {k} RangeEncoder streams (interleaved)",
            "us_encode": ms_enc * 1e3, "us_decode": ms_dec * 1e3,
            "Msymbols_per_s_round_trip": n / (ms_enc + ms_dec) / 1e3, "parity": ok,
            "bits_per_symbol": 32.0 * comp.total_words() / n}


def config5(images=64):
    """int32[images,192,32,32] latents, one QuantizedGaussian per channel, one ANS stream per (image, channel)."""
    import torch
    from constriction_b200 import batch as B
    from oracle import refapi as O
    rng = np.random.default_rng(5)
    C_, HW = 192, 32 * 32
    mu = rng.normal(0, 2, C_)
    sigma = np.exp(rng.uniform(np.log(0.3), np.log(12), C_))
    model = B.ModelTable.quantized_gaussian(-64, 64, mu, sigma)
    g = torch.Generator(device="cuda")
    g.manual_seed(5)
    t_mu = torch.from_numpy(mu).cuda().float()[None, :, None]
    t_sg = torch.from_numpy(sigma).cuda().float()[None, :, None]
    lat = torch.clamp(torch.round(torch.randn(images, C_, HW, device="cuda", generator=g) * t_sg + t_mu), -64, 64)
    syms = lat.to(torch.int32).reshape(-1).contiguous()
    k, n = images * C_, images * C_ * HW
    off = torch.arange(k + 1, device="cuda", dtype=torch.int64) * HW
    sidx = torch.arange(C_, device="cuda", dtype=torch.int32).repeat(images)
    bc = B.BatchCoder()
    state = {}

    def enc():
        state["c"] = bc.ans_encode(syms, model, sym_offsets=off, model_index=sidx, index_mode=2, out=state.get("c"))

    ms_enc = timed(enc)
    comp = state["c"]
    out = torch.empty_like(syms)
    ms_dec = timed(lambda: bc.ans_decode(comp, model, model_index=sidx, index_mode=2, out=out))
    bc.check()
    ok = bool(torch.equal(out, syms))
    cdfs = model.cdf()
    for s in (0, 191, k - 1):
        want = O.ans_encode_iid(syms[s * HW:(s + 1) * HW].cpu().numpy(), cdfs[s % C_], -64)
        ok &= bool(np.array_equal(comp.stream_words(s), want))
    # checkpoints every 128 symbols: the same words, 8 lanes per (image, channel) stream in the decoder
    ck = bc.ans_encode(syms, model, sym_offsets=off, model_index=sidx, index_mode=2, checkpoint_every=128)
    ok &= bool(torch.equal(ck.words[:ck.total_words()], comp.words[:comp.total_words()]))
    out.zero_()
    ms_dec_ck = timed(lambda: bc.ans_decode(ck, model, model_index=sidx, index_mode=2, out=out))
    bc.check()
    ok &= bool(torch.equal(out, syms))
    return {"config": 5, "workload": f"latents int32[{images},192,32,32], 192 per-channel QuantizedGaussian(-64,64), "
                                     f"{k} ANS streams x {HW} symbols",
            "us_encode": ms_enc * 1e3, "us_decode": ms_dec * 1e3,
            "Msymbols_per_s_round_trip": n / (ms_enc + ms_dec) / 1e3,
            "checkpoints_every_128": {"us_decode": ms_dec_ck * 1e3, "Msymbols_per_s_decode": n / ms_dec_ck / 1e3,
                                      "extra_bytes_per_symbol": 16.0 / 128},
            "parity": ok, "bits_per_symbol": 32.0 * comp.total_words() / n}


def config6(images=64, coder="ans"):
    """Config 5's latents with a (mean, std) pair per latent (hyperprior models): the table-free Gaussian kernels
    (SURVEY.md 8f rank 1).  Also reports the pre-pass on its own and the CPU oracle on a sample."""
    import time

    import torch
    from constriction_b200 import batch as B
    from oracle import refapi as O
    C_, HW = 192, 32 * 32
    k, n = images * C_, images * C_ * HW
    g = torch.Generator(device="cuda")
    g.manual_seed(6)
    means = torch.randn(n, device="cuda", generator=g, dtype=torch.float64) * 2.0
    stds = torch.exp(torch.rand(n, device="cuda", generator=g, dtype=torch.float64) * 3.7 - 1.2)  # 0.3 .. 12
    syms = torch.clamp(torch.round(means + stds * torch.randn(n, device="cuda", generator=g, dtype=torch.float64)), -64, 64).to(torch.int32)
    model = B.GaussianParams(-64, 64, means, stds)
    off = torch.arange(k + 1, device="cuda", dtype=torch.int64) * HW
    bc = B.BatchCoder()
    enc_fn, dec_fn = (bc.ans_encode, bc.ans_decode) if coder == "ans" else (bc.range_encode, bc.range_decode)
    state = {}

    def enc():
        state["c"] = enc_fn(syms, model, sym_offsets=off, out=state.get("c"))

    ms_enc = timed(enc)
    comp = state["c"]
    out = torch.empty_like(syms)
    ms_dec = timed(lambda: dec_fn(comp, model, out=out))
    bc.check()
    ok = bool(torch.equal(out, syms))
    # oracle: stream 0 and the last one, one quantised Gaussian per symbol
    t_cpu = 0.0
    for s in (0, k - 1):
        sl = slice(s * HW, (s + 1) * HW)
        hm, hs, hy = means[sl].cpu().numpy(), stds[sl].cpu().numpy(), syms[sl].cpu().numpy()
        t0 = time.perf_counter()
        cdfs = np.stack([O.qgauss_cdf(-64, 64, float(a), float(b)) for a, b in zip(hm, hs)])
        t_cpu += time.perf_counter() - t0
        if coder == "ans":
            ok &= bool(np.array_equal(comp.stream_words(s), O.ans_encode_indexed(hy, np.arange(HW), cdfs, -64)))
    # the same batch with checkpoints every 128 symbols: 8 chunks per stream decode on 8 lanes (the Gaussian decoder is
    # bound by the latency of its FP64 chain, so lanes are what it needs)
    ck = enc_fn(syms, model, sym_offsets=off, checkpoint_every=128)
    ok &= bool(torch.equal(ck.words[:ck.total_words()], comp.words[:comp.total_words()]))
    out.zero_()
    ms_dec_ck = timed(lambda: dec_fn(ck, model, out=out))
    bc.check()
    ok &= bool(torch.equal(out, syms))
    extra = {"checkpoints_every_128": {"us_decode": ms_dec_ck * 1e3, "Msymbols_per_s_decode": n / ms_dec_ck / 1e3}}
    return {**extra, "config": f"6-{coder}", "workload": f"latents int32[{images},192,32,32], one QuantizedGaussian(-64,64,mean,std) per latent "
                                              f"(f64 parameters on the device, no tables), {k} {coder} streams x {HW} symbols",
            "us_encode": ms_enc * 1e3, "us_decode": ms_dec * 1e3,
            "Msymbols_per_s_encode": n / ms_enc / 1e3, "Msymbols_per_s_decode": n / ms_dec / 1e3,
            "Msymbols_per_s_round_trip": n / (ms_enc + ms_dec) / 1e3, "parity": ok,
            "bits_per_symbol": 32.0 * comp.total_words() / n}


# ----------------------------------------------------------------------------------------------------
# The multi-GPU configs of BASELINE.json, callable under torch.distributed from bench.py (`extra_configs`)
# ----------------------------------------------------------------------------------------------------
def _max_over_ranks(values, world):
    import torch
    import torch.distributed as dist
    t = torch.tensor(list(values), dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(x) for x in t.tolist()]


def _all_ranks_ok(ok, world):
    import torch
    import torch.distributed as dist
    t = torch.tensor([1 if ok else 0], device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
    return bool(int(t.item()))


def _exchange(sg, bc, world, rank, encode_into, k_of, n_of, coder, decode_neighbour, steps=4):
    """Times `encode -> push -> wait for every peer's slot -> release` (the whole exchange exposed, nothing overlapped)
    and checks that the right neighbour's shard decodes out of the gathered container."""
    import torch
    if world == 1:
        return None, True
    nb = (rank + 1) % world
    turns = []

    def one():
        turn = sg.begin_turn(k_of(rank), n_of(rank), coder)
        encode_into(turn.out)
        sg.push(turn, k_of(rank))
        sg.wait(turn)
        turns.append(turn)
        if len(turns) > 1:
            sg.release(turns.pop(0))

    ms = timed(one, warmup=2, steps=steps)
    ok = decode_neighbour(sg.shard(turns[-1], nb, k_of(nb), n_of(nb), coder), nb)
    bc.check()
    torch.cuda.synchronize()
    sg.release(turns.pop())
    sg.sync()
    return ms, ok


def baseline_config3_sharded(world=1, rank=0, total_streams=8192, total_symbols=1_000_000_000):
    """BASELINE configs[3]: 1e9 symbols as 8192 independent RangeEncoder streams (122,070 symbols each + remainder),
    sharded over the ranks by blocks of streams (strong scaling: 8192 / N streams per GPU), containers gathered."""
    import torch
    from constriction_b200 import batch as B
    from constriction_b200 import dist as D
    from oracle import refapi as O
    base, extra = divmod(total_symbols, total_streams)

    def shard(r):
        lo, hi = D.shard_bounds(total_streams, world, r)
        lens = torch.full((hi - lo,), base, dtype=torch.int64)
        lens[: max(0, min(hi, extra) - lo)] += 1  # the first `extra` streams of the job carry one more symbol
        off = torch.cat([torch.zeros(1, dtype=torch.int64), torch.cumsum(lens, 0)])
        return hi - lo, int(off[-1]), off

    def symbols(r, n):
        g = torch.Generator(device="cuda")
        g.manual_seed(4 + r)
        return torch.clamp(torch.round(torch.randn(n, device="cuda", generator=g) * 9.6 + 3.2), -50, 50).to(torch.int32)

    k, n, off_h = shard(rank)
    off = off_h.cuda()
    syms = symbols(rank, n)
    model = B.ModelTable.quantized_gaussian(-50, 50, [3.2], [9.6])
    bc = B.BatchCoder()
    st = {}

    def enc():
        st["c"] = bc.range_encode(syms, model, sym_offsets=off, out=st.get("c"))

    ms_enc = timed(enc, warmup=2, steps=4)
    comp = st["c"]
    out = torch.empty_like(syms)
    ms_dec = timed(lambda: bc.range_decode(comp, model, out=out), warmup=1, steps=3)
    bc.check()
    ok = bool(torch.equal(out, syms))
    cdf = model.cdf()[0]
    for s_ in (0, k - 1):
        a, b = int(off_h[s_]), int(off_h[s_ + 1])
        ok &= bool(np.array_equal(comp.stream_words(s_), O.range_encode_iid(syms[a:b].cpu().numpy(), cdf, -50)))
    total_words = comp.total_words()
    # checkpoints every 1024 symbols: identical words, every chunk decodes on its own lane
    ck = bc.range_encode(syms, model, sym_offsets=off, checkpoint_every=1024)
    ok &= bool(torch.equal(ck.words[:ck.total_words()], comp.words[:total_words]))
    out.zero_()
    ms_dec_ck = timed(lambda: bc.range_decode(ck, model, out=out), warmup=2, steps=4)
    bc.check()
    ok &= bool(torch.equal(out, syms))
    del ck
    ms_x, sg = None, None
    if world > 1:
        sizes = [shard(r) for r in range(world)]
        sg = D.SlotGather(int(total_words * 1.1) + 4096, max(x[0] for x in sizes))

        def dec_nb(sh, nb):
            knb, nnb, offnb = sizes[nb]
            sh.sym_offsets = offnb.cuda()
            got = bc.range_decode(sh, model)
            return bool(torch.equal(got, symbols(nb, nnb)))

        ms_x, ok_x = _exchange(sg, bc, world, rank, lambda o: bc.range_encode(syms, model, sym_offsets=off, out=o),
                               lambda r: sizes[r][0], lambda r: sizes[r][1], "range", dec_nb)
        ok &= ok_x
        sg.close()
    ms_enc, ms_dec, ms_dec_ck, ms_xm = _max_over_ranks([ms_enc, ms_dec, ms_dec_ck, ms_x or 0.0], world)
    ok = _all_ranks_ok(ok, world)
    N = total_symbols
    return {"workload": f"{N} symbols as {total_streams} RangeEncoder streams (contiguous), QG(-50,50,3.2,9.6), "
                        f"{total_streams // world} streams per GPU on {world} GPU(s)",
            "ms_encode": ms_enc, "ms_decode": ms_dec, "ms_decode_with_checkpoints_every_1024": ms_dec_ck,
            "ms_encode_plus_gather_exposed": ms_xm if world > 1 else None,
            "Msymbols_per_s_encode": N / ms_enc / 1e3, "Msymbols_per_s_decode": N / ms_dec / 1e3,
            "Msymbols_per_s_round_trip": N / (ms_enc + ms_dec) / 1e3,
            "Msymbols_per_s_round_trip_checkpointed": N / (ms_enc + ms_dec_ck) / 1e3,
            "compressed_MB_per_gpu": total_words * 4 / 1e6, "parity": ok,
            "note": "times are the max over ranks; one dependent chain per stream: latency-bound, not HBM-bound"}


def baseline_config4_sharded(world=1, rank=0, images=64):
    """BASELINE configs[4]: int32[64,192,32,32] latents, 192 per-channel QuantizedGaussian models, one ANS stream per
    (image, channel); sharded by image (64 / N images per GPU), containers gathered."""
    import torch
    from constriction_b200 import batch as B
    from constriction_b200 import dist as D
    from oracle import refapi as O
    rng = np.random.default_rng(5)
    C_, HW = 192, 32 * 32
    mu = rng.normal(0, 2, C_)
    sigma = np.exp(rng.uniform(np.log(0.3), np.log(12), C_))
    model = B.ModelTable.quantized_gaussian(-64, 64, mu, sigma)
    t_mu = torch.from_numpy(mu).cuda().float()[None, :, None]
    t_sg = torch.from_numpy(sigma).cuda().float()[None, :, None]

    def images_of(r):
        lo, hi = D.shard_bounds(images, world, r)
        return hi - lo

    def symbols(r):
        g = torch.Generator(device="cuda")
        g.manual_seed(50 + r)
        lat = torch.clamp(torch.round(torch.randn(images_of(r), C_, HW, device="cuda", generator=g) * t_sg + t_mu), -64, 64)
        return lat.to(torch.int32).reshape(-1).contiguous()

    im = images_of(rank)
    syms = symbols(rank)
    k, n = im * C_, im * C_ * HW
    off = torch.arange(k + 1, device="cuda", dtype=torch.int64) * HW
    sidx = torch.arange(C_, device="cuda", dtype=torch.int32).repeat(im)
    bc = B.BatchCoder()
    st = {}

    def enc():
        st["c"] = bc.ans_encode(syms, model, sym_offsets=off, model_index=sidx, index_mode=2, out=st.get("c"))

    ms_enc = timed(enc, warmup=3, steps=10)
    comp = st["c"]
    out = torch.empty_like(syms)
    ms_dec = timed(lambda: bc.ans_decode(comp, model, model_index=sidx, index_mode=2, out=out), warmup=3, steps=10)
    bc.check()
    ok = bool(torch.equal(out, syms))
    cdfs = model.cdf()
    for s_ in (0, 191, k - 1):
        want = O.ans_encode_iid(syms[s_ * HW:(s_ + 1) * HW].cpu().numpy(), cdfs[s_ % C_], -64)
        ok &= bool(np.array_equal(comp.stream_words(s_), want))
    total_words = comp.total_words()
    ck = bc.ans_encode(syms, model, sym_offsets=off, model_index=sidx, index_mode=2, checkpoint_every=128)
    ok &= bool(torch.equal(ck.words[:ck.total_words()], comp.words[:total_words]))
    out.zero_()
    ms_dec_ck = timed(lambda: bc.ans_decode(ck, model, model_index=sidx, index_mode=2, out=out), warmup=3, steps=10)
    bc.check()
    ok &= bool(torch.equal(out, syms))
    ms_x = None
    if world > 1:
        ks = [images_of(r) * C_ for r in range(world)]
        sg = D.SlotGather(int(total_words * 1.2) + 4096, max(ks))

        def dec_nb(sh, nb):
            sh.sym_offsets = torch.arange(ks[nb] + 1, device="cuda", dtype=torch.int64) * HW
            idx = torch.arange(C_, device="cuda", dtype=torch.int32).repeat(images_of(nb))
            got = bc.ans_decode(sh, model, model_index=idx, index_mode=2)
            return bool(torch.equal(got, symbols(nb)))

        ms_x, ok_x = _exchange(sg, bc, world, rank,
                               lambda o: bc.ans_encode(syms, model, sym_offsets=off, model_index=sidx, index_mode=2, out=o),
                               lambda r: ks[r], lambda r: ks[r] * HW, "ans", dec_nb, steps=10)
        ok &= ok_x
        sg.close()
    ms_enc, ms_dec, ms_dec_ck, ms_xm = _max_over_ranks([ms_enc, ms_dec, ms_dec_ck, ms_x or 0.0], world)
    ok = _all_ranks_ok(ok, world)
    N = images * C_ * HW
    return {"workload": f"latents int32[{images},192,32,32], 192 per-channel QuantizedGaussian(-64,64), {images * C_} ANS streams "
                        f"x {HW} symbols, {images // world} images per GPU on {world} GPU(s)",
            "us_encode": ms_enc * 1e3, "us_decode": ms_dec * 1e3, "us_decode_with_checkpoints_every_128": ms_dec_ck * 1e3,
            "us_encode_plus_gather_exposed": ms_xm * 1e3 if world > 1 else None,
            "Msymbols_per_s_round_trip": N / (ms_enc + ms_dec) / 1e3,
            "Msymbols_per_s_round_trip_checkpointed": N / (ms_enc + ms_dec_ck) / 1e3,
            "bits_per_symbol": 32.0 * total_words / n, "parity": ok,
            "note": "times are the max over ranks; a 50 MB problem: launch- and latency-bound"}


def north_star_1e9(world=1, rank=0, n=1_000_000_000, k=148 * 1024):
    """The north star's size: 1e9 i.i.d. symbols per GPU, ANS encode + decode (configs[1]'s model and layout)."""
    import torch
    from constriction_b200 import batch as B
    from oracle import refapi as O
    g = torch.Generator(device="cuda")
    g.manual_seed(9 + rank)
    syms = torch.empty(n, dtype=torch.int32, device="cuda")
    for i in range(0, n, 1 << 27):  # generated in pieces: the float temporaries of one shot would need 12 GB
        m = min(1 << 27, n - i)
        syms[i:i + m] = torch.clamp(torch.round(torch.randn(m, device="cuda", generator=g) * 9.6 + 3.2), -50, 50).to(torch.int32)
    model = B.ModelTable.quantized_gaussian(-50, 50, [3.2], [9.6])
    bc = B.BatchCoder()
    st = {}

    def enc():
        st["c"] = bc.ans_encode(syms, model, n_streams=k, out=st.get("c"))

    ms_enc = timed(enc, warmup=2, steps=4)
    comp = st["c"]
    out = torch.empty_like(syms)
    ms_dec = timed(lambda: bc.ans_decode(comp, model, out=out), warmup=2, steps=4)
    bc.check()
    ok = bool(torch.equal(out, syms))
    cdf = model.cdf()[0]
    for s_ in (0, 4097, k - 1):
        ok &= bool(np.array_equal(comp.stream_words(s_), O.ans_encode_iid(syms[s_::k].cpu().numpy(), cdf, -50)))
    total_words = comp.total_words()
    ms_enc, ms_dec = _max_over_ranks([ms_enc, ms_dec], world)
    ok = _all_ranks_ok(ok, world)
    return {"workload": f"{n} i.i.d. symbols per GPU, QG(-50,50,3.2,9.6), {k} ANS streams (interleaved), {world} GPU(s)",
            "ms_encode": ms_enc, "ms_decode": ms_dec, "Msymbols_per_s_round_trip": world * n / (ms_enc + ms_dec) / 1e3,
            "hbm_frac_encode": (4.0 * n + 4.0 * total_words) / (ms_enc * 1e-3) / 1e9 / 6546.9,
            "hbm_frac_decode": (4.0 * n + 4.0 * total_words) / (ms_dec * 1e-3) / 1e9 / 6546.9,
            "bits_per_symbol": 32.0 * total_words / n, "parity": ok}


def lookup_bench():
    """benches/lookup.rs in batch form: 10,000 symbols from a 100-symbol categorical model at 12-bit precision,
    SmallAnsCoder / SmallRangeEncoder encode + decode with a lookup decoder model; one coder (the reference's shape: one
    kernel launch per call, latency-bound) and 65,536 coders of 10,000 symbols each."""
    import torch
    from constriction_b200 import small as S
    from oracle import refapi as O
    rng = np.random.default_rng(123)
    pmf = rng.dirichlet(0.5 * np.ones(100)).astype(np.float32)
    model = S.SmallModel.categorical(pmf, perfect=True)
    cdf = model.cdf()[0]
    p = np.diff(cdf.astype(np.int64)) / 4096.0
    bc = S.SmallBatchCoder()
    rows = []
    for k in (1, 65536):
        n = 10_000 * k
        syms = torch.multinomial(torch.tensor(p, device="cuda", dtype=torch.float32), n, replacement=True).to(torch.int32)
        for coder in ("ans", "range"):
            enc = bc.ans_encode if coder == "ans" else bc.range_encode
            dec = bc.ans_decode if coder == "ans" else bc.range_decode
            st = {}

            def e():
                st["c"] = enc(syms, model, n_streams=k)

            ms_e = timed(e, warmup=3, steps=5)
            out = torch.empty_like(syms)
            ms_d = timed(lambda: dec(st["c"], model, out=out), warmup=3, steps=5)
            bc.check()
            ok = bool(torch.equal(out, syms))
            enc1 = O.g_ans_encode if coder == "ans" else O.g_range_encode
            ok &= bool(np.array_equal(st["c"].stream_words(0), enc1("small", syms[0::k].cpu().numpy(), cdf).astype(np.uint16)))
            rows.append({"coder": coder, "streams": k, "symbols": n, "us_encode": ms_e * 1e3, "us_decode": ms_d * 1e3,
                         "ns_per_symbol_encode": ms_e * 1e6 / n, "ns_per_symbol_decode": ms_d * 1e6 / n, "parity": ok})
    return {"config": "lookup", "workload": "benches/lookup.rs shape: 10,000 symbols per coder, 100-symbol categorical, Small preset "
                                            "(u16 words, u32 state, 12 bits), lookup decoder model in shared memory",
            "reference_published": "README.md:202-206: ANS 6.1 ns/symbol decode with a lookup model (i7-7500U)", "rows": rows}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="3,4,5")
    ap.add_argument("--scale", type=float, default=1.0)
    args = ap.parse_args()
    import torch
    assert torch.cuda.is_available(), "needs a CUDA device"
    for c in [int(x) for x in args.configs.split(",")]:
        if c == 2:
            print(json.dumps(config2_range(args.scale)), flush=True)
        elif c == 3:
            print(json.dumps(config3(args.scale)), flush=True)
        elif c == 4:
            print(json.dumps(config4(args.scale, 1024)), flush=True)
            print(json.dumps(config4(args.scale, 8192)), flush=True)
        elif c == 5:
            print(json.dumps(config5(64)), flush=True)
            print(json.dumps(config5(8)), flush=True)
        elif c == 7:
            print(json.dumps(lookup_bench()), flush=True)
        elif c == 6:
            print(json.dumps(config6(64, "ans")), flush=True)
            print(json.dumps(config6(64, "range")), flush=True)
            print(json.dumps(config6(8, "ans")), flush=True)
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
