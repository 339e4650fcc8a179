#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric: Msymbols/s of bit-exact ANS encode+decode on B200(s).

Workload (BASELINE.json configs[1]): 1e8 i.i.d. int32 symbols ~ QuantizedGaussian(-50,50,3.2,9.6),
dealt round-robin to K lane-streams (one independent reference coder per GPU lane), CDF tables in
shared memory.  A "step" is one pass of the hot path over the batch: ANS encode of all streams
(coder kernel + compaction into the dense container), [N>1: all-gather of the compressed containers
through the C ABI's slotted copy-engine exchange, ctr_gather_*], ANS decode of all streams.  `value` is measured with inputs resident in HBM; `e2e` is
the same step through the host-buffer C ABI (pinned host buffers, H2D/D2H copies inside the timed
region).  Weak scaling: every GPU gets its own 1e8-symbol shard.  The line also carries `extra_configs`:
BASELINE.json configs[3] (1e9 symbols in 8192 RangeEncoder streams, sharded over the N ranks, gathered),
configs[4] (64x192x32x32 latents, sharded by image) and the north star's 1e9-symbol size, each with its
own timings and parity bit.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

`--impl reference` times the reference's CPU algorithm (the oracle's C restatement; the Rust crate
cannot be built in this image) on the host cores for the same config.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

MODEL = (-50, 50, 3.2, 9.6)
METRIC = "Msymbols/s ANS encode+decode (bit-exact)"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--symbols", type=int, default=100_000_000, help="symbols per GPU")
    ap.add_argument("--streams", type=int, default=148 * 1024, help="independent coders (lanes) per GPU")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--cpu-sample", type=int, default=50_000_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra-configs", action="store_true")
    ap.add_argument("--gather-lag", type=int, default=int(os.environ.get("CTR_GATHER_LAG", "1")),
                    help="N>1: a turn's gathered container is consumed this many steps after it was encoded")
    return ap.parse_args()


def measured_peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def synth_symbols_numpy(n, seed):
    rng = np.random.default_rng(seed)
    return np.clip(np.rint(rng.normal(MODEL[2], MODEL[3], size=n)), MODEL[0], MODEL[1]).astype(np.int32)


# ----------------------------------------------------------------------------------------------------
# clocks sampling (B200_PROFILING.md recipe)
# ----------------------------------------------------------------------------------------------------
_SAMPLER_SRC = r"""
import json, os, sys, time
idx = int(sys.argv[1])
samples, max_mhz, source = [], None, "nvml"
try:
    import pynvml as nv
    nv.nvmlInit()
    h = nv.nvmlDeviceGetHandleByIndex(idx)
    max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
    get = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
    def bit(a, b, d):
        return getattr(nv, a, getattr(nv, b, d))
    bits = [bit("nvmlClocksEventReasonHwSlowdown", "nvmlClocksThrottleReasonHwSlowdown", 0x8),
            bit("nvmlClocksEventReasonHwThermalSlowdown", "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
            bit("nvmlClocksEventReasonSwThermalSlowdown", "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
            bit("nvmlClocksEventReasonSwPowerCap", "nvmlClocksThrottleReasonSwPowerCap", 0x4)]
    def sample():
        r = int(get(h))
        return [time.time(), float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)), [bool(r & b) for b in bits]]
except Exception:
    import subprocess
    source = "nvidia-smi"
    F = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    def sample():
        global max_mhz
        o = subprocess.run(["nvidia-smi", "--id=%d" % idx, "--query-gpu=" + F, "--format=csv,noheader,nounits"],
                           capture_output=True, text=True, timeout=5).stdout.strip().split(",")
        max_mhz = float(o[1])
        return [time.time(), float(o[0]), [x.strip().lower().startswith("active") for x in o[2:6]]]
print("ready", flush=True)
os.set_blocking(0, False)
while True:
    try:
        samples.append(sample())
    except Exception:
        pass
    try:
        if os.read(0, 16):
            break
    except BlockingIOError:
        pass
    except Exception:
        break
    if len(samples) > 2000000:
        break
print(json.dumps({"samples": samples, "max": max_mhz, "source": source}), flush=True)
"""


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region by a child process that polls NVML
    back to back (one query takes ~0.1 ms, so a timed region of a few milliseconds gets tens of samples;
    a thread of this process would compete with the launch loop for the GIL).  nvidia-smi is the fallback."""
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, gpu_index: int):
        phys = gpu_index
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:  # NVML enumerates physical devices
            ids = [v.strip() for v in vis.split(",") if v.strip()]
            if gpu_index < len(ids) and ids[gpu_index].isdigit():
                phys = int(ids[gpu_index])
        self.result = None
        self.t0 = self.t1 = None
        try:
            self.proc = subprocess.Popen([sys.executable, "-c", _SAMPLER_SRC, str(phys)], stdin=subprocess.PIPE,
                                         stdout=subprocess.PIPE, text=True)
            self.proc.stdout.readline()  # "ready": NVML is initialised, sampling has started
        except Exception:
            self.proc = None

    def __enter__(self):
        self.t0 = time.time()
        return self

    def __exit__(self, *exc):
        self.t1 = time.time()
        if self.proc is None:
            return
        try:
            out, _ = self.proc.communicate("stop\n", timeout=30)
            self.result = json.loads(out.strip().splitlines()[-1])
        except Exception:
            self.proc.kill()

    def summary(self):
        if not self.result or not self.result["samples"]:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        inside = [s for s in self.result["samples"] if self.t0 <= s[0] <= self.t1]
        used = inside or self.result["samples"][-3:]
        sm = sorted(s[1] for s in used)
        reasons = [n for i, n in enumerate(self.NAMES) if any(s[2][i] for s in used)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": self.result["max"], "reasons": reasons,
                "samples": len(inside), "source": self.result["source"]}


# ----------------------------------------------------------------------------------------------------
# CPU legs (oracle; test infrastructure used only as the reported baseline / reference arm)
# ----------------------------------------------------------------------------------------------------
def cpu_round_trip(oracle, syms, k, cdf, threads):
    """K independent reference coders over contiguous chunks (the cache-friendly way to split a message
    on a CPU; same number of coders and symbols as the GPU batch)."""
    chunks = (np.arange(k + 1, dtype=np.uint64) * np.uint64(syms.size)) // np.uint64(k)
    t0 = time.perf_counter()
    words, off = oracle.multi_ans_encode(syms, k, cdf, MODEL[0], sym_offsets=chunks, threads=threads)
    t1 = time.perf_counter()
    out = oracle.multi_ans_decode(words, off, syms.size, k, cdf, MODEL[0], sym_offsets=chunks, threads=threads)
    t2 = time.perf_counter()
    assert np.array_equal(out, syms)
    return t1 - t0, t2 - t1


def cpu_model_string():
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.lower().startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


def cpu_baseline(n_sample, k, threads, small=False):
    """SURVEY 8(d)'s three rows of the restated reference on this box's host cores:
    (i) one thread, lazily evaluated QuantizedGaussian (what stock constriction computes for this model: 2 erf per
    encoded symbol, inverse + ~3 erf per decoded symbol); (ii) one thread, tabulated model; (iii) all cores, tabulated.
    `value` is row (iii), the fastest, so that every reported ratio is conservative."""
    from oracle import refapi as O
    cdf = O.qgauss_cdf(*MODEL)
    syms = synth_symbols_numpy(n_sample, 2)
    cpu_round_trip(O, syms[: n_sample // 10], k, cdf, threads)  # warm the threads / caches
    te, td = min((cpu_round_trip(O, syms, k, cdf, threads) for _ in range(2)), key=sum)
    rows = {"table_all_cores": {"value": n_sample / (te + td) / 1e6, "cores": threads, "sample": n_sample,
                                "encode_s": te, "decode_s": td}}
    n1 = min(n_sample, 4_000_000 if small else 20_000_000)
    one = syms[:n1]
    t0 = time.perf_counter()
    w = O.ans_encode_iid(one, cdf, MODEL[0])
    t1 = time.perf_counter()
    out = O.ans_decode_iid(w, n1, cdf, MODEL[0])
    t2 = time.perf_counter()
    assert np.array_equal(out, one)
    rows["table_1thread"] = {"value": n1 / (t2 - t0) / 1e6, "cores": 1, "sample": n1, "ns_per_symbol_encode": (t1 - t0) / n1 * 1e9,
                             "ns_per_symbol_decode": (t2 - t1) / n1 * 1e9}
    n2 = min(n_sample, 1_000_000 if small else 4_000_000)
    lazy = syms[:n2]
    t0 = time.perf_counter()
    w2 = O.ans_encode_qgauss_lazy(lazy, *MODEL)
    t1 = time.perf_counter()
    out = O.ans_decode_qgauss_lazy(w2, n2, *MODEL)
    t2 = time.perf_counter()
    assert np.array_equal(out, lazy) and np.array_equal(w2, O.ans_encode_iid(lazy, cdf, MODEL[0]))
    rows["lazy_erf_1thread"] = {"value": n2 / (t2 - t0) / 1e6, "cores": 1, "sample": n2, "ns_per_symbol_encode": (t1 - t0) / n2 * 1e9,
                                "ns_per_symbol_decode": (t2 - t1) / n2 * 1e9,
                                "note": "the work stock constriction does for QuantizedGaussian (quantize.rs:525-568,580-779)"}
    return {"value": rows["table_all_cores"]["value"], "unit": "Msymbols/s", "cores": threads, "kind": "port",
            "cpu_model": cpu_model_string(), "rows": rows,
            "reference_published": "README.md:202-206 (i7-7500U, Rust, table models): ANS 24.2 ns encode / 6.1 ns decode per symbol",
            "sample": f"{n_sample} symbols of the same workload in {k} streams, tabulated model, C restatement "
                      f"(oracle/) of stack.rs encode/decode, {threads} threads; encode {te:.3f}s decode {td:.3f}s; "
                      f"rows (i)/(ii) on one thread over {n2}/{n1} symbols"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import refapi as O
    threads = os.cpu_count() or 1
    cdf = O.qgauss_cdf(*MODEL)
    n_sample = min(args.symbols, args.cpu_sample)
    k = args.streams
    syms = synth_symbols_numpy(n_sample, 2)
    for _ in range(args.warmup):
        cpu_round_trip(O, syms, k, cdf, threads)
    times = []
    for _ in range(args.steps):
        te, td = cpu_round_trip(O, syms, k, cdf, threads)
        times.append(te + td)
    t = sum(times) / len(times)
    value = n_sample / t / 1e6
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "Msymbols/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": "configs[1]: 1e8 i.i.d. symbols, QuantizedGaussian(-50,50,3.2,9.6), ANS encode+decode",
                   "symbols_per_step": n_sample, "streams": k,
                   "note": "reference = CPU restatement of constriction's stack.rs loops (Rust toolchain absent); "
                           "each step is a bounded sample of the workload; tabulated model on all cores, i.e. faster "
                           "than stock constriction's lazily evaluated QuantizedGaussian"},
        "cpu_baseline": {"value": value, "unit": "Msymbols/s", "cores": threads, "kind": "port", "cpu_model": cpu_model_string(),
                         "sample": f"{n_sample} symbols per step, {k} streams, tabulated model, {threads} threads"},
        "e2e": {"value": value, "unit": "Msymbols/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------------
def pin_to_gpu_numa_node(local_rank):
    """Run this process (and the pinned buffers it allocates from now on) on the CPUs next to its GPU.  Returns the
    previous affinity mask so that the CPU baseline can have all cores back."""
    try:
        before = os.sched_getaffinity(0)
        import pynvml as nv
        nv.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        phys = local_rank
        if vis:
            ids = [v.strip() for v in vis.split(",") if v.strip()]
            if local_rank < len(ids) and ids[local_rank].isdigit():
                phys = int(ids[local_rank])
        h = nv.nvmlDeviceGetHandleByIndex(phys)
        n_words = (os.cpu_count() + 63) // 64
        mask = nv.nvmlDeviceGetCpuAffinity(h, n_words)
        cpus = {64 * i + b for i, w in enumerate(mask) for b in range(64) if (int(w) >> b) & 1} & before
        if cpus:
            os.sched_setaffinity(0, cpus)
            return before, len(cpus)
        return before, len(before)
    except Exception:
        return None, None


def run_ours(args):
    import torch
    import torch.distributed as dist

    from constriction_b200 import _native as N
    from constriction_b200 import batch as B
    from constriction_b200 import dist as D

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl ours needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    all_cpus, numa_cpus = pin_to_gpu_numa_node(local_rank)
    if world > 1:
        # NCCL prints its version banner to stdout when the first communicator is created; stdout must carry
        # exactly one JSON line, so file descriptor 1 points to stderr until the warm-up is over
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    lib = N.load()
    n, k = args.symbols, args.streams

    def synth(seed_rank):  # synthetic shard, generated on the device (seed differs per rank), resident in HBM
        g = torch.Generator(device="cuda")
        g.manual_seed(2 + seed_rank)
        return torch.clamp(torch.round(torch.randn(n, device="cuda", generator=g) * MODEL[3] + MODEL[2]), MODEL[0],
                           MODEL[1]).to(torch.int32)

    syms = synth(rank)
    model = B.ModelTable.quantized_gaussian(MODEL[0], MODEL[1], [MODEL[2]], [MODEL[3]])
    bc = B.BatchCoder()
    out = torch.empty_like(syms)
    L0 = N.Layout()
    L0.n_streams, L0.n_symbols = k, n
    cap_words = int(lib.ctr_ans_max_compressed_words(C.byref(L0)))
    state = {"comp": None, "prev": [], "sg": None}
    neighbour = (rank + 1) % world

    gather_kind = None
    if world > 1:
        if os.environ.get("CTR_GATHER", "peer") == "peer":
            try:
                state["sg"] = D.SlotGather(cap_words, k, n_buffers=args.gather_lag + 1)
                ok = 1
            except Exception as exc:  # no symmetric memory / stream memory operations here: NCCL all-gather instead
                sys.stderr.write(f"rank {rank}: SlotGather unavailable ({exc}); using the NCCL all-gather\n")
                ok = 0
            flag = torch.tensor([ok], device="cuda")
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)  # all ranks take the same path
            if int(flag.item()) == 0:
                state["sg"] = None
        gather_kind = ("ctr_gather_*: every rank encodes into its slot of a symmetric-memory container and a library thread pushes "
                       "it to all peers with copy-engine transfers over NVLink, ordered by stream memory operations (no kernel, "
                       "no cross-rank host wait)") if state["sg"] else "NCCL all-gather (torch.distributed), one host wait for the sizes"
    sg = state["sg"]

    def step():
        # N == 1: encode -> decode.
        # N > 1: encode my shard straight into my slot of the gathered container; push it to all peers (asynchronously,
        # no SM); wait for the PREVIOUS turn's container to be complete and decode my right neighbour's shard out of
        # it (so every decode reads words that crossed NVLink); release that turn.  In steady state a step costs
        # max(encode + decode, gather); drain() joins the last gather inside the timed region.
        if world == 1:
            state["comp"] = bc.ans_encode(syms, model, n_streams=k, out=state["comp"])
            bc.ans_decode(state["comp"], model, out=out)
            return
        if sg is None:
            state["comp"] = bc.ans_encode(syms, model, n_streams=k, out=state["comp"])
            gc = D.all_gather_compressed(state["comp"].words, state["comp"].offsets, stream_counts=[k] * world)
            state["gc"] = gc
            bc.ans_decode(state["comp"], model, out=out)
            return
        turn = sg.begin_turn(k, n, "ans")
        state["comp"] = bc.ans_encode(syms, model, n_streams=k, out=turn.out)
        sg.push(turn, k)
        pending = state["prev"]
        if len(pending) < args.gather_lag:
            bc.ans_decode(state["comp"], model, out=out)
        else:
            prev = pending.pop(0)
            sg.wait(prev)
            bc.ans_decode(sg.shard(prev, neighbour, k, n, "ans"), model, out=out)
            sg.release(prev)
        pending.append(turn)

    def drain():  # the last turn's gather completes inside the timed region
        while sg is not None and state["prev"]:
            prev = state["prev"].pop(0)
            sg.wait(prev)
            sg.release(prev)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)  # started before the warm-up so that the GPU is not idle right before the timed steps
    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()
    if world > 1:
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        os.close(saved_stdout)
    bc.check()
    total_words = state["comp"].total_words()
    if sg is not None:  # `out` = my neighbour's shard, decoded from the container that was gathered one step ago
        want = synth(neighbour)
        assert torch.equal(out, want), "decode of the neighbour's shard from the gathered container != its symbols"
        del want
        last = state["prev"][-1]
        drain()
        torch.cuda.synchronize()
        for r in range(world):  # every slot of the last turn: offsets sane, and my own slot equals what I encoded
            sh = sg.shard(last, r, k, n, "ans")
            assert int(sh.offsets[0].item()) == 0 and int(sh.offsets[-1].item()) > 0
        sg.sync()
    else:
        assert torch.equal(out, syms), "decode(encode(x)) != x"

    # ---- timed region: device-resident inputs --------------------------------------------------------
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    sync_all()
    for _ in range(3):  # the checks above left the GPU idle for a moment
        step()
    drain()
    sync_all()
    lib.ctr_profile_enable(1)
    lib.ctr_profile_read(0, None, None)
    lib.ctr_profile_read(1, None, None)
    launches0 = B.kernel_launch_count()
    with sampler as clocks:
        t_host0 = time.perf_counter()
        ev[0].record()
        for i in range(args.steps):
            step()
            if i + 1 == args.steps:
                drain()
            ev[i + 1].record()
        host_issue_ms = (time.perf_counter() - t_host0) * 1e3 / args.steps  # time the host needs to ENQUEUE a step
        sync_all()
    launches = B.kernel_launch_count() - launches0
    lib.ctr_profile_enable(0)
    if sg is not None:
        sg.sync()
    total_ms = ev[0].elapsed_time(ev[-1])
    if os.environ.get("CTR_BENCH_DEBUG"):
        per = [ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps)]
        sys.stderr.write(f"rank {rank}: steps(ms) " + " ".join(f"{x:.3f}" for x in per) + f" host_issue {host_issue_ms:.3f}\n")
    t = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    value = world * n / (ms_per_step * 1e-3) / 1e6
    bc.check()

    enc_ms, dec_ms = C.c_double(), C.c_double()
    enc_n, dec_n = C.c_uint64(), C.c_uint64()
    lib.ctr_profile_read(0, C.byref(enc_ms), C.byref(enc_n))
    lib.ctr_profile_read(1, C.byref(dec_ms), C.byref(dec_n))
    enc_kernel_ms = enc_ms.value / max(enc_n.value, 1)
    dec_kernel_ms = dec_ms.value / max(dec_n.value, 1)

    # ---- roofline of the dominant kernel ---------------------------------------------------------------
    peak, peak_src = measured_peak_gbs()
    bytes_per_launch = 4.0 * n + 4.0 * total_words  # symbols in/out + compressed words out/in (same for both kernels)
    dom_name, dom_ms = ("ans_encode_kernel", enc_kernel_ms) if enc_kernel_ms >= dec_kernel_ms else ("ans_decode_kernel", dec_kernel_ms)
    achieved = bytes_per_launch / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else 0.0
    traffic, traffic_src = None, None  # DRAM bytes per launch of that kernel from the committed ncu capture (same workload only)
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            tj = json.load(f)
        if tj["symbols"] == n and tj["streams"] == k:
            traffic = tj["dram_bytes_per_launch"].get(dom_name)
            traffic_src = "committed ncu --set full capture (profiles/traffic.json: %s), not measured in this run" % tj.get("source", "")
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": dom_name, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": bytes_per_launch,
                "kernel_ms": {"ans_encode_kernel": enc_kernel_ms, "ans_decode_kernel": dec_kernel_ms},
                "frac_by_kernel": {"ans_encode_kernel": bytes_per_launch / (enc_kernel_ms * 1e-3) / 1e9 / peak if enc_kernel_ms else None,
                                   "ans_decode_kernel": bytes_per_launch / (dec_kernel_ms * 1e-3) / 1e9 / peak if dec_kernel_ms else None},
                "frac_of_step": {"ans_encode_kernel": enc_kernel_ms / ms_per_step, "ans_decode_kernel": dec_kernel_ms / ms_per_step}}

    # ---- other BASELINE configs, each with its own timing and parity bit ----------------------------------
    extra = None
    if not args.no_extra_configs:
        import bench_configs as BC
        del out
        state["comp"] = None
        torch.cuda.empty_cache()
        extra = {}
        for name, fn in (("configs[3]", BC.baseline_config3_sharded), ("configs[4]", BC.baseline_config4_sharded),
                         ("north_star_1e9", BC.north_star_1e9)):
            try:
                sync_all()
                extra[name] = fn(world, rank)
            except Exception as exc:  # noqa: BLE001 -- an extra config must never take the contract line down
                extra[name] = {"error": f"{type(exc).__name__}: {exc}"}
            torch.cuda.empty_cache()
        out = torch.empty_like(syms)

    # ---- e2e: host buffers through the C ABI, copies inside the timed region ---------------------------
    # One host thread, the library's own pipeline (csrc/host_pipeline.cu): every step starts the encode of the batch
    # (ctr_ans_encode_reverse_host_async: upload-bound) and the decode of the container the previous step produced
    # (ctr_ans_decode_host_async: download-bound) and waits for both, so both directions of the bus are busy.  Every
    # step moves the whole batch up and down: 1e8 symbols encoded and 1e8 symbols decoded per step.
    # `serial_ms_per_step` is the same work with the synchronous calls, one after the other.
    def pinned(nel, dtype):
        return torch.empty(nel, dtype=dtype).pin_memory()

    h_syms = pinned(n, torch.int32)
    h_syms.copy_(syms)
    h_out = pinned(n, torch.int32)
    words_cap = min(cap_words, int(total_words * 1.25) + 4 * k + 1024)
    conts = [dict(words=pinned(words_cap, torch.int32), off=pinned(k + 1, torch.int64), status=C.c_int(), bad=C.c_uint64())
             for _ in range(2)]
    dstatus, dbad = C.c_int(), C.c_uint64()

    def enc_args(ct):
        return (model.handle, h_syms.data_ptr(), n, k, None, None, 0, ct["words"].data_ptr(), words_cap, ct["off"].data_ptr(),
                C.byref(ct["status"]), C.byref(ct["bad"]))

    def dec_args(ct):
        return (model.handle, ct["words"].data_ptr(), ct["off"].data_ptr(), n, k, None, None, 0, h_out.data_ptr(),
                C.byref(dstatus), C.byref(dbad))

    def e2e_serial_step(ct):
        rc = lib.ctr_ans_encode_reverse_host(*enc_args(ct))
        assert rc == 0 and ct["status"].value == 0, (rc, ct["status"].value)
        rc = lib.ctr_ans_decode_host(*dec_args(ct))
        assert rc == 0 and dstatus.value == 0, (rc, dstatus.value)

    def e2e_overlapped_step(i):
        j_enc, j_dec = C.c_void_p(), C.c_void_p()
        assert lib.ctr_ans_decode_host_async(*dec_args(conts[i & 1]), C.byref(j_dec)) == 0  # (its small word uploads go first)
        assert lib.ctr_ans_encode_reverse_host_async(*enc_args(conts[(i + 1) & 1]), C.byref(j_enc)) == 0
        rc_e, rc_d = lib.ctr_host_job_wait(j_enc), lib.ctr_host_job_wait(j_dec)
        assert rc_e == 0 and rc_d == 0 and conts[(i + 1) & 1]["status"].value == 0 and dstatus.value == 0, (rc_e, rc_d)

    e2e = None
    if args.e2e_steps:
        e2e_serial_step(conts[0])
        assert torch.equal(h_out, h_syms)
        e2e_overlapped_step(0)
        h_out.zero_()
        sync_all()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            e2e_serial_step(conts[0])
        serial_s = (time.perf_counter() - t0) / args.e2e_steps
        assert torch.equal(h_out, h_syms)
        h_out.zero_()
        e2e_serial_step(conts[0])  # container 0 is what the first overlapped step decodes
        sync_all()
        t0 = time.perf_counter()
        for i in range(args.e2e_steps):
            e2e_overlapped_step(i)
        e2e_s = (time.perf_counter() - t0) / args.e2e_steps
        assert torch.equal(h_out, h_syms)
        assert torch.equal(conts[0]["off"], conts[1]["off"])
        # what the bus allows: both directions at once, pinned, 4 bytes per symbol + the words each way
        d_a, d_b = torch.empty(n, dtype=torch.int32, device="cuda"), syms
        s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        with torch.cuda.stream(s1):
            d_a.copy_(h_syms, non_blocking=True)
        with torch.cuda.stream(s2):
            h_out.copy_(d_b, non_blocking=True)
        torch.cuda.synchronize()
        duplex_s = time.perf_counter() - t0
        t = torch.tensor([e2e_s, serial_s, duplex_s], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s, serial_s, duplex_s = (float(x) for x in t.tolist())
        words_bytes = 4 * int(conts[0]["off"][-1].item())
        h2d, d2h = 4 * n + words_bytes + 8 * (k + 1), words_bytes + 8 * (k + 1) + 4 * n + 64
        pcie_gbs = 4e-9 * n / duplex_s  # per direction, both directions busy
        bound_s = max(h2d, d2h) / (pcie_gbs * 1e9)
        e2e = {"value": world * n / e2e_s / 1e6, "unit": "Msymbols/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "ms_per_step": e2e_s * 1e3, "steps": args.e2e_steps, "serial_ms_per_step": serial_s * 1e3,
               "serial_value": world * n / serial_s / 1e6,
               "pcie_roofline": {"gbs_per_direction_full_duplex": pcie_gbs, "ms_per_step": bound_s * 1e3,
                                 "value": world * n / bound_s / 1e6, "frac": bound_s / e2e_s,
                                 "how": "one pinned 4n-byte copy per direction at the same time, timed in this run"},
               "numa": {"cpus_near_gpu": numa_cpus},
               "api": "one host thread: ctr_ans_encode_reverse_host_async(batch) + ctr_ans_decode_host_async(previous step's "
                      "container), then ctr_host_job_wait on both (pinned host buffers; the library pipelines every call as "
                      "chunks of streams over 3 CUDA streams); serial_* = the synchronous calls one after the other"}

    if rank == 0:
        cpu = None
        if not args.no_cpu_baseline:
            if all_cpus:
                try:
                    os.sched_setaffinity(0, all_cpus)
                except Exception:
                    pass
            small = world > 1
            cpu = cpu_baseline(min(n, args.cpu_sample // (5 if small else 1)), k, os.cpu_count() or 1, small=small)
        line = {
            "metric": METRIC, "value": value, "unit": "Msymbols/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {"workload": "configs[1]: 1e8 i.i.d. int32 symbols per GPU, QuantizedGaussian(-50,50,3.2,9.6), "
                                   "lane-interleaved rANS (one reference coder per lane), CDF in shared memory",
                       "symbols_per_gpu": n, "streams_per_gpu": k, "compressed_words_per_gpu": total_words,
                       "bits_per_symbol": 32.0 * total_words / n,
                       "l2": "inputs (400 MB symbols) larger than the 126 MB L2; no flush between steps",
                       "step": ("ANS encode (kernel with fused compaction) -> ANS decode" if world == 1 else
                                "ANS encode of the own shard into the own slot of the gathered container -> push to all peers "
                                "(side streams, overlapped with the following kernels) -> ANS decode of the right neighbour's "
                                "shard of the previous turn out of the gathered container; every gather joined inside the timed region"),
                       "gather_lag_steps": args.gather_lag if world > 1 else None,
                       "gather": gather_kind,
                       "parity_note": "erf/exp restate FreeBSD msun (what Rust libm 0.2.16 implements); equality with a real "
                                      "Rust build is pinned by the reference's golden vectors G1-G8 only (SURVEY 8c)"},
            "host_issue_ms_per_step": host_issue_ms,
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches),
            "clocks": clocks.summary(), "extra_configs": extra,
        }
        print(json.dumps(line))
    if sg is not None:
        sync_all()
        sg.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
