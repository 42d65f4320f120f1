#!/usr/bin/env python
"""bench.py — Mreads/s demultiplexed by the matcher core, on BASELINE.json's headline config.

One "step" = one pass of the hot path (barcode bytes in -> result word per read + per-sample counts) over one batch
of synthetic reads of the named config, per GPU.  Prints ONE JSON line on rank 0 (see the contract in the task).

  value      whole-job Mreads/s with the packed batch already resident in HBM (CUDA events, max over ranks,
             includes the final NCCL all-reduce of the per-sample count table when N > 1)
  e2e        same metric through the reference-facing C-ABI call fqtk_b200_matcher_assign_batch with HOST (pinned)
             buffers: ASCII barcode rows in, result words out, H2D/D2H inside the timed region
  roofline   dominant kernel vs the measured HBM copy peak, on ALGORITHMIC bytes B(L) = 4*ceil(L/8) + 4 per read
  cpu_baseline   the oracle (C restatement of the reference's matcher, memo cache on) timed on this box's host cores

  configs    the other single-GPU-sized configs of BASELINE.json next to the headline: cfg 2 (weak, 100 M reads per GPU),
             cfg 4 and cfg 5 (STRONG: 1 B reads sharded over the N ranks), each with its own kernel roofline
  parity_check   outside the timed region: per-rank histogram of the result words, all-reduced, == the all-reduced device
             counts; rank 0 replays a window of ANOTHER rank's shard through the oracle

`--impl reference` times the reference's CPU implementation of the path (the oracle port; the reference is Rust and
cannot be built in this image) on the same workload: each step is a bounded sample of the same read stream.  That arm
loads nothing of the product: panel and reads come from the oracle's own copy of the workload generator.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "Mreads/sec demuxed (dual 8+8bp, 384 samples) at 1/2/4/8 B200 vs ref CPU"
UNIT = "Mreads/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["b200", "reference"], default="b200")
    ap.add_argument("--config", type=int, default=3, help="BASELINE.json configs index + 1 (3 = the headline)")
    ap.add_argument("--reads", type=int, default=0, help="override reads per GPU per step (default: the config's N)")
    ap.add_argument("--e2e-reads", type=int, default=128 << 20, help="reads per GPU per e2e step (pinned host memory)")
    ap.add_argument("--mode", choices=["auto", "table", "brute"], default="auto")
    ap.add_argument("--cuckoo", type=int, default=-1, choices=[-1, 0, 1, 2, 3, 5],
                    help="packed-route kernel knob (fqtk_b200_set_cuckoo_arity): -1 auto, 0 k_probe2, 1 k_probe4, "
                         "2/3 k_probe3 with that many sub-tables; A/B timing")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-brute", action="store_true")
    ap.add_argument("--no-routing", action="store_true")
    ap.add_argument("--no-bgzf", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the cfg 2 / 4 / 5 entries of the line")
    ap.add_argument("--single-process", action="store_true",
                    help="ONE process driving --gpus N devices through the C ABI's group handle (fqtk_b200_group_*) "
                         "instead of one torchrun rank per GPU")
    ap.add_argument("--no-parity-check", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=10.0, help="target CPU seconds for the cpu_baseline sample")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------------------------
def host_panel(cfg):
    """The config's panel from the ORACLE's copy of the workload generator (byte-identical to fqtk_b200.synth.panel;
    tests/test_oracle_kats.py checks it) — the CPU legs load nothing of the product."""
    import oracle

    return oracle.synth_panel(cfg.seed_panel, cfg.n_samples, cfg.barcode_len, cfg.min_distance, cfg.n_degenerate)


def host_reads(panel, seed, first, n, threads=16):
    """Host replay of the synthetic stream, sliced over a few threads (ctypes releases the GIL)."""
    import oracle

    L = panel.shape[1]
    out = np.empty((n, L), dtype=np.uint8)
    threads = max(1, min(threads, os.cpu_count() or 1, n // 100_000 + 1))
    bounds = [n * t // threads for t in range(threads + 1)]

    def work(t):
        lo, hi = bounds[t], bounds[t + 1]
        if hi > lo:
            out[lo:hi] = oracle.synth_reads(panel, seed, first + lo, hi - lo)

    with ThreadPoolExecutor(threads) as ex:
        list(ex.map(work, range(threads)))
    return out


def time_oracle(cfg, panel, reads, threads):
    """Mreads/s of the oracle on `reads` (1 thread = the reference's model; >1 = OpenMP upper bound)."""
    import oracle

    bcs = [bytes(r) for r in panel]
    if threads == 1:
        m = oracle.OracleMatcher(bcs, cfg.max_mismatches, cfg.min_mismatch_delta, use_cache=True)
        m.assign_batch(reads[: min(len(reads), 200_000)], want_results=False)  # touch code + warm the memo cache
        t0 = time.perf_counter()
        _, counts = m.assign_batch(reads, want_results=False)
        dt = time.perf_counter() - t0
        return len(reads) / dt / 1e6, 1, counts
    t0 = time.perf_counter()
    _, counts, used = oracle.assign_batch_mt(panel, cfg.max_mismatches, cfg.min_mismatch_delta, reads,
                                             threads=threads, want_results=False)
    dt = time.perf_counter() - t0
    return len(reads) / dt / 1e6, used, counts


def cpu_baseline(cfg, panel, seconds, all_cores=True):
    import oracle

    calib = host_reads(panel, cfg.seed_reads, 0, 1_000_000)
    rate, _, _ = time_oracle(cfg, panel, calib, 1)
    n = int(min(max(rate * 1e6 * seconds, 2_000_000), 96_000_000, cfg.n_reads))
    reads = host_reads(panel, cfg.seed_reads, 0, n)
    v1, _, counts1 = time_oracle(cfg, panel, reads, 1)
    out = {
        "value": round(v1, 3), "unit": UNIT, "cores": 1, "kind": "port",
        "sample": f"first {n} reads of the same synthetic stream, held in host memory; oracle/fqtk_oracle.c literal "
                  f"restatement, memo cache on, 1 thread = the reference's threading model for the matcher "
                  f"(demux.rs:945-977 is serial)",
    }
    if all_cores:
        cores = os.cpu_count() or oracle.max_threads()  # explicit: torchrun exports OMP_NUM_THREADS=1
        vN, used, countsN = time_oracle(cfg, panel, reads, cores)
        assert np.array_equal(counts1, countsN)
        out["all_cores_upper_bound"] = {
            "value": round(vN, 3), "cores": used,
            "note": "reads sharded over all host threads with a private matcher + cache each (OpenMP); the reference "
                    "does NOT do this — upper bound only"}
    return out


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu_index = gpu_index
        self.proc = None
        self.path = f"/tmp/fqtk_b200_clocks_{os.getpid()}.csv"

    def start(self):
        try:
            self.fh = open(self.path, "w")
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.gpu_index)], stdout=self.fh, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.fh.close()
        sm, smax, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
                power.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        try:
            os.remove(self.path)
        except OSError:
            pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        busy = [s for s, p in zip(sm, power) if p >= 0.5 * max(power)] or sm
        return {"sm_mhz": statistics.median(busy), "sm_max_mhz": max(smax), "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": max(power)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(cfg_id, kernel, n_launch):
    """dram bytes per launch of the dominant kernel from the committed ncu capture, if one exists, scaled to the number of
    reads THIS launch processes (a strong-scaling shard is a fraction of the captured launch)."""
    p = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            key = f"cfg{cfg_id}_{kernel.split('<')[0]}"
            if key not in d:
                return None
            return int(d[key] * n_launch / d.get(key + "_reads", n_launch))
        except Exception:
            return None
    return None


# ---------------------------------------------------------------------------------------------------------------
def run_reference(args):
    """The reference arm: the reference's own CPU algorithm for the path (oracle port), bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from fqtk_b200.synth import CONFIGS  # plain dataclasses: importing them does not load the CUDA library

    cfg = CONFIGS[args.config]
    panel = host_panel(cfg)
    total_steps = args.steps + args.warmup
    calib = host_reads(panel, cfg.seed_reads, 0, 1_000_000)
    rate, _, _ = time_oracle(cfg, panel, calib, 1)
    per_step_s = min(10.0, max(0.5, 120.0 / max(1, total_steps)))
    n = int(min(max(rate * 1e6 * per_step_s, 1_000_000), 96_000_000, cfg.n_reads))
    reads = host_reads(panel, cfg.seed_reads, 0, n)
    import oracle

    bcs = [bytes(r) for r in panel]
    m = oracle.OracleMatcher(bcs, cfg.max_mismatches, cfg.min_mismatch_delta, use_cache=True)
    for _ in range(args.warmup):
        m.assign_batch(reads, want_results=False)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        m.assign_batch(reads, want_results=True)
    dt = time.perf_counter() - t0
    value = n * args.steps / dt / 1e6
    cores_all = os.cpu_count() or oracle.max_threads()  # explicit: torchrun exports OMP_NUM_THREADS=1
    vN, used, _ = time_oracle(cfg, panel, reads, cores_all)
    sample = (f"each step = the first {n} reads of the config's synthetic stream (of {cfg.n_reads}), in host memory; "
              f"literal C restatement of BarcodeMatcher::assign with memo cache on, 1 thread (the reference's matcher "
              f"is serial, demux.rs:945-977)")
    line = {
        "impl": "reference", "metric": METRIC, "value": round(value, 3), "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(dt / args.steps * 1e3, 3),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": dict(config_dict(cfg, n, "reference-cpu"), input="ASCII barcode rows in host memory",
                       sample_of_config_reads=cfg.n_reads,
                       note=f"a CPU rate: every step is a {n}-read sample of the config's {cfg.n_reads}-read stream"),
        "cpu_baseline": {"value": round(value, 3), "unit": UNIT, "cores": 1, "kind": "port", "sample": sample,
                         "all_cores_upper_bound": {"value": round(vN, 3), "cores": used,
                                                   "note": "OpenMP-sharded, one matcher + cache per thread; NOT what "
                                                           "the reference does"}},
        "e2e": {"value": round(value, 3), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def config_dict(cfg, n_reads, mode):
    return {
        "workload": cfg.name, "read_structures": cfg.read_structures, "n_samples": cfg.n_samples,
        "barcode_len": cfg.barcode_len, "max_mismatches": cfg.max_mismatches,
        "min_mismatch_delta": cfg.min_mismatch_delta, "reads_per_gpu_per_step": n_reads, "kernel_mode": mode,
        "input": "4-bit packed barcode words (BitEnc layout), HBM-resident", "l2": "inputs larger than L2 (no flush)",
    }


def pin_to_gpu_numa_node(local):
    """Bind this rank to the CPUs next to its GPU BEFORE any pinned buffer is allocated, so that first-touch puts the
    e2e buffers on the GPU's NUMA node.  Reports what happened instead of swallowing it."""
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local)
        n_words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, n_words)
        cpus = [64 * w + b for w, word in enumerate(mask) for b in range(64) if (int(word) >> b) & 1]
        if not cpus:
            return {"pinned": False, "why": "NVML reports no CPU affinity for this GPU"}
        os.sched_setaffinity(0, cpus)
        return {"pinned": True, "cpus": len(cpus), "first_cpu": cpus[0], "last_cpu": cpus[-1]}
    except Exception as e:  # noqa: BLE001 — report, do not hide
        sys.stderr.write(f"[bench] NUMA pinning of rank {local} failed: {e!r}\n")
        return {"pinned": False, "why": repr(e)}


def kernel_name(info, mode, W):
    if mode != "table":
        return "k_brute_sliced" if W <= 4 else "k_brute_long"
    if int(info.cuckoo_probes) and W <= 2 and int(info.l2_table_entries):
        return f"k_probe5<W={W},NP={int(info.cuckoo_probes)}>"
    if int(info.cuckoo_probes) and W <= 2:
        return f"k_probe3<W={W},NP={int(info.cuckoo_probes)}>"
    if int(info.l2_table_entries):
        return f"k_probe4<W={W}>"
    return f"k_probe2<W={W}>"


class Ctx:
    """What every measurement needs: torch, the process group, this rank's device and stream, the roofline peak."""

    def __init__(self, torch, dist, world, rank, local, dev, stream):
        self.torch, self.dist, self.world, self.rank, self.local, self.dev, self.stream = torch, dist, world, rank, local, dev, stream
        self.peak, self.peak_src = measured_peak()

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
            self.torch.cuda.synchronize()

    def max_over_ranks(self, x):
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def all_ok(self, ok):
        t = self.torch.tensor([1 if ok else 0], dtype=self.torch.int32, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MIN)
        return bool(t.item())


def parity_check(ctx, cfg, matcher, panel, d_packed, d_res, n_local, first_local, shard_of):
    """Outside the timed region.  (i) every rank histograms ITS result words on the device; the all-reduced histogram must
    equal the all-reduced device count table of one pass (a wrong shard offset or a double reduce that kept the sum would
    show).  (ii) rank 0 replays a strided 20 k-read window of ANOTHER rank's shard — regenerated on the host from the
    stream's definition — through the oracle and compares it with the result words that rank computed."""
    import oracle
    from fqtk_b200.distributed import all_reduce_counts, tensor_from_device_ptr

    torch, dist = ctx.torch, ctx.dist
    S = cfg.n_samples
    matcher.reset_counts()
    matcher.assign_packed_device(d_packed.data_ptr(), n_local, d_res.data_ptr(), ctx.stream)
    torch.cuda.synchronize()
    counts = tensor_from_device_ptr(matcher.counts_device_ptr(), S + 1, ctx.dev).clone()
    idx = torch.where(d_res == -1, torch.full_like(d_res, S), (d_res >> 16) & 0xFFFF)
    hist = torch.bincount(idx.to(torch.int64), minlength=S + 1)
    del idx
    local_ok = bool(torch.equal(hist, counts))
    all_reduce_counts(counts)
    all_reduce_counts(hist)
    ok = bool(torch.equal(hist, counts)) and local_ok
    # window replay: the source rank is the next one (rank 0 itself when alone)
    span = 20_000
    src_rank = 1 % ctx.world
    src_first, src_n = shard_of(src_rank)
    off = min(max(src_n // 3 + 17, 0), max(src_n - span, 0))
    span = min(span, src_n)
    window = torch.empty(span, dtype=torch.int32, device=ctx.dev)
    if ctx.rank == src_rank:
        window.copy_(d_res[off:off + span])
    if ctx.world > 1:
        dist.broadcast(window, src=src_rank)
    replay = None
    if ctx.rank == 0:
        reads = host_reads(panel, cfg.seed_reads, src_first + off, span)
        om = oracle.OracleMatcher([bytes(r) for r in panel], cfg.max_mismatches, cfg.min_mismatch_delta, use_cache=True)
        want, _ = om.assign_batch(reads)
        replay = bool(np.array_equal(window.cpu().numpy().view(np.uint32), want))
        ok = ok and replay
    ok = ctx.all_ok(ok)
    matcher.reset_counts()
    return {"ranks": ctx.world, "ok": ok, "reads_counted": int(counts.sum().item()),
            "window": {"of_rank": src_rank, "first_read": int(src_first + off), "reads": int(span)},
            "what": "all-reduced histogram of every rank's result words == all-reduced device counts; rank 0 replays a "
                    "window of another rank's shard through the CPU oracle"}


def measure_config(ctx, args, cfg_id, steps, warmup, strong, cuckoo=-1, mode="auto", keep=False, check=True):
    """One config on this rank's shard: device-timed passes over an HBM-resident packed batch + the count all-reduce.
    `strong`: the config's N reads are split over the ranks (contiguous shards); else every rank takes N (weak)."""
    from fqtk_b200 import BarcodeMatcher, synth
    from fqtk_b200.barcode_matching import kernel_launches
    from fqtk_b200.distributed import all_reduce_counts, shard_bounds, tensor_from_device_ptr, weak_shard_first_read

    torch, dist, dev, world, rank = ctx.torch, ctx.dist, ctx.dev, ctx.world, ctx.rank
    cfg = synth.CONFIGS[cfg_id]
    n_cfg = args.reads if (args.reads and cfg_id == args.config) else cfg.n_reads
    if strong:
        def shard_of(r):
            lo, hi = shard_bounds(n_cfg, r, world)
            return lo, hi - lo
    else:
        def shard_of(r):
            return weak_shard_first_read(n_cfg, r), n_cfg
    first, n = shard_of(rank)
    n_total = n_cfg if strong else n_cfg * world
    W = cfg.words_per_read
    panel = synth.panel(cfg)
    bcs = [bytes(r) for r in panel]
    d_packed = torch.empty((n, W), dtype=torch.int32, device=dev)
    d_res = torch.empty(n, dtype=torch.int32, device=dev)
    synth.reads_device(panel, cfg.seed_reads, first, n, 0, d_packed.data_ptr(), ctx.stream)
    opts = {} if cuckoo == -1 else {"kernel": cuckoo}
    matcher = BarcodeMatcher(bcs, cfg.max_mismatches, cfg.min_mismatch_delta, use_cache=(mode != "brute"),
                             device=ctx.local, **opts)
    if mode == "table" and matcher.mode != "table":
        raise SystemExit("memo table could not be built for this config")
    info = matcher.info()
    kernel = kernel_name(info, matcher.mode, W)
    counts_t = torch.zeros(cfg.n_samples + 1, dtype=torch.int64, device=dev)

    def one_pass():
        matcher.assign_packed_device(d_packed.data_ptr(), n, d_res.data_ptr(), ctx.stream)

    for _ in range(max(warmup, 0)):
        one_pass()
    ctx.barrier()
    matcher.reset_counts()
    launches0 = kernel_launches()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 2)]
    ctx.barrier()
    evs[0].record()
    for k in range(steps):
        one_pass()
        evs[k + 1].record()
    # the single collective of the path: the final per-sample count table (S+1 u64) summed over NVLink
    # (int64 view of the matcher's own device counters; same stream ordering, no host round trip)
    counts_t.copy_(tensor_from_device_ptr(matcher.counts_device_ptr(), cfg.n_samples + 1, dev))
    if world > 1:
        all_reduce_counts(counts_t)
    evs[steps + 1].record()
    ctx.barrier()
    launches = kernel_launches() - launches0
    total_ms = ctx.max_over_ranks(evs[0].elapsed_time(evs[steps + 1]))
    k_ms_local = statistics.mean(evs[k].elapsed_time(evs[k + 1]) for k in range(steps))
    k_ms = ctx.max_over_ranks(k_ms_local)
    counts = counts_t.cpu().numpy()
    assert int(counts.sum()) == n_total * steps, "per-sample counts must add up to every read processed"
    value = n_total * steps / (total_ms * 1e-3) / 1e6
    # roofline of the dominant kernel: this rank's launches, algorithmic bytes only, slowest rank's mean launch time
    bytes_per_launch = n * cfg.algorithmic_bytes_per_read
    achieved = bytes_per_launch / (k_ms * 1e-3) / 1e9
    roofline = {
        "bound": "hbm", "achieved": round(achieved, 2), "peak": ctx.peak, "unit": "GB/s",
        "frac": round(achieved / ctx.peak, 4), "traffic": ncu_traffic(cfg_id, kernel, n), "kernel": kernel,
        "kernel_ms": round(k_ms, 4), "algorithmic_bytes_per_read": cfg.algorithmic_bytes_per_read,
        "algorithmic_bytes_per_launch": bytes_per_launch, "peak_source": ctx.peak_src,
    }
    out = {
        "cfg": cfg, "n": n, "n_total": n_total, "first": first, "value": value, "total_ms": total_ms, "roofline": roofline,
        "launches": int(launches), "counts": counts, "info": info, "mode": matcher.mode, "kernel": kernel,
        "panel": panel, "shard_of": shard_of,
    }
    if check and not args.no_parity_check:
        out["parity_check"] = parity_check(ctx, cfg, matcher, panel, d_packed, d_res, n, first, shard_of)
    if keep:
        out.update(matcher=matcher, d_packed=d_packed, d_res=d_res)
    else:
        matcher.close()
        del d_packed, d_res
        torch.cuda.empty_cache()
    return out


def config_entry(m, steps, scaling):
    """One entry of the line's `configs` object."""
    cfg = m["cfg"]
    return {
        "workload": cfg.name, "n_reads": m["n_total"], "reads_per_gpu_per_step": m["n"], "scaling": scaling,
        "steps": steps, "value": round(m["value"], 2), "unit": UNIT, "ms": round(m["total_ms"] / steps, 4),
        "roofline": m["roofline"], "gpu_launches": m["launches"],
        "matched_fraction": round(1.0 - float(m["counts"][-1]) / float(m["counts"].sum()), 5),
        "parity_check": m.get("parity_check"),
    }


def host_pack_h2d_bytes(ne, L, W):
    """Bytes one rank's host-pack call moves host -> device: the library's own chunking rule (capi.cu
    assign_batch_host_pack): `mix` of every 8 chunks go as ASCII rows, the others as packed words."""
    mix = min(7, max(0, int(os.environ.get("FQTK_B200_HOST_PACK_MIX", "3"))))
    chunk_bytes = int(os.environ.get("FQTK_B200_CHUNK_MB", "0")) << 20 or (32 << 20)
    row = W * 4
    chunk = min(max(chunk_bytes // (4 * row), 65536) & ~3, ne)
    total = 0
    for c in range((ne + chunk - 1) // chunk):
        cnt = min(chunk, ne - c * chunk)
        total += cnt * (L if (c * mix) % 8 + mix >= 8 else row)
    return total


def measure_e2e(ctx, args, cfg, panel, matcher, d_packed, d_res, first, n):
    """The same metric through the reference-facing C-ABI calls on pinned HOST buffers, copies inside the timed region
    (wall clock around the synchronous calls, max over ranks), for both wire formats:
      ascii   fqtk_b200_matcher_assign_batch: L ASCII bytes per read in, 4-byte result words out (the drop-in form)
      packed  fqtk_b200_matcher_assign_batch_packed: the reference's own BitEnc words in (4*W bytes), u16 sample index out
    each next to the platform's ceiling for exactly those bytes (fqtk_b200_copy_ceiling: the same chunked pinned copies
    with no kernel in between, all ranks at once)."""
    import ctypes as C

    import psutil

    from fqtk_b200 import _lib, synth

    torch, world = ctx.torch, ctx.world
    lib = _lib.lib()
    L, W = cfg.barcode_len, cfg.words_per_read
    ne = min(args.e2e_reads, n)
    avail = psutil.virtual_memory().available
    while ne * (L + 4 * W + 8) * max(1, min(world, 8)) > 0.25 * avail and ne > (1 << 20):
        ne //= 2
    ne &= ~3

    def pinned(nbytes):
        p = C.c_void_p()
        _lib.check(lib.fqtk_b200_host_alloc(C.byref(p), nbytes))
        return p

    h_in, h_out, h_pk, h_idx = pinned(ne * L), pinned(ne * 4), pinned(ne * W * 4), pinned(ne * 2)
    h_in_np = np.ctypeslib.as_array(C.cast(h_in, C.POINTER(C.c_uint8)), shape=(ne, L))
    h_out_np = np.ctypeslib.as_array(C.cast(h_out, C.POINTER(C.c_uint32)), shape=(ne,))
    h_pk_np = np.ctypeslib.as_array(C.cast(h_pk, C.POINTER(C.c_uint32)), shape=(ne, W))
    h_idx_np = np.ctypeslib.as_array(C.cast(h_idx, C.POINTER(C.c_uint16)), shape=(ne,))
    d_ascii = torch.empty((ne, L), dtype=torch.uint8, device=ctx.dev)
    synth.reads_device(panel, cfg.seed_reads, first, ne, d_ascii.data_ptr(), 0, ctx.stream)
    h_in_np[:] = d_ascii.cpu().numpy()
    del d_ascii
    t0 = time.perf_counter()
    _lib.check(lib.fqtk_b200_pack_host(h_in.value, ne, L, L, h_pk.value, 0))  # the host's own encode() of the rows
    pack_s = time.perf_counter() - t0

    def timed(call):
        matcher.reset_counts()
        for _ in range(2):
            call()
        matcher.reset_counts()
        ctx.barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            call()  # synchronous: returns with the results on the host
        torch.cuda.synchronize()
        return ctx.max_over_ranks(time.perf_counter() - t0) / args.steps

    def ceiling(in_bytes, out_bytes):
        sec = C.c_double()
        ctx.barrier()
        _lib.check(lib.fqtk_b200_copy_ceiling(ctx.local, in_bytes, out_bytes, 32 << 20, 3, C.byref(sec)))
        return ctx.max_over_ranks(sec.value)

    dt_ascii = timed(lambda: matcher.assign_batch_ptr(h_in.value, ne, L, h_out.value))
    # the host results must be the device path's results for the same reads
    chk = min(ne, 1 << 20)
    matcher.reset_counts()
    matcher.assign_packed_device(d_packed.data_ptr(), chk, d_res.data_ptr(), ctx.stream)
    torch.cuda.synchronize()
    want = d_res[:chk].cpu().numpy().view(np.uint32)
    assert np.array_equal(want, h_out_np[:chk]), "e2e results differ"
    # the same call with encode() done by host threads while the batch is in flight (fqtk_b200_matcher_set_host_pack): the
    # packed words cross PCIe.  Tried where host cores and host memory bandwidth are to spare: one or two ranks per box.
    dt_hp, hp_threads = None, 0
    if world <= 2 or os.environ.get("FQTK_B200_BENCH_HOST_PACK"):
        hp_threads = max(1, min(16, len(os.sched_getaffinity(0)) // world))
        matcher.set_host_pack(hp_threads)
        h_out_np[:chk] = 0
        dt_hp = timed(lambda: matcher.assign_batch_ptr(h_in.value, ne, L, h_out.value))
        matcher.set_host_pack(0)
        assert np.array_equal(want, h_out_np[:chk]), "host-pack e2e results differ"
    dt_packed = timed(lambda: _lib.check(lib.fqtk_b200_matcher_assign_batch_packed(
        matcher._h, h_pk.value, ne, None, h_idx.value)))
    assert np.array_equal(np.where(want == 0xFFFFFFFF, 0xFFFF, want >> 16).astype(np.uint16), h_idx_np[:chk]), \
        "packed e2e results differ"
    assert np.array_equal(h_pk_np[:chk], synth.pack_host(h_in_np[:chk])), "host pack differs from encode()"
    ceil_ascii = ceiling(ne * L, ne * 4)
    ceil_packed = ceiling(ne * W * 4, ne * 2)
    hp_in = host_pack_h2d_bytes(ne, L, W)
    ceil_hp = ceiling(hp_in, ne * 4) if dt_hp is not None else None

    def mreads(sec):
        return round(ne * world / sec / 1e6, 2)

    out = {
        "value": mreads(dt_ascii), "unit": UNIT, "h2d_bytes_per_step": ne * L * world, "d2h_bytes_per_step": ne * 4 * world,
        "reads_per_gpu_per_step": ne, "ms_per_step": round(dt_ascii * 1e3, 3),
        "api": "fqtk_b200_matcher_assign_batch (ASCII rows in pinned host memory -> result words in pinned host memory; "
               "chunked H2D / kernel / D2H overlap inside the call)",
        "ceiling": mreads(ceil_ascii), "frac_of_ceiling": round(ceil_ascii / dt_ascii, 4),
        "ceiling_what": "fqtk_b200_copy_ceiling: the same bytes through the same chunked pinned copies, no kernel, all "
                        "ranks at once (tools/h2d_ceiling.cu is the stand-alone form)",
        "h2d_gb_per_s_per_gpu": round(ne * L / dt_ascii / 1e9, 2),
        "packed": {
            "value": mreads(dt_packed), "unit": UNIT, "h2d_bytes_per_step": ne * W * 4 * world,
            "d2h_bytes_per_step": ne * 2 * world, "ms_per_step": round(dt_packed * 1e3, 3),
            "api": "fqtk_b200_matcher_assign_batch_packed (the reference's BitEnc words in pinned host memory -> u16 "
                   "sample indices); the host's encode() of the rows (fqtk_b200_pack_host) is NOT in the timed region",
            "ceiling": mreads(ceil_packed), "frac_of_ceiling": round(ceil_packed / dt_packed, 4),
            "host_pack_mreads_per_s": round(ne / pack_s / 1e6, 1), "host_pack_threads": os.cpu_count(),
        },
    }
    if dt_hp is not None:
        hp = {"value": mreads(dt_hp), "unit": UNIT, "h2d_bytes_per_step": hp_in * world, "d2h_bytes_per_step": ne * 4 * world,
              "ms_per_step": round(dt_hp * 1e3, 3), "host_threads_per_rank": hp_threads,
              "api": "fqtk_b200_matcher_assign_batch after fqtk_b200_matcher_set_host_pack: the same ASCII rows in pinned host "
                     "memory -> result words in pinned host memory; encode() by host threads (AVX2) inside the call, inside the "
                     "timed region: 5 of every 8 chunks cross PCIe as the reference's BitEnc words, 3 as ASCII rows (encode() "
                     "in the kernel) so that host memory and PCIe are both loaded",
              "ceiling": mreads(ceil_hp), "frac_of_ceiling": round(ceil_hp / dt_hp, 4),
              "h2d_gb_per_s_per_gpu": round(hp_in / dt_hp / 1e9, 2)}
        plain = {k: out[k] for k in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step", "ms_per_step", "api", "ceiling",
                                     "frac_of_ceiling", "h2d_gb_per_s_per_gpu")}
        if dt_hp < dt_ascii:  # the faster form of the SAME call (ASCII rows in, result words out) is the headline
            out.update(hp)
            out["ascii_dma"] = plain
        else:
            out["host_packed"] = hp
    for p in (h_in, h_out, h_pk, h_idx):
        lib.fqtk_b200_host_free(p)
    matcher.reset_counts()
    return out


def measure_bgzf(ctx, args):
    """The output side (SURVEY 8f next #4 tail): BGZF compression of FASTQ text, the reference's pooled writers
    (demux.rs:755-798, level 5).  Device-resident figure (one 64 MiB chunk, CUDA events), the host-buffer call on pinned
    memory (H2D + kernels + D2H overlapped), and zlib level 5 on one host thread beside them (the reference runs libdeflate on
    a thread pool; this is the oracle's restatement, a bounded sample)."""
    import ctypes as C

    import oracle.bgzf as ob
    from fqtk_b200 import _lib
    from fqtk_b200.bgzf import BgzfCompressor

    torch = ctx.torch
    lib = _lib.lib()
    rng = np.random.default_rng(12345)
    n_rec, rl = 40_000, 150  # ~14 MB of distinct FASTQ text, repeated: BGZF blocks are independent, so repeats do not help
    seqs = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=(n_rec, rl))]
    seqs[rng.random((n_rec, rl)) < 0.002] = ord("N")
    quals = np.frombuffer(b"F:,#", dtype=np.uint8)[np.minimum(3, rng.geometric(0.75, size=(n_rec, rl)) - 1)]
    recs = []
    for i in range(n_rec):
        recs.append(b"@A00123:45:HXXXXXX:1:%d:%d:%d 1:N:0:ACGTACGT+TTGCAATC\n" % (1101 + i // 9000, 1000 + (i * 37) % 30000, 1000 + (i * 91) % 35000))
        recs.append(seqs[i].tobytes() + b"\n+\n" + quals[i].tobytes() + b"\n")
    unit = b"".join(recs)
    with BgzfCompressor(ctx.local) as z:
        chunk = z.chunk_bytes
        text = (unit * (chunk // len(unit) + 1))[:chunk]
        d_in = torch.frombuffer(bytearray(text), dtype=torch.uint8).to(ctx.dev)
        d_out = torch.empty(z.bound(chunk), dtype=torch.uint8, device=ctx.dev)
        d_n = torch.zeros(1, dtype=torch.int64, device=ctx.dev)
        stream = torch.cuda.current_stream().cuda_stream
        for _ in range(2):
            z.compress_device(d_in.data_ptr(), chunk, d_out.data_ptr(), d_out.numel(), d_n.data_ptr(), 5, stream)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 5
        e0.record()
        for _ in range(reps):
            z.compress_device(d_in.data_ptr(), chunk, d_out.data_ptr(), d_out.numel(), d_n.data_ptr(), 5, stream)
        e1.record()
        torch.cuda.synchronize()
        dms = e0.elapsed_time(e1) / reps
        out_n = int(d_n.item())
        img = d_out[:out_n].cpu().numpy().tobytes()
        payload, sizes = ob.parse(img)  # every member inflates to its piece (outside the timed region)
        assert payload == text and len(sizes) == (chunk + 65279) // 65280
        del d_in, d_out
        # host-buffer call: 4 chunks of pinned text in, the file image out
        n_host = 4 * chunk
        pin = []
        for nbytes in (n_host, z.bound(n_host)):
            q = C.c_void_p()
            _lib.check(lib.fqtk_b200_host_alloc(C.byref(q), nbytes))
            pin.append(q)
        src = np.ctypeslib.as_array(C.cast(pin[0], C.POINTER(C.c_uint8)), shape=(n_host,))
        dst = np.ctypeslib.as_array(C.cast(pin[1], C.POINTER(C.c_uint8)), shape=(z.bound(n_host),))
        src[:] = np.frombuffer((text * 4)[:n_host], dtype=np.uint8)
        z.compress_into(src, dst, 5)
        t0 = time.perf_counter()
        hreps = 3
        for _ in range(hreps):
            n_img = z.compress_into(src, dst, 5)
        hs = (time.perf_counter() - t0) / hreps
        for q in pin:
            lib.fqtk_b200_host_free(q)
    sample = text[: 24 << 20]
    t0 = time.perf_counter()
    ref_img = ob.compress(sample, 5)
    cs = time.perf_counter() - t0
    return {"what": "BGZF (65 280-byte members) of synthetic FASTQ text, compression level 5",
            "kernels": "k_bgzf_deflate + k_bgzf_scan + k_bgzf_gather",
            "device": {"input_bytes": chunk, "ms": round(dms, 4), "gb_per_s_in": round(chunk / (dms * 1e-3) / 1e9, 2),
                       "output_bytes": out_n, "ratio": round(chunk / out_n, 3),
                       "roofline_frac": round((chunk + out_n) / (dms * 1e-3) / 1e9 / ctx.peak, 5)},
            "host_call": {"api": "fqtk_b200_bgzf_compress (pinned text in, file image out, H2D / kernels / D2H overlapped)",
                          "input_bytes": n_host, "output_bytes": n_img, "ms": round(hs * 1e3, 3),
                          "gb_per_s_in": round(n_host / hs / 1e9, 2)},
            "cpu_zlib_level5_1_thread": {"input_bytes": len(sample), "seconds": round(cs, 3),
                                         "gb_per_s_in": round(len(sample) / cs / 1e9, 4),
                                         "ratio": round(len(sample) / len(ref_img), 3),
                                         "note": "oracle/bgzf.py (zlib; the reference uses libdeflate on its writer thread pool)"},
            "members_checked": len(sizes)}


def measure_gpu_demux(ctx, args):
    """The whole per-batch data path on the device (fqtk_b200.gpu_demux): paired-end 150 + dual 8 bp index FASTQ text of the
    headline panel in pinned host memory -> scan, match, route, records with rewritten headers, BGZF per sample -> the
    compressed images back on the host.  Wall clock around the call, PCIe copies inside."""
    import oracle.bgzf as ob
    from fqtk_b200 import BarcodeMatcher, synth
    from fqtk_b200.bgzf import BgzfCompressor
    from fqtk_b200.gpu_demux import demux_fastq_batch_gpu

    cfg = synth.CONFIGS[3]
    panel = synth.panel(cfg)
    bcs = [bytes(r).decode() for r in panel]
    ids = [f"S{j:03d}" for j in range(len(bcs))]
    n, rl = 400_000, 150
    reads = host_reads(panel, cfg.seed_reads, 0, n)
    rng = np.random.default_rng(99)
    head = np.frombuffer(b"@A00123:45:HXXXXXXXX:1:1101:", dtype=np.uint8)

    def fastq(seqs, comment):
        w = seqs.shape[1]
        coords = np.frombuffer(b"".join(b"%05d:%05d" % (1000 + i % 30000, 1000 + (i * 7) % 30000) for i in range(n)), dtype=np.uint8).reshape(n, 11)
        cm = np.frombuffer(comment, dtype=np.uint8)
        rec = np.empty((n, head.size + 11 + cm.size + 1 + w + 3 + w + 1), dtype=np.uint8)
        p0 = 0
        for piece in (head, coords, cm, b"\n", seqs, b"\n+\n", None, b"\n"):
            if piece is None:
                piece = np.frombuffer(b"F:,#", dtype=np.uint8)[np.minimum(3, rng.geometric(0.75, size=(n, w)) - 1)]
            arr = np.frombuffer(piece, dtype=np.uint8) if isinstance(piece, bytes) else piece
            rec[:, p0:p0 + arr.shape[-1]] = arr
            p0 += arr.shape[-1]
        return rec.reshape(-1)

    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    texts = [fastq(acgt[rng.integers(0, 4, size=(n, rl))], b" 1:N:0:0"), fastq(acgt[rng.integers(0, 4, size=(n, rl))], b" 2:N:0:0"),
             fastq(reads[:, :8], b" 1:N:0:0"), fastq(reads[:, 8:16], b" 2:N:0:0")]
    in_bytes = sum(int(t.size) for t in texts)
    import ctypes as C

    from fqtk_b200 import _lib
    pinned = []
    for k, t in enumerate(texts):  # the reader's chunk buffers live in pinned memory
        q = C.c_void_p()
        _lib.check(_lib.lib().fqtk_b200_host_alloc(C.byref(q), int(t.size)))
        pinned.append(q)
        v = np.ctypeslib.as_array(C.cast(q, C.POINTER(C.c_uint8)), shape=(int(t.size),))
        v[:] = t
        texts[k] = v
    with BarcodeMatcher(bcs, cfg.max_mismatches, cfg.min_mismatch_delta, device=ctx.local) as m, BgzfCompressor(ctx.local) as z:
        res = demux_fastq_batch_gpu(m, z, ids, bcs, ["+T", "+T", "8B", "8B"], texts, ["T"], device=ctx.local)
        m.reset_counts()
        ctx.torch.cuda.synchronize()
        t0 = time.perf_counter()
        reps = 3
        for _ in range(reps):
            res = demux_fastq_batch_gpu(m, z, ids, bcs, ["+T", "+T", "8B", "8B"], texts, ["T"], device=ctx.local)
            m.reset_counts()
        sec = (time.perf_counter() - t0) / reps
        # the same batch through the one-call C form (no Python between the device steps)
        from fqtk_b200.gpu_demux import demux_chunks
        one, used = demux_chunks(m, z, ids, bcs, ["+T", "+T", "8B", "8B"], texts, ["T"])
        m.reset_counts()
        assert one.files == res.files and used == [int(t.size) for t in texts]
        t0 = time.perf_counter()
        for _ in range(reps):
            (buf, offs, cnts), used = demux_chunks(m, z, ids, bcs, ["+T", "+T", "8B", "8B"], texts, ["T"], raw=True)
            m.reset_counts()
        sec_one = (time.perf_counter() - t0) / reps
        assert int(cnts.sum()) == n and int(offs[-1]) == sum(len(v) - 28 for v in one.files.values())
        want_offs = offs.copy()
    # two lanes (two matcher / compressor handles, two host threads, batches alternating): one batch's copies overlap the
    # other's kernels; results are taken in batch order
    from fqtk_b200.gpu_demux import DemuxLanes
    with DemuxLanes(lambda: (BarcodeMatcher(bcs, cfg.max_mismatches, cfg.min_mismatch_delta, device=ctx.local), BgzfCompressor(ctx.local)),
                    lanes=int(os.environ.get("FQTK_B200_BENCH_LANES", "2"))) as lanes:
        nb = 12
        for (buf, offs, cnts), used in lanes.map([texts] * (2 * len(lanes.lanes)), ids, bcs, ["+T", "+T", "8B", "8B"], ["T"], raw=True):
            pass  # warm-up: both lanes' buffers
        t0 = time.perf_counter()
        for (buf, offs, cnts), used in lanes.map([texts] * nb, ids, bcs, ["+T", "+T", "8B", "8B"], ["T"], raw=True):
            assert int(cnts.sum()) == n and np.array_equal(offs, want_offs)
        sec_lanes = (time.perf_counter() - t0) / nb
    del texts
    for q in pinned:
        _lib.lib().fqtk_b200_host_free(q)
    out_bytes = sum(len(v) for v in res.files.values())
    name = next(iter(sorted(res.files)))
    payload, _ = ob.parse(res.files[name])  # one file read back by the strict parser (outside the timed region)
    assert payload.count(b"\n") % 4 == 0 and int(res.counts.sum()) == n
    return {"what": "FASTQ text (R1 + R2 150 bases, I1 + I2 8 bases; pinned host memory) -> scan, match, route, records with rewritten "
                    "headers, BGZF level 5 per sample (385 x 2 files) -> compressed images on the host; one call of "
                    "fqtk_b200.gpu_demux.demux_fastq_batch_gpu, wall clock, copies inside",
            "reads": n, "input_bytes": in_bytes, "text_bytes": res.text_bytes, "output_bytes": out_bytes, "files": len(res.files),
            "ms": round(sec * 1e3, 2), "mreads_per_s": round(n / sec / 1e6, 2), "gb_per_s_in": round(in_bytes / sec / 1e9, 2),
            "one_call": {"api": "fqtk_b200_demux_chunks (the same batch, one C-ABI call: compressed members of all 770 runs in one pinned "
                                "buffer + their offsets; no per-file byte strings are built)",
                         "ms": round(sec_one * 1e3, 2), "mreads_per_s": round(n / sec_one / 1e6, 2),
                         "gb_per_s_in": round(in_bytes / sec_one / 1e9, 2)},
            "two_lanes": {"api": "fqtk_b200.gpu_demux.DemuxLanes: the same call from two host threads with their own matcher / compressor "
                                 "handles, batches alternating, results taken in batch order (one batch's copies overlap the other's kernels)",
                          "batches": nb, "ms_per_batch": round(sec_lanes * 1e3, 2), "mreads_per_s": round(n / sec_lanes / 1e6, 2),
                          "gb_per_s_in": round(in_bytes / sec_lanes / 1e9, 2)}}


def measure_fastq(ctx, args, cfg, panel, matcher):
    """The ingest side: the headline config's barcodes taken straight out of in-memory, uncompressed
    index FASTQ chunks (I1 and I2, 8 bases each at cfg 3): fqtk_b200_fastq_scan (host, one thread per chunk) -> per-read
    offsets -> fqtk_b200_matcher_assign_fastq (the chunks cross PCIe whole; gather + encode + match on the GPU)."""
    import ctypes as C

    from fqtk_b200 import _lib

    lib = _lib.lib()
    L = cfg.barcode_len
    n = 4 << 20
    reads = host_reads(panel, cfg.seed_reads, 0, n)
    halves = [(0, L // 2), (L // 2, L)] if L >= 2 else [(0, L)]
    rec_head = np.frombuffer(b"@A00000:000:HXXXXXXXX:1:1101:00000:00000 1:N:0:0\n", dtype=np.uint8)
    pinned = []

    def pinned_array(count, dtype):  # the reader's chunk buffers and the scanner's tables live in pinned memory
        p = C.c_void_p()
        _lib.check(lib.fqtk_b200_host_alloc(C.byref(p), count * np.dtype(dtype).itemsize))
        pinned.append(p)
        ct = {np.uint8: C.c_uint8, np.uint32: C.c_uint32, np.uint64: C.c_uint64}[dtype]
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(ct)), shape=(count,))

    texts = []
    for lo, hi in halves:  # one FASTQ chunk per index read: fixed-width records built with numpy
        w = hi - lo
        rec = pinned_array(n * (rec_head.size + w + 3 + w + 1), np.uint8).reshape(n, -1)
        rec[:, :rec_head.size] = rec_head
        rec[:, rec_head.size:rec_head.size + w] = reads[:, lo:hi]
        rec[:, rec_head.size + w:rec_head.size + w + 3] = np.frombuffer(b"\n+\n", dtype=np.uint8)
        rec[:, rec_head.size + w + 3:rec_head.size + 2 * w + 3] = ord("F")
        rec[:, -1] = ord("\n")
        texts.append(rec.reshape(-1))
    bufs = [(pinned_array(n, np.uint64), pinned_array(n, np.uint32)) for _ in texts]
    t0 = time.perf_counter()
    tables = []

    def scan_one(job):
        text, (seq, ln) = job
        k, used = C.c_uint64(), C.c_uint64()
        _lib.check(lib.fqtk_b200_fastq_scan(text.ctypes.data, text.size, n, None, seq.ctypes.data, ln.ctypes.data,
                                            C.byref(k), C.byref(used)))
        assert k.value == n and used.value == text.size
        return seq, ln

    with ThreadPoolExecutor(len(texts)) as ex:
        tables = list(ex.map(scan_one, zip(texts, bufs)))
    scan_s = time.perf_counter() - t0
    srcs = (_lib.FastqSource * len(texts))()
    for k, (text, (seq, ln)) in enumerate(zip(texts, tables)):
        srcs[k] = _lib.FastqSource(text.ctypes.data, text.size, seq.ctypes.data, ln.ctypes.data)
    segs = (_lib.FastqSegment * len(texts))(*[_lib.FastqSegment(k, 0, hi - lo) for k, (lo, hi) in enumerate(halves)])
    out = np.empty(n, dtype=np.uint32)
    matcher.reset_counts()
    _lib.check(lib.fqtk_b200_matcher_assign_fastq(matcher._h, srcs, len(texts), segs, len(texts), n, out.ctypes.data))
    t0 = time.perf_counter()
    reps = 3
    for _ in range(reps):
        _lib.check(lib.fqtk_b200_matcher_assign_fastq(matcher._h, srcs, len(texts), segs, len(texts), n, out.ctypes.data))
    gpu_s = (time.perf_counter() - t0) / reps
    want = matcher.assign_batch(reads[: 1 << 18])
    assert np.array_equal(out[: 1 << 18], want), "FASTQ ingest results differ from dense barcode rows"
    matcher.reset_counts()
    fastq_bytes = sum(int(t.size) for t in texts)
    # the same chunks through the one-call form: scanner, per-read rules, gather, encode and match all on the device
    chunks = (_lib.FastqChunk * len(texts))(*[_lib.FastqChunk(t.ctypes.data, t.size) for t in texts])
    out2 = np.empty(n, dtype=np.uint32)
    k2, used2 = C.c_uint64(), (C.c_uint64 * len(texts))()
    _lib.check(lib.fqtk_b200_matcher_assign_fastq_chunks(matcher._h, chunks, len(texts), segs, len(texts), n, out2.ctypes.data,
                                                         C.byref(k2), used2))
    t0 = time.perf_counter()
    for _ in range(reps):
        _lib.check(lib.fqtk_b200_matcher_assign_fastq_chunks(matcher._h, chunks, len(texts), segs, len(texts), n, out2.ctypes.data,
                                                             C.byref(k2), used2))
    chunks_s = (time.perf_counter() - t0) / reps
    assert k2.value == n and np.array_equal(out2, out), "device-scanned ingest differs from the host-scanned one"
    matcher.reset_counts()
    del texts, tables, srcs
    for p in pinned:
        lib.fqtk_b200_host_free(p)
    return {"reads": n, "fastq_bytes": fastq_bytes, "inputs": 2 if L >= 2 else 1,
            "scan_mreads_per_s": round(n / scan_s / 1e6, 2), "scan_threads": 2 if L >= 2 else 1,
            "assign_fastq_mreads_per_s": round(n / gpu_s / 1e6, 2),
            "end_to_end_mreads_per_s": round(n / (scan_s + gpu_s) / 1e6, 2),
            "device_scan_end_to_end_mreads_per_s": round(n / chunks_s / 1e6, 2),
            "device_scan_gb_per_s": round(fastq_bytes / chunks_s / 1e9, 2),
            "h2d_bytes_per_read": round((fastq_bytes + (2 if L >= 2 else 1) * 12 * n) / n, 1),
            "what": "uncompressed index FASTQ text in pinned host memory -> fqtk_b200_fastq_scan -> "
                    "fqtk_b200_matcher_assign_fastq (raw chunks + offset tables over PCIe, B segments gathered and encoded on "
                    "the GPU); scan and GPU call timed back to back, not overlapped.  device_scan_*: the same chunks through "
                    "fqtk_b200_matcher_assign_fastq_chunks (raw chunks over PCIe, records found and vetted ON the GPU: no host scan)"}


def run_single_process(args):
    """The headline config through fqtk_b200_group_*: one process, one matcher per GPU, contiguous shards, ONE count table
    summed on the first device over peer-mapped pointers — the process model of the Rust host."""
    import ctypes as C

    import torch

    from fqtk_b200 import MatcherGroup, _lib, synth

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: fqtk_b200 has no CPU fallback")
    G = min(args.gpus, torch.cuda.device_count())
    cfg = synth.CONFIGS[args.config]
    n = args.reads or cfg.n_reads
    W, L = cfg.words_per_read, cfg.barcode_len
    panel = synth.panel(cfg)
    bcs = [bytes(r) for r in panel]
    peak, peak_src = measured_peak()
    group = MatcherGroup(bcs, cfg.max_mismatches, cfg.min_mismatch_delta, True, devices=list(range(G)))
    d_packed, d_res, streams = [], [], []
    for k in range(G):  # weak scaling: device k holds reads [k*n, (k+1)*n) of the stream
        with torch.cuda.device(k):
            dev = torch.device("cuda", k)
            d_packed.append(torch.empty((n, W), dtype=torch.int32, device=dev))
            d_res.append(torch.empty(n, dtype=torch.int32, device=dev))
            streams.append(torch.cuda.current_stream(dev).cuda_stream)
            synth.reads_device(panel, cfg.seed_reads, k * n, n, 0, d_packed[k].data_ptr(), streams[k])
    pk, rs, ns = [t.data_ptr() for t in d_packed], [t.data_ptr() for t in d_res], [n] * G

    def sync_all():
        for k in range(G):
            torch.cuda.synchronize(k)

    for _ in range(max(args.warmup, 1)):
        group.assign_packed_device(pk, ns, rs, streams)
    sync_all()
    group.reset_counts()
    ev0 = [torch.cuda.Event(enable_timing=True) for _ in range(G)]
    ev1 = [torch.cuda.Event(enable_timing=True) for _ in range(G)]
    for k in range(G):
        with torch.cuda.device(k):
            ev0[k].record()
    for _ in range(args.steps):
        group.assign_packed_device(pk, ns, rs, streams)
    for k in range(G):
        with torch.cuda.device(k):
            ev1[k].record()
    sync_all()
    t0 = time.perf_counter()
    counts = group.counts()  # the reduce over the devices (peer-mapped loads on device 0) + D2H
    reduce_ms = (time.perf_counter() - t0) * 1e3
    total_ms = max(ev0[k].elapsed_time(ev1[k]) for k in range(G)) + reduce_ms
    assert int(counts.sum()) == n * G * args.steps
    # parity: every device's histogram of result words, summed, == the group's one table; a window against the oracle
    hist = np.zeros(cfg.n_samples + 1, dtype=np.int64)
    for k in range(G):
        idx = torch.where(d_res[k] == -1, torch.full_like(d_res[k], cfg.n_samples), (d_res[k] >> 16) & 0xFFFF)
        hist += torch.bincount(idx.to(torch.int64), minlength=cfg.n_samples + 1).cpu().numpy()
    ok = bool(np.array_equal(hist * args.steps, counts.astype(np.int64)))
    import oracle

    src = G - 1
    off, span = n // 3 + 17, 20_000
    want, _ = oracle.OracleMatcher(bcs, cfg.max_mismatches, cfg.min_mismatch_delta).assign_batch(
        host_reads(panel, cfg.seed_reads, src * n + off, span))
    ok = ok and bool(np.array_equal(d_res[src][off:off + span].cpu().numpy().view(np.uint32), want))
    value = n * G * args.steps / (total_ms * 1e-3) / 1e6
    k_ms = max(ev0[k].elapsed_time(ev1[k]) for k in range(G)) / args.steps
    roofline = {"bound": "hbm", "achieved": round(n * cfg.algorithmic_bytes_per_read / (k_ms * 1e-3) / 1e9, 2), "peak": peak,
                "unit": "GB/s", "frac": round(n * cfg.algorithmic_bytes_per_read / (k_ms * 1e-3) / 1e9 / peak, 4),
                "traffic": None, "kernel": "group of per-device matchers", "kernel_ms": round(k_ms, 4), "peak_source": peak_src}
    # e2e: ONE pinned host batch of G * ne reads through fqtk_b200_group_assign_batch
    e2e = None
    if not args.no_e2e:
        ne = min(args.e2e_reads, n) & ~3
        lib = _lib.lib()
        h_in, h_out = C.c_void_p(), C.c_void_p()
        _lib.check(lib.fqtk_b200_host_alloc(C.byref(h_in), G * ne * L))
        _lib.check(lib.fqtk_b200_host_alloc(C.byref(h_out), G * ne * 4))
        h_in_np = np.ctypeslib.as_array(C.cast(h_in, C.POINTER(C.c_uint8)), shape=(G * ne, L))
        h_in_np[:] = host_reads(panel, cfg.seed_reads, 0, G * ne)
        for _ in range(2):
            group.assign_batch_ptr(h_in.value, G * ne, L, h_out.value)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            group.assign_batch_ptr(h_in.value, G * ne, L, h_out.value)
        dt = (time.perf_counter() - t0) / args.steps
        e2e = {"value": round(G * ne / dt / 1e6, 2), "unit": UNIT, "h2d_bytes_per_step": G * ne * L,
               "d2h_bytes_per_step": G * ne * 4, "ms_per_step": round(dt * 1e3, 3),
               "api": "fqtk_b200_group_assign_batch (one pinned host batch, contiguous shards, one host thread per device)"}
        lib.fqtk_b200_host_free(h_in)
        lib.fqtk_b200_host_free(h_out)
    emit({"metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": G, "steps": args.steps, "warmup": args.warmup,
          "ms_per_step": round(total_ms / args.steps, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
          "dtype": "u32", "data": "synthetic", "config": dict(config_dict(cfg, n, "table"), process_model="single process, "
                                                              "fqtk_b200_group_* (one matcher per GPU, one count table)"),
          "roofline": roofline, "cpu_baseline": None, "e2e": e2e, "gpu_launches": int(_lib.lib().fqtk_b200_kernel_launches()),
          "count_reduce_ms": round(reduce_ms, 3), "parity_check": {"ranks": G, "ok": ok}})
    group.close()


def run_b200(args):
    import torch
    import torch.distributed as dist

    from fqtk_b200 import _lib, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: fqtk_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = pin_to_gpu_numa_node(local) if world > 1 else {"pinned": False, "why": "single rank"}
    if world > 1:
        # stdout carries exactly one JSON line: keep NCCL's version banner (NCCL_DEBUG=VERSION on some boxes) out of it
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=dev)
    ctx = Ctx(torch, dist, world, rank, local, dev, torch.cuda.current_stream().cuda_stream)
    stream = ctx.stream
    peak = ctx.peak

    # ---- the headline: args.config (cfg 3), weak scaling: each rank synthesises its own shard straight into HBM ----
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.15)
    head = measure_config(ctx, args, args.config, args.steps, args.warmup, strong=False, cuckoo=args.cuckoo,
                          mode=args.mode, keep=True)
    clocks = None
    if sampler:
        time.sleep(0.15)
        clocks = sampler.stop()
    cfg, n, W = head["cfg"], head["n"], head["cfg"].words_per_read
    matcher, d_packed, d_res = head["matcher"], head["d_packed"], head["d_res"]
    panel, info, mode, kernel, counts = head["panel"], head["info"], head["mode"], head["kernel"], head["counts"]
    value, total_ms, roofline, launches = head["value"], head["total_ms"], head["roofline"], head["launches"]
    roofline["pair_compares_per_s"] = (round(n * cfg.n_samples / (roofline["kernel_ms"] * 1e-3), 1)
                                       if mode == "brute" else None)

    # ---- brute-force kernel family on the same batch, for the record ----
    brute = None
    if mode == "table" and not args.no_brute:
        matcher.set_mode("brute")
        nb = min(n, 64 << 20)
        matcher.assign_packed_device(d_packed.data_ptr(), nb, d_res.data_ptr(), stream)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        reps = 3
        for _ in range(reps):
            matcher.assign_packed_device(d_packed.data_ptr(), nb, d_res.data_ptr(), stream)
        e1.record()
        torch.cuda.synchronize()
        bms = e0.elapsed_time(e1) / reps
        brute = {"kernel": "k_brute_sliced", "reads": nb, "ms": round(bms, 4), "value": round(nb / (bms * 1e-3) / 1e6, 2),
                 "unit": UNIT + " per GPU", "pair_compares_per_s": round(nb * cfg.n_samples / (bms * 1e-3), 1),
                 "roofline_frac": round(nb * cfg.algorithmic_bytes_per_read / (bms * 1e-3) / 1e9 / peak, 5)}
        matcher.set_mode("table")
        matcher.reset_counts()
        # the north star's kernel (every read x every barcode) in the headline block too, next to the kernel that ran
        roofline["pair_compares_per_s_brute_kernel"] = brute["pair_compares_per_s"]

    # ---- per-sample routing of the batch, for the record ----
    routing = None
    if not args.no_routing:
        matcher.assign_packed_device(d_packed.data_ptr(), n, d_res.data_ptr(), stream)
        d_order = torch.empty(n, dtype=torch.int32, device=dev)
        d_off = torch.zeros(cfg.n_samples + 2, dtype=torch.int64, device=dev)
        matcher.route_device(d_res.data_ptr(), n, d_order.data_ptr(), d_off.data_ptr(), stream)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            matcher.route_device(d_res.data_ptr(), n, d_order.data_ptr(), d_off.data_ptr(), stream)
        e1.record()
        torch.cuda.synchronize()
        rms = e0.elapsed_time(e1) / 3
        assert int(d_off[-1].item()) == n
        routing = {"kernels": "k_route_hist_cta4 + k_route_scan_* + k_route_scatter_tile2", "reads": n, "ms": round(rms, 4),
                   "value": round(n / (rms * 1e-3) / 1e6, 2), "unit": UNIT + " per GPU",
                   "algorithmic_bytes_per_read": 12,
                   "roofline_frac": round(n * 12 / (rms * 1e-3) / 1e9 / peak, 4)}
        del d_order, d_off
        matcher.reset_counts()

    # ---- e2e: the reference-facing C-ABI call on HOST buffers (pinned), H2D + kernel + D2H every step ----
    e2e = None
    if not args.no_e2e:
        e2e = measure_e2e(ctx, args, cfg, panel, matcher, d_packed, d_res, head["first"], n)

    ingest = None
    if rank == 0 and not args.no_e2e:
        ingest = measure_fastq(ctx, args, cfg, panel, matcher)
    matcher.close()
    del d_packed, d_res
    torch.cuda.empty_cache()
    bgzf = None
    if rank == 0 and not args.no_bgzf:
        bgzf = measure_bgzf(ctx, args)
        torch.cuda.empty_cache()
        bgzf["whole_data_path"] = measure_gpu_demux(ctx, args)
        torch.cuda.empty_cache()

    # ---- the other configs of BASELINE.json: cfg 2 weak; cfg 4 and cfg 5 STRONG over the N ranks ----
    configs = None
    if not args.no_configs:
        configs = {}
        csteps, cwarm = max(1, min(args.steps, 5)), max(1, min(args.warmup, 3))
        for cid, strong in ((2, False), (4, True), (5, True)):
            if cid == args.config:
                continue
            m = measure_config(ctx, args, cid, csteps, cwarm, strong=strong)
            entry = config_entry(m, csteps, "strong" if strong else "weak")
            if strong and world > 1:
                # the same N reads on ONE GPU of this box (rank 0, the others idle), for the strong-scaling efficiency
                t1 = None
                if rank == 0:
                    solo = Ctx(torch, dist, 1, 0, local, dev, stream)
                    m1 = measure_config(solo, args, cid, csteps, cwarm, strong=True, check=False)
                    t1 = m1["total_ms"] / csteps
                ctx.barrier()
                if rank == 0:
                    entry["strong_scaling"] = {"ms_1gpu_same_box": round(t1, 4), "ms_n_gpus": entry["ms"], "n_gpus": world,
                                               "speedup": round(t1 / entry["ms"], 3),
                                               "efficiency": round(t1 / entry["ms"] / world, 4)}
            configs[f"cfg{cid}"] = entry

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline(cfg, host_panel(cfg), args.cpu_seconds)

    if rank == 0:
        line = {
            "metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(total_ms / args.steps, 4), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": config_dict(cfg, n, mode),
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches),
            "clocks": clocks,
            "memo_table": {"entries": int(info.table_entries), "slots": int(info.table_slots),
                           "bytes": int(info.table_bytes), "candidates": int(info.table_candidates),
                           "cuckoo_entries": int(info.cuckoo_entries), "cuckoo_probes": int(info.cuckoo_probes),
                           "cuckoo_slots": int(info.cuckoo_slots), "l2_table_entries": int(info.l2_table_entries),
                           "l2_table_bytes": int(info.l2_table_bytes)},
            "brute_force": brute, "routing": routing,
            "matched_fraction": round(1.0 - float(counts[-1]) / float(counts.sum()), 5),
            "parity_check": head.get("parity_check"), "configs": configs, "numa": numa, "fastq_ingest": ingest, "bgzf": bgzf,
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()


_REAL_STDOUT = None


def quiet_stdout():
    """stdout carries exactly ONE JSON line: park the real fd 1 and point fd 1 at stderr, so that nothing a library
    prints from C (NCCL's version banner, for one) can land in front of it."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)


def emit(line: dict):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    args = parse_args()
    quiet_stdout()
    if args.impl == "reference":
        run_reference(args)
    elif args.single_process:
        run_single_process(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
