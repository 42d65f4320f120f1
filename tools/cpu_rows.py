#!/usr/bin/env python
"""CPU-baseline rows of BASELINE.md section 3: the oracle (C restatement of the reference matcher) on this host.
C1 = 1 thread, memo cache on (the reference's model); C2 = 1 thread, cache off; C3 = all cores, cache on (upper bound)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from bench import host_reads  # noqa: E402
from fqtk_b200 import synth  # noqa: E402


def rate(cfg, panel, reads, use_cache, threads):
    bcs = [bytes(r) for r in panel]
    if threads == 1:
        m = oracle.OracleMatcher(bcs, cfg.max_mismatches, cfg.min_mismatch_delta, use_cache=use_cache)
        t0 = time.perf_counter()
        m.assign_batch(reads, want_results=False)
        return len(reads) / (time.perf_counter() - t0) / 1e6
    t0 = time.perf_counter()
    oracle.assign_batch_mt(panel, cfg.max_mismatches, cfg.min_mismatch_delta, reads, threads=threads,
                           use_cache=use_cache, want_results=False)
    return len(reads) / (time.perf_counter() - t0) / 1e6


def main():
    secs = float(sys.argv[1]) if len(sys.argv) > 1 else 4.0
    cores = os.cpu_count()
    out = {"cores": cores, "rows": {}}
    for cid in (2, 3, 4, 5):
        cfg = synth.CONFIGS[cid]
        panel = synth.panel(cfg)
        probe = host_reads(panel, cfg.seed_reads, 0, 200_000)
        r1 = rate(cfg, panel, probe, True, 1)
        r2 = rate(cfg, panel, probe, False, 1)
        n1 = int(min(max(r1 * 1e6 * secs, 400_000), 40_000_000))
        reads = host_reads(panel, cfg.seed_reads, 0, n1)
        c1 = rate(cfg, panel, reads, True, 1)
        n2 = int(min(max(r2 * 1e6 * secs, 100_000), n1))
        c2 = rate(cfg, panel, reads[:n2], False, 1)
        n3 = int(min(n1 * min(cores, 8), 80_000_000))
        reads3 = reads if n3 <= n1 else host_reads(panel, cfg.seed_reads, 0, n3)
        c3 = rate(cfg, panel, reads3[:n3], True, cores)
        out["rows"][f"cfg{cid}"] = {"C1_1thread_cache_on": round(c1, 3), "n_C1": n1, "C2_1thread_cache_off": round(c2, 4),
                                    "n_C2": n2, "C3_all_cores_cache_on": round(c3, 2), "n_C3": n3}
        print(json.dumps({f"cfg{cid}": out["rows"][f"cfg{cid}"]}), flush=True)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
