#!/bin/bash
# compute-sanitizer memcheck + racecheck on the packed-route kernels (small batches); outputs -> gpurun_out/$1
OUT=gpurun_out/${1:-sanitize}; mkdir -p $OUT
cat > /tmp/san_case.py <<'PY'
import numpy as np, torch
import oracle
from fqtk_b200 import BarcodeMatcher, _lib, synth
L = _lib.lib()
for knob in (2, 3, 1, 0):
    L.fqtk_b200_set_cuckoo_arity(knob)
    for cfg_id, n in ((3, 40_037), (2, 30_011), (5, 20_005)):
        cfg = synth.CONFIGS[cfg_id]
        if cfg_id == 5 and knob in (2, 3):
            continue
        panel = synth.panel(cfg)
        bcs = [bytes(r) for r in panel]
        reads = synth.reads_host(panel, cfg.seed_reads, 11, n)
        reads[::29, 1] = ord("N"); reads[::53, 3] = ord("r"); reads[5000:5400, 0] = ord("N")
        want, wc = oracle.OracleMatcher(bcs, cfg.max_mismatches, cfg.min_mismatch_delta).assign_batch(reads)
        d_packed = torch.from_numpy(synth.pack_host(reads).view(np.int32)).cuda()
        d_res = torch.empty(n, dtype=torch.int32, device="cuda")
        with BarcodeMatcher(bcs, cfg.max_mismatches, cfg.min_mismatch_delta, True) as m:
            m.assign_packed_device(d_packed.data_ptr(), n, d_res.data_ptr(), torch.cuda.current_stream().cuda_stream)
            torch.cuda.synchronize()
            assert np.array_equal(d_res.cpu().numpy().view(np.uint32), want), (knob, cfg_id)
            assert np.array_equal(m.counts(), wc), (knob, cfg_id)
        print("ok knob", knob, "cfg", cfg_id, flush=True)
PY
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --kernel-regex kns=k_probe --log-file $OUT/$tool.log env PYTHONPATH=$PWD python /tmp/san_case.py > $OUT/$tool.out 2>&1
  echo "$tool rc=$?"; tail -3 $OUT/$tool.out; tail -4 $OUT/$tool.log
done
