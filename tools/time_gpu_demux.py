"""Stage timing of fqtk_b200.gpu_demux.demux_fastq_batch_gpu (tools; not part of the bench): wraps the C-ABI calls."""
import time, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from fqtk_b200 import _lib
L = _lib.lib()
acc = {}
def wrap(name):
    f = getattr(L, name)
    def g(*a):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        r = f(*a)
        torch.cuda.synchronize(); acc[name] = acc.get(name, 0) + time.perf_counter() - t0
        return r
    g.restype = f.restype; g.argtypes = f.argtypes
    return g
class Proxy:
    def __init__(self): self.cache = {}
    def __getattr__(self, k):
        if k not in self.cache: self.cache[k] = wrap(k) if k.startswith("fqtk_b200_") and k not in ("fqtk_b200_last_error",) else getattr(L, k)
        return self.cache[k]
_lib._lib = Proxy()
import bench
class Ctx: pass
ctx = Ctx(); ctx.local = 0; ctx.torch = torch
t0 = time.perf_counter()
out = bench.measure_gpu_demux(ctx, None)
print(out["ms"], out["mreads_per_s"])
for k, v in sorted(acc.items(), key=lambda x: -x[1])[:12]: print(f"{k:50s} {v*1e3/4:9.2f} ms per call-batch")
