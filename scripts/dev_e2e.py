"""Developer probe: where the end-to-end time of HOPE.learn_embedding goes (ctx, upload, solve, D2H, free)."""
import argparse, json, sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from gem_b200 import _native, synth
from gem_b200.graph import HostCSR

ap = argparse.ArgumentParser()
ap.add_argument('--n', type=int, default=1_000_000)
ap.add_argument('--d', type=int, default=128)
ap.add_argument('--reps', type=int, default=3)
a = ap.parse_args()
csr = synth.sbm(n=a.n)
ip = _native.pinned_empty(csr.n + 1, np.int32); ip[:] = csr.indptr
ix = _native.pinned_empty(csr.nnz, np.int32); ix[:] = csr.indices
out = _native.pinned_empty((csr.n, a.d), np.float32)
for r in range(a.reps):
    T = {}
    t0 = time.perf_counter(); ctx = _native.Context(0); T['ctx'] = time.perf_counter() - t0
    t0 = time.perf_counter(); g = _native.DeviceGraph(ctx, csr.n, ip, ix, None); T['upload'] = time.perf_counter() - t0
    t0 = time.perf_counter(); X, sig, st = g.hope(a.d, 0.01, tol=1e-3, max_iters=30, oversample=16, seed=1234, out=out); T['hope_call'] = time.perf_counter() - t0
    T['hope_device_ms'] = st['total_ms']; T['spmm_ms'] = st['spmm_ms'] / 1e3; T['dense_ms'] = st['dense_ms'] / 1e3
    t0 = time.perf_counter(); g.free(); T['gfree'] = time.perf_counter() - t0
    t0 = time.perf_counter(); ctx.close(); T['ctxclose'] = time.perf_counter() - t0
    print(json.dumps({k: round(v * 1e3, 2) if k != 'hope_device_ms' else round(v, 2) for k, v in T.items()}), flush=True)
