"""Developer A/B (not the bench): ms per SpMM sweep of the symmetric HOPE solve under GEMB_SPMM* environment knobs.
Run one process per setting (the knobs are read once per process)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gem_b200 import _native, synth
csr = synth.sbm(n=int(os.environ.get('EXP_N', '1000000')), block=1000, seed=42)
ctx = _native.Context(0)
g = _native.DeviceGraph(ctx, csr.n, csr.indptr, csr.indices, None)
osamp = int(os.environ.get('EXP_OS', '16'))
kw = dict(tol=1e-3, max_iters=30, min_iters=2, oversample=osamp, seed=1234)
g.hope(128, 0.01, want_output=False, **kw)
best = None
for _ in range(3):
    _, _, st = g.hope(128, 0.01, want_output=False, **kw)
    if best is None or st['spmm_ms'] < best['spmm_ms']:
        best = st
env = {k: v for k, v in os.environ.items() if k.startswith('GEMB_SPMM') or k == 'EXP_OS'}
print(json.dumps(dict(env=env, block=best['block'], sweeps=best['spmm_count'], ms_per_sweep=best['spmm_ms'] / best['spmm_count'],
                      GBps=best['spmm_bytes'] / (best['spmm_ms'] / best['spmm_count'] * 1e-3) / 1e9, total_ms=best['total_ms'],
                      dense_ms=best['dense_ms'])), flush=True)
