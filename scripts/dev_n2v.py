"""Developer probe: node2vec on a synthetic SBM, prints the phase timings."""
import argparse, json, sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from gem_b200 import _native, synth

ap = argparse.ArgumentParser()
ap.add_argument('--n', type=int, default=1_000_000)
ap.add_argument('--d', type=int, default=128)
ap.add_argument('--walk-len', type=int, default=80)
ap.add_argument('--num-walks', type=int, default=10)
ap.add_argument('--con-size', type=int, default=10)
ap.add_argument('--reps', type=int, default=1)
a = ap.parse_args()
t = time.time()
csr = synth.sbm(n=a.n)
nids = np.arange(csr.n, dtype=np.int32)
print('graph', csr.n, csr.nnz, 'gen s', round(time.time() - t, 2), flush=True)
ctx = _native.Context(0)
g = _native.DeviceGraph(ctx, csr.n, csr.indptr, csr.indices, None)
for r in range(a.reps):
    t = time.time()
    X, st = g.node2vec(nids, a.d, a.walk_len, a.num_walks, a.con_size, 1, seed=1, want_output=(r == a.reps - 1))
    st['wall_s'] = time.time() - t
    st['nodes_per_s_device'] = csr.n / (st['total_ms'] * 1e-3)
    st['sgns_TBps'] = st['sgns_bytes'] / (st['sgns_ms'] * 1e-3) / 1e12 if st['sgns_ms'] > 0 else 0
    print(json.dumps(st), flush=True)
print('X', X.shape, float(np.abs(X).mean()), bool(np.isfinite(X).all()))
