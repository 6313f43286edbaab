"""Developer probe (not the bench): HOPE on a synthetic SBM, prints the solver stats."""
import argparse, json, sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from gem_b200 import _native, synth

ap = argparse.ArgumentParser()
ap.add_argument('--n', type=int, default=1_000_000)
ap.add_argument('--d', type=int, default=128)
ap.add_argument('--beta', type=float, default=0.01)
ap.add_argument('--tol', type=float, default=1e-4)
ap.add_argument('--max-iters', type=int, default=30)
ap.add_argument('--reps', type=int, default=3)
ap.add_argument('--rmat', type=int, default=0)
ap.add_argument('--algorithm', type=int, default=0)
ap.add_argument('--cheb-degree', type=int, default=0)
a = ap.parse_args()
t = time.time()
csr = synth.rmat(scale=a.rmat) if a.rmat else synth.sbm(n=a.n)
print('graph', csr.n, csr.nnz, 'gen s', round(time.time() - t, 2), flush=True)
ctx = _native.Context(0)
t = time.time()
g = _native.DeviceGraph(ctx, csr.n, csr.indptr, csr.indices, None)
print('upload s', round(time.time() - t, 3), flush=True)
for r in range(a.reps):
    t = time.time()
    X, sig, st = g.hope(a.d, a.beta, tol=a.tol, max_iters=a.max_iters, want_output=(r == a.reps - 1), verbose=(r == 0), algorithm=a.algorithm, cheb_degree=a.cheb_degree,
                        compute_residual=int(r == a.reps - 1))
    wall = time.time() - t
    st['wall_s'] = wall
    st['nodes_per_s_device'] = csr.n / (st['total_ms'] * 1e-3)
    st['spmm_GBps'] = st['spmm_bytes'] * st['spmm_count'] / (st['spmm_ms'] * 1e-3) / 1e9 if st['spmm_ms'] > 0 else 0
    print(json.dumps(st), flush=True)
print('sigma', sig[:3], sig[-3:])
