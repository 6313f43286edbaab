"""Design prototype #2 (NumPy, fp32 blocks): lift the fp32 dynamic-range guard of the Chebyshev filter by
orthonormalising the filtered block in GROUPS of Ritz-ordered columns (block Gram-Schmidt, twice, CholQR inside a
group) instead of one Gram of the whole block; optionally lock converged vectors and project them out of every
filter step.   python scripts/proto_bcgs.py --scale 13"""
import argparse, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import scipy.sparse.linalg as sla
from gem_b200 import synth

ap = argparse.ArgumentParser()
ap.add_argument('--scale', type=int, default=13)
ap.add_argument('--d', type=int, default=128)
ap.add_argument('--tol', type=float, default=1e-3)
ap.add_argument('--max-iters', type=int, default=60)
ap.add_argument('--deg', type=int, default=8)
ap.add_argument('--group', type=int, default=16)
ap.add_argument('--sbm', type=int, default=0)
ap.add_argument('--oversample', type=int, default=16)
ap.add_argument('--warm', type=int, default=3)
ap.add_argument('--no-exact', action='store_true')
ap.add_argument('--fixture', action='store_true', help='the reference SBM fixture (1024 nodes), tests/golden/sbm1024.npz')
a = ap.parse_args()
if a.fixture:
    import scipy.sparse as sp
    z = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests/golden/sbm1024.npz'))
    nodes = z['nodes']; pos = {int(u): i for i, u in enumerate(nodes)}
    A = sp.csr_matrix((np.ones(len(z['src']), np.float32), ([pos[int(u)] for u in z['src']], [pos[int(v)] for v in z['dst']])), shape=(len(nodes),) * 2)
    a.sbm = 1
else:
    csr = synth.sbm(n=a.sbm, block=1000) if a.sbm else synth.rmat(scale=a.scale)
    A = csr.to_scipy().astype(np.float32)
n = A.shape[0]
rho = 1.0 if a.sbm else float(sla.eigsh(A.astype(np.float64), k=1, which='LA', return_eigenvectors=False)[0])
beta = 0.01 if a.sbm else 0.5 / rho
k = a.d // 2
b = k + a.oversample
f = lambda l: beta * l / (1.0 - beta * l)
print('n', n, 'nnz', A.nnz, 'rho %.4g beta %.4g' % (rho, beta))
if a.no_exact:
    ex = np.ones(k + 4)
elif n <= 8192:
    ex = np.linalg.eigvalsh(A.astype(np.float64).toarray())
else:
    ex = sla.eigsh(A.astype(np.float64), k=k + 4, which='LA', tol=1e-9, ncv=4 * k, return_eigenvectors=False)
exact = np.sort(np.abs(f(ex)))[::-1][:k]


def gram(P, Q):
    return (P.T @ Q).astype(np.float64)


PIVOT = float(os.environ.get('PIVOT', '0'))      # 1e-5 = the GPU kernel's rule: smaller pivots drop the column


def cholqr_pass(F):
    G = gram(F, F)
    dsc = 1.0 / np.sqrt(np.maximum(np.diag(G), 1e-300))
    Gs = G * dsc[:, None] * dsc[None, :]
    if PIVOT <= 0:
        L = np.linalg.cholesky(Gs + 1e-10 * np.eye(G.shape[0]))
        return F @ (np.linalg.inv(L).T * dsc[:, None]).astype(np.float32)
    m = Gs.shape[0]
    L = np.zeros_like(Gs); keep = np.ones(m, dtype=bool); Gw = Gs.copy()
    for j in range(m):                                  # right-looking Cholesky with the kernel's pivot rule
        d = Gw[j, j]
        if d > PIVOT:
            L[j, j] = np.sqrt(d)
            L[j + 1:, j] = Gw[j + 1:, j] / L[j, j]
            Gw[j + 1:, j + 1:] -= np.outer(L[j + 1:, j], L[j + 1:, j])
        else:
            keep[j] = False
            L[j, j] = 1.0
    Li = np.linalg.inv(L)
    Li[~keep, :] = 0.0; Li[:, ~keep] = 0.0
    if (~keep).any():
        print('    cholqr: dropped %d columns' % int((~keep).sum()), flush=True)
    return F @ (Li.T * dsc[:, None]).astype(np.float32)


def cholqr2(F):
    return cholqr_pass(cholqr_pass(F))


def bcgs_groups(F, group, Q):
    """columns of F are ordered by decreasing filter gain; orthonormalise group by group against Q and the groups done"""
    done = Q
    out = []
    for s in range(0, F.shape[1], group):
        P = F[:, s:s + group].copy()
        P /= np.maximum(np.linalg.norm(P, axis=0), 1e-30).astype(np.float32)
        for _ in range(2):
            if done.shape[1]:
                P = P - done @ (done.T @ P)
            P = cholqr_pass(P)
        out.append(P)
        done = np.concatenate((done, P), axis=1)
    return np.concatenate(out, axis=1)


def solve(mode):
    rng = np.random.default_rng(1234)
    V = cholqr2(rng.standard_normal((n, b)).astype(np.float32))
    sweeps = 0
    for _ in range(a.warm):
        V = cholqr2(A @ V); sweeps += 1
    Q = np.zeros((n, 0), dtype=np.float32); lamQ = np.zeros(0)
    sig_prev = np.zeros(k)
    hard = float(abs(A).sum(axis=1).max())
    for it in range(1, a.max_iters + 1):
        W = A @ V; sweeps += 1
        if Q.shape[1]:
            W -= Q @ (Q.T @ W)
        T = gram(V, W); T = 0.5 * (T + T.T)
        lam, Z = np.linalg.eigh(T)
        g = np.abs(f(lam))
        order = np.argsort(-g)
        sig = np.sort(np.concatenate((np.abs(f(lamQ)), g)))[::-1][:k]
        change = np.max(np.abs(sig - sig_prev) / np.maximum(sig, 1e-3 * sig[0]))
        err = np.max(np.abs(sig - exact) / exact)
        sig_prev = sig
        kk = k - Q.shape[1]
        Zt_ = Z[:, order[:kk]].astype(np.float32)
        lt_ = lam[order[:kk]]
        ra_ = np.linalg.norm(W @ Zt_ - (V @ Zt_) * lt_[None, :].astype(np.float32), axis=0)
        resid = float(np.max(beta / (1.0 - beta * lt_) ** 2 * ra_) / sig[0])
        print('  %s it %d change %.3g  resid %.3g  locked %d sweeps %d' % (mode, it, change, resid, Q.shape[1], sweeps), flush=True)
        if it >= 2 and change <= a.tol:
            Zt = Z[:, order[:k - Q.shape[1]]].astype(np.float32)
            Vt, Wt, lt = V @ Zt, W @ Zt, lam[order[:k - Q.shape[1]]]
            ra = np.linalg.norm(Wt - Vt * lt[None, :].astype(np.float32), axis=0)
            fp = beta / (1.0 - beta * lt) ** 2
            print('  %s residual ||f(A)v - f(l)v|| / sigma_max ~ %.3g  (max over the active top values)' % (mode, float(np.max(fp * ra) / sig[0])), flush=True)
            return it, sweeps, err
        Zr = Z[:, order].astype(np.float32)
        Vr = V @ Zr; Wr = W @ Zr; lamr = lam[order]; gr = g[order]
        if mode == 'bcgs+lock':
            res = np.linalg.norm(Wr - Vr * lamr[None, :].astype(np.float32), axis=0) / np.maximum(np.abs(lamr), 1e-30)
            m = 0
            while m < Vr.shape[1] - 24 and res[m] < 3e-5 and len(lamQ) + m < k:
                m += 1
            if m:
                Q = np.concatenate((Q, Vr[:, :m]), axis=1); lamQ = np.concatenate((lamQ, lamr[:m]))
                Vr, Wr, lamr, gr = Vr[:, m:], Wr[:, m:], lamr[m:], gr[m:]
        bound = min(hard * 1.02, 1.05 * np.max(np.abs(lamr)))
        tau = gr[-1]
        hi = tau / (beta * (1.0 + tau)); lo = -tau / (beta * (1.0 - tau)) if tau < 1 else -bound
        lo = max(lo, -bound); hi = min(hi, bound)
        e = 0.5 * (hi - lo); c0 = 0.5 * (hi + lo)
        aL = bound if lamr[0] >= c0 else -bound
        xL = abs(aL - c0) / e
        growth = xL + np.sqrt(max(xL * xL - 1.0, 0.0))
        deg = a.deg
        if mode == 'guard' and growth > 1 + 1e-9:
            deg = min(deg, int(np.floor(np.log(512.0) / np.log(growth))))
        elif growth > 1 + 1e-9:      # keep the total gain inside fp32's exponent range with a wide margin
            deg = min(deg, max(2, int(np.floor(np.log(1e12) / np.log(growth)))))
        if deg < 2:
            F = Wr
        else:
            sigma1 = e / (aL - c0); sigma = sigma1; tau2 = 2.0 / sigma1
            prev, cur = Vr, ((sigma / e) * (Wr - c0 * Vr)).astype(np.float32)
            for i in range(2, deg + 1):
                sn = 1.0 / (tau2 - sigma)
                nxt = ((2.0 * sn / e) * (A @ cur - c0 * cur) - (sigma * sn) * prev).astype(np.float32); sweeps += 1
                if Q.shape[1]:
                    nxt -= Q @ (Q.T @ nxt)
                sigma = sn; prev, cur = cur, nxt
            F = cur
        if mode == 'guard':
            V = cholqr2(F / np.maximum(np.linalg.norm(F, axis=0), 1e-30).astype(np.float32))
        else:
            V = bcgs_groups(F, a.group, Q)
    return a.max_iters, sweeps, err


import os
for mode in os.environ.get('MODES', 'guard,bcgs,bcgs+lock').split(','):
    t = time.time()
    it, sw, err = solve(mode)
    print('%s: %d rounds, %d sweeps, max rel sigma error %.3g  (%.1f s)' % (mode, it, sw, err, time.time() - t), flush=True)
