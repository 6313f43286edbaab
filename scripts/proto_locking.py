"""Design prototype (NumPy, fp32 blocks / fp64 small algebra like the GPU solver): Chebyshev-filtered subspace iteration
for HOPE on a symmetric A, WITHOUT and WITH locking of converged Ritz vectors.  Question it answers: on a power-law
graph (R-MAT, beta = 0.5/rho) the fp32 dynamic-range guard reduces the filter to power steps; does deflating the
locked vectors (project them out after every filter step, bound the filter by the largest UNLOCKED Ritz value)
restore the polynomial acceleration?   python scripts/proto_locking.py --scale 14"""
import argparse, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as sla
from gem_b200 import synth

ap = argparse.ArgumentParser()
ap.add_argument('--scale', type=int, default=14)
ap.add_argument('--d', type=int, default=128)
ap.add_argument('--tol', type=float, default=1e-3)
ap.add_argument('--max-iters', type=int, default=80)
ap.add_argument('--sbm', type=int, default=0)
a = ap.parse_args()

csr = synth.sbm(n=a.sbm, block=1000) if a.sbm else synth.rmat(scale=a.scale)
A = csr.to_scipy().astype(np.float32)
n = A.shape[0]
rho = float(sla.eigsh(A.astype(np.float64), k=1, which='LA', return_eigenvectors=False)[0])
beta = 0.01 if a.sbm else 0.5 / rho
k = a.d // 2
b = k + 16
f = lambda l: beta * l / (1.0 - beta * l)
print('n', n, 'nnz', A.nnz, 'rho', rho, 'beta', beta)


def gram(P, Q):                       # fp32 products, fp64 result (tensor-core Gram + fp64 atomics)
    return (P.T @ Q).astype(np.float64)


def cholqr_pass(F):
    G = gram(F, F)
    dsc = 1.0 / np.sqrt(np.maximum(np.diag(G), 1e-300))
    Gs = G * dsc[:, None] * dsc[None, :]
    L = np.linalg.cholesky(Gs + 1e-12 * np.eye(G.shape[0]))
    Rinv = (np.linalg.inv(L).T * dsc[:, None]).astype(np.float32)      # F R^-1 orthonormal
    return F @ Rinv


def cholqr2(F):
    return cholqr_pass(cholqr_pass(F))


def orth_rotated(F, Z):
    G = gram(F, F)
    Gp = Z.T @ G @ Z
    dsc = 1.0 / np.sqrt(np.maximum(np.diag(Gp), 1e-300))
    L = np.linalg.cholesky(Gp * dsc[:, None] * dsc[None, :] + 1e-12 * np.eye(G.shape[0]))
    M = (Z @ (np.linalg.inv(L).T * dsc[:, None])).astype(np.float32)
    return cholqr_pass(F @ M)


def solve(lock):
    rng = np.random.default_rng(1234)
    V = cholqr2(rng.standard_normal((n, b)).astype(np.float32))
    sweeps = 0
    for _ in range(3):
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
        allf = np.concatenate((np.abs(f(lamQ)), g[order]))
        sig = np.sort(allf)[::-1][:k]
        change = np.max(np.abs(sig - sig_prev) / np.maximum(sig, 1e-3 * sig[0]))
        sig_prev = sig
        if it >= 2 and change <= a.tol:
            return it, sweeps, sig, Q.shape[1]
        ba = V.shape[1]
        if lock:
            # lock the leading Ritz pairs whose residual is small: r_j = ||A v_j - lam_j v_j||
            Vr = (V @ Z[:, order].astype(np.float32))
            Wr = (W @ Z[:, order].astype(np.float32))
            res = np.linalg.norm(Wr - Vr * lam[order][None, :].astype(np.float32), axis=0) / np.maximum(np.abs(lam[order]), 1e-30)
            m = 0
            while m < ba - 24 and res[m] < 1e-4 and len(lamQ) + m < k:
                m += 1
            if m:
                Q = np.concatenate((Q, Vr[:, :m]), axis=1)
                lamQ = np.concatenate((lamQ, lam[order][:m]))
                keep = order[m:]
                V = Vr[:, m:]; W = Wr[:, m:]
                lam = lam[keep]; g = g[keep]
                Z = np.eye(V.shape[1]); order = np.arange(V.shape[1])
                ba = V.shape[1]
        amax = np.max(np.abs(lam)) if not lock or not len(lamQ) else np.max(np.abs(lam))
        bound = min(hard * 1.02, 1.05 * max(amax, 1e-30))
        tau = g[order[-1]]
        hi = tau / (beta * (1.0 + tau)); lo = -tau / (beta * (1.0 - tau)) if tau < 1 else -bound
        lo = max(lo, -bound); hi = min(hi, bound)
        e = 0.5 * (hi - lo); c0 = 0.5 * (hi + lo)
        aL = bound if lam[order[0]] >= c0 else -bound
        xL = abs(aL - c0) / e
        growth = xL + np.sqrt(max(xL * xL - 1.0, 0.0))
        deg = 8
        if growth > 1 + 1e-9:
            deg = min(deg, int(np.floor(np.log(512.0) / np.log(growth))))
        if deg < 2:
            F = W
        else:
            sigma1 = e / (aL - c0); sigma = sigma1; tau2 = 2.0 / sigma1
            prev, cur = V, ((sigma / e) * (W - c0 * V)).astype(np.float32)
            for i in range(2, deg + 1):
                sn = 1.0 / (tau2 - sigma)
                nxt = ((2.0 * sn / e) * (A @ cur - c0 * cur) - (sigma * sn) * prev).astype(np.float32); sweeps += 1
                if Q.shape[1]:
                    nxt -= Q @ (Q.T @ nxt)
                sigma = sn; prev, cur = cur, nxt
            F = cur
        if Q.shape[1]:
            F = F - Q @ (Q.T @ F)
        V = orth_rotated(F, Z if Z.shape[0] == F.shape[1] else np.eye(F.shape[1]))
        if it % 5 == 0 or lock:
            print('  it %d deg %d growth %.3g change %.3g locked %d sweeps %d' % (it, deg, growth, change, Q.shape[1], sweeps), flush=True)
    return a.max_iters, sweeps, sig, Q.shape[1]


for lock in (False, True):
    t = time.time()
    it, sw, sig, nl = solve(lock)
    print('lock=%s: %d rounds, %d sweeps, locked %d, sigma_1 %.4g sigma_k %.4g  (%.1f s)' % (lock, it, sw, nl, sig[0], sig[-1], time.time() - t), flush=True)
exact = np.sort(np.abs(f(sla.eigsh(A.astype(np.float64), k=k + 8, which='BE', return_eigenvectors=False))))[::-1]
print('reference sigma_1 %.4g  (top eigenvalues by magnitude, both ends)' % exact[0])
