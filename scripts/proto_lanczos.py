"""NumPy model of gem_b200/csrc/hope.cu::hope_lanczos (thick-restart block Lanczos on A, fp32 blocks / fp64 small
algebra), to check the host logic and the convergence on power-law and clustered spectra before the GPU run.
    python scripts/proto_lanczos.py --rmat 14 | --sbm 100000"""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import scipy.sparse.linalg as sla
from gem_b200 import synth

ap = argparse.ArgumentParser()
ap.add_argument('--rmat', type=int, default=0)
ap.add_argument('--sbm', type=int, default=0)
ap.add_argument('--d', type=int, default=128)
ap.add_argument('--tol', type=float, default=1e-3)
ap.add_argument('--mmax', type=int, default=0)
ap.add_argument('--p', type=int, default=16)
a = ap.parse_args()
csr = synth.rmat(scale=a.rmat, seed=3) if a.rmat else synth.sbm(n=a.sbm or 100000, seed=42)
A = csr.to_scipy().astype(np.float32)
n = csr.n
if a.rmat:
    rho = float(sla.eigsh(A.astype(np.float64), k=1, which='LA', return_eigenvectors=False)[0])
    beta = 0.5 / rho
else:
    beta = 0.01
k, p, cw = a.d // 2, a.p, 64
k_keep = (k + p + p - 1) // p * p
m_max = a.mmax or max(2 * k_keep, 160)
m_max = (m_max + p - 1) // p * p
f = lambda l: beta * l / (1 - beta * l)
fp = lambda l: beta / (1 - beta * l) ** 2
rng = np.random.default_rng(1)

def cholqr2(W):
    Rt = np.eye(W.shape[1])
    for _ in range(2):
        G = (W.T.astype(np.float64) @ W.astype(np.float64))
        R = np.linalg.cholesky(G).T
        W = (W.astype(np.float64) @ np.linalg.inv(R)).astype(np.float32)
        Rt = R @ Rt
    return W, Rt

Q = np.zeros((n, m_max), np.float32)
T = np.zeros((m_max + p, m_max + p))
V, _ = cholqr2(rng.standard_normal((n, p)).astype(np.float32))
m = steps = restarts = 0
t0 = time.time()
while True:
    Q[:, m:m + p] = V
    j0 = m; m += p; steps += 1
    W = (A @ V).astype(np.float32)
    H = np.zeros((m, p))
    for _ in range(2):
        Hc = Q[:, :m].T.astype(np.float64) @ W.astype(np.float64)
        W = (W - (Q[:, :m] @ Hc.astype(np.float32))).astype(np.float32)
        H += Hc
    T[:m, j0:j0 + p] = H; T[j0:j0 + p, :m] = H.T
    T[j0:j0 + p, j0:j0 + p] = 0.5 * (T[j0:j0 + p, j0:j0 + p] + T[j0:j0 + p, j0:j0 + p].T)
    V, R = cholqr2(W)
    if m + p <= m_max:
        continue
    th, Y = np.linalg.eigh(T[:m, :m])
    fa = np.abs(f(th))
    order = np.argsort(-fa)
    res = np.linalg.norm(R @ Y[m - p:m, :][:, order[:k]], axis=0)
    worst = float(np.max(fp(th[order[:k]]) * res) / fa[order[0]])
    restarts += 1
    print('restart %d steps %d matvecs %d  sigma_k/sigma_1 %.3g  residual %.3g' % (restarts, steps, steps * p, fa[order[k - 1]] / fa[order[0]], worst), flush=True)
    if worst <= a.tol or restarts >= 40:
        sel = order[:k]
        break
    keep = order[:k_keep]
    Q[:, :k_keep] = (Q[:, :m] @ Y[:, keep].astype(np.float32))
    Q[:, k_keep:] = 0
    T[:] = 0
    T[np.arange(k_keep), np.arange(k_keep)] = th[keep]
    m = k_keep
print('time %.1fs' % (time.time() - t0))
Vk = (Q[:, :m] @ Y[:, sel].astype(np.float32)).astype(np.float64)
lam = th[sel]
Ad = csr.to_scipy().astype(np.float64)
r = np.linalg.norm(Ad @ Vk - Vk * lam, axis=0)
print('true residual |A v - l v| * f\' / sigma_max: %.3g ; orth %.3g' % (np.max(fp(lam) * r) / np.abs(f(lam)).max(), np.abs(Vk.T @ Vk - np.eye(k)).max()))
ref = sla.eigsh(Ad, k=min(k + 20, n - 2), which='BE', return_eigenvectors=False, tol=1e-10)
sr = np.sort(np.abs(f(ref)))[::-1][:k]
print('sigma rel err vs eigsh: %.3g' % np.max(np.abs(np.sort(np.abs(f(lam)))[::-1] / sr - 1)))
