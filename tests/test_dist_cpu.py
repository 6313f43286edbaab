"""world_size-2 gloo tests (CPU) of the multi-GPU HOST logic: row / walk sharding, padding, gather order,
Gram all-reduce and vocabulary merge -- with the oracle standing in for the CUDA kernels.  The algorithm run
here is the same block subspace iteration libgemb200 runs (gem_b200/csrc/hope.cu), written with NumPy."""
import os
import sys

import numpy as np
import pytest

from conftest import REPO, load_sbm1024_nx, load_karate_nx


def _hope_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, 'oracle')); sys.path.insert(0, os.path.join(REPO, 'tests'))
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from gem_b200 import graph as hg, dist as gd
    import hope_oracle as ho
    G, _ = load_sbm1024_nx()
    G.remove_nodes_from(list(G.nodes)[1000:])                 # n = 1000 with P = 3 would pad; with P=2 use 999
    G.remove_node(list(G.nodes)[-1])                          # n = 999: odd -> last shard is short (padding path)
    csr = hg.from_networkx(G)
    n, d, beta, b, J = csr.n, 16, 0.01, 40, 24
    k = d // 2
    r0, ip, ix, _ = csr.row_shard(rank, world)
    assert (r0, r0 + len(ip) - 1) == gd.row_range(n, rank, world)
    import scipy.sparse as sp
    A_loc = sp.csr_matrix((np.ones(len(ix)), ix, ip), shape=(len(ip) - 1, n))
    n_shard = gd.rows_per_rank(n, world)

    def allgather(Xs):                                        # what ncclAllGather does in hope.cu::dist_spmm
        t = torch.from_numpy(gd.pad_rows(Xs, n_shard))
        out = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(out, t)
        return torch.cat(out).numpy()[:n]

    def allreduce(M):
        t = torch.from_numpy(np.ascontiguousarray(M)); dist.all_reduce(t); return t.numpy()

    def katz(Xs):                                             # symmetric A: A^T shard == A shard
        W = Xs
        for _ in range(J - 1):
            W = Xs + beta * (A_loc @ allgather(W))
        return beta * (A_loc @ allgather(W))

    def orth(Xs):
        for _ in range(2):
            R = np.linalg.cholesky(allreduce(Xs.T @ Xs)).T
            Xs = Xs @ np.linalg.inv(R)
        return Xs

    rng = np.random.default_rng(5)
    V = orth(rng.standard_normal((n, b))[r0:r0 + A_loc.shape[0]])
    for it in range(60):
        U = katz(V)
        T = allreduce(U.T @ U)
        th, Z = np.linalg.eigh(T)
        if it == 59:
            break
        V = orth(katz(orth(U)))
    Zk, thk = Z[:, -k:], th[-k:]
    Xs = np.concatenate((U @ Zk * thk ** -0.25, V @ Zk * thk ** 0.25), axis=1)
    X = allgather(Xs)
    if rank == 0:
        Xo, so = ho.hope_dense_lapack(csr.to_scipy(), d, beta)
        q.put((float(np.abs(np.sqrt(thk) / so - 1).max()), float(ho.recon_rel_err(X, Xo))))
    dist.barrier()
    dist.destroy_process_group()


def _n2v_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, 'oracle')); sys.path.insert(0, os.path.join(REPO, 'tests'))
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from gem_b200 import graph as hg, dist as gd
    import n2v_oracle_py as no
    csr, nids = hg.n2v_inputs_from_networkx(load_karate_nx())
    L, R = 12, 5
    Wall = no.walks(csr.indptr, csr.indices, csr.data, nids, L, R, seed=3, mode=1)
    w0, w1 = gd.walk_range(len(nids) * R, rank, world)
    mine = Wall[w0:w1]                                         # per-walk streams: a shard IS the slice
    # vocabulary merge (n2v.cu: vocab_kernel + ncclAllReduce min / sum)
    INF = np.iinfo(np.int64).max
    first = np.full(csr.n, INF, np.int64); cnt = np.zeros(csr.n, np.int64)
    flat = mine.ravel()
    pos = np.arange(flat.size, dtype=np.int64) + w0 * L
    np.minimum.at(first, flat, pos); np.add.at(cnt, flat, 1)
    tf, tc = torch.from_numpy(first), torch.from_numpy(cnt)
    dist.all_reduce(tf, op=dist.ReduceOp.MIN); dist.all_reduce(tc, op=dist.ReduceOp.SUM)
    if rank == 0:
        gf = np.full(csr.n, INF, np.int64); gc = np.zeros(csr.n, np.int64)
        fa = Wall.ravel()
        np.minimum.at(gf, fa, np.arange(fa.size, dtype=np.int64)); np.add.at(gc, fa, 1)
        ok = bool(np.array_equal(tf.numpy(), gf) and np.array_equal(tc.numpy(), gc))
        # token order by first appearance == the oracle's (SNAP's) renumbering
        tok = np.argsort(np.where(gc > 0, gf, INF), kind='stable')[:int((gc > 0).sum())]
        _, tok_o = no.learn(Wall, csr.n, 4, 3, 1, 3)
        q.put((ok, bool(np.array_equal(tok, tok_o)), int(w1 - w0)))
    dist.barrier()
    dist.destroy_process_group()


def _spawn(fn, world=2):
    import torch.multiprocessing as mp
    import socket
    s = socket.socket(); s.bind(('127.0.0.1', 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=fn, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return res


def test_sharding_rules():
    from gem_b200 import dist as gd
    for n in (1, 7, 999, 1000, 1024):
        for P in (1, 2, 3, 8):
            ranges = [gd.row_range(n, r, P) for r in range(P)]
            assert ranges[0][0] == 0 and ranges[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
            assert all(hi - lo <= gd.rows_per_rank(n, P) for lo, hi in ranges)
            assert all(lo == min(n, r * gd.rows_per_rank(n, P)) for r, (lo, hi) in enumerate(ranges))
            w = [gd.walk_range(n * 3, r, P) for r in range(P)]
            assert w[0][0] == 0 and w[-1][1] == n * 3 and all(a[1] == b[0] for a, b in zip(w, w[1:]))


@pytest.mark.parametrize('world', [2, 3])
def test_hope_row_sharded_gloo_equals_oracle(world):
    """world = 3 does not divide the row count: the last shard is zero padded, as on 8 GPUs with an odd n."""
    sig_err, recon = _spawn(_hope_worker, world)
    assert sig_err < 1e-8 and recon < 1e-6, (sig_err, recon)


def test_node2vec_walk_shards_and_vocab_merge_gloo():
    ok, tok_ok, nloc = _spawn(_n2v_worker)
    assert ok and tok_ok and nloc > 0


def _halo_worker(rank, world, port, q):
    """The needed-rows-only exchange of gem_b200/csrc/halo.cu, host logic with gloo: every rank pushes its rows into the
    peers' halo slots (here: point-to-point sends), then multiplies with the renumbered column ids -- the result must
    equal the all-gather form."""
    import torch
    import torch.distributed as dist
    sys.path.insert(0, REPO)
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from gem_b200 import synth, dist as gd
    import scipy.sparse as sp
    csr = synth.sbm(n=3003, block=231, seed=2)                        # 3003 rows over 2 ranks: padded last shard
    n, b = csr.n, 8
    n_shard = gd.rows_per_rank(n, world)
    r0, ip, ix, _ = csr.row_shard(rank, world)
    nl = len(ip) - 1
    H, ext = gd.halo_plan(n, r0, n_shard, ix)
    sizes = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([len(H)], dtype=torch.int64))
    maxH = int(max(s.item() for s in sizes))
    pad = torch.full((maxH,), 2 ** 31 - 1, dtype=torch.int32)
    pad[:len(H)] = torch.from_numpy(H)
    allH = [torch.empty(maxH, dtype=torch.int32) for _ in range(world)]
    dist.all_gather(allH, pad)
    H_all = [allH[qq][:int(sizes[qq].item())].numpy() for qq in range(world)]
    pushes = gd.push_lists(H_all, r0, n_shard, n, rank)
    rng = np.random.default_rng(7)
    Xfull = rng.standard_normal((n, b)).astype(np.float32)
    Xloc = Xfull[r0:r0 + nl]
    block = np.zeros((n_shard + len(H), b), np.float32)               # [local rows | halo]
    block[:nl] = Xloc
    # exchange: what the producing kernel's stores do over NVLink
    other = 1 - rank
    mine = [(row, slot) for row, qq, slot in pushes if qq == other]
    send = torch.from_numpy(np.stack([Xloc[row] for row, _ in mine]) if mine else np.zeros((0, b), np.float32))
    slots_out = torch.tensor([slot for _, slot in mine], dtype=torch.int64)
    cnt = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(cnt, torch.tensor([len(mine)], dtype=torch.int64))
    n_in = int(cnt[other].item())
    recv, slots_in = torch.empty((n_in, b), dtype=torch.float32), torch.empty(n_in, dtype=torch.int64)
    if rank == 0:
        dist.send(send, 1); dist.send(slots_out, 1); dist.recv(recv, 1); dist.recv(slots_in, 1)
    else:
        dist.recv(recv, 0); dist.recv(slots_in, 0); dist.send(send, 0); dist.send(slots_out, 0)
    block[n_shard + slots_in.numpy()] = recv.numpy()
    assert n_in == len(H)                                            # with 2 ranks every halo row comes from the peer
    A_ext = sp.csr_matrix((np.ones(len(ext), np.float32), ext, ip), shape=(nl, n_shard + len(H)))
    A_glb = sp.csr_matrix((np.ones(len(ix), np.float32), ix, ip), shape=(nl, n))
    ok = np.allclose(A_ext @ block, A_glb @ Xfull, atol=1e-5)
    q.put((rank, bool(ok), len(H), len(pushes)))
    dist.destroy_process_group()


def test_halo_exchange_equals_allgather_form():
    import socket
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(('127.0.0.1', 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    ps = [ctx.Process(target=_halo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = [q.get(timeout=300) for _ in ps]
    for p in ps:
        p.join(60)
    assert all(r[1] for r in res), res
    assert all(r[2] > 0 and r[3] > 0 for r in res)
