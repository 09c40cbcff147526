"""World-size-2 gloo tests (CPU) of the multi-GPU host logic: sharding, gather layout, colnorm statistics."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import dense_oracle as do
from oracle import qpp_oracle as qo
from qppfusion import distributed as qd


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(0)                      # same data on both ranks
        # ---- dense: row-sharded corpus, local top-k (oracle stands in for the kernel), gather, merge
        N, Q, D, k = 1001, 7, 32, 20
        corpus = rng.standard_normal((N, D)).astype(np.float32)
        queries = rng.standard_normal((Q, D)).astype(np.float32)
        lo, hi = qd.shard_range(N, world, rank)
        ids, sc = do.flat_ip_topk(queries, corpus[lo:hi], k)
        ids = np.where(ids >= 0, ids + lo, -1)
        gs, gi = qd.gather_topk(torch.from_numpy(sc), torch.from_numpy(ids.astype(np.int32)))
        assert gs.shape == (world, Q, k)
        ms, mi = do.merge_topk(gs.numpy(), gi.numpy(), k)
        fi, fs = do.flat_ip_topk(queries, corpus, k)
        assert np.array_equal(mi, fi) and np.array_equal(ms, fs)
        # all-to-all variant: this rank ends up with every shard's lists for its own slice of the queries
        rs, ri, qs = qd.exchange_topk(torch.from_numpy(sc), torch.from_numpy(ids.astype(np.int32)))
        assert rs.shape == (world, qs, k)
        ms2, mi2 = do.merge_topk(rs.numpy(), ri.numpy(), k)
        q0 = rank * qs
        n_mine = max(0, min(qs, Q - q0))
        assert np.array_equal(mi2[:n_mine], fi[q0:q0 + n_mine]) and np.array_equal(ms2[:n_mine], fs[q0:q0 + n_mine])
        # ---- packed ONE-exchange layout (dense_search_packed): [G, rows + 1, 2, k] int32, (score bits | ids) rows per
        # destination rank, padding rows (-inf, -1), every rank's status flag in the spare row of each chunk
        rows = (Q + world - 1) // world
        send = np.empty((world, rows + 1, 2, k), dtype=np.int32)
        send[:, :, 0, :] = np.float32(-np.inf).view(np.int32)
        send[:, :, 1, :] = -1
        for q in range(Q):
            send[q // rows, q % rows, 0] = sc[q].view(np.int32)
            send[q // rows, q % rows, 1] = ids[q].astype(np.int32)
        send[:, rows, 0, 0] = rank + 10                              # "flag" of this rank
        recv = torch.empty((world, rows + 1, 2, k), dtype=torch.int32)
        dist.all_to_all_single(recv, torch.from_numpy(send))
        recv = recv.numpy()
        assert recv[:, rows, 0, 0].tolist() == [10 + g for g in range(world)]      # everyone sees every flag
        ms3, mi3 = do.merge_topk(recv[:, :rows, 0, :].copy().view(np.float32), recv[:, :rows, 1, :].copy(), k)
        assert np.array_equal(mi3[:n_mine], fi[q0:q0 + n_mine]) and np.array_equal(ms3[:n_mine], fs[q0:q0 + n_mine])
        full = torch.empty((world * rows, 2, k), dtype=torch.int32)
        mine = np.stack([ms3.view(np.int32), mi3.astype(np.int32)], axis=1)
        dist.all_gather_into_tensor(full, torch.from_numpy(np.ascontiguousarray(mine)))
        full = full.numpy()[:Q]
        assert np.array_equal(full[:, 1, :], fi) and np.array_equal(full[:, 0, :].copy().view(np.float32), fs)
        # ---- QPP colnorm: query-sharded table, stats all-reduced, params as _normalize_qpp
        Qq, R, C = 53, 3, 13
        table = rng.normal(size=(Qq, R, C))
        table[:, 1, 4] = 0.25
        qlo, qhi = qd.shard_range(Qq, world, rank)
        loc = torch.from_numpy(table[qlo:qhi])
        for mode in ("minmax", "zscore"):
            st = torch.zeros((R, C, 4), dtype=torch.float64)
            st[..., 0] = loc.amin(dim=0); st[..., 1] = loc.amax(dim=0); st[..., 2] = loc.sum(dim=0)
            st = qd.reduce_colstats(st)
            if mode == "zscore":
                mean = st[..., 2] / Qq
                st[..., 3] = ((loc - mean) ** 2).sum(dim=0)
                s2 = st[..., 3].clone(); dist.all_reduce(s2); st[..., 3] = s2
            base, scale = qd.colnorm_params(st, Qq, mode)
            got = ((loc - base) / scale).numpy()
            for r in range(R):
                exp = qo.normalize_qpp({str(q): list(table[q, r]) for q in range(Qq)}, mode)
                e = np.array([exp[str(q)] for q in range(qlo, qhi)])
                assert np.allclose(got[:, r], e, rtol=1e-12, atol=1e-14), mode
        ret[rank] = "ok"
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert dict(ret) == {0: "ok", 1: "ok"}


def test_shard_range_covers_everything():
    for n in (0, 1, 7, 8, 1001, 2_681_468):
        for world in (1, 2, 3, 8):
            r = [qd.shard_range(n, world, i) for i in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[i][1] == r[i + 1][0] for i in range(world - 1))
