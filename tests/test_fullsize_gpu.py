"""Full-size parity (BASELINE.json configs[1]: 3,452 queries x 5 x top-1000) through the C ABI:
  * every query bit-exact against the C oracle (itself pinned on the Python oracle / reference goldens);
  * size-independent properties: ranked output is sorted, union size == number of distinct doc ids, counts sum
    to the number of postings, unranked + stable argsort == ranked, fusing the fused list again is the identity."""
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


@pytest.fixture(scope="module")
def nq_batch():
    import bench
    dev = torch.device("cuda:0")
    ids, micro, lens, feats = bench.synth_device(bench.Q_NQ, bench.R_NQ, bench.K_NQ, 20250103, dev)
    lens = lens.clone()
    lens[5, 2] = 0
    lens[6, :] = 0                       # a query with no postings at all
    lens[7, 1] = 137
    w = torch.round(torch.rand((bench.Q_NQ, bench.R_NQ), generator=torch.Generator().manual_seed(4), dtype=torch.float64) * 1e6) / 1e6
    return ids, micro, lens, feats, w.to(dev)


@pytest.mark.parametrize("method", [0, 1, 2])
def test_nq_shape_bit_exact_vs_c_oracle(nq_batch, method):
    from oracle import c_oracle as co
    from qppfusion import ops
    ids, micro, lens, feats, w = nq_batch
    out = ops.fuse(ids, micro, lens, w, None, method, 0, 60.0, 0)
    assert int(out.status.max()) == 0
    eid, esc, ecnt, epos, elen = co.fuse_weighted(ids.cpu().numpy(), micro.cpu().numpy(), lens.cpu().numpy(), w.cpu().numpy(),
                                                  method, 60.0, n_threads=os.cpu_count())
    glen = out.lens.cpu().numpy()
    assert np.array_equal(glen, elen)
    gid, gsc = out.ids.cpu().numpy(), out.scores.cpu().numpy()
    gcnt, gpos = out.counts.cpu().numpy(), out.pos.cpu().numpy()
    mask = np.arange(gid.shape[1])[None, :] < elen[:, None]
    assert np.array_equal(gid[mask], eid[mask])
    assert np.array_equal(gsc[mask].view(np.int64), esc[mask].view(np.int64))
    assert np.array_equal(gcnt[mask].astype(np.int32), ecnt[mask])
    assert np.array_equal(gpos[mask].astype(np.int32), epos[mask])


def test_nq_shape_properties(nq_batch):
    from qppfusion import ops
    ids, micro, lens, feats, w = nq_batch
    Q, R, K = ids.shape
    out = ops.fuse(ids, micro, lens, w, None, 1, 0, 60.0, 0)
    n = out.lens.long()
    col = torch.arange(R * K, device=ids.device)[None, :]
    valid = col < n[:, None]
    sc = torch.where(valid, out.scores, torch.full_like(out.scores, -float("inf")))
    assert bool((sc[:, 1:] <= sc[:, :-1]).all()), "ranked output not sorted"
    # union size == number of distinct doc ids among the valid postings
    pv = torch.arange(K, device=ids.device)[None, None, :] < lens[:, :, None]
    big = torch.where(pv, ids.long(), torch.full_like(ids.long(), -1)).reshape(Q, R * K)
    srt = torch.sort(big, dim=1).values
    distinct = ((srt[:, 1:] != srt[:, :-1]) & (srt[:, 1:] >= 0)).sum(dim=1) + (srt[:, 0] >= 0).long()
    assert torch.equal(distinct, n)
    cnt = torch.where(valid, out.counts.long(), torch.zeros_like(out.counts.long())).sum(dim=1)
    assert torch.equal(cnt, lens.long().sum(dim=1))
    # ties keep first-seen order
    tie = valid[:, 1:] & (out.scores[:, 1:] == out.scores[:, :-1])
    assert bool((out.pos.long()[:, 1:][tie] > out.pos.long()[:, :-1][tie]).all())
    # unranked output + stable sort == ranked output
    un = ops.fuse(ids, micro, lens, w, None, 1, 0, 60.0, 4)
    usc = torch.where(valid, un.scores, torch.full_like(un.scores, -float("inf")))
    order = torch.sort(usc, dim=1, descending=True, stable=True).indices
    assert torch.equal(torch.gather(un.ids, 1, order)[valid], out.ids[valid])
    # idempotence: fusing the fused list (one run, weight 1, float64 scores) returns it unchanged
    K2 = int(n.max())
    ids2 = out.ids[:, :K2].reshape(Q, 1, K2).contiguous()
    sc2 = out.scores[:, :K2].reshape(Q, 1, K2).contiguous()
    again = ops.fuse(ids2, sc2, out.lens.reshape(Q, 1).contiguous(), torch.ones((Q, 1), dtype=torch.float64, device=ids.device),
                     None, 0, 0, 60.0, 0)
    v2 = col[:, :K2] < n[:, None]
    assert torch.equal(again.ids[:, :K2][v2], out.ids[:, :K2][v2]) and torch.equal(again.scores[:, :K2][v2], out.scores[:, :K2][v2])


def test_qpp_full_size_vs_c_oracle(nq_batch):
    from oracle import c_oracle as co
    from qppfusion import ops
    ids, micro, lens, feats, w = nq_batch
    out, _ = ops.qpp_stats(micro, lens, feats, None, None, top_k=100, k_cut=50)
    exp = co.qpp_bridge_batch(micro.cpu().numpy(), lens.cpu().numpy(), feats.cpu().numpy(), 100, 50, n_threads=os.cpu_count())
    got = out.cpu().numpy()[:, :, :13]
    assert np.allclose(got, exp, rtol=1e-9, atol=1e-12)
