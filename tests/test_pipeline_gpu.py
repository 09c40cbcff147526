"""GPU: the CUDA-graph step (qppfusion.pipeline.FusionStepGraph) replays exactly what the kernel-by-kernel ops compute,
and follows in-place refills of its input tensors."""
import numpy as np
import pytest
import torch

from helpers import pack_lists, synth_lists

pytestmark = pytest.mark.gpu


def _inputs(seed, n_q, dev):
    lists = synth_lists(seed, n_q=n_q, n_r=5, k=200, pool=450, ragged=True, neg_ranker=2, ties=True)
    ids, sc, ranks, lens = pack_lists(lists, k=200, score_type="micro")
    rng = np.random.default_rng(seed)
    n_tok = rng.integers(1, 15, n_q)
    feats = np.stack([n_tok, np.maximum(n_tok - 1, 1), rng.integers(3, 12, n_q), n_tok * 5], axis=1).astype(np.int32)
    t = lambda a: torch.from_numpy(a).to(dev)
    return t(ids), t(sc), t(lens), t(feats)


def _eager(ids, sc, lens, feats, method):
    from qppfusion import ops
    qpp, _ = ops.qpp_stats(sc, lens, feats, None, None, top_k=100, k_cut=50)
    ops.qpp_colnorm_(qpp, "minmax", round6=True)
    w = ops.qpp_select_weight(qpp, 5)
    return qpp, w, ops.fuse(ids, sc, lens, w, None, method, 0, 60.0, 0)


@pytest.mark.parametrize("method", [0, 1, 2])
def test_graph_step_matches_eager_and_tracks_input_refills(method):
    from qppfusion.pipeline import FusionStepGraph
    dev = torch.device("cuda:0")
    a = _inputs(31, 64, dev)
    b = _inputs(32, 64, dev)
    g = FusionStepGraph(*a, qpp_index=5, method=method, want_counts=True, want_pos=True, compact=True)
    for src in (a, b, a):
        for dst, s_ in zip(g.inputs[:4], src):
            dst.copy_(s_)                               # refill the graph's inputs in place
        out = g.replay()
        torch.cuda.synchronize()
        qpp, w, exp = _eager(*src, method)
        assert torch.equal(g.qpp, qpp) and torch.equal(g.weights, w)
        n = exp.lens
        assert torch.equal(out.lens, n) and torch.equal(out.status, exp.status)
        m = torch.arange(out.ids.shape[1], device=dev)[None, :] < n[:, None]
        for x, y in ((out.ids, exp.ids), (out.scores, exp.scores), (out.counts.to(torch.int32), exp.counts.to(torch.int32)),
                     (out.pos.to(torch.int32), exp.pos.to(torch.int32))):
            assert torch.equal(x[m], y[m])
        off, cids, csc = g.csr
        assert int(off[-1]) == int(n.sum()) and torch.equal(cids[: int(off[-1])], exp.ids[m])


def test_graph_check_raises_on_flagged_lists():
    """a graph replay must not pass QSTAT_BAD_LEN / QSTAT_UNSORTED silently -- FusionStepGraph.check() raises
    like fusion.fuse_packed does (one small read when the result is consumed)."""
    from qppfusion.pipeline import FusionStepGraph
    dev = torch.device("cuda:0")
    ids, sc, lens, feats = _inputs(33, 16, dev)
    g = FusionStepGraph(ids, sc, lens, feats, qpp_index=5, method=0)
    g.replay()
    g.check()                                           # clean batch: nothing raised
    keep = sc[3, 1, :2].clone()
    sc[3, 1, 0], sc[3, 1, 1] = keep[1] - 5, keep[0] + 5  # ascending pair at the top of one list
    g.replay()
    with pytest.raises(ValueError, match="not sorted"):
        g.check()
    sc[3, 1, :2] = keep
    lens[5, 2] = 10_000                                 # longer than K
    g.replay()
    with pytest.raises(ValueError, match="invalid list length"):
        g.check()
