"""
Device-side step of the reference's 03_qpp -> 05_fusion pipeline, captured as a CUDA graph.

One step = qf_qpp_stats (13 QPPBridge predictors per run list, scripts/03_qpp.py -> src/qpp.py:245-337)
         -> qf_qpp_colnorm (cross-query min-max / z-score + "%.6f", src/qpp.py:340-374, :331-334)
         -> qf_qpp_select_weight (get_qpp_weight, src/fusion.py:116-158)
         -> qf_fuse (weighted_combsum / weighted_combmnz / weighted_rrf, src/fusion.py:237-380).
The kernels are short (tens to hundreds of microseconds), so issuing them one by one from Python leaves the GPU
waiting for the host; the graph replays the whole step with one launch.  Inputs live in caller-owned device tensors
whose storage the graph reads on every replay (refill them in place, e.g. with copy_ from pinned host memory).
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _native as N
from . import ops


class FusionStepGraph:
    def __init__(self, doc_ids: torch.Tensor, scores: torch.Tensor, lens: torch.Tensor, text_feats: torch.Tensor, *,
                 qpp_index: int = 5, top_k: int = 100, k_cut: int = 50, colnorm: Optional[str] = "minmax",
                 round6: bool = True, method: int = N.METHOD_SUM, norm: int = N.NORM_NONE, rrf_k: float = 60.0,
                 flags: int = 0, want_counts: bool = False, want_pos: bool = False, compact: bool = False,
                 ranks: Optional[torch.Tensor] = None, warmup: int = 2, wire: Optional["ops.Wire"] = None):
        """wire: a DEVICE ops.Wire whose words the step expands into doc_ids / scores first (qf_wire_unpack): the host
        then refills `wire.words` (5.25 instead of 8 bytes per posting over PCIe for the NQ shape) instead of the
        two int32 tensors."""
        self.dev = N.require_cuda(doc_ids, scores, lens, text_feats, ranks)
        self.inputs = (doc_ids, scores, lens, text_feats, ranks)
        self.wire = wire
        self._cfg = dict(qpp_index=int(qpp_index), top_k=int(top_k), k_cut=int(k_cut), colnorm=colnorm,
                         round6=bool(round6), method=int(method), norm=int(norm), rrf_k=float(rrf_k), flags=int(flags),
                         want_counts=bool(want_counts), want_pos=bool(want_pos), compact=bool(compact))
        cur = torch.cuda.current_stream(self.dev)
        side = torch.cuda.Stream(self.dev)
        side.wait_stream(cur)
        with torch.cuda.stream(side):          # warm-up outside capture: one-time kernel attribute setup, allocator
            for _ in range(max(1, warmup)):
                warm = self._body()
        cur.wait_stream(side)
        torch.cuda.synchronize(self.dev)
        # table-size hint from the warm-up batch (a performance choice only; see N.FLAG_LARGE_UNION)
        r, k = doc_ids.shape[1], doc_ids.shape[2]
        ratio = float(warm[2].lens.clamp(min=0).float().mean().item()) / float(max(1, r * k))
        if ratio > 0.46 and not (self._cfg["flags"] & N.FLAG_LARGE_UNION):
            self._cfg["flags"] |= N.FLAG_LARGE_UNION
            with torch.cuda.stream(side):
                self._body()                   # configure that kernel variant outside capture too
            cur.wait_stream(side)
            torch.cuda.synchronize(self.dev)
        del warm
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.qpp, self.weights, self.out, self.csr, self.qpp_status = self._body()

    def _body(self):
        ids, sc, lens, feats, ranks = self.inputs
        c = self._cfg
        if self.wire is not None:
            ops.wire_unpack_(self.wire, ids, sc)
        qpp, qpp_status = ops.qpp_stats(sc, lens, feats, None, None, top_k=c["top_k"], k_cut=c["k_cut"])
        if c["colnorm"] or c["round6"]:
            ops.qpp_colnorm_(qpp, c["colnorm"] or "none", round6=c["round6"])
        w = ops.qpp_select_weight(qpp, c["qpp_index"])
        out = ops.fuse(ids, sc, lens, w, ranks, c["method"], c["norm"], c["rrf_k"], c["flags"],
                       want_counts=c["want_counts"], want_pos=c["want_pos"])
        csr = ops.compact_fused(out) if c["compact"] else None
        return qpp, w, out, csr, qpp_status

    def replay(self):
        """Run the captured step on the current stream; returns the (static) FuseOut."""
        self.graph.replay()
        return self.out

    def check(self) -> None:
        """Raise like fusion.fuse_packed does when the last replay flagged a list (QSTAT_BAD_LEN: the query produced
        no output; QSTAT_UNSORTED: the ordering contract of ranx semantics / the rank column was violated) or the QPP
        kernel saw an unsorted list.  One small device->host read: call it when the result is consumed, not per replay."""
        st = torch.stack([(self.out.status & N.QSTAT_BAD_LEN).max(), (self.out.status & N.QSTAT_UNSORTED).max(),
                          self.qpp_status.max()]).cpu().tolist()
        if st[0]:
            raise ValueError("qf_fuse: invalid list length")
        if st[1]:
            bad = torch.nonzero(self.out.status & N.QSTAT_UNSORTED).flatten()[:5].tolist()
            raise ValueError(f"qf_fuse: run lists violate the ordering contract for query rows {bad}")
        if st[2]:
            raise ValueError("qf_qpp_stats: a run list is not sorted by score descending")


class ShardedFusionStep:
    """The same step on a QUERY-SHARDED run file (one rank per GPU): _normalize_qpp's per-column min / max run over ALL
    queries of the file (src/qpp.py:340-374), so the ranks exchange them -- the one collective of the QPP + fusion path:

        graph A  qf_qpp_stats -> qf_qpp_colstats (this rank's rows) -> [-min | max]
        NCCL     ONE all-reduce(MAX) of 2 x R x 13 doubles, enqueued on the same stream between the two replays
        graph B  (v - min) / (max - min or 1) + "%.6f" -> qf_qpp_select_weight -> qf_fuse

    min / max are exact under any reduction order, so the sharded result equals the single-GPU one bit for bit
    (checked by bench.py at N > 1 and by tools/dist_gpu_check.py).  Only colnorm="minmax" is offered here: the z-score
    variant's sums depend on the reduction order (use distributed.qpp_colnorm_sharded_ and a tolerance)."""

    def __init__(self, doc_ids, scores, lens, text_feats, *, group=None, qpp_index: int = 5, top_k: int = 100,
                 k_cut: int = 50, round6: bool = True, method: int = N.METHOD_SUM, rrf_k: float = 60.0, flags: int = 0,
                 ranks: Optional[torch.Tensor] = None, want_counts: bool = False, want_pos: bool = False,
                 compact: bool = False, wire: Optional["ops.Wire"] = None):
        import torch.distributed as dist
        self.dev = N.require_cuda(doc_ids, scores, lens, text_feats, ranks)
        self.group, self._dist = group, dist
        ids, sc, ln, feats = doc_ids, scores, lens, text_feats
        cur = torch.cuda.current_stream(self.dev)
        side = torch.cuda.Stream(self.dev)

        def part_a():
            if wire is not None:
                ops.wire_unpack_(wire, ids, sc)
            qpp, _st = ops.qpp_stats(sc, ln, feats, None, None, top_k=top_k, k_cut=k_cut)
            st = ops.qpp_colstats(qpp)                                   # [R, C, 4] = min, max, sum, sumsq
            return qpp, torch.stack([-st[..., 0], st[..., 1]]).contiguous()

        def part_b(qpp, mm):
            mn, mx = -mm[0], mm[1]
            scale = torch.where(mx > mn, mx - mn, torch.ones_like(mx))
            ops.qpp_colnorm_apply_(qpp, mn.contiguous(), scale.contiguous(), round6)
            w = ops.qpp_select_weight(qpp, qpp_index)
            out = ops.fuse(ids, sc, ln, w, ranks, method, N.NORM_NONE, rrf_k, flags, want_counts=want_counts,
                           want_pos=want_pos)
            return w, out, (ops.compact_fused(out) if compact else None)

        side.wait_stream(cur)
        with torch.cuda.stream(side):                   # warm-up outside capture (kernel attributes, allocator, NCCL)
            for _ in range(2):
                qpp, mm = part_a()
                dist.all_reduce(mm, op=dist.ReduceOp.MAX, group=group)
                part_b(qpp, mm)
        cur.wait_stream(side)
        torch.cuda.synchronize(self.dev)
        self.graph_a, self.graph_b = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph_a):
            self.qpp, self.mm = part_a()
        with torch.cuda.graph(self.graph_b):
            self.weights, self.out, self.csr = part_b(self.qpp, self.mm)

    def replay(self):
        self.graph_a.replay()
        self._dist.all_reduce(self.mm, op=self._dist.ReduceOp.MAX, group=self.group)
        self.graph_b.replay()
        return self.out
