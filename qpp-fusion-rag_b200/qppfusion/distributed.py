"""
Multi-GPU host logic (one process per GPU, torch.distributed; NCCL on the B200 box, gloo in CPU tests).

  * dense retrieval : corpus rows sharded contiguously; every rank scores all queries against its shard
                      (global ids via id_offset), one all-gather of the [Q, k] (score, id) pairs, k-way merge
                      (qf_topk_merge).  Precedent in the reference: the HNSW segment merge,
                      src/retrievers/bge.py:216-237.
  * QPP + fusion    : queries sharded contiguously, no data-path collective.  The only cross-query dependency
                      is _normalize_qpp's per-column min/max (or mean/std) over ALL queries of a run file
                      (src/qpp.py:340-374): three tiny all-reduces (MIN, MAX, SUM) of R*13 doubles.

The collective plumbing is separated from the kernel calls so it can be exercised on CPU tensors with gloo.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(n: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous shard [lo, hi) of n items for `rank` (ceil-sized shards, the last ones may be short/empty)."""
    per = (n + world - 1) // world
    lo = min(n, rank * per)
    return lo, min(n, lo + per)


def gather_topk(scores: torch.Tensor, ids: torch.Tensor, group=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """All-gather per-shard results [Q, k] -> [G, Q, k] (the layout qf_topk_merge consumes)."""
    world = dist.get_world_size(group)
    gs = torch.empty((world,) + tuple(scores.shape), dtype=scores.dtype, device=scores.device)
    gi = torch.empty((world,) + tuple(ids.shape), dtype=ids.dtype, device=ids.device)
    # views of the [G, Q, k] buffers: NCCL gathers in place, gloo (CPU tests) copies into them
    dist.all_gather(list(gs.unbind(0)), scores.contiguous(), group=group)
    dist.all_gather(list(gi.unbind(0)), ids.contiguous(), group=group)
    return gs, gi


def reduce_colstats(stats: torch.Tensor, group=None) -> torch.Tensor:
    """Combine per-rank [R, C, 4] = {min, max, sum, sumsq} tables across ranks (in place)."""
    mn, mx, sm = stats[..., 0].contiguous(), stats[..., 1].contiguous(), stats[..., 2:].contiguous()
    dist.all_reduce(mn, op=dist.ReduceOp.MIN, group=group)
    dist.all_reduce(mx, op=dist.ReduceOp.MAX, group=group)
    dist.all_reduce(sm, op=dist.ReduceOp.SUM, group=group)
    stats[..., 0], stats[..., 1], stats[..., 2:] = mn, mx, sm
    return stats


def colnorm_params(stats: torch.Tensor, q_total: int, mode: str) -> Tuple[torch.Tensor, torch.Tensor]:
    """(base, scale) per (retriever, column) exactly as _normalize_qpp derives them (src/qpp.py:355-361):
    minmax -> (min, max-min if max>min else 1.0); zscore -> (mean, std_pop if std_pop>0 else 1.0)."""
    if mode == "minmax":
        mn, mx = stats[..., 0], stats[..., 1]
        return mn.clone(), torch.where(mx > mn, mx - mn, torch.ones_like(mx))
    mean = stats[..., 2] / q_total
    std = torch.sqrt(stats[..., 3] / q_total)
    return mean, torch.where(std > 0, std, torch.ones_like(std))


def qpp_colnorm_sharded_(vals: torch.Tensor, mode: str, round6: bool, group=None) -> torch.Tensor:
    """_normalize_qpp over a query-sharded [Q_local, R, C] table (in place, on the GPU)."""
    from . import ops
    if mode == "none":
        return ops.qpp_colnorm_apply_(vals, None, None, round6) if round6 else vals
    q_total = torch.tensor([vals.shape[0]], dtype=torch.int64, device=vals.device)
    dist.all_reduce(q_total, group=group)
    q_total = int(q_total.item())
    stats = reduce_colstats(ops.qpp_colstats(vals), group)
    if mode == "zscore":
        mean = (stats[..., 2] / q_total).contiguous()
        stats = reduce_colstats(ops.qpp_colstats(vals, mean), group)
    base, scale = colnorm_params(stats, q_total, mode)
    return ops.qpp_colnorm_apply_(vals, base.contiguous(), scale.contiguous(), round6)


def exchange_topk(scores: torch.Tensor, ids: torch.Tensor, group=None):
    """All-to-all of per-shard results so that every rank holds ALL shards' lists for ITS slice of the queries:
    [Q, k] on each rank -> ([G, Qs, k], [G, Qs, k], Qs) with Qs = ceil(Q / G) (the query axis is padded with
    -inf / -1 rows).  The k-way merge then costs Q/G queries per rank instead of Q (all-gather + merge everywhere)."""
    world = dist.get_world_size(group)
    Q, k = scores.shape
    qs = (Q + world - 1) // world
    pad = qs * world - Q
    if pad:
        scores = torch.cat([scores, torch.full((pad, k), float("-inf"), dtype=scores.dtype, device=scores.device)])
        ids = torch.cat([ids, torch.full((pad, k), -1, dtype=ids.dtype, device=ids.device)])
    rs, ri = torch.empty_like(scores), torch.empty_like(ids)
    dist.all_to_all_single(rs, scores.contiguous(), group=group)     # chunk g of the send buffer goes to rank g
    dist.all_to_all_single(ri, ids.contiguous(), group=group)
    return rs.view(world, qs, k), ri.view(world, qs, k), qs


def send_depth(k: int, world: int) -> int:
    """Entries per shard list in the TRUNCATED exchange: a shard holds about k / G of the global top k (rows are
    sharded contiguously, documents are not sorted by topic), so twice that plus 64 is sent; the merge checks that
    nothing was lost (qf_topk_merge_check) and the search is repeated with full lists otherwise."""
    kk = 2 * ((k + world - 1) // world) + 64
    kk = (kk + 15) // 16 * 16
    return min(k, kk)


class ShardedSearchResult:
    """Result of dense_search_packed: everything stays on the device and NOTHING has been synchronised with the host.
    `pairs` is int32 [G, rows + 1, 2, k] after the all-gather (query q lives at [q // rows, q % rows]; [..., 0, :] are
    float32 score bits, [..., 1, :] int32 ids) or [rows + 1, 2, k] (this rank's merged query slice [lo, hi)) with
    gather=False; the spare row of every slice carries the merging rank's "truncation may have lost a document" flag.
    `flags` int32 [G] holds every rank's local-search threshold-failure flag (they travelled inside the all-to-all).
    Read `ok()` when the result is consumed (one tiny D2H, e.g. together with the result copy); False -- rare -- means
    the caller repeats the search on all ranks with exact=True (full-length lists, exact repair)."""

    def __init__(self, pairs, flags, rows, lo, hi, n_queries, gathered):
        self.pairs, self.flags, self.rows, self.lo, self.hi, self.Q, self.gathered = pairs, flags, rows, lo, hi, n_queries, gathered

    def _part(self, which):
        if self.gathered:
            v = self.pairs[:, :self.rows, which, :]
            return v.reshape(-1, v.shape[-1])[:self.Q]          # (a copy: the slices are separated by their spare rows)
        return self.pairs[:self.hi - self.lo, which, :]

    @property
    def scores(self):
        return self._part(0).view(torch.float32)

    @property
    def ids(self):
        return self._part(1)

    @property
    def incomplete(self):
        return self.pairs[:, self.rows, 0, 0] if self.gathered else self.pairs[self.rows, 0, :1]

    def ok(self) -> bool:
        return int(self.flags.max().item()) == 0 and int(self.incomplete.max().item()) == 0


def dense_search_packed(queries_bf16: torch.Tensor, corpus_shard_bf16: torch.Tensor, k: int, id_offset: int, group=None,
                        gather: bool = True, truncate: bool = True) -> ShardedSearchResult:
    """Sharded exact search with ONE exchange and no host synchronisation:
        local top-k_in written as packed (score, id) rows per destination rank (qf_dense_topk_packed); k_in =
        send_depth(k, G) < k when truncate (G >= 3): a cheaper local search and 2-4x fewer bytes on the wire
        -> ONE all-to-all (each rank receives all shards' lists for ITS Q/G queries, plus every rank's status flag in
           the spare row of its chunk)
        -> k-way merge of the slice to the top k (qf_topk_merge_packed) + completeness check of the truncation
           (qf_topk_merge_check -> flag in the slice's spare row)
        -> gather=True: ONE all-gather of the merged slices (every rank holds all Q results and all flags);
           gather=False: the rank keeps only its slice (a serving front end reads the G slices: Q*k*8 bytes in total
           leave the GPUs instead of G times that).
    Round 1 used five collectives (scores and ids separately, twice, plus a MAX all-reduce read back with .item())."""
    from . import ops
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    Q = queries_bf16.shape[0]
    k_in = send_depth(k, world) if truncate else k
    send, _status = ops.dense_topk_packed(queries_bf16, corpus_shard_bf16, k_in, id_offset, chunks=world)
    rows = send.shape[1] - 1
    recv = torch.empty_like(send)
    dist.all_to_all_single(recv, send, group=group)                    # chunk g of the send buffer goes to rank g
    flags = recv[:, rows, 0, 0].contiguous()                           # every rank's flag, identical on all ranks
    merged = torch.empty((rows + 1, 2, k), dtype=torch.int32, device=send.device)
    merged[rows, 0, :1] = 0                                            # this slice's "incomplete" flag
    ops.topk_merge_packed(recv, rows, k_out=k, out=merged, flag=merged[rows, 0, :1])
    lo = min(Q, rank * rows)
    hi = min(Q, lo + rows)
    if not gather:
        return ShardedSearchResult(merged, flags, rows, lo, hi, Q, False)
    full = torch.empty((world, rows + 1, 2, k), dtype=torch.int32, device=merged.device)
    dist.all_gather_into_tensor(full, merged, group=group)
    return ShardedSearchResult(full, flags, rows, 0, Q, Q, True)


def dense_search_sharded(queries_bf16: torch.Tensor, corpus_shard_bf16: torch.Tensor, k: int, id_offset: int, group=None,
                         merge: str = "all_to_all", exact: bool = False):
    """Every rank: local top-k over its corpus rows (global ids), exchange, k-way merge -> identical [Q, k] on all
    ranks.  -> (scores, ids, status).

    merge="all_to_all" (default): the packed ONE-exchange path (dense_search_packed) -- one all-to-all, one
    all-gather, nothing waits for the host.  `status` is then the int32 [G] vector of every rank's threshold-failure
    flag (identical on all ranks); the CALLER reads it when it consumes the result and, if any flag is set -- rare --
    repeats the call with exact=True on all ranks.
    exact=True or merge="all_gather": local search with the status-driven exact repair (reads the status back: a host
    synchronisation), scores and ids exchanged separately; `status` is the per-query vector (all zero)."""
    from . import ops
    world = dist.get_world_size(group)
    if merge == "all_to_all" and world > 1 and not exact:
        res = dense_search_packed(queries_bf16, corpus_shard_bf16, k, id_offset, group, gather=True)
        return res.scores, res.ids, torch.maximum(res.flags, res.incomplete.to(res.flags.dtype))
    scores, ids, status = ops.dense_topk(queries_bf16, corpus_shard_bf16, k, id_offset)
    Q = scores.shape[0]
    if merge == "all_gather" or world == 1:
        gs, gi = gather_topk(scores, ids, group)
        ms, mi = ops.topk_merge(gs, gi)
        return ms, mi, status
    rs, ri, qs = exchange_topk(scores, ids, group)
    ms, mi = ops.topk_merge(rs.contiguous(), ri.contiguous())                   # [qs, k] for this rank's query slice
    fs = torch.empty((world * qs, k), dtype=ms.dtype, device=ms.device)
    fi = torch.empty((world * qs, k), dtype=mi.dtype, device=mi.device)
    dist.all_gather_into_tensor(fs, ms.contiguous(), group=group)
    dist.all_gather_into_tensor(fi, mi.contiguous(), group=group)
    return fs[:Q], fi[:Q], status
