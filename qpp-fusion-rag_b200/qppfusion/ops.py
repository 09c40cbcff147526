"""
PyTorch custom ops over the C ABI (torch.ops.qppfusion.*).

Each op allocates its outputs with torch (device memory + current stream are PyTorch's; the arithmetic is
libqppfusion's hand-written sm_100a kernels) and calls the matching qf_* entry point through ctypes.
"""
from __future__ import annotations

import ctypes
from typing import NamedTuple, Optional, Tuple

import torch

from . import _native as N

_SCORE_DTYPES = {N.SCORE_F64: torch.float64, N.SCORE_F32: torch.float32, N.SCORE_MICRO: torch.int32}


def _score_type_of(t: torch.Tensor) -> int:
    for k, v in _SCORE_DTYPES.items():
        if t.dtype == v:
            return k
    raise TypeError(f"scores must be float64, float32 or int32 (micro-units), got {t.dtype}")


# ---------------------------------------------------------------------------------------------- fuse
@torch.library.custom_op("qppfusion::fuse", mutates_args=())
def _fuse_op(doc_ids: torch.Tensor, scores: torch.Tensor, lens: torch.Tensor,
             weights: Optional[torch.Tensor], ranks: Optional[torch.Tensor],
             method: int, norm: int, rrf_k: float, flags: int, want_counts: bool, want_pos: bool
             ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    dev = N.require_cuda(doc_ids, scores, lens, weights, ranks)
    if doc_ids.dim() != 3 or doc_ids.dtype != torch.int32:
        raise TypeError("doc_ids must be int32 [Q, R, K]")
    Q, R, K = doc_ids.shape
    if scores.shape != doc_ids.shape:
        raise ValueError("scores must have the shape of doc_ids")
    if lens.shape != (Q, R) or lens.dtype != torch.int32:
        raise TypeError("lens must be int32 [Q, R]")
    if weights is not None and (weights.shape != (Q, R) or weights.dtype != torch.float64):
        raise TypeError("weights must be float64 [Q, R]")
    if ranks is not None and (ranks.shape != doc_ids.shape or ranks.dtype != torch.int32):
        raise TypeError("ranks must be int32 [Q, R, K]")
    st = _score_type_of(scores)
    n = R * K
    out_ids = torch.empty((Q, n), dtype=torch.int32, device=dev)
    out_scores = torch.empty((Q, n), dtype=torch.float64, device=dev)
    out_counts = torch.empty((Q, n) if want_counts else (0, n), dtype=torch.uint16, device=dev)
    out_pos = torch.empty((Q, n) if want_pos else (0, n), dtype=torch.uint16, device=dev)
    out_lens = torch.empty((Q,), dtype=torch.int32, device=dev)
    out_status = torch.empty((Q,), dtype=torch.int32, device=dev)
    deep = n > fuse_envelope(st) or bool(flags & N.FLAG_DEBUG_DEEP)
    flags &= ~N.FLAG_DEBUG_DEEP
    args = (N.ptr(doc_ids), N.ptr(scores), st, N.ptr(ranks), N.ptr(lens), N.ptr(weights),
            Q, R, K, method, norm, float(rrf_k), flags, N.ptr(out_ids), N.ptr(out_scores),
            N.ptr(out_counts) if want_counts else None, N.ptr(out_pos) if want_pos else None,
            N.ptr(out_lens), N.ptr(out_status))
    with torch.cuda.device(dev):
        if deep:
            # deeper than the shared-memory kernels: per-query state in a device workspace (qf_fuse_deep)
            nbytes = N.lib().qf_fuse_deep_workspace_bytes(Q, R, K)
            if nbytes == 0 and Q > 0:
                raise ValueError(f"qf_fuse_deep: rankers x depth = {R} x {K} = {n} postings per query exceeds "
                                 f"{N.lib().qf_fuse_deep_max_postings()} (and 32 rankers)")
            ws = torch.empty((max(nbytes, 256),), dtype=torch.uint8, device=dev)
            N.check(N.lib().qf_fuse_deep(*args, N.ptr(ws), nbytes, N.stream_ptr(dev)))
        else:
            N.check(N.lib().qf_fuse(*args, N.stream_ptr(dev)))
    return out_ids, out_scores, out_counts, out_pos, out_lens, out_status


def fuse_envelope(score_type: int) -> int:
    """Postings per query (R*K) the shared-memory kernels of qf_fuse take; above it ops.fuse calls qf_fuse_deep."""
    cap = int(N.lib().qf_fuse_max_postings())
    return min(cap, 5120) if score_type == N.SCORE_F64 else cap


@_fuse_op.register_fake
def _(doc_ids, scores, lens, weights, ranks, method, norm, rrf_k, flags, want_counts, want_pos):
    Q, R, K = doc_ids.shape
    n = R * K
    e = lambda shape, dt: torch.empty(shape, dtype=dt, device=doc_ids.device)
    return (e((Q, n), torch.int32), e((Q, n), torch.float64), e((Q, n) if want_counts else (0, n), torch.uint16),
            e((Q, n) if want_pos else (0, n), torch.uint16),
            e((Q,), torch.int32), e((Q,), torch.int32))


class FuseOut(NamedTuple):
    ids: torch.Tensor       # int32  [Q, R*K]  ranked doc ids
    scores: torch.Tensor    # float64[Q, R*K]  fused scores
    counts: torch.Tensor    # uint16 [Q, R*K]  postings merged per doc (CombMNZ count); [0, R*K] if not requested
    pos: torch.Tensor       # uint16 [Q, R*K]  first-seen index of the doc;              [0, R*K] if not requested
    lens: torch.Tensor      # int32  [Q]       union size
    status: torch.Tensor    # int32  [Q]       QSTAT_* bits


def fuse(doc_ids, scores, lens, weights=None, ranks=None, method=N.METHOD_SUM, norm=N.NORM_NONE,
         rrf_k=60.0, flags=0, want_counts=True, want_pos=True) -> FuseOut:
    return FuseOut(*torch.ops.qppfusion.fuse(doc_ids, scores, lens, weights, ranks, int(method), int(norm),
                                             float(rrf_k), int(flags), bool(want_counts), bool(want_pos)))


def compact_fused(out: "FuseOut", cap: Optional[int] = None):
    """FuseOut (padded [Q, R*K]) -> CSR on the device: (offsets int64 [Q+1], ids int32 [cap], scores f64 [cap]).
    cap defaults to the padded size; entries beyond offsets[Q] are unspecified."""
    dev = N.require_cuda(out.ids, out.scores, out.lens)
    Q, stride = out.ids.shape
    cap = int(cap) if cap is not None else Q * stride
    offsets = torch.empty((Q + 1,), dtype=torch.int64, device=dev)
    cids = torch.empty((cap,), dtype=torch.int32, device=dev)
    csc = torch.empty((cap,), dtype=torch.float64, device=dev)
    nbytes = N.lib().qf_compact_workspace_bytes(Q)
    ws = torch.empty((max(nbytes, 8),), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        N.check(N.lib().qf_compact_fused(N.ptr(out.ids), N.ptr(out.scores), N.ptr(out.lens), Q, stride, N.ptr(offsets),
                                         N.ptr(cids), N.ptr(csc), N.ptr(ws), nbytes, N.stream_ptr(dev)))
    return offsets, cids, csc


# ---------------------------------------------------------------------------------------------- wire format
class Wire(NamedTuple):
    """Bit-packed postings (csrc/qf_wire.cu).  `words` is ONE int64 buffer (host, optionally pinned, or device):
    coding "fixed": the stream of (id_bits + sc_bits)-bit postings;
    coding "for"  : [ doc-id stream (id_bits each) | block bases int32 | block offsets uint32 | payload uint32 ] -- the
                    scores frame-of-reference coded per block of 32; the payload length depends on the data, `used`
                    (int64 elements) is what has to be uploaded, the buffer is sized for the worst case."""
    words: torch.Tensor
    shape: Tuple[int, int, int]      # (Q, R, K) of the arrays it expands to
    id_bits: int
    sc_bits: int                     # 0 for coding "for"
    sc_base: int
    coding: str = "fixed"
    used: int = -1                   # int64 elements in use (-1: all)

    def to(self, device, non_blocking: bool = False) -> "Wire":
        return self._replace(words=self.words.to(device, non_blocking=non_blocking))

    @property
    def used_words(self) -> int:
        return self.words.numel() if self.used < 0 else self.used


def _for_segments(Q: int, R: int, K: int, id_bits: int):
    """int64-element offsets of the four segments of a coding="for" buffer: (ids, base, off, payload, end)."""
    n = Q * R * K
    nb = int(N.lib().qf_wire_for_blocks(Q, R, K))
    ids_w = int(N.lib().qf_wire_words(n, id_bits, 0))
    base_w, off_w, pay_w = (nb * 4 + 7) // 8, ((nb + 1) * 4 + 7) // 8, ((nb * 32 + 1) * 4 + 7) // 8
    a = ids_w
    return 0, a, a + base_w, a + base_w + off_w, a + base_w + off_w + pay_w


def wire_pack(doc_ids: torch.Tensor, micro: torch.Tensor, lens: torch.Tensor, pin: bool = False,
              n_threads: int = 0, widths: Optional[Tuple[int, int, int]] = None, coding: str = "fixed") -> Wire:
    """HOST side: int32 [Q, R, K] doc ids (dictionary codes >= 0) + int32 micro-unit scores + lens -> Wire.
    coding="fixed": widths = (id_bits, sc_bits, sc_base) fixes the layout (e.g. to refill one pinned buffer step after
    step); default: the smallest that fits (qf_wire_plan).  coding="for": scores frame-of-reference coded per block of
    32 (sorted lists: ~8 bits per score fewer); only widths[0] = id_bits is used."""
    import os
    for t_ in (doc_ids, micro, lens):
        if t_.is_cuda or t_.dtype != torch.int32 or not t_.is_contiguous():
            raise TypeError("wire_pack takes contiguous int32 CPU tensors")
    Q, R, K = doc_ids.shape
    if micro.shape != doc_ids.shape or lens.shape != (Q, R):
        raise ValueError("micro must have the shape of doc_ids, lens [Q, R]")
    if coding not in ("fixed", "for"):
        raise ValueError("coding must be 'fixed' or 'for'")
    nt = n_threads or min(32, os.cpu_count() or 1)
    if widths is None:
        ib, sb, base = ctypes.c_int(0), ctypes.c_int(0), ctypes.c_int32(0)
        N.check(N.lib().qf_wire_plan(N.ptr(doc_ids), N.ptr(micro), N.ptr(lens), Q, R, K, nt, ctypes.addressof(ib),
                                     ctypes.addressof(sb), ctypes.addressof(base)))
        widths = (ib.value, sb.value, base.value)
    ib_, sb_, base_ = (int(x) for x in widths)
    if coding == "for":
        o_ids, o_base, o_off, o_pay, end = _for_segments(Q, R, K, ib_)
        if o_base == 0:
            raise ValueError(f"wire_pack: bad id_bits={ib_}")
        words = torch.empty((end,), dtype=torch.int64, pin_memory=pin)
        words[o_off - 1] = 0                      # the segments are padded to 8 bytes: keep the stream deterministic
        words[o_pay - 1] = 0
        N.check(N.lib().qf_wire_pack(N.ptr(doc_ids), None, N.ptr(lens), Q, R, K, ib_, 0, 0, N.ptr(words), o_base, nt))
        used = ctypes.c_size_t(0)
        p0 = words.data_ptr()
        N.check(N.lib().qf_wire_for_pack(N.ptr(micro), N.ptr(lens), Q, R, K, nt, p0 + 8 * o_base, p0 + 8 * o_off,
                                         p0 + 8 * o_pay, (end - o_pay) * 2, ctypes.addressof(used)))
        if used.value & 1:
            words.view(torch.int32)[2 * o_pay + int(used.value)] = 0
        return Wire(words, (Q, R, K), ib_, 0, 0, "for", o_pay + (int(used.value) + 1) // 2)
    nw = int(N.lib().qf_wire_words(Q * R * K, ib_, sb_))
    if nw == 0 or sb_ == 0:
        raise ValueError(f"wire_pack: bad widths id_bits={ib_} sc_bits={sb_}")
    words = torch.empty((nw,), dtype=torch.int64, pin_memory=pin)
    N.check(N.lib().qf_wire_pack(N.ptr(doc_ids), N.ptr(micro), N.ptr(lens), Q, R, K, ib_, sb_, base_, N.ptr(words), nw, nt))
    return Wire(words, (Q, R, K), ib_, sb_, base_)


def wire_unpack_(wire: Wire, doc_ids: torch.Tensor, micro: torch.Tensor) -> None:
    """DEVICE side: expand `wire` (words on the device) into the int32 [Q, R, K] tensors the kernels read (in place:
    these are the static inputs of a FusionStepGraph)."""
    dev = N.require_cuda(wire.words, doc_ids, micro)
    if tuple(doc_ids.shape) != tuple(wire.shape) or micro.shape != doc_ids.shape or \
            doc_ids.dtype != torch.int32 or micro.dtype != torch.int32:
        raise TypeError("wire_unpack_: outputs must be int32 tensors of the wire's [Q, R, K] shape")
    Q, R, K = wire.shape
    with torch.cuda.device(dev):
        if wire.coding == "for":
            _o_ids, o_base, o_off, o_pay, _end = _for_segments(Q, R, K, wire.id_bits)
            p0 = wire.words.data_ptr()
            N.check(N.lib().qf_wire_unpack(p0, doc_ids.numel(), wire.id_bits, 0, 0, N.ptr(doc_ids), None,
                                           N.stream_ptr(dev)))
            N.check(N.lib().qf_wire_for_unpack(p0 + 8 * o_base, p0 + 8 * o_off, p0 + 8 * o_pay, Q, R, K, N.ptr(micro),
                                               N.stream_ptr(dev)))
        else:
            N.check(N.lib().qf_wire_unpack(N.ptr(wire.words), doc_ids.numel(), wire.id_bits, wire.sc_bits, wire.sc_base,
                                           N.ptr(doc_ids), N.ptr(micro), N.stream_ptr(dev)))


# ---------------------------------------------------------------------------------------------- qpp
@torch.library.custom_op("qppfusion::qpp_stats", mutates_args=())
def _qpp_stats_op(scores: torch.Tensor, lens: torch.Tensor, text_feats: torch.Tensor,
                  query_idf: Optional[torch.Tensor], perms: Optional[torch.Tensor],
                  top_k: int, k_cut: int, wig_k: int) -> Tuple[torch.Tensor, torch.Tensor]:
    dev = N.require_cuda(scores, lens, text_feats, query_idf, perms)
    if scores.dim() != 3:
        raise TypeError("scores must be [Q, R, K]")
    Q, R, K = scores.shape
    if lens.shape != (Q, R) or lens.dtype != torch.int32:
        raise TypeError("lens must be int32 [Q, R]")
    if text_feats.shape != (Q, 4) or text_feats.dtype != torch.int32:
        raise TypeError("text_feats must be int32 [Q, 4]")
    if query_idf is not None and (query_idf.shape != (Q,) or query_idf.dtype != torch.float64):
        raise TypeError("query_idf must be float64 [Q]")
    n_samples = perm_len = 0
    if perms is not None:
        if perms.dim() != 2 or perms.dtype != torch.int32:
            raise TypeError("perms must be int32 [n_samples, perm_len]")
        n_samples, perm_len = perms.shape
    st = _score_type_of(scores)
    out = torch.empty((Q, R, N.QPP_COLS), dtype=torch.float64, device=dev)
    status = torch.empty((Q, R), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        N.check(N.lib().qf_qpp_stats(N.ptr(scores), st, N.ptr(lens), N.ptr(text_feats), N.ptr(query_idf),
                                     N.ptr(perms), n_samples, perm_len, Q, R, K, top_k, k_cut, wig_k,
                                     N.ptr(out), N.ptr(status), N.stream_ptr(dev)))
    return out, status


@_qpp_stats_op.register_fake
def _(scores, lens, text_feats, query_idf, perms, top_k, k_cut, wig_k):
    Q, R, K = scores.shape
    return (torch.empty((Q, R, N.QPP_COLS), dtype=torch.float64, device=scores.device),
            torch.empty((Q, R), dtype=torch.int32, device=scores.device))


def qpp_stats(scores, lens, text_feats, query_idf=None, perms=None, top_k=100, k_cut=50, wig_k=5):
    """-> (out float64 [Q, R, 18], status int32 [Q, R]).  Columns 0..12 = config.qpp.methods order."""
    return torch.ops.qppfusion.qpp_stats(scores, lens, text_feats, query_idf, perms, int(top_k), int(k_cut), int(wig_k))


@torch.library.custom_op("qppfusion::qpp_colnorm", mutates_args=("vals",))
def _colnorm_op(vals: torch.Tensor, mode: int, round6: bool) -> None:
    dev = N.require_cuda(vals)
    if vals.dim() != 3 or vals.dtype != torch.float64:
        raise TypeError("vals must be float64 [Q, R, ncols]")
    Q, R, C = vals.shape
    nbytes = N.lib().qf_qpp_colnorm_workspace_bytes(R, C)
    ws = torch.empty((max(nbytes, 8),), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        N.check(N.lib().qf_qpp_colnorm(N.ptr(vals), Q, R, C, mode, 1 if round6 else 0, N.ptr(ws), nbytes,
                                       N.stream_ptr(dev)))


def qpp_colnorm_(vals: torch.Tensor, mode: str = "minmax", round6: bool = False) -> torch.Tensor:
    """In-place _normalize_qpp over the query axis of vals [Q, R, ncols] (+ optional "%.6f" round trip)."""
    m = {"none": 0, "minmax": N.COLNORM_MINMAX, "zscore": N.COLNORM_ZSCORE}[mode]
    torch.ops.qppfusion.qpp_colnorm(vals, m, bool(round6))
    return vals


@torch.library.custom_op("qppfusion::qpp_select_weight", mutates_args=())
def _select_weight_op(vals: torch.Tensor, col: int, ncols_mean: int) -> torch.Tensor:
    dev = N.require_cuda(vals)
    if vals.dim() != 3 or vals.dtype != torch.float64:
        raise TypeError("vals must be float64 [Q, R, ncols]")
    Q, R, C = vals.shape
    w = torch.empty((Q, R), dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        N.check(N.lib().qf_qpp_select_weight(N.ptr(vals), Q, R, C, col, ncols_mean, N.ptr(w), N.stream_ptr(dev)))
    return w


@_select_weight_op.register_fake
def _(vals, col, ncols_mean):
    return torch.empty(vals.shape[:2], dtype=torch.float64, device=vals.device)


def qpp_select_weight(vals: torch.Tensor, col: int, ncols_mean: int = N.QPP_BRIDGE_COLS) -> torch.Tensor:
    """weights[q, r] = vals[q, r, col], or the mean of the first ncols_mean columns when col == -1."""
    return torch.ops.qppfusion.qpp_select_weight(vals, int(col), int(ncols_mean))


# ---------------------------------------------------------------------------------------------- dense
@torch.library.custom_op("qppfusion::dense_topk", mutates_args=())
def _dense_topk_op(queries: torch.Tensor, corpus: torch.Tensor, k: int, id_offset: int
                   ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    dev = N.require_cuda(queries, corpus)
    if queries.dtype != torch.bfloat16 or corpus.dtype != torch.bfloat16:
        raise TypeError("queries and corpus must be bfloat16 (convert the fp32 index once with corpus_to_device)")
    if queries.dim() != 2 or corpus.dim() != 2 or queries.shape[1] != corpus.shape[1]:
        raise ValueError("queries [Q, D] and corpus [N, D] must share D")
    Q, D = queries.shape
    Nn = corpus.shape[0]
    scores = torch.empty((Q, k), dtype=torch.float32, device=dev)
    ids = torch.empty((Q, k), dtype=torch.int32, device=dev)
    status = torch.zeros((Q,), dtype=torch.int32, device=dev)
    nbytes = N.lib().qf_dense_workspace_bytes(Q, Nn, k)
    ws = torch.empty((max(nbytes, 16),), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        N.check(N.lib().qf_dense_topk(N.ptr(queries), N.ptr(corpus), Q, Nn, D, k, id_offset, N.ptr(scores), N.ptr(ids),
                                      N.ptr(status), N.ptr(ws), nbytes, N.stream_ptr(dev)))
    return scores, ids, status


@_dense_topk_op.register_fake
def _(queries, corpus, k, id_offset):
    Q = queries.shape[0]
    return (torch.empty((Q, k), dtype=torch.float32, device=queries.device),
            torch.empty((Q, k), dtype=torch.int32, device=queries.device),
            torch.empty((Q,), dtype=torch.int32, device=queries.device))


def dense_rescore_topk(queries_f32: torch.Tensor, corpus_f32: torch.Tensor, cand_ids: torch.Tensor, k: int,
                       id_offset: int = 0):
    """float32 re-scoring of over-fetched candidates + exact top-k of the re-scored lists.
    queries_f32 [Q, D], corpus_f32 [N, D] (this shard's rows), cand_ids int32 [Q, kc] (global ids, -1 padding)
    -> (scores f32 [Q, k] desc, ids int32 [Q, k]); ties by candidate position (= the bf16 ranking)."""
    dev = N.require_cuda(queries_f32, corpus_f32, cand_ids)
    if queries_f32.dtype != torch.float32 or corpus_f32.dtype != torch.float32 or cand_ids.dtype != torch.int32:
        raise TypeError("queries/corpus must be float32 and cand_ids int32")
    Q, D = queries_f32.shape
    kc = cand_ids.shape[1]
    k = min(int(k), kc)
    queries_f32, corpus_f32, cand_ids = queries_f32.contiguous(), corpus_f32.contiguous(), cand_ids.contiguous()
    rs = torch.empty((Q, kc), dtype=torch.float32, device=dev)
    out_s = torch.empty((Q, k), dtype=torch.float32, device=dev)
    out_c = torch.empty((Q, k), dtype=torch.int32, device=dev)
    status = torch.zeros((Q,), dtype=torch.int32, device=dev)
    rows = torch.arange(Q, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        N.check(N.lib().qf_dense_rescore_f32(N.ptr(queries_f32), N.ptr(corpus_f32), Q, corpus_f32.shape[0], D,
                                             N.ptr(cand_ids), kc, id_offset, N.ptr(rs), N.stream_ptr(dev)))
        N.check(N.lib().qf_dense_row_topk(N.ptr(rs), kc, N.ptr(rows), Q, kc, k, 0, N.ptr(out_s), N.ptr(out_c),
                                          N.ptr(status), N.stream_ptr(dev)))
    ids = torch.gather(cand_ids, 1, out_c.clamp(min=0).long())
    ids = torch.where((out_c >= 0) & torch.isfinite(out_s), ids, torch.full_like(ids, -1))
    return out_s, ids


def dense_topk_packed(queries: torch.Tensor, corpus: torch.Tensor, k: int, id_offset: int = 0, chunks: int = 1):
    """Local search written straight into the layout the ONE-exchange multi-GPU search sends: int32
    [chunks, rows + 1, 2, k] with rows = ceil(Q / chunks); query q lives in chunk q // rows, row q % rows as
    (k float32 score bits, k int32 ids); padding rows hold (-inf, -1); row `rows` of every chunk is the spare row whose
    first word carries this rank's status flag (1 = some query's sampled threshold failed).  -> (pairs, status [Q])."""
    dev = N.require_cuda(queries, corpus)
    if queries.dtype != torch.bfloat16 or corpus.dtype != torch.bfloat16:
        raise TypeError("queries and corpus must be bfloat16 (convert the fp32 index once with corpus_to_device)")
    Q, D = queries.shape
    Nn = corpus.shape[0]
    rows = (Q + chunks - 1) // chunks
    pairs = torch.empty((chunks, rows + 1, 2, k), dtype=torch.int32, device=dev)
    pad = chunks * rows - Q
    flat = pairs.view(chunks * (rows + 1), 2, k)
    if pad:                                           # trailing padding rows of the last chunk(s): (-inf, -1)
        for c in range(chunks):
            lo = max(0, min(rows, Q - c * rows))
            if lo < rows:
                pairs[c, lo:rows, 0, :] = -8388608    # 0xFF800000 = float32 -inf
                pairs[c, lo:rows, 1, :] = -1
    status = torch.zeros((Q,), dtype=torch.int32, device=dev)
    nbytes = N.lib().qf_dense_workspace_bytes(Q, Nn, k)
    ws = torch.empty((max(nbytes, 16),), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        N.check(N.lib().qf_dense_topk_packed(N.ptr(queries), N.ptr(corpus), Q, Nn, D, k, id_offset, N.ptr(pairs), rows,
                                             (rows + 1) * 2 * k, N.ptr(status), N.ptr(ws), nbytes, N.stream_ptr(dev)))
    pairs[:, rows, 0, 0] = (status != 0).any().to(torch.int32)        # the rank's flag rides in every chunk's spare row
    del flat
    return pairs, status


def topk_merge_packed(pairs: torch.Tensor, rows: int, k_out: Optional[int] = None, out: Optional[torch.Tensor] = None,
                      flag: Optional[torch.Tensor] = None) -> torch.Tensor:
    """pairs int32 [G, rows(+spare), 2, k_in] (as the all-to-all leaves them) -> merged int32 [rows, 2, k_out]
    (written into `out[:rows]` when given).  With `flag` (device int32 [1]) and k_in < k_out the completeness check of
    the truncated exchange runs too (qf_topk_merge_check)."""
    dev = N.require_cuda(pairs)
    G, rr, two, k_in = pairs.shape
    k_out = int(k_out or k_in)
    if out is None:
        out = torch.empty((rows, 2, k_out), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        N.check(N.lib().qf_topk_merge_packed(N.ptr(pairs), G, rows, k_in, rr * 2 * k_in, k_out, N.ptr(out), N.stream_ptr(dev)))
        if flag is not None and k_in < k_out:
            N.check(N.lib().qf_topk_merge_check(N.ptr(pairs), G, rows, k_in, rr * 2 * k_in, N.ptr(out), k_out, N.ptr(flag),
                                                N.stream_ptr(dev)))
    return out


def dense_scores(queries: torch.Tensor, corpus: torch.Tensor) -> torch.Tensor:
    """Full fp32 score rows [ceil128(Q), 64*ceil(N/64)] (tcgen05 kernel, STORE epilogue); for small Q."""
    dev = N.require_cuda(queries, corpus)
    Q, D = queries.shape
    Nn = corpus.shape[0]
    ld = (Nn + 63) // 64 * 64
    out = torch.empty(((Q + 127) // 128 * 128, ld), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        N.check(N.lib().qf_dense_scores(N.ptr(queries), N.ptr(corpus), Q, Nn, D, N.ptr(out), ld, N.stream_ptr(dev)))
    return out


def dense_repair_(queries, corpus, k, id_offset, scores, ids, status, max_rows: int = 512) -> int:
    """Exact fallback: recompute every query whose status != 0 by streaming its full score row
    (qf_dense_scores + qf_dense_row_topk), in place.  Returns the number of repaired queries."""
    bad = torch.nonzero(status).flatten()
    n_bad = int(bad.numel())
    dev = queries.device
    for s0 in range(0, n_bad, max_rows):
        rows = bad[s0:s0 + max_rows].to(torch.int32).contiguous()
        qb = queries[rows.long()].contiguous()
        full = dense_scores(qb, corpus)
        with torch.cuda.device(dev):
            N.check(N.lib().qf_dense_row_topk(N.ptr(full), full.shape[1], N.ptr(rows), int(rows.numel()), corpus.shape[0],
                                              k, id_offset, N.ptr(scores), N.ptr(ids), N.ptr(status), N.stream_ptr(dev)))
        del full
    return n_bad


def dense_topk(queries: torch.Tensor, corpus: torch.Tensor, k: int, id_offset: int = 0, exact_fallback: bool = True):
    """-> (scores f32 [Q,k] desc, ids int32 [Q,k] (-1 padded), status int32 [Q]).
    With exact_fallback (default) queries whose sampled threshold failed are recomputed exactly, so status is 0
    on return; this reads the status vector back (one small D2H sync)."""
    scores, ids, status = torch.ops.qppfusion.dense_topk(queries, corpus, int(k), int(id_offset))
    if exact_fallback and bool(status.any().item()):
        dense_repair_(queries, corpus, int(k), int(id_offset), scores, ids, status)
    return scores, ids, status


@torch.library.custom_op("qppfusion::topk_merge", mutates_args=())
def _topk_merge_op(scores: torch.Tensor, ids: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    dev = N.require_cuda(scores, ids)
    if scores.dim() != 3 or scores.dtype != torch.float32 or ids.shape != scores.shape or ids.dtype != torch.int32:
        raise TypeError("scores f32 [G, Q, k] and ids int32 [G, Q, k] expected")
    G, Q, k = scores.shape
    out_s = torch.empty((Q, k), dtype=torch.float32, device=dev)
    out_i = torch.empty((Q, k), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        N.check(N.lib().qf_topk_merge(N.ptr(scores), N.ptr(ids), G, Q, k, N.ptr(out_s), N.ptr(out_i), N.stream_ptr(dev)))
    return out_s, out_i


@_topk_merge_op.register_fake
def _(scores, ids):
    G, Q, k = scores.shape
    return (torch.empty((Q, k), dtype=torch.float32, device=scores.device),
            torch.empty((Q, k), dtype=torch.int32, device=scores.device))


def topk_merge(scores: torch.Tensor, ids: torch.Tensor):
    """k-way merge of per-shard results [G, Q, k] (as all_gather leaves them) -> ([Q, k], [Q, k])."""
    return torch.ops.qppfusion.topk_merge(scores, ids)


# ---------------------------------------------------------------------------------------------- colstats (multi-GPU)
def qpp_colstats(vals: torch.Tensor, mean: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Per (retriever, column) {min, max, sum, sumsq-about-mean} of this rank's rows: float64 [R, ncols, 4]."""
    dev = N.require_cuda(vals, mean)
    if vals.dim() != 3 or vals.dtype != torch.float64:
        raise TypeError("vals must be float64 [Q, R, ncols]")
    Q, R, C = vals.shape
    stats = torch.empty((R, C, 4), dtype=torch.float64, device=dev)
    nbytes = N.lib().qf_qpp_colnorm_workspace_bytes(R, C)
    ws = torch.empty((max(nbytes, 8),), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        N.check(N.lib().qf_qpp_colstats(N.ptr(vals), Q, R, C, N.ptr(mean), N.ptr(stats), N.ptr(ws), nbytes,
                                        N.stream_ptr(dev)))
    return stats


def qpp_colnorm_apply_(vals: torch.Tensor, base: Optional[torch.Tensor], scale: Optional[torch.Tensor],
                       round6: bool = False) -> torch.Tensor:
    """In place (v - base[r,c]) / scale[r,c] (+ optional "%.6f" round trip)."""
    dev = N.require_cuda(vals, base, scale)
    Q, R, C = vals.shape
    with torch.cuda.device(dev):
        N.check(N.lib().qf_qpp_colnorm_apply(N.ptr(vals), Q, R, C, N.ptr(base), N.ptr(scale), 1 if round6 else 0,
                                             N.stream_ptr(dev)))
    return vals


# ---------------------------------------------------------------------------------------------- query encoder
@torch.library.custom_op("qppfusion::bert_encode", mutates_args=())
def _bert_encode_op(weights: torch.Tensor, tokens: torch.Tensor, lens: torch.Tensor, n_layers: int, hidden: int,
                    heads: int, ffn: int, vocab: int, max_pos: int, pooling: int, ln_eps: float
                    ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    dev = N.require_cuda(weights, tokens, lens)
    if weights.dtype != torch.bfloat16 or weights.dim() != 1:
        raise TypeError("weights must be the packed bf16 blob (qppfusion.encoder.pack_bert_weights)")
    if tokens.dtype != torch.int32 or tokens.dim() != 2 or lens.dtype != torch.int32 or lens.shape != (tokens.shape[0],):
        raise TypeError("tokens int32 [B, L] and lens int32 [B] expected")
    cfg = N.BertConfigC(n_layers, hidden, heads, ffn, vocab, max_pos, pooling, ln_eps)
    cp = ctypes.addressof(cfg)
    need = N.lib().qf_bert_weight_elems(cp)
    if need == 0:
        raise N.NativeError(-2, N.lib().qf_last_error().decode("utf-8", "replace"))
    if weights.numel() != need:
        raise ValueError(f"weight blob has {weights.numel()} elements, the config needs {need}")
    B, L = tokens.shape
    out_b = torch.empty((B, hidden), dtype=torch.bfloat16, device=dev)
    out_f = torch.empty((B, hidden), dtype=torch.float32, device=dev)
    status = torch.empty((B,), dtype=torch.int32, device=dev)
    ws = torch.empty((max(N.lib().qf_bert_workspace_bytes(cp, B, L), 8),), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        N.check(N.lib().qf_bert_encode(cp, N.ptr(weights), N.ptr(tokens), N.ptr(lens), B, L, N.ptr(out_b), N.ptr(out_f),
                                       N.ptr(status), N.ptr(ws), ws.numel(), N.stream_ptr(dev)))
    return out_b, out_f, status


@_bert_encode_op.register_fake
def _(weights, tokens, lens, n_layers, hidden, heads, ffn, vocab, max_pos, pooling, ln_eps):
    B = tokens.shape[0]
    return (torch.empty((B, hidden), dtype=torch.bfloat16, device=tokens.device),
            torch.empty((B, hidden), dtype=torch.float32, device=tokens.device),
            torch.empty((B,), dtype=torch.int32, device=tokens.device))


def bert_encode(weights: torch.Tensor, tokens: torch.Tensor, lens: torch.Tensor, *, n_layers: int, hidden: int, heads: int,
                ffn: int, vocab: int, max_pos: int, pooling: int = N.BERT_POOL_CLS, ln_eps: float = 1e-12):
    """-> (embeddings bf16 [B, H] -- the query operand of dense_topk, embeddings f32 [B, H], status int32 [B])."""
    return torch.ops.qppfusion.bert_encode(weights, tokens, lens, int(n_layers), int(hidden), int(heads), int(ffn),
                                           int(vocab), int(max_pos), int(pooling), float(ln_eps))


@torch.library.custom_op("qppfusion::bert_encode_packed", mutates_args=())
def _bert_encode_packed_op(weights: torch.Tensor, tokens: torch.Tensor, cu_seqlens: torch.Tensor, max_len: int,
                           n_layers: int, hidden: int, heads: int, ffn: int, vocab: int, max_pos: int, pooling: int,
                           ln_eps: float) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    dev = N.require_cuda(weights, tokens, cu_seqlens)
    if weights.dtype != torch.bfloat16 or weights.dim() != 1:
        raise TypeError("weights must be the packed bf16 blob (qppfusion.encoder.pack_bert_weights)")
    if tokens.dtype != torch.int32 or tokens.dim() != 1 or cu_seqlens.dtype != torch.int32 or cu_seqlens.dim() != 1 \
            or cu_seqlens.numel() < 1:
        raise TypeError("tokens int32 [M] and cu_seqlens int32 [B + 1] expected")
    cfg = N.BertConfigC(n_layers, hidden, heads, ffn, vocab, max_pos, pooling, ln_eps)
    cp = ctypes.addressof(cfg)
    need = N.lib().qf_bert_weight_elems(cp)
    if need == 0:
        raise N.NativeError(-2, N.lib().qf_last_error().decode("utf-8", "replace"))
    if weights.numel() != need:
        raise ValueError(f"weight blob has {weights.numel()} elements, the config needs {need}")
    B, M = cu_seqlens.numel() - 1, tokens.numel()
    out_b = torch.empty((B, hidden), dtype=torch.bfloat16, device=dev)
    out_f = torch.empty((B, hidden), dtype=torch.float32, device=dev)
    status = torch.empty((B,), dtype=torch.int32, device=dev)
    ws = torch.empty((max(N.lib().qf_bert_workspace_bytes(cp, 1, max(M, 1)), 8),), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        N.check(N.lib().qf_bert_encode_packed(cp, N.ptr(weights), N.ptr(tokens), N.ptr(cu_seqlens), B, M, max_len,
                                              N.ptr(out_b), N.ptr(out_f), N.ptr(status), N.ptr(ws), ws.numel(),
                                              N.stream_ptr(dev)))
    return out_b, out_f, status


@_bert_encode_packed_op.register_fake
def _(weights, tokens, cu_seqlens, max_len, n_layers, hidden, heads, ffn, vocab, max_pos, pooling, ln_eps):
    B = cu_seqlens.shape[0] - 1
    return (torch.empty((B, hidden), dtype=torch.bfloat16, device=tokens.device),
            torch.empty((B, hidden), dtype=torch.float32, device=tokens.device),
            torch.empty((B,), dtype=torch.int32, device=tokens.device))


def bert_encode_packed(weights: torch.Tensor, tokens: torch.Tensor, cu_seqlens: torch.Tensor, max_len: int, *,
                       n_layers: int, hidden: int, heads: int, ffn: int, vocab: int, max_pos: int,
                       pooling: int = N.BERT_POOL_CLS, ln_eps: float = 1e-12):
    """Variable-length form of bert_encode: tokens int32 [M] back to back, cu_seqlens int32 [B + 1] (prefix sums of the
    lengths), max_len >= the longest sequence.  -> (bf16 [B, H], f32 [B, H], status int32 [B])."""
    return torch.ops.qppfusion.bert_encode_packed(weights, tokens, cu_seqlens, int(max_len), int(n_layers), int(hidden),
                                                  int(heads), int(ffn), int(vocab), int(max_pos), int(pooling),
                                                  float(ln_eps))
