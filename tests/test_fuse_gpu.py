"""GPU parity: qf_fuse (through the C ABI / torch custom op) vs oracle/fusion_oracle.py -- bit exact."""
import gzip
import json
import os

import numpy as np
import pytest
import torch

from helpers import pack_lists, synth_lists
from oracle import fusion_oracle as fo

pytestmark = pytest.mark.gpu

METHODS = {"sum": 0, "mnz": 1, "rrf": 2}
EXTRA_FLAGS = [0]


KERNEL_PATHS = {"stream": 0, "stream-large": 8, "gather": 0x400, "gather-large": 0x408, "general": 0x200, "deep": 0x800}


@pytest.fixture(autouse=True, params=list(KERNEL_PATHS))
def fuse_kernel_path(request):
    """Every test runs six times: through the streaming kernel (+ general-kernel fix-up of deferred queries, the
    product path) and its larger-table configuration (QF_FLAG_LARGE_UNION, a performance hint that must not change any
    result), through the opt-in deposit + gather kernel (flag 0x400, round 2's measured alternative) in both table
    sizes, with both disabled (debug flag 0x200: general kernel alone), and through qf_fuse_deep (ops-level test flag
    0x800: the workspace-resident kernel for requests above the shared-memory envelope), so all kernels keep full
    parity coverage."""
    EXTRA_FLAGS[0] = KERNEL_PATHS[request.param]
    yield
    EXTRA_FLAGS[0] = 0


def run_gpu(lists, method, weights=None, norm=0, rrf_k=60.0, flags=0, score_type="f64", explicit_ranks=False, k=None):
    from qppfusion import ops
    ids, sc, ranks, lens = pack_lists(lists, k=k, score_type=score_type)
    dev = torch.device("cuda:0")
    t = lambda a: torch.from_numpy(a).to(dev)
    w = None if weights is None else t(np.asarray(weights, dtype=np.float64))
    out = ops.fuse(t(ids), t(sc), t(lens), w, t(ranks) if explicit_ranks else None, METHODS[method], norm, rrf_k,
                   flags | EXTRA_FLAGS[0])
    torch.cuda.synchronize()
    return [o.cpu().numpy() for o in out]


def check_against_oracle(lists, method, weights, gpu_out, ranx=False, rrf_k=60.0, mnz_skip_zero=False, norm="min-max"):
    oid, osc, ocnt, opos, olen, ostat = gpu_out
    for q, per_r in enumerate(lists):
        if ranx:
            docs, scores, counts = fo.fuse_ranx_query(per_r, METHODS[method], rrf_k, norm, mnz_skip_zero)
        else:
            docs, scores, counts = fo.fuse_weighted_query(per_r, weights[q], METHODS[method], rrf_k)
        n = len(docs)
        assert olen[q] == n, (q, olen[q], n)
        order = fo.rank_order(scores)
        exp_ids = np.array([docs[i] for i in order], dtype=np.int64)
        exp_sc = np.array([scores[i] for i in order], dtype=np.float64)
        exp_sc = exp_sc + 0.0
        assert np.array_equal(oid[q, :n].astype(np.int64), exp_ids), f"query {q}: ranking differs"
        assert np.array_equal(osc[q, :n].view(np.int64), exp_sc.view(np.int64)), f"query {q}: score bits differ"
        assert np.array_equal(ocnt[q, :n].astype(np.int64), np.array([counts[i] for i in order])), f"query {q}: counts"
        assert np.array_equal(opos[q, :n].astype(np.int64), np.array(order)), f"query {q}: first-seen positions"


@pytest.mark.parametrize("method", ["sum", "mnz", "rrf"])
@pytest.mark.parametrize("score_type", ["f64", "micro"])
def test_weighted_small_cap(method, score_type):
    lists = synth_lists(1, n_q=40, n_r=5, k=100, pool=230, ragged=True, neg_ranker=2, ties=True)
    rng = np.random.default_rng(2)
    w = np.round(rng.random((len(lists), 5)), 6)
    w[0, 1] = 0.0
    out = run_gpu(lists, method, w, score_type=score_type, k=100)
    check_against_oracle(lists, method, w, out)
    assert (out[5] == 0).all()


@pytest.mark.parametrize("method", ["sum", "mnz", "rrf"])
def test_weighted_nq_shape_k1000(method):
    lists = synth_lists(3, n_q=12, n_r=5, k=1000, pool=2200, neg_ranker=2, ties=True)
    rng = np.random.default_rng(4)
    w = np.round(rng.random((len(lists), 5)), 6)
    out = run_gpu(lists, method, w, score_type="micro", k=1000)
    check_against_oracle(lists, method, w, out)


def test_full_union_and_full_overlap_k1000():
    # pool == k: every list holds the same 1000 docs (union 1000); pool = 3200: union ~2700, the top of the streaming
    # kernel's envelope (table ~90 % full); pool = 4000 / 5000: union ~3050 / ~3400, deferred to the general kernel
    for pool in (1000, 3200, 4000, 5000):
        lists = synth_lists(5, n_q=4, n_r=5, k=1000, pool=pool)
        w = np.round(np.random.default_rng(6).random((4, 5)), 6)
        out = run_gpu(lists, "sum", w, score_type="f64", k=1000)
        check_against_oracle(lists, "sum", w, out)


def test_forced_full_sort_path_equals_fast_path():
    lists = synth_lists(7, n_q=6, n_r=5, k=300, pool=700, ties=True)
    w = np.round(np.random.default_rng(8).random((6, 5)), 6)
    a = run_gpu(lists, "sum", w, k=300)
    b = run_gpu(lists, "sum", w, k=300, flags=0x100)
    for x, y in zip(a, b):
        assert np.array_equal(x[:, :1] if x.ndim == 2 else x, y[:, :1] if y.ndim == 2 else y)
    check_against_oracle(lists, "sum", w, b)


def test_near_tie_scores_trigger_exact_fixup():
    # many distinct scores inside one 2^-32 relative window: equal 32-bit key prefix, order decided by
    # the float64 fix-up (short runs) or the 64-bit sort (long runs)
    rng = np.random.default_rng(9)
    lists = []
    for q in range(3):
        n = [40, 200, 900][q]
        docs = rng.choice(10**6, size=n, replace=False)
        base = 1.0
        sc = base + rng.permutation(n) * 2.0 ** -50
        rows = [(int(d), float(s), i + 1) for i, (d, s) in enumerate(zip(docs, sc))]
        anchor = [(int(10**6 + 1), 0.25, 1)]
        lists.append([rows, anchor])
    w = np.ones((3, 2))
    out = run_gpu(lists, "sum", w, k=1000)
    check_against_oracle(lists, "sum", w, out)


def test_duplicates_and_unsorted_rank_column():
    lists = synth_lists(10, n_q=5, n_r=4, k=50, pool=120)
    # duplicate doc id inside one list: every row accumulates and counts (reference weighted_* path)
    d0 = lists[1][2][3][0]
    lists[1][2][9] = (d0, lists[1][2][9][1], lists[1][2][9][2])
    w = np.round(np.random.default_rng(11).random((5, 4)), 6)
    for method in ("sum", "mnz", "rrf"):
        out = run_gpu(lists, method, w, k=50)
        check_against_oracle(lists, method, w, out)
        assert out[5][1] & 1 and not (out[5][0] & 1)
    # explicit rank column (W-RRF uses the file's rank, not the position)
    lists2 = [[[(d, s, rk + 3) for (d, s, rk) in rows] for rows in per_r] for per_r in lists]
    out = run_gpu(lists2, "rrf", w, k=50, explicit_ranks=True)
    check_against_oracle(lists2, "rrf", w, out)
    # descending ranks violate the contract -> flagged, never silently repaired
    lists3 = [[list(reversed(rows)) for rows in per_r] for per_r in lists2]
    out = run_gpu(lists3, "rrf", w, k=50, explicit_ranks=True)
    assert (out[5] & 2).all()


@pytest.mark.parametrize("method", ["sum", "mnz", "rrf"])
def test_ranx_semantics(method):
    lists = synth_lists(12, n_q=20, n_r=5, k=100, pool=230, ragged=True, ties=True)
    lists[0][1] = [(d, 0.7, rk) for (d, s, rk) in lists[0][1]]          # constant list -> 1e-9 guard
    lists[1][0] = lists[1][0][:1]                                        # single-doc list
    out = run_gpu(lists, method, None, norm=(0 if method == "rrf" else 1), flags=2, k=100)
    check_against_oracle(lists, method, None, out, ranx=True)
    if method == "mnz":
        out = run_gpu(lists, method, None, norm=1, flags=2 | 1, k=100)
        check_against_oracle(lists, method, None, out, ranx=True, mnz_skip_zero=True)


@pytest.mark.parametrize("method", ["sum", "mnz"])
@pytest.mark.parametrize("score_type", ["f64", "micro"])
def test_zscore_run_normalisation(method, score_type):
    """QF_NORM_ZSCORE (north star: "min-max/z-normalise each run"): per (query, run) zero mean / unit population
    variance with sequential float64 sums, max(std, 1e-9) guard; bit-exact against the oracle restatement."""
    lists = synth_lists(14, n_q=16, n_r=5, k=100, pool=230, ragged=True, ties=True, neg_ranker=2)
    lists[0][1] = [(d, 0.7, rk) for (d, s, rk) in lists[0][1]]          # constant list -> std 0 -> 1e-9 guard
    lists[1][0] = lists[1][0][:1]                                        # single-doc list
    lists[2][3] = []                                                     # empty list
    out = run_gpu(lists, method, None, norm=2, flags=2, k=100, score_type=score_type)
    check_against_oracle(lists, method, None, out, ranx=True, norm="zscore")
    if method == "mnz":
        out = run_gpu(lists, method, None, norm=2, flags=2 | 1, k=100, score_type=score_type)
        check_against_oracle(lists, method, None, out, ranx=True, mnz_skip_zero=True, norm="zscore")


def test_stream_kernel_envelope_edges():
    """Conditions the streaming kernel defers or handles specially: doc id -1 (its empty-slot sentinel), R = 16 / 17
    (list bit mask width), float32 scores, lists shorter than one 16-byte vector, an all-equal-score query (ranking by
    first-seen order only) and a K that is not a multiple of 4 (scalar loads)."""
    rng = np.random.default_rng(21)
    lists = synth_lists(20, n_q=6, n_r=3, k=64, pool=150, ragged=True, ties=True)
    r_long = max(range(3), key=lambda r: len(lists[2][r]))
    d, s, rk = lists[2][r_long][5]
    lists[2][r_long][5] = (-1, s, rk)                              # sentinel id
    lists[3] = [[(int(x), 0.5, i + 1) for i, x in enumerate(rng.choice(10**6, 40, replace=False))] for _ in range(3)]
    lists[4] = [rows[:3] for rows in lists[4]]
    w = np.round(rng.random((6, 3)), 6)
    w[3] = 0.5                                                     # query 3: every fused score equal
    for method in ("sum", "mnz", "rrf"):
        for st in ("f64", "f32", "micro"):
            ll = lists
            if st == "f32":     # values exactly representable in float32
                ll = [[[(d_, float(np.float32(s_)), r_) for (d_, s_, r_) in rows] for rows in per_r] for per_r in lists]
            out = run_gpu(ll, method, w, score_type=st, k=64)
            check_against_oracle(ll, method, w, out)
    out = run_gpu(lists, "sum", w, k=67)
    check_against_oracle(lists, "sum", w, out)
    for n_r in (16, 17):
        ll = synth_lists(22 + n_r, n_q=3, n_r=n_r, k=40, pool=200, ragged=True)
        ww = np.round(rng.random((3, n_r)), 6)
        for method in ("sum", "mnz"):
            out = run_gpu(ll, method, ww, k=40)
            check_against_oracle(ll, method, ww, out)


def test_unranked_output_is_first_seen_order():
    lists = synth_lists(23, n_q=5, n_r=5, k=200, pool=450, ragged=True, ties=True)
    w = np.round(np.random.default_rng(24).random((5, 5)), 6)
    oid, osc, ocnt, opos, olen, ostat = run_gpu(lists, "mnz", w, k=200, flags=4)
    for q, per_r in enumerate(lists):
        docs, scores, counts = fo.fuse_weighted_query(per_r, w[q], 1, 60.0)
        n = len(docs)
        assert olen[q] == n
        assert np.array_equal(oid[q, :n].astype(np.int64), np.array(docs, dtype=np.int64))
        assert np.array_equal(osc[q, :n].view(np.int64), (np.array(scores, dtype=np.float64) + 0.0).view(np.int64))
        assert np.array_equal(ocnt[q, :n].astype(np.int64), np.array(counts))
        assert np.array_equal(opos[q, :n].astype(np.int64), np.arange(n))


def test_long_tie_runs_and_skewed_scores():
    """Hundreds of exactly tied fused scores (order = first-seen) next to a heavy-tailed score distribution: the
    streaming kernel's bucket ranking must either resolve them exactly or defer to the general kernel."""
    rng = np.random.default_rng(25)
    lists = []
    for q in range(4):
        n = [100, 150, 400, 900][q]
        docs = rng.choice(10**6, size=n, replace=False)
        tied = [(int(d), 0.125, i + 1) for i, d in enumerate(docs)]
        hs = np.sort(np.round(1e3 * rng.pareto(1.5, size=500), 6))[::-1]
        heavy = [(int(d), float(s_), i + 1) for i, (d, s_) in enumerate(zip(rng.choice(10**6, size=500, replace=False), hs))]
        lists.append([tied, heavy, tied[: n // 2]])
    w = np.ones((4, 3))
    for method in ("sum", "mnz", "rrf"):
        out = run_gpu(lists, method, w, k=1000)
        check_against_oracle(lists, method, w, out)


def test_outliers_outside_the_sampled_score_range():
    """The streaming kernel takes the bucket map's range from a sample of the scores: queries whose sample is
    constant while a few unsampled scores differ, and scores far outside the sampled range, must still rank exactly."""
    rng = np.random.default_rng(26)
    lists = []
    for q in range(12):
        n = int(rng.integers(100, 155))
        docs = rng.choice(10**6, size=n + 3, replace=False)
        rows = [(int(docs[0]), 7.5, 1), (int(docs[1]), 0.9, 2)]
        rows += [(int(d), 0.5, i + 3) for i, d in enumerate(docs[3:])]
        rows.append((int(docs[2]), -4.0, n + 3))
        lists.append([rows, rows[:1] + rows[-1:]])
    for q in range(4):
        n = 1500 + 100 * q
        docs = rng.choice(10**6, size=n, replace=False)
        sc = np.sort(np.concatenate([[1e6, 5e5, -1e6], np.round(rng.random(n - 3), 6)]))[::-1]
        rows = [(int(d), float(s_), i + 1) for i, (d, s_) in enumerate(zip(docs, sc))]
        lists.append([rows, rows[::7]])
    w = np.ones((len(lists), 2))
    for method in ("sum", "mnz"):
        out = run_gpu(lists, method, w, k=2000)
        check_against_oracle(lists, method, w, out)


def test_empty_queries_and_bad_lens():
    lists = synth_lists(13, n_q=3, n_r=3, k=10, pool=30)
    lists[1] = [[], [], []]
    w = np.ones((3, 3))
    out = run_gpu(lists, "sum", w, k=10)
    check_against_oracle(lists, "sum", w, out)
    assert out[4][1] == 0


def test_golden_reference_vectors_through_cuda():
    """The real reference's outputs (tests/golden/fusion_weighted.json.gz) reproduced bit-exactly by CUDA."""
    path = os.path.join(os.path.dirname(__file__), "golden", "fusion_weighted.json.gz")
    with gzip.open(path, "rt") as f:
        gold = json.load(f)
    names = {"wcombsum": ("sum", 60), "wcombmnz": ("mnz", 60), "wrrf": ("rrf", 60), "wrrf_k10": ("rrf", 10)}
    for case in gold["cases"]:
        rankers = gold["rankers"]
        per = {rk: {} for rk in rankers}
        doc_index = {}
        for rk in rankers:
            for (qid, _it, docno, rank, score, _tag) in case["rows"][rk]:
                did = doc_index.setdefault(docno, len(doc_index))
                per[rk].setdefault(qid, []).append((did, float(score), int(rank)))
        inv = {v: k for k, v in doc_index.items()}
        qids = sorted(set().union(*[set(p.keys()) for p in per.values()]))
        for key, res in case["results"].items():
            mname, idx = key.split("@")
            idx = int(idx)
            method, k = names[mname]
            lists = []
            for qid in qids:
                per_r = [per[rk].get(qid, []) for rk in rankers]
                if method == "rrf":   # pandas sort_values("rank") before visiting (src/fusion.py:371)
                    per_r = [sorted(rows, key=lambda t: t[2]) for rows in per_r]
                lists.append(per_r)
            w = np.array([[(sum(case["qpp"][q][rk]) / 13) if idx == -1 else case["qpp"][q][rk][idx]
                           for rk in rankers] for q in qids])
            oid, osc, ocnt, opos, olen, ostat = run_gpu(lists, method, w, rrf_k=float(k), explicit_ranks=True)
            for qi, qid in enumerate(qids):
                ref = res["fused"].get(qid, [])
                n = olen[qi]
                assert n == len(ref)
                # rebuild the reference's first-seen-order return value from (pos, ids, scores)
                got = [None] * n
                for j in range(n):
                    got[opos[qi, j]] = (inv[int(oid[qi, j])], float(osc[qi, j]))
                exp = [(d, float(s) + 0.0) for d, s in ref]
                assert got == exp, (case["name"], key, qid)


def test_compact_fused_csr():
    from qppfusion import ops
    lists = synth_lists(41, n_q=300, n_r=5, k=40, pool=90, ragged=True)
    lists[7] = [[], [], [], [], []]
    ids, sc, ranks, lens = pack_lists(lists, k=40, score_type="micro")
    dev = torch.device("cuda:0")
    t = lambda a: torch.from_numpy(a).to(dev)
    w = torch.ones((300, 5), dtype=torch.float64, device=dev)
    out = ops.fuse(t(ids), t(sc), t(lens), w, None, 0, 0, 60.0, 0)
    off, cids, csc = ops.compact_fused(out)
    off = off.cpu().numpy()
    n = out.lens.cpu().numpy()
    assert off[0] == 0 and np.array_equal(np.diff(off), n)
    oi, os_ = out.ids.cpu().numpy(), out.scores.cpu().numpy()
    ci, cs = cids.cpu().numpy(), csc.cpu().numpy()
    for q in range(300):
        assert np.array_equal(ci[off[q]:off[q + 1]], oi[q, :n[q]]) and np.array_equal(cs[off[q]:off[q + 1]], os_[q, :n[q]])


def _same_group_ids(G, group, count, start=1):
    """doc ids whose home group in a streaming-kernel table of G groups is `group` (hash of qf_fuse_stream.cu)."""
    out, d = [], start
    while len(out) < count:
        if ((((d * 2654435761) & 0xFFFFFFFF) * G) >> 32) == group:
            out.append(d)
        d += 1
    return out


@pytest.mark.parametrize("k,groups", [(400, (768, 1280)), (100, (384,))])
def test_duplicate_doc_ids_in_the_first_list(k, groups):
    """The first non-empty list is inserted without probing (per-group fill counters); a doc id that occurs twice in
    it must still end in the exact sequential path: both copies inside one group, one inside and one overflowed, both
    overflowed (7 docs share a home group), and the same with a leading empty list."""
    rng = np.random.default_rng(77)
    base = synth_lists(31, n_q=8, n_r=3, k=k, pool=int(2.2 * k))
    lists = []
    for q, per_r in enumerate(base):
        per_r = [list(rows) for rows in per_r]
        G = groups[q % len(groups)]
        crowd = _same_group_ids(G, 5 + q, 7, start=2_000_000)
        rows0 = per_r[0]
        variant = q % 4
        if variant == 0:        # dup pair among the first four of the group
            pick = [crowd[0], crowd[1], crowd[0], crowd[2], crowd[3], crowd[4], crowd[5]]
        elif variant == 1:      # one copy inside the group, one overflowed
            pick = [crowd[0], crowd[1], crowd[2], crowd[3], crowd[4], crowd[0], crowd[5]]
        elif variant == 2:      # both copies overflowed
            pick = [crowd[0], crowd[1], crowd[2], crowd[3], crowd[4], crowd[5], crowd[4]]
        else:                   # crowded group without a duplicate, later lists hit the overflowed docs too
            pick = crowd
        at = sorted(rng.choice(len(rows0), size=len(pick), replace=False))
        for i, d in zip(at, pick):
            rows0[i] = (int(d), rows0[i][1], rows0[i][2])
        for r in (1, 2):
            j = int(rng.integers(0, len(per_r[r])))
            per_r[r][j] = (int(crowd[4 + (r - 1)]), per_r[r][j][1], per_r[r][j][2])
        if q >= 4:
            per_r = [[]] + per_r[:2]                    # first list empty: list 1 takes the no-probe path
        lists.append(per_r)
    w = np.round(rng.random((len(lists), 3)), 6)
    for method in ("sum", "mnz"):
        out = run_gpu(lists, method, w, k=k)
        check_against_oracle(lists, method, w, out)


def test_repeated_launches_are_bitwise_identical():
    """The join claims table keys with racy-by-design shared-memory CAS (stale volatile views, lost claims re-queued):
    the ARRIVAL order of postings differs from launch to launch, the result must not.  40 launches of the same
    NQ-shaped batch (1 024 queries x 5 x top-1000, ~44 % union, ties included) through the current kernel path must
    reproduce the first launch bit for bit (compute-sanitizer is closed on this pool; this is the stand-in for a
    racecheck run: a hazard that matters shows up as run-to-run differences)."""
    from qppfusion import ops
    import bench
    dev = torch.device("cuda:0")
    ids, micro, lens, _f = bench.synth_device(1024, 5, 1000, 77, dev)
    micro = (micro // 1000) * 1000                                      # 3-decimal scores: many exact ties
    w = (torch.rand((1024, 5), generator=torch.Generator().manual_seed(3), dtype=torch.float64) * 1e6).round() / 1e6
    w = w.to(dev)
    for method in (0, 1):
        first = None
        for it in range(40):
            out = ops.fuse(ids, micro, lens, w, None, method, 0, 60.0, EXTRA_FLAGS[0])
            cur = [out.ids.clone(), out.scores.clone().view(torch.int64), out.counts.to(torch.int32), out.pos.to(torch.int32), out.lens.clone()]
            n = out.lens.long()
            m = torch.arange(out.ids.shape[1], device=dev)[None, :] < n[:, None]
            cur = [c[m] if c.dim() == 2 else c for c in cur]
            if first is None:
                first = cur
            else:
                assert all(torch.equal(a, b) for a, b in zip(first, cur)), f"launch {it} differs (method {method})"


def test_six_runs_of_depth_1000_fit_the_wide_general_kernel():
    """6 retrievers x TREC depth 1000 = 6 000 postings per query worked in the reference and failed here.
    4-byte scores go up to 6 144 postings per query in shared memory (general kernel alone); float64 scores above
    5 120 take qf_fuse_deep (round 2) instead of raising."""
    lists = synth_lists(51, n_q=6, n_r=6, k=1000, pool=2600, ties=True)
    w = np.round(np.random.default_rng(2).random((6, 6)), 6)
    for method in ("sum", "mnz", "rrf"):
        out = run_gpu(lists, method, w, k=1000, score_type="micro")
        check_against_oracle(lists, method, w, out)
    out = run_gpu(lists, "sum", w, k=1000, score_type="f64")
    check_against_oracle(lists, "sum", w, out)


@pytest.mark.parametrize("n_r,k", [(8, 2000), (20, 1000), (32, 2047)])
def test_deep_runs_beyond_the_shared_memory_envelope(n_r, k):
    """the reference fuses arbitrary depth.  16 000 / 20 000 / 65 504 postings per query go through
    qf_fuse_deep (hash table and sort buffers in a device workspace): weighted sum / mnz / rrf with ties, a duplicate
    doc id inside a list, ragged and empty lists, and the ranx semantics (min-max, z-score, REQUIRE_SORTED), bit-exact."""
    if EXTRA_FLAGS[0] not in (0, 0x800):
        pytest.skip("size-routed: the kernel-path flags do not apply")
    nq = 3
    lists = synth_lists(60 + n_r, n_q=nq, n_r=n_r, k=k, pool=int(2.4 * k), ragged=True, ties=True, neg_ranker=1)
    lists[1][2] = []
    rl = max(range(n_r), key=lambda r: len(lists[2][r]))
    d0 = lists[2][rl][3][0]
    lists[2][rl][40] = (d0, lists[2][rl][40][1], lists[2][rl][40][2])       # duplicate inside one list
    w = np.round(np.random.default_rng(5).random((nq, n_r)), 6)
    for method in ("sum", "mnz", "rrf"):
        out = run_gpu(lists, method, w, k=k, score_type="micro" if method != "mnz" else "f64")
        check_against_oracle(lists, method, w, out)
        assert out[5][2] & 1 and not (out[5][0] & 1)
    clean = synth_lists(90 + n_r, n_q=2, n_r=n_r, k=k, pool=int(2.4 * k), ragged=True, ties=True)
    for method, norm in (("sum", 1), ("mnz", 2), ("rrf", 0)):
        out = run_gpu(clean, method, None, norm=norm, flags=2, k=k)
        check_against_oracle(clean, method, None, out, ranx=True, norm="zscore" if norm == 2 else "min-max")


def test_deep_limit_is_reported():
    from qppfusion import ops
    dev = torch.device("cuda:0")
    ids = torch.zeros((1, 33, 2000), dtype=torch.int32, device=dev)
    with pytest.raises(ValueError, match="65535"):
        ops.fuse(ids, ids.clone(), torch.zeros((1, 33), dtype=torch.int32, device=dev))
