#!/usr/bin/env python3
"""
bench.py -- headline benchmark of the B200-native QPP-Fusion-RAG hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one pass of the hot path over one NQ-shaped batch (BASELINE.json configs[1]):
    3,452 queries x 5 retrievers x top-1000 run lists ->
      qf_qpp_stats (13 QPPBridge predictors per list) -> qf_qpp_colnorm (cross-query min-max + "%.6f")
      -> qf_qpp_select_weight (RSD column, method id 5) -> qf_fuse (W-CombSUM, ranked)
metric = fused queries/s.  `value` has the inputs resident in HBM; `e2e` goes through host (pinned)
buffers with the H2D / D2H copies inside the timed region.  Synthetic data (no network): seeded, generated
on the device in the BEIR/NQ shape described in SURVEY.md section 8(d).

N > 1 (torchrun, one rank per GPU, NCCL): queries are sharded across ranks (weak scaling: every rank
fuses its own 3,452-query shard); QPP/fusion need no data-path collective.

--impl reference times the reference's CPU implementation of the same step (oracle port of
src/fusion.py weighted_combsum + QPPBridge.java; see oracle/) on a bounded query sample per step.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "qpp-fusion-rag_b200"), os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

Q_NQ, R_NQ, K_NQ, N_DOCS = 3452, 5, 1000, 2_681_468
POOL = 2200                      # per-query candidate pool -> union ratio ~0.45
METRIC = "fused queries/s (5x top-1000) + BGE top-1000 QPS @2.68M docs; % of roofline"


# ------------------------------------------------------------------------------------------------ data
def synth_device(Q, R, K, seed, device):
    """Seeded synthetic NQ-shaped runs, generated on the device.
    doc_ids int32 [Q,R,K] (distinct inside a list), scores int32 micro-units [Q,R,K] sorted descending
    (6-decimal values as .res files carry them), lens int32 [Q,R], text_feats int32 [Q,4]."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    stride = N_DOCS // POOL
    pool = (torch.arange(POOL, device=device, dtype=torch.int64)[None, :] * stride
            + torch.randint(0, stride, (Q, POOL), generator=g, device=device))            # distinct per query
    ids = torch.empty((Q, R, K), dtype=torch.int32, device=device)
    micro = torch.empty((Q, R, K), dtype=torch.int32, device=device)
    scale = [1.0, 30.0, 1.0, 1.0, 12.0]
    for r in range(R):
        pick = torch.rand((Q, POOL), generator=g, device=device).argsort(dim=1)[:, :K]
        ids[:, r] = torch.gather(pool, 1, pick).to(torch.int32)
        u1 = torch.rand((Q, K), generator=g, device=device, dtype=torch.float64)
        u2 = torch.rand((Q, K), generator=g, device=device, dtype=torch.float64)
        s = (1.0 - u1.sqrt()) * (0.35 + 0.65 * u2)          # right-skewed in (0,1), Beta(2,5)-like
        s = s * scale[r % 5]
        if r % 5 == 2:
            s = -3.0 * (-torch.log1p(-u2))                   # MonoT5-style negative log-probs
        s = torch.sort(s, dim=1, descending=True).values
        micro[:, r] = torch.round(s * 1e6).to(torch.int32)
    lens = torch.full((Q, R), K, dtype=torch.int32, device=device)
    n_tok = (torch.poisson(torch.full((Q,), 9.0, device=device), generator=g) + 1).to(torch.int32)
    n_unique = torch.clamp(n_tok - (torch.rand((Q,), generator=g, device=device) < 0.2).to(torch.int32), min=1)
    max_len = torch.randint(4, 14, (Q,), generator=g, device=device, dtype=torch.int32)
    sum_len = (n_tok.to(torch.float32) * (3.5 + 2.0 * torch.rand((Q,), generator=g, device=device))).to(torch.int32)
    feats = torch.stack([n_tok, n_unique, max_len, sum_len], dim=1).contiguous()
    return ids.contiguous(), micro.contiguous(), lens, feats


def write_synth_run_files(root, nq, seed=1, R=R_NQ, K=K_NQ):
    """Synthetic run / QPP files in the reference's wire formats (BASELINE configs[1] shape, a sample of `nq` queries):
    <root>/runs/<Ranker>.norm.res  "{qid} Q0 {docno} {rank} {score:.6f} {tag}"  (scripts/02_retrieve.py:80)
    <root>/qpp/<Ranker>.res.mmnorm.qpp  "qid\t13 x %.6f"                         (src/qpp.py:331-334)
    -> (runs_dir, qpp_dir).  Written with numpy/pandas on the host, outside every timed region."""
    import numpy as np
    import pandas as pd
    names = ["BGE", "BM25", "BM25_MonoT5", "BM25_TCT", "Splade"][:R]
    runs_dir, qpp_dir = os.path.join(root, "runs"), os.path.join(root, "qpp")
    os.makedirs(runs_dir, exist_ok=True)
    os.makedirs(qpp_dir, exist_ok=True)
    ids, micro, _lens, _feats = _cpu_sample_arrays(nq, seed)
    rng = np.random.default_rng(seed + 77)
    qid = np.repeat(np.arange(nq), K)
    rank = np.tile(np.arange(1, K + 1), nq)
    for r, name in enumerate(names):
        df = pd.DataFrame({"qid": np.char.add("test", qid.astype(str)), "q0": "Q0",
                           "docno": np.char.add("doc", ids[:, r].reshape(-1).astype(str)), "rank": rank,
                           "score": micro[:, r].reshape(-1) / 1e6, "tag": name})
        df.to_csv(os.path.join(runs_dir, f"{name}.norm.res"), sep=" ", header=False, index=False, float_format="%.6f")
        q = np.round(rng.random((nq, 13)), 6)
        with open(os.path.join(qpp_dir, f"{name}.res.mmnorm.qpp"), "w") as f:
            for i in range(nq):
                f.write(f"test{i}\t" + "\t".join(f"{v:.6f}" for v in q[i]) + "\n")
    return runs_dir, qpp_dir


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """SM clock + throttle reasons sampled WHILE the timed region runs: NVML polled every ~2 ms from a thread (the
    timed regions here last 10-150 ms, far below nvidia-smi's sampling period); falls back to `nvidia-smi -lms` when
    NVML is not importable."""
    REASONS = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}

    def __init__(self, index):
        self.index, self.rows, self.proc, self.thread, self.stop_flag = index, [], None, None, False
        self.nvml = None
        try:
            import pynvml
            import torch
            pynvml.nvmlInit()
            uuid = str(torch.cuda.get_device_properties(index).uuid)
            try:
                self.handle = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid).encode())
            except Exception:
                self.handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def _poll(self):
        nv = self.nvml
        while not self.stop_flag:
            try:
                mhz = float(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM))
                try:
                    mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
                except Exception:
                    mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
                self.rows.append((mhz, mask))
            except Exception:
                pass
            time.sleep(0.002)

    def start(self):
        if self.nvml is not None:
            self.stop_flag = False
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        try:
            fields = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
                      "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={fields}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.nvml is not None:
            self.stop_flag = True
            self.thread.join(timeout=1.0)
            sm = sorted(r[0] for r in self.rows)
            mask = 0
            for r in self.rows:
                mask |= r[1]
            reasons = sorted(n for n, bit in self.REASONS.items() if mask & bit)
            return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.max_mhz, "samples": len(sm),
                    "reasons": reasons, "source": "nvml polled during the timed region"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for row in self.rows:
            try:
                sm.append(float(row[0]))
                mx.append(float(row[1]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, row[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons), "source": "nvidia-smi -lms 100"}


# ------------------------------------------------------------------------------------------------ ours
def fuse_traffic():
    """DRAM bytes per launch of the fusion kernel from the committed ncu --set full capture (profiles/)."""
    p = os.path.join(ROOT, "profiles", "fuse_traffic.json")
    if os.path.isfile(p):
        return json.load(open(p)).get("dram_bytes_per_launch")
    return None


def dense_traffic():
    """DRAM bytes of the dense main pass (NQ shape, one GPU) from the committed ncu --set full capture."""
    p = os.path.join(ROOT, "profiles", "dense_traffic.json")
    if os.path.isfile(p):
        return json.load(open(p)).get("dram_bytes_per_launch")
    return None


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops", 1590.0), "measured"
    return 6650.0, 1590.0, "fallback"



N_DENSE, D_DENSE, K_DENSE = 2_681_468, 768, 1000


def synth_dense(Q, N, D, seed, device):
    """Unit-normalised N(0,1) rows -> bf16, generated on the device in blocks."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    corpus = torch.empty((N, D), dtype=torch.bfloat16, device=device)
    for s in range(0, N, 262144):
        x = torch.randn((min(262144, N - s), D), generator=g, device=device)
        corpus[s:s + x.shape[0]] = (x / x.norm(dim=1, keepdim=True)).to(torch.bfloat16)
    x = torch.randn((Q, D), generator=g, device=device)
    q32 = (x / x.norm(dim=1, keepdim=True)).contiguous()
    return q32, corpus


def dense_cpu_baseline(q32_cpu, corpus_rows_cpu, n_docs_total, k, seconds=6.0):
    """Flat inner-product search on the host cores, the way the reference's faiss.IndexFlatIP path computes it
    (src/retrievers/bge.py:188-194: fp32 GEMM + top-k; faiss itself is not installable offline): torch.mm fp32 +
    torch.topk on a bounded sample -- `q32_cpu` queries against `corpus_rows_cpu` corpus rows -- extrapolated linearly
    in the corpus size (a flat search is linear in N).  -> dict for the JSON line."""
    import time
    import torch
    nq, n_sub = q32_cpu.shape[0], corpus_rows_cpu.shape[0]
    c32 = corpus_rows_cpu.float()
    kk = min(k, n_sub)
    torch.topk(q32_cpu[:4] @ c32[:4096].T, min(kk, 4096), dim=1)          # warm the thread pool
    reps, t0 = 0, time.perf_counter()
    while True:
        torch.topk(q32_cpu @ c32.T, kk, dim=1)
        reps += 1
        dt = time.perf_counter() - t0
        if dt >= seconds or reps >= 20:
            break
    per_query_full = dt / reps / nq * (n_docs_total / n_sub)
    return {"value": 1.0 / per_query_full, "unit": "queries/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{nq} queries x {n_sub} of {n_docs_total} corpus rows, torch.mm fp32 + torch.topk on the host "
                      f"(restated flat search; faiss-cpu is not installable offline), {reps} repetitions, extrapolated "
                      f"linearly to the full corpus"}


def dense_leg(args, dev, world, rank, Q=Q_NQ, n_docs=None, label="BASELINE configs[2]", max_steps=10):
    """BGE-shaped flat inner-product top-1000 (BASELINE configs[2] by default; configs[3] = HotpotQA shape with
    Q=7405, n_docs=5,233,329); corpus row-sharded when world > 1."""
    import torch
    import torch.distributed as dist
    from qppfusion import _native as N_, ops
    N_DENSE = n_docs or globals()["N_DENSE"]
    D, k = D_DENSE, K_DENSE
    n_local = (N_DENSE + world - 1) // world
    lo = rank * n_local
    n_local = max(0, min(n_local, N_DENSE - lo))
    q32, corpus = synth_dense(Q, n_local, D, 20250103 + rank, dev)
    if world > 1:
        dist.broadcast(q32, src=0)
    qb = q32.to(torch.bfloat16)
    stream = torch.cuda.current_stream(dev)
    steps = max(3, min(args.steps, max_steps))

    from qppfusion import distributed as qd

    def step():
        if world > 1:        # local top-k_in -> ONE all-to-all -> merge Q/G queries per rank -> ONE all-gather
            res = qd.dense_search_packed(qb, corpus, k, lo, gather=True)
            return res, res.pairs, res.flags
        # the status vector is read when the result is consumed (after the timed region: every step searches the same
        # batch), as on the sharded path -- reading it inside the call would synchronise the host once per step
        return ops.dense_topk(qb, corpus, k, lo, exact_fallback=False)

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # consecutive searches are independent (different query batches in a serving loop): INFLIGHT of them are in flight on
    # their own streams, so that step i+1's sample / threshold / main pass runs while step i's exchange, merge and
    # launch gaps drain (the NCCL calls keep their issue order on every rank).  The per-kernel times behind the roofline
    # come from a separate SERIAL loop -- overlapped kernels would time each other's waiting.
    # Measured (tools/dense_leg_probe.py, 1 and 2 B200): two searches in flight LOWER the device-resident rate (312 k ->
    # 307 k, 604 k -> 589 k QPS: the persistent main-pass CTAs lose SMs to the next step's small kernels) and RAISE
    # the end-to-end rate (295 k -> 301 k, 560 k -> 585 k: query upload and result download overlap the next search),
    # so `value` runs one search at a time and `e2e` two.
    INFLIGHT_V = max(1, int(os.environ.get("BENCH_DENSE_INFLIGHT", "1")))
    INFLIGHT = max(2, INFLIGHT_V)
    run_streams = [torch.cuda.Stream(dev) for _ in range(INFLIGHT)]

    def run_steps(fn, n, inflight=INFLIGHT):
        for s_ in run_streams:
            s_.wait_stream(stream)
        last = None
        for i in range(n):
            with torch.cuda.stream(run_streams[i % inflight]):
                last = fn()
        for s_ in run_streams:
            stream.wait_stream(s_)
        return last

    sc, ids, st = run_steps(step, 3, INFLIGHT_V)
    barrier()
    status_bad = int((st != 0).sum().item())
    if world > 1:
        status_bad += int((sc.incomplete != 0).sum().item())      # the truncated exchange flagged a possibly lost document
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler = ClockSampler(dev.index)
    if rank == 0:
        sampler.start()
    barrier()
    e0.record(stream)
    sc, ids, st = run_steps(step, steps, INFLIGHT_V)
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1) / steps
    N_.profile_enable(True)          # averages the in-library event pairs of the serial calls from here on
    for _ in range(max(3, steps // 2)):
        sc, ids, st = step()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    phases = {n: N_.profile_read(i) for i, n in enumerate(["sample", "tau", "main", "select"])}
    N_.profile_enable(False)
    main_ms = phases["main"]

    # end to end: pinned fp32 queries -> H2D -> bf16 -> search -> D2H (ids, scores); the index stays resident
    h_q = q32.cpu().pin_memory()
    h_s = torch.empty((Q, k), dtype=torch.float32).pin_memory()
    h_i = torch.empty((Q, k), dtype=torch.int32).pin_memory()

    rows_slice = (Q + world - 1) // world
    h_pairs = torch.empty((rows_slice + 1, 2, k), dtype=torch.int32).pin_memory()
    h_flags = torch.empty((world,), dtype=torch.int32).pin_memory()

    def e2e_step():
        dq = h_q.to(dev, non_blocking=True).to(torch.bfloat16)
        if world > 1:
            # one exchange, and every rank downloads only ITS merged query slice (+ the G status flags): Q*k*8 bytes
            # leave the GPUs in total instead of G times that
            res = qd.dense_search_packed(dq, corpus, k, lo, gather=False)
            h_pairs.copy_(res.pairs, non_blocking=True)                # merged slice + its spare row (truncation flag)
            h_flags.copy_(res.flags, non_blocking=True)
        else:
            s_, i_, st_ = ops.dense_topk(dq, corpus, k, lo, exact_fallback=False)
            h_s.copy_(s_, non_blocking=True)
            h_i.copy_(i_, non_blocking=True)
            h_flags[:1].copy_(st_.max().reshape(1), non_blocking=True)     # status travels with the result

    h_bufs = [(h_s, h_i, h_pairs, h_flags)] + [tuple(torch.empty_like(x).pin_memory() for x in (h_s, h_i, h_pairs, h_flags))
                                                for _ in range(INFLIGHT - 1)]
    e2e_i = [0]

    def e2e_pipelined():
        nonlocal h_s, h_i, h_pairs, h_flags
        h_s, h_i, h_pairs, h_flags = h_bufs[e2e_i[0] % INFLIGHT]      # one set of pinned result buffers per stream
        e2e_i[0] += 1
        e2e_step()

    run_steps(e2e_pipelined, INFLIGHT)
    barrier()
    e0.record(stream)
    run_steps(e2e_pipelined, steps)
    e1.record(stream)
    barrier()
    e2e_ms = e0.elapsed_time(e1) / steps
    for hb in h_bufs:
        assert int(hb[3][:1 if world == 1 else world].max()) == 0, "a sampled threshold failed during the e2e steps"
        if world > 1:
            assert int(hb[2][rows_slice, 0, 0]) == 0, "the truncated exchange flagged a possibly lost document"

    # parity spot check on the device copy (torch fp32 matmul as an independent reference; world == 1 only)
    parity = None
    if world == 1:
        nq = 8
        ref = q32[:nq] @ corpus.float().T if n_local <= 3_000_000 else None
        if ref is not None:
            rs, ri = torch.topk(ref, k, dim=1)
            overlap = sum(len(set(ri[i].tolist()) & set(ids[i].tolist())) for i in range(nq)) / (nq * k)
            parity = {"queries_checked": nq, "top1000_overlap_vs_fp32": overlap,
                      "max_abs_score_diff_vs_fp32": float((rs - sc[:nq]).abs().max().item())}
    if world > 1 and n_local <= 3_000_000:
        # sharded search vs an independent fp32 reference: every rank scores 8 queries against its shard with a torch
        # matmul, the per-shard top-k are gathered and merged, and rank 0 compares with the sharded result
        nq = 8
        rs, ri = torch.topk(q32[:nq] @ corpus.float().T, min(k, n_local), dim=1)
        if rs.shape[1] < k:
            rs = torch.cat([rs, torch.full((nq, k - rs.shape[1]), float("-inf"), device=dev)], dim=1)
            ri = torch.cat([ri, torch.full((nq, k - ri.shape[1]), -1, device=dev, dtype=ri.dtype)], dim=1)
        ri = ri + lo
        all_s = [torch.empty_like(rs) for _ in range(world)]
        all_i = [torch.empty_like(ri) for _ in range(world)]
        dist.all_gather(all_s, rs.contiguous())
        dist.all_gather(all_i, ri.contiguous())
        if rank == 0:
            cs, ci = torch.cat(all_s, dim=1), torch.cat(all_i, dim=1)
            ms_, sel = torch.topk(cs, k, dim=1)
            mi_ = torch.gather(ci, 1, sel)
            got_s, got_i = sc.scores[:nq], sc.ids[:nq]
            overlap = sum(len(set(mi_[i].tolist()) & set(got_i[i].tolist())) for i in range(nq)) / (nq * k)
            parity = {"queries_checked": nq, "top1000_overlap_vs_fp32": overlap,
                      "max_abs_score_diff_vs_fp32": float((ms_ - got_s).abs().max().item()),
                      "reference": "per-shard torch fp32 matmul + top-k, gathered and merged on rank 0"}
    dense_cpu = None
    if world == 1 and n_docs is None and not args.no_cpu_baseline:      # configs[2] only, rank 0, bounded (~6 s)
        try:
            n_sub = min(n_local, 1 << 18)
            dense_cpu = dense_cpu_baseline(q32[:64].cpu(), corpus[:n_sub].cpu(), N_DENSE, k)
        except Exception as e:                                             # never let the side measurement sink the run
            dense_cpu = {"error": f"{type(e).__name__}: {e}"}
    t = torch.tensor([ms, e2e_ms, main_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, e2e_ms, main_ms = t.tolist()
    _hbm, tf_peak, which = load_peaks()
    flops_main = 2.0 * Q * n_local * D           # per GPU, main pass
    ach = flops_main / (main_ms * 1e-3) / 1e12
    return {"metric": "BGE top-1000 QPS", "value": Q / (ms * 1e-3), "unit": "queries/s", "ms_per_step": ms,
            "scaling": "strong" if world > 1 else "n/a",
            "config": {"workload": f"BGE-base-shaped flat IP top-1000 ({label})", "queries": Q,
                       "docs": N_DENSE, "docs_per_gpu": n_local, "dim": D, "k": k, "dtype": "bf16 in, fp32 accumulate",
                       "steps_in_flight": INFLIGHT_V, "e2e_steps_in_flight": INFLIGHT,
                       "sharding": f"corpus rows x{world}" + (f", packed (score,id) rows truncated to the shard's top {qd.send_depth(k, world)}"
                                                              ": ONE all_to_all + qf_topk_merge_packed(Q/G) + completeness check "
                                                              "+ ONE all_gather, no host sync" if world > 1 else "")},
            "kernel_ms": phases,
            "roofline": {"kernel": "dense_kernel (tcgen05 main pass)", "bound": "tensor", "achieved": ach, "peak": tf_peak,
                         "unit": "TFLOP/s", "frac": ach / tf_peak, "peak_source": which,
                         "algorithmic_flops": flops_main,
                         "traffic": dense_traffic() if (world == 1 and n_docs is None) else None},
            "whole_call_tflops": 2.0 * Q * N_DENSE * D / (ms * 1e-3) / 1e12,
            "e2e": {"value": Q / (e2e_ms * 1e-3), "unit": "queries/s", "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": Q * D * 4,
                    "d2h_bytes_per_step": (rows_slice * k * 8 + world * 4) if world > 1 else Q * k * 8,
                    "note": ("every rank uploads the queries and downloads its merged Q/G slice" if world > 1 else None)},
            "cpu_baseline": dense_cpu,
            "threshold_failures": status_bad, "parity": parity, "gpu_launches": (4 + (1 if world > 1 else 0)) * (2 * steps + max(3, steps // 2) + 3 + INFLIGHT), "clocks": clocks}


def encoder_leg(args, dev, world, rank, Q=Q_NQ):
    """BGE-base-shaped query encoder (12 x 768, random-init bf16 weights): Q synthetic queries of 5..32 tokens
    (Poisson(11) + 3, the NQ query length profile), query-sharded across ranks, no collective.  kernel = one padded
    qf_bert_encode call on resident tokens; e2e = host token rows -> packed tokens -> H2D -> qf_bert_encode_packed -> D2H."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from qppfusion import ops
    from qppfusion.encoder import BertConfig, BertQueryEncoder
    cfg = BertConfig()
    g = torch.Generator(device=dev).manual_seed(20250105)
    H, F = cfg.hidden, cfg.ffn

    def rnd(*shape, scale):
        return torch.randn(shape, generator=g, device=dev) * scale

    def ln(prefix, sd):
        sd[prefix + ".weight"] = 1.0 + rnd(H, scale=0.1)
        sd[prefix + ".bias"] = rnd(H, scale=0.1)

    sd = {"embeddings.word_embeddings.weight": rnd(cfg.vocab, H, scale=0.5),
          "embeddings.position_embeddings.weight": rnd(cfg.max_pos, H, scale=0.2),
          "embeddings.token_type_embeddings.weight": rnd(2, H, scale=0.2)}
    ln("embeddings.LayerNorm", sd)
    for l in range(cfg.n_layers):
        p = f"encoder.layer.{l}."
        for n_ in ("query", "key", "value"):
            sd[p + f"attention.self.{n_}.weight"] = rnd(H, H, scale=0.08)
            sd[p + f"attention.self.{n_}.bias"] = rnd(H, scale=0.05)
        sd[p + "attention.output.dense.weight"] = rnd(H, H, scale=0.05)
        sd[p + "attention.output.dense.bias"] = rnd(H, scale=0.05)
        ln(p + "attention.output.LayerNorm", sd)
        sd[p + "intermediate.dense.weight"] = rnd(F, H, scale=0.05)
        sd[p + "intermediate.dense.bias"] = rnd(F, scale=0.05)
        sd[p + "output.dense.weight"] = rnd(H, F, scale=0.05)
        sd[p + "output.dense.bias"] = rnd(H, scale=0.05)
        ln(p + "output.LayerNorm", sd)
    enc = BertQueryEncoder(sd, cfg, None, device=dev)
    del sd
    q_local = (Q + world - 1) // world
    q_local = max(0, min(q_local, Q - rank * q_local))
    rng = np.random.default_rng(20250105 + rank)
    L = 32
    lens = np.clip(rng.poisson(11, size=q_local) + 3, 3, L).astype(np.int32)
    tokens = rng.integers(0, cfg.vocab, size=(q_local, L)).astype(np.int32)
    rows = [tokens[i, :lens[i]] for i in range(q_local)]
    t_d, l_d = torch.from_numpy(tokens).to(dev), torch.from_numpy(lens).to(dev)
    stream = torch.cuda.current_stream(dev)
    steps = max(3, min(args.steps, 10))
    kw = dict(n_layers=cfg.n_layers, hidden=H, heads=cfg.heads, ffn=F, vocab=cfg.vocab, max_pos=cfg.max_pos)

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(3):
        ops.bert_encode(enc.weights, t_d, l_d, **kw)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for _ in range(steps):
        ops.bert_encode(enc.weights, t_d, l_d, **kw)
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1) / steps
    h_out = torch.empty((q_local, H), dtype=torch.float32).pin_memory()
    enc.encode_token_lists(rows)
    barrier()
    e0.record(stream)
    for _ in range(steps):
        h_out.copy_(enc.encode_token_lists(rows)[1], non_blocking=True)
    e1.record(stream)
    barrier()
    e2e_ms = e0.elapsed_time(e1) / steps
    t = torch.tensor([ms, e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, e2e_ms = t.tolist()
    flop_tok = 2.0 * cfg.n_layers * (4 * H * H + 2 * H * F)
    return {"metric": "BGE query-encoder QPS", "value": Q / (ms * 1e-3), "unit": "queries/s", "ms_per_step": ms,
            "config": {"workload": "BERT-base 12 x 768 forward + CLS pooling + L2 norm, random-init bf16 weights, "
                                   f"{Q} queries padded to {L} tokens (SURVEY 8f item 4)", "queries_per_gpu": q_local,
                       "real_tokens_per_gpu": int(lens.sum()), "sharding": f"query-sharded x{world}, no collective"},
            "linear_layer_tflops": flop_tok * q_local * L / (ms * 1e-3) / 1e12,
            "e2e": {"value": Q / (e2e_ms * 1e-3), "unit": "queries/s", "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": int(lens.sum()) * 4 + (q_local + 1) * 4,
                    "d2h_bytes_per_step": q_local * H * 4,
                    "note": "host token rows -> packed tokens + cu_seqlens -> H2D -> qf_bert_encode_packed (no padding "
                            "rows) -> D2H; tokenisation (host WordPiece) not included"},
            # own kernels per call: embed+LN, pool+norm, and per layer attention, 2 x add+LN, GELU (the 4 Linear layers per
            # block are cuBLASLt launches and not counted); the e2e path is one packed call per step
            "gpu_launches": (2 + 4 * cfg.n_layers) * (steps + 3) + (2 + 4 * cfg.n_layers) * (steps + 1)}



# ------------------------------------------------------------------------------------------------ extra legs (round 2)
def _oracle_check_weighted(ids, micro, lens, w, out, method, ranks=None, n=32, seed=0):
    """In-run parity sample, OUTSIDE every timed region: `n` queries of a device batch against the C restatement of
    the reference's weighted_* loops + write_runfile ranking (oracle/qf_oracle.c): ranked ids and float64 score bits."""
    import numpy as np
    import torch
    from oracle import c_oracle as co
    Q = ids.shape[0]
    rng = np.random.default_rng(seed)
    pick = torch.from_numpy(np.sort(rng.choice(Q, size=min(n, Q), replace=False))).to(ids.device)
    h = lambda t: t.index_select(0, pick).cpu().numpy()
    o_ids, o_sc, _c, _p, o_len = co.fuse_weighted(h(ids), h(micro), h(lens), h(w), method, 60.0,
                                                  None if ranks is None else h(ranks), n_threads=os.cpu_count() or 1)
    g_ids, g_sc, g_len = h(out.ids), h(out.scores), h(out.lens)
    ok = bool(np.array_equal(g_len, o_len))
    if ok:
        m = np.arange(g_ids.shape[1])[None, :] < o_len[:, None]
        ok = bool(np.array_equal(g_ids[m], o_ids[m]) and np.array_equal(g_sc[m].view(np.int64), o_sc[m].view(np.int64)))
    return {"queries_checked": int(len(pick)), "bit_exact_vs_oracle": ok}


def _time_replays(fn, steps, stream, dev, warm=3):
    import torch
    for _ in range(warm):
        fn()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for i in range(steps):
        fn(i) if fn.__code__.co_argcount else fn()
    e1.record(stream)
    torch.cuda.synchronize(dev)
    return e0.elapsed_time(e1) / steps


def wrrf_leg(args, dev, sets):
    """BASELINE configs[1] names W-CombSUM / W-RRF: the same step with weighted_rrf (src/fusion.py:333-380), i.e. the
    FILE's rank column (explicit int32 ranks, one more 69 MB array per input set) and one float64 division per posting."""
    import torch
    from qppfusion import _native as N_, ops
    from qppfusion.pipeline import FusionStepGraph
    Q, R, K = sets[0][0].shape
    stream = torch.cuda.current_stream(dev)
    ranks = [torch.arange(1, K + 1, dtype=torch.int32, device=dev).expand(Q, R, K).contiguous() for _ in sets]
    graphs = [FusionStepGraph(*st, qpp_index=5, top_k=100, k_cut=50, colnorm="minmax", round6=True, method=2, norm=0,
                              rrf_k=60.0, flags=0, want_counts=False, want_pos=False, compact=False, ranks=rk)
              for st, rk in zip(sets, ranks)]
    steps = max(4, min(args.steps, 20))
    ms = _time_replays(lambda i=0: graphs[i % len(graphs)].replay(), steps, stream, dev, warm=len(graphs))
    N_.profile_enable(True)
    for i in range(steps):
        ids, micro, lens, _f = sets[i % len(sets)]
        ops.fuse(ids, micro, lens, graphs[i % len(sets)].weights, ranks[i % len(sets)], 2, 0, 60.0, 0, want_counts=False, want_pos=False)
    torch.cuda.synchronize(dev)
    fuse_ms = N_.profile_read(N_.PROF_FUSE_STREAM)
    N_.profile_enable(False)
    g = graphs[0]
    g.replay()
    torch.cuda.synchronize(dev)
    check = _oracle_check_weighted(sets[0][0], sets[0][1], sets[0][2], g.weights, g.out, 2, ranks[0], seed=1)
    union = int(g.out.lens.sum().item())
    alg = 12 * Q * R * K + Q * R * 12 + 12 * union + 8 * Q           # ids + micro + ranks in
    hbm, _tf, _w = load_peaks()
    return {"value": Q / (ms * 1e-3), "unit": "fused queries/s", "ms_per_step": ms, "fuse_kernel_ms": fuse_ms,
            "hbm_frac": alg / (fuse_ms * 1e-3) / 1e9 / hbm, "algorithmic_bytes": alg, "parity": check,
            "workload": "configs[1] step with W-RRF on the file's rank column (one graph replay per step, one step in flight)"}


def c1_leg(args, dev):
    """BASELINE configs[0]: SciFact-shaped 300 queries x 5 retrievers x top-100, CombSUM / CombMNZ / RRF with ranx
    semantics (min-max per list, sorted-run positions), the three methods as ONE CUDA graph over one batch."""
    import numpy as np
    import torch
    from oracle import fusion_oracle as fo
    from qppfusion import ops
    global N_DOCS, POOL
    keep = (N_DOCS, POOL)
    N_DOCS, POOL = 5183, 230
    try:
        sets = [synth_device(300, 5, 100, 11 + i, dev) for i in range(4)]
    finally:
        N_DOCS, POOL = keep
    stream = torch.cuda.current_stream(dev)
    cfg = (("combsum", 0, 1), ("combmnz", 1, 1), ("rrf", 2, 0))

    def body(st):
        ids, micro, lens, _f = st
        return [ops.fuse(ids, micro, lens, None, None, m, norm, 60.0, 2, want_counts=False, want_pos=False) for _n, m, norm in cfg]

    side = torch.cuda.Stream(dev)
    side.wait_stream(stream)
    with torch.cuda.stream(side):
        for st in sets:
            body(st)
    stream.wait_stream(side)
    torch.cuda.synchronize(dev)
    graphs, outs = [], []
    for st in sets:
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            outs.append(body(st))
        graphs.append(g)
    steps = 200
    ms = _time_replays(lambda i=0: graphs[i % 4].replay(), steps, stream, dev, warm=8)
    # parity sample against the ranx restatement (oracle/fusion_oracle.py), 6 queries x 3 methods
    ids, micro, lens, _f = [t.cpu().numpy() for t in sets[0]]
    ok = True
    for (name, m, norm), out in zip(cfg, outs[0]):
        o_ids, o_sc, o_len = out.ids.cpu().numpy(), out.scores.cpu().numpy(), out.lens.cpu().numpy()
        for q in range(0, 300, 50):
            per_r = [[(int(ids[q, r, i]), micro[q, r, i] / 1e6, i + 1) for i in range(lens[q, r])] for r in range(5)]
            docs, sc, _c = fo.fuse_ranx_query(per_r, m, 60.0, "min-max", False)
            order = fo.rank_order(sc)
            n = len(docs)
            ok &= bool(o_len[q] == n and [int(x) for x in o_ids[q, :n]] == [docs[i] for i in order]
                       and np.array_equal(o_sc[q, :n], np.array([sc[i] for i in order]) + 0.0))
    return {"value": 300 / (ms * 1e-3), "unit": "queries/s (each fused by CombSUM, CombMNZ and RRF)", "ms_per_step": ms,
            "parity": {"queries_checked": 6, "methods": 3, "bit_exact_vs_oracle": ok},
            "workload": "configs[0]: 300 q x 5 x top-100, ranx-semantics CombSUM + CombMNZ + RRF as one CUDA graph of 6 "
                        "launches; 4 rotating input sets (1.2 MB each: L2-resident, launch-bound shape)"}


def c5_leg(args, dev, world, rank):
    """BASELINE configs[4]: throughput sweep, QPP-weighted fusion of 5 x top-1000 lists, query-sharded, no collective.
    200 000 synthetic queries per GPU (8 GB of postings + 12 GB of padded outputs), fused by ONE qf_fuse launch per
    step after qf_qpp_stats / colnorm / select_weight; inputs are 64x the L2."""
    import torch
    import torch.distributed as dist
    from qppfusion import _native as N_, ops
    Qc, chunk = 200_000, 25_000
    parts = [synth_device(chunk, R_NQ, K_NQ, 20250105 + 31 * rank + 1000 * i, dev) for i in range(Qc // chunk)]
    ids, micro, lens, feats = [torch.cat([p_[j] for p_ in parts]) for j in range(4)]
    del parts
    stream = torch.cuda.current_stream(dev)

    def step():
        qpp, _st = ops.qpp_stats(micro, lens, feats, None, None, top_k=100, k_cut=50)
        ops.qpp_colnorm_(qpp, "minmax", round6=True)
        w = ops.qpp_select_weight(qpp, 5)
        return w, ops.fuse(ids, micro, lens, w, None, 0, 0, 60.0, 0, want_counts=False, want_pos=False)

    for _ in range(2):              # two warm-up steps: both ping-pong sets of the 12 GB padded outputs exist before the
        w, out = step()             # timed region (a cudaMalloc of that size inside it costs ~100 ms)
    torch.cuda.synchronize(dev)
    N_.profile_enable(True)
    steps = 3
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        w, out = step()
    e1.record(stream)
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1) / steps
    fuse_ms = N_.profile_read(N_.PROF_FUSE_STREAM)
    N_.profile_enable(False)
    t = torch.tensor([ms, fuse_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, fuse_ms = t.tolist()
    deferred = int((out.lens < 0).sum().item())
    union = int(out.lens.clamp(min=0).sum().item())
    check = _oracle_check_weighted(ids, micro, lens, w, out, 0, None, n=32, seed=2) if rank == 0 else None
    alg = 8 * Qc * R_NQ * K_NQ + Qc * R_NQ * 12 + 12 * union + 8 * Qc
    hbm, _tf, _w = load_peaks()
    return {"value": world * Qc / (ms * 1e-3), "unit": "fused queries/s", "ms_per_step": ms, "fuse_kernel_ms": fuse_ms,
            "queries_per_gpu": Qc, "n_gpus": world, "hbm_frac": alg / (fuse_ms * 1e-3) / 1e9 / hbm, "deferred": deferred,
            "parity": check,
            "workload": "configs[4]: 200 000 queries x 5 x top-1000 per GPU, QPP (13 predictors) + min-max + W-CombSUM(RSD), "
                        "one launch per kernel per step; 1 M queries = 5 such steps per GPU"}


def api_e2e_leg(args, dev):
    """File to file through the reference-facing call: 5 x <Ranker>.norm.res + 5 x .qpp on disk ->
    qppfusion.fusion.run_fusion("wcombsum") -> fused .res on disk (native parse + pack + H2D + qf_fuse + CSR compaction
    + D2H + native writer).  A 512-query sample of the configs[1] shape (2.56 M run-file lines, ~106 MB of text)."""
    import contextlib
    import io
    import shutil
    import tempfile
    import torch
    from qppfusion import fusion
    nq = 512
    root = tempfile.mkdtemp(prefix="qf_api_e2e_")
    try:
        runs_dir, qpp_dir = write_synth_run_files(root, nq, seed=21)
        in_bytes = sum(os.path.getsize(os.path.join(runs_dir, f)) for f in os.listdir(runs_dir))
        out_path = os.path.join(root, "wcombsum_rsd.res")
        times = []
        for _ in range(4):
            torch.cuda.synchronize(dev)
            t0 = time.perf_counter()
            with contextlib.redirect_stdout(io.StringIO()):
                fusion.run_fusion("wcombsum", runs_dir, qpp_dir=qpp_dir, qpp_model="RSD", output_path=out_path, return_fused=False)
            torch.cuda.synchronize(dev)
            times.append(time.perf_counter() - t0)
        best = min(times[1:])
        out_bytes = os.path.getsize(out_path)
        return {"value": nq / best, "unit": "fused queries/s", "seconds": best, "queries": nq, "input_text_bytes": in_bytes,
                "output_text_bytes": out_bytes, "host_threads": os.cpu_count(),
                "workload": "run_fusion('wcombsum', runs_dir, qpp_dir, output_path): .norm.res + .qpp files -> fused .res "
                            "file, wall clock, best of 3 after one warm-up (page cache warm)"}
    finally:
        shutil.rmtree(root, ignore_errors=True)


def pcie_probe(dev, up_bytes, down_bytes):
    """Pinned-memory copy bandwidth with both directions active (the e2e step's pattern): the ceiling of `e2e`."""
    import torch
    h_up = torch.empty(up_bytes, dtype=torch.uint8).pin_memory()
    h_dn = torch.empty(down_bytes, dtype=torch.uint8).pin_memory()
    d_up = torch.empty(up_bytes, dtype=torch.uint8, device=dev)
    d_dn = torch.empty(down_bytes, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 6
    cur = torch.cuda.current_stream(dev)
    for it in range(2):
        e0.record(cur)
        s1.wait_stream(cur)
        s2.wait_stream(cur)
        for _ in range(reps):
            with torch.cuda.stream(s1):
                d_up.copy_(h_up, non_blocking=True)
            with torch.cuda.stream(s2):
                h_dn.copy_(d_dn, non_blocking=True)
        cur.wait_stream(s1)
        cur.wait_stream(s2)
        e1.record(cur)
        torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1) / reps
    return {"ms_per_step_copies_only": ms, "h2d_gbs": up_bytes / ms / 1e6, "d2h_gbs": down_bytes / ms / 1e6}


def recorded_reference():
    """The REAL reference (src/fusion.py weighted_combsum etc.) cannot travel to the GPU box (/root/reference exists
    only in the build container): its timing is recorded there by `python bench.py --impl reference` and committed."""
    p = os.path.join(ROOT, "profiles", "ref_real_fusion_cpu.json")
    if os.path.isfile(p):
        return json.load(open(p))
    return None


def run_ours(args):
    import torch
    import torch.distributed as dist
    from qppfusion import ops

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    Q, R, K = Q_NQ, R_NQ, K_NQ
    NSETS = 4     # rotating input sets: 138 MB each, so no step finds its inputs in the 126 MB L2
    INFLIGHT = int(os.environ.get("BENCH_INFLIGHT", "0")) or 4  # steps in flight (2 -> 4 in round 2: +1.5 % at N = 1) (one stream each; NSETS % INFLIGHT == 0 keeps a graph on one stream);
    # at N > 1 more steps in flight cover the all-reduce of the cross-query min/max (it waits for the slowest rank)
    sets = [synth_device(Q, R, K, 20250101 + 2 + 17 * rank + 1000 * i, dev) for i in range(NSETS)]
    stream = torch.cuda.current_stream(dev)
    from qppfusion.pipeline import FusionStepGraph

    def step(i, ev=None):
        ids, micro, lens, feats = sets[i % NSETS]
        if ev: ev[0].record(stream)
        qpp, _st = ops.qpp_stats(micro, lens, feats, None, None, top_k=100, k_cut=50)
        if ev: ev[1].record(stream)
        ops.qpp_colnorm_(qpp, "minmax", round6=True)
        w = ops.qpp_select_weight(qpp, 5)
        if ev: ev[2].record(stream)
        out = ops.fuse(ids, micro, lens, w, None, 0, 0, 60.0, 0, want_counts=False, want_pos=False)
        if ev: ev[3].record(stream)
        return out

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # The product path replays the step as ONE CUDA graph per input set (qppfusion.pipeline.FusionStepGraph): the
    # kernels are 10-400 us long and issuing seven of them from Python per step leaves the GPU waiting for the host.
    # N > 1: the ranks' shards form ONE logical run file of world x 3 452 queries, so the cross-query min / max of
    # _normalize_qpp (src/qpp.py:340-374) is all-reduced over the ranks inside every step (ShardedFusionStep).
    def make_graph(i, compact):
        if world > 1:
            from qppfusion.pipeline import ShardedFusionStep
            return ShardedFusionStep(*sets[i], qpp_index=5, top_k=100, k_cut=50, round6=True, method=0, rrf_k=60.0,
                                     flags=0, want_counts=False, want_pos=False, compact=compact)
        return FusionStepGraph(*sets[i], qpp_index=5, top_k=100, k_cut=50, colnorm="minmax", round6=True, method=0,
                               norm=0, rrf_k=60.0, flags=0, want_counts=False, want_pos=False, compact=compact)

    graphs = [make_graph(i, False) for i in range(NSETS)]
    run_streams = [torch.cuda.Stream(dev) for _ in range(INFLIGHT)]

    def run_steps(n):
        """n consecutive steps, INFLIGHT of them in flight: queries of different steps are independent, so step i+1's
        short QPP / normalisation kernels and the tail of step i's fusion kernel share the SMs."""
        for s_ in run_streams:
            s_.wait_stream(stream)
        out_ = None
        for i in range(n):
            with torch.cuda.stream(run_streams[i % INFLIGHT]):
                out_ = graphs[i % NSETS].replay()
        for s_ in run_streams:
            stream.wait_stream(s_)
        return out_

    out = run_steps(max(args.warmup, NSETS))
    barrier()
    union_total = int(graphs[0].out.lens.sum().item())
    assert all(int(g.out.status.max().item()) == 0 for g in graphs)
    sharded_check = None
    if world == 1:
        eager = step(0)
        barrier()
        assert torch.equal(eager.lens, graphs[0].out.lens) and torch.equal(eager.scores[0, :16], graphs[0].out.scores[0, :16])
    else:
        # sharded == single GPU, bit for bit: every rank's QPP weights of input set 0 are gathered on all ranks and
        # rank 0 recomputes the WHOLE world x Q-query file alone (it regenerates the other ranks' seeded shards)
        graphs[0].replay()
        mine = graphs[0].weights.contiguous()
        allw = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allw, mine)
        if rank == 0:
            ok = True
            qpp_all = []
            for r_ in range(world):
                ids_r, micro_r, lens_r, feats_r = sets[0] if r_ == 0 else synth_device(Q, R, K, 20250101 + 2 + 17 * r_, dev)
                qpp_all.append(ops.qpp_stats(micro_r, lens_r, feats_r, None, None, top_k=100, k_cut=50)[0])
                del ids_r, micro_r, lens_r, feats_r
            qpp_all = torch.cat(qpp_all)
            ops.qpp_colnorm_(qpp_all, "minmax", round6=True)
            w_all = ops.qpp_select_weight(qpp_all, 5)
            for r_ in range(world):
                ok &= bool(torch.equal(w_all[r_ * Q:(r_ + 1) * Q].view(torch.int64), allw[r_].view(torch.int64)))
            sharded_check = {"queries": world * Q, "weights_bit_equal_to_single_gpu": ok}
            del qpp_all, w_all
    barrier()

    # ---- timed region: K steps (graph replays), device-timed, max over ranks
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    t0.record(stream)
    run_steps(args.steps)
    t1.record(stream)
    barrier()
    ms_total = t0.elapsed_time(t1)
    # (the clock sampler keeps running through the per-kernel loop below: same kernels, same load, and the K-step timed
    # region alone lasts only ~10 ms)

    # ---- per-kernel durations: the same K steps issued kernel by kernel with CUDA events between them (events cannot
    # be read back from inside a graph replay); the fusion kernel's average is the roofline denominator
    from qppfusion import _native as N_
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(args.steps)]
    for i in range(3):
        step(i)
    barrier()
    N_.profile_enable(True)           # in-library event pairs right around the fusion / QPP kernel launches
    for i in range(args.steps):
        step(i, evs[i])
    barrier()
    fuse_ms = N_.profile_read(N_.PROF_FUSE_STREAM)            # streaming fusion kernel alone (average of K launches)
    fixup_ms = N_.profile_read(N_.PROF_FUSE_GENERAL)          # general kernel's fix-up pass over deferred queries
    qpp_ms = N_.profile_read(N_.PROF_QPP_STATS)
    N_.profile_enable(False)
    clocks = sampler.stop() if rank == 0 else None
    if clocks is not None:
        clocks["window"] = "timed region + the per-kernel loop that follows it"
    fuse_op_ms = sum(e[2].elapsed_time(e[3]) for e in evs) / args.steps
    norm_ms = sum(e[1].elapsed_time(e[2]) for e in evs) / args.steps
    assert int((graphs[0].out.lens < 0).sum().item()) == 0       # nothing left deferred
    main_check = _oracle_check_weighted(sets[0][0], sets[0][1], sets[0][2], graphs[0].weights, graphs[0].out, 0, None,
                                        n=32, seed=0) if rank == 0 else None
    wrrf = wrrf_leg(args, dev, sets) if (rank == 0 and world == 1) else None
    torch.cuda.empty_cache()

    # ---- end to end: pinned host buffers -> H2D -> graph replay (wire expansion + step + CSR compaction) -> D2H of the
    # fused run.  Three streams (H2D / compute / D2H) so step i+1's upload and step i-1's download overlap step i's
    # kernels; every step's copies are inside the timed region.  The postings travel bit-packed (csrc/qf_wire.cu:
    # id_bits + sc_bits per posting instead of 2 x int32) and the step's first kernel expands them; the plain int32
    # upload is measured next to it (`int32_upload`).
    ref_lens0 = graphs[0].out.lens.clone()
    ref_row0 = graphs[0].out.ids[0, :64].clone()
    del graphs
    torch.cuda.empty_cache()
    NSETS = 3
    n_post = Q * R * K
    NB = 2
    h_off = [torch.empty((Q + 1,), dtype=torch.int64).pin_memory() for _ in range(NB)]
    h_ids = [torch.empty((n_post,), dtype=torch.int32).pin_memory() for _ in range(NB)]
    h_sc = [torch.empty((n_post,), dtype=torch.float64).pin_memory() for _ in range(NB)]
    s_h2d, s_cmp, s_d2h, s_off = (torch.cuda.Stream(dev) for _ in range(4))
    e2e_steps = max(4, min(args.steps, 12))

    def e2e_measure(use_wire):
        if use_wire:
            h_wire = [ops.wire_pack(src_[0], src_[1], src_[2], pin=True, coding=use_wire) for src_ in cpu_sets]
            d_wire = [w_.to(dev) for w_ in h_wire]
            # (coding "for": the payload length depends on the data -- only the used prefix of the buffer travels)
            pairs = [[(d_wire[i].words[:h_wire[i].used_words], h_wire[i].words[:h_wire[i].used_words]),
                      (sets[i][2], sets[i][2].cpu().pin_memory()),
                      (sets[i][3], sets[i][3].cpu().pin_memory())] for i in range(NSETS)]
            for i in range(NSETS):                       # the graph must rebuild these from the wire on every replay
                sets[i][0].zero_()
                sets[i][1].zero_()
        else:
            h_wire = d_wire = [None] * NSETS
            pairs = [[(x, x.cpu().pin_memory()) for x in sets[i]] for i in range(NSETS)]
        graphs_ = []
        for i in range(NSETS):
            if world > 1:
                from qppfusion.pipeline import ShardedFusionStep
                graphs_.append(ShardedFusionStep(*sets[i], qpp_index=5, top_k=100, k_cut=50, round6=True, method=0,
                                                 rrf_k=60.0, flags=0, want_counts=False, want_pos=False, compact=True,
                                                 wire=d_wire[i]))
            else:
                graphs_.append(FusionStepGraph(*sets[i], qpp_index=5, top_k=100, k_cut=50, colnorm="minmax", round6=True,
                                               method=0, norm=0, rrf_k=60.0, flags=0, want_counts=False, want_pos=False,
                                               compact=True, wire=d_wire[i]))
        d2h_bytes = [0]

        def e2e_run(n_steps):
            ev_h2d = [torch.cuda.Event() for _ in range(n_steps)]
            ev_cmp = [torch.cuda.Event() for _ in range(n_steps)]
            ev_d2h = [torch.cuda.Event() for _ in range(n_steps)]
            ev_off = torch.cuda.Event()

            def upload(i):
                with torch.cuda.stream(s_h2d):
                    if i >= NSETS:
                        s_h2d.wait_event(ev_cmp[i - NSETS])     # that set's previous step has consumed its inputs
                    for dst, src in pairs[i % NSETS]:
                        dst.copy_(src, non_blocking=True)
                    ev_h2d[i].record(s_h2d)

            def compute(i):
                with torch.cuda.stream(s_cmp):
                    s_cmp.wait_event(ev_h2d[i])
                    if i >= NSETS:
                        s_cmp.wait_event(ev_d2h[i - NSETS])     # that graph's previous outputs have been downloaded
                    graphs_[i % NSETS].replay()
                    ev_cmp[i].record(s_cmp)

            # software pipeline: uploads run two steps ahead and the next step's kernels are queued BEFORE the host
            # waits for this step's CSR total (it needs the number of fused docs to size the download), so neither the
            # copy engine nor the SMs idle while the host blocks
            upload(0)
            if n_steps > 1:
                upload(1)
            compute(0)
            for i in range(n_steps):
                if i + 2 < n_steps:
                    upload(i + 2)
                if i + 1 < n_steps:
                    compute(i + 1)
                b = i % NB
                off, cids, csc = graphs_[i % NSETS].csr
                with torch.cuda.stream(s_off):                   # own stream: not queued behind step i-1's big download
                    s_off.wait_event(ev_cmp[i])
                    h_off[b].copy_(off, non_blocking=True)
                    ev_off.record(s_off)
                ev_off.synchronize()                             # host needs the total before sizing the copy
                total = int(h_off[b][-1])
                with torch.cuda.stream(s_d2h):
                    s_d2h.wait_event(ev_cmp[i])
                    h_ids[b][:total].copy_(cids[:total], non_blocking=True)
                    h_sc[b][:total].copy_(csc[:total], non_blocking=True)
                    ev_d2h[i].record(s_d2h)
                d2h_bytes[0] = total * 12 + (Q + 1) * 8

        e2e_run(3)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for s_ in (s_h2d, s_cmp, s_d2h, s_off):
            s_.wait_stream(stream)
        e2e_run(e2e_steps)
        for s_ in (s_h2d, s_cmp, s_d2h, s_off):
            stream.wait_stream(s_)
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1) / e2e_steps
        up = sum(src.numel() * src.element_size() for _d, src in pairs[0])
        # the downloaded run of the last step equals the device-resident result of the same input set (bit for bit)
        last = (e2e_steps - 1) % NSETS
        tot = int(h_off[(e2e_steps - 1) % NB][-1])
        same = bool(tot == int(graphs_[last].out.lens.sum().item()) and
                    torch.equal(h_ids[(e2e_steps - 1) % NB][:tot], graphs_[last].csr[1][:tot].cpu()) and
                    torch.equal(graphs_[0].out.lens, ref_lens0) and torch.equal(graphs_[0].out.ids[0, :64], ref_row0))
        wire_info = None
        if use_wire:
            wire_info = {"coding": use_wire, "id_bits": h_wire[0].id_bits,
                         "sc_bits": h_wire[0].sc_bits if use_wire == "fixed" else "frame of reference per block of 32",
                         "bytes_per_posting": h_wire[0].used_words * 8 / n_post}
        del graphs_
        return ms, up, d2h_bytes[0], same, wire_info

    cpu_sets = [[x.cpu() for x in st_[:3]] for st_ in sets[:NSETS]]      # (the wire variants zero the device copies)
    ms_plain, h2d_plain, d2h, same_plain, _ = e2e_measure(None)
    union_plain = d2h
    ms_fixed, h2d_fixed, d2h_f, same_fixed, wire_fixed = e2e_measure("fixed")
    e2e_ms, h2d, d2h, same_wire, wire_info = e2e_measure("for")
    assert same_plain and same_fixed and same_wire and d2h == union_plain and d2h_f == union_plain, \
        "e2e: downloaded run differs from the device-resident result"
    del cpu_sets
    h_out = (h_off, h_ids, h_sc)
    barrier()
    pcie = pcie_probe(dev, h2d, max(d2h, 1 << 20))       # the same bytes per step, copies only: the ceiling of e2e

    del sets, h_out
    torch.cuda.empty_cache()
    dense = None if args.no_dense else dense_leg(args, dev, world, rank)
    torch.cuda.empty_cache()
    # HotpotQA-shaped dense search (BASELINE configs[3]: the corpus-sharded multi-GPU config), fewer steps
    dense_c4 = None if args.no_dense else dense_leg(args, dev, world, rank, Q=7405, n_docs=5_233_329,
                                                    label="BASELINE configs[3], HotpotQA shape", max_steps=4)

    torch.cuda.empty_cache()
    encoder = None if args.no_dense else encoder_leg(args, dev, world, rank)
    torch.cuda.empty_cache()
    c1 = c1_leg(args, dev) if (rank == 0 and not args.no_extra) else None
    torch.cuda.empty_cache()
    c5 = None if args.no_extra else c5_leg(args, dev, world, rank)
    torch.cuda.empty_cache()
    api = api_e2e_leg(args, dev) if (rank == 0 and world == 1 and not args.no_extra) else None

    times = torch.tensor([ms_total, e2e_ms, ms_plain, ms_fixed], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    ms_total, e2e_ms, ms_plain, ms_fixed = times.tolist()

    if rank == 0:
        hbm_peak, tf_peak, which = load_peaks()
        n_post = Q * R * K
        alg_bytes = 8 * n_post + Q * R * 12 + 12 * union_total + 8 * Q   # ids+micro in, lens+weights, ids+f64 out
        achieved = alg_bytes / (fuse_ms * 1e-3) / 1e9
        # second roofline: instruction issue.  The fusion kernel is bound by the warp schedulers, not by HBM or the
        # shared-memory pipe: warp-instructions per query from the committed ncu capture x
        # queries per launch / measured kernel time, against 4 schedulers x 148 SMs x the SM clock seen under load.
        ft = {}
        try:
            ft = json.load(open(os.path.join(ROOT, "profiles", "fuse_traffic.json")))
        except Exception:
            pass
        issue = None
        if ft.get("warp_inst_per_query") and clocks and clocks.get("sm_mhz"):
            peak_i = 4 * 148 * clocks["sm_mhz"] * 1e6
            ach_i = ft["warp_inst_per_query"] * Q / (fuse_ms * 1e-3)
            issue = {"kernel": "fuse_stream_kernel", "bound": "issue", "achieved": ach_i / 1e9, "peak": peak_i / 1e9,
                     "unit": "G warp-instructions/s", "frac": ach_i / peak_i,
                     "warp_inst_per_query": ft["warp_inst_per_query"],
                     "smem_wavefronts_per_query": ft.get("smem_wavefronts_per_query"),
                     "lsu_frac": (ft.get("smem_wavefronts_per_query") or 0) * Q / (fuse_ms * 1e-3) / (148 * clocks["sm_mhz"] * 1e6),
                     "source": "profiles/fuse_traffic.json (ncu --set full of the same kernel) x in-run kernel time"}
        line = {
            "metric": METRIC, "value": world * Q * args.steps / (ms_total * 1e-3), "unit": "fused queries/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "NQ-shaped 3452 queries x 5 retrievers x top-1000: 13 QPP predictors + "
                                   "cross-query min-max + W-CombSUM(RSD) per step (BASELINE configs[1]); each step is "
                                   "one CUDA-graph replay of the 7 kernels",
                       "queries_per_gpu": Q, "retrievers": R, "depth": K, "score_format": "int32 micro-units",
                       "union_ratio": union_total / n_post,
                       "sharding": (f"query-sharded x{world}; one all-reduce(MAX) of 2x5x13 doubles per step for the "
                                    "cross-query min/max of _normalize_qpp (ShardedFusionStep)") if world > 1 else "1 GPU",
                       "l2": "4 rotating input sets of 138 MB (> 126 MB L2)",
                       "steps_in_flight": INFLIGHT},
            "kernel_ms": {"qpp_stats": qpp_ms, "colnorm+select": norm_ms, "fuse_stream": fuse_ms,
                          "fuse_fixup": fixup_ms, "fuse_op_eager": fuse_op_ms,
                          "note": "per-kernel times from a second K-step loop issued kernel by kernel (in-library "
                                  "CUDA events around the launches); the timed region itself replays CUDA graphs"},
            "roofline": {"kernel": "fuse_stream_kernel<768,2816,4096,5120,256,3,micro,plain>", "bound": "hbm",
                         "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                         "peak_source": which, "algorithmic_bytes": alg_bytes, "traffic": fuse_traffic(),
                         "frac_of_nominal_8TBs": achieved / 8000.0},
            "e2e": {"value": world * Q / (e2e_ms * 1e-3), "unit": "fused queries/s", "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "pcie_ceiling": dict(pcie, queries_per_s=world * Q / (pcie["ms_per_step_copies_only"] * 1e-3),
                                         note="the same H2D + D2H bytes per step, copies only, both directions at once "
                                              "(this rank; measured while the other ranks do the same at N > 1)"),
                    "wire": wire_info,
                    "int32_upload": {"value": world * Q / (ms_plain * 1e-3), "ms_per_step": ms_plain,
                                     "h2d_bytes_per_step": h2d_plain,
                                     "note": "the same step fed with the plain int32 [Q,R,K] id / score arrays"},
                    "fixed_width_upload": {"value": world * Q / (ms_fixed * 1e-3), "ms_per_step": ms_fixed,
                                           "h2d_bytes_per_step": h2d_fixed, "wire": wire_fixed},
                    "note": "pinned host buffers (postings bit-packed by qf_wire_pack / qf_wire_for_pack: id_bits per doc id, "
                            "scores frame-of-reference coded per block of 32) -> "
                            "H2D -> graph replay (qf_wire_unpack + qf_wire_for_unpack + step + CSR compaction) -> D2H of the fused run; "
                            "consecutive steps overlap on 3 streams; the downloaded run is checked against the "
                            "device-resident result"},
            "roofline_issue": issue,
            "parity": main_check,
            "sharded_colnorm": sharded_check,
            "wrrf": wrrf, "c1": c1, "c5": c5, "api_e2e": api,
            "cpu_reference_recorded": recorded_reference(),
            # per step: qpp 1 + colnorm 3 + select 1 + fuse 2 (streaming + fix-up) = 7; CSR compaction adds 4, the wire
            # expansion 1 (e2e is measured twice: int32 upload and bit-packed upload)
            "gpu_launches": (7 * (args.steps + max(args.warmup, 4)) + 7 * (args.steps + 4) + (11 + 12 + 13) * (e2e_steps + 3 + 6)
                             + (dense["gpu_launches"] if dense else 0) + (dense_c4["gpu_launches"] if dense_c4 else 0)
                             + (encoder["gpu_launches"] if encoder else 0)),
            "clocks": clocks,
            "dense": dense,
            "dense_hotpotqa": dense_c4,
            "encoder": encoder,
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(seconds=10.0)
        # the contract keys first, the large nested legs last: a reader that truncates the line keeps what it grades on
        first = ["metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                 "vs_baseline", "dtype", "data", "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks",
                 "kernel_ms", "roofline_issue", "parity", "sharded_colnorm"]
        line = {**{k_: line[k_] for k_ in first if k_ in line}, **{k_: v_ for k_, v_ in line.items() if k_ not in first}}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------ CPU arm
def _cpu_sample_arrays(nq, seed):
    """Host-side sample of the same synthetic NQ shape in the packed layout ([nq,5,1000] int32 micro-units)."""
    import numpy as np
    rng = np.random.default_rng(seed)
    stride = N_DOCS // POOL
    scale = [1.0, 30.0, 1.0, 1.0, 12.0]
    ids = np.empty((nq, R_NQ, K_NQ), dtype=np.int32)
    micro = np.empty((nq, R_NQ, K_NQ), dtype=np.int32)
    for q in range(nq):
        pool = np.arange(POOL) * stride + rng.integers(0, stride, POOL)
        for r in range(R_NQ):
            ids[q, r] = rng.choice(pool, size=K_NQ, replace=False)
            u1, u2 = rng.random(K_NQ), rng.random(K_NQ)
            s = (1.0 - np.sqrt(u1)) * (0.35 + 0.65 * u2) * scale[r]
            if r == 2:
                s = -3.0 * (-np.log1p(-u2))
            micro[q, r] = np.round(np.sort(s)[::-1] * 1e6).astype(np.int32)
    lens = np.full((nq, R_NQ), K_NQ, dtype=np.int32)
    n_tok = rng.poisson(9, nq).astype(np.int32) + 1
    feats = np.stack([n_tok, np.maximum(n_tok - 1, 1), rng.integers(4, 14, nq).astype(np.int32),
                      (n_tok * 4.5).astype(np.int32)], axis=1).astype(np.int32)
    return ids, micro, lens, feats


def _cpu_step_c(ids, micro, lens, feats, threads):
    """The reference's step restated in C (oracle/qf_oracle.c): QPPBridge columns -> cross-query min-max -> %.6f ->
    W-CombSUM(RSD) -> stable ranking.  Returns the number of fused documents."""
    import numpy as np
    from oracle import c_oracle as co
    qpp = co.qpp_bridge_batch(micro, lens, feats, top_k=100, k_cut=50, n_threads=threads)      # [Q,R,13]
    col = qpp[:, :, 5]
    mn, mx = col.min(axis=0), col.max(axis=0)
    w = (col - mn) / np.where(mx > mn, mx - mn, 1.0)
    w = np.round(w * 1e6) / 1e6                                   # "%.6f" round trip (vectorised; sample-size only)
    out = co.fuse_weighted(ids, micro, lens, w, 0, 60.0, n_threads=threads)
    return int(out[4].sum())


def cpu_baseline(seconds=10.0):
    threads = os.cpu_count() or 1
    ids, micro, lens, feats = _cpu_sample_arrays(64, 7)
    _cpu_step_c(ids, micro, lens, feats, threads)                 # warm-up (+ builds the oracle .so if needed)
    t = time.perf_counter()
    _cpu_step_c(ids, micro, lens, feats, threads)
    per_q = (time.perf_counter() - t) / 64
    nq = int(max(64, min(4096, seconds / max(per_q, 1e-6))))
    ids, micro, lens, feats = _cpu_sample_arrays(nq, 8)
    t = time.perf_counter()
    _cpu_step_c(ids, micro, lens, feats, threads)
    dt = time.perf_counter() - t
    t = time.perf_counter()
    _cpu_step_c(ids[:256], micro[:256], lens[:256], feats[:256], 1)
    dt1 = time.perf_counter() - t
    return {"value": nq / dt, "unit": "fused queries/s", "cores": threads, "kind": "port",
            "sample": f"{nq} queries of the same NQ shape (5 x top-1000); C/OpenMP restatement (oracle/qf_oracle.c) of "
                      f"QPPBridge.java + src/fusion.py weighted_combsum + write_runfile ranking on {threads} threads; "
                      f"1 thread: {min(256, nq) / dt1:.0f} q/s. The reference itself is single-threaded Python "
                      f"(pandas iterrows, ~5 q/s measured at survey time)"}


def run_reference_real(args):
    """The REAL reference: /root/reference/src/fusion.py weighted_combsum / weighted_rrf (single-threaded pandas
    loops, src/fusion.py:237-380) + write_runfile on a bounded subsample of the configs[1] shape, and its file-to-file
    run_fusion on the same files.  Build container only (no GPU needed); the result is committed under profiles/ and
    quoted by the GPU arm as `cpu_reference_recorded`."""
    import shutil
    import tempfile
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import ref_import
    import contextlib
    import io
    rf = ref_import.import_reference_fusion()
    nq = 40
    root = tempfile.mkdtemp(prefix="qf_ref_real_")
    try:
        runs_dir, qpp_dir = write_synth_run_files(root, nq, seed=21)
        out = {}
        with contextlib.redirect_stdout(io.StringIO()):
            t0 = time.perf_counter()
            runs = rf.load_runs(runs_dir, use_normalized=True)
            qpp = rf.load_qpp_scores(qpp_dir)
            t_load = time.perf_counter() - t0
            for name, fn in (("weighted_combsum", rf.weighted_combsum), ("weighted_rrf", rf.weighted_rrf)):
                t0 = time.perf_counter()
                fused = fn(runs, qpp, 5)
                dt = time.perf_counter() - t0
                t0 = time.perf_counter()
                rf.write_runfile(fused, os.path.join(root, name + ".res"), name)
                out[name] = {"queries_per_s": nq / dt, "seconds": dt, "write_runfile_seconds": time.perf_counter() - t0}
            t0 = time.perf_counter()
            rf.run_fusion("wcombsum", runs_dir, qpp_dir=qpp_dir, qpp_model="RSD", output_path=os.path.join(root, "f.res"))
            t_api = time.perf_counter() - t0
        rec = {"kind": "reference", "what": "REAL /root/reference/src/fusion.py (ranx stubbed, src/__init__ bypassed: "
                                            "tools/ref_import.py), single-threaded by construction",
               "sample": f"{nq} queries x 5 retrievers x top-1000 (configs[1] shape), synthetic .norm.res / .qpp files",
               "cores": 1, "host_cpus": os.cpu_count(), "load_runs_and_qpp_seconds": t_load, "functions": out,
               "run_fusion_wcombsum_file_to_file": {"queries_per_s": nq / t_api, "seconds": t_api},
               "where": "build container (no GPU); /root/reference does not exist on the GPU box"}
        os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
        json.dump(rec, open(os.path.join(ROOT, "profiles", "ref_real_fusion_cpu.json"), "w"), indent=1)
        print(json.dumps(dict(rec, impl="reference", metric=METRIC, value=out["weighted_combsum"]["queries_per_s"],
                              unit="fused queries/s")), flush=True)
    finally:
        shutil.rmtree(root, ignore_errors=True)


def run_reference(args):
    """Reference arm: the CPU implementation of the step (oracle port, all host threads) on a bounded sample."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if args.real:
        return run_reference_real(args)
    threads = os.cpu_count() or 1
    nq_step = 1024
    ids, micro, lens, feats = _cpu_sample_arrays(nq_step, 9)
    for _ in range(max(1, args.warmup)):
        _cpu_step_c(ids[:128], micro[:128], lens[:128], feats[:128], threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        _cpu_step_c(ids, micro, lens, feats, threads)
    dt = time.perf_counter() - t0
    value = nq_step * args.steps / dt
    world = int(os.environ.get("WORLD_SIZE", "1"))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "fused queries/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "NQ-shaped 5 retrievers x top-1000: 13 QPP predictors + cross-query min-max + "
                                   "W-CombSUM(RSD) (BASELINE configs[1]); bounded sample per step",
                       "queries_per_step": nq_step},
            "cpu_baseline": {"value": value, "unit": "fused queries/s", "cores": threads, "kind": "port",
                             "sample": f"{nq_step} queries/step; C/OpenMP restatement (oracle/qf_oracle.c) of "
                                       "QPPBridge.java + src/fusion.py weighted_combsum + write_runfile ranking "
                                       "(the reference's engines -- JVM, ranx, faiss -- are not installable offline)"},
            "e2e": {"value": value, "unit": "fused queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    # stdout carries exactly ONE JSON line: everything else written to fd 1 while the benchmark runs (NCCL prints its
    # version banner there from C) is sent to stderr; the line itself goes to the saved descriptor.
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real_stdout, "w")
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-dense", action="store_true", help="skip the BGE dense-retrieval leg")
    ap.add_argument("--no-extra", action="store_true", help="skip the configs[0] / configs[4] / file-to-file legs")
    ap.add_argument("--real", action="store_true", help="--impl reference: time the REAL /root/reference/src/fusion.py "
                    "(build container only) and write profiles/ref_real_fusion_cpu.json")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
