/*
 * qppfusion.h -- C ABI of libqppfusion.so: the B200 (sm_100a) implementation of QPP-Fusion-RAG's
 * retrieval-scoring -> QPP -> rank-fusion hot path.
 *
 * The reference (PoisedDok/QPP-Fusion-RAG) has no FFI / plugin registry for this path: its boundary is
 * a set of Python functions (src/fusion.py, src/qpp.py, src/retrievers/bge.py) over three foreign
 * engines (ranx/numba, a JVM subprocess, FAISS).  Each entry point below replaces the arithmetic
 * inside one of those call sites; the reference-side binding is a ctypes stub.
 *
 * Conventions
 *   - every function returns 0 (QF_OK) or a negative QF_E_* code; qf_last_error() gives the message
 *     for the calling thread.  No C++ exception crosses this boundary.
 *   - all buffers are CALLER-OWNED DEVICE pointers (PyTorch allocates them); functions never allocate
 *     device memory, never synchronise the device, and enqueue all work on `stream`
 *     (a cudaStream_t passed as void*).
 *   - there is no CPU fallback: without an sm_100 device every compute entry point fails with
 *     QF_E_NO_DEVICE.
 *   - row-major, innermost dimension last.  Q = queries, R = retrievers (runs), K = list capacity.
 */
#ifndef QPPFUSION_H_
#define QPPFUSION_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QF_ABI_VERSION 1

#if defined(__GNUC__)
#define QF_API __attribute__((visibility("default")))
#else
#define QF_API
#endif

/* status codes */
#define QF_OK               0
#define QF_E_INVALID       -1   /* bad argument (null pointer, negative size, unknown enum) */
#define QF_E_UNSUPPORTED   -2   /* valid request outside the implemented envelope (e.g. R*K too large) */
#define QF_E_CUDA          -3   /* a CUDA runtime call failed; message has cudaGetErrorString */
#define QF_E_NO_DEVICE     -4   /* no CUDA device, or device is not compute capability 10.x */
#define QF_E_WORKSPACE     -5   /* workspace too small; call the matching *_workspace_bytes first */

QF_API const char* qf_last_error(void);
QF_API int         qf_abi_version(void);
/* 0 if `device` is a compute-capability-10.x GPU this library can run on. */
QF_API int         qf_device_check(int device);

/*
 * Optional device-side timing of individual kernels (used by bench.py for the per-kernel roofline): while enabled,
 * CUDA events are recorded on the caller's stream around the named phase of every call (up to 512 calls);
 * qf_profile_read synchronises on the last closing event and returns the AVERAGE elapsed milliseconds of the calls
 * recorded since the last qf_profile_enable(1).  Not for use during CUDA-graph capture.
 */
#define QF_PROF_DENSE_SAMPLE 0   /* strided-sample scoring pass of qf_dense_topk            */
#define QF_PROF_DENSE_TAU    1   /* per-query threshold selection                           */
#define QF_PROF_DENSE_MAIN   2   /* tcgen05 GEMM + fused threshold filter over the corpus   */
#define QF_PROF_DENSE_SELECT 3   /* per-query candidate ranking                             */
#define QF_PROF_FUSE_STREAM  4   /* qf_fuse: streaming fusion kernel                        */
#define QF_PROF_FUSE_GENERAL 5   /* qf_fuse: general kernel (whole launch, or the fix-up pass after the streaming one) */
#define QF_PROF_QPP_STATS    6   /* qf_qpp_stats kernel                                     */
#define QF_PROF_PHASES       7
QF_API int qf_profile_enable(int on);
QF_API int qf_profile_read(int phase, float* ms);

/* ------------------------------------------------------------------------------------------------
 * Rank fusion  --  replaces ranx.fuse (src/fusion.py:182,201,226) and the weighted_* Python loops
 *                  (src/fusion.py:266-278, 309-327, 361-377)
 * ---------------------------------------------------------------------------------------------- */
#define QF_METHOD_SUM 0    /* CombSUM  / W-CombSUM : sum_r w_r * s_r(d)                               */
#define QF_METHOD_MNZ 1    /* CombMNZ  / W-CombMNZ : count(d) * sum_r w_r * s_r(d)                    */
#define QF_METHOD_RRF 2    /* RRF      / W-RRF     : sum_r w_r / (rrf_k + rank_r(d))                  */

#define QF_NORM_NONE   0   /* weighted_* path: scores used as given (src/fusion.py:275)               */
#define QF_NORM_MINMAX 1   /* ranx default norm="min-max": (s-min)/max(max-min,1e-9) per (q, r) list  */
#define QF_NORM_ZSCORE 2   /* ranx norm="zscore": (s-mean)/max(std,1e-9) per (q, r) list; mean and the   */
                           /* population variance are SEQUENTIAL float64 sums in list order (what numba's */
                           /* np.mean / np.std compile to): mean = (((0+s0)+s1)+...)/n,                   */
                           /* var = (((0+d0*d0)+d1*d1)+...)/n with d = s-mean, std = sqrt(var)            */

#define QF_SCORE_F64   0   /* double scores[Q][R][K]                                                  */
#define QF_SCORE_F32   1   /* float  scores[Q][R][K]  (widened exactly to double)                     */
#define QF_SCORE_MICRO 2   /* int32  scores[Q][R][K] = round(score * 1e6): the ".6f" text value;      */
                           /* the kernel reconstructs exactly the double float("x.xxxxxx") parses to */

/* flags */
#define QF_FLAG_MNZ_SKIP_ZERO     1u  /* MNZ count ignores postings whose (normalised) score == 0.0   */
#define QF_FLAG_REQUIRE_SORTED    2u  /* ranx semantics: each list must be sorted by score descending */
                                      /* and duplicate-free (the packer guarantees it); violations    */
                                      /* are reported in out_status, never silently repaired          */
#define QF_FLAG_UNRANKED_OUTPUT   4u  /* skip the final ranking: outputs stay in first-seen order     */
#define QF_FLAG_LARGE_UNION       8u  /* performance hint, results unchanged: most queries merge to   */
                                      /* more than ~45 % of R*K distinct docs (little overlap between  */
                                      /* the runs) -> use the larger shared-memory table configuration */

/* per-query status bits written to out_status[Q] */
#define QF_QSTAT_DUPLICATES   1   /* a doc id occurs twice inside one list (handled exactly; info)    */
#define QF_QSTAT_UNSORTED     2   /* QF_FLAG_REQUIRE_SORTED violated, or RRF ranks not ascending      */
#define QF_QSTAT_BAD_LEN      4   /* lens[q][r] < 0 or > K; the query produced no output              */

/*
 * One launch fuses Q queries.  For query q, list r holds lens[q][r] postings (doc id, score[, rank])
 * in visiting order; rank defaults to position+1 when `ranks` is NULL (RRF with explicit ranks requires
 * them ascending within each list, as pandas sort_values("rank") would leave them, src/fusion.py:371).
 * weights == NULL means unweighted (w = 1).
 *
 * Semantics (bit-exact with the reference's float64 arithmetic):
 *   visiting order  = r ascending, then row order; union kept in FIRST-SEEN order;
 *   accumulation    = acc <- fl(acc + fl(w * s))      (SUM/MNZ, no FMA contraction)
 *                     acc <- fl(acc + fl(w / (rrf_k + rank)))  (RRF)
 *   MNZ             = fl(acc * count), count = postings of the doc (all rows, duplicates included)
 *   ranking         = score descending, ties by first-seen order (Python's stable sort,
 *                     src/fusion.py:488); -0.0 == +0.0.
 *
 * Outputs, per query q (row stride = R*K):
 *   out_ids[q][j], out_scores[q][j] : j-th ranked doc and its fused float64 score, j < out_lens[q]
 *   out_counts[q][j] (nullable)     : number of postings merged into that doc
 *   out_pos[q][j]    (nullable)     : the doc's index in first-seen order (to rebuild the reference's
 *                                     unsorted return value)
 *   out_status[q]    (nullable)     : QF_QSTAT_* bits
 *
 * Envelope: R*K <= qf_fuse_max_postings() = 6144 postings per query (shared-memory resident; 5120 for
 * QF_SCORE_F64), R <= 32.  Up to 5120 the streaming kernel runs; 5121..6144 (e.g. 6 runs x depth 1000) take the
 * general kernel alone (about 3x slower per query).  Deeper requests: qf_fuse_deep below.
 */
QF_API int qf_fuse_max_postings(void);
QF_API int qf_fuse(const int32_t* doc_ids, const void* scores, int score_type, const int32_t* ranks,
            const int32_t* lens, const double* weights,
            int Q, int R, int K, int method, int norm, double rrf_k, uint32_t flags,
            int32_t* out_ids, double* out_scores, uint16_t* out_counts, uint16_t* out_pos,
            int32_t* out_lens, int32_t* out_status, void* stream);

/*
 * The same fusion for run lists deeper than the shared-memory envelope (the reference's dict loops and ranx.fuse take
 * any depth: src/fusion.py:266-278, 182): R*K <= qf_fuse_deep_max_postings() = 65535 postings per query (the 16-bit
 * out_counts / out_pos bound), R <= 32, per-query state in `ws` (qf_fuse_deep_workspace_bytes(Q, R, K) bytes of
 * 256-byte aligned device scratch; a smaller workspace that holds at least one query just runs fewer CTAs).
 * Arguments, semantics and outputs are exactly qf_fuse's; out_counts saturates at 65535.  Several times slower per
 * posting than qf_fuse (hash table and sort buffers in L2 instead of shared memory) -- the host layers call it only
 * above qf_fuse's envelope.
 */
QF_API int    qf_fuse_deep_max_postings(void);
QF_API size_t qf_fuse_deep_workspace_bytes(int Q, int R, int K);
QF_API int    qf_fuse_deep(const int32_t* doc_ids, const void* scores, int score_type, const int32_t* ranks,
                           const int32_t* lens, const double* weights,
                           int Q, int R, int K, int method, int norm, double rrf_k, uint32_t flags,
                           int32_t* out_ids, double* out_scores, uint16_t* out_counts, uint16_t* out_pos,
                           int32_t* out_lens, int32_t* out_status, void* ws, size_t ws_bytes, void* stream);

/*
 * Padded fusion outputs -> CSR: offsets[Q+1] (exclusive scan of out_lens, offsets[Q] = total), then the first
 * out_lens[q] entries of every row copied to out_ids/out_scores[offsets[q] ...).  Used before the device->host
 * copy / the .res writer so that |union| instead of R*K entries per query move.  out_* need capacity
 * sum(out_lens) (<= Q*stride).  ws: qf_compact_workspace_bytes(Q).
 */
QF_API size_t qf_compact_workspace_bytes(int Q);
QF_API int    qf_compact_fused(const int32_t* ids, const double* scores, const int32_t* lens, int Q, long long stride,
                               long long* offsets, int32_t* out_ids, double* out_scores, void* ws, size_t ws_bytes,
                               void* stream);

/* ------------------------------------------------------------------------------------------------
 * Run-file wire format (host side, multi-threaded, no CUDA)  --  replaces pd.read_csv(sep=r"\s+") in
 * load_runs (src/fusion.py:67-72), the row loop of _df_runs_to_ranx (:455-466) and the per-line formatting of
 * write_runfile (:480-492)
 * ---------------------------------------------------------------------------------------------- */
typedef struct qf_runs qf_runs;
/* Parse TREC run files ("{qid} Q0 {docno} {rank} {score} {tag}", whitespace separated) into dictionary-encoded
 * columns; one shared qid dictionary and one shared docno dictionary (first-seen order) for all files.  Scores are
 * kept as strtod doubles and, when every score has <= 6 decimals and fits, as exact int32 micro-units. */
QF_API int  qf_runs_parse(const char* const* paths, int n_files, int n_threads, qf_runs** out);
QF_API void qf_runs_free(qf_runs* runs);
QF_API int  qf_runs_info(const qf_runs* runs, int64_t* rows_per_file, int32_t* n_qids, int32_t* n_docs, int32_t* all_micro,
                         int64_t* qid_arena_bytes, int64_t* doc_arena_bytes);
QF_API int  qf_runs_rows(const qf_runs* runs, int file, int32_t* qid_idx, int32_t* doc_idx, int32_t* rank, int32_t* micro,
                         double* score);
QF_API int  qf_runs_dict(const qf_runs* runs, int which /*0 qid, 1 docno*/, char* arena, int64_t* offsets);
/* Host-side packing of the parsed columns into the dense [Q][R][K] layout qf_fuse reads, "weighted" semantics (every
 * row kept, file order inside a query: the per-query groupby of src/fusion.py:260-275).  qid_rank[code] = row of that
 * query in the packed arrays; outputs pre-initialised by the caller (doc_ids -1, the rest 0); pass micro OR scores. */
QF_API int  qf_pack_weighted(const int32_t* const* qid_cols, const int32_t* const* doc_cols, const int32_t* const* micro_cols,
                             const double* const* score_cols, const int32_t* const* rank_cols, const int64_t* n_rows, int R,
                             const int64_t* qid_rank, int n_qid_codes, int Q, int K, int32_t* doc_ids, int32_t* micro,
                             double* scores, int32_t* ranks);
/* Write "{qid} Q0 {docno} {rank} {score:.6f} {tag}" lines from CSR results (query j of the file = CSR row qsel[j],
 * already ranked); strings come from the dictionaries (arena + offsets). */
QF_API int  qf_runfile_write(const char* path, const char* qid_arena, const int64_t* qid_off, const int32_t* qsel,
                             int n_queries, const char* doc_arena, const int64_t* doc_off, const int64_t* offsets,
                             const int32_t* ids, const double* scores, const char* tag, int n_threads);

/* ------------------------------------------------------------------------------------------------
 * Bit-packed host->device transport of the run postings (csrc/qf_wire.cu).  The end-to-end step is bound by the PCIe
 * upload of 8 bytes per posting; dictionary-coded doc ids and 6-decimal scores need id_bits + sc_bits bits (22 + 20
 * for the NQ shape: 5.25 bytes).  Posting i (row-major position in [Q][R][K]) occupies bits [i*W, (i+1)*W) of a
 * little-endian uint64 stream, W = id_bits + sc_bits; low id_bits = doc id (>= 0), next sc_bits = micro - sc_base;
 * positions beyond lens[q][r] are zero.  Lossless.  Sits between the packer above and qf_qpp_stats / qf_fuse; the
 * reference has no counterpart (it reads its run files from disk, src/fusion.py:46-76).
 *   qf_wire_plan   (host) : smallest id_bits / sc_bits / sc_base for the valid postings
 *   qf_wire_words         : uint64 words of the stream (0 for a bad width: id_bits in [1,31], sc_bits in [1,32])
 *   qf_wire_pack   (host, multi-threaded) : QF_E_INVALID if a value does not fit
 *   qf_wire_unpack (device): expands into int32 doc_ids[n] / micro[n] (n = Q*R*K; 16-byte aligned)
 * sc_bits == 0 carries the doc ids alone (micro may be NULL); the scores then travel frame-of-reference coded:
 * a run list is sorted by score, so 32 consecutive scores span a small range.  Block b = list * ceil(K/32) + i/32
 * stores base[b] = its smallest valid score and 32 offsets of w_b = bits(max - min) bits each (0 for a constant or
 * empty block) as w_b uint32 words at payload[off[b] .. off[b+1]) -- 19 instead of 27 bits per score on bench.py's
 * synthetic runs, header included.
 *   qf_wire_for_blocks         : number of blocks = Q*R*ceil(K/32)  (base has that many entries, off one more)
 *   qf_wire_for_pack   (host)  : payload capacity >= 32 * blocks + 1 words; *words_used = words to upload
 *   qf_wire_for_unpack (device): expands into int32 micro[Q*R*K]
 * ---------------------------------------------------------------------------------------------- */
QF_API size_t qf_wire_words(long long n, int id_bits, int sc_bits);
QF_API int    qf_wire_plan(const int32_t* doc_ids, const int32_t* micro, const int32_t* lens, int Q, int R, int K,
                           int n_threads, int* id_bits, int* sc_bits, int32_t* sc_base);
QF_API int    qf_wire_pack(const int32_t* doc_ids, const int32_t* micro, const int32_t* lens, int Q, int R, int K,
                           int id_bits, int sc_bits, int32_t sc_base, uint64_t* out, size_t out_words, int n_threads);
QF_API int    qf_wire_unpack(const uint64_t* wire, long long n, int id_bits, int sc_bits, int32_t sc_base,
                             int32_t* doc_ids, int32_t* micro, void* stream);
QF_API long long qf_wire_for_blocks(int Q, int R, int K);
QF_API int    qf_wire_for_pack(const int32_t* micro, const int32_t* lens, int Q, int R, int K, int n_threads,
                               int32_t* base, uint32_t* off, uint32_t* payload, size_t payload_cap_words,
                               size_t* words_used);
QF_API int    qf_wire_for_unpack(const int32_t* base, const uint32_t* off, const uint32_t* payload, int Q, int R, int K,
                                 int32_t* micro, void* stream);

/* ------------------------------------------------------------------------------------------------
 * QPP statistics  --  replaces the JVM round trip QPPBridge.compute_batch -> qpp.QPPBridge
 *                     (src/qpp.py:117-171; src/qpp/java/qpp/QPPBridge.java:134-453)
 * ---------------------------------------------------------------------------------------------- */
#define QF_QPP_BRIDGE_COLS   13  /* nqc smv wig SigmaMax SigmaX RSD UEF MaxIDF avgidf cumnqc snqc dense-qpp dense-qpp-m */
#define QF_QPP_RESEARCH_COLS  5  /* rsd_bootstrap smv_research wig_research sigmax_research scnqc_research */
#define QF_QPP_COLS          (QF_QPP_BRIDGE_COLS + QF_QPP_RESEARCH_COLS)

/*
 * For each of the Q*R lists: scores[q][r][0..len) must be sorted descending (what
 * compute_qpp_for_res_file's sorted(reverse=True), src/qpp.py:304-305, produces; violations set
 * out_status).  The 13 bridge columns use the first min(k_cut, len, top_k) scores rounded to float32
 * (Gson getAsFloat, QPPBridge.java:69) and widened to double; out[q][r][0..13) is in
 * config.qpp.methods order (config/defaults.yaml:164-177).
 *
 * text_feats[q] = {n_terms, n_unique, max_len, sum_len} of queryText.toLowerCase().split("\\s+")
 * (Java semantics; computed on the host, QPPBridge.java:250,272,369,386).
 *
 * Research columns out[q][r][13..18) (RSDSpecificity.java:22-52 bootstrap over NQCSpecificity,
 * SMVSpecificity, WIGSpecificity, SigmaXSpecificity, NQCCalibratedSpecificity) are computed when
 * `perms` != NULL: perms[s][0..perm_len) are caller-supplied permutations of range(perm_len)
 * (perm_len >= K), one per bootstrap sample (n_samples, reference uses 10); query_idf[q] (nullable,
 * default 1.0) stands in for maxIDF(q), which needs a Lucene index.  Without perms the 5 columns are 0.
 */
QF_API int qf_qpp_stats(const void* scores, int score_type, const int32_t* lens, const int32_t* text_feats,
                 const double* query_idf, const int32_t* perms, int n_samples, int perm_len,
                 int Q, int R, int K, int top_k, int k_cut, int wig_k,
                 double* out, int32_t* out_status, void* stream);

/*
 * Cross-query column normalisation of one retriever's [Q][ncols] block, _normalize_qpp
 * (src/qpp.py:340-374): mode 1 = minmax ((v-min)/(max-min if max>min else 1)), 2 = zscore
 * ((v-mean)/(std_pop or 1)).  `vals` is [Q][R][ncols]; each (r, col) is normalised over q, in place.
 * When `round6` != 0 the result is additionally passed through the "%.6f" text round trip
 * (src/qpp.py:333 -> src/data_utils.py:290) so it equals the value fusion reads back from the .qpp file.
 * ws: qf_qpp_colnorm_workspace_bytes(R, ncols) bytes of device scratch.
 */
#define QF_COLNORM_MINMAX 1
#define QF_COLNORM_ZSCORE 2
QF_API size_t qf_qpp_colnorm_workspace_bytes(int R, int ncols);
QF_API int qf_qpp_colnorm(double* vals, int Q, int R, int ncols, int mode, int round6,
                   void* ws, size_t ws_bytes, void* stream);

/*
 * The same normalisation split in two for query-sharded multi-GPU runs (the ONE cross-query dependency of
 * the QPP/fusion path): qf_qpp_colstats writes stats[r*ncols+c] = {min, max, sum, sum of squared deviations
 * from mean[] (0 when mean == NULL)} of this rank's rows; the host all-reduces (MIN, MAX, SUM) over ranks,
 * derives (base, scale) exactly as _normalize_qpp does and calls qf_qpp_colnorm_apply on every rank.
 * base == scale == NULL applies only the "%.6f" round trip.
 */
QF_API int qf_qpp_colstats(const double* vals, int Q, int R, int ncols, const double* mean, double* stats,
                           void* ws, size_t ws_bytes, void* stream);
QF_API int qf_qpp_colnorm_apply(double* vals, int Q, int R, int ncols, const double* base, const double* scale,
                                int round6, void* stream);

/*
 * weights[q][r] = vals[q][r][col]  (col >= 0)  or  mean over the first `ncols_mean` columns
 * (col == -1, get_qpp_weight's fusion mode, src/fusion.py:152-153: sum(scores)/len(scores)).
 */
QF_API int qf_qpp_select_weight(const double* vals, int Q, int R, int ncols, int col, int ncols_mean,
                         double* weights, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Dense retrieval  --  replaces faiss.IndexFlatIP.search behind BGERetriever._search_batch /
 *                      _search_faiss_batch (src/retrievers/bge.py:179-194; also the flat-IP search of
 *                      TCTColBERTRetriever, src/retrievers/tct_colbert.py:139-151)
 * ---------------------------------------------------------------------------------------------- */
/*
 * Exact inner-product top-k of queries[Q][D] against corpus[N][D], both bf16 row-major, fp32 accumulate
 * on the tcgen05 tensor cores (the reference scores in fp32; see DESIGN.md for the bf16 tolerance).
 * Output rows are sorted by (score descending, id ascending); ids are corpus row + id_offset (the
 * shard offset under corpus sharding); rows with fewer than k hits are padded with id -1 / score -inf,
 * as FAISS does.  out_status[q] (nullable): 0 ok, bit 0 = the candidate threshold admitted fewer than
 * k documents, bit 1 = more than the candidate capacity (results for that query are then NOT exact and
 * the host wrapper re-runs it; never silently accepted).
 * Envelope: D multiple of 64 and <= 768 (the query tile is held in tensor memory), k <= qf_dense_max_k().
 * ws: qf_dense_workspace_bytes(Q, N, k) bytes of device scratch.
 */
QF_API size_t qf_dense_workspace_bytes(int Q, int N, int k);
QF_API int    qf_dense_max_k(void);
QF_API int    qf_dense_topk(const void* queries_bf16, const void* corpus_bf16, int Q, int N, int D, int k,
                            long long id_offset, float* out_scores, int32_t* out_ids, int32_t* out_status,
                            void* ws, size_t ws_bytes, void* stream);

/*
 * Exact fallback for queries whose out_status is non-zero (and a building block for tests):
 * qf_dense_scores writes the full fp32 score rows out[q][d] (rows padded to a multiple of 128, row stride ld >=
 * 64*ceil(N/64)) with the same tcgen05 kernel; qf_dense_row_topk streams stored rows and writes the exact top-k
 * (score desc, id asc; exact under arbitrary ties) of stored row r into output row rows[r], clearing its status.
 */
QF_API int    qf_dense_scores(const void* queries_bf16, const void* corpus_bf16, int Q, int N, int D, float* out,
                              long long ld, void* stream);
QF_API int    qf_dense_row_topk(const float* scores, long long ld, const int32_t* rows, int n_rows, int N, int k,
                                long long id_offset, float* out_scores, int32_t* out_ids, int32_t* out_status,
                                void* stream);

/*
 * float32 re-scoring of an over-fetched candidate list (closes the bf16-vs-fp32 gap to the reference's float32
 * faiss.IndexFlatIP, src/retrievers/bge.py:188-194): out_scores[q][j] = <queries_f32[q], corpus_f32[cand_ids[q][j] -
 * id_offset]> with float32 FMAs (fixed summation order), -inf for padding ids (< 0) or ids outside [id_offset,
 * id_offset + N).  The host layer searches top (k + margin) in bf16, re-scores, and keeps the exact top k
 * (qf_dense_row_topk over the re-scored rows).
 */
QF_API int    qf_dense_rescore_f32(const float* queries_f32, const float* corpus_f32, int Q, long long N, int D,
                                   const int32_t* cand_ids, int kc, long long id_offset, float* out_scores, void* stream);

/*
 * k-way merge after the all-gather of per-shard results (multi-GPU dense retrieval; precedent: the
 * reference's HNSW segment merge, src/retrievers/bge.py:216-237): scores/ids are [G][Q][k] as
 * ncclAllGather leaves them, ids global, -1 = padding.  Output [Q][k], (score desc, id asc).
 * Envelope: G*k <= 8192.
 */
QF_API int    qf_topk_merge(const float* scores, const int32_t* ids, int G, int Q, int k,
                            float* out_scores, int32_t* out_ids, void* stream);

/*
 * Packed variants for the ONE-exchange multi-GPU search (qppfusion/distributed.py): a query's result is a pair of
 * adjacent rows, k float32 score bits followed by k int32 ids ("[2][k] int32"), so scores and ids travel in one
 * all-to-all and one all-gather instead of two each.
 *   qf_dense_topk_packed : as qf_dense_topk, query q written at
 *                          out_pairs + (q / chunk_rows) * chunk_stride + (q % chunk_rows) * 2k      (int32 elements)
 *                          (chunk_rows = 0: plain [Q][2][k]).  chunk_rows = ceil(Q / G) with chunk_stride =
 *                          (chunk_rows + 1) * 2k leaves one spare row per destination rank, which the host layer
 *                          uses to ship the rank's status flag inside the same all-to-all.
 *   qf_topk_merge_packed : segment g, query j read at in_pairs + g * seg_stride + j * 2 k_in (k_in entries per shard
 *                          list); output [Q][2][k_out].  k_in < k_out is the TRUNCATED exchange: with G shards a
 *                          shard holds about k_out / G of the global top k_out, so it sends only its top
 *                          k_in ~ 2 k_out / G + 64 (4x fewer bytes and a cheaper local search at G = 8).
 *   qf_topk_merge_check  : sets *flag (device int32) when that truncation may have lost a document: some shard's
 *                          list is full and its last score reaches the merged k_out-th score.  The host layer then
 *                          repeats the search with full-length lists (exactness is never traded).
 */
QF_API int    qf_dense_topk_packed(const void* queries_bf16, const void* corpus_bf16, int Q, int N, int D, int k,
                                   long long id_offset, int32_t* out_pairs, int chunk_rows, long long chunk_stride,
                                   int32_t* out_status, void* ws, size_t ws_bytes, void* stream);
QF_API int    qf_topk_merge_packed(const int32_t* in_pairs, int G, int Q, int k_in, long long seg_stride, int k_out,
                                   int32_t* out_pairs, void* stream);
QF_API int    qf_topk_merge_check(const int32_t* in_pairs, int G, int Q, int k_in, long long seg_stride,
                                  const int32_t* merged_pairs, int k_out, int32_t* flag, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Query encoder  --  replaces SentenceTransformer(model).encode(queries, normalize_embeddings=True) behind
 *                    BGERetriever._encode_queries (src/retrievers/bge.py:153-170): HF BertModel (post-LN encoder,
 *                    exact erf GELU, LayerNorm eps in the config) + CLS pooling + L2 normalisation, which is what
 *                    BAAI/bge-base-en-v1.5 (config/defaults.yaml:60-69) is.  Tokenisation stays on the host.
 * ---------------------------------------------------------------------------------------------- */
#define QF_BERT_POOL_CLS  0
#define QF_BERT_POOL_MEAN 1   /* mean over the unpadded tokens (other sentence-transformers checkpoints) */

typedef struct qf_bert_config {
  int   n_layers;   /* 12 for bge-base */
  int   hidden;     /* 768; envelope: 256, 512, 768 or 1024, head dimension 64 */
  int   heads;      /* hidden / 64 */
  int   ffn;        /* 3072 */
  int   vocab;      /* 30522 */
  int   max_pos;    /* 512 */
  int   pooling;    /* QF_BERT_POOL_* */
  float ln_eps;     /* 1e-12 */
} qf_bert_config;

/*
 * Weights: ONE packed device blob of bf16 elements, tensors in this order, each row-major as PyTorch stores them:
 *   word_embeddings [vocab, H], position_embeddings [max_pos, H], token_type_embeddings [2, H],
 *   embeddings.LayerNorm weight [H], bias [H], then per layer:
 *   W_qkv [3H, H] (query, key, value weights stacked), b_qkv [3H], W_o [H, H], b_o [H], LN1 weight [H], bias [H],
 *   W_1 [ffn, H], b_1 [ffn], W_2 [H, ffn], b_2 [H], LN2 weight [H], bias [H].
 * qf_bert_weight_elems returns the element count (0 for a config outside the envelope).
 */
QF_API size_t qf_bert_weight_elems(const qf_bert_config* cfg);
QF_API size_t qf_bert_workspace_bytes(const qf_bert_config* cfg, int B, int L);
/*
 * One Linear layer of the encoder on the tensor cores (own tcgen05 kernel, csrc/qf_gemm.cu; round 1 called cuBLASLt):
 *   Y[M][N] (bf16) = act( X[M][K] (bf16) . W[N][K]^T (bf16, the nn.Linear weight) + bias[N] (bf16) ),
 *   act: 0 = identity, 1 = erf-GELU (HF BERT's "gelu").  fp32 accumulation in tensor memory.
 * Envelope: M % 128 == 0, N % 128 == 0, K % 64 == 0, operands 16-byte aligned.  qf_bert_encode uses it for every
 * Linear layer (the FFN-1 GELU is fused into its epilogue).
 */
QF_API int    qf_linear_bf16(const void* W, const void* X, const void* bias, void* Y, long long M, int N, int K, int act,
                             void* stream);

/*
 * tokens[B][L] int32 (padded with anything; positions >= lens[b] are masked as keys and never pooled), lens[B]
 * (nullable = all L; clamped to [1, L]).  Outputs (either may be null): out_bf16[B][H] -- directly the `queries_bf16`
 * of qf_dense_topk -- and out_f32[B][H], the L2-normalised embedding.  out_status[b] (nullable): bit 0 = a token id
 * outside [0, vocab) (replaced by id 0, never read out of bounds).  bf16 activations, fp32 accumulation and
 * statistics; the four Linear layers per block run through qf_linear_bf16's tcgen05 kernel (cuBLASLt only for shapes
 * outside its envelope), everything else is in csrc/qf_bert.cu.
 */
QF_API int    qf_bert_encode(const qf_bert_config* cfg, const void* weights_bf16, const int32_t* tokens,
                             const int32_t* lens, int B, int L, void* out_bf16, float* out_f32, int32_t* out_status,
                             void* ws, size_t ws_bytes, void* stream);
/*
 * Packed (variable-length) form: the B sequences lie back to back, tokens[M], sequence b = tokens[cu_seqlens[b] ..
 * cu_seqlens[b+1]) (cu_seqlens int32 [B+1] on the device, cu_seqlens[0] = 0, cu_seqlens[B] = M), no padding rows at
 * all -- every GEMM row is a real token.  max_len bounds the sequence lengths (sizes the attention kernel's shared
 * memory); workspace = qf_bert_workspace_bytes(cfg, 1, M).  out_status[b]: bit 0 as above, bit 1 = the sequence's
 * length is outside [1, max_len] (tokens beyond max_len are ignored, an empty sequence yields a zero vector).
 */
QF_API int    qf_bert_encode_packed(const qf_bert_config* cfg, const void* weights_bf16, const int32_t* tokens,
                                    const int32_t* cu_seqlens, int B, long long M, int max_len, void* out_bf16,
                                    float* out_f32, int32_t* out_status, void* ws, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * WordPiece tokeniser (host side, multi-threaded, no CUDA)  --  replaces the HF BertTokenizer call inside
 * SentenceTransformer.encode (BGERetriever._encode_queries, src/retrievers/bge.py:160-170) for texts below U+3000
 * that NFC leaves alone (ASCII, accented Latin, Greek, Cyrillic, punctuation); other texts are flagged, not tokenised
 * ---------------------------------------------------------------------------------------------- */
typedef struct qf_wordpiece qf_wordpiece;
/* vocab = the bytes of a BERT vocab.txt (one token per line, id = line number); needs [CLS] [SEP] [UNK] [PAD]. */
QF_API int  qf_wordpiece_create(const char* vocab, size_t vocab_bytes, int do_lower_case, qf_wordpiece** out);
QF_API void qf_wordpiece_free(qf_wordpiece* wp);
/* texts: n UTF-8 strings back to back, text i = bytes [offsets[i], offsets[i+1]).  out_tokens[n][row_cap]
 * ([CLS] pieces [SEP], truncated to max_len <= row_cap tokens, padded with [PAD]), out_lens[n].  out_flags[i] = 1:
 * the text holds a character outside the native scope (combining mark, CJK, malformed UTF-8 ...) and was NOT tokenised
 * (out_lens[i] = 0) -- the caller runs its full Unicode path for it. */
QF_API int  qf_wordpiece_encode(const qf_wordpiece* wp, const char* texts, const int64_t* offsets, int n, int max_len,
                                int row_cap, int n_threads, int32_t* out_tokens, int32_t* out_lens, uint8_t* out_flags);

#ifdef __cplusplus
}
#endif
#endif /* QPPFUSION_H_ */
