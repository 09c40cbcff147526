// qf_io: native (host, multi-threaded) TREC run-file parser and writer -- the wire formats either side of the
// fusion kernel (reference: pd.read_csv(sep=r"\s+") in load_runs, src/fusion.py:67-72; the row loop of
// _df_runs_to_ranx, :455-466; write_runfile, :480-492).  Once qf_fuse runs at millions of queries per second the
// regex CSV reader, the docno string factorisation and the per-line Python formatting dominate a file-to-file run
// by four orders of magnitude; this module parses "{qid} Q0 {docno} {rank} {score} {tag}" lines straight into
// dictionary-encoded integer arrays (scores as exact int32 micro-units when the text has <= 6 decimals, and as
// strtod doubles always) and writes "{qid} Q0 {docid} {rank} {score:.6f} {tag}" lines from CSR results.
// Pure host code: no CUDA calls, usable on a CPU-only box.
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "qf_common.cuh"

namespace {

struct Tok {
  const char* p;
  int len;
};

struct Row {
  Tok qid, doc;
  uint64_t doc_hash;
  int32_t rank;
  int32_t micro;
  double score;
  bool micro_ok;
};

inline bool is_ws(char c) { return c == ' ' || c == '\t' || c == '\r' || c == '\f' || c == '\v'; }

// exact decimal -> micro-units: [+-]digits[.digits{0,6}] with |value| < 2147.483647, else false
inline bool parse_micro(const char* s, int len, int32_t* out) {
  int i = 0;
  bool neg = false;
  if (i < len && (s[i] == '-' || s[i] == '+')) neg = s[i++] == '-';
  long long ip = 0;
  int nd = 0;
  while (i < len && s[i] >= '0' && s[i] <= '9') {
    ip = ip * 10 + (s[i++] - '0');
    if (++nd > 6) return false;
  }
  long long fp = 0;
  int nf = 0;
  if (i < len && s[i] == '.') {
    ++i;
    while (i < len && s[i] >= '0' && s[i] <= '9') {
      if (nf == 6) return false;
      fp = fp * 10 + (s[i++] - '0');
      ++nf;
    }
  }
  if (i != len || (nd == 0 && nf == 0)) return false;
  for (int k = nf; k < 6; ++k) fp *= 10;
  const long long v = ip * 1000000LL + fp;
  if (v > 2147483647LL) return false;
  *out = (int32_t)(neg ? -v : v);
  return true;
}

struct Dict {
  // open addressing over (hash, arena offset, length); arena stores the strings back to back
  std::vector<uint64_t> hashes;
  std::vector<int32_t> slots;   // index into entries or -1
  std::vector<int64_t> off;     // entry -> arena offset (off.size() == n+1)
  std::string arena;
  size_t mask = 0;
  Dict() { rehash(1 << 16); off.push_back(0); }
  static uint64_t hash(const char* p, int len) {
    uint64_t h = 1469598103934665603ull;
    for (int i = 0; i < len; ++i) { h ^= (unsigned char)p[i]; h *= 1099511628211ull; }
    return h ^ (h >> 29);
  }
  void rehash(size_t cap) {
    std::vector<int32_t> ns(cap, -1);
    const size_t nm = cap - 1;
    for (size_t e = 0; e < hashes.size(); ++e) {
      size_t i = hashes[e] & nm;
      while (ns[i] >= 0) i = (i + 1) & nm;
      ns[i] = (int32_t)e;
    }
    slots.swap(ns);
    mask = nm;
  }
  void prefetch(uint64_t h) const { __builtin_prefetch(&slots[h & mask]); }
  int32_t get(const char* p, int len) { return get_hashed(p, len, hash(p, len)); }
  int32_t get_hashed(const char* p, int len, uint64_t h) {
    size_t i = h & mask;
    while (true) {
      const int32_t e = slots[i];
      if (e < 0) break;
      if (hashes[e] == h && off[e + 1] - off[e] == len && memcmp(arena.data() + off[e], p, (size_t)len) == 0) return e;
      i = (i + 1) & mask;
    }
    const int32_t e = (int32_t)hashes.size();
    hashes.push_back(h);
    arena.append(p, (size_t)len);
    off.push_back((int64_t)arena.size());
    slots[i] = e;
    if (hashes.size() * 2 > slots.size()) rehash(slots.size() * 2);
    return e;
  }
  int32_t size() const { return (int32_t)hashes.size(); }
};

struct FileRows {
  std::vector<int32_t> qid, doc, rank, micro;
  std::vector<double> score;
};

constexpr int DOC_SHARDS = 256;        // docno dictionary shards (top 8 bits of the hash), each owned by one thread
constexpr int LOCAL_BITS = 23;         // (shard << 23 | code inside the shard) until the shard bases are known

}  // namespace

struct qf_runs {
  std::vector<FileRows> files;
  Dict qids;
  std::vector<Dict> doc_shards;
  std::vector<int32_t> shard_base;     // global docno code = shard_base[shard] + code inside the shard
  std::string doc_arena;               // merged docno dictionary (shard by shard)
  std::vector<int64_t> doc_off;
  bool all_micro = true;
  std::string error;
  qf_runs() : doc_shards(DOC_SHARDS) {}
};

namespace {

inline int doc_shard(uint64_t h) { return (int)(h >> 56); }

// digits with an optional sign -> int32; false on anything else (caller falls back to strtol)
inline bool parse_int(const char* s, int len, int32_t* out) {
  int i = 0;
  bool neg = false;
  if (i < len && (s[i] == '-' || s[i] == '+')) neg = s[i++] == '-';
  if (i >= len || len - i > 9) return false;
  int32_t v = 0;
  for (; i < len; ++i) {
    if (s[i] < '0' || s[i] > '9') return false;
    v = v * 10 + (s[i] - '0');
  }
  *out = neg ? -v : v;
  return true;
}

// `owned[t]` collects the indices of this chunk's rows whose docno shard belongs to thread t (phase 3 then touches
// only its own rows instead of scanning everybody's)
bool parse_chunk(const char* b, const char* e, std::vector<Row>& rows, std::vector<std::vector<uint32_t>>& owned, int nt,
                 std::string& err) {
  rows.reserve((size_t)(e - b) / 32 + 16);
  owned.assign((size_t)nt, std::vector<uint32_t>());
  for (auto& o : owned) o.reserve((size_t)(e - b) / 32 / (size_t)nt + 16);
  const char* p = b;
  while (p < e) {
    const char* eol = (const char*)memchr(p, '\n', (size_t)(e - p));
    if (!eol) eol = e;
    Tok f[6];
    int nf = 0;
    const char* c = p;
    while (c < eol && nf < 6) {
      while (c < eol && is_ws(*c)) ++c;
      if (c >= eol) break;
      const char* s = c;
      while (c < eol && !is_ws(*c)) ++c;
      f[nf].p = s;
      f[nf].len = (int)(c - s);
      ++nf;
    }
    if (nf >= 5) {
      Row r;
      r.qid = f[0];
      r.doc = f[2];
      r.doc_hash = Dict::hash(f[2].p, f[2].len);
      char buf[64];
      char* endp = nullptr;
      if (!parse_int(f[3].p, f[3].len, &r.rank)) {
        const int rl = f[3].len < 63 ? f[3].len : 63;
        memcpy(buf, f[3].p, (size_t)rl);
        buf[rl] = 0;
        const long rk = strtol(buf, &endp, 10);
        if (*endp != 0 || rk < -2147483647L || rk > 2147483647L) { err = "bad rank field '" + std::string(buf) + "'"; return false; }
        r.rank = (int32_t)rk;
      }
      r.micro_ok = parse_micro(f[4].p, f[4].len, &r.micro);
      if (r.micro_ok) {
        // <= 6 decimals: the value is micro / 10^6 exactly, and the IEEE quotient of two exactly representable
        // numbers is the double nearest to it -- the same double strtod returns (sign of a zero kept)
        r.score = (double)r.micro / 1.0e6;
        if (r.micro == 0 && f[4].p[0] == '-') r.score = -0.0;
      } else {
        r.micro = 0;
        const int sl = f[4].len < 63 ? f[4].len : 63;
        memcpy(buf, f[4].p, (size_t)sl);
        buf[sl] = 0;
        r.score = strtod(buf, &endp);
        if (*endp != 0) { err = "bad score field '" + std::string(buf) + "'"; return false; }
      }
      owned[(size_t)(doc_shard(r.doc_hash) % nt)].push_back((uint32_t)rows.size());
      rows.push_back(r);
    } else if (nf != 0) {
      err = "line with " + std::to_string(nf) + " fields (need >= 5)";
      return false;
    }
    p = eol + 1;
  }
  return true;
}

template <typename F>
void parallel_for(int nt, F fn) {
  if (nt <= 1) { fn(0); return; }
  std::vector<std::thread> th;
  for (int t = 0; t < nt; ++t) th.emplace_back([&, t]() { fn(t); });
  for (auto& x : th) x.join();
}

}  // namespace

extern "C" int qf_runs_parse(const char* const* paths, int n_files, int n_threads, qf_runs** out) {
  QF_REQUIRE(paths && out && n_files >= 1, QF_E_INVALID, "qf_runs_parse: bad arguments");
  if (n_threads < 1) n_threads = 1;
  if (n_threads > DOC_SHARDS) n_threads = DOC_SHARDS;
  qf_runs* R = new qf_runs();
  R->files.resize((size_t)n_files);
  for (int fi = 0; fi < n_files; ++fi) {
    const int fd = open(paths[fi], O_RDONLY);
    if (fd < 0) { qf::set_error("qf_runs_parse: cannot open %s", paths[fi]); delete R; return QF_E_INVALID; }
    struct stat st;
    fstat(fd, &st);
    const size_t sz = (size_t)st.st_size;
    const char* base = sz ? (const char*)mmap(nullptr, sz, PROT_READ, MAP_PRIVATE, fd, 0) : nullptr;
    if (sz && base == MAP_FAILED) { close(fd); qf::set_error("qf_runs_parse: mmap failed for %s", paths[fi]); delete R; return QF_E_INVALID; }
    // 1. split at newline boundaries, tokenise + convert numbers in parallel
    const int nt = sz < (1u << 20) ? 1 : n_threads;
    std::vector<const char*> cut((size_t)nt + 1);
    cut[0] = base;
    cut[(size_t)nt] = base + sz;
    for (int t = 1; t < nt; ++t) {
      const char* p = base + sz * (size_t)t / (size_t)nt;
      const char* nl = (const char*)memchr(p, '\n', (size_t)(base + sz - p));
      cut[(size_t)t] = nl ? nl + 1 : base + sz;
    }
    std::vector<std::vector<Row>> parts((size_t)nt);
    std::vector<std::vector<std::vector<uint32_t>>> owned((size_t)nt);
    std::vector<std::string> errs((size_t)nt);
    std::vector<char> ok((size_t)nt, 1);
    parallel_for(nt, [&](int t) {
      ok[(size_t)t] = parse_chunk(cut[(size_t)t], cut[(size_t)t + 1], parts[(size_t)t], owned[(size_t)t], nt,
                                  errs[(size_t)t]) ? 1 : 0;
    });
    for (int t = 0; t < nt; ++t)
      if (!ok[(size_t)t]) {
        qf::set_error("qf_runs_parse: %s: %s", paths[fi], errs[(size_t)t].c_str());
        if (sz) munmap((void*)base, sz);
        close(fd);
        delete R;
        return QF_E_INVALID;
      }
    std::vector<size_t> part_off((size_t)nt + 1, 0);
    for (int t = 0; t < nt; ++t) part_off[(size_t)t + 1] = part_off[(size_t)t] + parts[(size_t)t].size();
    const size_t n = part_off[(size_t)nt];
    FileRows& F = R->files[(size_t)fi];
    F.qid.resize(n); F.doc.resize(n); F.rank.resize(n); F.micro.resize(n); F.score.resize(n);
    // 2. qid codes: serial, but almost every row repeats the previous row's qid (one memcmp)
    {
      Tok last_q{nullptr, 0};
      int32_t last_qi = -1;
      for (int t = 0; t < nt; ++t) {
        const std::vector<Row>& v = parts[(size_t)t];
        int32_t* dst = F.qid.data() + part_off[(size_t)t];
        for (size_t ri = 0; ri < v.size(); ++ri) {
          const Row& r = v[ri];
          if (last_qi < 0 || r.qid.len != last_q.len || memcmp(r.qid.p, last_q.p, (size_t)r.qid.len) != 0) {
            last_qi = R->qids.get(r.qid.p, r.qid.len);
            last_q = r.qid;
          }
          dst[ri] = last_qi;
        }
      }
    }
    // 3. numeric columns (chunk-parallel) and docno codes (shard-parallel: thread t owns the dictionary shards
    //    s with s % nt == t and visits the rows in file order, so codes are deterministic)
    std::vector<char> micro_bad((size_t)nt, 0);
    parallel_for(nt, [&](int t) {
      const std::vector<Row>& mine = parts[(size_t)t];
      const size_t o = part_off[(size_t)t];
      for (size_t ri = 0; ri < mine.size(); ++ri) {
        F.rank[o + ri] = mine[ri].rank;
        F.micro[o + ri] = mine[ri].micro;
        F.score[o + ri] = mine[ri].score;
        if (!mine[ri].micro_ok) micro_bad[(size_t)t] = 1;
      }
      for (int pt = 0; pt < nt; ++pt) {
        const std::vector<Row>& v = parts[(size_t)pt];
        const std::vector<uint32_t>& idx = owned[(size_t)pt][(size_t)t];
        int32_t* dst = F.doc.data() + part_off[(size_t)pt];
        for (size_t k = 0; k < idx.size(); ++k) {
          const Row& r = v[idx[k]];
          const int s_ = doc_shard(r.doc_hash);
          if (k + 8 < idx.size()) R->doc_shards[(size_t)doc_shard(v[idx[k + 8]].doc_hash)].prefetch(v[idx[k + 8]].doc_hash);
          const int32_t code = R->doc_shards[(size_t)s_].get_hashed(r.doc.p, r.doc.len, r.doc_hash);
          dst[idx[k]] = (int32_t)(((uint32_t)s_ << LOCAL_BITS) | (uint32_t)code);
        }
      }
    });
    for (int t = 0; t < nt; ++t)
      if (micro_bad[(size_t)t]) R->all_micro = false;
    if (sz) munmap((void*)base, sz);
    close(fd);
  }
  // 4. shard bases, merged dictionary, final codes
  R->shard_base.assign(DOC_SHARDS + 1, 0);
  for (int s_ = 0; s_ < DOC_SHARDS; ++s_) {
    if (R->doc_shards[(size_t)s_].size() >= (1 << LOCAL_BITS)) {
      qf::set_error("qf_runs_parse: more than %d distinct docnos in one dictionary shard", 1 << LOCAL_BITS);
      delete R;
      return QF_E_UNSUPPORTED;
    }
    R->shard_base[(size_t)s_ + 1] = R->shard_base[(size_t)s_] + R->doc_shards[(size_t)s_].size();
  }
  {
    size_t bytes = 0;
    for (auto& d : R->doc_shards) bytes += d.arena.size();
    R->doc_arena.reserve(bytes);
    R->doc_off.reserve((size_t)R->shard_base[DOC_SHARDS] + 1);
    R->doc_off.push_back(0);
    for (auto& d : R->doc_shards) {
      const int64_t at = (int64_t)R->doc_arena.size();
      R->doc_arena.append(d.arena);
      for (size_t e = 1; e < d.off.size(); ++e) R->doc_off.push_back(at + d.off[e]);
      Dict().hashes.swap(d.hashes);          // the shard tables are no longer needed
      std::vector<int32_t>().swap(d.slots);
      std::string().swap(d.arena);
    }
  }
  for (auto& F : R->files) {
    const size_t n = F.doc.size();
    const int nt = n < (1u << 16) ? 1 : n_threads;
    parallel_for(nt, [&](int t) {
      const size_t i0 = n * (size_t)t / (size_t)nt, i1 = n * (size_t)(t + 1) / (size_t)nt;
      for (size_t i = i0; i < i1; ++i) {
        const uint32_t v = (uint32_t)F.doc[i];
        F.doc[i] = R->shard_base[v >> LOCAL_BITS] + (int32_t)(v & ((1u << LOCAL_BITS) - 1u));
      }
    });
  }
  *out = R;
  return QF_OK;
}

extern "C" void qf_runs_free(qf_runs* r) { delete r; }

extern "C" int qf_runs_info(const qf_runs* r, int64_t* rows_per_file, int32_t* n_qids, int32_t* n_docs, int32_t* all_micro,
                            int64_t* qid_arena_bytes, int64_t* doc_arena_bytes) {
  QF_REQUIRE(r, QF_E_INVALID, "qf_runs_info: null handle");
  if (rows_per_file)
    for (size_t i = 0; i < r->files.size(); ++i) rows_per_file[i] = (int64_t)r->files[i].qid.size();
  if (n_qids) *n_qids = r->qids.size();
  if (n_docs) *n_docs = (int32_t)(r->doc_off.size() - 1);
  if (all_micro) *all_micro = r->all_micro ? 1 : 0;
  if (qid_arena_bytes) *qid_arena_bytes = (int64_t)r->qids.arena.size();
  if (doc_arena_bytes) *doc_arena_bytes = (int64_t)r->doc_arena.size();
  return QF_OK;
}

// copy one file's columns into caller buffers (each nullable), rows_per_file[file] elements
extern "C" int qf_runs_rows(const qf_runs* r, int file, int32_t* qid_idx, int32_t* doc_idx, int32_t* rank, int32_t* micro,
                            double* score) {
  QF_REQUIRE(r && file >= 0 && (size_t)file < r->files.size(), QF_E_INVALID, "qf_runs_rows: bad arguments");
  const FileRows& F = r->files[(size_t)file];
  const size_t n = F.qid.size();
  if (qid_idx) memcpy(qid_idx, F.qid.data(), n * 4);
  if (doc_idx) memcpy(doc_idx, F.doc.data(), n * 4);
  if (rank) memcpy(rank, F.rank.data(), n * 4);
  if (micro) memcpy(micro, F.micro.data(), n * 4);
  if (score) memcpy(score, F.score.data(), n * 8);
  return QF_OK;
}

// which = 0: qid dictionary, 1: docno dictionary.  arena: concatenated strings, offsets: n+1 entries.
extern "C" int qf_runs_dict(const qf_runs* r, int which, char* arena, int64_t* offsets) {
  QF_REQUIRE(r && (which == 0 || which == 1), QF_E_INVALID, "qf_runs_dict: bad arguments");
  if (which == 0) {
    if (arena) memcpy(arena, r->qids.arena.data(), r->qids.arena.size());
    if (offsets) memcpy(offsets, r->qids.off.data(), r->qids.off.size() * 8);
  } else {
    if (arena) memcpy(arena, r->doc_arena.data(), r->doc_arena.size());
    if (offsets) memcpy(offsets, r->doc_off.data(), r->doc_off.size() * 8);
  }
  return QF_OK;
}

// Dense packing of parsed run files for qf_fuse, "weighted" semantics (rows kept as they are, in file order, grouped
// by query: src/fusion.py:260-275): file r's row i goes to [q][r][w] with q = qid_rank[qid[i]] and w = the number of
// earlier rows of that file with the same query.  One thread per file; outputs are pre-initialised by the caller
// (doc ids -1, scores / ranks 0).  Replaces a numpy argsort + fancy-index scatter that cost 75 ns per row.
extern "C" int qf_pack_weighted(const int32_t* const* qid_cols, const int32_t* const* doc_cols, const int32_t* const* micro_cols,
                                const double* const* score_cols, const int32_t* const* rank_cols, const int64_t* n_rows, int R,
                                const int64_t* qid_rank, int n_qid_codes, int Q, int K, int32_t* doc_ids, int32_t* micro,
                                double* scores, int32_t* ranks) {
  QF_REQUIRE(qid_cols && doc_cols && n_rows && qid_rank && doc_ids && R >= 1 && Q >= 0 && K >= 1 && (micro || scores),
             QF_E_INVALID, "qf_pack_weighted: bad arguments");
  std::vector<std::thread> th;
  std::vector<int> bad((size_t)R, 0);
  for (int r = 0; r < R; ++r)
    th.emplace_back([&, r]() {
      std::vector<int32_t> cursor((size_t)Q, 0);
      const int32_t* qc = qid_cols[r];
      const int32_t* dc = doc_cols[r];
      const int64_t n = n_rows[r];
      for (int64_t i = 0; i < n; ++i) {
        const int32_t code = qc[i];
        if (code < 0 || code >= n_qid_codes) { bad[(size_t)r] = 1; return; }
        const int64_t q = qid_rank[code];
        const int32_t w = cursor[(size_t)q]++;
        if (w >= K) { bad[(size_t)r] = 1; return; }
        const size_t at = ((size_t)q * (size_t)R + (size_t)r) * (size_t)K + (size_t)w;
        doc_ids[at] = dc[i];
        if (micro) micro[at] = micro_cols[r][i];
        else scores[at] = score_cols[r][i];
        if (ranks) ranks[at] = rank_cols[r][i];
      }
    });
  for (auto& t : th) t.join();
  for (int r = 0; r < R; ++r)
    QF_REQUIRE(!bad[(size_t)r], QF_E_INVALID, "qf_pack_weighted: file %d has a row outside the packed shape", r);
  return QF_OK;
}

// "%d" / "%.6f" without printf.  fmt_fixed6 reproduces glibc's (= Python's) correctly rounded "%.6f" for
// |x| < 1e9: x * 10^6 is rounded to the nearest integer (ties to even) EXACTLY -- pr = fl(x * 1e6) and the FMA
// residual err = x * 1e6 - pr give the true product as pr + err, so the distance to the rounding boundary is decided
// without error (same construction as the device-side round6, csrc/qf_qpp.cu).  Anything else falls back to snprintf.
inline int fmt_uint(char* out, unsigned long long v) {
  char tmp[24];
  int n = 0;
  do { tmp[n++] = (char)('0' + v % 10); v /= 10; } while (v);
  for (int i = 0; i < n; ++i) out[i] = tmp[n - 1 - i];
  return n;
}
inline int fmt_fixed6(char* out, double x) {
  if (!(std::fabs(x) < 1.0e9)) return snprintf(out, 48, "%.6f", x);
  const bool neg = std::signbit(x);
  const double a = std::fabs(x);
  const double pr = a * 1.0e6;
  const double err = std::fma(a, 1.0e6, -pr);          // a * 1e6 == pr + err exactly
  double r = std::nearbyint(pr);                        // ties to even (default rounding mode)
  const double d = pr - r;                              // exact, in [-0.5, 0.5]
  if (d > 0) {
    const double c = 0.5 - d;
    if (err > c || (err == c && std::fmod(r, 2.0) != 0.0)) r += 1.0;
  } else if (d < 0) {
    const double c = 0.5 + d;
    if (-err > c || (-err == c && std::fmod(r, 2.0) != 0.0)) r -= 1.0;
  }
  // d == 0: |err| < 0.5, r stands.  d == +-0.5 (pr itself on a tie, already rounded to even) is covered above with c == 0.
  const unsigned long long u = (unsigned long long)r;
  int n = 0;
  if (neg) out[n++] = '-';
  n += fmt_uint(out + n, u / 1000000ull);
  out[n++] = '.';
  unsigned long long f = u % 1000000ull;
  for (int i = 5; i >= 0; --i) { out[n + i] = (char)('0' + f % 10); f /= 10; }
  return n + 6;
}

// Write a TREC run file from CSR results.  Query j (output order) = rows [offsets[qsel[j]], offsets[qsel[j]+1]) of
// ids/scores, already ranked; lines "{qid} Q0 {docno} {rank} {score:.6f} {tag}\n" (reference src/fusion.py:488-490;
// glibc's %.6f is correctly rounded, like Python's).
extern "C" int qf_runfile_write(const char* path, const char* qid_arena, const int64_t* qid_off, const int32_t* qsel,
                                int n_queries, const char* doc_arena, const int64_t* doc_off, const int64_t* offsets,
                                const int32_t* ids, const double* scores, const char* tag, int n_threads) {
  QF_REQUIRE(path && qid_arena && qid_off && qsel && doc_arena && doc_off && offsets && ids && scores && tag && n_queries >= 0,
             QF_E_INVALID, "qf_runfile_write: bad arguments");
  if (n_threads < 1) n_threads = 1;
  if (n_threads > n_queries) n_threads = n_queries > 0 ? n_queries : 1;
  std::vector<std::string> bufs((size_t)n_threads);
  std::vector<std::thread> th;
  const size_t tag_len = strlen(tag);
  for (int t = 0; t < n_threads; ++t)
    th.emplace_back([&, t]() {
      const int j0 = (int)((long long)n_queries * t / n_threads), j1 = (int)((long long)n_queries * (t + 1) / n_threads);
      std::string& out = bufs[(size_t)t];
      char num[96];
      {
        size_t lines = 0;
        for (int j = j0; j < j1; ++j) lines += (size_t)(offsets[qsel[j] + 1] - offsets[qsel[j]]);
        out.reserve(lines * (48 + tag_len));
      }
      for (int j = j0; j < j1; ++j) {
        const int q = qsel[j];
        const char* qs = qid_arena + qid_off[q];
        const size_t ql = (size_t)(qid_off[q + 1] - qid_off[q]);
        int rank = 1;
        for (int64_t e = offsets[q]; e < offsets[q + 1]; ++e, ++rank) {
          const int32_t d = ids[e];
          out.append(qs, ql);
          out.append(" Q0 ", 4);
          out.append(doc_arena + doc_off[d], (size_t)(doc_off[d + 1] - doc_off[d]));
          int nl = 0;
          num[nl++] = ' ';
          nl += fmt_uint(num + nl, (unsigned long long)rank);
          num[nl++] = ' ';
          nl += fmt_fixed6(num + nl, scores[e]);
          num[nl++] = ' ';
          out.append(num, (size_t)nl);
          out.append(tag, tag_len);
          out.push_back('\n');
        }
      }
    });
  for (auto& x : th) x.join();
  FILE* f = fopen(path, "wb");
  if (!f) { qf::set_error("qf_runfile_write: cannot open %s", path); return QF_E_INVALID; }
  for (auto& b : bufs)
    if (!b.empty() && fwrite(b.data(), 1, b.size(), f) != b.size()) { fclose(f); qf::set_error("qf_runfile_write: short write"); return QF_E_INVALID; }
  fclose(f);
  return QF_OK;
}
