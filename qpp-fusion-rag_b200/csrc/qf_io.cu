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
#include <utility>
#include <vector>
#include <chrono>

#include "qf_common.cuh"
static double now_ms(){return std::chrono::duration<double,std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();}
static bool g_trace = getenv("QF_IO_TRACE")!=nullptr;
#define TR(label) do{ if(g_trace){ double t_=now_ms(); fprintf(stderr,"  [io] %-18s %.1f ms\n", label, t_-t_last); t_last=t_; } }while(0)


namespace {

struct Tok {
  const char* p;
  int len;
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

// Large uninitialised buffers: anonymous mappings with transparent huge pages requested, so that the first touch
// (by the parser threads, spread over all cores) costs one fault per 2 MB instead of one per 4 KB -- on a cold call the
// 4 KB faults of the column arrays used to cost more than the parsing itself.
template <typename T>
struct Big {
  T* p = nullptr;
  size_t n = 0, bytes = 0;
  bool mapped = false;
  Big() = default;
  Big(const Big&) = delete;
  Big& operator=(const Big&) = delete;
  Big(Big&& o) noexcept { *this = std::move(o); }
  Big& operator=(Big&& o) noexcept {
    if (this != &o) {
      release();
      p = o.p; n = o.n; bytes = o.bytes; mapped = o.mapped;
      o.p = nullptr; o.n = o.bytes = 0;
    }
    return *this;
  }
  ~Big() { release(); }
  void release() {
    if (!p) return;
    if (mapped) munmap(p, bytes); else free(p);
    p = nullptr;
    n = bytes = 0;
  }
  bool alloc(size_t count) {
    release();
    n = count;
    bytes = count * sizeof(T);
    if (bytes == 0) { bytes = sizeof(T); }
    if (bytes >= (size_t(4) << 20)) {
      bytes = (bytes + (size_t(2) << 20) - 1) & ~((size_t(2) << 20) - 1);
      void* m = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
      if (m == MAP_FAILED) { p = nullptr; return false; }
#ifdef MADV_HUGEPAGE
      if (!getenv("QF_IO_NOHUGE")) madvise(m, bytes, MADV_HUGEPAGE);
#endif
      p = static_cast<T*>(m);
      mapped = true;
    } else {
      p = static_cast<T*>(malloc(bytes));
      mapped = false;
    }
    return p != nullptr;
  }
  T* data() { return p; }
  const T* data() const { return p; }
  size_t size() const { return n; }
  T& operator[](size_t i) { return p[i]; }
  const T& operator[](size_t i) const { return p[i]; }
};

struct Dict {
  // open addressing; a slot holds (entry, upper hash half), so a probe touches the table line and, on a tag match,
  // the arena -- not a third array.  The arena stores the strings back to back.
  struct Slot { int32_t e; uint32_t tag; };
  std::vector<uint64_t> hashes;
  std::vector<Slot> slots;
  std::vector<int64_t> off;     // entry -> arena offset (off.size() == n+1)
  std::string arena;
  size_t mask = 0;
  Dict() { rehash(1 << 10); off.push_back(0); }
  static uint64_t hash(const char* p, int len) {
    uint64_t h = 1469598103934665603ull;
    for (int i = 0; i < len; ++i) { h ^= (unsigned char)p[i]; h *= 1099511628211ull; }
    return h ^ (h >> 29);
  }
  void rehash(size_t cap) {
    std::vector<Slot> ns(cap, Slot{-1, 0u});
    const size_t nm = cap - 1;
    for (size_t e = 0; e < hashes.size(); ++e) {
      size_t i = hashes[e] & nm;
      while (ns[i].e >= 0) i = (i + 1) & nm;
      ns[i] = Slot{(int32_t)e, (uint32_t)(hashes[e] >> 32)};
    }
    slots.swap(ns);
    mask = nm;
  }
  void prefetch(uint64_t h) const { __builtin_prefetch(&slots[h & mask]); }
  int32_t get(const char* p, int len) { return get_hashed(p, len, hash(p, len)); }
  int32_t get_hashed(const char* p, int len, uint64_t h) {
    size_t i = h & mask;
    const uint32_t tag = (uint32_t)(h >> 32);
    while (true) {
      const Slot s = slots[i];
      if (s.e < 0) break;
      if (s.tag == tag && off[s.e + 1] - off[s.e] == len && memcmp(arena.data() + off[s.e], p, (size_t)len) == 0) return s.e;
      i = (i + 1) & mask;
    }
    const int32_t e = (int32_t)hashes.size();
    hashes.push_back(h);
    arena.append(p, (size_t)len);
    off.push_back((int64_t)arena.size());
    slots[i] = Slot{e, tag};
    if (hashes.size() * 2 > slots.size()) rehash(slots.size() * 2);
    return e;
  }
  int32_t size() const { return (int32_t)hashes.size(); }
};

struct FileRows {
  Big<int32_t> qid, doc, rank, micro;
  Big<double> score;
  size_t n = 0;
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

struct QStart {            // first row (chunk-local index) of a run of rows with the same qid token
  uint32_t row;
  Tok tok;
  int32_t code;
};

// Per-file transient row data the docno phase needs (the numeric columns are written straight into FileRows).
struct DocTok {
  const char* p;
  uint64_t hash;
  int32_t len;
};

struct Chunk {
  const char *b, *e;
  size_t cap = 0, rows = 0, off = 0;          // line count bound, parsed rows, first row in the file's columns
  std::vector<QStart> qstarts;
  std::vector<std::vector<uint32_t>> owned;   // owned[t]: chunk-local rows whose docno shard belongs to thread t
  bool micro_bad = false;
  std::string err;
};

size_t count_lines(const char* b, const char* e) {
  size_t n = 0;
  const char* p = b;
  while (p < e) {
    const char* nl = (const char*)memchr(p, '\n', (size_t)(e - p));
    ++n;
    if (!nl) break;
    p = nl + 1;
  }
  return n;
}

// Tokenise + convert one chunk straight into the file's columns at [c.off, c.off + c.cap).
bool parse_chunk(Chunk& c, FileRows& F, DocTok* dt, int nt) {
  c.owned.assign((size_t)nt, std::vector<uint32_t>());
  for (auto& o : c.owned) o.reserve(c.cap / (size_t)nt + c.cap / (size_t)(8 * nt) + 16);
  int32_t* rank = F.rank.data() + c.off;
  int32_t* micro = F.micro.data() + c.off;
  double* score = F.score.data() + c.off;
  DocTok* d = dt + c.off;
  Tok last_q{nullptr, 0};
  size_t n = 0;
  const char* p = c.b;
  while (p < c.e) {
    const char* eol = (const char*)memchr(p, '\n', (size_t)(c.e - p));
    if (!eol) eol = c.e;
    Tok f[6];
    int nf = 0;
    const char* s = p;
    while (s < eol && nf < 6) {
      while (s < eol && is_ws(*s)) ++s;
      if (s >= eol) break;
      const char* t0 = s;
      while (s < eol && !is_ws(*s)) ++s;
      f[nf].p = t0;
      f[nf].len = (int)(s - t0);
      ++nf;
    }
    if (nf >= 5) {
      if (!last_q.p || f[0].len != last_q.len || memcmp(f[0].p, last_q.p, (size_t)f[0].len) != 0) {
        c.qstarts.push_back(QStart{(uint32_t)n, f[0], -1});
        last_q = f[0];
      }
      const uint64_t h = Dict::hash(f[2].p, f[2].len);
      d[n] = DocTok{f[2].p, h, f[2].len};
      char buf[64];
      char* endp = nullptr;
      if (!parse_int(f[3].p, f[3].len, &rank[n])) {
        const int rl = f[3].len < 63 ? f[3].len : 63;
        memcpy(buf, f[3].p, (size_t)rl);
        buf[rl] = 0;
        const long rk = strtol(buf, &endp, 10);
        if (*endp != 0 || rk < -2147483647L || rk > 2147483647L) { c.err = "bad rank field '" + std::string(buf) + "'"; return false; }
        rank[n] = (int32_t)rk;
      }
      int32_t mi = 0;
      if (parse_micro(f[4].p, f[4].len, &mi)) {
        // <= 6 decimals: the value is micro / 10^6 exactly, and the IEEE quotient of two exactly representable
        // numbers is the double nearest to it -- the same double strtod returns (sign of a zero kept)
        micro[n] = mi;
        score[n] = (mi == 0 && f[4].p[0] == '-') ? -0.0 : (double)mi / 1.0e6;
      } else {
        c.micro_bad = true;
        micro[n] = 0;
        const int sl = f[4].len < 63 ? f[4].len : 63;
        memcpy(buf, f[4].p, (size_t)sl);
        buf[sl] = 0;
        score[n] = strtod(buf, &endp);
        if (*endp != 0) { c.err = "bad score field '" + std::string(buf) + "'"; return false; }
      }
      c.owned[(size_t)(doc_shard(h) % nt)].push_back((uint32_t)n);
      ++n;
    } else if (nf != 0) {
      // a short line is rejected, not padded with NaN as pandas would: a run file with a truncated line is broken
      c.err = "line with " + std::to_string(nf) + " fields (need >= 5)";
      return false;
    }
    p = eol + 1;
  }
  c.rows = n;
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
  double t_last = now_ms();
  qf_runs* R = new qf_runs();
  R->files.resize((size_t)n_files);
  TR("ctor");
  auto fail = [&](int code, const char* fmt, const char* a, const char* b) {
    qf::set_error(fmt, a, b);
    delete R;
    return code;
  };
  for (int fi = 0; fi < n_files; ++fi) {
    const int fd = open(paths[fi], O_RDONLY);
    if (fd < 0) return fail(QF_E_INVALID, "qf_runs_parse: cannot open %s%s", paths[fi], "");
    struct stat st;
    if (fstat(fd, &st) != 0) { close(fd); return fail(QF_E_INVALID, "qf_runs_parse: cannot stat %s%s", paths[fi], ""); }
    const size_t sz = (size_t)st.st_size;
    const char* base = sz ? (const char*)mmap(nullptr, sz, PROT_READ, MAP_PRIVATE | MAP_POPULATE, fd, 0) : nullptr;
    if (sz && base == MAP_FAILED) { close(fd); return fail(QF_E_INVALID, "qf_runs_parse: mmap failed for %s%s", paths[fi], ""); }
    auto unmap = [&]() { if (sz) munmap((void*)base, sz); close(fd); };
    // 1. split at newline boundaries; count the lines of every chunk (bound on its rows) so that the columns can be
    //    allocated once and written in place
    const int nt = sz < (1u << 20) ? 1 : n_threads;
    std::vector<Chunk> chunks((size_t)nt);
    {
      std::vector<const char*> cut((size_t)nt + 1);
      cut[0] = base;
      cut[(size_t)nt] = base + sz;
      for (int t = 1; t < nt; ++t) {
        const char* p = base + sz * (size_t)t / (size_t)nt;
        const char* nl = (const char*)memchr(p, '\n', (size_t)(base + sz - p));
        cut[(size_t)t] = nl ? nl + 1 : base + sz;
      }
      for (int t = 0; t < nt; ++t) { chunks[(size_t)t].b = cut[(size_t)t]; chunks[(size_t)t].e = cut[(size_t)t + 1]; }
    }
    parallel_for(nt, [&](int t) { chunks[(size_t)t].cap = count_lines(chunks[(size_t)t].b, chunks[(size_t)t].e); });
    TR("mmap+count");
    size_t cap = 0;
    for (auto& c : chunks) { c.off = cap; cap += c.cap; }
    FileRows& F = R->files[(size_t)fi];
    Big<DocTok> dt;
    if (!F.qid.alloc(cap) || !F.doc.alloc(cap) || !F.rank.alloc(cap) || !F.micro.alloc(cap) || !F.score.alloc(cap) || !dt.alloc(cap)) {
      unmap();
      return fail(QF_E_INVALID, "qf_runs_parse: out of memory for %s%s", paths[fi], "");
    }
    TR("alloc");
    // 2. tokenise + convert numbers in parallel, straight into the columns
    std::vector<char> ok((size_t)nt, 1);
    parallel_for(nt, [&](int t) { ok[(size_t)t] = parse_chunk(chunks[(size_t)t], F, dt.data(), nt) ? 1 : 0; });
    for (int t = 0; t < nt; ++t)
      if (!ok[(size_t)t]) {
        const std::string e = chunks[(size_t)t].err;
        unmap();
        return fail(QF_E_INVALID, "qf_runs_parse: %s: %s", paths[fi], e.c_str());
      }
    TR("tokenise");
    // blank lines leave gaps between the chunks' rows: close them (rare; the common file has none)
    size_t n = 0;
    for (auto& c : chunks) {
      if (c.off != n) {
        memmove(F.rank.data() + n, F.rank.data() + c.off, c.rows * 4);
        memmove(F.micro.data() + n, F.micro.data() + c.off, c.rows * 4);
        memmove(F.score.data() + n, F.score.data() + c.off, c.rows * 8);
        memmove(dt.data() + n, dt.data() + c.off, c.rows * sizeof(DocTok));
        c.off = n;
      }
      n += c.rows;
      if (c.micro_bad) R->all_micro = false;
    }
    F.n = n;
    // 3. qid codes: only the first row of every run of equal qids goes through the dictionary (serial, in file
    //    order, so codes are first-seen ordered); the runs are filled in parallel
    for (auto& c : chunks)
      for (auto& q : c.qstarts) q.code = R->qids.get(q.tok.p, q.tok.len);
    TR("qid serial");
    // 4. qid fill + docno codes (shard-parallel: thread t owns the dictionary shards s with s % nt == t and visits
    //    the rows in file order, so codes are deterministic)
    parallel_for(nt, [&](int t) {
      Chunk& c = chunks[(size_t)t];
      int32_t* qd = F.qid.data() + c.off;
      for (size_t k = 0; k < c.qstarts.size(); ++k) {
        const size_t r0 = c.qstarts[k].row, r1 = k + 1 < c.qstarts.size() ? c.qstarts[k + 1].row : c.rows;
        const int32_t code = c.qstarts[k].code;
        for (size_t r = r0; r < r1; ++r) qd[r] = code;
      }
      for (int pt = 0; pt < nt; ++pt) {
        const Chunk& pc = chunks[(size_t)pt];
        const std::vector<uint32_t>& idx = pc.owned[(size_t)t];
        const DocTok* v = dt.data() + pc.off;
        int32_t* dst = F.doc.data() + pc.off;
        for (size_t k = 0; k < idx.size(); ++k) {
          const DocTok& r = v[idx[k]];
          const int s_ = doc_shard(r.hash);
          if (k + 8 < idx.size()) {
            const DocTok& nx = v[idx[k + 8]];
            R->doc_shards[(size_t)doc_shard(nx.hash)].prefetch(nx.hash);
          }
          const int32_t code = R->doc_shards[(size_t)s_].get_hashed(r.p, r.len, r.hash);
          dst[idx[k]] = (int32_t)(((uint32_t)s_ << LOCAL_BITS) | (uint32_t)code);
        }
      }
    });
    TR("docno dict");
    unmap();
    TR("unmap");
  }
  // 5. shard bases, merged dictionary, final codes
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
    std::vector<size_t> at((size_t)DOC_SHARDS + 1, 0);
    for (int s_ = 0; s_ < DOC_SHARDS; ++s_) at[(size_t)s_ + 1] = at[(size_t)s_] + R->doc_shards[(size_t)s_].arena.size();
    R->doc_arena.resize(at[DOC_SHARDS]);
    R->doc_off.resize((size_t)R->shard_base[DOC_SHARDS] + 1);
    R->doc_off[0] = 0;
    parallel_for(n_threads, [&](int t) {
      for (int s_ = t; s_ < DOC_SHARDS; s_ += n_threads) {
        Dict& d = R->doc_shards[(size_t)s_];
        memcpy(&R->doc_arena[at[(size_t)s_]], d.arena.data(), d.arena.size());
        int64_t* o = R->doc_off.data() + R->shard_base[(size_t)s_];
        for (size_t e = 1; e < d.off.size(); ++e) o[e] = (int64_t)at[(size_t)s_] + d.off[e];
        std::vector<uint64_t>().swap(d.hashes);          // the shard tables are no longer needed
        std::vector<Dict::Slot>().swap(d.slots);
        std::string().swap(d.arena);
      }
    });
  }
  for (auto& F : R->files) {
    const size_t n = F.n;
    const int nt = n < (1u << 16) ? 1 : n_threads;
    parallel_for(nt, [&](int t) {
      const size_t i0 = n * (size_t)t / (size_t)nt, i1 = n * (size_t)(t + 1) / (size_t)nt;
      for (size_t i = i0; i < i1; ++i) {
        const uint32_t v = (uint32_t)F.doc[i];
        F.doc[i] = R->shard_base[v >> LOCAL_BITS] + (int32_t)(v & ((1u << LOCAL_BITS) - 1u));
      }
    });
  }
  TR("merge+recode");
  *out = R;
  return QF_OK;
}

extern "C" void qf_runs_free(qf_runs* r) { delete r; }

extern "C" int qf_runs_info(const qf_runs* r, int64_t* rows_per_file, int32_t* n_qids, int32_t* n_docs, int32_t* all_micro,
                            int64_t* qid_arena_bytes, int64_t* doc_arena_bytes) {
  QF_REQUIRE(r, QF_E_INVALID, "qf_runs_info: null handle");
  if (rows_per_file)
    for (size_t i = 0; i < r->files.size(); ++i) rows_per_file[i] = (int64_t)r->files[i].n;
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
  const size_t n = F.n;
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
  if (!(std::fabs(x) < 1.0e9)) return snprintf(out, 400, "%.6f", x);   // (the writer reserves 400 bytes for each)
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
  const size_t tag_len = strlen(tag);
  // Every thread formats a contiguous range of queries into its own buffer (sized by an upper bound per line; the
  // untouched tail of the anonymous mapping costs nothing) and then writes it at its final file offset with pwrite,
  // so neither the formatting nor the copy into the page cache is serial.
  struct Part { Big<char> buf; size_t used = 0; int bad = 0; };
  std::vector<Part> parts((size_t)n_threads);
  auto range = [&](int t, int& j0, int& j1) {
    j0 = (int)((long long)n_queries * t / n_threads);
    j1 = (int)((long long)n_queries * (t + 1) / n_threads);
  };
  parallel_for(n_threads, [&](int t) {
    int j0, j1;
    range(t, j0, j1);
    Part& P = parts[(size_t)t];
    size_t bound = 0;
    for (int j = j0; j < j1; ++j) {
      const int q = qsel[j];
      const size_t ql = (size_t)(qid_off[q + 1] - qid_off[q]);
      size_t docs = 0;
      for (int64_t e = offsets[q]; e < offsets[q + 1]; ++e) {
        docs += (size_t)(doc_off[ids[e] + 1] - doc_off[ids[e]]);
        if (!(std::fabs(scores[e]) < 1.0e9)) docs += 400;       // "%.6f" of a huge double is up to ~320 characters
      }
      bound += docs + (size_t)(offsets[q + 1] - offsets[q]) * (ql + 4 + 1 + 11 + 1 + 48 + 1 + tag_len + 1);
    }
    if (!P.buf.alloc(bound + 64)) { P.bad = 1; return; }
    char* o = P.buf.data();
    for (int j = j0; j < j1; ++j) {
      const int q = qsel[j];
      const char* qs = qid_arena + qid_off[q];
      const size_t ql = (size_t)(qid_off[q + 1] - qid_off[q]);
      unsigned long long rank = 1;
      for (int64_t e = offsets[q]; e < offsets[q + 1]; ++e, ++rank) {
        const int32_t d = ids[e];
        if (e + 4 < offsets[q + 1]) __builtin_prefetch(doc_arena + doc_off[ids[e + 4]]);
        memcpy(o, qs, ql); o += ql;
        memcpy(o, " Q0 ", 4); o += 4;
        const size_t dl = (size_t)(doc_off[d + 1] - doc_off[d]);
        memcpy(o, doc_arena + doc_off[d], dl); o += dl;
        *o++ = ' ';
        o += fmt_uint(o, rank);
        *o++ = ' ';
        o += fmt_fixed6(o, scores[e]);
        *o++ = ' ';
        memcpy(o, tag, tag_len); o += tag_len;
        *o++ = '\n';
      }
    }
    P.used = (size_t)(o - P.buf.data());
  });
  for (auto& P : parts) QF_REQUIRE(!P.bad, QF_E_INVALID, "qf_runfile_write: out of memory");
  const int fd = open(path, O_WRONLY | O_CREAT | O_TRUNC, 0644);
  if (fd < 0) { qf::set_error("qf_runfile_write: cannot open %s", path); return QF_E_INVALID; }
  std::vector<size_t> at((size_t)n_threads + 1, 0);
  for (int t = 0; t < n_threads; ++t) at[(size_t)t + 1] = at[(size_t)t] + parts[(size_t)t].used;
  if (at[(size_t)n_threads] && ftruncate(fd, (off_t)at[(size_t)n_threads]) != 0) {
    close(fd);
    qf::set_error("qf_runfile_write: cannot size %s", path);
    return QF_E_INVALID;
  }
  std::vector<int> short_write((size_t)n_threads, 0);
  parallel_for(n_threads, [&](int t) {
    const Part& P = parts[(size_t)t];
    size_t done = 0;
    while (done < P.used) {
      const ssize_t w = pwrite(fd, P.buf.data() + done, P.used - done, (off_t)(at[(size_t)t] + done));
      if (w <= 0) { short_write[(size_t)t] = 1; return; }
      done += (size_t)w;
    }
  });
  close(fd);
  for (int t = 0; t < n_threads; ++t)
    if (short_write[(size_t)t]) { qf::set_error("qf_runfile_write: short write to %s", path); return QF_E_INVALID; }
  return QF_OK;
}
