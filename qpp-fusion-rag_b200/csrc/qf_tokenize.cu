// qf_wordpiece: BERT WordPiece tokenisation of query strings on the host, multi-threaded (no CUDA) -- the step in front
// of qf_bert_encode.  Replaces the HF tokenizer that SentenceTransformer.encode runs inside
// BGERetriever._encode_queries (reference src/retrievers/bge.py:160-170).
//
// Scope: text whose characters lie below U+3000 and are stable under NFC -- ASCII, Latin with accents, Greek, Cyrillic,
// general punctuation.  For those BertTokenizer's pipeline (clean -> [NFC] -> whitespace split -> lower-case -> NFD +
// strip Mn marks -> split punctuation -> greedy longest-match-first word pieces) is a per-character table lookup
// (csrc/qf_unicode_table.h, generated from Python's unicodedata by tools/gen_unicode_table.py): class word / space /
// removed / punctuation and the replacement text NFD(lower(c)) without marks.  A text holding anything else (combining
// marks, characters NFC rewrites, capital sigma whose lower case depends on context, CJK, malformed UTF-8) is NOT
// tokenised here: its flag is set and the host layer (qppfusion.encoder.WordPieceTokenizer) runs the full Unicode
// path for that text alone.  Words longer than 100 characters become [UNK].
#include <stdint.h>
#include <string.h>

#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "qf_common.cuh"
#include "qf_unicode_table.h"

struct qf_wordpiece {
  std::unordered_map<std::string, int32_t> start, cont;   // pieces that start a word / continue one (stored without "##")
  int32_t cls = -1, sep = -1, unk = -1, pad = -1;
  int lower = 1;
  size_t longest = 0;                                     // longest piece: bounds the match window
};

namespace {

// greedy longest-match-first pieces of one word; false = no segmentation ([UNK])
bool pieces_of(const qf_wordpiece& wp, const char* w, size_t n, std::string& key, std::vector<int32_t>& out) {
  const size_t mark = out.size();
  size_t s = 0;
  while (s < n) {
    const auto& map = s == 0 ? wp.start : wp.cont;
    size_t e = n - s > wp.longest ? s + wp.longest : n;
    int32_t found = -1;
    for (; e > s; --e) {
      key.assign(w + s, e - s);
      const auto it = map.find(key);
      if (it != map.end()) {
        found = it->second;
        break;
      }
    }
    if (found < 0) {
      out.resize(mark);
      return false;
    }
    out.push_back(found);
    s = e;
  }
  return true;
}

// Decode one UTF-8 sequence at t[j..n); returns the code point and advances j, or 0xFFFFFFFF for malformed input.
inline uint32_t utf8_next(const unsigned char* t, size_t n, size_t& j) {
  const unsigned char c = t[j];
  if (c < 0x80) { ++j; return c; }
  int len;
  uint32_t cp;
  if ((c & 0xE0) == 0xC0) { len = 2; cp = c & 0x1F; }
  else if ((c & 0xF0) == 0xE0) { len = 3; cp = c & 0x0F; }
  else if ((c & 0xF8) == 0xF0) { len = 4; cp = c & 0x07; }
  else return 0xFFFFFFFFu;
  if (j + len > n) return 0xFFFFFFFFu;
  for (int k = 1; k < len; ++k) {
    if ((t[j + k] & 0xC0) != 0x80) return 0xFFFFFFFFu;
    cp = (cp << 6) | (t[j + k] & 0x3F);
  }
  static const uint32_t min_cp[5] = {0, 0, 0x80, 0x800, 0x10000};
  if (cp < min_cp[len] || cp > 0x10FFFF || (cp >= 0xD800 && cp <= 0xDFFF)) return 0xFFFFFFFFu;   // overlong / surrogate
  j += len;
  return cp;
}

void encode_range(const qf_wordpiece& wp, const char* texts, const int64_t* off, int lo, int hi, int max_len,
                  int row_cap, int32_t* out_tokens, int32_t* out_lens, uint8_t* out_flags) {
  std::vector<int32_t> ids;
  std::string word, key;
  for (int i = lo; i < hi; ++i) {
    const unsigned char* t = reinterpret_cast<const unsigned char*>(texts + off[i]);
    const size_t n = (size_t)(off[i + 1] - off[i]);
    int32_t* row = out_tokens + (size_t)i * row_cap;
    ids.clear();
    word.clear();
    size_t word_chars = 0;                                   // characters (not bytes) of the current word
    const size_t budget = (size_t)max_len - 2;
    bool handled = true;
    auto flush = [&]() {
      if (word.empty()) return;
      if (word_chars > 100 || !pieces_of(wp, word.data(), word.size(), key, ids)) ids.push_back(wp.unk);
      word.clear();
      word_chars = 0;
    };
    // NOTE: the whole text is scanned even when the token budget is reached: an unhandled character anywhere sends the
    // text to the Python path, whose truncated result must not depend on where the scan stopped
    for (size_t j = 0; j < n;) {
      const uint32_t cp = utf8_next(t, n, j);
      if (cp >= qf_unicode::kLimit) {                        // includes malformed input (0xFFFFFFFF)
        handled = false;
        break;
      }
      const uint8_t kind = qf_unicode::kKind[cp];
      const int cls = kind & 15, mlen = kind >> 4;
      if (cls == 4) {
        handled = false;
        break;
      }
      if (ids.size() >= budget) continue;                    // only validating the rest of the text
      if (cls == 1) {
        flush();
      } else if (cls == 2) {
        // removed: does not separate words
      } else {
        const char* m = reinterpret_cast<const char*>(qf_unicode::kMapBlob + qf_unicode::kMapOff[cp]);
        if (cls == 3) {
          flush();
          word.assign(m, (size_t)mlen);
          word_chars = 1;
          flush();
        } else if (cp < 0x80) {
          word.push_back(wp.lower ? m[0] : (char)cp);
          ++word_chars;
        } else if (wp.lower) {
          word.append(m, (size_t)mlen);
          for (int k = 0; k < mlen; ++k) word_chars += ((unsigned char)m[k] & 0xC0) != 0x80;
        } else {                                             // cased model: characters kept as they are
          handled = false;                                   // (accent handling of cased checkpoints: Python path)
          break;
        }
      }
    }
    out_flags[i] = handled ? 0 : 1;
    if (!handled) {
      out_lens[i] = 0;
      continue;
    }
    if (ids.size() < budget) flush();
    const size_t m = ids.size() < budget ? ids.size() : budget;
    row[0] = wp.cls;
    memcpy(row + 1, ids.data(), m * sizeof(int32_t));
    row[m + 1] = wp.sep;
    for (size_t j = m + 2; j < (size_t)row_cap; ++j) row[j] = wp.pad;
    out_lens[i] = (int32_t)(m + 2);
  }
}

}  // namespace

extern "C" int qf_wordpiece_create(const char* vocab, size_t vocab_bytes, int do_lower_case, qf_wordpiece** out) {
  QF_REQUIRE(vocab && out, QF_E_INVALID, "qf_wordpiece_create: null argument");
  qf_wordpiece* wp = new qf_wordpiece();
  wp->lower = do_lower_case ? 1 : 0;
  int32_t id = 0;
  size_t s = 0;
  while (s < vocab_bytes) {
    size_t e = s;
    while (e < vocab_bytes && vocab[e] != '\n') ++e;
    size_t len = e - s;
    if (len && vocab[s + len - 1] == '\r') --len;
    std::string tok(vocab + s, len);
    if (tok == "[CLS]") wp->cls = id;
    if (tok == "[SEP]") wp->sep = id;
    if (tok == "[UNK]") wp->unk = id;
    if (tok == "[PAD]") wp->pad = id;
    if (len > 2 && tok[0] == '#' && tok[1] == '#') {
      wp->cont.emplace(tok.substr(2), id);                 // first occurrence wins, as a Python dict built line by line
      if (len - 2 > wp->longest) wp->longest = len - 2;    //   would keep the LAST: duplicates are rejected below
    } else {
      wp->start.emplace(tok, id);
      if (len > wp->longest) wp->longest = len;
    }
    ++id;
    s = e + 1;
  }
  if (wp->cls < 0 || wp->sep < 0 || wp->unk < 0 || wp->pad < 0) {
    delete wp;
    qf::set_error("qf_wordpiece_create: the vocabulary lacks [CLS], [SEP], [UNK] or [PAD]");
    return QF_E_INVALID;
  }
  if ((size_t)id != wp->start.size() + wp->cont.size()) {
    delete wp;
    qf::set_error("qf_wordpiece_create: duplicate token in the vocabulary");
    return QF_E_INVALID;
  }
  *out = wp;
  return QF_OK;
}

extern "C" void qf_wordpiece_free(qf_wordpiece* wp) { delete wp; }

extern "C" int qf_wordpiece_encode(const qf_wordpiece* wp, const char* texts, const int64_t* offsets, int n, int max_len,
                                   int row_cap, int n_threads, int32_t* out_tokens, int32_t* out_lens,
                                   uint8_t* out_flags) {
  QF_REQUIRE(wp && offsets && out_tokens && out_lens && out_flags && (texts || n == 0), QF_E_INVALID,
             "qf_wordpiece_encode: null argument");
  QF_REQUIRE(n >= 0 && max_len >= 2 && row_cap >= max_len, QF_E_INVALID,
             "qf_wordpiece_encode: bad sizes n=%d max_len=%d row_cap=%d", n, max_len, row_cap);
  if (n == 0) return QF_OK;
  int nt = n_threads > 0 ? n_threads : (int)std::thread::hardware_concurrency();
  nt = nt < 1 ? 1 : (nt > 64 ? 64 : nt);
  if (n < 256) nt = 1;
  if (nt == 1) {
    encode_range(*wp, texts, offsets, 0, n, max_len, row_cap, out_tokens, out_lens, out_flags);
    return QF_OK;
  }
  std::vector<std::thread> th;
  const int per = (n + nt - 1) / nt;
  for (int t = 0; t < nt; ++t) {
    const int lo = t * per, hi = lo + per < n ? lo + per : n;
    if (lo >= hi) break;
    th.emplace_back(encode_range, std::cref(*wp), texts, offsets, lo, hi, max_len, row_cap, out_tokens, out_lens,
                    out_flags);
  }
  for (auto& x : th) x.join();
  return QF_OK;
}
