"""
Query encoder on the device -- the step right before the dense search.

Reference: BGERetriever._load_encoder / _encode_queries (src/retrievers/bge.py:153-170):
    SentenceTransformer("BAAI/bge-base-en-v1.5").encode(queries, batch_size=..., normalize_embeddings=True)
i.e. BERT WordPiece tokenisation (lower-cased, [CLS] ... [SEP], truncated to 512 tokens) -> HF BertModel ->
CLS pooling -> L2 normalisation -> float32 [n, 768].  No instruction prefix is added to the queries (:160-169).

Here: tokenisation on the host (WordPieceTokenizer below, a plain-Python restatement of HF's BertTokenizer),
everything else in ONE C-ABI call, qf_bert_encode (csrc/qf_bert.cu), on bf16 weights packed into one device blob.
`BertQueryEncoder.__call__` is the drop-in for `_encode_queries` (List[str] -> float32 [n, 768]); `encode_device`
returns the bf16 rows that feed qf_dense_topk directly, so a retrieve_batch never moves embeddings through the host.
There is no CPU fallback: without the CUDA library or an sm_100 device every call raises.
"""
from __future__ import annotations

import json
import unicodedata
from dataclasses import dataclass
from pathlib import Path
from typing import Dict, List, Mapping, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _native as N
from . import ops


@dataclass(frozen=True)
class BertConfig:
    """The fields of an HF BertConfig the encoder needs (defaults: bert-base / bge-base-en-v1.5)."""
    n_layers: int = 12
    hidden: int = 768
    heads: int = 12
    ffn: int = 3072
    vocab: int = 30522
    max_pos: int = 512
    ln_eps: float = 1e-12
    pooling: str = "cls"              # "cls" (bge) or "mean" (other sentence-transformers checkpoints)

    @classmethod
    def from_hf(cls, cfg: Mapping, pooling: str = "cls") -> "BertConfig":
        if cfg.get("hidden_act", "gelu") != "gelu":
            raise ValueError(f"hidden_act={cfg.get('hidden_act')!r}: only the exact erf GELU is implemented")
        if cfg.get("position_embedding_type", "absolute") != "absolute":
            raise ValueError("only absolute position embeddings are implemented")
        return cls(n_layers=int(cfg["num_hidden_layers"]), hidden=int(cfg["hidden_size"]),
                   heads=int(cfg["num_attention_heads"]), ffn=int(cfg["intermediate_size"]),
                   vocab=int(cfg["vocab_size"]), max_pos=int(cfg["max_position_embeddings"]),
                   ln_eps=float(cfg.get("layer_norm_eps", 1e-12)), pooling=pooling)

    @property
    def pooling_id(self) -> int:
        try:
            return {"cls": N.BERT_POOL_CLS, "mean": N.BERT_POOL_MEAN}[self.pooling]
        except KeyError:
            raise ValueError(f"unknown pooling {self.pooling!r}") from None


def pack_bert_weights(state_dict: Mapping[str, torch.Tensor], cfg: BertConfig) -> torch.Tensor:
    """HF BertModel state dict (with or without a 'bert.' / '0.auto_model.' prefix) -> the packed bf16 blob
    qf_bert_encode reads (layout: include/qppfusion.h).  Raises KeyError for a missing tensor, ValueError for a
    wrong shape -- never fills anything in."""
    keys = list(state_dict.keys())
    probe = "embeddings.word_embeddings.weight"
    hit = [k for k in keys if k.endswith(probe)]
    if not hit:
        raise KeyError(f"no '{probe}' in the state dict")
    prefix = hit[0][:-len(probe)]

    def get(name: str, shape: Tuple[int, ...]) -> torch.Tensor:
        t = state_dict[prefix + name]
        if tuple(t.shape) != shape:
            raise ValueError(f"{name}: shape {tuple(t.shape)}, expected {shape}")
        return t.detach().to(torch.float32).reshape(-1)

    H, F = cfg.hidden, cfg.ffn
    parts = [get("embeddings.word_embeddings.weight", (cfg.vocab, H)),
             get("embeddings.position_embeddings.weight", (cfg.max_pos, H)),
             get("embeddings.token_type_embeddings.weight", (2, H)),
             get("embeddings.LayerNorm.weight", (H,)), get("embeddings.LayerNorm.bias", (H,))]
    for l in range(cfg.n_layers):
        p = f"encoder.layer.{l}."
        parts += [get(p + f"attention.self.{n}.weight", (H, H)) for n in ("query", "key", "value")]
        parts += [get(p + f"attention.self.{n}.bias", (H,)) for n in ("query", "key", "value")]
        parts += [get(p + "attention.output.dense.weight", (H, H)), get(p + "attention.output.dense.bias", (H,)),
                  get(p + "attention.output.LayerNorm.weight", (H,)), get(p + "attention.output.LayerNorm.bias", (H,)),
                  get(p + "intermediate.dense.weight", (F, H)), get(p + "intermediate.dense.bias", (F,)),
                  get(p + "output.dense.weight", (H, F)), get(p + "output.dense.bias", (H,)),
                  get(p + "output.LayerNorm.weight", (H,)), get(p + "output.LayerNorm.bias", (H,))]
    return torch.cat(parts).to(torch.bfloat16)


# ------------------------------------------------------------------------------------------------ tokeniser
def _is_punct(ch: str) -> bool:
    cp = ord(ch)
    if 33 <= cp <= 47 or 58 <= cp <= 64 or 91 <= cp <= 96 or 123 <= cp <= 126:
        return True
    return unicodedata.category(ch).startswith("P")


def _is_cjk(cp: int) -> bool:
    return (0x4E00 <= cp <= 0x9FFF or 0x3400 <= cp <= 0x4DBF or 0x20000 <= cp <= 0x2A6DF or 0x2A700 <= cp <= 0x2B73F
            or 0x2B740 <= cp <= 0x2B81F or 0x2B820 <= cp <= 0x2CEAF or 0xF900 <= cp <= 0xFAFF or 0x2F800 <= cp <= 0x2FA1F)


class WordPieceTokenizer:
    """BERT tokenisation as HF's BertTokenizer does it (BasicTokenizer + WordpieceTokenizer): drop control
    characters, CJK characters as words of their own, NFC, split on whitespace, lower-case + strip accents
    (do_lower_case), split punctuation off, then greedy longest-match-first word pieces ('##' continuation,
    words longer than 100 characters -> [UNK])."""

    def __init__(self, vocab, do_lower_case: bool = True, max_length: int = 512):
        if isinstance(vocab, (str, Path)):
            with open(vocab, "r", encoding="utf-8") as f:
                vocab = {line.rstrip("\n"): i for i, line in enumerate(f)}
        self.vocab: Dict[str, int] = dict(vocab)
        self.do_lower_case = do_lower_case
        self.max_length = int(max_length)
        try:
            self.cls_id, self.sep_id, self.unk_id = self.vocab["[CLS]"], self.vocab["[SEP]"], self.vocab["[UNK]"]
            self.pad_id = self.vocab["[PAD]"]
        except KeyError as e:
            raise ValueError(f"vocabulary lacks the special token {e}") from None

    def words(self, text: str) -> List[str]:
        out = []
        for ch in text:
            cp = ord(ch)
            if ch in "\t\n\r" or ch == " " or unicodedata.category(ch) == "Zs":
                out.append(" ")
            elif cp == 0 or cp == 0xFFFD or unicodedata.category(ch).startswith("C"):
                continue
            elif _is_cjk(cp):
                out.append(f" {ch} ")
            else:
                out.append(ch)
        words: List[str] = []
        for tok in unicodedata.normalize("NFC", "".join(out)).split():
            if self.do_lower_case:
                tok = tok.lower()
                tok = "".join(c for c in unicodedata.normalize("NFD", tok) if unicodedata.category(c) != "Mn")
            cur = ""
            for ch in tok:
                if _is_punct(ch):
                    if cur:
                        words.append(cur)
                    words.append(ch)
                    cur = ""
                else:
                    cur += ch
            if cur:
                words.append(cur)
        return words

    def pieces(self, word: str) -> List[int]:
        if len(word) > 100:
            return [self.unk_id]
        start, ids = 0, []
        while start < len(word):
            end, found = len(word), None
            while start < end:
                sub = ("##" if start > 0 else "") + word[start:end]
                found = self.vocab.get(sub)
                if found is not None:
                    break
                end -= 1
            if found is None:
                return [self.unk_id]
            ids.append(found)
            start = end
        return ids

    def encode(self, text: str) -> List[int]:
        """[CLS] pieces [SEP], truncated to max_length tokens (the pieces are cut, the specials stay)."""
        ids = [i for w in self.words(text) for i in self.pieces(w)][: self.max_length - 2]
        return [self.cls_id] + ids + [self.sep_id]

    # -- native fast path -------------------------------------------------------------------------------------
    def _native_handle(self):
        h = getattr(self, "_wp", None)
        if h is None:
            import ctypes
            by_id = [None] * (max(self.vocab.values()) + 1)
            for t, i in self.vocab.items():
                if i < 0 or by_id[i] is not None:
                    raise ValueError("vocabulary ids must be distinct and non-negative")
                by_id[i] = t
            # ids no token maps to (a vocab.txt line shadowed by a later duplicate): a placeholder no word can match
            blob = "\n".join(t if t is not None else f"\x00unused{i}" for i, t in enumerate(by_id)).encode("utf-8")
            h = ctypes.c_void_p()
            N.check(N.lib().qf_wordpiece_create(blob, len(blob), 1 if self.do_lower_case else 0, ctypes.byref(h)))
            self._wp = h
        return h

    def __del__(self):
        h = getattr(self, "_wp", None)
        if h is not None:
            try:
                N.lib().qf_wordpiece_free(h)
            except Exception:
                pass

    def encode_batch(self, texts: Sequence[str], n_threads: int = 0) -> List[np.ndarray]:
        """[CLS] ... [SEP] id rows (int32 arrays) for many texts: texts in the native scope (below U+3000, NFC-stable) through the multi-threaded C++
        tokeniser (qf_wordpiece_encode), the rest through `encode` (full Unicode handling) one by one."""
        n = len(texts)
        if n == 0:
            return []
        toks, lens, flags = self._native_encode(texts, n_threads)
        rows = [toks[i, :lens[i]] for i in range(n)]
        for i in np.flatnonzero(flags):
            rows[i] = np.asarray(self.encode(texts[i]), dtype=np.int32)
        return rows

    def _native_encode(self, texts: Sequence[str], n_threads: int = 0):
        """qf_wordpiece_encode over all texts -> (tokens int32 [n, cap] padded, lens int32 [n], flags uint8 [n];
        flag 1 = a byte >= 0x80, text not tokenised)."""
        n = len(texts)
        enc = [t.encode("utf-8") for t in texts]
        offs = np.zeros(n + 1, dtype=np.int64)
        np.cumsum([len(b) for b in enc], out=offs[1:])
        blob = b"".join(enc)
        cap = int(min(self.max_length, max(len(b) for b in enc) + 2))
        cap = max(cap, 2)
        toks = np.empty((n, cap), dtype=np.int32)
        lens = np.empty(n, dtype=np.int32)
        flags = np.empty(n, dtype=np.uint8)
        N.check(N.lib().qf_wordpiece_encode(self._native_handle(), blob, offs.ctypes.data, n, cap, cap, int(n_threads),
                                            toks.ctypes.data, lens.ctypes.data, flags.ctypes.data))
        return toks, lens, flags

    def encode_batch_packed(self, texts: Sequence[str], n_threads: int = 0) -> Tuple[np.ndarray, np.ndarray]:
        """-> (flat int32 token ids of all texts back to back, lens int64 [n]): the packed layout qf_bert_encode_packed
        takes, built without a Python object per text (non-ASCII texts go through `encode` one by one)."""
        n = len(texts)
        if n == 0:
            return np.zeros(0, dtype=np.int32), np.zeros(0, dtype=np.int64)
        toks, lens, flags = self._native_encode(texts, n_threads)
        slow = np.flatnonzero(flags)
        if slow.size == 0:
            keep = np.arange(toks.shape[1])[None, :] < lens[:, None]
            return toks[keep], lens.astype(np.int64)
        rows = [toks[i, :lens[i]] for i in range(n)]
        for i in slow:
            rows[i] = np.asarray(self.encode(texts[i]), dtype=np.int32)
        return np.concatenate(rows), np.fromiter((len(r) for r in rows), dtype=np.int64, count=n)

    def batch(self, texts: Sequence[str], pad_multiple: int = 8) -> Tuple[np.ndarray, np.ndarray]:
        """-> (tokens int32 [n, L] padded with [PAD], lens int32 [n]); L = longest row rounded up to pad_multiple."""
        rows = [self.encode(t) for t in texts]
        lens = np.array([len(r) for r in rows], dtype=np.int32)
        L = int(lens.max()) if len(rows) else 1
        L = min(-(-L // pad_multiple) * pad_multiple, max(self.max_length, L))
        toks = np.full((len(rows), L), self.pad_id, dtype=np.int32)
        for i, r in enumerate(rows):
            toks[i, : len(r)] = r
        return toks, lens


def bucket_token_rows(rows: Sequence[Sequence[int]], bucket: int, max_len: int, batch_size: int, pad_id: int = 0,
                      max_tokens: int = 1 << 18):
    """Group token rows by their length rounded up to `bucket` (host side, vectorised): yields
    (row indices int64 [m], tokens int32 [m, L] padded with pad_id, lens int32 [m]) per group of at most batch_size
    rows and at most max_tokens padded tokens (the encoder's workspace is ~15 kB per token for bert-base)."""
    n = len(rows)
    lens = np.fromiter((len(r) for r in rows), dtype=np.int64, count=n)
    if n and (lens.min() < 1 or lens.max() > max_len):
        raise ValueError(f"row lengths must lie in [1, {max_len}]")
    if n and isinstance(rows[0], np.ndarray):
        flat = np.concatenate(rows).astype(np.int64, copy=False)
    else:
        import itertools
        flat = np.fromiter(itertools.chain.from_iterable(rows), dtype=np.int64, count=int(lens.sum()))
    if flat.size and (flat.min() < -2**31 or flat.max() >= 2**31):
        raise ValueError("token id does not fit int32")
    offs = np.concatenate(([0], np.cumsum(lens)))
    padded = np.minimum(-(-lens // bucket) * bucket, max_len)
    order = np.argsort(padded, kind="stable")
    bounds = np.flatnonzero(np.diff(padded[order])) + 1
    for grp in np.split(order, bounds) if n else []:
        L = int(padded[grp[0]])
        step = max(1, min(batch_size, max_tokens // L))
        for s in range(0, len(grp), step):
            idx = grp[s:s + step]
            ln = lens[idx]
            toks = np.full((len(idx), L), pad_id, dtype=np.int32)
            r = np.repeat(np.arange(len(idx)), ln)                         # destination row of every real token
            c = np.arange(int(ln.sum())) - np.repeat(np.cumsum(ln) - ln, ln)   # and its column
            toks[r, c] = flat[np.repeat(offs[idx], ln) + c]
            yield idx, toks, ln.astype(np.int32)


# ------------------------------------------------------------------------------------------------ encoder
class BertQueryEncoder:
    """Callable List[str] -> float32 [n, hidden] with L2-normalised rows: what BGERetriever(encoder=...) expects,
    and what the reference's `self.encoder.encode(..., normalize_embeddings=True)` returns."""

    def __init__(self, state_dict: Mapping[str, torch.Tensor], cfg: BertConfig, tokenizer: WordPieceTokenizer,
                 device: Optional[torch.device] = None, batch_size: int = 8192):
        if not torch.cuda.is_available():
            raise RuntimeError("BertQueryEncoder needs a CUDA (sm_100) device: there is no CPU fallback")
        self.cfg = cfg
        self.tokenizer = tokenizer
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        N.check(N.lib().qf_device_check(self.device.index or 0))
        self.batch_size = int(batch_size)
        self.weights = pack_bert_weights(state_dict, cfg).to(self.device)

    @classmethod
    def from_pretrained(cls, model_dir: str, **kwargs) -> "BertQueryEncoder":
        """A local HF / sentence-transformers checkpoint directory: config.json, vocab.txt, model.safetensors or
        pytorch_model.bin, optionally 1_Pooling/config.json (pooling mode) and tokenizer_config.json (lower-casing)."""
        d = Path(model_dir)
        if not (d / "config.json").exists():
            raise FileNotFoundError(f"No config.json under {d}")
        hf = json.loads((d / "config.json").read_text())
        # a plain HF checkpoint (no sentence-transformers modules): SentenceTransformer(name) adds MEAN pooling;
        # bge-base-en-v1.5 ships 1_Pooling/config.json with pooling_mode_cls_token = true
        pooling = "mean"
        if (d / "1_Pooling" / "config.json").exists():
            pooling = "cls"
            pc = json.loads((d / "1_Pooling" / "config.json").read_text())
            if pc.get("pooling_mode_mean_tokens"):
                pooling = "mean"
            elif not pc.get("pooling_mode_cls_token", True):
                raise ValueError(f"{d}: pooling mode not implemented: {pc}")
        lower, max_len = True, int(hf["max_position_embeddings"])
        if (d / "tokenizer_config.json").exists():
            tc = json.loads((d / "tokenizer_config.json").read_text())
            lower = bool(tc.get("do_lower_case", True))
        if (d / "sentence_bert_config.json").exists():
            max_len = min(max_len, int(json.loads((d / "sentence_bert_config.json").read_text()).get("max_seq_length", max_len)))
        if (d / "model.safetensors").exists():
            from safetensors.torch import load_file
            sd = load_file(str(d / "model.safetensors"))
        elif (d / "pytorch_model.bin").exists():
            sd = torch.load(str(d / "pytorch_model.bin"), map_location="cpu", weights_only=True)
        else:
            raise FileNotFoundError(f"No model.safetensors / pytorch_model.bin under {d}")
        return cls(sd, BertConfig.from_hf(hf, pooling), WordPieceTokenizer(d / "vocab.txt", lower, max_len), **kwargs)

    def encode_tokens(self, tokens: torch.Tensor, lens: torch.Tensor, _status_out: Optional[list] = None
                      ) -> Tuple[torch.Tensor, torch.Tensor]:
        """tokens int32 [n, L], lens int32 [n] on the device -> (bf16 [n, H], float32 [n, H]); raises on a token id
        outside the vocabulary."""
        c = self.cfg
        out_b, out_f, status = ops.bert_encode(self.weights, tokens, lens, n_layers=c.n_layers, hidden=c.hidden,
                                               heads=c.heads, ffn=c.ffn, vocab=c.vocab, max_pos=c.max_pos,
                                               pooling=c.pooling_id, ln_eps=c.ln_eps)
        if _status_out is not None:                     # the caller checks all its calls with one host sync
            _status_out.append(status)
        elif int(status.max().item()) != 0:
            bad = torch.nonzero(status).flatten()[:5].tolist()
            raise ValueError(f"token id outside [0, {c.vocab}) in rows {bad}")
        return out_b, out_f

    def encode_token_lists(self, rows: Sequence[Sequence[int]], max_tokens: int = 1 << 18
                           ) -> Tuple[torch.Tensor, torch.Tensor]:
        """Token id rows ([CLS] ... [SEP] included) -> (bf16 [n, H], float32 [n, H]) on the device.
        The rows go to the device PACKED back to back (qf_bert_encode_packed: tokens [M] + prefix sums of the
        lengths), so every GEMM row is a real token -- a query batch whose lengths spread from 5 to 32 tokens costs
        its 14-token average, not the longest query.  Batches are cut at max_tokens tokens (workspace ~15 kB per
        token for bert-base) and at batch_size rows."""
        n = len(rows)
        if n == 0:
            return self.encode_packed(np.zeros(0, dtype=np.int32), np.zeros(0, dtype=np.int64))
        lens = np.fromiter((len(r) for r in rows), dtype=np.int64, count=n)
        if lens.min() < 1 or lens.max() > self.cfg.max_pos:
            raise ValueError(f"row lengths must lie in [1, {self.cfg.max_pos}]")
        if isinstance(rows[0], np.ndarray):
            flat = np.concatenate(rows).astype(np.int32, copy=False)
        else:
            import itertools
            flat64 = np.fromiter(itertools.chain.from_iterable(rows), dtype=np.int64, count=int(lens.sum()))
            if flat64.size and (flat64.min() < -2**31 or flat64.max() >= 2**31):
                raise ValueError("token id does not fit int32")
            flat = flat64.astype(np.int32)
        return self.encode_packed(flat, lens, max_tokens)

    def encode_packed(self, flat: np.ndarray, lens: np.ndarray, max_tokens: int = 1 << 18
                      ) -> Tuple[torch.Tensor, torch.Tensor]:
        """Host arrays already in the packed layout -- flat int32 [sum(lens)] token ids back to back, lens [n] --
        -> (bf16 [n, H], float32 [n, H]) on the device (what WordPieceTokenizer.encode_batch_packed returns)."""
        lens = np.asarray(lens, dtype=np.int64)
        flat = np.ascontiguousarray(flat, dtype=np.int32)
        n, H = len(lens), self.cfg.hidden
        out_b = torch.empty((n, H), dtype=torch.bfloat16, device=self.device)
        out_f = torch.empty((n, H), dtype=torch.float32, device=self.device)
        if n == 0:
            return out_b, out_f
        if lens.min() < 1 or lens.max() > self.cfg.max_pos or int(lens.sum()) != flat.size:
            raise ValueError(f"row lengths must lie in [1, {self.cfg.max_pos}] and add up to the number of tokens")
        offs = np.concatenate(([0], np.cumsum(lens)))
        c = self.cfg
        kw = dict(n_layers=c.n_layers, hidden=c.hidden, heads=c.heads, ffn=c.ffn, vocab=c.vocab, max_pos=c.max_pos,
                  pooling=c.pooling_id, ln_eps=c.ln_eps)
        statuses = []
        s = 0
        while s < n:                                   # greedy cut: <= batch_size rows and <= max_tokens tokens
            e = int(np.searchsorted(offs, offs[s] + max_tokens, side="right")) - 1
            e = min(max(e, s + 1), s + self.batch_size, n)
            cu = (offs[s:e + 1] - offs[s]).astype(np.int32)
            toks = flat[offs[s]:offs[e]]
            b, f, st = ops.bert_encode_packed(self.weights, torch.from_numpy(toks).to(self.device, non_blocking=True),
                                              torch.from_numpy(cu).to(self.device, non_blocking=True),
                                              int(lens[s:e].max()), **kw)
            out_b[s:e] = b
            out_f[s:e] = f
            statuses.append(st.max())
            s = e
        worst = int(torch.stack(statuses).max().item())
        if worst & 1:
            raise ValueError(f"token id outside [0, {self.cfg.vocab}) in the batch")
        if worst:
            raise RuntimeError("qf_bert_encode_packed: inconsistent sequence lengths")
        return out_b, out_f

    def _encode(self, queries: Sequence[str]) -> Tuple[torch.Tensor, torch.Tensor]:
        return self.encode_packed(*self.tokenizer.encode_batch_packed(list(queries)))

    def encode_device(self, queries: Sequence[str]) -> torch.Tensor:
        """bf16 [n, H] on the device: the `queries` operand of ops.dense_topk."""
        return self._encode(queries)[0]

    def __call__(self, queries: Sequence[str]) -> np.ndarray:
        return self._encode(queries)[1].cpu().numpy()
