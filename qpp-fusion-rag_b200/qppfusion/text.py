"""
Query-text features consumed by the WIG / SigmaMax / MaxIDF / avgidf estimators.

The reference tokenises inside the JVM with queryText.toLowerCase().split("\\s+")
(src/qpp/java/qpp/QPPBridge.java:250,272,369,386).  String work stays on the host; the kernel receives
four integers per query: (n_terms, n_unique, max_len, sum_len).  Java semantics reproduced:
  * \\s = [ \\t\\n\\x0B\\f\\r] (no UNICODE_CHARACTER_CLASS);
  * a leading whitespace run yields a leading empty token, trailing empty tokens are dropped,
    "" -> [""] (one token), an all-whitespace string -> [] (zero tokens);
  * String.length() counts UTF-16 code units.
"""
from typing import Iterable, List, Tuple

import numpy as np

_WS = frozenset(" \t\n\x0b\f\r")


def java_split_whitespace(text: str) -> List[str]:
    out: List[str] = []
    start, i, n, seen = 0, 0, len(text), False
    while i < n:
        if text[i] in _WS:
            out.append(text[start:i])
            while i < n and text[i] in _WS:
                i += 1
            start, seen = i, True
        else:
            i += 1
    if not seen:
        return [text]
    out.append(text[start:])
    while out and out[-1] == "":
        out.pop()
    return out


def _len16(s: str) -> int:
    return len(s) + sum(1 for ch in s if ord(ch) > 0xFFFF)


def text_features(query_text: str) -> Tuple[int, int, int, int]:
    terms = java_split_whitespace(query_text.lower())
    uniq = set(terms)
    return (len(terms), len(uniq), max((_len16(t) for t in uniq), default=0), sum(_len16(t) for t in terms))


def text_features_array(texts: Iterable[str]) -> np.ndarray:
    """int32 [Q, 4] for qf_qpp_stats."""
    return np.array([text_features(t) for t in texts], dtype=np.int32).reshape(-1, 4)
