"""k-mer and tryptic lookups, oracle restatement (test infrastructure only).

Follows /root/reference/src/commands/prot2kmer2lca.rs:115,163-192 and
src/commands/prot2tryp2lca.rs:95-137 (digest + filters + lookup).  The index is any object
with `.get(bytes) -> Optional[int]` (oracle.fstv2.Fst or a dict wrapper).
"""
from __future__ import annotations

import re
from typing import Iterable, List, Optional, Tuple

from .fasta import Record


class DictIndex:
    def __init__(self, d):
        self.d = d

    def get(self, key: bytes) -> Optional[int]:
        return self.d.get(bytes(key))


def prot2kmer2lca(records: Iterable[Record], index, k: int = 9, one_on_one: bool = False
                  ) -> List[Tuple[str, List[int]]]:
    """prot2kmer2lca.rs:168-185.  Input records must be read with unwrap=True.
    A record whose peptide is shorter than k is dropped entirely (no header)."""
    default = 0 if one_on_one else None  # :115
    out = []
    for header, seq in records:
        prot = seq[0].encode() if seq else None
        if prot is None or len(prot) < k:  # :172 (byte length)
            continue
        ids = []
        for i in range(len(prot) - k + 1):  # :174-175
            v = index.get(prot[i:i + k])
            if v is None:
                v = default
            if v is not None:
                ids.append(v)
        out.append((header, ids))
    return out


_PATTERN = re.compile(r"([KR])([^P])")


def tryptic_digest_regex(line: str) -> List[str]:
    """prot2tryp2lca.rs:112-117 verbatim: two regex passes, '*' -> newline, non-empty lines.
    Python's `re.sub` has the same leftmost, non-overlapping semantics as regex::replace_all
    and `[^P]` matches '\\n' in both."""
    first = _PATTERN.sub(r"\1\n\2", line)
    second = _PATTERN.sub(r"\1\n\2", first).replace("*", "\n")
    return [x for x in second.split("\n") if x != ""]


def tryptic_digest(line: str) -> List[str]:
    """Closed form: cleave after K/R unless the next char is P; '*' separates."""
    out, cur = [], []
    n = len(line)
    for i, c in enumerate(line):
        if c == "*":
            if cur:
                out.append("".join(cur))
                cur = []
            continue
        cur.append(c)
        if c in "KR" and i + 1 < n and line[i + 1] != "P":
            out.append("".join(cur))
            cur = []
    if cur:
        out.append("".join(cur))
    return out


def tryptic_filter(peps: Iterable[str], minlen: int = 5, maxlen: int = 50,
                   keep: str = "", drop: str = "") -> List[str]:
    """prot2tryp2lca.rs:118-128: byte length window, then keep/drop residue sets."""
    ks, ds = set(keep), set(drop)
    out = []
    for p in peps:
        L = len(p.encode())
        if not (minlen <= L <= maxlen):
            continue
        if ks or ds:
            s = set(p)
            if not (len(ks & s) == len(ks) and len(ds & s) == 0):
                continue
        out.append(p)
    return out


def prot2tryp2lca(records: Iterable[Record], index, one_on_one: bool = False,
                  minlen: int = 5, maxlen: int = 50, keep: str = "", drop: str = ""
                  ) -> List[Tuple[str, List[int]]]:
    """prot2tryp2lca.rs:105-134.  Input records must be read with unwrap=False: every
    physical line is digested on its own; the header is always emitted."""
    default = 0 if one_on_one else None
    out = []
    for header, seq in records:
        ids = []
        for line in seq:
            for pep in tryptic_filter(tryptic_digest(line), minlen, maxlen, keep, drop):
                v = index.get(pep.encode())
                if v is None:
                    v = default
                if v is not None:
                    ids.append(v)
        out.append((header, ids))
    return out
