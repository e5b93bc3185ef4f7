"""FASTA / FASTQ stream model of the reference (oracle; test infrastructure only).

Follows /root/reference/src/io/fasta.rs:30-67 (reader), :164-180 (writer),
src/io/fastq.rs:26-87 (FASTQ reader), src/commands/fastq2fasta.rs:62-84 and
src/commands/uniq.rs:56-84.
"""
from __future__ import annotations

from typing import Iterable, Iterator, List, Tuple

Record = Tuple[str, List[str]]  # (header without '>', items)


class FastaError(Exception):
    pass


def _lines(text: str) -> List[str]:
    # BufRead::lines(): split on '\n', strip one trailing '\r' (fasta.rs:36).
    if text == "":
        return []
    parts = text.split("\n")
    if parts[-1] == "":
        parts.pop()
    return [p[:-1] if p.endswith("\r") else p for p in parts]


def read_records(text: str, unwrap: bool) -> Iterator[Record]:
    """fasta.rs:38-67.  `unwrap` concatenates the item lines into exactly one item."""
    lines = _lines(text)
    i = 0
    while i < len(lines):
        header = lines[i]
        i += 1
        if not header.startswith(">"):
            raise FastaError("Expected > at beginning of fasta header.")  # fasta.rs:44-49
        header = header[1:]
        seq: List[str] = []
        while i < len(lines) and not lines[i].startswith(">"):
            seq.append(lines[i])
            i += 1
        if unwrap:
            seq = ["".join(seq)]  # fasta.rs:62-64 (always one element, possibly "")
        yield header, seq


def write_record(header: str, items: Iterable[str], sep: str, wrap: bool = False) -> str:
    """fasta.rs:164-180."""
    out = [">" + header]
    s = sep.join(items)
    if not wrap:
        out.append("\n")
        out.append(s)
    else:
        b = s.encode()
        for k in range(0, len(b), 70):
            out.append("\n")
            out.append(b[k:k + 70].decode())
    if s != "":
        out.append("\n")
    return "".join(out)


def read_fastq(text: str) -> Iterator[Tuple[str, str]]:
    """fastq.rs:26-87: '@header', n sequence lines until a '+' line, then n quality lines."""
    lines = _lines(text)
    i = 0
    while i < len(lines):
        header = lines[i]
        i += 1
        if not header.startswith("@"):
            raise FastaError("Expected @ at beginning of fastq header.")
        header = header[1:]
        seq = []
        while i < len(lines) and not lines[i].startswith("+"):
            seq.append(lines[i])
            i += 1
        n = len(seq)
        if i < len(lines):
            i += 1  # the '+' line (fastq.rs:54-66; EOF here is not an error)
        if i + n > len(lines):
            raise FastaError("Expected as many quality lines as sequence lines.")
        i += n
        yield header, "".join(seq)


def fastq2fasta(texts: List[str]) -> str:
    """fastq2fasta.rs:62-84 + utils.rs:4-21: interleave; stop when any input is exhausted."""
    its = [read_fastq(t) for t in texts]
    out = []
    while True:
        recs = []
        for it in its:
            r = next(it, None)
            if r is None:
                return "".join(out)
            recs.append(r)
        for h, s in recs:
            out.append(write_record(h, [s], "", False))


def uniq(records: Iterable[Record], delimiter: str | None) -> List[Record]:
    """uniq.rs:56-84: truncate header at first delimiter, merge consecutive equal headers."""
    out: List[Record] = []
    last = None
    for header, seq in records:
        if delimiter is not None:
            p = header.find(delimiter)
            if p >= 0:
                header = header[:p]
        if last is not None:
            if last[0] == header:
                last[1].extend(seq)
            else:
                out.append(last)
                last = (header, list(seq))
        else:
            last = (header, list(seq))
    if last is not None:
        out.append(last)
    return out
