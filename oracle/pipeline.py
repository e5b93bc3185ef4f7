"""Stage-by-stage text pipelines exactly as `umgap-analyse.sh` composes them (oracle; test
infrastructure only).  Each function maps a stage's stdin text to its stdout text and follows
the command loop it names; see the per-module docstrings for the reference line numbers.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Set, Tuple

from . import agg, fasta, lookup, seedextend as se, translate as tr
from .taxonomy import Taxonomy


def translate_text(text: str, table: int = 1, methionine: bool = False,
                   frames: Sequence[str] = tuple(tr.FRAME_NAMES), append_name: bool = False) -> str:
    """src/commands/translate.rs:99-133."""
    tr.get_table(table)
    out = []
    for header, seq in fasta.read_records(text, unwrap=True):
        for name, pep in tr.translate_record(seq[0], table, methionine, frames):
            h = header + "|" + name if append_name else header
            out.append(fasta.write_record(h, [pep], "", False))
    return "".join(out)


def _ids_text(records) -> str:
    # prot2kmer2lca.rs:173-183 / prot2tryp2lca.rs:108,131: ">{header}\n" then "{id}\n" per id
    out = []
    for header, ids in records:
        out.append(">" + header + "\n")
        for v in ids:
            out.append(f"{v}\n")
    return "".join(out)


def prot2kmer2lca_text(text: str, index, k: int = 9, one_on_one: bool = False) -> str:
    return _ids_text(lookup.prot2kmer2lca(fasta.read_records(text, True), index, k, one_on_one))


def prot2tryp2lca_text(text: str, index, one_on_one: bool = False, minlen: int = 5,
                       maxlen: int = 50, keep: str = "", drop: str = "") -> str:
    return _ids_text(lookup.prot2tryp2lca(fasta.read_records(text, False), index, one_on_one,
                                          minlen, maxlen, keep, drop))


def _parse_ids(seq: List[str]) -> List[int]:
    from .taxonomy import _parse_usize
    return [_parse_usize(s) for s in seq]


def seedextend_text(text: str, min_seed_size: int = 2, max_gap_size: int = 0) -> str:
    """src/commands/seedextend.rs:92-176."""
    out = []
    for header, seq in fasta.read_records(text, unwrap=False):
        ids = se.seedextend(_parse_ids(seq), min_seed_size, max_gap_size)
        out.append(fasta.write_record(header, [str(i) for i in ids], "\n", False))
    return "".join(out)


def seedextend_ranked_text(text: str, tax: Taxonomy, min_seed_size: int = 2, max_gap_size: int = 0, penalty: int = 5) -> str:
    """seedextend -r <taxon file> -p <penalty>, src/commands/seedextend.rs:84-176."""
    out = []
    for header, seq in fasta.read_records(text, unwrap=False):
        ids = se.seedextend_ranked(_parse_ids(seq), tax, min_seed_size, max_gap_size, penalty)
        out.append(fasta.write_record(header, [str(i) for i in ids], "\n", False))
    return "".join(out)


def uniq_text(text: str, delimiter: Optional[str] = None, separator: str = "\n",
              wrap: bool = False) -> str:
    recs = fasta.uniq(fasta.read_records(text, unwrap=False), delimiter)
    return "".join(fasta.write_record(h, s, separator, wrap) for h, s in recs)


def taxa2agg_sets(text: str, tax: Taxonomy, strategy: int, factor: float = 0.25,
                  lower_bound: float = 0.0, ranked_only: bool = False
                  ) -> List[Tuple[str, Set[int]]]:
    """src/commands/taxa2agg.rs:159-181; per record the admissible answer set."""
    snapping = tax.snapping(ranked_only)
    out = []
    for header, seq in fasta.read_records(text, unwrap=False):
        out.append((header, agg.taxa2agg_record(tax, snapping, _parse_ids(seq), strategy,
                                                factor, lower_bound)))
    return out


def classify_reads(reads: Sequence[Tuple[str, str]], index, tax: Taxonomy, *, table: int = 1,
                   methionine: bool = False, k: int = 9, one_on_one: bool = True, use_seedextend: bool = True,
                   min_seed_size: int = 2, max_gap_size: int = 0, delimiter: Optional[str] = "/",
                   strategy: int = agg.HYBRID, factor: float = 0.25, lower_bound: float = 0.0,
                   ranked_only: bool = False) -> List[Tuple[str, Set[int]]]:
    """`translate -a | prot2kmer2lca -o | seedextend | uniq -d / | taxa2agg` on (header, nt)
    reads, composed through the text stages above so that every stream-format quirk applies."""
    text = "".join(fasta.write_record(h, [s], "", False) for h, s in reads)
    text = translate_text(text, table, methionine)
    text = prot2kmer2lca_text(text, index, k, one_on_one=one_on_one)
    if use_seedextend:
        text = seedextend_text(text, min_seed_size, max_gap_size)
    text = uniq_text(text, delimiter)
    return taxa2agg_sets(text, tax, strategy, factor, lower_bound, ranked_only)
