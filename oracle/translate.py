"""Six-frame translation, oracle restatement (test infrastructure only).

Follows /root/reference/src/dna/mod.rs:23-44,78-103 (Nucleotide, Strand),
src/dna/translation.rs:20,33-45 (codon order T,C,A,G), :47-104 (NCBI genetic codes; the
strings below are the public NCBI `gc.prt` AAs / Starts rows), :125-144 (translate,
translate_frame) and src/commands/translate.rs:82-133 (frame order and record emission).
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

# NCBI genetic codes (AAs, Starts) in TCAG order; ids 7, 8, 17-20 do not exist
# (translation.rs:47-104 holds the same 19 tables).
TABLES = {
    1: ("FFLLSSSSYY**CC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG",
        "---M---------------M---------------M----------------------------"),
    2: ("FFLLSSSSYY**CCWWLLLLPPPPHHQQRRRRIIMMTTTTNNKKSS**VVVVAAAADDEEGGGG",
        "--------------------------------MMMM---------------M------------"),
    3: ("FFLLSSSSYY**CCWWTTTTPPPPHHQQRRRRIIMMTTTTNNKKSSRRVVVVAAAADDEEGGGG",
        "----------------------------------MM----------------------------"),
    4: ("FFLLSSSSYY**CCWWLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG",
        "--MM---------------M------------MMMM---------------M------------"),
    5: ("FFLLSSSSYY**CCWWLLLLPPPPHHQQRRRRIIMMTTTTNNKKSSSSVVVVAAAADDEEGGGG",
        "---M----------------------------MMMM---------------M------------"),
    6: ("FFLLSSSSYYQQCC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG",
        "-----------------------------------M----------------------------"),
    9: ("FFLLSSSSYY**CCWWLLLLPPPPHHQQRRRRIIIMTTTTNNNKSSSSVVVVAAAADDEEGGGG",
        "-----------------------------------M---------------M------------"),
    10: ("FFLLSSSSYY**CCCWLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG",
         "-----------------------------------M----------------------------"),
    11: ("FFLLSSSSYY**CC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG",
         "---M---------------M------------MMMM---------------M------------"),
    12: ("FFLLSSSSYY**CC*WLLLSPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG",
         "-------------------M---------------M----------------------------"),
    13: ("FFLLSSSSYY**CCWWLLLLPPPPHHQQRRRRIIMMTTTTNNKKSSGGVVVVAAAADDEEGGGG",
         "---M------------------------------MM---------------M------------"),
    14: ("FFLLSSSSYYY*CCWWLLLLPPPPHHQQRRRRIIIMTTTTNNNKSSSSVVVVAAAADDEEGGGG",
         "-----------------------------------M----------------------------"),
    15: ("FFLLSSSSYY*QCC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG",
         "-----------------------------------M----------------------------"),
    16: ("FFLLSSSSYY*LCC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG",
         "-----------------------------------M----------------------------"),
    21: ("FFLLSSSSYY**CCWWLLLLPPPPHHQQRRRRIIMMTTTTNNNKSSSSVVVVAAAADDEEGGGG",
         "-----------------------------------M---------------M------------"),
    22: ("FFLLSS*SYY*LCC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG",
         "-----------------------------------M----------------------------"),
    23: ("FF*LSSSSYY**CC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG",
         "--------------------------------M--M---------------M------------"),
}

FRAME_NAMES = ["1", "2", "3", "1R", "2R", "3R"]  # translate.rs:83-90 order
_ORDER = {"T": 0, "C": 1, "A": 2, "G": 3}       # translation.rs:20
_COMP = {"A": "T", "C": "G", "G": "C", "T": "A", "N": "N"}  # dna/mod.rs:23-31


class UnknownTable(Exception):
    pass


def get_table(table: int) -> Tuple[str, str]:
    """translation.rs:176-185: TABLES[id-1] or `Unknown table`."""
    if table not in TABLES:
        raise UnknownTable(f"Unknown table: {table}")
    return TABLES[table]


def strand(seq: str) -> List[str]:
    """dna/mod.rs:34-44,78-88: only uppercase A C G T survive, everything else is N."""
    return [c if c in "ACGT" else "N" for c in seq]


def reversed_strand(fwd: Sequence[str]) -> List[str]:
    """dna/mod.rs:101-103."""
    return [_COMP[c] for c in reversed(fwd)]


def translate_codon(aas: str, starts: str, methionine: bool, b0: str, b1: str, b2: str) -> str:
    """translation.rs:125-132: a codon holding an N is absent from the map -> '-'."""
    if "N" in (b0, b1, b2):
        return "-"
    idx = 16 * _ORDER[b0] + 4 * _ORDER[b1] + _ORDER[b2]
    if methionine and starts[idx] == "M":
        return "M"
    return aas[idx]


def translate_frame(table: int, methionine: bool, s: Sequence[str], frame: int) -> str:
    """dna/mod.rs:92-98 + translation.rs:136-144: suffix from frame-1, whole codons only."""
    aas, starts = get_table(table)
    x = s[frame - 1:] if len(s) > frame - 1 else []
    return "".join(
        translate_codon(aas, starts, methionine, x[3 * j], x[3 * j + 1], x[3 * j + 2])
        for j in range(len(x) // 3)
    )


def translate_record(seq: str, table: int = 1, methionine: bool = False,
                     frames: Sequence[str] = tuple(FRAME_NAMES)) -> List[Tuple[str, str]]:
    """translate.rs:114-133: returns [(frame name, peptide)] in the order of `frames`."""
    fwd = strand(seq)
    rev = reversed_strand(fwd)
    out = []
    for name in frames:
        f = int(name[0])
        s = rev if name.endswith("R") else fwd
        out.append((name, translate_frame(table, methionine, s, f)))
    return out
