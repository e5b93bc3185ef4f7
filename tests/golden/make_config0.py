"""BASELINE.json configs[0] on the reference's own input fixture.

Reads /root/reference/testdata/A1.fq + A2.fq (the one input fixture the reference ships; 100 paired
100-nt reads, testdata/README.md:1-11) and freezes, under tests/golden/config0.json.gz:

  * `fasta`     what `umgap fastq2fasta A1.fq A2.fq` prints (fastq2fasta.rs:62-84) -- the reads travel to
                the GPU box in this form (/root/reference does not exist there);
  * `seeded`    the part of the config-1 index that SURVEY 8(d) seeds from the reads themselves: the reads are
                six-frame translated with the oracle, a seeded 60 % sample of their stop-free 9-mers is kept,
                every pair gets a "home" taxon and every sampled k-mer the home taxon (70 %), one of its
                ancestors (20 %) or an unrelated taxon (10 %).  `config0_index()` pads this to 1e6 k-mers
                with seeded random 9-mers at test time (the padding is not stored);
  * `expected`  what the ORACLE (the restatement pinned to the reference's unit-test vectors) prints for
                `translate -a | prot2kmer2lca [-o] <fst> | taxa2agg -a 'lca*' <taxa>` (one record per frame,
                the configuration as BASELINE.json states it) and for the same with `uniq -d /` in front of
                taxa2agg, plus SHA-256 digests of the intermediate streams.

The reference binary cannot be built here (no cargo), so `expected` is oracle output, frozen.  LCA* has no
ties (tree/lca.rs:34-40), so every expected answer is a single id.

Run:  python tests/golden/make_config0.py
"""
import gzip
import hashlib
import json
import os
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import numpy as np  # noqa: E402

import datagen  # noqa: E402
from oracle import fasta as ofasta, lookup as olookup, pipeline as opipe, translate as otr  # noqa: E402
from oracle.taxonomy import Taxonomy  # noqa: E402

TESTDATA = "/root/reference/testdata"
N_TAXA = 5000
TAXONOMY_SEED = 1
INDEX_KEYS = 1_000_000
PAD_SEED = 20
OUT = os.path.join(HERE, "config0.json.gz")


def sha(text: str) -> str:
    return hashlib.sha256(text.encode()).hexdigest()


def config0_taxa():
    return datagen.make_taxonomy(N_TAXA, seed=TAXONOMY_SEED)


def seeded_entries(fasta_text: str, tax: Taxonomy, seed: int = 2):
    """(k-mer, taxon) pairs drawn from the reads' own frames; duplicates merge to the LCA as joinkmers would."""
    rng = random.Random(seed)
    all_ids = sorted(t[0] for t in tax.by_id if t is not None)
    leaves = sorted(set(all_ids) - {t[3] for t in tax.by_id if t is not None})
    entries = {}
    home = None
    last_pair = None
    for header, seq in ofasta.read_records(fasta_text, unwrap=True):
        pair = header.split("/")[0]
        if pair != last_pair:
            home = rng.choice(leaves)
            last_pair = pair
        path = tax.root_path(home)
        for _name, pep in otr.translate_record(seq[0], 1, False, otr.FRAME_NAMES):
            for i in range(len(pep) - 8):
                kmer = pep[i:i + 9]
                if "*" in kmer or "-" in kmer or rng.random() >= 0.6:
                    continue
                u = rng.random()
                v = home if u < 0.7 else (rng.choice(path) if u < 0.9 else rng.choice(all_ids))
                if kmer in entries and entries[kmer] != v:
                    a, b = tax.root_path(entries[kmer]), tax.root_path(v)
                    j = 0
                    while j < min(len(a), len(b)) and a[j] == b[j]:
                        j += 1
                    v = a[j - 1]
                entries[kmer] = v
    return sorted(entries.items())


def config0_index(seeded, taxa, total: int = INDEX_KEYS):
    """The 1e6-key index of config 1: the seeded entries plus seeded random 9-mers over the 20 amino acids with
    random taxa.  Returns (sorted list of key bytes, list of values)."""
    rng = np.random.default_rng(PAD_SEED)
    ids = np.array(sorted(t[0] for t in taxa), dtype=np.uint64)
    aas = np.frombuffer(datagen.AAS.encode(), dtype=np.uint8)
    index = {k.encode(): int(v) for k, v in seeded}
    while len(index) < total:
        need = total - len(index)
        keys = aas[rng.integers(0, 20, size=(need + 16, 9))]
        vals = ids[rng.integers(0, len(ids), size=need + 16)]
        for kb, v in zip(keys, vals):
            kb = kb.tobytes()
            if kb not in index:
                index[kb] = int(v)
                if len(index) == total:
                    break
    keys = sorted(index)
    return keys, [index[k] for k in keys]


def expected_outputs(fasta_text: str, index: dict, tax: Taxonomy):
    oidx = olookup.DictIndex(index)
    t_out = opipe.translate_text(fasta_text)
    exp = {"translate_sha256": sha(t_out), "translate_records": t_out.count(">")}
    for one in (False, True):
        k_out = opipe.prot2kmer2lca_text(t_out, oidx, 9, one)
        tag = "o" if one else "plain"
        exp[f"kmer_{tag}_sha256"] = sha(k_out)
        exp[f"kmer_{tag}_ids"] = k_out.count("\n") - k_out.count(">")
        per_frame = opipe.taxa2agg_sets(k_out, tax, 0)
        assert all(len(s) == 1 for _, s in per_frame)
        exp[f"per_frame_{tag}"] = [[h, next(iter(s))] for h, s in per_frame]
        per_pair = opipe.taxa2agg_sets(opipe.uniq_text(k_out, "/"), tax, 0)
        assert all(len(s) == 1 for _, s in per_pair)
        exp[f"per_pair_{tag}"] = [[h, next(iter(s))] for h, s in per_pair]
    # the 9-mer presets of scripts/umgap-analyse.sh:276-290 on the same reads (admissible sets: hybrid / MRTL ties)
    reads = [(h, q[0]) for h, q in ofasta.read_records(fasta_text, unwrap=True)]
    for name, kw in PRESETS.items():
        out = opipe.classify_reads(reads, oidx, tax, **kw)
        exp["preset_" + name] = [[h, sorted(s)] for h, s in out]
    return exp


PRESETS = {"high_sensitivity": dict(min_seed_size=3, max_gap_size=1, strategy=1, factor=0.25, lower_bound=1.0),
           "max_sensitivity": dict(min_seed_size=2, max_gap_size=1, strategy=2, lower_bound=1.0),
           "bench_configuration": dict(min_seed_size=3, max_gap_size=0, strategy=1, factor=0.25)}


def main():
    texts = [open(os.path.join(TESTDATA, f)).read() for f in ("A1.fq", "A2.fq")]
    fasta_text = ofasta.fastq2fasta(texts)
    taxa = config0_taxa()
    tax = Taxonomy(taxa)
    seeded = seeded_entries(fasta_text, tax)
    keys, vals = config0_index(seeded, taxa)
    exp = expected_outputs(fasta_text, dict(zip(keys, vals)), tax)
    doc = {"source": "testdata/A1.fq + testdata/A2.fq of unipept/umgap through fastq2fasta (fastq2fasta.rs:62-84)",
           "fastq_sha256": [hashlib.sha256(t.encode()).hexdigest() for t in texts],
           "fasta": fasta_text, "n_taxa": N_TAXA, "taxonomy_seed": TAXONOMY_SEED, "index_keys": INDEX_KEYS,
           "pad_seed": PAD_SEED, "seeded": seeded, "index_sha256": hashlib.sha256(b"".join(
               k + v.to_bytes(4, "little") for k, v in zip(keys, vals))).hexdigest(), "expected": exp}
    with gzip.GzipFile(OUT, "wb", mtime=0) as f:
        f.write(json.dumps(doc, separators=(",", ":")).encode())
    hits = exp["kmer_plain_ids"]
    below = sum(1 for _, v in exp["per_pair_plain"] if v != 1)
    print(f"written {os.path.getsize(OUT)} bytes: {fasta_text.count('>')} reads, {len(seeded)} seeded k-mers, "
          f"{exp['kmer_o_ids']} lookups, {hits} hits, {below} of {len(exp['per_pair_plain'])} pairs below the root")


if __name__ == "__main__":
    main()
