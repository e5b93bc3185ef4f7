"""Regenerates tests/golden/pipeline_small.json from the Python oracle (the restatement of the reference
that tests/test_oracle_golden.py pins to the reference's own unit-test vectors and doc examples).

The reference is a Rust binary that cannot be built in this image, so these are ORACLE outputs, frozen
so that neither the CUDA path nor the oracle can drift unnoticed.  Run:  python tests/golden/make_golden.py
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import datagen  # noqa: E402
from oracle import lookup as olookup, pipeline as opipe  # noqa: E402
from oracle.taxonomy import Taxonomy  # noqa: E402


def build():
    taxa = datagen.make_taxonomy(120, seed=201)
    tax = Taxonomy(taxa)
    proteins = datagen.make_proteome(25, seed=202, lo=60, hi=90)
    index = datagen.make_index(proteins, tax, seed=203)
    reads = datagen.make_reads(proteins, 40, seed=204)
    reads += [("short/1", "ACGTACGTAC"), ("short/2", "ACG"), ("n/1", "N" * 45), ("n/2", "ACGT" * 12)]
    cases = []
    for name, kw in [("high_precision_like", dict(min_seed_size=3, max_gap_size=0, strategy=1, factor=0.25)),
                     ("max_sensitivity", dict(min_seed_size=2, max_gap_size=1, strategy=2, lower_bound=1.0)),
                     ("lca_star_l2", dict(min_seed_size=3, max_gap_size=1, strategy=0, lower_bound=2.0)),
                     ("no_seedextend_no_o", dict(use_seedextend=False, one_on_one=False, strategy=1, factor=0.5))]:
        out = opipe.classify_reads(reads, olookup.DictIndex(index), tax, **kw)
        cases.append({"name": name, "options": kw, "expected": [[h, sorted(s)] for h, s in out]})
    return {"taxa": [[t[0], t[2], t[3], int(t[4])] for t in taxa],
            "index": sorted([k.decode(), v] for k, v in index.items()),
            "reads": reads, "cases": cases}


if __name__ == "__main__":
    with open(os.path.join(HERE, "pipeline_small.json"), "w") as f:
        json.dump(build(), f, separators=(",", ":"))
    print("written", os.path.getsize(os.path.join(HERE, "pipeline_small.json")), "bytes")
