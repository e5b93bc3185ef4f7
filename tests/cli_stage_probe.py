"""Wall clock of every stage of the five-stage pipe on its own (files in, files out), and of the pipe: where the
stage-wise CLI spends its time.  python tests/cli_stage_probe.py [npairs]"""
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))  # datagen
import numpy as np
import datagen
from oracle import cport, synth
from oracle.taxonomy import format_taxon

UMGAP = os.path.join(ROOT, "umgap_b200", "bin", "umgap")
npairs = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
with tempfile.TemporaryDirectory() as d:
    taxa = datagen.make_taxonomy(5000, seed=1)
    pre = synth.Preorder(taxa)
    n_prot, plen, rlen = 5000, 408, 150
    keys, vals = synth.build_index(2, n_prot, plen, 70, 20, pre)
    open(f"{d}/nine.fst", "wb").write(cport.fst_build_blob(keys.reshape(-1), np.arange(0, 9 * len(keys) + 1, 9, dtype=np.uint64), vals))
    open(f"{d}/taxons.tsv", "wb").write(("\n".join(format_taxon(t) for t in taxa) + "\n").encode("latin-1"))
    nt = cport.synth_reads(2, n_prot, plen, 3, 0, npairs, rlen, 70)
    with open(f"{d}/reads.fa", "wb") as f:
        for i in range(0, 2 * npairs, 2):
            f.write(b">r%d/1\n%s\n>r%d/2\n%s\n" % (i // 2, nt[i].tobytes(), i // 2, nt[i + 1].tobytes()))
    stages = [("startup (translate of nothing)", f"{UMGAP} translate -a < /dev/null > /dev/null"),
              ("translate -a", f"{UMGAP} translate -a < {d}/reads.fa > {d}/t.out"),
              ("prot2kmer2lca -o", f"{UMGAP} prot2kmer2lca -o {d}/nine.fst < {d}/t.out > {d}/k.out"),
              ("seedextend -s 3", f"{UMGAP} seedextend -s 3 < {d}/k.out > {d}/s.out"),
              ("uniq -d /", f"{UMGAP} uniq -d / < {d}/s.out > {d}/u.out"),
              ("taxa2agg -a hybrid", f"{UMGAP} taxa2agg -a hybrid {d}/taxons.tsv < {d}/u.out > {d}/a.out"),
              ("the five-stage pipe", f"{UMGAP} translate -a < {d}/reads.fa | {UMGAP} prot2kmer2lca -o {d}/nine.fst | {UMGAP} seedextend -s 3 | "
                                      f"{UMGAP} uniq -d / | {UMGAP} taxa2agg -a hybrid {d}/taxons.tsv > {d}/p.out"),
              ("umgap classify", f"{UMGAP} classify -s 3 -a hybrid {d}/nine.fst {d}/taxons.tsv < {d}/reads.fa > {d}/c.out")]
    for rep in range(2):
        for name, cmd in stages:
            t0 = time.perf_counter()
            subprocess.run(["bash", "-o", "pipefail", "-c", cmd], check=True)
            dt = time.perf_counter() - t0
            if rep:
                print(f"{name:32s} {dt:6.2f} s", flush=True)
    for f in ("t.out", "k.out", "s.out", "u.out", "a.out"):
        print(f, os.path.getsize(f"{d}/{f}") / 1e6, "MB")
    print("pipe == staged files:", open(f"{d}/p.out", "rb").read() == open(f"{d}/a.out", "rb").read(),
          " classify == pipe:", open(f"{d}/c.out", "rb").read() == open(f"{d}/p.out", "rb").read())
