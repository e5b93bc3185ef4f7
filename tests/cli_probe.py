"""Steady-state rate of `umgap classify` from a FASTA file under its measurement knobs (UMGAP_CLI_READ, UMGAP_CLI_DEPTH,
--parser-threads): the reads of tests/test_gpu_cli.py::test_cli_wall_clock_through_pipes, many times over.
  python tests/cli_probe.py [reps] [variant,variant,...]     variant = read:depth:parsers, e.g. pread:2:12
"""
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

UMGAP = os.path.join(ROOT, "umgap_b200", "bin", "umgap")
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
variants = sys.argv[2].split(",") if len(sys.argv) > 2 else ["mmap:1:12", "mmap:2:12", "populate:2:12", "pread:2:12", "pread:2:14", "pread:3:14"]


with tempfile.TemporaryDirectory(dir="/dev/shm" if os.path.isdir("/dev/shm") and os.environ.get("CLI_PROBE_SHM") else None) as d:
    taxa = datagen.make_taxonomy(5000, seed=1)
    pre = synth.Preorder(taxa)
    n_prot, plen, npairs, rlen = 5000, 408, 100_000, 150
    keys, vals = synth.build_index(2, n_prot, plen, 70, 20, pre)
    fst = cport.fst_build_blob(keys.reshape(-1), np.arange(0, 9 * len(keys) + 1, 9, dtype=np.uint64), vals)
    open(f"{d}/nine.fst", "wb").write(fst)
    from oracle.taxonomy import format_taxon as ft
    open(f"{d}/taxons.tsv", "wb").write(("\n".join(ft(t) for t in taxa) + "\n").encode("latin-1"))
    nt = cport.synth_reads(2, n_prot, plen, 3, 0, npairs, rlen, 70)
    with open(f"{d}/reads.fa", "wb") as f:
        for i in range(0, 2 * npairs, 2):
            f.write(b">r%d/1\n%s\n>r%d/2\n%s\n" % (i // 2, nt[i].tobytes(), i // 2, nt[i + 1].tobytes()))
    subprocess.run(["bash", "-c", f"for i in $(seq {reps}); do cat {d}/reads.fa; done > {d}/big.fa"], check=True)
    subprocess.run(["bash", "-c", f"cat {d}/big.fa > /dev/null"], check=True)   # page cache warm
    print(f"{2 * reps * npairs} reads, {os.path.getsize(f'{d}/big.fa') / 1e9:.1f} GB of FASTA, {os.cpu_count()} host threads", flush=True)
    for v in variants:
        rd, depth, par = v.split(":")[:3]
        env = dict(os.environ, UMGAP_CLI_VERBOSE="1", UMGAP_CLI_READ=rd, UMGAP_CLI_DEPTH=depth)
        if len(v.split(":")) > 3:
            env["UMGAP_CLI_BLOCK"] = str(int(v.split(":")[3]) << 20)
        if len(v.split(":")) > 4:
            env["UMGAP_DEVICES"] = v.split(":")[4].replace("+", ",")
        cmd = f"{UMGAP} classify -s 3 -a hybrid --parser-threads {par} {d}/nine.fst {d}/taxons.tsv < {d}/big.fa | wc -l"
        t0 = time.perf_counter()
        p = subprocess.run(["bash", "-o", "pipefail", "-c", cmd], stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=env, timeout=600)
        dt = time.perf_counter() - t0
        line = [l for l in p.stderr.decode().split("\n") if l.startswith("umgap classify")]
        ok = p.returncode == 0 and int(p.stdout) == 2 * reps * npairs
        print(f"{v:18s} wall {dt:5.2f} s  ok={ok}  " + (" | ".join(line) if line else p.stderr.decode()[-300:]), flush=True)
