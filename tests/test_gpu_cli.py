"""The `umgap` CLI against the oracle's text pipeline, stage by stage, on the B200 box."""
import os
import socket
import subprocess
import time

import pytest

import datagen
from oracle import fasta as ofasta, fstv2, lookup as olookup, pipeline as opipe
from oracle.taxonomy import Taxonomy as OTaxonomy, format_taxon

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UMGAP = os.path.join(ROOT, "umgap_b200", "bin", "umgap")


def run(args, stdin: str = ""):
    p = subprocess.run([UMGAP] + args, input=stdin.encode(), stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=300)
    return p.returncode, p.stdout.decode(), p.stderr.decode()


@pytest.fixture(scope="module")
def files(tmp_path_factory):
    import umgap_b200.capi as c
    if c.device_count() <= 0:
        pytest.fail("no CUDA device")
    d = tmp_path_factory.mktemp("cli")
    taxa = datagen.make_taxonomy(300, seed=71)
    otax = OTaxonomy(taxa)
    proteins = datagen.make_proteome(60, seed=72)
    index = datagen.make_index(proteins, otax, seed=73)
    (d / "taxons.tsv").write_bytes(("\n".join(format_taxon(t) for t in taxa) + "\n").encode("latin-1"))
    (d / "nine.fst").write_bytes(fstv2.build(sorted(index.items())))
    tryp = {}
    for i, p in enumerate(proteins):
        for pep in olookup.tryptic_filter(olookup.tryptic_digest(p), 5, 50):
            tryp.setdefault(pep.encode(), taxa[i % len(taxa)][0])
    (d / "tryp.fst").write_bytes(fstv2.build(sorted(tryp.items())))
    reads = datagen.make_reads(proteins, 80, seed=74)
    reads += [("short/1", "ACGTAC"), ("short/2", "ACG"), ("n/1", "N" * 60), ("n/2", "ACGT" * 20)]
    fasta = "".join(f">{h}\n{s[:70]}\n{s[70:]}\n" for h, s in reads)   # hard-wrapped input, unwrapped by the reader
    return dict(dir=d, taxa=taxa, otax=otax, index=index, tryp=tryp, reads=reads, fasta=fasta, proteins=proteins)


def test_stagewise_pipeline_matches_oracle_text(files):
    d = files["dir"]
    oidx = olookup.DictIndex(files["index"])
    rc, t_out, err = run(["translate", "-a"], files["fasta"])
    assert rc == 0, err
    assert t_out == opipe.translate_text(files["fasta"])
    rc, n_out, _ = run(["translate", "-n", "-f", "1", "-f", "3R", "-m", "-t", "11"], files["fasta"])
    assert n_out == opipe.translate_text(files["fasta"], 11, True, ["1", "3R"], True)
    for one in (True, False):
        rc, k_out, err = run(["prot2kmer2lca"] + (["-o"] if one else []) + [str(d / "nine.fst")], t_out)
        assert rc == 0, err
        assert k_out == opipe.prot2kmer2lca_text(t_out, oidx, 9, one)
    rc, k_out, _ = run(["prot2kmer2lca", "-o", "-m", "-c", "120", str(d / "nine.fst")], t_out)
    for s, g in ((2, 0), (3, 0), (3, 1), (4, 2)):
        rc, s_out, err = run(["seedextend", "-s", str(s), "-g", str(g)], k_out)
        assert rc == 0, err
        assert s_out == opipe.seedextend_text(k_out, s, g)
    for s, g, pen in ((2, 0, 5), (3, 1, 7)):   # -r: the best-scoring extended seed only
        rc, r_out, err = run(["seedextend", "-s", str(s), "-g", str(g), "-r", str(d / "taxons.tsv"), "-p", str(pen)], k_out)
        assert rc == 0, err
        assert r_out == opipe.seedextend_ranked_text(k_out, files["otax"], s, g, pen)
    rc, s_out, _ = run(["seedextend", "-s", "3"], k_out)
    rc, u_out, _ = run(["uniq", "-d", "/"], s_out)
    assert u_out == opipe.uniq_text(s_out, "/")
    for flags, strategy, factor, lb, ranked in [(["-a", "lca*"], 0, 0.25, 0.0, False), ([], 1, 0.25, 0.0, False),
                                                (["-m", "rmq", "-a", "mrtl", "-l", "1"], 2, 0.25, 1.0, False),
                                                (["-a", "hybrid", "-f", "0.5", "-l", "2", "-r"], 1, 0.5, 2.0, True)]:
        rc, a_out, err = run(["taxa2agg"] + flags + [str(d / "taxons.tsv")], u_out)
        assert rc == 0, err
        want = opipe.taxa2agg_sets(u_out, files["otax"], strategy, factor, lb, ranked)
        got = list(ofasta.read_records(a_out, False))
        assert len(got) == len(want)
        for (gh, gs), (wh, ws) in zip(got, want):
            assert gh == wh and len(gs) == 1 and int(gs[0]) in ws, (gh, gs, ws)
    # taxa2agg -s: "taxon=score" lines (taxa2agg.rs:141-148); scores derived from the ids, short decimals
    scored_in = "".join(f">{h}\n" + "".join(f"{t}={(int(t) % 7 + 1) / 4}\n" for t in seq) for h, seq in ofasta.read_records(u_out, False))
    from oracle import agg as oagg
    for flags, strategy, lb in [(["-s", "-a", "lca*", "-l", "1.5"], 0, 1.5), (["-s"], 1, 0.0), (["-s", "-m", "rmq", "-a", "mrtl", "-l", "0.5"], 2, 0.5)]:
        rc, a_out, err = run(["taxa2agg"] + flags + [str(d / "taxons.tsv")], scored_in)
        assert rc == 0, err
        got = list(ofasta.read_records(a_out, False))
        recs = list(ofasta.read_records(scored_in, False))
        assert len(got) == len(recs)
        snapping = files["otax"].snapping(False)
        for (gh, gs), (wh, ws) in zip(got, recs):
            pairs = [(int(x.split("=")[0]), float(x.split("=")[1])) for x in ws]
            assert gh == wh and len(gs) == 1 and int(gs[0]) in oagg.taxa2agg_record_scored(files["otax"], snapping, pairs, strategy, 0.25, lb), gh
    # -m rmq -a lca*: the fold of rmq/lca.rs:60-90 over the record's distinct taxa, restated with its Euler tour and RMQ
    from oracle import rmq as ormq
    calc = ormq.LCACalculator(files["otax"])
    rc, a_out, err = run(["taxa2agg", "-m", "rmq", "-a", "lca*", "-l", "2", str(d / "taxons.tsv")], u_out)
    assert rc == 0, err
    snapping = files["otax"].snapping(False)
    for (gh, gs), (wh, ws) in zip(ofasta.read_records(a_out, False), ofasta.read_records(u_out, False)):
        counts = oagg.filter_counts(oagg.count(int(x) for x in ws if int(x) != 0), 2.0)
        want = snapping[calc.aggregate(list(counts))] if counts else 1
        assert gh == wh and int(gs[0]) == want, gh
    rc, a_out, err = run(["taxa2agg", "-m", "rmq", "-a", "hybrid", "-f", "0.5", str(d / "taxons.tsv")], u_out)   # rmq/mix.rs
    assert rc == 0 and "Warning: this is a hybrid between LCA/MRTL" in err
    snapping = files["otax"].snapping(False)
    for (gh, gs), (wh, ws) in zip(ofasta.read_records(a_out, False), ofasta.read_records(u_out, False)):
        assert gh == wh and int(gs[0]) in oagg.taxa2agg_record_scored(files["otax"], snapping, [(int(x), 1.0) for x in ws], oagg.RMQ_HYBRID, 0.5), gh
    rc, out, err = run(["taxa2agg", "-s", str(d / "taxons.tsv")], ">r\n5\n")
    assert rc == 1 and "Taxon without score" in err
    # the fused command prints what the five-stage pipe prints
    rc, c_out, err = run(["classify", "-s", "3", "-a", "lca*", str(d / "nine.fst"), str(d / "taxons.tsv")], files["fasta"])
    assert rc == 0, err
    rc, ref_out, _ = run(["taxa2agg", "-a", "lca*", str(d / "taxons.tsv")], u_out)
    assert c_out == ref_out
    # errors: message on stderr, exit 1
    rc, out, err = run(["taxa2agg", str(d / "taxons.tsv")], ">r\n999999999\n")
    assert rc == 1 and "Unknown Taxon ID: 999999999" in err
    rc, out, err = run(["seedextend"], ">r\nabc\n")
    assert rc == 1 and err.startswith("Error:")
    rc, out, err = run(["translate", "-a", "-t", "8"], ">r\nACG\n")
    assert rc == 1 and "Unknown table" in err
    rc, out, err = run(["prot2kmer2lca", str(d / "missing.fst")], "")
    assert rc == 1


def test_tryptic_cli_matches_oracle_text(files):
    d = files["dir"]
    text = "".join(f">p{i}\n{p[:80]}\n{p[80:]}\n" for i, p in enumerate(files["proteins"][:30])) + ">e\n>s\n*K*\n"
    oidx = olookup.DictIndex(files["tryp"])
    for flags, kw in [([], {}), (["-o"], {"one_on_one": True}), (["-o", "-l", "9", "-L", "45"], {"one_on_one": True, "minlen": 9, "maxlen": 45}),
                      (["-k", "L", "-d", "W"], {"keep": "L", "drop": "W"})]:
        rc, out, err = run(["prot2tryp2lca"] + flags + [str(d / "tryp.fst")], text)
        assert rc == 0, err
        assert out == opipe.prot2tryp2lca_text(text, oidx, **kw), flags
    rc, out, err = run(["prot2tryp2lca", "-p", "([KR])", str(d / "tryp.fst")], text)
    assert rc == 1 and "pattern" in err


def test_socket_server_mode(files, tmp_path):
    """prot2kmer2lca -m -o -s <socket> serves FASTA over a Unix socket until client EOF
    (prot2kmer2lca.rs:116-137; `nc -NU` in scripts/umgap-analyse.sh:279)."""
    d = files["dir"]
    sock = str(tmp_path / "umgap.sock")
    srv = subprocess.Popen([UMGAP, "prot2kmer2lca", "-m", "-o", "-s", sock, str(d / "nine.fst")], stdout=subprocess.PIPE)
    try:
        assert srv.stdout.readline().decode() == "Socket created, listening for connections.\n"
        t_out = opipe.translate_text(files["fasta"])
        want = opipe.prot2kmer2lca_text(t_out, olookup.DictIndex(files["index"]), 9, True)
        for _ in range(2):  # sequential connections
            c = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
            c.connect(sock)
            c.sendall(t_out.encode())
            c.shutdown(socket.SHUT_WR)
            chunks = []
            while True:
                b = c.recv(1 << 16)
                if not b:
                    break
                chunks.append(b)
            c.close()
            assert b"".join(chunks).decode() == want
    finally:
        srv.kill()
        srv.wait()


NC_STANDIN = """#!/usr/bin/env python3
# stand-in for `nc -NU <socket>` (scripts/umgap-analyse.sh:279): stdin -> Unix socket, half-close at EOF, socket -> stdout
import socket, sys, threading
c = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
c.connect(sys.argv[-1])
def up():
    while True:
        b = sys.stdin.buffer.read(1 << 16)
        if not b:
            break
        c.sendall(b)
    c.shutdown(socket.SHUT_WR)
t = threading.Thread(target=up)
t.start()
while True:
    b = c.recv(1 << 16)
    if not b:
        break
    sys.stdout.buffer.write(b)
sys.stdout.buffer.flush()
t.join()
"""


def test_analyse_script_presets_through_the_socket_server(files, tmp_path):
    """The pipeline section of scripts/umgap-analyse.sh (:257-264 the index server, :276-311 the six case arms), arm by arm
    as shell pipes of `umgap` commands, with stand-ins for what the image lacks: `nc -NU` (a Python socket client) and
    FGSpp (gene prediction: its output is a FASTA of predicted peptides, here fragments of the proteome).  Every arm's
    text is the oracle's text pipeline on the same input (ties of hybrid / mrtl: membership)."""
    d = files["dir"]
    bindir = tmp_path / "bin"
    bindir.mkdir()
    (bindir / "nc").write_text(NC_STANDIN)
    os.chmod(bindir / "nc", 0o755)
    os.symlink(UMGAP, bindir / "umgap")
    env = dict(os.environ, PATH=str(bindir) + os.pathsep + os.environ.get("PATH", ""))
    sock = str(tmp_path / "socket")
    taxons, ninemers, tryptics = str(d / "taxons.tsv"), str(d / "nine.fst"), str(d / "tryp.fst")
    genes = []   # what FGSpp prints for paired reads: >header/mate_start_end_strand, one predicted peptide each
    for i, p in enumerate(files["proteins"][:40]):
        genes.append((f"r{i}/1_1_150_+", p[:50]))
        genes.append((f"r{i}/2_4_150_-", p[60:108] + ("*" if i % 5 == 0 else "")))
    genes += [("lone/1_1_30_+", "MKRW"), ("none/1_1_30_+", "")]
    (tmp_path / "genes.fa").write_text("".join(f">{h}\n" + (f"{s}\n" if s else "") for h, s in genes))
    (tmp_path / "reads.fa").write_text(files["fasta"])
    srv = subprocess.Popen(["umgap", "prot2kmer2lca", "-m", "-o", "-s", sock, ninemers], stdout=subprocess.DEVNULL, env=env)
    try:
        for _ in range(600):   # `while [ ! -S "$socket" ] && sleep 1` of the script
            if os.path.exists(sock):
                break
            time.sleep(0.1)
        assert os.path.exists(sock)
        arms = {
            "max-sensitivity": ("reads.fa", "umgap translate -a | nc -NU $socket | umgap seedextend -g1 -s2 | umgap uniq -d / | umgap taxa2agg -l1 -m rmq -a mrtl $taxons"),
            "high-sensitivity": ("reads.fa", "umgap translate -a | nc -NU $socket | umgap seedextend -g1 -s3 | umgap uniq -d / | umgap taxa2agg -l1 -a hybrid -f 0.25 $taxons"),
            "tryptic-sensitivity": ("genes.fa", "cat | umgap prot2tryp2lca -l9 -L45 $tryptics | umgap uniq -d / | umgap taxa2agg -l1 -m rmq -a mrtl $taxons"),
            "tryptic-precision": ("genes.fa", "cat | umgap prot2tryp2lca -l9 -L45 $tryptics | umgap uniq -d / | umgap taxa2agg -l5 -m rmq -a mrtl $taxons"),
            "high-precision": ("genes.fa", "cat | nc -NU $socket | umgap seedextend -g1 -s3 | umgap uniq -d / | umgap taxa2agg -l2 -a lca\\* $taxons"),
            "max-precision": ("genes.fa", "cat | nc -NU $socket | umgap seedextend -g1 -s4 | umgap uniq -d / | umgap taxa2agg -l5 -a lca\\* $taxons"),
        }
        nine, tryp = olookup.DictIndex(files["index"]), olookup.DictIndex(files["tryp"])
        otax = files["otax"]
        below_root = 0
        for name, (infile, pipe) in arms.items():
            cmd = f"set -o pipefail; socket={sock}; taxons={taxons}; tryptics={tryptics}; {{ {pipe}; }} < {tmp_path / infile}"
            p = subprocess.run(["bash", "-c", cmd], stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=300, env=env)
            assert p.returncode == 0, (name, p.stderr.decode())
            text = (tmp_path / infile).read_text()
            if name.endswith("sensitivity") and infile == "reads.fa":
                text = opipe.translate_text(text)
            if name.startswith("tryptic"):
                ids = opipe.prot2tryp2lca_text(text, tryp, False, 9, 45)
            else:
                g, s = {"max-sensitivity": (1, 2), "high-sensitivity": (1, 3), "high-precision": (1, 3), "max-precision": (1, 4)}[name]
                ids = opipe.seedextend_text(opipe.prot2kmer2lca_text(text, nine, 9, True), s, g)
            joined = opipe.uniq_text(ids, "/")
            strategy, lb = {"max-sensitivity": (2, 1.0), "high-sensitivity": (1, 1.0), "tryptic-sensitivity": (2, 1.0),
                            "tryptic-precision": (2, 5.0), "high-precision": (0, 2.0), "max-precision": (0, 5.0)}[name]
            want = opipe.taxa2agg_sets(joined, otax, strategy, 0.25, lb, False)
            got = list(ofasta.read_records(p.stdout.decode(), False))
            assert len(got) == len(want) and len(got) > 0, name
            for (gh, gs), (wh, ws) in zip(got, want):
                assert gh == wh and len(gs) == 1 and int(gs[0]) in ws, (name, gh, gs, ws)
                below_root += int(gs[0]) != 1
        assert below_root > 50
    finally:
        srv.kill()
        srv.wait()


def test_fused_peptide_cli_matches_the_three_stage_pipe(files):
    """`umgap classify-peptides` prints what `prot2tryp2lca | uniq -d / | taxa2agg` prints (tryptic presets)."""
    d = files["dir"]
    recs = []
    for i, p in enumerate(files["proteins"][:40]):
        recs.append((f"g{i}/1", [p[:120]]))
        recs.append((f"g{i}/2", [p[100:200], p[200:260] + "*" + p[260:300]]))     # two physical lines in one record
    recs += [("e/1", []), ("e/2", ["*"]), ("solo", ["MKR" * 20])]
    text = "".join(f">{h}\n" + "".join(l + "\n" for l in ls) for h, ls in recs)
    for tflags, aflags in ((["-l", "9", "-L", "45"], ["-m", "rmq", "-a", "mrtl", "-l", "1"]),
                           (["-l", "5", "-L", "50", "-d", "CW"], ["-a", "lca*"])):
        rc, t_out, err = run(["prot2tryp2lca"] + tflags + [str(d / "tryp.fst")], text)
        assert rc == 0, err
        rc, u_out, _ = run(["uniq", "-d", "/"], t_out)
        rc, want, err = run(["taxa2agg"] + aflags + [str(d / "taxons.tsv")], u_out)
        assert rc == 0, err
        fused_flags = tflags + [("-b" if f == "-l" else f) for f in aflags]
        rc, got, err = run(["classify-peptides"] + fused_flags + [str(d / "tryp.fst"), str(d / "taxons.tsv")], text)
        assert rc == 0, err
        assert got == want, (tflags, aflags)
        assert got.count(">") == 42 and sum(l != "1" for l in got.split("\n")[1::2]) > 10
        # the command parses blocks of whole groups on several threads: seams at every place, from a pipe and from a file
        # (mapped and cut in place), CRLF line ends, no final newline, two classifier threads
        odd = text.replace("\n", "\r\n", 30).rstrip("\n")
        (d / "pep.fa").write_text(odd)
        for block, devices, extra in (("16", None, ["-P", "1"]), ("300", "0,0", ["-P", "3"]), ("5000", None, [])):
            env = dict(os.environ, UMGAP_CLI_BLOCK=block)
            if devices:
                env["UMGAP_DEVICES"] = devices
            cmd = [UMGAP, "classify-peptides"] + extra + fused_flags + [str(d / "tryp.fst"), str(d / "taxons.tsv")]
            p = subprocess.run(cmd, input=odd.encode(), stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=300, env=env)
            assert p.returncode == 0, p.stderr.decode()
            assert p.stdout.decode() == want, (block, devices)
            with open(d / "pep.fa", "rb") as f:
                p = subprocess.run(cmd, stdin=f, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=300, env=env)
            assert p.returncode == 0, p.stderr.decode()
            assert p.stdout.decode() == want, (block, devices, "file")


def test_fused_cli_block_parser_edge_cases(files):
    """`umgap classify` parses the stream a block at a time: CRLF line ends, hard-wrapped and empty records, a last
    record without a newline, block seams at every place (UMGAP_CLI_BLOCK), several parser threads and index replicas -- always the bytes of the five-stage pipe."""
    d = files["dir"]
    reads = files["reads"]
    text = files["fasta"].replace("\n", "\r\n", 40)                      # some CRLF line ends
    text += ">empty/1\n>empty/2\n\n>tail/1\n" + reads[0][1][:75] + "\n" + reads[0][1][75:] + "\n>tail/2\n" + reads[1][1]   # no final newline
    rc, t_out, err = run(["translate", "-a"], text)
    assert rc == 0, err
    rc, k_out, _ = run(["prot2kmer2lca", "-o", str(d / "nine.fst")], t_out)
    rc, s_out, _ = run(["seedextend", "-s", "3"], k_out)
    rc, u_out, _ = run(["uniq", "-d", "/"], s_out)
    rc, want, _ = run(["taxa2agg", "-a", "lca*", str(d / "taxons.tsv")], u_out)
    args = ["classify", "-s", "3", "-a", "lca*", str(d / "nine.fst"), str(d / "taxons.tsv")]
    # block seams (UMGAP_CLI_BLOCK bytes: blocks of one group up to the whole stream), parser threads, and two replicas
    # of the index driven by two classifier threads (UMGAP_DEVICES=0,0: both on this box's GPU)
    for block, devices, extra in ((None, None, []), ("16", None, ["-P", "1"]), ("300", "0,0", ["-P", "3"]), ("2000", None, []),
                                  ("70000", "0,0", ["--gpus", "2"])):
        env = dict(os.environ)
        if block:
            env["UMGAP_CLI_BLOCK"] = block
        if devices:
            env["UMGAP_DEVICES"] = devices
        p = subprocess.run([UMGAP] + args[:1] + extra + args[1:], input=text.encode(), stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                           timeout=300, env=env)
        assert p.returncode == 0, p.stderr.decode()
        assert p.stdout.decode() == want, (block, devices)
    # a regular file on stdin is mapped and cut in place; the parser threads populate, read or fault their blocks' pages
    (d / "edge.fa").write_text(text)
    for block, mode in ((None, None), ("16", "populate"), ("300", "pread"), ("2000", "mmap"), ("70000", "populate")):
        env = dict(os.environ)
        if block:
            env["UMGAP_CLI_BLOCK"] = block
        if mode:
            env["UMGAP_CLI_READ"] = mode
        with open(d / "edge.fa", "rb") as f:
            p = subprocess.run([UMGAP] + args, stdin=f, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=300, env=env)
        assert p.returncode == 0, p.stderr.decode()
        assert p.stdout.decode() == want, (block, mode, "file")
    # a short read between the mates of a pair is dropped before uniq sees it (prot2kmer2lca.rs:172): the mates still join,
    # also when a block seam falls next to it
    lines = text.split("\n>")
    odd = "\n>".join(lines[:9] + ["short/1\nACGT"] + lines[9:])
    rc, t_out, _ = run(["translate", "-a"], odd)
    rc, k_out, _ = run(["prot2kmer2lca", "-o", str(d / "nine.fst")], t_out)
    rc, s_out, _ = run(["seedextend", "-s", "3"], k_out)
    rc, u_out, _ = run(["uniq", "-d", "/"], s_out)
    rc, want_odd, _ = run(["taxa2agg", "-a", "lca*", str(d / "taxons.tsv")], u_out)
    for block in ("16", "500", "1500"):
        env = dict(os.environ, UMGAP_CLI_BLOCK=block)
        p = subprocess.run([UMGAP] + args, input=odd.encode(), stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=300, env=env)
        assert p.returncode == 0, p.stderr.decode()
        assert p.stdout.decode() == want_odd, block
    assert want.count(">") >= 80
    # the same stream from a regular file: the command maps it and cuts the blocks in place
    (d / "edge.fa").write_text(text)
    for block in (None, "16", "300", "70000"):
        env = dict(os.environ)
        if block:
            env["UMGAP_CLI_BLOCK"] = block
        with open(d / "edge.fa", "rb") as fh:
            p = subprocess.run([UMGAP] + args, stdin=fh, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=300, env=env)
        assert p.returncode == 0, p.stderr.decode()
        assert p.stdout.decode() == want, ("file", block)
    rc, out, err = run(args, "ACGT\n>r\nACGT\n")
    assert rc == 1 and "Expected > at beginning of fasta header." in err
    rc, out, err = run(args, "")
    assert rc == 0 and out == ""


def test_cli_wall_clock_through_pipes(tmp_path):
    """SURVEY 8(d) timing method, last item: wall-clock runs of the CLI through pipes.  200 000 synthetic reads
    (the bench generator) against a 2e6-key fst file: the five-stage pipe of `umgap-analyse.sh` with our binaries,
    the fused `umgap classify`, and the C restatement of the reference on the host cores; the two CLI forms must
    print the same bytes.  Timings go to stdout and gpurun_out/cli_wall_clock.log (a record, not an assertion)."""
    import numpy as np
    from oracle import cport, synth
    import umgap_b200.capi as capi
    NGPU = capi.device_count()
    taxa = datagen.make_taxonomy(5000, seed=1)
    pre = synth.Preorder(taxa)
    n_prot, plen, npairs, rlen = 5000, 408, 100_000, 150
    keys, vals = synth.build_index(2, n_prot, plen, 70, 20, pre)
    fst = cport.fst_build_blob(keys.reshape(-1), np.arange(0, 9 * len(keys) + 1, 9, dtype=np.uint64), vals)
    (tmp_path / "nine.fst").write_bytes(fst)
    (tmp_path / "taxons.tsv").write_bytes(("\n".join(format_taxon(t) for t in taxa) + "\n").encode("latin-1"))
    nt = cport.synth_reads(2, n_prot, plen, 3, 0, npairs, rlen, 70)
    with open(tmp_path / "reads.fa", "wb") as f:
        for i in range(0, 2 * npairs, 2):
            f.write(b">r%d/1\n%s\n>r%d/2\n%s\n" % (i // 2, nt[i].tobytes(), i // 2, nt[i + 1].tobytes()))
    d = str(tmp_path)
    pipe = (f"{UMGAP} translate -a < {d}/reads.fa | {UMGAP} prot2kmer2lca -o {d}/nine.fst | {UMGAP} seedextend -s 3 | "
            f"{UMGAP} uniq -d / | {UMGAP} taxa2agg -a hybrid {d}/taxons.tsv > {d}/staged.out")
    fused = f"{UMGAP} classify -s 3 -a hybrid {d}/nine.fst {d}/taxons.tsv < {d}/reads.fa > {d}/fused.out"
    times = {}
    for name, cmd in (("fused_cold", fused), ("staged", pipe), ("fused", fused)):
        t0 = time.perf_counter()
        p = subprocess.run(["bash", "-o", "pipefail", "-c", cmd], stderr=subprocess.PIPE, timeout=600)
        times[name] = time.perf_counter() - t0
        assert p.returncode == 0, p.stderr.decode()[-2000:]
    # steady state of the fused command: the same reads many times over, from a pipe and from a file, on one GPU and on two
    # replicas; the start-up (CUDA context, index load) is timed on its own with an empty input and subtracted
    import shutil
    reps = int(max(50, min(400, shutil.disk_usage(d).free // 4 // os.path.getsize(tmp_path / "reads.fa"))))
    subprocess.run(["bash", "-c", f"for i in $(seq {reps}); do cat {d}/reads.fa; done > {d}/big.fa"], check=True)
    base = f"UMGAP_CLI_VERBOSE=1 {UMGAP} classify -s 3 -a hybrid {d}/nine.fst {d}/taxons.tsv"
    rates = {}
    for name, cmd in (("fused_big", f"cat {d}/big.fa | {base} | wc -l > {d}/big.count"),
                      ("fused_big_file", f"{base} < {d}/big.fa | wc -l > {d}/big.count"),
                      ("fused_big_file_2", f"UMGAP_DEVICES=0,{1 if NGPU > 1 else 0} {base} < {d}/big.fa | wc -l > {d}/big.count"),
                      ("pipe_alone", f"cat {d}/big.fa | cat > /dev/null")):
        t0 = time.perf_counter()
        p = subprocess.run(["bash", "-o", "pipefail", "-c", cmd], stderr=subprocess.PIPE, timeout=900)
        times[name] = time.perf_counter() - t0
        assert p.returncode == 0, p.stderr.decode()[-2000:]
        if name != "pipe_alone":
            assert int((tmp_path / "big.count").read_text()) == 2 * reps * npairs
            rates[name] = [l for l in p.stderr.decode().split("\n") if l.startswith("umgap classify:")][-1]
    os.unlink(tmp_path / "big.fa")
    staged, fused_out = (tmp_path / "staged.out").read_bytes(), (tmp_path / "fused.out").read_bytes()
    assert staged == fused_out
    assert staged.count(b">") == npairs
    # the reference's algorithm on the host cores (C restatement, in-memory arrays: no text parsing at all)
    img, ctax = cport.FstImage(fst), cport.RefTaxonomy(taxa)
    opts = cport.RefOpts(table=1, methionine=0, one_on_one=1, seedextend=1, min_seed_size=3, max_gap_size=0, strategy=1,
                         factor=0.25, lower_bound=0.0, ranked_only=0, k=9)
    flat = np.ascontiguousarray(nt.reshape(-1))
    off = np.arange(0, flat.size + 1, rlen, dtype=np.uint64)
    goff = np.arange(0, 2 * npairs + 1, 2, dtype=np.uint64)
    threads = os.cpu_count() or 1
    t0 = time.perf_counter()
    ref, _, _ = cport.classify(img, ctax, opts, flat, off, goff, threads=threads)
    times["c_port"] = time.perf_counter() - t0
    got = np.array([int(x) for x in fused_out.split(b"\n")[1::2]], dtype=np.uint64)
    agree = float((got == np.asarray(ref, dtype=np.uint64)).mean())
    assert agree > 0.97     # hybrid ties may differ (the C port takes one admissible answer, the GPU another)
    lines = [f"CLI wall clock, {2 * npairs} reads of {rlen} nt ({os.path.getsize(tmp_path / 'reads.fa') / 1e6:.0f} MB of FASTA), "
             f"{len(keys)}-key fst file ({len(fst) / 1e6:.0f} MB), {threads} host threads",
             f"  five-stage pipe (translate | prot2kmer2lca -o | seedextend | uniq | taxa2agg): {times['staged']:.2f} s = "
             f"{2 * npairs / times['staged'] / 1e3:.0f} k reads/s",
             f"  fused `umgap classify`: {times['fused']:.2f} s = {2 * npairs / times['fused'] / 1e3:.0f} k reads/s "
             f"(first run, cold: {times['fused_cold']:.2f} s) -- index load, FASTA parsing and CUDA start-up included",
             f"  fused `umgap classify`, the same reads {reps} times ({2 * reps * npairs} reads, {reps * os.path.getsize(tmp_path / 'reads.fa') / 1e9:.1f} GB of FASTA); "
             f"the command's own clock starts when index and taxonomy are loaded (UMGAP_CLI_VERBOSE)",
             f"    through a pipe (cat |), {times['fused_big']:.2f} s wall: {rates['fused_big']}",
             f"      (cat | cat > /dev/null alone moves the same bytes in {times['pipe_alone']:.2f} s = "
             f"{reps * os.path.getsize(tmp_path / 'reads.fa') / times['pipe_alone'] / 1e9:.2f} GB/s)",
             f"    from a file (stdin redirected), {times['fused_big_file']:.2f} s wall: {rates['fused_big_file']}",
             f"    two index replicas (UMGAP_DEVICES, {NGPU} GPU(s) on this box), from the file, {times['fused_big_file_2']:.2f} s wall: {rates['fused_big_file_2']}",
             f"  C restatement of the reference, arrays in memory, {threads} threads: {times['c_port']:.2f} s = "
             f"{2 * npairs / times['c_port'] / 1e3:.0f} k reads/s; answers equal to the CLI's on {agree:.4f} of the pairs"]
    print("\n".join(lines))
    out_dir = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out_dir):
        with open(os.path.join(out_dir, "cli_wall_clock.log"), "w") as f:
            f.write("\n".join(lines) + "\n")
