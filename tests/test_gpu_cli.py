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
