"""BASELINE.json configs[0]: the reference's own input fixture (testdata/A1.fq + A2.fq) through
`fastq2fasta | translate -a | prot2kmer2lca [-o] | taxa2agg -a 'lca*'` against the frozen oracle output of
tests/golden/config0.json.gz (made by tests/golden/make_config0.py, which documents the index recipe).

CPU half: the golden is what the oracle and the C restatement print today, and — where /root/reference is
present — the committed reads are byte for byte what fastq2fasta makes of the reference's FASTQ files.
GPU half (-m gpu): the `umgap` CLI stage by stage, the fused `umgap classify`, and the C ABI entry points, all
against the same golden.  The reads reach the GPU box inside the golden (the reference tree does not travel).
"""
import gzip
import hashlib
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))

import make_config0 as mk  # noqa: E402
from oracle import cport, fasta as ofasta, fstv2, pipeline as opipe  # noqa: E402
from oracle.taxonomy import Taxonomy as OTaxonomy, format_taxon  # noqa: E402

UMGAP = os.path.join(ROOT, "umgap_b200", "bin", "umgap")
TESTDATA = "/root/reference/testdata"


def sha(text: str) -> str:
    return hashlib.sha256(text.encode()).hexdigest()


@pytest.fixture(scope="module")
def cfg0():
    with gzip.open(os.path.join(ROOT, "tests", "golden", "config0.json.gz")) as f:
        doc = json.loads(f.read().decode())
    taxa = mk.config0_taxa()
    keys, vals = mk.config0_index([tuple(e) for e in doc["seeded"]], taxa, doc["index_keys"])
    assert len(keys) == doc["index_keys"] == 1_000_000
    digest = hashlib.sha256(b"".join(k + v.to_bytes(4, "little") for k, v in zip(keys, vals))).hexdigest()
    assert digest == doc["index_sha256"], "the seeded padding of the config-1 index is not reproducible here"
    blob = np.frombuffer(b"".join(keys), dtype=np.uint8)
    fst_bytes = cport.fst_build_blob(blob, np.arange(0, 9 * len(keys) + 1, 9, dtype=np.uint64), np.array(vals, dtype=np.uint64))
    return dict(doc=doc, exp=doc["expected"], taxa=taxa, otax=OTaxonomy(taxa), keys=keys, vals=vals, fst=fst_bytes,
                fasta=doc["fasta"])


def run(args, stdin: str = ""):
    p = subprocess.run([UMGAP] + args, input=stdin.encode(), stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=300)
    return p.returncode, p.stdout.decode(), p.stderr.decode()


def records(text):
    return [(h, s) for h, s in ofasta.read_records(text, False)]


# ------------------------------------------------------------------------------------------------ CPU

def test_config0_reads_are_the_reference_fixture(cfg0, tmp_path):
    """fastq2fasta.rs:62-84 on the reference's files gives the committed FASTA: 200 reads of 100 nt, no N."""
    doc = cfg0["doc"]
    assert doc["fasta"].count(">") == 200
    if os.path.isdir(TESTDATA):
        texts = [open(os.path.join(TESTDATA, f)).read() for f in ("A1.fq", "A2.fq")]
        assert [hashlib.sha256(t.encode()).hexdigest() for t in texts] == doc["fastq_sha256"]
        assert ofasta.fastq2fasta(texts) == doc["fasta"]
        rc, out, err = run(["fastq2fasta", os.path.join(TESTDATA, "A1.fq"), os.path.join(TESTDATA, "A2.fq")])
        assert rc == 0, err
        assert out == doc["fasta"]
    # the same through FASTQ files rebuilt from the committed reads (what the GPU box sees)
    fq = _write_fastq(doc["fasta"], tmp_path)
    rc, out, err = run(["fastq2fasta"] + fq)
    assert rc == 0, err
    assert out == doc["fasta"]


def _write_fastq(fasta_text, d):
    recs = list(ofasta.read_records(fasta_text, True))
    paths = []
    for m in (0, 1):
        p = d / f"A{m + 1}.fq"
        p.write_text("".join(f"@{h}\n{s[0]}\n+\n{'I' * len(s[0])}\n" for h, s in recs[m::2]))
        paths.append(str(p))
    return paths


def test_config0_oracle_reproduces_the_golden(cfg0):
    exp = mk.expected_outputs(cfg0["fasta"], dict(zip(cfg0["keys"], cfg0["vals"])), cfg0["otax"])
    assert exp == cfg0["exp"]
    assert exp["translate_records"] == 1200 and exp["kmer_o_ids"] == 29600   # 148 nine-mers per 100-nt read


def test_config0_c_restatement_and_fst_image(cfg0):
    """The timed CPU baseline (oracle/c) on the same inputs: its fst codec at 1e6 keys and its pipeline."""
    img = cport.FstImage(cfg0["fst"])
    for k, v in list(zip(cfg0["keys"], cfg0["vals"]))[::4999]:
        assert img.get(k) == v
    assert img.get(b"AAAAAAAAB") is None
    pyfst = fstv2.Fst(cfg0["fst"])   # the Python codec reads the C codec's image
    for kmer, v in cfg0["doc"]["seeded"][::97]:
        assert pyfst.get(kmer.encode()) == v
    recs = list(ofasta.read_records(cfg0["fasta"], True))
    nt = np.frombuffer("".join(s[0] for _, s in recs).encode(), dtype=np.uint8)
    off = np.zeros(len(recs) + 1, dtype=np.uint64)
    off[1:] = np.cumsum([len(s[0]) for _, s in recs])
    goff = np.arange(0, len(recs) + 1, 2, dtype=np.uint64)
    ctax = cport.RefTaxonomy(cfg0["taxa"])
    for one, tag in ((0, "plain"), (1, "o")):
        opts = cport.RefOpts(table=1, methionine=0, one_on_one=one, seedextend=0, min_seed_size=2, max_gap_size=0, strategy=0,
                             factor=0.25, lower_bound=0.0, ranked_only=0, k=9)
        out, nl, nh = cport.classify(img, ctax, opts, nt, off, goff, threads=2)
        assert [int(x) for x in out] == [v for _, v in cfg0["exp"][f"per_pair_{tag}"]]
        assert nl == 29600 and nh == cfg0["exp"]["kmer_plain_ids"]


# ------------------------------------------------------------------------------------------------ GPU

@pytest.fixture(scope="module")
def files(cfg0, tmp_path_factory):
    import umgap_b200.capi as c
    if c.device_count() <= 0:
        pytest.fail("no CUDA device")
    d = tmp_path_factory.mktemp("config0")
    (d / "taxons.tsv").write_bytes(("\n".join(format_taxon(t) for t in cfg0["taxa"]) + "\n").encode("latin-1"))
    (d / "nine.fst").write_bytes(cfg0["fst"])
    fq = [os.path.join(TESTDATA, f) for f in ("A1.fq", "A2.fq")] if os.path.isdir(TESTDATA) else _write_fastq(cfg0["fasta"], d)
    return dict(dir=d, fq=fq)


@pytest.mark.gpu
def test_config0_stagewise_cli(cfg0, files):
    """umgap fastq2fasta A1.fq A2.fq | umgap translate -a | umgap prot2kmer2lca [-o] <fst> | umgap taxa2agg -a 'lca*'."""
    d, exp = files["dir"], cfg0["exp"]
    rc, fa, err = run(["fastq2fasta"] + files["fq"])
    assert rc == 0, err
    assert fa == cfg0["fasta"]
    rc, t_out, err = run(["translate", "-a"], fa)
    assert rc == 0, err
    assert sha(t_out) == exp["translate_sha256"]
    for flags, tag in (([], "plain"), (["-o"], "o")):
        rc, k_out, err = run(["prot2kmer2lca"] + flags + [str(d / "nine.fst")], t_out)
        assert rc == 0, err
        assert sha(k_out) == exp[f"kmer_{tag}_sha256"]
        rc, a_out, err = run(["taxa2agg", "-a", "lca*", str(d / "taxons.tsv")], k_out)
        assert rc == 0, err
        assert [[h, int(s[0])] for h, s in records(a_out)] == exp[f"per_frame_{tag}"]
        rc, u_out, _ = run(["uniq", "-d", "/"], k_out)
        rc, a_out, err = run(["taxa2agg", "-m", "tree", "-a", "lca*", str(d / "taxons.tsv")], u_out)
        assert rc == 0, err
        assert [[h, int(s[0])] for h, s in records(a_out)] == exp[f"per_pair_{tag}"]


@pytest.mark.gpu
def test_config0_fused_cli_and_presets(cfg0, files):
    d, exp = files["dir"], cfg0["exp"]
    fst, tsv = str(d / "nine.fst"), str(d / "taxons.tsv")
    for flags, tag in ((["-O"], "plain"), ([], "o")):
        rc, out, err = run(["classify", "-S"] + flags + ["-a", "lca*", fst, tsv], cfg0["fasta"])
        assert rc == 0, err
        assert [[h, int(s[0])] for h, s in records(out)] == exp[f"per_pair_{tag}"]
    for name, flags in (("high_sensitivity", ["-s", "3", "-g", "1", "-l", "1", "-a", "hybrid", "-f", "0.25"]),
                        ("max_sensitivity", ["-s", "2", "-g", "1", "-l", "1", "-m", "rmq", "-a", "mrtl"]),
                        ("bench_configuration", ["-s", "3", "-g", "0", "-a", "hybrid", "-f", "0.25"])):
        rc, out, err = run(["classify"] + flags + [fst, tsv], cfg0["fasta"])
        assert rc == 0, err
        got = records(out)
        want = exp["preset_" + name]
        assert [h for h, _ in got] == [h for h, _ in want]
        for (h, s), (_, adm) in zip(got, want):
            assert int(s[0]) in adm, (name, h, s, adm)
        # the five-stage pipe of scripts/umgap-analyse.sh:276-290 prints the same bytes as the fused command
        s_, g_ = flags[1], flags[3]
        rc, t_out, _ = run(["translate", "-a"], cfg0["fasta"])
        rc, k_out, _ = run(["prot2kmer2lca", "-o", fst], t_out)
        rc, s_out, _ = run(["seedextend", "-s", s_, "-g", g_], k_out)
        rc, u_out, _ = run(["uniq", "-d", "/"], s_out)
        rc, a_out, err = run(["taxa2agg"] + flags[4:] + [tsv], u_out)
        assert rc == 0, err
        assert a_out == out, name


@pytest.mark.gpu
def test_config0_c_abi(cfg0, files):
    """The library entry points on the same inputs: fst loader, translate, kmer lookup (every k-mer of every frame),
    aggregate per frame record, and the fused call in its sampled and every-position forms."""
    import umgap_b200.capi as capi
    exp = cfg0["exp"]
    d = files["dir"]
    gidx = capi.Index.load_fst(str(d / "nine.fst"), k=9)
    assert gidx.info().n_keys == 1_000_000
    gtax = capi.Taxonomy.load(str(d / "taxons.tsv"))
    recs = list(ofasta.read_records(cfg0["fasta"], True))
    nt, off = capi.pack_strings([s[0].encode() for _, s in recs])
    aa, aoff = capi.translate(nt, off)
    assert len(aoff) - 1 == 1200
    text = "".join(f">{recs[i // 6][0]}\n{bytes(aa[int(aoff[i]):int(aoff[i + 1])]).decode()}\n" for i in range(1200))
    assert sha(text) == exp["translate_sha256"]
    for one, tag in ((False, "plain"), (True, "o")):
        ids, ioff = capi.kmer_lookup(gidx, aa, aoff, one)[:2]
        assert len(ids) == exp[f"kmer_{tag}_ids"]
        got = capi.aggregate(gtax, ids, ioff, capi.AGG_LCA_STAR, 0.25, 0.0, False)
        assert [int(x) for x in got] == [v for _, v in exp[f"per_frame_{tag}"]]
        goff = np.arange(0, len(recs) + 1, 2, dtype=np.uint64)
        opts = capi.default_opts(one_on_one=int(one), seedextend=0, strategy=capi.AGG_LCA_STAR)
        out, nl = capi.classify_reads(gidx, gtax, opts, nt, off, goff)
        assert nl == 29600
        assert [int(x) for x in out] == [v for _, v in exp[f"per_pair_{tag}"]]
    goff = np.arange(0, len(recs) + 1, 2, dtype=np.uint64)
    for name, kw in mk.PRESETS.items():
        opts = capi.default_opts(seedextend=1, min_seed_size=kw["min_seed_size"], max_gap_size=kw["max_gap_size"],
                                 strategy=kw["strategy"], factor=kw.get("factor", 0.25), lower_bound=kw.get("lower_bound", 0.0))
        outs = []
        for sampling in (1, 0):
            prev = capi.pipeline_sampling(sampling)
            try:
                out, _ = capi.classify_reads(gidx, gtax, opts, nt, off, goff)
            finally:
                capi.pipeline_sampling(prev)
            outs.append([int(x) for x in out])
            for g, (_, adm) in zip(outs[-1], exp["preset_" + name]):
                assert g in adm, (name, sampling, g, adm)
        assert outs[0] == outs[1], name
