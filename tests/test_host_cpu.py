"""CPU-side checks that need no GPU: the C ABI library loads and exports every declared symbol, the
header and the ctypes mirror agree, compute entry points fail loudly without a device, the host-only
CLI commands match the oracle's stream model, and the read partitioner works across 2 gloo ranks."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UMGAP = os.path.join(ROOT, "umgap_b200", "bin", "umgap")


@pytest.fixture(scope="module")
def built():
    if not (os.path.exists(os.path.join(ROOT, "umgap_b200", "lib", "libumgap_gpu.so")) and os.path.exists(UMGAP)):
        subprocess.run(["make", "-j8", "all"], cwd=ROOT, check=True, stdout=subprocess.DEVNULL)
    from umgap_b200 import capi
    return capi


def test_library_exports_every_header_symbol(built):
    header = open(os.path.join(ROOT, "include", "umgap_gpu.h")).read()
    declared = sorted(set(re.findall(r"\b(umgap_[a-z0-9_]+)\s*\(", header)))
    assert declared, "no declarations found"
    assert sorted(built.SYMBOLS) == declared
    lib = built.load_library()
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.umgap_abi_version() == 1


def test_structs_match_header_layout(built):
    import ctypes as C
    assert C.sizeof(built.PipelineOpts) == 40
    assert C.sizeof(built.SynthSpec) == 32
    o = built.default_opts()
    assert (o.table, o.one_on_one, o.seedextend, o.min_seed_size, o.max_gap_size, o.strategy) == (1, 1, 1, 2, 0, 1)
    assert abs(o.factor - 0.25) < 1e-9 and o.lower_bound == 0.0


def test_no_cpu_fallback(built):
    """Without a CUDA device the compute entry points must fail, not fall back."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    assert built.device_count() < 0
    nt, off = built.pack_strings([b"ACGTACGTAC"])
    with pytest.raises(built.UmgapError) as e:
        built.translate(nt, off)
    assert e.value.code == -3
    with pytest.raises(built.UmgapError):
        built.Index.from_pairs([b"ACDEFGHIK"], [5])
    with pytest.raises(built.UmgapError):
        built.Taxonomy.from_arrays([1], [1], [0], [1])


def _run(args, stdin=b""):
    p = subprocess.run([UMGAP] + args, input=stdin, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    return p.returncode, p.stdout.decode(), p.stderr.decode()


def test_cli_host_commands_match_oracle_streams(built, tmp_path):
    from oracle import fasta as ofasta, pipeline as opipe
    rc, out, _ = _run(["-V"])
    assert rc == 0 and out == "umgap 1.1.1\n"       # scripts/umgap-analyse.sh:106-109 probes this
    text = ">header1/1\n1\n2\n>header1/2\n3\n>x\n>y/1\n\n9\n>y/2\n>header1/1\n4\n"
    for flags, kw in [([], {}), (["-d", "/"], {"delimiter": "/"}), (["-d", "/", "-s", " "], {"delimiter": "/", "separator": " "}),
                      (["-d", "er", "-w"], {"delimiter": "er", "wrap": True})]:
        rc, out, err = _run(["uniq"] + flags, text.encode())
        assert rc == 0 and out == opipe.uniq_text(text, **kw), (flags, err)
    # uniq.rs:22-40 doc example
    rc, out, _ = _run(["uniq", "-d", "/"], b">header1/1\nsequence1\n>header1/2\nsequence2\n")
    assert out == ">header1\nsequence1\nsequence2\n"
    long_text = ">w\n" + "A" * 150 + "\n" + "C" * 10 + "\n"
    rc, out, _ = _run(["uniq", "-s", "", "-w"], long_text.encode())
    assert out == opipe.uniq_text(long_text, separator="", wrap=True)
    rc, out, err = _run(["uniq"], b"no header\n")
    assert rc == 1 and err == "Error: Expected > at beginning of fasta header.\n" and out == ""
    f1, f2 = tmp_path / "a.fq", tmp_path / "b.fq"
    q1 = "@r1/1\nACGT\nAC\n+\nIIII\nII\n@r2/1\nGGGG\n+\nIIII\n@r3/1\nTT\n+\nII\n"
    q2 = "@r1/2\nTTTT\n+r1/2\nIIII\n@r2/2\nCCCC\n+\nIIII\n"
    f1.write_text(q1)
    f2.write_text(q2)
    rc, out, _ = _run(["fastq2fasta", str(f1), str(f2)])
    assert rc == 0 and out == ofasta.fastq2fasta([q1, q2])
    rc, _, err = _run(["taxa2agg", "-m", "tree", "-a", "mrtl", "taxons.tsv"])
    assert rc == 1 and "cannot be combined" in err   # taxa2agg.rs:134-138
    rc, _, err = _run(["translate", "-a", "-f", "1"])
    assert rc == 1 and "cannot be used with" in err   # translate.rs:52
    rc, _, err = _run(["nonsense"])
    assert rc == 1 and err.startswith("Error:")


def test_split_groups_properties():
    from umgap_b200.partition import slice_batch, split_groups
    rng = np.random.default_rng(5)
    for _ in range(50):
        nreads = int(rng.integers(0, 200))
        lens = rng.integers(0, 300, size=nreads)
        read_off = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
        cuts = np.sort(rng.choice(np.arange(1, max(nreads, 1)), size=min(max(nreads - 1, 0), int(rng.integers(0, 40))), replace=False)) if nreads > 1 else []
        group_off = np.array([0] + list(cuts) + [nreads], dtype=np.uint64) if nreads else np.array([0], dtype=np.uint64)
        nt = rng.integers(65, 70, size=int(read_off[-1]), dtype=np.uint8)
        for world in (1, 2, 3, 8):
            parts = split_groups(group_off, read_off, world)
            assert parts[0][0] == 0 and parts[-1][1] == len(group_off) - 1
            assert all(a <= b for a, b in parts) and all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
            cat = b"".join(bytes(slice_batch(nt, read_off, group_off, a, b)[0]) for a, b in parts)
            assert cat == bytes(nt)


_WORKER = r'''
import os, sys
import numpy as np
import torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from umgap_b200.partition import classify_partitioned
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:" + sys.argv[2], rank=int(sys.argv[3]), world_size=2)
rank = dist.get_rank()
rng = np.random.default_rng(7)
lens = rng.integers(27, 200, size=501)
read_off = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
group_off = np.array(list(range(0, 501, 2)) + [501], dtype=np.uint64)
nt = rng.integers(65, 70, size=int(read_off[-1]), dtype=np.uint8)
def fake_classify(nt_s, roff, goff):   # stands in for umgap_classify_reads on this rank's GPU
    return np.array([int(nt_s[int(roff[int(goff[g])]):int(roff[int(goff[g + 1])])].astype(np.uint64).sum() % 100003)
                     for g in range(len(goff) - 1)], dtype=np.uint32)
got = classify_partitioned(fake_classify, nt, read_off, group_off, rank, 2, dist)
want = fake_classify(nt, read_off, group_off)
assert np.array_equal(got, want), "rank %d: partitioned result differs" % rank
t = __import__("torch").tensor([float(rank + 1)])
dist.all_reduce(t, op=dist.ReduceOp.MAX)      # the bench's max-over-ranks timing reduction
assert t.item() == 2.0
dist.destroy_process_group()
print("ok", rank)
'''


def test_partitioned_classification_two_gloo_ranks(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    port = str(29500 + os.getpid() % 2000)
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, port, str(r)], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
             for r in range(2)]
    for r, p in enumerate(procs):
        out, err = p.communicate(timeout=180)
        assert p.returncode == 0, err.decode()[-2000:]
        assert out.decode().strip() == f"ok {r}"


def test_seedextend_closed_form():
    """The classify kernel runs seedextend in closed form on bit masks (pipeline.cu: seedextend_frame): ranges =
    stretches between gaps longer than -g, selected iff they hold -s equal consecutive non-zero ids, plus the machine's
    one irregularity (a record that begins with 1..g zeros loses its first non-zero id).  This is the same closed form in
    Python against the line-by-line restatement of seedextend.rs:94-149 on random lists."""
    import random
    from collections import Counter
    from oracle.seedextend import seedextend
    M64 = (1 << 64) - 1

    def ffs(x):
        return (x & -x).bit_length()

    def closed_form(ids, S, G):
        cnt = len(ids)
        nz = eq = 0
        prev = None
        for i, v in enumerate(ids):
            if v != 0:
                nz |= 1 << i
            if prev is not None and v == prev:
                eq |= 1 << i
            prev = v
        if not nz:
            return Counter()
        valid = (1 << cnt) - 1
        inr = nz
        if G:
            g_eff = min(G, 63)
            z = ffs(nz) - 1
            if 1 <= z <= g_eff:
                above = ~((2 << z) - 1) & M64
                nz &= above
                valid &= above
                eq &= above & ~(1 << (z + 1))
            zero = ~nz & valid
            lng = zero
            g = 1
            while g <= g_eff and lng:
                lng &= zero >> g
                g += 1
            cover = lng
            for g in range(1, g_eff + 1):
                cover |= (lng << g) & M64
            inr = valid & ~cover
        seed = nz & eq
        k = 1
        while k + 1 < S and seed:
            seed &= (eq << k) & M64
            k += 1
        kept = 0
        while seed:
            sp = ffs(seed) - 1
            up = (((inr + (1 << sp)) & M64) ^ inr) & inr
            below = ~inr & ((1 << sp) - 1)
            first = below.bit_length() if below else 0
            rng_ = up | (((1 << sp) - 1) & ~((1 << first) - 1))
            kept |= rng_
            seed &= ~rng_
        out = Counter()
        ends = ~(kept & nz & eq) & M64
        heads = kept & nz & ~eq
        while heads:
            hp = ffs(heads) - 1
            heads &= heads - 1
            out[ids[hp]] += ffs(ends >> (hp + 1))
        return out

    rnd = random.Random(5)
    for _ in range(30000):
        cnt = rnd.choice([1, 2, 3, 5, 8, 13, 21, 33, 41, 42, 63])
        alpha = rnd.choice([2, 3, 4])
        pz = rnd.choice([0.1, 0.3, 0.5, 0.8])
        ids = [0 if rnd.random() < pz else rnd.randrange(1, alpha + 1) for _ in range(cnt)]
        S = rnd.choice([2, 2, 3, 3, 4, 5, 7])
        G = rnd.choice([0, 1, 1, 2, 3, 5, 70])
        want = Counter(x for x in seedextend(ids, S, G) if x != 0)
        assert closed_form(ids, S, G) == want, (ids, S, G)


def test_oracle_joinkmers_on_a_hand_made_tree():
    """oracle.indexbuild (splitkmers.rs:44-66, joinkmers.rs:53-105) on a tree small enough to check by hand:
    1 (root, no rank) - 2 (superkingdom) - 3 (genus) - {4 (species), 5 (species, invalid)}; 6 (species) under 2."""
    from oracle import indexbuild
    from oracle.taxonomy import Taxonomy
    taxa = [(1, "root", "no rank", 1, True), (2, "sk", "superkingdom", 1, True), (3, "g", "genus", 2, True),
            (4, "s1", "species", 3, True), (5, "s2", "species", 3, False), (6, "s3", "species", 2, True)]
    tax = Taxonomy(taxa)
    rows = [(4, "AAAAAC"), (5, "AAAAAD"), (6, "AAAAA"), (4, "AAAA"), (7, "CCCCC"), (5, "DDDDD")]
    assert indexbuild.splitkmers(rows, 5) == [("AAAAA", 4), ("AAAAC", 4), ("AAAAA", 5), ("AAAAD", 5), ("AAAAA", 6),
                                              ("CCCCC", 7), ("DDDDD", 5)]
    assert indexbuild.splitkmers([(4, "ABCDEF")], 5, "B") == [("CDEF", 4)]
    got = indexbuild.build(rows, tax, 5)
    # AAAAA: taxa 4, 3 (5 is invalid -> its valid ancestor 3) and 6: no child of 2 holds 95 % -> 2
    assert got["AAAAA"] == {2}
    assert got["AAAAC"] == {4}
    assert got["AAAAD"] == {3} and got["DDDDD"] == {3}
    assert "CCCCC" not in got          # taxon 7 is unknown: dropped, nothing left, nothing printed
    assert len(got) == 4


def test_oracle_ranked_seedextend_and_rank_score_ladder():
    """oracle.seedextend ranked mode (seedextend.rs:151-164, taxon.rs:181-191, rank.rs:86-99)."""
    from oracle import seedextend as ose
    from oracle.taxonomy import Taxonomy, RANKS
    # the ladder as written: everything above species scores 12, species and below and "no rank" nothing
    for i, r in enumerate(RANKS):
        want = 12 if 0 < i < RANKS.index("species") else None
        assert ose.rank_score(i) == want, r
    R = {r: i for i, r in enumerate(RANKS)}
    taxa = [(1, "root", 0, 1, True), (2, "sk", R["superkingdom"], 1, True), (3, "g", R["genus"], 2, True),
            (4, "s", R["species"], 3, True), (5, "nr", 0, 3, True), (6, "nr2", 0, 1, True)]
    tax = Taxonomy(taxa)
    assert [ose.taxon_score(tax, t) for t in (0, 1, 2, 3, 4, 5, 6, 7)] == [None, None, 12, 12, None, 12, None, None]
    # two extended seeds: 3 3 3 (36) and 4 4 4 4 (4 x penalty 5 = 20) -> the first; with penalty 9 they tie -> the last
    ids = [3, 3, 3, 0, 0, 4, 4, 4, 4]
    assert ose.seedextend(ids, 2, 0) == [3, 3, 3, 4, 4, 4, 4]
    assert ose.seedextend_ranked(ids, tax, 2, 0, 5) == [3, 3, 3]
    assert ose.seedextend_ranked(ids, tax, 2, 0, 9) == [4, 4, 4, 4]
    assert ose.seedextend_ranked(ids, tax, 2, 0, 10) == [4, 4, 4, 4]
    assert ose.seedextend_ranked([1, 2, 3], tax, 2, 0, 5) == []


def test_pack_reads_host_packer():
    """umgap_pack_reads (host only): 2-bit codes (A, C, G, T = 0..3) and one ascending entry per 16-nucleotide word
    that holds any other byte -- lower case, N, IUPAC, NUL and 0xFF included (dna/mod.rs:34-44) -- plus the padding of
    the last word; capacity errors."""
    import numpy as np
    from umgap_b200 import capi
    rng = np.random.default_rng(0)
    alphabet = np.frombuffer(b"ACGTNacgtRYKM*-\x00\xff", dtype=np.uint8)
    p = np.array([0.2495] * 4 + [0.002 / (len(alphabet) - 4)] * (len(alphabet) - 4))
    lut = np.full(256, 4, dtype=np.uint8)
    for i, c in enumerate(b"ACGT"):
        lut[c] = i
    for n in (0, 1, 15, 16, 17, 1000, 100003, 2_100_000):
        nt = alphabet[rng.choice(len(alphabet), size=n, p=p / p.sum())]
        codes, ent = capi.pack_reads(nt, threads=3)
        nw = (n + 15) // 16
        pad = np.full(nw * 16, 4, dtype=np.uint8)
        pad[:n] = lut[nt]
        pad = pad.reshape(nw, 16)
        want_mask = np.zeros(nw, dtype=np.uint64)
        for j in range(16):
            want_mask |= (pad[:, j] == 4).astype(np.uint64) << np.uint64(j)
            ok = pad[:, j] != 4
            assert np.array_equal(((codes[:nw].astype(np.uint64) >> np.uint64(2 * j)) & np.uint64(3))[ok], pad[:, j][ok].astype(np.uint64))
        nzw = np.nonzero(want_mask)[0].astype(np.uint64)
        assert np.array_equal(ent, (nzw << np.uint64(16)) | want_mask[nzw]), n
    nt = np.frombuffer(b"ACGTN" * 100, dtype=np.uint8)
    with pytest.raises(capi.UmgapError) as e:
        capi.pack_reads(nt, entries=np.zeros(3, dtype=np.uint64))
    assert e.value.code == -5   # UMGAP_ERR_CAPACITY
    # the packer takes 32 nucleotides per step where the host has AVX2 and 8 otherwise (chosen once per process):
    # the other form, in a process of its own, packs the same words
    import subprocess
    import sys
    code = ("import sys, numpy as np; sys.path.insert(0, %r); from umgap_b200 import capi; "
            "rng = np.random.default_rng(3); a = np.frombuffer(b'ACGTNacgtRY\\x00\\xff', dtype=np.uint8); "
            "nt = a[rng.choice(len(a), size=200003, p=[0.24] * 4 + [0.04 / 9] * 9)]; c, e = capi.pack_reads(nt, threads=2); "
            "print(int(c.sum(dtype=np.uint64)), len(e), int(e.sum(dtype=np.uint64)))" % ROOT)
    outs = []
    for scalar in (False, True):
        env = dict(os.environ)
        env.pop("UMGAP_PACK_SCALAR", None)
        if scalar:
            env["UMGAP_PACK_SCALAR"] = "1"
        outs.append(subprocess.run([sys.executable, "-c", code], env=env, check=True, stdout=subprocess.PIPE).stdout)
    assert outs[0] == outs[1] and len(outs[0].split()) == 3


def test_cli_buildindex_printindex_roundtrip(built, tmp_path):
    """`umgap buildindex` (buildindex.rs:32-48) and `umgap printindex` (printindex.rs:38-51) on the host: the doc example
    round trip, the bytes of the doc example against the oracle's codec, and a larger TSV written by the product's
    writer and read back by the three independent readers (product, oracle/fstv2.py, oracle/c)."""
    import random
    from oracle import cport, fstv2
    tsv = b"AAAAA\t2759\nBBBBBB\t9153\n"
    fst = tmp_path / "tiny.index"
    p = subprocess.run([UMGAP, "buildindex"], input=tsv, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert p.returncode == 0
    fst.write_bytes(p.stdout)
    assert p.stdout == fstv2.build([(b"AAAAA", 2759), (b"BBBBBB", 9153)])
    rc, out, err = _run(["printindex", str(fst)])
    assert rc == 0 and out == tsv.decode(), err
    rng = random.Random(3)
    items = {}
    for _ in range(30000):
        L = rng.choice([1, 2, 3, 5, 9, 9, 9, 9, 15, 40])
        items[bytes(rng.choice(b"ACDEFGHIKLMNPQRSTVWY") for _ in range(L))] = rng.choice([0, 1, 7, 300, 2 ** 31, 2 ** 40 + 3, 2759])
    for a in range(33, 127):
        if a != ord('"'):
            items[bytes([ord("Z"), a])] = a      # a node with more than 32 transitions
    items = sorted(items.items())
    big_tsv = b"".join(k + b"\t" + str(v).encode() + b"\n" for k, v in items)
    p = subprocess.run([UMGAP, "buildindex"], input=big_tsv, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert p.returncode == 0, p.stderr.decode()
    big = tmp_path / "big.index"
    big.write_bytes(p.stdout)
    assert list(fstv2.Fst(p.stdout).stream()) == items
    img = cport.FstImage(p.stdout)
    assert all(img.get(k) == v for k, v in rng.sample(items, 3000))
    rc, out, err = _run(["printindex", str(big)])
    assert rc == 0 and out.encode() == big_tsv, err
    # files of the oracle's two writers (fstv2.build minimises: shared suffixes) through printindex
    for data in (fstv2.build(items), cport.fst_build([k for k, _ in items], [v for _, v in items])):
        big.write_bytes(data)
        rc, out, _ = _run(["printindex", str(big)])
        assert rc == 0 and out.encode() == big_tsv
    rc, _, err = _run(["buildindex"], b"B\t1\nA\t2\n")
    assert rc == 1 and "out of order" in err
    rc, _, err = _run(["buildindex"], b"A\t1\nA\t2\n")
    assert rc == 1 and "duplicate" in err
    rc, _, err = _run(["buildindex"], b"A 1\n")
    assert rc == 1


def test_cli_reporting_commands_match_oracle(built, tmp_path):
    """snaptaxon (snaptaxon.rs:66-108), taxa2freq (taxa2freq.rs:86-169) and bestof (bestof.rs:50-79) against their
    line-by-line restatements (oracle/reporting.py) on a synthetic taxonomy."""
    import random
    from oracle import reporting as orep
    from oracle.taxonomy import Taxonomy as OTaxonomy, format_taxon
    import datagen
    taxa = datagen.make_taxonomy(300, seed=5)
    otax = OTaxonomy(taxa)
    tfile = tmp_path / "taxons.tsv"
    tfile.write_bytes(("\n".join(format_taxon(t) for t in taxa) + "\n").encode("latin-1"))
    rng = random.Random(9)
    ids = [t[0] for t in taxa]
    lines = []
    for i in range(400):
        if i % 7 == 0:
            lines.append(f">read{i}")
        lines.append(str(rng.choice(ids)))
    text = "\n".join(lines) + "\n"
    picks = rng.sample(ids, 5)
    for flags, kw in ((["-r", "genus"], dict(rank="genus")), (["-r", "species", "-i"], dict(rank="species", invalid=True)),
                      (["-t", str(picks[0]), "-t", str(picks[1])], dict(taxons=picks[:2])),
                      (["-r", "family", "-t", str(picks[2])], dict(rank="family", taxons=picks[2:3])), ([], {})):
        rc, out, err = _run(["snaptaxon", str(tfile)] + flags, text.encode())
        assert rc == 0, err
        assert out == orep.snaptaxon_text(text, otax, **kw), flags
    rc, _, err = _run(["snaptaxon", str(tfile), "-r", "no rank"], text.encode())
    assert rc == 1
    rc, _, err = _run(["snaptaxon", str(tfile), "-r", "genus"], b"notanumber\n")
    assert rc == 1 and "invalid digit" in err
    # taxa2freq: stdin and two files; rows ordered by descending total, equal totals in any order (HashMap order there)
    other = "".join(str(rng.choice(ids)) + "\n" for _ in range(300))
    (tmp_path / "a.txt").write_text(text)
    (tmp_path / "b.txt").write_text(other)
    for rank, minf in (("species", 1), ("genus", 0), ("phylum", 3)):
        for files, inputs in (([], [text]), ([str(tmp_path / "a.txt"), str(tmp_path / "b.txt")], [text, other])):
            rc, out, err = _run(["taxa2freq", "-r", rank, "-f", str(minf), str(tfile)] + files, text.encode() if not files else b"")
            assert rc == 0, err
            rows, _ = orep.taxa2freq_rows(inputs, otax, rank, minf)
            got = out.split("\n")
            assert got[0] == "taxon id,taxon name," + (",".join(files) if files else "stdin") and got[-1] == ""
            got_rows = [g.split(",") for g in got[1:-1]]
            assert sorted((int(g[0]), g[1], [int(x) for x in g[2:]]) for g in got_rows) == sorted(rows)
            totals = [sum(int(x) for x in g[2:]) for g in got_rows]
            assert totals == sorted(totals, reverse=True)
    rc, _, err = _run(["taxa2freq", "-r", "no rank", str(tfile)], b"")
    assert rc == 1 and "Snap to an actual rank." in err
    # bestof: groups of six frame records (and of three), ids 0 and 1 do not count, ties go to the later record
    recs = []
    for i in range(50):
        for fr in range(6):
            n = rng.choice([0, 1, 3, 3, 8])
            recs.append((f"r{i}|{fr}", [str(rng.choice([0, 1, 1, rng.choice(ids)])) for _ in range(n)]))
    btext = "".join(f">{h}\n" + "".join(x + "\n" for x in s) for h, s in recs)
    for frames in (6, 3, 2):
        rc, out, err = _run(["bestof", "-f", str(frames)], btext.encode())
        assert rc == 0, err
        assert out == orep.bestof_text(btext, frames), frames
    # the doc example of bestof.rs:24-41 (frames 1, 2, 3, 1R, 2R, 3R of one read): frame 1 wins
    doc = (">header1|1\n" + "\n".join("9606 9606 2759 9606 9606 9606 9606 9606 9606 9606 8287".split()) + "\n>header1|2\n" +
           "\n".join("2026807 888268 186802 1598 1883".split()) + "\n>header1|3\n1883\n>header1|1R\n" +
           "\n".join("27342 2759 155619 1133106 38033 2".split()) + "\n>header1|2R\n>header1|3R\n2951\n")
    rc, out, _ = _run(["bestof"], doc.encode())
    assert rc == 0 and out == ">header1|1\n" + "\n".join("9606 9606 2759 9606 9606 9606 9606 9606 9606 9606 8287".split()) + "\n"
