"""The C restatement (oracle/c) against the line-by-line Python oracle and the reference's own
known-answer vectors.  CPU only."""
import random

import numpy as np
import pytest

import datagen
from oracle import agg as oagg, cport, fstv2, lookup as olookup, pipeline as opipe, seedextend as ose, translate as otr
from oracle.taxonomy import Taxonomy as OTaxonomy


@pytest.fixture(scope="module")
def world():
    taxa = datagen.make_taxonomy(300, seed=41)
    otax = OTaxonomy(taxa)
    proteins = datagen.make_proteome(60, seed=42)
    index = datagen.make_index(proteins, otax, seed=43)
    return dict(taxa=taxa, otax=otax, proteins=proteins, index=index, ctax=cport.RefTaxonomy(taxa))


def test_fst_buildindex_example_bytes():
    # SURVEY Appendix B known answer for `AAAAA 2759`, `BBBBBB 9153` (buildindex.rs:20-28)
    data = cport.fst_build([b"AAAAA", b"BBBBBB"], [2759, 9153])
    assert data == fstv2.build([(b"AAAAA", 2759), (b"BBBBBB", 9153)])
    img = cport.FstImage(data)
    assert img.get(b"AAAAA") == 2759 and img.get(b"BBBBBB") == 9153
    assert img.get(b"AAAA") is None and img.get(b"AAAAAA") is None and img.get(b"") is None


def test_fst_c_and_python_codecs_agree(world):
    items = sorted(world["index"].items())
    extra = [(b"A", 7), (b"AC", 2 ** 40 + 5), (b"ACD", 0), (bytes(range(60, 110)), 12345), (b"\xff\xfe", 1)]
    allitems = sorted(dict(items[:6000] + extra).items())
    c_img = cport.fst_build([k for k, _ in allitems], [v for _, v in allitems])
    py = fstv2.Fst(c_img)                      # Python reader on the C-built image
    assert list(py.stream()) == allitems
    py_img = fstv2.build(allitems)
    cimg_on_py = cport.FstImage(py_img)         # C reader on the Python-built (minimised) image
    cimg = cport.FstImage(c_img)
    rng = random.Random(1)
    for k, v in rng.sample(allitems, 1500):
        assert cimg.get(k) == v and cimg_on_py.get(k) == v
    for k, _ in rng.sample(allitems, 500):
        miss = k[:-1] + bytes([k[-1] ^ 1])
        if miss not in dict(allitems):
            assert cimg.get(miss) is None and cimg_on_py.get(miss) is None
    with pytest.raises(ValueError):
        cport.fst_build([b"B", b"A"], [1, 2])


def test_fst_wide_nodes():
    # nodes with more than 32 transitions use the 256-byte index
    keys = sorted(bytes([a, b]) for a in range(1, 200, 3) for b in (5, 9))
    vals = [i * 77 + 1 for i in range(len(keys))]
    img = cport.fst_build(keys, vals)
    assert list(fstv2.Fst(img).stream()) == list(zip(keys, vals))
    c = cport.FstImage(img)
    assert all(c.get(k) == v for k, v in zip(keys, vals))


def test_translate_matches_python_oracle():
    rng = random.Random(2)
    for _ in range(200):
        n = rng.choice([0, 1, 2, 3, 4, 26, 27, 100, 151])
        s = "".join(rng.choice("ACGTACGTNacgtX") for _ in range(n))
        table = rng.choice([1, 2, 11, 23])
        meth = rng.random() < 0.5
        want = otr.translate_record(s, table, meth)
        for i, (_, pep) in enumerate(want):
            assert cport.translate(s.encode(), i, table, meth) == pep
    assert cport.translate(b"GATTACAAA", 0) == "DYK" and cport.translate(b"GATTACAAA", 3) == "FVI"
    with pytest.raises(ValueError):
        cport.translate(b"ACG", 0, 7)


def test_seedextend_matches_python_oracle():
    rng = random.Random(3)
    for _ in range(3000):
        L = rng.choice([0, 1, 2, 3, 5, 8, 13, 42])
        pool = [0, 0, 0] + [rng.randrange(1, 5) for _ in range(3)]
        ids, cur = [], 0
        for _ in range(L):
            if rng.random() < 0.45:
                cur = rng.choice(pool)
            ids.append(cur)
        for s in (2, 3):
            for g in (0, 1, 2):
                assert cport.seedextend(ids, s, g) == ose.seedextend(ids, s, g), (ids, s, g)


def test_aggregate_within_python_oracle_sets(world):
    rng = random.Random(4)
    otax, ctax = world["otax"], world["ctax"]
    ids = [t[0] for t in otax.by_id if t is not None]
    exact = 0
    for _ in range(400):
        home = rng.choice(ids)
        path = otax.root_path(home)
        rec = []
        for _ in range(rng.choice([0, 1, 2, 5, 12, 60, 300])):
            u = rng.random()
            rec.append(0 if u < 0.2 else home if u < 0.6 else rng.choice(path) if u < 0.85 else rng.choice(ids))
        for strategy in (0, 1, 2):
            for lb in (0.0, 2.0):
                for ranked in (False, True):
                    want = oagg.taxa2agg_record(otax, otax.snapping(ranked), rec, strategy, 0.25, lb)
                    got = cport.aggregate(ctax, rec, strategy, 0.25, lb, ranked)
                    assert got in want, (rec, strategy, lb, ranked, got, want)
                    exact += len(want) == 1
    assert exact > 2000
    with pytest.raises(KeyError):
        cport.aggregate(ctax, [max(ids) + 7], 0)


def test_pipeline_matches_python_oracle(world):
    reads = datagen.make_reads(world["proteins"], 60, seed=44)
    reads += [("s0/1", "ACGT"), ("s0/2", "ACGTACGTACGTACGTACGTACGTAC"), ("s1/1", "ACG" * 9), ("s1/2", "N" * 40)]
    items = sorted(world["index"].items())
    img = cport.FstImage(cport.fst_build([k for k, _ in items], [v for _, v in items]))
    nt = np.frombuffer("".join(r[1] for r in reads).encode(), dtype=np.uint8)
    off = np.zeros(len(reads) + 1, dtype=np.uint64)
    off[1:] = np.cumsum([len(r[1]) for r in reads])
    goff = np.arange(0, len(reads) + 1, 2, dtype=np.uint64)
    for strategy, s, g, lb, use_se in [(0, 2, 0, 0.0, 1), (1, 3, 0, 0.0, 1), (2, 2, 1, 1.0, 1), (1, 2, 0, 2.0, 0)]:
        opts = cport.RefOpts(table=1, methionine=0, one_on_one=1, seedextend=use_se, min_seed_size=s, max_gap_size=g,
                             strategy=strategy, factor=0.25, lower_bound=lb, ranked_only=0, k=9)
        got, nl, nh = cport.classify(img, world["ctax"], opts, nt, off, goff, threads=3)
        want = dict(opipe.classify_reads(reads, olookup.DictIndex(world["index"]), world["otax"],
                                         use_seedextend=bool(use_se), min_seed_size=s, max_gap_size=g,
                                         strategy=strategy, factor=0.25, lower_bound=lb))
        for gi in range(len(goff) - 1):
            h = reads[2 * gi][0].split("/")[0]
            if h in want:
                assert int(got[gi]) in want[h], (h, int(got[gi]), want[h])
            else:
                assert int(got[gi]) == 0xFFFFFFFF
        assert nl == sum(2 * (len(r[1]) - 26) for r in reads if len(r[1]) >= 27) and nh > 0


def test_peptide_pipeline_matches_python_oracle(world):
    """ref_classify_peptides (the CPU baseline of the tryptic presets) against the oracle's text pipeline
    prot2tryp2lca | uniq -d / | taxa2agg, with the digest's edge cases (KP, trailing K, '*', empty lines)."""
    import random
    rng = random.Random(91)
    tryp = {}
    ids = [t[0] for t in world["taxa"]]
    for p in world["proteins"]:
        for pep in olookup.tryptic_filter(olookup.tryptic_digest(p), 5, 50):
            tryp.setdefault(pep.encode(), rng.choice(ids))
    items = sorted(tryp.items())
    img = cport.FstImage(cport.fst_build([k for k, _ in items], [v for _, v in items]))
    lines, heads = [], []
    for i, p in enumerate(world["proteins"][:50]):
        a = rng.randrange(0, len(p) // 2)
        for m, piece in enumerate((p[a:a + 70], p[a + 20:a + 110]), 1):
            heads.append(f"g{i}/{m}")
            lines.append(piece)
    extra = ["", "K", "KP", "AAAAKPAAAAK", "*K*", "AAAAAK*RRRRRP", "PEPTIDEK" * 9]
    for j, e in enumerate(extra):
        heads.append(f"x{j}/1")
        lines.append(e)
    text = "".join(f">{h}\n{l}\n" if l else f">{h}\n" for h, l in zip(heads, lines))
    # groups as uniq -d / forms them; a record without item lines contributes no line (fasta.rs:38-67)
    keep_lines = list(zip(heads, lines))
    ne_lines = [l for l in lines if l]
    aa = np.frombuffer("".join(ne_lines).encode(), dtype=np.uint8)
    off = np.zeros(len(ne_lines) + 1, dtype=np.uint64)
    off[1:] = np.cumsum([len(l) for l in ne_lines])
    group_names, goff, cnt = [], [0], 0
    for h, l in keep_lines:
        g = h.split("/")[0]
        if not group_names or group_names[-1] != g:
            if group_names:
                goff.append(cnt)
            group_names.append(g)
        cnt += 1 if l else 0
    goff.append(cnt)
    goff = np.array(goff, dtype=np.uint64)
    for minlen, maxlen, strategy, lb, keep, drop in [(5, 50, 2, 1.0, b"", b""), (9, 45, 2, 1.0, b"", b""), (9, 45, 2, 5.0, b"", b""),
                                                     (5, 50, 1, 0.0, b"L", b"W"), (5, 50, 0, 0.0, b"", b"")]:
        opts = cport.RefTrypOpts(minlen=minlen, maxlen=maxlen, keep=keep, drop=drop, strategy=strategy, factor=0.25,
                                 lower_bound=lb, ranked_only=0)
        got, nl, nh = cport.classify_peptides(img, world["ctax"], opts, aa, off, goff, threads=3)
        t = opipe.prot2tryp2lca_text(text, olookup.DictIndex(tryp), False, minlen, maxlen, keep.decode(), drop.decode())
        want = opipe.taxa2agg_sets(opipe.uniq_text(t, "/"), world["otax"], strategy, 0.25, lb)
        assert [h for h, _ in want] == group_names
        for gi, (h, adm) in enumerate(want):
            if goff[gi + 1] == goff[gi]:
                assert adm == {1}    # the reference prints the root for a record without peptides
                continue
            assert int(got[gi]) in adm, (h, int(got[gi]), adm)
        assert nh > 20


def test_staged_text_pipeline_matches_python_oracle(world):
    """ref_pipeline_staged (the reference's five-process structure: text between the stages, only the lookups
    multi-threaded -- the `reference_structure` CPU figure of bench.py) prints what the oracle's text pipeline prints."""
    reads = datagen.make_reads(world["proteins"], 150, seed=45) + [("s0/1", "ACGT"), ("s0/2", "ACG" * 9), ("s1/1", "N" * 40)]
    items = sorted(world["index"].items())
    img = cport.FstImage(cport.fst_build([k for k, _ in items], [v for _, v in items]))
    fa = "".join(f">{h}\n{s}\n" for h, s in reads).encode()
    for strategy, s, g, lb in [(1, 3, 0, 0.0), (2, 2, 1, 1.0), (0, 3, 1, 2.0)]:
        opts = cport.RefOpts(table=1, methionine=0, one_on_one=1, seedextend=1, min_seed_size=s, max_gap_size=g,
                             strategy=strategy, factor=0.25, lower_bound=lb, ranked_only=0, k=9)
        out, stage_s, nl = cport.pipeline_staged(img, world["ctax"], opts, fa, threads=3)
        want = opipe.classify_reads(reads, olookup.DictIndex(world["index"]), world["otax"], min_seed_size=s, max_gap_size=g,
                                    strategy=strategy, factor=0.25, lower_bound=lb)
        assert len(out) == len(want) and len(stage_s) == 5
        for (h, adm), o in zip(want, out):
            assert int(o) in adm, (h, int(o), adm)
        assert nl == sum(2 * (len(r[1]) - 26) for r in reads if len(r[1]) >= 27)


def test_c_workload_generator_matches_numpy_mirror():
    """ref_synth_fst / ref_synth_reads (bench.py's CPU legs) against oracle/synth.py, which the GPU tests compare with the
    device generator bit for bit: the same fst image (keys, LCA-merged values) and the same reads."""
    from oracle import synth
    taxa = datagen.make_taxonomy(500, seed=1)
    pre = synth.Preorder(taxa)
    for n_prot, plen in ((700, 408), (40, 60)):
        keys, vals = synth.build_index(2, n_prot, plen, 70, 20, pre)
        want = cport.fst_build_blob(keys.reshape(-1), np.arange(0, 9 * len(keys) + 1, 9, dtype=np.uint64), vals)
        got, nkeys = cport.synth_fst(2, n_prot, plen, 70, 20, pre, threads=3)
        assert nkeys == len(keys) and got == want
    for first, n, L, pct in ((0, 64, 150, 70), (12345, 33, 101, 50), (7, 5, 27, 100)):
        assert np.array_equal(synth.reads(2, 700, 408, 3, first, n, L, pct), cport.synth_reads(2, 700, 408, 3, first, n, L, pct, threads=3))
