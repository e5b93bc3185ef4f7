"""Pins the oracle against every in-scope known-answer vector the reference holds
(SURVEY Appendix E): unit tests and doc-comment examples, cited per test."""
import itertools
import random

import pytest

from oracle import agg, fasta, fstv2, lookup, pipeline, seedextend, taxonomy, translate

# /root/reference/src/fixtures.rs:4-21
FIXTURE = [
    (1, "root", 0, 1, True),
    (2, "Bacteria", 1, 1, True),
    (10239, "Viruses", 1, 1, True),
    (12884, "Viroids", 1, 1, True),
    (185751, "Pospiviroidae", 19, 12884, True),
    (185752, "Avsunviroidae", 19, 12884, True),
]


@pytest.fixture(scope="module")
def tax():
    return taxonomy.Taxonomy(FIXTURE)


def cagg(ids):
    return agg.count(ids)


# ---- dna/mod.rs:114-135
def test_strand_from():
    assert translate.strand("ACGT*TCGA") == list("ACGTNTCGA")


def test_strand_reversed():
    assert translate.reversed_strand(list("TGCANACGT")) == list("ACGTNTGCA")


def test_strand_frames():
    s = list("ACGT")
    assert s[0:] == list("ACGT") and s[1:] == list("CGT") and s[2:] == list("GT")
    # frame() of a too-short strand is empty (dna/mod.rs:92-98)
    assert translate.translate_frame(1, False, list("AC"), 3) == ""


# ---- dna/translation.rs:205-232
def test_translate_codon():
    aas, starts = translate.get_table(1)
    assert translate.translate_codon(aas, starts, False, "T", "T", "G") == "L"
    assert translate.translate_codon(aas, starts, True, "T", "T", "G") == "M"


def test_translation_table_parsing():
    for i in range(1, 24):
        if i in (7, 8, 17, 18, 19, 20):
            with pytest.raises(translate.UnknownTable):
                translate.get_table(i)
        else:
            aas, starts = translate.get_table(i)
            assert len(aas) == 64 and len(starts) == 64


# ---- commands/translate.rs:21-40 (doc example)
def test_translate_doc_example():
    out = pipeline.translate_text(">header1\nGATTACAAA\n", frames=["1", "1R"])
    assert out == ">header1\nDYK\n>header1\nFVI\n"
    out = pipeline.translate_text(">header1\nGATTACAAA\n", frames=["1", "1R"], append_name=True)
    assert out == ">header1|1\nDYK\n>header1|1R\nFVI\n"


def test_translate_all_frames_shapes():
    # a 150-nt read gives 50/49/49/50/49/49 aa
    recs = translate.translate_record("ACGT" * 37 + "AC")
    assert [len(p) for _, p in recs] == [50, 49, 49, 50, 49, 49]
    assert [n for n, _ in recs] == ["1", "2", "3", "1R", "2R", "3R"]
    # N anywhere in a codon -> '-'
    assert translate.translate_record("ACNGGG", frames=["1"])[0][1] == "-G"
    # lowercase is N (dna/mod.rs:34-44)
    assert translate.translate_record("acgGGG", frames=["1"])[0][1] == "-G"


# ---- commands/prot2kmer.rs:21-35, prot2kmer2lca.rs:32-59 (shape)
def test_kmers_doc_example():
    pep = "DAIGDVAKAYKKAG*S"
    kmers = [pep[i:i + 9] for i in range(len(pep) - 8)]
    assert kmers == ["DAIGDVAKA", "AIGDVAKAY", "IGDVAKAYK", "GDVAKAYKK", "DVAKAYKKA",
                     "VAKAYKKAG", "AKAYKKAG*", "KAYKKAG*S"]
    vals = [571525, 571525, 6920, 6920, 1, 6920]
    idx = lookup.DictIndex({k.encode(): v for k, v in zip(kmers, vals)})
    text = ">header1\n" + pep + "\n"
    assert pipeline.prot2kmer2lca_text(text, idx) == \
        ">header1\n571525\n571525\n6920\n6920\n1\n6920\n"
    assert pipeline.prot2kmer2lca_text(text, idx, one_on_one=True) == \
        ">header1\n571525\n571525\n6920\n6920\n1\n6920\n0\n0\n"
    # short peptides vanish entirely (prot2kmer2lca.rs:172)
    assert pipeline.prot2kmer2lca_text(">a\nDAIGDVAK\n>b\n", idx, one_on_one=True) == ""


# ---- commands/prot2tryp.rs:22-36, prot2tryp2lca.rs:30-40 (shape)
TRYP_IN = "AYKKAGVSGHVWQSDGITNCLLRGLTRVKEAVANRDSGNGYINKVYYWTVDKRATTRDALDAGVDGIMTNYPDVITDVLN"
TRYP_OUT = ["AYK", "K", "AGVSGHVWQSDGITNCLLR", "GLTR", "VK", "EAVANR", "DSGNGYINK", "VYYWTVDK",
            "R", "ATTR", "DALDAGVDGIMTNYPDVITDVLN"]


def test_tryptic_doc_example():
    assert lookup.tryptic_digest_regex(TRYP_IN) == TRYP_OUT
    assert lookup.tryptic_digest(TRYP_IN) == TRYP_OUT
    kept = lookup.tryptic_filter(TRYP_OUT)
    assert kept == ["AGVSGHVWQSDGITNCLLR", "EAVANR", "DSGNGYINK", "VYYWTVDK",
                    "DALDAGVDGIMTNYPDVITDVLN"]
    idx = lookup.DictIndex({b"AGVSGHVWQSDGITNCLLR": 571525, b"EAVANR": 1, b"DSGNGYINK": 571525,
                            b"VYYWTVDK": 6920})
    assert pipeline.prot2tryp2lca_text(">header1\n" + TRYP_IN + "\n", idx) == \
        ">header1\n571525\n1\n571525\n6920\n"


def test_tryptic_closed_form_matches_regex():
    import random
    rnd = random.Random(7)
    for _ in range(20000):
        s = "".join(rnd.choice("KRPA*G") for _ in range(rnd.randint(0, 24)))
        assert lookup.tryptic_digest(s) == lookup.tryptic_digest_regex(s), s


# ---- commands/seedextend.rs:26-50 (doc example)
def test_seedextend_doc_example():
    inp = {
        "header1|1": [9606, 9606, 2759, 9606, 9606, 9606, 9606, 9606, 9606, 9606, 8287],
        "header1|2": [2026807, 888268, 186802, 1598, 1883],
        "header1|3": [1883],
        "header1|1R": [27342, 2759, 155619, 1133106, 38033, 2],
        "header1|2R": [],
        "header1|3R": [2951],
    }
    text = "".join(fasta.write_record(h, [str(x) for x in v], "\n") for h, v in inp.items())
    out = pipeline.seedextend_text(text)
    exp = ">header1|1\n" + "".join(f"{x}\n" for x in inp["header1|1"]) + \
        ">header1|2\n>header1|3\n>header1|1R\n>header1|2R\n>header1|3R\n"
    assert out == exp


def test_seedextend_quirks():
    # SURVEY Appendix A.4 observed consequences of the verbatim machine
    assert seedextend.seedextend([0, 5, 5, 5], 2, 1) == [5, 5]
    assert seedextend.seedextend([0, 5, 5], 2, 1) == []
    assert seedextend.seedextend([0, 0, 5, 5, 5], 2, 1) == [5, 5, 5]
    assert seedextend.seedextend([7, 7, 7, 0, 1, 1], 3, 1) == [7, 7, 7, 0, 1, 1]
    assert seedextend.seedextend([5, 5, 0, 6, 6], 2, 0) == [5, 5, 6, 6]
    assert seedextend.seedextend([5, 5, 0, 6, 6], 2, 1) == [5, 5, 0, 6, 6]
    assert seedextend.seedextend([], 2, 0) == []


# ---- commands/uniq.rs:22-40, fastq2fasta.rs:26-54
def test_uniq_doc_example():
    text = ">header1/1\n147206\n240495\n>header1/2\n1883\n1\n1883\n1883\n"
    assert pipeline.uniq_text(text, "/") == ">header1\n147206\n240495\n1883\n1\n1883\n1883\n"


def test_fastq2fasta_doc_example():
    a = "@header1/1\nGATAAACAAAACACTCATCC\n+\nAAAAAAAAAAAAAAAAAAAA\n@header2/1\nACCC\n+\nAAAA\n"
    b = "@header1/2\nGGGTTT\n+\nAAAAAA\n@header2/2\nTTT\n+\nAAA\n"
    assert fasta.fastq2fasta([a, b]) == \
        ">header1/1\nGATAAACAAAACACTCATCC\n>header1/2\nGGGTTT\n>header2/1\nACCC\n>header2/2\nTTT\n"


# ---- commands/buildindex.rs:20-28
def test_index_round_trip():
    data = fstv2.build([(b"AAAAA", 2759), (b"BBBBBB", 9153)])
    f = fstv2.Fst(data)
    assert list(f.stream()) == [(b"AAAAA", 2759), (b"BBBBBB", 9153)]
    # SURVEY Appendix B self-consistent known answer (55 bytes)
    assert data.hex() == ("0200000000000000" "0000000000000000" "0010ad" "ededed" "0010b1"
                          "f1f1f1f1" "c123c70a0108424112" "02" "0200000000000000"
                          "2600000000000000")


# ---- taxon.rs:417-429
def test_taxon_parsing():
    assert taxonomy.parse_taxon("1\tFelis catus\tspecies\t4\t\x01") == \
        (1, "Felis catus", taxonomy.RANKS.index("species"), 4, True)
    assert taxonomy.parse_taxon("1\tFelis catus\tspecies\t4\t\x00")[4] is False
    for bad in ["hello world", "a\tFelis catus\tspecies\t4\t\x01", "1\tFelis catus\tspecies\tb\t\x01",
                "1\tFelis catus\tspecies\t4\tz", "1\tFelis catus\tnorank\t4\t\x01"]:
        with pytest.raises(taxonomy.TaxonError):
            taxonomy.parse_taxon(bad)


def test_taxon_list_and_root(tax):
    # taxon.rs:449-464
    assert tax.root == 1
    assert tax.parents[185751] == 12884 and tax.parents[1] == 1 and tax.parents[3] is None
    snap = tax.snapping(False)
    assert snap[185751] == 185751 and snap[1] == 1 and snap[3] is None


# ---- agg/mod.rs:82-118 over the in-scope aggregators
def test_empty_singleton_unknown(tax):
    with pytest.raises(agg.EmptyInput):
        agg.lca_star(tax, {})
    with pytest.raises(agg.EmptyInput):
        agg.hybrid(tax, {}, 0.5)
    with pytest.raises(agg.EmptyInput):
        agg.mrtl(tax, {})
    for t in FIXTURE:
        c = cagg([t[0]])
        assert agg.lca_star(tax, c) == t[0]
        for f in (0.0, 0.5, 1.0):
            assert agg.hybrid(tax, c, f) == {t[0]}
        assert agg.mrtl(tax, c) == {t[0]}
    for ids in ([5], [1, 2, 5, 1]):
        for fn in (lambda c: agg.lca_star(tax, c), lambda c: agg.hybrid(tax, c, 0.5),
                   lambda c: agg.mrtl(tax, c)):
            with pytest.raises(taxonomy.UnknownTaxon) as e:
                fn(cagg(ids))
            assert e.value.tid == 5


# ---- tree/lca.rs:51-77
def test_lca_star_vectors(tax):
    L = lambda ids: agg.lca_star(tax, cagg(ids))
    assert L([12884, 185752]) == 185752 and L([185752, 12884]) == 185752
    assert L([1, 2]) == 2 and L([2, 1]) == 2
    assert L([2, 10239]) == 1 and L([10239, 2]) == 1
    assert L([185751, 185752]) == 12884 and L([185752, 185751]) == 12884
    for p in itertools.permutations([12884, 185751, 185752]):
        assert L(list(p)) == 12884


# ---- tree/mix.rs:75-97
def test_hybrid_vectors(tax):
    H = lambda ids, f: agg.hybrid(tax, cagg(ids), f)
    assert H([12884, 185751], 0.0) == {185751}
    assert H([12884, 185751, 185752, 185752], 0.0) == {185752}
    # the reference's own test accepts either child (tree/mix.rs:79): the tie set is both
    assert H([1, 1, 10239, 10239, 12884, 185751, 185752], 0.0) == {185751, 185752}
    assert H([12884, 185751], 1.0) == {185751}
    assert H([12884, 185751, 185752, 185752], 1.0) == {12884}
    assert H([1, 1, 10239, 10239, 10239, 12884, 185751, 185752], 1.0) == {1}
    assert H([12884, 185751], 0.66) == {185751}
    assert H([1, 12884, 12884, 185751], 0.66) == {185751}
    assert H([1, 12884, 10239, 185751, 185751, 185752], 0.66) == {12884}


# ---- rmq/rtl.rs:69-92
def test_mrtl_vectors(tax):
    M = lambda ids: agg.mrtl(tax, cagg(ids))
    assert M([1]) == {1}
    assert M([1, 12884]) == {12884}
    assert M([1, 12884, 185751]) == {185751}
    assert M([1, 1, 1, 185751, 1, 1]) == {185751}
    assert M([1, 1, 185752, 185751, 185751, 1]) == {185751}
    assert M([1, 1, 185752, 185751, 1]) == {185751, 185752}


# ---- rmq/mix.rs:102-126 (`-m rmq -a hybrid`)
def test_rmq_mix_vectors(tax):
    X = lambda ids, f: agg.rmq_mix(tax, cagg(ids), f)
    assert X([12884, 185751], 0.0) == {185751}
    assert X([12884, 185751, 185752, 185752], 0.0) == {185752}
    assert X([1, 1, 10239, 10239, 10239, 12884, 185751, 185752], 0.0) == {10239}
    assert X([12884, 185751], 1.0) == {12884}
    assert X([12884, 185751, 185752, 185752], 1.0) == {12884}
    assert X([1, 1, 10239, 10239, 10239, 12884, 185751, 185752], 1.0) == {1}
    assert X([12884, 12884, 185751], 0.5) == {12884}
    assert X([12884, 185751, 185751], 0.5) == {185751}
    assert X([1, 12884, 12884, 185751, 185752], 0.5) == {12884}   # 3.5 against 2.5: no tie, whatever rmq/mix.rs:119 fears
    with pytest.raises(agg.EmptyInput):
        agg.rmq_mix(tax, {}, 0.5)


# ---- rmq/mod.rs:171-260, rmq/lca.rs:105-172 (`-m rmq -a lca*`)
def test_rmq_structure_and_fold_vectors(tax):
    from oracle import rmq
    arr = [12, 17, 23, 2, 20, 4, 8, 27, 26, 19, 31, 22, 28, 16, 24, 14, 5, 29, 32, 11, 7, 9, 25, 30, 21, 13, 6, 18, 15, 33, 10, 3, 33, 1]
    assert rmq.RMQ(arr).block_min == [33]                     # rmq/mod.rs:177-182 on a 64-bit usize
    assert rmq.euler_tour(tax) == [(1, 0), (2, 1), (1, 0), (10239, 1), (1, 0), (12884, 1), (185751, 2), (12884, 1), (185752, 2),
                                   (12884, 1), (1, 0)]            # taxon.rs:433-446
    rng = random.Random(3)
    for _ in range(100):                                      # query returns a position of the range's minimum
        a = [rng.randrange(0, 5) for _ in range(rng.randrange(1, 300))]
        R = rmq.RMQ(a)
        for _ in range(100):
            i, j = rng.randrange(len(a)), rng.randrange(len(a))
            q = R.query(i, j)
            assert min(i, j) <= q <= max(i, j) and a[q] == min(a[min(i, j):max(i, j) + 1])
    calc = rmq.LCACalculator(tax)
    A = lambda ids: calc.aggregate(list(dict.fromkeys(ids)))
    assert A([12884, 185752]) == 185752 and A([185752, 12884]) == 185752 and A([1, 2]) == 2 and A([2, 1]) == 2
    assert A([2, 10239]) == 1 and A([10239, 2]) == 1 and A([185751, 185752]) == 12884 and A([185752, 185751]) == 12884
    for p in itertools.permutations([12884, 185751, 185752]):
        assert A(list(p)) == 12884
    large = taxonomy.Taxonomy([(i, "", 0, p, True) for i, p in [(1, 1), (2, 1), (5, 2), (6, 2), (3, 1), (7, 3), (10, 7), (13, 10), (14, 13),
                                                              (15, 3), (8, 3), (11, 8), (12, 8), (9, 3), (4, 1)]])
    lc = rmq.LCACalculator(large)
    assert lc.aggregate([9, 7]) == 3 and lc.aggregate([9, 10]) == 3 and lc.aggregate([7, 9]) == 3 and lc.aggregate([14, 8]) == 3


def test_rmq_fold_is_the_tree_form_lca_star():
    """Why the product answers `-m rmq -a lca*` with the kernel of `-m tree -a lca*`: over every order of a record's
    distinct taxa (the reference folds HashMap keys) the fold of rmq/lca.rs:60-90, restated with its RMQ as written,
    gives one answer, and it is tree/lca.rs:34-40's."""
    from oracle import rmq
    import datagen
    rng = random.Random(77)
    checked = 0
    for seed in range(24):
        taxa = datagen.make_taxonomy(rng.choice([8, 30, 120, 400]), seed=100 + seed)
        tx = taxonomy.Taxonomy(taxa)
        calc = rmq.LCACalculator(tx)
        ids = [t[0] for t in taxa]
        for _ in range(40):
            path = tx.root_path(rng.choice(ids))
            keys = list({rng.choice(path) if rng.random() < 0.6 else rng.choice(ids) for _ in range(rng.randrange(1, 7))})
            got = {calc.aggregate(list(p)) for p in itertools.permutations(keys)}
            assert got == {agg.lca_star(tx, {t: agg.f32(1) for t in keys})}, keys
            checked += 1
    assert checked == 960


# ---- taxa2agg -s (taxa2agg.rs:141-148, agg/mod.rs:27-44)
def test_scored_counts_and_filter(tax):
    c = agg.count_scored([(185751, 0.5), (12884, 0.25), (185751, 0.75), (0, 9.0)])
    assert c[185751] == agg.f32(1.25) and c[12884] == agg.f32(0.25)
    snap = tax.snapping(False)
    rec = [(185751, 0.5), (185752, 0.25), (185751, 0.75), (0, 9.0)]
    assert agg.taxa2agg_record_scored(tax, snap, rec, agg.LCA_STAR) == {12884}
    assert agg.taxa2agg_record_scored(tax, snap, rec, agg.LCA_STAR, lower_bound=1.0) == {185751}   # 185752 sums to 0.25
    assert agg.taxa2agg_record_scored(tax, snap, rec, agg.LCA_STAR, lower_bound=2.0) == {1}       # nothing left: literal 1
    assert agg.taxa2agg_record_scored(tax, snap, rec, agg.MRTL) == {185751}
    assert agg.taxa2agg_record_scored(tax, snap, [(t, 1.0) for t in (12884, 185751)], agg.HYBRID, 0.0) == {185751}


# ---- commands/taxa2agg.rs:159-181
def test_taxa2agg_record_loop(tax):
    out = pipeline.taxa2agg_sets(">a\n185751\n0\n185751\n12884\n>b\n0\n0\n>c\n", tax, agg.LCA_STAR)
    assert out == [("a", {185751}), ("b", {1}), ("c", {1})]
    # lower bound drops singletons (agg/mod.rs:39-44)
    out = pipeline.taxa2agg_sets(">a\n185751\n185751\n185752\n", tax, agg.LCA_STAR, lower_bound=2)
    assert out == [("a", {185751})]


def test_oracle_reproduces_committed_golden_fixture():
    """tests/golden/pipeline_small.json (frozen oracle outputs) still matches the oracle."""
    import importlib.util
    import json
    import os
    here = os.path.dirname(os.path.abspath(__file__))
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(here, "golden", "make_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    frozen = json.load(open(os.path.join(here, "golden", "pipeline_small.json")))
    fresh = json.loads(json.dumps(mod.build()))
    assert fresh == frozen
