"""GPU parity: every C-ABI stage and the fused path against the CPU oracle, bit-exact.

Runs on the B200 box only (`-m gpu`).  Hybrid / MRTL answers are compared against the oracle's
admissible SET (the reference breaks ties by hash iteration order): equality
whenever the set is a singleton, membership otherwise.
"""
import os
import random

import numpy as np
import pytest

from oracle import agg as oagg
from oracle import fstv2, lookup as olookup, pipeline as opipe, seedextend as ose, translate as otr
from oracle.taxonomy import Taxonomy as OTaxonomy

import datagen

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def capi():
    import umgap_b200.capi as c
    if c.device_count() <= 0:
        pytest.fail("no CUDA device: the gpu-marked tests must run on the B200 box")
    return c


@pytest.fixture(scope="module")
def world(capi):
    taxa = datagen.make_taxonomy(400, seed=11)
    otax = OTaxonomy(taxa)
    proteins = datagen.make_proteome(120, seed=12)
    index = datagen.make_index(proteins, otax, seed=13)
    gtax = capi.Taxonomy.from_arrays(*datagen.taxonomy_arrays(taxa))
    keys = sorted(index)
    gidx = capi.Index.from_pairs(keys, [index[k] for k in keys], k=9)
    return dict(taxa=taxa, otax=otax, proteins=proteins, index=index, gtax=gtax, gidx=gidx)


def _random_reads(rng, n, maxlen=200):
    out = []
    for i in range(n):
        L = rng.choice([0, 1, 2, 3, 5, 26, 27, 28, 29, 30, 31, 32, 33, 100, 150, 151, 152, maxlen, 300, 701])
        s = "".join(rng.choice("ACGTACGTACGTNacgtRY*") for _ in range(L))
        out.append(s)
    return out


def test_translate_all_tables_and_frames(capi):
    rng = random.Random(5)
    reads = _random_reads(rng, 60)
    nt, off = capi.pack_strings([r.encode() for r in reads])
    for table in (1, 2, 4, 11, 23):
        for meth in (False, True):
            for mask in (0x3F, 0x01, 0x28, 0x15):
                aa, aa_off = capi.translate(nt, off, table, meth, mask)
                frames = [f for i, f in enumerate(otr.FRAME_NAMES) if mask >> i & 1]
                j = 0
                for r in reads:
                    for name, pep in otr.translate_record(r, table, meth, frames):
                        got = bytes(aa[int(aa_off[j]):int(aa_off[j + 1])]).decode()
                        assert got == pep, (table, meth, mask, r, name)
                        j += 1
                assert j == len(aa_off) - 1


def test_translate_known_answers(capi):
    # translate.rs:21-40 (GATTACAAA -> DYK / FVI) and dna/translation.rs:205-209 (TTG -> L, -m M)
    nt, off = capi.pack_strings([b"GATTACAAA", b"TTG"])
    aa, aa_off = capi.translate(nt, off, 1, False, 0x09)
    peps = [bytes(aa[int(aa_off[i]):int(aa_off[i + 1])]) for i in range(4)]
    assert peps[0] == b"DYK" and peps[1] == b"FVI" and peps[2] == b"L"
    aa, aa_off = capi.translate(nt, off, 1, True, 0x01)
    assert bytes(aa[int(aa_off[1]):int(aa_off[2])]) == b"M"
    with pytest.raises(capi.UmgapError) as e:
        capi.translate(nt, off, 7, False, 0x3F)
    assert "Unknown table" in str(e.value)


def test_kmer_lookup_matches_oracle(capi, world):
    rng = random.Random(6)
    peps = []
    for p in world["proteins"][:40]:
        s = list(p)
        for _ in range(3):
            s[rng.randrange(len(s))] = rng.choice("*-Xa")
        peps.append("".join(s))
    peps += ["", "ACD", "ACDEFGHI", "ACDEFGHIK", world["proteins"][0][:9], world["proteins"][1][5:30]]
    oidx = olookup.DictIndex(world["index"])
    aa, off = capi.pack_strings([p.encode() for p in peps])
    for one in (True, False):
        taxa, toff, kept = capi.kmer_lookup(world["gidx"], aa, off, one)
        expect = {h: ids for h, ids in olookup.prot2kmer2lca([(str(i), [p]) for i, p in enumerate(peps)], oidx, 9, one)}
        hits = 0
        for i, p in enumerate(peps):
            got = [int(x) for x in taxa[int(toff[i]):int(toff[i + 1])]]
            if str(i) in expect:
                assert kept[i] == 1
                assert got == expect[str(i)], (one, p)
                hits += sum(1 for g in got if g)
            else:
                assert kept[i] == 0 and got == []
        assert hits > 500


def test_index_build_stats_and_misses(capi, world):
    info = world["gidx"].info()
    assert info.n_keys == len(world["index"])
    assert info.k == 9 and info.n_buckets >= 1 << 17
    # keys absent from the index are misses, also when they share 8 of 9 residues with a key
    rng = random.Random(7)
    keys = list(world["index"])
    probes = []
    for key in rng.sample(keys, 300):
        s = bytearray(key)
        s[rng.randrange(9)] = ord(rng.choice(datagen.AAS))
        probes.append(bytes(s))
    aa, off = capi.pack_strings(probes)
    taxa, toff, _ = capi.kmer_lookup(world["gidx"], aa, off, True)
    for i, p in enumerate(probes):
        assert int(taxa[int(toff[i])]) == world["index"].get(p, 0)


def test_dense_table_overflow_levels(capi):
    # a load factor of 1.0 forces displaced keys, flagged buckets and overflow levels
    rng = random.Random(8)
    n = 700000
    keys = set()
    while len(keys) < n:
        keys.add(bytes(rng.choice(b"ACDEFGHIKLMNPQRSTVWY") for _ in range(9)))
    keys = sorted(keys)
    vals = [(i * 2654435761) % 1000003 + 1 for i in range(n)]
    idx = capi.Index.from_pairs(keys, vals, k=9, load_factor=1.0)
    info = idx.info()
    assert info.n_keys == n and info.n_displaced > 0 and info.n_flagged > 0
    sample = rng.sample(range(n), 20000)
    probes = [keys[i] for i in sample]
    misses = []
    while len(misses) < 5000:
        m = bytes(rng.choice(b"ACDEFGHIKLMNPQRSTVWY") for _ in range(9))
        if m not in set(probes):
            misses.append(m)
    keyset = set(keys)
    aa, off = capi.pack_strings(probes + misses)
    taxa, toff, _ = capi.kmer_lookup(idx, aa, off, True)
    for j, i in enumerate(sample):
        assert int(taxa[j]) == vals[i]
    for j, m in enumerate(misses):
        if m not in keyset:
            assert int(taxa[len(probes) + j]) == 0
    idx.close()


def test_fst_loader_roundtrip(capi, world, tmp_path):
    items = sorted(world["index"].items())[:20000]
    extra = [(b"AAAAA", 2759), (b"BBBBBB", 9153)]  # buildindex.rs:20-28 keys (other lengths: skipped)
    data = fstv2.build(sorted(items + extra))
    path = tmp_path / "nine.fst"
    path.write_bytes(data)
    idx = capi.Index.load_fst(str(path), k=9)
    info = idx.info()
    assert info.n_keys == len(items) and info.n_skipped == 2
    rng = random.Random(9)
    probe = [k for k, _ in rng.sample(items, 3000)]
    aa, off = capi.pack_strings(probe)
    taxa, toff, _ = capi.kmer_lookup(idx, aa, off, True)
    d = dict(items)
    assert [int(x) for x in taxa] == [d[k] for k in probe]
    idx.close()
    with pytest.raises(capi.UmgapError):
        capi.Index.load_fst(str(tmp_path / "missing.fst"), k=9)


def _random_id_lists(rng, n):
    out = []
    for _ in range(n):
        L = rng.choice([0, 1, 2, 3, 5, 8, 13, 42, 42, 42, 90])
        pool = [0, 0, 0] + [rng.randrange(1, 6) for _ in range(3)]
        ids, cur = [], 0
        for _ in range(L):
            if rng.random() < 0.45:
                cur = rng.choice(pool)
            ids.append(cur)
        out.append(ids)
    return out


def test_seedextend_matches_oracle(capi):
    rng = random.Random(10)
    recs = _random_id_lists(rng, 600)
    recs += [[0, 5, 5, 5], [0, 5, 5], [0, 0, 5, 5, 5], [7, 7, 7, 0, 1, 1], [5, 5, 0, 6, 6],
             [9606, 9606, 2759, 9606, 9606, 9606, 9606, 9606, 9606, 9606, 8287],
             [0, 5, 0, 7], [0, 5, 0, 0, 7, 7], [0, 0, 5, 0, 0, 0, 7]]   # inverted range after a leading gap (-s1 -g1)
    flat = np.array([x for r in recs for x in r], dtype=np.uint32)
    off = np.zeros(len(recs) + 1, dtype=np.uint64)
    off[1:] = np.cumsum([len(r) for r in recs])
    for s in (0, 1, 2, 3, 4):
        for g in (0, 1, 2):
            out, ooff = capi.seedextend(flat, off, s, g)
            for i, r in enumerate(recs):
                got = [int(x) for x in out[int(ooff[i]):int(ooff[i + 1])]]
                assert got == ose.seedextend(r, s, g), (s, g, r)


def _check_agg(capi, world, recs, strategy, factor, lb, ranked):
    otax = world["otax"]
    flat = np.array([x for r in recs for x in r], dtype=np.uint32)
    off = np.zeros(len(recs) + 1, dtype=np.uint64)
    off[1:] = np.cumsum([len(r) for r in recs])
    got = capi.aggregate(world["gtax"], flat, off, strategy, factor, lb, ranked)
    snapping = otax.snapping(ranked)
    singles = 0
    for i, r in enumerate(recs):
        want = oagg.taxa2agg_record(otax, snapping, r, strategy, factor, lb)
        assert int(got[i]) in want, (strategy, factor, lb, ranked, r, int(got[i]), want)
        singles += len(want) == 1
    return singles


def test_ranked_seedextend_matches_oracle(capi, world):
    """umgap_seedextend_ranked (seedextend -r, seedextend.rs:151-164) against the oracle on random id lists over the
    world's taxonomy (ranked, unranked and unknown ids, misses, runs and gaps)."""
    rng = random.Random(202)
    otax = world["otax"]
    known = [t[0] for t in otax.by_id if t is not None]
    unknown = [i for i in range(1, len(otax.by_id)) if otax.by_id[i] is None][:3] + [len(otax.by_id) + 50]
    recs = [[], [0], [0, 0, 0], [known[0]] * 5]
    for _ in range(600):
        pool = [rng.choice(known) for _ in range(rng.randrange(1, 5))] + [0, 0, rng.choice(unknown)]
        rec = []
        for _ in range(rng.randrange(1, 14)):
            rec += [rng.choice(pool)] * rng.randrange(1, 6)
        recs.append(rec)
    flat = np.array([x for r in recs for x in r] or [0], dtype=np.uint32)
    off = np.zeros(len(recs) + 1, dtype=np.uint64)
    np.cumsum([len(r) for r in recs], out=off[1:])
    picked = 0
    for s, g, pen in [(2, 0, 5), (3, 1, 5), (2, 2, 12), (4, 0, 0), (1, 0, 5)]:
        got, goff = capi.seedextend_ranked(world["gtax"], flat, off, s, g, pen)
        plain, poff = capi.seedextend(flat, off, s, g)[:2]
        for i, rec in enumerate(recs):
            want = ose.seedextend_ranked(rec, otax, s, g, pen)
            mine = [int(x) for x in got[int(goff[i]):int(goff[i + 1])]]
            assert mine == want, (s, g, pen, rec, mine, want)
            picked += len(mine) < int(poff[i + 1] - poff[i])
    assert picked > 200      # the ranked mode really drops seeds the plain mode keeps


def test_aggregate_matches_oracle(capi, world):
    rng = random.Random(14)
    otax = world["otax"]
    ids = [t[0] for t in otax.by_id if t is not None]
    recs = [[], [0, 0], [ids[3]], [ids[3]] * 5]
    for _ in range(500):
        home = rng.choice(ids)
        path = otax.root_path(home)
        L = rng.choice([1, 2, 3, 6, 12, 40, 120, 496])
        r = []
        for _ in range(L):
            u = rng.random()
            r.append(0 if u < 0.2 else home if u < 0.6 else rng.choice(path) if u < 0.85 else rng.choice(ids))
        recs.append(r)
    recs.append([rng.choice(ids) for _ in range(3000)])      # beyond the shared-memory list
    recs.append([rng.choice(ids[:40]) for _ in range(1500)])
    total = 0
    for strategy in (capi.AGG_LCA_STAR, capi.AGG_HYBRID, capi.AGG_MRTL):
        for factor in ((0.25, 0.0, 0.5, 1.0, 0.66) if strategy == capi.AGG_HYBRID else (0.25,)):
            for lb in (0.0, 1.0, 2.0, 5.0):
                for ranked in (False, True):
                    total += _check_agg(capi, world, recs, strategy, factor, lb, ranked)
    assert total > 1000


def test_aggregate_reference_fixture(capi):
    # the 6-taxon tree of src/fixtures.rs:4-21 and the vectors of tree/lca.rs:51-77,
    # tree/mix.rs:75-97, rmq/rtl.rs:69-92
    taxa = [(1, "root", 0, 1, True), (2, "Bacteria", 1, 1, True), (10239, "Viruses", 1, 1, True),
            (12884, "Viroids", 1, 1, True), (185751, "Pospiviroidae", 19, 12884, True),
            (185752, "Avsunviroidae", 19, 12884, True)]
    gtax = capi.Taxonomy.from_arrays(*datagen.taxonomy_arrays(taxa))

    def run(ids, strategy, factor=0.25):
        flat = np.array(ids, dtype=np.uint32)
        return int(capi.aggregate(gtax, flat, np.array([0, len(ids)], dtype=np.uint64), strategy, factor)[0])

    assert run([12884, 185751], capi.AGG_LCA_STAR) == 185751          # same path -> deeper
    assert run([185751, 185752], capi.AGG_LCA_STAR) == 12884          # fork -> parent
    assert run([12884, 185751, 185752], capi.AGG_LCA_STAR) == 12884
    assert run([1, 185751, 185752, 2], capi.AGG_LCA_STAR) == 1
    assert run([12884, 185751, 185751, 185752], capi.AGG_HYBRID, 0.0) == 185751
    assert run([12884, 185751, 185751, 185752], capi.AGG_HYBRID, 1.0) == 12884
    assert run([1, 12884, 12884, 185751, 2], capi.AGG_MRTL) == 185751
    assert run([1, 1, 1, 185751, 1, 1], capi.AGG_MRTL) == 185751
    assert run([], capi.AGG_MRTL) == 1
    with pytest.raises(capi.UmgapError) as e:
        run([5], capi.AGG_LCA_STAR)                                   # agg/mod.rs:104-118
    assert "Unknown Taxon ID: 5" in str(e.value)


@pytest.mark.parametrize("strategy,factor,lb,seed_opts", [
    (0, 0.25, 0.0, (1, 2, 0)), (1, 0.25, 0.0, (1, 3, 0)), (2, 0.25, 1.0, (1, 2, 1)),
    (1, 0.25, 1.0, (1, 3, 1)), (0, 0.25, 2.0, (1, 3, 1)), (2, 0.25, 5.0, (0, 0, 0)),
    (1, 0.5, 0.0, (1, 4, 2)),
])
def test_classify_reads_matches_oracle_pipeline(capi, world, strategy, factor, lb, seed_opts):
    use_se, s, g = seed_opts
    reads = datagen.make_reads(world["proteins"], 120, seed=21 + strategy)
    # ragged extras: short reads (dropped records), an N-rich read, a long read, a triple group
    reads += [("s0/1", "ACGT"), ("s0/2", "ACGTACGTACGTACGTACGTACGTAC"), ("s1/1", "ACG" * 9), ("s1/2", "N" * 40),
              ("e0/1", ""), ("e0/2", "")]
    long_src = world["proteins"][3]
    long_nt = "".join(datagen.CODONS.get(a, ["GCT"])[0] for a in long_src)
    reads += [("L0/1", long_nt), ("L0/2", datagen.revcomp(long_nt))]
    # six long reads joined into one group: more kept ids than the shared-memory list holds
    reads += [(f"M0/{i}", long_nt if i % 2 else datagen.revcomp(long_nt)) for i in range(1, 7)]
    # groups with 39 / 68 / 184 distinct kept taxa: more than one per lane, more than the per-group
    # shared-memory table takes (list path through global scratch)
    def nt_of(p):
        return "".join(datagen.CODONS.get(a, ["GCT"])[0] for a in p)
    for name, cnt in (("D2", 2), ("D4", 4), ("D12", 12)):
        reads += [(f"{name}/{i}", nt_of(world["proteins"][10 + i]) if i % 2 else datagen.revcomp(nt_of(world["proteins"][10 + i])))
                  for i in range(cnt)]
    oidx = olookup.DictIndex(world["index"])
    want = opipe.classify_reads(reads, oidx, world["otax"], use_seedextend=bool(use_se), min_seed_size=s,
                                max_gap_size=g, strategy=strategy, factor=factor, lower_bound=lb)
    want = dict(want)
    # groups exactly as `uniq -d /` would form them from the headers
    nt, off = capi.pack_strings([r[1].encode() for r in reads])
    heads = [h.split("/")[0] for h, _ in reads]
    goff = [0]
    for i in range(1, len(reads) + 1):
        if i == len(reads) or heads[i] != heads[i - 1]:
            goff.append(i)
    opts = capi.default_opts(seedextend=use_se, min_seed_size=s, max_gap_size=g, strategy=strategy,
                             factor=factor, lower_bound=lb)
    got, nlook = capi.classify_reads(world["gidx"], world["gtax"], opts, nt, off, np.array(goff, dtype=np.uint64))
    assert nlook == sum(2 * (len(r[1]) - 26) for r in reads if len(r[1]) >= 27)
    non_root = 0
    for gi in range(len(goff) - 1):
        h = heads[goff[gi]]
        if h not in want:
            assert int(got[gi]) == capi.ABSENT, h
            continue
        assert int(got[gi]) in want[h], (h, int(got[gi]), want[h])
        non_root += int(got[gi]) != 1
    assert non_root > (40 if strategy else 3)  # LCA* lands on the root whenever a read carries noise
    assert "e0" not in want and "s0" not in want and "s1" in want and "L0" in want and "M0" in want and "D12" in want


def test_classify_dev_entry_matches_host_entry(capi, world):
    torch = pytest.importorskip("torch")
    reads = datagen.make_reads(world["proteins"], 300, seed=33)
    nt, off = capi.pack_strings([r[1].encode() for r in reads])
    goff = np.arange(0, len(reads) + 1, 2, dtype=np.uint64)
    opts = capi.default_opts(min_seed_size=3, strategy=capi.AGG_HYBRID)
    host, _ = capi.classify_reads(world["gidx"], world["gtax"], opts, nt, off, goff)
    d_nt = torch.from_numpy(nt).cuda()
    d_off = torch.from_numpy(off.astype(np.int64)).cuda()
    d_goff = torch.from_numpy(goff.astype(np.int64)).cuda()
    d_out = torch.zeros(len(goff) - 1, dtype=torch.int32, device="cuda")
    capi.classify_reads_dev(world["gidx"], world["gtax"], opts, d_nt.data_ptr(), d_off.data_ptr(), len(reads),
                            int(off[-1]), d_goff.data_ptr(), len(goff) - 1, d_out.data_ptr(),
                            torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert np.array_equal(d_out.cpu().numpy().view(np.uint32), host)


def test_sliced_device_path_matches_unsliced(capi, world):
    """umgap_classify_reads_dev cuts a large batch into slices on two internal streams (lookup kernel of one slice
    beside the classify kernel of the previous one); same answers as one slice and as the host-buffer entry."""
    torch = pytest.importorskip("torch")
    base = datagen.make_reads(world["proteins"], 400, seed=77, hit_frac=0.8)
    base += [("t/1", "ACGT" * 10), ("t/2", ""), ("u/1", "ACG" * 400), ("u/2", "TTGACC" * 30)]   # ragged, one read beyond a warp batch
    reads = base * 80                                                                                # 32 320 groups
    nt, off = capi.pack_strings([r[1].encode() for r in reads])
    goff = np.arange(0, len(reads) + 1, 2, dtype=np.uint64)
    opts = capi.default_opts(min_seed_size=3, strategy=capi.AGG_HYBRID)
    host, _ = capi.classify_reads(world["gidx"], world["gtax"], opts, nt, off, goff)
    assert np.array_equal(host[:len(base) // 2], host[len(base) // 2:len(base)])
    d_nt = torch.from_numpy(nt).cuda()
    d_off = torch.from_numpy(off.astype(np.int64)).cuda()
    d_goff = torch.from_numpy(goff.astype(np.int64)).cuda()
    before = capi.pipeline_slices(0)
    try:
        for slices in (1, 2, 6, 7):
            capi.pipeline_slices(slices)
            d_out = torch.zeros(len(goff) - 1, dtype=torch.int32, device="cuda")
            for _ in range(2):   # twice: the second call reuses every buffer while the first may still be in flight
                capi.classify_reads_dev(world["gidx"], world["gtax"], opts, d_nt.data_ptr(), d_off.data_ptr(), len(reads),
                                        int(off[-1]), d_goff.data_ptr(), len(goff) - 1, d_out.data_ptr(),
                                        torch.cuda.current_stream().cuda_stream)
            torch.cuda.synchronize()
            assert np.array_equal(d_out.cpu().numpy().view(np.uint32), host), slices
    finally:
        capi.pipeline_slices(before)


def test_synthetic_generators_match_numpy_mirror(capi):
    """The device-side workload generator (bench aid) against oracle/synth.py, bit for bit."""
    torch = pytest.importorskip("torch")
    from oracle import synth
    taxa = datagen.make_taxonomy(300, seed=51)
    pre = synth.Preorder(taxa)
    gtax = capi.Taxonomy.from_arrays(*datagen.taxonomy_arrays(taxa))
    spec = capi.SynthSpec(seed=77, n_proteins=300, protein_len=120, home_pct=70, ancestor_pct=20)
    gidx = capi.Index.build_synthetic(spec, gtax)
    keys, vals = synth.build_index(77, 300, 120, 70, 20, pre)
    assert gidx.info().n_keys == len(keys)
    taxa_out, toff, _ = capi.kmer_lookup(gidx, keys.reshape(-1), np.arange(0, 9 * len(keys) + 1, 9, dtype=np.uint64), True)
    assert np.array_equal(taxa_out.astype(np.uint64), vals)
    for read_len, hit_pct in ((150, 70), (100, 100), (151, 0)):
        npairs = 257
        want = synth.reads(77, 300, 120, 5, 11, npairs, read_len, hit_pct)
        d = torch.zeros(npairs * 2 * read_len, dtype=torch.uint8, device="cuda")
        capi.synth_reads_dev(spec, 5, 11, npairs, read_len, hit_pct, d.data_ptr())
        torch.cuda.synchronize()
        got = d.cpu().numpy().reshape(npairs * 2, read_len)
        assert np.array_equal(got, want), (read_len, hit_pct)
    # the reads really come from the proteome: the pipeline classifies most hit pairs below the root
    reads = synth.reads(77, 300, 120, 5, 0, 200, 150, 100)
    nt = reads.reshape(-1)
    off = np.arange(0, len(nt) + 1, 150, dtype=np.uint64)
    goff = np.arange(0, 401, 2, dtype=np.uint64)
    got, _ = capi.classify_reads(gidx, gtax, capi.default_opts(min_seed_size=3, strategy=capi.AGG_MRTL), nt, off, goff)
    assert (got != 1).mean() > 0.8


def test_tryptic_lookup_matches_oracle(capi, world, tmp_path):
    """prot2tryp2lca: digest + length/keep/drop filters + variable-length lookup (prot2tryp2lca.rs:105-134)."""
    rng = random.Random(61)
    proteins = world["proteins"]
    # tryptic index: digest the proteome, keep 5..50, value = a taxon id
    ids = [t[0] for t in world["otax"].by_id if t is not None]
    tryp = {}
    for p in proteins:
        for pep in olookup.tryptic_filter(olookup.tryptic_digest(p), 5, 50):
            tryp.setdefault(pep.encode(), rng.choice(ids))
    keys = sorted(tryp)
    gidx = capi.Index.from_pairs(keys, [tryp[k] for k in keys], k=0)
    assert gidx.info().n_keys == len(keys) and gidx.info().k == 0
    # same table through the fst loader
    path = tmp_path / "tryp.fst"
    path.write_bytes(fstv2.build([(k, tryp[k]) for k in keys]))
    fidx = capi.Index.load_fst(str(path), k=0)
    assert fidx.info().n_keys == len(keys)
    lines = []
    for p in proteins[:50]:
        s = list(p)
        for _ in range(4):
            s[rng.randrange(len(s))] = rng.choice("*KRPX")
        lines.append("".join(s))
    lines += ["", "*", "K", "KP", "KKKK", "AAAAAKPAAAAAR*", "*" * 7, "MKR" * 30, "A" * 120, proteins[51][:200] + "*" + proteins[52][:150],
              "MVRFKHVQLVKLNSLMFSKEIFTRRVLGYERPLEEIKEAYSKLVHQYHPDRNPNEGRA"]  # shape of prot2tryp.rs:22-36
    aa, off = capi.pack_strings([l.encode() for l in lines])
    oidx = olookup.DictIndex(tryp)
    for idx in (gidx, fidx):
        for one, mn, mx, keep, drop in [(False, 5, 50, "", ""), (True, 5, 50, "", ""), (True, 9, 45, "", ""), (False, 1, 7, "", ""),
                                        (True, 5, 50, "L", ""), (True, 5, 50, "", "CW"), (False, 6, 30, "AE", "P")]:
            taxa, toff = capi.tryp_lookup(idx, aa, off, mn, mx, keep, drop, one)
            for i, line in enumerate(lines):
                want = olookup.prot2tryp2lca([("h", [line])], oidx, one, mn, mx, keep, drop)[0][1]
                got = [int(x) for x in taxa[int(toff[i]):int(toff[i + 1])]]
                assert got == want, (one, mn, mx, keep, drop, line)
    # the digest closed form equals the reference's double regex pass on every test line
    for line in lines:
        assert olookup.tryptic_digest(line) == olookup.tryptic_digest_regex(line)
    with pytest.raises(capi.UmgapError):
        capi.tryp_lookup(world["gidx"], aa, off)   # a k-mer table is not a peptide table


@pytest.mark.parametrize("strategy,lb,mn,mx,keep,drop,ranked", [
    (2, 1.0, 9, 45, "", "", False),     # tryptic-sensitivity: prot2tryp2lca -l9 -L45 | uniq | taxa2agg -l1 -a mrtl
    (2, 5.0, 9, 45, "", "", False),     # tryptic-precision: -l5
    (1, 0.0, 5, 50, "", "", False),     # command defaults, hybrid
    (0, 0.0, 5, 50, "", "CW", True),    # drop set, ranked snapping, LCA*
    (1, 2.0, 6, 30, "L", "", False),    # keep set
    (1, 0.0, 1, 50, "", "", False),     # -l 1 and -l 3: one slot of the per-group hit array per residue / per two residues
    (2, 1.0, 3, 50, "", "", False),
])
def test_fused_peptide_path_matches_oracle_text_pipeline(capi, world, monkeypatch, strategy, lb, mn, mx, keep, drop, ranked):
    """umgap_classify_peptides (digest + lookup + uniq join + aggregation on the device) against the oracle's text
    stages `prot2tryp2lca | uniq -d / | taxa2agg` (scripts/umgap-analyse.sh:291-300), and the device-buffer entry
    point against the host-buffer one."""
    rng = random.Random(77 + strategy)
    proteins = world["proteins"]
    otax = world["otax"]
    ids = [t[0] for t in otax.by_id if t is not None]
    tryp = {}
    for i, p in enumerate(proteins):   # peptides of one protein share a neighbourhood of the tree now and then
        home = rng.choice(ids)
        for pep in olookup.tryptic_filter(olookup.tryptic_digest(p), 5, 50):
            tryp.setdefault(pep.encode(), home if rng.random() < 0.6 else rng.choice(ids))
    keys = sorted(tryp)
    gidx = capi.Index.from_pairs(keys, [tryp[k] for k in keys], k=0)
    # predicted-gene style input: fragments of proteins, two records per fragment pair (`/1`, `/2`), some noise
    recs = []
    for g in range(300):
        p = rng.choice(proteins)
        for mate in (1, 2):
            a = rng.randrange(0, max(1, len(p) - 60))
            frag = list(p[a:a + rng.randrange(20, 140)])
            for _ in range(rng.randrange(0, 3)):
                frag[rng.randrange(len(frag))] = rng.choice("*KRPX")
            recs.append((f"g{g}/{mate}", "".join(frag)))
    recs += [("e0/1", ""), ("e0/2", "*"), ("e1/1", "K"), ("e1/2", "MKR" * 30), ("solo", proteins[3][:200]), ("e2/1", "A" * 120), ("e2/2", "KP")]
    text = "".join(f">{h}\n{sq}\n" if sq else f">{h}\n" for h, sq in recs)
    oidx = olookup.DictIndex(tryp)
    ids_text = opipe.prot2tryp2lca_text(text, oidx, False, mn, mx, keep, drop)
    want = opipe.taxa2agg_sets(opipe.uniq_text(ids_text, "/"), otax, strategy, 0.25, lb, ranked)
    aa, off = capi.pack_strings([sq.encode() for _, sq in recs])
    heads = [h.split("/")[0] for h, _ in recs]
    goff = np.array([0] + [i for i in range(1, len(recs) + 1) if i == len(recs) or heads[i] != heads[i - 1]], dtype=np.uint64)
    opts = capi.tryp_opts(minlen=mn, maxlen=mx, keep=keep, drop=drop, strategy=strategy, factor=0.25, lower_bound=lb,
                          ranked_only=int(ranked))
    got = capi.classify_peptides(gidx, world["gtax"], opts, aa, off, goff)
    assert len(want) == len(got) == len(goff) - 1
    below = 0
    for (h, adm), g_ in zip(want, got):
        assert int(g_) in adm, (h, int(g_), adm)
        below += int(g_) != 1
    assert below > (100 if lb < 5 else 10)
    # the host-buffer call sends the batch in ranges of whole groups on rotating streams: move the seams
    for chunk in ("64", "777", "5000"):
        monkeypatch.setenv("UMGAP_PEP_CHUNK_BYTES", chunk)
        assert np.array_equal(capi.classify_peptides(gidx, world["gtax"], opts, aa, off, goff), got), chunk
    monkeypatch.delenv("UMGAP_PEP_CHUNK_BYTES")
    # device-buffer entry point, asynchronous on the current stream
    import torch
    d_aa = torch.from_numpy(aa).cuda()
    d_off = torch.from_numpy(off.astype(np.int64)).cuda()
    d_goff = torch.from_numpy(goff.astype(np.int64)).cuda()
    d_out = torch.zeros(len(goff) - 1, dtype=torch.int32, device="cuda")
    capi.classify_peptides_dev(gidx, world["gtax"], opts, d_aa.data_ptr(), d_off.data_ptr(), len(off) - 1, int(off[-1]),
                               d_goff.data_ptr(), len(goff) - 1, d_out.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert np.array_equal(d_out.cpu().numpy().view(np.uint32), got)
    # a group without lines has no record; a k-mer table is not a peptide table
    g2 = capi.classify_peptides(gidx, world["gtax"], opts, aa, off, np.array([0, 2, 2, 4], dtype=np.uint64))
    assert int(g2[1]) == capi.ABSENT and int(g2[0]) == int(got[0])
    with pytest.raises(capi.UmgapError):
        capi.classify_peptides(world["gidx"], world["gtax"], opts, aa, off, goff)
    gidx.close()


@pytest.mark.parametrize("k", [9, 5])
def test_index_built_from_proteins_matches_oracle_joinkmers(capi, world, k):
    """umgap_index_build_from_proteins (windows, radix sort, hybrid-0.95 aggregation per k-mer, ranked snapping, table
    insert on the device) against the oracle's `splitkmers | sort | joinkmers`: every k-mer the reference would
    emit is in the table with an admissible value, nothing else is."""
    from oracle import indexbuild
    rng = random.Random(400 + k)
    otax = world["otax"]
    known = [t[0] for t in otax.by_id if t is not None]
    absent = [i for i in range(1, len(otax.by_id)) if otax.by_id[i] is None][:5]
    assert absent
    base = world["proteins"][:60]
    rows = []
    for i, p in enumerate(base):
        home = rng.choice(known)
        rows.append((home, p))
        for _ in range(rng.randrange(0, 4)):     # homologues: shared stretches under related and unrelated taxa
            a = rng.randrange(0, len(p) - 30)
            frag = p[a:a + rng.randrange(20, 120)]
            tid = rng.choice([home, home, rng.choice(known), rng.choice(absent)])
            rows.append((tid, frag + "".join(rng.choice("ACDEFGHIKLMNPQRSTVWY") for _ in range(rng.randrange(0, 12)))))
    rows += [(rng.choice(known), "ACDEFGH"[:k - 1]), (rng.choice(known), ""), (rng.choice(known), "MKX*UB" * 4), (absent[0], base[0])]
    rows += [(rng.choice(known), base[1])] * 3   # the same protein under three more taxa
    want = indexbuild.build(rows, otax, k)
    assert len(want) > 3000
    gidx = capi.Index.build_from_proteins(world["gtax"], [sq.encode() for _, sq in rows], [t for t, _ in rows], k=k)
    info = gidx.info()
    assert info.n_keys == len(want) and info.k == k
    kmers = sorted(want)
    # the same table built in three passes over hash-prefix ranges of the k-mers (protein tables too large to sort at once)
    os.environ["UMGAP_BUILD_PASSES"] = "3"
    try:
        pidx = capi.Index.build_from_proteins(world["gtax"], [sq.encode() for _, sq in rows], [t for t, _ in rows], k=k)
    finally:
        del os.environ["UMGAP_BUILD_PASSES"]
    assert pidx.info().n_keys == len(want)
    aa_k, off_k = capi.pack_strings([x.encode() for x in kmers])
    one_pass, _, _ = capi.kmer_lookup(gidx, aa_k, off_k, True)
    three_pass, _, _ = capi.kmer_lookup(pidx, aa_k, off_k, True)
    assert np.array_equal(one_pass, three_pass)
    pidx.close()
    misses = ["".join(rng.choice("ACDEFGHIKLMNPQRSTVWY") for _ in range(k)) for _ in range(3000)]
    misses = [m for m in misses if m not in want]
    aa, off = capi.pack_strings([x.encode() for x in kmers + misses])
    got, goff, _ = capi.kmer_lookup(gidx, aa, off, True)
    assert len(got) == len(kmers) + len(misses)
    multi = 0
    for kmer, v in zip(kmers, got[:len(kmers)]):
        assert int(v) in want[kmer], (kmer, int(v), want[kmer])
        multi += int(v) != 1
    assert multi > 1000
    assert not np.any(got[len(kmers):])          # -o: a miss is 0
    # a taxon id beyond the taxonomy's id range is an error (an index panic in the reference)
    with pytest.raises(capi.UmgapError):
        capi.Index.build_from_proteins(world["gtax"], [base[0].encode()], [len(otax.by_id) + 1000], k=k)
    gidx.close()


def test_large_batch_matches_c_port(capi, tmp_path):
    """50 000 synthetic pairs against a 2e6-key index: the CUDA path (through the fst loader) and the C
    restatement of the reference algorithm agree on every pair for LCA* (deterministic), and on every
    pair whose hybrid / MRTL answer is unique; the others lie in the Python oracle's admissible set."""
    from oracle import cport, synth
    taxa = datagen.make_taxonomy(2000, seed=81)
    otax = OTaxonomy(taxa)
    pre = synth.Preorder(taxa)
    n_prot, plen = 5000, 408
    keys, vals = synth.build_index(9, n_prot, plen, 70, 20, pre)
    img_bytes = cport.fst_build_blob(keys.reshape(-1), np.arange(0, 9 * len(keys) + 1, 9, dtype=np.uint64), vals)
    path = tmp_path / "synth.fst"
    path.write_bytes(img_bytes)
    gidx = capi.Index.load_fst(str(path), k=9)
    assert gidx.info().n_keys == len(keys)
    gtax = capi.Taxonomy.from_arrays(*datagen.taxonomy_arrays(taxa))
    ctax = cport.RefTaxonomy(taxa)
    img = cport.FstImage(img_bytes)
    npairs = 50000
    reads = synth.reads(9, n_prot, plen, 4, 0, npairs, 150, 70)
    nt = reads.reshape(-1)
    off = np.arange(0, len(nt) + 1, 150, dtype=np.uint64)
    goff = np.arange(0, 2 * npairs + 1, 2, dtype=np.uint64)
    index_dict = None
    for strategy, s, g, lb in [(0, 3, 0, 0.0), (1, 3, 0, 0.0), (2, 2, 1, 1.0), (1, 2, 1, 2.0)]:
        gopts = capi.default_opts(min_seed_size=s, max_gap_size=g, strategy=strategy, factor=0.25, lower_bound=lb)
        got, nlook = capi.classify_reads(gidx, gtax, gopts, nt, off, goff)
        copts = cport.RefOpts(table=1, methionine=0, one_on_one=1, seedextend=1, min_seed_size=s, max_gap_size=g,
                              strategy=strategy, factor=0.25, lower_bound=lb, ranked_only=0, k=9)
        want, nl, nh = cport.classify(img, ctax, copts, nt, off, goff, threads=os.cpu_count() or 1)
        assert nlook == nl == npairs * 2 * 248
        diff = np.nonzero(got != want)[0]
        assert (got != 1).mean() > (0.5 if strategy else 0.05)   # LCA* falls to the root on any unrelated hit
        if strategy == 0:
            assert len(diff) == 0
            continue
        assert len(diff) < npairs * 0.02, len(diff)   # only tie-breaks may differ
        if index_dict is None:
            index_dict = {bytes(k): int(v) for k, v in zip(keys, vals)}
        oidx = olookup.DictIndex(index_dict)
        for gi in diff[:200]:
            pair = [(f"r{gi}/1", bytes(reads[2 * gi]).decode()), (f"r{gi}/2", bytes(reads[2 * gi + 1]).decode())]
            admissible = opipe.classify_reads(pair, oidx, otax, min_seed_size=s, max_gap_size=g, strategy=strategy,
                                              factor=0.25, lower_bound=lb)[0][1]
            assert len(admissible) > 1 and int(got[gi]) in admissible and int(want[gi]) in admissible, (gi, admissible)


def test_full_size_index_properties(capi):
    """BASELINE-size table (1e9 windows, built on the device): membership and values of sampled keys
    re-derived on the host, misses for perturbed keys, and a stable checksum over two builds."""
    torch = pytest.importorskip("torch")
    free, _ = torch.cuda.mem_get_info()
    if free < 40e9:
        pytest.skip("needs 40 GB of free HBM")
    from oracle import synth
    taxa = datagen.make_taxonomy(5000, seed=1)
    pre = synth.Preorder(taxa)
    gtax = capi.Taxonomy.from_arrays(*datagen.taxonomy_arrays(taxa))
    n_prot, plen = 2_500_000, 408
    spec = capi.SynthSpec(seed=2, n_proteins=n_prot, protein_len=plen, home_pct=70, ancestor_pct=20)
    gidx = capi.Index.build_synthetic(spec, gtax)
    info = gidx.info()
    nwin = n_prot * (plen - 8)
    assert 0.99 * nwin < info.n_keys <= nwin             # a few 9-mers occur twice and merge
    assert info.load_factor == 0.5 and info.bytes < 17e9 and info.n_flagged / info.n_buckets < 0.12   # default policy: spare HBM buys a sparser table
    # 2000 proteins sampled across the proteome: every window is a key; its value is the window's own
    # value or (merged duplicates) an ancestor of it
    rng = np.random.default_rng(3)
    sample = np.sort(rng.choice(n_prot, size=500, replace=False))
    exact = total = 0
    for j in sample:
        k, v = synth.windows(2, n_prot, plen, 70, 20, pre, int(j), int(j) + 1)
        taxa_out, toff, _ = capi.kmer_lookup(gidx, k.reshape(-1), np.arange(0, 9 * len(k) + 1, 9, dtype=np.uint64), False)
        assert len(taxa_out) == len(k)                   # no window is missing (misses would be omitted)
        same = taxa_out.astype(np.uint64) == v
        exact += int(same.sum())
        total += len(k)
        for got, own in zip(taxa_out[~same], v[~same]):
            a, b = pre.dense_of[int(got)], pre.dense_of[int(own)]
            assert pre.lca(a, b) == a                    # the stored value is an ancestor of this occurrence's
    assert exact > 0.985 * total                         # ~1 % of the windows share their 9-mer with another one
    # perturbing one residue of a key gives a miss (the key space is 5e11, the table holds 1e9)
    k, _ = synth.windows(2, n_prot, plen, 70, 20, pre, 7, 8)
    k = k.copy()
    k[:, 4] = np.where(k[:, 4] == ord("W"), ord("C"), ord("W"))
    taxa_out, _, _ = capi.kmer_lookup(gidx, k.reshape(-1), np.arange(0, 9 * len(k) + 1, 9, dtype=np.uint64), True)
    assert (taxa_out == 0).mean() > 0.99
    # the classification of device-generated reads is reproducible and mostly below the root
    B = 200_000
    nt = torch.empty(B * 300, dtype=torch.uint8, device="cuda")
    capi.synth_reads_dev(spec, 3, 0, B, 150, 70, nt.data_ptr())
    roff = torch.arange(0, 2 * B + 1, dtype=torch.int64, device="cuda") * 150
    goff = torch.arange(0, 2 * B + 1, 2, dtype=torch.int64, device="cuda")
    outs = []
    for _ in range(2):
        out = torch.zeros(B, dtype=torch.int32, device="cuda")
        capi.classify_reads_dev(gidx, gtax, capi.default_opts(min_seed_size=3), nt.data_ptr(), roff.data_ptr(), 2 * B,
                                B * 300, goff.data_ptr(), B, out.data_ptr(), torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        outs.append(out.cpu().numpy())
    assert np.array_equal(outs[0], outs[1])
    host, _ = capi.classify_reads(gidx, gtax, capi.default_opts(min_seed_size=3), nt.cpu().numpy(),
                                  roff.cpu().numpy().astype(np.uint64), goff.cpu().numpy().astype(np.uint64))
    assert np.array_equal(host, outs[0].view(np.uint32))
    assert 0.6 < (host != 1).mean() < 0.8                # 70 % of the pairs come from the proteome
    # sampled lookups (every min(S,4)-th position first, sliced over two streams) against the every-position kernel,
    # full-size table, the presets of umgap-analyse.sh and the bench configuration: identical taxa for every pair
    for s_, g_, strat, lb in ((3, 0, capi.AGG_HYBRID, 0.0), (2, 1, capi.AGG_MRTL, 1.0), (3, 1, capi.AGG_HYBRID, 1.0),
                              (4, 1, capi.AGG_LCA_STAR, 5.0)):
        o = capi.default_opts(min_seed_size=s_, max_gap_size=g_, strategy=strat, lower_bound=lb)
        got = {}
        for sampling in (1, 0):
            before = capi.pipeline_sampling(sampling)
            try:
                out = torch.zeros(B, dtype=torch.int32, device="cuda")
                capi.classify_reads_dev(gidx, gtax, o, nt.data_ptr(), roff.data_ptr(), 2 * B, B * 300, goff.data_ptr(), B,
                                        out.data_ptr(), torch.cuda.current_stream().cuda_stream)
                torch.cuda.synchronize()
                got[sampling] = out.cpu().numpy()
            finally:
                capi.pipeline_sampling(before)
        assert np.array_equal(got[0], got[1]), (s_, g_, strat)
        assert (got[1] != 1).mean() > 0.3
    gidx.close()


def test_region_partitioned_probing_gives_identical_results(capi, world):
    """Tables beyond the probe-region size are looked up one hash-prefix region per kernel pass; forcing
    4 MiB regions on the 16 MiB test table (4 passes) must not change a single answer."""
    reads = datagen.make_reads(world["proteins"], 400, seed=101)
    reads += [("s/1", "ACGT" * 5), ("s/2", "N" * 90)]
    nt, off = capi.pack_strings([r[1].encode() for r in reads])
    goff = np.arange(0, len(reads) + 1, 2, dtype=np.uint64)
    keys = sorted(world["index"])
    gidx = capi.Index.from_pairs(keys, [world["index"][k] for k in keys], k=9)
    for strategy in (0, 1, 2):
        opts = capi.default_opts(min_seed_size=2, max_gap_size=1, strategy=strategy)
        gidx.set_probe_region(0)
        a, _ = capi.classify_reads(gidx, world["gtax"], opts, nt, off, goff)
        gidx.set_probe_region(4 << 20)
        b, _ = capi.classify_reads(gidx, world["gtax"], opts, nt, off, goff)
        assert np.array_equal(a, b)
        assert (a != 1).sum() > (100 if strategy else 10)
    gidx.close()


@pytest.mark.parametrize("table,meth,one,use_se,s,g,strategy,ranked", [
    (1, False, False, True, 2, 0, 1, False),    # misses omitted before seedextend (no -o)
    (1, False, False, True, 3, 1, 2, False),
    (1, False, False, False, 2, 0, 0, False),   # no seedextend, no -o (tryptic-style aggregation of all hits)
    (11, True, True, True, 2, 1, 1, True),      # bacterial code, -m, ranked snapping
    (4, True, True, False, 2, 0, 2, False),
])
def test_classify_reads_option_matrix(capi, world, table, meth, one, use_se, s, g, strategy, ranked):
    """Every switch of the fused entry point against the oracle's text pipeline."""
    reads = datagen.make_reads(world["proteins"], 150, seed=111 + table)
    reads += [("x/1", "ATG" * 20 + "TTG" * 15), ("x/2", "CTG" * 30 + "NNN" + "GTG" * 9)]   # start codons that -m rewrites
    oidx = olookup.DictIndex(world["index"])
    want = dict(opipe.classify_reads(reads, oidx, world["otax"], table=table, methionine=meth, one_on_one=one,
                                     use_seedextend=use_se, min_seed_size=s, max_gap_size=g, strategy=strategy,
                                     factor=0.25, lower_bound=0.0, ranked_only=ranked))
    nt, off = capi.pack_strings([r[1].encode() for r in reads])
    goff = np.arange(0, len(reads) + 1, 2, dtype=np.uint64)
    opts = capi.default_opts(table=table, methionine=int(meth), one_on_one=int(one), seedextend=int(use_se), min_seed_size=s,
                             max_gap_size=g, strategy=strategy, ranked_only=int(ranked))
    got, _ = capi.classify_reads(world["gidx"], world["gtax"], opts, nt, off, goff)
    below = 0
    for gi in range(len(goff) - 1):
        h = reads[2 * gi][0].split("/")[0]
        assert int(got[gi]) in want[h], (h, int(got[gi]), want[h])
        below += int(got[gi]) != 1
    assert below > 5


def test_committed_golden_fixture(capi):
    """The CUDA path against tests/golden/pipeline_small.json (frozen oracle outputs; the script that
    made them is tests/golden/make_golden.py)."""
    import json
    g = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "pipeline_small.json")))
    ids = np.array([t[0] for t in g["taxa"]], dtype=np.uint64)
    gtax = capi.Taxonomy.from_arrays(ids, np.array([t[2] for t in g["taxa"]], dtype=np.uint64),
                                     np.array([t[1] for t in g["taxa"]], dtype=np.uint8),
                                     np.array([t[3] for t in g["taxa"]], dtype=np.uint8))
    gidx = capi.Index.from_pairs([k.encode() for k, _ in g["index"]], [v for _, v in g["index"]], k=9)
    reads = g["reads"]
    nt, off = capi.pack_strings([r[1].encode() for r in reads])
    heads = [h.split("/")[0] for h, _ in reads]
    goff = [0] + [i for i in range(1, len(reads) + 1) if i == len(reads) or heads[i] != heads[i - 1]]
    for case in g["cases"]:
        o = case["options"]
        opts = capi.default_opts(one_on_one=int(o.get("one_on_one", True)), seedextend=int(o.get("use_seedextend", True)),
                                 min_seed_size=o.get("min_seed_size", 2), max_gap_size=o.get("max_gap_size", 0),
                                 strategy=o["strategy"], factor=o.get("factor", 0.25), lower_bound=o.get("lower_bound", 0.0))
        got, _ = capi.classify_reads(gidx, gtax, opts, nt, off, np.array(goff, dtype=np.uint64))
        want = dict((h, s) for h, s in case["expected"])
        for gi in range(len(goff) - 1):
            h = heads[goff[gi]]
            if h in want:
                assert int(got[gi]) in want[h], (case["name"], h, int(got[gi]), want[h])
            else:
                assert int(got[gi]) == capi.ABSENT, (case["name"], h)


@pytest.mark.parametrize("s,g,strategy", [(2, 0, 0), (3, 0, 1), (3, 2, 2), (4, 1, 1), (7, 0, 1)])
def test_sampled_lookup_ragged_batches(capi, world, s, g, strategy):
    """`-o | seedextend -s S` lets the lookup kernel probe every min(S,4)-th position first and the rest only for
    frames with a hit (pipeline.cu: lookup_sampled_kernel).  Its warp batches mix reads of every length here:
    empty, shorter than one k-mer, 27..33 nt, 100..320 nt with hits on either strand, reads longer than a
    batch (plain path) -- all against the oracle's text pipeline."""
    rng = random.Random(1000 + 7 * s + g)
    prots = world["proteins"]
    reads = []
    gi = 0
    def add(seqs):
        nonlocal gi
        for m, sq in enumerate(seqs, 1):
            reads.append((f"q{gi}/{m}", sq))
        gi += 1
    for L in (100, 126, 150, 151, 152, 153, 250, 301, 320):
        for pair in datagen.make_reads(prots, 12, seed=500 + L + s, read_len=L, hit_frac=0.8)[::1]:
            reads.append(pair)
    # regroup the make_reads pairs under fresh headers, interleaved with the odd sizes
    paired = [(reads[i][1], reads[i + 1][1]) for i in range(0, len(reads), 2)]
    reads.clear()
    rng.shuffle(paired)
    long_nt = "".join(rng.choice(datagen.CODONS.get(a, ["GCT"])) for p in prots[:6] for a in p)   # > 960 nt
    assert len(long_nt) > 1200
    odd = ["", "A", "ACGTACGTACGTACGTACGTACGTAC", "ACG" * 9, "N" * 33, long_nt, datagen.revcomp(long_nt)[:1100],
           long_nt[:961], long_nt[:960], long_nt[5:965], "acgt" * 40]
    for a, b in paired:
        add([a, b])
        if rng.random() < 0.5:
            k = rng.randrange(1, 4)
            add([rng.choice(odd) for _ in range(k)])
    oidx = olookup.DictIndex(world["index"])
    want = dict(opipe.classify_reads(reads, oidx, world["otax"], use_seedextend=True, min_seed_size=s, max_gap_size=g,
                                     strategy=strategy, factor=0.25, lower_bound=0.0))
    nt, off = capi.pack_strings([r[1].encode() for r in reads])
    heads = [h.split("/")[0] for h, _ in reads]
    goff = [0] + [i for i in range(1, len(reads) + 1) if i == len(reads) or heads[i] != heads[i - 1]]
    opts = capi.default_opts(seedextend=1, min_seed_size=s, max_gap_size=g, strategy=strategy)
    got, _ = capi.classify_reads(world["gidx"], world["gtax"], opts, nt, off, np.array(goff, dtype=np.uint64))
    below = 0
    for k in range(len(goff) - 1):
        h = heads[goff[k]]
        if h not in want:
            assert int(got[k]) == capi.ABSENT, h
            continue
        assert int(got[k]) in want[h], (h, int(got[k]), want[h])
        below += int(got[k]) != 1
    assert below > (20 if strategy else 2)
    # the same batch with the table probed in hash-prefix regions: the two phases run as separate launches per region
    try:
        world["gidx"].set_probe_region(4 << 20)
        again, _ = capi.classify_reads(world["gidx"], world["gtax"], opts, nt, off, np.array(goff, dtype=np.uint64))
    finally:
        world["gidx"].set_probe_region(0)
    assert np.array_equal(np.asarray(got), np.asarray(again))


def test_packed_reads_match_byte_form(capi, world, monkeypatch):
    """umgap_classify_reads_packed (2-bit nucleotides + the list of words holding an N, packed on the host by
    umgap_pack_reads) against umgap_classify_reads and the oracle: ragged reads with N, lower case and other bytes,
    chunk seams that fall inside a 16-nucleotide word (UMGAP_CHUNK_NT), the sampled and the every-position kernels,
    a read longer than a sampled batch."""
    rng = random.Random(77)
    prots = world["proteins"]
    reads = []
    for L in (100, 149, 150, 151, 301):
        reads += [(f"L{L}{h}", sq) for h, sq in datagen.make_reads(prots, 40, seed=900 + L, read_len=L, hit_frac=0.8)]
    long_nt = "".join(rng.choice(datagen.CODONS.get(a, ["GCT"])) for p in prots[:5] for a in p)
    odd = _random_reads(rng, 40) + [long_nt[:1500]]
    groups = [reads[i:i + 2] for i in range(0, len(reads), 2)]   # single reads go between the pairs, never inside one
    for i, sq in enumerate(odd):
        groups.insert(rng.randrange(0, len(groups)), [(f"odd{i}/1", sq)])
    reads = [r for g in groups for r in g]
    nt, off = capi.pack_strings([r[1].encode() for r in reads])
    heads = [h.split("/")[0] for h, _ in reads]
    goff = np.array([0] + [i for i in range(1, len(reads) + 1) if i == len(reads) or heads[i] != heads[i - 1]], dtype=np.uint64)
    codes, entries = capi.pack_reads(nt)
    assert len(entries) > 10 and np.all(np.diff(entries >> np.uint64(16)) > 0)
    oidx = olookup.DictIndex(world["index"])
    for chunk_nt in ("0", "777", "5000"):
        monkeypatch.setenv("UMGAP_CHUNK_NT", chunk_nt)
        for kw in (dict(seedextend=1, min_seed_size=3, max_gap_size=0, strategy=capi.AGG_HYBRID),
                   dict(seedextend=1, min_seed_size=2, max_gap_size=1, strategy=capi.AGG_MRTL, lower_bound=1.0),
                   dict(seedextend=0, one_on_one=0, strategy=capi.AGG_LCA_STAR)):
            opts = capi.default_opts(**kw)
            want, nl = capi.classify_reads(world["gidx"], world["gtax"], opts, nt, off, goff)
            got, nl2 = capi.classify_reads_packed(world["gidx"], world["gtax"], opts, codes, entries, off, goff)
            assert nl == nl2
            assert np.array_equal(want, got), (chunk_nt, kw)
            if chunk_nt == "777" and kw["strategy"] == capi.AGG_LCA_STAR:
                ref = dict(opipe.classify_reads(reads, oidx, world["otax"], use_seedextend=False, one_on_one=False, strategy=0))
                for k in range(len(goff) - 1):
                    h = heads[int(goff[k])]
                    assert (int(got[k]) in ref[h]) if h in ref else int(got[k]) == capi.ABSENT, h
    monkeypatch.delenv("UMGAP_CHUNK_NT")
    # no N entries: every byte is one of A, C, G, T
    clean = np.frombuffer(bytes(nt).translate(bytes.maketrans(b"NacgtRY*", b"AACGTAAA")), dtype=np.uint8)
    c2, e2 = capi.pack_reads(clean)
    assert len(e2) == (1 if len(clean) % 16 else 0)   # only the padding of the last word
    opts = capi.default_opts(seedextend=1, min_seed_size=3)
    a, _ = capi.classify_reads(world["gidx"], world["gtax"], opts, clean, off, goff)
    b, _ = capi.classify_reads_packed(world["gidx"], world["gtax"], opts, c2, None, off, goff)
    assert np.array_equal(a, b)


def test_multi_replica_classification_matches_single(capi, world):
    """umgap_classify_reads_multi (index and taxonomy replicated, the groups of a batch cut into one range per replica,
    one host thread each) returns what umgap_classify_reads returns: two and three replicas (on this box's GPUs, or all
    on device 0 when it has one), ragged reads, the byte and the packed form, and an error raised inside one range."""
    rng = random.Random(31)
    reads = []
    for L in (100, 150, 151, 301):
        reads += [(f"L{L}{h}", sq) for h, sq in datagen.make_reads(world["proteins"], 60, seed=700 + L, read_len=L, hit_frac=0.8)]
    groups = [reads[i:i + 2] for i in range(0, len(reads), 2)]
    for i, sq in enumerate(_random_reads(rng, 30)):
        groups.insert(rng.randrange(0, len(groups)), [(f"odd{i}/1", sq)])
    reads = [r for g in groups for r in g]
    nt, off = capi.pack_strings([r[1].encode() for r in reads])
    heads = [h.split("/")[0] for h, _ in reads]
    goff = np.array([0] + [i for i in range(1, len(reads) + 1) if i == len(reads) or heads[i] != heads[i - 1]], dtype=np.uint64)
    ndev = capi.device_count()
    codes, entries = capi.pack_reads(nt)
    for nrep in (2, 3):
        reps = [(world["gidx"], world["gtax"])] + capi.replicate(world["gidx"], world["gtax"], [(i + 1) % ndev for i in range(nrep - 1)])
        for kw in (dict(seedextend=1, min_seed_size=3, strategy=capi.AGG_HYBRID), dict(seedextend=0, one_on_one=0, strategy=capi.AGG_LCA_STAR)):
            opts = capi.default_opts(**kw)
            want, nl = capi.classify_reads(world["gidx"], world["gtax"], opts, nt, off, goff)
            got, nl2 = capi.classify_reads_multi(reps, opts, nt, off, goff)
            assert nl == nl2 and np.array_equal(want, got), (nrep, kw)
            got, _ = capi.classify_reads_multi(reps, opts, None, off, goff, packed=(codes, entries))
            assert np.array_equal(want, got), (nrep, kw, "packed")
        # fewer groups than replicas, and none
        got, _ = capi.classify_reads_multi(reps, opts, nt, off[:3], np.array([0, 2], dtype=np.uint64))
        assert np.array_equal(got, want[:1])
        got, _ = capi.classify_reads_multi(reps, opts, nt, off[:1], np.array([0], dtype=np.uint64))
        assert len(got) == 0
        for i, t in reps[1:]:
            i.close()
            t.close()
    with pytest.raises(capi.UmgapError):
        capi.classify_reads_multi([(world["gidx"], world["gtax"]), (world["gidx"], world["gtax"])], opts, nt, off, goff)


def test_async_host_calls_match_the_synchronous_ones(capi, world):
    """umgap_classify_reads_async / _packed_async + umgap_pending_wait: several batches in flight on one index (their
    chunks share the index's streams and workspaces in stream order) give what the synchronous calls give, complete in
    order, keep each batch's error to that batch (Unknown Taxon ID is raised by the wait of the batch that held it and
    by no other), and the 33rd batch in flight is refused."""
    prots = world["proteins"]
    gidx, gtax = world["gidx"], world["gtax"]
    batches = []
    for b, L in enumerate((150, 100, 251, 150, 301, 120)):
        reads = datagen.make_reads(prots, 40 + 10 * b, seed=7700 + b, read_len=L, hit_frac=0.8)
        nt, off = capi.pack_strings([r[1].encode() for r in reads])
        goff = np.arange(0, len(reads) + 1, 2, dtype=np.uint64)
        batches.append((nt, off, goff, capi.pack_reads(nt)))
    for kw in (dict(min_seed_size=3, strategy=capi.AGG_HYBRID), dict(seedextend=0, strategy=capi.AGG_MRTL),
               dict(min_seed_size=2, max_gap_size=1, strategy=capi.AGG_LCA_STAR, lower_bound=2.0)):
        opts = capi.default_opts(**kw)
        want = [capi.classify_reads(gidx, gtax, opts, nt, off, goff)[0].copy() for nt, off, goff, _ in batches]
        for chunk_nt in (None, "3000"):  # several chunks per batch: the seams of consecutive batches interleave
            if chunk_nt:
                os.environ["UMGAP_CHUNK_NT"] = chunk_nt
            try:
                tickets = [capi.classify_reads_async(gidx, gtax, opts, nt, off, goff) if i % 2 == 0 else
                           capi.classify_reads_packed_async(gidx, gtax, opts, pk[0], pk[1], off, goff)
                           for i, (nt, off, goff, pk) in enumerate(batches)]
                for i in reversed(range(len(tickets))):  # waiting out of order is allowed
                    assert np.array_equal(tickets[i].wait(), want[i]), (kw, chunk_nt, i)
            finally:
                os.environ.pop("UMGAP_CHUNK_NT", None)
    # an index value the taxonomy does not hold: only the batch that meets it fails
    nt0, off0 = batches[0][0], batches[0][1]
    nt1, off1 = batches[1][0], batches[1][1]
    in_batch1 = set()
    for r in range(len(off1) - 1):
        for _, pep in otr.translate_record(bytes(nt1[int(off1[r]):int(off1[r + 1])]).decode(), 1, False):
            in_batch1.update(pep[i:i + 9] for i in range(len(pep) - 8))
    probe = None
    for r in range(len(off0) - 1):
        pep = otr.translate_record(bytes(nt0[int(off0[r]):int(off0[r + 1])]).decode(), 1, False, ["1"])[0][1][:9]
        if len(pep) == 9 and "*" not in pep and "-" not in pep and pep not in in_batch1:
            probe = pep.encode()
            break
    assert probe is not None
    idx_map = dict(world["index"])
    idx_map[probe] = 4_000_000_000
    k2 = sorted(idx_map)
    bidx = capi.Index.from_pairs(k2, [idx_map[k] for k in k2], k=9)
    try:
        opts = capi.default_opts(seedextend=0, strategy=capi.AGG_LCA_STAR)
        # batch 1 (reads of another length and seed) does not hold the probe k-mer
        clean = capi.classify_reads(gidx, gtax, opts, *batches[1][:3])[0].copy()
        t0 = capi.classify_reads_async(bidx, gtax, opts, *batches[1][:3])
        t1 = capi.classify_reads_async(bidx, gtax, opts, *batches[0][:3])
        t2 = capi.classify_reads_async(bidx, gtax, opts, *batches[1][:3])
        assert np.array_equal(t0.wait(), clean)
        with pytest.raises(capi.UmgapError) as e:
            t1.wait()
        assert "Unknown Taxon ID: 4000000000" in str(e.value)
        assert np.array_equal(t2.wait(), clean)
    finally:
        bidx.close()
    tickets = [capi.classify_reads_async(gidx, gtax, capi.default_opts(), *batches[0][:3]) for _ in range(32)]
    with pytest.raises(capi.UmgapError):
        capi.classify_reads_async(gidx, gtax, capi.default_opts(), *batches[0][:3])
    for t in tickets:
        t.wait()
    capi.classify_reads_async(gidx, gtax, capi.default_opts(), *batches[0][:3]).wait()


def test_scored_aggregation_matches_oracle(capi, world):
    """umgap_aggregate_scored (taxa2agg -s, taxa2agg.rs:141-148): f32 scores summed in the reference's orders.  Against the
    oracle's literal restatement: equality where its answer is unique, membership in the set of tied maxima / child
    orders otherwise; with all scores 1.0 the unscored kernels' answers; an unknown taxon raises only when its sum
    reaches the lower bound (the reference filters before the aggregator sees it, taxa2agg.rs:169-170), in both kernels."""
    rng = random.Random(515)
    otax = world["otax"]
    ids = [t[0] for t in otax.by_id if t is not None]
    recs = [[], [(0, 1.5)], [(ids[3], 0.25)], [(ids[3], 0.1)] * 7]
    for _ in range(400):
        home = rng.choice(ids)
        path = otax.root_path(home)
        r = []
        for _ in range(rng.choice([1, 2, 3, 6, 12, 40, 150])):
            u = rng.random()
            t = 0 if u < 0.15 else home if u < 0.5 else rng.choice(path) if u < 0.8 else rng.choice(ids)
            sc = rng.choice([1.0, 0.5, 0.1, 0.3, 2.75, float(np.float32(rng.random())), float(np.float32(rng.random() * 1e-3)), 1e6 + 0.5])
            r.append((t, sc))
        recs.append(r)
    flat = np.array([t for r in recs for t, _ in r] or [0], dtype=np.uint32)
    sc = np.array([s for r in recs for _, s in r] or [0], dtype=np.float32)
    off = np.zeros(len(recs) + 1, dtype=np.uint64)
    np.cumsum([len(r) for r in recs], out=off[1:])
    shuffles = [sorted] + [(lambda seed: (lambda xs: random.Random(seed).sample(sorted(xs), len(xs))))(k) for k in range(3)]
    unique = 0
    # `-m rmq -a hybrid` (rmq/mix.rs:56-93) through the same entry point: unscored input and dyadic scores (its sums run
    # in HashMap order in the reference: exact sums make every order agree), incl. the reference's own vectors
    fx = [(1, "root", 0, 1, True), (2, "Bacteria", 1, 1, True), (10239, "Viruses", 1, 1, True), (12884, "Viroids", 1, 1, True),
          (185751, "Pospiviroidae", 19, 12884, True), (185752, "Avsunviroidae", 19, 12884, True)]
    ftax = capi.Taxonomy.from_arrays(*datagen.taxonomy_arrays(fx))
    def mix(idl, f):
        a = np.array(idl, dtype=np.uint32)
        return int(capi.aggregate_scored(ftax, a, np.ones(len(idl), dtype=np.float32), np.array([0, len(idl)], dtype=np.uint64), capi.AGG_RMQ_HYBRID, f)[0])
    assert mix([12884, 185751], 0.0) == 185751 and mix([12884, 185751, 185752, 185752], 0.0) == 185752
    assert mix([1, 1, 10239, 10239, 10239, 12884, 185751, 185752], 0.0) == 10239
    assert mix([12884, 185751], 1.0) == 12884 and mix([1, 1, 10239, 10239, 10239, 12884, 185751, 185752], 1.0) == 1
    assert mix([12884, 12884, 185751], 0.5) == 12884 and mix([12884, 185751, 185751], 0.5) == 185751
    assert mix([1, 12884, 12884, 185751, 185752], 0.5) == 12884
    ftax.close()
    dyadic = np.array([float(int(x * 8) % 5 + 1) / 4 for x in sc], dtype=np.float32)
    for scores in (np.ones_like(sc), dyadic):
        for factor in (0.25, 0.0, 0.5, 1.0):
            for lb in (0.0, 1.0):
                got = capi.aggregate_scored(world["gtax"], flat, scores, off, capi.AGG_RMQ_HYBRID, factor, lb)
                snapping = otax.snapping(False)
                k = 0
                for i, r in enumerate(recs):
                    pairs = [(t, float(scores[k + j])) for j, (t, _) in enumerate(r)]
                    k += len(r)
                    want = oagg.taxa2agg_record_scored(otax, snapping, pairs, oagg.RMQ_HYBRID, factor, lb)
                    assert int(got[i]) in want, (factor, lb, pairs, int(got[i]), want)
    for strategy in (capi.AGG_LCA_STAR, capi.AGG_HYBRID, capi.AGG_MRTL):
        for factor in ((0.25, 0.0, 0.5, 1.0) if strategy == capi.AGG_HYBRID else (0.25,)):
            for lb in (0.0, 0.75, 2.0):
                for ranked in (False, True):
                    got = capi.aggregate_scored(world["gtax"], flat, sc, off, strategy, factor, lb, ranked)
                    snapping = otax.snapping(ranked)
                    for i, r in enumerate(recs):
                        mine = oagg.taxa2agg_record_scored(otax, snapping, r, strategy, factor, lb)          # the kernel's child order
                        anyorder = oagg.taxa2agg_record_scored(otax, snapping, r, strategy, factor, lb, shuffles)
                        assert int(got[i]) in mine and mine <= anyorder, (strategy, factor, lb, ranked, r, int(got[i]), mine)
                        unique += len(anyorder) == 1
    assert unique > 5000
    # scores of 1.0 are the unscored input (taxa2agg.rs:150-152)
    ones = np.ones_like(sc)
    for strategy in (capi.AGG_LCA_STAR, capi.AGG_HYBRID, capi.AGG_MRTL):
        a = capi.aggregate_scored(world["gtax"], flat, ones, off, strategy, 0.25, 1.0)
        b = capi.aggregate(world["gtax"], flat, off, strategy, 0.25, 1.0)
        assert np.array_equal(a, b), strategy
    # an id the tree does not hold: filtered away below the lower bound, an error when it survives
    unknown = max(ids) + 77
    rec_t = np.array([ids[5], ids[5], unknown, ids[9], ids[9]], dtype=np.uint32)
    rec_s = np.array([1.0, 1.0, 1.0, 1.0, 1.0], dtype=np.float32)
    o2 = np.array([0, 5], dtype=np.uint64)
    want = oagg.taxa2agg_record(otax, otax.snapping(False), [int(x) for x in rec_t], oagg.LCA_STAR, 0.25, 2.0)
    assert int(capi.aggregate_scored(world["gtax"], rec_t, rec_s, o2, capi.AGG_LCA_STAR, 0.25, 2.0)[0]) in want
    assert int(capi.aggregate(world["gtax"], rec_t, o2, capi.AGG_LCA_STAR, 0.25, 2.0)[0]) in want
    big = np.array([ids[5], ids[5], unknown] + [ids[k] for k in range(10, 60)] * 2, dtype=np.uint32)   # the list path of the warp kernel
    o3 = np.array([0, len(big)], dtype=np.uint64)
    want = oagg.taxa2agg_record(otax, otax.snapping(False), [int(x) for x in big], oagg.MRTL, 0.25, 2.0)
    assert int(capi.aggregate(world["gtax"], big, o3, capi.AGG_MRTL, 0.25, 2.0)[0]) in want
    for fn in (lambda: capi.aggregate_scored(world["gtax"], rec_t, rec_s, o2, capi.AGG_LCA_STAR, 0.25, 1.0),
               lambda: capi.aggregate(world["gtax"], rec_t, o2, capi.AGG_LCA_STAR, 0.25, 1.0),
               lambda: capi.aggregate(world["gtax"], big, o3, capi.AGG_MRTL, 0.25, 1.0)):
        with pytest.raises(capi.UmgapError) as e:
            fn()
        assert f"Unknown Taxon ID: {unknown}" in str(e.value)
