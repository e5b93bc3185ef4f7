"""Multi-GPU checks (need >= 2 visible GPUs: `gpurun --gpus 2`); skipped on a single-GPU box."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpus():
    import umgap_b200.capi as c
    return c.device_count()


def test_key_range_sharded_index_over_peer_memory():
    n = _ngpus()
    if n < 2:
        pytest.skip("needs 2 GPUs")
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", str(29600 + os.getpid() % 300),
                        os.path.join(ROOT, "tests", "dist_sharded_check.py")], stdout=subprocess.PIPE,
                       stderr=subprocess.PIPE, timeout=600)
    out = p.stdout.decode()
    assert p.returncode == 0, p.stderr.decode()[-3000:]
    assert "sharded ok rank 0/2" in out and "sharded ok rank 1/2" in out


def _torchrun(script, nproc):
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc),
                        "--master-addr", "127.0.0.1", "--master-port", str(29900 + os.getpid() % 300 + nproc),
                        os.path.join(ROOT, "tests", script)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=600)
    assert p.returncode == 0, p.stderr.decode()[-3000:] + p.stdout.decode()[-2000:]
    return p.stdout.decode()


def test_routed_exchange_single_rank():
    """The exchange path with one rank (one shard = the whole table, the exchange degenerates to the local bucket copy):
    pack / lookup / scatter kernels and the two-round sampled form against the fused path, on any GPU box."""
    if _ngpus() < 1:
        pytest.skip("needs a GPU")
    assert "routed ok rank 0/1" in _torchrun("dist_routed_check.py", 1)


def test_routed_exchange_two_ranks():
    if _ngpus() < 2:
        pytest.skip("needs 2 GPUs")
    out = _torchrun("dist_routed_check.py", 2)
    assert "routed ok rank 0/2" in out and "routed ok rank 1/2" in out


def test_in_kernel_exchange_matches_unsharded_table():
    """umgap_sharded (exchange.cu): one process, a shard per GPU, the pack kernels storing hashes into the owners' inboxes
    and the lookup kernels storing answers into the requesters' boxes over NVLink peer mappings -- against the unsharded
    table on the same reads: sampled (two rounds) and every-position (one round) option sets, ragged reads incl. reads
    longer than a sampled batch, several passes (buffers smaller than the batch), a GPU without reads, the oracle."""
    import random
    import numpy as np
    import datagen
    import umgap_b200.capi as capi
    from oracle import lookup as olookup, pipeline as opipe
    from oracle.taxonomy import Taxonomy as OTaxonomy
    n = min(_ngpus(), 4)
    if n < 2:
        pytest.skip("needs 2 GPUs")
    taxa = datagen.make_taxonomy(400, seed=91)
    otax = OTaxonomy(taxa)
    proteins = datagen.make_proteome(150, seed=92)
    index = datagen.make_index(proteins, otax, seed=93)
    keys = sorted(index)
    vals = [index[k] for k in keys]
    tax_arrays = datagen.taxonomy_arrays(taxa)
    gtax = [capi.Taxonomy.from_arrays(*tax_arrays, device=d) for d in range(n)]
    full = capi.Index.from_pairs(keys, vals, k=9, device=0)
    shards = [capi.Index.from_pairs(keys, vals, k=9, device=d, shard=d, nshards=n) for d in range(n)]
    rng = random.Random(7)
    reads = []
    for L in (100, 150, 151, 250, 301):
        reads += datagen.make_reads(proteins, 60, seed=300 + L, read_len=L, hit_frac=0.8)
    long_nt = "".join(rng.choice(datagen.CODONS.get(a, ["GCT"])) for p in proteins[:8] for a in p)
    odd = ["", "A", "ACG" * 9, "N" * 33, long_nt, datagen.revcomp(long_nt)[:1300], long_nt[:1281], "acgt" * 40]
    for i, sq in enumerate(odd):
        reads += [(f"odd{i}/1", sq), (f"odd{i}/2", odd[(i + 3) % len(odd)])]
    reads = [(f"q{i // 2}/{i % 2 + 1}", sq) for i, (_, sq) in enumerate(reads)]
    nt, roff = capi.pack_strings([r[1].encode() for r in reads])
    goff = np.arange(0, len(reads) + 1, 2, dtype=np.uint64)
    total_nt = int(roff[-1])
    cases = [dict(seedextend=1, min_seed_size=3, max_gap_size=0, strategy=1), dict(seedextend=1, min_seed_size=2, max_gap_size=1, strategy=2),
             dict(seedextend=1, min_seed_size=6, max_gap_size=2, strategy=0), dict(seedextend=0, one_on_one=0, strategy=0),
             dict(seedextend=1, min_seed_size=1, max_gap_size=0, strategy=1)]
    for max_nt in (total_nt + 1000, total_nt // (2 * n) + 2 * len(long_nt)):   # one pass; several passes
        ctx = capi.Sharded(shards, gtax, max_nt)
        for kw in cases:
            opts = capi.default_opts(**kw)
            want, nl = capi.classify_reads(full, gtax[0], opts, nt, roff, goff)
            got, nl2 = ctx.classify_reads(opts, nt, roff, goff)
            assert nl == nl2 and np.array_equal(want, got), (max_nt, kw)
        # a batch smaller than the number of GPUs (some GPUs take part in the rounds without reads), and an empty one
        got, _ = ctx.classify_reads(opts, nt, roff[:3], np.array([0, 2], dtype=np.uint64))
        assert np.array_equal(got, want[:1])
        got, _ = ctx.classify_reads(opts, nt, roff[:1], np.array([0], dtype=np.uint64))
        assert len(got) == 0
        ctx.close()
    # the oracle on the last case's pairs
    ref = opipe.classify_reads(reads, olookup.DictIndex(index), otax, min_seed_size=1, max_gap_size=0, strategy=1)
    ref = dict(ref)
    for g in range(len(goff) - 1):
        h = reads[2 * g][0].split("/")[0]
        assert (int(want[g]) in ref[h]) if h in ref else int(want[g]) == capi.ABSENT
