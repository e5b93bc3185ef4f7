"""Run under torchrun with one rank per GPU: builds a key-range-sharded index, attaches the peers'
shards over CUDA IPC and checks lookups and fused classification against the replicated index and
the oracle.  Used by tests/test_gpu_multi.py and runnable by hand:
  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tests/dist_sharded_check.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import datagen  # noqa: E402
from oracle import lookup as olookup, pipeline as opipe  # noqa: E402
from oracle.taxonomy import Taxonomy as OTaxonomy  # noqa: E402
from umgap_b200 import capi, sharded  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    taxa = datagen.make_taxonomy(400, seed=91)
    otax = OTaxonomy(taxa)
    proteins = datagen.make_proteome(150, seed=92)
    index = datagen.make_index(proteins, otax, seed=93)
    keys = sorted(index)
    vals = [index[k] for k in keys]
    gtax = capi.Taxonomy.from_arrays(*datagen.taxonomy_arrays(taxa), device=local)
    full = capi.Index.from_pairs(keys, vals, k=9, device=local)
    shard = capi.Index.from_pairs(keys, vals, k=9, device=local, shard=rank, nshards=world)
    n_mine = shard.info().n_keys
    t = torch.tensor([n_mine], dtype=torch.int64, device="cuda")
    dist.all_reduce(t)
    assert int(t.item()) == len(keys), (int(t.item()), len(keys))          # the shards partition the keys
    assert 0 < n_mine < len(keys)
    try:
        capi.kmer_lookup(shard, *capi.pack_strings([keys[0]]), True)
        raise SystemExit("lookup on an unattached shard must fail")
    except capi.UmgapError:
        pass
    sharded.attach_all(shard, dist)
    # every key and some misses through the sharded view == through the replicated table
    probes = keys[rank::3] + [bytes(reversed(k)) for k in keys[:500]]
    aa, off = capi.pack_strings(probes)
    a, _, _ = capi.kmer_lookup(full, aa, off, True)
    b, _, _ = capi.kmer_lookup(shard, aa, off, True)
    assert np.array_equal(a, b)
    assert [int(x) for x in b[:len(keys[rank::3])]] == [index[k] for k in keys[rank::3]]
    # fused classification: each rank takes its own reads, sharded == replicated, and in the oracle's set
    reads = datagen.make_reads(proteins, 150, seed=94 + rank)
    nt, roff = capi.pack_strings([r[1].encode() for r in reads])
    goff = np.arange(0, len(reads) + 1, 2, dtype=np.uint64)
    for strategy in (0, 1, 2):
        opts = capi.default_opts(min_seed_size=3, strategy=strategy)
        x, _ = capi.classify_reads(full, gtax, opts, nt, roff, goff)
        y, _ = capi.classify_reads(shard, gtax, opts, nt, roff, goff)
        assert np.array_equal(x, y)
        want = opipe.classify_reads(reads, olookup.DictIndex(index), otax, min_seed_size=3, strategy=strategy)
        for (h, adm), g in zip(want, y):
            assert int(g) in adm, (h, int(g), adm)
    # routed variant: NCCL all-to-all of packed hashes and answers, lookups in the local shard only
    rc = sharded.RoutedClassifier(shard, gtax, dist, max_total_nt=int(roff[-1]) + 1000)
    d_nt = torch.from_numpy(nt).cuda()
    d_roff = torch.from_numpy(roff.astype(np.int64)).cuda()
    d_goff = torch.from_numpy(goff.astype(np.int64)).cuda()
    for strategy in (0, 1, 2):
        opts = capi.default_opts(min_seed_size=3, strategy=strategy)
        x, _ = capi.classify_reads(full, gtax, opts, nt, roff, goff)
        d_out = torch.zeros(len(goff) - 1, dtype=torch.int32, device="cuda")
        rc.classify(opts, d_nt, d_roff, d_goff, d_out, int(roff[-1]))
        torch.cuda.synchronize()
        assert not rc.overflowed()
        assert np.array_equal(d_out.cpu().numpy().view(np.uint32), x), strategy
    dist.barrier()
    torch.cuda.synchronize()
    shard.close()
    full.close()
    dist.destroy_process_group()
    print(f"sharded ok rank {rank}/{world}: {n_mine} of {len(keys)} keys local")


if __name__ == "__main__":
    main()
