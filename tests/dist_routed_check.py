"""Run under torchrun with one rank per GPU (any world size >= 1): the routed form of the key-range-sharded mode
(umgap_b200/sharded.py: RoutedClassifier -- pack, exchange, local lookups, exchange back, scatter, classify) against
the fused path over a replicated table, on ragged batches with reads of every length, for the sampled two-round
form (`-o | seedextend -s S`, S >= 2) and the every-position form.  Used by tests/test_gpu_multi.py:
  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tests/dist_routed_check.py
"""
import os
import random
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
    # ragged batch: pairs of several lengths, odd sizes, reads longer than a warp batch of the sampled kernel
    rng = random.Random(7 + rank)
    reads = []
    for L in (100, 150, 151, 250, 301):
        reads += datagen.make_reads(proteins, 40, seed=300 + L + rank, read_len=L, hit_frac=0.8)
    long_nt = "".join(rng.choice(datagen.CODONS.get(a, ["GCT"])) for p in proteins[:8] for a in p)
    assert len(long_nt) > 1400
    odd = ["", "A", "ACG" * 9, "N" * 33, long_nt, datagen.revcomp(long_nt)[:1300], long_nt[:1281], long_nt[3:1283], "acgt" * 40]
    for i, sq in enumerate(odd):
        reads += [(f"odd{i}/1", sq), (f"odd{i}/2", odd[(i + 3) % len(odd)])]
    reads = [(f"q{i // 2}/{i % 2 + 1}", sq) for i, (_, sq) in enumerate(reads)]   # unique headers, pairs stay pairs
    nt, roff = capi.pack_strings([r[1].encode() for r in reads])
    goff = np.arange(0, len(reads) + 1, 2, dtype=np.uint64)
    total_nt = int(roff[-1])
    rc = sharded.RoutedClassifier(shard, gtax, dist, max_total_nt=total_nt + 1000, lanes=2)
    d_nt = torch.zeros(total_nt + 64, dtype=torch.uint8, device="cuda")
    d_nt[:total_nt] = torch.from_numpy(nt)
    d_roff = torch.from_numpy(roff.astype(np.int64)).cuda()
    d_goff = torch.from_numpy(goff.astype(np.int64)).cuda()
    oidx = olookup.DictIndex(index)
    cases = [dict(seedextend=1, min_seed_size=3, max_gap_size=0, strategy=1),    # sampled, stride 3
             dict(seedextend=1, min_seed_size=2, max_gap_size=1, strategy=2),    # sampled, stride 2
             dict(seedextend=1, min_seed_size=6, max_gap_size=2, strategy=0),    # sampled, stride 4
             dict(seedextend=0, min_seed_size=2, max_gap_size=0, strategy=1),    # every position
             dict(seedextend=1, min_seed_size=1, max_gap_size=0, strategy=1)]    # every position (S < 2)
    for case, cut in [(c, cut) for cut in (False, True) for c in cases]:
        # cut: the batch is split into two group ranges on two streams whose exchange rounds alternate (sampled form only)
        rc.min_groups_per_lane = 16 if cut else 1 << 30
        opts = capi.default_opts(**case)
        sampled = capi.route_sampled_applies(shard, opts)
        assert sampled == (case["seedextend"] == 1 and case["min_seed_size"] >= 2), case
        x, _ = capi.classify_reads(full, gtax, opts, nt, roff, goff)
        d_out = torch.zeros(len(goff) - 1, dtype=torch.int32, device="cuda")
        rc.classify(opts, d_nt[:total_nt], d_roff, d_goff, d_out, total_nt)
        torch.cuda.synchronize()
        assert not rc.overflowed()
        y = d_out.cpu().numpy().view(np.uint32)
        assert np.array_equal(y, x), (case, np.nonzero(y != x)[0][:10], y[y != x][:10], x[y != x][:10])
        want = dict(opipe.classify_reads(reads, oidx, otax, use_seedextend=bool(case["seedextend"]),
                                         min_seed_size=case["min_seed_size"], max_gap_size=case["max_gap_size"],
                                         strategy=case["strategy"], factor=0.25, lower_bound=0.0))
        heads = [h.split("/")[0] for h, _ in reads[::2]]
        below = 0
        for h, g in zip(heads, y):
            if h in want:
                assert int(g) in want[h], (case, h, int(g), want[h])
                below += int(g) != 1
        assert below > (20 if case["strategy"] else 5), (case, below)
        n_routed = torch.tensor([rc.lookups_routed], dtype=torch.int64, device="cuda")
        dist.all_reduce(n_routed)
        if rank == 0:
            print(f"routed case {case} cut={cut}: sampled={sampled} lookups routed {int(n_routed.item())}", flush=True)
    # a bucket that overflows is reported, not silently dropped
    tiny = sharded.RoutedClassifier(shard, gtax, dist, max_total_nt=total_nt + 1000, slack=0.01)
    for lane in tiny.lanes:
        lane.cap = 64
    d_out = torch.zeros(len(goff) - 1, dtype=torch.int32, device="cuda")
    try:   # every rank learns of the overflow in the same exchange and raises: none is left behind in a collective
        tiny.classify(capi.default_opts(seedextend=1, min_seed_size=3), d_nt[:total_nt], d_roff, d_goff, d_out, total_nt)
        raise AssertionError("a bucket overflow went unreported")
    except RuntimeError as e:
        assert "overflowed" in str(e)
    assert tiny.overflowed()
    dist.barrier()
    torch.cuda.synchronize()
    shard.close()
    full.close()
    dist.destroy_process_group()
    print(f"routed ok rank {rank}/{world}")


if __name__ == "__main__":
    main()
