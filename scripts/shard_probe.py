"""Single-process probe of the sharded lookup kernel: two shards on one GPU, or on two GPUs (peer access)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import datagen
from umgap_b200 import capi

two_gpus = len(sys.argv) > 1 and sys.argv[1] == "2"
nprot, npairs, L = 2500000, 1000000, 150
taxa = datagen.make_taxonomy(5000, seed=1)
spec = capi.SynthSpec(seed=2, n_proteins=nprot, protein_len=408, home_pct=70, ancestor_pct=20)
devs = [0, 1] if two_gpus else [0, 0]
taxs = {d: capi.Taxonomy.from_arrays(*datagen.taxonomy_arrays(taxa), device=d) for d in set(devs)}
shards = [capi.Index.build_synthetic(spec, taxs[devs[s]], devs[s], 0.5, shard=s, nshards=2) for s in range(2)]
for s in shards: s.attach_shards_local(shards)
print([s.info().n_keys for s in shards], flush=True)
torch.cuda.set_device(0)
nt = torch.empty(npairs * 2 * L, dtype=torch.uint8, device="cuda:0")
capi.synth_reads_dev(spec, 3, 0, npairs, L, 70, nt.data_ptr())
roff = torch.arange(0, npairs * 2 + 1, dtype=torch.int64, device="cuda:0") * L
ids = torch.empty(2 * npairs * 2 * L + 64, dtype=torch.int32, device="cuda:0")
opts = capi.default_opts(min_seed_size=3)
st = torch.cuda.current_stream().cuda_stream
f = lambda: capi.translate_lookup_dev(shards[0], opts, nt.data_ptr(), roff.data_ptr(), npairs * 2, npairs * 2 * L, ids.data_ptr(), st)
f(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
e0.record(); f(); f(); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 2
print(f"sharded lookup kernel ({'2 GPUs, peer access' if two_gpus else '2 shards on one GPU'}): {ms:.2f} ms, {npairs*2*248/ms/1e6:.1f} G lookups/s")
