"""Host-buffer entry point (umgap_classify_reads) timing: pinned inputs, 1e9-key index, 1 M pairs per call."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import datagen
from umgap_b200 import capi
nprot = int(float(sys.argv[1])) if len(sys.argv) > 1 else 2500000
npairs = int(float(sys.argv[2])) if len(sys.argv) > 2 else 1000000
taxa = datagen.make_taxonomy(5000, seed=1)
gtax = capi.Taxonomy.from_arrays(*datagen.taxonomy_arrays(taxa))
spec = capi.SynthSpec(seed=2, n_proteins=nprot, protein_len=408, home_pct=70, ancestor_pct=20)
gidx = capi.Index.build_synthetic(spec, gtax, 0, 0.0)
L = 150
nt = torch.empty(npairs * 2 * L, dtype=torch.uint8, device="cuda")
capi.synth_reads_dev(spec, 3, 0, npairs, L, 70, nt.data_ptr())
h_nt = torch.empty(npairs * 2 * L, dtype=torch.uint8).pin_memory(); h_nt.copy_(nt)
def pinned(a):
    t = torch.from_numpy(a.view(np.int64) if a.dtype == np.uint64 else a.view(np.int32)).pin_memory()
    return t.numpy().view(a.dtype), t
h_roff, k1 = pinned(np.arange(0, 2 * npairs + 1, dtype=np.uint64) * L)
h_goff, k2 = pinned(np.arange(0, 2 * npairs + 1, 2, dtype=np.uint64))
h_out, k3 = pinned(np.zeros(npairs, dtype=np.uint32))
opts = capi.default_opts(min_seed_size=3, strategy=1)
nt_np = h_nt.numpy()
for _ in range(2):
    capi.classify_reads(gidx, gtax, opts, nt_np, h_roff, h_goff, count_lookups=False, out=h_out)
torch.cuda.synchronize()
n = 8
t0 = time.perf_counter()
for _ in range(n):
    capi.classify_reads(gidx, gtax, opts, nt_np, h_roff, h_goff, count_lookups=False, out=h_out)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / n
print(f"e2e {os.environ.get('UMGAP_CHUNK_MB','default')} MB chunks ordered={os.environ.get('UMGAP_CHUNK_ORDERED','0')}: {dt*1e3:.3f} ms per call  {2*npairs/dt/1e6:.1f} M reads/s  checksum {int(h_out.astype(np.uint64).sum())}", flush=True)
