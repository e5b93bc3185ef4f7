"""Quick device probe: synthetic index + reads, kernel timings, random-sector rate."""
import sys, time, json, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import datagen
from umgap_b200 import capi

nprot = int(float(sys.argv[1])) if len(sys.argv) > 1 else 250000
npairs = int(float(sys.argv[2])) if len(sys.argv) > 2 else 1000000
lf = float(sys.argv[3]) if len(sys.argv) > 3 else 0.0
taxa = datagen.make_taxonomy(5000, seed=1)
gtax = capi.Taxonomy.from_arrays(*datagen.taxonomy_arrays(taxa))
spec = capi.SynthSpec(seed=2, n_proteins=nprot, protein_len=408, home_pct=70, ancestor_pct=20)
t0 = time.time()
gidx = capi.Index.build_synthetic(spec, gtax, 0, lf)
torch.cuda.synchronize()
info = gidx.info()
print("index build s", time.time() - t0, {f[0]: getattr(info, f[0]) for f in info._fields_}, flush=True)
L = int(os.environ.get("PROBE_LEN", "150"))
nt = torch.empty(npairs * 2 * L, dtype=torch.uint8, device="cuda")
capi.synth_reads_dev(spec, 3, 0, npairs, L, 70, nt.data_ptr())
roff = (torch.arange(0, npairs * 2 + 1, dtype=torch.int64, device="cuda") * L)
goff = torch.arange(0, npairs * 2 + 1, 2, dtype=torch.int64, device="cuda")
out = torch.zeros(npairs, dtype=torch.int32, device="cuda")
ids = torch.empty(2 * npairs * 2 * L + 64, dtype=torch.int32, device="cuda")
opts = capi.default_opts(min_seed_size=int(os.environ.get("PROBE_SEED", "3")), max_gap_size=int(os.environ.get("PROBE_GAP", "0")), strategy=int(os.environ.get("PROBE_STRATEGY", "1")))
st = torch.cuda.current_stream().cuda_stream
def timeit(f, n=5):
    for _ in range(2): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
nlook = npairs * 2 * 2 * (L - 26)
ms1 = timeit(lambda: capi.translate_lookup_dev(gidx, opts, nt.data_ptr(), roff.data_ptr(), npairs * 2, npairs * 2 * L, ids.data_ptr(), st))
print(f"translate_lookup: {ms1:.3f} ms  {nlook/ms1/1e6:.2f} G lookups/s  {nlook*32/ms1/1e6:.1f} GB/s algorithmic", flush=True)
ms2 = timeit(lambda: capi.classify_reads_dev(gidx, gtax, opts, nt.data_ptr(), roff.data_ptr(), npairs * 2, npairs * 2 * L, goff.data_ptr(), npairs, out.data_ptr(), st))
print(f"classify (both kernels): {ms2:.3f} ms  {npairs*2/ms2/1e3:.2f} M reads/s", flush=True)
for sl in [int(x) for x in os.environ.get("PROBE_SLICES", "").split(",") if x]:
    capi.pipeline_slices(sl)
    ms3 = timeit(lambda: capi.classify_reads_dev(gidx, gtax, opts, nt.data_ptr(), roff.data_ptr(), npairs * 2, npairs * 2 * L, goff.data_ptr(), npairs, out.data_ptr(), st), 8)
    print(f"slices {sl}: {ms3:.3f} ms  {npairs*2/ms3/1e3:.2f} M reads/s", flush=True)
capi.pipeline_slices(1)
for gib in [float(x) for x in os.environ.get("PROBE_REGIONS", "").split(",") if x]:   # probe-region sizes in GiB
    gidx.set_probe_region(int(gib * (1 << 30)))
    for sampling in (1, 0):
        capi.pipeline_sampling(sampling)
        ms4 = timeit(lambda: capi.classify_reads_dev(gidx, gtax, opts, nt.data_ptr(), roff.data_ptr(), npairs * 2, npairs * 2 * L, goff.data_ptr(), npairs, out.data_ptr(), st), 4)
        capi.kernel_timing(True); capi.kernel_times()
        capi.classify_reads_dev(gidx, gtax, opts, nt.data_ptr(), roff.data_ptr(), npairs * 2, npairs * 2 * L, goff.data_ptr(), npairs, out.data_ptr(), st)
        torch.cuda.synchronize()
        lm, ln, cm, cn = capi.kernel_times(); capi.kernel_timing(False)
        print(f"region {gib} GiB sampling {sampling}: step {ms4:.3f} ms  {npairs*2/ms4/1e3:.2f} M reads/s  lookup stage {lm:.3f} ms ({ln} brackets) classify {cm:.3f} ms  sum(out) {int(out.sum())}", flush=True)
    capi.pipeline_sampling(1)
gidx.set_probe_region(0)
capi.kernel_timing(True); capi.kernel_times()
for _ in range(5):
    capi.classify_reads_dev(gidx, gtax, opts, nt.data_ptr(), roff.data_ptr(), npairs * 2, npairs * 2 * L, goff.data_ptr(), npairs, out.data_ptr(), st)
torch.cuda.synchronize()
lm, ln, cm, cn = capi.kernel_times(); capi.kernel_timing(False)
print(f"kernels: lookup {lm/max(ln,1):.3f} ms  classify {cm/max(cn,1):.3f} ms", flush=True)
o = out.cpu().numpy().view(np.uint32)
print("classified below root:", float((o != 1).mean()), "absent:", int((o == 0xFFFFFFFF).sum()))
hits = (ids[: 2 * npairs * 2 * L].view(torch.int32) != -1)
print("hit-rate over id slots (incl. unused slots):", float(hits.float().mean()))
for n in (1 << 26, 1 << 28):
    print("randsector", n, gidx.randsector_rate(n, 5) / 1e9, "G sectors/s", flush=True)
