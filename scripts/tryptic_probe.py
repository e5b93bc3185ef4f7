"""Device probe of the fused peptide path (BASELINE configs[2]: prot2tryp2lca -l9 -L45 | uniq -d / | taxa2agg -l1 -a mrtl):
synthetic proteome -> tryptic index (peptides of 9..45 residues, value = the protein's taxon), predicted-gene style
fragments in pairs -> umgap_classify_peptides_dev / umgap_classify_peptides timed with CUDA events / wall clock, and
a property check: every group's answer is the root or the (snapped) taxon of the protein its fragments come from.
  python scripts/tryptic_probe.py [n_proteins] [n_pairs] [fragment_aa]
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import datagen
from umgap_b200 import capi

nprot = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000
npairs = int(float(sys.argv[2])) if len(sys.argv) > 2 else 1_000_000
frag = int(sys.argv[3]) if len(sys.argv) > 3 else 50
L = 400
rng = np.random.default_rng(5)
taxa = datagen.make_taxonomy(5000, seed=1)
gtax = capi.Taxonomy.from_arrays(*datagen.taxonomy_arrays(taxa))
ids = np.array([t[0] for t in taxa], dtype=np.uint64)
# residue frequencies close to UniProt's (K + R = 11 %)
letters = np.frombuffer(b"ACDEFGHIKLMNPQRSTVWY", dtype=np.uint8)
freq = np.array([8.3, 1.4, 5.5, 6.8, 3.9, 7.1, 2.3, 5.9, 5.8, 9.7, 2.4, 4.1, 4.7, 3.9, 5.5, 6.6, 5.4, 6.9, 1.1, 2.9])
t0 = time.time()
prot = letters[rng.choice(20, size=(nprot, L), p=freq / freq.sum())]
home = ids[rng.integers(0, len(ids), nprot)]
flat = prot.reshape(-1)
prev = np.empty_like(flat)
prev[1:] = flat[:-1]
start = ((prev == ord("K")) | (prev == ord("R"))) & (flat != ord("P"))
start[::L] = True
pos = np.flatnonzero(start)
nxt = np.empty_like(pos)
nxt[:-1] = pos[1:]
nxt[-1] = flat.size
row_end = (pos // L + 1) * L
end = np.minimum(nxt, row_end)
ln = end - pos
keep = (ln >= 9) & (ln <= 45)
kpos, klen = pos[keep], ln[keep].astype(np.uint64)
koff = np.zeros(len(kpos) + 1, dtype=np.uint64)
np.cumsum(klen, out=koff[1:])
# gather the key bytes: index = start + (0..len)
rep = np.repeat(kpos - koff[:-1].astype(np.int64), klen.astype(np.int64)) + np.arange(int(koff[-1]), dtype=np.int64)
blob = flat[rep]
vals = home[kpos // L]
print(f"proteome {nprot} x {L} aa, {len(kpos)} tryptic peptides of 9..45 residues ({blob.nbytes / 1e6:.0f} MB of key bytes), {time.time() - t0:.1f} s", flush=True)
t0 = time.time()
gidx = capi.Index.from_blob(blob, koff, vals, k=0)
info = gidx.info()
print(f"peptide table: {info.n_keys} keys, {info.bytes / 1e6:.0f} MB in HBM, built in {time.time() - t0:.1f} s", flush=True)
# fragments: pairs of windows of one protein (70 %) or random residues (30 %)
nlines = 2 * npairs
src = rng.integers(0, nprot, npairs)
hit = rng.random(npairs) < 0.7
a = rng.integers(0, L - frag, nlines)
lines = prot[np.repeat(src, 2)[:, None], (a[:, None] + np.arange(frag)[None, :])]
noise = letters[rng.integers(0, 20, size=(nlines, frag))]
lines = np.where(np.repeat(hit, 2)[:, None], lines, noise)
aa = np.ascontiguousarray(lines.reshape(-1))
loff = (np.arange(nlines + 1, dtype=np.uint64) * frag)
goff = np.arange(0, nlines + 1, 2, dtype=np.uint64)
opts = capi.tryp_opts(minlen=9, maxlen=45, strategy=capi.AGG_MRTL, lower_bound=1.0)
d_aa = torch.from_numpy(aa).cuda()
d_loff = torch.from_numpy(loff.astype(np.int64)).cuda()
d_goff = torch.from_numpy(goff.astype(np.int64)).cuda()
d_out = torch.zeros(npairs, dtype=torch.int32, device="cuda")
st = torch.cuda.current_stream().cuda_stream
def run():
    capi.classify_peptides_dev(gidx, gtax, opts, d_aa.data_ptr(), d_loff.data_ptr(), nlines, aa.size, d_goff.data_ptr(), npairs,
                               d_out.data_ptr(), st)
for _ in range(3):
    run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
e0.record()
for _ in range(10):
    run()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
out = d_out.cpu().numpy().view(np.uint32)
print(f"device-resident: {ms:.3f} ms per {npairs} pairs of {frag}-aa fragments = {nlines / ms / 1e3:.1f} M peptide lines/s, "
      f"{aa.size / ms / 1e6:.1f} G residues/s", flush=True)
h = capi.classify_peptides(gidx, gtax, opts, aa, loff, goff)
t0 = time.perf_counter()
for _ in range(3):
    h = capi.classify_peptides(gidx, gtax, opts, aa, loff, goff)
dt = (time.perf_counter() - t0) / 3
assert np.array_equal(h, out)
print(f"host buffers (pageable, H2D + D2H inside): {dt * 1e3:.2f} ms = {nlines / dt / 1e6:.1f} M peptide lines/s", flush=True)
def pin(x):
    t = torch.from_numpy(x.view(np.int64) if x.dtype == np.uint64 else x).pin_memory()
    return t.numpy().view(x.dtype), t
p_aa, _k1 = pin(aa)
p_loff, _k2 = pin(loff)
p_out, _k4 = pin(np.zeros(npairs, dtype=np.uint32))
p_goff, _k3 = pin(goff)
for chunk in (sys.argv[4].split(",") if len(sys.argv) > 4 else ["4194304", "16777216", "67108864", "1073741824"]):
    os.environ["UMGAP_PEP_CHUNK_BYTES"] = chunk
    h = capi.classify_peptides(gidx, gtax, opts, p_aa, p_loff, p_goff, out=p_out)
    t0 = time.perf_counter()
    for _ in range(5):
        h = capi.classify_peptides(gidx, gtax, opts, p_aa, p_loff, p_goff, out=p_out)
    dt = (time.perf_counter() - t0) / 5
    assert np.array_equal(h, out)
    print(f"host buffers (pinned), ranges of {int(chunk) >> 20} MB: {dt * 1e3:.2f} ms = {nlines / dt / 1e6:.1f} M peptide lines/s", flush=True)
os.environ.pop("UMGAP_PEP_CHUNK_BYTES")
# property: the answer of a group is the root or the snapped taxon of its protein (= the aggregate of that taxon alone)
want = capi.aggregate(gtax, home[src].astype(np.uint32), np.arange(npairs + 1, dtype=np.uint64), capi.AGG_MRTL)
ok = (out == 1) | (out == want)
print(f"classified below root: {(out != 1).mean():.3f}; groups answering root or their protein's taxon: {ok.mean():.4f}", flush=True)
assert ok.mean() > 0.999
