"""Times umgap_index_build_from_proteins (splitkmers | sort | joinkmers | buildindex on the device) on a synthetic
protein table: n_proteins x 400 residues, a fraction of the proteins are copies of others under related taxa so that
k-mers with several taxa occur.   python scripts/build_probe.py [n_proteins] [copy_fraction]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import ctypes as C
import numpy as np
import datagen
from umgap_b200 import capi

nprot = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000
copies = float(sys.argv[2]) if len(sys.argv) > 2 else 0.3
L = 400
rng = np.random.default_rng(11)
taxa = datagen.make_taxonomy(5000, seed=1)
gtax = capi.Taxonomy.from_arrays(*datagen.taxonomy_arrays(taxa))
ids = np.array([t[0] for t in taxa], dtype=np.uint64)
letters = np.frombuffer(b"ACDEFGHIKLMNPQRSTVWY", dtype=np.uint8)
t0 = time.time()
norig = max(1, int(nprot * (1 - copies)))
prot = np.empty((nprot, L), dtype=np.uint8)
prot[:norig] = letters[rng.integers(0, 20, size=(norig, L))]
src = rng.integers(0, norig, nprot - norig)
prot[norig:] = prot[src]
tid = ids[rng.integers(0, len(ids), nprot)]
tid[norig:] = np.where(rng.random(nprot - norig) < 0.5, tid[src], tid[norig:])
aa = prot.reshape(-1)
off = np.arange(nprot + 1, dtype=np.uint64) * L
print(f"{nprot} proteins, {nprot * (L - 8)} windows generated in {time.time() - t0:.1f} s", flush=True)
lib = capi.load_library()
for rep in range(2):
    h = C.c_void_p()
    t0 = time.perf_counter()
    capi._check(lib.umgap_index_build_from_proteins(gtax._h, capi._p(aa), capi._p(off), capi._p(tid), C.c_uint64(nprot), C.c_int(9),
                                                    C.c_double(0.0), C.byref(h)))
    dt = time.perf_counter() - t0
    gidx = capi.Index(h)
    info = gidx.info()
    print(f"build {rep}: {dt:.2f} s for {nprot * (L - 8) / 1e6:.0f} M windows -> {info.n_keys / 1e6:.1f} M keys, table {info.bytes / 1e9:.2f} GB "
          f"({nprot * (L - 8) / dt / 1e6:.0f} M windows/s, host copy and alphabet scan included)", flush=True)
    gidx.close()
