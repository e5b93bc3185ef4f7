#!/usr/bin/env python
"""Benchmark of the per-read classification hot path (BASELINE.json configs[1]).

Workload: the high-precision 9-mer pipeline `translate -a | prot2kmer2lca -o | seedextend -s3 -g0 |
uniq -d / | taxa2agg -a hybrid -f 0.25` on synthetic 150-nt paired reads against a synthetic
1e9-entry 9-mer index (SURVEY 8(d) recipe, generated on the device from a counter-based hash).
A "step" is one pass of the hot path over one batch of pairs; with the defaults 10 steps cover
the 10 M pairs of the configuration.

  python bench.py [--gpus N] [--steps K] [--warmup W]            our arm (CUDA, one process per GPU)
  python bench.py --impl reference ...                           reference arm: the CPU port on all host cores

One JSON line on stdout (rank 0).  `value` is device-resident throughput (CUDA events, max over
ranks); `e2e` goes through the host-buffer C ABI call (umgap_classify_reads) with pinned host
input, H2D and D2H inside the timed region; `roofline` is the lookup kernel against the measured
HBM copy peak at 32 algorithmic bytes per lookup; `cpu_baseline` is the C restatement of the
reference algorithm (oracle/c) on the host cores over a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

METRIC = "reads_per_second"
UNIT = "reads/s"
READ_LEN = 150
K = 9
LOOKUPS_PER_READ = 2 * (READ_LEN - 3 * K + 1)  # 248
BYTES_PER_LOOKUP = 32  # one DRAM sector per probe
HIT_PCT = 70
N_TAXA = 5000
PROTEIN_LEN = 408  # 400 nine-mer windows per protein


_REAL_STDOUT = None


def emit(line: str) -> None:
    """The JSON line is the only thing that reaches the real stdout: fd 1 is pointed at stderr for the
    whole run so that library chatter (NCCL prints its version banner on stdout) cannot precede it."""
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, (line + "\n").encode())


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--pairs-per-step", type=int, default=1_000_000)
    ap.add_argument("--index-keys", type=float, default=1e9, help="synthetic index size (9-mer windows)")
    ap.add_argument("--cpu-index-keys", type=float, default=1e8, help="index size of the host-resident CPU legs")
    ap.add_argument("--legs", default="all", help="extra legs after the headline: all | none | comma list of sustained,every_position,parity,sharded,large_index,tryptic,loader")
    ap.add_argument("--large-index-keys", type=float, default=6.12e9, help="windows of the large-index leg (6.12e9 -> a 100 GB table)")
    ap.add_argument("--load-factor", type=float, default=0.0, help="table load factor (0 = library default policy)")
    ap.add_argument("--sharded", action="store_true", help="key-range-shard the index over the GPUs (peer-memory lookups) instead of replicating it")
    ap.add_argument("--peer-loads", action="store_true", help="with --sharded: read remote shards through IPC peer mappings instead of the all-to-all exchange")
    ap.add_argument("--routed-lanes", type=int, default=1, help="with --sharded: group ranges of a batch whose exchange rounds alternate on separate streams")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-depth", type=int, default=2, help="batches in flight in the asynchronous end-to-end measurement")
    return ap.parse_args()


def pipeline_config(args, extra=None):
    cfg = {
        "workload": "high-precision 9-mer pipeline: translate -a | prot2kmer2lca -o | seedextend -s3 -g0 | uniq -d / | "
                    "taxa2agg -a hybrid -f 0.25; synthetic 150-nt paired reads (70 % drawn from the indexed proteome), "
                    f"synthetic {args.index_keys:.0e}-window 9-mer index, {N_TAXA}-taxon tree",
        "pairs_per_step": args.pairs_per_step,
        "read_len": READ_LEN,
        "k": K,
        "index": ("key-range sharded over the GPUs, " + ("remote sectors read over NVLink peer mappings" if getattr(args, "peer_loads", False)
                                                           else "packed k-mer hashes and answers exchanged by NCCL all-to-all") if getattr(args, "sharded", False)
                  else "replicated per GPU") if args.gpus > 1 else "single GPU",
        "partitioning": f"reads partitioned over {args.gpus} GPU(s), " + (
            "two exchange rounds per batch (sampled positions, then the live frames): grouped NCCL send/recv of the filled bucket parts"
            if getattr(args, "sharded", False) and not getattr(args, "peer_loads", False) and args.gpus > 1 else "no data-path collective"),
        "cache": "inputs larger than L2: every step reads a different 300 MB batch and probes a table of GBs",
    }
    if extra:
        cfg.update(extra)
    return cfg


# ----------------------------------------------------------------------------------------- CPU legs

def cpu_instance(n_keys: float, threads: int):
    """Host-resident instance of the same generator (oracle/c mirror of the device generator): taxonomy, fst image
    of the proteome's 9-mer index, and a reads() closure."""
    import datagen
    from oracle import cport, synth
    taxa = datagen.make_taxonomy(N_TAXA, seed=1)
    pre = synth.Preorder(taxa)
    n_prot = max(1, int(n_keys // (PROTEIN_LEN - 8)))
    t0 = time.perf_counter()
    data, nkeys = cport.synth_fst(2, n_prot, PROTEIN_LEN, 70, 20, pre, threads=threads)
    build_s = time.perf_counter() - t0
    img = cport.FstImage(data)
    ctax = cport.RefTaxonomy(taxa)
    opts = cport.RefOpts(table=1, methionine=0, one_on_one=1, seedextend=1, min_seed_size=3, max_gap_size=0,
                         strategy=1, factor=0.25, lower_bound=0.0, ranked_only=0, k=K)

    def reads(first_pair: int, npairs: int):
        nt = cport.synth_reads(2, n_prot, PROTEIN_LEN, 3, first_pair, npairs, READ_LEN, HIT_PCT, threads=threads).reshape(-1)
        off = np.arange(0, len(nt) + 1, READ_LEN, dtype=np.uint64)
        goff = np.arange(0, 2 * npairs + 1, 2, dtype=np.uint64)
        return nt, off, goff

    return dict(img=img, tax=ctax, taxa=taxa, opts=opts, reads=reads, n_keys=int(nkeys), n_prot=n_prot, fst_bytes=len(img.data),
                fst_build_s=build_s, raw=data)


def cpu_run(inst, first_pair: int, npairs: int, threads: int, lookups_only: bool = False):
    from oracle import cport
    nt, off, goff = inst["reads"](first_pair, npairs)
    t0 = time.perf_counter()
    out, nl, nh = cport.classify(inst["img"], inst["tax"], inst["opts"], nt, off, goff, threads=threads, lookups_only=lookups_only)
    dt = time.perf_counter() - t0
    return dt, nl, out


def cpu_staged(inst, first_pair: int, npairs: int, threads: int):
    """The reference's own structure (scripts/umgap-analyse.sh:276-311): five stages with FASTA text between them, only
    prot2kmer2lca multi-threaded (prot2kmer2lca.rs:163-166).  The stages run one after another here; as concurrent
    processes the pipe's rate is that of its slowest stage."""
    from oracle import cport
    nt, off, _ = inst["reads"](first_pair, npairs)
    rows = nt.reshape(-1, READ_LEN)
    fa = b"".join(b">r%d/%d\n%s\n" % (first_pair + i // 2, 1 + (i & 1), rows[i].tobytes()) for i in range(len(rows)))
    out, stage_s, nl = cport.pipeline_staged(inst["img"], inst["tax"], inst["opts"], fa, threads=threads)
    names = ("translate", "prot2kmer2lca", "seedextend", "uniq", "taxa2agg")
    nreads = 2 * npairs
    return {"reads_per_second_stages_in_sequence": nreads / sum(stage_s),
            "reads_per_second_as_concurrent_pipe": nreads / max(stage_s),
            "stage_seconds": {n: round(s, 4) for n, s in zip(names, stage_s)}, "pairs": npairs,
            "what": "ref_pipeline_staged: FASTA text between the stages, prot2kmer2lca on all threads in 240-record chunks, "
                    "every other stage single-threaded (the reference's process structure)"}


def cpu_baseline(args, inst=None, budget_s: float = 12.0):
    threads = os.cpu_count() or 1
    inst = inst or cpu_instance(args.cpu_index_keys, threads)
    dt, _, _ = cpu_run(inst, 0, 2000, threads)  # calibration, also warms the image
    npairs = int(min(400_000, max(2000, 2000 * budget_s / max(dt, 1e-4))))
    dt, nl, _ = cpu_run(inst, 10_000_000, npairs, threads)
    dl, nll, _ = cpu_run(inst, 10_000_000, npairs, threads, lookups_only=True)
    staged = cpu_staged(inst, 11_000_000, max(1000, npairs // 10), threads)
    return {
        "value": 2 * npairs / dt,
        "unit": UNIT,
        "lookups_per_second": nl / dt,
        "lookup_stage_alone_lookups_per_second": nll / dl,
        "cores": threads,
        "kind": "port",
        "sample": f"{npairs} pairs of the same generator against a host-resident fst image of {inst['n_keys']} keys "
                  f"({inst['fst_bytes'] / 1e6:.0f} MB, built in {inst['fst_build_s']:.0f} s; the 1e9-key index is built on the device only), "
                  f"whole pipeline chunk-parallel over {threads} threads (best-effort CPU: more parallel than the reference's "
                  f"process pipeline, no text between the stages), {dt:.1f} s",
        "reference_structure": staged,
    }


def run_reference(args):
    """Reference arm: the CPU restatement of the reference pipeline on every host thread, on our arm's config; each
    step is a bounded sample of the step's 1 M pairs."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    inst = cpu_instance(args.cpu_index_keys, threads)
    dt, _, _ = cpu_run(inst, 0, 2000, threads)
    # bounded sample per step: the whole run (warmup + steps) stays around a minute
    per_step = int(min(args.pairs_per_step, 100_000, max(1000, 2000 * (60.0 / max(1, args.steps + args.warmup)) / max(dt, 1e-4))))
    for w in range(args.warmup):
        cpu_run(inst, 1_000_000 + w * per_step, per_step, threads)
    total_t, total_l = 0.0, 0
    for s in range(args.steps):
        dt, nl, _ = cpu_run(inst, 20_000_000 + s * per_step, per_step, threads)
        total_t += dt
        total_l += nl
    value = 2 * per_step * args.steps / total_t
    staged = cpu_staged(inst, 30_000_000, max(1000, per_step // 10), threads)
    sample = (f"{per_step} pairs per step (a bounded sample of the step's {args.pairs_per_step} pairs) against a host-resident fst image of "
              f"{inst['n_keys']} keys ({inst['fst_bytes'] / 1e6:.0f} MB; the arm's 1e9-key index lives on the device only), {threads} threads; "
              "C restatement of the reference algorithm (the Rust binary cannot be built here: no cargo/rustc), whole chain chunk-parallel")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total_t / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8/u32 integer", "data": "synthetic",
        "lookups_per_second": total_l / total_t,
        "config": pipeline_config(args),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                         "pairs_per_step_sampled": per_step, "index_keys": inst["n_keys"], "reference_structure": staged},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(json.dumps(line))


# -------------------------------------------------------------------------------------------- clocks

class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int):
        self.device = device
        self.rows = []
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.device}", f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        sm = []
        reasons = set()
        for r in self.rows:
            if len(r) < 7:
                continue
            try:
                sm.append(float(r[0]))
                out["sm_max_mhz"] = float(r[1])
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if sm:
            busy = sorted(sm)[len(sm) // 2:]  # upper half: samples taken under load
            out["sm_mhz"] = busy[len(busy) // 2]
        out["reasons"] = sorted(reasons)
        out["samples"] = len(sm)
        return out


# --------------------------------------------------------------------------------------------- legs
# Extra measurements after the headline (BASELINE.json configs[2..4], the every-position kernel, parity checks); each
# returns a dict for the JSON line.  A leg that fails reports {"error": ...} instead of taking the headline down.

class Ctx:
    pass


def timed_steps(ctx, fn, steps, warmup=2):
    """CUDA events on the current stream around `steps` calls of fn(i), max over ranks; lookup / classify brackets."""
    torch, capi = ctx.torch, ctx.capi
    for w in range(warmup):
        fn(w)
    ctx.barrier()
    capi.kernel_timing(True)
    capi.kernel_times()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        fn(i)
    e1.record()
    ctx.barrier()
    ms = e0.elapsed_time(e1)
    lookup_ms, lookup_n, classify_ms, _ = capi.kernel_times()
    capi.kernel_timing(False)
    if ctx.world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=ctx.dev)
        ctx.dist.all_reduce(t, op=ctx.dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms / steps, lookup_ms / steps, classify_ms / steps


def all_ranks_true(ctx, ok: bool) -> bool:
    if ctx.world == 1:
        return bool(ok)
    t = ctx.torch.tensor([1 if ok else 0], dtype=ctx.torch.int32, device=ctx.dev)
    ctx.dist.all_reduce(t, op=ctx.dist.ReduceOp.MIN)
    return bool(t.item())


def leg_every_position(ctx, gidx, peak):
    """The plain kernel (translate_lookup_kernel: every k-mer position probed, the path of configs[0], of pipelines
    without seedextend / -o, and of k != 9) on the headline workload: umgap_pipeline_sampling(0)."""
    capi = ctx.capi
    before = capi.pipeline_sampling(0)
    try:
        steps = max(3, min(5, ctx.args.steps))
        step_ms, lookup_ms, classify_ms = timed_steps(
            ctx, lambda i: capi.classify_reads_dev(gidx, ctx.gtax, ctx.opts, ctx.batches[i % len(ctx.batches)].data_ptr(), ctx.roff.data_ptr(),
                                                   ctx.nreads, ctx.total_nt, ctx.goff.data_ptr(), ctx.B, ctx.out_b.data_ptr(), ctx.stream), steps)
    finally:
        capi.pipeline_sampling(before)
    alg = ctx.nreads * LOOKUPS_PER_READ * BYTES_PER_LOOKUP
    return {"value": ctx.world * ctx.nreads / (step_ms * 1e-3), "unit": UNIT, "ms_per_step": step_ms, "steps": steps,
            "lookup_stage_ms": lookup_ms, "classify_ms": classify_ms,
            "roofline_frac": alg / (lookup_ms * 1e-3) / 1e9 / peak, "lookups_per_second_kernel": ctx.nreads * LOOKUPS_PER_READ / (lookup_ms * 1e-3),
            "kernel": "translate_lookup_kernel<9,TableView> (all 248 positions of a read probed)"}


def leg_sustained(ctx, gidx):
    """The headline step repeated for seconds, not milliseconds: does the figure hold in power / thermal steady state?
    Clocks and throttle reasons are sampled over the whole leg."""
    capi, torch = ctx.capi, ctx.torch
    seconds = 3.0

    def step(i):
        capi.classify_reads_dev(gidx, ctx.gtax, ctx.opts, ctx.batches[i % len(ctx.batches)].data_ptr(), ctx.roff.data_ptr(), ctx.nreads,
                                ctx.total_nt, ctx.goff.data_ptr(), ctx.B, ctx.out_b.data_ptr(), ctx.stream)
    ms1, _, _ = timed_steps(ctx, step, 5)
    steps = int(max(50, min(2000, seconds * 1e3 / ms1)))
    sampler = ClockSampler(ctx.local)
    if ctx.rank == 0:
        sampler.start()
        time.sleep(0.2)
    ms, lookup_ms, classify_ms = timed_steps(ctx, step, steps, warmup=0)
    clocks = sampler.stop() if ctx.rank == 0 else None
    return {"value": ctx.world * ctx.nreads / (ms * 1e-3), "unit": UNIT, "steps": steps, "ms_per_step": ms, "timed_region_s": ms * steps * 1e-3,
            "lookup_stage_ms": lookup_ms, "classify_ms": classify_ms, "clocks": clocks}


def leg_parity_device(ctx, gidx):
    """Bit-equality on batch 0 of every rank: the sampled lookups against every position probed (seedextend sees the
    same extended seeds), and the packed host form against the device-resident form."""
    capi, torch = ctx.capi, ctx.torch
    def run():
        capi.classify_reads_dev(gidx, ctx.gtax, ctx.opts, ctx.batches[0].data_ptr(), ctx.roff.data_ptr(), ctx.nreads, ctx.total_nt,
                                ctx.goff.data_ptr(), ctx.B, ctx.out_b.data_ptr(), ctx.stream)
        torch.cuda.synchronize()
        return ctx.out_b.clone()
    sampled = run()
    before = capi.pipeline_sampling(0)
    try:
        plain = run()
    finally:
        capi.pipeline_sampling(before)
    ctx.replicated_out0 = sampled
    return {"sampled_equals_every_position": all_ranks_true(ctx, bool(torch.equal(sampled, plain))),
            "groups_compared_per_rank": int(ctx.B)}


def leg_sharded(ctx, gidx_repl):
    """BASELINE configs[4]: the index key-range-sharded over the GPUs, every lookup answered by the GPU that owns the
    hash.  The exchange step runs inside the kernels (exchange.cu: the pack kernels store hashes into the owners'
    inboxes over NVLink peer mappings, the lookup kernels store the answers back; epoch flags order the rounds) --
    no collective library and no host synchronisation inside a batch.  One process drives all the GPUs through
    umgap_sharded (the form the CLI / a Rust host uses): rank 0 does, the other ranks wait on the host."""
    capi, torch, dist = ctx.capi, ctx.torch, ctx.dist
    args = ctx.args
    if ctx.rank != 0:
        dist.barrier(group=ctx.gloo)
        return None
    res = None
    try:
        res = _sharded_rank0(ctx, gidx_repl)
    finally:
        torch.cuda.set_device(ctx.local)
        dist.barrier(group=ctx.gloo)
    return res


def _sharded_rank0(ctx, gidx_repl):
    import datagen
    capi, torch = ctx.capi, ctx.torch
    args = ctx.args
    G = ctx.world
    taxa_arrays = datagen.taxonomy_arrays(datagen.make_taxonomy(N_TAXA, seed=1))
    t0 = time.perf_counter()
    gtax, shards = [], []
    for d in range(G):
        gtax.append(capi.Taxonomy.from_arrays(*taxa_arrays, device=d))
        shards.append(capi.Index.build_synthetic(ctx.spec, gtax[d], device=d, load_factor=args.load_factor, shard=d, nshards=G))
    build_s = time.perf_counter() - t0
    nb = 3
    batches, roffs, goffs, outs = [], [], [], []
    for d in range(G):
        with torch.cuda.device(d):
            dev = torch.device("cuda", d)
            bs = []
            for b in range(nb):
                nt = torch.empty(ctx.total_nt, dtype=torch.uint8, device=dev)
                capi.synth_reads_dev(ctx.spec, 3, (d * 64 + b) * ctx.B, ctx.B, READ_LEN, HIT_PCT, nt.data_ptr())
                bs.append(nt)
            batches.append(bs)
            roffs.append(torch.arange(0, ctx.nreads + 1, dtype=torch.int64, device=dev) * READ_LEN)
            goffs.append(torch.arange(0, ctx.nreads + 1, 2, dtype=torch.int64, device=dev))
            outs.append([torch.zeros(ctx.B, dtype=torch.int32, device=dev) for _ in range(2)])   # consecutive batches are in flight together
            torch.cuda.synchronize()
    S = capi.Sharded(shards, gtax, ctx.total_nt)

    def step(i):
        S.classify_reads_dev(ctx.opts, [batches[d][i % nb].data_ptr() for d in range(G)], [r.data_ptr() for r in roffs],
                             [ctx.nreads] * G, [ctx.total_nt] * G, [g.data_ptr() for g in goffs], [ctx.B] * G,
                             [o[i % 2].data_ptr() for o in outs])

    # parity: batch 0 of every GPU against the replicated table (rank 0's replica)
    step(0)
    routed0 = S.sync()
    same = True
    torch.cuda.set_device(ctx.local)
    for d in range(G):
        nt0 = batches[d][0].to(ctx.dev)
        capi.classify_reads_dev(gidx_repl, ctx.gtax, ctx.opts, nt0.data_ptr(), ctx.roff.data_ptr(), ctx.nreads, ctx.total_nt,
                                ctx.goff.data_ptr(), ctx.B, ctx.out_b.data_ptr(), ctx.stream)
        torch.cuda.synchronize()
        same = same and bool(torch.equal(ctx.out_b, outs[d][0].to(ctx.dev)))
        del nt0
    steps = max(3, min(20, 2 * args.steps))   # consecutive batches overlap (lanes): enough of them for the steady state
    for i in range(2):
        step(i)
    S.sync()
    t0 = time.perf_counter()
    for i in range(steps):
        step(i)
    S.sync()
    step_ms = 1e3 * (time.perf_counter() - t0) / steps
    # where the time goes: CUDA-event brackets around the stages of two more batches, averaged over the GPUs
    capi.kernel_timing(True)
    capi.kernel_times_ex()
    for i in range(2):
        step(i)
    S.sync()
    stages = {k: round(ms / (2 * G), 3) for k, (ms, _) in capi.kernel_times_ex().items()}
    capi.kernel_timing(False)
    res = {"value": G * ctx.nreads / (step_ms * 1e-3), "unit": UNIT, "ms_per_step": step_ms, "steps": steps,
           "timing": "wall clock between umgap_sharded_sync calls around the timed batches (one process enqueues for every GPU; nothing "
                     "synchronises inside)",
           "stages_ms_per_step_per_gpu": stages, "lookups_routed_per_step": int(routed0),
           "shard_build_s_all": build_s, "shard_bytes": int(shards[0].info().bytes), "equals_replicated": bool(same),
           "exchange": "in-kernel: peer stores over NVLink (8 B per lookup out, 4 B back), epoch flags, no NCCL, no host read-back",
           "lanes": int(os.environ.get("UMGAP_SHARDED_LANES", "2")),
           "what": "index key-range-sharded over the GPUs; two exchange rounds per batch (sampled positions, then the live frames)"}
    S.close()
    for x in shards + gtax:
        x.close()
    return res


def leg_large_index(ctx, peak):
    """BASELINE configs[3]: a ~100 GB table replicated per GPU, reads partitioned (every rank classifies its own batches)."""
    capi, torch = ctx.capi, ctx.torch
    args = ctx.args
    n_prot = max(1, int(args.large_index_keys // (PROTEIN_LEN - 8)))
    spec = capi.SynthSpec(seed=2, n_proteins=n_prot, protein_len=PROTEIN_LEN, home_pct=70, ancestor_pct=20)
    free_b, _ = torch.cuda.mem_get_info()
    need = int(args.large_index_keys * 16.6) + 24 * ctx.total_nt * 4
    if free_b < need:
        return {"skipped": f"needs ~{need / 1e9:.0f} GB of free HBM, {free_b / 1e9:.0f} GB free"}
    t0 = time.perf_counter()
    big = capi.Index.build_synthetic(spec, ctx.gtax, device=ctx.local, load_factor=args.load_factor or 0.5)
    torch.cuda.synchronize()
    build_s = time.perf_counter() - t0
    info = big.info()
    nb = 3
    batches = []
    for b in range(nb):
        nt = torch.empty(ctx.total_nt, dtype=torch.uint8, device=ctx.dev)
        capi.synth_reads_dev(spec, 3, (ctx.rank * 64 + b) * ctx.B, ctx.B, READ_LEN, HIT_PCT, nt.data_ptr())
        batches.append(nt)
    torch.cuda.synchronize()
    # BASELINE configs[3] names the max-sensitivity preset (scripts/umgap-analyse.sh:277-282): seedextend -g1 -s2, taxa2agg -l1 -a mrtl
    maxsens = capi.default_opts(min_seed_size=2, max_gap_size=1, strategy=capi.AGG_MRTL, lower_bound=1.0)
    use = [maxsens]

    def step(i):
        capi.classify_reads_dev(big, ctx.gtax, use[0], batches[i % nb].data_ptr(), ctx.roff.data_ptr(), ctx.nreads, ctx.total_nt,
                                ctx.goff.data_ptr(), ctx.B, ctx.out_b.data_ptr(), ctx.stream)
    steps = max(3, min(5, args.steps))
    use[0] = ctx.opts   # the headline's options on the large table, for comparison with the 16 GB table
    hp_ms, hp_lookup_ms, hp_classify_ms = timed_steps(ctx, step, steps)
    use[0] = maxsens
    step_ms, lookup_ms, classify_ms = timed_steps(ctx, step, steps)
    sampled = ctx.out_b.clone()
    classified = float((ctx.out_b != 1).float().mean().item())
    before = capi.pipeline_sampling(0)
    try:
        step(steps - 1)
        torch.cuda.synchronize()
        same = bool(torch.equal(sampled, ctx.out_b))
    finally:
        capi.pipeline_sampling(before)
    region = 60 << 30
    alg = ctx.nreads * LOOKUPS_PER_READ * BYTES_PER_LOOKUP
    res = {"value": ctx.world * ctx.nreads / (step_ms * 1e-3), "unit": UNIT, "ms_per_step": step_ms, "steps": steps,
           "lookup_stage_ms": lookup_ms, "classify_ms": classify_ms, "index_bytes": int(info.bytes), "index_keys_resident": int(info.n_keys),
           "index_build_s": build_s, "index_load_factor": info.load_factor,
           "probe_regions": int(max(1, -(-int(info.bytes) // region))), "roofline_frac": alg / (lookup_ms * 1e-3) / 1e9 / peak,
           "classified_below_root_frac": classified, "sampled_equals_every_position": all_ranks_true(ctx, same),
           "pipeline": "max-sensitivity preset: translate -a | prot2kmer2lca -o | seedextend -g1 -s2 | uniq -d / | taxa2agg -l1 -m rmq -a mrtl",
           "with_the_headline_options": {"value": ctx.world * ctx.nreads / (hp_ms * 1e-3), "ms_per_step": hp_ms, "lookup_stage_ms": hp_lookup_ms,
                                         "classify_ms": hp_classify_ms, "roofline_frac": alg / (hp_lookup_ms * 1e-3) / 1e9 / peak},
           "what": "replicated per GPU, reads partitioned; level 0 probed one <= 60 GiB hash-prefix region per launch (address-translation reach)"}
    del batches
    big.close()
    return res


def tryptic_instance(n_prot: int, npairs: int, frag: int, taxa, seed: int = 5):
    """Synthetic proteome -> tryptic peptides of 9..45 residues valued with the protein's taxon (keys sorted, first value
    of a duplicate kept), and predicted-gene style fragments in pairs (70 % windows of one protein, 30 % random residues)."""
    rng = np.random.default_rng(seed)
    L = 400
    ids = np.array([t[0] for t in taxa], dtype=np.uint64)
    letters = np.frombuffer(b"ACDEFGHIKLMNPQRSTVWY", dtype=np.uint8)
    freq = np.array([8.3, 1.4, 5.5, 6.8, 3.9, 7.1, 2.3, 5.9, 5.8, 9.7, 2.4, 4.1, 4.7, 3.9, 5.5, 6.6, 5.4, 6.9, 1.1, 2.9])
    prot = letters[rng.choice(20, size=(n_prot, L), p=freq / freq.sum())]
    home = ids[rng.integers(0, len(ids), n_prot)]
    flat = prot.reshape(-1)
    prev = np.empty_like(flat)
    prev[1:] = flat[:-1]
    prev[0] = 0
    start = ((prev == ord("K")) | (prev == ord("R"))) & (flat != ord("P"))
    start[::L] = True
    pos = np.flatnonzero(start)
    nxt = np.empty_like(pos)
    nxt[:-1] = pos[1:]
    nxt[-1] = flat.size
    end = np.minimum(nxt, (pos // L + 1) * L)
    ln = end - pos
    keep = (ln >= 9) & (ln <= 45)
    kpos, klen = pos[keep], ln[keep]
    peps = {}
    fb = flat.tobytes()
    for p0, l0 in zip(kpos.tolist(), klen.tolist()):
        peps.setdefault(fb[p0:p0 + l0], int(home[p0 // L]))
    keys = sorted(peps)
    vals = np.array([peps[k] for k in keys], dtype=np.uint64)
    koff = np.zeros(len(keys) + 1, dtype=np.uint64)
    np.cumsum(np.fromiter((len(k) for k in keys), dtype=np.uint64, count=len(keys)), out=koff[1:])
    blob = np.frombuffer(b"".join(keys), dtype=np.uint8)
    nlines = 2 * npairs
    src = rng.integers(0, n_prot, npairs)
    hit = rng.random(npairs) < 0.7
    a = rng.integers(0, L - frag, nlines)
    lines = prot[np.repeat(src, 2)[:, None], (a[:, None] + np.arange(frag)[None, :])]
    noise = letters[rng.integers(0, 20, size=(nlines, frag))]
    lines = np.where(np.repeat(hit, 2)[:, None], lines, noise)
    aa = np.ascontiguousarray(lines.reshape(-1))
    loff = np.arange(nlines + 1, dtype=np.uint64) * frag
    goff = np.arange(0, nlines + 1, 2, dtype=np.uint64)
    return dict(blob=blob, koff=koff, vals=vals, aa=aa, loff=loff, goff=goff, n_keys=len(keys), nlines=nlines, frag=frag)


def leg_tryptic(ctx, taxa):
    """BASELINE configs[2]: prot2tryp2lca -l9 -L45 | uniq -d / | taxa2agg -l1 -a mrtl (the tryptic presets of
    scripts/umgap-analyse.sh:291-300) through umgap_classify_peptides[_dev]; the C port of the same chain as the CPU
    baseline, and a parity sample of the device results against it."""
    capi, torch = ctx.capi, ctx.torch
    from oracle import cport
    npairs, frag = 1_000_000, 50
    t0 = time.perf_counter()
    inst = tryptic_instance(250_000, npairs, frag, taxa)
    gen_s = time.perf_counter() - t0
    t0 = time.perf_counter()
    tidx = capi.Index.from_blob(inst["blob"], inst["koff"], inst["vals"], k=0, device=ctx.local)
    build_s = time.perf_counter() - t0
    opts = capi.tryp_opts(minlen=9, maxlen=45, strategy=capi.AGG_MRTL, lower_bound=1.0)
    aa, loff, goff, nlines = inst["aa"], inst["loff"], inst["goff"], inst["nlines"]
    d_aa = torch.from_numpy(aa).to(ctx.dev)
    d_loff = torch.from_numpy(loff.astype(np.int64)).to(ctx.dev)
    d_goff = torch.from_numpy(goff.astype(np.int64)).to(ctx.dev)
    d_out = torch.zeros(npairs, dtype=torch.int32, device=ctx.dev)
    def run(_i=0):
        capi.classify_peptides_dev(tidx, ctx.gtax, opts, d_aa.data_ptr(), d_loff.data_ptr(), nlines, aa.size, d_goff.data_ptr(), npairs,
                                   d_out.data_ptr(), ctx.stream)
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    dev_out = d_out.cpu().numpy().view(np.uint32)
    # end to end from pinned host buffers
    def pin(a):
        t = torch.from_numpy(a.view(np.int64) if a.dtype == np.uint64 else a).pin_memory()
        return t.numpy().view(a.dtype), t
    h_aa, _k1 = pin(aa)
    h_loff, _k2 = pin(loff)
    h_goff, _k3 = pin(goff)
    h_out, _k4 = pin(np.zeros(npairs, dtype=np.uint32))
    host_out = capi.classify_peptides(tidx, ctx.gtax, opts, h_aa, h_loff, h_goff, out=h_out)
    t0 = time.perf_counter()
    for _ in range(5):
        host_out = capi.classify_peptides(tidx, ctx.gtax, opts, h_aa, h_loff, h_goff, out=h_out)
    e2e_s = (time.perf_counter() - t0) / 5
    # CPU baseline + parity: the C port on the same peptides against an fst image of the same keys
    threads = os.cpu_count() or 1
    t0 = time.perf_counter()
    img = cport.FstImage(cport.fst_build_blob(inst["blob"], inst["koff"], inst["vals"]))
    fst_s = time.perf_counter() - t0
    copts = cport.RefTrypOpts(minlen=9, maxlen=45, keep=b"", drop=b"", strategy=2, factor=0.25, lower_bound=1.0, ranked_only=0)
    ns = npairs
    sub_aa, sub_loff, sub_goff = aa[: 2 * ns * frag], loff[: 2 * ns + 1], goff[: ns + 1]
    ctax = cport.RefTaxonomy(taxa)
    cport.classify_peptides(img, ctax, copts, sub_aa[: 2000 * frag], sub_loff[:2001], sub_goff[:1001], threads=threads)  # warm
    reps = 10
    t0 = time.perf_counter()
    for _ in range(reps):
        want, nl, nh = cport.classify_peptides(img, ctax, copts, sub_aa, sub_loff, sub_goff, threads=threads)
    cpu_s = (time.perf_counter() - t0) / reps
    # MRTL ties (equal maxima) are broken by HashMap order in the reference: count agreement, and require that a
    # disagreement is a tie (both answers have a hit and differ) -- never root-vs-taxon
    agree = float((want == dev_out[:ns]).mean())
    hard = int(((want != dev_out[:ns]) & ((want == 1) | (dev_out[:ns] == 1))).sum())
    res = {"workload": "tryptic presets: prot2tryp2lca -l9 -L45 | uniq -d / | taxa2agg -l1 -a mrtl on pairs of 50-residue predicted-gene "
                       "fragments (70 % from the indexed proteome), tryptic index of the synthetic proteome",
           "value": nlines / (ms * 1e-3), "unit": "peptide lines/s", "pairs_per_second": npairs / (ms * 1e-3), "ms_per_step": ms,
           "pairs_per_step": npairs, "index_keys_resident": int(tidx.info().n_keys), "index_bytes": int(tidx.info().bytes),
           "index_build_s": build_s, "instance_generation_s": gen_s,
           "e2e": {"value": nlines / e2e_s, "unit": "peptide lines/s", "ms_per_step": 1e3 * e2e_s, "h2d_bytes_per_step": int(aa.nbytes + loff.nbytes + goff.nbytes),
                   "d2h_bytes_per_step": int(4 * npairs), "entry_point": "umgap_classify_peptides: page-locked host arrays in and out, ranges of whole groups on rotating streams"},
           "host_equals_device": bool(np.array_equal(host_out, dev_out)),
           "cpu_baseline": {"value": 2 * ns / cpu_s, "unit": "peptide lines/s", "cores": threads, "kind": "port", "lookups_per_second": nl / cpu_s,
                            "sample": f"{reps} x {ns} pairs, ref_classify_peptides (oracle/c) on {threads} threads against an fst image of the same "
                                      f"{inst['n_keys']} peptides (built in {fst_s:.1f} s), {cpu_s * reps:.1f} s"},
           "parity_vs_c_port": {"groups": ns, "equal_frac": agree, "root_vs_taxon_disagreements": hard,
                                "note": "differences are MRTL ties (rmq/rtl.rs:52-55 breaks them by HashMap order)"},
           "classified_below_root_frac": float((dev_out != 1).mean())}
    tidx.close()
    return res


def leg_cpu_parity_and_loader(ctx, inst):
    """(1) The fst loader: the CPU legs' image (1e8 keys) written to a file and streamed into a device table by
    umgap_index_load_fst -- keys/s of the loader.  (2) Parity against the oracle on the same index: a 2 000-pair slice
    through umgap_classify_reads against oracle/c (LCA*: no ties, must be equal; hybrid: equal or a tie) and 300 pairs
    against the line-by-line Python oracle's admissible sets, which also gives the fraction of reads with a tie."""
    capi, torch = ctx.capi, ctx.torch
    from oracle import cport, lookup as olookup, pipeline as opipe
    from oracle.taxonomy import Taxonomy as OTaxonomy
    import tempfile
    res = {}
    with tempfile.NamedTemporaryFile(suffix=".fst", dir="/tmp") as f:
        f.write(inst["raw"])
        f.flush()
        t0 = time.perf_counter()
        lidx = capi.Index.load_fst(f.name, k=9, device=ctx.local)
        torch.cuda.synchronize()
        load_s = time.perf_counter() - t0
    info = lidx.info()
    res["fst_loader"] = {"index_load_s": load_s, "keys": int(info.n_keys), "keys_per_second": info.n_keys / load_s,
                         "fst_bytes": inst["fst_bytes"], "bytes_per_second": inst["fst_bytes"] / load_s,
                         "what": "umgap_index_load_fst on the CPU legs' fst image (page cache warm)"}
    npairs = 2000
    nt, off, goff = inst["reads"](40_000_000, npairs)
    checks = {}
    for name, strategy in (("lca_star", 0), ("hybrid", 1)):
        o = capi.default_opts(min_seed_size=3, max_gap_size=0, strategy=strategy, factor=0.25)
        got, _ = capi.classify_reads(lidx, ctx.gtax, o, nt, off, goff)
        inst["opts"].strategy = strategy
        want, _, _ = cport.classify(inst["img"], inst["tax"], inst["opts"], nt, off, goff, threads=os.cpu_count() or 1)
        checks[name] = {"groups": npairs, "equal": int((got == want).sum())}
    inst["opts"].strategy = 1
    # the line-by-line oracle on the first 300 pairs: admissible sets (ties of tree/mix.rs:52-55)
    ns = 300
    otax = OTaxonomy(inst["taxa"])
    rows = nt.reshape(-1, READ_LEN)
    reads = [(f"r{i // 2}/{1 + (i & 1)}", rows[i].tobytes().decode()) for i in range(2 * ns)]
    sets = opipe.classify_reads(reads, inst["img"], otax, min_seed_size=3, max_gap_size=0, strategy=1, factor=0.25)
    o = capi.default_opts(min_seed_size=3, max_gap_size=0, strategy=capi.AGG_HYBRID, factor=0.25)
    got, _ = capi.classify_reads(lidx, ctx.gtax, o, nt[: 2 * ns * READ_LEN], off[: 2 * ns + 1], goff[: ns + 1])
    inside = sum(int(g) in adm for g, (_, adm) in zip(got, sets))
    ties = sum(len(adm) > 1 for _, adm in sets)
    checks["hybrid_vs_python_oracle_sets"] = {"groups": ns, "inside_admissible_set": inside, "groups_with_a_tie": ties}
    res["oracle_slice"] = checks
    res["tie_fraction"] = ties / ns
    res["ok"] = (checks["lca_star"]["equal"] == npairs and inside == ns)
    lidx.close()
    return res


# --------------------------------------------------------------------------------------------- ours

def run_ours(args):
    import torch
    import torch.distributed as dist
    import datagen
    from umgap_b200 import capi

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if capi.device_count() <= 0:
        raise SystemExit("bench.py needs a CUDA device: " + capi.load_library().umgap_last_error().decode())
    torch.cuda.set_device(local)
    gloo = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        gloo = dist.new_group(backend="gloo")   # host-side waits that keep the GPUs free (the sharded leg)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- instance: taxonomy, device-built index (replicated), device-generated reads
    taxa = datagen.make_taxonomy(N_TAXA, seed=1)
    gtax = capi.Taxonomy.from_arrays(*datagen.taxonomy_arrays(taxa), device=local)
    n_prot = max(1, int(args.index_keys // (PROTEIN_LEN - 8)))
    spec = capi.SynthSpec(seed=2, n_proteins=n_prot, protein_len=PROTEIN_LEN, home_pct=70, ancestor_pct=20)
    t0 = time.perf_counter()
    shard_mode = args.sharded and world > 1
    gidx = capi.Index.build_synthetic(spec, gtax, device=local, load_factor=args.load_factor,
                                      shard=rank if shard_mode else 0, nshards=world if shard_mode else 1)
    if shard_mode:
        from umgap_b200 import sharded
        if args.peer_loads:
            sharded.attach_all(gidx, dist)
    torch.cuda.synchronize()
    build_s = time.perf_counter() - t0
    info = gidx.info()
    B = args.pairs_per_step
    nreads = 2 * B
    total_nt = nreads * READ_LEN
    nbatches = max(1, min(10, args.steps))  # distinct resident batches, cycled by the steps
    dev = torch.device("cuda", local)
    batches = []
    for b in range(nbatches):
        nt = torch.empty(total_nt, dtype=torch.uint8, device=dev)
        first_pair = (rank * 64 + b) * B  # every rank and batch draws different pairs
        capi.synth_reads_dev(spec, 3, first_pair, B, READ_LEN, HIT_PCT, nt.data_ptr())
        batches.append(nt)
    roff = torch.arange(0, nreads + 1, dtype=torch.int64, device=dev) * READ_LEN
    goff = torch.arange(0, nreads + 1, 2, dtype=torch.int64, device=dev)
    out = torch.zeros(B, dtype=torch.int32, device=dev)
    opts = capi.default_opts(min_seed_size=3, max_gap_size=0, strategy=capi.AGG_HYBRID, factor=0.25)
    stream = torch.cuda.current_stream().cuda_stream
    torch.cuda.synchronize()

    routed = None
    if shard_mode and not args.peer_loads:
        routed = sharded.RoutedClassifier(gidx, gtax, dist, total_nt, lanes=args.routed_lanes)

    def step(i):
        if routed is not None:  # hashes and answers cross NVLink in two all-to-alls, lookups stay local
            routed.classify(opts, batches[i % nbatches], roff, goff, out, total_nt)
            return
        capi.classify_reads_dev(gidx, gtax, opts, batches[i % nbatches].data_ptr(), roff.data_ptr(), nreads, total_nt,
                                goff.data_ptr(), B, out.data_ptr(), stream)

    # ---- device-resident throughput (clocks are sampled from the warm-up to the end of the timed region)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)  # let nvidia-smi start polling
    for w in range(args.warmup):
        step(w)
    barrier()
    capi.kernel_timing(True)
    capi.kernel_times()
    launches0 = capi.kernel_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for s in range(args.steps):
        step(s)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    lookup_ms, lookup_n, classify_ms, classify_n = capi.kernel_times()
    launches = capi.kernel_launch_count() - launches0
    # the same kernels timed alone (one lookup stage and one classify launch per step, nothing beside them)
    slices = capi.pipeline_slices(1)
    alone_steps = max(1, min(3, args.steps))
    for s in range(alone_steps):
        step(s)
    barrier()
    capi.kernel_times()
    for s in range(alone_steps):
        step(s)
    barrier()
    alone_lookup_ms, _, alone_classify_ms, _ = capi.kernel_times()
    capi.pipeline_slices(slices)
    capi.kernel_timing(False)
    routed_stages = None
    if routed is not None:   # where a routed step spends its time: a few extra steps with event brackets around the stages
        routed.profile = True
        for s in range(alone_steps):
            step(s)
        routed.profile = False
        routed_stages = {k: round(v / alone_steps, 3) for k, v in routed.stage_ms.items()}
        routed_stages["lookups_routed_per_step"] = routed.lookups_routed
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    classified = float((out != 1).float().mean().item())
    # the observed k-mer hit rate of the workload (every position of batch 0 looked up once, untimed)
    kmer_hit_rate = None
    if rank == 0 and not shard_mode:
        ids_all = torch.full((2 * total_nt + 64,), -1, dtype=torch.int32, device=dev)
        capi.translate_lookup_dev(gidx, opts, batches[0].data_ptr(), roff.data_ptr(), nreads, total_nt, ids_all.data_ptr(), stream)
        torch.cuda.synchronize()
        per_read = ids_all[: 2 * total_nt].view(nreads, 2 * READ_LEN)
        npos = READ_LEN - 3 * K + 1
        hits = int((per_read[:, :npos] != -1).sum().item()) + int((per_read[:, READ_LEN:READ_LEN + npos] != -1).sum().item())
        kmer_hit_rate = hits / float(nreads * LOOKUPS_PER_READ)
        del ids_all, per_read
    # the random-sector ceiling of this table on this GPU (uniform random 32-byte gathers over level 0)
    rand_sectors_per_s = gidx.randsector_rate(1 << 28, 3) if rank == 0 and not shard_mode else None

    # ---- end to end through the host-buffer C ABI call (pinned host input, H2D + D2H timed)
    e2e = None
    if not args.no_e2e and routed is None:
        h_nt = torch.empty(total_nt, dtype=torch.uint8).pin_memory()
        h_nt.copy_(batches[0])
        def pinned(a):  # page-locked copy, so that every transfer of the timed region is truly asynchronous
            t = torch.from_numpy(a.view(np.int64) if a.dtype == np.uint64 else a.view(np.int32)).pin_memory()
            return t.numpy().view(a.dtype), t

        h_roff, _k1 = pinned(np.arange(0, nreads + 1, dtype=np.uint64) * READ_LEN)
        h_goff, _k2 = pinned(np.arange(0, nreads + 1, 2, dtype=np.uint64))
        h_out, _k3 = pinned(np.zeros(B, dtype=np.uint32))
        nt_np = h_nt.numpy()
        host_out, _ = capi.classify_reads(gidx, gtax, opts, nt_np, h_roff, h_goff)  # warm-up (allocates workspaces)
        dev_out = None
        step(0)
        torch.cuda.synchronize()
        dev_out = out.cpu().numpy().view(np.uint32)
        if not np.array_equal(dev_out, host_out):
            raise SystemExit("host-buffer and device-resident entry points disagree")
        e2e_steps = max(3, min(args.steps, 5))

        def timed_e2e(call, calls=None, steps_per_call=1):
            calls = e2e_steps if calls is None else calls
            n = calls * steps_per_call
            barrier()
            moved0 = capi.transfer_bytes()
            t0 = time.perf_counter()
            for _ in range(calls):
                call()
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            moved1 = capi.transfer_bytes()
            if world > 1:
                t = torch.tensor([dt], dtype=torch.float64, device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                dt = float(t.item())
            return {"value": world * nreads * n / dt, "unit": UNIT,
                    "h2d_bytes_per_step": int((moved1[0] - moved0[0]) // n), "d2h_bytes_per_step": int((moved1[1] - moved0[1]) // n),
                    "steps": n, "ms_per_step": 1e3 * dt / n}

        # (a) the byte form: one byte per nucleotide over PCIe (umgap_classify_reads)
        e2e_bytes = timed_e2e(lambda: capi.classify_reads(gidx, gtax, opts, nt_np, h_roff, h_goff, count_lookups=False, out=h_out))
        e2e_bytes["entry_point"] = "umgap_classify_reads: pinned host arrays, one byte per nucleotide"
        e2e_bytes["host_input_bytes_per_step"] = int(total_nt + h_roff.nbytes + h_goff.nbytes)
        # (b) the packed form (umgap_classify_reads_packed): 2 bits per nucleotide + N flags, a form a parser can
        #     emit; the packing (umgap_pack_reads) is timed beside it, not inside: it is the parser's job, once per read
        nw = (total_nt + 15) // 16
        p_codes = torch.empty(nw, dtype=torch.int32).pin_memory()
        p_entries = torch.empty(max(1, nw // 8), dtype=torch.int64).pin_memory()
        codes_np = p_codes.numpy().view(np.uint32)
        t0 = time.perf_counter()
        codes_np, entries_np = capi.pack_reads(nt_np, 0, codes=codes_np, entries=p_entries.numpy().view(np.uint64))
        pack_s = time.perf_counter() - t0
        packed_out, _ = capi.classify_reads_packed(gidx, gtax, opts, codes_np, entries_np, h_roff, h_goff)
        if not np.array_equal(dev_out, packed_out):
            raise SystemExit("packed and device-resident entry points disagree")
        e2e_sync = timed_e2e(lambda: capi.classify_reads_packed(gidx, gtax, opts, codes_np, entries_np, h_roff, h_goff, count_lookups=False, out=h_out))
        e2e_sync["entry_point"] = "umgap_classify_reads_packed, one batch at a time (the call returns with the results)"
        # (c) the same batches through the asynchronous form, two in flight: batch i + 1 is handed over before the wait
        #     for batch i, as a streaming host does (the CLI's parser threads) -- every step still uploads its own input
        #     from pinned host memory and its results are read back and waited for inside the timed region
        h_out2, _k4 = pinned(np.zeros(B, dtype=np.uint32))
        outs = [h_out, h_out2]
        n_pipe = max(20, args.steps)
        depth = max(2, args.e2e_depth)
        _extra = [pinned(np.zeros(B, dtype=np.uint32)) for _ in range(depth - 2)]
        outs = outs + [a_ for a_, _t in _extra]
        submit_s = [0.0]

        def pipelined():
            pend = []
            for i in range(n_pipe):
                t_s = time.perf_counter()
                pend.append(capi.classify_reads_packed_async(gidx, gtax, opts, codes_np, entries_np, h_roff, h_goff, out=outs[i % depth]))
                submit_s[0] += time.perf_counter() - t_s
                if len(pend) >= depth:
                    pend.pop(0).wait()
            for p_ in pend:
                p_.wait()

        pipelined()  # warm-up
        if not all(np.array_equal(o_, dev_out) for o_ in outs):
            raise SystemExit("asynchronous and device-resident entry points disagree")
        submit_s[0] = 0.0
        e2e = timed_e2e(pipelined, calls=1, steps_per_call=n_pipe)
        e2e["batches_in_flight"] = depth
        e2e["host_submit_ms_per_step"] = 1e3 * submit_s[0] / n_pipe
        e2e["entry_point"] = ("umgap_classify_reads_packed_async + umgap_pending_wait, two batches in flight: pinned host arrays, 2-bit "
                              "nucleotides + the list of 16-nucleotide words that hold an N (a form a parser can emit directly; "
                              "umgap_pack_reads makes it from bytes)")
        e2e["one_batch_at_a_time"] = e2e_sync
        e2e["host_input_bytes_per_step"] = int(codes_np.nbytes + entries_np.nbytes + h_roff.nbytes + h_goff.nbytes)
        e2e["note"] = ("bytes counted by the library around its cudaMemcpyAsync calls; the offset arrays of this workload are arithmetic "
                       "progressions (reads of one length, pairs), which the library detects and regenerates on the device instead of uploading")
        e2e["pack_reads_ms_per_step_untimed"] = 1e3 * pack_s
        e2e["pack_reads_threads"] = os.cpu_count()
        e2e["byte_form"] = e2e_bytes
        if world == 1:
            # (d) from bytes with the packing inside the timed region: umgap_pack_reads (all host threads) into one of three
            #     page-locked sets, the packed batch handed over asynchronously, the batch that used the set before waited
            #     for first -- the host packs batch i while the GPU classifies batch i - 1
            sets = [(codes_np, p_entries.numpy().view(np.uint64))]
            _keep = []
            for _ in range(2):
                c_t, e_t = torch.empty(nw, dtype=torch.int32).pin_memory(), torch.empty(max(1, nw // 8), dtype=torch.int64).pin_memory()
                _keep.append((c_t, e_t))
                sets.append((c_t.numpy().view(np.uint32), e_t.numpy().view(np.uint64)))
            outs3 = [h_out, h_out2, pinned(np.zeros(B, dtype=np.uint32))[0]]
            pack_in_s = [0.0, 0.0, 0.0]
            pack_threads = int(os.environ.get("UMGAP_BENCH_PACK_THREADS", "0"))

            def from_bytes():
                # a packer thread (umgap_pack_reads spreads over the host threads itself) and this thread, which hands the
                # packed sets over and waits for their results: the structure of a streaming host (the CLI's parser and
                # classifier threads)
                import queue
                import threading
                free_q, ready_q = queue.Queue(), queue.Queue()
                for k_ in range(3):
                    free_q.put(k_)

                def packer():
                    for _i in range(n_pipe):
                        k_ = free_q.get()
                        t_s = time.perf_counter()
                        c_, e_ = capi.pack_reads(nt_np, pack_threads, codes=sets[k_][0], entries=sets[k_][1])
                        pack_in_s[0] += time.perf_counter() - t_s
                        ready_q.put((k_, c_, e_))

                th = threading.Thread(target=packer)
                th.start()
                pend = []
                for _i in range(n_pipe):
                    k_, c_, e_ = ready_q.get()
                    t_p = time.perf_counter()
                    pend.append((k_, capi.classify_reads_packed_async(gidx, gtax, opts, c_, e_, h_roff, h_goff, out=outs3[k_])))
                    t_w = time.perf_counter()
                    pack_in_s[2] += t_w - t_p
                    if len(pend) >= 2:
                        k0, p0 = pend.pop(0)
                        p0.wait()
                        free_q.put(k0)
                    pack_in_s[1] += time.perf_counter() - t_w
                for k0, p0 in pend:
                    p0.wait()
                th.join()

            from_bytes()
            if not all(np.array_equal(o_, dev_out) for o_ in outs3):
                raise SystemExit("the packed-on-the-fly path and the device-resident entry point disagree")
            pack_in_s[:] = [0.0, 0.0, 0.0]
            fb = timed_e2e(from_bytes, calls=1, steps_per_call=n_pipe)
            fb["pack_ms_per_step_inside"] = 1e3 * pack_in_s[0] / n_pipe
            fb["host_wait_ms_per_step"] = 1e3 * pack_in_s[1] / n_pipe
            fb["host_submit_ms_per_step"] = 1e3 * pack_in_s[2] / n_pipe
            fb["entry_point"] = ("umgap_pack_reads (bytes -> 2-bit codes + N words, all host threads) on a packer thread + "
                                 "umgap_classify_reads_packed_async on this one, three page-locked sets, two batches in flight: the host "
                                 "packs a batch while the GPU classifies the ones before it")
            e2e["from_bytes_packed_on_the_fly"] = fb

    was_routed = routed is not None
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak = float(json.load(open(peaks_path))["hbm_gbs"])
        peak_src = "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)"
    else:
        peak, peak_src = 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"

    # ---- legs beyond the headline: every rank takes part in the device legs, rank 0 alone runs the CPU-side ones
    want = {"sustained", "every_position", "parity", "sharded", "large_index", "tryptic", "loader"} if args.legs == "all" else \
        set() if args.legs == "none" else set(args.legs.split(","))
    ctx = Ctx()
    ctx.torch, ctx.capi, ctx.dist, ctx.args = torch, capi, dist, args
    ctx.rank, ctx.world, ctx.local, ctx.dev, ctx.barrier = rank, world, local, dev, barrier
    ctx.gtax, ctx.spec, ctx.opts, ctx.batches, ctx.roff, ctx.goff = gtax, spec, opts, batches, roff, goff
    ctx.nreads, ctx.total_nt, ctx.B, ctx.stream = nreads, total_nt, B, stream
    ctx.out_b = torch.zeros(B, dtype=torch.int32, device=dev)
    ctx.replicated_out0 = None
    ctx.gloo = gloo
    legs = {}

    def leg(name, fn, *a):
        if name not in want:
            return
        t0 = time.perf_counter()
        try:
            legs[name] = fn(*a)
        except Exception as e:  # a failed leg is reported, the headline stands
            legs[name] = {"error": f"{type(e).__name__}: {e}"}
        if isinstance(legs[name], dict):
            legs[name]["leg_wall_s"] = round(time.perf_counter() - t0, 1)
        else:
            legs.pop(name)
        barrier()

    if not shard_mode:
        leg("sustained", leg_sustained, ctx, gidx)
        leg("every_position", leg_every_position, ctx, gidx, peak)
        leg("parity", leg_parity_device, ctx, gidx)
        if world > 1:
            leg("sharded", leg_sharded, ctx, gidx)
    # the large table needs the HBM the headline instance holds
    routed = None
    del batches
    ctx.batches = None
    gidx.close()
    torch.cuda.empty_cache()
    if not shard_mode:
        leg("large_index", leg_large_index, ctx, peak)
    torch.cuda.empty_cache()

    if rank != 0:
        if world > 1:
            dist.barrier()   # rank 0's CPU-side legs
            dist.destroy_process_group()
        return

    leg_inst = None
    if "tryptic" in want:
        t0 = time.perf_counter()
        try:
            legs["tryptic"] = leg_tryptic(ctx, taxa)
        except Exception as e:
            legs["tryptic"] = {"error": f"{type(e).__name__}: {e}"}
        legs["tryptic"]["leg_wall_s"] = round(time.perf_counter() - t0, 1)
    if not args.no_cpu_baseline:
        leg_inst = cpu_instance(args.cpu_index_keys, os.cpu_count() or 1)
    if "loader" in want and leg_inst is not None:
        t0 = time.perf_counter()
        try:
            legs["oracle_parity_and_loader"] = leg_cpu_parity_and_loader(ctx, leg_inst)
        except Exception as e:
            legs["oracle_parity_and_loader"] = {"error": f"{type(e).__name__}: {e}"}
        legs["oracle_parity_and_loader"]["leg_wall_s"] = round(time.perf_counter() - t0, 1)

    # ---- roofline of the dominant kernel (lookup stage)
    lookups_per_launch = nreads * LOOKUPS_PER_READ
    # a step's lookup stage = the once-per-batch bracket plus one bracket per slice (routed mode: one launch per exchange round)
    avg_ms = max(lookup_ms / max(1, args.steps), 1e-9)   # routed mode: the local-shard lookups of both exchange rounds
    achieved = lookups_per_launch * BYTES_PER_LOOKUP / (avg_ms * 1e-3) / 1e9
    alone_ms = max(alone_lookup_ms / alone_steps, 1e-9)
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "k1_traffic.json")
    if os.path.exists(tpath):
        tj = json.load(open(tpath))
        if tj.get("pairs_per_launch") == B:
            traffic = tj.get("dram_bytes_per_launch")
    lines = None
    if os.path.exists(tpath) and traffic:
        lines = tj.get("lookup_kernel_dram_bytes_per_launch", traffic) / 128.0 / nreads
    alg_bytes = lookups_per_launch * BYTES_PER_LOOKUP
    if was_routed or slices == 1:
        use_ms = avg_ms if was_routed else max(lookup_ms / max(1, lookup_n), 1e-9)
        mode = "CUDA-event brackets around every launch of the lookup stage in the timed region, on its stream"
    else:
        # In the timed region a step is cut into slices whose lookup kernels run beside the previous slice's classify
        # kernel on two streams: their event brackets overlap each other and include the time a kernel waits for SMs the
        # other stream holds, so they do not measure the kernel.  The roofline therefore uses the same kernels on the
        # same batches timed alone (umgap_pipeline_slices(1): one lookup stage, then one classify launch per step),
        # measured in this run right after the timed region; the in-region bracket sum is reported beside it.
        use_ms, mode = alone_ms, "lookup stage timed alone on the timed region's batches (umgap_pipeline_slices(1)), CUDA events on its stream"
    achieved = alg_bytes / (use_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": "lookup_hashes_kernel (local shard)" if was_routed
                else "lookup stage = lookup_sampled_kernel<9,TableView,3,0> (translation inside the kernel; + the long-read pass of translate_lookup_kernel), one event bracket",
                "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic,
                "traffic_source": None if not traffic else "committed ncu --set full capture, not measured in this run: " + str(tj.get("source", ""))[:160],
                "peak_source": peak_src, "timing": mode,
                "note": "algorithmic bytes = 248 k-mers x 32 B per read (SURVEY 8(d)); in front of seedextend -s3 the kernel probes every "
                        "third position first and the rest only for frames with a hit (bit-identical output), so the HBM actually moves "
                        "`traffic` bytes: 128-byte line fills at the random-line ceiling (profiles/README.md)",
                "dram_lines_per_read": lines,
                "algorithmic_bytes_per_launch": alg_bytes, "avg_launch_ms": use_ms,
                "classify_kernel_ms_alone": alone_classify_ms / alone_steps if not was_routed else None,
                "in_timed_region": {"slices_per_step": slices, "lookup_bracket_sum_ms_per_step": avg_ms,
                                    "classify_bracket_sum_ms_per_step": classify_ms / args.steps,
                                    "step_ms": ms / args.steps,
                                    "what": "one lookup stage and one classify launch per step, one after the other" if slices == 1
                                            else "brackets of concurrent streams overlap: their sum exceeds the step"},
                "random_sector_roofline": None if not rand_sectors_per_s or not lines else {
                    "gathers_per_second": rand_sectors_per_s,
                    "what": "bench/randsector microbenchmark (umgap_randsector_bench): uniform random 32-byte gathers over this table; "
                            "every gather fills a 128-byte line",
                    "kernel_line_fills_per_second": lines * nreads / (use_ms * 1e-3),
                    "frac_line_fills": lines * nreads / (use_ms * 1e-3) / rand_sectors_per_s,
                    "frac_effective_lookups": lookups_per_launch / (use_ms * 1e-3) / rand_sectors_per_s},
                "kernel_share_of_step": min(1.0, use_ms / (ms / args.steps)) if ms else None,
                "lookups_per_second_kernel": lookups_per_launch / (use_ms * 1e-3)}
    cb = None if args.no_cpu_baseline else cpu_baseline(args, leg_inst)
    reads_total = world * nreads * args.steps
    line = {
        "metric": METRIC, "value": reads_total / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8/u32 integer", "data": "synthetic",
        "lookups_per_second": reads_total * LOOKUPS_PER_READ / (ms * 1e-3),
        "config": pipeline_config(args),
        "workload_stats": {"index_keys_resident": int(info.n_keys), "index_bytes": int(info.bytes), "index_build_s": build_s,
                           "index_load_factor": info.load_factor, "flagged_sector_frac": info.n_flagged / max(1, info.n_buckets),
                           "classified_below_root_frac": classified, "kmer_hit_rate": kmer_hit_rate,
                           "tie_fraction": (legs.get("oracle_parity_and_loader") or {}).get("tie_fraction")},
        "roofline": roofline, "cpu_baseline": cb, "e2e": e2e, "clocks": clocks,
        "gpu_launches": int(launches),
        "kernel_ms": {"translate_lookup": lookup_ms, "classify": classify_ms},
    }
    if routed_stages is not None:
        line["routed_stages_ms_per_step"] = routed_stages
    line.update({("parity_check" if k == "parity" else k): v for k, v in legs.items()})
    if "parity_check" in line and isinstance(line["parity_check"], dict):
        pc = line["parity_check"]
        if isinstance(line.get("sharded"), dict):
            pc["sharded_equals_replicated"] = line["sharded"].get("equals_replicated")
        if isinstance(line.get("oracle_parity_and_loader"), dict):
            pc["oracle_slice"] = line["oracle_parity_and_loader"].get("oracle_slice")
            pc["oracle_slice_ok"] = line["oracle_parity_and_loader"].get("ok")
    emit(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    global _REAL_STDOUT
    args = parse_args()
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
