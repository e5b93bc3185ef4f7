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
    ap.add_argument("--cpu-index-keys", type=float, default=2e7, help="index size of the host-resident CPU legs")
    ap.add_argument("--load-factor", type=float, default=0.0, help="table load factor (0 = library default policy)")
    ap.add_argument("--sharded", action="store_true", help="key-range-shard the index over the GPUs (peer-memory lookups) instead of replicating it")
    ap.add_argument("--peer-loads", action="store_true", help="with --sharded: read remote shards through IPC peer mappings instead of the all-to-all exchange")
    ap.add_argument("--routed-lanes", type=int, default=1, help="with --sharded: group ranges of a batch whose exchange rounds alternate on separate streams")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
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

def cpu_instance(n_keys: float):
    """Host-resident instance of the same generator: taxonomy, fst image, and a reads() closure."""
    import datagen
    from oracle import cport, synth
    taxa = datagen.make_taxonomy(N_TAXA, seed=1)
    pre = synth.Preorder(taxa)
    n_prot = max(1, int(n_keys // (PROTEIN_LEN - 8)))
    keys, vals = synth.build_index(2, n_prot, PROTEIN_LEN, 70, 20, pre)
    img = cport.FstImage(cport.fst_build_blob(keys.reshape(-1), np.arange(0, 9 * len(keys) + 1, 9, dtype=np.uint64), vals))
    ctax = cport.RefTaxonomy(taxa)
    opts = cport.RefOpts(table=1, methionine=0, one_on_one=1, seedextend=1, min_seed_size=3, max_gap_size=0,
                         strategy=1, factor=0.25, lower_bound=0.0, ranked_only=0, k=K)

    def reads(first_pair: int, npairs: int):
        parts = [synth.reads(2, n_prot, PROTEIN_LEN, 3, first_pair + c, min(50_000, npairs - c), READ_LEN, HIT_PCT)
                 for c in range(0, npairs, 50_000)]
        nt = np.concatenate(parts).reshape(-1)
        off = np.arange(0, len(nt) + 1, READ_LEN, dtype=np.uint64)
        goff = np.arange(0, 2 * npairs + 1, 2, dtype=np.uint64)
        return nt, off, goff

    return dict(img=img, tax=ctax, opts=opts, reads=reads, n_keys=len(keys), fst_bytes=len(img.data))


def cpu_run(inst, first_pair: int, npairs: int, threads: int):
    from oracle import cport
    nt, off, goff = inst["reads"](first_pair, npairs)
    t0 = time.perf_counter()
    out, nl, nh = cport.classify(inst["img"], inst["tax"], inst["opts"], nt, off, goff, threads=threads)
    dt = time.perf_counter() - t0
    return dt, nl, out


def cpu_baseline(args, budget_s: float = 12.0):
    threads = os.cpu_count() or 1
    inst = cpu_instance(args.cpu_index_keys)
    dt, _, _ = cpu_run(inst, 0, 2000, threads)  # calibration, also warms the image
    npairs = int(min(400_000, max(2000, 2000 * budget_s / max(dt, 1e-4))))
    dt, nl, _ = cpu_run(inst, 10_000_000, npairs, threads)
    return {
        "value": 2 * npairs / dt,
        "unit": UNIT,
        "lookups_per_second": nl / dt,
        "cores": threads,
        "kind": "port",
        "sample": f"{npairs} pairs of the same generator against a host-resident fst image of {inst['n_keys']} keys "
                  f"({inst['fst_bytes'] / 1e6:.0f} MB; the 1e9-key index is built on the device only), whole pipeline "
                  f"chunk-parallel over {threads} threads, {dt:.1f} s",
    }


def run_reference(args):
    """Reference arm: the CPU restatement of the reference pipeline on every host thread."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    inst = cpu_instance(args.cpu_index_keys)
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
    sample = (f"{per_step} pairs per step against a host-resident fst image of {inst['n_keys']} keys "
              f"({inst['fst_bytes'] / 1e6:.0f} MB), {threads} threads")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total_t / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8/u32 integer", "data": "synthetic",
        "lookups_per_second": total_l / total_t,
        "config": pipeline_config(args, {"pairs_per_step": per_step, "note": "C restatement of the reference algorithm "
                                         "(the Rust binary cannot be built here: no cargo/rustc); " + sample}),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
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
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

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

        def timed_e2e(call):
            barrier()
            moved0 = capi.transfer_bytes()
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                call()
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            moved1 = capi.transfer_bytes()
            if world > 1:
                t = torch.tensor([dt], dtype=torch.float64, device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                dt = float(t.item())
            return {"value": world * nreads * e2e_steps / dt, "unit": UNIT,
                    "h2d_bytes_per_step": int((moved1[0] - moved0[0]) // e2e_steps), "d2h_bytes_per_step": int((moved1[1] - moved0[1]) // e2e_steps),
                    "steps": e2e_steps, "ms_per_step": 1e3 * dt / e2e_steps}

        # (a) the byte form: one byte per nucleotide over PCIe (umgap_classify_reads)
        e2e_bytes = timed_e2e(lambda: capi.classify_reads(gidx, gtax, opts, nt_np, h_roff, h_goff, count_lookups=False, out=h_out))
        e2e_bytes["entry_point"] = "umgap_classify_reads: pinned host arrays, one byte per nucleotide"
        e2e_bytes["host_input_bytes_per_step"] = int(total_nt + h_roff.nbytes + h_goff.nbytes)
        # (b) the packed form (umgap_classify_reads_packed): 2 bits per nucleotide + N flags, what the CLI's block parser
        #     emits; the packing (umgap_pack_reads) is timed beside it, not inside: it is the parser's job, once per read
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
        e2e = timed_e2e(lambda: capi.classify_reads_packed(gidx, gtax, opts, codes_np, entries_np, h_roff, h_goff, count_lookups=False, out=h_out))
        e2e["entry_point"] = ("umgap_classify_reads_packed: pinned host arrays, 2-bit nucleotides + the list of 16-nucleotide words that hold an N "
                              "(a form a parser can emit directly; umgap_pack_reads makes it from bytes)")
        e2e["host_input_bytes_per_step"] = int(codes_np.nbytes + entries_np.nbytes + h_roff.nbytes + h_goff.nbytes)
        e2e["note"] = ("bytes counted by the library around its cudaMemcpyAsync calls; the offset arrays of this workload are arithmetic "
                       "progressions (reads of one length, pairs), which the library detects and regenerates on the device instead of uploading")
        e2e["pack_reads_ms_per_step_untimed"] = 1e3 * pack_s
        e2e["pack_reads_threads"] = os.cpu_count()
        e2e["byte_form"] = e2e_bytes

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (translate_lookup_kernel)
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak = float(json.load(open(peaks_path))["hbm_gbs"])
        peak_src = "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)"
    else:
        peak, peak_src = 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"
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
    if routed is not None or slices == 1:
        use_ms = avg_ms if routed is not None else max(lookup_ms / max(1, lookup_n), 1e-9)
        mode = "CUDA-event brackets around every launch of the lookup stage in the timed region, on its stream"
    else:
        # In the timed region a step is cut into slices whose lookup kernels run beside the previous slice's classify
        # kernel on two streams: their event brackets overlap each other and include the time a kernel waits for SMs the
        # other stream holds, so they do not measure the kernel.  The roofline therefore uses the same kernels on the
        # same batches timed alone (umgap_pipeline_slices(1): one lookup stage, then one classify launch per step),
        # measured in this run right after the timed region; the in-region bracket sum is reported beside it.
        use_ms, mode = alone_ms, "lookup stage timed alone on the timed region's batches (umgap_pipeline_slices(1)), CUDA events on its stream"
    achieved = alg_bytes / (use_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": "lookup_hashes_kernel (local shard)" if routed is not None
                else "lookup stage = lookup_sampled_kernel<9,TableView,3,0> (translation inside the kernel; + the long-read pass of translate_lookup_kernel), one event bracket",
                "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src, "timing": mode,
                "note": "algorithmic bytes = 248 k-mers x 32 B per read (SURVEY 8(d)); in front of seedextend -s3 the kernel probes every "
                        "third position first and the rest only for frames with a hit (bit-identical output), so the HBM actually moves "
                        "`traffic` bytes: 128-byte line fills at the random-line ceiling (profiles/README.md)",
                "dram_lines_per_read": lines,
                "algorithmic_bytes_per_launch": alg_bytes, "avg_launch_ms": use_ms,
                "classify_kernel_ms_alone": alone_classify_ms / alone_steps if routed is None else None,
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
    cb = None if args.no_cpu_baseline else cpu_baseline(args)
    reads_total = world * nreads * args.steps
    line = {
        "metric": METRIC, "value": reads_total / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8/u32 integer", "data": "synthetic",
        "lookups_per_second": reads_total * LOOKUPS_PER_READ / (ms * 1e-3),
        "config": pipeline_config(args, {"index_keys_resident": int(info.n_keys), "index_bytes": int(info.bytes),
                                         "index_build_s": build_s, "index_load_factor": info.load_factor, "flagged_sector_frac": info.n_flagged / max(1, info.n_buckets),
                                         "classified_below_root_frac": classified, "kmer_hit_rate": kmer_hit_rate}),
        "roofline": roofline, "cpu_baseline": cb, "e2e": e2e, "clocks": clocks,
        "gpu_launches": int(launches),
        "kernel_ms": {"translate_lookup": lookup_ms, "classify": classify_ms},
    }
    if routed_stages is not None:
        line["routed_stages_ms_per_step"] = routed_stages
    emit(json.dumps(line))
    if world > 1:
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
