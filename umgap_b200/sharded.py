"""Key-range-sharded index across the GPUs of one node (mode 2).

Every rank (one process per GPU) builds the shard of the table that owns its hash range.  Two ways to reach a key
another rank owns:

* ``RoutedClassifier`` (the default of ``bench.py --sharded``): the exchange step -- packed k-mer hashes travel to
  their owners and the answers back (NCCL over NVLink), every lookup runs against local HBM;
* ``attach_all``: the ranks swap the shards' CUDA IPC descriptors once and the lookup kernel loads remote sectors
  through peer mappings (no collective on the data path; request-rate bound, kept for comparison and tests).
"""
from __future__ import annotations

import ctypes as C

from . import capi


def attach_all(index: capi.Index, dist, device=None) -> None:
    """all_gather the shard descriptors over `dist` (torch.distributed, NCCL or gloo) and attach."""
    import torch
    world = dist.get_world_size()
    mine = index.shard_desc()
    raw = bytes(C.string_at(C.byref(mine), C.sizeof(mine)))
    dev = device if device is not None else ("cuda" if dist.get_backend() == "nccl" else "cpu")
    t = torch.frombuffer(bytearray(raw), dtype=torch.uint8).to(dev)
    out = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(out, t)
    descs = [capi.ShardDesc.from_buffer_copy(bytes(o.cpu().numpy().tobytes())) for o in out]
    index.attach_shards(descs)


class _Lane:
    """Buckets, counters and stream of one group range of a batch (RoutedClassifier cuts a batch into `lanes` ranges)."""

    def __init__(self, torch, G: int, cap: int, dev, slot: int):
        self.cap, self.slot = cap, slot
        self.send_h = torch.empty(G * cap, dtype=torch.int64, device=dev)
        self.recv_h = torch.empty(G * cap, dtype=torch.int64, device=dev)
        self.send_pos = torch.empty(G * cap, dtype=torch.int32, device=dev)
        self.ans = torch.empty(G * cap, dtype=torch.int32, device=dev)
        self.ans_back = torch.empty(G * cap, dtype=torch.int32, device=dev)
        self.cursors = torch.zeros(2 * G, dtype=torch.int64, device=dev)
        self.recv_counts = torch.zeros(G, dtype=torch.int64, device=dev)
        self.overflow = torch.zeros(1, dtype=torch.bool, device=dev)
        self.fills = None
        self.stream = torch.cuda.Stream(device=dev)


class RoutedClassifier:
    """Per-batch driver of the routed sharded mode.  One instance per rank; `index` is this rank's shard.

    Behind `-o | seedextend -s S` (S >= 2; capi.route_sampled_applies) the batch takes the two phases of the sampled
    lookups (pipeline.cu: lookup_sampled_kernel), each an exchange round:
      phase 1  every min(S,4)-th position of every frame record -> owners' buckets -> exchange -> local lookups ->
               exchange back -> frame masks;
      phase 2  every position of the frames with a sampled hit -> buckets -> exchange -> lookups -> exchange back ->
               ids;  then the classify kernel over the flagged frames.
    Otherwise one round with every position.  An exchange round moves only the filled part of each bucket: the fills
    are swapped first (one small all-to-all, read on the host), then grouped NCCL send/recv of the hashes (8 B per
    lookup) and, after the lookups, of the answers (4 B back).  These are the only collectives on the data path.

    A large batch is cut into `lanes` group ranges, each with its own buckets and stream; the host issues their stages
    alternately, so the transfers of one range run beside the pack / lookup / scatter kernels of the other."""

    def __init__(self, index: capi.Index, tax: capi.Taxonomy, dist, max_total_nt: int, slack: float = 1.08,
                 max_reads: int | None = None, lanes: int = 1):
        import torch
        self.index, self.tax, self.dist, self.torch = index, tax, dist, torch
        self.world, self.rank = dist.get_world_size(), dist.get_rank()
        self.max_total_nt = int(max_total_nt)
        # bucket capacity: every one of the 2*nt lookups valid and evenly spread, plus slack for skew
        self.cap = int(2 * self.max_total_nt / self.world * slack) + 4096
        dev = torch.device("cuda", torch.cuda.current_device())
        G = self.world
        self.nlanes = max(1, min(int(lanes), 4))
        # lane 0 holds full-size buckets (a batch that is not cut uses it alone), the others their share
        self.lanes = [_Lane(torch, G, self.cap if i == 0 else int(self.cap / self.nlanes * 1.05) + 4096, dev, i)
                      for i in range(self.nlanes)]
        self.ids = torch.empty(2 * self.max_total_nt + 64, dtype=torch.int32, device=dev)
        self.max_reads = int(max_reads) if max_reads is not None else self.max_total_nt // 27 + 1
        self.frame_hits = torch.zeros(self.max_reads + 8, dtype=torch.uint8, device=dev)
        self.lookups_routed = 0   # hashes this rank sent in the last batch (both phases)
        self.min_groups_per_lane = 65536
        self.profile = False      # True: CUDA-event brackets around the stages of a batch (uncut), summed in self.stage_ms
        self.stage_ms: dict = {}
        self._marks: list = []
        self._used: list = []

    def _mark(self, name: str) -> None:
        if self.profile:
            e = self.torch.cuda.Event(enable_timing=True)
            e.record()
            self._marks.append((name, e))

    def _collect(self) -> None:
        if not self.profile or not self._marks:
            return
        self.torch.cuda.synchronize()
        for (_, a), (name, b) in zip(self._marks, self._marks[1:]):
            self.stage_ms[name] = self.stage_ms.get(name, 0.0) + a.elapsed_time(b)
        self._marks = []

    # -- one exchange round of a lane, in three host steps (a generator: the caller alternates the lanes between them).
    #    This rank's own bucket is neither sent nor copied: it is looked up in place while the other buckets travel.
    def _round(self, lane: _Lane, stream):
        torch, dist = self.torch, self.dist
        G, cap, me = self.world, lane.cap, self.rank
        st = stream.cuda_stream
        with torch.cuda.stream(stream):
            lane.overflow |= lane.cursors[G:].any()
            lane.fills = lane.cursors[:G].clamp(max=cap)
            # bucket fills, and beside each this rank's overflow flag: every rank learns of an overflow anywhere in the
            # same exchange, so that all of them raise together (one rank leaving a collective alone would hang the rest)
            send = torch.stack([lane.fills, lane.overflow.to(torch.int64).expand(G)], dim=1).contiguous()
            recv = torch.empty_like(send)
            dist.all_to_all_single(recv, send)
            lane.recv_counts.copy_(recv[:, 0])
            both = torch.stack([lane.fills, recv[:, 0], recv[:, 1]]).cpu()             # the one host read of a round
            sc, rc = both[0].tolist(), both[1].tolist()
            if any(both[2].tolist()):
                # a pack kernel dropped the lookups that did not fit: the results would silently miss k-mers
                bad = [r for r, f in enumerate(both[2].tolist()) if f]
                raise RuntimeError(f"RoutedClassifier: a bucket overflowed on rank(s) {bad} (capacity {cap} lookups per owner; key skew "
                                   "beyond the slack) -- construct it with a larger `slack`")
            self.lookups_routed += sum(sc)
            self._mark("counts")
            works = self._swap(lane, lane.send_h, sc, lane.recv_h, rc)                 # hashes: 8 B per lookup
            if sc[me]:
                capi.lookup_hashes_dev(self.index, lane.send_h.data_ptr() + 8 * me * cap, lane.fills.data_ptr() + 8 * me, 1, cap,
                                       lane.ans_back.data_ptr() + 4 * me * cap, st)
            self._mark("lookup_own_bucket")
        yield
        with torch.cuda.stream(stream):
            for w in works:
                w.wait()
            self._mark("swap_hashes_exposed")
            lane.recv_counts[me] = 0
            if G > 1:
                capi.lookup_hashes_dev(self.index, lane.recv_h.data_ptr(), lane.recv_counts.data_ptr(), G, cap,
                                       lane.ans.data_ptr(), st)
            self._mark("lookup_received")
            works = self._swap(lane, lane.ans, rc, lane.ans_back, sc)                  # answers: 4 B per lookup
        yield
        with torch.cuda.stream(stream):
            for w in works:
                w.wait()
            self._mark("swap_answers")

    def _swap(self, lane: _Lane, send, send_counts, recv, recv_counts):
        """Grouped send/recv of the filled part of every other rank's bucket; returns the work handles."""
        dist = self.dist
        cap, me = lane.cap, self.rank
        ops = []
        for peer in range(self.world):
            if peer == me:
                continue
            if send_counts[peer]:
                ops.append(dist.P2POp(dist.isend, send[peer * cap: peer * cap + send_counts[peer]], peer))
            if recv_counts[peer]:
                ops.append(dist.P2POp(dist.irecv, recv[peer * cap: peer * cap + recv_counts[peer]], peer))
        return dist.batch_isend_irecv(ops) if ops else []

    def _sampled_steps(self, lane: _Lane, stream, opts, nt, read_off, group_off, nreads, total_nt, g_lo, g_hi, ranged):
        """The two exchange rounds of one group range (generator of host steps)."""
        torch = self.torch
        st = stream.cuda_stream
        for phase in (1, 2):
            with torch.cuda.stream(stream):
                capi.route_pack_sampled_dev(self.index, opts, phase, nt.data_ptr(), read_off.data_ptr(), nreads, total_nt, lane.cap,
                                            lane.send_h.data_ptr(), lane.send_pos.data_ptr(), lane.cursors.data_ptr(),
                                            self.frame_hits.data_ptr(), self.ids.data_ptr(), st,
                                            group_off.data_ptr() if ranged else 0, g_lo, g_hi, lane.slot)
                self._mark(f"pack{phase}")
            yield
            yield from self._round(lane, stream)
            with torch.cuda.stream(stream):
                if phase == 1:
                    capi.route_scatter_hits_dev(self.index, lane.ans_back.data_ptr(), lane.send_pos.data_ptr(),
                                                lane.cursors.data_ptr(), lane.cap, self.frame_hits.data_ptr(), st)
                else:
                    capi.route_scatter_dev(self.index, lane.ans_back.data_ptr(), lane.send_pos.data_ptr(),
                                           lane.cursors.data_ptr(), lane.cap, self.ids.data_ptr(), st)
                self._mark(f"scatter{phase}")
            yield

    def classify(self, opts, nt, read_off, group_off, out, total_nt: int) -> None:
        """nt (uint8), read_off / group_off (int64), out (int32): CUDA tensors of this rank's batch."""
        torch = self.torch
        if total_nt > self.max_total_nt:
            raise ValueError("batch larger than the buffers of this RoutedClassifier")
        main = torch.cuda.current_stream()
        st = main.cuda_stream
        nreads, ngroups = read_off.numel() - 1, group_off.numel() - 1
        self.lookups_routed = 0
        for lane in self.lanes:
            lane.overflow.zero_()
        if capi.route_sampled_applies(self.index, opts) and nt.data_ptr() % 16 == 0:
            if nreads > self.max_reads:
                raise ValueError("more reads than the frame-mask buffer of this RoutedClassifier holds")
            nl = self.nlanes if not self.profile and ngroups >= self.min_groups_per_lane * self.nlanes else 1
            self._used = self.lanes[:nl]
            self._mark("start")
            if nl == 1:
                steps = [self._sampled_steps(self.lanes[0], main, opts, nt, read_off, group_off, nreads, total_nt, 0, 0, False)]
            else:
                self.frame_hits[:nreads + 8].zero_()
                fork = torch.cuda.Event()
                fork.record(main)
                steps = []
                for i, lane in enumerate(self._used):
                    lane.stream.wait_event(fork)
                    steps.append(self._sampled_steps(lane, lane.stream, opts, nt, read_off, group_off, nreads, total_nt,
                                                     ngroups * i // nl, ngroups * (i + 1) // nl, True))
            while steps:   # the lanes' host steps alternate (same order on every rank: the collectives match up)
                for g in list(steps):
                    try:
                        next(g)
                    except StopIteration:
                        steps.remove(g)
            if nl > 1:
                for lane in self._used:
                    main.wait_stream(lane.stream)
            capi.classify_ids_masked_dev(self.index, self.tax, opts, self.ids.data_ptr(), read_off.data_ptr(), total_nt,
                                         group_off.data_ptr(), ngroups, self.frame_hits.data_ptr(), True, out.data_ptr(), st)
            self._mark("classify")
            self._collect()
            return
        lane = self.lanes[0]
        self._used = [lane]
        capi.route_pack_dev(self.index, opts, nt.data_ptr(), read_off.data_ptr(), nreads, total_nt, lane.cap,
                            lane.send_h.data_ptr(), lane.send_pos.data_ptr(), lane.cursors.data_ptr(), self.ids.data_ptr(), st)
        for _ in self._round(lane, main):
            pass
        capi.route_scatter_dev(self.index, lane.ans_back.data_ptr(), lane.send_pos.data_ptr(), lane.cursors.data_ptr(),
                               lane.cap, self.ids.data_ptr(), st)
        capi.classify_ids_dev(self.index, self.tax, opts, self.ids.data_ptr(), read_off.data_ptr(), total_nt,
                              group_off.data_ptr(), ngroups, out.data_ptr(), st)

    def overflowed(self) -> bool:
        """True when a bucket overflowed in the last batch (extreme key skew): rebuild with more slack."""
        return any(bool(lane.overflow.item()) for lane in self._used)
