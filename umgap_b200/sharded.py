"""Key-range-sharded index across the GPUs of one node (mode 2).

Every rank (one process per GPU) builds the shard of the table that owns its hash range, the ranks
exchange the shards' CUDA IPC descriptors with one ``all_gather`` at start-up and attach them; from
then on the lookup kernel loads remote sectors straight from the owning GPU's HBM over NVLink peer
mappings.  The exchange below is control-plane only -- there is no collective on the data path.
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


class RoutedClassifier:
    """Per-batch driver of the routed sharded mode: pack + bucket -> all-to-all (hashes) -> local
    lookups -> all-to-all (answers) -> scatter -> classify.  One instance per rank; `index` is this
    rank's shard.  The two all-to-alls are the only collectives on the data path."""

    def __init__(self, index: capi.Index, tax: capi.Taxonomy, dist, max_total_nt: int, slack: float = 1.08):
        import torch
        self.index, self.tax, self.dist, self.torch = index, tax, dist, torch
        self.world, self.rank = dist.get_world_size(), dist.get_rank()
        self.max_total_nt = int(max_total_nt)
        # bucket capacity: every one of the 2*nt lookups valid and evenly spread, plus slack for skew
        self.cap = int(2 * self.max_total_nt / self.world * slack) + 4096
        dev = torch.device("cuda", torch.cuda.current_device())
        G, cap = self.world, self.cap
        self.send_h = torch.empty(G * cap, dtype=torch.int64, device=dev)
        self.recv_h = torch.empty(G * cap, dtype=torch.int64, device=dev)
        self.send_pos = torch.empty(G * cap, dtype=torch.int32, device=dev)
        self.ans = torch.empty(G * cap, dtype=torch.int32, device=dev)
        self.ans_back = torch.empty(G * cap, dtype=torch.int32, device=dev)
        self.cursors = torch.zeros(2 * G, dtype=torch.int64, device=dev)
        self.recv_counts = torch.zeros(G, dtype=torch.int64, device=dev)
        self.ids = torch.empty(2 * self.max_total_nt + 64, dtype=torch.int32, device=dev)

    def classify(self, opts, nt, read_off, group_off, out, total_nt: int) -> None:
        """nt (uint8), read_off / group_off (int64), out (int32): CUDA tensors of this rank's batch."""
        torch, dist = self.torch, self.dist
        if total_nt > self.max_total_nt:
            raise ValueError("batch larger than the buffers of this RoutedClassifier")
        G, cap = self.world, self.cap
        st = torch.cuda.current_stream().cuda_stream
        nreads, ngroups = read_off.numel() - 1, group_off.numel() - 1
        capi.route_pack_dev(self.index, opts, nt.data_ptr(), read_off.data_ptr(), nreads, total_nt, cap,
                            self.send_h.data_ptr(), self.send_pos.data_ptr(), self.cursors.data_ptr(), self.ids.data_ptr(), st)
        dist.all_to_all_single(self.recv_counts, self.cursors[:G].contiguous())      # bucket fills
        dist.all_to_all_single(self.recv_h, self.send_h)                              # hashes: G equal buckets of cap
        capi.lookup_hashes_dev(self.index, self.recv_h.data_ptr(), self.recv_counts.data_ptr(), G, cap,
                               self.ans.data_ptr(), st)
        dist.all_to_all_single(self.ans_back, self.ans)                               # answers
        capi.route_scatter_dev(self.index, self.ans_back.data_ptr(), self.send_pos.data_ptr(), self.cursors.data_ptr(),
                               cap, self.ids.data_ptr(), st)
        capi.classify_ids_dev(self.index, self.tax, opts, self.ids.data_ptr(), read_off.data_ptr(), total_nt,
                              group_off.data_ptr(), ngroups, out.data_ptr(), st)

    def overflowed(self) -> bool:
        """True when a bucket overflowed in the last batch (extreme key skew): rebuild with more slack."""
        return bool(self.cursors[self.world:].any().item())
