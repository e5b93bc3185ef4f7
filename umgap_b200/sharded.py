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
