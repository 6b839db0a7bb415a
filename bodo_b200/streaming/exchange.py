"""Receive slabs of the fused partial-aggregate exchange (csrc/groupby.cu: xchg_pack_remote_kernel): symmetric memory that
every rank of a process group can store into over NVLink.

Plumbing only: torch.distributed's symmetric-memory allocator maps one buffer per rank into all peers (CUDA IPC / fabric
handles, one rendezvous per process group, cached for the life of the process) and provides the device-side barrier between
the pack and the combine kernels.  Two slabs are used alternately so that a rank that is already packing the NEXT exchange
never writes into a slab a slower peer is still combining from (a rank cannot get two exchanges ahead: the barrier of the
exchange in between needs every rank).
"""

from __future__ import annotations

import os

_CACHE: dict = {}
HDR_BYTES = 256  # XCHG_HDR_BYTES in csrc/groupby.cu


class Slabs:
    def __init__(self, group, device: int, slab_bytes: int):
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem

        self.device = device
        self.slab_bytes = slab_bytes
        self.n_pes = dist.get_world_size(group)
        gname = (group if group is not None else dist.group.WORLD).group_name
        import warnings

        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            try:
                symm_mem.enable_symm_mem_for_group(gname)  # needed by older torch builds, a deprecated no-op on newer ones
            except Exception:
                pass
        dev = torch.device("cuda", device)
        self.bufs, self.hdls = [], []
        for _ in range(2):
            t = symm_mem.empty(slab_bytes, dtype=torch.uint8, device=dev)
            h = symm_mem.rendezvous(t, gname)
            t.zero_()
            self.bufs.append(t)
            self.hdls.append(h)
        torch.cuda.synchronize(dev)
        dist.barrier(group=group)
        self.parity = 0

    def next(self):
        """(peer pointer array on the device, my slab address, handle) of the slab this exchange uses."""
        p = self.parity
        self.parity ^= 1
        h = self.hdls[p]
        return int(h.buffer_ptrs_dev), int(self.bufs[p].data_ptr()), h


def get_slabs(group, device: int):
    """Slabs of (group, device), or None when symmetric memory is unavailable (the caller then uses the NCCL exchange)."""
    if os.environ.get("B200_XCHG_FUSED", "1") == "0":
        return None
    key = (id(group) if group is not None else 0, device)
    if key not in _CACHE:
        try:
            _CACHE[key] = Slabs(group, device, int(os.environ.get("B200_XCHG_SLAB_BYTES", 256 << 20)))
        except Exception as ex:  # no symmetric memory on this build / topology
            if os.environ.get("B200_TRACE"):
                print(f"[b200 exchange] symmetric memory unavailable ({type(ex).__name__}: {ex}); using the NCCL exchange", flush=True)
            _CACHE[key] = None
    return _CACHE[key]
