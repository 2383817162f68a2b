"""SyncBatchNorm statistic exchange over NVLink peer memory (reference: ``nn.SyncBatchNorm`` under
``train_ddp.py:270-280``): plumbing for the ``tcv_peer_allreduce_f64`` kernel (csrc/peer_reduce.cu).

torch provides the plumbing only -- a symmetric allocation that every rank of the node maps
(``torch.distributed._symmetric_memory``) and the rendezvous that exchanges the handles; the reduction itself is the
kernel, launched on the engine's compute stream.  One instance per (process group, device)."""
from __future__ import annotations

import ctypes as C
import os
import sys

import torch

from . import _cabi


class PeerReducer:
    SLOT_DOUBLES = 8192          # >= S groups x 512 channels x 2 sums of the widest BatchNorm (5 x 512 x 2 = 5120)

    def __init__(self, group, device: torch.device):
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem
        nbytes = C.c_longlong(0)
        _cabi.check(_cabi.lib().tcv_peer_buffer_bytes(self.SLOT_DOUBLES, C.byref(nbytes)), "peer_buffer_bytes")
        group = group if group is not None else dist.group.WORLD
        with torch.cuda.device(device):
            self.buf = symm_mem.empty(nbytes.value // 8, dtype=torch.float64, device=device)
            self.hdl = symm_mem.rendezvous(self.buf, group)
            self.buf.zero_()
            torch.cuda.synchronize(device)
        dist.barrier(group)      # nobody signals before every flag block is zero
        self.rank, self.world = int(self.hdl.rank), int(self.hdl.world_size)
        if self.world > 16:
            raise RuntimeError("tcv_peer_allreduce_f64 supports up to 16 ranks per node")
        self.peers_dev = int(self.hdl.buffer_ptrs_dev)
        self.epoch = 0
        self.calls = 0

    def allreduce_(self, t: torch.Tensor, stream_ptr: int) -> None:
        """In-place sum over the ranks of the group (fp64, contiguous, <= SLOT_DOUBLES values), rank order, on `stream_ptr`."""
        assert t.dtype == torch.float64 and t.is_contiguous() and t.numel() <= self.SLOT_DOUBLES
        self.epoch += 1
        self.calls += 1
        _cabi.check(_cabi.lib().tcv_peer_allreduce_f64(t.data_ptr(), t.numel(), self.peers_dev, self.rank, self.world,
                                                        self.epoch, self.SLOT_DOUBLES, stream_ptr), "peer_allreduce_f64")


def make_peer_reducer(group, device: torch.device):
    """PeerReducer, or None when peer memory is not available (then the caller keeps NCCL).  Collective: every rank of
    the group must call it, and all ranks agree on the outcome (a rank-local failure must not split the job into
    ranks that spin on flags and ranks that wait in NCCL)."""
    import torch.distributed as dist
    if os.environ.get("TCV_SYNCBN_P2P", "1") != "1":
        return None
    red, err = None, None
    try:
        red = PeerReducer(group, device)
    except Exception as e:                                     # noqa: BLE001 - reported once, NCCL stays the transport
        err = f"{type(e).__name__}: {e}"
    ok = torch.tensor([1 if red is not None else 0], dtype=torch.int32, device=device)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
    if int(ok.item()) != 1:
        if dist.get_rank(group) == 0:
            print(f"tcvom_b200: SyncBatchNorm statistics stay on NCCL (peer memory unavailable: {err})", file=sys.stderr)
        return None
    return red
