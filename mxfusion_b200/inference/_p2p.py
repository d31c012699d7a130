"""The data-parallel gradient exchange as one kernel over NVLink peer memory (csrc/allreduce_p2p.cu).

`torch.distributed._symmetric_memory` is the plumbing (it allocates the bucket with the driver's virtual-memory API,
exchanges the handles between the ranks of one node and maps every peer's bucket into this process); the data path --
barrier, reduce-scatter by peer loads, all-gather by peer stores, barrier -- is `mxf_allreduce_p2p`, a plain kernel
launch, so it is captured in the step's CUDA graph (rows shard, gradients are averaged)."""
import ctypes
import os

import torch
import torch.distributed as dist

from .. import _lib
from ..common.exceptions import InferenceError


class PeerBucket(object):
    """A gradient bucket of `n` elements in symmetric memory + its in-place all-reduce."""

    def __init__(self, n, dtype, device, group=None):
        import torch.distributed._symmetric_memory as symm
        group = group if group is not None else dist.group.WORLD
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        esz = torch.empty((), dtype=dtype).element_size()
        data_bytes = ((n * esz + 15) // 16) * 16
        flag_bytes = int(_lib.lib().mxf_allreduce_p2p_flag_bytes())
        self.n, self.dtype, self.flag_off = int(n), dtype, data_bytes
        self.flag_bytes, self.esz = flag_bytes, esz
        # two flag regions: the early exchange of one segment may still be in flight when the exchange of the rest starts
        self.buf = symm.empty((data_bytes + 2 * flag_bytes) // esz, dtype=dtype, device=device)
        self.buf.zero_()
        self.handle = symm.rendezvous(self.buf, group)
        ptrs = list(self.handle.buffer_ptrs)
        if len(ptrs) != self.world or int(ptrs[self.rank]) != self.buf.data_ptr():
            raise InferenceError("symmetric memory: unexpected peer pointer table")
        self._base = [int(p) for p in ptrs]
        self._ptrs = (ctypes.c_void_p * self.world)(*self._base)
        self._range_ptrs = {}
        self.grad = self.buf[:n]
        self.err = torch.zeros((1,), dtype=torch.int32, device=device)
        self.timeout_s = float(os.environ.get('MXF_P2P_TIMEOUT_S', '20'))
        # every rank's flags are zero before any peer raises one
        torch.cuda.synchronize(device)
        dist.barrier(group)

    def all_reduce_(self, scale=1.0):
        """In place on `self.grad`: scale * sum over the ranks (identical bits on every rank).  Enqueued on the current
        stream; every rank must call it the same number of times."""
        _lib.check(_lib.lib().mxf_allreduce_p2p(_lib.dtype_code(self.buf), self._ptrs, self.rank, self.world, self.n,
                                                float(scale), self.flag_off, self.timeout_s, _lib.ptr(self.err),
                                                _lib.stream_ptr()), 'mxf_allreduce_p2p')
        return self.grad

    def all_reduce_range_(self, off, n, slot=0, scale=1.0):
        """The same on elements [off, off + n) only (off * element size a multiple of 16); `slot` selects the flag region, so
        that two exchanges of disjoint ranges may overlap in time.  Every rank must issue the same sequence of calls."""
        if (off * self.esz) % 16 != 0 or off < 0 or n < 0 or off + n > self.n or slot not in (0, 1):
            raise InferenceError("all_reduce_range_: bad range (%d, %d)" % (off, n))
        ptrs = self._range_ptrs.get(off)
        if ptrs is None:
            ptrs = (ctypes.c_void_p * self.world)(*[b + off * self.esz for b in self._base])
            self._range_ptrs[off] = ptrs
        flag_rel = self.flag_off + slot * self.flag_bytes - off * self.esz
        _lib.check(_lib.lib().mxf_allreduce_p2p(_lib.dtype_code(self.buf), ptrs, self.rank, self.world, int(n), float(scale),
                                                flag_rel, self.timeout_s, _lib.ptr(self.err), _lib.stream_ptr()),
                   'mxf_allreduce_p2p')

    def check(self):
        if int(self.err.item()) != 0:
            self.err.zero_()
            raise InferenceError("data-parallel gradient exchange: a peer rank did not reach the all-reduce within %.0f s"
                                 % self.timeout_s)


def try_peer_bucket(n, dtype, device):
    """A PeerBucket when every rank could build one (MXF_DP_P2P=0 disables), else None (the caller uses NCCL)."""
    if os.environ.get('MXF_DP_P2P', '1') == '0' or device.type != 'cuda':
        return None
    bucket, ok = None, 1
    try:
        bucket = PeerBucket(n, dtype, device)
    except Exception as e:          # no P2P between the GPUs, an older torch, ...
        ok = 0
        if os.environ.get('MXF_DP_P2P', '1') == '2':
            raise
        print("mxfusion_b200: peer-memory all-reduce unavailable on rank %d (%s: %s); using NCCL" %
              (dist.get_rank(), type(e).__name__, e), flush=True)
    flag = torch.tensor([ok], dtype=torch.int32, device=device)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    return bucket if int(flag.item()) == 1 else None
