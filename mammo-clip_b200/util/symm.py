"""Peer-mapped gather buffers for the fused all-gather + InfoNCE kernel (world_size > 1, one node, NVLink/NVSwitch).

Replaces the NCCL transport of `DistAutogradAllGatherFunction` (util/dist_autograd.py:4-26): every rank cudaMallocs two
buffer sets (double buffering by step parity, see csrc/loss.cu), exports them with CUDA IPC, and the 64-byte handles are
exchanged ONCE through torch.distributed.  Afterwards no collective call runs on the loss path: the kernel stores its
embeddings straight into every peer's buffer and signals arrival counters."""
import ctypes as C
import os

import torch
import torch.distributed as dist

from .. import _lib
from .._lib import check, lib


class SymmetricGather:
    def __init__(self, n_tensors, batch, dim, group=None):
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.K, self.B, self.D = n_tensors, batch, dim
        self.device = torch.device("cuda", torch.cuda.current_device())
        W = self.world
        self.slab_floats = W * batch * dim
        # LSE exchange (csrc/loss.cu, ABI 5): each rank scores only its own rows/columns and the 2*P*B log-sum-exps travel instead.
        # It pays once the fp32 score tiles dominate the call (W/2 times less of them) and costs one more flag round trip.
        mode = os.environ.get("MCLIP_LOSS_LSE", "auto")
        self.lse_exchange = mode == "1" or (mode == "auto" and W >= 4 and W * batch >= 1024)
        self.lse_floats = W * batch * 2 * _lib.MAX_PAIRS if self.lse_exchange else 0
        self.flags_off = n_tensors * self.slab_floats * 4 + self.lse_floats * 4
        self.set_bytes = self.flags_off + 256                             # K gather buffers (+ LSE table) + W arrival counters (padded)
        self.total_bytes = 2 * self.set_bytes
        base = C.c_void_p()
        check(lib().mclip_ipc_alloc(C.c_longlong(self.total_bytes), C.byref(base)), "mclip_ipc_alloc")
        self.base = base.value
        handle = (C.c_ubyte * 64)()
        check(lib().mclip_ipc_get_handle(C.c_void_p(self.base), handle), "mclip_ipc_get_handle")
        handles = [None] * W
        dist.all_gather_object(handles, bytes(handle), group=group)
        self.peer_base = []
        for r in range(W):
            if r == self.rank:
                self.peer_base.append(self.base)
            else:
                p = C.c_void_p()
                buf = (C.c_ubyte * 64).from_buffer_copy(handles[r])
                check(lib().mclip_ipc_open_handle(buf, C.byref(p)), "mclip_ipc_open_handle")
                self.peer_base.append(p.value)
        # device pointer tables per set: [W][K] gather buffers, [W] flag arrays
        self.tables = []
        for s in range(2):
            off = s * self.set_bytes
            gath = torch.tensor([[self.peer_base[r] + off + k * self.slab_floats * 4 for k in range(n_tensors)] for r in range(W)],
                                dtype=torch.int64, device=self.device)
            flags = torch.tensor([self.peer_base[r] + off + self.flags_off for r in range(W)], dtype=torch.int64, device=self.device)
            lse = torch.tensor([self.peer_base[r] + off + n_tensors * self.slab_floats * 4 for r in range(W)], dtype=torch.int64, device=self.device)
            self.tables.append((gath, flags, lse))
        self.calls = 0
        dist.barrier(group=group)

    def fill_args(self, args, n_tensors):
        assert n_tensors == self.K
        s = self.calls % 2
        off = s * self.set_bytes
        gath, flags, lse = self.tables[s]
        for k in range(self.K):
            args.gathered[k] = self.base + off + k * self.slab_floats * 4
        args.peer_gathered = gath.data_ptr()
        args.peer_flags = flags.data_ptr()
        args.my_flags = self.base + off + self.flags_off
        if self.lse_exchange:
            args.lse_all = self.base + off + self.K * self.slab_floats * 4
            args.peer_lse = lse.data_ptr()
        args.epoch = self.calls // 2 + 1
        self.calls += 1
