"""Transport of the multi-GPU path: torch.distributed behind the cb_exchange callbacks.

One process per GPU (torchrun).  The library partitions records on the device and calls back for
the transfers (include/cloudbrush_gpu.h, cb_exchange); this module implements those callbacks:

  backend "nccl"  -- device buffers are wrapped zero-copy as torch tensors
                     (__cuda_array_interface__) and moved with all_to_all_single / broadcast over
                     NVLink/NVSwitch;
  backend "gloo"  -- the same calls staged through host memory (CPU tests, and 2 ranks sharing one
                     GPU in the GPU test-suite, where NCCL refuses duplicate devices).

The tensor-level methods (exchange_counts / alltoallv / allgatherv / allreduce) are what the CPU
tests exercise with world_size 2 on gloo.
"""
import ctypes as C
import sys
import traceback

import torch
import torch.distributed as dist

CB_COUNTS = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.c_int32)
CB_A2A = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.POINTER(C.c_uint64), C.c_void_p, C.POINTER(C.c_uint64))
CB_AGV = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.POINTER(C.c_uint64))
CB_ARED = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(C.c_uint64), C.c_int32, C.c_int32)
CB_AMAX8 = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_uint64)


class cb_exchange(C.Structure):
    _fields_ = [("user", C.c_void_p), ("exchange_counts", CB_COUNTS), ("alltoallv", CB_A2A),
                ("allgatherv", CB_AGV), ("allreduce_u64", CB_ARED), ("allreduce_max_u8_dev", CB_AMAX8)]


class _DevMem:
    """Zero-copy view of raw device memory for torch.as_tensor."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False),
                                         "version": 2}


def _dev_tensor(ptr, nbytes, device):
    if nbytes == 0 or not ptr:
        return torch.empty(0, dtype=torch.uint8, device=device)
    return torch.as_tensor(_DevMem(ptr, nbytes), device=device)


class TorchExchange:
    def __init__(self, group=None, device=None):
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed is not initialised")
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.backend = dist.get_backend(group)
        self.device = device  # torch.device of the GPU this rank drives (None on CPU-only runs)
        self.on_device = self.backend == "nccl"
        self.bytes_moved = 0
        self._struct = None

    # ---------------------------------------------------------------- tensor-level API
    def exchange_counts(self, send):
        """send: int64 tensor of world*n values (n per destination) -> what every rank sent here."""
        send = torch.as_tensor(send, dtype=torch.int64)
        n = send.numel() // self.world
        if self.on_device:
            t_in = send.to(self.device)
            t_out = torch.empty_like(t_in)
            dist.all_to_all_single(t_out, t_in, group=self.group)
            return t_out.cpu()
        gathered = [torch.empty_like(send) for _ in range(self.world)]
        dist.all_gather(gathered, send.contiguous(), group=self.group)
        return torch.cat([g[self.rank * n:(self.rank + 1) * n] for g in gathered])

    def alltoallv(self, send, send_bytes, recv, recv_bytes):
        """send/recv: uint8 tensors holding world consecutive blocks of the given byte sizes."""
        send_bytes = [int(x) for x in send_bytes]
        recv_bytes = [int(x) for x in recv_bytes]
        self.bytes_moved += sum(send_bytes) - send_bytes[self.rank]
        if self.on_device:
            dist.all_to_all_single(recv, send, output_split_sizes=recv_bytes, input_split_sizes=send_bytes,
                                   group=self.group)
            return
        # gloo has no all_to_all: pairwise isend / irecv on host tensors
        h_send = send.cpu() if send.is_cuda else send
        h_recv = torch.empty(sum(recv_bytes), dtype=torch.uint8)
        so = [0]
        ro = [0]
        for b in send_bytes:
            so.append(so[-1] + b)
        for b in recv_bytes:
            ro.append(ro[-1] + b)
        reqs = []
        for r in range(self.world):
            if r == self.rank:
                h_recv[ro[r]:ro[r + 1]] = h_send[so[r]:so[r + 1]]
                continue
            if recv_bytes[r]:
                reqs.append(dist.irecv(h_recv[ro[r]:ro[r + 1]], src=self._global(r), group=self.group))
            if send_bytes[r]:
                reqs.append(dist.isend(h_send[so[r]:so[r + 1]].contiguous(), dst=self._global(r), group=self.group))
        for q in reqs:
            q.wait()
        recv.copy_(h_recv)

    def allgatherv(self, send, recv, recv_bytes):
        """recv receives the world blocks (recv_bytes[r] bytes from rank r) in rank order."""
        recv_bytes = [int(x) for x in recv_bytes]
        self.bytes_moved += sum(recv_bytes) - recv_bytes[self.rank]
        off = 0
        stage = None if self.on_device else torch.empty(sum(recv_bytes), dtype=torch.uint8)
        for r in range(self.world):
            view = (recv if self.on_device else stage)[off:off + recv_bytes[r]]
            if r == self.rank:
                view.copy_(send[:recv_bytes[r]])
            if recv_bytes[r]:
                dist.broadcast(view, src=self._global(r), group=self.group)
            off += recv_bytes[r]
        if not self.on_device:
            recv.copy_(stage)

    def allreduce(self, vals, op=0):
        t = torch.as_tensor(vals, dtype=torch.int64).clone()
        if self.on_device:
            t = t.to(self.device)
        dist.all_reduce(t, op=dist.ReduceOp.SUM if op == 0 else dist.ReduceOp.MAX, group=self.group)
        return t.cpu()

    def allreduce_max_u8(self, t):
        """in-place element-wise max of a uint8 tensor (device tensor on nccl, any tensor on gloo)."""
        if self.on_device:
            dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
            return t
        h = t.cpu() if t.is_cuda else t
        dist.all_reduce(h, op=dist.ReduceOp.MAX, group=self.group)
        if h is not t:
            t.copy_(h)
        return t

    def _global(self, r):
        return dist.get_global_rank(self.group, r) if self.group is not None else r

    # ---------------------------------------------------------------- C callbacks
    def _guard(self, fn):
        def wrapped(*a):
            try:
                fn(*a)
                if self.device is not None and self.device.type == "cuda":
                    torch.cuda.current_stream(self.device).synchronize()
                return 0
            except Exception:  # never let an exception cross the C boundary
                traceback.print_exc(file=sys.stderr)
                return 1
        return wrapped

    def as_struct(self):
        """The cb_exchange structure for cb_set_exchange (kept alive by this object)."""
        if self._struct is not None:
            return self._struct
        W = self.world

        def counts(user, send, recv, n):
            out = self.exchange_counts([int(send[i]) for i in range(W * n)])
            for i in range(W * n):
                recv[i] = int(out[i])

        def a2a(user, send, sb, recv, rb):
            sbl = [int(sb[i]) for i in range(W)]
            rbl = [int(rb[i]) for i in range(W)]
            self.alltoallv(_dev_tensor(send, sum(sbl), self.device), sbl, _dev_tensor(recv, sum(rbl), self.device), rbl)

        def agv(user, send, nbytes, recv, rb):
            rbl = [int(rb[i]) for i in range(W)]
            self.allgatherv(_dev_tensor(send, int(nbytes), self.device), _dev_tensor(recv, sum(rbl), self.device), rbl)

        def ared(user, vals, n, op):
            out = self.allreduce([int(vals[i]) for i in range(n)], op)
            for i in range(n):
                vals[i] = int(out[i])

        def amax8(user, dev, n):
            self.allreduce_max_u8(_dev_tensor(dev, int(n), self.device))

        self._cbs = (CB_COUNTS(self._guard(counts)), CB_A2A(self._guard(a2a)), CB_AGV(self._guard(agv)),
                     CB_ARED(self._guard(ared)), CB_AMAX8(self._guard(amax8)))
        self._struct = cb_exchange(None, *self._cbs)
        return self._struct


def shard_lines(sfa: bytes, rank: int, world: int) -> bytes:
    """Consecutive, line-aligned shards of an SFA buffer (what every rank loads)."""
    lines = sfa.split(b"\n")
    if lines and lines[-1] == b"":
        lines.pop()
    per = (len(lines) + world - 1) // world
    part = lines[rank * per:(rank + 1) * per]
    return b"".join(l + b"\n" for l in part)
