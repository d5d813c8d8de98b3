"""Data-parallel plumbing for the sampled path (one process per GPU, torch.distributed).

The path shards by mini-batch: every rank samples / gathers / aggregates its own batches with no data-path
collective. What is exchanged:
  * the dense weight gradients, summed (NOT averaged) over ranks once per step -- Parameter::reduce_multi_gpu_gradient,
    core/NtsScheduler.hpp:830-836 -> ncclAllReduce(sum) per tensor; here one bucketed all_reduce for all tensors;
  * optionally, peer mappings of a row-sharded HBM feature table (replaces the per-GPU replicated cache of
    toolkits/GS_SAMPLE_PC_MULTI.hpp:916-1015): row v lives on rank v % N at local row v // N.
"""
import ctypes as C

import numpy as np
import torch
import torch.distributed as dist

from ._capi import check, lib


def shard_seeds(ids, rank, world):
    """Contiguous split of the training ids over the GPUs (toolkits/GAT_SAMPLE_ALL_MULTI.hpp:513-527); the last rank takes the tail."""
    ids = np.asarray(ids)
    per = ids.size // world
    return ids[rank * per:(rank + 1) * per if rank < world - 1 else ids.size]


def interleave_seeds(ids, rank, world, global_batch):
    """Per-global-batch interleave (toolkits/GS_SAMPLE_PC_MULTI.hpp:1033-1131): global batch b is cut into `world` equal local
    batches of global_batch // world seeds; rank r takes slice r of every global batch."""
    ids = np.asarray(ids)
    local = global_batch // world
    out = []
    for start in range(0, ids.size, global_batch):
        chunk = ids[start:start + global_batch]
        out.append(chunk[rank * local:(rank + 1) * local])
    return np.concatenate(out) if out else ids[:0]


class PeerAllReduce:
    """In-place SUM all-reduce of a small fp32 CUDA tensor over the ranks of one node by this library's own kernels over NVLink
    peer memory (csrc/peer.cu) -- no NCCL call on the step's critical path. Each rank's block (arrival flags + two slots of
    `world` regions) is a cuMemCreate allocation shared by POSIX descriptor, like the sharded table's shards. Push based:
    begin() writes this rank's buffer into every rank's slot and signals, end() waits for every rank's flags and sums the local
    slot in rank order (bit-identical on every rank); all_reduce() does both in one launch. Enqueued on `cuda_stream`'s stream."""

    def __init__(self, cuda_stream, max_floats):
        import os
        self.cs, self.max_floats = cuda_stream, int(max_floats)
        rank, world = dist.get_rank(), dist.get_world_size()
        nbytes = int(lib().nb_peer_comm_block_bytes(self.max_floats, world))
        self._local, fd = C.c_void_p(), C.c_int(-1)
        check(lib().nb_vmm_alloc(cuda_stream._h, nbytes, C.byref(self._local), C.byref(fd)))
        peer_fds = exchange_fds(fd.value)
        self._peers, blocks = [], []
        for r in range(world):
            if r == rank:
                blocks.append(self._local.value)
            else:
                p = C.c_void_p()
                check(lib().nb_vmm_import(cuda_stream._h, peer_fds[r], nbytes, C.byref(p)))
                os.close(peer_fds[r])
                self._peers.append(p)
                blocks.append(p.value)
        arr = (C.c_void_p * world)(*blocks)
        self._h = C.c_void_p()
        check(lib().nb_peer_comm_create(cuda_stream._h, rank, world, self.max_floats, arr, C.byref(self._h)))   # zeroes this rank's flags
        dist.barrier()                     # every rank's flags are zero before anyone signals

    def all_reduce(self, t):
        assert t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() and t.numel() <= self.max_floats
        check(lib().nb_peer_allreduce_sum(self._h, t.data_ptr(), t.numel()))
        return t

    def begin(self, t):
        """push phase: nothing waits; enqueue the work that should hide the rank skew, then end()"""
        assert t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() and t.numel() <= self.max_floats
        check(lib().nb_peer_allreduce_begin(self._h, t.data_ptr(), t.numel()))

    def end(self, t):
        check(lib().nb_peer_allreduce_end(self._h, t.data_ptr(), t.numel()))
        return t

    def stats(self, reset=True):
        """(exchanges ended, mean us, max us) the reduce phases spent waiting for the slowest rank since the last reset"""
        n, s, m = C.c_uint64(), C.c_uint64(), C.c_uint64()
        check(lib().nb_peer_comm_stats(self._h, C.byref(n), C.byref(s), C.byref(m), 1 if reset else 0))
        return n.value, (s.value / n.value / 1e3 if n.value else 0.0), m.value / 1e3

    def timed_out(self):
        e = C.c_int(0)
        check(lib().nb_peer_comm_check(self._h, C.byref(e)))
        return bool(e.value)

    def close(self):
        torch.cuda.synchronize()
        dist.barrier()
        lib().nb_peer_comm_destroy(self._h)
        for p in self._peers:
            check(lib().nb_vmm_free(p))
        self._peers = []
        dist.barrier()
        if self._local:
            check(lib().nb_vmm_free(self._local))
            self._local = None


class GradBucket:
    """All dense gradients of a step in one flat buffer -> one sum-allreduce (latency-bound: ~330 KB for 602-128-41):
    `peer` (a PeerAllReduce) does it as one kernel over NVLink peer memory, otherwise one NCCL / gloo all_reduce."""

    def __init__(self, params, peer=None):
        self.params = list(params)
        self.peer = peer
        n = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(n, dtype=torch.float32, device=self.params[0].device)

    def all_reduce(self):
        off = 0
        for p in self.params:
            g = p.grad if p.grad is not None else torch.zeros_like(p)
            self.flat[off:off + p.numel()].copy_(g.reshape(-1))
            off += p.numel()
        if self.peer is not None:
            # the pack / unpack copies run on torch's current stream, the exchange on the peer's own: order them both ways
            cur = torch.cuda.current_stream(self.flat.device)
            peer_stream = torch.cuda.ExternalStream(self.peer.cs.stream or 0, device=self.flat.device)
            if peer_stream.cuda_stream != cur.cuda_stream:
                peer_stream.wait_stream(cur)
            self.peer.all_reduce(self.flat)
            if peer_stream.cuda_stream != cur.cuda_stream:
                cur.wait_stream(peer_stream)
        elif dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
        off = 0
        for p in self.params:
            if p.grad is None:
                p.grad = torch.empty_like(p)
            p.grad.copy_(self.flat[off:off + p.numel()].view_as(p))
            off += p.numel()


def exchange_fds(fd):
    """Every rank hands `fd` to every other rank of the node; returns {rank: local duplicate of that rank's descriptor}.
    Descriptors cross processes as SCM_RIGHTS ancillary data on unix stream sockets (abstract names, nothing left on disk).
    A connect() completes once it is queued on the peer's listen backlog, so connect-all, then accept-and-send, then receive
    cannot deadlock."""
    import os
    import socket
    import uuid
    rank, world = dist.get_rank(), dist.get_world_size()
    name = f"\0nts_b200_{os.getpid()}_{uuid.uuid4().hex}"
    listener = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
    listener.bind(name)
    listener.listen(world)
    names = [None] * world
    dist.all_gather_object(names, name)      # also orders every listen() before any connect()
    links = {}
    for r in range(world):
        if r != rank:
            c = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
            c.connect(names[r])
            links[r] = c
    served = []
    for _ in range(world - 1):
        conn, _addr = listener.accept()
        socket.send_fds(conn, [b"f"], [fd])
        served.append(conn)
    got = {}
    for r, c in links.items():
        _msg, fds, _flags, _a = socket.recv_fds(c, 1, 1)
        assert len(fds) == 1, f"rank {rank}: no descriptor from rank {r}"
        got[r] = fds[0]
        c.close()
    for conn in served:
        conn.close()
    listener.close()
    return got


class ShardedTable:
    """Row-sharded fp32 feature table over the ranks of one node, read peer-to-peer inside the gather kernel. Shards live in
    cuMemCreate allocations shared through POSIX descriptors (nb_vmm_*): cudaIpc mappings of multi-GB shards are TLB-miss bound
    for random rows (~45 GB/s per peer measured), these are not (link rate)."""

    def __init__(self, cuda_stream, rows_of_this_rank, n_rows_total, feature_size, pitch=None):
        import os
        from . import FeatureTable
        rank, world = dist.get_rank(), dist.get_world_size()
        pitch = pitch or feature_size
        rows_of = lambda r: (n_rows_total - r + world - 1) // world
        n_local = rows_of(rank)
        assert rows_of_this_rank.shape == (n_local, feature_size), (rows_of_this_rank.shape, n_local)
        self._local, fd = C.c_void_p(), C.c_int(-1)
        check(lib().nb_vmm_alloc(cuda_stream._h, n_local * pitch * 4, C.byref(self._local), C.byref(fd)))
        torch.cuda.current_stream().synchronize()
        cuda_stream.CUDA_DEVICE_SYNCHRONIZE()
        if n_local:
            # strided upload into the pitched shard
            shard = torch.as_tensor(_Raw(self._local.value, n_local * pitch), device=cuda_stream.device).view(n_local, pitch)
            if pitch != feature_size:
                shard.zero_()
            shard[:, :feature_size] = rows_of_this_rank
        torch.cuda.synchronize()
        peer_fds = exchange_fds(fd.value)
        self._peers, ptrs = [], []
        for r in range(world):
            if r == rank:
                ptrs.append(self._local.value)
            else:
                p = C.c_void_p()
                check(lib().nb_vmm_import(cuda_stream._h, peer_fds[r], rows_of(r) * pitch * 4, C.byref(p)))
                os.close(peer_fds[r])     # the mapping keeps the allocation alive
                self._peers.append(p)
                ptrs.append(p.value)
        self.table = FeatureTable(cuda_stream, ptrs, feature_size, pitch, n_rows_total, keepalive=self)
        dist.barrier()

    def gather(self, out, ids, n_rows):
        return self.table.gather(out, ids, n_rows)

    def close(self):
        torch.cuda.synchronize()
        dist.barrier()
        for p in self._peers:
            check(lib().nb_vmm_free(p))
        self._peers = []
        dist.barrier()
        if self._local:
            check(lib().nb_vmm_free(self._local))
            self._local = None


class _Raw:
    def __init__(self, address, n_floats):
        self.__cuda_array_interface__ = {"shape": (n_floats,), "typestr": "<f4", "data": (address, False), "version": 2}
