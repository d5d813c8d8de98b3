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


class GradBucket:
    """All dense gradients of a step in one flat buffer -> one sum-allreduce (latency-bound: ~330 KB for 602-128-41)."""

    def __init__(self, params):
        self.params = list(params)
        n = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(n, dtype=torch.float32, device=self.params[0].device)

    def all_reduce(self):
        off = 0
        for p in self.params:
            g = p.grad if p.grad is not None else torch.zeros_like(p)
            self.flat[off:off + p.numel()].copy_(g.reshape(-1))
            off += p.numel()
        if dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
        off = 0
        for p in self.params:
            if p.grad is None:
                p.grad = torch.empty_like(p)
            p.grad.copy_(self.flat[off:off + p.numel()].view_as(p))
            off += p.numel()


class ShardedTable:
    """Row-sharded fp32 feature table over the ranks of one node, read peer-to-peer inside the gather kernel."""

    def __init__(self, cuda_stream, rows_of_this_rank, n_rows_total, feature_size, pitch=None):
        from . import FeatureTable
        rank, world = dist.get_rank(), dist.get_world_size()
        pitch = pitch or feature_size
        n_local = (n_rows_total - rank + world - 1) // world
        assert rows_of_this_rank.shape == (n_local, feature_size), (rows_of_this_rank.shape, n_local)
        self._local = C.c_void_p()
        check(lib().nb_malloc_device(max(n_local, 1) * pitch * 4, C.byref(self._local)))
        check(lib().nb_memset_async(cuda_stream._h, self._local, 0, max(n_local, 1) * pitch * 4))
        src = rows_of_this_rank.contiguous()
        # strided upload into the pitched shard
        torch.cuda.current_stream().synchronize()
        cuda_stream.CUDA_DEVICE_SYNCHRONIZE()
        shard = torch.as_tensor(_Raw(self._local.value, n_local * pitch), device=cuda_stream.device).view(n_local, pitch)
        shard[:, :feature_size] = src
        torch.cuda.synchronize()
        handle = (C.c_char * 64)()
        check(lib().nb_ipc_get_handle(self._local, handle))
        handles = [None] * world
        dist.all_gather_object(handles, bytes(handle.raw))
        self._peers, ptrs = [], []
        for r in range(world):
            if r == rank:
                ptrs.append(self._local.value)
            else:
                p = C.c_void_p()
                check(lib().nb_ipc_open_handle(handles[r], C.byref(p)))
                self._peers.append(p)
                ptrs.append(p.value)
        self.table = FeatureTable(cuda_stream, ptrs, feature_size, pitch, n_rows_total, keepalive=self)
        dist.barrier()

    def gather(self, out, ids, n_rows):
        return self.table.gather(out, ids, n_rows)

    def close(self):
        dist.barrier()
        for p in self._peers:
            lib().nb_ipc_close_handle(p)
        self._peers = []
        lib().nb_free_device(self._local)


class _Raw:
    def __init__(self, address, n_floats):
        self.__cuda_array_interface__ = {"shape": (n_floats,), "typestr": "<f4", "data": (address, False), "version": 2}
