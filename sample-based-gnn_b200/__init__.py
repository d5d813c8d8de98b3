"""sample-based-gnn_b200 -- host-side mirror of the reference's operator surface over libnts_b200.so.

The names, argument meaning and layer/shape conventions follow the reference so that callers
and tests read like the reference's own code:

  Cuda_Stream                      cuda/ntsCUDA.hpp:177-595 (the methods on the sampled hot path)
  FullyRepGraph                    core/FullyRepGraph.hpp:682-799
  sampCSC / SampledSubgraph        core/coocsc.hpp:24-462 / core/FullyRepGraph.hpp:30-681 (device members)
  FastSampler                      core/ntsFastSampler.hpp (GPU ctor :125-176, sample_gpu_fast[_omit] :648-915,
                                   load_feature_gpu[_cache] :227-317, load_label_gpu :400-426, load_share_embedding* :472-529)
  SingleGPUAllSampleGraphOp, SingleGPUSampleGraphOp   core/ntsSingleGPUSampleGraphOp.hpp:50-294
  BatchGPUSrcDstScatterOp / BatchGPUEdgeSoftMax / BatchGPUAggregateDst   core/ntsPushdownGraphOp.hpp:490-747
  GATFusedOp                       the five-op chain of toolkits/GAT_SAMPLE_ALL_MULTI.hpp:383-464 as one op

torch is used for device memory, streams and autograd glue only; every operator body is a
hand-written sm_100a kernel behind the C ABI (include/nts_b200.h). There is no CPU path.
"""
import ctypes as C

import numpy as np
import torch

from . import _capi
from ._capi import (NB_SAMPLER_BUILD_CSR, NB_SAMPLER_MERGE_SRC_DST, NB_SAMPLER_NO_BOTTOM_CSR, NB_SAMPLER_UP_DEGREE, NB_WEIGHT_MEAN,
                    NB_WEIGHT_MEAN_SAMPLED, NB_WEIGHT_NONE, NB_WEIGHT_SUM, LayerView, NtsError, check, lib, ptr)

__all__ = ["ColdStage", "LazyFeature", "preSample", "write_pre_sample_file", "read_pre_sample_file", "set_cache_index", "Cuda_Stream", "FullyRepGraph", "FastSampler", "SampledSubgraph", "sampCSC", "WeightType",
           "SingleGPUAllSampleGraphOp", "SingleGPUSampleGraphOp", "GATFusedOp", "BatchGPUSrcDstScatterOp",
           "BatchGPUEdgeSoftMax", "BatchGPUAggregateDst", "FeatureTable", "NtsError"]


class WeightType:  # core/ntsFastSampler.hpp:27
    Sum, Mean, None_, MeanSampled = NB_WEIGHT_SUM, NB_WEIGHT_MEAN, NB_WEIGHT_NONE, NB_WEIGHT_MEAN_SAMPLED


def _pitch(t, F):
    """row pitch in floats of a [rows, F] tensor (a [:, :F] view of a wider, row-padded allocation is allowed)"""
    if hasattr(t, "stride") and t.dim() == 2:
        assert t.stride(1) == 1, "rows must be contiguous"
        return int(t.stride(0)) if t.shape[0] > 1 else max(int(t.stride(0)), F)
    return F


def _alloc_like_rows(n, F, like):
    """[n, F] output with the same row pitch as `like` (padded in -> padded out)"""
    p = _pitch(like, F)
    if p == F:
        return torch.empty((n, F), dtype=torch.float32, device=like.device)
    return torch.empty((n, p), dtype=torch.float32, device=like.device)[:, :F]


class _DevArray:
    """zero-copy view of sampler-owned device memory for torch.as_tensor (CUDA array interface)."""

    def __init__(self, address, n, typestr, owner):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (address, False), "version": 2}
        self._owner = owner


def _view(address, n, kind, device, owner):
    if not address or n == 0:
        return torch.empty(0, dtype=torch.int32 if kind == "u" else torch.float32, device=device)
    return torch.as_tensor(_DevArray(address, n, "<i4" if kind == "u" else "<f4", owner), device=device)


class Cuda_Stream:
    """cuda/ntsCUDA.hpp:177-595. One per (device, pipeline slot); all work is ordered on `stream`."""

    def __init__(self, device=0, stream=None, adopt=None):
        """stream=None: own non-blocking stream (the reference's default ctor). A torch.cuda.Stream or a raw
        cudaStream_t handle is adopted (handle 0 = the legacy default stream needs adopt=True)."""
        h = C.c_void_p()
        if hasattr(stream, "cuda_stream"):
            stream, adopt = stream.cuda_stream, True
        if adopt is None:
            adopt = stream is not None
        check(lib().nb_ctx_create(int(device), ptr(stream) or None, 1 if adopt else 0, C.byref(h)))
        self._h = h
        self.device = torch.device("cuda", int(device))

    @classmethod
    def on_torch_stream(cls, device=0):
        """ordered with torch's current stream, as the toolkits do (setCurrentCUDAStream + setNewStream)"""
        return cls(device, torch.cuda.current_stream(device))

    def __del__(self):
        try:
            lib().nb_ctx_destroy(self._h)
        except Exception:
            pass

    # -- stream management
    def getStream(self):
        return lib().nb_ctx_stream(self._h)

    @property
    def stream(self):
        return self.getStream()

    def setNewStream(self, cuda_stream):
        check(lib().nb_ctx_set_stream(self._h, ptr(cuda_stream)))

    def CUDA_DEVICE_SYNCHRONIZE(self):
        check(lib().nb_ctx_synchronize(self._h))

    def launch_count(self):
        return int(lib().nb_ctx_launch_count(self._h))

    # -- aggregation (cuda/ntsCUDA.hpp:223-288). The unused src/dst range arguments of the reference are kept.
    def Gather_By_Dst_From_Src_Spmm(self, input, output, weight_forward, row_indices, column_offset, column_num,
                                    src_start=0, src_end=0, dst_start=0, dst_end=0, edges=0, batch_size=0,
                                    feature_size=0, with_weight=False, tensor_weight=False):
        check(lib().nb_aggregate_csc_fwd(self._h, ptr(input), ptr(output), ptr(weight_forward) if with_weight else None,
                                         ptr(row_indices), ptr(column_offset), batch_size, column_num, feature_size))

    Gather_By_Dst_From_Src = lambda self, input, output, weight_forward, row_indices, column_offset, src_start=0, src_end=0, dst_start=0, dst_end=0, edges=0, batch_size=0, feature_size=0, with_weight=False, tensor_weight=False: \
        self.Gather_By_Dst_From_Src_Spmm(input, output, weight_forward, row_indices, column_offset, 0, 0, 0, 0, 0, edges, batch_size, feature_size, with_weight, tensor_weight)
    Gather_By_Dst_From_Src_Optim = Gather_By_Dst_From_Src

    def Gather_By_Src_From_Dst_Spmm(self, input, output, weight_backward, row_offset, column_indices, column_num,
                                    src_start=0, src_end=0, dst_start=0, dst_end=0, edges=0, batch_size=0,
                                    feature_size=0, with_weight=False, tensor_weight=False):
        check(lib().nb_aggregate_csr_bwd(self._h, ptr(input), ptr(output), ptr(weight_backward) if with_weight else None,
                                         ptr(row_offset), ptr(column_indices), batch_size, column_num, feature_size))

    Gather_By_Src_From_Dst = lambda self, input, output, weight_backward, row_offset, column_indices, src_start=0, src_end=0, dst_start=0, dst_end=0, edges=0, batch_size=0, feature_size=0, with_weight=False, tensor_weight=False: \
        self.Gather_By_Src_From_Dst_Spmm(input, output, weight_backward, row_offset, column_indices, 0, 0, 0, 0, 0, edges, batch_size, feature_size, with_weight, tensor_weight)
    Gather_By_Src_From_Dst_Optim = Gather_By_Src_From_Dst

    def aggregate_fwd_pitched(self, input, output, weight, row_indices, column_offset, n_dst, feature_size, in_pitch, out_pitch,
                              n_dst_dev=None):
        """CSC forward on row-padded tensors (pitch in floats); n_dst_dev: device address of the row count, or None."""
        check(lib().nb_aggregate_csc_fwd_dyn(self._h, ptr(input), ptr(output), ptr(weight), ptr(row_indices), ptr(column_offset),
                                             ptr(n_dst_dev), n_dst, feature_size, in_pitch, out_pitch))

    def aggregate_gathered_fwd(self, table, gather_index, output, weight, column_offset, n_dst, feature_size, table_pitch, out_pitch,
                               n_dst_dev=None):
        """load_feature_gpu + the bottom hop's forward in one kernel: output[d] = sum_e w[e] * table[gather_index[e] & 0x7fffffff]
        (bit 31 of a gather_index entry = the L2 hint; sampCSC.dev_gather_index)"""
        check(lib().nb_aggregate_gathered_fwd_dyn(self._h, ptr(table), table_pitch, ptr(gather_index), ptr(output), ptr(weight),
                                                  ptr(column_offset), ptr(n_dst_dev), n_dst, feature_size, out_pitch))

    def aggregate_bwd_pitched(self, input, output, weight_b, row_offset, column_indices, n_src, feature_size, in_pitch, out_pitch,
                              n_src_dev=None):
        check(lib().nb_aggregate_csr_bwd_dyn(self._h, ptr(input), ptr(output), ptr(weight_b), ptr(row_offset), ptr(column_indices),
                                             ptr(n_src_dev), n_src, feature_size, in_pitch, out_pitch))

    def Push_From_Dst_To_Src_Spmm(self, input, output, weight, row_indices, column_offset, column_num, src_start=0,
                                  src_end=0, dst_start=0, dst_end=0, edges=0, batch_size=0, feature_size=0,
                                  with_weight=False, tensor_weight=False):
        check(lib().nb_aggregate_push_bwd(self._h, ptr(input), ptr(output), ptr(weight) if with_weight else None,
                                          ptr(row_indices), ptr(column_offset), batch_size, column_num, feature_size))

    # -- gathers (cuda/ntsCUDA.hpp:370-387)
    def zero_copy_feature_move_gpu(self, dev_feature, pinned_host_feature, src_vertex, feature_size, vertex_size,
                                   table_pitch=None, out_pitch=None):
        check(lib().nb_gather_rows(self._h, ptr(dev_feature), ptr(pinned_host_feature), ptr(src_vertex), vertex_size,
                                   feature_size, table_pitch or feature_size, out_pitch or feature_size))

    def gather_feature_cached(self, dev_feature, cold_feature, dev_cache_feature, cache_node_hashmap, src_vertex,
                              feature_size, vertex_size, hit_count=None):
        """zero_copy_feature_move_gpu_cache + gather_feature_from_gpu_cache in one kernel (:378-383)."""
        check(lib().nb_gather_rows_cached(self._h, ptr(dev_feature), ptr(cold_feature), feature_size,
                                          ptr(dev_cache_feature), feature_size, ptr(cache_node_hashmap), ptr(src_vertex),
                                          vertex_size, feature_size, feature_size, ptr(hit_count)))

    def zero_copy_feature_move_gpu_cache(self, dev_feature, host_pinned_feature, src_vertex, feature_size, vertex_size, local_idx):
        """:378-380 -- row local_idx[i] of dev_feature <- table row src_vertex[local_idx[i]] (the cold list of
        FastSampler::load_feature_gpu_cache); synchronises like the reference (cuda/ntsCUDAGraphOP.cu:1744-1756)."""
        check(lib().nb_gather_rows_indexed(self._h, ptr(dev_feature), _pitch(dev_feature, feature_size), ptr(host_pinned_feature),
                                           _pitch(host_pinned_feature, feature_size), ptr(src_vertex), ptr(local_idx), None,
                                           vertex_size, feature_size))
        self.CUDA_DEVICE_SYNCHRONIZE()

    def gather_feature_from_gpu_cache(self, dev_feature, dev_cache_feature, src_vertex, feature_size, vertex_size, local_idx,
                                      cache_node_hashmap):
        """:381-383 -- row local_idx[i] of dev_feature <- cache row cache_node_hashmap[src_vertex[local_idx[i]]] (the hot list)."""
        check(lib().nb_gather_rows_indexed(self._h, ptr(dev_feature), _pitch(dev_feature, feature_size), ptr(dev_cache_feature),
                                           _pitch(dev_cache_feature, feature_size), ptr(src_vertex), ptr(local_idx),
                                           ptr(cache_node_hashmap), vertex_size, feature_size))
        self.CUDA_DEVICE_SYNCHRONIZE()

    def global_copy_label_move_gpu(self, dev_label, global_dev_label, dst_vertex, vertex_size):
        check(lib().nb_gather_labels(self._h, ptr(dev_label), ptr(global_dev_label), ptr(dst_vertex), vertex_size))

    # -- hot-row override (cuda/ntsCUDA.hpp:495-497, 521-537)
    def dev_load_share_embedding(self, dev_embedding, share_embedding, dev_cacheflag, dev_cachelocation, embedding_size,
                                 destination_vertex, vertex_size, super_batch_id):
        check(lib().nb_row_override(self._h, ptr(dev_embedding), ptr(share_embedding), ptr(dev_cacheflag),
                                    ptr(dev_cachelocation), ptr(destination_vertex), vertex_size, embedding_size,
                                    super_batch_id))

    def dev_load_share_aggregate(self, dev_feature, share_feature, dev_cacheflag, dev_cachemap, feature_size,
                                 destination_vertex, vertex_size):
        check(lib().nb_row_override(self._h, ptr(dev_feature), ptr(share_feature), ptr(dev_cacheflag), ptr(dev_cachemap),
                                    ptr(destination_vertex), vertex_size, feature_size, 0xFFFFFFFF))

    def dev_load_share_embedding_and_feature(self, dev_feature, dev_embedding, share_feature, share_embedding,
                                             dev_cacheflag, dev_cachelocation, feature_size, embedding_size,
                                             destination_vertex, vertex_size, super_batch_id=0xFFFFFFFF):
        check(lib().nb_row_override2(self._h, ptr(dev_feature), ptr(dev_embedding), ptr(share_feature),
                                     ptr(share_embedding), ptr(dev_cacheflag), ptr(dev_cachelocation),
                                     ptr(destination_vertex), vertex_size, feature_size, embedding_size, super_batch_id))

    # -- GAT edge ops (cuda/ntsCUDA.hpp:308-329, 567-573)
    def Scatter_Src_Dst_to_Msg(self, message, src_mirror_feature, row_indices, column_offset, batch_size, feature_size,
                               dst_to_local):
        check(lib().nb_scatter_src_dst_to_msg(self._h, ptr(message), ptr(src_mirror_feature), ptr(row_indices),
                                              ptr(column_offset), batch_size, feature_size, ptr(dst_to_local)))

    def Gather_Msg_To_Src_Dst(self, src_mirror_feature, message, row_indices, column_offset, batch_size, feature_size,
                              dst_to_local, src_size=None):
        n_src = src_mirror_feature.shape[0] if src_size is None else src_size
        check(lib().nb_gather_msg_to_src_dst(self._h, ptr(src_mirror_feature), ptr(message), ptr(row_indices),
                                             ptr(column_offset), batch_size, n_src, feature_size, ptr(dst_to_local)))

    def Edge_Softmax_Forward_Norm_Block(self, msg_output, msg_input, msg_cached, row_indices, column_offset, batch_size,
                                        feature_size):
        assert feature_size == 1, "the reference's GAT has one head (SURVEY section 8 a14)"
        check(lib().nb_edge_softmax_fwd(self._h, ptr(msg_output), ptr(msg_input), ptr(msg_cached), ptr(column_offset),
                                        batch_size))

    def Edge_Softmax_Backward_Block(self, msg_input_grad, msg_output_grad, msg_cached, row_indices, column_offset,
                                    batch_size, feature_size):
        assert feature_size == 1
        check(lib().nb_edge_softmax_bwd(self._h, ptr(msg_input_grad), ptr(msg_output_grad), ptr(msg_cached),
                                        ptr(column_offset), batch_size))

    def Gather_Msg_to_Dst(self, dst_feature, message, row_indices, column_offset, batch_size, feature_size):
        check(lib().nb_gather_msg_to_dst(self._h, ptr(dst_feature), ptr(message), ptr(column_offset), batch_size,
                                         feature_size))

    def Scatter_Dst_to_Msg(self, message, dst_feature, row_indices, column_offset, batch_size, feature_size):
        check(lib().nb_scatter_dst_to_msg(self._h, ptr(message), ptr(dst_feature), ptr(column_offset), batch_size,
                                          feature_size))


class FullyRepGraph:
    """core/FullyRepGraph.hpp:682-799: the global in-edge CSC, resident in HBM (nb_graph). From an edge list it is built on
    the device (`GenerateAll`'s two-pass counting sort -- column = dst, entries in file order -- as a stable radix sort by
    dst, csrc/ingest.cu); `build_on_host=True` keeps the host restatement (stable argsort), which the tests use as the
    second opinion."""

    def __init__(self, cuda_stream, global_vertices, edge_pairs=None, column_offset=None, row_indices=None,
                 in_degree=None, out_degree=None, build_on_host=False):
        self.cs = cuda_stream
        self.global_vertices = int(global_vertices)
        if edge_pairs is not None and not build_on_host:
            # the CSC is built on the device (stable radix sort by dst, csrc/ingest.cu); edge_pairs: numpy [E,2] / flat, or a
            # CUDA int32 tensor of the same layout
            on_dev = hasattr(edge_pairs, "is_cuda") and edge_pairs.is_cuda
            pairs = edge_pairs.contiguous() if on_dev else np.ascontiguousarray(edge_pairs, dtype=np.uint32).reshape(-1, 2)
            n_edges = int(pairs.numel() // 2) if on_dev else int(pairs.shape[0])
            h = C.c_void_p()
            check(lib().nb_graph_create_from_pairs(self.cs._h, self.global_vertices, n_edges, ptr(pairs), 1 if on_dev else 0, C.byref(h)))
            self._h = h
            self.global_edges = n_edges
            self.column_offset = self.row_indices = self.in_degree = self.out_degree = None   # device resident: device_arrays()
            return
        if column_offset is not None and hasattr(column_offset, "is_cuda") and column_offset.is_cuda:
            # CSC arrays already on the device (int32 tensors holding u32 values): adopted without a host round trip
            co, ri = column_offset.contiguous(), row_indices.contiguous()
            assert co.dtype == torch.int32 and ri.dtype == torch.int32 and co.numel() == self.global_vertices + 1
            torch.cuda.synchronize(co.device)     # the arrays may still be being produced on another stream than cuda_stream's
            h = C.c_void_p()
            check(lib().nb_graph_create_from_device(self.cs._h, self.global_vertices, int(ri.numel()), ptr(co), ptr(ri), C.byref(h)))
            self._h = h
            self.global_edges = int(ri.numel())
            self.column_offset = self.row_indices = self.in_degree = self.out_degree = None
            return
        if edge_pairs is not None:
            column_offset, row_indices, ind, outd = self.build_csc_host(edge_pairs, self.global_vertices)
            if in_degree is None:
                in_degree, out_degree = ind, outd
        self.column_offset = np.ascontiguousarray(column_offset, dtype=np.uint32)
        self.row_indices = np.ascontiguousarray(row_indices, dtype=np.uint32)
        self.global_edges = int(self.row_indices.size)
        self.in_degree = None if in_degree is None else np.ascontiguousarray(in_degree, dtype=np.uint32)
        self.out_degree = None if out_degree is None else np.ascontiguousarray(out_degree, dtype=np.uint32)
        h = C.c_void_p()
        check(lib().nb_graph_create(self.cs._h, self.global_vertices, self.global_edges, ptr(self.column_offset),
                                    ptr(self.row_indices), ptr(self.in_degree), ptr(self.out_degree), C.byref(h)))
        self._h = h

    @staticmethod
    def build_csc_host(edge_pairs, global_vertices):
        """(column_offset, row_indices, in_degree, out_degree) from (src,dst) pairs: column = dst, entries in
        file order (FullyRepGraph.hpp:761-794); degrees clamped >= 1 (core/graph.hpp:4525-4530)."""
        pairs = np.ascontiguousarray(edge_pairs, dtype=np.uint32).reshape(-1, 2)
        order = np.argsort(pairs[:, 1], kind="stable")
        row_indices = pairs[order, 0].copy()
        cnt_in = np.bincount(pairs[:, 1], minlength=global_vertices)
        column_offset = np.zeros(global_vertices + 1, np.uint32)
        np.cumsum(cnt_in, out=column_offset[1:])
        in_degree = np.maximum(cnt_in, 1).astype(np.uint32)
        out_degree = np.maximum(np.bincount(pairs[:, 0], minlength=global_vertices), 1).astype(np.uint32)
        return column_offset, row_indices, in_degree, out_degree

    @classmethod
    def from_edge_file(cls, cuda_stream, path, global_vertices):
        """EDGE_FILE format: raw little-endian (u32 src, u32 dst) pairs (core/FullyRepGraph.hpp:738-795)."""
        return cls(cuda_stream, global_vertices, edge_pairs=np.fromfile(path, dtype=np.uint32))

    def device_arrays(self):
        v, e = C.c_uint32(), C.c_uint64()
        p = [C.c_void_p() for _ in range(4)]
        check(lib().nb_graph_info(self._h, C.byref(v), C.byref(e), *[C.byref(x) for x in p]))
        dev = self.cs.device
        return (_view(p[0].value, v.value + 1, "u", dev, self), _view(p[1].value, e.value, "u", dev, self),
                _view(p[2].value, v.value, "u", dev, self), _view(p[3].value, v.value, "u", dev, self))

    def __del__(self):
        try:
            lib().nb_graph_destroy(self._h)
        except Exception:
            pass


class sampCSC:
    """One sampled layer (core/coocsc.hpp:24-462), device members only; tensors are zero-copy views of the sampler's arena and
    stay valid until the sampler's next batch. The views are built on first access (a training step touches four or five of the
    fourteen arrays; building all of them eagerly cost more host time per step than the GPU work of the step)."""

    _ARRAYS = {  # attribute -> (view field, length attribute, +1, kind)
        "dev_destination": ("destination", "v_size", 0, "u"), "dev_column_offset": ("column_offset", "v_size", 1, "u"),
        "dev_sample_ans": ("sample_ans", "e_size", 0, "u"), "dev_row_indices": ("row_indices", "e_size", 0, "u"),
        "dev_source": ("source", "src_size", 0, "u"), "dev_row_offset": ("row_offset", "src_size", 1, "u"),
        "dev_column_indices": ("column_indices", "e_size", 0, "u"), "dev_csr_to_csc": ("csr_to_csc", "e_size", 0, "u"),
        "dev_edge_weight_forward": ("edge_weight_forward", "e_size", 0, "f"),
        "dev_edge_weight_backward": ("edge_weight_backward", "e_size", 0, "f"),
        "dev_dst_local_id": ("dst_local_id", "v_size", 0, "u"), "dev_src_to_dst": ("src_to_dst", "src_size", 0, "u"),
        "dev_source_use_count": ("source_use_count", "src_size", 0, "u"), "dev_gather_index": ("gather_index", "e_size", 0, "u")}
    _OPTIONAL = {"dev_row_offset", "dev_column_indices", "dev_csr_to_csc", "dev_edge_weight_backward", "dev_dst_local_id",
                 "dev_src_to_dst", "dev_source_use_count", "dev_gather_index"}

    def __init__(self, view, device, owner):
        self.v_size, self.e_size, self.src_size = view.n_dst, view.n_edges, view.n_src
        self._ptr = {f: getattr(view, f) for f, _, _, _ in self._ARRAYS.values()}
        self._device, self._owner = device, owner

    def __getattr__(self, name):     # only reached for attributes not set yet
        spec = sampCSC._ARRAYS.get(name)
        if spec is None:
            if name == "edge_weight":   # the GPU-sampled path's name (coocsc.hpp:440)
                return self.dev_edge_weight_forward
            raise AttributeError(name)
        field, n_attr, plus, kind = spec
        addr = self._ptr[field]
        if not addr and name in sampCSC._OPTIONAL:
            t = None
        else:
            t = _view(addr, getattr(self, n_attr) + plus, kind, self._device, self._owner)
        setattr(self, name, t)
        return t

    def address(self, name):
        """raw device address of an array (no tensor is built): for callers that only forward pointers to the C ABI"""
        return self._ptr[sampCSC._ARRAYS[name][0]]

    # accessor names of the reference
    def dev_dst(self): return self.dev_destination
    def dev_src(self): return self.dev_source
    def dev_c_o(self): return self.dev_column_offset
    def dev_r_i(self): return self.dev_row_indices
    def dev_e_w_f(self): return self.dev_edge_weight_forward
    def dev_c_i(self): return self.dev_column_indices
    def dev_r_o(self): return self.dev_row_offset
    def dev_e_w_b(self): return self.dev_edge_weight_backward
    def dev_e_w(self): return self.edge_weight


class SampledSubgraph:
    """core/FullyRepGraph.hpp:30-681 (GPU members): `sampled_sgs[i]` is sampling layer i
    (layer 0 = the seeds' layer); sampled_sgs[i+1].dev_destination aliases sampled_sgs[i].dev_source."""

    def __init__(self, layers, cs):
        self.sampled_sgs = layers
        self.layers = len(layers)
        self.cs = cs


class FastSampler:
    """GPU FastSampler (core/ntsFastSampler.hpp:125-176, 648-915). One nb_sampler per pipeline slot."""

    def __init__(self, whole_graph, index, layers, batch_size, fanout, pipeline_num=1, cuda_stream=None,
                 merge_src_dst=False, up_degree=False, build_csr=True, rng_seed=0x5EED0004, bottom_csr=True):
        assert len(index) > 0
        self.whole_graph = whole_graph
        self.sample_nids = np.ascontiguousarray(index, dtype=np.uint32).copy()
        self.work_range = [0, self.sample_nids.size]
        self.work_offset = 0
        self.layer = int(layers)
        self.fanout = [int(f) for f in fanout]
        assert len(self.fanout) == self.layer
        self.batch_size = int(batch_size)
        streams = cuda_stream if isinstance(cuda_stream, (list, tuple)) else [cuda_stream or whole_graph.cs]
        self.cs_array = list(streams) + [streams[0]] * max(0, pipeline_num - len(streams))
        self.cs = self.cs_array[0]
        self.flags = (NB_SAMPLER_MERGE_SRC_DST if merge_src_dst else 0) | (NB_SAMPLER_UP_DEGREE if up_degree else 0) | \
                     (NB_SAMPLER_BUILD_CSR if build_csr else 0) | (0 if bottom_csr else NB_SAMPLER_NO_BOTTOM_CSR)
        fan = (C.c_int * self.layer)(*self.fanout)
        self._samplers = []
        for i in range(max(1, pipeline_num)):
            h = C.c_void_p()
            check(lib().nb_sampler_create(self.cs_array[i]._h, whole_graph._h, self.layer, fan, self.batch_size,
                                          self.flags, 0, C.byref(h)))
            self._samplers.append(h)
        self.rng_seed = int(rng_seed)
        self.batch_counter = 0
        self.ssg = None

    def __del__(self):
        try:
            for h in self._samplers:
                lib().nb_sampler_destroy(h)
        except Exception:
            pass

    def sample_not_finished(self):
        return self.work_offset < self.work_range[1]

    def restart(self):
        self.work_offset = self.work_range[0]

    def set_merge_src_dst(self, sg_num=1):
        raise NtsError("pass merge_src_dst=True to the constructor (arenas are sized once)")

    def _finish(self, ssg_id, views):
        cs = self.cs_array[ssg_id]
        owner = self
        self.ssg = SampledSubgraph([sampCSC(v, cs.device, owner) for v in views], cs)
        return self.ssg

    def sample_gpu_fast(self, batch_size_, ssg_id=0, weightType=WeightType.Sum, CacheFlag=None, super_batch_id=0xFFFFFFFF,
                        sync=True):
        """One mini-batch, all layers, no host round trip (replaces :648-709 and, with CacheFlag, :711-915)."""
        assert self.work_offset < self.work_range[1]
        n = min(int(batch_size_), self.work_range[1] - self.work_offset)
        seeds = self.sample_nids[self.work_offset:self.work_offset + n]
        views = (LayerView * self.layer)()
        check(lib().nb_sampler_sample(self._samplers[ssg_id], ptr(seeds), n, 0, self.rng_seed, self.batch_counter,
                                      int(weightType), ptr(CacheFlag), super_batch_id, views, 1 if sync else 0))
        self.work_offset += n
        self.batch_counter += 1
        return self._finish(ssg_id, views) if sync else None

    def wait(self, ssg_id=0):
        """Completes a sample_gpu_fast(..., sync=False) on pipeline slot ssg_id and returns its SampledSubgraph."""
        views = (LayerView * self.layer)()
        check(lib().nb_sampler_wait(self._samplers[ssg_id], views))
        return self._finish(ssg_id, views)

    def sample_gpu_fast_omit(self, batch_size_, CacheFlag, super_batch_id=0xFFFFFFFF, weightType=WeightType.Sum, ssg_id=0):
        return self.sample_gpu_fast(batch_size_, ssg_id, weightType, CacheFlag, super_batch_id)

    def replay(self, seeds, sample_ans_per_layer, weightType=WeightType.Sum, ssg_id=0):
        """Replays recorded neighbour draws (sampCSC::sample_ans per layer): the bit-exact path."""
        seeds = np.ascontiguousarray(seeds, dtype=np.uint32)
        arrs = [np.ascontiguousarray(a, dtype=np.uint32) for a in sample_ans_per_layer]
        pp = (C.c_void_p * self.layer)(*[a.ctypes.data if a.size else None for a in arrs])
        ne = (C.c_uint32 * self.layer)(*[a.size for a in arrs])
        views = (LayerView * self.layer)()
        check(lib().nb_sampler_replay(self._samplers[ssg_id], ptr(seeds), seeds.size, pp, ne, int(weightType), views))
        return self._finish(ssg_id, views)

    # -- loaders (ntsFastSampler.hpp:227-317, 400-426, 472-529)
    def load_feature_gpu(self, cuda_stream, subgraph, local_feature, global_feature_buffer, lazy=False):
        """X0[i,:] = table[source_bottom[i],:] (core/ntsFastSampler.hpp:244-261). lazy=True returns a LazyFeature instead of
        copying: the bottom hop's SingleGPU[All]SampleGraphOp.forward then aggregates straight from the table through the global
        ids (identical bits), and X0 is written only if somebody asks for it (LazyFeature.materialize())."""
        l = subgraph.sampled_sgs[self.layer - 1]
        if lazy:
            return LazyFeature(cuda_stream, l, local_feature, global_feature_buffer)
        F = local_feature.shape[1]
        if local_feature.shape[0] != l.src_size:
            local_feature.resize_(l.src_size, F)
        cuda_stream.zero_copy_feature_move_gpu(local_feature, global_feature_buffer, l.dev_source, F, l.src_size,
                                               _pitch(global_feature_buffer, F), _pitch(local_feature, F))
        return local_feature

    def load_feature_gpu_cache(self, cuda_stream, subgraph, local_feature, global_feature_buffer, dev_cache_feature,
                               dev_cache_node_hashmap, hit_count=None):
        l = subgraph.sampled_sgs[self.layer - 1]
        F = local_feature.shape[1]
        if local_feature.shape[0] != l.src_size:
            local_feature.resize_(l.src_size, F)
        cuda_stream.gather_feature_cached(local_feature, global_feature_buffer, dev_cache_feature, dev_cache_node_hashmap,
                                          l.dev_source, F, l.src_size, hit_count)
        return local_feature

    def load_label_gpu(self, cuda_stream, subgraph, local_label, global_label_buffer):
        l = subgraph.sampled_sgs[0]
        if local_label.shape[0] != l.v_size:
            local_label.resize_(l.v_size)
        cuda_stream.global_copy_label_move_gpu(local_label, global_label_buffer, l.dev_destination, l.v_size)
        return local_label

    def load_share_embedding(self, cuda_stream, subgraph, dev_embedding, share_embedding, dev_cache_map,
                             dev_cache_location, super_batch_id):
        l = subgraph.sampled_sgs[self.layer - 1]
        cuda_stream.dev_load_share_embedding(dev_embedding, share_embedding, dev_cache_map, dev_cache_location,
                                             dev_embedding.shape[1], l.dev_destination, l.v_size, super_batch_id)


class LazyFeature:
    """The gathered bottom-layer input X0 = table[source,:] as a promise. The only consumer in the GCN toolkits is the bottom
    hop's aggregation (toolkits/GCN_SAMPLE_GPU.hpp:360-374), which can read the table rows through the layer's global ids
    (sample_ans) directly: same rows, same order, same bits, and the [S,F] copy (write S*4F + read E*4F) never happens.
    `materialize()` performs the ordinary gather for any other reader (GraphSAGE's self term, debugging)."""

    def __init__(self, cuda_stream, layer, local_feature, table):
        self.cs, self.layer_csc, self.buffer, self.table = cuda_stream, layer, local_feature, table
        self.shape = (layer.src_size, table.shape[1])
        self._x0 = None

    def materialize(self):
        if self._x0 is None:
            l, F = self.layer_csc, self.shape[1]
            buf = self.buffer
            if buf.shape[0] != l.src_size:
                buf = buf[:l.src_size] if buf.shape[0] > l.src_size else buf.resize_(l.src_size, F)
            self.cs.zero_copy_feature_move_gpu(buf, self.table, l.dev_source, F, l.src_size, _pitch(self.table, F), _pitch(buf, F))
            self._x0 = buf
        return self._x0


def preSample(train_ids, batch_size, pipeline_num, layers, whole_graph, cache_rate=0.8, cuda_stream=None):
    """nts::op::preSample (core/ntsBaseOp.hpp:415-470) on the GPU: for every super-batch (batch_size * pipeline_num seeds) the hot
    vertices of its (layers-1)-hop in-neighbourhood. Returns (batch_cache_num u32[#super_batches], batch_cache_ids u32[sum])."""
    cs = cuda_stream or whole_graph.cs
    train_ids = np.ascontiguousarray(train_ids, dtype=np.uint32)
    sb = int(batch_size) * int(pipeline_num)
    V = whole_graph.global_vertices
    out = torch.empty(V, dtype=torch.int32, device=cs.device)
    counts, ids = [], []
    for start in range(0, train_ids.size, sb):
        seeds = train_ids[start:start + sb]
        n = C.c_uint32()
        check(lib().nb_hotness(cs._h, whole_graph._h, ptr(seeds), seeds.size, 0, int(layers), float(cache_rate), 0xFFFFFFFF,
                               ptr(out), V, C.byref(n), None))
        counts.append(n.value)
        ids.append(out[:n.value].cpu().numpy().view(np.uint32).copy())
    return np.array(counts, np.uint32), (np.concatenate(ids) if ids else np.zeros(0, np.uint32))


def write_pre_sample_file(path, batch_cache_num, batch_cache_ids):
    """PRE_SAMPLE_FILE layout (core/ntsBaseOp.hpp:477-495): u32 counts[#super_batches] || u32 ids[sum(counts)]"""
    np.concatenate([np.asarray(batch_cache_num, np.uint32), np.asarray(batch_cache_ids, np.uint32)]).tofile(path)


def read_pre_sample_file(path, n_super_batches, of_rate=1.0):
    """reader (core/ntsBaseOp.hpp:497-538): the first counts[i]*of_rate ids of every super-batch group"""
    raw = np.fromfile(path, dtype=np.uint32)
    counts = raw[:n_super_batches]
    take = (counts.astype(np.float32) * np.float32(of_rate)).astype(np.uint32)
    out, pos = [], n_super_batches
    for c, t in zip(counts, take):
        out.append(raw[pos:pos + t])
        pos += int(c)
    return take, (np.concatenate(out) if out else np.zeros(0, np.uint32))


def set_cache_index(cuda_stream, cache_map, cache_location, super_batch_id, cache_ids_dev, n):
    """GNNDatum::set_cache_index (core/ntsDataloador.hpp:440-478) on device arrays"""
    check(lib().nb_set_cache_index(cuda_stream._h, ptr(cache_map), ptr(cache_location), super_batch_id, ptr(cache_ids_dev), n))


class ColdStage:
    """Cold rows from a host-resident feature table, staged per batch through pinned memory on a side stream while the previous
    batch trains; hot rows from an HBM cache table (replaces load_feature_gpu_cache's CPU split + zero-copy, ntsFastSampler.hpp:263-317)."""

    def __init__(self, cuda_stream, host_table, max_rows):
        assert host_table.dim() == 2 and host_table.stride(1) == 1 and not host_table.is_cuda
        self.cs, self.host_table, self.F = cuda_stream, host_table, host_table.shape[1]
        h = C.c_void_p()
        check(lib().nb_stage_create(cuda_stream._h, ptr(host_table), host_table.stride(0), self.F, int(max_rows), C.byref(h)))
        self._h = h

    def submit(self, slot, ids_dev, n_rows, cache_node_hashmap_dev):
        check(lib().nb_stage_submit(self._h, slot, ptr(ids_dev), n_rows, ptr(cache_node_hashmap_dev)))

    def gather(self, slot, out, cache_table, cache_node_hashmap_dev, ids_dev):
        n_cold = C.c_uint32()
        check(lib().nb_stage_gather(self._h, slot, ptr(out), _pitch(out, self.F), ptr(cache_table), _pitch(cache_table, self.F),
                                    ptr(cache_node_hashmap_dev), ptr(ids_dev), C.byref(n_cold)))
        return n_cold.value

    def gather_table(self, slot, out, hot_table, cache_node_hashmap_dev, ids_dev):
        """hot rows from a (GPU-sharded) FeatureTable indexed by cache slot, cold rows from the staged block"""
        n_cold = C.c_uint32()
        check(lib().nb_stage_gather_table(self._h, slot, ptr(out), _pitch(out, self.F), hot_table._h, ptr(cache_node_hashmap_dev),
                                          ptr(ids_dev), C.byref(n_cold)))
        return n_cold.value

    def __del__(self):
        try:
            lib().nb_stage_destroy(self._h)
        except Exception:
            pass


class FeatureTable:
    """HBM-resident feature table, optionally row-sharded over the GPUs of a node (row v on shard
    v % n at local row v // n) and read peer-to-peer inside the gather kernel."""

    def __init__(self, cuda_stream, shard_ptrs, feature_size, pitch, n_rows_total, keepalive=None):
        self.cs = cuda_stream
        arr = (C.c_void_p * len(shard_ptrs))(*[ptr(p) for p in shard_ptrs])
        h = C.c_void_p()
        check(lib().nb_table_create(cuda_stream._h, len(shard_ptrs), arr, feature_size, pitch, n_rows_total, C.byref(h)))
        self._h = h
        self.feature_size = feature_size
        self._keep = keepalive

    def gather(self, out, ids, n_rows):
        check(lib().nb_table_gather(self.cs._h, self._h, ptr(out), ptr(ids), n_rows, _pitch(out, self.feature_size)))
        return out

    def __del__(self):
        try:
            lib().nb_table_destroy(self._h)
        except Exception:
            pass


# ---------------------------------------------------------------------------------------------
# graph operators: forward(X[S_hop,F]) -> Y[V_hop,F]; backward(dY[V_hop,F]) -> dX[S_hop,F]
# (core/ntsBaseOp.hpp:28-45 is the interface; NtsContext::runGraphOp/self_backward drive it).
class _AggFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, op):
        ctx.op = op
        return op.forward(x)

    @staticmethod
    def backward(ctx, dy):
        return ctx.op.backward(dy if dy.stride(-1) == 1 else dy.contiguous()), None


class SingleGPUAllSampleGraphOp:
    """core/ntsSingleGPUSampleGraphOp.hpp:195-294 (GPU-sampled graph). Forward = CSC segment reduce;
    backward = CSR segment reduce when the sampler built the CSR (deterministic, write-once),
    otherwise the CSC push with vector reductions (what the reference's Push_From_Dst_To_Src_Spmm computes)."""

    def __init__(self, subgraphs, layer, cuda_stream, with_weight=True):
        self.subgraphs, self.layer, self.cuda_stream, self.with_weight = subgraphs, layer, cuda_stream, with_weight

    def forward(self, f_input):
        l = self.subgraphs.sampled_sgs[self.layer]
        if isinstance(f_input, LazyFeature):
            if f_input.layer_csc is not l:
                f_input = f_input.materialize()
            else:   # bottom hop straight from the feature table: input row of edge e = table[sample_ans[e]]
                table, F = f_input.table, f_input.shape[1]
                out = _alloc_like_rows(l.v_size, F, table)
                w = l.address("dev_edge_weight_forward") if self.with_weight else None
                if l.address("dev_gather_index"):
                    self.cuda_stream.aggregate_gathered_fwd(table, l.address("dev_gather_index"), out, w, l.address("dev_column_offset"),
                                                            l.v_size, F, _pitch(table, F), _pitch(out, F))
                else:   # no packed index (|V| >= 2^31): plain aggregation through the global ids
                    self.cuda_stream.aggregate_fwd_pitched(table, out, w, l.address("dev_sample_ans"), l.address("dev_column_offset"),
                                                           l.v_size, F, _pitch(table, F), _pitch(out, F))
                return out
        F = f_input.shape[1]
        assert f_input.shape[0] == l.src_size
        out = _alloc_like_rows(l.v_size, F, f_input)
        # raw arena addresses (no tensor views are built on the per-step path); the math is Cuda_Stream::Gather_By_Dst_From_Src_Spmm's
        check(lib().nb_aggregate_csc_fwd_dyn(self.cuda_stream._h, f_input.data_ptr(), out.data_ptr(),
                                             l.address("dev_edge_weight_forward") if self.with_weight else None,
                                             l.address("dev_row_indices"), l.address("dev_column_offset"), None, l.v_size, F,
                                             _pitch(f_input, F), _pitch(out, F)))
        return out

    def backward(self, f_output_grad):
        l = self.subgraphs.sampled_sgs[self.layer]
        F = f_output_grad.shape[1]
        assert f_output_grad.shape[0] == l.v_size
        grad = _alloc_like_rows(l.src_size, F, f_output_grad)
        if l.address("dev_row_offset"):      # CSR segment reduce: deterministic, every row written once (Gather_By_Src_From_Dst_Spmm)
            check(lib().nb_aggregate_csr_bwd_dyn(self.cuda_stream._h, f_output_grad.data_ptr(), grad.data_ptr(),
                                                 l.address("dev_edge_weight_backward") if self.with_weight else None,
                                                 l.address("dev_row_offset"), l.address("dev_column_indices"), None, l.src_size, F,
                                                 _pitch(f_output_grad, F), _pitch(grad, F)))
        else:
            assert _pitch(f_output_grad, F) == F, "padded tensors need the CSR (build_csr=True)"
            self.cuda_stream.Push_From_Dst_To_Src_Spmm(f_output_grad, grad, l.dev_e_w(), l.dev_r_i(), l.dev_c_o(),
                                                       l.src_size, 0, 0, 0, 0, l.e_size, l.v_size, F, self.with_weight, False)
        return grad

    def __call__(self, x):
        if isinstance(x, LazyFeature):     # the gathered features are a leaf without a gradient (core/ntsContext.hpp:443)
            return self.forward(x)
        return _AggFn.apply(x, self)


SingleGPUSampleGraphOp = SingleGPUAllSampleGraphOp  # :50-176: same math (CSC forward, CSR backward)


class _GatFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, h, att, op):
        out = op.forward(h, att)
        ctx.op = op
        ctx.save_for_backward(h, att)
        return out

    @staticmethod
    def backward(ctx, dout):
        h, att = ctx.saved_tensors
        dh, datt = ctx.op.backward(h, att, dout.contiguous())
        return dh, datt, None


class GATFusedOp:
    """One op for the GAT layer body between X*W and relu (toolkits/GAT_SAMPLE_ALL_MULTI.hpp:383-464):
    out[d] = sum_e softmax_d(leaky_relu([h_src(e), h_dst(d)] . att, 0.2))[e] * h_src(e).
    Needs a sampler built with merge_src_dst=True (dst_local_id) and build_csr=True (backward)."""

    def __init__(self, subgraphs, layer, cuda_stream, negative_slope=0.2):
        self.subgraphs, self.layer, self.cs, self.slope = subgraphs, layer, cuda_stream, float(negative_slope)
        self.score_pre = self.alpha = None

    def forward(self, h, att):
        l = self.subgraphs.sampled_sgs[self.layer]
        F = h.shape[1]
        assert h.is_contiguous() and h.shape[0] == l.src_size and att.numel() == 2 * F
        self.score_pre = torch.empty(l.e_size, dtype=torch.float32, device=h.device)
        self.alpha = torch.empty(l.e_size, dtype=torch.float32, device=h.device)
        out = torch.empty((l.v_size, F), dtype=torch.float32, device=h.device)
        check(lib().nb_gat_fwd(self.cs._h, ptr(h), ptr(att.contiguous()), self.slope, ptr(l.dev_c_o()), ptr(l.dev_r_i()),
                               ptr(l.dev_dst_local_id), l.v_size, l.src_size, F, ptr(self.score_pre), ptr(self.alpha), ptr(out)))
        return out

    def backward(self, h, att, dout):
        l = self.subgraphs.sampled_sgs[self.layer]
        F = h.shape[1]
        dh = torch.empty_like(h)
        datt = torch.empty(2 * F, dtype=torch.float32, device=h.device)
        check(lib().nb_gat_bwd(self.cs._h, ptr(h), ptr(att.contiguous()), self.slope, ptr(dout), ptr(self.score_pre),
                               ptr(self.alpha), ptr(l.dev_c_o()), ptr(l.dev_r_i()), ptr(l.dev_dst_local_id), ptr(l.dev_r_o()),
                               ptr(l.dev_c_i()), ptr(l.dev_csr_to_csc), ptr(l.dev_src_to_dst), l.v_size, l.src_size, l.e_size, F,
                               ptr(dh), ptr(datt)))
        return dh, datt.view_as(att)

    def __call__(self, h, att):
        return _GatFn.apply(h, att, self)


class BatchGPUSrcDstScatterOp:
    """core/ntsPushdownGraphOp.hpp:490-576"""

    def __init__(self, subgraphs, layer, cuda_stream):
        self.subgraphs, self.layer, self.cs = subgraphs, layer, cuda_stream

    def forward(self, f_input):
        l = self.subgraphs.sampled_sgs[self.layer]
        F = f_input.shape[1]
        out = torch.empty((l.e_size, 2 * F), dtype=torch.float32, device=f_input.device)
        self.cs.Scatter_Src_Dst_to_Msg(out, f_input, l.dev_r_i(), l.dev_c_o(), l.v_size, F, l.dev_dst_local_id)
        return out

    def backward(self, f_output_grad):
        l = self.subgraphs.sampled_sgs[self.layer]
        F = f_output_grad.shape[1] // 2
        grad = torch.empty((l.src_size, F), dtype=torch.float32, device=f_output_grad.device)
        self.cs.Gather_Msg_To_Src_Dst(grad, f_output_grad, l.dev_r_i(), l.dev_c_o(), l.v_size, F, l.dev_dst_local_id)
        return grad


class BatchGPUEdgeSoftMax:
    """core/ntsPushdownGraphOp.hpp:578-667"""

    def __init__(self, subgraphs, layer, cuda_stream):
        self.subgraphs, self.layer, self.cs = subgraphs, layer, cuda_stream
        self.IntermediateResult = None

    def forward(self, f_input):
        l = self.subgraphs.sampled_sgs[self.layer]
        out = torch.empty_like(f_input)
        self.IntermediateResult = torch.empty_like(f_input)
        self.cs.Edge_Softmax_Forward_Norm_Block(out, f_input, self.IntermediateResult, l.dev_r_i(), l.dev_c_o(), l.v_size, 1)
        return out

    def backward(self, f_output_grad):
        l = self.subgraphs.sampled_sgs[self.layer]
        grad = torch.empty_like(f_output_grad)
        self.cs.Edge_Softmax_Backward_Block(grad, f_output_grad, self.IntermediateResult, l.dev_r_i(), l.dev_c_o(), l.v_size, 1)
        return grad


class BatchGPUAggregateDst:
    """core/ntsPushdownGraphOp.hpp:670-747"""

    def __init__(self, subgraphs, layer, cuda_stream):
        self.subgraphs, self.layer, self.cs = subgraphs, layer, cuda_stream

    def forward(self, f_input):
        l = self.subgraphs.sampled_sgs[self.layer]
        F = f_input.shape[1]
        out = torch.empty((l.v_size, F), dtype=torch.float32, device=f_input.device)
        self.cs.Gather_Msg_to_Dst(out, f_input, l.dev_r_i(), l.dev_c_o(), l.v_size, F)
        return out

    def backward(self, f_output_grad):
        l = self.subgraphs.sampled_sgs[self.layer]
        F = f_output_grad.shape[1]
        grad = torch.empty((l.e_size, F), dtype=torch.float32, device=f_output_grad.device)
        self.cs.Scatter_Dst_to_Msg(grad, f_output_grad, l.dev_r_i(), l.dev_c_o(), l.v_size, F)
        return grad
