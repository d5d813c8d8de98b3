import sys; sys.path.insert(0,'.')
import numpy as np, torch, ctypes as C
import __graft_entry__ as ge
import bench as B
nts = ge.load_package()
lib, check, ptr = nts._capi.lib(), nts._capi.check, nts._capi.ptr
v, col_off, src = B.reddit_shaped_graph(1.0)
stream = torch.cuda.Stream()
with torch.cuda.stream(stream):
    cs = nts.Cuda_Stream(0, stream)
    graph = nts.FullyRepGraph(cs, v, column_offset=col_off, row_indices=src)
    seeds = B.train_seeds(v)
    fs = nts.FastSampler(graph, seeds, 2, 1024, [25, 10], cuda_stream=cs)
    for i in range(6):
        sg = fs.sample_gpu_fast(1024)
    print([ (l.v_size,l.e_size,l.src_size) for l in sg.sampled_sgs])
