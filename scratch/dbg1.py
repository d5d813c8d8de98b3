import sys; sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import numpy as np, torch
import __graft_entry__ as ge
import oracle
from golden_util import load
nts = ge.load_package()
cs = nts.Cuda_Stream(0)
g = load("synth300_takeall_3layer")
graph = nts.FullyRepGraph(cs, g["V"], column_offset=g["col_off"], row_indices=g["row_idx"], in_degree=g["in_deg"], out_degree=g["out_deg"])
b = g["batches"][0]
seeds = b["layers"][0]["destination"]
sampler = nts.FastSampler(graph, seeds, g["L"], len(seeds), g["fanout"], cuda_stream=cs)
sg = sampler.replay(seeds, [l["sample_ans"] for l in b["layers"]])
lay = sg.sampled_sgs[0]
X = torch.from_numpy(b["Y1"].reshape(-1, g["F"])).cuda()
op = nts.SingleGPUAllSampleGraphOp(sg, 0, cs)
Y = op.forward(X).cpu().numpy()
ref = b["Y0"].reshape(Y.shape)
bad = np.argwhere(Y != ref)
print(bad)
co = lay.dev_column_offset.cpu().numpy()
print("lens", np.diff(co)[np.unique(bad[:,0])])
print(Y[np.unique(bad[:,0])], ref[np.unique(bad[:,0])])
