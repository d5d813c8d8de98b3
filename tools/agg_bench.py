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
    sgs = []
    F = 602
    dense = torch.rand((v, 602), device='cuda')
    padded = torch.zeros((v, 608), device='cuda'); padded[:, :602] = dense
    def run(name, fn, nbytes, reps=8):
        for _ in range(2): fn(0)
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps+1)]
        for i in range(reps):
            ev[i].record(stream); fn(i)
        ev[reps].record(stream); torch.cuda.synchronize()
        ms = np.median([ev[i].elapsed_time(ev[i+1]) for i in range(reps)])
        print(f"{name:44s} {ms*1e3:8.1f} us {nbytes/ms/1e6:8.1f} GB/s(alg)")
    # several batches' layers kept alive: need separate samplers (arena reuse) -> just use one batch, rotate inputs
    sg = fs.sample_gpu_fast(1024)
    bot = sg.sampled_sgs[1]
    S, E, V1 = bot.src_size, bot.e_size, bot.v_size
    print("S,E,V1", S, E, V1)
    xs_d = [torch.rand((S, 602), device='cuda') for _ in range(3)]
    xs_p = [torch.rand((S, 608), device='cuda') for _ in range(3)]
    y_d = torch.empty((V1, 602), device='cuda'); y_p = torch.empty((V1, 608), device='cuda')
    nb = E*(8+4*F) + 4*(V1+1) + 4*V1*F
    agg = lambda x, y, ri, pin, pout: check(lib.nb_aggregate_csc_fwd_dyn(cs._h, ptr(x), ptr(y), ptr(bot.dev_e_w()), ptr(ri), ptr(bot.dev_c_o()),
                                       ptr(bot.dev_c_o()) + 0*4 if False else _nd, V1, F, pin, pout))
    nd = C.c_void_p(); check(lib.nb_sampler_sizes_dev(fs._samplers[0], 1, C.byref(nd), None, None, None, None, None)); _nd = nd.value
    run("agg dense 602 (float2, CHUNK10)", lambda i: agg(xs_d[i%3], y_d, bot.dev_r_i(), 602, 602), nb)
    run("agg padded 608 (float4, F_eff 604)", lambda i: agg(xs_p[i%3], y_p, bot.dev_r_i(), 608, 608), nb)
    run("agg fused from table dense (no X0)", lambda i: agg(dense, y_d, bot.dev_sample_ans, 602, 602), nb)
    run("agg fused from table padded (no X0)", lambda i: agg(padded, y_p, bot.dev_sample_ans, 608, 608), nb)
    # correctness of fused vs separate
    x0 = torch.empty((S, 608), device='cuda')
    check(lib.nb_gather_rows(cs._h, ptr(x0), ptr(padded), ptr(bot.dev_source), S, 602, 608, 608))
    ya = torch.empty((V1, 608), device='cuda'); yb = torch.empty((V1, 608), device='cuda')
    agg(x0, ya, bot.dev_r_i(), 608, 608); agg(padded, yb, bot.dev_sample_ans, 608, 608)
    torch.cuda.synchronize()
    print("fused == separate:", torch.equal(ya[:, :602], yb[:, :602]))
    # top layer F=128
    top = sg.sampled_sgs[0]
    h = torch.rand((top.src_size, 128), device='cuda'); y0 = torch.empty((1024, 128), device='cuda'); dh = torch.empty_like(h); dy = torch.rand((1024,128), device='cuda')
    nb0 = top.e_size*(8+512) + 4*1025 + 4*1024*128
    run("agg fwd top F=128", lambda i: check(lib.nb_aggregate_csc_fwd(cs._h, ptr(h), ptr(y0), ptr(top.dev_e_w()), ptr(top.dev_r_i()), ptr(top.dev_c_o()), 1024, top.src_size, 128)), nb0)
    nb1 = top.e_size*(8+512) + 4*(top.src_size+1) + 4*top.src_size*128
    run("agg bwd top F=128 (CSR)", lambda i: check(lib.nb_aggregate_csr_bwd(cs._h, ptr(dy), ptr(dh), ptr(top.dev_e_w_b()), ptr(top.dev_r_o()), ptr(top.dev_c_i()), top.src_size, 1024, 128)), nb1)
    # sampler alone
    def samp(i):
        fs.work_offset = (i % 50) * 1024
        check(lib.nb_sampler_sample(fs._samplers[0], ptr(fs.sample_nids[fs.work_offset:fs.work_offset+1024]), 1024, 0, 1, i, 0, None, 0xFFFFFFFF, None, 0))
    run("sampler batch (graph, async)", samp, 1, reps=20)
