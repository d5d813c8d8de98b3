import sys; sys.path.insert(0,'.')
import numpy as np, torch
import __graft_entry__ as ge
nts = ge.load_package()
lib, check, ptr = nts._capi.lib(), nts._capi.check, nts._capi.ptr
V, F, N = 232965, 602, 140000
stream = torch.cuda.Stream()
with torch.cuda.stream(stream):
    cs = nts.Cuda_Stream(0, stream)
    g = torch.Generator(device='cuda').manual_seed(1)
    def bench(name, table, out, tp, op, variant, F_eff=F):
        check(lib.nb_set_option(b"gather_variant", variant))
        ids = [torch.randint(0, V, (N,), device='cuda', dtype=torch.int32, generator=g) for _ in range(12)]
        for i in range(3):
            check(lib.nb_gather_rows(cs._h, ptr(out), ptr(table), ptr(ids[i]), N, F_eff, tp, op))
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(10)]
        for i in range(9):
            ev[i].record(stream)
            check(lib.nb_gather_rows(cs._h, ptr(out), ptr(table), ptr(ids[3+i]), N, F_eff, tp, op))
        ev[9].record(stream)
        torch.cuda.synchronize()
        ms = np.median([ev[i].elapsed_time(ev[i+1]) for i in range(9)])
        ok = torch.equal(out[:, :F], table[ids[11].long()][:, :F])
        print(f"{name:62s} {ms*1e3:8.1f} us  {N*(4+8*F)/ms/1e6:8.1f} GB/s (alg)  ok={ok}")
    dense = torch.rand((V, 602), device='cuda')
    out_dense = torch.empty((N, 602), device='cuda')
    padded = torch.zeros((V, 608), device='cuda'); padded[:, :602] = dense
    out_pad = torch.empty((N, 608), device='cuda')
    bench("LSU dense 602->602 (float2)", dense, out_dense, 602, 602, 0)
    bench("LSU pitch 608->608 F=602 (float2)", padded, out_pad, 608, 608, 0)
    bench("LSU pitch 608->608 F=604 (float4)", padded, out_pad, 608, 608, 0, 604)
    bench("LSU pitch 608->608 F=608 (float4)", padded, out_pad, 608, 608, 0, 608)
    bench("TMA pitch 608->608 F=602", padded, out_pad, 608, 608, 1)
    bench("TMA pitch 608->608 F=608", padded, out_pad, 608, 608, 1, 608)
    bench("TMA tensor map, tile::gather4 608->608 F=602 (4 boxes of 152)", padded, out_pad, 608, 608, 2)
    bench("TMA tensor map, tile::gather4 608->608 F=608", padded, out_pad, 608, 608, 2, 608)
    t128 = torch.rand((V, 128), device='cuda'); o128 = torch.empty((N, 128), device='cuda')
    F=128
    bench("LSU F=128", t128, o128, 128, 128, 0, 128)
    bench("TMA F=128", t128, o128, 128, 128, 1, 128)
    bench("tile::gather4 F=128", t128, o128, 128, 128, 2, 128)
    # narrow rows (products F=100, papers F=128) on a table larger than L2: TMA bulk rows vs one row per warp vs 4 rows per warp
    V = 2449029
    for Fn in (100, 128, 64, 256):
        F = Fn
        tn = torch.rand((V, Fn), device='cuda'); on = torch.empty((N, Fn), device='cuda')
        bench(f"TMA F={Fn} (V=2.4M)", tn, on, Fn, Fn, 1, Fn)
        bench(f"tile::gather4 F={Fn} (4 rows per instruction)", tn, on, Fn, Fn, 2, Fn)
        check(lib.nb_set_option(b"gather_narrow_rows", 1))
        bench(f"LSU F={Fn} 1 row/warp", tn, on, Fn, Fn, 0, Fn)
        check(lib.nb_set_option(b"gather_narrow_rows", 4))
        bench(f"LSU F={Fn} 4 rows/warp", tn, on, Fn, Fn, 0, Fn)
        del tn, on
    V, F = 232965, 602
    # copy peak for context
    a = torch.empty(N*608, device='cuda'); b = torch.empty_like(a)
    for _ in range(3): b.copy_(a)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(10): b.copy_(a)
    e1.record(stream); torch.cuda.synchronize()
    print("torch copy same bytes", 10*2*a.numel()*4/e0.elapsed_time(e1)/1e6, "GB/s")
