"""SURVEY section 8(d) sweep of the K1 kernel pair through the C ABI: r in {4, 24, 48, 96, 192} x L in {56, 92, 320, 664} x gate in
{large, middle_x, middle_y, small}, B chosen so that M = B*L is about 96 000 tokens, d = 768, bf16.  CUDA events around
graph replays of vlpet_k1_fwd / vlpet_k1_bwd, L2 flushed between replays, median of --iters.  Writes one JSON document
(default profiles/r2_micro_sweep.json): per case the path taken (1 = fused tcgen05, 2 = row-wise, 0 = generic), the time
and the achieved ALGORITHMIC bandwidth (3 d e bytes/token forward, 5 d e backward) as a fraction of the measured HBM peak."""
import argparse, ctypes as C, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vlpet_b200 as V  # noqa: F401
from vlpet_b200 import _lib as L

ap = argparse.ArgumentParser()
ap.add_argument("--iters", type=int, default=5)
ap.add_argument("--out", default=os.path.join(os.path.dirname(__file__), "..", "profiles", "r2_micro_sweep.json"))
ap.add_argument("--ranks", type=int, nargs="+", default=[4, 24, 48, 96, 192])
ap.add_argument("--lens", type=int, nargs="+", default=[56, 92, 320, 664])
ap.add_argument("--gates", nargs="+", default=["large", "middle_x", "middle_y", "small"])
ap.add_argument("--tokens", type=int, default=96000)
a = ap.parse_args()
d, bf = 768, torch.bfloat16
peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
g = torch.Generator(device="cuda").manual_seed(0)
junk = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
p_ = lambda t: C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)  # noqa: E731
mk = lambda *sh, std: (torch.randn(*sh, device="cuda", generator=g) * std).to(bf)  # noqa: E731


def timed(fn, iters):
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        fn(side); side.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=side):
            fn(torch.cuda.current_stream())
    ts = []
    for _ in range(iters):
        junk.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); graph.replay(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


results = []
for L_seq in a.lens:
    B = max(1, round(a.tokens / L_seq))
    M = B * L_seq
    x1 = torch.randn(M, d, device="cuda", generator=g).to(bf)
    x2 = (0.5 * torch.randn(M, d, device="cuda", generator=g)).to(bf)
    dout = torch.randn(M, d, device="cuda", generator=g).to(bf)
    out, dx1, dx2 = torch.empty_like(x1), torch.empty_like(x1), torch.empty_like(x1)
    for r in a.ranks:
        Wd, bd, Wu, bu = mk(r, d, std=0.05), mk(r, std=0.02), mk(d, r, std=0.05), mk(d, std=0.02)
        Gd, gbd, Gu, gbu = mk(r, d, std=0.05), mk(r, std=0.02), mk(d, r, std=0.05), mk(d, std=0.02)
        gw1, gw2, gb, gz = mk(d, std=0.05), mk(2 * d, std=0.05), mk(1, std=0.02), mk(d, std=0.05)
        f32 = lambda t: torch.zeros(t.shape, dtype=torch.float32, device="cuda")  # noqa: E731
        for gate in a.gates:
            desc = L.K1Desc(M=M, L=L_seq, d=d, r=r, rg=r if gate == "large" else 0, gate=L.GATE_IDS[gate], add_gate=0, dtype=L.BF16,
                            impl=L.IMPL_AUTO, s=1.0, alpha=1.0, kappa=1.0, p_drop=0.0, seed=0, seed_dev=None)
            w = L.K1Params(Wd=p_(Wd), bd=p_(bd), Wu=p_(Wu), bu=p_(bu))
            gkeep = [f32(Wd), f32(bd), f32(Wu), f32(bu)]          # (the tensors must outlive the calls: only pointers travel)
            gr = L.K1Grads(dWd=p_(gkeep[0]), dbd=p_(gkeep[1]), dWu=p_(gkeep[2]), dbu=p_(gkeep[3]))
            keep = []
            if gate == "large":
                w.Gd, w.gbd, w.Gu, w.gbu = p_(Gd), p_(gbd), p_(Gu), p_(gbu)
                keep = [f32(Gd), f32(gbd), f32(Gu), f32(gbu)]
                gr.dGd, gr.dgbd, gr.dGu, gr.dgbu = (p_(t) for t in keep)
            elif gate in ("middle_x", "small"):
                w.gw, w.gb = p_(gw1 if gate == "middle_x" else gw2), p_(gb)
                keep = [f32(gw1 if gate == "middle_x" else gw2), f32(gb)]
                gr.dgw, gr.dgb = p_(keep[0]), p_(keep[1])
            else:
                w.gz = p_(gz)
                keep = [f32(gz)]
                gr.dgz = p_(keep[0])
            pf, pb = L.lib.vlpet_k1_fwd_is_fused(C.byref(desc)), L.lib.vlpet_k1_bwd_is_fused(C.byref(desc))
            wsf = torch.empty(L.lib.vlpet_k1_fwd_workspace_bytes(C.byref(desc)) + 256, dtype=torch.uint8, device="cuda")
            wsb = torch.empty(L.lib.vlpet_k1_bwd_workspace_bytes(C.byref(desc)) + 256, dtype=torch.uint8, device="cuda")

            def fwd(st):
                L.check(L.lib.vlpet_k1_fwd(C.byref(desc), p_(x1), p_(x2), C.byref(w), p_(out), p_(wsf), wsf.numel(),
                                           C.c_void_p(st.cuda_stream)), "vlpet_k1_fwd")

            def bwd(st):
                L.check(L.lib.vlpet_k1_bwd(C.byref(desc), p_(x1), p_(x2), p_(dout), C.byref(w), p_(dx1), p_(dx2), C.byref(gr), p_(wsb),
                                           wsb.numel(), C.c_void_p(st.cuda_stream)), "vlpet_k1_bwd")
            it = a.iters if (pf and pb) else max(2, a.iters // 2)
            tf, tb = timed(fwd, it), timed(bwd, it)
            rec = {"L": L_seq, "B": B, "M": M, "r": r, "gate": gate, "path_fwd": int(pf), "path_bwd": int(pb),
                   "fwd_us": round(tf, 1), "bwd_us": round(tb, 1),
                   "fwd_frac": round(3 * M * d * 2 / (tf * 1e-6) / 1e9 / peak, 3),
                   "bwd_frac": round(5 * M * d * 2 / (tb * 1e-6) / 1e9 / peak, 3),
                   "fwd_bwd_frac": round(8 * M * d * 2 / ((tf + tb) * 1e-6) / 1e9 / peak, 3)}
            results.append(rec)
            print(json.dumps(rec), flush=True)
            del wsf, wsb
doc = {"what": "K1 micro sweep (SURVEY 8d): C-ABI calls replayed from CUDA graphs, L2 flushed between replays, median",
       "d": d, "dtype": "bf16", "hbm_peak_GBps": peak, "paths": {"1": "fused tcgen05", "2": "row-wise (+ tcgen05 adapter kernel at r >= 24, two rank halves at r > 96)",
                                                               "3": "large gate at 96 < r <= 192: ungated tcgen05 kernels per rank half + element-wise gate kernel",
                                                               "0": "generic CUDA-core path"},
       "results": results}
with open(a.out, "w") as fh:
    json.dump(doc, fh, indent=1)
