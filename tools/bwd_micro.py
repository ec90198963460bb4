"""Developer tool: time the launches of vlpet_k1_bwd separately (tile kernel / column sums / weight-gradient GEMM) through
the C ABI, CUDA events on the launch stream, L2 flushed between launches (--flush 0: warm).  One JSON line per (M, part)."""
import argparse, ctypes as C, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vlpet_b200 as V  # noqa: F401
from vlpet_b200 import _lib as L

ap = argparse.ArgumentParser()
ap.add_argument("--M", type=int, nargs="+", default=[96000, 16800, 2128])
ap.add_argument("--r", type=int, default=96)
ap.add_argument("--iters", type=int, default=20)
ap.add_argument("--flush", type=int, default=1)
ap.add_argument("--gate", default="large")
ap.add_argument("--p-drop", type=float, default=0.0, help="dropout probability of the PET output (training: 0.1)")
a = ap.parse_args()
d, r, bf = 768, a.r, torch.bfloat16
peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
g = torch.Generator(device="cuda").manual_seed(0)
junk = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
p_ = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
for M in a.M:
    x1 = torch.randn(M, d, device="cuda", generator=g).to(bf)
    x2 = (0.5 * torch.randn(M, d, device="cuda", generator=g)).to(bf)
    dout = torch.randn(M, d, device="cuda", generator=g).to(bf)
    mk = lambda *sh, std: (torch.randn(*sh, device="cuda", generator=g) * std).to(bf)  # noqa: E731
    W = [mk(r, d, std=0.05), mk(r, std=0.02), mk(d, r, std=0.05), mk(d, std=0.02),
         mk(r, d, std=0.05), mk(r, std=0.02), mk(d, r, std=0.05), mk(d, std=0.02)]
    G = [torch.zeros(w.shape, dtype=torch.float32, device="cuda") for w in W]
    desc = L.K1Desc(M=M, L=0, d=d, r=r, rg=r, gate=L.GATE_IDS[a.gate], add_gate=0, dtype=L.BF16, impl=L.IMPL_AUTO, s=1.0, alpha=1.0,
                    kappa=1.0, p_drop=a.p_drop, seed=1234, seed_dev=None)
    w = L.K1Params(Wd=p_(W[0]), bd=p_(W[1]), Wu=p_(W[2]), bu=p_(W[3]), Gd=p_(W[4]), gbd=p_(W[5]), Gu=p_(W[6]), gbu=p_(W[7]))
    gr = L.K1Grads(dWd=p_(G[0]), dbd=p_(G[1]), dWu=p_(G[2]), dbu=p_(G[3]), dGd=p_(G[4]), dgbd=p_(G[5]), dGu=p_(G[6]), dgbu=p_(G[7]))
    dx1, dx2 = torch.empty_like(x1), torch.empty_like(x2)
    ws = torch.empty(L.lib.vlpet_k1_bwd_workspace_bytes(C.byref(desc)) + 256, dtype=torch.uint8, device="cuda")
    fused = L.lib.vlpet_k1_bwd_is_fused(C.byref(desc))

    def run():
        L.check(L.lib.vlpet_k1_bwd(C.byref(desc), p_(x1), p_(x2), p_(dout), C.byref(w), p_(dx1), p_(dx2), C.byref(gr), p_(ws),
                                   ws.numel(), st), "vlpet_k1_bwd")
    for parts, name in ((7, "all"), (1, "tile_kernel"), (4, "wgrad"), (2, "colsums")):
        L.lib.vlpet_debug_set_k1_bwd_parts(parts)
        run(); torch.cuda.synchronize()
        # replay from a CUDA graph: at small M the host side of the call (19 tensor maps, 3 launches) is longer than the kernels
        side = torch.cuda.Stream()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.stream(side):
            st = C.c_void_p(side.cuda_stream)
            run(); side.synchronize()
            with torch.cuda.graph(graph, stream=side):
                st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
                run()
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        ts = []
        for _ in range(a.iters):
            if a.flush:
                junk.zero_()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(); graph.replay(); e.record(); torch.cuda.synchronize()
            ts.append(s.elapsed_time(e) * 1e3)
        ts.sort()
        med = ts[len(ts) // 2]
        gbs = 5 * M * d * 2 / (med * 1e-6) / 1e9
        print(json.dumps({"M": M, "r": r, "part": name, "fused": int(fused), "us_median": round(med, 1), "us_best": round(ts[0], 1),
                          "algorithmic_GBps_if_whole_bwd": round(gbs, 1), "frac": round(gbs / peak, 3), "l2_flushed": bool(a.flush)}))
    L.lib.vlpet_debug_set_k1_bwd_parts(7)
