"""Phase timeline of the fused K1 backward activation-gradient kernel (developer tool)."""
import ctypes as C, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import _variant
_variant.use_trace_lib()   # the stamps exist only in the -DVLPET_TRACE build
import vlpet_b200 as V
import vlpet_b200.functional as F_
from vlpet_b200 import _lib as L

def run(M):
    d, r = 768, 96
    bf = torch.bfloat16
    g = torch.Generator(device="cuda").manual_seed(0)
    x1 = torch.randn(M, d, device="cuda", generator=g).to(bf).requires_grad_(); x2 = torch.randn(M, d, device="cuda", generator=g).to(bf).requires_grad_()
    dout = torch.randn(M, d, device="cuda", generator=g).to(bf)
    W = [(torch.randn(*s, device="cuda", generator=g) * 0.05).to(bf).float().requires_grad_() for s in ((r, d), (r,), (d, r), (d,), (r, d), (r,), (d, r), (d,))]
    cfg = V.PetSiteConfig(gate="large")
    buf = torch.zeros(148 * 128, dtype=torch.int64, device="cuda")
    L.lib.vlpet_debug_set_k1_bwd_trace.argtypes = [C.c_void_p]
    for _ in range(3):
        F_.GatedPETFn.apply(cfg, 0, 0, 1, x1, x2, *W).backward(dout)
    torch.cuda.synchronize()
    L.lib.vlpet_debug_set_k1_bwd_trace(C.c_void_p(buf.data_ptr()))
    F_.GatedPETFn.apply(cfg, 0, 0, 1, x1, x2, *W).backward(dout)
    torch.cuda.synchronize()
    L.lib.vlpet_debug_set_k1_bwd_trace(C.c_void_p(0))
    t = buf.view(148, 128).cpu().numpy()
    row = t[0]; rel = row - row[0]
    print(f"M={M}: CTA 0 first tile, ns from tile start: A/P ready {rel[1]}, epi1 done {rel[2]}")
    for c in range(12):
        if rel[3 + 3 * c] < 0: break
        a, b, cc = rel[3 + 3 * c], rel[4 + 3 * c], rel[5 + 3 * c]
        print(f"   p2 chunk {c:2d}: UT ready {a:7d}  ld+wait x {b - a:5d}  math {cc - b:6d} -> {cc}")
    print(f"   dz ready {rel[41]} (waited {rel[41]-rel[40]}), epi3 done {rel[42]}")
    for c in range(12):
        if rel[43 + 3 * c] < 0: break
        a, b, cc = rel[43 + 3 * c], rel[44 + 3 * c], rel[45 + 3 * c]
        print(f"   p3 chunk {c:2d}: ACC ready {a:7d}  ld+wait x {b - a:5d}  math {cc - b:6d} -> {cc}")
import sys as _s
for M in ([int(x) for x in _s.argv[1:]] or (128 * 100, 96000)):
    run(M)
