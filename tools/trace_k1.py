"""Phase timeline of the fused K1 forward (developer tool): where does a tile's time go?"""
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
    x1 = torch.randn(M, d, device="cuda", generator=g).to(bf); x2 = torch.randn(M, d, device="cuda", generator=g).to(bf)
    W = [(torch.randn(*s, device="cuda", generator=g) * 0.05).to(bf) for s in ((r, d), (r,), (d, r), (d,), (r, d), (r,), (d, r), (d,))]
    cfg = V.PetSiteConfig(gate="large")
    buf = torch.zeros(148 * 256, dtype=torch.int64, device="cuda")
    L.lib.vlpet_debug_set_k1_trace.argtypes = [C.c_void_p]
    for _ in range(3):
        with torch.no_grad():
            F_.GatedPETFn.apply(cfg, 0, 0, 1, x1, x2, *W)
    torch.cuda.synchronize()
    L.lib.vlpet_debug_set_k1_trace(C.c_void_p(buf.data_ptr()))
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    with torch.no_grad():
        F_.GatedPETFn.apply(cfg, 0, 0, 1, x1, x2, *W)
    e.record()
    torch.cuda.synchronize()
    L.lib.vlpet_debug_set_k1_trace(C.c_void_p(0))
    t = buf.view(148, 256).cpu().numpy()
    print(f"M={M}: kernel {s.elapsed_time(e)*1e3:.1f} us (first tile of CTA 0 and CTA 5, ns relative to tile start)")
    for cta in (0, 5):
        row = t[cta]
        if row[0] == 0: continue
        rel = (row - row[0])
        print(f"  cta {cta}: wait A/P {rel[1]} | epi1 done {rel[2]}")
        for c in range(12):
            a, b, cc, dd = rel[3 + 4 * c], rel[4 + 4 * c], rel[5 + 4 * c], rel[6 + 4 * c]
            print(f"    chunk {c:2d}: UT ready {a:7d}  tmem read {b - a:5d}  wait x {cc - b:6d}  math {dd - cc:6d}  -> {dd}")
    if M >= 128 * 148 * 2:
        r0 = t[0]
        z = r0[0]
        print("  cta 0 MMA thread: standalone phase A steps issued at", [int(v - z) for v in r0[100:112]])
        print("  cta 0 MMA thread, item 0: chunk_b issued at", [int(v - z) for v in r0[64:88:2]])
        print("  cta 0 MMA thread, item 0: step_a (next item) issued at", [int(v - z) for v in r0[65:88:2]])
        print("  cta 0 x manager: OUTRDY seen at", [int(v - z) for v in r0[128:152:2]])
        print("  cta 0 x manager: store read done at", [int(v - z) for v in r0[129:152:2]])
        import numpy as np
        t0 = t[:, 0][t[:, 0] > 0].min()
        ends = t[:, 52:64].astype(np.int64)
        print("  per work item (ns since the earliest CTA start): min / median / max end over CTAs, median duration")
        prev = (t[:, 0] - t0).astype(np.int64)
        for k in range(12):
            col = ends[:, k]
            ok = col > 0
            if not ok.any():
                break
            e = col[ok] - t0
            dur = e - prev[ok]
            print(f"    item {k}: ctas {ok.sum():3d}  end {e.min():7d} / {int(np.median(e)):7d} / {e.max():7d}   duration median {int(np.median(dur)):6d} max {dur.max():6d}")
            prev = np.where(ok, col - t0, prev)
for M in (128 * 16, 128 * 148, 96000):
    run(M)
