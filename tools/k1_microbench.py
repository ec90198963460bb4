"""Kernel micro-benchmark for K1: B=300, L=320, d=768, r=rg=96, bf16.  CUDA events on the launch
stream, L2 flushed between iterations (also reports the warm number).  Prints one JSON line per measurement."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vlpet_b200 as V  # noqa: E402
import vlpet_b200.functional as F_  # noqa: E402


def time_it(fn, iters, flush):
    fn()
    torch.cuda.synchronize()
    ts = []
    junk = torch.empty(256 << 20, dtype=torch.uint8, device="cuda") if flush else None
    for _ in range(iters):
        if flush:
            junk.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--B", type=int, default=300)
    ap.add_argument("--L", type=int, default=320)
    ap.add_argument("--d", type=int, default=768)
    ap.add_argument("--r", type=int, default=96)
    ap.add_argument("--iters", type=int, default=30)
    ap.add_argument("--impl", default="auto")
    ap.add_argument("--bwd", action="store_true")
    a = ap.parse_args()
    M, d, r = a.B * a.L, a.d, a.r
    bf = torch.bfloat16
    g = torch.Generator(device="cuda").manual_seed(0)
    x1 = torch.randn(a.B, a.L, d, device="cuda", generator=g).to(bf)
    x2 = (0.5 * torch.randn(a.B, a.L, d, device="cuda", generator=g)).to(bf)
    dout = torch.randn(a.B, a.L, d, device="cuda", generator=g).to(bf)
    mk = lambda *sh, std: (torch.randn(*sh, device="cuda", generator=g) * std).to(bf)  # noqa: E731
    W = [mk(r, d, std=0.05), mk(r, std=0.02), mk(d, r, std=0.05), mk(d, std=0.02),
         mk(r, d, std=0.05), mk(r, std=0.02), mk(d, r, std=0.05), mk(d, std=0.02)]
    cfg = V.PetSiteConfig(gate="large", impl=a.impl)
    peak = 6454.0
    try:
        peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass

    def fwd():
        with torch.no_grad():
            return F_.GatedPETFn.apply(cfg, 0, a.L, 1, x1, x2, *W)

    for flush in (True, False):
        med, best = time_it(fwd, a.iters, flush)
        gbs = 3 * M * d * 2 / (med * 1e-3) / 1e9
        print(json.dumps({"kernel": "k1_fwd", "impl": a.impl, "fused": V.fwd_is_fused(M, d, r, r), "M": M, "d": d, "r": r,
                          "l2_flushed": flush, "ms_median": round(med, 4), "ms_best": round(best, 4),
                          "algorithmic_GBps": round(gbs, 1), "frac_of_measured_hbm": round(gbs / peak, 3)}))
    if a.bwd:
        Wg = [w.clone().requires_grad_() for w in W]
        x1g, x2g = x1.clone().requires_grad_(), x2.clone().requires_grad_()

        def fb():
            out = F_.GatedPETFn.apply(cfg, 0, a.L, 1, x1g, x2g, *Wg)
            out.backward(dout)

        med, best = time_it(fb, max(3, a.iters // 3), True)
        gbs = 8 * M * d * 2 / (med * 1e-3) / 1e9
        print(json.dumps({"kernel": "k1_fwd+bwd", "impl": a.impl, "M": M, "ms_median": round(med, 4),
                          "algorithmic_GBps": round(gbs, 1), "frac_of_measured_hbm": round(gbs / peak, 3)}))


if __name__ == "__main__":
    main()
