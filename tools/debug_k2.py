import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vlpet_b200 as V
from oracle import pet_oracle as O
from tests.helpers import bf16_round, rel
def dev(a, dt): return torch.tensor(np.asarray(a), dtype=dt, device="cuda")
def run(M, d, r, sf):
    rng = np.random.default_rng(M + d + r + 7)
    kv, y, dout = rng.standard_normal((M, d)), 0.5 * rng.standard_normal((M, d)), rng.standard_normal((M, d))
    p = {"Wd": rng.standard_normal((r, d)) * 0.05, "bd": rng.standard_normal(r) * 0.02,
         "Wu": rng.standard_normal((d, r)) * 0.05, "bu": rng.standard_normal(d) * 0.02}
    bf = torch.bfloat16
    P = {k: dev(v, bf) for k, v in p.items()}
    with torch.no_grad():
        out = V.vpa(dev(kv, bf), dev(y, bf), P["Wd"], P["bd"], P["Wu"], P["bu"], sf).double().cpu().numpy()
        outg = V.vpa(dev(kv, bf), dev(y, bf), P["Wd"], P["bd"], P["Wu"], P["bu"], sf, impl="generic").double().cpu().numpy()
    kvr, yr = bf16_round(kv), bf16_round(y)
    pr = {k: bf16_round(v).reshape(np.shape(v)) for k, v in p.items()}
    ref, c = O.vpa_fwd(kvr, yr, pr, sf)
    err = np.abs(out - ref); errg = np.abs(outg - ref)
    print(f"M={M} r={r}: fused rel {rel(out, ref):.3e} max {err.max():.4f}; generic rel {rel(outg, ref):.3e} max {errg.max():.4f}; max|a|={np.abs(c['a']).max():.2f}")
    bad = np.argwhere(err > 3e-3 + 4e-3 * np.abs(ref))
    print("  bad:", len(bad), "cols", np.unique(bad[:, 1])[:30], "rows", np.unique(bad[:, 0])[:10])
    for (i, j) in bad[:5]:
        print(f"   [{i},{j}] ours={out[i,j]:.5f} gen={outg[i,j]:.5f} ref={ref[i,j]:.5f} y={yr[i,j]:.4f}")
for a in [(1, 768, 96, 1.0), (200, 768, 96, 1.0), (333, 768, 48, 1.0)]:
    run(*a)
