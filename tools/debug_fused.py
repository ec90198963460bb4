import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vlpet_b200 as V
from oracle import pet_oracle as O
from tests.helpers import bf16_round
from tests.test_gpu_parity import random_large_case, dev

def run(M, d, r, rg, add_gate=False, s=1.0):
    rng = np.random.default_rng(M + d + r)
    x1, x2, _, p = random_large_case(rng, M, d, r, rg)
    bf = torch.bfloat16
    scfg = V.PetSiteConfig(gate="large", add_gate=add_gate, s=s, impl="fused")
    with torch.no_grad():
        out = V.gated_pet(dev(x1, bf), dev(x2, bf), [dev(p["Wd"], bf)], [dev(p["bd"], bf)], dev(p["Wu"], bf),
                          dev(p["bu"], bf), [dev(p[k], bf) for k in ("Gd", "gbd", "Gu", "gbu")], scfg)
    out = out.to(torch.float64).cpu().numpy()
    pr = {k: bf16_round(v).reshape(np.shape(v)) for k, v in p.items()}
    ref, c = O.gated_pet_fwd(bf16_round(x1), bf16_round(x2), pr, O.PetConfig(gate="large", add_gate=add_gate, s=s))
    err = np.abs(out - ref)
    bad = np.argwhere(err > 5e-3 + 4e-3 * np.abs(ref))
    print(f"M={M} d={d} r={r} rg={rg}: {len(bad)} bad elements of {out.size}; max err {err.max():.4f}")
    if len(bad):
        rows, cols = bad[:, 0], bad[:, 1]
        print("  rows:", np.unique(rows)[:40], " row%128:", np.unique(rows % 128)[:40])
        print("  cols:", np.unique(cols)[:60])
        for (i, j) in bad[:12]:
            print(f"   [{i},{j}] ours={out[i,j]:.5f} ref={ref[i,j]:.5f} x1={x1[i,j]:.4f} y1={c['y1'][i,j]:.4f} G={c['G'][i,j]:.4f} maxabs_a={np.abs(c['a'][i]).max():.2f} maxabs_p={np.abs(c['p'][i]).max():.2f}")

for args in [(129, 768, 96, 96), (260, 768, 128, 128), (1000, 768, 96, 96), (777, 768, 96, 96, True), (148*128*2+77, 768, 96, 96)]:
    run(*args)
