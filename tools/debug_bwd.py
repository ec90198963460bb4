import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vlpet_b200 as V
from oracle import pet_oracle as O
from tests.helpers import bf16_round, rel
from tests.test_gpu_parity import random_large_case, run_k1

def run(M, d, r, rg, add_gate=False, s=1.0):
    rng = np.random.default_rng(M + d + r + 1)
    x1, x2, dout, p = random_large_case(rng, M, d, r, rg)
    cfg = O.PetConfig(gate="large", add_gate=add_gate, s=s)
    out, dx1, dx2, gr = run_k1(V, x1, x2, dout, p, cfg, 1, torch.bfloat16, "auto", (1, M, d))
    x1r, x2r, dor = bf16_round(x1), bf16_round(x2), bf16_round(dout)
    pr = {k: bf16_round(v).reshape(np.shape(v)) for k, v in p.items()}
    _, c = O.gated_pet_fwd(x1r, x2r, pr, cfg, rnd=bf16_round)
    e1, e2, g = O.gated_pet_bwd(dor, pr, cfg, c, rnd=bf16_round)
    print(f"M={M} d={d} r={r} rg={rg}: dx1 {rel(dx1, e1):.3e} dx2 {rel(dx2, e2):.3e}", {k: float('%.2e' % rel(v, g[k].reshape(np.shape(v)))) for k, v in gr.items()})
    t1 = dx1 - dor; o1 = e1 - dor
    print("  term dp*Gd: rel", rel(t1, o1), " by row%8:", [float('%.2f' % rel(t1[i::8], o1[i::8])) for i in range(8)])
    print("   by col chunk:", [float('%.2f' % rel(t1[:, i*64:(i+1)*64], o1[:, i*64:(i+1)*64])) for i in range(d // 64)])
    G = c["G"]; dy1 = (cfg.s * dor * G) if not add_gate else cfg.s * dor
    t2 = dx2 - cfg.kappa * dy1; o2 = e2 - cfg.kappa * dy1
    print("  term da*Wd: rel", rel(t2, o2), " by row%8:", [float('%.2f' % rel(t2[i::8], o2[i::8])) for i in range(8)])
    _, cx = O.gated_pet_fwd(x1r, x2r, pr, cfg)
    x1e, x2e, _ = O.gated_pet_bwd(dor, pr, cfg, cx)
    print("  rnd-vs-exact oracle: max abs diff dx1 %.2e dx2 %.2e" % (np.abs(x1e - e1).max(), np.abs(x2e - e2).max()))
    for name, ours, ref in (("dx1", dx1, e1), ("dx2", dx2, e2), ("dx1-exact", dx1, x1e), ("dx2-exact", dx2, x2e)):
        err = np.abs(ours - bf16_round(ref).reshape(ref.shape))
        ulp = 2.0 ** (np.floor(np.log2(np.maximum(np.maximum(np.abs(ours), np.abs(ref)), 1e-30))) - 7)
        bad = np.argwhere((err > ulp * 1.0001) & (err > 5e-3 * np.sqrt(np.mean(ref * ref))))
        print(f"  {name}: {len(bad)} outliers; rows {np.unique(bad[:,0])[:20]} cols {np.unique(bad[:,1])[:20]}")
        for (i, j) in bad[:6]:
            print(f"     [{i},{j}] ours={ours[i,j]:.5f} ref={ref[i,j]:.5f} dout={dor[i,j]:.5f} G={c['G'][i,j]:.4f} y1={c['y1'][i,j]:.3f} max|da_row|={np.abs(c['a'][i]).max():.2f}")
    # ratio check
    print("   sample ours/ref dp*Gd:", (t1[-1, :4]), (o1[-1, :4]))

for a in [(1000, 768, 96, 96, False, 0.3)]:
    run(*a)
