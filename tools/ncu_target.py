"""Target for the ncu captures of round 2 (profiles/README.md): K1 forward + backward at the north-star shape (M = 96 000, d = 768,
r = rg = 96, bf16, large gate) and one middleX step on the row-wise path, through the autograd binding."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vlpet_b200 as V
import vlpet_b200.functional as F_

M, d, r, bf = 96000, 768, 96, torch.bfloat16
g = torch.Generator(device="cuda").manual_seed(0)
x1 = torch.randn(300, 320, d, device="cuda", generator=g).to(bf).requires_grad_()
x2 = (0.5 * torch.randn(300, 320, d, device="cuda", generator=g)).to(bf).requires_grad_()
dout = torch.randn(300, 320, d, device="cuda", generator=g).to(bf)
mk = lambda *sh, std: (torch.randn(*sh, device="cuda", generator=g) * std).to(bf).float().requires_grad_()  # noqa: E731
W = [mk(r, d, std=0.05), mk(r, std=0.02), mk(d, r, std=0.05), mk(d, std=0.02), mk(r, d, std=0.05), mk(r, std=0.02), mk(d, r, std=0.05), mk(d, std=0.02)]
for _ in range(3):
    F_.GatedPETFn.apply(V.PetSiteConfig(gate="large"), 0, 320, 1, x1, x2, *W).backward(dout)
gw, gb = mk(d, std=0.05), mk(1, std=0.02)
for _ in range(2):
    F_.GatedPETFn.apply(V.PetSiteConfig(gate="middle_x"), 0, 320, 1, x1, x2, *W[:4], gw, gb).backward(dout)
torch.cuda.synchronize()
print("done")
