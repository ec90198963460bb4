"""Developer tool: short-sequence attention (vlpet_attn_*) against torch SDPA at the workload's shapes, CUDA-event times."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
import vlpet_b200.functional as F_

def bench(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / n * 1e3

for (B, H, Lq, Lk, causal) in ((300, 12, 56, 56, False), (500, 12, 56, 56, False), (100, 12, 92, 92, False), (250, 12, 40, 40, True), (300, 12, 5, 56, False)):
    p = 0.1
    q = torch.randn(B, Lq, H * 64, device="cuda").bfloat16().requires_grad_()
    k = torch.randn(B, Lk, H * 64, device="cuda").bfloat16().requires_grad_()
    v = torch.randn(B, Lk, H * 64, device="cuda").bfloat16().requires_grad_()
    do = torch.randn(B, Lq, H * 64, device="cuda").bfloat16()
    def ours_f():
        return F_.short_attention(q, k, v, H, causal, p, True)
    def ours_fb():
        o = ours_f(); torch.autograd.grad(o, (q, k, v), do)
    hv = lambda t, L: t.view(B, L, H, 64).transpose(1, 2)
    def sdpa_f():
        return F.scaled_dot_product_attention(hv(q, Lq), hv(k, Lk), hv(v, Lk), dropout_p=p, is_causal=causal)
    def sdpa_fb():
        o = sdpa_f(); torch.autograd.grad(o, (q, k, v), hv(do, Lq))
    print(f"B={B} H={H} Lq={Lq} Lk={Lk} causal={causal}: ours fwd {bench(ours_f):6.1f} us  fwd+bwd {bench(ours_fb):6.1f} us | "
          f"sdpa fwd {bench(sdpa_f):6.1f} us  fwd+bwd {bench(sdpa_fb):6.1f} us", flush=True)
