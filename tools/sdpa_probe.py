"""Which SDPA backend is fastest at the encoder shapes of the bench workload?  (host-side choice, stock PyTorch)"""
import torch
from torch.nn.attention import sdpa_kernel, SDPBackend
import torch.nn.functional as F
def bench(B, H, L, S, D, backend, causal=False, p=0.1):
    q = torch.randn(B, H, L, D, device="cuda", dtype=torch.bfloat16, requires_grad=True)
    k = torch.randn(B, H, S, D, device="cuda", dtype=torch.bfloat16, requires_grad=True)
    v = torch.randn(B, H, S, D, device="cuda", dtype=torch.bfloat16, requires_grad=True)
    try:
        with sdpa_kernel([backend]):
            for _ in range(3):
                o = F.scaled_dot_product_attention(q, k, v, dropout_p=p, is_causal=causal); o.sum().backward()
            torch.cuda.synchronize()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            for _ in range(10):
                o = F.scaled_dot_product_attention(q, k, v, dropout_p=p, is_causal=causal); o.sum().backward()
            e.record(); torch.cuda.synchronize()
            return s.elapsed_time(e) / 10
    except Exception as ex:
        return str(ex)[:60]
for name, shape in [("enc gqa", (500, 12, 56, 56, 64, False)), ("enc vqa", (300, 12, 56, 56, 64, False)), ("dec self cap", (250, 12, 40, 40, 64, True)), ("dec cross gqa", (500, 12, 5, 56, 64, False))]:
    B, H, L, S, D, c = shape
    for be in (SDPBackend.CUDNN_ATTENTION, SDPBackend.FLASH_ATTENTION, SDPBackend.EFFICIENT_ATTENTION, SDPBackend.MATH):
        print(name, be.name, bench(B, H, L, S, D, be, c))
