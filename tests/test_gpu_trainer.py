"""GPU: the PET-only trainer end to end on a small model -- eager and CUDA-graph replay must agree, the loss must
go down, and dropout masks must differ between graph replays (device-side seed, VlpetK1Desc.seed_dev)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _run(H, cls, steps=6, p=0.0, lr=1e-2):
    torch.manual_seed(0)
    cfg = H.tiny_test_config(d_model=128, adapter_down_dim=32, adapter_gating_down_dim=32,
                             decoder_enc_attn_value_parallel_adapter_down_dim=32, assume_no_padding=True, dropout=p,
                             attention_dropout=p, activation_dropout=p)
    model = H.VLBart(cfg).train()
    tr = cls(model, cfg, "cuda", lr=lr, total_steps=100)
    tr.set_step(10)
    batches = [{k: (v.cuda() if torch.is_tensor(v) else v)
                for k, v in H.make_task_batch(t, 16, feat_dim=128, vocab_hi=300, seed=3).items()} for t in ("vqa", "nlvr")]
    return [float(tr.train_step(batches[i % 2]).item()) for i in range(steps)], tr


def test_graphed_trainer_tracks_eager_trainer():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import vlpet_b200 as V
    import vlpet_b200.host as H
    n0 = V.launch_count()
    eager, _ = _run(H, H.PetTrainer)
    assert V.launch_count() > n0
    graphed, tr = _run(H, H.GraphedPetTrainer)
    assert all(abs(a - b) < 5e-3 * abs(a) for a, b in zip(eager, graphed)), (eager, graphed)
    assert eager[4] < eager[0] and eager[5] < eager[1], "loss does not decrease"
    assert tr.step_idx == 16 and abs(float(tr._t.item()) - 16.0) < 1e-6
    V.functional.set_device_seed(None)


def test_graph_replays_draw_fresh_dropout_masks():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import vlpet_b200 as V
    import vlpet_b200.functional as F_
    M, d, r = 256, 256, 32
    g = torch.Generator(device="cuda").manual_seed(1)
    bf = torch.bfloat16
    x1 = torch.randn(M, d, device="cuda", generator=g).to(bf)
    x2 = torch.randn(M, d, device="cuda", generator=g).to(bf)
    W = [(torch.randn(*s, device="cuda", generator=g) * 0.05).to(bf) for s in ((r, d), (r,), (d, r), (d,), (r, d), (r,), (d, r), (d,))]
    seed = torch.zeros(1, dtype=torch.int64, device="cuda")
    F_.set_device_seed(seed)
    try:
        cfg = V.PetSiteConfig(gate="large", p_drop=0.5)
        outs = []
        for i in range(3):
            with torch.no_grad():
                outs.append(F_.GatedPETFn.apply(cfg, 77, 0, 1, x1, x2, *W).clone())
            if i == 0:
                seed.add_(1000003)          # what the graphed trainer does at the head of every replay
        assert not torch.equal(outs[0], outs[1])      # new device seed -> new mask
        assert torch.equal(outs[1], outs[2])          # same device seed -> same mask (forward/backward agreement)
    finally:
        F_.set_device_seed(None)


def test_fused_qkv_projection_equals_three_linears():
    """Frozen self-attention: the one-GEMM q/k/v projection (host.vlbart.BartAttention._fused_qkv) must give the same
    output and input gradient as the reference's three Linears (my_transformers/modeling_bart.py:143-280)."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import vlpet_b200.host as H
    from vlpet_b200.host.vlbart import BartAttention
    torch.manual_seed(0)
    cfg = H.tiny_test_config(d_model=128, adapter_down_dim=32, adapter_gating_down_dim=32,
                             decoder_enc_attn_value_parallel_adapter_down_dim=32, assume_no_padding=True, dropout=0.0,
                             attention_dropout=0.0, activation_dropout=0.0)
    att = BartAttention(cfg, 4, is_decoder=False, value_adapter=False).cuda().to(torch.bfloat16).eval()
    for p_ in att.parameters():
        p_.requires_grad_(False)
    x = torch.randn(5, 37, 128, device="cuda").to(torch.bfloat16)
    dy = torch.randn(5, 37, 128, device="cuda").to(torch.bfloat16)
    outs = []
    for fused in (True, False):
        xi = x.clone().requires_grad_()
        if not fused:
            att._fused_qkv = lambda: None
        assert (att._fused_qkv() is not None) == fused
        y = att(xi)
        y.backward(dy)
        outs.append((y.float(), xi.grad.float()))
    for a, b in zip(outs[0], outs[1]):
        assert (a - b).norm() <= 1e-2 * b.norm()        # two bf16 GEMM schedules: a few ulps apart
