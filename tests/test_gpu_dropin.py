"""The drop-in claim (SURVEY section 8 b / f-2), on the GPU: the reference's OWN model -- built by the reference's code from
the sources staged under baseline/_ref/src -- runs forward + backward unpatched (stock PyTorch eager) and then again after
``vlpet_b200.patch_reference_model`` routed its PET sites through libvlpet.so; loss and every trainable gradient must agree.
Also the LAST.pth round trip (trainer_base.py:764-781) through the patched model with the aliased single-adapter keys
(adapters/adapter_controller.py:49-58).  Skipped only where the staged reference is absent."""
import io
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from baseline import ref_arm as RA  # noqa: E402

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(RA.available() is None, reason="reference sources not staged")]
TASKS = ["vqa", "gqa", "nlvr", "caption"]


def _trained_like_(model, seed=1):
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for n, p in model.named_parameters():
            if any(t in n for t in ("adapter", "gating")):
                p.copy_(torch.randn(p.shape, generator=g) * (0.02 if p.dim() == 1 else 0.05))


def _run(model, params, batches, autocast=None):
    losses, grads = [], []
    for b in batches:
        model.zero_grad(set_to_none=True)
        if autocast is not None:
            with torch.autocast("cuda", dtype=autocast):
                loss = model.train_step(b)["loss"]
        else:
            loss = model.train_step(b)["loss"]
        loss.backward()
        losses.append(float(loss))
        grads.append([p.grad.detach().double().clone() for p in params])
    return losses, grads


def _rel(a, b):
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


@pytest.mark.parametrize("kind", ["bart", "t5"])
def test_patched_reference_model_matches_unpatched_fp32(kind):
    import vlpet_b200 as V
    model, _ = RA.build_model(kind, r=96, dropout=0.0, layers=2)
    _trained_like_(model)
    model = model.cuda().train()
    params, _ = RA.prepare_training(model)
    names = [n for n, p in model.named_parameters() if p.requires_grad]
    cycle, _ = RA.reference_batches(6, TASKS, vocab_hi=32000)
    l0, g0 = _run(model, params, cycle)
    keys0 = list(model.state_dict().keys())
    counts = V.patch_reference_model(model)
    assert counts["vpa"] == 2 and counts["visual_embedding"] == 1
    assert counts["bart_encoder_layer"] == 2 if kind == "bart" else (counts["t5_self_attention"] == 2 and counts["t5_ff"] == 2)
    assert list(model.state_dict().keys()) == keys0
    n0 = V.launch_count()
    l1, g1 = _run(model, params, cycle)
    assert V.launch_count() > n0, "the patched model did not launch a vlpet kernel"
    for a, b in zip(l0, l1):
        assert abs(a - b) <= 2e-5 * abs(a), (a, b)
    for t in range(len(cycle)):
        scale = max(float(g.norm()) for g in g0[t])          # gradients ~1e-3 of the largest one are cancelling sums: fp32 noise
        for n, ga, gb in zip(names, g0[t], g1[t]):
            if ga.norm() == 0:
                assert gb.norm() == 0, n
                continue
            assert _rel(gb, ga) < 2e-4 or float((gb - ga).norm()) < 1e-5 * scale, (TASKS[t], n, _rel(gb, ga))


def test_patched_reference_bf16_is_no_worse_than_reference_bf16():
    """bf16: the fused tcgen05 path against the reference's own bf16 autocast run, both measured against the reference in
    fp32 on the same inputs."""
    import vlpet_b200 as V
    model, _ = RA.build_model("bart", r=96, dropout=0.0, layers=2)
    _trained_like_(model)
    model = model.cuda().train()
    params, _ = RA.prepare_training(model)
    cycle, _ = RA.reference_batches(12, TASKS[:2])
    l32, g32 = _run(model, params, cycle)
    lref, gref = _run(model, params, cycle, autocast=torch.bfloat16)
    V.patch_reference_model(model)
    lours, gours = _run(model, params, cycle, autocast=torch.bfloat16)
    for t in range(len(cycle)):
        assert abs(lours[t] - l32[t]) <= max(2.0 * abs(lref[t] - l32[t]), 2e-3 * abs(l32[t])), (lours[t], lref[t], l32[t])
        num = sum(float((a - b).norm() ** 2) for a, b in zip(gours[t], g32[t])) ** 0.5
        den = sum(float((a - b).norm() ** 2) for a, b in zip(gref[t], g32[t])) ** 0.5
        assert num <= 1.5 * den, ("gradient error ours vs reference-bf16", num, den)


def test_last_pth_round_trip_through_patched_model():
    import vlpet_b200 as V
    model, _ = RA.build_model("bart", r=96, dropout=0.0, layers=2)
    _trained_like_(model)
    model = model.cuda().eval()
    V.patch_reference_model(model)
    sd = model.state_dict()
    stem = "model.decoder.layers.0.encoder_attn.attn_value_parallel_adapter.adapters."
    for task in TASKS:     # use_single_adapter: one Adapter object under every task key
        assert stem + task + ".down_sampler.weight" in sd
    assert sd[stem + "vqa.down_sampler.weight"].data_ptr() == sd[stem + "caption.down_sampler.weight"].data_ptr()
    buf = io.BytesIO()
    torch.save(sd, buf)                      # trainer_base.py:764-771 save("LAST")
    buf.seek(0)
    fresh, _ = RA.build_model("bart", r=96, dropout=0.0, layers=2, seed=123)
    fresh = fresh.cuda().eval()
    V.patch_reference_model(fresh)
    missing = fresh.load_state_dict(torch.load(buf), strict=False)   # trainer_base.py:734-740
    assert not missing.missing_keys and not missing.unexpected_keys
    cycle, _ = RA.reference_batches(4, ["vqa"])
    with torch.no_grad():
        a = float(model.train_step(cycle[0])["loss"])
        b = float(fresh.train_step(cycle[0])["loss"])
    assert a == b
