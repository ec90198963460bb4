"""Host models against the LIVE reference for the flag sets that have no committed host-model fixture: BART with the middleX
and middleY gates (VL-PET-middleX.sh / VL-PET-middleY.sh; plus the large gate as a check of the fixtures' generator) and T5 with
the small, middleX and middleY gates.  The reference's own ``VLBart`` / ``VLT5MultiTask`` are built by the golden generators'
``build`` from /root/reference (or the sources staged under baseline/_ref) -- so these run in the build container and skip on
a box without the reference.  CPU, fp64, the PET
sites through the eager restatement: checks the host plumbing and the per-gate flag routing, loss and every trainable gradient."""
import os
import sys

import pytest
import torch

from tests.conftest import ROOT
from tests.helpers import GOLDEN, rel


def _reference_src():
    for p in (os.environ.get("VLPET_REFERENCE_SRC"), "/root/reference/src", os.path.join(ROOT, "baseline", "_ref", "src")):
        if p and os.path.isdir(p):
            return p
    return None


@pytest.fixture(scope="module")
def MV():
    src = _reference_src()
    if src is None:
        pytest.skip("reference sources not available on this box")
    os.environ["VLPET_REFERENCE_SRC"] = src
    sys.dont_write_bytecode = True
    if GOLDEN not in sys.path:
        sys.path.insert(0, GOLDEN)
    try:
        import make_golden_vlbart as MV_
    except Exception as e:
        pytest.skip(f"reference not importable here: {e!r}")
    yield MV_
    MV_.R.remove_shims()                         # stock transformers models built by later tests tie their weights again


@pytest.mark.parametrize("gate", ["middle_x", "middle_y", "large"])
def test_host_vlbart_matches_live_reference(MV, gate):
    import vlpet_b200.host as H
    from oracle.eager_ref import use_eager_pet
    ref, _ = MV.build(gate)
    flag = {"large": "use_encoder_adapter_gating_large_x_lowrank", "middle_x": "use_encoder_adapter_gating_middle_xy_add",
            "middle_y": "use_encoder_adapter_gating_middle_ia3_add"}[gate]
    flags = dict(use_encoder_adapter_gating_large_x_lowrank=False, dropout=0.0, attention_dropout=0.0, activation_dropout=0.0)
    flags[flag] = True
    cfg = H.tiny_test_config(**flags)
    model = use_eager_pet(H.VLBart(cfg).double().eval())
    sd = ref.state_dict()
    assert set(model.state_dict()) == set(sd), sorted(set(model.state_dict()) ^ set(sd))[:6]
    model.load_state_dict(sd, strict=True)
    names = H.trainable_names(model, cfg)
    g = torch.Generator().manual_seed(5)
    B, Lt, T = 3, 6, 4
    ids = torch.randint(3, 300, (B, Lt), generator=g)
    ids[1, 4:] = cfg.pad_token_id
    tgt = torch.randint(3, 300, (B, T), generator=g)
    tgt[2, 2:] = -100
    feats = torch.randn(B, 49, MV.FEAT, generator=g, dtype=torch.float64)
    boxes = torch.zeros(B, 49, 4, dtype=torch.float64)
    scores = torch.rand(B, generator=g, dtype=torch.float64)
    ref.zero_grad()
    o = ref(input_ids=ids, vis_inputs=(feats, boxes), labels=tgt, return_dict=True, task="vqa")
    mask = (tgt != -100).double()
    want = ((o["loss"].view(B, T) * mask).sum(1) / mask.sum(1).clamp(min=1) * scores).mean()      # vqa_model.py:211-227
    want.backward()
    loss = model.train_step({"task": "vqa", "input_ids": ids, "target_ids": tgt, "vis_feats": feats, "boxes": boxes,
                             "scores": scores})["loss"]
    loss.backward()
    assert abs(loss.item() - want.item()) < 1e-10
    rp, hp = dict(ref.named_parameters()), dict(model.named_parameters())
    gate_names = [n for n in names if "gating" in n]
    assert gate_names, "the gate's parameters are not trainable"
    for n in names:
        assert rp[n].grad is not None, n
        assert rel(hp[n].grad.numpy(), rp[n].grad.numpy()) < 1e-8, (gate, n)


@pytest.fixture(scope="module")
def MT():
    if _reference_src() is None:
        pytest.skip("reference sources not available on this box")
    sys.dont_write_bytecode = True
    if GOLDEN not in sys.path:
        sys.path.insert(0, GOLDEN)
    try:
        import make_golden_vlt5 as MT_
    except Exception as e:
        pytest.skip(f"reference not importable here: {e!r}")
    yield MT_
    MT_.RA.remove_shims()


@pytest.mark.parametrize("gate", ["small", "middle_x", "middle_y"])
def test_host_vlt5_matches_live_reference(MT, gate):
    """T5 twin (T5-VL-PET-small / -middleX / -middleY flag sets; the large gate has the committed fixture): the reference's
    ``VLT5MultiTask.train_step`` against ``host.VLT5.train_step`` on the same VQA batch, same state_dict."""
    import vlpet_b200.host as H
    from oracle.eager_ref import use_eager_pet
    ref, _ = MT.build(gate)
    flag = {"small": "use_encoder_adapter_gating_small_xy_cat", "middle_x": "use_encoder_adapter_gating_middle_xy_add",
            "middle_y": "use_encoder_adapter_gating_middle_ia3_add"}[gate]
    flags = dict(use_encoder_adapter_gating_large_x_lowrank=False, dropout_rate=0.0, dropout=0.0)
    flags[flag] = True
    cfg = H.tiny_t5_test_config(**flags)
    model = use_eager_pet(H.VLT5(cfg).double().eval())
    sd = ref.state_dict()
    assert set(model.state_dict()) == set(sd), sorted(set(model.state_dict()) ^ set(sd))[:6]
    model.load_state_dict(sd, strict=True)
    names = H.trainable_names(model, cfg)
    g = torch.Generator().manual_seed(6)
    B, Lt, T = 3, 6, 4
    ids = torch.randint(3, 300, (B, Lt), generator=g)
    ids[1, 4:] = cfg.pad_token_id
    tgt = torch.randint(3, 300, (B, T), generator=g)
    tgt[2, 2:] = -100
    batch = {"task": "vqa", "input_ids": ids, "target_ids": tgt, "vis_feats": torch.randn(B, 49, MT.FEAT, generator=g, dtype=torch.float64),
             "boxes": torch.zeros(B, 49, 4, dtype=torch.float64), "scores": torch.rand(B, generator=g, dtype=torch.float64)}
    ref.zero_grad()
    want = ref.train_step(dict(batch))["loss"]
    want.backward()
    loss = model.train_step(dict(batch))["loss"]
    loss.backward()
    assert abs(loss.item() - want.item()) < 1e-7
    rp, hp = dict(ref.named_parameters()), dict(model.named_parameters())
    assert [n for n in names if "gating" in n], "the gate's parameters are not trainable"
    for n in names:
        assert rp[n].grad is not None, n
        assert rel(hp[n].grad.numpy(), rp[n].grad.numpy()) < 2e-6, (gate, n)
