"""CPU, build container only (skipped where /root/reference is absent): the drop-in entry ``patch_reference_model`` finds
every VL-PET site of the reference's OWN VLBart (built through the import shims of tests/golden/ref_import.py) and leaves
its parameters / state_dict untouched.  The patched forwards need CUDA, so only the wiring is checked here; numerical
parity of the same forwards is covered by the golden-vector tests."""
import os
import sys

import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import ref_import as R  # noqa: E402

pytestmark = pytest.mark.skipif(not R.available(), reason="reference sources not present (GPU box)")


def test_patch_reference_vlbart_wiring():
    import make_golden_vlbart as MG
    import vlpet_b200 as V
    model, config = MG.build("large")
    before = {k: v.data_ptr() for k, v in model.state_dict().items()}
    names_before = [n for n, _ in model.named_parameters()]
    counts = V.patch_reference_model(model)
    assert counts["bart_encoder_layer"] == config.encoder_layers
    assert counts["vpa"] == config.decoder_layers
    assert counts["visual_embedding"] == 1
    after = {k: v.data_ptr() for k, v in model.state_dict().items()}
    assert before == after and names_before == [n for n, _ in model.named_parameters()]
    layer = model.model.encoder.layers[0]
    assert layer._vlpet_site_cfg.gate == "large" and layer._vlpet_site_cfg.p_drop == 0.0
    # the patched forward refuses to run without CUDA (no CPU fallback), loudly
    x = torch.zeros(1, 4, config.d_model, dtype=torch.float64)
    with pytest.raises(RuntimeError, match="CUDA tensors required"):
        V.encoder_pet(layer, "attn", x, x)


def test_site_config_follows_reference_flags():
    import make_golden as G
    import vlpet_b200 as V
    cfg = G.make_config("bart", 64, 16, 4, 16, "small", add_gate=True, s=0.3)
    sc = V.site_config(cfg, is_t5=False)
    assert (sc.gate, sc.add_gate, abs(sc.s - 0.3) < 1e-12, sc.alpha, sc.kappa) == ("small", True, True, 1.0, 1.0)
    cfg_t5 = G.make_config("t5", 64, 16, 4, 16, "large", s=0.3, alpha=0.5)
    st = V.site_config(cfg_t5, is_t5=True)
    assert st.gate == "large" and abs(st.s - 0.3) < 1e-12 and abs(st.alpha - 0.5) < 1e-12 and not st.add_gate
