"""CPU: pin the oracle (oracle/pet_oracle.py) against golden vectors produced by the reference's own layer
classes (tests/golden/make_golden.py).  fp64, tolerance 1e-12 relative (Frobenius)."""
import os

import numpy as np
import pytest

from oracle import pet_oracle as O
from tests.helpers import golden_files, k1_case, load, rel

TOL = 1e-12


@pytest.mark.parametrize("path", golden_files("k1_"), ids=os.path.basename)
def test_k1_matches_reference(path):
    g = load(path)
    x1, x2, dout, p, cfg = k1_case(g)
    out, cache = O.gated_pet_fwd(x1, x2, p, cfg)
    assert rel(out, g["out"].reshape(out.shape)) < TOL
    dx1, dx2, gr = O.gated_pet_bwd(dout, p, cfg, cache)
    assert rel(dx1, g["dx1"].reshape(dx1.shape)) < TOL
    assert rel(dx2, g["dx2"].reshape(dx2.shape)) < TOL
    assert set(gr) == set(p)
    for k, v in gr.items():
        assert rel(v, g["d" + k].reshape(np.shape(v))) < TOL, k


def test_k1_covers_every_granularity():
    kinds = {str(load(p)["meta_gate"]) for p in golden_files("k1_")}
    assert kinds == {"large", "middle_x", "middle_y", "small"}


@pytest.mark.parametrize("path", golden_files("k2_"), ids=os.path.basename)
def test_k2_matches_reference(path):
    g = load(path)
    d = int(g["meta_d"])
    p = {k: g[k] for k in ("Wd", "bd", "Wu", "bu")}
    sf = float(g["meta_sf"])
    out, c = O.vpa_fwd(g["kv"].reshape(-1, d), g["y"].reshape(-1, d), p, sf)
    assert rel(out, g["out"].reshape(out.shape)) < TOL
    dkv, dy, gr = O.vpa_bwd(g["dout"].reshape(-1, d), p, c, sf)
    assert rel(dkv, g["dkv"].reshape(dkv.shape)) < TOL
    assert rel(dy, g["dy"].reshape(dy.shape)) < TOL
    for k, v in gr.items():
        assert rel(v, g["d" + k]) < TOL, k


def test_k2_single_adapter_aliasing_contract():
    """use_single_adapter registers ONE Adapter under every task key (adapter_controller.py:49-58):
    state_dict lists all tasks, named_parameters only the first."""
    g = load(golden_files("k2_vpa_d64")[0])
    keys = list(g["meta_state_keys"])
    assert len(keys) == 16 and all(k.startswith("adapters.") for k in keys)
    assert list(g["meta_param_names"]) == ["adapters.vqa.down_sampler.weight", "adapters.vqa.down_sampler.bias",
                                           "adapters.vqa.up_sampler.weight", "adapters.vqa.up_sampler.bias"]


@pytest.mark.parametrize("path", golden_files("k3_"), ids=os.path.basename)
def test_k3_matches_reference(path):
    g = load(path)
    rms = str(g["meta_kind"]) == "t5"
    keys = ["Wf", "bf", "ln_f_w", "Wp", "bp", "ln_p_w", "E_img", "E_obj"] + ([] if rms else ["ln_f_b", "ln_p_b"])
    p = {k: g[k] for k in keys}
    out, c = O.visproj_fwd(g["feats"], g["pos"], p, g.get("img_ids"), g.get("obj_ids"), rms=rms, eps=float(g["meta_eps"]))
    # T5LayerNorm computes its variance in float32 even for float64 inputs (my_transformers/modeling_t5.py:246),
    # so the reference's own fp64 run carries ~1e-7 rounding there; the oracle stays in the input dtype.
    tol = 1e-6 if rms else TOL
    assert rel(out, g["out"]) < tol
    dfeats, gr = O.visproj_bwd(g["dout"], p, c, rms=rms)
    assert rel(dfeats, g["dfeats"]) < tol
    for k, v in gr.items():
        assert rel(v, g["d" + k]) < tol, k


def test_gelu_new_grad_finite_difference():
    t = np.linspace(-6, 6, 1001)
    h = 1e-6
    fd = (O.gelu_new(t + h) - O.gelu_new(t - h)) / (2 * h)
    assert np.max(np.abs(fd - O.gelu_new_grad(t))) < 1e-8


def test_multihead_down_is_one_linear():
    """cat_i(Linear_i(x)) == Linear(rowcat W_i)."""
    rng = np.random.default_rng(0)
    x = rng.standard_normal((5, 32))
    Ws = [rng.standard_normal((4, 32)) for _ in range(3)]
    bs = [rng.standard_normal(4) for _ in range(3)]
    W, b = O.stack_heads(Ws, bs)
    ref = np.concatenate([x @ w.T + bb for w, bb in zip(Ws, bs)], axis=-1)
    assert np.allclose(x @ W.T + b, ref, atol=1e-14)


LR_KEYS = ["Wd", "bd", "Wu", "bu", "ln_f_w", "ln_f_b", "Wp", "bp", "ln_p_w", "ln_p_b", "E_img", "E_obj"]


@pytest.mark.parametrize("path", golden_files("k3lr_"), ids=os.path.basename)
def test_lowrank_visual_projector_matches_reference(path):
    """LowRankVisualEmbedding (src/modeling_bart.py:195-334)."""
    g = load(path)
    gated, residual = bool(int(g["meta_gated"])), bool(int(g["meta_residual"]))
    p = {k: g[k] for k in LR_KEYS + (["Gd", "gbd", "Gu", "gbu"] if gated else [])}
    out, c = O.lowrank_visproj_fwd(g["feats"], g["pos"], p, g.get("img_ids"), g.get("obj_ids"), gated=gated, residual=residual)
    assert rel(out, g["out"]) < TOL
    dfeats, gr = O.lowrank_visproj_bwd(g["dout"], p, c)
    assert rel(dfeats, g["dfeats"]) < TOL
    assert set(gr) == set(p) - {"E_obj"}
    for k, v in gr.items():
        assert rel(v, g["d" + k]) < TOL, k


def test_dropout_stream_restatement_statistics():
    """oracle.pet_oracle.dropout_mask restates the kernels' counter-based dropout stream (csrc/vlpet_common.cuh drop_hash4); the
    GPU test test_k1_dropout_matches_oracle_with_the_same_mask holds the kernels to it bit for bit.  Here: kept fraction,
    scale, determinism per seed, independence of consecutive seeds (CUDA-graph replays bump the seed by one)."""
    from oracle import pet_oracle as O
    m = O.dropout_mask(1234, 0.1, 512, 768)
    assert m.shape == (512, 768) and abs((m == 0).mean() - 0.1) < 2e-3
    assert np.allclose(np.unique(m), [0.0, 1.0 / (1.0 - 6554 / 65536.0)], rtol=1e-6)
    assert np.array_equal(m, O.dropout_mask(1234, 0.1, 512, 768))
    m2 = O.dropout_mask(1235, 0.1, 512, 768)
    assert abs(((m == 0) & (m2 == 0)).mean() - 0.01) < 1e-3
    assert np.array_equal(O.dropout_mask(7, 0.0, 3, 8), np.ones((3, 8)))
    x = np.arange(24.0).reshape(3, 8)
    p = {"Wd": np.zeros((2, 8)), "bd": np.zeros(2), "Wu": np.zeros((8, 2)), "bu": np.ones(8)}
    mk = O.dropout_mask(3, 0.5, 3, 8)
    out, c = O.gated_pet_fwd(x, x, p, O.PetConfig(gate="none", kappa=0.0), mask=mk)     # h = bu = 1 -> out - x1 = mask
    assert np.array_equal(out - x, mk)
    dx1, dx2, gr = O.gated_pet_bwd(np.ones((3, 8)), p, O.PetConfig(gate="none", kappa=0.0), c)
    assert np.array_equal(gr["bu"], mk.sum(0))

