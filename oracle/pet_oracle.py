"""ORACLE — test infrastructure, NOT product code.

CPU (numpy) restatement of the reference's PET hot path (HenryHZY/VL-PET), forward and analytic
backward, in whatever float dtype the inputs carry (tests run it in float64).  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` leg may import
this module; the product path (``vl-pet_b200``) never does and fails loudly without its CUDA library.

Parity status: the reference has no tests and no golden vectors for this path, so the
pin is our own: ``tests/golden/*.npz`` were produced by running the reference's *own* layer classes
(imported unmodified from /root/reference/src with the import shims of ``tests/golden/ref_import.py``)
and ``tests/test_oracle.py`` checks every function below against them (fwd and all grads, fp64,
<= 1e-12 relative).

Every function cites the reference lines it restates (paths relative to /root/reference/src).

Notation: M = B*L tokens, d = d_model, r = adapter rank, rg = gate rank.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict

import numpy as np

GATE_NONE = "none"
GATE_LARGE = "large"          # --use_encoder_adapter_gating_large_x_lowrank
GATE_MIDDLE_X = "middle_x"    # --use_encoder_adapter_gating_middle_xy_add
GATE_MIDDLE_Y = "middle_y"    # --use_encoder_adapter_gating_middle_ia3_add
GATE_SMALL = "small"          # --use_encoder_adapter_gating_small_xy_cat
GATE_KINDS = (GATE_NONE, GATE_LARGE, GATE_MIDDLE_X, GATE_MIDDLE_Y, GATE_SMALL)

_C = math.sqrt(2.0 / math.pi)
_K = 0.044715


def gelu_new(t):
    """transformers.activations.NewGELUActivation (third party, pinned transformers==4.2.1; reached via
    get_activation('gelu_new') at my_transformers/modeling_bart.py:1000,1044 and adapters/adapter_utils.py:10):
    0.5*x*(1+tanh(sqrt(2/pi)*(x+0.044715*x^3)))."""
    return 0.5 * t * (1.0 + np.tanh(_C * (t + _K * t * t * t)))


def gelu_new_grad(t):
    u = _C * (t + _K * t * t * t)
    th = np.tanh(u)
    return 0.5 * (1.0 + th) + 0.5 * t * (1.0 - th * th) * _C * (1.0 + 3.0 * _K * t * t)


def sigmoid(t):
    return 1.0 / (1.0 + np.exp(-t))


@dataclass
class PetConfig:
    """Flags read off ``layer.config`` by the reference (param.py:262-376)."""
    gate: str = GATE_LARGE
    add_gate: bool = False        # --use_encoder_adapter_gating_add   (modeling_bart.py:1206-1207)
    s: float = 1.0                # --encoder_gating_scaling_factor    (modeling_bart.py:1256-1257)
    alpha: float = 1.0            # --encoder_adapter_scaling_factor   (modeling_t5.py:789-790)
    kappa: float = 1.0            # --encoder_x2_scaling_factor        (modeling_t5.py:792-793)
    seq_len: int = 0              # L; needed by the small gate (mean over dim=1)


def stack_heads(head_weights, head_biases):
    """Multi-head down projection == one Linear whose weight is the row-concatenation of the heads
    (my_transformers/modeling_bart.py:1045-1051,1149-1150)."""
    return np.concatenate(list(head_weights), axis=0), np.concatenate(list(head_biases), axis=0)


# --------------------------------------------------------------------------------------------------------
# Dropout stream of the CUDA kernels (vl-pet_b200/csrc/vlpet_common.cuh: drop_hash4 / drop_thr16), restated in numpy so that
# a parity test can hand the oracle the SAME mask the kernels draw.  The reference's own dropout is torch's Philox stream
# (modeling_bart.py:1259 `F.dropout`); which elements are dropped is not part of parity, the arithmetic around the mask is.
# --------------------------------------------------------------------------------------------------------
def dropout_mask(seed: int, p_drop: float, M: int, d: int) -> np.ndarray:
    """Multiplicative mask [M, d] (0 or 1/(1-p')) of the counter-based stream: one 64-bit hash per 4 consecutive elements of
    the flat [M*d] index, 16 bits per element, keep iff bits >= thr16 = round(p * 65536); p' = thr16 / 65536."""
    if p_drop <= 0.0:
        return np.ones((M, d))
    t = np.float32(p_drop) * np.float32(65536.0) + np.float32(0.5)
    thr16 = 0 if t <= 0 else (65535 if t >= 65535 else int(t))
    n4 = (M * d + 3) // 4
    idx4 = np.arange(n4, dtype=np.uint64)
    m32 = np.uint64(0xFFFFFFFF)
    c0 = idx4 & m32
    c1 = (idx4 >> np.uint64(32)) ^ np.uint64((seed >> 32) & 0xFFFFFFFF)
    k = seed & 0xFFFFFFFF
    for _ in range(5):                                   # Philox-2x32-style rounds: 32x32 -> 64 multiply, xor, swap
        prod = c0 * np.uint64(0xD256D193)
        c0 = ((prod >> np.uint64(32)) ^ np.uint64(k) ^ c1) & m32
        c1 = prod & m32
        k = (k + 0x9E3779B9) & 0xFFFFFFFF
    h = (c1 << np.uint64(32)) | c0
    bits = np.stack([(h >> np.uint64(16 * j)) & np.uint64(0xFFFF) for j in range(4)], axis=1).reshape(-1)[:M * d]
    keep = bits >= np.uint64(thr16)
    inv_keep = 1.0 / (1.0 - thr16 / 65536.0) if thr16 else 1.0
    return np.where(keep, np.float32(inv_keep).astype(np.float64), 0.0).reshape(M, d)


# --------------------------------------------------------------------------------------------------------
# K1: granularity-controlled PET module (encoder, after self-attention and after the FFN)
# --------------------------------------------------------------------------------------------------------
def gated_pet_fwd(x1, x2, p: Dict[str, np.ndarray], cfg: PetConfig, rnd=None, mask=None):
    """out = x1 + s * gate(x1, y1),  y1 = kappa*x2 + alpha*Up(gelu_new(Down(x2))).

    Restates my_transformers/modeling_bart.py:1145-1155 (adapter), 1195-1231 (gates), 1256-1260 (scale,
    dropout=identity, residual) and the T5 twins my_transformers/modeling_t5.py:777-824, 359-409.
    The BART LayerNorm that follows (1261) is NOT included.

    x1, x2: [M, d].  p: Wd [r,d], bd [r], Wu [d,r], bu [d] and, per gate,
      large:    Gd [rg,d], gbd [rg], Gu [d,rg], gbu [d]
      middle_x: gw [d], gb []           (Linear(d,1))
      middle_y: gz [d]                  (bare parameter)
      small:    gw [2d], gb []          (Linear(2d,1))
    Returns (out [M,d], cache).

    ``rnd`` (optional, large gate and ungated form only): a rounding function applied to the activations that feed the second GEMM
    of each branch (z, q) -- and in the backward to du, dt, da, dp -- i.e. the points where a reference run in
    bf16 stores an activation before the next nn.Linear consumes it.  rnd=None is exact arithmetic in the input dtype.
    """
    R = rnd if (rnd is not None and cfg.gate in (GATE_LARGE, GATE_NONE)) else (lambda t: t)
    a = x2 @ p["Wd"].T + p["bd"]
    z = R(gelu_new(a))
    u = z @ p["Wu"].T + p["bu"]
    y1 = cfg.kappa * x2 + cfg.alpha * u
    c = dict(x1=x1, x2=x2, a=a, z=z, y1=y1)
    g = cfg.gate
    if g == GATE_LARGE:
        pp = x1 @ p["Gd"].T + p["gbd"]
        q = R(gelu_new(pp))
        t = q @ p["Gu"].T + p["gbu"]
        G = sigmoid(t)
        c.update(p=pp, q=q, G=G)
        h = y1 + G if cfg.add_gate else y1 * G
    elif g == GATE_MIDDLE_X:
        tm = (x1 + y1) @ p["gw"] + p["gb"]                    # [M]
        G = sigmoid(tm)[:, None]
        c.update(G=G)
        h = y1 + G if cfg.add_gate else y1 * G
    elif g == GATE_MIDDLE_Y:
        h = (y1 + 1.0 + p["gz"]) if cfg.add_gate else (y1 + y1 * p["gz"])
    elif g == GATE_SMALL:
        L = cfg.seq_len
        assert L > 0 and x1.shape[0] % L == 0
        d = x1.shape[1]
        ts = x1 @ p["gw"][:d] + y1 @ p["gw"][d:] + p["gb"]    # [M]
        sg = sigmoid(ts).reshape(-1, L)
        Gb = sg.mean(axis=1)                                   # [B]
        G = np.repeat(Gb, L)[:, None]
        c.update(sg=sg, G=G)
        h = y1 + G if cfg.add_gate else y1 * G
    elif g == GATE_NONE:
        h = y1
    else:
        raise ValueError(g)
    if mask is not None:
        c["mask"] = mask
        h = h * mask
    out = x1 + cfg.s * h
    return out, c


def gated_pet_bwd(dout, p: Dict[str, np.ndarray], cfg: PetConfig, c, rnd=None):
    """Analytic backward of gated_pet_fwd (SURVEY Appendix A for the large gate; the middle/small gates
    route an extra gradient through y1 and x1 into the gate).  Returns (dx1, dx2, grads dict)."""
    x1, x2, a, z, y1 = c["x1"], c["x2"], c["a"], c["z"], c["y1"]
    R = rnd if (rnd is not None and cfg.gate in (GATE_LARGE, GATE_NONE)) else (lambda t: t)
    dh = cfg.s * dout
    if c.get("mask") is not None:
        dh = dh * c["mask"]
    gr: Dict[str, np.ndarray] = {}
    dx1 = dout.copy()
    g = cfg.gate
    if g == GATE_LARGE:
        G = c["G"]
        if cfg.add_gate:
            dy1, dG = dh, dh
        else:
            dy1, dG = dh * G, dh * y1
        dt = R(dG * G * (1.0 - G))
        gr["Gu"] = dt.T @ c["q"]
        gr["gbu"] = dt.sum(0)
        dq = dt @ p["Gu"]
        dp_exact = dq * gelu_new_grad(c["p"])
        dp = R(dp_exact)
        gr["Gd"] = dp.T @ x1
        gr["gbd"] = dp_exact.sum(0)       # bias sums are taken before the storage rounding (fp32 in the kernel's epilogue)
        dx1 = dx1 + dp @ p["Gd"]
    elif g == GATE_MIDDLE_X:
        G = c["G"]
        if cfg.add_gate:
            dy1 = dh.copy()
            dG = dh.sum(1, keepdims=True)
        else:
            dy1 = dh * G
            dG = (dh * y1).sum(1, keepdims=True)
        dtm = (dG * G * (1.0 - G))[:, 0]
        gr["gw"] = dtm @ (x1 + y1)
        gr["gb"] = dtm.sum()
        dy1 = dy1 + dtm[:, None] * p["gw"][None, :]
        dx1 = dx1 + dtm[:, None] * p["gw"][None, :]
    elif g == GATE_MIDDLE_Y:
        if cfg.add_gate:
            dy1 = dh
            gr["gz"] = dh.sum(0)
        else:
            dy1 = dh * (1.0 + p["gz"])
            gr["gz"] = (dh * y1).sum(0)
    elif g == GATE_SMALL:
        L = cfg.seq_len
        d = x1.shape[1]
        G, sg = c["G"], c["sg"]
        if cfg.add_gate:
            dy1 = dh.copy()
            dGb = dh.reshape(-1, L, d).sum((1, 2))
        else:
            dy1 = dh * G
            dGb = (dh * y1).reshape(-1, L, d).sum((1, 2))
        dts = ((dGb / L)[:, None] * sg * (1.0 - sg)).reshape(-1)
        gr["gw"] = np.concatenate([dts @ x1, dts @ y1])
        gr["gb"] = dts.sum()
        dx1 = dx1 + dts[:, None] * p["gw"][None, :d]
        dy1 = dy1 + dts[:, None] * p["gw"][None, d:]
    elif g == GATE_NONE:
        dy1 = dh
    else:
        raise ValueError(g)
    # ungated form (the K2 value parallel adapter): du = alpha * dout is never stored, dout itself feeds the GEMMs
    du = R(cfg.alpha * dy1) if g != GATE_NONE else cfg.alpha * dy1
    gr["Wu"] = du.T @ z
    gr["bu"] = du.sum(0)
    dz = du @ p["Wu"]
    da_exact = dz * gelu_new_grad(a)
    da = R(da_exact)
    gr["Wd"] = da.T @ x2
    gr["bd"] = da_exact.sum(0)            # as above: only the GEMM operand da is stored in bf16
    dx2 = cfg.kappa * dy1 + da @ p["Wd"]
    return dx1, dx2, gr


# --------------------------------------------------------------------------------------------------------
# K2: decoder cross-attention value parallel adapter
# --------------------------------------------------------------------------------------------------------
def vpa_fwd(kv, y, p, scaling_factor: float = 1.0):
    """AdapterController.forward(inputs=kv, task, y=v) with use_parallel_adapter=True:
    out = y + sf * Up(gelu_new(Down(kv)))   (adapters/adapter_controller.py:131-162,
    adapters/adapter_modeling.py:55-61; call site my_transformers/modeling_bart.py:427-430).
    With y=None and use_parallel_adapter=False the residual is ``inputs`` (adapter_controller.py:160-161):
    pass y=kv for that."""
    a = kv @ p["Wd"].T + p["bd"]
    z = gelu_new(a)
    out = y + scaling_factor * (z @ p["Wu"].T + p["bu"])
    return out, dict(kv=kv, a=a, z=z)


def vpa_bwd(dout, p, c, scaling_factor: float = 1.0):
    """Returns (dkv [adapter path only], dy, grads)."""
    du = scaling_factor * dout
    gr = {"Wu": du.T @ c["z"], "bu": du.sum(0)}
    da = (du @ p["Wu"]) * gelu_new_grad(c["a"])
    gr["Wd"] = da.T @ c["kv"]
    gr["bd"] = da.sum(0)
    return da @ p["Wd"], dout, gr


# --------------------------------------------------------------------------------------------------------
# K3: visual projection (CLIP feature -> d_model)
# --------------------------------------------------------------------------------------------------------
def layer_norm_fwd(x, w, b, eps):
    mu = x.mean(-1, keepdims=True)
    xc = x - mu
    rstd = 1.0 / np.sqrt((xc * xc).mean(-1, keepdims=True) + eps)
    xh = xc * rstd
    return xh * w + b, (xh, rstd)


def layer_norm_bwd(dy, w, cache):
    xh, rstd = cache
    dxh = dy * w
    dx = rstd * (dxh - dxh.mean(-1, keepdims=True) - xh * (dxh * xh).mean(-1, keepdims=True))
    return dx, (dy * xh).reshape(-1, xh.shape[-1]).sum(0), dy.reshape(-1, xh.shape[-1]).sum(0)


def rms_norm_fwd(x, w, eps):
    """T5LayerNorm (my_transformers/modeling_t5.py:235-252): x * rsqrt(mean(x^2)+eps) * w, no bias, no mean."""
    rstd = 1.0 / np.sqrt((x * x).mean(-1, keepdims=True) + eps)
    xh = x * rstd
    return xh * w, (xh, rstd)


def rms_norm_bwd(dy, w, cache):
    xh, rstd = cache
    dxh = dy * w
    dx = rstd * (dxh - xh * (dxh * xh).mean(-1, keepdims=True))
    return dx, (dy * xh).reshape(-1, xh.shape[-1]).sum(0)


def visproj_fwd(feats, pos, p, img_order_ids=None, obj_order_ids=None, rms: bool = False, eps: float = 1e-5):
    """VisualEmbedding.forward (src/modeling_bart.py:143-192; T5: src/modeling_t5.py:124-174, rms=True,
    eps=config.layer_norm_epsilon) with use_vis_layer_norm & individual_vis_layer_norm (param.py defaults):

        LN(feats Wf^T + bf) + LN([pos, area] Wp^T + bp) + E_img[img_ids] + E_obj[V-1-obj_ids]

    feats [B,N,F], pos [B,N,4] (x1,x2,y1,y2); p: Wf [d,F], bf, ln_f_w, (ln_f_b), Wp [d,5], bp, ln_p_w,
    (ln_p_b), E_img [n_images,d], E_obj [V,d] (aliases the token embedding, frozen)."""
    B, N, _ = feats.shape
    f = feats @ p["Wf"].T + p["bf"]
    area = ((pos[:, :, 3] - pos[:, :, 2]) * (pos[:, :, 1] - pos[:, :, 0]))[:, :, None]
    pos5 = np.concatenate([pos, area], axis=2)
    a = pos5 @ p["Wp"].T + p["bp"]
    if rms:
        fe, cf = rms_norm_fwd(f, p["ln_f_w"], eps)
        ae, ca = rms_norm_fwd(a, p["ln_p_w"], eps)
    else:
        fe, cf = layer_norm_fwd(f, p["ln_f_w"], p["ln_f_b"], eps)
        ae, ca = layer_norm_fwd(a, p["ln_p_w"], p["ln_p_b"], eps)
    if img_order_ids is None:
        img_order_ids = np.zeros((1, N), dtype=np.int64)
    if obj_order_ids is None:
        obj_order_ids = np.arange(N, dtype=np.int64)[None]
    V = p["E_obj"].shape[0]
    obj_rows = V - obj_order_ids - 1
    out = fe + ae + p["E_img"][img_order_ids] + p["E_obj"][obj_rows]
    return out, dict(feats=feats, pos5=pos5, cf=cf, ca=ca, img_ids=np.broadcast_to(img_order_ids, (B, N)))


def visproj_bwd(dout, p, c, rms: bool = False):
    """Grads for the trainable visual_embedding parameters (E_obj aliases model.shared and stays frozen,
    trainer_base.py:497-533).  Returns (dfeats, grads)."""
    gr = {}
    d = dout.shape[-1]
    if rms:
        df, gr["ln_f_w"] = rms_norm_bwd(dout, p["ln_f_w"], c["cf"])
        da, gr["ln_p_w"] = rms_norm_bwd(dout, p["ln_p_w"], c["ca"])
    else:
        df, gr["ln_f_w"], gr["ln_f_b"] = layer_norm_bwd(dout, p["ln_f_w"], c["cf"])
        da, gr["ln_p_w"], gr["ln_p_b"] = layer_norm_bwd(dout, p["ln_p_w"], c["ca"])
    df2 = df.reshape(-1, d)
    da2 = da.reshape(-1, d)
    gr["Wf"] = df2.T @ c["feats"].reshape(df2.shape[0], -1)
    gr["bf"] = df2.sum(0)
    gr["Wp"] = da2.T @ c["pos5"].reshape(da2.shape[0], -1)
    gr["bp"] = da2.sum(0)
    E = np.zeros_like(p["E_img"])
    np.add.at(E, c["img_ids"].reshape(-1), dout.reshape(-1, d))
    gr["E_img"] = E
    return df @ p["Wf"], gr


# --------------------------------------------------------------------------------------------------------
# K3-LR: the PET-shaped visual projector (flag-gated: --use_lowrank_visual_projector)
# --------------------------------------------------------------------------------------------------------
def lowrank_visproj_fwd(feats, pos, p, img_order_ids=None, obj_order_ids=None, gated: bool = True,
                        residual: bool = False, eps: float = 1e-5):
    """LowRankVisualEmbedding.forward (src/modeling_bart.py:263-334):

        e   = Up(gelu_new(cat_h Down_h(feats)))                                         (275-281)
        e   = e (+ e) * sigmoid(GUp(gelu_new(GDown(feats))))   with the gate flag        (283-293)
        out = LN(e) + LN([pos, area] Wp^T + bp) + E_img[img_ids] + E_obj[V-1-obj_ids]   (296-327)

    p: Wd [r,F] (row-concatenated heads), bd, Wu [d,r], bu, (Gd [rg,F], gbd, Gu [d,rg], gbu), ln_f_w, ln_f_b
    (visual_projector_layer_norm), Wp, bp, ln_p_w, ln_p_b, E_img, E_obj."""
    B, N, F = feats.shape
    f2 = feats.reshape(-1, F)
    a = f2 @ p["Wd"].T + p["bd"]
    z = gelu_new(a)
    u = z @ p["Wu"].T + p["bu"]
    c = dict(f2=f2, a=a, z=z, u=u)
    if gated:
        pp = f2 @ p["Gd"].T + p["gbd"]
        q = gelu_new(pp)
        G = sigmoid(q @ p["Gu"].T + p["gbu"])
        e = u + u * G if residual else u * G
        c.update(p=pp, q=q, G=G)
    else:
        e = u
    e = e.reshape(B, N, -1)
    area = ((pos[:, :, 3] - pos[:, :, 2]) * (pos[:, :, 1] - pos[:, :, 0]))[:, :, None]
    pos5 = np.concatenate([pos, area], axis=2)
    apos = pos5 @ p["Wp"].T + p["bp"]
    fe, cf = layer_norm_fwd(e, p["ln_f_w"], p["ln_f_b"], eps)
    ae, ca = layer_norm_fwd(apos, p["ln_p_w"], p["ln_p_b"], eps)
    if img_order_ids is None:
        img_order_ids = np.zeros((1, N), dtype=np.int64)
    if obj_order_ids is None:
        obj_order_ids = np.arange(N, dtype=np.int64)[None]
    V = p["E_obj"].shape[0]
    out = fe + ae + p["E_img"][img_order_ids] + p["E_obj"][V - obj_order_ids - 1]
    c.update(pos5=pos5, cf=cf, ca=ca, img_ids=np.broadcast_to(img_order_ids, (B, N)), gated=gated, residual=residual)
    return out, c


def lowrank_visproj_bwd(dout, p, c):
    """Grads of every trainable parameter of the low-rank visual projector.  Returns (dfeats, grads)."""
    gr = {}
    d = dout.shape[-1]
    de, gr["ln_f_w"], gr["ln_f_b"] = layer_norm_bwd(dout, p["ln_f_w"], c["cf"])
    da_pos, gr["ln_p_w"], gr["ln_p_b"] = layer_norm_bwd(dout, p["ln_p_w"], c["ca"])
    de = de.reshape(-1, d)
    da_pos = da_pos.reshape(-1, d)
    gr["Wp"] = da_pos.T @ c["pos5"].reshape(da_pos.shape[0], -1)
    gr["bp"] = da_pos.sum(0)
    E = np.zeros_like(p["E_img"])
    np.add.at(E, c["img_ids"].reshape(-1), dout.reshape(-1, d))
    gr["E_img"] = E
    f2 = c["f2"]
    dfeats = 0.0
    if c["gated"]:
        G, u = c["G"], c["u"]
        du = de * (1.0 + G) if c["residual"] else de * G
        dt = de * u * G * (1.0 - G)
        gr["Gu"] = dt.T @ c["q"]
        gr["gbu"] = dt.sum(0)
        dp = (dt @ p["Gu"]) * gelu_new_grad(c["p"])
        gr["Gd"] = dp.T @ f2
        gr["gbd"] = dp.sum(0)
        dfeats = dp @ p["Gd"]
    else:
        du = de
    gr["Wu"] = du.T @ c["z"]
    gr["bu"] = du.sum(0)
    da = (du @ p["Wu"]) * gelu_new_grad(c["a"])
    gr["Wd"] = da.T @ f2
    gr["bd"] = da.sum(0)
    dfeats = dfeats + da @ p["Wd"]
    return dfeats.reshape(dout.shape[0], dout.shape[1], -1), gr
