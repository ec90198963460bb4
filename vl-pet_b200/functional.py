"""torch.autograd bindings of the libvlpet.so kernels (PyTorch is plumbing here: device memory, streams, autograd
graph).  Every function requires CUDA tensors and raises otherwise -- there is no CPU path in the product."""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass
from typing import List, Optional, Sequence

import torch

from . import _lib as L

_DT = {torch.float32: L.F32, torch.bfloat16: L.BF16}
_workspaces = {}
_seed_counter = [0]
_seed_dev = None   # optional device int64 scalar added to every dropout seed (CUDA-graph replays), see set_device_seed
_prof = None   # when profiling: list of (kernel name, algorithmic bytes, start event, end event)


_direct_grads = False


def set_direct_grad_accumulation(flag: bool):
    """When on, the backward kernels accumulate the fp32 weight gradients STRAIGHT into ``param.grad`` (the C ABI's
    `+=` contract: with PetBucket those are views of the flat all-reduce payload) and autograd receives ``None`` for
    the parameters -- no temporary gradient buffer, no per-parameter AccumulateGrad kernels.  Requires every parameter
    of the call to own a pre-allocated contiguous fp32 ``.grad`` (heads adjacent); otherwise the call falls back to
    returning gradients.  Off by default (hooks / torch.autograd.grad need the returned tensors)."""
    global _direct_grads
    _direct_grads = bool(flag)


def _direct_targets(params, groups):
    """groups: list of index lists into params that must form ONE contiguous fp32 buffer each (e.g. the heads of a
    multi-head down projection).  -> list of data pointers (one per group) or None."""
    if not _direct_grads:
        return None
    ptrs = []
    for idxs in groups:
        g0 = params[idxs[0]].grad
        if g0 is None or g0.dtype != torch.float32 or not g0.is_contiguous() or (g0.data_ptr() & 15):
            return None
        off = g0.data_ptr()
        for i in idxs:
            gi = params[i].grad
            if gi is None or gi.dtype != torch.float32 or not gi.is_contiguous() or gi.data_ptr() != off or \
                    not params[i].requires_grad:
                return None
            off += gi.numel() * 4
        ptrs.append(g0.data_ptr())
    return ptrs


def set_device_seed(t: Optional[torch.Tensor]):
    """Register a device int64 scalar that the K1 kernels add to their dropout seed at run time (VlpetK1Desc.seed_dev).
    A CUDA graph that captured the calls then draws a fresh mask on every replay once the caller bumps the scalar
    (on the stream, between replays).  None switches back to host-side seeds."""
    global _seed_dev
    if t is not None and (not t.is_cuda or t.dtype != torch.int64 or t.numel() != 1):
        raise ValueError("vlpet.set_device_seed: expected a CUDA int64 scalar tensor")
    _seed_dev = t


def profile_kernels(enable: bool):
    """Bracket every C-ABI compute call with CUDA events on the launching stream (bench.py's roofline leg).
    Returns the records collected so far when disabling."""
    global _prof
    old, _prof = _prof, ([] if enable else None)
    return old


def profile_summary(records):
    """-> {kernel: {"launches", "ms", "bytes"}} (call after torch.cuda.synchronize())."""
    out = {}
    for name, nbytes, s, e in records or []:
        d = out.setdefault(name, {"launches": 0, "ms": 0.0, "bytes": 0})
        d["launches"] += 1
        d["ms"] += s.elapsed_time(e)
        d["bytes"] += nbytes
    return out


_NVTX = os.environ.get("VLPET_NVTX", "0") not in ("", "0")     # NVTX ranges around every C-ABI compute call (ncu --nvtx filters on them)


def _call(name: str, nbytes: int, fn, *args):
    if _NVTX:
        torch.cuda.nvtx.range_push("vlpet." + name)
        try:
            return _call_inner(name, nbytes, fn, *args)
        finally:
            torch.cuda.nvtx.range_pop()
    return _call_inner(name, nbytes, fn, *args)


def _call_inner(name: str, nbytes: int, fn, *args):
    if _prof is None:
        return fn(*args)
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    r = fn(*args)
    e.record()
    _prof.append((name, nbytes, s, e))
    return r


def _require_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("vlpet: CUDA tensors required -- the PET kernels have no CPU fallback")


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _workspace(nbytes: int, device) -> torch.Tensor:
    """Grow-only per-(device, stream) scratch buffer; safe because every call is ordered on the current stream.
    Under CUDA-graph capture the buffer comes from the graph's private pool instead (its address is baked into the
    captured launches, so it must not be a cached tensor that a later, larger request could replace)."""
    if torch.cuda.is_current_stream_capturing():
        return torch.empty(max(nbytes, 256), dtype=torch.uint8, device=device)
    key = (device.index, torch.cuda.current_stream(device).cuda_stream)
    buf = _workspaces.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, device=device)
        _workspaces[key] = buf
    return buf


def _p(t: Optional[torch.Tensor]):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _as(t: Optional[torch.Tensor], dtype) -> Optional[torch.Tensor]:
    if t is None:
        return None
    sh = getattr(t, "_vlpet_shadow", None)  # bf16 shadow kept fresh by PetBucket (refresh_shadow / the fused AdamW kernel)
    if sh is not None and sh.dtype == dtype and getattr(t, "_vlpet_shadow_version", None) == t._version:
        return sh      # an in-place write through torch since the last refresh (load_state_dict, copy_) bumps _version: cast instead
    t = t.detach()
    return (t if t.dtype == dtype else t.to(dtype)).contiguous()


def _stack_rows(ts: Sequence[torch.Tensor], dtype) -> torch.Tensor:
    """Row-concatenate head tensors.  Free when the heads are adjacent slices of one buffer."""
    ts = [_as(t, dtype) for t in ts]
    if len(ts) == 1:
        return ts[0]
    t0 = ts[0]
    step = t0.numel() * t0.element_size()
    if all(t.shape == t0.shape and t.data_ptr() == t0.data_ptr() + i * step for i, t in enumerate(ts)):
        rows = t0.shape[0] * len(ts)
        try:
            return torch.as_strided(t0, (rows,) + tuple(t0.shape[1:]), t0.stride())
        except RuntimeError:
            pass
    return torch.cat(ts, dim=0)


def next_dropout_seed() -> int:
    _seed_counter[0] += 1
    return (torch.initial_seed() * 0x9E3779B1 + _seed_counter[0]) & 0xFFFFFFFFFFFFFFFF


@dataclass(frozen=True)
class PetSiteConfig:
    """Flags the reference reads off ``layer.config`` at a PET site (param.py:262-376)."""
    gate: str = "large"         # large | middle_x | middle_y | small | none
    add_gate: bool = False      # use_encoder_adapter_gating_add
    s: float = 1.0              # encoder_gating_scaling_factor (1 when use_encoder_gating_scaling is off)
    alpha: float = 1.0          # encoder_adapter_scaling_factor
    kappa: float = 1.0          # encoder_x2_scaling_factor
    p_drop: float = 0.0         # dropout between gate and residual; applied only when training=True
    impl: str = "auto"          # forward kernel: auto | generic | fused
    bwd_impl: str = "auto"      # backward kernel: auto | generic | fused


_GATE_NPARAMS = {"none": 0, "large": 4, "middle_x": 2, "small": 2, "middle_y": 1}


class GatedPETFn(torch.autograd.Function):
    """out = x1 + dropout(s * gate(x1, kappa*x2 + alpha*Up(gelu_new(Down(x2)))))   -- include/vlpet.h K1.

    Tensor args after ``nheads``: down_w[0..h), down_b[0..h), up_w, up_b, then the gate parameters
    (large: gd_w, gd_b, gu_w, gu_b; middle_x / small: gw, gb; middle_y: gz)."""

    @staticmethod
    def forward(ctx, cfg: PetSiteConfig, seed: int, seq_len: int, nheads: int, x1, x2, *params):
        _require_cuda(x1, x2, *params)
        if x1.shape != x2.shape or x1.dtype != x2.dtype or x1.dtype not in _DT:
            raise ValueError(f"vlpet.gated_pet: x1/x2 must share shape and be fp32/bf16, got {x1.shape} {x1.dtype} / {x2.shape} {x2.dtype}")
        ng = _GATE_NPARAMS[cfg.gate]
        if len(params) != 2 * nheads + 2 + ng:
            raise ValueError(f"vlpet.gated_pet: expected {2 * nheads + 2 + ng} parameter tensors, got {len(params)}")
        dt = x1.dtype
        d = x1.shape[-1]
        x1c, x2c = x1.contiguous(), x2.contiguous()
        M = x1c.numel() // d
        Wd = _stack_rows(params[:nheads], dt)
        bd = _stack_rows(params[nheads:2 * nheads], dt)
        Wu, bu = _as(params[2 * nheads], dt), _as(params[2 * nheads + 1], dt)
        gp = [_as(t, dt) for t in params[2 * nheads + 2:]]
        r = Wd.shape[0]
        if Wd.shape != (r, d) or Wu.shape != (d, r) or bd.numel() != r or bu.numel() != d:
            raise ValueError(f"vlpet.gated_pet: inconsistent adapter shapes Wd{tuple(Wd.shape)} Wu{tuple(Wu.shape)}")
        rg = gp[0].shape[0] if cfg.gate == "large" else 0
        desc = L.K1Desc(M=M, L=seq_len, d=d, r=r, rg=rg, gate=L.GATE_IDS[cfg.gate], add_gate=int(cfg.add_gate),
                        dtype=_DT[dt], impl=L.IMPL_IDS[cfg.impl], s=cfg.s, alpha=cfg.alpha, kappa=cfg.kappa,
                        p_drop=cfg.p_drop if seed else 0.0, seed=seed,
                        seed_dev=(_seed_dev.data_ptr() if (seed and _seed_dev is not None) else None))
        w = L.K1Params(Wd=_p(Wd), bd=_p(bd), Wu=_p(Wu), bu=_p(bu))
        if cfg.gate == "large":
            w.Gd, w.gbd, w.Gu, w.gbu = _p(gp[0]), _p(gp[1]), _p(gp[2]), _p(gp[3])
        elif cfg.gate in ("middle_x", "small"):
            w.gw, w.gb = _p(gp[0]), _p(gp[1])
        elif cfg.gate == "middle_y":
            w.gz = _p(gp[0])
        out = torch.empty_like(x1c)
        nws = L.lib.vlpet_k1_fwd_workspace_bytes(C.byref(desc))
        ws = _workspace(nws, x1.device)
        L.check(_call("k1_fwd", 3 * x1c.numel() * x1c.element_size(), L.lib.vlpet_k1_fwd, C.byref(desc), _p(x1c), _p(x2c),
                      C.byref(w), _p(out), _p(ws), ws.numel(), _stream()), "vlpet_k1_fwd")
        ctx.desc, ctx.nheads, ctx.cfg = desc, nheads, cfg
        ctx.param_meta = [(tuple(t.shape), t.dtype) for t in params]
        ctx.param_refs = params
        ctx.save_for_backward(x1c, x2c, Wd, bd, Wu, bu, *gp)
        return out

    @staticmethod
    def backward(ctx, dout):
        x1, x2, Wd, bd, Wu, bu, *gp = ctx.saved_tensors
        desc, nheads, cfg = ctx.desc, ctx.nheads, ctx.cfg
        desc.impl = L.IMPL_IDS[cfg.bwd_impl]
        dout = dout.contiguous()
        if dout.dtype != x1.dtype:
            dout = dout.to(x1.dtype)
        d, r, rg = desc.d, desc.r, desc.rg
        sizes = [r * d, r, d * r, d]
        if cfg.gate == "large":
            sizes += [rg * d, rg, d * rg, d]
        elif cfg.gate == "middle_x":
            sizes += [d, 1]
        elif cfg.gate == "small":
            sizes += [2 * d, 1]
        elif cfg.gate == "middle_y":
            sizes += [d]
        offs = [0]
        for n in sizes:
            offs.append(offs[-1] + (n + 3) // 4 * 4)      # keep every grad 16-byte aligned
        nh = nheads
        groups = [list(range(nh)), list(range(nh, 2 * nh))] + [[i] for i in range(2 * nh, len(ctx.param_refs))]
        direct = _direct_targets(ctx.param_refs, groups)
        if direct is not None:
            vp = [C.c_void_p(a) for a in direct]
            gv = None
        else:
            gbuf = torch.zeros(offs[-1], dtype=torch.float32, device=x1.device)
            gv = [gbuf[offs[i]:offs[i] + sizes[i]] for i in range(len(sizes))]
            vp = [_p(t) for t in gv]
        g = L.K1Grads(dWd=vp[0], dbd=vp[1], dWu=vp[2], dbu=vp[3])
        if cfg.gate == "large":
            g.dGd, g.dgbd, g.dGu, g.dgbu = vp[4], vp[5], vp[6], vp[7]
        elif cfg.gate in ("middle_x", "small"):
            g.dgw, g.dgb = vp[4], vp[5]
        elif cfg.gate == "middle_y":
            g.dgz = vp[4]
        w = L.K1Params(Wd=_p(Wd), bd=_p(bd), Wu=_p(Wu), bu=_p(bu))
        if cfg.gate == "large":
            w.Gd, w.gbd, w.Gu, w.gbu = _p(gp[0]), _p(gp[1]), _p(gp[2]), _p(gp[3])
        elif cfg.gate in ("middle_x", "small"):
            w.gw, w.gb = _p(gp[0]), _p(gp[1])
        elif cfg.gate == "middle_y":
            w.gz = _p(gp[0])
        dx1, dx2 = torch.empty_like(x1), torch.empty_like(x2)
        nws = L.lib.vlpet_k1_bwd_workspace_bytes(C.byref(desc))
        ws = _workspace(nws, x1.device)
        L.check(_call("k1_bwd", 5 * x1.numel() * x1.element_size(), L.lib.vlpet_k1_bwd, C.byref(desc), _p(x1), _p(x2),
                      _p(dout), C.byref(w), _p(dx1), _p(dx2), C.byref(g), _p(ws), ws.numel(), _stream()), "vlpet_k1_bwd")
        if gv is None:       # gradients were accumulated into param.grad by the kernels
            return (None, None, None, None, dx1, dx2) + (None,) * len(ctx.param_refs)
        # scatter the flat fp32 grads back onto the parameter list (heads are row slices of dWd / dbd)
        meta = ctx.param_meta
        grads: List[Optional[torch.Tensor]] = []
        hr = r // nheads
        dWd, dbd = gv[0].view(r, d), gv[1]
        for h in range(nheads):
            grads.append(dWd[h * hr:(h + 1) * hr])
        for h in range(nheads):
            grads.append(dbd[h * hr:(h + 1) * hr])
        grads += [gv[2].view(d, r), gv[3]]
        for i in range(4, len(sizes)):
            grads.append(gv[i])
        outg = []
        for gt, (shape, dtype) in zip(grads, meta):
            gt = gt.reshape(shape)
            outg.append(gt if dtype == torch.float32 else gt.to(dtype))
        return (None, None, None, None, dx1, dx2, *outg)


def gated_pet(x1, x2, down_ws, down_bs, up_w, up_b, gate_params=(), cfg: PetSiteConfig = PetSiteConfig(),
              training: bool = False):
    """Functional form of one encoder PET site (include/vlpet.h K1).  x1, x2: [B, L, d] (or [M, d])."""
    seed = next_dropout_seed() if (training and cfg.p_drop > 0.0) else 0
    seq_len = x1.shape[-2] if x1.dim() >= 2 else 0
    if cfg.gate == "small" and x1.dim() < 3:
        raise ValueError("vlpet.gated_pet: the small gate averages over the sequence: pass [B, L, d]")
    return GatedPETFn.apply(cfg, seed, int(seq_len), len(down_ws), x1, x2, *down_ws, *down_bs, up_w, up_b, *gate_params)


class VpaFn(torch.autograd.Function):
    """out = y + sf * Up(gelu_new(Down(kv)))   -- include/vlpet.h K2 (y may be None: no residual)."""

    @staticmethod
    def forward(ctx, sf: float, impl: str, kv, y, Wd, bd, Wu, bu):
        _require_cuda(kv, y, Wd, bd, Wu, bu)
        if kv.dtype not in _DT or (y is not None and (y.shape != kv.shape or y.dtype != kv.dtype)):
            raise ValueError("vlpet.vpa: kv / y must share shape and be fp32/bf16")
        dt, d = kv.dtype, kv.shape[-1]
        kvc = kv.contiguous()
        yc = y.contiguous() if y is not None else None
        Wdc, bdc, Wuc, buc = (_as(t, dt) for t in (Wd, bd, Wu, bu))
        r = Wdc.shape[0]
        if Wdc.shape != (r, d) or Wuc.shape != (d, r):
            raise ValueError(f"vlpet.vpa: inconsistent shapes Wd{tuple(Wdc.shape)} Wu{tuple(Wuc.shape)} d={d}")
        desc = L.K2Desc(M=kvc.numel() // d, d=d, r=r, dtype=_DT[dt], impl=L.IMPL_IDS[impl], sf=sf)
        w = L.K2Params(Wd=_p(Wdc), bd=_p(bdc), Wu=_p(Wuc), bu=_p(buc))
        out = torch.empty_like(kvc)
        ws = _workspace(L.lib.vlpet_k2_fwd_workspace_bytes(C.byref(desc)), kv.device)
        L.check(_call("k2_fwd", (3 if yc is not None else 2) * kvc.numel() * kvc.element_size(), L.lib.vlpet_k2_fwd,
                      C.byref(desc), _p(kvc), _p(yc), C.byref(w), _p(out), _p(ws), ws.numel(), _stream()), "vlpet_k2_fwd")
        ctx.desc, ctx.has_y = desc, y is not None
        ctx.param_meta = [(tuple(t.shape), t.dtype) for t in (Wd, bd, Wu, bu)]
        ctx.param_refs = (Wd, bd, Wu, bu)
        ctx.save_for_backward(kvc, Wdc, bdc, Wuc, buc)
        return out

    @staticmethod
    def backward(ctx, dout):
        kv, Wd, bd, Wu, bu = ctx.saved_tensors
        desc = ctx.desc
        dout = dout.contiguous()
        if dout.dtype != kv.dtype:
            dout = dout.to(kv.dtype)
        d, r = desc.d, desc.r
        sizes = [r * d, r, d * r, d]
        offs = [0]
        for n in sizes:
            offs.append(offs[-1] + (n + 3) // 4 * 4)
        direct = _direct_targets(ctx.param_refs, [[0], [1], [2], [3]])
        if direct is not None:
            gv = None
            g = L.K2Grads(*[C.c_void_p(a) for a in direct])
        else:
            gbuf = torch.zeros(offs[-1], dtype=torch.float32, device=kv.device)
            gv = [gbuf[offs[i]:offs[i] + sizes[i]] for i in range(4)]
            g = L.K2Grads(dWd=_p(gv[0]), dbd=_p(gv[1]), dWu=_p(gv[2]), dbu=_p(gv[3]))
        w = L.K2Params(Wd=_p(Wd), bd=_p(bd), Wu=_p(Wu), bu=_p(bu))
        dkv = torch.empty_like(kv) if ctx.needs_input_grad[2] else None
        ws = _workspace(L.lib.vlpet_k2_bwd_workspace_bytes(C.byref(desc)), kv.device)
        L.check(_call("k2_bwd", 3 * kv.numel() * kv.element_size(), L.lib.vlpet_k2_bwd, C.byref(desc), _p(kv), _p(dout),
                      C.byref(w), _p(dkv), C.byref(g), _p(ws), ws.numel(), _stream()), "vlpet_k2_bwd")
        if gv is None:
            return (None, None, dkv, dout if ctx.has_y else None, None, None, None, None)
        outg = []
        for gt, (shape, dtype) in zip(gv, ctx.param_meta):
            gt = gt.reshape(shape)
            outg.append(gt if dtype == torch.float32 else gt.to(dtype))
        return (None, None, dkv, dout if ctx.has_y else None, *outg)


def vpa(kv, y, down_w, down_b, up_w, up_b, scaling_factor: float = 1.0, impl: str = "auto"):
    """Decoder value-parallel-adapter (adapter_controller.py:149-162): y + sf * Up(gelu_new(Down(kv)))."""
    return VpaFn.apply(float(scaling_factor), impl, kv, y, down_w, down_b, up_w, up_b)


class VisProjFn(torch.autograd.Function):
    """VisualEmbedding.forward (src/modeling_bart.py:143-192) -- include/vlpet.h K3."""

    @staticmethod
    def forward(ctx, rms: bool, eps: float, impl: str, feats, pos, img_ids, obj_ids, Wf, bf, lnfw, lnfb, Wp, bp, lnpw,
                lnpb, E_img, E_obj):
        _require_cuda(feats, pos, Wf, E_img, E_obj)
        if feats.dim() != 3 or feats.dtype not in _DT:
            raise ValueError("vlpet.visual_projection: feats must be [B, N, F] fp32/bf16")
        B, N, F = feats.shape
        if tuple(pos.shape) != (B, N, 4):
            raise ValueError(f"vlpet.visual_projection: pos must be [{B}, {N}, 4], got {tuple(pos.shape)}")
        dt = feats.dtype
        d = Wf.shape[0]
        fc = feats.contiguous()
        pc = pos.to(dt).contiguous()

        def ids(t):
            if t is None:
                return None
            return t.to(torch.int64).expand(B, N).contiguous()

        img, obj = ids(img_ids), ids(obj_ids)
        ws_ = [_as(t, dt) for t in (Wf, bf, lnfw, lnfb, Wp, bp, lnpw, lnpb, E_img, E_obj)]
        desc = L.K3Desc(M=B * N, N=N, F=F, d=d, V=E_obj.shape[0], n_img=E_img.shape[0], rms=int(rms), dtype=_DT[dt],
                        impl=L.IMPL_IDS[impl], eps=eps)
        w = L.K3Params(*[_p(t) for t in ws_])
        out = torch.empty(B, N, d, dtype=dt, device=feats.device)
        save = torch.empty(L.lib.vlpet_k3_save_floats(C.byref(desc)), dtype=torch.float32, device=feats.device)
        ws = _workspace(L.lib.vlpet_k3_fwd_workspace_bytes(C.byref(desc)), feats.device)
        L.check(_call("k3_fwd", (fc.numel() + out.numel()) * fc.element_size(), L.lib.vlpet_k3_fwd, C.byref(desc), _p(fc),
                      _p(pc), _p(img), _p(obj), C.byref(w), _p(out), _p(save), _p(ws), ws.numel(), _stream()), "vlpet_k3_fwd")
        ctx.desc = desc
        ctx.param_meta = [None if t is None else (tuple(t.shape), t.dtype) for t in (Wf, bf, lnfw, lnfb, Wp, bp, lnpw, lnpb, E_img)]
        ctx.save_for_backward(fc, pc, img, save, *[t for t in ws_ if t is not None])
        ctx.present = [t is not None for t in ws_]
        return out

    @staticmethod
    def backward(ctx, dout):
        fc, pc, img, save, *rest = ctx.saved_tensors
        it = iter(rest)
        ws_ = [next(it) if p else None for p in ctx.present]
        desc = ctx.desc
        dout = dout.contiguous()
        if dout.dtype != fc.dtype:
            dout = dout.to(fc.dtype)
        d, F = desc.d, desc.F
        sizes = [d * F, d, d, d, d * 5, d, d, d, desc.n_img * d]
        offs = [0]
        for n in sizes:
            offs.append(offs[-1] + (n + 3) // 4 * 4)
        gbuf = torch.zeros(offs[-1], dtype=torch.float32, device=fc.device)
        gv = [gbuf[offs[i]:offs[i] + sizes[i]] for i in range(len(sizes))]
        present = ctx.present
        g = L.K3Grads(*[_p(gv[i]) if present[i] else C.c_void_p(0) for i in range(9)])
        w = L.K3Params(*[_p(t) for t in ws_])
        dfeats = torch.empty_like(fc) if ctx.needs_input_grad[3] else None
        ws = _workspace(L.lib.vlpet_k3_bwd_workspace_bytes(C.byref(desc)), fc.device)
        L.check(_call("k3_bwd", (fc.numel() + dout.numel()) * fc.element_size(), L.lib.vlpet_k3_bwd, C.byref(desc), _p(fc),
                      _p(pc), _p(img), _p(dout), C.byref(w), _p(save), _p(dfeats), C.byref(g), _p(ws), ws.numel(), _stream()),
                "vlpet_k3_bwd")
        outg = []
        for gt, meta in zip(gv, ctx.param_meta):
            if meta is None:
                outg.append(None)
                continue
            shape, dtype = meta
            gt = gt.reshape(shape)
            outg.append(gt if dtype == torch.float32 else gt.to(dtype))
        return (None, None, None, dfeats, None, None, None, *outg, None)


def visual_projection(feats, pos, img_order_ids, obj_order_ids, Wf, bf, ln_f_w, ln_f_b, Wp, bp, ln_p_w, ln_p_b, E_img,
                      E_obj, rms: bool = False, eps: float = 1e-5, impl: str = "auto"):
    return VisProjFn.apply(bool(rms), float(eps), impl, feats, pos, img_order_ids, obj_order_ids, Wf, bf, ln_f_w, ln_f_b,
                           Wp, bp, ln_p_w, ln_p_b, E_img, E_obj.detach())


class LowRankVisProjFn(torch.autograd.Function):
    """LowRankVisualEmbedding.forward (src/modeling_bart.py:263-334) -- include/vlpet.h K3-LR.  Parameter order:
    Wd (row-concatenated heads), bd, Wu, bu, Gd, gbd, Gu, gbu (None x4 when not gated), ln_f_w, ln_f_b, Wp, bp, ln_p_w,
    ln_p_b, E_img, E_obj."""

    @staticmethod
    def forward(ctx, gated: bool, residual: bool, eps: float, feats, pos, img_ids, obj_ids, *params):
        _require_cuda(feats, pos)
        if feats.dim() != 3 or feats.dtype not in _DT or len(params) != 16:
            raise ValueError("vlpet.lowrank_visual_projection: feats must be [B, N, F] fp32/bf16 with 16 parameter slots")
        B, N, Fd = feats.shape
        dt = feats.dtype
        fc, pc = feats.contiguous(), pos.to(dt).contiguous()
        ids = lambda t: None if t is None else t.to(torch.int64).expand(B, N).contiguous()  # noqa: E731
        img, obj = ids(img_ids), ids(obj_ids)
        ws_ = [_as(t, dt) for t in params]
        d, r = ws_[2].shape[0], ws_[0].shape[0]
        rg = ws_[4].shape[0] if gated else 0
        desc = L.K3LRDesc(M=B * N, N=N, F=Fd, d=d, r=r, rg=rg, V=params[15].shape[0], n_img=params[14].shape[0],
                          gated=int(gated), residual=int(residual), dtype=_DT[dt], impl=L.IMPL_AUTO, eps=eps)
        w = L.K3LRParams(*[_p(t) for t in ws_])
        out = torch.empty(B, N, d, dtype=dt, device=feats.device)
        save = torch.empty(B * N * d, dtype=torch.float32, device=feats.device)
        ws = _workspace(L.lib.vlpet_k3lr_fwd_workspace_bytes(C.byref(desc)), feats.device)
        L.check(L.lib.vlpet_k3lr_fwd(C.byref(desc), _p(fc), _p(pc), _p(img), _p(obj), C.byref(w), _p(out), _p(save), _p(ws),
                                     ws.numel(), _stream()), "vlpet_k3lr_fwd")
        ctx.desc = desc
        ctx.present = [t is not None for t in ws_]
        ctx.param_meta = [None if t is None else (tuple(t.shape), t.dtype) for t in params]
        ctx.save_for_backward(fc, pc, img, save, *[t for t in ws_ if t is not None])
        return out

    @staticmethod
    def backward(ctx, dout):
        fc, pc, img, save, *rest = ctx.saved_tensors
        it = iter(rest)
        ws_ = [next(it) if p else None for p in ctx.present]
        desc = ctx.desc
        dout = dout.contiguous()
        if dout.dtype != fc.dtype:
            dout = dout.to(fc.dtype)
        sizes = [0 if m is None else int(torch.Size(m[0]).numel()) for m in ctx.param_meta[:15]]
        offs = [0]
        for n in sizes:
            offs.append(offs[-1] + (n + 3) // 4 * 4)
        gbuf = torch.zeros(max(offs[-1], 4), dtype=torch.float32, device=fc.device)
        gv = [gbuf[offs[i]:offs[i] + sizes[i]] if sizes[i] else None for i in range(15)]
        g = L.K3LRGrads(*[_p(t) for t in gv])
        w = L.K3LRParams(*[_p(t) for t in ws_])
        ws = _workspace(L.lib.vlpet_k3lr_bwd_workspace_bytes(C.byref(desc)), fc.device)
        L.check(L.lib.vlpet_k3lr_bwd(C.byref(desc), _p(fc), _p(pc), _p(img), _p(dout), C.byref(w), _p(save), C.byref(g), _p(ws),
                                     ws.numel(), _stream()), "vlpet_k3lr_bwd")
        outg = []
        for gt, meta in zip(gv, ctx.param_meta[:15]):
            if meta is None:
                outg.append(None)
            else:
                gt = gt.reshape(meta[0])
                outg.append(gt if meta[1] == torch.float32 else gt.to(meta[1]))
        return (None, None, None, None, None, None, None, *outg, None)


def lowrank_visual_projection(feats, pos, img_order_ids, obj_order_ids, params, gated: bool, residual: bool, eps: float = 1e-5):
    """params: the 16 tensors of LowRankVisProjFn (gate slots None when not gated).  Features are inputs: no dfeats."""
    params = list(params)
    params[15] = params[15].detach()
    return LowRankVisProjFn.apply(bool(gated), bool(residual), float(eps), feats, pos, img_order_ids, obj_order_ids, *params)


class LayerNormFn(torch.autograd.Function):
    """nn.LayerNorm on bf16 activations with fp32 affine parameters (include/vlpet.h vlpet_layernorm_fwd / _bwd)."""

    @staticmethod
    def forward(ctx, x, weight, bias, eps: float):
        _require_cuda(x, weight, bias)
        d = x.shape[-1]
        xc = x.contiguous()
        M = xc.numel() // d
        w32 = weight.detach() if weight.dtype == torch.float32 else weight.detach().float()
        b32 = bias.detach() if bias.dtype == torch.float32 else bias.detach().float()
        y = torch.empty_like(xc)
        stats = torch.empty(2, M, dtype=torch.float32, device=x.device)
        L.check(_call("ln_fwd", 2 * xc.numel() * xc.element_size(), L.lib.vlpet_layernorm_fwd, _p(xc), _p(w32.contiguous()),
                      _p(b32.contiguous()), _p(y), _p(stats[0]), _p(stats[1]), M, d, float(eps), _DT[xc.dtype], _stream()),
                "vlpet_layernorm_fwd")
        ctx.save_for_backward(xc, w32, stats)
        ctx.meta = (weight.dtype, bias.dtype)
        ctx.param_refs = (weight, bias)
        return y

    @staticmethod
    def backward(ctx, dy):
        xc, w32, stats = ctx.saved_tensors
        d = xc.shape[-1]
        M = xc.numel() // d
        dy = dy.contiguous()
        if dy.dtype != xc.dtype:
            dy = dy.to(xc.dtype)
        dx = torch.empty_like(xc)
        need_w, need_b = ctx.needs_input_grad[1], ctx.needs_input_grad[2]
        direct = _direct_targets(ctx.param_refs, [[0], [1]]) if (need_w and need_b) else None
        if direct is not None:
            L.check(_call("ln_bwd", 3 * xc.numel() * xc.element_size(), L.lib.vlpet_layernorm_bwd, _p(xc), _p(dy), _p(w32.contiguous()),
                          _p(stats[0]), _p(stats[1]), _p(dx), C.c_void_p(direct[0]), C.c_void_p(direct[1]), M, d, _DT[xc.dtype],
                          _stream()), "vlpet_layernorm_bwd")
            return dx, None, None, None
        gbuf = torch.zeros(2, d, dtype=torch.float32, device=xc.device) if (need_w or need_b) else None
        L.check(_call("ln_bwd", 3 * xc.numel() * xc.element_size(), L.lib.vlpet_layernorm_bwd, _p(xc), _p(dy), _p(w32.contiguous()),
                      _p(stats[0]), _p(stats[1]), _p(dx), _p(gbuf[0]) if need_w else C.c_void_p(0),
                      _p(gbuf[1]) if need_b else C.c_void_p(0), M, d, _DT[xc.dtype], _stream()), "vlpet_layernorm_bwd")
        dw = gbuf[0].to(ctx.meta[0]) if need_w else None
        db = gbuf[1].to(ctx.meta[1]) if need_b else None
        return dx, dw, db, None


def layer_norm(x, weight, bias, eps: float = 1e-5):
    """LayerNorm over the last dimension through the CUDA kernels (bf16 activations, d % 256 == 0, d <= 1024)."""
    return LayerNormFn.apply(x, weight, bias, eps)


class DropoutAddLayerNormFn(torch.autograd.Function):
    """LayerNorm(res + dropout_p(h)) in one pass each way (include/vlpet.h vlpet_dropout_add_layernorm_fwd / _bwd)."""

    @staticmethod
    def forward(ctx, h, res, weight, bias, eps: float, p: float, seed: int):
        _require_cuda(h, res, weight, bias)
        d = h.shape[-1]
        hc, rc = h.contiguous(), res.contiguous()
        M = hc.numel() // d
        w32 = weight.detach() if weight.dtype == torch.float32 else weight.detach().float()
        b32 = bias.detach() if bias.dtype == torch.float32 else bias.detach().float()
        y, xs = torch.empty_like(hc), torch.empty_like(hc)
        stats = torch.empty(2, M, dtype=torch.float32, device=h.device)
        sd = _seed_dev.data_ptr() if (seed and _seed_dev is not None) else None
        pd = float(p if seed else 0.0)
        L.check(_call("ln_fwd", 4 * hc.numel() * 2, L.lib.vlpet_dropout_add_layernorm_fwd, _p(hc), _p(rc), _p(w32.contiguous()),
                      _p(b32.contiguous()), _p(y), _p(xs), _p(stats[0]), _p(stats[1]), M, d, float(eps), pd, seed,
                      C.c_void_p(sd) if sd else C.c_void_p(0), _stream()), "vlpet_dropout_add_layernorm_fwd")
        ctx.save_for_backward(xs, w32, stats)
        ctx.args = (pd, seed, sd)
        ctx.meta = (weight.dtype, bias.dtype)
        ctx.param_refs = (weight, bias)
        return y

    @staticmethod
    def backward(ctx, dy):
        xs, w32, stats = ctx.saved_tensors
        pd, seed, sd = ctx.args
        d = xs.shape[-1]
        M = xs.numel() // d
        dy = dy.contiguous()
        if dy.dtype != xs.dtype:
            dy = dy.to(xs.dtype)
        dres, dh = torch.empty_like(xs), torch.empty_like(xs)
        need_w, need_b = ctx.needs_input_grad[2], ctx.needs_input_grad[3]
        direct = _direct_targets(ctx.param_refs, [[0], [1]]) if (need_w and need_b) else None
        gbuf = None
        if direct is not None:
            pw, pb = C.c_void_p(direct[0]), C.c_void_p(direct[1])
        else:
            gbuf = torch.zeros(2, d, dtype=torch.float32, device=xs.device) if (need_w or need_b) else None
            pw = _p(gbuf[0]) if need_w else C.c_void_p(0)
            pb = _p(gbuf[1]) if need_b else C.c_void_p(0)
        L.check(_call("ln_bwd", 4 * xs.numel() * 2, L.lib.vlpet_dropout_add_layernorm_bwd, _p(xs), _p(dy), _p(w32.contiguous()),
                      _p(stats[0]), _p(stats[1]), _p(dres), _p(dh), pw, pb, M, d, pd, seed, C.c_void_p(sd) if sd else C.c_void_p(0),
                      _stream()), "vlpet_dropout_add_layernorm_bwd")
        dw = gbuf[0].to(ctx.meta[0]) if (gbuf is not None and need_w) else None
        db = gbuf[1].to(ctx.meta[1]) if (gbuf is not None and need_b) else None
        return dh, dres, dw, db, None, None, None


def dropout_add_layer_norm(h, res, weight, bias, eps: float, p: float, training: bool):
    """LayerNorm(res + dropout(h, p)) of a post-LN residual step, one kernel forward and one backward (bf16 CUDA tensors)."""
    seed = next_dropout_seed() if (training and p > 0.0) else 0
    return DropoutAddLayerNormFn.apply(h, res, weight, bias, eps, p, seed)


def layer_norm_supported(x: torch.Tensor) -> bool:
    d = x.shape[-1]
    return x.is_cuda and x.dtype == torch.bfloat16 and d % 256 == 0 and 256 <= d <= 1024


class GeluDropoutFn(torch.autograd.Function):
    """dropout_p(gelu(x)) in one pass each way (include/vlpet.h vlpet_gelu_dropout_fwd / _bwd); bf16, exact erf GELU."""

    @staticmethod
    def forward(ctx, x, p: float, seed: int):
        _require_cuda(x)
        xc = x.contiguous()
        y = torch.empty_like(xc)
        sd = _seed_dev.data_ptr() if (seed and _seed_dev is not None) else None
        L.check(_call("gelu_drop_fwd", 2 * xc.numel() * 2, L.lib.vlpet_gelu_dropout_fwd, _p(xc), _p(y), xc.numel(), float(p if seed else 0.0),
                      seed, C.c_void_p(sd) if sd else C.c_void_p(0), _stream()), "vlpet_gelu_dropout_fwd")
        ctx.save_for_backward(xc)
        ctx.args = (float(p if seed else 0.0), seed, sd)
        return y

    @staticmethod
    def backward(ctx, dy):
        (xc,) = ctx.saved_tensors
        p, seed, sd = ctx.args
        dy = dy.contiguous()
        dx = torch.empty_like(xc)
        L.check(_call("gelu_drop_bwd", 3 * xc.numel() * 2, L.lib.vlpet_gelu_dropout_bwd, _p(xc), _p(dy), _p(dx), xc.numel(), p, seed,
                      C.c_void_p(sd) if sd else C.c_void_p(0), _stream()), "vlpet_gelu_dropout_bwd")
        return dx, None, None


def gelu_dropout(x: torch.Tensor, p: float, training: bool) -> torch.Tensor:
    """dropout(gelu(x), p) of the frozen FFN, one kernel forward and one backward (bf16 CUDA tensors, numel % 8 == 0)."""
    seed = next_dropout_seed() if (training and p > 0.0) else 0
    return GeluDropoutFn.apply(x, p, seed)


def gelu_dropout_supported(x: torch.Tensor) -> bool:
    return x.is_cuda and x.dtype == torch.bfloat16 and x.numel() % 8 == 0


class ShortAttentionFn(torch.autograd.Function):
    """softmax(q k^T / 8 [causal]) (dropout) v for head_dim 64 and L <= 128 (include/vlpet.h vlpet_attn_fwd / _bwd).
    q, k, v: [B, L, H*64] bf16 with unit inner stride and batch stride L * row stride (views of a fused projection ok)."""

    @staticmethod
    def forward(ctx, q, k, v, H: int, causal: bool, p: float, seed: int):
        _require_cuda(q, k, v)
        B, Lq, _ = q.shape
        Lk = k.shape[1]
        out = torch.empty(B, Lq, H * 64, dtype=q.dtype, device=q.device)
        lse = torch.empty(B, H, Lq, dtype=torch.float32, device=q.device)
        sd = _seed_dev.data_ptr() if (seed and _seed_dev is not None) else None
        pd = float(p if seed else 0.0)
        nb = 2 * (q.numel() + k.numel() + v.numel() + out.numel())
        L.check(_call("attn_fwd", nb, L.lib.vlpet_attn_fwd, _p(q), _p(k), _p(v), q.stride(1), k.stride(1), v.stride(1), _p(out), _p(lse),
                      B, H, Lq, Lk, int(causal), pd, seed, C.c_void_p(sd) if sd else C.c_void_p(0), _stream()), "vlpet_attn_fwd")
        ctx.save_for_backward(q, k, v, out, lse)
        ctx.args = (H, int(causal), pd, seed, sd)
        return out

    @staticmethod
    def backward(ctx, dout):
        q, k, v, out, lse = ctx.saved_tensors
        H, causal, pd, seed, sd = ctx.args
        B, Lq, _ = q.shape
        Lk = k.shape[1]
        dout = dout.contiguous()
        dq = torch.empty(B, Lq, H * 64, dtype=q.dtype, device=q.device)
        dk = torch.empty(B, Lk, H * 64, dtype=q.dtype, device=q.device)
        dv = torch.empty(B, Lk, H * 64, dtype=q.dtype, device=q.device)
        nb = 2 * (2 * q.numel() + 2 * k.numel() + 2 * v.numel() + 2 * out.numel())
        L.check(_call("attn_bwd", nb, L.lib.vlpet_attn_bwd, _p(q), _p(k), _p(v), q.stride(1), k.stride(1), v.stride(1), _p(out),
                      _p(dout), _p(lse), _p(dq), _p(dk), _p(dv), H * 64, H * 64, H * 64, B, H, Lq, Lk, causal, pd, seed,
                      C.c_void_p(sd) if sd else C.c_void_p(0), _stream()), "vlpet_attn_bwd")
        return dq, dk, dv, None, None, None, None


class ShortSelfAttentionFn(torch.autograd.Function):
    """Self-attention straight on the output of the one-GEMM q/k/v projection: qkv [B, L, 3, H*64] contiguous.  The
    backward writes dq / dk / dv into the thirds of ONE [B, L, 3, H*64] gradient (no concatenation pass)."""

    @staticmethod
    def forward(ctx, qkv, H: int, causal: bool, p: float, seed: int):
        _require_cuda(qkv)
        B, Lq, _, d = qkv.shape
        q, k, v = qkv.unbind(2)
        out = torch.empty(B, Lq, d, dtype=qkv.dtype, device=qkv.device)
        lse = torch.empty(B, H, Lq, dtype=torch.float32, device=qkv.device)
        sd = _seed_dev.data_ptr() if (seed and _seed_dev is not None) else None
        pd = float(p if seed else 0.0)
        L.check(_call("attn_fwd", 2 * (qkv.numel() + out.numel()), L.lib.vlpet_attn_fwd, _p(q), _p(k), _p(v), 3 * d, 3 * d, 3 * d,
                      _p(out), _p(lse), B, H, Lq, Lq, int(causal), pd, seed, C.c_void_p(sd) if sd else C.c_void_p(0), _stream()),
                "vlpet_attn_fwd")
        ctx.save_for_backward(qkv, out, lse)
        ctx.args = (H, int(causal), pd, seed, sd)
        return out

    @staticmethod
    def backward(ctx, dout):
        qkv, out, lse = ctx.saved_tensors
        H, causal, pd, seed, sd = ctx.args
        B, Lq, _, d = qkv.shape
        q, k, v = qkv.unbind(2)
        dout = dout.contiguous()
        dqkv = torch.empty_like(qkv)
        dq, dk, dv = dqkv.unbind(2)
        L.check(_call("attn_bwd", 2 * (2 * qkv.numel() + 2 * out.numel()), L.lib.vlpet_attn_bwd, _p(q), _p(k), _p(v), 3 * d, 3 * d,
                      3 * d, _p(out), _p(dout), _p(lse), _p(dq), _p(dk), _p(dv), 3 * d, 3 * d, 3 * d, B, H, Lq, Lq, causal, pd, seed,
                      C.c_void_p(sd) if sd else C.c_void_p(0), _stream()), "vlpet_attn_bwd")
        return dqkv, None, None, None, None


def short_self_attention(qkv, H: int, causal: bool, p: float, training: bool) -> torch.Tensor:
    """qkv [B, L, 3, H*64] (fused projection) -> [B, L, H*64]."""
    seed = next_dropout_seed() if (training and p > 0.0) else 0
    return ShortSelfAttentionFn.apply(qkv, H, causal, p, seed)


def short_attention_supported(q, k, v, H: int) -> bool:
    def ok(t):
        return (t.is_cuda and t.dtype == torch.bfloat16 and t.dim() == 3 and t.shape[2] == H * 64 and t.stride(2) == 1
                and t.stride(1) % 8 == 0 and t.stride(0) == t.shape[1] * t.stride(1) and t.data_ptr() % 16 == 0)
    return ok(q) and ok(k) and ok(v) and q.shape[1] <= 128 and k.shape[1] <= 128 and k.shape[1] == v.shape[1] \
        and q.shape[0] == k.shape[0]


def short_attention_profitable(q, k) -> bool:
    """L <= 64 runs the register-resident kernels (1.7-2.4x faster than torch SDPA forward + backward on B200 at the
    workload's shapes, tools/attn_probe.py); 65..128 runs the shared-memory (wmma) kernels, which are slower than SDPA."""
    return q.shape[1] <= 64 and k.shape[1] <= 64


def short_attention(q, k, v, H: int, causal: bool, p: float, training: bool) -> torch.Tensor:
    """[B, Lq, H*64] x [B, Lk, H*64]^2 -> [B, Lq, H*64]: attention of the frozen BART blocks in one launch each way."""
    seed = next_dropout_seed() if (training and p > 0.0) else 0
    return ShortAttentionFn.apply(q, k, v, H, causal, p, seed)


class CrossEntropyBf16Fn(torch.autograd.Function):
    """Per-token cross-entropy (reduction='none') straight from bf16 logits (include/vlpet.h vlpet_ce_fwd / _bwd):
    no fp32 copy of the [tokens, vocab] logits, one pass forward, one pass backward."""

    @staticmethod
    def forward(ctx, logits2d, labels, ignore_index: int):
        _require_cuda(logits2d, labels)
        rows, ncols = logits2d.shape
        ld = logits2d.stride(0)
        lab = labels.reshape(-1).contiguous()
        loss = torch.empty(rows, dtype=torch.float32, device=logits2d.device)
        lse = torch.empty(rows, dtype=torch.float32, device=logits2d.device)
        L.check(_call("ce_fwd", rows * ncols * 2, L.lib.vlpet_ce_fwd, _p(logits2d), ld, _p(lab), _p(loss), _p(lse), rows, ncols,
                      int(ignore_index), _stream()), "vlpet_ce_fwd")
        ctx.save_for_backward(logits2d, lab, lse)
        ctx.ignore_index = int(ignore_index)
        return loss

    @staticmethod
    def backward(ctx, dloss):
        logits2d, lab, lse = ctx.saved_tensors
        rows, ncols = logits2d.shape
        ld = logits2d.stride(0)
        dl = dloss.contiguous().float()
        dlogits = torch.empty_strided(logits2d.shape, logits2d.stride(), dtype=logits2d.dtype, device=logits2d.device)
        L.check(_call("ce_bwd", 2 * rows * ncols * 2, L.lib.vlpet_ce_bwd, _p(logits2d), ld, _p(lab), _p(lse), _p(dl), _p(dlogits),
                      rows, ncols, ctx.ignore_index, _stream()), "vlpet_ce_bwd")
        return dlogits, None, None


def cross_entropy_supported(logits2d: torch.Tensor) -> bool:
    return (logits2d.is_cuda and logits2d.dtype == torch.bfloat16 and logits2d.dim() == 2 and logits2d.stride(1) == 1
            and logits2d.shape[1] % 8 == 0 and logits2d.stride(0) % 8 == 0 and logits2d.stride(0) >= logits2d.shape[1]
            and logits2d.data_ptr() % 16 == 0)


def cross_entropy_bf16(logits2d: torch.Tensor, labels: torch.Tensor, ignore_index: int = -100) -> torch.Tensor:
    """F.cross_entropy(logits.float(), labels, ignore_index, reduction='none') without the fp32 logits."""
    return CrossEntropyBf16Fn.apply(logits2d, labels, ignore_index)


def grid_maxpool(feats: torch.Tensor, out_size: int, out_dtype: Optional[torch.dtype] = None) -> torch.Tensor:
    """[B, g*g, F] CLIP grid features -> [B, o*o, F] by adaptive max-pool (src/modeling_bart.py:556-613 Downsample),
    fused with the cast to ``out_dtype`` (include/vlpet.h vlpet_grid_maxpool).  Inputs are data: no autograd."""
    _require_cuda(feats)
    if feats.dim() != 3 or feats.dtype not in _DT:
        raise ValueError("vlpet.grid_maxpool: feats must be [B, g*g, F] fp32/bf16")
    B, G, Fd = feats.shape
    g = int(round(G ** 0.5))
    if g * g != G:
        raise ValueError(f"vlpet.grid_maxpool: {G} grid cells is not a square")
    out_dtype = out_dtype or feats.dtype
    fc = feats.detach().contiguous()
    out = torch.empty(B, out_size * out_size, Fd, dtype=out_dtype, device=feats.device)
    L.check(L.lib.vlpet_grid_maxpool(_p(fc), _DT[fc.dtype], _p(out), _DT[out_dtype], B, g, out_size, Fd, _stream()),
            "vlpet_grid_maxpool")
    return out


def fwd_is_fused(M: int, d: int, r: int, rg: int, dtype=torch.bfloat16, gate: str = "large") -> bool:
    desc = L.K1Desc(M=M, L=0, d=d, r=r, rg=rg, gate=L.GATE_IDS[gate], add_gate=0, dtype=_DT[dtype], impl=L.IMPL_AUTO,
                    s=1.0, alpha=1.0, kappa=1.0, p_drop=0.0, seed=0)
    return bool(L.lib.vlpet_k1_fwd_is_fused(C.byref(desc)))
