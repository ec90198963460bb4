"""Shared test helpers: golden loading and oracle-parameter plumbing (test infrastructure)."""
import glob
import os

import numpy as np

from oracle import pet_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
K1_PARAM_KEYS = ("Wd", "bd", "Wu", "bu", "Gd", "gbd", "Gu", "gbu", "gw", "gb", "gz")


def golden_files(prefix):
    return sorted(glob.glob(os.path.join(GOLDEN, prefix + "*.npz")))


def load(path):
    z = np.load(path, allow_pickle=False)
    return {k: z[k] for k in z.files}


def k1_case(g):
    """-> (x1[M,d], x2[M,d], dout[M,d], params, cfg) from a k1 golden record."""
    d = int(g["meta_d"])
    cfg = O.PetConfig(gate=str(g["meta_gate"]), add_gate=bool(int(g["meta_add_gate"])), s=float(g["meta_s"]),
                      alpha=float(g["meta_alpha"]), kappa=float(g["meta_kappa"]), seq_len=int(g["meta_L"]))
    p = {k: (g[k].reshape(()) if k == "gb" else g[k]) for k in K1_PARAM_KEYS if k in g}
    return g["x1"].reshape(-1, d), g["x2"].reshape(-1, d), g["dout"].reshape(-1, d), p, cfg


def rel(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def bf16_round(a):
    import torch
    return torch.tensor(np.asarray(a), dtype=torch.float32).to(torch.bfloat16).to(torch.float64).numpy()


def bf16_check(ours, ref, tol, outlier=5.0):
    """Parity of a bf16-typed result `ours` (as float64) against the unrounded fp64 oracle `ref`.

    Storing a value in bf16 costs ~1.6e-3 rms relative by itself, so a 1e-3 bar can only be about the value BEFORE
    the final rounding.  That error delta is observable through the rounding flips: ours != round_bf16(ref) happens
    only when a rounding boundary lies between ref and ref+delta (so |ref - boundary| <= |delta|), and for a boundary
    uniformly placed inside that interval E|ref - boundary| = E|delta| / 2.  We therefore require
      (1) every mismatch is between ADJACENT bf16 values, or (near-zero outputs) smaller than outlier*tol*rms(ref) -- no
          outliers (outlier = 5; the tensor-core backward uses 20: a one-ulp flip of a bf16-STORED intermediate such as
          da (|da| ~ 2 -> ulp 0.016) times |Wd| ~ 0.05 legitimately moves a whole row of dx2 by ~1e-3),
      (2) the estimated mean pre-rounding error  2 * mean_mismatch |ref - boundary| / rms(ref)  <= tol,
      (3) Frobenius error vs the correctly-rounded oracle <= 2 * tol  (a flip costs a whole ulp, which amplifies
          delta to ~sqrt(delta * ulp); two exact-to-1e-4 bf16 kernels disagree at the 1e-3 level for this reason).
    Returns (estimated_pre_rounding_error, mismatch_fraction) for reporting."""
    ours = np.asarray(ours, dtype=np.float64).reshape(-1)
    ref = np.asarray(ref, dtype=np.float64).reshape(-1)
    assert np.all(np.isfinite(ours))
    r = bf16_round(ref).reshape(-1)
    rms = float(np.sqrt(np.mean(ref * ref))) + 1e-30
    mism = ours != r
    if not mism.any():
        return 0.0, 0.0
    a, b = ours[mism], r[mism]
    big = np.maximum(np.abs(a), np.abs(b))
    ulp = 2.0 ** (np.floor(np.log2(np.maximum(big, 1e-38))) - 7)
    diff = np.abs(a - b)
    assert np.all((diff <= ulp * 1.0001) | (diff <= outlier * tol * rms)), \
        "bf16 outlier: %g ulp / %g of rms" % (float(np.max(diff / ulp)), float(np.max(diff) / rms))
    eff = 2.0 * float(np.mean(np.abs(ref[mism] - 0.5 * (a + b)))) / rms
    assert eff <= tol, "estimated pre-rounding error %g exceeds %g (relative to rms)" % (eff, tol)
    assert rel(ours, r) <= 2 * tol
    return eff, float(mism.mean())
