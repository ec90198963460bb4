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
