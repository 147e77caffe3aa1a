"""MANO asset loading without chumpy.

The reference unpickles ``MANO_LEFT/RIGHT.pkl`` through chumpy (mano/manolayer.py:350-394), which
is neither installable here nor needed: only the raw arrays are used.  ``load_mano_pkl`` reads the
official pickles with a stub for every ``chumpy.*`` class and pulls the ndarray out of each;
``synthetic`` assets (acr_b200.synth.make_synthetic_mano) have the same keys and shapes.
"""
from __future__ import annotations

import os
import pickle
from typing import Dict

import numpy as np

KEYS = ("hands_components", "hands_mean", "shapedirs", "posedirs", "v_template", "J_regressor", "weights", "f",
        "kintree_table")


class _ChStub:
    """Stands in for chumpy objects; keeps whatever state the pickle restores."""

    def __init__(self, *a, **k):
        pass

    def __setstate__(self, state):
        self.__dict__.update(state if isinstance(state, dict) else {"state": state})

    def array(self) -> np.ndarray:
        for k in ("x", "_x", "r", "a"):
            v = self.__dict__.get(k)
            if v is not None:
                return np.asarray(v.array() if isinstance(v, _ChStub) else v)
        raise ValueError("cannot extract an array from a chumpy object")


class _Unpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if module.split(".")[0] == "chumpy":
            return _ChStub
        return super().find_class(module, name)


def _arr(v) -> np.ndarray:
    if isinstance(v, _ChStub):
        return v.array()
    if hasattr(v, "toarray"):
        return np.asarray(v.toarray())
    return np.asarray(v)


def load_mano_pkl(path: str) -> Dict[str, np.ndarray]:
    with open(path, "rb") as f:
        dd = _Unpickler(f, encoding="latin1").load()
    out = {k: _arr(dd[k]) for k in KEYS if k in dd}
    out["betas"] = np.zeros(10, np.float32)
    for k in ("hands_components", "hands_mean", "shapedirs", "posedirs", "v_template", "J_regressor", "weights"):
        out[k] = np.ascontiguousarray(out[k], np.float32)
    out["f"] = np.asarray(out["f"]).astype(np.int64)
    return out


def get_asset(mano_root: str, side: str) -> Dict[str, np.ndarray]:
    """Real pickle if present under ``mano_root``; seeded synthetic asset only when explicitly
    allowed through ACR_B200_SYNTHETIC_MANO=1 (tests / benchmark)."""
    path = os.path.join(mano_root, "MANO_RIGHT.pkl" if side == "right" else "MANO_LEFT.pkl")
    if os.path.exists(path):
        return load_mano_pkl(path)
    if os.environ.get("ACR_B200_SYNTHETIC_MANO", "0") == "1":
        from acr_b200.synth import make_synthetic_mano
        return make_synthetic_mano(side)
    raise FileNotFoundError(f"{path} not found (licence-gated, see the reference README); set "
                            "ACR_B200_SYNTHETIC_MANO=1 to run with the seeded synthetic hand model")
