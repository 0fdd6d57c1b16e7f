"""Locates and imports the UNMODIFIED reference (autonomousvision/gta): $GTA_REF, /root/reference (build container) or
baseline/_ref (GPU box; produced by baseline/install_ref.py).  Not imported by the product package.

Import quirks handled without editing the tree: `J_dense.pt` is loaded from a CWD-relative path at
import time (source/utils/wigner_d.py:8-9), and source.encoder / source.decoder import a symbol `ray2rotation` that is
not defined at this commit (only used under the `ray_to_se3` flag, which no shipped config sets).
"""
from __future__ import annotations

import contextlib
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
_mods = None


def root() -> str | None:
    for cand in (os.environ.get("GTA_REF"), "/root/reference", os.path.join(HERE, "_ref")):
        if cand and os.path.isfile(os.path.join(cand, "source", "utils", "gta.py")):
            return cand
    return None


def available() -> bool:
    return root() is not None


@contextlib.contextmanager
def _cwd(path):
    old = os.getcwd()
    os.chdir(path)
    try:
        yield
    finally:
        os.chdir(old)


def load():
    """Import the reference modules once -> namespace(gta, wigner_d, layers, encoder, decoder, models_nvs, root)."""
    global _mods
    if _mods is not None:
        return _mods
    r = root()
    if r is None:
        raise RuntimeError("reference tree not found ($GTA_REF, /root/reference, baseline/_ref)")
    sys.dont_write_bytecode = True
    with _cwd(r):
        sys.path.insert(0, r)
        try:
            import source.utils.gta as rgta
            if not hasattr(rgta, "ray2rotation"):
                def _stub(*a, **k):
                    raise NotImplementedError("ray2rotation is undefined in the reference")
                rgta.ray2rotation = _stub
            import source.utils.wigner_d as rwig
            import source.layers as rlay
            import source.encoder as renc
            import source.decoder as rdec
            import source.models_nvs as rmod
        finally:
            sys.path.remove(r)
    _mods = types.SimpleNamespace(gta=rgta, wigner_d=rwig, layers=rlay, encoder=renc, decoder=rdec, models_nvs=rmod, root=r)
    return _mods


def config(run: str) -> dict:
    """runs/<run>/config.yaml of the reference, e.g. 'msn/GTA/gta_so3'."""
    import yaml
    with open(os.path.join(root(), "runs", run, "config.yaml")) as f:
        return yaml.safe_load(f)
