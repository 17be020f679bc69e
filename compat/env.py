"""Paths and environment for running the unmodified reference (baseline/_ref, staged by __graft_entry__.build()) in this image."""
from __future__ import annotations

import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIMS = os.path.join(ROOT, "compat", "shims")
DROPIN = os.path.join(ROOT, "compat", "dropin")
REF_TULIP = os.path.join(ROOT, "baseline", "_ref", "tulip")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REF_TULIP, "model", "tulip.py"))


def pythonpath(dropin: bool) -> str:
    """PYTHONPATH for a driver subprocess: the reference's own `tulip/` directory is the script directory (sys.path[0])."""
    parts = ([DROPIN] if dropin else []) + [SHIMS, ROOT]
    return os.pathsep.join(parts)


def driver_env(dropin: bool) -> dict:
    env = dict(os.environ)
    env["PYTHONPATH"] = pythonpath(dropin) + (os.pathsep + env["PYTHONPATH"] if env.get("PYTHONPATH") else "")
    env["WANDB_MODE"] = "disabled"
    env["TORCH_FORCE_NO_WEIGHTS_ONLY_LOAD"] = "1"      # misc.load_model (util/misc.py:366) loads a checkpoint that pickles `args`
    return env


def import_reference_model():
    """In-process import of the reference's `model.tulip` (for bench.py's reference arms).  Returns the module."""
    if not reference_available():
        raise FileNotFoundError(f"{REF_TULIP} missing: run `python -c 'import __graft_entry__ as g; g.build()'` in the build container")
    if SHIMS not in sys.path:
        sys.path.insert(0, SHIMS)
    if not any(type(f).__name__ == "_TorchSix" for f in sys.meta_path):      # interpreter did not start with SHIMS on its path
        import importlib.util
        spec = importlib.util.spec_from_file_location("_tulip_compat_sitecustomize", os.path.join(SHIMS, "sitecustomize.py"))
        spec.loader.exec_module(importlib.util.module_from_spec(spec))
    if REF_TULIP not in sys.path:
        sys.path.insert(0, REF_TULIP)
    import model.tulip as T
    if not os.path.abspath(T.__file__).startswith(os.path.abspath(REF_TULIP)):
        raise RuntimeError(f"`model.tulip` resolved to {T.__file__}, not to the staged reference")
    return T
