"""
Import the REAL reference modules from /root/reference (build container only -- the GPU box has no
/root/reference).  Used by tools/make_golden_*.py to generate tests/golden/ fixtures and by the
optional differential tests (skipped when the reference tree is absent).

Recipe: bypass src/__init__.py (it imports ir_measures etc.), stub `ranx`
(not installed; only the weighted functions are exercised), and point PROJECT_ROOT at a writable dir
because src/config.py:202-203 mkdirs ${PROJECT_ROOT}/cache at import time.
"""
import os
import sys
import tempfile
import types

REFERENCE_ROOT = "/root/reference"


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "src", "fusion.py"))


def import_reference_fusion():
    if not reference_available():
        raise RuntimeError("reference tree not present")
    os.environ.setdefault("PROJECT_ROOT", tempfile.mkdtemp(prefix="qf_ref_"))
    if "src" not in sys.modules or getattr(sys.modules["src"], "__path__", None) != [REFERENCE_ROOT + "/src"]:
        pkg = types.ModuleType("src")
        pkg.__path__ = [REFERENCE_ROOT + "/src"]
        sys.modules["src"] = pkg
    if "ranx" not in sys.modules:
        stub = types.ModuleType("ranx")
        stub.Run = object
        stub.Qrels = object
        stub.fuse = lambda *a, **k: (_ for _ in ()).throw(RuntimeError("ranx is stubbed"))
        sys.modules["ranx"] = stub
    from src import fusion  # type: ignore
    return fusion
