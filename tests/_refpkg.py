"""Import the UNMODIFIED reference package (orchestration / topology / networking — everything around the hot path) for the
drop-in tests: from `baseline/_ref` (pip-installed once from /root/reference with --no-deps, git-ignored, travels to the GPU
box) or, in the build container, from /root/reference itself.  `scapy` is stubbed (xotorch/helpers.py imports it at module
top and it is not installed); the CWD is moved because device_capabilities.py writes a log file there."""
import sys
import types
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def reference_root():
  for p in (ROOT / "baseline" / "_ref", Path("/root/reference")):
    if (p / "xotorch" / "orchestration" / "node.py").exists():
      return p
  return None


def import_reference(monkeypatch, tmp_path):
  root = reference_root()
  assert root is not None
  stub, stub_all = types.ModuleType("scapy"), types.ModuleType("scapy.all")
  stub_all.get_if_addr = lambda *a, **k: "127.0.0.1"
  stub_all.get_if_list = lambda *a, **k: []
  monkeypatch.setitem(sys.modules, "scapy", stub)
  monkeypatch.setitem(sys.modules, "scapy.all", stub_all)
  monkeypatch.syspath_prepend(str(root))
  monkeypatch.chdir(tmp_path)
  monkeypatch.setattr(sys, "dont_write_bytecode", True)
  return root
