"""TEST INFRASTRUCTURE: builds the oracle's C helpers (gcc only).  /root/reference is pure Python and its hot-path
arithmetic lives in an absent third-party package (torchtune), so there is nothing to compile into oracle/_ref —
see DESIGN.md "Oracle"."""
import subprocess
from pathlib import Path

HERE = Path(__file__).resolve().parent
OUT = HERE / "_build" / "libsynthfast.so"


def build(force: bool = False) -> Path:
  src = HERE / "synth_fast.c"
  if force or not OUT.exists() or OUT.stat().st_mtime < src.stat().st_mtime:
    OUT.parent.mkdir(exist_ok=True)
    subprocess.check_call(["gcc", "-O3", "-fopenmp", "-ffp-contract=off", "-shared", "-fPIC", "-o", str(OUT), str(src)])
  return OUT


if __name__ == "__main__":
  print(build(True))
