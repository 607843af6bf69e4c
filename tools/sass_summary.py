"""Per-kernel counts of the SASS mnemonics that prove the Blackwell-native claims (B200_PROFILING.md evidence table):
UTCHMMA (tcgen05.mma), LDTM / STTM (tcgen05.ld / st), UTMALDG (TMA tensor loads), UBLKCP (cp.async.bulk), UTCBAR
(tcgen05.commit -> mbarrier), SYNCS (mbarrier ops), FFMA2 (packed fp32 FMA), HMMA (legacy mma.sync), LDL/STL (spills).

  python tools/sass_summary.py xotorch_b200/lib/libxotb200.so > profiles/r2_sass_summary.txt"""
import collections
import re
import subprocess
import sys

OPS = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "SYNCS", "FFMA2", "HMMA", "LDGSTS", "LDL", "STL",
       "REDG", "ATOMG", "MEMBAR", "ERRBAR", "CCTL"]


def main(path):
  out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True, check=True).stdout
  per, cur, arch = collections.OrderedDict(), None, None
  for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
      cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
      per[cur] = collections.Counter()
      continue
    m = re.search(r"arch = (sm_\w+)", line)
    if m:
      arch = m.group(1)
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and cur:
      op = m.group(1)
      per[cur]["_total"] += 1
      for o in OPS:
        if op == o or op.startswith(o + "."):
          per[cur][o] += 1
  print(f"# {path}  ({arch}): SASS instruction counts per kernel (cuobjdump -sass | tools/sass_summary.py)")
  print(f"{'kernel':100s} {'instrs':>7s} " + " ".join(f"{o:>7s}" for o in OPS))
  for k, c in per.items():
    if c["_total"] < 50:
      continue
    print(f"{k[:100]:100s} {c['_total']:7d} " + " ".join(f"{c[o]:7d}" for o in OPS))


if __name__ == "__main__":
  main(sys.argv[1] if len(sys.argv) > 1 else "xotorch_b200/lib/libxotb200.so")
