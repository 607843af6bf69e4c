"""Turn ncu output into the summaries kept under profiles/.

  launch list  : ncu --metrics gpu__time_duration.sum --clock-control none -c N --csv --log-file L.csv <cmd>
                 python tools/ncu_summarize.py launches L.csv > profiles/<round>_launch_list_summary.txt
  full capture : ncu --set full --clock-control none --import-source on -k regex:<kernels> -o R <cmd>
                 ncu -i R.ncu-rep --page raw --csv > R.csv ; python tools/ncu_summarize.py full R.csv > profiles/<round>_ncu_full_metrics.txt
(per-launch times under ncu are cold-cache and serialised: read the SHARES, not the absolutes — B200_PROFILING.md)"""
import collections
import csv
import re
import sys

FULL_KEYS = [
  "gpu__time_duration.sum", "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
  "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
  "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
  "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__occupancy_limit_shared_mem",
  "sm__warps_active.avg.pct_of_peak_sustained_active",
]


def launches(path):
  with open(path) as f:
    rows = csv.DictReader([l for l in f if not l.startswith("==")])
    agg = collections.OrderedDict()
    for r in rows:
      name, val = r.get("Kernel Name"), r.get("Metric Value")
      if not name or not val:
        continue
      t = float(val.replace(",", "")) * {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(r.get("Metric Unit", "ns"), 1.0)
      a = agg.setdefault(re.sub(r"\(.*", "", name), [0, 0.0])
      a[0] += 1
      a[1] += t
  tot = sum(a[1] for a in agg.values())
  print(f"# {path}: launches and gpu__time_duration.sum per kernel (ncu, --clock-control none)")
  for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{t / 1e3:12.1f} us  {100 * t / tot:5.1f} %  x{n:<5d} {t / n / 1e3:10.2f} us/launch  {k}")
  print(f"total {tot / 1e6:.3f} ms over {sum(a[0] for a in agg.values())} launches")


def full(path):
  r = list(csv.reader(open(path)))
  h, units = r[0], r[1]
  for row in r[2:]:
    print(row[h.index("Kernel Name")].split("(")[0], " grid", row[h.index("Grid Size")], "block", row[h.index("Block Size")])
    for k in FULL_KEYS:
      if k in h:
        print(f"    {k:95s} {row[h.index(k)]} {units[h.index(k)]}")


if __name__ == "__main__":
  {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
