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


def traffic(path, tokens, command):
  """profiles/r2_decode_ncu_traffic.json (what bench.py reports as roofline.traffic): DRAM bytes per decoded token of the
  persistent decode kernel = (dram__bytes_read.sum + dram__bytes_write.sum) of ONE launch / the tokens that launch decoded."""
  import json
  r = list(csv.reader(open(path)))
  h, units = r[0], r[1]
  row = [x for x in r[2:] if "shard_step_kernel" in x[h.index("Kernel Name")]][0]

  def val(k):
    v, u = float(row[h.index(k)].replace(",", "")), units[h.index(k)]
    return v * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(u, 1.0)
  rd, wr = val("dram__bytes_read.sum"), val("dram__bytes_write.sum")
  dur = float(row[h.index("gpu__time_duration.sum")].replace(",", "")) * {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}[units[h.index("gpu__time_duration.sum")]]
  out = {"bytes_per_token": int((rd + wr) / tokens), "dram_read_bytes": int(rd), "dram_write_bytes": int(wr), "tokens_in_launch": tokens,
         "kernel": row[h.index("Kernel Name")].split("(")[0], "duration_ms_under_ncu": round(dur * 1e3, 3), "command": command,
         "note": f"bytes per token = (dram__bytes_read.sum + dram__bytes_write.sum) / {tokens} tokens of one persistent-kernel launch, ncu --set full "
                 f"--clock-control none of the committed build (profiles/r2_decode_ncu_full_metrics.txt)"}
  print(json.dumps(out, indent=1))


if __name__ == "__main__":
  if sys.argv[1] == "traffic":
    traffic(sys.argv[2], int(sys.argv[3]), sys.argv[4] if len(sys.argv) > 4 else "")
  else:
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
