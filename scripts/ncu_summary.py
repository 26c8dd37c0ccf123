"""Summarise an .ncu-rep (read here with `ncu -i`) into a compact per-kernel table (markdown)."""
import csv, io, subprocess, sys
rep = sys.argv[1]
METRICS = [
  ("gpu__time_duration.sum", "time"),
  ("dram__bytes_read.sum", "dram_rd"),
  ("dram__bytes_write.sum", "dram_wr"),
  ("dram__cycles_active.avg.pct_of_peak_sustained_elapsed", "dram_active%"),
  ("dram__bytes.sum.per_second", "dram_bw"),
  ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
  ("smsp__inst_executed.sum", "warp_inst"),
  ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
  ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
  ("launch__registers_per_thread", "regs"),
  ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "st_lsb/iss"),
]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}
print("| kernel | " + " | ".join(n for _, n in METRICS) + " |")
print("|---|" + "---|" * len(METRICS))
for r in rows[2:]:
  name = r[col["Kernel Name"]].split("(")[0].replace("void ", "")
  cells = []
  for m, _ in METRICS:
    if m in col:
      v = r[col[m]]
      u = units[col[m]]
      try:
        f = float(v.replace(",", ""))
        v = f"{f:.3g}"
      except ValueError:
        pass
      cells.append(f"{v} {u}".strip())
    else:
      cells.append("-")
  print(f"| {name} | " + " | ".join(cells) + " |")
