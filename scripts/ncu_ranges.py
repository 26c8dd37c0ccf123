"""Instruction / sample share of line ranges. python scripts/ncu_ranges.py REP KERNEL 'name:file:lo:hi' ..."""
import csv, io, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
ranges = [a.split(':') for a in sys.argv[3:]]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kern, "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
cur = None; agg = []
for r in csv.reader(io.StringIO(out)):
  if not r: continue
  if r[0] == "File Path": cur = r[1].split('/')[-1]; continue
  if r[0] == "Function Name": continue
  if r[0] == "Line No":
    hdr = r; ci = hdr.index('Instructions Executed'); cs = hdr.index('# Samples'); ct = hdr.index('Thread Instructions Executed'); continue
  if r[0] not in ("", "..."):
    try: agg.append((cur, int(r[0]), float(r[ci]), float(r[cs]), float(r[ct])))
    except ValueError: continue
tot = sum(a[2] for a in agg); tots = sum(a[3] for a in agg)
acc = {}
for a in agg:
  k = "other:" + a[0]
  for nm, f, lo, hi in ranges:
    if f in a[0] and int(lo) <= a[1] <= int(hi): k = nm; break
  x = acc.setdefault(k, [0, 0, 0]); x[0] += a[2]; x[1] += a[3]; x[2] += a[4]
print(f"total warp inst {tot:.4g}")
for k, v in sorted(acc.items(), key=lambda kv: -kv[1][0]):
  print(f"{k:28} inst {100*v[0]/tot:5.1f}% ({v[0]/1e6:6.1f} M)  samples {100*v[1]/tots:5.1f}%  act {v[2]/max(v[0],1):4.1f}")
