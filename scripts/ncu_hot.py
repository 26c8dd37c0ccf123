"""Hot-spot summary of one kernel from `ncu -i rep --page source --csv --kernel-name regex:NAME > file.csv`."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ci = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) > ci["Instructions Executed"] and r[0].startswith("0x")]
I = lambda r, k: int(r[ci[k]] or 0)
tot_inst = sum(I(r, "Instructions Executed") for r in data)
tot_samp = sum(I(r, "# Samples") for r in data)
print("total warp inst", tot_inst, "samples", tot_samp, "n sass", len(data))
acc = samp = seg = 0
for r in data:
  acc += I(r, "Instructions Executed"); samp += I(r, "# Samples")
  if "BAR.SYNC" in r[ci["Source"]]:
    print(f"  segment {seg}: inst {acc} ({100*acc/tot_inst:.1f}%) samples {samp} ({100*samp/max(1,tot_samp):.1f}%)"); seg += 1; acc = samp = 0
print(f"  segment {seg}: inst {acc} ({100*acc/tot_inst:.1f}%) samples {samp} ({100*samp/max(1,tot_samp):.1f}%)")
n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
print("top by samples: samples, warp-inst, avg-threads, sass")
for r in sorted(data, key=lambda r: -I(r, "# Samples"))[:n]:
  print(f"  {I(r,'# Samples'):6d} {I(r,'Instructions Executed'):10d} {r[ci['Avg. Threads Executed']]:>5} {r[ci['Source']].strip()[:100]}")
