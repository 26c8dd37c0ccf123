"""Per-source-line summary of one kernel of an .ncu-rep (read here with `ncu -i`): share of warp instructions, of the
stall samples, average active threads. python scripts/ncu_lines.py REP KERNEL_REGEX [min_pct]"""
import csv, io, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
thr = float(sys.argv[3]) if len(sys.argv) > 3 else 0.6
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kern, "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
cur = None; agg = []
for r in csv.reader(io.StringIO(out)):
  if not r: continue
  if r[0] == "File Path": cur = r[1].split('/')[-1]; continue
  if r[0] == "Function Name": continue
  if r[0] == "Line No":
    hdr = r; ci = hdr.index('Instructions Executed'); cs = hdr.index('# Samples'); ct = hdr.index('Thread Instructions Executed'); cb = hdr.index('stall_barrier')
    stalls = [(i, h) for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
    continue
  if r[0] not in ("", "..."):
    try: n = float(r[ci]); s = float(r[cs]); t = float(r[ct])
    except ValueError: continue
    agg.append((cur, int(r[0]), r[1], n, s, t, [float(r[i] or 0) for i, _ in stalls]))
tot = sum(a[3] for a in agg); tots = sum(a[4] for a in agg)
print(f"total warp inst {tot:.4g} samples {tots:.0f}")
st = [sum(a[6][k] for a in agg) for k in range(len(stalls))]
print("stalls: " + " ".join(f"{h[6:]}={100*v/max(1,sum(st)):.1f}%" for (_, h), v in sorted(zip(stalls, st), key=lambda x: -x[1]) if v / max(1, sum(st)) > 0.01))
for a in agg:
  if a[3] / tot * 100 > thr or a[4] / max(1, tots) * 100 > thr:
    top = max(range(len(stalls)), key=lambda k: a[6][k])
    print(f"{a[0][5:14]:9}{a[1]:>5} inst {100*a[3]/tot:5.1f}% smp {100*a[4]/max(1,tots):5.1f}% act {a[5]/max(a[3],1):4.1f} {stalls[top][1][6:]:>10} | {a[2].strip()[:96]}")
