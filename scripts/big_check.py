"""Full-size sanity + timing of the BASELINE configs that fit one GPU (size-independent properties:
idempotence of the labelling, N consistency between connectivities, statistics sums)."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "connected-components-3d_b200")); sys.path.insert(0, ROOT)
import cc3d_b200, benchdata

def timed(x, **kw):
    out, N = cc3d_b200.connected_components(x, return_N=True, **kw)
    torch.cuda.synchronize()
    ts = []
    for _ in range(3):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); out, N = cc3d_b200.connected_components(x, return_N=True, **kw); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return out, N, min(ts)

def check_idempotent(out, N, **kw):
    # labelling the label image again must reproduce it exactly (same partition, same numbering)
    o = out.view(torch.int32) if out.dtype == torch.uint32 else out
    out2, N2 = cc3d_b200.connected_components(o, return_N=True, **kw)
    same = bool(torch.equal(out2.view(o.dtype) if out2.element_size() == o.element_size() else out2.to(o.dtype), o))
    return N2 == N and same

n = int(os.environ.get("N", "1024"))
x = benchdata.voronoi_multilabel((n, n, n), cell=64, seed=5, device="cuda", dtype=torch.int32)
out, N, t = timed(x, connectivity=26)
print(f"voronoi {n}^3 u32 26: N={N} {t:.3f} ms {x.numel()/t/1e6:.1f} GVx/s idempotent={check_idempotent(out, N, connectivity=26)}", flush=True)
out6, N6, t6 = timed(x, connectivity=6)
print(f"voronoi {n}^3 u32 6: N={N6} (>= N26: {N6 >= N}) {t6:.3f} ms {x.numel()/t6/1e6:.1f} GVx/s", flush=True)
del out, out6
xf = benchdata.three_tone_noise((n, n, n), cell=64, seed=3, device="cuda")
outf, Nf, tf = timed(xf, connectivity=26, delta=10)
print(f"three-tone {n}^3 f32 delta=10 26: N={Nf} {tf:.3f} ms {xf.numel()/tf/1e6:.1f} GVx/s", flush=True)
del xf, outf
g = torch.Generator(device="cuda"); g.manual_seed(4)
xp = torch.randint(0, 4, (n, n, n), generator=g, device="cuda", dtype=torch.int32)
outp, Np, tp = timed(xp, connectivity=6, periodic_boundary=True)
outn, Nn, tn = timed(xp, connectivity=6)
print(f"random 0..3 {n}^3 u32 6 periodic: N={Np} (<= non-periodic {Nn}: {Np <= Nn}) {tp:.3f} ms {xp.numel()/tp/1e6:.1f} GVx/s", flush=True)
del xp, outp, outn
m = 16384
g.manual_seed(5)
x2 = (torch.rand((m, m), generator=g, device="cuda") < 0.5).to(torch.uint8)
o2, N2, t2 = timed(x2, connectivity=8)
o2b, N2b, t2b = timed(x2, connectivity=8, binary_image=True)
print(f"random binary {m}^2 u8 8-conn: multilabel N={N2} {t2:.3f} ms {x2.numel()/t2/1e6:.1f} GVx/s; binary N={N2b} {t2b:.3f} ms; same N: {N2 == N2b}", flush=True)
# statistics on a 512^3 labelling: counts sum to the volume
x5 = benchdata.voronoi_multilabel((512, 512, 512), cell=40, seed=2, device="cuda", dtype=torch.int32)
o5, N5 = cc3d_b200.connected_components(x5, return_N=True)
st = cc3d_b200.statistics(o5.cpu().numpy(), no_slice_conversion=True)
print("statistics: sum(counts) == voxels:", int(st["voxel_counts"].astype(np.int64).sum()) == x5.numel(), "labels:", len(st["voxel_counts"]) - 1 == N5)
