"""Per-kernel timing of the label pipeline on device-resident workloads (debug helper for gpurun)."""
import os, sys, time, ctypes
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "connected-components-3d_b200")); sys.path.insert(0, ROOT)
import cc3d_b200, benchdata
from cc3d_b200 import _lib

n = int(os.environ.get("N", "512"))
dev = "cuda"
def run(name, x, **kw):
    cc3d_b200.set_timing(False)
    for _ in range(2):
        out, N = cc3d_b200.connected_components(x, return_N=True, **kw)
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); out, N = cc3d_b200.connected_components(x, return_N=True, **kw); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    cc3d_b200.set_timing(True)
    out, N = cc3d_b200.connected_components(x, return_N=True, **kw)
    tm = cc3d_b200.last_timings()
    cc3d_b200.set_timing(False)
    vox = x.numel()
    print(f"== {name}: N={N} out={out.dtype} best {min(ts):.3f} ms  median {sorted(ts)[2]:.3f} ms -> {vox/min(ts)/1e6:.1f} GVx/s", flush=True)
    print("   " + "  ".join(f"{k}={v:.3f}" for k, v in tm), flush=True)
    return out, N

x = benchdata.voronoi_multilabel((n, n, n), cell=40, seed=2, device=dev, dtype=torch.int32)
run("voronoi u32 26", x, connectivity=26)
run("voronoi u32 6", x, connectivity=6)
run("voronoi u32 18", x, connectivity=18)
if os.environ.get("U64"):
    x64 = benchdata.voronoi_multilabel((256, 1024, 1024), cell=80, seed=2, device=dev, dtype=torch.int64, id_bits=62)
    run("voronoi u64 256x1024x1024 26", x64, connectivity=26)
    del x64
xb = benchdata.random_binary((n, n, n), 0.5, 1, dev)
run("random binary u8 26 (multilabel path)", xb, connectivity=26)
run("random binary u8 26 (binary)", xb, connectivity=26, binary_image=True)
run("random binary u8 6 (binary)", xb, connectivity=6, binary_image=True)
xf = benchdata.three_tone_noise((n, n, n), cell=64, seed=3, device=dev)
run("three-tone f32 delta=10 26", xf, connectivity=26, delta=10)
from oracle import decode_connectomics
vol = decode_connectomics.load_fixture()
if vol is not None:
    xc = torch.from_numpy(np.ascontiguousarray(vol.transpose(2, 1, 0)).view(np.int32)).to(dev)
    out, N = run("connectomics u32 26", xc, connectivity=26)
    print("N (expect 3619):", N)
    run("connectomics u32 6", xc, connectivity=6)
