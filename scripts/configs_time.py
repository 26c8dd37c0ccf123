"""BASELINE.json configs[3] and configs[4] at full size on one B200 (device-resident inputs, CUDA-event times):
continuous 1024^3 f32 delta=10; 6-connected periodic 1024^3; 16384^2 2D 8-connected (binary and multilabel call);
statistics + dust(threshold=100) on a 2048x2048xZ uint64 Voronoi labelling (Z = 512: one call below 2^32 voxels)."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "connected-components-3d_b200")); sys.path.insert(0, ROOT)
import cc3d_b200, benchdata
dev = "cuda"

def marks(fn):
    cc3d_b200.set_timing(True); fn(); tm = cc3d_b200.last_timings(); cc3d_b200.set_timing(False)
    print("   kernels: " + " ".join(f"{k}={v:.3f}" for k, v in tm), flush=True)

def timed(name, fn, vox, reps=4, bytes_per_vox=None):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); r = fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    t = min(ts)
    extra = f"  {bytes_per_vox * vox / t / 1e6:.0f} GB/s of compulsory bytes" if bytes_per_vox else ""
    print(f"{name}: best {t:.3f} ms -> {vox / t / 1e6:.1f} GVx/s{extra}", flush=True)
    return r

n = int(os.environ.get("N", "1024"))
x = benchdata.three_tone_noise((n, n, n), cell=64, seed=3, device=dev)
r = timed(f"configs[3] continuous {n}^3 f32 delta=10 conn26", lambda: cc3d_b200.connected_components(x, connectivity=26, delta=10, return_N=True), x.numel(), bytes_per_vox=8)
print("   N =", r[1], r[0].dtype); marks(lambda: cc3d_b200.connected_components(x, connectivity=26, delta=10, return_N=True)); del x, r
g = torch.Generator(device=dev); g.manual_seed(4)
x = torch.randint(0, 4, (n, n, n), generator=g, device=dev, dtype=torch.int32)
r = timed(f"configs[4] periodic 6-conn {n}^3 u32 random labels 0..3", lambda: cc3d_b200.connected_components(x, connectivity=6, periodic_boundary=True, return_N=True), x.numel(), bytes_per_vox=8)
print("   N =", r[1], r[0].dtype); marks(lambda: cc3d_b200.connected_components(x, connectivity=6, periodic_boundary=True, return_N=True)); del x, r
m = int(os.environ.get("M", "16384"))
x = benchdata.random_binary((m, m), 0.5, 5, dev)
r = timed(f"configs[4] 2D {m}^2 u8 8-conn binary_image=True", lambda: cc3d_b200.connected_components(x, connectivity=8, binary_image=True, return_N=True), x.numel(), bytes_per_vox=5)
print("   N =", r[1], r[0].dtype); marks(lambda: cc3d_b200.connected_components(x, connectivity=8, binary_image=True, return_N=True))
r = timed(f"configs[4] 2D {m}^2 u8 8-conn multilabel call", lambda: cc3d_b200.connected_components(x, connectivity=8, return_N=True), x.numel(), bytes_per_vox=5)
print("   N =", r[1], r[0].dtype); marks(lambda: cc3d_b200.connected_components(x, connectivity=8, return_N=True)); del x, r
zs = int(os.environ.get("Z", "512"))
x = benchdata.voronoi_multilabel((2048, 2048, 2048), cell=160, seed=2, device=dev, dtype=torch.int64, id_bits=62, z_range=(0, zs))
lab, N = cc3d_b200.connected_components(x, connectivity=26, return_N=True)
print(f"labelling 2048x2048x{zs} u64: N = {N}", flush=True)
st = timed(f"configs[4] statistics on 2048x2048x{zs} u32 labels", lambda: cc3d_b200.statistics(lab, no_slice_conversion=True), lab.numel(), bytes_per_vox=4)
print("   counts[:4] =", st["voxel_counts"][:4])
timed(f"configs[4] dust(threshold=100) on 2048x2048x{zs} u64 (CCL + statistics + mask)", lambda: cc3d_b200.dust(x, threshold=100, connectivity=26), x.numel(), reps=2, bytes_per_vox=28)
