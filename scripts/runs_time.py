"""runs / draw on a device-resident 512^3 uint32 labelling: the C-ABI calls alone (count-only call =
R1 + scan; full call = R1 + scan + R2) and a whole-volume draw of one label's runs, with HBM fractions, next to the
reference's extract_runs on one host core (oracle/_ref, bounded sample)."""
import os, sys, ctypes, time, json
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "connected-components-3d_b200")); sys.path.insert(0, ROOT)
import cc3d_b200, benchdata
from cc3d_b200 import _lib
L = _lib.lib()
PEAK = 6456.2
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
x = benchdata.voronoi_multilabel((512, 512, 512), cell=40, seed=1, device="cuda", dtype=torch.int32)
lab, N = cc3d_b200.connected_components(x, connectivity=26, return_N=True)
flat = lab.reshape(-1)
n = flat.numel()
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
count = ctypes.c_uint64(0)
_lib.check(L.cc3d_b200_runs(flat.data_ptr(), _lib.U32, n, None, None, None, 0, ctypes.byref(count), _lib.DEVICE, st))
k = count.value
tab = torch.empty((3, k), dtype=torch.int64, device="cuda")
def count_only():
    _lib.check(L.cc3d_b200_runs(flat.data_ptr(), _lib.U32, n, None, None, None, 0, ctypes.byref(count), _lib.DEVICE, st))
def full():
    _lib.check(L.cc3d_b200_runs(flat.data_ptr(), _lib.U32, n, tab[0].data_ptr(), tab[1].data_ptr(), tab[2].data_ptr(), k, ctypes.byref(count), _lib.DEVICE, st))
def t(fn, name, bytes_):
    fn(); torch.cuda.synchronize(); ts = []
    for _ in range(7):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ms = min(ts)
    print(f"{name}: best {ms:.3f} ms = {bytes_ / ms / 1e6:.0f} GB/s = {bytes_ / ms / 1e6 / PEAK:.2f} of HBM peak ({PEAK:.0f} GB/s)", flush=True)
    return ms
print(f"512^3 uint32 labels, N = {N}, runs = {k} (mean length {(flat != 0).sum().item() / max(k, 1):.1f})")
a = t(count_only, "cc3d_b200_runs count only (R1 + scan)", n * 4)
b = t(full, "cc3d_b200_runs full (R1 + scan + R2)", 2 * n * 4 + k * 24)
print(f"R2 alone ~ {b - a:.3f} ms = {(n * 4 + k * 24) / (b - a) / 1e6:.0f} GB/s")
# draw: all runs of the volume drawn into an image (every run written once)
img = torch.zeros(n, dtype=torch.int32, device="cuda")
nz = int((flat != 0).sum().item())
def draw_all():
    _lib.check(L.cc3d_b200_draw(img.data_ptr(), _lib.U32, n, 1, tab[1].data_ptr(), tab[2].data_ptr(), k, _lib.DEVICE, st))
t(draw_all, "cc3d_b200_draw of every run (check + draw kernels)", nz * 4 + k * 16 * 2)
assert int((img != 0).sum().item()) == nz
t0 = time.perf_counter(); r = cc3d_b200.runs(lab); t1 = time.perf_counter()
print(f"cc3d_b200.runs public call (GPU table + D2H + host grouping into a dict of {len(r)} labels): {(t1 - t0) * 1e3:.0f} ms")
t0 = time.perf_counter(); m = sum(1 for _ in cc3d_b200.each(lab[:, :, :64].contiguous(), in_place=True)); t1 = time.perf_counter()
print(f"cc3d_b200.each over 512x512x64 CUDA labels, in_place: {m} images in {(t1 - t0) * 1e3:.0f} ms")
try:
    from oracle import oracle
    ref = oracle.reference_module()
    sample = lab[:64].cpu().numpy()
    t0 = time.perf_counter(); rr = ref.runs(sample); t1 = time.perf_counter()
    print(f"reference runs on 64x512x512 (1 host core): {(t1 - t0) * 1e3:.0f} ms = {sample.size / (t1 - t0) / 1e9:.3f} GVx/s; "
          f"same table: {rr == cc3d_b200.runs(sample)}")
except Exception as e:
    print("reference timing skipped:", e)
