"""Statistics kernel alone (C-ABI call on device buffers) vs the public call, 2048x2048xZ u32 labels."""
import os, sys, ctypes
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "connected-components-3d_b200")); sys.path.insert(0, ROOT)
import cc3d_b200, benchdata
from cc3d_b200 import _lib
L = _lib.lib()
zs = int(os.environ.get("Z", "512"))
x = benchdata.voronoi_multilabel((2048, 2048, 2048), cell=160, seed=2, device="cuda", dtype=torch.int64, id_bits=62, z_range=(0, zs))
lab, N = cc3d_b200.connected_components(x, connectivity=26, return_N=True)
del x
counts = torch.empty((N + 1,), dtype=torch.int32, device="cuda"); bbox = torch.empty((N + 1, 6), dtype=torch.int32, device="cuda")
sums = torch.empty((N + 1, 3), dtype=torch.int64, device="cuda")
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
def kern():
    _lib.check(L.cc3d_b200_statistics(lab.data_ptr(), _lib.U32, 2048, 2048, zs, N, counts.data_ptr(), bbox.data_ptr(), sums.data_ptr(), _lib.DEVICE, st))
def t(fn, name, bytes_):
    fn(); torch.cuda.synchronize(); ts = []
    for _ in range(5):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    print(f"{name}: best {min(ts):.3f} ms = {bytes_ / min(ts) / 1e6:.0f} GB/s", flush=True)
t(kern, "cc3d_b200_statistics (C-ABI, device buffers, N known)", lab.numel() * 4)
t(lambda: cc3d_b200.statistics(lab, no_slice_conversion=True), "cc3d_b200.statistics (public call: max + kernel + D2H + finalise)", lab.numel() * 4)
t(lambda: torch.aminmax(lab.view(torch.int32)), "torch.aminmax", lab.numel() * 4)
ref = cc3d_b200.statistics(lab, no_slice_conversion=True)
print("sum counts == voxels:", int(ref["voxel_counts"].astype(np.int64).sum()) == lab.numel())
