"""statistics: the public call on (a) the connectomics labelling 512^3, (b) Voronoi 512^3 labelling, (c) random-noise labels;
prints time per call for the kernel selected by CC3D_B200_STATS (unset: x-run kernel, v1: vertical-run kernel)."""
import os, sys, ctypes
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "connected-components-3d_b200")); sys.path.insert(0, ROOT)
import cc3d_b200, benchdata
from oracle import decode_connectomics
from cc3d_b200 import _lib
L = _lib.lib()
tag = os.environ.get("CC3D_B200_STATS", "x")
vol = decode_connectomics.load_fixture()
cases = []
if vol is not None:
    x = torch.from_numpy(np.ascontiguousarray(vol.transpose(2, 1, 0)).view(np.int32)).cuda()
    cases.append(("connectomics512", cc3d_b200.connected_components(x, connectivity=26, return_N=True)))
x = benchdata.voronoi_multilabel((512, 512, 512), cell=40, seed=2, device="cuda", dtype=torch.int32)
cases.append(("voronoi512", cc3d_b200.connected_components(x, connectivity=26, return_N=True)))
x = benchdata.random_binary((256, 512, 512), 0.5, 1, "cuda")
cases.append(("binary_noise_6conn_256x512x512", cc3d_b200.connected_components(x, connectivity=6, return_N=True, binary_image=True)))
del x
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
for name, (lab, N) in cases:
    sz, sy, sx = lab.shape
    counts = torch.empty((N + 1,), dtype=torch.int32, device="cuda"); bbox = torch.empty((N + 1, 6), dtype=torch.int32, device="cuda")
    sums = torch.empty((N + 1, 3), dtype=torch.int64, device="cuda")
    kind = {torch.int16: _lib.U16, torch.uint16: _lib.U16, torch.int32: _lib.U32, torch.uint32: _lib.U32}.get(lab.dtype, _lib.U32)
    def kern():
        _lib.check(L.cc3d_b200_statistics(lab.data_ptr(), kind, sx, sy, sz, N, counts.data_ptr(), bbox.data_ptr(), sums.data_ptr(), _lib.DEVICE, st))
    def t(fn):
        fn(); torch.cuda.synchronize(); ts = []
        for _ in range(7):
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
        return min(ts)
    tk = t(kern); tp = t(lambda: cc3d_b200.statistics(lab, no_slice_conversion=True))
    chk = int(counts.cpu().numpy().astype(np.uint32).astype(np.int64).sum()) == lab.numel()
    print(f"[{tag}] {name} N={N} {lab.dtype}: C-ABI kernel call {tk:.3f} ms = {lab.numel()*lab.element_size()/tk/1e6:.0f} GB/s; public call {tp:.3f} ms; counts sum ok: {chk}", flush=True)
