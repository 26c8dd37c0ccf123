"""statistics of the connectomics labelling (target of ncu runs under gpurun)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "connected-components-3d_b200")); sys.path.insert(0, ROOT)
import cc3d_b200, benchdata
from oracle import decode_connectomics
vol = decode_connectomics.load_fixture()
if vol is not None and os.environ.get("WL", "connectomics") == "connectomics":
    x = torch.from_numpy(np.ascontiguousarray(vol.transpose(2, 1, 0)).view(np.int32)).cuda()
else:
    x = benchdata.voronoi_multilabel((512, 512, 512), cell=40, seed=2, device="cuda", dtype=torch.int32)
lab, N = cc3d_b200.connected_components(x, connectivity=26, return_N=True)
torch.cuda.synchronize()
import time
for _ in range(3):
    t0 = time.perf_counter(); st = cc3d_b200.statistics(lab, no_slice_conversion=True); torch.cuda.synchronize(); print("statistics", (time.perf_counter() - t0) * 1e3, "ms")
print("N", N, int(st["voxel_counts"].astype(np.int64).sum()))
