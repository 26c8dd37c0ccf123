"""Kernel A timing on the connectomics volume for whatever CC3D_B200_* knobs are set (A/B helper)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "connected-components-3d_b200")); sys.path.insert(0, ROOT)
import cc3d_b200
from oracle import decode_connectomics
vol = decode_connectomics.load_fixture()
x = torch.from_numpy(np.ascontiguousarray(vol.transpose(2, 1, 0)).view(np.int32)).cuda()
for _ in range(3): cc3d_b200.connected_components(x, connectivity=26)
cc3d_b200.set_timing(True)
acc = {}
for _ in range(20):
    cc3d_b200.connected_components(x, connectivity=26)
    for k, v in cc3d_b200.last_timings(): acc.setdefault(k, []).append(v)
print(os.environ.get("CC3D_B200_ZCHUNK", "-"), " ".join(f"{k}={sorted(v)[len(v)//2]:.4f}" for k, v in acc.items()), flush=True)
