"""One labelling call on a device-resident workload (target of ncu runs under gpurun)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "connected-components-3d_b200")); sys.path.insert(0, ROOT)
import cc3d_b200, benchdata
n = int(os.environ.get("N", "512"))
wl = os.environ.get("WL", "voronoi")
conn = int(os.environ.get("CONN", "26"))
reps = int(os.environ.get("REPS", "2"))
if wl == "voronoi":
    x = benchdata.voronoi_multilabel((n, n, n), cell=40, seed=2, device="cuda", dtype=torch.int32)
elif wl == "binary":
    x = benchdata.random_binary((n, n, n), 0.5, 1, "cuda")
elif wl == "tone":
    x = benchdata.three_tone_noise((n, n, n), cell=64, seed=3, device="cuda")
elif wl == "connectomics":
    from oracle import decode_connectomics
    vol = decode_connectomics.load_fixture()
    x = torch.from_numpy(np.ascontiguousarray(vol.transpose(2, 1, 0)).view(np.int32)).cuda()
kw = {}
if os.environ.get("DELTA"): kw["delta"] = float(os.environ["DELTA"])
if os.environ.get("BINARY"): kw["binary_image"] = True
torch.cuda.synchronize()
for _ in range(reps):
    out, N = cc3d_b200.connected_components(x, connectivity=conn, return_N=True, **kw)
torch.cuda.synchronize()
print("N", N)
