"""Small workload for compute-sanitizer (memcheck / racecheck): every kernel family once on modest volumes."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "connected-components-3d_b200")); sys.path.insert(0, ROOT)
import cc3d_b200
rng = np.random.default_rng(0)
def blobs(shape, nvals, scale):
    coarse = rng.integers(0, nvals, tuple((s + scale - 1) // scale for s in shape))
    for ax in range(len(shape)):
        coarse = np.repeat(coarse, scale, ax)
    return coarse[tuple(slice(0, s) for s in shape)]
vols = [blobs((40, 50, 256), 5, 6).astype(np.uint32), (rng.random((33, 47, 130)) < 0.5).astype(np.uint8),
        blobs((20, 30, 97), 4, 3).astype(np.uint64), blobs((300, 384), 4, 7).astype(np.uint16),
        ((rng.random((21, 30, 96)) < 0.4) * rng.integers(1, 200, (21, 30, 96))).astype(np.uint8)]   # 1-byte binary kernel A, 16-byte rows
for v in vols:
    conns = (6, 18, 26) if v.ndim == 3 else (4, 8)
    for c in conns:
        out, N = cc3d_b200.connected_components(v, connectivity=c, return_N=True)
        out, N = cc3d_b200.connected_components(v, connectivity=c, return_N=True, binary_image=True)
        if c in (4, 8, 6):
            cc3d_b200.connected_components(v, connectivity=c, periodic_boundary=True)
    f = (v.astype(np.float32) * 3 + rng.random(v.shape).astype(np.float32))
    cc3d_b200.connected_components(f, connectivity=conns[-1], delta=1.5)
    lab = cc3d_b200.connected_components(v, connectivity=conns[-1])
    cc3d_b200.statistics(lab)
    cc3d_b200.dust(v, threshold=20, connectivity=conns[-1])
    cc3d_b200.dust(v, threshold=5, connectivity=conns[-1], binary_image=True)     # block-path session: run labels on demand
    cc3d_b200.largest_k(v, 3, connectivity=conns[-1])
    g = cc3d_b200.voxel_connectivity_graph(v, connectivity=conns[-1])
    cc3d_b200.color_connectivity_graph(g, connectivity=conns[-1])
    cc3d_b200.contacts(v, connectivity=conns[-1])
t = torch.from_numpy(vols[0].view(np.int32)).cuda()
cc3d_b200.connected_components(t, return_N=True); cc3d_b200.statistics(cc3d_b200.connected_components(t))
from cc3d_b200 import sharded
sharded.connected_components_slabs([t[:20], t[20:]], connectivity=26, return_N=True)
torch.cuda.synchronize()
print("sanitize workload done")
