"""configs[2] on ONE GPU: the 2048^3 uint64 Voronoi volume as 8 virtual z-slabs
(cc3d_b200.sharded.connected_components_slabs) - the 1-GPU denominator of the 8-GPU scaling figure."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "connected-components-3d_b200")); sys.path.insert(0, ROOT)
import benchdata
from cc3d_b200 import sharded
n = int(os.environ.get("N", "2048")); ns = int(os.environ.get("SLABS", "8")); cell = int(os.environ.get("CELL", "160"))
szr = n // ns
slabs = [benchdata.voronoi_multilabel((n, n, n), cell=cell, seed=2, device="cuda", dtype=torch.int64, id_bits=62,
                                      z_range=(r * szr, (r + 1) * szr)) for r in range(ns)]
torch.cuda.synchronize()
print("generated", torch.cuda.memory_allocated() / 2**30, "GiB", flush=True)
for conn in (26, 6):
    outs, N = sharded.connected_components_slabs(slabs, connectivity=conn, return_N=True)
    torch.cuda.synchronize()
    del outs
    ts = []
    for _ in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        outs, N = sharded.connected_components_slabs(slabs, connectivity=conn, return_N=True)
        torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
        mx = max(int(o.view(torch.int32).max()) for o in outs)
        del outs
    print(f"voronoi u64 {n}^3 conn={conn} 1 GPU, {ns} virtual slabs: N={N} max_label==N={mx == N} best {min(ts)*1e3:.2f} ms -> {n**3/min(ts)/1e9:.1f} GVx/s "
          f"peak mem {torch.cuda.max_memory_allocated()/2**30:.1f} GiB (torch only)", flush=True)
