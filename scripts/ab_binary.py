"""Random binary 512^3 (BASELINE configs[1]) step times + per-kernel marks (A/B helper)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "connected-components-3d_b200")); sys.path.insert(0, ROOT)
import cc3d_b200, benchdata
x = benchdata.random_binary((512, 512, 512), 0.5, 1, "cuda")
for conn, kw in ((26, dict(binary_image=True)), (26, {}), (18, dict(binary_image=True)), (6, dict(binary_image=True))):
    for _ in range(3): out, N = cc3d_b200.connected_components(x, connectivity=conn, return_N=True, **kw)
    torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); out, N = cc3d_b200.connected_components(x, connectivity=conn, return_N=True, **kw); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    cc3d_b200.set_timing(True); cc3d_b200.connected_components(x, connectivity=conn, return_N=True, **kw); tm = cc3d_b200.last_timings(); cc3d_b200.set_timing(False)
    print(f"binary512 conn={conn} {kw}: N={N} median {ts[5]:.4f} ms | " + " ".join(f"{k}={v:.3f}" for k, v in tm), flush=True)
