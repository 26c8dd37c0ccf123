"""Step times + per-kernel marks for the non-headline predicates (A/B helper): binary noise 6/18/26, continuous."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "connected-components-3d_b200")); sys.path.insert(0, ROOT)
import cc3d_b200, benchdata
xb = benchdata.random_binary((512, 512, 512), 0.5, 1, "cuda")
xt = benchdata.three_tone_noise((512, 512, 512), cell=64, seed=3, device="cuda")
cases = [("binary512", xb, 26, dict(binary_image=True)), ("binary512", xb, 18, dict(binary_image=True)), ("binary512", xb, 6, dict(binary_image=True)),
         ("binary512_multilabel_call", xb, 26, {}), ("tone512_f32", xt, 26, dict(delta=10)), ("tone512_f32", xt, 6, dict(delta=10))]
for name, x, conn, kw in cases:
    for _ in range(3): out, N = cc3d_b200.connected_components(x, connectivity=conn, return_N=True, **kw)
    torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); out, N = cc3d_b200.connected_components(x, connectivity=conn, return_N=True, **kw); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    cc3d_b200.set_timing(True); cc3d_b200.connected_components(x, connectivity=conn, return_N=True, **kw); tm = cc3d_b200.last_timings(); cc3d_b200.set_timing(False)
    print(f"{name} conn={conn} {kw}: N={N} median {ts[5]:.4f} ms | " + " ".join(f"{k}={v:.3f}" for k, v in tm), flush=True)
