"""Compare the labelling of wide volumes between the library's default B1 and CC3D_B200_B1=phased (two processes)."""
import os, sys, subprocess, hashlib
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "connected-components-3d_b200")); sys.path.insert(0, ROOT)
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import cc3d_b200, benchdata
    for shape, cell, dt in [((32, 2048, 2048), 160, torch.int64), ((8, 64, 2048), 40, torch.int32), ((8, 64, 1536), 24, torch.int32),
                            ((4, 40, 1024), 16, torch.int32), ((16, 16, 4096), 30, torch.int32), ((1, 300, 2048), 20, torch.int32)]:
        x = benchdata.voronoi_multilabel(shape, cell=cell, seed=2, device="cuda", dtype=dt, id_bits=62 if dt == torch.int64 else 30)
        for conn in ((26, 18, 6) if shape[0] > 1 else (8, 4)):
            out, N = cc3d_b200.connected_components(x, connectivity=conn, return_N=True)
            print(shape, conn, N, hashlib.sha256(out.cpu().numpy().tobytes()).hexdigest()[:16], flush=True)
else:
    a = subprocess.run([sys.executable, __file__, "child"], capture_output=True, text=True, env=dict(os.environ)).stdout
    b = subprocess.run([sys.executable, __file__, "child"], capture_output=True, text=True, env=dict(os.environ, CC3D_B200_B1="phased")).stdout
    for la, lb in zip(a.splitlines(), b.splitlines()):
        print("OK  " if la == lb else "DIFF", la, "|", lb)
