import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "connected-components-3d_b200")); sys.path.insert(0, ROOT)
import cc3d_b200
rng = np.random.default_rng(77)
for it in range(47):
    dims = int(rng.integers(2, 4))
    shape = tuple(int(rng.integers(2, 40)) for _ in range(dims))
    dt = [np.float16, np.float32, np.float64][it % 3]
    vals = np.array([0.0, -0.0, 1.0, 2.0, 2.5, np.inf, -np.inf, np.nan, 1e-3, 65000.0], dtype=dt)
    x = vals[rng.integers(0, len(vals), shape)]
    if rng.random() < 0.5:
        x = np.repeat(np.repeat(x, 3, 0), 3, 1)
    x = np.asarray(x, order="F" if rng.random() < 0.5 else "C")
    conns = [4, 8, 6, 18, 26] if x.ndim == 2 else [6, 18, 26]
    c = int(conns[rng.integers(len(conns))])
    if dt != np.float16:
        rng.choice([0.5, 1.0, 1e30])
        if c == 26 and it == 46:
            xin = np.asarray((x != 0).astype(dt), order="F" if x.flags.f_contiguous else "C")
            np.save(os.path.join(ROOT, "gpurun_out", "case46.npy"), xin)
            print(it, x.shape, dt.__name__, "F" if x.flags.f_contiguous else "C", flush=True)
            out, N = cc3d_b200.connected_components(xin, connectivity=c, return_N=True, binary_image=True)
print("done")
