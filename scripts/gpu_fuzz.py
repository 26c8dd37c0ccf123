"""Differential fuzz of the CUDA path against the reference build (oracle/_ref) or the C oracle.
Debug helper for gpurun; the real parity tests live in tests/."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "connected-components-3d_b200"))
sys.path.insert(0, ROOT)
import cc3d_b200
from oracle import oracle

ref = oracle.reference_module()
print("reference module:", ref is not None, flush=True)
truth = ref if ref is not None else oracle
rng = np.random.default_rng(int(os.environ.get("SEED", "1")))
NCASE = int(os.environ.get("NCASE", "600"))
MAXDIM = int(os.environ.get("MAXDIM", "80"))
dtypes = [np.uint8, np.uint16, np.uint32, np.uint64, np.int8, np.int32, np.int64, np.float32, np.float64, bool]
bad = 0; n = 0; t0 = time.time()
for it in range(NCASE):
    dims = int(rng.integers(1, 4))
    big = rng.random() < 0.3
    hi = MAXDIM if big else 20
    shape = tuple(int(rng.integers(1, hi)) for _ in range(dims))
    if dims == 2 and rng.random() < 0.3:
        shape = (int(rng.integers(1, 300)), int(rng.integers(1, 300)))
    dt = dtypes[rng.integers(len(dtypes))]
    order = 'F' if rng.random() < 0.5 else 'C'
    nvals = int(rng.integers(2, 6))
    if dt == bool:
        x = rng.random(shape) < rng.random()
    elif rng.random() < 0.3:
        # blobby data: coarse random upsampled
        coarse = rng.integers(0, nvals, tuple((s + 3) // 4 for s in shape))
        x = coarse
        for ax in range(dims):
            x = np.repeat(x, 4, axis=ax)
        x = x[tuple(slice(0, s) for s in shape)].astype(dt)
    else:
        x = rng.integers(0, nvals, shape).astype(dt)
    x = np.asarray(x, order=order)
    conns = [4, 8, 6, 18, 26] if dims == 2 else [6, 18, 26]
    c = conns[rng.integers(len(conns))]
    mode = int(rng.integers(0, 4))
    kw = {}
    fast = x.shape[0] if order == 'F' else x.shape[-1]
    if mode == 1:
        kw['binary_image'] = True
        if dt != bool:
            x = np.asarray((x != 0).astype(dt), order=order)
    elif mode == 2 and dt != bool:
        if np.issubdtype(dt, np.floating):
            x = np.asarray(((x * 3 + rng.random(shape) * 2.5) * (x != 0)).astype(dt), order=order)
            kw['delta'] = float(rng.random() * 3)
        else:
            x = np.asarray((x * 3 + rng.integers(0, 3, shape) * (x != 0)).astype(dt), order=order)
            kw['delta'] = int(rng.integers(1, 4))
    elif mode == 3 and c in (4, 8, 6):
        kw['periodic_boundary'] = True
    is_bin = kw.get('binary_image', False) or dt == bool
    if is_bin and c == 8 and fast % 2 == 1:
        continue  # reference defect D1 (stale labels for odd sx)
    try:
        a, Na = truth.connected_components(x, connectivity=c, return_N=True, **kw)
    except Exception as e:
        continue
    try:
        b, Nb = cc3d_b200.connected_components(x, connectivity=c, return_N=True, **kw)
    except Exception as e:
        bad += 1
        print("EXC", repr(e), shape, dt.__name__, order, c, kw, flush=True)
        continue
    n += 1
    if Na != Nb or a.dtype != b.dtype or a.shape != b.shape or not np.array_equal(a, b):
        bad += 1
        if bad < 25:
            diff = int(np.count_nonzero(np.asarray(a) != np.asarray(b))) if a.shape == b.shape else -1
            print("MISMATCH", shape, dt.__name__, order, c, kw, "N", Na, Nb, a.dtype, b.dtype, "ndiff", diff, flush=True)
            if x.size <= 64:
                print(x); print(a); print(b)
print("cases", n, "bad", bad, "%.1fs" % (time.time() - t0), flush=True)
sys.exit(1 if bad else 0)
