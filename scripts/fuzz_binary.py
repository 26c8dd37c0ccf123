"""Time-boxed differential fuzz of the round-2d binary paths against the reference build (oracle/_ref): the 1-byte kernel A
(k_fg_bitmap_u8 / k_faces_from_fg), the block-label path of 26-connected volumes (k_block_minrun, k_block_labels,
k_expand_blocks, k_block_fill_L through fused dust), dense edge queues (B2 dedupe) and the four-runs-per-lane C stage.
FUZZ_SECONDS=60 SEED=1 python scripts/fuzz_binary.py"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "connected-components-3d_b200")); sys.path.insert(0, ROOT)
import cc3d_b200
from oracle import oracle

truth = oracle.reference_module() or oracle
pkg = oracle.reference_package() or oracle
rng = np.random.default_rng(int(os.environ.get("SEED", "1")))
budget = float(os.environ.get("FUZZ_SECONDS", "60"))
t0 = time.time(); n = 0; bad = 0
while time.time() - t0 < budget:
    dims = 2 if rng.random() < 0.25 else 3
    sx = int(rng.choice([16, 32, 64, 96, 128, 144, 160, 256, 512, 33, 100, 130, int(rng.integers(1, 300))]))
    rest = (int(rng.integers(1, 120)),) if dims == 2 else (int(rng.integers(1, 40)), int(rng.integers(1, 24)))
    shape = (sx,) + rest
    dt = [np.uint8, bool, np.int8, np.uint16, np.uint32][int(rng.integers(0, 5))]
    p = float(rng.choice([0.02, 0.1, 0.25, 0.4, 0.5, 0.6, 0.85, 1.0]))
    if rng.random() < 0.3:      # blobs with holes instead of noise
        coarse = rng.random(tuple((s + 4) // 5 for s in shape)) < p
        m = coarse
        for ax in range(dims):
            m = np.repeat(m, 5, axis=ax)
        m = m[tuple(slice(0, s) for s in shape)] & (rng.random(shape) < 0.97)
    else:
        m = rng.random(shape) < p
    x = np.asfortranarray(m) if dt == bool else np.asfortranarray((m * rng.integers(1, 100, shape)).astype(dt))
    if rng.random() < 0.5:
        x = np.ascontiguousarray(x.T)      # C order: the fast axis is the last one
    fast = sx
    conns = [4, 8] if dims == 2 else [6, 18, 26]
    c = int(conns[int(rng.integers(0, len(conns)))])
    ref_fast = x.shape[0] if x.flags.f_contiguous else x.shape[-1]      # (1, n) / (n, 1) arrays are both C and F contiguous: the reference reads them as F
    if c == 8 and (fast % 2 == 1 or ref_fast % 2 == 1):
        continue      # reference defect D1 (stale labels when ITS fast axis is odd)
    if c == 4 and dt != bool:
        x = ((x != 0) * 3).astype(x.dtype)      # reference defect: first row of the binary 2D-4 kernel compares values
    try:
        a, Na = truth.connected_components(x, connectivity=c, return_N=True, binary_image=True)
    except RuntimeError:
        continue      # reference defect D3
    b, Nb = cc3d_b200.connected_components(x, connectivity=c, return_N=True, binary_image=True)
    ok = Na == Nb and a.dtype == b.dtype and np.array_equal(a, b)
    if ok and c == 26 and dt in (np.uint8, np.uint16, np.uint32) and rng.random() < 0.4:
        thr = int(rng.integers(2, 30))
        da, dNa = pkg.dust(x, thr, connectivity=26, binary_image=True, return_N=True)
        db, dNb = cc3d_b200.dust(x, thr, connectivity=26, binary_image=True, return_N=True)
        ok = dNa == dNb and np.array_equal(da, db)
    n += 1
    if not ok:
        bad += 1
        print("MISMATCH", shape, np.dtype(dt), "C" if x.flags.c_contiguous else "F", c, p, Na, Nb, flush=True)
print(f"fuzz_binary: {n} cases, {bad} mismatches, {time.time() - t0:.1f} s", flush=True)
sys.exit(1 if bad else 0)
