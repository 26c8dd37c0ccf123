"""Host-side slab merge alone (no GPU needed): cc3d_b200_merge_slabs through both Python entry points on a synthetic
8-slab interface graph shaped like the 8 x 512^3 bench step (2 900 labels per slab, 6 700 reported pairs per
interface, a few hundred distinct ones: the face kernel reports a pair once per touching voxel pair)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "connected-components-3d_b200"))
from cc3d_b200 import sharded

rng = np.random.default_rng(0)
world, cap, n_lab, n_pairs = 8, 8192, 2900, 6700
for distinct in (300, 1000, 6700):
    facts = np.zeros((world, 4 + cap), dtype=np.int64)
    facts[:, 0] = n_lab
    for r in range(1, world):
        lo, up = rng.integers(1, n_lab + 1, distinct), rng.integers(1, n_lab + 1, distinct)
        idx = np.clip(np.sort(rng.integers(0, distinct, n_pairs)) + rng.integers(-2, 3, n_pairs), 0, distinct - 1)
        facts[r, 3] = n_pairs
        facts[r, 4:4 + n_pairs] = (lo[idx] << 32) | up[idx]
    lists = [facts[r, 4:4 + int(facts[r, 3])] for r in range(world)]
    Nw, remaps = sharded._global_numbering(facts[:, 0], lists, range(world))
    res = {}
    for name, fn in (("_merge_native (per-slab views)", lambda: sharded._merge_native(facts[:, 0], lists, 3)),
                     ("_merge_gathered (gathered buffer)", lambda: sharded._merge_gathered(facts, 3))):
        ts = []
        for _ in range(300):
            t = time.perf_counter(); out = fn(); ts.append(time.perf_counter() - t)
        assert out[0] == Nw and np.array_equal(out[1], remaps[3])
        res[name] = 1e6 * sorted(ts)[150]
    print(f"{world} slabs x {n_lab} labels, {n_pairs} pairs per interface, {distinct} distinct: "
          + ", ".join(f"{k} {v:.0f} us" for k, v in res.items()) + f" (N = {Nw})", flush=True)
