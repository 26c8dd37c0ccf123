"""Generates tests/golden/runs_*.npz from the UNMODIFIED reference build (oracle/_ref) for SURVEY 8(f)4:
fastcc3d.runs (run table per label), fastcc3d.draw on a non-empty canvas and the images cc3d.each yields.

    bash oracle/build_ref.sh && python tests/golden/make_golden_runs.py

The GPU box has no /root/reference: tests/test_runs_gpu.py checks cc3d_b200 against these."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import oracle  # noqa: E402

ext = oracle.reference_module()
assert ext is not None, "needs oracle/_ref (bash oracle/build_ref.sh)"


def blobs(rng, shape, nvals, scale):
  coarse = rng.integers(0, nvals, tuple((s + scale - 1) // scale for s in shape))
  for ax in range(len(shape)):
    coarse = np.repeat(coarse, scale, axis=ax)
  return coarse[tuple(slice(0, s) for s in shape)]


rng = np.random.default_rng(20241018)
cases = [((40, 33, 21), np.uint32, 6, 4), ((64, 20, 9), np.uint16, 4, 1), ((130, 70), np.uint8, 3, 5), ((5000,), np.uint64, 5, 7),
         ((33, 17, 5), np.bool_, 2, 2), ((70, 66, 3), np.uint64, 9, 8)]
n = 0
for i, (shape, dt, nvals, scale) in enumerate(cases):
  for order in "CF":
    x = blobs(rng, shape, nvals, scale)
    if dt == np.uint64:
      x = x.astype(np.uint64) * np.uint64(0x1F23456789ABCDEF)   # values above 2^32 and above 2^63
    x = np.asarray(x.astype(dt), order=order)
    r = ext.runs(x)
    keys = np.array(list(r.keys()), dtype=np.uint64)
    offsets = np.cumsum([0] + [len(v) for v in r.values()]).astype(np.int64)
    table = np.array([p for v in r.values() for p in v], dtype=np.uint64).reshape(-1, 2)
    out = dict(x=x, f_order=(order == "F"), keys=keys, offsets=offsets, table=table)
    # draw: the runs of the largest key drawn over a canvas of sevens, in the image dtype
    if len(keys):
      canvas = np.full(x.shape, 7 if dt != np.bool_ else 0, dtype=dt, order=order)
      k = int(keys[-1])
      val = k if dt == np.bool_ or k <= np.iinfo(dt).max else 1
      ext.draw(val, r[k], canvas)
      out.update(draw_key=np.uint64(k), draw_value=np.uint64(val), drawn=canvas)
    # each: label, checksum of the yielded image (position-weighted) for both modes
    for binary in (False, True):
      sums = []
      for label, img in ext.each(x, binary=binary, in_place=False):
        f = img.reshape(-1, order=order)
        nz = np.flatnonzero(f)
        sums.append((label, nz.size, int(nz.sum() % (1 << 61)), int(f[nz[0]]) if nz.size else 0))
      out[f"each_{int(binary)}"] = np.array(sums, dtype=np.uint64).reshape(-1, 4)
    np.savez_compressed(os.path.join(HERE, f"runs_{i}_{order}.npz"), **out)
    n += 1
print("wrote", n, "run fixtures")
