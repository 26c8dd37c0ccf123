"""Localise the segfault seen in the compiled-binding test (estimate_provisional_labels)."""
import os, sys, faulthandler
faulthandler.enable()
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "connected-components-3d_b200")); sys.path.insert(0, ROOT)
import cc3d_b200
from oracle import oracle
ref = oracle.reference_module()
fc = cc3d_b200.fastcc3d
rng = np.random.default_rng(41)
for it in range(30):
  dims = int(rng.integers(1, 4))
  shape = tuple(int(rng.integers(1, 60)) for _ in range(dims))
  dt = [np.uint8, np.uint16, np.uint32, np.uint64, np.int8, np.int32, np.int64, np.float32, np.float64, bool][it % 10]
  x = (rng.random(shape) < 0.5) if dt == bool else rng.integers(0, 4, shape).astype(dt)
  x = np.asarray(x, order="F" if it % 2 else "C")
  print(it, shape, np.dtype(dt), flush=True)
  print("  ctypes", cc3d_b200.estimate_provisional_labels(x), flush=True)
  print("  ref   ", ref.estimate_provisional_labels(x), flush=True)
  print("  cython", fc.estimate_provisional_labels(x), flush=True)
print("epl debug done")
