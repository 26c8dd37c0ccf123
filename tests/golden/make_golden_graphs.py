"""Generates tests/golden/graphs_*.npz from the UNMODIFIED reference for the SURVEY 8(f) rows: largest_k (the
reference's Python layer, executed in place from /root/reference through oracle.reference_package()),
voxel_connectivity_graph, color_connectivity_graph and contacts (the built extension, oracle/_ref).

    bash oracle/build_ref.sh && python tests/golden/make_golden_graphs.py

The GPU box has no /root/reference: tests/test_parity_gpu.py::test_graph_goldens checks cc3d_b200 against these."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import oracle  # noqa: E402

ext, pkg = oracle.reference_module(), oracle.reference_package()
assert ext is not None and pkg is not None, "needs /root/reference and oracle/_ref (bash oracle/build_ref.sh)"


def blobs(rng, shape, nvals, scale):
  coarse = rng.integers(0, nvals, tuple((s + scale - 1) // scale for s in shape))
  for ax in range(len(shape)):
    coarse = np.repeat(coarse, scale, axis=ax)
  return coarse[tuple(slice(0, s) for s in shape)]


rng = np.random.default_rng(20241017)
n = 0
for i, (shape, conn) in enumerate([((40, 33, 21), 26), ((64, 20, 9), 6), ((37, 29, 18), 18), ((90, 70), 8), ((77, 50), 4)]):
  for order in "CF":
    x = np.asarray(blobs(rng, shape, 6, 4).astype([np.uint32, np.uint16, np.uint64, np.uint8, np.int32][i]), order=order)
    out = dict(x=x, connectivity=conn)
    out["vcg"] = ext.voxel_connectivity_graph(x, connectivity=conn)
    if conn in (4, 8, 6, 26):
      g = out["vcg"].copy()
      g[rng.random(g.shape) < 0.05] &= g.dtype.type(0x2AAAAAA & (0xFF if g.dtype == np.uint8 else 0x3FFFFFF))
      col, N = ext.color_connectivity_graph(g, connectivity=conn, return_N=True)
      out.update(vcg_cut=g, colors=col, colors_N=N)
    ct = ext.contacts(x, connectivity=conn, surface_area=True, anisotropy=(4, 4, 40))
    keys = np.array(sorted(ct.keys()), dtype=np.uint64).reshape(-1, 2)
    out.update(contact_pairs=keys, contact_areas=np.array([ct[tuple(int(v) for v in k)] for k in keys], dtype=np.float32))
    if len(shape) == 3:
      for k in (1, 3):
        lk, lkN = pkg.largest_k(x, k, connectivity=conn, return_N=True)
        out[f"largest_{k}"] = lk
        out[f"largest_{k}_N"] = lkN
    np.savez_compressed(os.path.join(HERE, f"graphs_{i}_{order}.npz"), **out)
    n += 1
print("wrote", n, "graph fixtures")
