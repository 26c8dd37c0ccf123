import json
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_manifest():
  with open(os.path.join(GOLDEN_DIR, "manifest.json")) as f:
    return json.load(f)


def load_golden(name):
  z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
  x = z["x"]
  labels = z["labels"]
  if bool(z["f_order"]):
    x = np.asfortranarray(x)
    labels = np.asfortranarray(labels)
  return x, labels, int(z["N"]), z


def call_kwargs(kw):
  kw = dict(kw)
  if "out_dtype" in kw:
    kw["out_dtype"] = np.dtype(kw["out_dtype"])
  return kw


def blobs(rng, shape, nvals, scale):
  coarse = rng.integers(0, nvals, tuple((s + scale - 1) // scale for s in shape))
  x = coarse
  for ax in range(len(shape)):
    x = np.repeat(x, scale, axis=ax)
  return x[tuple(slice(0, s) for s in shape)]


def assert_same_labels(a, Na, b, Nb, ctx=""):
  assert Na == Nb, f"N differs: {Na} vs {Nb} {ctx}"
  assert a.dtype == b.dtype, f"dtype differs: {a.dtype} vs {b.dtype} {ctx}"
  assert a.shape == b.shape, f"shape differs: {a.shape} vs {b.shape} {ctx}"
  assert np.array_equal(a, b), f"labels differ in {np.count_nonzero(a != b)} voxels {ctx}"
  assert a.flags.f_contiguous == b.flags.f_contiguous and a.flags.c_contiguous == b.flags.c_contiguous, f"layout differs {ctx}"


def spec_oracle(x, connectivity, mode="eq", delta=0, periodic=False):
  """Independent specification: connected components of the voxel adjacency graph (scipy csgraph),
  numbered by first appearance in memory order. x must be F-ordered 1-3D."""
  import scipy.sparse as sp
  from scipy.sparse.csgraph import connected_components as cc
  v = x.reshape(x.shape + (1,) * (3 - x.ndim), order="F")
  sx, sy, sz = v.shape
  idx = np.arange(v.size, dtype=np.int64).reshape(v.shape, order="F")
  if connectivity in (4, 6):
    offs = [(-1, 0, 0), (0, -1, 0), (0, 0, -1)]
  elif connectivity == 8:
    offs = [(-1, 0, 0), (0, -1, 0), (-1, -1, 0), (1, -1, 0)]
  else:
    offs = [(dx, dy, dz) for dz in (-1, 0) for dy in (-1, 0, 1) for dx in (-1, 0, 1)
            if (dz, dy, dx) < (0, 0, 0) and (connectivity == 26 or abs(dx) + abs(dy) + abs(dz) <= 2)]
  rows, cols = [], []
  fg = v != 0
  for dx, dy, dz in offs:
    if periodic:
      q = np.roll(v, (-dx, -dy, -dz), axis=(0, 1, 2)); qi = np.roll(idx, (-dx, -dy, -dz), axis=(0, 1, 2))
      p, pi = v, idx
    else:
      def sl(d, n):
        return (slice(max(0, -d), n - max(0, d)), slice(max(0, d), n - max(0, -d)))
      (px, qx), (py, qy), (pz, qz) = sl(dx, sx), sl(dy, sy), sl(dz, sz)
      p, pi = v[px, py, pz], idx[px, py, pz]
      q, qi = v[qx, qy, qz], idx[qx, qy, qz]
    both = (p != 0) & (q != 0)
    if mode == "eq":
      e = both & (p == q)
    elif mode == "nonzero":
      e = both
    else:
      if np.issubdtype(v.dtype, np.floating):
        e = both & (np.abs(p - q) <= v.dtype.type(delta))
      else:
        pp, qq = p.astype(np.int64) if v.dtype.itemsize < 8 else p, q.astype(np.int64) if v.dtype.itemsize < 8 else q
        e = both & (np.where(pp > qq, pp - qq, qq - pp) <= delta)
    rows.append(pi[e]); cols.append(qi[e])
  r = np.concatenate(rows); c = np.concatenate(cols)
  g = sp.coo_matrix((np.ones(r.size, np.int8), (r, c)), shape=(v.size, v.size))
  _, lab = cc(g, directed=False)
  lab = lab.reshape(v.shape, order="F") + 1
  lab[~fg] = 0
  flat = lab.reshape(-1, order="F")
  u, first = np.unique(flat, return_index=True)
  u, first = u[u != 0], first[u != 0]
  remap = np.zeros(flat.max() + 1, dtype=np.int64)
  remap[u[np.argsort(first)]] = np.arange(1, u.size + 1)
  return remap[lab].reshape(x.shape, order="F"), int(u.size)
