"""Generates tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref, built by
oracle/build_ref.sh from /root/reference). Run in the build container:

    bash oracle/build_ref.sh && python tests/golden/make_golden.py

Each case stores the input, the call arguments and the reference's outputs (labels, N, and for
some cases statistics / dust), so the GPU box (which has no /root/reference) can check parity.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import oracle  # noqa: E402

ref = oracle.reference_module()
assert ref is not None, "build oracle/_ref first (bash oracle/build_ref.sh)"


def blobs(rng, shape, nvals, scale):
  coarse = rng.integers(0, nvals, tuple((s + scale - 1) // scale for s in shape))
  x = coarse
  for ax in range(len(shape)):
    x = np.repeat(x, scale, axis=ax)
  return x[tuple(slice(0, s) for s in shape)]


def cases():
  rng = np.random.default_rng(20240517)
  out = []
  dtypes = [np.uint8, np.uint16, np.uint32, np.uint64, np.int16, np.int64, np.float32, np.float64]
  # multilabel, all connectivities, both orders, odd shapes crossing tile seams
  for i, (shape, conn) in enumerate([((70, 19, 11), 26), ((130, 9, 17), 18), ((65, 33, 9), 6), ((97, 71), 8),
                                     ((140, 67), 4), ((33, 20, 18), 26), ((257,), 26), ((129, 70), 26)]):
    for order in "CF":
      dt = dtypes[(i * 2 + (order == "F")) % len(dtypes)]
      x = np.asarray(blobs(rng, shape, 4, 3).astype(dt), order=order)
      out.append(dict(name=f"multi_{i}_{order}", x=x, kw=dict(connectivity=conn)))
  # random noise multilabel (worst case for unions)
  for i, (shape, conn) in enumerate([((66, 18, 10), 26), ((40, 40, 9), 6), ((100, 80), 8), ((90, 77), 4)]):
    x = rng.integers(0, 3, shape).astype(np.uint32)
    out.append(dict(name=f"noise_{i}", x=np.asarray(x, order="F"), kw=dict(connectivity=conn)))
  # binary
  for i, (shape, conn) in enumerate([((66, 20, 12), 26), ((66, 20, 12), 18), ((66, 20, 12), 6), ((130, 66), 8), ((131, 66), 4)]):
    x = (rng.random(shape) < 0.5)
    out.append(dict(name=f"binary_bool_{i}", x=np.asarray(x, order="F"), kw=dict(connectivity=conn)))
    out.append(dict(name=f"binary_u8_{i}", x=np.asarray(x.astype(np.uint8), order="C" if shape[-1] % 2 == 0 else "F"),
                    kw=dict(connectivity=conn, binary_image=True)))
  # continuous
  for i, (shape, conn, dt, delta) in enumerate([((66, 20, 12), 26, np.float32, 10.0), ((40, 30, 9), 18, np.float64, 7.5),
                                                ((70, 17, 9), 6, np.uint8, 3), ((120, 90), 8, np.float32, 10.0),
                                                ((120, 90), 8, np.uint8, 12), ((100, 75), 4, np.uint16, 5),
                                                ((64, 64), 8, np.float64, 2.5)]):
    tones = blobs(rng, shape, 4, 5) * 64
    if np.issubdtype(dt, np.floating):
      x = (tones + rng.uniform(-4, 4, shape)) * (tones > 0)
    else:
      x = (tones + rng.integers(0, 5, shape)) * (tones > 0)
    out.append(dict(name=f"continuous_{i}", x=np.asarray(x.astype(dt), order="F"), kw=dict(connectivity=conn, delta=delta)))
  # periodic
  for i, (shape, conn) in enumerate([((66, 19, 9), 6), ((70, 66), 4), ((70, 66), 8), ((5, 4), 8), ((64, 8, 8), 6)]):
    x = np.asarray(blobs(rng, shape, 3, 2).astype(np.uint16), order="F")
    out.append(dict(name=f"periodic_{i}", x=x, kw=dict(connectivity=conn, periodic_boundary=True)))
    out.append(dict(name=f"periodic_bin_{i}", x=np.asarray(x != 0, order="F"), kw=dict(connectivity=conn, periodic_boundary=True)))
  # dtype rule / special cases
  out.append(dict(name="all_zero", x=np.zeros((20, 20, 20), np.uint8), kw=dict(connectivity=26)))
  out.append(dict(name="all_one", x=np.ones((20, 20, 20), np.uint8), kw=dict(connectivity=6)))
  out.append(dict(name="all_distinct_u32", x=(np.arange(41 ** 3, dtype=np.uint32) + 1).reshape((41, 41, 41), order="F"), kw=dict(connectivity=26)))
  one_row = np.zeros((40, 7, 5), np.uint32, order="F"); one_row[3:9, 2, 1] = 5; one_row[9:12, 2, 1] = 6; one_row[30:, 2, 1] = 5; one_row[0, 2, 1] = 5
  out.append(dict(name="single_row_periodic", x=one_row, kw=dict(connectivity=6, periodic_boundary=True)))
  out.append(dict(name="out_dtype_u64", x=np.asarray(blobs(rng, (30, 30, 30), 3, 4).astype(np.uint8), order="F"), kw=dict(connectivity=26, out_dtype="uint64")))
  return out


def main():
  manifest = []
  for c in cases():
    x, kw = c["x"], c["kw"]
    call_kw = dict(kw)
    if "out_dtype" in call_kw:
      call_kw["out_dtype"] = np.dtype(call_kw["out_dtype"])
    labels, N = ref.connected_components(x, return_N=True, **call_kw)
    st = ref.statistics(labels, no_slice_conversion=True)
    np.savez_compressed(os.path.join(HERE, c["name"] + ".npz"), x=x, labels=labels, N=np.int64(N),
                        voxel_counts=st["voxel_counts"], bounding_boxes=st["bounding_boxes"], centroids=st["centroids"],
                        f_order=np.bool_(x.flags.f_contiguous and not x.flags.c_contiguous))
    manifest.append(dict(name=c["name"], kw=kw, shape=list(x.shape), dtype=str(x.dtype), N=int(N), out_dtype=str(labels.dtype)))
  with open(os.path.join(HERE, "manifest.json"), "w") as f:
    json.dump(manifest, f, indent=1)
  print(len(manifest), "golden cases written")


if __name__ == "__main__":
  main()
