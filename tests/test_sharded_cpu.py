"""CPU tests of the sharded (z-slab) host logic: two gloo ranks label one slab each through
cc3d_b200.sharded.connected_components_slab with an ORACLE backend (oracle/ is test infrastructure),
and the concatenation must equal the monolithic oracle labelling bit for bit: same partition, same
first-appearance numbering, same N, same out dtype. The compute steps of the product path (CudaBackend)
are covered by the -m gpu tests; this file covers the exchange / union-find / renumbering plumbing."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class OracleBackend:
  """Same interface as cc3d_b200.sharded.CudaBackend, computed with the CPU oracle + numpy."""

  def __init__(self):
    import torch
    self.torch = torch
    from oracle import oracle
    self.oracle = oracle

  def _pred(self, kind_dtype, delta_arr, binary_image):
    d = delta_arr[0]
    if binary_image:
      return lambda p, q: (p != 0) & (q != 0)
    if d == 0:
      return lambda p, q: (p == q) & (p != 0)
    def f(p, q):
      hi, lo = np.maximum(p, q), np.minimum(p, q)
      return (p != 0) & (q != 0) & ((hi - lo) <= d)
    return f

  def resolve(self, slab, kind, connectivity, delta_arr, binary_image):
    x = slab.numpy()
    kw = dict(connectivity=connectivity, return_N=True, out_dtype=np.uint32)
    if binary_image:
      kw["binary_image"] = True
    elif delta_arr[0] != 0:
      kw["delta"] = delta_arr[0].item()
    labels, N = self.oracle.connected_components(x, **kw)
    epl = self.oracle.estimate_provisional_labels(x)[0]
    return {"labels": labels, "N": int(N), "epl": int(epl), "shape": tuple(x.shape), "device": slab.device,
            "values": x, "delta": delta_arr, "binary": binary_image}

  def plane_labels(self, h, z):
    return self.torch.from_numpy(h["labels"][z].astype(np.int32))

  def partial_statistics(self, labels, N):
    """Per-slab partial sums in array axes (numpy restatement of the statistics kernel's outputs)."""
    lab = labels.numpy().astype(np.int64)
    counts = np.bincount(lab.ravel(), minlength=N + 1)[: N + 1].astype(np.uint32)
    bbox = np.zeros((N + 1, 3, 2), dtype=np.uint32)
    bbox[:, :, 0] = np.iinfo(np.uint32).max
    sums = np.zeros((N + 1, 3), dtype=np.uint64)
    idx = np.indices(lab.shape)
    for ax in range(3):
      np.minimum.at(bbox[:, ax, 0], lab.ravel(), idx[ax].ravel().astype(np.uint32))
      np.maximum.at(bbox[:, ax, 1], lab.ravel(), idx[ax].ravel().astype(np.uint32))
      sums[:, ax] = np.bincount(lab.ravel(), weights=idx[ax].ravel(), minlength=N + 1)[: N + 1].astype(np.uint64)
    return counts, bbox.reshape(N + 1, 6), sums

  def face_pairs(self, vals_upper, labs_upper, vals_lower, labs_lower, kind, connectivity, delta_arr, binary_image):
    P, Q = vals_upper.numpy(), vals_lower.numpy()
    lP, lQ = labs_upper.numpy().astype(np.int64), labs_lower.numpy().astype(np.int64)
    pred = self._pred(P.dtype, delta_arr, binary_image)
    sy, sx = P.shape
    out = []
    for dy in (-1, 0, 1):
      for dx in (-1, 0, 1):
        nz = (dx != 0) + (dy != 0)
        if connectivity == 6 and nz > 0:
          continue
        if connectivity == 18 and nz > 1:
          continue
        ys = slice(max(0, -dy), sy - max(0, dy)); yq = slice(max(0, dy), sy - max(0, -dy))
        xs = slice(max(0, -dx), sx - max(0, dx)); xq = slice(max(0, dx), sx - max(0, -dx))
        m = pred(P[ys, xs], Q[yq, xq])
        out.append(((lQ[yq, xq][m] << 32) | lP[ys, xs][m]).ravel())
    return self.torch.from_numpy(np.concatenate(out) if out else np.zeros(0, np.int64))

  def solve_pairs(self, n_nodes, a, b):
    parent = list(range(n_nodes))
    def find(i):
      while parent[i] != i:
        parent[i] = parent[parent[i]]
        i = parent[i]
      return i
    for x, y in zip(a.tolist(), b.tolist()):
      rx, ry = find(x), find(y)
      if rx < ry:
        parent[ry] = rx
      elif ry < rx:
        parent[rx] = ry
    return self.torch.tensor([find(i) for i in range(n_nodes)], dtype=self.torch.int64)

  def write_remap(self, h, remap, max_label, out_dtype):
    r = remap.numpy()
    assert int(r.max(initial=0)) <= np.iinfo(out_dtype).max
    out = r[h["labels"].astype(np.int64)].astype(out_dtype)
    return self.torch.from_numpy(out.astype(np.int64))  # torch has limited unsigned support on CPU

  def release(self, h):
    pass

  # streaming front end (connected_components_stack)
  def to_device(self, slab_np):
    return self.torch.from_numpy(np.ascontiguousarray(slab_np))

  def local_labels_host(self, h):
    return np.ascontiguousarray(h["labels"].astype(np.uint32))

  def remap_host(self, local, table, out):
    out[...] = table[local.astype(np.int64)].astype(out.dtype)


def _free_port():
  s = socket.socket()
  s.bind(("127.0.0.1", 0))
  p = s.getsockname()[1]
  s.close()
  return p


def _worker(rank, world, port, cases, results):
  sys.path.insert(0, os.path.join(ROOT, "connected-components-3d_b200"))
  sys.path.insert(0, ROOT)
  sys.path.insert(0, os.path.join(ROOT, "tests"))
  import torch
  import torch.distributed as dist
  os.environ["MASTER_ADDR"] = "127.0.0.1"
  os.environ["MASTER_PORT"] = str(port)
  dist.init_process_group("gloo", rank=rank, world_size=world)
  from cc3d_b200 import sharded
  backend = OracleBackend()
  try:
    for ci, (vol, kw) in enumerate(cases):
      bounds = np.linspace(0, vol.shape[0], world + 1).astype(int)
      slab = torch.from_numpy(np.ascontiguousarray(vol[bounds[rank]:bounds[rank + 1]]))
      out, N = sharded.connected_components_slab(slab, return_N=True, backend=backend, **kw)
      st = sharded.statistics_slab(out.to(torch.int64), N, no_slice_conversion=True, backend=backend)
      results.put((ci, rank, out.numpy(), N, st))
  finally:
    dist.destroy_process_group()


def _make_cases():
  rng = np.random.default_rng(7)
  cases = []
  for it, conn in enumerate((6, 18, 26, 26, 6)):
    shape = (int(rng.integers(6, 14)), int(rng.integers(5, 20)), int(rng.integers(5, 40)))
    coarse = rng.integers(0, 4, tuple((s + 2) // 3 for s in shape))
    vol = np.repeat(np.repeat(np.repeat(coarse, 3, 0), 3, 1), 3, 2)[:shape[0], :shape[1], :shape[2]]
    kw = dict(connectivity=conn)
    if it == 3:
      vol = (rng.random(shape) < 0.45).astype(np.uint8)
      kw["binary_image"] = True
    elif it == 4:
      vol = (vol * 10 + rng.integers(0, 3, shape)) * (vol != 0)
      kw["delta"] = 2
    cases.append((np.ascontiguousarray(vol.astype(np.uint8 if it == 3 else np.int32)), kw))
  return cases


@pytest.mark.parametrize("world", [2, 3])
def test_slabs_equal_monolithic_labelling(world):
  import torch.multiprocessing as mp
  sys.path.insert(0, ROOT)
  from oracle import oracle
  cases = _make_cases()
  ctx = mp.get_context("spawn")
  results = ctx.Queue()
  port = _free_port()
  procs = [ctx.Process(target=_worker, args=(r, world, port, cases, results)) for r in range(world)]
  for p in procs:
    p.start()
  got = {}
  for _ in range(world * len(cases)):
    ci, rank, out, N, st = results.get(timeout=120)
    got[(ci, rank)] = (out, N, st)
  for p in procs:
    p.join(timeout=60)
    assert p.exitcode == 0
  for ci, (vol, kw) in enumerate(cases):
    want, Nw = oracle.connected_components(vol, return_N=True, **kw)
    parts = [got[(ci, r)][0] for r in range(world)]
    for r in range(world):
      assert got[(ci, r)][1] == Nw, f"case {ci} rank {r}: N {got[(ci, r)][1]} != {Nw}"
    whole = np.concatenate(parts, axis=0)
    assert np.array_equal(whole, want.astype(np.int64)), f"case {ci} ({kw}) differs from the monolithic labelling"
    # statistics_slab: every rank holds the statistics of the whole volume
    ref_st = oracle.statistics(want, no_slice_conversion=True)
    for r in range(world):
      st = got[(ci, r)][2]
      assert np.array_equal(st["voxel_counts"], ref_st["voxel_counts"]), f"case {ci} rank {r}: counts"
      assert st["bounding_boxes"].dtype == ref_st["bounding_boxes"].dtype
      assert np.array_equal(st["bounding_boxes"], ref_st["bounding_boxes"]), f"case {ci} rank {r}: boxes"
      assert np.array_equal(st["centroids"], ref_st["centroids"], equal_nan=True), f"case {ci} rank {r}: centroids"


def test_single_process_slabs_equal_monolithic_labelling():
  """connected_components_slabs: same merge without a process group (virtual slabs on one device)."""
  sys.path.insert(0, os.path.join(ROOT, "connected-components-3d_b200"))
  sys.path.insert(0, ROOT)
  import torch
  from cc3d_b200 import sharded
  from oracle import oracle
  backend = OracleBackend()
  for vol, kw in _make_cases():
    for nslab in (1, 2, 4):
      bounds = np.linspace(0, vol.shape[0], nslab + 1).astype(int)
      slabs = [torch.from_numpy(np.ascontiguousarray(vol[bounds[r]:bounds[r + 1]])) for r in range(nslab)]
      outs, N = sharded.connected_components_slabs(slabs, return_N=True, backend=backend, **kw)
      want, Nw = oracle.connected_components(vol, return_N=True, **kw)
      assert N == Nw
      assert np.array_equal(np.concatenate([o.numpy() for o in outs], 0), want.astype(np.int64)), (kw, nslab)


def test_native_merge_equals_numpy_merge():
  """cc3d_b200_merge_slabs (C++ host function in libcc3d_b200.so) against the numpy restatement in
  cc3d_b200.sharded._global_numbering on random interface graphs."""
  sys.path.insert(0, os.path.join(ROOT, "connected-components-3d_b200"))
  from cc3d_b200 import sharded
  rng = np.random.default_rng(3)
  for it in range(200):
    world = int(rng.integers(1, 6))
    N_r = rng.integers(0, 12, world)
    pair_lists = [np.zeros(0, dtype=np.int64)]
    for r in range(1, world):
      n = int(rng.integers(0, 15)) if N_r[r - 1] > 0 and N_r[r] > 0 else 0
      lo = rng.integers(1, N_r[r - 1] + 1, n) if n else np.zeros(0, dtype=np.int64)
      up = rng.integers(1, N_r[r] + 1, n) if n else np.zeros(0, dtype=np.int64)
      pair_lists.append(((lo.astype(np.int64) << 32) | up.astype(np.int64)))
    Nw, remaps = sharded._global_numbering(N_r, pair_lists, range(world))
    for r in range(world):
      Ng, remap = sharded._merge_native(N_r, pair_lists, r)
      assert Ng == Nw and np.array_equal(remap, remaps[r]), (it, r, N_r, pair_lists)


def _stack_cases():
  rng = np.random.default_rng(11)
  for it, conn in enumerate((26, 6, 18, 26, 6, 26)):
    sx, sy = int(rng.integers(3, 20)), int(rng.integers(3, 16))
    depths = [int(d) for d in rng.integers(1, 7, int(rng.integers(1, 6)))]
    sz = sum(depths)
    coarse = rng.integers(0, 4, ((sx + 2) // 3, (sy + 2) // 3, (sz + 2) // 3))
    vol = np.repeat(np.repeat(np.repeat(coarse, 3, 0), 3, 1), 3, 2)[:sx, :sy, :sz]
    kw = dict(connectivity=conn)
    if it == 3:
      vol = rng.random((sx, sy, sz)) < 0.45
      kw["binary_image"] = True
    vol = vol.astype([np.uint32, np.uint8, np.int64, np.bool_, np.uint16, np.uint64][it])
    cuts = np.cumsum([0] + depths)
    images = [np.asarray(vol[:, :, a:b], order="C" if (i + it) % 2 else "F") for i, (a, b) in enumerate(zip(cuts[:-1], cuts[1:]))]
    if it == 4:
      images.insert(1, np.zeros((sx, sy, 0), vol.dtype))          # empty image: skipped
      images.append(vol[:, :, -1].copy() * 0 + vol[:, :, -1])     # a 2-D image is a slab of depth 1
      vol = np.concatenate([vol, vol[:, :, -1:]], axis=2)
    yield vol, images, kw


def test_stack_streaming_equals_monolithic_labelling():
  """connected_components_stack (streaming front end): the host logic (slab bookkeeping, face
  interfaces, merge, remap, out-dtype rule) with the oracle backend; equals the monolithic labelling of the
  Fortran-ordered concatenation bit for bit."""
  sys.path.insert(0, os.path.join(ROOT, "connected-components-3d_b200"))
  sys.path.insert(0, ROOT)
  from cc3d_b200 import sharded
  from oracle import oracle
  backend = OracleBackend()
  n = 0
  for vol, images, kw in _stack_cases():
    want, Nw = oracle.connected_components(np.asfortranarray(vol), return_N=True, **kw)
    got, N = sharded.connected_components_stack(iter(images), return_N=True, backend=backend, order="F", **kw)
    assert N == Nw and got.dtype == want.dtype and got.shape == want.shape and got.flags.f_contiguous, kw
    assert np.array_equal(got, want), kw
    # default: the result's memory order follows the first image; scratch_dir spills the local labels to memmaps
    import tempfile
    with tempfile.TemporaryDirectory() as td:
      got_k = sharded.connected_components_stack(iter(images), backend=backend, scratch_dir=td, **kw)
      assert os.listdir(td) == []
    first = next(im for im in images if im.size)
    assert got_k.flags.f_contiguous == bool(first.flags.f_contiguous) and np.array_equal(got_k, want), kw
    mm = np.zeros(want.shape, dtype=np.uint64, order="F")
    res = sharded.connected_components_stack(images, out_dtype=np.uint64, out=mm, backend=backend, **kw)
    assert res is mm and np.array_equal(mm, want)
    n += 1
  assert n == 6
  with pytest.raises(ValueError):
    sharded.connected_components_stack([np.zeros((3, 3, 2), np.uint8), np.zeros((3, 4, 2), np.uint8)], backend=backend)
  with pytest.raises(ValueError):
    sharded.connected_components_stack([np.zeros((3, 3, 2), np.uint8)], connectivity=8, backend=backend)
  with pytest.raises(ValueError):
    sharded.connected_components_stack([], backend=backend)


def test_merge_on_the_gathered_buffer_equals_native_merge():
  """_merge_gathered (pointer arithmetic on the all-gathered [N, epl, sz, n_pairs, pairs...] rows of the CUDA fast
  path) against _merge_native on per-slab views and the numpy merge; heavy duplication exercises the pair cache of
  cc3d_b200_merge_slabs."""
  sys.path.insert(0, os.path.join(ROOT, "connected-components-3d_b200"))
  import torch
  from cc3d_b200 import sharded
  rng = np.random.default_rng(5)
  for it in range(120):
    world = int(rng.integers(1, 7))
    cap = int(rng.choice([0, 8, 64, 9000]))
    facts = torch.zeros((world, 4 + cap), dtype=torch.int64).numpy()
    facts[:, 0] = rng.integers(0, 40 if cap < 9000 else 3000, world)
    facts[:, 1] = rng.integers(0, 1000, world)
    facts[:, 2] = rng.integers(1, 9, world)
    for r in range(1, world):
      n = int(rng.integers(0, cap + 1)) if facts[r, 0] and facts[r - 1, 0] else 0
      if n:
        distinct = max(1, n // int(rng.choice([1, 3, 30])))
        lo, up = rng.integers(1, facts[r - 1, 0] + 1, distinct), rng.integers(1, facts[r, 0] + 1, distinct)
        idx = rng.integers(0, distinct, n)
        facts[r, 3] = n
        facts[r, 4:4 + n] = (lo[idx] << 32) | up[idx]
    lists = [facts[r, 4:4 + int(facts[r, 3])] for r in range(world)]
    Nw, remaps = sharded._global_numbering(facts[:, 0], lists, range(world))
    for rank in range(world):
      a = sharded._merge_native(facts[:, 0], lists, rank)
      b = sharded._merge_gathered(facts, rank)
      assert a[0] == b[0] == Nw and np.array_equal(a[1], b[1]) and np.array_equal(a[1], remaps[rank]), (it, rank)
      c = sharded._merge_fast_path(facts, facts[:, 3], rank, world)
      assert c[0] == Nw and np.array_equal(c[1], remaps[rank])
  assert sharded._gathered_merge_state == {"checked": True, "ok": True}
