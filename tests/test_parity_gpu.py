"""GPU parity tests (run with -m gpu on the B200 box). Every call goes through the C-ABI
(libcc3d_b200.so via the Python host layer); results are compared bit-exactly with
 - the golden vectors generated from the unmodified reference (tests/golden/),
 - the plain-C oracle (oracle/cc3d_oracle.c) on seeded random inputs,
 - the reference build itself (oracle/_ref) when it travelled with the repo,
 - size-independent properties at BASELINE.json's full sizes."""
import hashlib

import numpy as np
import pytest

from helpers import assert_same_labels, blobs, call_kwargs, golden_manifest, load_golden, spec_oracle

pytestmark = pytest.mark.gpu
MANIFEST = golden_manifest()


@pytest.mark.parametrize("case", MANIFEST, ids=[c["name"] for c in MANIFEST])
def test_golden_labels(cc3d, case):
  x, labels, N, z = load_golden(case["name"])
  out, No = cc3d.connected_components(x, return_N=True, **call_kwargs(case["kw"]))
  assert_same_labels(labels, N, out, No, case["name"])


@pytest.mark.parametrize("case", MANIFEST, ids=[c["name"] for c in MANIFEST])
def test_golden_statistics(cc3d, case):
  x, labels, N, z = load_golden(case["name"])
  st = cc3d.statistics(labels, no_slice_conversion=True)
  assert st["voxel_counts"].dtype == z["voxel_counts"].dtype
  assert np.array_equal(st["voxel_counts"], z["voxel_counts"])
  assert st["bounding_boxes"].dtype == z["bounding_boxes"].dtype
  assert np.array_equal(st["bounding_boxes"], z["bounding_boxes"])
  assert np.array_equal(st["centroids"], z["centroids"], equal_nan=True)


def _truth(oracle_mod):
  ref = oracle_mod.reference_module()
  return ref if ref is not None else oracle_mod


def _fuzz(cc3d, truth, seed, ncase, maxdim, wide=False):
  rng = np.random.default_rng(seed)
  dtypes = [np.uint8, np.uint16, np.uint32, np.uint64, np.int8, np.int16, np.int32, np.int64, np.float32, np.float64, bool]
  checked = 0
  for it in range(ncase):
    dims = int(rng.integers(1, 4))
    shape = tuple(int(rng.integers(1, maxdim)) for _ in range(dims))
    if dims == 2 and rng.random() < 0.4:
      shape = (int(rng.integers(1, 400)), int(rng.integers(1, 400)))
    dt = dtypes[rng.integers(len(dtypes))]
    order = "F" if rng.random() < 0.5 else "C"
    if wide:
      # long fast axis: full-width union tiles (16 words) with staged halos; multiples of 128 take the staged
      # (cp.async) face kernel, the others its bounds-checked path
      dims = int(rng.integers(2, 4))
      fast = int(rng.choice([128, 256, 384, 512, 640, 500, 530, 700, int(rng.integers(130, 720))]))
      rest = (int(rng.integers(1, 40)),) if dims == 2 else (int(rng.integers(1, 22)), int(rng.integers(1, 14)))
      shape = (fast,) + rest if order == "F" else rest[::-1] + (fast,)
    nvals = int(rng.integers(2, 6))
    if dt == bool:
      x = rng.random(shape) < rng.random()
    elif rng.random() < 0.4:
      x = blobs(rng, shape, nvals, int(rng.integers(2, 6))).astype(dt)
    else:
      x = rng.integers(0, nvals, shape).astype(dt)
    conns = [4, 8, 6, 18, 26] if dims == 2 else [6, 18, 26]
    c = conns[rng.integers(len(conns))]
    mode = int(rng.integers(0, 4))
    kw = {}
    if mode == 1:
      kw["binary_image"] = True
      x = (x != 0).astype(x.dtype)
    elif mode == 2 and dt != bool:
      if np.issubdtype(dt, np.floating):
        x = ((x * 3 + rng.random(shape) * 2.5) * (x != 0)).astype(dt)
        kw["delta"] = float(rng.random() * 3)
      else:
        x = (x * 3 + rng.integers(0, 3, shape) * (x != 0)).astype(dt)
        kw["delta"] = int(rng.integers(1, 4))
    elif mode == 3 and c in (4, 8, 6):
      kw["periodic_boundary"] = True
    x = np.asarray(x, order=order)
    fast = x.shape[0] if order == "F" else x.shape[-1]
    if (kw.get("binary_image") or dt == bool) and c == 8 and fast % 2 == 1:
      continue  # reference defect D1: stale labels for odd sx
    try:
      a, Na = truth.connected_components(x, connectivity=c, return_N=True, **kw)
    except RuntimeError:
      continue  # reference union-find overflow (defect D3)
    b, Nb = cc3d.connected_components(x, connectivity=c, return_N=True, **kw)
    assert_same_labels(a, Na, b, Nb, f"{shape} {np.dtype(dt)} {order} conn={c} {kw}")
    checked += 1
  return checked


def test_differential_fuzz_small(cc3d, oracle_mod):
  assert _fuzz(cc3d, _truth(oracle_mod), seed=101, ncase=500, maxdim=24) > 400


def test_differential_fuzz_seams(cc3d, oracle_mod):
  """Shapes that cross several 64x8x8 tiles in every axis."""
  assert _fuzz(cc3d, _truth(oracle_mod), seed=202, ncase=150, maxdim=150) > 100


def test_differential_fuzz_wide(cc3d, oracle_mod):
  """Rows of 128..720 voxels: the staged face kernel, full-width union tiles with staged halos, partial last words."""
  assert _fuzz(cc3d, _truth(oracle_mod), seed=505, ncase=160, maxdim=24, wide=True) > 120


def test_long_runs_across_x_tiles(cc3d, oracle_mod):
  """Rows of 1 100 .. 2 100 voxels (three to five union tiles side by side) with runs of hundreds of voxels: a run that
  enters a tile from the left keeps owning the leading bits of the tile's second, third ... word (regression: the
  warp-owned B1 treated those bits as tile-local; only a 2048-wide full-size test saw it)."""
  truth = _truth(oracle_mod)
  rng = np.random.default_rng(77)
  checked = 0
  for it in range(24):
    sx = int(rng.choice([1100, 1536, 2048, 2100, int(rng.integers(1025, 2200))]))
    sy, sz = int(rng.integers(9, 40)), int(rng.integers(1, 11))
    coarse = rng.integers(0 if it % 3 == 0 else 1, 5, (sz // 3 + 1, sy // 5 + 1, sx // int(rng.integers(90, 400)) + 1))
    x = np.repeat(np.repeat(np.repeat(coarse, 3, 0), 5, 1), 400, 2)
    shift = rng.integers(0, 300, (x.shape[0], x.shape[1]))
    rows = np.arange(x.shape[2])[None, None, :] + shift[:, :, None]
    x = np.take_along_axis(x, np.minimum(rows, x.shape[2] - 1), axis=2)[:sz, :sy, :sx]
    x = np.ascontiguousarray(x).astype([np.uint32, np.uint64, np.uint8, np.uint16][it % 4])
    for conn in ((6, 18, 26) if sz > 1 else (4, 8)):
      xx = x if sz > 1 else x[0]
      a, Na = truth.connected_components(xx, connectivity=conn, return_N=True)
      b, Nb = cc3d.connected_components(xx, connectivity=conn, return_N=True)
      assert_same_labels(a, Na, b, Nb, f"{xx.shape} conn={conn} case {it}")
      checked += 1
  assert checked >= 48


def test_binary26_block_nodes(cc3d, oracle_mod):
  """Binary 26-connected volumes solve their unions on the grid of 2x2x2 blocks (cc3d_blocks.cuh): odd and tiny extents
  (partial blocks, one block per row), sparse and dense occupancies, numbering by the first VOXEL of every component."""
  truth = _truth(oracle_mod)
  rng = np.random.default_rng(2626)
  checked = 0
  for it in range(220):
    if it % 4 == 0:
      shape = (int(rng.integers(1, 6)), int(rng.integers(1, 70)), int(rng.integers(1, 70)))
    elif it % 4 == 1:
      shape = (int(rng.integers(1, 140)), int(rng.integers(1, 30)), int(rng.integers(1, 6)))
    else:
      shape = tuple(int(rng.integers(1, 75)) for _ in range(3))
    if it % 7 == 0:
      shape = shape[:2]     # a 2D array labelled with a 3D connectivity
    p = float(rng.choice([0.03, 0.1, 0.2, 0.35, 0.5, 0.8, 0.97]))
    dt = [np.uint8, bool, np.uint16, np.float32, np.int64][it % 5]
    x = np.asarray((rng.random(shape) < p).astype(dt), order="F" if it % 2 else "C")
    try:
      a, Na = truth.connected_components(x, connectivity=26, return_N=True, binary_image=True)
    except RuntimeError:
      continue   # reference union-find overflow (defect D3)
    b, Nb = cc3d.connected_components(x, connectivity=26, return_N=True, binary_image=True)
    assert_same_labels(a, Na, b, Nb, f"{shape} {np.dtype(dt)} p={p} case {it}")
    checked += 1
  assert checked > 150


def test_binary_one_byte_kernel_a(cc3d, oracle_mod):
  """Binary images of 1-byte elements take their own kernel A (k_fg_bitmap_u8 + k_faces_from_fg: byte-parallel foreground
  bits, links as word logic). RAW values (any non-zero byte is foreground), rows that are / are not multiples of 16 and 32
  voxels, misaligned base pointers, every connectivity, the maximum delta (which the reference turns into the binary
  kernel AFTER it has estimated epl on the raw values, fastcc3d.pyx:388-395), and epl itself through the C-ABI."""
  import ctypes
  from cc3d_b200 import _lib
  L = _lib.lib()
  truth = _truth(oracle_mod)
  rng = np.random.default_rng(8080)
  checked = 0
  for it in range(120):
    sx = int(rng.choice([16, 32, 48, 64, 96, 128, 160, 37, 100, 130, 1, 7, int(rng.integers(1, 200))]))
    dims = 2 if it % 5 == 0 else 3
    rest = (int(rng.integers(1, 60)),) if dims == 2 else (int(rng.integers(1, 20)), int(rng.integers(1, 12)))
    dt = [np.uint8, np.int8, bool][it % 3]
    p = float(rng.choice([0.05, 0.3, 0.5, 0.8, 1.0]))
    vals = rng.integers(1, 6 if it % 2 else 256, (sx,) + rest) * (rng.random((sx,) + rest) < p)
    base = np.zeros(vals.size + 3, dtype=np.uint8)
    off = int(rng.integers(0, 4)) if it % 4 == 0 else 0           # misaligned base pointer
    view = base[off:off + vals.size].reshape((sx,) + rest, order="F")
    view[...] = vals.astype(np.uint8)
    x = view.view(np.int8) if dt == np.int8 else ((view != 0) if dt == bool else view)
    conns = [4, 8] if dims == 2 else [6, 18, 26]
    c = conns[it % len(conns)]
    if c == 8 and sx % 2 == 1:
      continue  # reference defect D1
    if c == 4:
      # reference defect: the first row of the binary 2D 4-connected kernel joins EQUAL values instead of non-zero ones
      # (cc3d_binary.hpp:832 `in_labels[loc] == in_labels[loc + B]`), so raw values split runs of row 0 that every other
      # row - and scipy - join; the drop-in joins them. One non-zero value keeps the comparison meaningful.
      view[...] = (view != 0) * np.uint8(7)
      x = view.view(np.int8) if dt == np.int8 else ((view != 0) if dt == bool else view)
    kws = [dict(binary_image=True)]
    if dt != bool:
      kws.append(dict(delta=int(np.iinfo(dt).max)))
    for kw in kws:
      try:
        a, Na = truth.connected_components(x, connectivity=c, return_N=True, **kw)
      except RuntimeError:
        continue  # reference union-find overflow (defect D3)
      b, Nb = cc3d.connected_components(x, connectivity=c, return_N=True, **kw)
      assert_same_labels(a, Na, b, Nb, f"{x.shape} {np.dtype(dt)} off={off} conn={c} {kw} case {it}")
      checked += 1
    if dt == np.uint8:
      # info.epl of the binary call = the reference's transition count on the raw values (cc3d.hpp:287-315)
      sy, sz = (rest[0], 1) if dims == 2 else rest
      info, sess = _lib.ResolveInfo(), ctypes.c_void_p()
      zero = np.zeros(1, np.uint8)
      assert L.cc3d_b200_label_resolve(x.ctypes.data, _lib.U8, sx, sy, sz, 6 if dims == 3 else 4, zero.ctypes.data, 1, 0,
                                       _lib.HOST, None, ctypes.byref(info), ctypes.byref(sess)) == 0
      L.cc3d_b200_session_release(sess)
      assert int(info.epl) == oracle_mod.estimate_provisional_labels(x)[0], f"epl {x.shape} off={off} case {it}"
  assert checked > 120


def test_binary26_block_path_consumers(cc3d, oracle_mod):
  """Block-path sessions (binary 26-connected) keep their labels on the block runs; the consumers of the voxel-level run
  table (fused dust: component sizes from the run table, masked expansion) fill the run labels in on demand."""
  truth = oracle_mod.reference_package() or oracle_mod
  rng = np.random.default_rng(2627)
  for it in range(12):
    shape = tuple(int(rng.integers(3, 90)) for _ in range(3))
    p = float(rng.choice([0.05, 0.15, 0.3, 0.5]))
    img = np.asarray((rng.integers(1, 200, shape) * (rng.random(shape) < p)).astype([np.uint8, np.uint16, np.uint32][it % 3]),
                     order="F" if it % 2 else "C")
    for thr, inv in ((3, False), (6, True), ((2, 40), False)):
      a, Na = truth.dust(img, thr, connectivity=26, binary_image=True, invert=inv, return_N=True)
      b, Nb = cc3d.dust(img, thr, connectivity=26, binary_image=True, invert=inv, return_N=True)
      assert Na == Nb and a.dtype == b.dtype and np.array_equal(a, b), (shape, p, thr, inv)


def test_c_oracle_agrees_too(cc3d, oracle_mod):
  assert _fuzz(cc3d, oracle_mod, seed=303, ncase=200, maxdim=40) > 150


@pytest.mark.parametrize("conn", [6, 18, 26])
@pytest.mark.parametrize("order", ["C", "F"])
def test_3d_all_different(cc3d, conn, order):
  # automated_test.py:261-270
  x = (np.arange(100 * 99 * 98, dtype=np.uint32) + 1).reshape((100, 99, 98), order=order)
  out, N = cc3d.connected_components(x, connectivity=conn, return_N=True)
  assert N == x.size and out.dtype == np.uint32 and out.shape == x.shape
  assert np.array_equal(out.reshape(-1, order=order), np.arange(1, x.size + 1, dtype=np.uint32))


@pytest.mark.parametrize("conn", [6, 18, 26])
def test_scipy_equality_random_binary(cc3d, conn):
  # automated_test.py:499-554: numbering equals scipy.ndimage.label on C-ordered bool volumes
  import scipy.ndimage
  rng = np.random.default_rng(conn)
  x = rng.random((128, 128, 128)) < 0.5
  structure = {6: None, 18: scipy.ndimage.generate_binary_structure(3, 2), 26: np.ones((3, 3, 3))}[conn]
  want, n = scipy.ndimage.label(x, structure=structure)
  out, N = cc3d.connected_components(x, connectivity=conn, return_N=True)
  assert N == n and np.array_equal(out, want)


def test_binary_equals_multilabel_on_mask(cc3d):
  # automated_test.py:1644-1661
  rng = np.random.default_rng(5)
  x = rng.integers(0, 4, (96, 80, 72)).astype(np.uint8)
  for conn in (6, 18, 26):
    a, Na = cc3d.connected_components(x, connectivity=conn, binary_image=True, return_N=True)
    b, Nb = cc3d.connected_components((x > 0).astype(np.uint8), connectivity=conn, return_N=True)
    assert Na == Nb and np.array_equal(a, b)


def test_graph_spec_multilabel_and_continuous(cc3d):
  rng = np.random.default_rng(9)
  x = np.asfortranarray(blobs(rng, (70, 40, 30), 4, 3).astype(np.uint32))
  for conn in (6, 18, 26):
    a, Na = cc3d.connected_components(x, connectivity=conn, return_N=True)
    b, Nb = spec_oracle(x, conn, "eq")
    assert Na == Nb and np.array_equal(a, b)
  xf = np.asfortranarray(((blobs(rng, (70, 40, 30), 3, 5) + 1) * 64 + rng.uniform(-4, 4, (70, 40, 30))).astype(np.float32))
  for conn in (6, 26):
    a, Na = cc3d.connected_components(xf, connectivity=conn, delta=10, return_N=True)
    b, Nb = spec_oracle(xf, conn, "delta", 10)
    assert Na == Nb and np.array_equal(a, b)


def test_dtype_rule_and_special_cases(cc3d):
  assert cc3d.connected_components(np.zeros((64, 64, 64), np.uint8)).dtype == np.uint16
  for od in (np.uint16, np.uint32, np.uint64):
    assert cc3d.connected_components(np.zeros((32, 32, 32), np.uint8), out_dtype=od).dtype == od
  with pytest.raises(ValueError):
    cc3d.connected_components(np.zeros((8, 8, 8), np.uint8), out_dtype=np.uint8)
  with pytest.raises(ValueError):
    cc3d.connected_components((np.arange(41 ** 3, dtype=np.uint32) + 1).reshape(41, 41, 41), out_dtype=np.uint16)
  out = cc3d.connected_components((np.arange(40 ** 3, dtype=np.uint32) + 1).reshape(40, 40, 40))
  assert out.dtype == np.uint16
  out, N = cc3d.connected_components(np.array([[[1, 1, 1, 1]]]), return_N=True)
  assert N == 1 and np.all(out == 1)
  # test_epl_special_case (automated_test.py:453-483)
  x = np.zeros((10, 10, 10), np.uint8, order="F"); x[:, 5, 5] = 1
  assert cc3d.estimate_provisional_labels(x) == (1, 55, 55)
  out = cc3d.connected_components(x)
  assert out.dtype == np.uint16 and np.array_equal(out, x)
  with pytest.raises(TypeError):
    cc3d.connected_components(np.ones((4, 4), np.float16), delta=1)


def test_estimate_provisional_labels_matches_oracle(cc3d, oracle_mod):
  rng = np.random.default_rng(13)
  for dt in (np.uint8, np.uint32, np.int64, np.float32, np.float64):
    for order in "CF":
      x = np.asarray(blobs(rng, (70, 33, 9), 4, 3).astype(dt), order=order)
      assert cc3d.estimate_provisional_labels(x) == oracle_mod.estimate_provisional_labels(x)
  assert cc3d.estimate_provisional_labels(np.zeros((5, 5, 5), np.uint8)) == (0, -1, -1)


def test_statistics_and_dust_against_oracle(cc3d, oracle_mod):
  rng = np.random.default_rng(17)
  for shape, order in [((50, 40, 30), "F"), ((50, 40, 30), "C"), ((300, 200), "F"), ((300, 200), "C")]:
    img = np.asarray(blobs(rng, shape, 5, 3).astype(np.uint32), order=order)
    lab, N = cc3d.connected_components(img, connectivity=6, return_N=True)
    a, b = cc3d.statistics(lab), oracle_mod.statistics(lab)
    assert np.array_equal(a["voxel_counts"], b["voxel_counts"])
    assert a["bounding_boxes"] == b["bounding_boxes"]
    assert np.array_equal(a["centroids"], b["centroids"], equal_nan=True)
    for thr, inv in [(30, False), (30, True), ((10, 200), False), ((10, 200), True), (0, False)]:
      da, na = cc3d.dust(img, thr, connectivity=6, invert=inv, return_N=True)
      db, nb = oracle_mod.dust(img, thr, connectivity=6, invert=inv, return_N=True)
      assert na == nb and da.dtype == db.dtype and np.array_equal(da, db), (shape, order, thr, inv)
  # signed input comes back signed (cc3d/__init__.py:102-103,151)
  img = rng.integers(0, 3, (40, 40, 40)).astype(np.int16)
  out = cc3d.dust(img, 20)
  assert out.dtype == np.int16 and np.array_equal(out, oracle_mod.dust(img, 20))


def test_statistics_absent_labels_and_big_dims(cc3d, oracle_mod):
  lab = np.zeros((10, 10), np.uint32); lab[2:4, 5:9] = 7
  a, b = cc3d.statistics(lab), oracle_mod.statistics(lab)
  assert np.array_equal(a["voxel_counts"], b["voxel_counts"]) and a["bounding_boxes"] == b["bounding_boxes"]
  assert np.array_equal(a["centroids"], b["centroids"], equal_nan=True)
  wide = np.zeros((70000, 2), np.uint8); wide[66000:, 1] = 1
  a = cc3d.statistics(wide, no_slice_conversion=True)
  assert a["bounding_boxes"].dtype == np.uint32 and list(a["bounding_boxes"][1]) == [66000, 69999, 1, 1]


def test_torch_cpu_and_cuda_tensors(cc3d):
  import torch
  rng = np.random.default_rng(21)
  x = blobs(rng, (40, 50, 60), 4, 3).astype(np.int32)
  want, N = cc3d.connected_components(x, return_N=True)
  t = torch.from_numpy(x)
  got, n = cc3d.connected_components(t, return_N=True)          # CPU tensor in -> CPU tensor out
  assert isinstance(got, torch.Tensor) and n == N and np.array_equal(got.numpy(), want)
  got, n = cc3d.connected_components(t.cuda(), return_N=True)   # device tensor: zero copy, stays on device
  assert got.is_cuda and n == N and np.array_equal(got.cpu().numpy(), want)
  tf = t.cuda().permute(2, 1, 0)                                # Fortran-ordered view
  wantf, Nf = cc3d.connected_components(np.asfortranarray(x.transpose(2, 1, 0)), return_N=True)
  gotf, nf = cc3d.connected_components(tf, return_N=True)
  assert nf == Nf and np.array_equal(gotf.cpu().numpy(), wantf)


def test_full_size_properties_512(cc3d):
  """BASELINE configs at full size: size-independent properties (idempotence, binary==multilabel,
  label range, N == number of distinct labels)."""
  import torch
  import benchdata
  x = benchdata.random_binary((512, 512, 512), 0.5, 1, "cuda")
  for conn in (6, 26):
    lab, N = cc3d.connected_components(x, connectivity=conn, return_N=True)
    labb, Nb = cc3d.connected_components(x, connectivity=conn, binary_image=True, return_N=True)
    l64 = lab.to(torch.int64)
    assert N == Nb and torch.equal(l64, labb.to(torch.int64))
    assert int(l64.max()) == N and bool(((l64 > 0) == (x > 0)).all())
    again, N2 = cc3d.connected_components(lab.to(torch.int32), connectivity=conn, return_N=True)
    assert N2 == N and torch.equal(again.to(torch.int64), l64)   # idempotent: relabelling a labelling is the identity
    first = torch.unique(l64[l64 > 0], sorted=False)
    assert first.numel() == N
  v = benchdata.voronoi_multilabel((512, 512, 512), cell=40, seed=2, device="cuda", dtype=torch.int32)
  lab, N = cc3d.connected_components(v, connectivity=26, return_N=True)
  l64 = lab.to(torch.int64).reshape(-1)
  # first-appearance numbering: the running maximum of the labels in memory order increases by at most 1
  cm = torch.cummax(l64, 0).values
  assert int(cm[0]) <= 1 and int((cm[1:] - cm[:-1]).max()) <= 1 and int(cm[-1]) == N


def test_connectomics_known_answers(cc3d):
  """config #1: N and sha256 of the labelling of the reference's own benchmark volume."""
  from oracle import decode_connectomics
  vol = decode_connectomics.load_fixture()
  if vol is None:
    pytest.skip("oracle/_ref/connectomics_512_u32.npz did not travel")
  assert hashlib.sha256(vol.tobytes(order="F")).hexdigest() == decode_connectomics.SHA
  assert cc3d.estimate_provisional_labels(vol) == (4730127, 0, 262143)
  out, N = cc3d.connected_components(vol, connectivity=26, return_N=True)
  assert N == 3619 and out.dtype == np.uint32
  assert hashlib.sha256(out.tobytes(order="F")).hexdigest() == "8e11e41a9a83a8fe3f59f84f016665e6a8d1f9036b9533286f12e04b421c8d49"
  st = cc3d.statistics(out, no_slice_conversion=True)
  assert list(st["voxel_counts"][:5]) == [1018433, 36827584, 3695, 151, 4336]
  assert list(st["bounding_boxes"][1]) == [0, 511, 0, 511, 0, 357]
  assert np.allclose(st["centroids"][1], [281.75905283, 164.67780751, 144.57447961], rtol=0, atol=1e-8)
  out6, N6 = cc3d.connected_components(vol, connectivity=6, return_N=True)
  assert N6 == 4744
  assert hashlib.sha256(out6.tobytes(order="F")).hexdigest() == "cef5d8cecd2ac9b119323dfce62df3020d3be2520e4672003f50c864013a0ccc"
  assert cc3d.connected_components(vol, connectivity=18, return_N=True)[1] == 3657
  _, kept = cc3d.dust(vol, threshold=100, connectivity=26, return_N=True)
  assert kept == 2810


@pytest.mark.parametrize("conn", [6, 18, 26])
def test_virtual_slabs_equal_single_call(cc3d, conn):
  """Sharded machinery on one GPU (cc3d_b200.sharded.connected_components_slabs): per-slab labelling, face pairs,
  global renumbering and the remapped expand must reproduce the monolithic labelling bit for bit."""
  import torch
  from cc3d_b200 import sharded
  rng = np.random.default_rng(conn)
  coarse = rng.integers(0, 5, (20, 17, 23))
  vol = np.repeat(np.repeat(np.repeat(coarse, 5, 0), 5, 1), 5, 2)[:97, :83, :111].astype(np.int32)
  for kw in (dict(), dict(binary_image=True), dict(delta=1)):
    x = vol if "delta" not in kw else (vol * 3 + rng.integers(0, 2, vol.shape)).astype(np.int32) * (vol != 0)
    t = torch.from_numpy(np.ascontiguousarray(x)).cuda()
    want, Nw = cc3d.connected_components(t, connectivity=conn, return_N=True, **kw)
    for cuts in ([0, 97], [0, 40, 97], [0, 1, 2, 50, 96, 97]):
      slabs = [t[a:b] for a, b in zip(cuts[:-1], cuts[1:])]
      outs, N = sharded.connected_components_slabs(slabs, connectivity=conn, return_N=True, **kw)
      got = torch.cat(outs, 0)
      assert N == Nw and got.dtype == want.dtype
      assert torch.equal(got.view(torch.int32) if got.dtype == torch.uint32 else got.to(torch.int64),
                         want.view(torch.int32) if want.dtype == torch.uint32 else want.to(torch.int64)), (conn, kw, cuts)


def test_statistics_and_dust_on_device_tensors(cc3d, oracle_mod):
  """CUDA tensors stay on the device for statistics / dust; results equal the host path and the oracle."""
  import torch
  rng = np.random.default_rng(11)
  coarse = rng.integers(0, 6, (14, 15, 16))
  vol = np.repeat(np.repeat(np.repeat(coarse, 4, 0), 4, 1), 4, 2)[:53, :57, :61].astype(np.uint32)
  vol[rng.random(vol.shape) < 0.02] = 7          # speckle -> small components for dust
  labels, N = oracle_mod.connected_components(vol, return_N=True)
  want = oracle_mod.statistics(labels, no_slice_conversion=True)
  got = cc3d.statistics(torch.from_numpy(labels.astype(np.int32)).cuda(), no_slice_conversion=True)
  assert np.array_equal(got["voxel_counts"], want["voxel_counts"])
  assert np.array_equal(got["bounding_boxes"], want["bounding_boxes"]) and got["bounding_boxes"].dtype == want["bounding_boxes"].dtype
  assert np.array_equal(got["centroids"], want["centroids"], equal_nan=True)
  for kw in (dict(threshold=20), dict(threshold=(5, 50)), dict(threshold=20, invert=True), dict(threshold=30, connectivity=6)):
    a, Na = oracle_mod.dust(vol, return_N=True, **kw)
    b, Nb = cc3d.dust(torch.from_numpy(vol.view(np.int32)).cuda(), return_N=True, **kw)
    assert Na == Nb and np.array_equal(a.view(np.int32), b.cpu().numpy()), kw


def test_edge_queue_overflow_fallback(cc3d, oracle_mod):
  """With a global edge queue of 8 entries every volume that spans several union tiles overflows it; the
  fallback kernel (every edge on the global forest) must give the same labelling."""
  from cc3d_b200 import _lib
  L = _lib.lib()
  L.cc3d_b200_debug_set_queue_capacity(8)
  try:
    assert _fuzz(cc3d, _truth(oracle_mod), seed=404, ncase=120, maxdim=150) > 80
    rng = np.random.default_rng(5)
    x = rng.integers(0, 9, (70, 90, 130)).astype(np.uint32)      # > 16 runs per word: whole tiles go global
    for conn in (6, 26):
      a, Na = _truth(oracle_mod).connected_components(x, connectivity=conn, return_N=True)
      b, Nb = cc3d.connected_components(x, connectivity=conn, return_N=True)
      assert_same_labels(a, Na, b, Nb, f"overflow conn={conn}")
  finally:
    L.cc3d_b200_debug_set_queue_capacity(0)


def test_many_runs_per_word_tiles(cc3d, oracle_mod):
  """Multilabel rows with more than 16 runs per 32-voxel word do not fit the shared-memory forest of a tile;
  those tiles send every edge to the global phase."""
  from cc3d_b200 import _lib
  L = _lib.lib()
  rng = np.random.default_rng(6)
  x = rng.integers(1, 4, (64, 96, 160)).astype(np.uint16)
  x[:, :48] = np.repeat(np.repeat(np.repeat(rng.integers(0, 3, (16, 12, 40)), 4, 0), 4, 1), 4, 2)  # mixed: coarse half
  x2 = rng.integers(0, 3, (300, 700)).astype(np.uint8)                                            # 2D noise
  try:
    for seen in (0, 1):     # 0: every edge of such a tile goes through the global queue; 1: second launch, 32 runs per word
      for conn in (6, 18, 26):
        L.cc3d_b200_debug_set_big_tiles(seen)
        a, Na = _truth(oracle_mod).connected_components(x, connectivity=conn, return_N=True)
        b, Nb = cc3d.connected_components(x, connectivity=conn, return_N=True)
        assert_same_labels(a, Na, b, Nb, f"dense runs conn={conn} seen={seen}")
      for conn, kw in ((4, {}), (8, {}), (8, dict(periodic_boundary=True)), (6, dict(periodic_boundary=True))):
        L.cc3d_b200_debug_set_big_tiles(seen)
        a, Na = _truth(oracle_mod).connected_components(x2, connectivity=conn, return_N=True, **kw)
        b, Nb = cc3d.connected_components(x2, connectivity=conn, return_N=True, **kw)
        assert_same_labels(a, Na, b, Nb, f"dense 2D conn={conn} {kw} seen={seen}")
    # the first noisy volume switches the second launch on for the calls that follow
    L.cc3d_b200_debug_set_big_tiles(0)
    cc3d.connected_components(x, connectivity=26)
    b, Nb = cc3d.connected_components(x, connectivity=26, return_N=True)
    a, Na = _truth(oracle_mod).connected_components(x, connectivity=26, return_N=True)
    assert_same_labels(a, Na, b, Nb, "dense runs after the switch")
  finally:
    L.cc3d_b200_debug_set_big_tiles(0)


def test_voxel_and_color_connectivity_graph(cc3d, oracle_mod):
  """voxel_connectivity_graph and color_connectivity_graph against the reference (bit-exact)."""
  truth = _truth(oracle_mod)
  rng = np.random.default_rng(21)
  dts = [np.uint8, np.uint16, np.uint32, np.uint64, np.int16, np.int64, bool]
  n = 0
  for it in range(160):
    dims = int(rng.integers(2, 4))
    shape = tuple(int(rng.integers(1, 40)) for _ in range(dims)) if it % 8 else ((130, 70, 33) if dims == 3 else (300, 257))
    dt = dts[rng.integers(len(dts))]
    x = (rng.random(shape) < 0.5) if dt == bool else blobs(rng, shape, 4, int(rng.integers(1, 5))).astype(dt)
    x = np.asarray(x, order="F" if rng.random() < 0.5 else "C")
    conns = [4, 8, 6, 18, 26] if dims == 2 else [6, 18, 26]
    c = conns[rng.integers(len(conns))]
    a, b = truth.voxel_connectivity_graph(x, connectivity=c), cc3d.voxel_connectivity_graph(x, connectivity=c)
    assert a.dtype == b.dtype and a.shape == b.shape and np.array_equal(a, b), (shape, np.dtype(dt), c)
    if c in (4, 8, 6, 26):
      # knock a few links out so that the graph is not just the label image again
      g = a.copy()
      g[rng.random(g.shape) < 0.05] &= g.dtype.type(rng.integers(0, 1 << 26) & (0xFF if g.dtype == np.uint8 else 0x3FFFFFF))
      ca, Na = truth.color_connectivity_graph(g, connectivity=c, return_N=True)
      cb, Nb = cc3d.color_connectivity_graph(g, connectivity=c, return_N=True)
      assert Na == Nb and ca.dtype == cb.dtype and np.array_equal(ca, cb), (shape, np.dtype(dt), c)
      n += 1
  assert n > 60
  import torch
  x = blobs(rng, (40, 50, 60), 5, 3).astype(np.int32)
  for c in (6, 26):
    want = truth.voxel_connectivity_graph(x, connectivity=c)
    got = cc3d.voxel_connectivity_graph(torch.from_numpy(x).cuda(), connectivity=c)
    assert got.is_cuda and np.array_equal(got.cpu().numpy().view(want.dtype), want)


def test_largest_k(cc3d, oracle_mod):
  """largest_k (reference path without fastremap) on numpy arrays and CUDA tensors."""
  import torch
  truth = oracle_mod.reference_package() or oracle_mod   # the reference's Python layer when /root/reference is here
  rng = np.random.default_rng(22)
  for it in range(40):
    shape = tuple(int(rng.integers(4, 60)) for _ in range(3))
    x = np.asarray(blobs(rng, shape, 7, int(rng.integers(2, 5))).astype(np.uint32), order="F" if it % 2 else "C")
    for k in (0, 1, 2, 5, 1000):
      a = truth.largest_k(x, k, connectivity=26, return_N=True) if k else (truth.largest_k(x, k), 0)
      b = cc3d.largest_k(x, k, connectivity=26, return_N=True) if k else (cc3d.largest_k(x, k), 0)
      assert a[1] == b[1] and a[0].dtype == b[0].dtype and np.array_equal(a[0], b[0]), (shape, k)
      assert a[0].flags.c_contiguous == b[0].flags.c_contiguous
    if it % 8 == 0:
      t = torch.from_numpy(np.ascontiguousarray(x).view(np.int32)).cuda()
      a = truth.largest_k(np.ascontiguousarray(x), 3, return_N=True)
      b = cc3d.largest_k(t, 3, return_N=True)
      assert a[1] == b[1] and np.array_equal(b[0].cpu().numpy().view(a[0].dtype), a[0])


def test_contacts_and_region_graph(cc3d, oracle_mod):
  """contacts / region_graph against the reference (exact: integer face areas)."""
  truth = _truth(oracle_mod)
  rng = np.random.default_rng(23)
  for it in range(120):
    dims = int(rng.integers(2, 4))
    shape = tuple(int(rng.integers(1, 30)) for _ in range(dims)) if it % 10 else ((90, 70, 40) if dims == 3 else (200, 150))
    dt = [np.uint8, np.uint16, np.uint32, np.uint64, np.int32][rng.integers(5)]
    x = rng.integers(0, 4, shape).astype(dt) if it % 3 else blobs(rng, shape, 6, int(rng.integers(1, 5))).astype(dt)
    x = np.asarray(x, order="F" if rng.random() < 0.5 else "C")
    conns = [4, 8, 6, 18, 26] if dims == 2 else [6, 18, 26]
    c = conns[rng.integers(len(conns))]
    sa = bool(rng.integers(0, 2))
    an = [(1, 1, 1), (4, 4, 40), (2, 3, 5)][rng.integers(3)]
    a = truth.contacts(x, connectivity=c, surface_area=sa, anisotropy=an)
    b = cc3d.contacts(x, connectivity=c, surface_area=sa, anisotropy=an)
    assert a == b, (shape, np.dtype(dt), c, sa, an)
    assert truth.region_graph(x, connectivity=c) == cc3d.region_graph(x, connectivity=c)
  big = np.arange(1, 40 * 40 * 40 + 1, dtype=np.uint32).reshape(40, 40, 40)      # 64 000 labels: the hash table has to grow
  assert truth.contacts(big, connectivity=26) == cc3d.contacts(big, connectivity=26)


def _graph_goldens():
  import glob, os
  return sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "graphs_*.npz")))


@pytest.mark.parametrize("path", _graph_goldens(), ids=[p.split("/")[-1][:-4] for p in _graph_goldens()])
def test_graph_goldens(cc3d, path):
  """SURVEY 8(f) rows against fixtures generated from the reference (incl. largest_k from its Python layer)."""
  z = np.load(path)
  x, c = z["x"], int(z["connectivity"])
  g = cc3d.voxel_connectivity_graph(x, connectivity=c)
  assert g.dtype == z["vcg"].dtype and np.array_equal(g, z["vcg"])
  if "colors" in z:
    col, N = cc3d.color_connectivity_graph(z["vcg_cut"], connectivity=c, return_N=True)
    assert N == int(z["colors_N"]) and col.dtype == z["colors"].dtype and np.array_equal(col, z["colors"])
  ct = cc3d.contacts(x, connectivity=c, surface_area=True, anisotropy=(4, 4, 40))
  assert ct == {(int(a), int(b)): float(v) for (a, b), v in zip(z["contact_pairs"], z["contact_areas"])}
  for k in (1, 3):
    if f"largest_{k}" in z:
      got, N = cc3d.largest_k(x, k, connectivity=c, return_N=True)
      assert N == int(z[f"largest_{k}_N"]) and got.dtype == z[f"largest_{k}"].dtype and np.array_equal(got, z[f"largest_{k}"])
      assert got.flags.c_contiguous == z[f"largest_{k}"].flags.c_contiguous or got.ndim < 2
