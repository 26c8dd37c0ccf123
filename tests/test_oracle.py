"""CPU tests: the oracle (plain-C restatement, oracle/cc3d_oracle.c) is pinned against
(1) golden vectors generated from the unmodified reference (tests/golden/, make_golden.py),
(2) the reference itself when oracle/_ref is built here, and (3) an independent graph spec."""
import numpy as np
import pytest

from helpers import assert_same_labels, call_kwargs, golden_manifest, load_golden, spec_oracle, blobs

MANIFEST = golden_manifest()


@pytest.mark.parametrize("case", MANIFEST, ids=[c["name"] for c in MANIFEST])
def test_oracle_matches_golden(oracle_mod, case):
  x, labels, N, z = load_golden(case["name"])
  out, No = oracle_mod.connected_components(x, return_N=True, **call_kwargs(case["kw"]))
  assert_same_labels(labels, N, out, No, case["name"])


@pytest.mark.parametrize("case", MANIFEST[::3], ids=[c["name"] for c in MANIFEST[::3]])
def test_oracle_statistics_match_golden(oracle_mod, case):
  x, labels, N, z = load_golden(case["name"])
  st = oracle_mod.statistics(labels, no_slice_conversion=True)
  assert np.array_equal(st["voxel_counts"], z["voxel_counts"])
  assert np.array_equal(st["bounding_boxes"], z["bounding_boxes"])
  assert st["bounding_boxes"].dtype == z["bounding_boxes"].dtype
  assert np.array_equal(st["centroids"], z["centroids"], equal_nan=True)


def test_oracle_vs_reference_fuzz(oracle_mod):
  ref = oracle_mod.reference_module()
  if ref is None:
    pytest.skip("oracle/_ref not built (needs /root/reference)")
  rng = np.random.default_rng(7)
  dtypes = [np.uint8, np.uint16, np.uint32, np.uint64, np.int8, np.int64, np.float32, np.float64, bool]
  checked = 0
  for it in range(400):
    dims = int(rng.integers(1, 4))
    shape = tuple(int(rng.integers(1, 13)) for _ in range(dims))
    dt = dtypes[rng.integers(len(dtypes))]
    order = "F" if rng.random() < 0.5 else "C"
    x = (rng.random(shape) < 0.5) if dt == bool else rng.integers(0, 4, shape).astype(dt)
    x = np.asarray(x, order=order)
    conns = [4, 8, 6, 18, 26] if dims == 2 else [6, 18, 26]
    c = conns[rng.integers(len(conns))]
    kw = {}
    mode = int(rng.integers(0, 4))
    fast = x.shape[0] if order == "F" else x.shape[-1]
    if mode == 1:
      kw["binary_image"] = True
      x = np.asarray((x != 0).astype(x.dtype), order=order)
    elif mode == 2 and dt != bool:
      kw["delta"] = int(rng.integers(1, 3))
    elif mode == 3 and c in (4, 8, 6):
      kw["periodic_boundary"] = True
    if (kw.get("binary_image") or dt == bool) and c == 8 and fast % 2 == 1:
      continue  # reference defect D1
    try:
      a, Na = ref.connected_components(x, connectivity=c, return_N=True, **kw)
    except RuntimeError:
      continue  # reference union-find overflow on tiny binary inputs (defect D3)
    b, Nb = oracle_mod.connected_components(x, connectivity=c, return_N=True, **kw)
    assert_same_labels(a, Na, b, Nb, f"{shape} {dt} {order} {c} {kw}")
    checked += 1
  assert checked > 300


@pytest.mark.parametrize("conn,mode", [(26, "eq"), (18, "eq"), (6, "eq"), (26, "nonzero"), (6, "delta"), (26, "delta")])
def test_oracle_vs_graph_spec(oracle_mod, conn, mode):
  rng = np.random.default_rng(conn * 7 + len(mode))
  for shape in [(9, 8, 7), (16, 5, 6), (5, 17, 3)]:
    if mode == "delta":
      x = np.asfortranarray((blobs(rng, shape, 3, 2) * 10 + rng.integers(0, 4, shape)).astype(np.uint8))
      a, Na = oracle_mod.connected_components(x, connectivity=conn, delta=3, return_N=True)
      b, Nb = spec_oracle(x, conn, "delta", 3)
    elif mode == "nonzero":
      x = np.asfortranarray((rng.random(shape) < 0.4).astype(np.uint8))
      a, Na = oracle_mod.connected_components(x, connectivity=conn, binary_image=True, return_N=True)
      b, Nb = spec_oracle(x, conn, "nonzero")
    else:
      x = np.asfortranarray(rng.integers(0, 3, shape).astype(np.uint16))
      a, Na = oracle_mod.connected_components(x, connectivity=conn, return_N=True)
      b, Nb = spec_oracle(x, conn, "eq")
    assert Na == Nb
    assert np.array_equal(a, b)


def test_oracle_periodic_vs_torus_spec(oracle_mod):
  rng = np.random.default_rng(3)
  for conn, shape in [(6, (7, 6, 5)), (4, (9, 8)), (8, (9, 8))]:
    x = np.asfortranarray(rng.integers(0, 3, shape).astype(np.uint8))
    if x.reshape(-1, order="F")[: shape[0]].any() and np.count_nonzero(x.any(axis=0)) <= 1:
      continue
    a, Na = oracle_mod.connected_components(x, connectivity=conn, periodic_boundary=True, return_N=True)
    b, Nb = spec_oracle(x, conn, "eq", periodic=True)
    assert Na == Nb and np.array_equal(a, b)


def test_oracle_dtype_rule_and_errors(oracle_mod):
  x = np.zeros((8, 8, 8), np.uint8)
  assert oracle_mod.connected_components(x).dtype == np.uint16
  assert oracle_mod.connected_components(x, out_dtype=np.uint64).dtype == np.uint64
  with pytest.raises(ValueError):
    oracle_mod.connected_components(x, out_dtype=np.uint8)
  big = (np.arange(41 ** 3, dtype=np.uint32) + 1).reshape(41, 41, 41)
  with pytest.raises(ValueError):
    oracle_mod.connected_components(big, out_dtype=np.uint16)
  assert oracle_mod.connected_components(big).dtype == np.uint32
  with pytest.raises(TypeError):
    oracle_mod.connected_components(np.ones((4, 4), np.float16), delta=1)
  out, N = oracle_mod.connected_components(np.zeros((0, 0), np.uint32), return_N=True)
  assert out.size == 0 and N == 0


def test_oracle_dust_matches_reference(oracle_mod):
  ref = oracle_mod.reference_module()
  rng = np.random.default_rng(11)
  x = np.asfortranarray(rng.integers(0, 3, (20, 20, 20)).astype(np.uint8))
  out, dN = oracle_mod.dust(x, threshold=5, connectivity=6, return_N=True)
  lab, N = oracle_mod.connected_components(x, connectivity=6, return_N=True)
  cnt = oracle_mod.statistics(lab, no_slice_conversion=True)["voxel_counts"]
  keep = cnt >= 5
  keep[0] = False
  assert np.array_equal(out, x * keep[lab])
  assert dN == int(np.count_nonzero(keep))
  if ref is not None:
    assert np.array_equal(ref.statistics(lab, no_slice_conversion=True)["voxel_counts"], cnt)


def _graph_cases(rng, n):
  for it in range(n):
    dims = int(rng.integers(2, 4))
    shape = tuple(int(rng.integers(1, 11)) for _ in range(dims))
    dt = [np.uint8, np.uint16, np.uint32, np.uint64, np.int32, bool][rng.integers(6)]
    x = (rng.random(shape) < 0.5) if dt == bool else rng.integers(0, 3, shape).astype(dt)
    x = np.asarray(x, order="F" if rng.random() < 0.5 else "C")
    conns = [4, 8, 6, 18, 26] if dims == 2 else [6, 18, 26]
    yield x, conns[rng.integers(len(conns))]


def test_oracle_graph_rows_vs_reference(oracle_mod):
  """voxel_connectivity_graph, color_connectivity_graph, largest_k against the reference build."""
  ref = oracle_mod.reference_module()
  if ref is None:
    pytest.skip("oracle/_ref not built (needs /root/reference)")
  rng = np.random.default_rng(17)
  n = 0
  for x, c in _graph_cases(rng, 150):
    a, b = ref.voxel_connectivity_graph(x, connectivity=c), oracle_mod.voxel_connectivity_graph(x, connectivity=c)
    assert a.dtype == b.dtype and a.shape == b.shape and np.array_equal(a, b), (x.shape, x.dtype, c)
    if c in (4, 8, 6, 26):
      ca, Na = ref.color_connectivity_graph(a, connectivity=c, return_N=True)
      cb, Nb = oracle_mod.color_connectivity_graph(a, connectivity=c, return_N=True)
      assert Na == Nb and np.array_equal(ca, cb), (x.shape, x.dtype, c)
      n += 1
  assert n > 60


def test_oracle_largest_k_vs_reference_python_layer(oracle_mod):
  """largest_k lives in the reference's Python layer (cc3d/__init__.py:199-279); it is executed in place from
  /root/reference (nothing copied) to pin the oracle's restatement."""
  pkg = oracle_mod.reference_package()
  if pkg is None:
    pytest.skip("reference Python layer not available (needs /root/reference and oracle/_ref)")
  rng = np.random.default_rng(18)
  checked = 0
  for it in range(80):
    shape = tuple(int(rng.integers(2, 14)) for _ in range(3))
    x = np.asarray(blobs(rng, shape, 6, 2).astype(np.uint32), order="F" if it % 2 else "C")
    for k in (0, 1, 2, 3, 50):
      a = pkg.largest_k(x, k, connectivity=26, return_N=True) if k else (pkg.largest_k(x, k), 0)
      b = oracle_mod.largest_k(x, k, connectivity=26, return_N=True) if k else (oracle_mod.largest_k(x, k), 0)
      assert a[1] == b[1] and a[0].dtype == b[0].dtype and a[0].shape == b[0].shape, (shape, k)
      assert np.array_equal(a[0], b[0]), (shape, k)
      assert a[0].flags.c_contiguous == b[0].flags.c_contiguous and a[0].flags.f_contiguous == b[0].flags.f_contiguous
      checked += 1
    d1, n1 = pkg.dust(x, threshold=5, return_N=True), None
    d2 = oracle_mod.dust(x, threshold=5, return_N=True)
    assert d1[1] == d2[1] and np.array_equal(d1[0], d2[0])
  assert checked == 400


def test_oracle_contacts_vs_reference(oracle_mod):
  """contacts / region_graph against the reference build, including its border behaviour."""
  ref = oracle_mod.reference_module()
  if ref is None:
    pytest.skip("oracle/_ref not built (needs /root/reference)")
  rng = np.random.default_rng(19)
  n = 0
  for it in range(120):
    dims = int(rng.integers(2, 4))
    shape = tuple(int(rng.integers(1, 9)) for _ in range(dims))
    dt = [np.uint8, np.uint16, np.uint32, np.uint64, np.int32][rng.integers(5)]
    x = np.asarray(rng.integers(0, 4, shape).astype(dt), order="F" if rng.random() < 0.5 else "C")
    conns = [4, 8, 6, 18, 26] if dims == 2 else [6, 18, 26]
    c = conns[rng.integers(len(conns))]
    sa = bool(rng.integers(0, 2))
    an = [(1, 1, 1), (4, 4, 40), (2, 3, 5)][rng.integers(3)]
    a = ref.contacts(x, connectivity=c, surface_area=sa, anisotropy=an)
    b = oracle_mod.contacts(x, connectivity=c, surface_area=sa, anisotropy=an)
    assert a == b, (shape, np.dtype(dt), c, sa, an)
    assert ref.region_graph(x, connectivity=c) == oracle_mod.region_graph(x, connectivity=c)
    n += 1
  assert n == 120


def _graph_goldens():
  import glob, os
  return sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "graphs_*.npz")))


@pytest.mark.parametrize("path", _graph_goldens(), ids=[p.split("/")[-1][:-4] for p in _graph_goldens()])
def test_oracle_matches_graph_goldens(oracle_mod, path):
  """SURVEY 8(f) rows against fixtures generated from the reference (tests/golden/make_golden_graphs.py)."""
  z = np.load(path)
  x, c = z["x"], int(z["connectivity"])
  assert np.array_equal(oracle_mod.voxel_connectivity_graph(x, connectivity=c), z["vcg"])
  if "colors" in z:
    col, N = oracle_mod.color_connectivity_graph(z["vcg_cut"], connectivity=c, return_N=True)
    assert N == int(z["colors_N"]) and np.array_equal(col, z["colors"])
  ct = oracle_mod.contacts(x, connectivity=c, surface_area=True, anisotropy=(4, 4, 40))
  want = {(int(a), int(b)): float(v) for (a, b), v in zip(z["contact_pairs"], z["contact_areas"])}
  assert ct == want
  for k in (1, 3):
    if f"largest_{k}" in z:
      got, N = oracle_mod.largest_k(x, k, connectivity=c, return_N=True)
      assert N == int(z[f"largest_{k}_N"]) and got.dtype == z[f"largest_{k}"].dtype and np.array_equal(got, z[f"largest_{k}"])


def _run_cases(rng, n):
  for it in range(n):
    dims = int(rng.integers(1, 4))
    shape = tuple(int(rng.integers(1, 12)) for _ in range(dims))
    dt = [np.uint8, np.uint16, np.uint32, np.uint64, bool][rng.integers(5)]
    x = (rng.random(shape) < 0.5) if dt == bool else blobs(rng, shape, 4, int(rng.integers(1, 4))).astype(dt)
    if it % 7 == 0:
      x = (np.zeros(shape, dtype=np.uint8) + (it % 2)).astype(dt)   # constant volumes: all background / one run
    yield np.asarray(x, order="F" if rng.random() < 0.5 else "C")


def test_oracle_runs_draw_each_vs_reference(oracle_mod):
  """SURVEY 8(f)4 (runs / draw / each) against the reference build, incl. the one-voxel quirk and invalid runs."""
  ref = oracle_mod.reference_module()
  if ref is None:
    pytest.skip("oracle/_ref not built (needs /root/reference)")
  rng = np.random.default_rng(23)
  n = 0
  for x in _run_cases(rng, 200):
    a, b = ref.runs(x), oracle_mod.runs(x)
    assert a == b and list(a.keys()) == list(b.keys()), (x.shape, x.dtype)
    for binary in (False, True):
      for in_place in (False, True):
        got = [(k, im.copy()) for k, im in oracle_mod.each(x, binary=binary, in_place=in_place)]
        want = [(k, im.copy()) for k, im in ref.each(x, binary=binary, in_place=in_place)]
        assert len(got) == len(want)
        for (ka, ia), (kb, ib) in zip(got, want):
          assert ka == kb and ia.dtype == ib.dtype and ia.shape == ib.shape and np.array_equal(ia, ib)
          assert ia.flags.f_contiguous == ib.flags.f_contiguous and ia.flags.c_contiguous == ib.flags.c_contiguous
    if a:
      k = list(a.keys())[-1]
      c1 = np.full(x.shape, 1, dtype=x.dtype, order="F" if x.flags.f_contiguous else "C")
      c2 = c1.copy(order="K")
      r1, r2 = ref.draw(0, a[k], c1), oracle_mod.draw(0, a[k], c2)
      assert np.array_equal(c1, c2) and r1.shape == r2.shape and np.array_equal(r1, r2)
    n += 1
  assert n == 200
  one = np.zeros((1, 1, 1), np.uint16)
  assert ref.runs(one) == oracle_mod.runs(one) == {0: [(0, 1)]}
  for f in (ref.runs, oracle_mod.runs):
    with pytest.raises(IndexError):
      f(np.zeros((0,), np.uint8))
  for bad in ([(3, 3)], [(5, 2)], [(0, 65)], [(0, 4), (70, 71)]):
    for f in (ref.draw, oracle_mod.draw):
      with pytest.raises(RuntimeError, match="Invalid run"):
        f(1, bad, np.zeros((8, 8), np.uint8))
  with pytest.raises(TypeError):
    oracle_mod.runs(np.zeros((3, 3), np.float32))
  with pytest.raises(TypeError):
    ref.runs(np.zeros((3, 3), np.float32))


def _run_goldens():
  import glob, os
  return sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "runs_*.npz")))


def load_run_golden(path):
  z = np.load(path)
  x = np.asarray(z["x"], order="F" if bool(z["f_order"]) else "C")
  keys, off, tab = z["keys"].tolist(), z["offsets"].tolist(), z["table"].tolist()
  want = {int(k): [tuple(p) for p in tab[off[i]:off[i + 1]]] for i, k in enumerate(keys)}
  return z, x, want


def each_checksums(it, order):
  rows = []
  for label, img in it:
    f = np.asarray(img).reshape(-1, order=order)
    nz = np.flatnonzero(f)
    rows.append((label, nz.size, int(nz.sum() % (1 << 61)), int(f[nz[0]]) if nz.size else 0))
  return np.array(rows, dtype=np.uint64).reshape(-1, 4)


@pytest.mark.parametrize("path", _run_goldens(), ids=[p.split("/")[-1][:-4] for p in _run_goldens()])
def test_oracle_matches_run_goldens(oracle_mod, path):
  """runs / draw / each against fixtures generated from the reference (tests/golden/make_golden_runs.py)."""
  z, x, want = load_run_golden(path)
  order = "F" if bool(z["f_order"]) else "C"
  got = oracle_mod.runs(x)
  assert got == want and list(got.keys()) == list(want.keys())
  canvas = np.full(x.shape, 7 if x.dtype != np.bool_ else 0, dtype=x.dtype, order=order)
  oracle_mod.draw(int(z["draw_value"]), want[int(z["draw_key"])], canvas)
  assert np.array_equal(canvas, z["drawn"])
  for binary in (False, True):
    for in_place in (False, True):
      assert np.array_equal(each_checksums(oracle_mod.each(x, binary=binary, in_place=in_place), order), z[f"each_{int(binary)}"])


def test_oracle_float_special_values_vs_reference(oracle_mod):
  """NaN / +-inf / -0.0 / float16 inputs: the C restatement follows the reference, including the 26-connected
  continuous scan's equal-voxel-below shortcut (cc3d_continuous.hpp:147-150) that joins inf with inf."""
  ref = oracle_mod.reference_module()
  if ref is None:
    pytest.skip("oracle/_ref not built")
  rng = np.random.default_rng(77)
  n = 0
  for it in range(45):
    dims = int(rng.integers(2, 4))
    shape = tuple(int(rng.integers(2, 30)) for _ in range(dims))
    dt = [np.float16, np.float32, np.float64][it % 3]
    vals = np.array([0.0, -0.0, 1.0, 2.0, 2.5, np.inf, -np.inf, np.nan, 1e-3, 65000.0], dtype=dt)
    x = np.asarray(vals[rng.integers(0, len(vals), shape)], order="F" if rng.random() < 0.5 else "C")
    conns = [4, 6, 18, 26] if x.ndim == 2 else [6, 18, 26]
    c = int(conns[rng.integers(len(conns))])
    for kw in ([dict()] if dt == np.float16 else [dict(), dict(delta=float(rng.choice([0.5, 1.0, 1e30])))]):
      try:
        want, Nw = ref.connected_components(x, connectivity=c, return_N=True, **kw)
      except (RuntimeError, ValueError):
        continue
      got, N = oracle_mod.connected_components(x, connectivity=c, return_N=True, **kw)
      assert N == Nw and got.dtype == want.dtype and np.array_equal(got, want), (shape, dt, c, kw)
      n += 1
  assert n > 50
