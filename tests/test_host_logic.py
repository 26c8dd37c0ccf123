"""CPU tests of the host-side logic around the run index: grouping of the run table by label, the
memory-order flattening rule, argument conversion. No GPU work: the kernels are covered by tests/test_runs_gpu.py."""
import numpy as np
import pytest


def test_stable_argsort_matches_numpy(cc3d):
  rng = np.random.default_rng(1)
  for hi in (3, 1 << 8, (1 << 16) - 1, 1 << 16, 70000, 1 << 31, (1 << 32) - 1, 1 << 32, 1 << 40, (1 << 64) - 1):
    for n in (0, 1, 2, 1000, 40000):
      v = rng.integers(0, hi, n, dtype=np.uint64, endpoint=True)
      if n > 10:
        v[:5] = hi   # the largest value decides the number of digit passes
      got = cc3d._stable_argsort_u64(v)
      assert np.array_equal(got, np.argsort(v, kind="stable")), (hi, n)


def test_group_runs_is_the_std_map_order(cc3d):
  rng = np.random.default_rng(2)
  for hi in (5, 70000, 1 << 45):
    k = 5000
    values = rng.integers(1, hi, k).astype(np.uint64)
    starts = np.arange(k, dtype=np.uint64) * 3
    ends = starts + 2
    lab, off, s, e = cc3d._group_runs(values, starts, ends)
    assert np.array_equal(lab, np.unique(values)) and off[0] == 0 and off[-1] == k and len(off) == len(lab) + 1
    for i, l in enumerate(lab):
      m = values == l
      assert np.array_equal(s[off[i]:off[i + 1]], starts[m]) and np.array_equal(e[off[i]:off[i + 1]], ends[m])
  lab, off, s, e = cc3d._group_runs(np.zeros(0, np.uint64), np.zeros(0, np.uint64), np.zeros(0, np.uint64))
  assert lab.size == 0 and off.tolist() == [0]


def test_flat_memory_order_follows_the_reference_reshape(cc3d, oracle_mod):
  ref = oracle_mod.reference_module()
  base = np.arange(4 * 5 * 6, dtype=np.uint16).reshape(4, 5, 6)
  cases = [base, np.asfortranarray(base), base[::2], base[:, 1::2, :], base.transpose(2, 0, 1), base[0], base[0, 0],
           np.asfortranarray(base)[:, :, 2], base.reshape(4, 30)[:, :1]]
  for a in cases:
    flat = cc3d._flat_memory_order(a)
    assert flat.ndim == 1 and flat.size == a.size
    assert np.array_equal(flat, oracle_mod._flat_memory_order(a))
    if a.flags.f_contiguous or a.flags.c_contiguous:
      assert np.shares_memory(flat, a)
    if ref is not None:
      assert np.array_equal(flat, ref._reshape(a, (a.size,)))


def test_draw_argument_conversion(cc3d):
  s, e = cc3d._runs_as_arrays([(1, 4), (10, 12)])
  assert s.dtype == e.dtype == np.uint64 and s.tolist() == [1, 10] and e.tolist() == [4, 12]
  s, e = cc3d._runs_as_arrays([])
  assert s.size == 0 and e.size == 0
  s, e = cc3d._runs_as_arrays(np.array([[3, 9]], dtype=np.int64))
  assert s.tolist() == [3] and e.tolist() == [9]
  with pytest.raises(OverflowError):
    cc3d._runs_as_arrays([(-1, 4)])     # size_t conversion in the reference raises too
  assert cc3d._label_for_image(7, np.bool_) == 1 and cc3d._label_for_image(0, np.bool_) == 0
  assert cc3d._label_for_image(255, np.uint8) == 255
  for bad in (256, -1):
    with pytest.raises(OverflowError):
      cc3d._label_for_image(bad, np.uint8)
  for dt in (np.float32, np.int16):
    with pytest.raises(TypeError):
      cc3d._run_dtype(dt)
  assert cc3d._run_dtype(np.bool_) == np.uint8


def test_fastcc3d_namespace(cc3d):
  """cc3d.fastcc3d is the compiled Cython boundary (fastcc3d.pyx); like the reference's extension it carries the entry
  points callers reach for there (cc3d/__init__.py:6-17, 268-275), runs / draw / _erase forwarding to the package."""
  fc = cc3d.fastcc3d
  assert fc is not None and fc.__file__.endswith(".so")
  for name in ("connected_components", "statistics", "estimate_provisional_labels", "runs", "draw", "erase", "_erase"):
    assert callable(getattr(fc, name)), name


def test_runs_and_draw_fail_loudly_without_gpu(cc3d):
  try:
    import torch
    if torch.cuda.is_available():
      pytest.skip("GPU present")
  except ImportError:
    pass
  x = np.ones((4, 4), np.uint8)
  with pytest.raises(cc3d.CC3DB200Error):
    cc3d.runs(x)
  with pytest.raises(cc3d.CC3DB200Error):
    cc3d.draw(1, [(0, 3)], x)
  with pytest.raises(cc3d.CC3DB200Error):
    list(cc3d.each(x))
  with pytest.raises(RuntimeError):   # the upload of the first slab already fails (torch), nothing is computed on the CPU
    cc3d.connected_components_stack([np.ones((4, 4, 2), np.uint8)])


class _DLPackOnly:
  """An array type of another library: nothing but the DLPack protocol."""

  def __init__(self, t):
    self._t = t

  def __dlpack__(self, *args, **kwargs):
    return self._t.__dlpack__(*args, **kwargs)

  def __dlpack_device__(self):
    return self._t.__dlpack_device__()


def test_foreign_arrays_are_adopted_through_dlpack(cc3d):
  import torch
  t = torch.arange(24, dtype=torch.int32).reshape(2, 3, 4)
  got = cc3d._adopt_device_array(_DLPackOnly(t))
  assert isinstance(got, torch.Tensor) and got.data_ptr() == t.data_ptr() and torch.equal(got, t)   # zero copy
  x = np.zeros((3, 3), np.uint8)
  assert cc3d._adopt_device_array(x) is x and cc3d._adopt_device_array(t) is t
  assert cc3d._adopt_device_array([1, 2]) == [1, 2]
  if not torch.cuda.is_available():
    # the adopted tensor takes the normal path: on a box without a GPU that is the loud failure, not an AttributeError
    with pytest.raises(cc3d.CC3DB200Error):
      cc3d.connected_components(_DLPackOnly(torch.ones((4, 4, 4), dtype=torch.uint8)))


def test_dust_bounds_accept_non_finite_thresholds():
  """the reference compares sizes with the threshold directly, so inf / nan bounds are legal."""
  import numpy as np
  from cc3d_b200 import _dust_bounds
  big = 1 << 62
  assert _dust_bounds(100) == (100, big)
  assert _dust_bounds((10, np.inf)) == (10, big) and _dust_bounds((10.5, float("inf"))) == (11, big)
  assert _dust_bounds((-np.inf, 5)) == (-big, 5)
  assert _dust_bounds(np.float32(3.2)) == (4, big)
  lo, hi = _dust_bounds((np.nan, 5))
  assert lo > hi                       # empty range: every component is outside it
  assert _dust_bounds(float("nan")) == (-big, big)   # `size < nan` is false: nothing is dust
