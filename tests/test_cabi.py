"""CPU tests of the drop-in boundary: the C-ABI library loads and exports every symbol that
include/cc3d_b200.h declares (no compute calls - there is no GPU here), and argument validation
that happens before any CUDA call behaves like the reference's."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "cc3d_b200.h")


def declared_symbols():
  src = open(HEADER).read()
  src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
  return sorted(set(re.findall(r"\b(cc3d_b200_[a-z_0-9]+)\s*\(", src)))


def test_header_declares_expected_entry_points():
  syms = declared_symbols()
  for s in ["cc3d_b200_prepass", "cc3d_b200_label_resolve", "cc3d_b200_label_write", "cc3d_b200_label",
            "cc3d_b200_statistics", "cc3d_b200_mask_by_label", "cc3d_b200_last_error",
            "cc3d_b200_dust", "cc3d_b200_face_pairs", "cc3d_b200_merge_slabs", "cc3d_b200_slab_begin",
            "cc3d_b200_face_pairs_async", "cc3d_b200_slab_finish", "cc3d_b200_merge_slabs_device", "cc3d_b200_merge_slabs_device_small", "cc3d_b200_voxel_connectivity_graph",
            "cc3d_b200_color_connectivity_graph", "cc3d_b200_contacts", "cc3d_b200_remap_labels",
            "cc3d_b200_runs", "cc3d_b200_draw"]:
    assert s in syms


def test_library_exports_every_declared_symbol():
  from cc3d_b200 import _lib
  assert os.path.exists(_lib.LIB_PATH), "build with python connected-components-3d_b200/build.py"
  L = ctypes.CDLL(_lib.LIB_PATH)
  for s in declared_symbols():
    assert hasattr(L, s), f"{s} declared in include/cc3d_b200.h but not exported"
  assert b"sm_100a" in L.cc3d_b200_version.__call__.__self__.restype.__class__.__name__.encode() or True
  _lib.lib()
  assert "sm_100a" in _lib.lib().cc3d_b200_version().decode()


def test_library_contains_sm100a_code():
  from cc3d_b200 import _lib
  import subprocess
  out = subprocess.run(["cuobjdump", "--list-elf", _lib.LIB_PATH], capture_output=True, text=True)
  if out.returncode != 0:
    pytest.skip("cuobjdump unavailable")
  assert "sm_100a" in out.stdout


def test_python_validation_matches_reference_errors(cc3d):
  with pytest.raises(cc3d.DimensionError):
    cc3d.connected_components(np.zeros((2, 2, 2, 2), np.uint8))
  with pytest.raises(ValueError):
    cc3d.connected_components(np.zeros((4, 4, 4), np.uint8), connectivity=8)
  with pytest.raises(ValueError):
    cc3d.connected_components(np.zeros((4, 4), np.uint8), connectivity=5)
  with pytest.raises(ValueError):
    cc3d.connected_components(np.zeros((4, 4, 4), np.uint8), connectivity=26, periodic_boundary=True)
  with pytest.raises(ValueError):
    cc3d.connected_components(np.zeros((4, 4, 4), np.uint8), connectivity=6, periodic_boundary=True, delta=1)
  out, N = cc3d.connected_components(np.zeros((0, 0), np.uint32), return_N=True)
  assert out.size == 0 and N == 0 and out.dtype == np.uint32
  assert cc3d.statistics(np.zeros((0, 0), np.uint32)) == {"voxel_counts": None, "bounding_boxes": None, "centroids": None}


def test_no_cpu_fallback_without_gpu(cc3d):
  """On a box without a CUDA device the product path must fail loudly, never compute on the CPU."""
  try:
    import torch
    if torch.cuda.is_available():
      pytest.skip("GPU present")
  except ImportError:
    pass
  with pytest.raises(cc3d.CC3DB200Error):
    cc3d.connected_components(np.ones((4, 4, 4), np.uint8))


def test_compiled_cython_binding_loads_and_validates():
  """The Cython boundary (cc3d_b200/fastcc3d.pyx) is compiled by build(), binds the C-ABI through
  `cdef extern from "cc3d_b200.h"` and keeps the reference's argument validation (no compute without a GPU)."""
  import numpy as np
  import pytest
  import cc3d_b200
  fc = cc3d_b200.fastcc3d
  assert fc is not None, "cc3d_b200/fastcc3d extension not built (python connected-components-3d_b200/build.py)"
  assert fc.__file__.endswith(".so") and "sm_100a" in fc.version()
  src = open(os.path.join(os.path.dirname(cc3d_b200.__file__), "fastcc3d.pyx")).read()
  assert 'cdef extern from "cc3d_b200.h"' in src
  for name in ("connected_components", "statistics", "estimate_provisional_labels", "runs", "draw", "erase", "DimensionError"):
    assert hasattr(fc, name), name
  x = np.ones((4, 4, 4), np.uint8)
  with pytest.raises(ValueError, match="Only 6, 18, and 26 connectivities are supported for 3D images"):
    fc.connected_components(x, connectivity=5)
  with pytest.raises(ValueError, match="periodic_boundary is not yet implemented for 26-connectivity"):
    fc.connected_components(x, connectivity=26, periodic_boundary=True)
  with pytest.raises(fc.DimensionError):
    fc.connected_components(np.ones((2, 2, 2, 2), np.uint8))
  with pytest.raises(TypeError):
    fc.connected_components(np.ones((4, 4), np.float16), delta=1)
  try:   # C-ordered input indexes shape[-1] (a module-wide wraparound=False once turned this into a wild read)
    fc.estimate_provisional_labels(np.ones((5, 4), np.uint8))
  except RuntimeError:
    pass   # no GPU here: the C-ABI reports it
  out, N = fc.connected_components(np.zeros((0, 0, 0), np.uint32), return_N=True)      # empty: no GPU needed
  assert N == 0 and out.size == 0
  assert fc.statistics(np.zeros((0, 0), np.uint8)) == {"voxel_counts": None, "bounding_boxes": None, "centroids": None}
