"""Runs the REFERENCE's own test-suite (automated_test.py) against a `cc3d` module of our choosing.

The reference's test file is not part of this repository: oracle/build_ref.sh places a copy next to the reference
build in the git-ignored oracle/_ref/ (which travels to the GPU box). Here we only provide
  * a package `cc3d` that re-exports either cc3d_b200 (the drop-in under test) or the reference itself (to validate
    this harness on a machine without a GPU), and
  * a 20-line stand-in for `fastremap` (not installed): `renumber` = first-appearance renumbering in memory order,
    `unique`; any other attribute raises ImportError so that cc3d.largest_k takes its pure-numpy path
    (cc3d/__init__.py:262-266), exactly as SURVEY.md 8(c) describes.
"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SUITE = os.path.join(ROOT, "oracle", "_ref", "automated_test.py")

FASTREMAP_STUB = '''
import numpy as np
def renumber(arr, start=1, preserve_zero=True, in_place=False):
  flat = np.asarray(arr).ravel(order="K")
  u, first = np.unique(flat, return_index=True)
  order = np.argsort(first, kind="stable")
  u = u[order]
  if preserve_zero:
    u = u[u != 0]
  lut = {int(v): i + start for i, v in enumerate(u)}
  if preserve_zero:
    lut[0] = 0
  keys = np.array(sorted(lut), dtype=flat.dtype)
  vals = np.array([lut[int(k)] for k in keys], dtype=np.uint64)
  out = vals[np.searchsorted(keys, flat)].astype(np.uint32 if len(lut) < 2**32 else np.uint64)
  a = np.asarray(arr)
  out = out.reshape(a.shape, order="F" if (a.flags.f_contiguous and not a.flags.c_contiguous) else "C")
  return out, lut
def unique(arr, return_counts=False, **kw):
  return np.unique(arr, return_counts=return_counts)
def __getattr__(name):
  raise ImportError("fastremap stand-in: " + name + " is not provided")
'''

SHIM_B200 = '''
import sys
sys.path.insert(0, {pkg!r})
import cc3d_b200 as _impl
from cc3d_b200 import *
from cc3d_b200 import DimensionError
connected_components_stack = _impl.connected_components_stack
'''

SHIM_REFERENCE = '''
import sys
sys.path.insert(0, {root!r})
from oracle import oracle as _o
_pkg = _o.reference_package()
if _pkg is None:
  raise ImportError("reference package unavailable")
globals().update({{k: v for k, v in _pkg.__dict__.items() if not k.startswith("__")}})
'''


def run(tmpdir, target="b200", k=None, timeout=1500):
  """Returns (returncode, passed, failed, tail of the pytest output)."""
  if not os.path.exists(SUITE):
    return None
  shim = os.path.join(str(tmpdir), "shim")
  os.makedirs(os.path.join(shim, "cc3d"), exist_ok=True)
  with open(os.path.join(shim, "fastremap.py"), "w") as f:
    f.write(FASTREMAP_STUB)
  with open(os.path.join(shim, "cc3d", "__init__.py"), "w") as f:
    f.write((SHIM_B200 if target == "b200" else SHIM_REFERENCE).format(
      pkg=os.path.join(ROOT, "connected-components-3d_b200"), root=ROOT))
  suite = os.path.join(str(tmpdir), "reference_automated_test.py")
  with open(SUITE) as src, open(suite, "w") as dst:
    dst.write(src.read())
  env = dict(os.environ)
  env["PYTHONPATH"] = shim + os.pathsep + env.get("PYTHONPATH", "")
  cmd = [sys.executable, "-m", "pytest", suite, "-q", "-x", "-p", "no:cacheprovider", "--rootdir", str(tmpdir),
         "-o", "python_files=reference_automated_test.py"]
  if k:
    cmd += ["-k", k]
  r = subprocess.run(cmd, capture_output=True, text=True, env=env, cwd=str(tmpdir), timeout=timeout)
  out = r.stdout + r.stderr
  m = re.search(r"(\d+) passed", out)
  f = re.search(r"(\d+) failed", out)
  return r.returncode, int(m.group(1)) if m else 0, int(f.group(1)) if f else 0, out[-3000:]
