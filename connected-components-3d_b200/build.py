"""Builds libcc3d_b200.so (hand-written sm_100a CUDA + C-ABI) in-tree with nvcc.

  python connected-components-3d_b200/build.py [--force]

The .so lands next to the Python host layer (cc3d_b200/libcc3d_b200.so); it is git-ignored but
travels to the GPU box with gpurun. Translation units compile in parallel.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
# CC3D_BUILD_TAG=<tag> builds an experimental variant next to the product library (A/B runs on the GPU box:
# CC3D_NVCC_EXTRA="-DFOO=1" CC3D_BUILD_TAG=foo python build.py; CC3D_B200_LIB=.../libcc3d_b200_foo.so python ...)
TAG = os.environ.get("CC3D_BUILD_TAG", "")
OBJ = os.path.join(HERE, "build" + ("_" + TAG if TAG else ""))
OUT = os.path.join(HERE, "cc3d_b200", "libcc3d_b200" + ("_" + TAG if TAG else "") + ".so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
EXTRA = os.environ.get("CC3D_NVCC_EXTRA", "").split()
FLAGS = EXTRA + ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]
UNITS = ["cc3d_b200.cu", "inst_u8.cu", "inst_u16.cu", "inst_u32.cu", "inst_u64.cu", "inst_f32.cu", "inst_f64.cu"]


def _stale(target, deps):
  if not os.path.exists(target):
    return True
  t = os.path.getmtime(target)
  return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
  os.makedirs(OBJ, exist_ok=True)
  headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
  headers.append(os.path.join(HERE, "..", "include", "cc3d_b200.h"))
  jobs = []
  for u in UNITS:
    src = os.path.join(CSRC, u)
    obj = os.path.join(OBJ, u.replace(".cu", ".o"))
    if force or _stale(obj, [src] + headers):
      jobs.append((src, obj))

  def compile_one(job):
    src, obj = job
    cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
      raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
    return r.stderr

  with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
    logs = list(ex.map(compile_one, jobs))
  objs = [os.path.join(OBJ, u.replace(".cu", ".o")) for u in UNITS]
  if jobs or force or _stale(OUT, objs):
    cmd = [NVCC, "-shared", "-o", OUT] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
      raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
  if verbose:
    for l in logs:
      print(l)
  return OUT


def build_binding(force=False):
  """The compiled Cython boundary cc3d_b200/fastcc3d.pyx -> cc3d_b200/fastcc3d<EXT_SUFFIX> (cdef extern from
  include/cc3d_b200.h, linked against libcc3d_b200.so next to it through an $ORIGIN rpath)."""
  import sysconfig
  pkg = os.path.join(HERE, "cc3d_b200")
  pyx = os.path.join(pkg, "fastcc3d.pyx")
  out = os.path.join(pkg, "fastcc3d" + sysconfig.get_config_var("EXT_SUFFIX"))
  hdr = os.path.join(HERE, "..", "include", "cc3d_b200.h")
  lib = os.path.join(pkg, "libcc3d_b200.so")
  if not force and not _stale(out, [pyx, hdr]) and os.path.getmtime(out) >= os.path.getmtime(pyx):
    return out
  os.makedirs(OBJ, exist_ok=True)
  cpp = os.path.join(OBJ, "fastcc3d.cpp")
  r = subprocess.run([sys.executable, "-m", "cython", "-3", "--cplus", pyx, "-o", cpp], capture_output=True, text=True)
  if r.returncode != 0:
    raise RuntimeError(f"cython failed:\n{r.stdout}\n{r.stderr}")
  cmd = ["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-w", "-I", os.path.join(HERE, "..", "include"),
         "-I", sysconfig.get_paths()["include"], cpp, "-o", out, "-L", pkg, "-l:libcc3d_b200.so", "-Wl,-rpath,$ORIGIN"]
  r = subprocess.run(cmd, capture_output=True, text=True)
  if r.returncode != 0:
    raise RuntimeError(f"g++ failed for the Cython binding:\n{r.stdout}\n{r.stderr}")
  return out


if __name__ == "__main__":
  print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
  if not TAG:
    print(build_binding(force="--force" in sys.argv))
