#!/usr/bin/env python
"""bench.py — gigavoxels/s of 26-connected CCL on a 512^3 volume (BASELINE.json metric) on B200.
Headline workload = BASELINE.json configs[0]: the reference's own connectomics.npy.ckl.gz fixture (512^3 uint32).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl reference]

A "step" is one full labelling call (resolve + write) on the whole volume.
  value      device-resident input/output, CUDA events around K steps, max over ranks.
  e2e        same call through the public API on pinned HOST buffers (H2D + kernels + D2H per step).
  roofline   dominant kernel (tile labelling) against the measured HBM peak (MEASURED_PEAKS.json).
  cpu_baseline  the unmodified reference (oracle/_ref) on this box's host, 1 core, on the SAME whole volume.
`--impl reference` times the reference's own CPU implementation instead (same metric/config/volume).
--gpus 1 also reports BASELINE configs[1..3] and configs[2] (2048^3 uint64, 8 virtual slabs on one GPU) under `also`;
--gpus N > 1 first checks the sharded labelling bit for bit against the monolithic call (`parity_checked`).
Under torchrun (N > 1) the ranks label ONE volume of N 512^3 z-slabs (cc3d_b200.sharded: per-slab labelling,
NVLink face exchange, allgather of the face equivalences, global renumbering) - "weak" scaling, fixed work per GPU.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "connected-components-3d_b200"))
sys.path.insert(0, ROOT)

METRIC = "gigavoxels/s, 26-connected CCL, 512^3"
UNIT = "GVx/s"

WORKLOADS = {
  # name: generator kind, shape, call kwargs, algorithmic bytes/voxel = sizeof(in) + sizeof(out)
  "multilabel_512_u32_conn26": dict(kind="voronoi", shape=(512, 512, 512), kw=dict(connectivity=26), in_bytes=4, alg_bytes=8,
                                    desc="512^3 uint32 multilabel (synthetic Voronoi cells, ~2.9k labels), 26-connected, return_N=True: "
                                         "the configuration BASELINE.json's roofline target is quoted on (configs[0] shape)"),
  "connectomics_512_u32_conn26": dict(kind="connectomics", shape=(512, 512, 512), kw=dict(connectivity=26), in_bytes=4, alg_bytes=8,
                                      desc="configs[0]: the reference's connectomics.npy.ckl.gz (decoded fixture), 26-connected"),
  "random_binary_512_u8_conn26": dict(kind="binary", shape=(512, 512, 512), kw=dict(connectivity=26, binary_image=True), in_bytes=1, alg_bytes=5,
                                      desc="configs[1]: random 0/1 uint8 512^3 at 50% density, 26-connected, binary_image=True"),
  "random_binary_512_u8_conn26_multilabel_call": dict(kind="binary", shape=(512, 512, 512), kw=dict(connectivity=26), in_bytes=1, alg_bytes=5,
                                                      desc="configs[1]: same volume through the default (multilabel) call"),
  "random_binary_512_u8_conn6": dict(kind="binary", shape=(512, 512, 512), kw=dict(connectivity=6, binary_image=True), in_bytes=1, alg_bytes=5,
                                     desc="configs[1]: random 0/1 uint8 512^3 at 50% density, 6-connected, binary_image=True"),
  "continuous_512_f32_conn26": dict(kind="tone", shape=(512, 512, 512), kw=dict(connectivity=26, delta=10), in_bytes=4, alg_bytes=8,
                                    desc="configs[3] at 512^3: three-tone float32 + noise, delta=10, 26-connected"),
}
DEFAULT_WORKLOAD = "connectomics_512_u32_conn26"


def fixture_available():
  from oracle import decode_connectomics
  return os.path.exists(decode_connectomics.DST)


def make_volume(wl, device, rank=0, world=1):
  """One rank's volume. world > 1: rank r holds z-slab r of ONE volume of shape (world*sz, sy, sx).
  connectomics: the volume of N ranks is N copies of the fixture stacked along z (every interface joins plane 511 of
  one copy to plane 0 of the next, so the face merge has real work)."""
  import torch
  import benchdata
  sz, sy, sx = wl["shape"]
  if wl["kind"] == "binary":
    return benchdata.random_binary(wl["shape"], 0.5, 1 + rank, device)
  if wl["kind"] == "voronoi":
    return benchdata.voronoi_multilabel((sz * world, sy, sx), cell=40, seed=2, device=device, dtype=torch.int32,
                                        z_range=(rank * sz, (rank + 1) * sz))
  if wl["kind"] == "tone":
    return benchdata.three_tone_noise(wl["shape"], cell=64, seed=3 + rank, device=device)
  if wl["kind"] == "connectomics":
    from oracle import decode_connectomics
    vol = decode_connectomics.load_fixture()
    if vol is None:
      raise SystemExit("connectomics fixture missing (oracle/_ref/connectomics_512_u32.npz)")
    return torch.from_numpy(np.ascontiguousarray(vol.transpose(2, 1, 0)).view(np.int32)).to(device)
  raise ValueError(wl["kind"])


class ClockSampler:
  """Samples SM clock and throttle reasons through NVML while the timed region runs."""

  def __init__(self, index):
    self.samples, self.reasons, self.max_mhz = [], set(), None
    self._stop = threading.Event()
    self._thr = None
    try:
      import pynvml
      pynvml.nvmlInit()
      self.nv = pynvml
      self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
      self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
    except Exception:
      self.nv = None

  def _run(self):
    nv = self.nv
    names = {
      getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8): "hw_slowdown",
      getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
      getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
      getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4): "sw_power_cap",
    }
    while not self._stop.is_set():
      try:
        self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
        r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
        for bit, name in names.items():
          if r & bit:
            self.reasons.add(name)
      except Exception:
        pass
      time.sleep(0.002)

  def __enter__(self):
    if self.nv is not None:
      self._thr = threading.Thread(target=self._run, daemon=True)
      self._thr.start()
    return self

  def __exit__(self, *a):
    self._stop.set()
    if self._thr is not None:
      self._thr.join()

  def summary(self):
    if not self.samples:
      return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
    return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
            "reasons": sorted(self.reasons), "samples": len(self.samples)}


def hbm_peak():
  p = os.path.join(ROOT, "MEASURED_PEAKS.json")
  if os.path.exists(p):
    try:
      return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
      pass
  return 6650.0, "fallback (B200_PROFILING.md)"


def reference_labeller():
  """(callable, kind): the unmodified reference when oracle/_ref travelled, else the C oracle port."""
  from oracle import oracle
  ref = oracle.reference_module()
  if ref is not None:
    return ref.connected_components, "reference"
  return oracle.connected_components, "port"


def cpu_volume(wl):
  """The WHOLE single-GPU workload volume as a host array for the CPU arms (same generator, same seed as rank 0)."""
  import torch
  dev = "cuda" if torch.cuda.is_available() else "cpu"
  return np.ascontiguousarray(make_volume(wl, dev).cpu().numpy())


def time_cpu(fn, x, kw, repeats):
  best = float("inf")
  for _ in range(repeats):
    t0 = time.perf_counter()
    fn(x, return_N=True, **kw)
    best = min(best, time.perf_counter() - t0)
  return best


def run_reference_arm(args, wl):
  """The reference's own CPU implementation (oracle/_ref, unmodified cc3d) on the SAME full volume the GPU arm labels at
  N=1 (under torchrun only rank 0 works; cc3d is single threaded, so one slab of the N-slab volume is the bounded
  sample of the N > 1 workload and is named as such)."""
  rank = int(os.environ.get("RANK", "0"))
  if rank != 0:
    return
  fn, kind = reference_labeller()
  x = cpu_volume(wl)
  steps, warmup = args.steps, args.warmup
  # bounded: a whole-volume step is 0.6 s (multilabel) to 5 s (random binary); keep the arm within a few minutes
  t0 = time.perf_counter()
  fn(x, return_N=True, **wl["kw"])
  t_one = time.perf_counter() - t0
  budget = 150.0
  if (steps + warmup) * t_one > budget:
    warmup = min(warmup, 1)
    steps = max(1, min(steps, int(budget / t_one) - warmup))
  for _ in range(max(0, warmup - 1)):
    fn(x, return_N=True, **wl["kw"])
  t0 = time.perf_counter()
  for _ in range(steps):
    out, N = fn(x, return_N=True, **wl["kw"])
  dt = time.perf_counter() - t0
  value = x.size * steps / dt / 1e9
  sample = (f"the whole {'x'.join(str(v) for v in x.shape)} {args.workload} volume ({x.size} voxels) per step, "
            f"{steps} timed steps" + (f" (requested {args.steps}; bounded to ~{int(budget)} s of CPU work)" if steps != args.steps else "")
            + (f"; N={args.gpus} workload = {args.gpus} such slabs, cc3d is single threaded" if args.gpus > 1 else ""))
  line = {
    "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
    "warmup": warmup, "ms_per_step": dt / steps * 1e3, "higher_is_better": True, "scaling": "weak",
    "vs_baseline": None, "dtype": "u8" if wl["in_bytes"] == 1 else ("f32" if wl["kind"] == "tone" else "u32"),
    "data": "fixture" if wl["kind"] == "connectomics" else "synthetic",
    "config": {"workload": args.workload, "description": wl["desc"], "sample": sample, "N": int(N),
               "same_volume_as_gpu_arm": True},
    "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": kind, "sample": sample,
                     "host_cores": os.cpu_count()},
    "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    "gpu_launches": 0,
  }
  print(json.dumps(line), flush=True)


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument("--gpus", type=int, default=1)
  ap.add_argument("--steps", type=int, default=30)
  ap.add_argument("--warmup", type=int, default=5)
  ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
  ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
  ap.add_argument("--no-cpu-baseline", action="store_true")
  ap.add_argument("--no-extra", action="store_true", help="skip the additional workloads reported under 'also'")
  ap.add_argument("--no-big", action="store_true", help="skip the 1-GPU configs[2] leg (2048^3 uint64, 96 GiB)")
  args = ap.parse_args()
  args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
  if WORKLOADS[args.workload]["kind"] == "connectomics" and not fixture_available():
    print("bench.py: oracle/_ref/connectomics_512_u32.npz is missing (python oracle/decode_connectomics.py builds it "
          "where /root/reference exists); falling back to the synthetic Voronoi volume of the same shape", file=sys.stderr)
    args.workload = "multilabel_512_u32_conn26"
  wl = WORKLOADS[args.workload]

  if args.impl == "reference":
    run_reference_arm(args, wl)
    return

  import torch
  import torch.distributed as dist
  import cc3d_b200
  from cc3d_b200 import _lib

  rank = int(os.environ.get("RANK", "0"))
  world = int(os.environ.get("WORLD_SIZE", "1"))
  local_rank = int(os.environ.get("LOCAL_RANK", "0"))
  if not torch.cuda.is_available():
    raise SystemExit("bench.py needs a CUDA device (cc3d_b200 has no CPU fallback)")
  torch.cuda.set_device(local_rank)
  dev = torch.device("cuda", local_rank)
  if world > 1:
    dist.init_process_group("nccl", device_id=dev)

  def barrier():
    if world > 1:
      dist.barrier()
    torch.cuda.synchronize()

  def max_over_ranks(v):
    if world == 1:
      return v
    t = torch.tensor([v], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())

  L = _lib.lib()
  x = make_volume(wl, dev, rank, world)
  voxels = x.numel()
  kw = wl["kw"]

  # N = 1: the plain call. N > 1: ONE volume of N z-slabs, one per rank, labelled globally (face exchange +
  # allgather of the face equivalences); the result is identical to the single-GPU labelling of the whole volume.
  if world > 1:
    from cc3d_b200 import sharded
    def label(vol, **kwargs):
      return sharded.connected_components_slab(vol, return_N=True, **kwargs)
  else:
    def label(vol, **kwargs):
      return cc3d_b200.connected_components(vol, return_N=True, **kwargs)

  def timed_device_loop(vol, kwargs, steps, warmup):
    for _ in range(warmup):
      out, N = label(vol, **kwargs)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
      out, N = label(vol, **kwargs)
    e1.record()
    barrier()
    return max_over_ranks(e0.elapsed_time(e1)), N, out

  # ---- N > 1: the sharded labelling must equal the monolithic labelling of the whole volume, bit for bit ----
  parity = None
  if world > 1:
    import hashlib
    out_s, N_s = label(x, **kw)
    full_in = torch.empty((world,) + tuple(x.shape), dtype=x.dtype, device=dev)
    dist.all_gather_into_tensor(full_in, x)
    signed = {2: torch.int16, 4: torch.int32, 8: torch.int64}[out_s.element_size()]
    out_s = out_s.contiguous().view(signed)                  # NCCL has no unsigned 16/32/64-bit types in every build
    full_out = torch.empty((world,) + tuple(out_s.shape), dtype=signed, device=dev)
    dist.all_gather_into_tensor(full_out, out_s)
    torch.cuda.synchronize()
    if rank == 0:
      whole = full_in.reshape((-1,) + tuple(x.shape[1:]))
      mono, N_m = cc3d_b200.connected_components(whole, return_N=True, **kw)
      same = bool(N_m == N_s and mono.element_size() == full_out.element_size()
                  and torch.equal(mono.reshape(-1).view(signed), full_out.reshape(-1)))
      sha = hashlib.sha256(mono.reshape(-1).cpu().numpy().tobytes()).hexdigest()
      parity = {"parity_checked": same, "N_sharded": int(N_s), "N_monolithic": int(N_m), "labels_sha256": sha,
                "how": f"rank 0 labelled the whole {tuple(whole.shape)} volume with the single-GPU call and compared it bit for bit "
                       "with the all-gathered sharded result before the timed region"}
      if not same:
        print("bench.py: SHARDED RESULT DIFFERS FROM THE MONOLITHIC LABELLING", file=sys.stderr)
      del mono, whole
    del full_in, full_out, out_s
    L.cc3d_b200_release_workspace()
    torch.cuda.empty_cache()
    barrier()

  # ---- value: device-resident ----
  launches0 = L.cc3d_b200_launch_count()
  with ClockSampler(local_rank) as clk:
    ms, N, out = timed_device_loop(x, kw, args.steps, args.warmup)
  # warm-up launches are excluded: count again over a clean run of `steps`
  per_step_launches = (L.cc3d_b200_launch_count() - launches0) // (args.steps + args.warmup)
  value = world * voxels * args.steps / (ms / 1e3) / 1e9
  out_dtype = str(out.dtype).replace("torch.", "")
  out_bytes = out.element_size()
  del out

  # ---- roofline of the dominant kernel: CUDA events recorded by the library on the launching stream ----
  # (N > 1: the single-synchronisation slab path records no per-kernel events; the stepwise path runs the same kernels)
  cc3d_b200.set_timing(True)
  if world > 1:
    os.environ["CC3D_SHARDED_STEPWISE"] = "1"
  ktimes = {}
  for _ in range(args.steps):
    label(x, **kw)
    for name, t in cc3d_b200.last_timings():
      ktimes.setdefault(name, []).append(t)
  os.environ.pop("CC3D_SHARDED_STEPWISE", None)
  cc3d_b200.set_timing(False)
  kavg = {k: float(np.mean(v)) for k, v in ktimes.items()}
  dom = max(kavg, key=kavg.get)
  peak, peak_src = hbm_peak()
  alg_bytes = wl["alg_bytes"] * voxels
  achieved = alg_bytes / (kavg[dom] / 1e3) / 1e9
  traffic = None
  tpath = os.path.join(ROOT, "profiles", "traffic.json")
  if os.path.exists(tpath):
    try:
      traffic = json.load(open(tpath)).get(args.workload, {}).get(dom)
    except Exception:
      traffic = None
  # each dense kernel against its OWN compulsory traffic: A reads the input once, D writes the output once
  own = {}
  for kname, nbytes in (("A_faces", voxels * x.element_size()), ("D_expand", voxels * out_bytes)):
    if kname in kavg:
      gbs = nbytes / (kavg[kname] / 1e3) / 1e9
      own[kname] = {"bytes": nbytes, "ms": kavg[kname], "GB/s": gbs, "frac": gbs / peak}
  # the same kernel against the DRAM bytes it really moves (ncu dram__bytes_read + write per launch, profiles/traffic.json)
  dram_frac = (traffic / (kavg[dom] / 1e3) / 1e9 / peak) if traffic else None
  all_traffic = {}
  try:
    all_traffic = {k: v for k, v in json.load(open(tpath)).get(args.workload, {}).items() if k in kavg}
  except Exception:
    pass
  roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
              "dram_frac": dram_frac,
              "per_kernel_dram_frac": {k: v / (kavg[k] / 1e3) / 1e9 / peak for k, v in all_traffic.items()},
              "per_kernel_own_bytes": own,
              "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes,
              "note": "achieved = (sizeof(in)+sizeof(out)) * voxels / duration of the dominant kernel (the prescribed formula; it credits that "
                      "kernel with bytes it does not move); dram_frac = that kernel's measured DRAM traffic / its duration / peak; "
                      "pipeline_frac = compulsory bytes / whole step / peak (the honest end-to-end figure)",
              "kernel_ms": kavg, "kernel_share": {k: v / sum(kavg.values()) for k, v in kavg.items()},
              "pipeline_frac": alg_bytes / (ms / args.steps / 1e3) / 1e9 / peak}

  # ---- e2e: public API on pinned host buffers (H2D + kernels + D2H inside the timed region) ----
  # NUMA: the pinned staging buffers should live on the memory node of this rank's GPU (r01: with every rank on node 0
  # the 8-GPU e2e collapsed). Bind this process to the CPUs of the GPU's node before allocating them, if the cpuset
  # allows it; report what was possible.
  numa = {"gpu_node": None, "bound": False, "affinity_before": len(os.sched_getaffinity(0))}
  try:
    props = torch.cuda.get_device_properties(local_rank)
    bus = f"{props.pci_domain_id:04x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0"
    node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read().strip())
    numa["gpu_node"] = node
    if node >= 0:
      cpus = set()
      for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
        a, _, b = part.partition("-")
        cpus.update(range(int(a), int(b or a) + 1))
      allowed = cpus & os.sched_getaffinity(0)
      if allowed:
        os.sched_setaffinity(0, allowed)
        numa["bound"] = True
      numa["node_cpus_allowed"] = len(allowed)
  except Exception as e:   # noqa: BLE001 - informational only
    numa["error"] = repr(e)[:120]
  xh = torch.empty(x.shape, dtype=x.dtype, pin_memory=True)
  xh.copy_(x)
  np_out_dtype = {"uint16": np.uint16, "uint32": np.uint32, "uint64": np.uint64}[out_dtype]
  oh = torch.empty((voxels * out_bytes,), dtype=torch.uint8, pin_memory=True)
  e2e_steps = max(3, min(args.steps, 10))
  if world == 1:
    x_np = xh.numpy()
    out_np = oh.numpy().view(np_out_dtype)
    def e2e_step():
      return cc3d_b200.connected_components(x_np, return_N=True, out=out_np, **kw)[1]
  else:
    def e2e_step():
      xd = xh.to(dev, non_blocking=True)
      o, n = label(xd, **kw)
      oh.copy_(o.reshape(-1).view(torch.uint8), non_blocking=True)
      torch.cuda.current_stream().synchronize()
      return n
  for _ in range(2):
    e2e_step()
  barrier()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(e2e_steps):
    N2 = e2e_step()
  e1.record()
  barrier()
  e2e_ms = max_over_ranks(e0.elapsed_time(e1))
  assert N2 == N
  # the two copies alone, every rank at the same time (what the PCIe / host-memory path of this box gives each GPU)
  def copy_gbs(fn, nbytes):
    fn(); barrier()
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    c0.record()
    for _ in range(3):
      fn()
    c1.record(); barrier()
    return nbytes * 3 / (c0.elapsed_time(c1) / 1e3) / 1e9
  xd_tmp = torch.empty_like(x)
  od_tmp = torch.empty((voxels * out_bytes,), dtype=torch.uint8, device=dev)
  h2d_gbs = copy_gbs(lambda: xd_tmp.copy_(xh, non_blocking=True), voxels * x.element_size())
  d2h_gbs = copy_gbs(lambda: oh.copy_(od_tmp, non_blocking=True), voxels * out_bytes)
  del xd_tmp, od_tmp
  if world > 1:
    tt = torch.tensor([h2d_gbs, d2h_gbs], dtype=torch.float64, device=dev)
    allt = [torch.empty_like(tt) for _ in range(world)]
    dist.all_gather(allt, tt)
    per_rank = [[round(float(v), 1) for v in t_.tolist()] for t_ in allt]
  else:
    per_rank = [[round(h2d_gbs, 1), round(d2h_gbs, 1)]]
  e2e = {"value": world * voxels * e2e_steps / (e2e_ms / 1e3) / 1e9, "unit": UNIT,
         "h2d_bytes_per_step": voxels * x.element_size(), "d2h_bytes_per_step": voxels * out_bytes + 32,
         "ms_per_step": e2e_ms / e2e_steps, "steps": e2e_steps, "host_memory": "pinned",
         "copy_GBps_per_rank_h2d_d2h": per_rank, "numa": numa,
         "note": "e2e is bounded by the two PCIe copies (h2d + d2h bytes / the per-rank copy rates above); "
                 "copy rates are measured with every rank copying at the same time"}

  # ---- other BASELINE workloads, same timing method (N=1 only) ----
  also = []
  if world == 1 and not args.no_extra:
    for name, w in WORKLOADS.items():
      if name == args.workload:
        continue
      try:
        v = make_volume(w, dev)
      except SystemExit:
        continue
      m, n_, o_ = timed_device_loop(v, w["kw"], max(5, args.steps // 3), 3)
      steps_ = max(5, args.steps // 3)
      also.append({"workload": name, "value": v.numel() * steps_ / (m / 1e3) / 1e9, "unit": UNIT,
                   "ms_per_step": m / steps_, "N": int(n_),
                   "compulsory_roofline_frac": w["alg_bytes"] * v.numel() / (m / steps_ / 1e3) / 1e9 / peak})
      del v, o_

  # ---- rows a13 / a14 on the headline volume: statistics of its labelling, dust(threshold=100) of the volume ----
  if world == 1 and not args.no_extra:
    lab_t, n_lab = cc3d_b200.connected_components(x, return_N=True, **kw)
    def timed_call(fn, steps_=10):
      for _ in range(3):
        fn()
      torch.cuda.synchronize()
      a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
      a0.record()
      for _ in range(steps_):
        fn()
      a1.record()
      torch.cuda.synchronize()
      return a0.elapsed_time(a1) / steps_
    m_st = timed_call(lambda: cc3d_b200.statistics(lab_t, no_slice_conversion=True))
    also.append({"workload": "statistics_of_" + args.workload, "value": voxels / (m_st / 1e3) / 1e9, "unit": UNIT, "ms_per_step": m_st,
                 "N": int(n_lab), "compulsory_roofline_frac": out_bytes * voxels / (m_st / 1e3) / 1e9 / peak,
                 "note": "public call on the CUDA label tensor: label max + statistics kernel + D2H of the per-label arrays + host finalisation"})
    m_du = timed_call(lambda: cc3d_b200.dust(x, threshold=100, connectivity=kw.get("connectivity", 26)))
    also.append({"workload": "dust100_of_" + args.workload, "value": voxels / (m_du / 1e3) / 1e9, "unit": UNIT, "ms_per_step": m_du,
                 "compulsory_roofline_frac": 3 * x.element_size() * voxels / (m_du / 1e3) / 1e9 / peak,
                 "note": "fused dust: image read twice and written once, no label volume"})
    # ---- SURVEY 8(f)4 on the same labelling: run table through the C-ABI on device buffers (R1 + scan + R2) ----
    try:
      import ctypes
      from cc3d_b200 import _lib as _cl
      Lc = _cl.lib()
      flat = lab_t.reshape(-1)
      kind_l = {1: _cl.U8, 2: _cl.U16, 4: _cl.U32, 8: _cl.U64}[flat.element_size()]
      st_ = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
      cnt_ = ctypes.c_uint64(0)
      _cl.check(Lc.cc3d_b200_runs(flat.data_ptr(), kind_l, flat.numel(), None, None, None, 0, ctypes.byref(cnt_), _cl.DEVICE, st_))
      k_runs = int(cnt_.value)
      tab = torch.empty((3, max(k_runs, 1)), dtype=torch.int64, device=flat.device)
      m_rn = timed_call(lambda: _cl.check(Lc.cc3d_b200_runs(flat.data_ptr(), kind_l, flat.numel(), tab[0].data_ptr(), tab[1].data_ptr(),
                                                             tab[2].data_ptr(), k_runs, ctypes.byref(cnt_), _cl.DEVICE, st_)))
      also.append({"workload": "runs_of_" + args.workload, "value": voxels / (m_rn / 1e3) / 1e9, "unit": UNIT, "ms_per_step": m_rn,
                   "runs": k_runs, "compulsory_roofline_frac": (flat.element_size() * voxels + 24 * k_runs) / (m_rn / 1e3) / 1e9 / peak,
                   "note": "cc3d_b200_runs on the CUDA label tensor: count sweep + scan + emit sweep (labels read twice, 24 B per run written), one 8-byte read back"})
      del tab
    except Exception as e:   # the extra line must never cost the headline
      also.append({"workload": "runs_of_" + args.workload, "error": repr(e)[:200]})
    del lab_t

  # ---- configs[2]: ONE 2048^3 uint64 Voronoi volume, z-slabs over the ranks (strong scaling; needs slabs < 2^32 voxels) ----
  if world >= 4 and not args.no_extra:
    import benchdata
    n3 = 2048
    szr = n3 // world
    slab = benchdata.voronoi_multilabel((n3, n3, n3), cell=160, seed=2, device=dev, dtype=torch.int64, id_bits=62,
                                        z_range=(rank * szr, (rank + 1) * szr))
    m, n_, o_ = timed_device_loop(slab, dict(connectivity=26), 5, 2)
    also.append({"workload": "voronoi_2048_u64_conn26_sharded", "description": "configs[2]: 2048^3 uint64 Voronoi (~2.9k labels), 26-connected, "
                 f"z-slabs of {szr} planes over {world} GPUs, face exchange + allgather (strong scaling; 1 GPU with 8 virtual slabs: see DESIGN.md)",
                 "value": float(n3) ** 3 * 5 / (m / 1e3) / 1e9, "unit": UNIT, "ms_per_step": m / 5, "N": int(n_),
                 "compulsory_roofline_frac": 12 * float(n3) ** 3 / (m / 5 / 1e3) / 1e9 / (peak * world)})
    del slab, o_

  # ---- CPU baseline (rank 0, N=1) ----
  cpu = None
  if rank == 0 and world == 1 and not args.no_cpu_baseline:
    fn, kind = reference_labeller()
    xs = np.ascontiguousarray(x.cpu().numpy())
    t = time_cpu(fn, xs, kw, repeats=3)
    cpu = {"value": xs.size / t / 1e9, "unit": UNIT, "cores": 1, "kind": kind, "host_cores": os.cpu_count(),
           "sample": f"the whole {'x'.join(str(v) for v in xs.shape)} volume the GPU arm labels ({xs.size} voxels), best of 3"}

  # ---- configs[2] on ONE GPU: the 2048^3 uint64 Voronoi volume as 8 virtual z-slabs (64 GiB in, 32 GiB out) - the
  #      1-GPU denominator of the 1 -> 8 GPU scaling figure, measured by the same driver run ----
  if world == 1 and not args.no_extra and not args.no_big:
    try:
      import benchdata
      from cc3d_b200 import sharded
      del x, xh, oh
      if "x_np" in dir():
        del x_np, out_np
      L.cc3d_b200_release_workspace()
      torch.cuda.empty_cache()
      free_b, _tot = torch.cuda.mem_get_info()
      n3, ns = 2048, 8
      if free_b < 150 * 2**30:
        also.append({"workload": "voronoi_2048_u64_conn26_1gpu", "skipped": f"only {free_b / 2**30:.0f} GiB free"})
      else:
        szr = n3 // ns
        slabs = [benchdata.voronoi_multilabel((n3, n3, n3), cell=160, seed=2, device=dev, dtype=torch.int64, id_bits=62,
                                              z_range=(r * szr, (r + 1) * szr)) for r in range(ns)]
        outs, n_ = sharded.connected_components_slabs(slabs, connectivity=26, return_N=True)
        del outs
        torch.cuda.synchronize()
        b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        b0.record()
        for _ in range(3):
          outs, n_ = sharded.connected_components_slabs(slabs, connectivity=26, return_N=True)
          del outs
        b1.record()
        torch.cuda.synchronize()
        m = b0.elapsed_time(b1) / 3
        also.append({"workload": "voronoi_2048_u64_conn26_1gpu", "description": "configs[2]: 2048^3 uint64 Voronoi, 26-connected, ONE GPU, "
                     f"{ns} virtual z-slabs of {szr} planes through cc3d_b200.sharded.connected_components_slabs (a single call is limited "
                     "to < 2^32-1 voxels); the N >= 4 runs report the same volume sharded over the GPUs",
                     "value": float(n3) ** 3 / (m / 1e3) / 1e9, "unit": UNIT, "ms_per_step": m, "N": int(n_),
                     "compulsory_roofline_frac": 12 * float(n3) ** 3 / (m / 1e3) / 1e9 / peak})
        del slabs
      L.cc3d_b200_release_workspace()
      torch.cuda.empty_cache()
    except Exception as e:   # never lose the headline line to the 96 GiB leg
      also.append({"workload": "voronoi_2048_u64_conn26_1gpu", "error": repr(e)[:300]})

  if rank == 0:
    line = {
      "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
      "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
      "dtype": "u8" if wl["in_bytes"] == 1 else ("f32" if wl["kind"] == "tone" else "u32"),
      "data": "fixture (the reference's benchmarks/connectomics.npy.ckl.gz, decoded)" if wl["kind"] == "connectomics" else "synthetic",
      "config": {"workload": args.workload, "description": wl["desc"], "out_dtype": out_dtype, "N": int(N),
                 "l2": "input + output of one step (>= 640 MB) are larger than the 126 MB L2",
                 "parallelism": (f"one ({world}*512)x512x512 volume, one 512^3 z-slab per GPU (NVLink plane exchange + one all-gather of the face equivalences)"
                                 if world > 1 else "single GPU")},
      "clocks": clk.summary(), "e2e": e2e, "gpu_launches": int(per_step_launches) * args.steps,
      "gpu_launches_per_step": int(per_step_launches),
      "roofline": roofline, "cpu_baseline": cpu, "also": also,
    }
    if parity is not None:
      line.update(parity)
    print(json.dumps(line), flush=True)
  if world > 1:
    dist.destroy_process_group()


if __name__ == "__main__":
  main()
