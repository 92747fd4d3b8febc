#!/usr/bin/env python
"""bench.py -- matvec matrix-elements/s of the symmetry-adapted Hamiltonian on B200 (+ basis-build
time and time-to-ground-state), with the CPU restatement timed beside it.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config DECK]

A step is one application y = H x of the deck's Hamiltonian to one Krylov vector (block size 1):
at N > 1 the step is what the solver does per matvec (`sped_operator_matvec_sharded`) -- NCCL
all-gather of the row-sharded vector, overlapped with the pass over the elements whose sources the
rank owns, then the pass over the remote-source elements.  `value` = (N_rows + E_offdiag) / t with inputs
resident in HBM; `e2e` = the same through the reference-facing `ls_operator_matmat` with HOST
buffers (H2D of x and D2H of y inside the timed region).  One JSON line on stdout (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

DEFAULT_DECK = "heisenberg_square_6x6"
METRIC = "matvec_matrix_elements_per_s"
UNIT = "matrix-elements/s"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def algorithmic_bytes(n_rows, n_off, elem_size, symmetric):
    """B = N (8 rep + 8 diag + T y + 8 norm if symmetric) + E T."""
    return n_rows * (8 + 8 + elem_size + (8 if symmetric else 0)) + n_off * elem_size


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region: an NVML polling thread (about
    one sample per millisecond -- the timed region of a 2 ms matvec is far shorter than nvidia-smi's
    100 ms period); `nvidia-smi -lms` is the fallback when NVML cannot be loaded."""
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index=0):
        self.proc = None
        self.device_index = device_index
        self.thread = None
        self.samples = []
        self.stop_flag = False
        self.nvml = None

    def _physical_index(self):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            try:
                return int(vis.split(",")[self.device_index])
            except Exception:
                pass
        return self.device_index

    def _poll(self, handle):
        nv = self.nvml
        while not self.stop_flag:
            try:
                sm = nv.nvmlDeviceGetClockInfo(handle, nv.NVML_CLOCK_SM)
                reasons = nv.nvmlDeviceGetCurrentClocksEventReasons(handle)
                self.samples.append((sm, reasons))
            except Exception:
                break
            time.sleep(0.0005)

    def start(self):
        try:
            import threading

            import pynvml as nv

            nv.nvmlInit()
            handle = nv.nvmlDeviceGetHandleByIndex(self._physical_index())
            self.max_sm = nv.nvmlDeviceGetMaxClockInfo(handle, nv.NVML_CLOCK_SM)
            self.nvml = nv
            self.thread = threading.Thread(target=self._poll, args=(handle,), daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
            self.thread = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self._physical_index())], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.thread is not None:
            self.stop_flag = True
            self.thread.join(timeout=2)
            nv = self.nvml
            bits = {"hw_slowdown": nv.nvmlClocksEventReasonHwSlowdown,
                    "hw_thermal_slowdown": nv.nvmlClocksEventReasonHwThermalSlowdown,
                    "sw_thermal_slowdown": nv.nvmlClocksEventReasonSwThermalSlowdown,
                    "sw_power_cap": nv.nvmlClocksEventReasonSwPowerCap}
            sm = [float(v) for v, _ in self.samples]
            reasons = sorted(name for name, bit in bits.items() if any(r & bit for _, r in self.samples))
            return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": float(self.max_sm),
                    "samples": len(sm), "reasons": reasons, "source": "nvml"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                mx.append(float(parts[2]))
            except ValueError:
                continue
            for name, val in zip(names, parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons), "source": "nvidia-smi"}


def cpu_sample_rate(O, oop, n, seconds, dtype=np.float64):
    """Time the oracle matvec on a strided row sample sized for about `seconds` of CPU work."""
    from helpers import splitmix_vector

    x = splitmix_vector(n, 0x5EED0001, np.float64).astype(dtype)  # real-valued, like the device generator
    x /= np.linalg.norm(x)
    y = np.zeros_like(x)
    probe = min(n, 512 * O.num_threads())
    stride = max(1, n // probe)
    t0 = time.perf_counter()
    e = oop.matmat_rows(x, y, 0, n, stride)
    dt = time.perf_counter() - t0
    rows = len(range(0, n, stride))
    rate_rows = rows / max(dt, 1e-9)
    target = int(max(rows, min(n, rate_rows * seconds)))
    stride = max(1, n // target)
    return x, y, stride


def run_reference(args):
    """--impl reference: the CPU restatement (oracle port) of the same path on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from helpers import oracle_problem
    from oracle import oracle as O
    from spin_ed_b200 import decks

    O.build()
    cfg = decks.load(args.config)
    ob, terms = oracle_problem(O, cfg)
    cache = os.path.join("/tmp", f"sped_oracle_reps_{args.config}.npy")
    t0 = time.perf_counter()
    if os.path.exists(cache):
        ob.build(np.load(cache))
    else:
        ob.build()
        try:
            np.save(cache, ob.states)
        except Exception:
            pass
    build_s = time.perf_counter() - t0
    oop = O.Operator(ob, terms)
    n = ob.number_states
    dtype = np.float64 if oop.is_real else np.complex128
    x, y, stride = cpu_sample_rate(O, oop, n, args.cpu_seconds / 3.0, dtype)
    rows = len(range(0, n, stride))
    times, elems = [], 0
    for it in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        e = oop.matmat_rows(x, y, 0, n, stride)
        dt = time.perf_counter() - t0
        if it >= args.warmup:
            times.append(dt)
            elems = rows + e
    t = sum(times) / len(times)
    value = elems / t
    sample = f"{rows} of {n} rows (every {stride}-th), {elems} matrix elements per step, float64, OpenMP"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": args.config, "rows": n, "note": "CPU restatement of the reference path (oracle port), not the "
                   "upstream binary: lattice-symmetries/PRIMME are absent and cannot be built offline"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": O.num_threads(), "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "extra": {"basis_build_s_cpu": build_s},
    }
    print(json.dumps(line), flush=True)
    return 0


def run_ours(args):
    import torch
    import torch.distributed as dist

    from helpers import splitmix_vector
    from spin_ed_b200 import config as sconfig
    from spin_ed_b200 import decks, ffi

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path")
    torch.cuda.set_device(local)
    ffi.setDevice(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
        box = [ffi.commUniqueId() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        ffi.commInit(world, rank, box[0])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    cfg = decks.load(args.config)
    spec = sconfig.parseConfig(cfg)
    uc = sconfig.toConfig(spec)
    basis, op = uc.cBasis, uc.cHamiltonian.operatorObject
    barrier()
    t0 = time.perf_counter()
    ffi.buildBasis(basis)
    barrier()
    build_wall = time.perf_counter() - t0
    n = ffi.getNumberStates(basis)
    is_real = ffi.isOperatorReal(op)
    np_dtype = np.float64 if is_real else np.complex128
    t_dtype = torch.float64 if is_real else torch.complex128
    tag = ffi.DTYPE_TAGS[np.dtype(np_dtype)]
    es = np.dtype(np_dtype).itemsize
    rows = n
    rd = ffi.basisRowDistribution(basis)  # block-cyclic rows of this rank, [rank][local] vector layout
    n_local, chunk = int(rd.n_local), int(rd.chunk)
    symmetric = ffi.basisProgramStats(basis)["steps"] > 1 or cfg["basis"].get("spin_inversion") is not None
    log(f"[rank {rank}] {args.config}: N={n} local rows {n_local} (blocks of {1 << rd.log2_block}) build {build_wall:.3f}s")

    # device-resident inputs: the replicated vector (padded to world * chunk) and the local output
    # x[i] = uniform(-1,1) from splitmix64(seed ^ global row), generated on the device
    # shard by shard so that 40-spin vectors never exist on the host
    def device_splitmix(lo, hi, seed):
        """values of the local rows lo..hi-1 (local indices), keyed by their GLOBAL row"""
        def s64(v):
            return v - (1 << 64) if v >= (1 << 63) else v

        out = torch.empty(hi - lo, dtype=torch.float64, device=dev)
        lb, bmask = int(rd.log2_block), (1 << int(rd.log2_block)) - 1
        for c0 in range(lo, hi, 1 << 24):
            c1 = min(hi, c0 + (1 << 24))
            i = torch.arange(c0, c1, dtype=torch.int64, device=dev)
            g = i if world == 1 else ((((i >> lb) * world + rank) << lb) + (i & bmask))
            z = (g ^ s64(seed)) + s64(0x9E3779B97F4A7C15)
            z = (z ^ ((z >> 30) & ((1 << 34) - 1))) * s64(0xBF58476D1CE4E5B9)
            z = (z ^ ((z >> 27) & ((1 << 37) - 1))) * s64(0x94D049BB133111EB)
            z = z ^ ((z >> 31) & ((1 << 33) - 1))
            out[c0 - lo:c1 - lo] = ((z >> 11) & ((1 << 53) - 1)).to(torch.float64) / 9007199254740992.0 * 2.0 - 1.0
        return out

    xshard = torch.zeros(chunk, dtype=t_dtype, device=dev)
    if n_local:
        xshard[:n_local] = device_splitmix(0, n_local, 0x5EED0001).to(t_dtype)
    nrm2 = (xshard.abs() ** 2).sum().to(torch.float64).reshape(1)
    if world > 1:
        dist.all_reduce(nrm2)
    xshard /= float(nrm2.sqrt().item())
    xfull = torch.zeros(chunk * world, dtype=t_dtype, device=dev)
    if world > 1:
        dist.all_gather_into_tensor(xfull, xshard)
    else:
        xfull.copy_(xshard)
    if n <= 4096 and world == 1:  # the generator must reproduce the host definition used by the tests
        ref = splitmix_vector(n, 0x5EED0001, np.float64)
        assert np.allclose(xfull[:n].cpu().numpy().real * float(nrm2.sqrt().item()), ref, rtol=0, atol=1e-15)
    ylocal = torch.zeros(max(n_local, 1), dtype=t_dtype, device=dev)
    stream = torch.cuda.current_stream()

    def step():
        if world > 1:  # what sped_eigh does per matvec: all-gather of the shards overlapped with the local-source pass
            ffi.operatorMatvecSharded(op, tag, xshard.data_ptr(), ylocal.data_ptr(), xfull.data_ptr(), stream.cuda_stream)
        else:
            ffi.operatorMatmatDevice(op, tag, 1, xfull.data_ptr(), chunk * world, ylocal.data_ptr(), max(n_local, 1),
                                     stream.cuda_stream)

    # (a) matrix-free kernel alone (what the first application of an operator costs, and the only
    #     mode when the element cache does not fit): a few steps, device events
    ffi.operatorSetCache(op, 0)
    step()
    barrier()
    mf_steps = max(2, min(args.steps, 3))
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(mf_steps):
        step()
    b.record()
    barrier()
    mf_t = torch.tensor([a.elapsed_time(b) / mf_steps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(mf_t, op=dist.ReduceOp.MAX)
    matrix_free_ms = float(mf_t.item())
    # (b) the default path: elements cached in HBM by the first application (if they fit)
    ffi.operatorSetCache(op, -1)
    for _ in range(args.warmup):
        step()
    barrier()
    cache_info = ffi.operatorCacheInfo(op)
    rows, n_off = ffi.operatorCountElements(op)  # E: off-diagonal elements one application touches
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = ffi.kernelLaunches()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    barrier()
    ev[0].record()
    for i in range(args.steps):
        step()
        ev[i + 1].record()
    barrier()
    launches = ffi.kernelLaunches() - launches0
    clocks = sampler.stop() if rank == 0 else None
    total_ms = ev[0].elapsed_time(ev[-1])
    t_local = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_local, op=dist.ReduceOp.MAX)
    total_ms = float(t_local.item())
    ms_per_step = total_ms / args.steps
    value = (rows + n_off) / (ms_per_step * 1e-3)
    # dominant kernel: the matvec kernel is the only kernel of ours in a step
    kern_ms = ms_per_step if world == 1 else None
    if world > 1:
        # time the kernel alone (no all-gather) for the per-GPU roofline
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(args.steps):
            ffi.operatorMatmatDevice(op, tag, 1, xfull.data_ptr(), chunk * world, ylocal.data_ptr(), max(n_local, 1),
                                     stream.cuda_stream)
        b.record()
        barrier()
        kern_ms = a.elapsed_time(b) / args.steps
    kern_all = [kern_ms]
    if world > 1:
        kern_all = [None] * world
        dist.all_gather_object(kern_all, kern_ms)
        kern_ms = max(kern_all)  # the slowest rank bounds the step
    peak, peak_src = measured_peak()
    local_off = n_off * n_local / max(n, 1)
    alg_bytes = algorithmic_bytes(n_local, local_off, es, symmetric)
    achieved = alg_bytes / (kern_ms * 1e-3) / 1e9

    # the same kernel with single-precision storage (what `datatype: float32` decks such as 6x6 ask
    # for; accumulation stays f64): kernel only, reported beside the f64 headline
    f32_leg = None
    if is_real:
        x32 = xfull.to(torch.float32)
        y32 = torch.zeros(max(n_local, 1), dtype=torch.float32, device=dev)
        tag32 = ffi.DTYPE_TAGS[np.dtype(np.float32)]
        for _ in range(3):
            ffi.operatorMatmatDevice(op, tag32, 1, x32.data_ptr(), chunk * world, y32.data_ptr(), max(n_local, 1),
                                     stream.cuda_stream)
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(args.steps):
            ffi.operatorMatmatDevice(op, tag32, 1, x32.data_ptr(), chunk * world, y32.data_ptr(), max(n_local, 1),
                                     stream.cuda_stream)
        b.record()
        barrier()
        t32 = torch.tensor([a.elapsed_time(b) / args.steps], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t32, op=dist.ReduceOp.MAX)
        ms32 = float(t32.item())
        bytes32 = algorithmic_bytes(n_local, local_off, 4, symmetric)
        f32_leg = {"kernel_ms": ms32, "value": (rows + n_off) / (ms32 * 1e-3) if world == 1 else None, "unit": UNIT,
                   "algorithmic_bytes_per_launch": bytes32, "roofline_frac_hbm": bytes32 / (ms32 * 1e-3) / 1e9 / peak,
                   "rel_l2_vs_f64": float(((y32[:n_local].double() - ylocal[:n_local]).norm() /
                                           ylocal[:n_local].norm().clamp_min(1e-300)).item()) if n_local else 0.0}
        del x32, y32
    traffic = None
    try:  # DRAM bytes per launch of the dominant kernel from the committed ncu capture (one-GPU runs only)
        if world > 1:
            raise LookupError
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            traffic = json.load(f).get(args.config, {}).get("dram_bytes_per_launch")
    except Exception:
        pass

    # end-to-end through the reference-facing C ABI with host buffers
    e2e = None
    y_host = None
    if 2 * n * es * world <= args.e2e_host_gb * 1e9:
        y_host = torch.zeros(n, dtype=t_dtype).pin_memory().numpy()
        x_host_t = torch.empty(n, dtype=t_dtype).pin_memory()
        if world == 1:
            x_host_t.copy_(xfull[:n])
        else:  # back to global row order for the host-pointer call
            pos = torch.from_numpy(rd.global_to_position(np.arange(n, dtype=np.uint64)).astype(np.int64)).to(dev)
            x_host_t.copy_(xfull[pos])
            del pos
        x_host = x_host_t.numpy()
        e2e_steps = max(3, min(args.steps, 10))
        for _ in range(2):
            ffi.inplaceApply(op, x_host, y_host)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            ffi.inplaceApply(op, x_host, y_host)
        barrier()
        e2e_t = torch.tensor([(time.perf_counter() - t0) / e2e_steps], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
        e2e = {"value": (rows + n_off) / float(e2e_t.item()), "unit": UNIT, "h2d_bytes_per_step": n * es,
               "d2h_bytes_per_step": n * es}
        # consistency of the two paths (same rows, same data)
        dev_y = ylocal[:n_local].cpu().numpy()
        if n_local and not np.array_equal(dev_y, y_host[rd.local_rows().astype(np.int64)]):
            raise SystemExit("device-resident and host-pointer matvec disagree")
    else:
        e2e = {"value": None, "unit": UNIT, "h2d_bytes_per_step": n * es, "d2h_bytes_per_step": n * es,
               "skipped": f"pinned host buffers would need {2 * n * es * world / 1e9:.0f} GB on this node (--e2e-host-gb)"}

    mf_achieved = alg_bytes / (matrix_free_ms * 1e-3) / 1e9
    extra = {"basis_build_s": build_wall, "basis_build_device_s": ffi.basisBuildSeconds(basis), "rows": rows,
             "offdiag_elements": n_off, "program": ffi.basisProgramStats(basis), "peak_source": peak_src,
             "kernel_ms": kern_ms, "kernel_ms_per_rank": kern_all, "operator_cache": cache_info,
             "float32_storage": f32_leg, "cached_variant": os.environ.get("SPED_CACHED_VARIANT", "default"),
             "matrix_free": {"ms_per_step": matrix_free_ms, "value": (rows + n_off) / (matrix_free_ms * 1e-3),
                             "unit": UNIT, "roofline_frac_hbm": mf_achieved / peak,
                             "note": "integer-ALU bound: canonicalisation over the symmetry group per element"}}
    if not args.no_eigh:
        ffi.operatorSetCache(op, -1)  # drop the cache: time-to-ground-state includes building it
        del xfull, xshard, ylocal      # the solver allocates its own vectors (42 spins: 25.6 GB each)
        torch.cuda.empty_cache()
        barrier()
        t0 = time.perf_counter()
        evals, _, rnorms = ffi.eigh(op, np.dtype(np_dtype), spec.number_vectors, spec.precision, spec.max_primme_basis_size,
                                    spec.max_primme_block_size, spec.min_primme_restart_size, want_vectors=False)
        barrier()
        st = ffi.eighLastStats(op)
        extra.update({"time_to_ground_state_s": time.perf_counter() - t0, "eigenvalues": [float(v) for v in evals],
                      "residual_norms": [float(v) for v in rnorms], "eigh_matvecs": st["matvecs"],
                      "eigh_restarts": st["restarts"], "eigh_seconds_matvec": st["seconds_matvec"], "eigh_stats": st,
                      "eigh_dtype": "f64 (deck asks " + spec.datatype + ")"})

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu and y_host is not None:
        from helpers import oracle_problem
        from oracle import oracle as O

        O.build()
        ob, terms = oracle_problem(O, cfg)
        ob.build(np.array(ffi.basisGetStates(basis)))  # adopt the representatives (oracle computes its own norms)
        oop = O.Operator(ob, terms)
        xs, ys, stride = cpu_sample_rate(O, oop, n, args.cpu_seconds, np_dtype)
        t0 = time.perf_counter()
        e = oop.matmat_rows(xs, ys, 0, n, stride)
        dt = time.perf_counter() - t0
        srows = len(range(0, n, stride))
        # the sampled rows must agree with the GPU result (full-size parity on the sample)
        gpu_rows = y_host[0:n:stride]
        err = np.linalg.norm(ys[0:n:stride] - gpu_rows) / max(np.linalg.norm(gpu_rows), 1e-300)
        extra["sample_parity_rel_l2"] = float(err)
        cpu_baseline = {"value": (srows + e) / dt, "unit": UNIT, "cores": O.num_threads(), "kind": "port",
                        "sample": f"{srows} of {n} rows (every {stride}-th), {srows + e} matrix elements, {dt:.1f} s; "
                                  "oracle port (OpenMP), not the upstream binary"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64" if is_real else "c128", "data": "synthetic",
            "config": {"workload": args.config, "rows": rows, "offdiag_elements": n_off, "block_size": 1,
                       "path": ("operator elements cached in HBM by the first (matrix-free) application; steady-state "
                                "matvec streams them" if cache_info["ready"] else "matrix-free every application"),
                       "parallelism": f"rows dealt block-cyclically over {world} GPU(s); per matvec one NCCL all-gather of the "
                                      "Krylov vector, overlapped with the local-source part of the product",
                       "l2": "inputs larger than L2 (no flush)" if alg_bytes > 126e6 else "inputs fit in L2 (no flush)"},
            "clocks": clocks,
            "e2e": e2e,
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "algorithmic_bytes_per_launch": alg_bytes},
            "cpu_baseline": cpu_baseline, "extra": extra,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        ffi.commFinalize()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default=DEFAULT_DECK)
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--e2e-host-gb", type=float, default=24.0, help="skip the host-buffer leg above this much pinned memory")
    ap.add_argument("--watchdog-seconds", type=float, default=1500.0,
                    help="abort the process if the whole run takes longer (a mismatched collective hangs every rank)")
    ap.add_argument("--no-eigh", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else max(args.warmup, 1)
    if args.watchdog_seconds > 0:
        import threading

        def _abort():
            log(f"bench.py: no result after {args.watchdog_seconds:.0f} s, aborting")
            os._exit(3)

        timer = threading.Timer(args.watchdog_seconds, _abort)
        timer.daemon = True
        timer.start()
    # exactly one JSON line on stdout: native libraries (NCCL's version banner) write to fd 1, so
    # fd 1 is pointed at stderr for the run and the result line goes to the saved descriptor
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(saved, "w")
    rc = run_reference(args) if args.impl == "reference" else run_ours(args)
    sys.stdout.flush()
    sys.exit(rc)


if __name__ == "__main__":
    main()
