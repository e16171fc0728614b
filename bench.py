#!/usr/bin/env python
"""Benchmark of the accessor-BLAS hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
                    [--workload config2|config5] [--no-detail]

A "step" is one GEMV pass y = A x + y over one resident matrix
(BASELINE.json configs[1]: m = n = 16384, fp64 arithmetic on fp32 storage,
uniform(-1,1) data generated on the device from the reference's stream).
`value` is algorithmic GB/s (m*n*s + n*s + 2*m*s bytes per step, SURVEY.md
section 8(d)) with the operands resident in HBM; `e2e` is the same metric
through the host-buffer C-ABI call (pinned host memory -> H2D -> kernel ->
D2H inside the timed region).  With N > 1 (torchrun) every rank owns one such
slab of a (16384*N) x 16384 row-sharded matrix (weak scaling, x broadcast
once, no data-path collective).

`--impl reference` times the CPU port of the reference's accessor kernels
(oracle/cpu_baseline.cpp) on the host cores for the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

M = N_COLS = 16384
METRIC = json.loads((ROOT / "BASELINE.json").read_text())["metric"] \
    if (ROOT / "BASELINE.json").exists() else "GEMV/DOT GB/s"
NOMINAL_HBM_GBS = 8000.0
FALLBACK_HBM_GBS = 6650.0


def gemv_bytes(m, n, s):
    return m * n * s + n * s + 2 * m * s


def dot_bytes(n, s, r):
    return 2 * n * s + r


def trsv_bytes(n, s):
    return n * (n + 1) // 2 * s + 2 * n * s


def measured_peak():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured"
        except Exception:
            pass
    return FALLBACK_HBM_GBS, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu_index = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.gpu_index}", f"--query-gpu={self.QUERY}",
                 "--format=csv,noheader,nounits", "-lms", "50"],
                stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            for line in Path(self.path).read_text().splitlines():
                f = [c.strip() for c in line.split(",")]
                if len(f) < 9:
                    continue
                sm.append(float(f[1]))
                smax.append(float(f[2]))
                for name, val in zip(names, f[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
        except Exception:
            pass
        finally:
            try:
                os.unlink(self.path)
            except OSError:
                pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(smax),
                "reasons": sorted(reasons), "samples": len(sm)}


def bind_to_gpu_numa_node(torch, local_rank):
    """N > 1: run this rank on the cores next to its GPU, so that the pinned
    staging buffers of the host-buffer path (first touch) and the copy-engine
    traffic stay on the GPU's own socket -- what an MPI launcher with GPU
    affinity does.  Round 1 measured N concurrent 1 GiB uploads from wherever
    torchrun happened to place the ranks.  Best effort; returns what was done
    (and why not) for the bench line."""
    info = {"bound": False, "source": None, "cpus": None, "node": None, "notes": []}
    allowed = os.sched_getaffinity(0)

    def apply(cpus, source, node=None):
        cpus = set(cpus) & allowed
        if cpus and cpus != allowed:
            os.sched_setaffinity(0, cpus)
            info.update(bound=True, source=source, cpus=len(cpus), node=node)
            return True
        info["notes"].append(f"{source}: no narrower set than the {len(allowed)} allowed cpus")
        return False

    bdf = None
    try:
        prop = torch.cuda.get_device_properties(local_rank)
        bdf = f"{prop.pci_domain_id:04x}:{prop.pci_bus_id:02x}:{prop.pci_device_id:02x}.0"
        node = int(Path(f"/sys/bus/pci/devices/{bdf}/numa_node").read_text())
        if node < 0:
            info["notes"].append(f"sysfs numa_node of {bdf} is {node}")
        else:
            cpus = set()
            for part in Path(f"/sys/devices/system/node/node{node}/cpulist").read_text().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
            if apply(cpus, "sysfs", node):
                return info
    except Exception as exc:  # noqa: BLE001 - diagnostics only
        info["notes"].append(f"sysfs: {exc!r}")
    try:
        import pynvml
        pynvml.nvmlInit()
        hnd = (pynvml.nvmlDeviceGetHandleByPciBusId(bdf.encode()) if bdf
               else pynvml.nvmlDeviceGetHandleByIndex(local_rank))
        words = (max(allowed) + 64) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(hnd, words)
        cpus = {64 * w + bit for w, word in enumerate(mask) for bit in range(64) if (int(word) >> bit) & 1}
        if apply(cpus, "nvml"):
            return info
    except Exception as exc:  # noqa: BLE001
        info["notes"].append(f"nvml: {exc!r}")
    return info


def time_launches(fn, steps, warmup, torch, barrier=None):
    """ms per step: `warmup` untimed calls, then exactly `steps` calls between
    two CUDA events on the launching (current) stream, synchronised both sides."""
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    if barrier:
        barrier()
        torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    if barrier:
        barrier()
    return e0.elapsed_time(e1) / steps


def min_of_10(fn, torch):
    """The reference's protocol (cuda/utils.cuh:236-262): one warm-up, ten
    single calls bracketed by events, minimum."""
    fn()
    torch.cuda.synchronize()
    best = float("inf")
    for _ in range(10):
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        # keep the GPU busy for ~20 us while the host enqueues event, call and
        # event: otherwise the interval also counts the Python/ctypes time
        # between record() and the launch (5-10 us, 10 % of a 80 us kernel),
        # which the reference's C++ loop does not have
        torch.cuda._sleep(40_000)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


# ---------------------------------------------------------------------------
# reference arm / CPU baseline: the oracle's port on the host cores
# ---------------------------------------------------------------------------
def host_threads():
    """Host threads this process may use.  torchrun exports OMP_NUM_THREADS=1 to
    its workers; the CPU arm sets its thread count explicitly instead."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def cpu_gemv_run(steps, warmup, budget_s=25.0, slabs=1):
    """Times the CPU port on the SAME workload (GEMV of `slabs` 16384 x 16384
    slabs = the row-sharded matrix the accblas arm spreads over `slabs` GPUs,
    Acc<fp64,fp32>, all host threads).  The only place bench.py executes
    anything under oracle/."""
    import numpy as np
    sys.path.insert(0, str(ROOT / "tests"))
    from oracle_binding import Oracle
    orc = Oracle()
    threads = host_threads()
    orc.L.oracle_set_num_threads(threads)
    n = N_COLS
    m = M * slabs
    A = np.empty(m * n, dtype=np.float32)
    for sl in range(slabs):       # slab by slab: the fp64 draws are 2 GiB each
        A[sl * M * n:(sl + 1) * M * n] = orc.uniform(M * n, seed=42, first_draw=sl * M * n)
    x = orc.uniform(n, seed=42, first_draw=m * n).astype(np.float32)
    y = orc.uniform(m, seed=42, first_draw=m * n + n).astype(np.float32)
    for _ in range(max(1, min(warmup, 2))):
        orc.cpu_gemv(np.float64, A, m, n, n, x, 1.0, 1.0, y)
    times = []
    t_begin = time.perf_counter()
    for _ in range(steps):
        t0 = time.perf_counter()
        orc.cpu_gemv(np.float64, A, m, n, n, x, 1.0, 1.0, y)
        times.append(time.perf_counter() - t0)
        if time.perf_counter() - t_begin > budget_s:
            break
    ms = 1e3 * sum(times) / len(times)
    total_bytes = slabs * gemv_bytes(M, n, 4)
    gbs = total_bytes / (ms * 1e-3) / 1e9
    return {"value": gbs, "unit": "GB/s", "cores": orc.num_threads, "kind": "port",
            "sample": f"full workload (GEMV {m}x{n} Acc<fp64,fp32>), {len(times)} passes, "
                      f"OpenMP over rows, {orc.num_threads} threads",
            "ms_per_step": ms, "steps_timed": len(times), "total_bytes": total_bytes}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = max(1, args.gpus)
    base = cpu_gemv_run(args.steps, args.warmup, budget_s=120.0, slabs=world)
    line = {
        "impl": "reference", "metric": METRIC, "value": base["value"], "unit": "GB/s",
        "n_gpus": args.gpus, "steps": base["steps_timed"], "warmup": args.warmup,
        "ms_per_step": base["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(world, M, base["total_bytes"]),
        "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": base["value"], "unit": "GB/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "the reference ships no CPU build of its kernels; this is the host port "
                "of the accessor kernel bodies (oracle/cpu_baseline.cpp) on the same "
                f"{world} x (16384 x 16384) row slabs the accblas arm spreads over {world} GPU(s)",
    }
    print(json.dumps(line), flush=True)


def workload_name():
    return ("configs[1]: GEMV m=n=16384, fp64 arithmetic on fp32 storage "
            "(Acc<fp64,fp32>), alpha=beta=1, uniform(-1,1) seed 42")


def workload_config(world, rows_per_gpu, total_bytes):
    """`config` of the JSON line: identical keys and values for both arms."""
    return {"workload": workload_name(), "m_per_gpu": rows_per_gpu, "n": N_COLS,
            "arithmetic": "fp64", "storage": "fp32",
            "parallelism": f"row-sharded x{world}" if world > 1 else "single GPU",
            "l2": "inputs exceed L2 (1 GiB matrix per GPU vs 126 MB L2)",
            "algorithmic_bytes_per_step": total_bytes}


# ---------------------------------------------------------------------------
# detail: every (op, arithmetic, storage) pair + cuBLAS + TRSV (N = 1 only)
# ---------------------------------------------------------------------------
def detail_pairs(ab, h, torch, peak):
    import ctypes
    from accessor_blas_b200 import capi
    out = {}
    dev = torch.device("cuda", torch.cuda.current_device())
    stream = torch.cuda.current_stream().cuda_stream
    name = {torch.float64: "fp64", torch.float32: "fp32", torch.float16: "fp16"}
    size = {torch.float64: 8, torch.float32: 4, torch.float16: 2}
    try:
        bl = capi.load_baselines()
    except Exception:
        bl = None
    # the reference's OWN CUDA kernels on this GPU (oracle/_ref, compiled from
    # /root/reference/cuda by oracle/Makefile): a labelled baseline row next to
    # every pair, never part of the accblas path
    refk = None
    try:
        sys.path.insert(0, str(ROOT / "tests"))
        from oracle_binding import REF_LIB, RefKernels
        if REF_LIB.exists():
            refk = RefKernels()
    except Exception:
        refk = None

    # ---- GEMV 16384^2: fp64 master, converted on the device ----------------
    m = n = M
    A64 = torch.empty(m * n, dtype=torch.float64, device=dev)
    x64 = torch.empty(n, dtype=torch.float64, device=dev)
    y64 = torch.empty(m, dtype=torch.float64, device=dev)
    h.fill_uniform(m, n, A64, n, 42, 0)
    h.fill_uniform(n, 1, x64, 1, 42, m * n)
    h.fill_uniform(m, 1, y64, 1, 42, m * n + n)
    ref = y64.clone()
    h.gemv(torch.float64, m, n, 1.0, A64, n, x64, 1, 1.0, ref, 1)  # plain fp64 kernel
    gemv = {}
    for st in (torch.float64, torch.float32, torch.float16):
        if st == torch.float64:
            A, x, y0 = A64, x64, y64
        else:
            A = torch.empty(m * n, dtype=st, device=dev)
            x = torch.empty(n, dtype=st, device=dev)
            y0 = torch.empty(m, dtype=st, device=dev)
            h.convert(m, n, A64, n, A, n)
            h.convert(n, 1, x64, 1, x, 1)
            h.convert(m, 1, y64, 1, y0, 1)
        for ar in (torch.float64, torch.float32):
            y = y0.clone()
            h.gemv(ar, m, n, 1.0, A, n, x, 1, 1.0, y, 1)
            err = h.l1_error(m, ref, 1, y, 1)
            ms = min_of_10(lambda: h.gemv(ar, m, n, 1.0, A, n, x, 1, 0.0, y, 1), torch)
            gbs = gemv_bytes(m, n, size[st]) / (ms * 1e-3) / 1e9
            gemv[f"Acc<{name[ar]},{name[st]}>"] = {
                "ms": ms, "GBps": gbs, "frac_measured_peak": gbs / peak,
                "frac_nominal_8TBps": gbs / NOMINAL_HBM_GBS,
                "GFLOPps": (2 * m * n + 3 * m) / (ms * 1e-3) / 1e9,
                "rel_error_vs_fp64_kernel": err}
            if refk is not None:
                y = y0.clone()
                refk.gemv(ar, m, n, 1.0, A, n, x, 1, 1.0, y, 1)
                refk.sync()
                err = h.l1_error(m, ref, 1, y, 1)
                ms = min_of_10(lambda: refk.gemv(ar, m, n, 1.0, A, n, x, 1, 0.0, y, 1), torch)
                gbs = gemv_bytes(m, n, size[st]) / (ms * 1e-3) / 1e9
                gemv[f"reference kernel Acc<{name[ar]},{name[st]}>"] = {
                    "ms": ms, "GBps": gbs, "frac_measured_peak": gbs / peak,
                    "rel_error_vs_fp64_kernel": err}
        if bl is not None and st != torch.float16:
            y = y0.clone()
            code = 0 if st == torch.float64 else 1
            bl.accblas_baseline_cublas_gemv(code, m, n, 1.0, A.data_ptr(), n, x.data_ptr(),
                                            1, 1.0, y.data_ptr(), 1, stream)
            err = h.l1_error(m, ref, 1, y, 1)
            ms = min_of_10(lambda: bl.accblas_baseline_cublas_gemv(
                code, m, n, 1.0, A.data_ptr(), n, x.data_ptr(), 1, 0.0, y.data_ptr(), 1,
                stream), torch)
            gbs = gemv_bytes(m, n, size[st]) / (ms * 1e-3) / 1e9
            gemv[f"cuBLAS {name[st]}"] = {"ms": ms, "GBps": gbs,
                                          "frac_measured_peak": gbs / peak,
                                          "rel_error_vs_fp64_kernel": err}
        if st != torch.float64:
            del A
    out["gemv_16384"] = gemv
    del A64
    torch.cuda.empty_cache()

    # ---- DOT n = 2^28 ---------------------------------------------------------
    nd = 2 ** 28
    x64 = torch.empty(nd, dtype=torch.float64, device=dev)
    y64 = torch.empty(nd, dtype=torch.float64, device=dev)
    h.fill_uniform(1, nd, x64, nd, 42, 0)
    h.fill_uniform(1, nd, y64, nd, 42, nd)
    ref = torch.zeros(1, dtype=torch.float64, device=dev)
    h.dot(torch.float64, nd, x64, 1, y64, 1, ref)
    ref_v = ref.item()
    dot = {}
    for st in (torch.float64, torch.float32, torch.float16):
        if st == torch.float64:
            x, y = x64, y64
        else:
            x = torch.empty(nd, dtype=st, device=dev)
            y = torch.empty(nd, dtype=st, device=dev)
            h.convert(1, nd, x64, nd, x, nd)
            h.convert(1, nd, y64, nd, y, nd)
        for ar in (torch.float64, torch.float32):
            # result stored in the storage type when it is fp32 (as the
            # reference driver does), else in the arithmetic type
            res_t = torch.float32 if st == torch.float32 else ar
            res = torch.zeros(1, dtype=res_t, device=dev)
            h.dot(ar, nd, x, 1, y, 1, res)
            got = float(res.item())
            ms = min_of_10(lambda: h.dot(ar, nd, x, 1, y, 1, res), torch)
            gbs = dot_bytes(nd, size[st], size[res_t]) / (ms * 1e-3) / 1e9
            dot[f"Acc<{name[ar]},{name[st]}>"] = {
                "ms": ms, "GBps": gbs, "frac_measured_peak": gbs / peak,
                "frac_nominal_8TBps": gbs / NOMINAL_HBM_GBS,
                "GFLOPps": 2 * nd / (ms * 1e-3) / 1e9,
                "rel_error_vs_fp64_kernel": abs(got - ref_v) / abs(ref_v)}
            if refk is not None:
                refk.dot(ar, nd, x, 1, y, 1, res)
                refk.sync()
                got = float(res.item())
                ms = min_of_10(lambda: refk.dot(ar, nd, x, 1, y, 1, res), torch)
                gbs = dot_bytes(nd, size[st], size[res_t]) / (ms * 1e-3) / 1e9
                dot[f"reference kernel Acc<{name[ar]},{name[st]}>"] = {
                    "ms": ms, "GBps": gbs, "frac_measured_peak": gbs / peak,
                    "rel_error_vs_fp64_kernel": abs(got - ref_v) / abs(ref_v)}
        if bl is not None and st != torch.float16:
            code = 0 if st == torch.float64 else 1
            res = torch.zeros(1, dtype=st, device=dev)
            bl.accblas_baseline_cublas_dot(code, nd, x.data_ptr(), 1, y.data_ptr(), 1,
                                           res.data_ptr(), stream)
            got = float(res.item())
            ms = min_of_10(lambda: bl.accblas_baseline_cublas_dot(
                code, nd, x.data_ptr(), 1, y.data_ptr(), 1, res.data_ptr(), stream), torch)
            gbs = dot_bytes(nd, size[st], size[st]) / (ms * 1e-3) / 1e9
            dot[f"cuBLAS {name[st]}"] = {"ms": ms, "GBps": gbs,
                                         "frac_measured_peak": gbs / peak,
                                         "rel_error_vs_fp64_kernel": abs(got - ref_v) / abs(ref_v)}
        if st != torch.float64:
            del x, y
    out["dot_2^28"] = dot
    del x64, y64
    torch.cuda.empty_cache()

    # ---- TRSV n = 16384 on the pivoted LU of the uniform(-1,1) fixture ----------
    # Row-major LU = [L \\ U]: lower/unit = L (BASELINE configs[3]).  Its
    # transpose has the layout the reference's fixture has after cuSOLVER's
    # column-major getrf (cuda/trsv_memory.cuh:131-168): upper/unit = L^T, the
    # reference driver's default (cuda/trsv_benchmark.cu:26-27), and
    # lower/non-unit = U^T, conditioned like the random matrix itself.
    nt = 16384
    g = torch.empty(nt * nt, dtype=torch.float64, device=dev)
    h.fill_uniform(nt, nt, g, nt, 42, 0)
    LU, _ = torch.linalg.lu_factor(g.view(nt, nt))
    del g
    b64 = torch.empty(nt, dtype=torch.float64, device=dev)
    h.fill_uniform(nt, 1, b64, 1, 42, nt * nt)
    b32 = torch.empty(nt, dtype=torch.float32, device=dev)
    h.convert(nt, 1, b64, 1, b32, 1)
    trsv_all = {}
    cases = (("lower_unit", "lower, unit (L of LU)", False, ab.LOWER, ab.UNIT),
             ("upper_unit", "upper, unit (L^T; the reference driver's default)", True,
              ab.UPPER, ab.UNIT),
             ("lower_nonunit", "lower, non-unit (U^T; ill-conditioned for every solver)", True,
              ab.LOWER, ab.NON_UNIT))
    for key, label, transposed, uplo, diag in cases:
        M64 = (LU.t().contiguous() if transposed else LU.contiguous()).view(-1)
        A32 = torch.empty(nt * nt, dtype=torch.float32, device=dev)
        h.convert(nt, nt, M64, nt, A32, nt)
        trsv = {}
        xref = b64.clone()
        h.trsv(torch.float64, uplo, diag, nt, M64, nt, xref, 1)
        pairs = (("Acc<fp64,fp64>", torch.float64, M64, b64),
                 ("Acc<fp64,fp32>", torch.float64, A32, b32),
                 ("Acc<fp32,fp32>", torch.float32, A32, b32))
        if key != "lower_unit":
            pairs = pairs[1:2]
        for plabel, ar, A, b in pairs:
            xw = b.clone()
            h.trsv(ar, uplo, diag, nt, A, nt, xw, 1)
            err = h.l1_error(nt, xref, 1, xw, 1)

            def call():
                xw.copy_(b)
                h.trsv(ar, uplo, diag, nt, A, nt, xw, 1)
            ms = min_of_10(call, torch) - min_of_10(lambda: xw.copy_(b), torch)
            sz = A.element_size()
            nbytes = trsv_bytes(nt, sz) - (nt * sz if diag == ab.UNIT else 0)
            trsv[plabel] = {"ms": ms, "GBps": nbytes / (ms * 1e-3) / 1e9,
                            "frac_measured_peak": nbytes / (ms * 1e-3) / 1e9 / peak,
                            "rel_error_vs_fp64_kernel": err, "triangle": label}
            if refk is not None:
                xr = b.clone()
                refk.trsv(ar, uplo == ab.UPPER, diag == ab.UNIT, nt, A, nt, xr, 1)
                refk.sync()
                err = h.l1_error(nt, xref, 1, xr, 1)

                def rcall():
                    xr.copy_(b)
                    refk.trsv(ar, uplo == ab.UPPER, diag == ab.UNIT, nt, A, nt, xr, 1)
                ms = min_of_10(rcall, torch) - min_of_10(lambda: xr.copy_(b), torch)
                trsv[f"reference kernel {plabel}"] = {"ms": ms, "rel_error_vs_fp64_kernel": err}
        if bl is not None:
            cub = (("cuBLAS fp64", 0, M64, b64), ("cuBLAS fp32", 1, A32, b32))
            for clabel, code, A, b in (cub if key == "lower_unit" else cub[:1]):
                xw = b.clone()
                upper = 1 if uplo == ab.UPPER else 0
                unit = 1 if diag == ab.UNIT else 0
                bl.accblas_baseline_cublas_trsv(code, upper, unit, nt, A.data_ptr(), nt,
                                                xw.data_ptr(), 1, stream)
                err = h.l1_error(nt, xref, 1, xw, 1)

                def ccall():
                    xw.copy_(b)
                    bl.accblas_baseline_cublas_trsv(code, upper, unit, nt, A.data_ptr(), nt,
                                                    xw.data_ptr(), 1, stream)
                ms = min_of_10(ccall, torch) - min_of_10(lambda: xw.copy_(b), torch)
                trsv[clabel] = {"ms": ms, "rel_error_vs_fp64_kernel": err}
        trsv_all[key] = trsv
        del M64, A32
    out["trsv_16384_lower_unit"] = trsv_all["lower_unit"]
    out["trsv_16384_upper_unit"] = trsv_all["upper_unit"]
    out["trsv_16384_lower_nonunit"] = trsv_all["lower_nonunit"]
    return out


# ---------------------------------------------------------------------------
# main arm
# ---------------------------------------------------------------------------
def run_accblas_arm(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    import accessor_blas_b200 as ab
    from accessor_blas_b200 import sharded

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the accblas path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa_node = None
    if world > 1:
        numa_node = bind_to_gpu_numa_node(torch, local_rank)
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    barrier = (lambda: dist.barrier()) if world > 1 else None

    h = ab.Handle(local_rank)
    peak, peak_kind = measured_peak()

    if args.workload == "config5":
        return run_config5(args, ab, sharded, h, torch, dist, world, rank, dev, barrier, peak)

    # -- workload: this rank's slab of the row-sharded (M*world) x N matrix ------
    m_total, n = M * world, N_COLS
    first, rows = sharded.row_partition(m_total, world, rank)
    st, ar, s = torch.float32, torch.float64, 4
    A = torch.empty(rows * n, dtype=st, device=dev)
    x = torch.empty(n, dtype=st, device=dev)
    y = torch.empty(rows, dtype=st, device=dev)
    h.fill_uniform(rows, n, A, n, 42, first * n)       # draw r*n + c of the global stream
    if rank == 0:
        h.fill_uniform(n, 1, x, 1, 42, m_total * n)
    sharded.broadcast_vector(x)                          # once, outside the timed region
    h.fill_uniform(rows, 1, y, 1, 42, m_total * n + n + first)
    gemv = sharded.ShardedGemv(h, ar, m_total, n, n)

    def step():
        gemv(1.0, A, x, 1.0, y)

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms = time_launches(step, args.steps, args.warmup, torch, barrier)
    # the timed region can be far shorter than nvidia-smi's sampling period:
    # keep the identical back-to-back load running until 0.5 s have been sampled
    t_probe = time.perf_counter()
    while time.perf_counter() - t_probe < 0.5:
        for _ in range(50):
            step()
        torch.cuda.synchronize()
    clocks = sampler.stop() if rank == 0 else None
    if clocks is not None:
        clocks["window"] = "timed region + 0.5 s of the same back-to-back GEMV steps"
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    total_bytes = sum(gemv_bytes(sharded.row_partition(m_total, world, r)[1], n, s)
                      for r in range(world))
    value = total_bytes / (ms * 1e-3) / 1e9
    kernel_gbs = gemv_bytes(rows, n, s) / (ms * 1e-3) / 1e9   # one launch per step per GPU

    # -- e2e: host buffers through the C-ABI host entry point ---------------------
    A_h = torch.empty(rows * n, dtype=st).pin_memory()
    x_h = torch.empty(n, dtype=st).pin_memory()
    y_h = torch.empty(rows, dtype=st).pin_memory()
    A_h.copy_(A)
    x_h.copy_(x)
    y_h.copy_(y)
    A_np, x_np, y_np = A_h.numpy(), x_h.numpy(), y_h.numpy()

    def e2e_step():
        h.gemv_host(ar, rows, n, 1.0, A_np, n, x_np, 1, 1.0, y_np, 1)

    e2e_steps = max(3, min(args.steps, 10))
    for _ in range(2):
        e2e_step()
    torch.cuda.synchronize()
    if barrier:
        barrier()
    t0 = time.perf_counter()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(e2e_steps):
        e2e_step()
    e1.record()
    torch.cuda.synchronize()
    e2e_ms = max(e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3) / e2e_steps
    if world > 1:
        t = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
    e2e_value = total_bytes / (e2e_ms * 1e-3) / 1e9
    h2d = (rows * n + n + rows) * s
    d2h = rows * s

    # -- sharded DOT with one all-reduce (reported next to the headline) ---------
    extra = {}
    nd_total = 2 ** 28 * world
    d_first, d_count = sharded.range_partition(nd_total, world, rank)
    xd = torch.empty(d_count, dtype=st, device=dev)
    yd = torch.empty(d_count, dtype=st, device=dev)
    h.fill_uniform(1, d_count, xd, d_count, 42, d_first)
    h.fill_uniform(1, d_count, yd, d_count, 42, nd_total + d_first)
    # partials combined inside the DOT kernel over peer memory (NVLink); the
    # NCCL single-element all-reduce is timed next to it
    sdot = sharded.ShardedDot(h, ar, nd_total, fused=True)
    dot_ms = time_launches(lambda: sdot(xd, yd, torch.float32), args.steps, args.warmup,
                           torch, barrier)
    nccl_ms = None
    if world > 1:
        t = torch.tensor([dot_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dot_ms = float(t.item())
        ndot = sharded.ShardedDot(h, ar, nd_total, fused=False)
        nccl_ms = time_launches(lambda: ndot(xd, yd, torch.float32), args.steps,
                                args.warmup, torch, barrier)
        t = torch.tensor([nccl_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        nccl_ms = float(t.item())
    extra["dot_sharded"] = {
        "workload": f"DOT n=2^28 per GPU (total {nd_total}), Acc<fp64,fp32>, " +
                    (("partials exchanged inside the kernel over peer memory"
                      if sdot.fused else "one 1-element NCCL all-reduce per step")
                     if world > 1 else "single rank"),
        "ms_per_step": dot_ms,
        "GBps": dot_bytes(nd_total, s, 4) / (dot_ms * 1e-3) / 1e9,
        "ms_per_step_nccl_allreduce": nccl_ms,
        "result": float(sdot(xd, yd, torch.float32).item())}
    del xd, yd
    torch.cuda.empty_cache()

    # -- e2e with the matrix RESIDENT: how the reference's drivers (and an
    #    iterative solver) run -- the matrix is uploaded once, every step moves
    #    x in and y out through pinned host memory around one GEMV
    #    (cuda/utils.cuh:236-262 times only the launch)
    y_res = y.clone()

    def e2e_resident_step():
        x.copy_(x_h, non_blocking=True)
        gemv(1.0, A, x, 1.0, y_res)
        y_h.copy_(y_res, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    res_steps = max(10, min(args.steps, 100))
    for _ in range(3):
        e2e_resident_step()
    if barrier:
        barrier()
    t0 = time.perf_counter()
    for _ in range(res_steps):
        e2e_resident_step()
    res_ms = (time.perf_counter() - t0) * 1e3 / res_steps
    if world > 1:
        t = torch.tensor([res_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        res_ms = float(t.item())
    extra["e2e_resident"] = {
        "what": "matrix uploaded once; per step x H2D + GEMV + y D2H through pinned host "
                "memory, host-synchronised (wall clock)",
        "value": total_bytes / (res_ms * 1e-3) / 1e9, "unit": "GB/s", "ms_per_step": res_ms,
        "h2d_bytes_per_step": n * s, "d2h_bytes_per_step": rows * s}
    del y_res

    # -- BASELINE configs[4] (the multi-GPU row): strong scaling of a FIXED
    #    64 GiB GEMV and a 2^32 DOT with its exchange, at every N
    if not args.no_config5:
        del A, A_h, A_np
        A = A_h = A_np = None
        torch.cuda.empty_cache()
        extra["config5"] = config5_numbers(args, sharded, h, torch, dist, world, rank, dev,
                                           barrier, peak)

    detail = None
    cpu = None
    if rank == 0 and world == 1:
        A = A_h = A_np = None
        torch.cuda.empty_cache()
        if not args.no_detail:
            detail = detail_pairs(ab, h, torch, peak)
        cpu = cpu_gemv_run(steps=8, warmup=1, budget_s=20.0)

    if rank == 0:
        traffic = None
        tp = ROOT / "profiles" / "roofline_traffic.json"
        if tp.exists():
            try:
                traffic = json.loads(tp.read_text()).get("gemv_f64_f32_16384_dram_bytes")
            except Exception:
                traffic = None
        line = {
            "metric": METRIC, "value": value, "unit": "GB/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": workload_config(world, rows, total_bytes),
            "roofline": {"bound": "hbm", "achieved": kernel_gbs, "peak": peak, "unit": "GB/s",
                         "frac": kernel_gbs / peak, "traffic": traffic,
                         "peak_source": f"MEASURED_PEAKS.json hbm_gbs ({peak_kind})",
                         "frac_of_nominal_8TBps": kernel_gbs / NOMINAL_HBM_GBS,
                         "kernel": "accblas::gemv_stream_kernel<float,double,ROWS=4,UNROLL=2,RG=1,COLW=8>",
                         "timing": "launches BACK TO BACK on one stream between two CUDA events: consecutive "
                                   "launches overlap their tail and ramp through programmatic dependent launch; "
                                   "isolated calls (min of 10) are in `pairs`"},
            "e2e": {"value": e2e_value, "unit": "GB/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms,
                    "api": "accblas_gemv_host (pinned host buffers)",
                    "host_affinity_rank0": numa_node},
            "gpu_launches": args.steps * world,
            "clocks": clocks,
            "cpu_baseline": ({k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}
                             if cpu else None),
            "extra": extra,
        }
        if detail is not None:
            line["pairs"] = detail
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def config5_numbers(args, sharded, h, torch, dist, world, rank, dev, barrier, peak):
    """BASELINE.json configs[4]: FIXED total problem, sharded over the ranks
    (strong scaling): GEMV m = n = 131072 fp32 storage (64 GiB) row-sharded with
    x broadcast once, and DOT n = 2^32 fp32 range-sharded with the partials
    combined inside the kernel over peer memory (the 1-element NCCL all-reduce
    is timed next to it).  Acc<fp64, fp32>.  Returns the numbers (all ranks)."""
    st, ar, s = torch.float32, torch.float64, 4
    m_total = n = 131072
    first, rows = sharded.row_partition(m_total, world, rank)
    A = torch.empty(rows * n, dtype=st, device=dev)
    x = torch.empty(n, dtype=st, device=dev)
    y = torch.empty(rows, dtype=st, device=dev)
    h.fill_uniform(rows, n, A, n, 42, first * n)
    if rank == 0:
        h.fill_uniform(n, 1, x, 1, 42, m_total * n)
    sharded.broadcast_vector(x)
    h.fill_uniform(rows, 1, y, 1, 42, m_total * n + n + first)
    gemv = sharded.ShardedGemv(h, ar, m_total, n, n)
    steps, warmup = min(args.steps, 20), max(3, min(args.warmup, 5))

    def max_over_ranks(v):
        if world > 1:
            t = torch.tensor([v], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return v

    ms = max_over_ranks(time_launches(lambda: gemv(1.0, A, x, 1.0, y), steps, warmup, torch,
                                      barrier))
    gemv_total = gemv_bytes(m_total, n, s)
    # checksum of this rank's slab of y after warmup + steps accumulating GEMVs
    # is not comparable across N; a single GEMV into a fresh y is
    y1 = torch.empty(rows, dtype=st, device=dev)
    h.fill_uniform(rows, 1, y1, 1, 42, m_total * n + n + first)
    gemv(1.0, A, x, 1.0, y1)
    checksum = y1.double().sum()
    if world > 1:
        dist.all_reduce(checksum, op=dist.ReduceOp.SUM)
    checksum = float(checksum.item())
    del A, y1
    torch.cuda.empty_cache()

    nd = 2 ** 32
    d_first, d_count = sharded.range_partition(nd, world, rank)
    xd = torch.empty(d_count, dtype=st, device=dev)
    yd = torch.empty(d_count, dtype=st, device=dev)
    h.fill_uniform(1, d_count, xd, d_count, 42, d_first)
    h.fill_uniform(1, d_count, yd, d_count, 42, nd + d_first)
    sdot = sharded.ShardedDot(h, ar, nd, fused=True)
    dot_ms = max_over_ranks(time_launches(lambda: sdot(xd, yd, torch.float32), steps, warmup,
                                          torch, barrier))
    dot_value = float(sdot(xd, yd, torch.float64).item())
    nccl_ms = None
    if world > 1:
        ndot = sharded.ShardedDot(h, ar, nd, fused=False)
        nccl_ms = max_over_ranks(time_launches(lambda: ndot(xd, yd, torch.float32), steps,
                                               warmup, torch, barrier))
    del xd, yd
    torch.cuda.empty_cache()
    return {
        "workload": "configs[4]: row-sharded GEMV m=n=131072 fp32 storage (64 GiB, FIXED total: "
                    "strong scaling) + DOT n=2^32 range-sharded, Acc<fp64,fp32>",
        "steps": steps, "warmup": warmup, "rows_per_gpu": rows,
        "gemv": {"ms_per_step": ms, "GBps": gemv_total / (ms * 1e-3) / 1e9,
                 "GBps_per_gpu": gemv_bytes(rows, n, s) / (ms * 1e-3) / 1e9,
                 "frac_measured_peak_per_gpu": gemv_bytes(rows, n, s) / (ms * 1e-3) / 1e9 / peak,
                 "collective": "none in the timed region (x broadcast once before it)",
                 "sum_of_y_after_one_gemv": checksum},
        "dot": {"n": nd, "ms_per_step": dot_ms,
                "GBps": dot_bytes(nd, s, 4) / (dot_ms * 1e-3) / 1e9,
                "ms_per_step_nccl_allreduce": nccl_ms,
                "result": dot_value,
                "collective": ("partials exchanged inside the kernel over peer memory (NVLink)"
                               if sdot.fused else "1-element all_reduce(SUM) per call")
                if world > 1 else "none (N=1)"},
    }


def run_config5(args, ab, sharded, h, torch, dist, world, rank, dev, barrier, peak):
    """`--workload config5`: the config-5 numbers as the main line."""
    c5 = config5_numbers(args, sharded, h, torch, dist, world, rank, dev, barrier, peak)
    if rank == 0:
        g = c5["gemv"]
        print(json.dumps({
            "metric": METRIC, "value": g["GBps"], "unit": "GB/s",
            "n_gpus": world, "steps": c5["steps"], "warmup": c5["warmup"],
            "ms_per_step": g["ms_per_step"],
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": c5["workload"], "rows_per_gpu": c5["rows_per_gpu"],
                       "l2": "inputs exceed L2"},
            "roofline": {"bound": "hbm", "achieved": g["GBps_per_gpu"],
                         "peak": peak, "unit": "GB/s",
                         "frac": g["frac_measured_peak_per_gpu"], "traffic": None},
            "gpu_launches": c5["steps"] * world,
            "extra": {"gemv_sum_of_y_after_one_gemv": g["sum_of_y_after_one_gemv"],
                      "dot": c5["dot"]},
        }), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="accblas", choices=["accblas", "reference"])
    ap.add_argument("--workload", default="config2", choices=["config2", "config5"],
                    help="config2 = BASELINE configs[1] (default, weak scaling); "
                         "config5 = BASELINE configs[4] (fixed 64 GiB GEMV + 2^32 DOT, strong)")
    ap.add_argument("--no-detail", action="store_true",
                    help="skip the per-pair / cuBLAS / TRSV detail table")
    ap.add_argument("--no-config5", action="store_true",
                    help="skip the configs[4] strong-scaling numbers in `extra`")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_accblas_arm(args)


if __name__ == "__main__":
    main()
