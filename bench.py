#!/usr/bin/env python
"""Benchmark of the rSVD hot path (BASELINE.json: rSVD time-to-rank-k and GFLOP/s at 1/2/4/8 B200 vs the CPU path).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

Workload (config 5 of BASELINE.json): synthetic 200000 x 20000 FP64 matrix A = X diag(s) Y^T + 1e-6 N,
rank-100 sketch (l = 100), q = 2 power iterations, Jacobi back-end, host-supplied Omega.  One "step" = one complete rSVD.
With N GPUs A is row-sharded (strong scaling: the matrix is fixed), one process per GPU, NCCL for the two exchanges.

  value        algorithmic GFLOP/s = 12 m n l / time, whole job, A and Omega resident in HBM when the clock starts
  e2e          the same metric through rsvdb_rsvd_host: A (pinned host memory) -> device -> U, S, V back on the host,
               copies inside the timed region
  roofline     the six skinny GEMM passes over A (DMMA kernels): algorithmic flops / their device time, measured with
               CUDA events inside the timed steps, against the measured FP64 DMMA peak (profiles/FP64_PEAKS.json;
               MEASURED_PEAKS.json carries no FP64 number)
  cpu_baseline the CPU oracle (numpy/OpenBLAS + LAPACK QR + C Jacobi: an Eigen-free restatement of the reference
               algorithm -- the reference itself needs Eigen and MPI, absent from this image): ONE full rSVD of the
               SAME 200000 x 20000 matrix (about 20 s on the box's cores), whose singular values are also compared
               with the device path's (sigma_parity_vs_oracle, tolerance 1e-8 relative, north_star)

--impl reference times that same CPU restatement as the reference arm, on the full matrix (same config as our arm), with
the BLAS thread count forced to the CPUs this process may run on (torchrun exports OMP_NUM_THREADS=1) and reported as
the number of threads actually used.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))


def _host_threads() -> int:
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return os.cpu_count() or 1


# The CPU legs use every host core this process may run on.  torch.distributed.run exports OMP_NUM_THREADS=1 to its
# workers, which silently starved the reference arm at N >= 2 in round 1: override it BEFORE numpy/OpenBLAS load.
if "reference" in sys.argv[1:] or any(a.startswith("--impl=reference") for a in sys.argv[1:]):
    for _v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[_v] = str(_host_threads())
# the contract is ONE JSON line on stdout.  NCCL writes its "NCCL version ..." banner and its log to stdout; NCCL_DEBUG_FILE
# redirects them, but only for levels above VERSION -- so VERSION is raised to WARN (same banner, no extra output)
if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
    os.environ["NCCL_DEBUG"] = "WARN"
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")

M_FULL, N_FULL, L_FULL, Q_FULL = 200000, 20000, 100, 2
REF_WALL_BUDGET_S = 1200.0     # the reference arm stops early (and says so) rather than run past the driver's limit


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rows", type=int, default=M_FULL)      # development overrides; the driver never passes them
    ap.add_argument("--cols", type=int, default=N_FULL)
    ap.add_argument("--l", type=int, default=L_FULL)
    ap.add_argument("--q", type=int, default=Q_FULL)
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    return ap.parse_args()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons under load (B200_PROFILING.md).  The sampler runs from the warm-up steps to
    the end of the timed steps (same kernels, same load); samples taken while the GPU is busy (utilization >= 50 %) count."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,utilization.gpu,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20", "-i", str(self.index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True); self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        def num(x):
            try:
                return float(x)
            except Exception:
                return None
        rows = [r for r in self.rows if len(r) >= 9 and num(r[1]) is not None]
        busy = [r for r in rows if (num(r[4]) or 0) >= 50] or rows
        sm = sorted(num(r[1]) for r in busy)
        mx = [num(r[2]) for r in rows if num(r[2]) is not None]
        pw = [num(r[3]) for r in busy if num(r[3]) is not None]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [nm for i, nm in enumerate(names) if any(r[5 + i].lower().startswith("active") for r in busy)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "power_w_max": max(pw) if pw else None,
                "samples": len(rows), "samples_under_load": len(busy) if busy is not rows else 0, "reasons": reasons}


def workload_name(m, n, l, q):
    """The same string in both arms: BASELINE.json configs[4]."""
    return f"c5: rSVD rank-{l} q={q} of synthetic {m}x{n} FP64, Jacobi back-end, host-supplied Omega"


def flops(m, n, l, q):
    return (2 * q + 2) * 2.0 * m * n * l          # GEMM passes only (QR, small SVD, U = Q*Ut are overhead)


def fp64_peak():
    try:
        d = json.loads((ROOT / "profiles" / "FP64_PEAKS.json").read_text())
        return float(d["fp64_dmma_tflops"]), "measured FP64 DMMA issue rate, profiles/FP64_PEAKS.json (MEASURED_PEAKS.json has bf16/HBM only)"
    except Exception:
        return 37.2, "nominal 148 SM x 1.965 GHz x 128 flop/clk (profiles/FP64_PEAKS.json missing)"


# ---------------------------------------------------------------------------------------------------------------------
def blas_threads_in_use():
    """Threads the BLAS behind numpy will actually use (threadpoolctl), not os.cpu_count()."""
    try:
        from threadpoolctl import threadpool_info
        n = [int(d.get("num_threads", 0)) for d in threadpool_info() if d.get("user_api") == "blas"]
        return max(n) if n else None
    except Exception:
        return None


def cpu_oracle_gflops(A_host, Om, l, q, steps=1, warmup=0, wall_budget_s=None):
    """Time the CPU restatement on a host matrix (F-order view, never copied); returns (GFLOP/s, s per step, threads, S,
    steps actually timed, warm-up passes actually run).  With a wall budget the warm-up count, then the step count, are
    cut (and reported) so that the run ends inside it."""
    from oracle import rsvd_oracle as O
    from threadpoolctl import threadpool_limits
    O.build()
    want = _host_threads()
    with threadpool_limits(limits=want, user_api="blas"):
        threads = blas_threads_in_use() or want
        t_start = time.perf_counter()
        warmed = 0
        for i in range(warmup):
            tw = time.perf_counter()
            O.rsvd(A_host, Om, l, q, O.JACOBI)
            warmed += 1
            per = time.perf_counter() - tw
            if wall_budget_s is not None and (time.perf_counter() - t_start) + (warmup - warmed + steps) * per > wall_budget_s:
                break                                             # keep the budget for the timed steps
        done, S = 0, None
        t0 = time.perf_counter()
        for _ in range(steps):
            U, S, V = O.rsvd(A_host, Om, l, q, O.JACOBI)
            done += 1
            if wall_budget_s is not None and done < steps and (time.perf_counter() - t_start) + (time.perf_counter() - t0) / done > wall_budget_s:
                break
        dt = (time.perf_counter() - t0) / done
    m, n = A_host.shape
    return flops(m, n, l, q) / dt * 1e-9, dt, threads, S, done, warmed


def host_matrix_c5(m, n, dev, chunk=20000):
    """The full C5 matrix on the host as an F-order (column-major) m x n numpy view, generated block-wise on `dev`
    (bit-identical to what the GPU arm holds) -- 32 GB at full size, allocated once, never copied."""
    import numpy as np
    import torch
    from rsvd_kamaneh_raganato_terrana_b200 import workloads as W
    At = torch.empty((n, m), dtype=torch.float64)                   # (n, m) C-order == column-major m x n
    for r0 in range(0, m, chunk):
        r = min(chunk, m - r0)
        At[:, r0:r0 + r] = W.c5_shard_torch(m, n, r0, r, dev).cpu()
    return At.numpy().T


def run_reference(args):
    """Reference arm: the reference's algorithm (CPU restatement) on the host cores, full matrix, rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import numpy as np
    import torch
    from rsvd_kamaneh_raganato_terrana_b200 import workloads as W
    m, n, l, q = args.rows, args.cols, args.l, args.q
    dev = torch.device("cuda:0") if torch.cuda.is_available() else torch.device("cpu")
    torch.set_num_threads(_host_threads())
    A = host_matrix_c5(m, n, dev)
    if dev.type == "cuda":
        torch.cuda.empty_cache()
    Om = W.omega(n, l)
    g, dt, threads, _, done, warm = cpu_oracle_gflops(A, Om, l, q, steps=max(1, args.steps), warmup=max(0, args.warmup), wall_budget_s=REF_WALL_BUDGET_S)
    line = {
        "impl": "reference", "metric": "rsvd_gflops", "value": round(g, 2), "unit": "GFLOP/s", "n_gpus": args.gpus, "steps": done,
        "warmup": warm, "ms_per_step": round(dt * 1e3, 2), "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(m, n, l, q), "m": m, "n": n, "l": l, "q": q,
                   "parallelism": "host cores (the reference replicates the whole computation on every MPI rank)",
                   "steps_requested": args.steps, "warmup_requested": args.warmup},
        "cpu_baseline": {"value": round(g, 2), "unit": "GFLOP/s", "cores": threads, "kind": "port",
                         "host_cpus_available": _host_threads(), "omp_num_threads_env": os.environ.get("OMP_NUM_THREADS"),
                         "sample": f"the full {m}x{n} matrix, {done} complete rSVD(s) (numpy/OpenBLAS GEMM, LAPACK Householder QR, C Jacobi); the reference itself cannot be built (Eigen, MPI absent)"},
        "e2e": {"value": round(g, 2), "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def bind_to_gpu_cpus(torch, local):
    """Pin this rank to the CPUs next to its GPU (NVML's ideal affinity) so that the pinned host buffers of the end-to-end
    leg are allocated on the GPU's NUMA node -- what an MPI launcher's binding does for the reference.  Best effort."""
    try:
        import pynvml
        pynvml.nvmlInit()
        p = torch.cuda.get_device_properties(local)
        h = pynvml.nvmlDeviceGetHandleByPciBusId(f"{p.pci_domain_id:08x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0".encode())
        pynvml.nvmlDeviceSetCpuAffinity(h)
        return len(os.sched_getaffinity(0))
    except Exception:
        return None


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from rsvd_kamaneh_raganato_terrana_b200 import Engine, SVDMethod, workloads as W

    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    bound_cpus = bind_to_gpu_cpus(torch, local) if world > 1 else None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    m, n, l, q = args.rows, args.cols, args.l, args.q
    off, rows = W.row_split(m, world, rank)

    eng = Engine(local)                                   # raises without librsvdb.so / a B200: no fallback
    if world > 1:
        uid = torch.zeros(128, dtype=torch.uint8, device=dev)
        if rank == 0:
            uid = torch.frombuffer(bytearray(eng.comm_unique_id()), dtype=torch.uint8).to(dev)
        dist.broadcast(uid, 0)
        eng.comm_init(world, rank, bytes(uid.cpu().numpy().tobytes()))
    stream = torch.cuda.current_stream()
    eng.set_stream(stream.cuda_stream)

    A = W.c5_shard_torch(m, n, off, rows, dev)            # (n, rows): column-major rows x n shard
    Om = torch.from_numpy(W.omega(n, l).T.copy()).to(dev) # (l, n): column-major n x l
    U = torch.empty((l, rows), dtype=torch.float64, device=dev); V = torch.empty((l, n), dtype=torch.float64, device=dev)
    S = torch.empty(l, dtype=torch.float64, device=dev)

    def step():
        eng.rsvd_dev(A.data_ptr(), rows, n, rows, Om.data_ptr(), n, l, q, SVDMethod.Jacobi, U.data_ptr(), rows, S.data_ptr(), V.data_ptr(), n)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    sampler = ClockSampler(local); sampler.start()
    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    eng.set_profiling(True); eng.phase_ms()
    launches0 = eng.launches
    qr0 = eng.qr_path_counts()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    barrier()
    clocks = sampler.stop()
    total_ms = e0.elapsed_time(e1)
    phases = eng.phase_ms(); eng.set_profiling(False)
    launches = eng.launches - launches0
    qr1 = eng.qr_path_counts()
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    F = flops(m, n, l, q)
    value = F / (ms_per_step * 1e-3) * 1e-9
    s_dev = S.cpu().numpy()

    # roofline of the dominant kernels: the (2q+2) skinny GEMM passes of this rank's shard
    gemm_ms = (phases["gemm_an"] + phases["gemm_at"]) / args.steps
    peak, peak_src = fp64_peak()
    achieved = flops(rows, n, l, q) / (gemm_ms * 1e-3) * 1e-12 if gemm_ms > 0 else 0.0
    traffic = None
    try:   # dram__bytes_read + write of one k_gemm_an launch from the committed ncu --set full capture, scaled to this shard
        tj = json.loads((ROOT / "profiles" / "NCU_TRAFFIC.json").read_text())
        per = [tj[k]["dram_bytes_read"] + tj[k]["dram_bytes_write"] for k in ("k_gemm_an<13>", "k_gemm_at<13>") if k in tj]
        traffic = sum(per) / len(per) * (rows * n) / (200000.0 * 20000.0)          # mean over the two kernels (3 launches each per step)
    except Exception:
        pass
    roofline = {"bound": "tensor", "achieved": round(achieved, 3), "peak": peak, "unit": "TFLOP/s", "frac": round(achieved / peak, 4),
                "traffic": traffic, "traffic_note": "DRAM bytes per GEMM launch, mean of k_gemm_an / k_gemm_at (ncu --set full, profiles/r02_ncu_full_summary.txt, NCU_TRAFFIC.json), scaled to this shard; algorithmic bytes per launch = 8*rows*n",
                "algorithmic_bytes_per_launch": 8.0 * rows * n, "kernel": "k_gemm_an<13> / k_gemm_at<13> (FP64 DMMA + TMA), 6 passes per step",
                "peak_source": peak_src, "gemm_ms_per_step": round(gemm_ms, 3),
                "whole_step_frac_of_peak": round(value * 1e-3 / (world * peak), 4),
                "whole_step_note": "value / (n_gpus * peak): every phase (TSQR, small SVD, NCCL) counted; 'frac' covers the GEMM passes only",
                "phase_ms_per_step": {k: round(v / args.steps, 3) for k, v in phases.items()},
                "algorithmic_flops_per_launch": 2.0 * rows * n * l,
                "hbm_GBps_of_A_stream": round((2 * q + 2) * 8.0 * rows * n / (gemm_ms * 1e-3) * 1e-9, 1) if gemm_ms > 0 else None}

    # end to end through the host-pointer C ABI call: pinned host A -> device -> U, S, V on the host
    e2e = None
    if not args.no_e2e:
        Ah = torch.empty((n, rows), dtype=torch.float64, pin_memory=True); Ah.copy_(A)
        Omh = torch.empty((l, n), dtype=torch.float64, pin_memory=True); Omh.copy_(Om)
        Uh = torch.empty((l, rows), dtype=torch.float64, pin_memory=True); Vh = torch.empty((l, n), dtype=torch.float64, pin_memory=True)
        Sh = torch.empty(l, dtype=torch.float64, pin_memory=True)
        del A, U, V
        torch.cuda.empty_cache()

        def step_host():
            eng.rsvd_host_raw(Ah.data_ptr(), rows, n, rows, Omh.data_ptr(), n, 0, l, q, SVDMethod.Jacobi, Uh.data_ptr(), rows, Sh.data_ptr(), Vh.data_ptr(), n)

        step_host()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            step_host()                                   # synchronous: returns with U, S, V on the host
        barrier()
        dt = (time.perf_counter() - t0) / args.e2e_steps
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
        e2e = {"value": round(F / dt * 1e-9, 2), "unit": "GFLOP/s", "h2d_bytes_per_step": int(8 * (rows * n + n * l)),
               "d2h_bytes_per_step": int(8 * (rows * l + l + n * l)), "ms_per_step": round(dt * 1e3, 2), "steps": args.e2e_steps,
               "sigma_matches_device_path": bool(np.max(np.abs(Sh.numpy() - s_dev) / s_dev[0]) < 1e-12),
               "host_cpus_bound_per_rank": bound_cpus}

    cpu = None
    parity = None
    if rank == 0 and world == 1 and not args.no_cpu:
        # ONE full rSVD of the same matrix on the host cores (about 20 s): the CPU baseline and the parity check in one
        if e2e is not None:
            A_h = Ah.numpy().T                                         # F-order m x n view of the pinned buffer
        else:
            A_h = A.cpu().numpy().T
        g, dts, threads, S_cpu, _, _ = cpu_oracle_gflops(A_h, W.omega(n, l), l, q, steps=1, warmup=0)
        cpu = {"value": round(g, 2), "unit": "GFLOP/s", "cores": threads, "kind": "port", "seconds": round(dts, 2),
               "sample": f"the full {m}x{n} matrix, one complete rSVD (numpy/OpenBLAS GEMM, LAPACK Householder QR, C Jacobi); reference itself unbuildable here (Eigen, MPI absent)"}
        tol = 1e-8 * np.maximum(S_cpu, 1e-6 * S_cpu[0])
        rel = np.abs(s_dev - S_cpu) / np.maximum(S_cpu, 1e-6 * S_cpu[0])
        parity = {"max_rel_err": float(rel.max()), "tolerance": 1e-8, "ok": bool(np.all(np.abs(s_dev - S_cpu) <= tol)),
                  "n_sigma": int(len(S_cpu)), "same_matrix": True}

    if rank == 0:
        line = {
            "metric": "rsvd_gflops", "value": round(value, 2), "unit": "GFLOP/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": round(ms_per_step, 3), "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": workload_name(m, n, l, q),
                       "m": m, "n": n, "l": l, "q": q, "rows_per_gpu": rows, "parallelism": f"row-shard x{world}",
                       "l2": "inputs larger than L2 (A shard is %.1f GB)" % (rows * n * 8 / 1e9), "time_to_rank_k_ms": round(ms_per_step, 3)},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
            "sigma_head": [float(x) for x in s_dev[:4]], "sigma_parity_vs_oracle": parity,
            # how the 2q + 2 sketches per step were orthonormalised inside the timed region (csrc/cholqr.cu: guarded CholeskyQR2,
            # Householder TSQR when the measured guard refuses a sketch)
            "qr_paths": {"cholqr2": int(qr1[0] - qr0[0]), "householder_tsqr": int(qr1[1] - qr0[1])},
        }
        print(json.dumps(line), flush=True)
    eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
