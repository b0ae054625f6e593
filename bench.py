#!/usr/bin/env python
"""Benchmark of the PINN training hot path (BASELINE.json): collocation points / s for one
residual-loss + weight-gradient evaluation, and full Adam steps / s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload ev|ns] [--n-f N]

Workload at every N (weak scaling): ev-NSFnet Re=2000 (main net 2-80x6-3, EVM net 2-40x4-1, EVM frozen
as on 9 999 of 10 000 steps), 1 000 000 synthetic uniform collocation points PER GPU + the reference's
2052 boundary points (sharded), fp32.  One "step" = one `nsf_step` (+ the gradient all-reduce when
N > 1).  `--workload ns` runs BASELINE config[1] (NSFnet 4x120, Re=1000, 1M points) instead.

Prints ONE JSON line (rank 0).  Keys beyond the base contract: `roofline` (dominant kernel, timed by
CUDA events inside the library on the launching stream), `cpu_baseline` (the autograd port of the
reference timed on this box's host cores), `e2e` (public solver API with pinned HOST buffers: H2D of
the points and D2H of the loss inside every timed step), `clocks`, `gpu_launches`, and metric (ii) of
BASELINE.json: `adam_steps_per_s` (fused iteration: nsf_step + device-resident Adam, one CUDA graph
replay), `adam_steps_per_s_torch_optim_loop` (the reference's loop body on the same kernels) and
`adam_small_batch` (both again at the shipped 120 000 points per GPU, where launch overhead decides).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

FLOPS_PER_PT = {"ev": 976_400.0, "ns": 1_305_840.0}      # algorithmic (5-stream) fwd + reverse
EXEC_FLOPS_PER_PT = {"ev": 30 * 80 * 80 * 5 * 0.8 + 9840.0, "ns": 30 * 120 * 120 * 3 * 0.8}  # 4 streams carried (laplacian merged)




def read_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["bf16_tflops"]), float(p["bf16_tflops_sustained"]), float(p["hbm_gbs"]), "measured"
    except Exception:
        return 1590.0, 1400.0, 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index), "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for nme, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nme)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def cpu_reference(workload, n_f, steps, warmup):
    """The reference on the host cores: the UNMODIFIED solvers under baseline/_ref (baseline/ref_arm.py, kind "reference") when that
    copy is present, else the pinned restatement oracle/autograd_port.py (kind "port": torch modules + 7 autograd.grad sweeps +
    backward + Adam, bit-identical loss curves to the reference, tests/test_oracle_golden.py).  Returns (pts/s, s/step, threads, kind)."""
    try:
        from baseline import ref_arm
        if ref_arm.available():
            sec, threads, _ = ref_arm.time_steps(workload, "cpu", n_f, steps, warmup)
            return n_f / sec, sec, threads, "reference"
    except Exception as e:  # noqa: BLE001
        sys.stderr.write(f"bench: unmodified reference arm unavailable ({e!r}); timing the pinned port\n")
    from oracle.autograd_port import time_reference_step
    sec, threads = time_reference_step(workload, n_f, steps=steps, warmup=warmup)
    return n_f / sec, sec, threads, "port"


def reference_cuda_eager(workload):
    """The unmodified reference's own CUDA path (eager autograd) on this GPU: the same-box competitor."""
    out = {}
    try:
        import torch
        from baseline import ref_arm
        if not ref_arm.available():
            return {"unavailable": "baseline/_ref not present"}
        for n_f, steps in ((100_000, 5), (500_000, 3)):
            torch.cuda.reset_peak_memory_stats()
            sec, _, _ = ref_arm.time_steps(workload, "cuda:0", n_f, steps, 2)
            out[f"n_f_{n_f}"] = {"ms_per_step": sec * 1e3, "pts_per_s": n_f / sec, "max_mem_GB": torch.cuda.max_memory_allocated() / 2 ** 30}
        out["what"] = "baseline/_ref (verbatim copy of the reference) through its own API: fwd_computing_loss_2d + backward + Adam, full step"
    except Exception as e:  # noqa: BLE001
        out["error"] = repr(e)[:200]
    return out


def measure_tf32_peak():
    """Dense TF32 tensor throughput of THIS GPU (torch.matmul 8192^3, allow_tf32): best of 10 (burst) and back to back for 1.5 s."""
    import torch
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        n = 8192
        a = torch.randn(n, n, device="cuda"); b = torch.randn(n, n, device="cuda")
        for _ in range(3):
            a @ b
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(10):
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); a @ b; e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        t0 = time.time(); k = 0
        e0.record()
        while time.time() - t0 < 1.5:
            for _ in range(20):
                a @ b
            k += 20
            torch.cuda.synchronize()
        e1.record(); torch.cuda.synchronize()
        return 2 * n ** 3 / (best * 1e-3) / 1e12, 2 * n ** 3 * k / (e0.elapsed_time(e1) * 1e-3) / 1e12
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old


def ncu_traffic(workload, path, n):
    """dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the dominant kernel, from the committed `ncu --set full` capture
    of this command line (profiles/r2_ncu_traffic.json, written by scripts/ncu_summary.py); None when no capture matches."""
    try:
        with open(os.path.join(ROOT, "profiles", "r2_ncu_traffic.json")) as f:
            t = json.load(f)
        e = t.get(f"{workload}/path{path}/{n}")
        return (float(e["bytes"]), e.get("source")) if e else (None, None)
    except Exception:
        return None, None


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    n_f = args.cpu_n_f
    t0 = time.time()
    pts_s, sec, threads, kind = cpu_reference(args.workload, n_f, max(1, args.steps), max(1, args.warmup))
    what = "the unmodified reference (baseline/_ref) on the host cores" if kind == "reference" else "reference algorithm on the host cores (torch autograd, oracle/autograd_port.py)"
    line = {"impl": "reference", "metric": "collocation_pts_per_s (residual + weight gradient)", "value": pts_s, "unit": "pts/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": dict(workload_config(args, n_f_override=n_f), kernel_path=what,
                           l2="n/a (CPU run)", sample_of=f"the {args.n_f}-point workload: {n_f} collocation points per step"),
            "cpu_baseline": {"value": pts_s, "unit": "pts/s", "cores": threads, "kind": kind,
                             "sample": f"{n_f} collocation pts + 2052 boundary pts per step, full Adam step (N_f=1e6 does not fit host memory for the reference's retained graph)"},
            "e2e": {"value": pts_s, "unit": "pts/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "wall_s": time.time() - t0}
    print(json.dumps(line), flush=True)


def workload_config(args, n_f_override=None):
    n_f = n_f_override if n_f_override is not None else args.n_f
    if args.workload == "ev":
        name = f"ev-NSFnet Re=2000 cavity, main 2-80x6-3 + EVM 2-40x4-1 (frozen), {n_f} uniform collocation pts/GPU + 2052 boundary pts"
    else:
        name = f"NSFnet Re=1000 cavity, 2-120x4-3, {n_f} uniform collocation pts/GPU + 2052 boundary pts"
    return {"workload": name, "n_f_per_gpu": n_f, "n_b": 2052, "parallelism": f"dp{args.gpus} over points",
            "l2": "flushed between timed iterations (256 MiB write)", "kernel_path": "ffma"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="ev", choices=["ev", "ns"])
    ap.add_argument("--n-f", type=int, default=1_000_000)
    ap.add_argument("--cpu-n-f", type=int, default=100_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-small", action="store_true", help="skip the 120 000-point Adam-iteration measurement")
    ap.add_argument("--path", type=int, default=0, help="0 auto, 1 FFMA, 3 tcgen05 (points on M)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import torch.distributed as dist
    from nsfnet_b200.cavity_data import cavity_boundary

    world = int(os.environ.get("WORLD_SIZE", 1))
    rank = int(os.environ.get("RANK", 0))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    dev = torch.device(f"cuda:{local}")

    if args.workload == "ev":
        from nsfnet_b200.ev_nsfnet import PysicsInformedNeuralNetwork
        torch.manual_seed(0)
        P = PysicsInformedNeuralNetwork(Re=2000, layers=6, hidden_size=80, layers_1=4, hidden_size_1=40, N_f=args.n_f * world,
                                        alpha_evm=0.05, bc_weight=10, eq_weight=1, supervised_data_weight=0.0)
    else:
        from nsfnet_b200.nsfnet import PysicsInformedNeuralNetwork
        torch.manual_seed(0)
        P = PysicsInformedNeuralNetwork(Re=1000, layers=4, hidden_size=120, N_f=args.n_f * world, bc_weight=10, eq_weight=1)
    if args.path:
        P._ctx.set_path(args.path)
    P.log_interval = 10 ** 9
    P.checkpoints = False
    P.set_boundary_data(cavity_boundary(513))
    # every rank owns its own synthetic shard, generated on the device; the solver's sharding setter is
    # bypassed for the collocation set because a W*1e6-point host array per rank is exactly what DP avoids
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    n = args.n_f
    x = torch.rand(n, device=dev, generator=g); y = torch.rand(n, device=dev, generator=g)
    P.set_eq_training_shard((x, y), n_global=n * world)
    if args.workload == "ev":
        P.freeze_evm_net(0)
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- data-parallel identity (W > 1): the all-reduced gradient of W shards == one GPU on the union ----------------------
    dp_identity = None
    if world > 1:
        from nsfnet_b200.solver_core import shard_bounds
        n_g = 262_144
        gg = torch.Generator(device=dev).manual_seed(4321)           # the same global set on every rank
        gx = torch.rand(n_g, device=dev, generator=gg); gy = torch.rand(n_g, device=dev, generator=gg)
        lo, hi = shard_bounds(n_g, rank, world)
        P.set_eq_training_shard((gx[lo:hi].clone(), gy[lo:hi].clone()), n_global=n_g)
        if args.workload == "ev":
            P.freeze_evm_net(0)
        loss_dp = float(P._launch_step())
        g_dp = P._buf[:P._n_main].clone()
        sync_all()
        if rank == 0:                                                # one process, whole set, whole boundary, no collective
            ws, rk, isd = P.world_size, P.rank, P.is_distributed
            P.world_size, P.rank, P.is_distributed = 1, 0, False
            try:
                P.set_boundary_data(cavity_boundary(513))
                P.set_eq_training_shard((gx, gy), n_global=n_g)
                if args.workload == "ev":
                    P.freeze_evm_net(0)
                loss_1 = float(P._launch_step())
                g_1 = P._buf[:P._n_main].clone()
            finally:
                P.world_size, P.rank, P.is_distributed = ws, rk, isd
            dp_identity = {"n_f_global": n_g, "grad_rel": float((g_dp - g_1).norm() / g_1.norm()), "loss_rel": abs(loss_dp - loss_1) / abs(loss_1),
                           "what": f"all-reduced gradient of {world} shards vs the same {n_g} points + 2052 boundary points on one GPU"}
        sync_all()
        P.set_boundary_data(cavity_boundary(513))
        P.set_eq_training_shard((x, y), n_global=n * world)
        if args.workload == "ev":
            P.freeze_evm_net(0)
        if rank == 0 and not (dp_identity["grad_rel"] < 1e-6 and dp_identity["loss_rel"] < 1e-6):
            raise RuntimeError(f"data-parallel identity violated: {dp_identity}")

    # ---- metric (i): residual + weight gradient, inputs resident in HBM -------------------------
    P._ctx.set_timing(True)
    for _ in range(args.warmup):
        P._launch_step()
    sync_all()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    kern_ms = []
    launches = 0
    sync_all()
    for i in range(args.steps):
        flush.fill_(float(i))
        ev[i][0].record()
        P._launch_step()
        ev[i][1].record()
        launches += P._ctx.info()["launches"]
        kern_ms.append(P._ctx.last_kernel_ms())      # syncs on the jet kernel's closing event only
    sync_all()
    step_ms = [a.elapsed_time(b) for a, b in ev]
    tot = torch.tensor([sum(step_ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tot, op=dist.ReduceOp.MAX)
    total_ms = float(tot.item())
    clocks = sampler.stop() if rank == 0 else None
    ms_per_step = total_ms / args.steps
    value = n * world / (ms_per_step * 1e-3)

    # ---- metric (ii): full Adam iterations (loss, backward, optimizer) ---------------------------
    # (a) the reference's own loop body on the drop-in classes: fwd_computing_loss_2d + loss.backward() + torch.optim.Adam
    # (b) the fused iteration: nsf_step + device-resident Adam replayed as one CUDA graph (one process;
    #     launched eagerly under data parallelism because of the NCCL all-reduce in the middle)
    def adam_iter():
        loss, _ = P.fwd_computing_loss_2d()
        P.opt.zero_grad()
        loss.backward()
        P.opt.step()
        return loss

    def time_iters(fn, steps):
        for _ in range(3):
            fn()
        sync_all()
        t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
        w0 = time.perf_counter()
        t0.record()
        for _ in range(steps):
            fn()
        t1.record()
        sync_all()
        wall_ms = (time.perf_counter() - w0) * 1e3 / steps
        ms = torch.tensor([max(t0.elapsed_time(t1) / steps, wall_ms)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return 1e3 / float(ms.item())

    def fused_on(S):
        S.enable_fused_step(True)
        if args.workload == "ev":
            S.freeze_evm_net(0)
        else:
            S._adam_reset()
        S._adam_set_lr(1e-3)

    P._ctx.set_timing(False)
    adam_torch_steps_s = time_iters(adam_iter, args.steps)
    fused_on(P)
    adam_steps_s = time_iters(P._fused_step_replayable, args.steps)
    P.enable_fused_step(False)
    small = None
    if not args.no_small:
        # the shipped configuration's point count (ev-NSFnet/configs/production.yaml: N_f = 120 000), where the iteration is
        # launch / Python bound rather than kernel bound
        n_s = 120_000
        x_s = torch.rand(n_s, device=dev, generator=g); y_s = torch.rand(n_s, device=dev, generator=g)
        P.set_eq_training_shard((x_s, y_s), n_global=n_s * world)
        if args.workload == "ev":
            P.freeze_evm_net(0)
        a = time_iters(adam_iter, 200)
        fused_on(P)
        b = time_iters(P._fused_step_replayable, 200)
        P.enable_fused_step(False)
        graphed = world == 1 or os.environ.get("NSF_FUSED_GRAPH_DDP", "1") == "1"
        small = {"n_f_per_gpu": n_s, "adam_steps_per_s_torch_optim_loop": a,
                 "adam_steps_per_s_fused_graph" if graphed else "adam_steps_per_s_fused_eager": b}
        P.set_eq_training_shard((x, y), n_global=n * world)
        if args.workload == "ev":
            P.freeze_evm_net(0)

    # ---- strong scaling: 1e6 collocation points IN TOTAL (BASELINE config 2), fused Adam iterations / s at this N ----------
    strong = None
    if not args.no_small:
        n_t = 1_000_000
        lo, hi = (rank * (n_t // world), (rank + 1) * (n_t // world) if rank < world - 1 else n_t)
        x_t = torch.rand(hi - lo, device=dev, generator=g); y_t = torch.rand(hi - lo, device=dev, generator=g)
        P.set_eq_training_shard((x_t, y_t), n_global=n_t)
        if args.workload == "ev":
            P.freeze_evm_net(0)
        fused_on(P)
        sps = time_iters(P._fused_step_replayable, 100)
        P.enable_fused_step(False)
        strong = {"n_f_total": n_t, "adam_steps_per_s": sps, "pts_per_s": sps * n_t,
                  "limiter": "fixed per-iteration part (EVM forward, boundary blocks, finalize, Adam, all-reduce), not NCCL bandwidth (152 KB per step)"}
        P.set_eq_training_shard((x, y), n_global=n * world)
        if args.workload == "ev":
            P.freeze_evm_net(0)

    # ---- e2e: public API with HOST buffers (pinned), H2D of the points + D2H of the loss every step
    e2e = None
    if not args.no_e2e:
        hx = torch.rand(n, 1).pin_memory(); hy = torch.rand(n, 1).pin_memory()

        def e2e_iter():
            P.set_eq_training_shard((hx, hy), n_global=n * world)   # H2D of this step's points (+ the reference's init_vis_t)
            loss, _ = P.fwd_computing_loss_2d()
            P.opt.zero_grad()
            loss.backward()
            return float(loss)                          # D2H of the result
        P.verbose = False
        for _ in range(2):
            e2e_iter()
        sync_all()
        w0 = time.perf_counter()
        for _ in range(args.steps):
            e2e_iter()
        sync_all()
        e2e_s = (time.perf_counter() - w0) / args.steps
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e = {"value": n * world / float(t.item()), "unit": "pts/s", "h2d_bytes_per_step": 8 * n, "d2h_bytes_per_step": 4,
               "ms_per_step": float(t.item()) * 1e3, "api": "set_eq_training_shard(pinned host x,y) + fwd_computing_loss_2d() + loss.backward() + float(loss)"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    bf16_burst, bf16_sust, hbm, how = read_peaks()
    k_ms = float(np.mean(kern_ms))
    flops = FLOPS_PER_PT[args.workload] * n
    achieved = flops / (k_ms * 1e-3) / 1e12
    # dense TF32 peak MEASURED on this GPU in this run.  The jet kernel runs at the full SM clock without a power cap (see
    # `clocks`), so the burst figure is its denominator; the sustained one (power-capped GEMM) is reported beside it.
    try:
        tf32_burst, tf32_sust = measure_tf32_peak()
        peak, peak_note = tf32_burst, "dense TF32 measured in this run: torch.matmul 8192^3, allow_tf32, best of 10 (burst)"
    except Exception as e:  # noqa: BLE001
        tf32_burst = tf32_sust = None
        peak, peak_note = 0.5 * bf16_burst, f"0.5 x bf16_tflops ({how}); the live TF32 measurement failed: {e!r}"
    info = P._ctx.info()
    traffic, traffic_src = ncu_traffic(args.workload, info["path"], n)
    cfg = workload_config(args)
    cfg["kernel_path"] = {1: "ffma", 3: "tcgen05-3xtf32, points on M (32 / 16-point tiles)"}[info["path"]]
    line = {"metric": "collocation_pts_per_s (residual + weight gradient)", "value": value, "unit": "pts/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
            "adam_steps_per_s": adam_steps_s, "adam_steps_per_s_torch_optim_loop": adam_torch_steps_s,
            "adam_iteration": "nsf_step + device-resident Adam (nsf_adam_dev), one CUDA graph replay per iteration" if world == 1
                              else ("nsf_step + NCCL all-reduce + device-resident Adam (nsf_adam_dev), " +
                                    ("one CUDA graph replay per iteration" if os.environ.get("NSF_FUSED_GRAPH_DDP", "1") == "1" else "launched eagerly")),
            "adam_small_batch": small, "strong": strong,
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_source": traffic_src,
                         "tf32_tflops_burst": tf32_burst, "tf32_tflops_sustained": tf32_sust,
                         "kernel": "collocation jet step (fwd jet + residuals + reverse)", "kernel_ms": k_ms,
                         "kernel_share_of_step": k_ms / ms_per_step, "flops_per_pt_algorithmic": FLOPS_PER_PT[args.workload],
                         "flops_per_pt_executed": EXEC_FLOPS_PER_PT[args.workload],
                         "peak_note": peak_note + "; fp32 FFMA peak 148*128*2*1.965GHz = 74.5 TFLOP/s",
                         "frac_of_ffma_peak": (EXEC_FLOPS_PER_PT[args.workload] * n / (k_ms * 1e-3) / 1e12) / 74.5,
                         "hbm_algorithmic_GBps": 20.0 * n / (k_ms * 1e-3) / 1e9, "hbm_peak_GBps": hbm},
            "gpu_launches": launches, "clocks": clocks}
    if e2e is not None:
        line["e2e"] = e2e
    if dp_identity is not None:
        line["dp_identity"] = dp_identity
    if not args.no_cpu_baseline:
        if world == 1:
            del P
            torch.cuda.empty_cache()
            line["reference_cuda_eager"] = reference_cuda_eager(args.workload)
        pts_s, sec, threads, kind = cpu_reference(args.workload, args.cpu_n_f, 3, 1)
        line["cpu_baseline"] = {"value": pts_s, "unit": "pts/s", "cores": threads, "kind": kind,
                                "sample": f"{args.cpu_n_f} collocation pts + 2052 boundary pts, 3 full Adam steps after 1 warm-up ({sec * 1e3:.0f} ms/step)"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
