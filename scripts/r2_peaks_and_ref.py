"""One-off measurements for round 2 (run under gpurun, results -> gpurun_out/r2_peaks_and_ref.json):
dense TF32 tensor peak (8192^3 torch.matmul, allow_tf32) as a burst and sustained for 4 s, and the UNMODIFIED
reference (baseline/_ref) timed on this box: CUDA eager on the B200 and on the host cores."""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def tf32_peak():
    torch.backends.cuda.matmul.allow_tf32 = True
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
    burst = 2 * n ** 3 / (best * 1e-3) / 1e12
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    t0 = time.time(); k = 0
    e0.record()
    while time.time() - t0 < 4.0:
        for _ in range(20):
            a @ b
        k += 20
        torch.cuda.synchronize()
    e1.record(); torch.cuda.synchronize()
    sust = 2 * n ** 3 * k / (e0.elapsed_time(e1) * 1e-3) / 1e12
    torch.backends.cuda.matmul.allow_tf32 = False
    return {"tf32_tflops": burst, "tf32_tflops_sustained": sust, "how": "torch.matmul fp32 8192^3, allow_tf32=True: best of 10 (burst), back to back for 4 s (sustained)"}


def main():
    from baseline import ref_arm
    out = {"gpu": torch.cuda.get_device_name(0), "host_cores": os.cpu_count()}
    out.update(tf32_peak())
    print(json.dumps(out), flush=True)
    ref = {}
    for wl, dev, n_f, steps in (("ev", "cuda:0", 100_000, 5), ("ev", "cuda:0", 500_000, 3), ("ns", "cuda:0", 100_000, 5), ("ns", "cuda:0", 500_000, 3),
                                ("ev", "cpu", 100_000, 2), ("ns", "cpu", 100_000, 2)):
        try:
            sec, thr, loss = ref_arm.time_steps(wl, dev, n_f, steps, 2)
            ref[f"{wl}_{dev.split(':')[0]}_{n_f}"] = {"sec_per_step": sec, "pts_per_s": n_f / sec, "threads": thr, "loss": loss,
                                                      "max_mem_GB": torch.cuda.max_memory_allocated() / 2 ** 30 if dev != "cpu" else None}
        except Exception as e:  # noqa: BLE001
            ref[f"{wl}_{dev.split(':')[0]}_{n_f}"] = {"error": repr(e)[:200]}
        torch.cuda.reset_peak_memory_stats()
        print(json.dumps(ref), flush=True)
    out["reference"] = ref
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "r2_peaks_and_ref.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
