#!/usr/bin/env python
"""bench.py — DualDiff denoising-step throughput on B200 (contract in the task statement / DESIGN.md §Measurement).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--scenes B] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" = one loop body of the sampler (pipeline_bev_controlnet.py:381-504) over a batch of B six-view
224x400 scenes per GPU: ControlNet-bg + ControlNet-fg + multi-view UNet + CFG + UniPC update, bf16 kernels.
Workload at N=1 = BASELINE.json configs[1] (B=8 scenes, CFG, UniPC) — `value` is scene-steps/s over all ranks
(weak scaling: B scenes per GPU, scenes are independent, no collectives).  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "denoise steps/sec (6-view 224x400 scene)"
UNIT = "scene-steps/s"
H, W = 28, 50
GFLOP_PER_IMAGE = 485.22          # UNet 307.84 + ControlNet-bg 91.20 + ControlNet-fg 86.17
TFLOP_PER_SCENE_STEP_CFG = 6 * 2 * GFLOP_PER_IMAGE / 1e3


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(tf_burst=d["bf16_tflops"], tf_sustained=d["bf16_tflops_sustained"], hbm=d["hbm_gbs"], src="measured")
    return dict(tf_burst=1590.0, tf_sustained=1400.0, hbm=6650.0, src="fallback")


def ncu_traffic(kernel):
    """DRAM bytes per launch of `kernel`, from the committed ncu launch list of one step (profiles/); None if absent"""
    p = os.path.join(ROOT, "profiles", "r01_launches_traffic.json")
    try:
        return round(json.load(open(p))[kernel]["traffic_bytes_per_launch"])
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks/throttle reasons DURING the timed region (B200_PROFILING.md recipe)"""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = sorted(int(float(r[1])) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit())
        mx = [int(float(r[2])) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


# -------------------------------------------------------------------------------------------------------
# our arm
# -------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import common
    from dualdiff_b200 import _lib, ops, synthetic as S
    from dualdiff_b200.pipeline import DualDiffDenoiser

    rank, world, local = dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — dualdiff_b200 has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    B, K, Wm = args.scenes, args.steps, args.warmup
    unet, nets, sds = common.build_models()
    for m in [unet] + nets:
        m.pack(dev)
    inp_cpu = S.make_inputs(B, H, W, seed=1 + rank, L_bg=28, L_fg=32)
    inp = common.to_dev(inp_cpu, dev)
    total_steps = Wm + K
    den = DualDiffDenoiser(unet, nets, guidance_scale=2.0, use_cuda_graph=True)

    def prepare():
        den.prepare(inp["latents"], inp["prompt_embeds"], inp["camera_param"], [inp["boxes_bg"], inp["boxes_fg"]],
                    [inp["cond_bg"], inp["cond_fg"]], num_inference_steps=max(total_steps, 4))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timed region: K sampler steps, latents already in HBM -----------------------
    prepare()
    for i in range(Wm):
        den.step(i)
    sampler = ClockSampler(local)
    barrier()
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(Wm, Wm + K):
        den.step(i)
    e1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms = e0.elapsed_time(e1)
    launches = den.launches_per_step * K
    final_lat = den.latents.float().cpu()
    assert torch.isfinite(final_lat).all(), "non-finite latents after the timed steps"

    # ---- end-to-end through the public API with HOST buffers: per step H2D(latents) + step + D2H(latents)
    host_lat = inp_cpu["latents"].reshape(B * 6, 4, H, W).float().contiguous().pin_memory()
    host_out = torch.empty_like(host_lat).pin_memory()
    prepare()
    for i in range(Wm):
        den.latents.copy_(host_lat, non_blocking=True)
        den.step(i)
        host_out.copy_(den.latents, non_blocking=True)
    barrier()
    t0 = time.perf_counter()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for i in range(Wm, Wm + K):
        den.latents.copy_(host_lat if i == Wm else host_out, non_blocking=True)   # step input from pinned host memory
        den.step(i)
        host_out.copy_(den.latents, non_blocking=True)                            # step result back to the host
        torch.cuda.current_stream().synchronize()                                 # the host really has it
    f1.record()
    barrier()
    ms_e2e = f0.elapsed_time(f1)
    wall_e2e = (time.perf_counter() - t0) * 1e3
    ms_e2e = max(ms_e2e, wall_e2e)

    if world > 1:
        tt = torch.tensor([ms, ms_e2e], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms, ms_e2e = float(tt[0]), float(tt[1])

    # ---- roofline pass (rank 0): per-launch CUDA events on the launching stream, eager replay of one step
    roof, breakdown = None, None
    if rank == 0:
        den2 = DualDiffDenoiser(unet, nets, guidance_scale=2.0, use_cuda_graph=False)
        den2.parallel_branches = False   # serial launch order: clean per-kernel device times
        den2.prepare(inp["latents"], inp["prompt_embeds"], inp["camera_param"], [inp["boxes_bg"], inp["boxes_fg"]],
                     [inp["cond_bg"], inp["cond_fg"]], num_inference_steps=4)
        for _ in range(2):
            den2.step(0)
        torch.cuda.synchronize()
        # give the host a head start so the per-launch event intervals are pure device time (kernels back to back)
        torch.cuda._sleep(int(0.08 * 1.9e9))
        ops.profile_start()
        den2.step(1)
        rec = ops.profile_stop()
        if os.environ.get("DD_BENCH_SHAPES"):
            agg = {}
            for name, t_ms, fl, by, tag in rec:
                a = agg.setdefault((name, tag), [0.0, 0, 0.0])
                a[0] += t_ms; a[1] += 1; a[2] += fl
            with open(os.environ["DD_BENCH_SHAPES"], "w") as fh:
                for (name, tag), (t_ms, cnt, fl) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
                    fh.write(f"{name:16s} {tag:28s} x{cnt:3d} {t_ms:8.3f} ms  {fl / max(t_ms, 1e-9) / 1e9:8.1f} TFLOP/s\n")
        fam = {}
        for name, t_ms, fl, by, tag in rec:
            f = fam.setdefault(name, dict(ms=0.0, flops=0.0, bytes=0.0, launches=0))
            f["ms"] += t_ms; f["flops"] += fl; f["bytes"] += by; f["launches"] += 1
        tot = sum(f["ms"] for f in fam.values())
        pk = peaks()
        top = max(fam.items(), key=lambda kv: kv[1]["ms"])
        g = fam.get("gemm_tcgen05", top[1])
        ach = g["flops"] / (g["ms"] * 1e-3) / 1e12
        roof = {"kernel": "gemm_tcgen05_kernel (Linear / 1x1 / implicit-GEMM 3x3 conv)", "bound": "tensor",
                "achieved": round(ach, 1), "peak": pk["tf_sustained"], "unit": "TFLOP/s",
                "frac": round(ach / pk["tf_sustained"], 4), "peak_source": f"{pk['src']} (bf16 sustained)",
                "traffic": ncu_traffic("gemm_tcgen05_kernel"), "traffic_unit": "bytes per launch (ncu dram read+write, profiles/r01_launches_traffic.json)",
                "launches_per_step": g["launches"], "avg_launch_ms": round(g["ms"] / g["launches"], 4),
                "share_of_step": round(g["ms"] / tot, 3),
                "flops_per_step": g["flops"], "note": "algorithmic 2*M*N*K per launch summed over one step / summed CUDA-event durations"}
        breakdown = {k: {"ms": round(v["ms"], 3), "share": round(v["ms"] / tot, 3), "launches": v["launches"],
                         "tflops": round(v["flops"] / (v["ms"] * 1e-3) / 1e12, 1) if v["flops"] else None,
                         "gbs": round(v["bytes"] / (v["ms"] * 1e-3) / 1e9, 1) if v["bytes"] and not v["flops"] else None}
                     for k, v in sorted(fam.items(), key=lambda kv: -kv[1]["ms"])}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    value = world * B * K / (ms * 1e-3)
    pk = peaks()
    out = {
        "metric": METRIC, "value": round(value, 3), "unit": UNIT, "n_gpus": world, "steps": K, "warmup": Wm,
        "ms_per_step": round(ms / K, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": {"workload": f"BASELINE.json configs[1]: UniPC+CFG sampling steps, batch {B} six-view 224x400 scenes per GPU "
                               f"(latent 28x50, n={12 * B} images/step), bf16 kernels, random-init SDv1.5-shaped weights",
                   "scenes_per_gpu": B, "cfg": True, "guidance_scale": 2.0, "scheduler": "UniPC(bh2, order 2)",
                   "l2": "inputs larger than L2 (3.3 GB bf16 weights + >1 GB activations per step vs 126 MB L2)",
                   "cuda_graph": True, "parallelism": f"scene-sharded x{world}, no collectives"},
        "tflops_per_scene_step": TFLOP_PER_SCENE_STEP_CFG,
        "model_tflops": round(value * TFLOP_PER_SCENE_STEP_CFG, 1),
        "model_frac_of_peak": round(value * TFLOP_PER_SCENE_STEP_CFG / world / pk["tf_sustained"], 4),
        "e2e": {"value": round(world * B * K / (ms_e2e * 1e-3), 3), "unit": UNIT,
                "h2d_bytes_per_step": host_lat.numel() * 4, "d2h_bytes_per_step": host_out.numel() * 4,
                "ms_per_step": round(ms_e2e / K, 3)},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roof,
        "kernel_breakdown": breakdown,
    }
    if world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline(sds, budget_s=args.cpu_budget)
    emit(out)
    if world > 1:
        dist.destroy_process_group()


# -------------------------------------------------------------------------------------------------------
# CPU arm: the oracle port (the reference is pure Python on un-vendored diffusers/xformers and cannot travel to
# the GPU box; oracle/dualdiff_oracle.py restates it and is pinned to the reference's own classes by the goldens)
# -------------------------------------------------------------------------------------------------------
def cpu_step_time(sds, scenes=1, cfg=True, repeats=1):
    import torch
    from dualdiff_b200 import synthetic as S
    from oracle import dualdiff_oracle as O
    torch.set_num_threads(os.cpu_count())
    inp = S.make_inputs(scenes, H, W, seed=1, L_bg=28, L_fg=32)
    sch = O.UniPC()
    sch.set_timesteps(25)
    lat = inp["latents"]
    times = []
    with torch.no_grad():
        for r in range(repeats):
            t0 = time.perf_counter()
            lat, _ = O.denoise_step(sds["unet"], sds["bg"], sds["fg"], sch, lat, int(sch.timesteps[r]), inp, 2.0, cfg)
            times.append(time.perf_counter() - t0)
    return times


def cpu_baseline(sds, budget_s=30.0):
    import torch
    t = cpu_step_time(sds, scenes=1, cfg=True, repeats=1)[0]
    return {"value": round(1.0 / t, 5), "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
            "sample": f"1 CFG scene-step (B=1: 12 images, 5.82 TFLOP fp32) of the same workload with the oracle port "
                      f"(oracle/dualdiff_oracle.py, torch {torch.__version__} CPU, {torch.get_num_threads()} threads): {t:.1f} s",
            "seconds": round(t, 2)}


def run_reference(args):
    rank, world, _ = dist_env()
    if rank != 0:
        return
    import torch
    import common
    from dualdiff_b200 import synthetic as S
    sds = {}
    unet, nets, sds = common.build_models()
    del unet, nets
    K, Wm = args.steps, args.warmup
    # bounded sample per step: ONE six-view scene with CFG (12 images).  Probe once, then cap the number of steps so
    # the arm ends within a few minutes on this host; the line reports the steps actually timed.
    probe = cpu_step_time(sds, 1, True, 1)[0]
    budget = float(args.cpu_budget_total)
    k_run = max(1, min(K, int(budget / max(probe, 1e-3))))
    w_run = 0 if probe * (k_run + 1) > budget else min(Wm, 1)
    times = cpu_step_time(sds, 1, True, w_run + k_run)[w_run:]
    total = sum(times)
    value = k_run / total
    cores = os.cpu_count()
    sample = (f"each step = 1 six-view 224x400 scene with CFG (12 images, 5.82 TFLOP fp32) through the oracle port of the "
              f"reference's CPU path; {k_run} of the requested {K} steps timed ({w_run} warm-up) to stay within {budget:.0f} s")
    out = {"impl": "reference", "metric": METRIC, "value": round(value, 5), "unit": UNIT, "n_gpus": args.gpus,
           "steps": k_run, "warmup": w_run, "ms_per_step": round(1e3 * total / k_run, 1), "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": "BASELINE.json configs[0]/[1] sampled: UniPC+CFG sampling step of one six-view 224x400 scene "
                                  "per step on the host CPU (fp32)", "scenes_per_step": 1, "cfg": True},
           "cpu_baseline": {"value": round(value, 5), "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
           "e2e": {"value": round(value, 5), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    emit(out)


_REAL_STDOUT = None


def protect_stdout():
    """the contract is ONE JSON line on stdout: route everything else written to fd 1 (NCCL's version banner,
    library chatter) to stderr and keep the real stdout for the final line"""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)


def emit(obj):
    out = _REAL_STDOUT if _REAL_STDOUT is not None else sys.stdout
    out.write(json.dumps(obj) + "\n")
    out.flush()


def main():
    protect_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--scenes", type=int, default=8, help="six-view scenes per GPU per step (BASELINE configs[1]: 8)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-budget", type=float, default=30.0)
    ap.add_argument("--cpu-budget-total", type=float, default=150.0)
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
