#!/usr/bin/env python
"""bench.py — DualDiff denoising-step throughput on B200 (contract in the task statement / DESIGN.md §Measurement).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--scenes B] [--impl ours|reference]
                    [--workload scenes|viewshard|frameshard] [--total-scenes S] [--no-extra]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" = one loop body of the sampler (pipeline_bev_controlnet.py:381-504) over a batch of B six-view
224x400 scenes per GPU: ControlNet-bg + ControlNet-fg + multi-view UNet + CFG + UniPC update, bf16 kernels.
Workload at N=1 = BASELINE.json configs[1] (B=8 scenes, CFG, UniPC) — `value` is scene-steps/s over all ranks
(weak scaling: B scenes per GPU, scenes are independent, no collectives; `--total-scenes S` fixes the job instead:
strong scaling, BASELINE configs[2] as written with S = 64).  Prints ONE JSON line on rank 0.

The two configurations that need a collective are measured by the same code: `--workload viewshard` (configs[3]: 448x800,
camera views split over the ranks of a scene group, neighbour-view exchange over NCCL) and `--workload frameshard`
(configs[4]: 16-frame x 6-view clips, frames split over the ranks, temporal attention between two all-to-alls).  The default run appends a
short measurement of both under `extra_workloads` (skip with --no-extra), so the driver's 1/2/4/8 scaling runs record them.
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
    for name in ("r02_launches_traffic.json", "r01_launches_traffic.json"):
        try:
            return round(json.load(open(os.path.join(ROOT, "profiles", name)))[kernel]["traffic_bytes_per_launch"])
        except Exception:
            continue
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
WORKLOADS = {
    "scenes": "BASELINE.json configs[1]: UniPC+CFG sampling steps, batch {B} six-view 224x400 scenes per GPU (latent 28x50, "
              "n={n} images/step), bf16 kernels, random-init SDv1.5-shaped weights",
    "viewshard": "BASELINE.json configs[3]: UniPC+CFG sampling steps of six-view 448x800 scenes (latent 56x100), camera views "
                 "split over the {R} rank(s) of a scene group, {B} scene(s) per group, n={n} images/step per GPU, neighbour-view "
                 "exchange for the cross-view attention over NCCL/NVLink",
    "frameshard": "BASELINE.json configs[4]: UniPC+CFG sampling steps of {clips} video clip(s) of 16 frames x 6 views (224x400), "
                  "frames split over the ranks ({fl} frames of every clip per GPU, n={n} images/step per GPU), temporal "
                  "attention token-sharded between two all-to-alls over NCCL/NVLink",
}


class Workload:
    """one configuration of the hot path on this rank: the denoiser, how to (re)load its inputs, and how many six-view
    scenes the whole job advances per step"""

    def __init__(self, name, args, rank, world, dev, models):
        import torch
        import common
        from dualdiff_b200 import synthetic as S
        from dualdiff_b200.pipeline import DualDiffDenoiser
        from dualdiff_b200.sharding import FrameShard, ViewShard, shard_scenes
        self.name, self.world = name, world
        unet, nets = models["unet"], models["nets"]
        self.parallelism, self.scaling, kw, prep_kw = f"scene-sharded x{world}, no collectives", "weak", {}, {}
        if name == "scenes":
            self.h, self.w = H, W
            if args.total_scenes:
                mine = shard_scenes(args.total_scenes, rank, world)
                B, self.scenes_per_step, self.scaling = len(mine), args.total_scenes, "strong"
                if B == 0:
                    raise SystemExit(f"--total-scenes {args.total_scenes} leaves rank {rank} of {world} without a scene")
            else:
                B, self.scenes_per_step = args.scenes, world * args.scenes
            seed = 1 + rank
            self.desc = WORKLOADS[name].format(B=B, n=12 * B)
        elif name == "viewshard":
            self.h, self.w = 2 * H, 2 * W
            vs = ViewShard(rank, world) if world > 1 else None
            R = vs.ranks_per_scene if vs is not None else 1
            B = R * args.hd_scenes                     # scenes of this rank's group; every rank holds B * 6 / R views (x2 CFG)
            self.scenes_per_step = world * args.hd_scenes
            seed = 101 + (vs.group_index if vs is not None else 0)
            kw, prep_kw = dict(view_shard=vs), (dict(scenes_sliced=True) if vs is not None else {})
            self.parallelism = (f"{world // R} group(s) x {R} rank(s) per scene: (scene, view) units, LayerNorm rows of the 2 halo "
                                f"views exchanged per cross-view block by NCCL send/recv on a side stream") if vs is not None else \
                "one GPU: all six views local, no exchange"
            self.desc = WORKLOADS[name].format(B=B, R=R, n=12 * args.hd_scenes)
        elif name == "frameshard":
            self.h, self.w = H, W
            F = 16
            if F % world:
                raise SystemExit(f"frameshard: world size {world} must divide {F} frames")
            unet = models["video_unet"]()
            unet.frame_shard = FrameShard(rank, world, F) if world > 1 else None
            clips, fl = world * args.clips, F // world
            B = clips * fl                             # (clip, frame) pairs of this rank, frame-minor
            self.scenes_per_step = clips * F
            seed = 201 + rank
            self.parallelism = f"frames of every clip split over {world} rank(s); per temporal block the LayerNorm rows go all-to-all to " \
                f"token shards (all frames, 1/{world} of the tokens), attention output all-to-all back" \
                if world > 1 else "one GPU: all 16 frames local, no exchange"
            self.desc = WORKLOADS[name].format(clips=clips, fl=fl, n=12 * B)
        else:
            raise SystemExit(f"unknown workload {name}")
        self.B = B
        self.inp_cpu = S.make_inputs(B, self.h, self.w, seed=seed, L_bg=28, L_fg=32)
        self.inp = common.to_dev(self.inp_cpu, dev)
        self.den = DualDiffDenoiser(unet, nets, guidance_scale=2.0, use_cuda_graph=True, **kw)
        if args.serial_branches:
            self.den.parallel_branches = False
        self.unet, self.nets, self.prep_kw = unet, nets, prep_kw

    def prepare(self, den=None, steps=4):
        i = self.inp
        (den or self.den).prepare(i["latents"], i["prompt_embeds"], i["camera_param"], [i["boxes_bg"], i["boxes_fg"]],
                                  [i["cond_bg"], i["cond_fg"]], num_inference_steps=max(steps, 4),
                                  **(self.prep_kw if den is None else {}))

    def host_latents(self):
        """this rank's latents as the denoiser keeps them ((scene, local view) images, fp32 NCHW), in pinned host memory"""
        return self.den.latents.detach().float().cpu().contiguous().pin_memory()


def measure(wl, K, Wm, barrier, rank, local, sample_clocks=True):
    """device-resident and end-to-end timing of K sampler steps of one workload (W warm-up steps each); returns a dict"""
    import torch
    den = wl.den
    # ---- device-resident timed region: K sampler steps, latents already in HBM -----------------------
    wl.prepare(steps=Wm + K)
    for i in range(Wm):
        den.step(i)
    sampler = ClockSampler(local)
    barrier()
    if rank == 0 and sample_clocks:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(Wm, Wm + K):
        den.step(i)
    e1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 and sample_clocks else None
    ms = e0.elapsed_time(e1)
    final_lat = den.latents.float().cpu()
    assert torch.isfinite(final_lat).all(), "non-finite latents after the timed steps"

    # ---- end-to-end through the public API with HOST buffers: per step H2D(latents) + step + D2H(latents)
    wl.prepare(steps=Wm + K)
    host_lat = wl.host_latents()
    host_out = torch.empty_like(host_lat).pin_memory()
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
    ms_e2e = max(f0.elapsed_time(f1), (time.perf_counter() - t0) * 1e3)
    return dict(ms=ms, ms_e2e=ms_e2e, clocks=clocks, launches=den.launches_per_step * K if den._graph is not None else None,
                h2d=host_lat.numel() * 4, d2h=host_out.numel() * 4, cuda_graph=den._graph is not None, graph_note=den.graph_note)


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
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    K, Wm = args.steps, args.warmup
    unet, nets, sds = common.build_models()
    for m in [unet] + nets:
        m.pack(dev)

    def video_unet():
        """the multi-view UNet with temporal blocks (16-frame clips): same seeded weights plus the temporal projections"""
        from dualdiff_b200.networks import UNet2DConditionModelMultiview
        with torch.device("meta"):
            vu = UNet2DConditionModelMultiview(cross_attention_dim=768, neighboring_view_pair=common.NEIGHBORS, temporal_frames=16)
        man = S.manifest_of(vu)
        sd = {k: (sds["unet"][k] if k in sds["unet"] else S.init_tensor(k, tuple(shape), common.SEEDS["unet"]))
              for k, shape in man.items()}
        vu.load_state_dict(sd, strict=True, assign=True)
        vu.eval()
        vu.pack(dev)
        return vu

    models = dict(unet=unet, nets=nets, video_unet=video_unet)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_max(*vals):
        if world == 1:
            return vals
        tt = torch.tensor(list(vals), device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return tuple(float(x) for x in tt)

    wl = Workload(args.workload, args, rank, world, dev, models)
    m = measure(wl, K, Wm, barrier, rank, local)
    ms, ms_e2e = reduce_max(m["ms"], m["ms_e2e"])
    launches = m["launches"]
    if launches is None:      # eager launches (a refused capture): count one step
        n0 = _lib.lib().dd_launch_count()
        wl.den.step(0)
        torch.cuda.synchronize()
        launches = (_lib.lib().dd_launch_count() - n0) * K

    # ---- roofline pass (rank 0): per-launch CUDA events on the launching stream, eager SERIAL replay of one step of the
    #      main workload (not the timed graph: the graph overlaps the three branches on three streams)
    roof, breakdown, sub = None, None, None
    if rank == 0 and wl.den.view_shard is None and getattr(wl.unet, "frame_shard", None) is None:
        den2 = DualDiffDenoiser(wl.unet, wl.nets, guidance_scale=2.0, use_cuda_graph=False)
        den2.parallel_branches = False   # serial launch order: clean per-kernel device times
        wl.prepare(den2)
        for _ in range(2):
            den2.step(0)
        torch.cuda.synchronize()
        # sub-metrics: CUDA events around the three sub-networks and the CFG + scheduler kernel
        den2.phase_events = []
        torch.cuda._sleep(int(0.08 * 1.9e9))
        den2.step(1)
        torch.cuda.synchronize()
        ev = den2.phase_events
        den2.phase_events = None
        names = ("ms_controlnet_bg", "ms_controlnet_fg", "ms_unet", "ms_cfg_sched")
        sub = {n: round(ev[i].elapsed_time(ev[i + 1]), 3) for i, n in enumerate(names)}
        sub["note"] = ("serial eager replay of one step (branches one after the other); the timed CUDA graph overlaps the two "
                       "ControlNet branches with the UNet encoder on three streams, so ms_per_step < the sum")
        # give the host a head start so the per-launch event intervals are pure device time (kernels back to back)
        torch.cuda._sleep(int(0.08 * 1.9e9))
        ops.profile_start()
        den2.step(2)
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
                "traffic": ncu_traffic("gemm_tcgen05_kernel"), "traffic_unit": "bytes per launch (ncu dram read+write, profiles/r02_launches_traffic.json)",
                "launches_per_step": g["launches"], "avg_launch_ms": round(g["ms"] / g["launches"], 4),
                "share_of_step": round(g["ms"] / tot, 3),
                "flops_per_step": g["flops"],
                "note": "algorithmic 2*M*N*K per launch summed over one step / summed CUDA-event durations, measured in a serial "
                        "EAGER replay of one step after the timed region (per-launch events cannot be recorded inside the timed graph)"}
        breakdown = {k: {"ms": round(v["ms"], 3), "share": round(v["ms"] / tot, 3), "launches": v["launches"],
                         "tflops": round(v["flops"] / (v["ms"] * 1e-3) / 1e12, 1) if v["flops"] else None,
                         "gbs": round(v["bytes"] / (v["ms"] * 1e-3) / 1e9, 1) if v["bytes"] and not v["flops"] else None}
                     for k, v in sorted(fam.items(), key=lambda kv: -kv[1]["ms"])}
        del den2

    # ---- the line of the main workload is complete here (rank 0); what follows can only add to it
    out = None
    if rank == 0:
        value = wl.scenes_per_step * K / (ms * 1e-3)
        pk = peaks()
        out = {
            "metric": METRIC if wl.h == H else METRIC.replace("224x400", f"{8 * wl.h}x{8 * wl.w}"),
            "value": round(value, 3), "unit": UNIT, "n_gpus": world, "steps": K, "warmup": Wm,
            "ms_per_step": round(ms / K, 3), "higher_is_better": True, "scaling": wl.scaling, "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": wl.desc, "scenes_per_gpu": wl.B if args.workload == "scenes" else None,
                       "scenes_per_step_all_gpus": wl.scenes_per_step, "cfg": True, "guidance_scale": 2.0,
                       "scheduler": "UniPC(bh2, order 2)",
                       "l2": "inputs larger than L2 (3.3 GB bf16 weights + >1 GB activations per step vs 126 MB L2)",
                       "cuda_graph": m["cuda_graph"], "parallelism": wl.parallelism,
                       "branch_streams": 1 if args.serial_branches else 3},
            "e2e": {"value": round(wl.scenes_per_step * K / (ms_e2e * 1e-3), 3), "unit": UNIT,
                    "h2d_bytes_per_step": m["h2d"], "d2h_bytes_per_step": m["d2h"],
                    "ms_per_step": round(ms_e2e / K, 3)},
            "gpu_launches": int(launches),
            "clocks": m["clocks"],
            "roofline": roof,
            "kernel_breakdown": breakdown,
            "sub_metrics": sub,
        }
        if m["graph_note"]:
            out["config"]["graph_note"] = m["graph_note"]
        if args.warmup_requested != Wm:
            out["config"]["warmup_note"] = f"--warmup {args.warmup_requested} raised to {Wm} (minimum of the timing rules)"
        if args.workload == "scenes":
            out["tflops_per_scene_step"] = TFLOP_PER_SCENE_STEP_CFG
            out["model_tflops"] = round(value * TFLOP_PER_SCENE_STEP_CFG, 1)
            out["model_frac_of_peak"] = round(value * TFLOP_PER_SCENE_STEP_CFG / world / pk["tf_sustained"], 4)

    # A captured step with NCCL nodes keeps the communicator busy (destroy_process_group() waits for it): release it now.
    wl.den.release_graph()

    # Watchdog: the additions below (multi-rank extras, process-group teardown) must never lose the line above.  When it
    # fires, rank 0 prints the line as it stands and every rank leaves without further collectives.
    import threading
    state = {"done": False}

    def bail():
        if state["done"]:
            return
        state["done"] = True
        if out is not None:
            out.setdefault("extra_workloads", {})["watchdog"] = f"extras / teardown exceeded {args.extra_budget:.0f} s; line printed by the watchdog"
            emit(out)
        os._exit(0)

    dog = threading.Timer(args.extra_budget, bail)
    dog.daemon = True
    dog.start()

    # ---- the two configurations with a collective, measured briefly in the same run (all ranks take part)
    extra = None
    if args.workload == "scenes" and not args.no_extra and not args.total_scenes:
        extra = {}
        if out is not None:
            out["extra_workloads"] = extra
        wl.den = None
        torch.cuda.empty_cache()
        for name in ("viewshard", "frameshard"):
            try:
                if name == "frameshard" and 16 % world:
                    raise RuntimeError(f"world size {world} does not divide 16 frames")
                w2 = Workload(name, args, rank, world, dev, models)
                k2 = max(3, min(K, args.extra_steps))
                m2 = measure(w2, k2, 3, barrier, rank, local, sample_clocks=False)
                ms2, ms2e = reduce_max(m2["ms"], m2["ms_e2e"])
                extra[name] = {"value": round(w2.scenes_per_step * k2 / (ms2 * 1e-3), 3), "unit": UNIT, "n_gpus": world,
                               "steps": k2, "warmup": 3, "ms_per_step": round(ms2 / k2, 3), "scaling": w2.scaling,
                               "e2e": {"value": round(w2.scenes_per_step * k2 / (ms2e * 1e-3), 3), "unit": UNIT,
                                       "h2d_bytes_per_step": m2["h2d"], "d2h_bytes_per_step": m2["d2h"]},
                               "config": {"workload": w2.desc, "scenes_per_step_all_gpus": w2.scenes_per_step,
                                          "parallelism": w2.parallelism, "cuda_graph": m2["cuda_graph"]}}
                if m2["graph_note"]:
                    extra[name]["config"]["graph_note"] = m2["graph_note"]
                w2.den.release_graph()
                del w2
                torch.cuda.empty_cache()
            except Exception as e:     # an extra must never lose the main line
                extra[name] = {"unavailable": f"{type(e).__name__}: {e}"[:300]}
                if world > 1:
                    break              # the ranks may no longer be in step: no further collectives

    if rank == 0:
        if world == 1 and args.workload == "scenes" and not args.no_library_baseline:
            dog.cancel()               # single process: nothing below can wait for another rank
            out["library_gpu_baseline"] = library_gpu_baseline(sds, dev, wl.B, budget_s=args.library_budget)
        if world == 1 and not args.no_cpu_baseline and args.workload == "scenes":
            dog.cancel()
            out["cpu_baseline"] = cpu_baseline(sds, budget_s=args.cpu_budget)
        if not state["done"]:
            state["done"] = True
            emit(out)
    if world > 1:
        # the line is out; a teardown that hangs is cut short by a second timer
        t2 = threading.Timer(30.0, lambda: os._exit(0))
        t2.daemon = True
        t2.start()
        dist.destroy_process_group()
        t2.cancel()
    dog.cancel()


# -------------------------------------------------------------------------------------------------------
# "library GPU" line (SURVEY §8d last row): the SAME step through the oracle restatement, moved to the B200 in bf16 with
# torch's own kernels (cuDNN convolutions, cuBLAS GEMMs, F.scaled_dot_product_attention = flash attention) -- what the
# reference's PyTorch path would get from stock libraries on this box.  Test infrastructure, timed after the product
# numbers; never on the product path.
# -------------------------------------------------------------------------------------------------------
def library_gpu_baseline(sds, dev, scenes, budget_s=60.0):
    import torch
    import torch.nn.functional as F
    from dualdiff_b200 import synthetic as S
    import common
    try:
        from oracle import dualdiff_oracle as O
    except Exception as e:
        return {"unavailable": f"oracle not importable: {e}"}
    keep = {n: getattr(O, n) for n in ("mha", "_lin", "_conv", "_gn", "_ln")}
    try:
        def cast_first(fn):      # the oracle is written for one dtype: bring the input to the layer's (bf16) weight dtype
            def wrapped(sd, p, x, *a, **k):
                return fn(sd, p, x.to(sd[p + ".weight"].dtype), *a, **k)
            return wrapped
        for n in ("_lin", "_conv", "_gn", "_ln"):
            setattr(O, n, cast_first(keep[n]))

        def sdpa(q, k, v, heads):
            b, lq, c = q.shape
            d = c // heads
            qh, kh, vh = (t.reshape(b, t.shape[1], heads, d).transpose(1, 2) for t in (q, k, v))
            return F.scaled_dot_product_attention(qh, kh, vh).transpose(1, 2).reshape(b, lq, c)
        O.mha = sdpa
        bf = torch.bfloat16
        gsd = {n: {k: v.to(dev, bf) for k, v in sd.items()} for n, sd in sds.items()}
        with torch.no_grad(), torch.device(dev):
            sch = O.UniPC()
            sch.set_timesteps(25)
            times, used = [], scenes
            while used >= 1:
                try:
                    with torch.device("cpu"):
                        inp = S.make_inputs(used, H, W, seed=1, L_bg=28, L_fg=32)
                    inp = common.to_dev(inp, dev)
                    inp = {k: ({kk: (vv.to(bf) if vv.is_floating_point() else vv) for kk, vv in v.items()} if isinstance(v, dict)
                               else v.to(bf)) for k, v in inp.items()}
                    lat = inp["latents"]
                    t_start = time.perf_counter()
                    for r in range(4):
                        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        e0.record()
                        out = O.noise_prediction(gsd["unet"], gsd["bg"], gsd["fg"], lat, int(sch.timesteps[r]), inp, 2.0, True)
                        e1.record()
                        torch.cuda.synchronize()
                        times.append(e0.elapsed_time(e1))
                        if time.perf_counter() - t_start > budget_s:
                            break
                    finite = bool(torch.isfinite(out["eps"].float()).all())
                    break
                except torch.cuda.OutOfMemoryError:
                    times, used = [], used // 2
                    torch.cuda.empty_cache()
        if not times:
            return {"unavailable": "out of memory at one scene"}
        t = min(times[1:]) if len(times) > 1 else times[0]
        return {"value": round(used / (t * 1e-3), 3), "unit": UNIT, "ms_per_step": round(t, 2), "scenes_per_step": used,
                "kind": "oracle restatement on the GPU: bf16, eager torch ops (cuDNN conv, cuBLAS GEMM, SDPA flash attention), "
                        "no CUDA graph, ControlNet K/V and condition embedding recomputed every step as the reference does; "
                        "noise prediction + CFG only (scheduler update excluded)",
                "finite": finite, "torch": torch.__version__, "note": "test infrastructure, timed on the same B200 after the product numbers; "
                                                    "best of the steps after the first"}
    except Exception as e:
        return {"unavailable": f"{type(e).__name__}: {e}"[:300]}
    finally:
        for n, fn in keep.items():
            setattr(O, n, fn)
        torch.cuda.empty_cache()


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
    ap.add_argument("--workload", default="scenes", choices=["scenes", "viewshard", "frameshard"])
    ap.add_argument("--total-scenes", type=int, default=0, help="strong scaling: this many scenes per step over ALL ranks "
                    "(BASELINE configs[2]: 64) instead of --scenes per GPU")
    ap.add_argument("--hd-scenes", type=int, default=2, help="viewshard: 448x800 scenes per step per GPU-equivalent of work")
    ap.add_argument("--clips", type=int, default=1, help="frameshard: 16-frame clips per step per GPU-equivalent of work")
    ap.add_argument("--no-extra", action="store_true", help="skip the short viewshard / frameshard measurements of the default run")
    ap.add_argument("--extra-steps", type=int, default=5)
    ap.add_argument("--extra-budget", type=float, default=240.0, help="seconds the extras + teardown may take before the "
                    "watchdog prints the main line and exits")
    ap.add_argument("--serial-branches", action="store_true", help="A/B: launch the two condition branches and the UNet encoder "
                    "on ONE stream instead of three (default: three streams inside the CUDA graph)")
    ap.add_argument("--no-library-baseline", action="store_true")
    ap.add_argument("--library-budget", type=float, default=45.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-budget", type=float, default=30.0)
    ap.add_argument("--cpu-budget-total", type=float, default=150.0)
    args = ap.parse_args()
    args.warmup_requested = args.warmup
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3      # timing rule: at least 3 warm-up steps; the line reports the value used and the one requested
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
