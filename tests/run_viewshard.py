"""torchrun entry: camera views sharded over WORLD_SIZE GPUs with the neighbour K/V exchange over NCCL (config 4)
must reproduce the single-GPU step.  Rank 0 prints 'VIEWSHARD OK ...' on success."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import common  # noqa: E402
from dualdiff_b200 import synthetic as S  # noqa: E402
from dualdiff_b200.pipeline import DualDiffDenoiser  # noqa: E402
from dualdiff_b200.sharding import ViewShard  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    h, w = (int(x) for x in os.environ.get("LATENT", "28x50").split("x"))
    vs = ViewShard(rank, world)
    B, steps = int(os.environ.get("SCENES", str(vs.n_groups))), 2      # at least one scene per group of ranks
    unet, nets, _ = common.build_models()
    for m in [unet] + nets:
        m.pack(dev)
    inp = common.to_dev(S.make_inputs(B, h, w, seed=1, L_bg=28, L_fg=32), dev)
    args = (inp["latents"], inp["prompt_embeds"], inp["camera_param"], [inp["boxes_bg"], inp["boxes_fg"]],
            [inp["cond_bg"], inp["cond_fg"]])
    den = DualDiffDenoiser(unet, nets, guidance_scale=2.0, view_shard=vs)
    den.prepare(*args, num_inference_steps=8)
    for i in range(steps):
        den.step(i)                                  # the two steps that are compared with the unsharded run
    shard_lat = den.latents.clone()
    torch.cuda.synchronize()
    dist.barrier()
    # timing: three further steps, after the warm-up above (NCCL connections, lazy kernel attributes)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps, steps + 3):
        den.step(i)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    den.latents.copy_(shard_lat)
    mine = vs.scenes(B)
    shard = den.latents.reshape(len(mine), vs.v_loc, 4, h, w)
    # reference: the same two steps unsharded on this rank's GPU
    ref = DualDiffDenoiser(unet, nets, guidance_scale=2.0, use_cuda_graph=False)
    ref.prepare(*args, num_inference_steps=8)
    for i in range(steps):
        ref.step(i)
    full = ref.latents.reshape(B, 6, 4, h, w)[mine.start:mine.stop][:, vs.views].clone()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for i in range(steps, steps + 3):
        ref.step(i)                                  # the unsharded step on ONE GPU (eager launches), for scale
    f1.record()
    torch.cuda.synchronize()
    ms_one = f0.elapsed_time(f1) / 3
    m = common.metrics(shard.float().cpu(), full.float().cpu())
    t = torch.tensor([m["rel_l2"], ms, ms_one], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        ok = float(t[0]) < 2e-2
        print(f"VIEWSHARD {'OK' if ok else 'FAIL'} world={world} ranks_per_scene={vs.ranks_per_scene} latent={h}x{w} scenes={B} "
              f"cuda_graph={den._graph is not None} rel_l2(max over ranks)={float(t[0]):.3e} "
              f"ms/step(max over ranks)={float(t[1]):.2f} all scenes unsharded on one GPU (eager)={float(t[2]):.2f}", flush=True)
        if den.graph_note:
            print(den.graph_note, flush=True)
    # a captured graph with NCCL send / recv nodes keeps the communicator busy: release it before the group is destroyed
    # (round-2 finding: destroy_process_group() never returned while the graph was alive)
    den.release_graph()
    del den, ref
    torch.cuda.synchronize()
    common_shutdown()


def common_shutdown():
    import threading
    t = threading.Timer(30.0, lambda: os._exit(0))     # teardown must not be able to hang the test
    t.daemon = True
    t.start()
    dist.destroy_process_group()
    t.cancel()


if __name__ == "__main__":
    main()
