"""Multi-GPU partitioning of the denoising path.

DualDiff's step shards by scene with NO data-path collective: cross-view attention regroups '(b n)' per scene
(networks/blocks.py:196-197), so a scene never reads another scene's activations.  This mirrors how the reference
runs multi-GPU inference (`accelerator.prepare(val_dataloader)`, perception/data_prepare/val_set_gen.py:121;
manual `shard_volumn`, tools/downstream_v3_batched.py:120).  One process per GPU; the process group is used only
for the barrier and the max-over-ranks timing reduction.
"""
import torch


def shard_scenes(total_scenes: int, rank: int, world_size: int) -> range:
    """contiguous, balanced partition of scene indices [0, total_scenes)"""
    base, rem = divmod(total_scenes, world_size)
    start = rank * base + min(rank, rem)
    return range(start, start + base + (1 if rank < rem else 0))


def reduce_max_ms(ms: float, device=None) -> float:
    """max over ranks of a device-measured duration (every multi-GPU number is timed as the slowest rank)"""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return ms
    t = torch.tensor([ms], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])


def slice_scenes(obj, scenes: range, total_scenes: int, n_cam: int = 6):
    """this rank's scenes of a batched pipeline input: tensors whose leading dimension is `total_scenes` (or
    `total_scenes * n_cam`, the '(b n) ...' layout of the ORS tensor) are cut, lists of per-scene prompts are cut, dicts /
    lists / tuples are walked, everything else is passed through"""
    if torch.is_tensor(obj):
        if obj.dim() > 0 and obj.shape[0] == total_scenes:
            return obj[scenes.start:scenes.stop]
        if obj.dim() > 0 and obj.shape[0] == total_scenes * n_cam:
            return obj[scenes.start * n_cam:scenes.stop * n_cam]
        return obj
    if isinstance(obj, dict):
        return {k: slice_scenes(v, scenes, total_scenes, n_cam) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        if len(obj) == total_scenes and all(isinstance(x, (str, torch.Generator)) for x in obj):
            return type(obj)(obj[scenes.start:scenes.stop])
        return type(obj)(slice_scenes(v, scenes, total_scenes, n_cam) for v in obj)
    return obj


def run_scene_sharded(pipe, *, prompt, image, camera_param, bev_controlnet_kwargs, gather: bool = True, n_cam: int = 6,
                      **call_kwargs):
    """BASELINE config 3 at the pipeline level: every rank runs `pipe(...)` on its contiguous share of the scenes
    (weights replicated, no data-path collective); with `gather`, rank 0 receives the per-scene results of all ranks in
    scene order (one gather of the finished outputs, the only communication) and the other ranks get None.
    Mirrors how the reference shards generation (`accelerator.prepare(val_dataloader)`, val_set_gen.py:121)."""
    import torch.distributed as dist
    ready = dist.is_available() and dist.is_initialized()
    rank, world = (dist.get_rank(), dist.get_world_size()) if ready else (0, 1)
    total = len(prompt)
    mine = shard_scenes(total, rank, world)
    local = None
    if len(mine) > 0:
        out = pipe(prompt=slice_scenes(list(prompt), mine, total, n_cam), image=slice_scenes(image, mine, total, n_cam),
                   camera_param=slice_scenes(camera_param, mine, total, n_cam),
                   bev_controlnet_kwargs=slice_scenes(bev_controlnet_kwargs, mine, total, n_cam),
                   **slice_scenes(call_kwargs, mine, total, n_cam))
        local = out.images if hasattr(out, "images") else out[0]
        if torch.is_tensor(local):
            local = local.cpu()
    if not gather or world == 1:
        return local
    parts = [None] * world if rank == 0 else None
    dist.gather_object(local, parts, dst=0)
    if rank != 0:
        return None
    parts = [p for p in parts if p is not None]
    if not parts:
        return None
    if torch.is_tensor(parts[0]):
        return torch.cat(parts)
    if hasattr(parts[0], "shape"):                     # numpy images
        import numpy as np
        return np.concatenate(parts)
    return [scene for part in parts for scene in part]  # PIL: [scene][view]


def default_ranks_per_scene(world: int, n_cam: int = 6) -> int:
    """largest divisor of the camera count that also divides the world size: 1 -> 1, 2 / 4 / 8 -> 2, 3 -> 3, 6 / 12 -> 6"""
    for r in range(min(world, n_cam), 0, -1):
        if n_cam % r == 0 and world % r == 0:
            return r
    return 1


class ViewShard:
    """Camera-view sharding inside a scene (BASELINE config 4), for ANY world size.

    The unit of work is (scene, camera view), both CFG halves of a view staying together (so the CFG combine and the
    scheduler update need no communication).  The `world` ranks form world / R GROUPS of R ranks (R = `ranks_per_scene`, a
    divisor of the camera count: 2 on 2 / 4 / 8 GPUs by default); the scenes of a call are dealt out to the groups
    (`shard_scenes`, no communication between groups -- scenes are independent, networks/blocks.py:196-197), and inside a
    group rank g owns the V_loc = n_cam / R consecutive views [g V_loc, (g+1) V_loc) of each of the group's scenes.
    8 GPUs therefore run 4 (or 8, 12 ...) scenes with 3 views x 2 CFG halves per scene on every GPU.

    Convolutions, norms, self- and text-attention and the ControlNet branches are per image -> unit-local.  The one
    exchange step of the path is the cross-view attention (networks/blocks.py:190-222): view v attends to views v-1 and v+1
    (ring, configs/dataset/Nuscenes.yaml:27-33), so per transformer block each rank sends the norm4 output rows of its FIRST
    view to the left rank of its group and of its LAST view to the right rank (grouped NCCL send/recv over NVLink) and
    receives the two halo views behind its own rows.  The rows exchanged are the LayerNorm output (C columns), not the
    projections (K + V = 2C columns): half the bytes, and the K/V projection of the two halo views is recomputed locally.
    `exchange_async` runs the transfer on a side stream so that it overlaps the Q/K/V projection of the local views.
    """

    def __init__(self, rank: int, world: int, n_cam: int = 6, ranks_per_scene: int = None, group=None):
        R = default_ranks_per_scene(world, n_cam) if ranks_per_scene is None else int(ranks_per_scene)
        if R < 1 or n_cam % R != 0 or world % R != 0:
            raise ValueError(f"ranks_per_scene={R} must divide both the number of camera views {n_cam} and the world size {world}")
        self.rank, self.world, self.n_cam, self.group = rank, world, n_cam, group
        self.ranks_per_scene = R
        self.n_groups = world // R
        self.group_index = rank // R
        self.group_rank = rank % R
        self.v_loc = n_cam // R
        base = self.group_index * R
        self.left = base + (self.group_rank - 1) % R
        self.right = base + (self.group_rank + 1) % R
        self._side = None

    @property
    def views(self):
        return list(range(self.group_rank * self.v_loc, (self.group_rank + 1) * self.v_loc))

    def scenes(self, total_scenes: int) -> range:
        """the scenes of a call this rank's group works on"""
        if total_scenes < self.n_groups:
            raise ValueError(f"{total_scenes} scene(s) cannot feed {self.n_groups} groups of {self.ranks_per_scene} rank(s): "
                             f"pass at least one scene per group (8 GPUs with 2 ranks per scene need 4 scenes)")
        return shard_scenes(total_scenes, self.group_index, self.n_groups)

    def kv_rows(self, n_outer: int) -> int:
        """images in the K/V buffer: local images followed by the left and right halo views of every scene"""
        return n_outer * self.v_loc + 2 * n_outer

    def kv_map(self, n_outer: int, device=None) -> torch.Tensor:
        """int32 [n_outer*V_loc, 2]: K/V image index of the (left, right) neighbour of every local image"""
        v, n_loc = self.v_loc, n_outer * self.v_loc
        rows = []
        for o in range(n_outer):
            for j in range(v):
                left = o * v + (j - 1) if j > 0 else n_loc + o                   # left halo block
                right = o * v + (j + 1) if j < v - 1 else n_loc + n_outer + o     # right halo block
                rows.append([left, right])
        return torch.tensor(rows, dtype=torch.int32, device=device)

    def exchange(self, buf: torch.Tensor, n_outer: int, T: int) -> None:
        """buf: [(n_loc + 2*n_outer) * T, W] rows (local images, then the halo tail); fills the halo tail in place."""
        import torch.distributed as dist
        v, W = self.v_loc, buf.shape[1]
        n_loc = n_outer * v
        local = buf[: n_loc * T].view(n_outer, v, T, W)
        send_first = local[:, 0].contiguous()        # my first view  -> right halo of the left rank
        send_last = local[:, v - 1].contiguous()     # my last view   -> left halo of the right rank
        halo_left = buf[n_loc * T:(n_loc + n_outer) * T].view(n_outer, T, W)
        halo_right = buf[(n_loc + n_outer) * T:].view(n_outer, T, W)
        if self.ranks_per_scene == 1:
            halo_left.copy_(send_last)
            halo_right.copy_(send_first)
            return
        # posting order matters when left == right (2 ranks per scene): the peer's sends arrive in the order (first view,
        # last view), which are my (right halo, left halo)
        ops = [dist.P2POp(dist.isend, send_first, self.left, self.group),
               dist.P2POp(dist.isend, send_last, self.right, self.group),
               dist.P2POp(dist.irecv, halo_right, self.right, self.group),
               dist.P2POp(dist.irecv, halo_left, self.left, self.group)]
        for r in dist.batch_isend_irecv(ops):
            r.wait()

    def exchange_async(self, buf: torch.Tensor, n_outer: int, T: int) -> None:
        """start `exchange` on a side stream ordered after the work already queued on the current stream (the LayerNorm that
        produced `buf`); the caller keeps launching kernels that do not read the halo tail and calls `exchange_wait()` before
        the first one that does.  CPU tensors (gloo tests) exchange synchronously."""
        if not buf.is_cuda:
            self.exchange(buf, n_outer, T)
            return
        main = torch.cuda.current_stream(buf.device)
        if self._side is None:
            self._side = torch.cuda.Stream(buf.device)
        self._side.wait_stream(main)
        with torch.cuda.stream(self._side):
            self.exchange(buf, n_outer, T)
        buf.record_stream(self._side)

    def exchange_wait(self, device=None) -> None:
        if self._side is not None:
            torch.cuda.current_stream(device).wait_stream(self._side)


class FrameShard:
    """Frame sharding of a video clip (BASELINE config 5: 16 frames x 6 views, frames split over the GPUs).

    Rank r owns the F_loc = n_frames / world consecutive frames [r*F_loc, (r+1)*F_loc) of every clip, with all six camera
    views of those frames: convolutions, norms, self-, text- and CROSS-VIEW attention stay rank-local.  The one exchange
    step is the temporal attention, which needs every frame of a (clip, view, token) position.  It runs TOKEN-sharded:
    an all-to-all (`to_token_shards`) re-partitions the LayerNorm rows from "my frames, all tokens" to "all frames, my
    share of the tokens"; Q/K/V projection and the attention over frames then run locally (the kernel addresses the received
    [rank][clip][F_loc][view] blocks in place through its rank strides), and a second all-to-all (`from_token_shards`) brings
    the attention output back to frame-sharded rows for the output projection and the residual.  Per block a rank sends
    (world-1)/world * 2C columns per row -- against the (world-1) * 2C' of the K/V all-gather this design replaces
    (`gather`, kept for comparison; measured 0.77 weak-scaling efficiency on 4 B200 because of its volume).
    The reference has no temporal block or frame sharding: this design is dualdiff_b200's own.
    """

    def __init__(self, rank: int, world: int, n_frames: int = 16, group=None):
        if n_frames % world != 0:
            raise ValueError(f"world size {world} must divide the number of frames {n_frames}")
        self.rank, self.world, self.n_frames, self.group = rank, world, n_frames, group
        self.f_loc = n_frames // world

    @property
    def frames(self):
        return list(range(self.rank * self.f_loc, (self.rank + 1) * self.f_loc))

    def token_range(self, T: int, rank: int = None) -> range:
        """the tokens of an image whose temporal attention `rank` (default: this rank) computes"""
        return shard_scenes(T, self.rank if rank is None else rank, self.world)

    def to_token_shards(self, rows: torch.Tensor, n_img: int, T: int) -> torch.Tensor:
        """rows [n_img*T, W] (this rank's frames, every token) -> [world * n_img * t_me, W]: for every source rank s (major)
        its n_img images restricted to MY token range (t_me tokens) -- all frames of my share of the tokens."""
        import torch.distributed as dist
        W = rows.shape[1]
        x = rows.view(n_img, T, W)
        ranges = [self.token_range(T, r) for r in range(self.world)]
        send = torch.cat([x[:, r.start:r.stop].reshape(-1, W) for r in ranges])      # packing pass: [dest rank][image][token]
        t_me = len(ranges[self.rank])
        recv = torch.empty((self.world * n_img * t_me, W), device=rows.device, dtype=rows.dtype)
        dist.all_to_all_single(recv, send, [n_img * t_me] * self.world, [n_img * len(r) for r in ranges], group=self.group)
        return recv

    def from_token_shards(self, rows_tok: torch.Tensor, n_img: int, T: int) -> torch.Tensor:
        """inverse of `to_token_shards` for the attention output: [world * n_img * t_me, W] -> [n_img*T, W]"""
        import torch.distributed as dist
        W = rows_tok.shape[1]
        ranges = [self.token_range(T, r) for r in range(self.world)]
        t_me = len(ranges[self.rank])
        recv = torch.empty((n_img * T, W), device=rows_tok.device, dtype=rows_tok.dtype)   # [source rank][image][its tokens]
        dist.all_to_all_single(recv, rows_tok.contiguous(), [n_img * len(r) for r in ranges], [n_img * t_me] * self.world,
                               group=self.group)
        out = torch.empty((n_img, T, W), device=rows_tok.device, dtype=rows_tok.dtype)
        off = 0
        for r in ranges:
            out[:, r.start:r.stop] = recv[off:off + n_img * len(r)].view(n_img, len(r), W)
            off += n_img * len(r)
        return out.view(n_img * T, W)

    def gather(self, rows: torch.Tensor) -> torch.Tensor:
        """rows: this rank's projection rows [n_loc_img*T, W] -> [world * n_loc_img*T, W], rank-major (the all-gather form of
        the exchange: simpler, (world-1) x the volume)"""
        import torch.distributed as dist
        if self.world == 1:
            return rows
        out = torch.empty((self.world * rows.shape[0], rows.shape[1]), device=rows.device, dtype=rows.dtype)
        dist.all_gather_into_tensor(out, rows.contiguous(), group=self.group)
        return out


def slice_frames(x: torch.Tensor, frames, n_frames: int, n_view: int = 6) -> torch.Tensor:
    """x: [(clip, frame, view), ...] image-major tensor -> the images of `frames` (same ordering)"""
    n_clip = x.shape[0] // (n_frames * n_view)
    idx = torch.as_tensor(list(frames))
    return x.reshape(n_clip, n_frames, n_view, *x.shape[1:])[:, idx].reshape(n_clip * len(idx) * n_view, *x.shape[1:]).contiguous()


def slice_views(inputs: dict, views, n_cam: int = 6, scenes: range = None) -> dict:
    """select the scenes (optional) and the camera views of one rank from full step inputs (layouts of synthetic.make_inputs
    / dataset/utils.py:390-445): per-view tensors are sliced, view-shared ones kept.  `prompt_embeds` (uncond rows first,
    then cond rows when it holds 2B rows) follows the scene selection."""
    if scenes is not None:
        B = inputs["latents"].shape[0]
        inputs = slice_scenes(dict(inputs), scenes, B, n_cam)
        pe = inputs.get("prompt_embeds")
        if pe is not None and pe.shape[0] == 2 * B:
            inputs["prompt_embeds"] = torch.cat([pe[scenes.start:scenes.stop], pe[B + scenes.start:B + scenes.stop]])
    idx = torch.as_tensor(list(views))
    out = dict(inputs)
    out["latents"] = inputs["latents"][:, idx].contiguous()
    out["camera_param"] = inputs["camera_param"][:, idx].contiguous()
    bb = inputs["boxes_bg"]
    out["boxes_bg"] = None if bb is None else {k: (v[:, idx].contiguous() if v.shape[1] == n_cam else v) for k, v in bb.items()}
    cb = inputs["cond_bg"]                                   # (B, 3, H, n_cam*W) panorama
    w = cb.shape[-1] // n_cam
    out["cond_bg"] = torch.cat([cb[..., v * w:(v + 1) * w] for v in views], dim=-1).contiguous()
    cf = inputs["cond_fg"]                                   # (B*n_cam, 320, h, w)
    B = cf.shape[0] // n_cam
    out["cond_fg"] = cf.reshape(B, n_cam, *cf.shape[1:])[:, idx].reshape(B * len(views), *cf.shape[1:]).contiguous()
    return out


def slice_step_inputs(vs: "ViewShard", latents, prompt_embeds, camera_param, bboxes_3d_data, images, scenes_sliced: bool = False):
    """what `DualDiffDenoiser.prepare` keeps on a view-sharded rank: the scenes of the rank's group (all of them when the
    caller already passed only those) and the rank's camera views of latents, camera parameters, box tokens and both
    condition inputs.  Prompt embeddings (uncond rows first) follow the scenes -- and the views when there is one prompt
    per view (use_aug_text); the view-shared map vectors only follow the scenes."""
    B_all, n_all = latents.shape[0], latents.shape[1]
    mine = range(B_all) if scenes_sliced else vs.scenes(B_all)
    full = dict(latents=latents, camera_param=camera_param, boxes_bg=bboxes_3d_data[0], cond_bg=images[0],
                cond_fg=images[1], prompt_embeds=prompt_embeds)
    pe = None
    if prompt_embeds.shape[0] % (B_all * n_all) == 0 and n_all > 1:   # one prompt per view: keep this rank's scenes AND views
        pe = prompt_embeds.reshape(-1, B_all, n_all, *prompt_embeds.shape[1:])[:, mine.start:mine.stop][:, :, vs.views]
        pe = pe.reshape(-1, *pe.shape[3:]).contiguous()
        full["prompt_embeds"] = None
    loc = slice_views(full, vs.views, n_all, scenes=mine)
    fg_boxes = bboxes_3d_data[1]
    fg_boxes = None if fg_boxes is None else slice_scenes(fg_boxes, mine, B_all, vs.n_cam)
    return (loc["latents"], pe if pe is not None else loc["prompt_embeds"], loc["camera_param"], [loc["boxes_bg"], fg_boxes],
            [loc["cond_bg"], loc["cond_fg"]])


def gathered_kv_image(clip: int, frame: int, view: int, n_clip: int, f_loc: int, n_view: int = 6) -> int:
    """image index of (clip, global frame, view) inside FrameShard.gather's rank-major buffer -- the address arithmetic
    of csrc/dd_temporal.cu (kv_rank_stride = n_clip * f_loc * n_view images)"""
    rank, fl = divmod(frame, f_loc)
    return rank * (n_clip * f_loc * n_view) + (clip * f_loc + fl) * n_view + view
