"""N>1 path on CPU: scenes are sharded across ranks with no data-path collective; only the timing
reduction (max over ranks) and a barrier use the process group.  world_size 2, gloo."""
import os
import socket
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dualdiff_b200.sharding import shard_scenes, reduce_max_ms  # noqa: E402


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, total_scenes, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = shard_scenes(total_scenes, rank, world)
    dist.barrier()
    ms = reduce_max_ms(10.0 + 5.0 * rank)      # rank 1 is slower
    gathered = [None] * world
    dist.all_gather_object(gathered, list(mine))
    q.put((rank, list(mine), ms, gathered))
    dist.destroy_process_group()


def test_scene_sharding_two_ranks_gloo():
    world, total = 2, 11
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    shards = [r[1] for r in res]
    assert sorted(shards[0] + shards[1]) == list(range(total))            # a partition of the scenes
    assert abs(len(shards[0]) - len(shards[1])) <= 1                      # balanced
    assert all(abs(r[2] - 15.0) < 1e-6 for r in res)                      # time = max over ranks


def test_shard_scenes_properties():
    for total in (0, 1, 7, 64):
        for world in (1, 2, 3, 8):
            parts = [list(shard_scenes(total, r, world)) for r in range(world)]
            assert sorted(sum(parts, [])) == list(range(total))
            assert max(map(len, parts)) - min(map(len, parts)) <= 1


class _StubPipe:
    """records what it was called with and returns one 'image' per scene that encodes the scene's inputs"""

    def __call__(self, prompt, image, camera_param, bev_controlnet_kwargs, generator=None, **kw):
        from types import SimpleNamespace
        b = len(prompt)
        assert camera_param.shape[0] == b and image[0].shape[0] == b and image[1].shape[0] == 6 * b
        assert bev_controlnet_kwargs["bboxes_3d_data"][0]["bboxes"].shape[0] == b and len(generator) == b
        assert kw == {"height": 64, "num_inference_steps": 3}
        ids = torch.tensor([float(p.split()[-1]) for p in prompt])
        assert torch.equal(camera_param[:, 0, 0, 0], ids) and torch.equal(image[1][::6, 0, 0, 0], ids)
        return SimpleNamespace(images=ids[:, None].repeat(1, 6))


def _pipe_worker(rank, world, port, total, q):
    from dualdiff_b200.sharding import run_scene_sharded
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ids = torch.arange(total, dtype=torch.float32)
    out = run_scene_sharded(
        _StubPipe(), prompt=[f"scene {i}" for i in range(total)],
        image=[ids.view(-1, 1, 1, 1).expand(total, 3, 2, 12), ids.repeat_interleave(6).view(-1, 1, 1, 1).expand(6 * total, 320, 1, 2)],
        camera_param=ids.view(-1, 1, 1, 1).expand(total, 6, 3, 7),
        bev_controlnet_kwargs={"bboxes_3d_data": [{"bboxes": torch.zeros(total, 6, 2, 8, 3)}, {"bboxes": torch.zeros(total, 1, 2, 8, 3)}],
                               "use_aug_text": False},
        generator=[torch.Generator().manual_seed(i) for i in range(total)], height=64, num_inference_steps=3)
    q.put((rank, None if out is None else out.tolist()))
    dist.destroy_process_group()


def test_pipeline_scene_sharding_two_ranks_gloo():
    """config 3 at the pipeline level: each rank runs the pipeline on its scenes, rank 0 gathers the results in scene order"""
    world, total = 2, 5
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_pipe_worker, args=(r, world, port, total, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[1] is None and res[0] == [[float(i)] * 6 for i in range(total)]


# ---------------------------------------------------------------------------------------------------
# camera-view sharding: halo exchange + kv_map reproduce the unsharded cross-view attention (blocks.py:190-222)
# ---------------------------------------------------------------------------------------------------
NEIGHBORS = {0: [5, 1], 1: [0, 2], 2: [1, 3], 3: [2, 4], 4: [3, 5], 5: [4, 0]}


def _xview_reference(q, k, v, n_outer, n_cam, heads):
    from oracle.dualdiff_oracle import mha
    out = torch.zeros_like(q)
    qv, kv, vv = (t.reshape(n_outer, n_cam, *t.shape[1:]) for t in (q, k, v))
    ov = out.reshape(n_outer, n_cam, *out.shape[1:])
    for cam, nbrs in NEIGHBORS.items():
        for nb in nbrs:
            ov[:, cam] += mha(qv[:, cam], kv[:, nb], vv[:, nb], heads)
    return out


def _view_worker(rank, world, port, q_out):
    from dualdiff_b200.sharding import ViewShard
    from oracle.dualdiff_oracle import mha
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n_outer, n_cam, T, C, heads = 3, 6, 5, 16, 2
    g = torch.Generator().manual_seed(0)
    q, k, v = (torch.randn(n_outer * n_cam, T, C, generator=g) for _ in range(3))
    ref = _xview_reference(q, k, v, n_outer, n_cam, heads)
    vs = ViewShard(rank, world, n_cam)
    sel = torch.tensor([o * n_cam + c for o in range(n_outer) for c in vs.views])
    n_loc = len(sel)
    qkv = torch.cat([q[sel], k[sel], v[sel]], dim=-1).reshape(n_loc * T, 3 * C)
    buf = torch.zeros(vs.kv_rows(n_outer) * T, 3 * C)
    buf[: n_loc * T] = qkv
    vs.exchange(buf, n_outer, T)
    kv_map = vs.kv_map(n_outer)
    b3 = buf.reshape(-1, T, 3 * C)
    out = torch.zeros(n_loc, T, C)
    for s in range(2):
        idx = kv_map[:, s].long()
        out += mha(b3[:n_loc, :, :C], b3[idx][:, :, C:2 * C], b3[idx][:, :, 2 * C:], heads)
    err = (out - ref[sel]).abs().max().item()
    q_out.put((rank, err))
    dist.destroy_process_group()


def _run_view_sharding(world):
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_view_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, err in res:
        assert err < 1e-5, (world, rank, err)


def test_view_sharded_crossview_two_ranks_gloo():
    _run_view_sharding(2)     # left == right peer: exercises the send/recv posting order


def test_view_sharded_crossview_three_ranks_gloo():
    _run_view_sharding(3)


def _group_worker(rank, world, port, q_out):
    """ranks in groups of R per scene (ViewShard with world 4 -> 2 groups of 2): the scenes are dealt to the groups, the
    LayerNorm rows of the halo views are exchanged inside the group and their K/V projections recomputed locally"""
    from dualdiff_b200.sharding import ViewShard, default_ranks_per_scene
    from oracle.dualdiff_oracle import mha
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    scenes, n_cam, T, C, heads = 4, 6, 5, 16, 2
    g = torch.Generator().manual_seed(0)
    ln = torch.randn(scenes * n_cam, T, C, generator=g)
    wq, wk, wv = (torch.randn(C, C, generator=g) * 0.3 for _ in range(3))
    ref = _xview_reference(ln @ wq.T, ln @ wk.T, ln @ wv.T, scenes, n_cam, heads)
    vs = ViewShard(rank, world, n_cam)
    assert vs.ranks_per_scene == default_ranks_per_scene(world, n_cam)
    mine = vs.scenes(scenes)
    sel = torch.tensor([o * n_cam + c for o in mine for c in vs.views])
    n_outer, n_loc = len(mine), len(sel)
    ln_ext = torch.zeros(vs.kv_rows(n_outer) * T, C)
    ln_ext[: n_loc * T] = ln[sel].reshape(n_loc * T, C)
    vs.exchange_async(ln_ext, n_outer, T)        # CPU tensors: synchronous
    vs.exchange_wait()
    ext = ln_ext.reshape(-1, T, C)
    q, k, v = ext[:n_loc] @ wq.T, ext @ wk.T, ext @ wv.T    # K/V of the halo views recomputed from their LayerNorm rows
    kv_map = vs.kv_map(n_outer)
    out = torch.zeros(n_loc, T, C)
    for s in range(2):
        idx = kv_map[:, s].long()
        out += mha(q, k[idx], v[idx], heads)
    q_out.put((rank, (out - ref[sel]).abs().max().item(), vs.group_index, list(mine), vs.views, (vs.left, vs.right)))
    dist.destroy_process_group()


def test_view_shard_groups_four_ranks_gloo():
    world = 4
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_group_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, err, grp, mine, views, peers in res:
        assert err < 1e-4, (rank, err)
        assert grp == rank // 2 and mine == [2 * grp, 2 * grp + 1] and views == [3 * (rank % 2) + i for i in range(3)]
        assert peers == (rank ^ 1, rank ^ 1)          # exchanges stay inside the group


def test_view_sharded_step_inputs_follow_scenes_and_views():
    """what a rank of a 4-GPU job (2 groups x 2 ranks) keeps of a 4-scene call: its group's scenes, its 3 views; prompts per
    scene and per view (use_aug_text)"""
    from dualdiff_b200.sharding import ViewShard, slice_step_inputs
    from dualdiff_b200 import synthetic as S
    B, h, w = 4, 4, 6
    inp = S.make_inputs(B, h, w, seed=3, L_bg=3, L_fg=4, same_noise_across_views=False)
    vs = ViewShard(3, 4)                               # group 1 (scenes 2, 3), second half of the views
    assert list(vs.scenes(B)) == [2, 3] and vs.views == [3, 4, 5]
    args = (inp["latents"], inp["prompt_embeds"], inp["camera_param"], [inp["boxes_bg"], inp["boxes_fg"]], [inp["cond_bg"], inp["cond_fg"]])
    lat, pe, cam, boxes, images = slice_step_inputs(vs, *args)
    assert torch.equal(lat, inp["latents"][2:4, 3:6]) and torch.equal(cam, inp["camera_param"][2:4, 3:6])
    assert torch.equal(pe, torch.cat([inp["prompt_embeds"][2:4], inp["prompt_embeds"][B + 2:B + 4]]))       # uncond rows, then cond rows
    assert torch.equal(boxes[0]["bboxes"], inp["boxes_bg"]["bboxes"][2:4, 3:6]) and torch.equal(boxes[1]["bboxes"], inp["boxes_fg"]["bboxes"][2:4])
    assert torch.equal(images[0], inp["cond_bg"][2:4][..., 3 * 48:6 * 48])
    assert torch.equal(images[1], inp["cond_fg"].reshape(B, 6, 320, h, w)[2:4, 3:6].reshape(6, 320, h, w))
    # one prompt per view: rows are (cfg half, scene, view)
    pv = torch.arange(2 * B * 6, dtype=torch.float32).view(-1, 1, 1).expand(2 * B * 6, 2, 3).contiguous()
    _, pe2, _, _, _ = slice_step_inputs(vs, inp["latents"], pv, *args[2:])
    want = [half * B * 6 + s * 6 + v for half in range(2) for s in (2, 3) for v in (3, 4, 5)]
    assert pe2[:, 0, 0].tolist() == [float(x) for x in want]
    # the caller already cut the scenes of the group
    lat3, _, _, _, _ = slice_step_inputs(vs, inp["latents"][2:4], inp["prompt_embeds"][[2, 3, 6, 7]], inp["camera_param"][2:4],
                                         [{k: v[2:4] for k, v in inp["boxes_bg"].items()}, {k: v[2:4] for k, v in inp["boxes_fg"].items()}],
                                         [inp["cond_bg"][2:4], inp["cond_fg"][12:24]], scenes_sliced=True)
    assert torch.equal(lat3, lat)


def test_view_shard_world_sizes():
    from dualdiff_b200.sharding import ViewShard, default_ranks_per_scene
    import pytest
    assert [default_ranks_per_scene(w) for w in (1, 2, 3, 4, 6, 8, 12)] == [1, 2, 3, 2, 6, 2, 6]
    for world in (1, 2, 3, 4, 6, 8):
        shards = [ViewShard(r, world) for r in range(world)]
        R = shards[0].ranks_per_scene
        scenes = 2 * (world // R)
        units = sorted((s, v) for vs in shards for s in vs.scenes(scenes) for v in vs.views)
        assert units == [(s, v) for s in range(scenes) for v in range(6)]          # every (scene, view) exactly once
    with pytest.raises(ValueError):
        ViewShard(0, 8).scenes(3)                    # 4 groups need at least 4 scenes
    with pytest.raises(ValueError):
        ViewShard(0, 8, ranks_per_scene=3)


def test_view_shard_kv_map_and_slicing():
    from dualdiff_b200.sharding import ViewShard, slice_views
    from dualdiff_b200 import synthetic as S
    vs = ViewShard(1, 3)
    assert vs.views == [2, 3] and vs.left == 0 and vs.right == 2
    m = vs.kv_map(2)
    # image (scene 0, local view 0 = global 2): left neighbour = halo-left of scene 0, right = local view 1
    assert m.tolist() == [[4, 1], [0, 6], [5, 3], [2, 7]]
    inp = S.make_inputs(2, 4, 6, seed=3, L_bg=3, L_fg=4)
    loc = slice_views(inp, vs.views)
    assert loc["latents"].shape == (2, 2, 4, 4, 6) and loc["cond_bg"].shape[-1] == 2 * 8 * 6
    assert torch.equal(loc["cond_bg"][..., :48], inp["cond_bg"][..., 2 * 48:3 * 48])
    assert torch.equal(loc["cond_fg"].reshape(2, 2, 320, 4, 6)[:, 1], inp["cond_fg"].reshape(2, 6, 320, 4, 6)[:, 3])
    assert loc["boxes_fg"]["bboxes"].shape[1] == 1   # view-shared map vectors stay whole


# ---------------------------------------------------------------------------------------------------
# frame sharding of a video clip (BASELINE config 5): all-gather of the projected rows + the rank-major address
# arithmetic of dd_temporal_attention reproduce the unsharded temporal attention (oracle.temporal_attention's core)
# ---------------------------------------------------------------------------------------------------
def _frame_worker(rank, world, port, q_out):
    from dualdiff_b200.sharding import FrameShard, slice_frames, gathered_kv_image
    from oracle.dualdiff_oracle import mha
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n_clip, F, V, T, C, heads = 2, 4, 3, 5, 16, 2
    g = torch.Generator().manual_seed(0)
    q, k, v = (torch.randn(n_clip * F * V, T, C, generator=g) for _ in range(3))
    # unsharded reference: sequences over frames at every (clip, view, token)
    def seq(t):
        return t.reshape(n_clip, F, V, T, C).permute(0, 2, 3, 1, 4).reshape(n_clip * V * T, F, C)
    ref = mha(seq(q), seq(k), seq(v), heads).reshape(n_clip, V, T, F, C).permute(0, 3, 1, 2, 4).reshape(n_clip * F * V, T, C)
    fs = FrameShard(rank, world, F)
    loc = [slice_frames(t, fs.frames, F, V) for t in (q, k, v)]
    rows = torch.cat(loc, dim=-1).reshape(-1, 3 * C)                     # this rank's fused projection rows
    allkv = fs.gather(rows).reshape(-1, T, 3 * C)                        # [world * n_loc_img, T, 3C]
    n_loc = n_clip * fs.f_loc * V
    out = torch.zeros(n_loc, T, C)
    for c in range(n_clip):
        for view in range(V):
            idx = torch.tensor([gathered_kv_image(c, f, view, n_clip, fs.f_loc, V) for f in range(F)])
            kk = allkv[idx][:, :, C:2 * C].permute(1, 0, 2)              # [T, F, C]
            vv = allkv[idx][:, :, 2 * C:].permute(1, 0, 2)
            for fl in range(fs.f_loc):
                img = (c * fs.f_loc + fl) * V + view
                qq = loc[0][img][:, None, :]                             # [T, 1, C]
                out[img] = mha(qq, kk, vv, heads)[:, 0]
    want = slice_frames(ref, fs.frames, F, V)
    q_out.put((rank, (out - want).abs().max().item()))
    dist.destroy_process_group()


def test_frame_sharded_temporal_attention_two_ranks_gloo():
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_frame_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, err in res:
        assert err < 1e-5, (rank, err)


def _frame_a2a_worker(rank, world, port, q_out):
    """token-sharded temporal attention: all-to-all of the LayerNorm rows, projection + attention over ALL frames on this
    rank's share of the tokens (addressing the received rank blocks like csrc/dd_temporal.cu), all-to-all back"""
    from dualdiff_b200.sharding import FrameShard, slice_frames
    from oracle.dualdiff_oracle import mha
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n_clip, F, V, T, C, heads = 2, 4, 3, 5, 16, 2          # T = 5 over 2 ranks: uneven token shares (3 + 2)
    g = torch.Generator().manual_seed(0)
    ln = torch.randn(n_clip * F * V, T, C, generator=g)
    wq, wk, wv = (torch.randn(C, C, generator=g) * 0.3 for _ in range(3))

    def seq(t):
        return t.reshape(n_clip, F, V, T, C).permute(0, 2, 3, 1, 4).reshape(n_clip * V * T, F, C)
    ref = mha(seq(ln @ wq.T), seq(ln @ wk.T), seq(ln @ wv.T), heads)
    ref = ref.reshape(n_clip, V, T, F, C).permute(0, 3, 1, 2, 4).reshape(n_clip * F * V, T, C)
    fs = FrameShard(rank, world, F)
    loc = slice_frames(ln, fs.frames, F, V)                  # [n_loc, T, C], images (clip, local frame, view)
    n_loc = loc.shape[0]
    t_me = len(fs.token_range(T))
    tok = fs.to_token_shards(loc.reshape(n_loc * T, C), n_loc, T)          # [world * n_loc * t_me, C]
    blk = tok.reshape(world, n_clip, fs.f_loc, V, t_me, C)                   # rank blocks, as the kernel addresses them
    allf = blk.permute(1, 3, 4, 0, 2, 5).reshape(n_clip * V * t_me, F, C)    # sequences over global frames (rank-major = frame order)
    o = mha(allf @ wq.T, allf @ wk.T, allf @ wv.T, heads)
    a_tok = o.reshape(n_clip, V, t_me, world, fs.f_loc, C).permute(3, 0, 4, 1, 2, 5).reshape(world * n_loc * t_me, C)
    out = fs.from_token_shards(a_tok, n_loc, T).reshape(n_loc, T, C)
    want = slice_frames(ref, fs.frames, F, V)
    q_out.put((rank, (out - want).abs().max().item(), t_me))
    dist.destroy_process_group()


def test_frame_sharded_temporal_attention_all_to_all_gloo():
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_frame_a2a_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [r[2] for r in res] == [3, 2]
    for rank, err, _ in res:
        assert err < 1e-5, (rank, err)


def test_frame_shard_partition():
    from dualdiff_b200.sharding import FrameShard
    import pytest
    assert sum((FrameShard(r, 8, 16).frames for r in range(8)), []) == list(range(16))
    with pytest.raises(ValueError):
        FrameShard(0, 3, 16)
