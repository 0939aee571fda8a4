"""The sampler loop body of DualDiff as one B200-resident object.

Reference: pipeline/pipeline_bev_controlnet.py:349-512 (`StableDiffusionBEVControlNetPipeline.__call__`):
per step, [ControlNet-bg, ControlNet-fg] -> residual sum -> multi-view UNet -> classifier-free guidance ->
scheduler.step.  `DualDiffDenoiser` keeps that contract (uncond half first, tokens from branch 0, residuals
summed across branches, same noise handling) but
  * hoists the timestep-invariant work into `prepare()` (tokens, text K/V of all 30 cross-attentions,
    condition embedding, Semantic Fusion Attention),
  * runs a step as a CUDA graph of hand-written kernels (no host sync, no torch math),
  * fuses CFG + UniPC/DDIM into one kernel.
Scenes are independent (cross-view attention never crosses scenes, blocks.py:196-197), so multi-GPU use is
one denoiser per rank over its shard of scenes — no collectives.
"""
from typing import Dict, List, Optional

import torch

from . import engine, ops
from .scheduler import UniPCMultistepScheduler


class DualDiffDenoiser:
    def __init__(self, unet, controlnets, scheduler=None, guidance_scale: float = 2.0, use_cuda_graph: bool = True,
                 view_shard=None):
        assert len(controlnets) == 2, "dual branch: [controlnet_bg, controlnet_fg]"
        self.unet, self.nets = unet, list(controlnets)
        self.guidance_scale = guidance_scale
        self.cfg = guidance_scale > 1.0          # pipeline: do_classifier_free_guidance = guidance_scale > 1
        self.scheduler = scheduler if scheduler is not None else UniPCMultistepScheduler()
        self.scheduler.guidance_scale = guidance_scale
        self.view_shard = view_shard               # sharding.ViewShard: camera views split across ranks (config 4)
        # The view- and frame-sharded steps are captured like the unsharded one: NCCL send/recv, all-gather and the
        # side-stream fork/join are capturable.  A capture the NCCL build refuses falls back to eager launches (graph_note).
        self.use_cuda_graph = use_cuda_graph
        self.graph_note = None                     # why a requested CUDA graph was not used (capture failed), if so
        self._graph = None
        self.device = None
        # The two condition branches and the UNet encoder are mutually independent until the skip/mid residual adds
        # (unet_2d_condition_multiview.py:464-488): run them on three streams so the small low-resolution kernels
        # (grids below one wave of 148 SMs) overlap instead of serialising.
        self.parallel_branches = True
        self._side = None
        # measurement hook: when a list, the serial (parallel_branches = False) step records a CUDA event at the start and
        # after ControlNet-bg, ControlNet-fg, the UNet and the CFG + scheduler kernel (bench.py's sub_metrics)
        self.phase_events = None

    # -------------------------------------------------------------------------------------------------
    def prepare(self, latents, prompt_embeds, camera_param, bboxes_3d_data: List[Dict[str, torch.Tensor]], images,
                num_inference_steps: int, scenes_sliced: bool = False):
        """latents (B, 6, 4, h, w) fp32; prompt_embeds (2B, 77, 768) uncond first (or (B, ...) without CFG);
        camera_param (B, 6, 3, 7); bboxes_3d_data = [bg boxes, fg map vectors]; images = [bg panorama
        (B, 3, 8h, 48w), fg ORS (B*6, 320, h, w)].
        View-sharded: every rank passes the full inputs of the call (`scenes_sliced`: only the scenes of its group)."""
        if self.view_shard is not None:
            # every rank passes the FULL inputs of the call and keeps the scenes of its group and its own camera views
            from .sharding import slice_step_inputs
            latents, prompt_embeds, camera_param, bboxes_3d_data, images = slice_step_inputs(
                self.view_shard, latents, prompt_embeds, camera_param, bboxes_3d_data, images, scenes_sliced)
        dev = latents.device
        if dev.type != "cuda":
            raise RuntimeError("dualdiff_b200 has no CPU path: inputs must be CUDA tensors")
        self.device = dev
        B, n_cam, c, H, W = latents.shape
        self.B, self.n_cam, self.H, self.W = B, n_cam, H, W
        G = 2 if self.cfg else 1
        self.G = G
        n = G * B * n_cam
        self.n = n
        for m in [self.unet] + self.nets:
            m.ensure_packed(dev)          # packs on first use and re-packs when the weights changed since
        if self.cfg:  # uncond in the front, cond in the tail (pipeline:349-375; camera from nets[0])
            kw = self.nets[0].add_uncond_to_kwargs(camera_param=camera_param, bboxes_3d_data=bboxes_3d_data, image=None)
            cam, boxes = kw["camera_param"], kw["bboxes_3d_data"]
            text = prompt_embeds
            # one prompt per scene, or one per view (use_aug_text, unet_addon_rawbox.py:351-352): uncond rows first either way
            assert text.shape[0] in (2 * B, 2 * B * n_cam), (text.shape, B, n_cam)
        else:
            cam, boxes = camera_param, bboxes_3d_data
            per = n_cam if prompt_embeds.shape[0] % (B * n_cam) == 0 and n_cam > 1 else 1
            text = prompt_embeds[-B * per:]
        # condition image / ORS tensor are identical in both CFG halves (pipeline:351-373 skips the uncond map)
        conds = [torch.cat([images[0]] * G) if G > 1 else images[0], torch.cat([images[1]] * G) if G > 1 else images[1]]
        preps = [net.prepare_condition(cam, text, boxes[i], conds[i], H, W) for i, net in enumerate(self.nets)]
        Pu = self.unet._packed
        unet_text_kv = engine.prepare_text(Pu, engine.ATTN2_LAYERS_UNET, preps[0].enc_rows)  # tokens of branch 0
        if self.view_shard is None:
            kv_map = engine.make_kv_map(n, Pu["view_pairs"], dev, Pu["xview_mode"])
        else:
            if Pu["xview_mode"] == "self":
                raise NotImplementedError("camera-view sharding exchanges the two ring neighbours only: neighboring_attn_type='self' "
                                          "needs every view of the scene on the rank")
            assert self.view_shard.v_loc == n_cam
            kv_map = self.view_shard.kv_map(G * B, dev)
        self.scheduler.set_timesteps(num_inference_steps, device=dev)
        self.coef_table = self.scheduler.coef_table(dev)
        self.t_table = self.scheduler.timesteps.to(device=dev, dtype=torch.float32)
        # state (fp32, NCHW-flat): latents, last corrected sample, x0 history
        lat = latents.reshape(B * n_cam, c, H, W).float().contiguous()
        # every tensor the captured step reads: when a graph of the same geometry exists (the pipeline calling again with
        # another prompt / scene), the new values are copied into the buffers it was captured on and the graph is kept
        sig = (self.cfg, B, n_cam, c, H, W, tuple(p.lk for p in preps), str(dev))
        new_in = _graph_inputs(preps, unet_text_kv, kv_map)
        # the graph also reads the packed weights: a re-pack (weights reloaded) must re-capture.  The references held here
        # keep the buffers a kept graph points at alive.
        packs = [m._packed for m in [self.unet] + self.nets]
        same_weights = len(packs) == len(getattr(self, "_packs", [])) and all(a is b for a, b in zip(packs, self._packs))
        if self._graph is not None and same_weights and getattr(self, "_sig", None) == sig and \
                [(t.shape, t.dtype) for t in new_in] == [(t.shape, t.dtype) for t in self._static_in]:
            for dst, src in zip(self._static_in, new_in):
                dst.copy_(src)
            self.latents.copy_(lat)
            for t in (self.last, self.m0, self.m1):
                t.zero_()
        else:
            self.preps, self.unet_text_kv, self.kv_map = preps, unet_text_kv, kv_map
            self._static_in, self._sig, self._packs = new_in, sig, packs
            self.latents = lat.clone()
            self.last = torch.zeros_like(self.latents)
            self.m0 = torch.zeros_like(self.latents)
            self.m1 = torch.zeros_like(self.latents)
            self.t_cur = torch.zeros(1, device=dev, dtype=torch.float32)
            self.coef_cur = torch.zeros(16, device=dev, dtype=torch.float32)
            self._graph = None
        self.eps_rows = None
        self.step_index = 0
        return self

    # -------------------------------------------------------------------------------------------------
    def _step_kernels(self):
        """one loop body (pipeline:381-504) on the current stream; reads t / coefficients from device buffers"""
        B6, H, W, G = self.B * self.n_cam, self.H, self.W, self.G
        Pu = self.unet._packed
        ev = self.phase_events if not self.parallel_branches else None

        def mark():
            if ev is not None:
                e = torch.cuda.Event(enable_timing=True)
                e.record()
                ev.append(e)

        if not self.parallel_branches:
            acc = None
            mark()
            for net, prep in zip(self.nets, self.preps):
                down, mid = engine.controlnet_forward(net._packed, prep, self.latents, G, B6, H, W, self.t_cur,
                                                      None if acc is None else acc)
                acc = down + [mid]
                mark()
            temb = engine.time_embedding(Pu, self.t_cur)
            ctx = engine.StepCtx(n=self.n, temb=temb, temb_rows_per_img_factor=self.n, text_kv=self.unet_text_kv,
                                 lk=self.preps[0].lk, kv_map=self.kv_map, n_nbr=Pu["n_nbr"], xview_concat=Pu["xview_mode"] != "add",
                                 view_shard=self.view_shard,
                                 n_outer=self.G * self.B)
            self.unet.video_ctx(ctx)   # video configuration: scenes are (clip, frame) pairs, frame-minor
            eps = engine.unet_forward(Pu, self.latents, G, B6, H, W, ctx, acc[:12], acc[12])
            mark()
        else:
            main = torch.cuda.current_stream()
            if self._side is None:
                self._side = [torch.cuda.Stream(), torch.cuda.Stream()]
            res = []
            for net, prep, st in zip(self.nets, self.preps, self._side):
                st.wait_stream(main)                       # fork
                with torch.cuda.stream(st):
                    down, mid = engine.controlnet_forward(net._packed, prep, self.latents, G, B6, H, W, self.t_cur, None)
                for t in down + [mid]:
                    t.record_stream(main)                  # consumed on the main stream after the join
                res.append((down, mid))
            temb = engine.time_embedding(Pu, self.t_cur)
            ctx = engine.StepCtx(n=self.n, temb=temb, temb_rows_per_img_factor=self.n, text_kv=self.unet_text_kv,
                                 lk=self.preps[0].lk, kv_map=self.kv_map, n_nbr=Pu["n_nbr"], xview_concat=Pu["xview_mode"] != "add",
                                 view_shard=self.view_shard,
                                 n_outer=self.G * self.B)
            self.unet.video_ctx(ctx)   # video configuration: scenes are (clip, frame) pairs, frame-minor

            def join():
                for st in self._side:
                    main.wait_stream(st)

            eps = engine.unet_forward(Pu, self.latents, G, B6, H, W, ctx, res[0][0], res[0][1], res[1][0], res[1][1],
                                      before_residuals=join)
        ops.cfg_sched_step(eps, self.latents, self.last, self.m0, self.m1, self.coef_cur, n_img=B6, c=4, hw=H * W,
                           cfg=self.cfg, eps_nchw=False)
        mark()
        self.eps_rows = eps
        return eps

    def step(self, i: Optional[int] = None):
        """advance the latents by one sampler step (index i of scheduler.timesteps)"""
        i = self.step_index if i is None else i
        self.t_cur.copy_(self.t_table[i:i + 1], non_blocking=True)
        self.coef_cur.copy_(self.coef_table[i], non_blocking=True)
        if not self.use_cuda_graph:
            self._step_kernels()
        else:
            if self._graph is None:
                # warm-up on a side stream (lazy kernel-attribute init, allocator warm-up), restoring the state after
                saved = [t.clone() for t in (self.latents, self.last, self.m0, self.m1)]
                s = torch.cuda.Stream()
                s.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(s):
                    self._step_kernels()
                torch.cuda.current_stream().wait_stream(s)
                for t, sv in zip((self.latents, self.last, self.m0, self.m1), saved):
                    t.copy_(sv)
                graph = torch.cuda.CUDAGraph()
                n0 = _launches()
                try:
                    with torch.cuda.graph(graph):
                        self._step_kernels()
                    self._graph = graph
                except Exception as e:      # a capture the driver / NCCL build refuses: launch eagerly, and say so
                    if self.view_shard is None and getattr(self.unet, "frame_shard", None) is None:
                        raise
                    import logging
                    self.graph_note = f"CUDA graph capture of the sharded step failed ({type(e).__name__}: {e}); eager launches"
                    logging.getLogger(__name__).warning(self.graph_note)
                    self.use_cuda_graph = False
                    torch.cuda.synchronize()
                self.launches_per_step = _launches() - n0
                for t, sv in zip((self.latents, self.last, self.m0, self.m1), saved):
                    t.copy_(sv)
                if self._graph is None:
                    self._step_kernels()
                    self.step_index = i + 1
                    return self.latents
            self._graph.replay()
        self.step_index = i + 1
        return self.latents

    def release_graph(self):
        """drop the captured step.  A graph that contains NCCL operations (view- / frame-sharded steps) keeps the communicator
        busy: `dist.destroy_process_group()` waits for it, so sharded callers release the graph before tearing the group down."""
        if self._graph is not None:
            torch.cuda.synchronize()
            self._graph = None
            torch.cuda.synchronize()

    def run(self):
        for i in range(len(self.scheduler.timesteps)):
            self.step(i)
        return self.latents.reshape(self.B, self.n_cam, 4, self.H, self.W)


def _graph_inputs(preps, unet_text_kv, kv_map):
    """the tensors prepare() produces and the step kernels read, in a fixed order"""
    ts = []
    for p in preps:
        ts += [p.enc_rows, p.cond] + [p.text_kv[k] for k in sorted(p.text_kv)]
    ts += [unet_text_kv[k] for k in sorted(unet_text_kv)]
    ts.append(kv_map)
    return ts


def _launches():
    from . import _lib
    return _lib.lib().dd_launch_count()
