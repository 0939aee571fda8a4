"""shared helpers for the parity tests (builds the product modules with seeded synthetic weights)"""
import json
import os

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
NEIGHBORS = {0: [5, 1], 1: [0, 2], 2: [1, 3], 3: [2, 4], 4: [3, 5], 5: [4, 0]}
SEEDS = {"unet": 0, "bg": 1, "fg": 2}

# sd-controlnet-seg/config.json of the reference (the configuration both branches are built from)
CONTROLNET_CONFIG = dict(
    act_fn="silu", attention_head_dim=8,
    bbox_embedder_cls="magicdrive.networks.bbox_embedder.ContinuousBBoxWithTextEmbedding",
    bbox_embedder_param=dict(class_token_dim=768, embedder_num_freq=4, minmax_normalize=False, mode="all-xyz",
                             n_classes=10, proj_dims=[768, 512, 512, 768], trainable_class_token=False,
                             use_text_encoder_init=True),
    block_out_channels=[320, 640, 1280, 1280],
    cam_embedder_param=dict(include_input=True, input_dims=3, log_sampling=True, num_freqs=4),
    camera_in_dim=189, camera_out_dim=768, class_embed_type=None,
    conditioning_embedding_out_channels=[16, 32, 96, 256], controlnet_conditioning_channel_order="rgb",
    cross_attention_dim=768,
    down_block_types=["CrossAttnDownBlock2D", "CrossAttnDownBlock2D", "CrossAttnDownBlock2D", "DownBlock2D"],
    downsample_padding=1, drop_cam_num=6, drop_cam_with_box=False, drop_cond_ratio=0.25, flip_sin_to_cos=True,
    freq_shift=0, global_pool_conditions=False, in_channels=4, layers_per_block=2,
    map_embedder_cls="magicdrive.networks.map_embedder.ControlNetConditioningEmbedding",
    map_embedder_param=dict(block_out_channels=[16, 32, 96, 256]), map_size=[8, 200, 200],
    mid_block_scale_factor=1, norm_eps=1e-05, norm_num_groups=32, num_class_embeds=None,
    only_cross_attention=False, projection_class_embeddings_input_dim=None, resnet_time_scale_shift="default",
    uncond_cam_in_dim=[3, 7], upcast_attention=False, use_linear_projection=False, use_uncond_map=None)


def build_models(device="cpu", load=True):
    """product modules wired like misc/test_utils.py:97-171, with the seeded synthetic weights"""
    from dualdiff_b200 import synthetic as S
    from dualdiff_b200.networks import BEVControlNetModel, UNet2DConditionModelMultiview
    with torch.device("meta"):
        unet = UNet2DConditionModelMultiview(cross_attention_dim=768, neighboring_view_pair=NEIGHBORS)
        nets = [BEVControlNetModel(**CONTROLNET_CONFIG) for _ in range(2)]
    for i, c in enumerate(nets):
        c.use_cam_in_temb = False
        c.use_box_adapter = False
        c.adm_proj = None
        c.use_txt_con_fusion = True
        c.use_txt_con_fusionp = False
        c.txt_con_fusionp = None
        c.use_occ_3d = i == 1
        if c.use_occ_3d:
            c.controlnet_cond_embedding = None
    sds = {}
    for name, m in (("unet", unet), ("bg", nets[0]), ("fg", nets[1])):
        man = S.manifest_of(m)
        if load:
            sd = S.init_state_dict(man, SEEDS[name])
            m.load_state_dict(sd, strict=True, assign=True)
            sds[name] = sd
        m.eval()
    return unet, nets, sds


def to_dev(obj, dev):
    if torch.is_tensor(obj):
        return obj.to(dev)
    if isinstance(obj, dict):
        return {k: to_dev(v, dev) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return type(obj)(to_dev(v, dev) for v in obj)
    return obj


def metrics(out, ref):
    out, ref = out.double().flatten(), ref.double().flatten()
    cos = torch.nn.functional.cosine_similarity(out, ref, dim=0).item()
    rel_l2 = ((out - ref).norm() / ref.norm()).item()
    max_rel = ((out - ref).abs().max() / ref.abs().max()).item()
    return dict(cos=cos, rel_l2=rel_l2, max_rel=max_rel)


def strided_sample(t, n):
    """every k-th element of the flattened tensor, k chosen so that about n elements are kept (k is odd, so the sample
    walks through all channels / rows / columns of an NCHW tensor).  Used by oracle/make_golden.py to keep large
    reference tensors as small fixtures and by the tests to sample the CUDA result the same way."""
    f = t.reshape(-1)
    k = max(1, f.numel() // n) | 1
    return f[::k].clone()


def make_scene_batch(seeds, h, w, L_bg=28, L_fg=32):
    """step inputs of len(seeds) scenes where scene i is exactly synthetic.make_inputs(1, h, w, seed=seeds[i]) -- so that a
    scene computed inside a batch can be compared with the same scene computed alone (make_inputs(B, ...) draws all scenes
    from one generator stream).  prompt_embeds keeps the pipeline's layout: all uncond rows first, then all cond rows."""
    from dualdiff_b200 import synthetic as S
    per = [S.make_inputs(1, h, w, seed=s, L_bg=L_bg, L_fg=L_fg) for s in seeds]
    out = {}
    for k in ("latents", "camera_param", "cond_bg", "cond_fg"):
        out[k] = torch.cat([p[k] for p in per])
    out["prompt_embeds"] = torch.cat([p["prompt_embeds"][:1] for p in per] + [p["prompt_embeds"][1:] for p in per])
    for k in ("boxes_bg", "boxes_fg"):
        out[k] = {kk: torch.cat([p[k][kk] for p in per]) for kk in per[0][k]}
    return out


def scene_of(inputs, i, n_cam=6):
    """scene i of a make_scene_batch() dict, in the layout of synthetic.make_inputs(1, ...)"""
    B = inputs["latents"].shape[0]
    out = {k: inputs[k][i:i + 1] for k in ("latents", "camera_param", "cond_bg")}
    out["cond_fg"] = inputs["cond_fg"][i * n_cam:(i + 1) * n_cam]
    out["prompt_embeds"] = torch.stack([inputs["prompt_embeds"][i], inputs["prompt_embeds"][B + i]])
    for k in ("boxes_bg", "boxes_fg"):
        out[k] = {kk: v[i:i + 1] for kk, v in inputs[k].items()}
    return out
