"""The drop-in modules that were parameter containers in round 1 called as modules (reference call sites
unet_addon_rawbox.py:967-978), the 40-point map-vector embedder (misc/test_utils.py:116-121) and a non-default cross-view
topology, each against the oracle.  Tolerance: bf16 kernels vs fp32 oracle, max-abs <= 2e-2 * max|ref| / cosine >= 0.999."""
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import common  # noqa: E402

pytestmark = pytest.mark.gpu


def _seeded(module, seed):
    from dualdiff_b200 import synthetic as S
    sd = S.init_state_dict(S.manifest_of(module), seed)
    module.load_state_dict(sd, strict=True, assign=True)
    return sd


def test_sfa_module_forward_matches_oracle():
    from dualdiff_b200.networks.txt_con_fusion import txt_con_XFormersAttn
    from oracle import dualdiff_oracle as O
    with torch.device("meta"):
        m = txt_con_XFormersAttn()
    sd = _seeded(m, 3)
    g = torch.Generator().manual_seed(0)
    cond = torch.randn(4, 320, 9, 14, generator=g)
    txt = torch.randn(4, 77, 768, generator=g)
    with torch.no_grad():
        ref = O.sfa({"txt_con_fusion." + k: v for k, v in sd.items()}, cond, txt)
    out = m.cuda()(attn=None, hidden_states=cond.cuda(), encoder_hidden_states=txt.cuda())
    assert out.shape == ref.shape and out.dtype == torch.float32
    r = common.metrics(out.cpu(), ref)
    print("SFA module vs oracle:", r)
    assert r["cos"] >= 0.999 and r["max_rel"] <= 2e-2, r
    with pytest.raises(NotImplementedError):
        m(attn=None, hidden_states=cond.cuda(), encoder_hidden_states=txt.cuda(), attention_mask=torch.zeros(1).cuda())


def test_sfa_plus_module_forward_matches_reference_golden():
    """txt_con_XFormersAttn_plus (txt_con_fusion.py:184-337) as a module against the golden of the reference's own class;
    bf16 kernels vs fp32: cosine >= 0.999, max-rel <= 2e-2"""
    sys.path.insert(0, os.path.join(common.ROOT, "oracle"))
    import make_golden_sfa_plus as MG
    from dualdiff_b200.networks.txt_con_fusion import txt_con_XFormersAttn_plus
    fix = torch.load(os.path.join(common.GOLDEN, "sfa_plus_small.pt"))
    with torch.device("meta"):
        m = txt_con_XFormersAttn_plus()
    _seeded(m, fix["weight_seed"])
    cond, txt = MG.inputs(fix["input_seed"])
    out = m.cuda()(attn=None, hidden_states=cond.cuda(), encoder_hidden_states=txt.cuda())
    assert out.shape == fix["out"].shape and out.dtype == torch.float32
    r = common.metrics(out.cpu(), fix["out"])
    print("SFA+ module vs reference golden:", r)
    assert r["cos"] >= 0.999 and r["max_rel"] <= 2e-2, r
    # the text really reaches the output through the chained attention
    out2 = m(attn=None, hidden_states=cond.cuda(), encoder_hidden_states=torch.flip(txt, dims=[0]).cuda())
    assert (out2 - out).abs().max() > 1e-3


def test_branch_with_sfa_plus_matches_oracle():
    """a condition branch configured with use_txt_con_fusionp (configs/exp/occ_bg_fusionp.yaml) against oracle.controlnet_forward"""
    from dualdiff_b200 import synthetic as S
    from dualdiff_b200.networks import BEVControlNetModel
    from oracle import dualdiff_oracle as O
    with torch.device("meta"):
        net = BEVControlNetModel(**common.CONTROLNET_CONFIG)
    net.use_cam_in_temb = False
    net.use_box_adapter = False
    net.adm_proj = None
    net.use_txt_con_fusion, net.txt_con_fusion = False, None
    net.use_txt_con_fusionp = True
    net.use_occ_3d = False
    sd = _seeded(net, 1)
    assert any(k.startswith("txt_con_fusionp.") for k in sd) and not any(k.startswith("txt_con_fusion.") for k in sd)
    h, w = 8, 12
    inp = S.make_inputs(1, h, w, seed=2, L_bg=5, L_fg=4)
    t = torch.tensor([500])
    with torch.no_grad():
        d_ref, m_ref, _ = O.controlnet_forward(sd, inp["latents"], t, inp["camera_param"], inp["boxes_bg"], inp["prompt_embeds"][1:],
                                               inp["cond_bg"], use_occ_3d=False)
    net = net.cuda()
    dev = torch.device("cuda")
    out = net(inp["latents"].to(dev), 500, camera_param=inp["camera_param"].to(dev), bboxes_3d_data=common.to_dev(inp["boxes_bg"], dev),
              encoder_hidden_states=inp["prompt_embeds"][1:].to(dev), controlnet_cond=inp["cond_bg"].to(dev), use_aug_text=False)
    r = common.metrics(out.mid_block_res_sample.float().cpu(), m_ref)
    print("branch with SFA+ mid residual vs oracle:", r)
    assert r["cos"] >= 0.999 and r["rel_l2"] <= 2e-2, r
    r0 = common.metrics(out.down_block_res_samples[0].float().cpu(), d_ref[0])
    assert r0["cos"] >= 0.999 and r0["rel_l2"] <= 2e-2, r0


def test_cond_embedding_module_forward_matches_oracle():
    from dualdiff_b200.networks.map_embedder import ControlNetConditioningEmbedding
    from oracle import dualdiff_oracle as O
    with torch.device("meta"):
        m = ControlNetConditioningEmbedding(320, block_out_channels=(16, 32, 96, 256))
    sd = _seeded(m, 4)
    for k in ("conv_out.weight", "conv_out.bias"):       # zero-initialised upstream: randomise or the check is vacuous
        sd[k] = torch.randn(sd[k].shape, generator=torch.Generator().manual_seed(1)) * 0.02
    m.load_state_dict(sd, strict=True, assign=True)
    pano = torch.rand(2, 3, 64, 6 * 96, generator=torch.Generator().manual_seed(2))
    with torch.no_grad():
        ref = O.cond_embedding({"controlnet_cond_embedding." + k: v for k, v in sd.items()}, pano)
    out = m.cuda()(pano.cuda())
    assert out.shape == ref.shape == (12, 320, 8, 12)
    r = common.metrics(out.cpu(), ref)
    print("ControlNetConditioningEmbedding module vs oracle:", r)
    assert r["cos"] >= 0.999 and r["rel_l2"] <= 2e-2, r


def test_box_tokens_with_40_point_map_vectors():
    """`bbox_embedder.reinitialize()` (40 points per map vector) followed by a reload, as misc/test_utils.py:116-121 does"""
    from dualdiff_b200 import engine, synthetic as S
    from dualdiff_b200.networks.bbox_embedder import ContinuousBBoxWithTextEmbedding
    from oracle import dualdiff_oracle as O
    e = ContinuousBBoxWithTextEmbedding(n_classes=3, mode="all-xyz", minmax_normalize=False, embedder_num_freq=4,
                                        proj_dims=[768, 512, 512, 768])
    e.reinitialize()
    sd = S.init_state_dict(S.manifest_of(e), 9)
    e.load_state_dict(sd)
    sdp = {"bbox_embedder." + k: v for k, v in sd.items()}
    g = torch.Generator().manual_seed(3)
    R, L = 2, 11
    vec = torch.rand(R, L, 40, 3, generator=g) * 100 - 50
    cls = torch.randint(0, 3, (R, L), generator=g)
    msk = torch.rand(R, L, generator=g) < 0.7
    with torch.no_grad():
        ref = O.box_tokens(sdp, vec, cls, msk)
    pk = engine.Packer(sdp, torch.device("cuda:0"))
    for n in ("bbox_proj", "second_linear.0", "second_linear.2", "second_linear.4"):
        pk.lin32("bbox_embedder." + n)
    for n in ("_class_tokens", "null_class_feature", "null_pos_feature"):
        pk.put("bbox_embedder." + n, pk.f32(sdp["bbox_embedder." + n]))
    out = engine.box_tokens(pk.out, vec.cuda(), cls.cuda(), msk.cuda()).reshape(R, L, 768).cpu()
    assert (out - ref).abs().max() <= 1e-3 * ref.abs().max()
    with pytest.raises(ValueError, match="points per box"):
        engine.box_tokens(pk.out, vec[:, :, :8].contiguous().cuda(), cls.cuda(), msk.cuda())


@pytest.mark.parametrize("pairs", [{0: [1], 1: [2], 2: [0]}, {0: [3, 1], 1: [0, 2], 2: [1, 3], 3: [2, 0]}])
def test_block_with_another_cross_view_topology(pairs):
    """neighboring_view_pair is read from the model: 3 views with one neighbour each, 4 views in a ring -- kv_map and the
    fused to_out bias multiplier (one b_o per neighbour) follow the table (networks/blocks.py:106-121,203-217)"""
    from dualdiff_b200 import synthetic as S
    from dualdiff_b200.networks import BasicMultiviewTransformerBlock
    from oracle import dualdiff_oracle as O
    n_cam = len(pairs)
    with torch.device("meta"):
        blk = BasicMultiviewTransformerBlock(320, 8, 40, cross_attention_dim=768, neighboring_view_pair=pairs)
    sd = _seeded(blk, 7)
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2 * n_cam, 96, 320, generator=g)
    enc = torch.randn(2 * n_cam, 83, 768, generator=g)
    with torch.no_grad():
        ref = O.transformer_block({"b." + k: v for k, v in sd.items()}, "b", x, enc, True, n_cam=n_cam, neighbors=pairs)
    out = blk.to("cuda:0")(x.cuda(), encoder_hidden_states=enc.cuda()).float().cpu()
    r = common.metrics(out, ref)
    print(f"block with {n_cam} views / {len(pairs[0])} neighbour(s) vs oracle:", r)
    assert r["cos"] >= 0.999 and r["rel_l2"] <= 2e-2, r


@pytest.mark.parametrize("mode,dim,T", [("concat", 320, 150), ("self", 320, 150), ("concat", 640, 91), ("self", 1280, 28), ("add", 320, 150)])
def test_block_neighboring_attn_types(mode, dim, T):
    """neighboring_attn_type (networks/blocks.py:112-140): "concat" puts the two neighbours' tokens under ONE softmax, "self"
    attends over all six views of the scene, "add" sums one attention per neighbour -- the same kernel with its K/V sources
    concatenated or summed, against the oracle (itself equal to the reference's class in all three modes)"""
    from dualdiff_b200.networks import BasicMultiviewTransformerBlock
    from oracle import dualdiff_oracle as O
    with torch.device("meta"):
        blk = BasicMultiviewTransformerBlock(dim, 8, dim // 8, cross_attention_dim=768, neighboring_view_pair=common.NEIGHBORS,
                                             neighboring_attn_type=mode)
    sd = _seeded(blk, 9)
    g = torch.Generator().manual_seed(4)
    x = torch.randn(12, T, dim, generator=g)
    enc = torch.randn(12, 83, 768, generator=g)
    with torch.no_grad():
        ref = O.transformer_block({"b." + k: v for k, v in sd.items()}, "b", x, enc, True, attn_type=mode)
        other = O.transformer_block({"b." + k: v for k, v in sd.items()}, "b", x, enc, True, attn_type="add" if mode != "add" else "concat")
    assert (ref - other).abs().max() > 1e-2 * ref.abs().max()          # the modes really differ on these inputs
    out = blk.to("cuda:0")(x.cuda(), encoder_hidden_states=enc.cuda()).float().cpu()
    r = common.metrics(out, ref)
    print(f"block with neighboring_attn_type={mode} (C={dim}) vs oracle:", r)
    assert r["cos"] >= 0.999 and r["rel_l2"] <= 2e-2, r


def test_unknown_neighboring_attn_type_raises():
    from dualdiff_b200.networks import BasicMultiviewTransformerBlock
    with pytest.raises(NotImplementedError, match="Unknown type"):
        with torch.device("meta"):
            BasicMultiviewTransformerBlock(320, 8, 40, cross_attention_dim=768, neighboring_view_pair=common.NEIGHBORS,
                                           neighboring_attn_type="mean")


def test_per_view_prompts_use_aug_text():
    """use_aug_text (configs/exp/occ_bg_augtext.yaml; unet_addon_rawbox.py:351-352, pipeline_bev_controlnet.py:250): one prompt
    per camera VIEW instead of one per scene -- a branch forward and a whole CFG noise prediction against the oracle"""
    from dualdiff_b200 import synthetic as S
    from dualdiff_b200.pipeline import DualDiffDenoiser
    from oracle import dualdiff_oracle as O
    unet, nets, sds = common.build_models()
    h, w, B = 8, 12, 1
    inp = S.make_inputs(B, h, w, seed=6, L_bg=5, L_fg=4)
    g = torch.Generator().manual_seed(8)
    inp["prompt_embeds"] = torch.randn(2 * B * 6, 77, 768, generator=g)          # uncond rows of all views first, then cond rows
    dev = torch.device("cuda")
    t = torch.tensor([500])
    with torch.no_grad():
        d_ref, m_ref, enc_ref = O.controlnet_forward(sds["bg"], inp["latents"], t, inp["camera_param"], inp["boxes_bg"],
                                                     inp["prompt_embeds"][6:], inp["cond_bg"], use_occ_3d=False)
        shared, _, _ = O.controlnet_forward(sds["bg"], inp["latents"], t, inp["camera_param"], inp["boxes_bg"],
                                            inp["prompt_embeds"][6:7], inp["cond_bg"], use_occ_3d=False)
    assert (d_ref[3] - shared[3]).abs().max() > 1e-3 * d_ref[3].abs().max()        # the views' prompts really differ
    net = nets[0].cuda()
    out = net(inp["latents"].to(dev), 500, camera_param=inp["camera_param"].to(dev), bboxes_3d_data=common.to_dev(inp["boxes_bg"], dev),
              encoder_hidden_states=inp["prompt_embeds"][6:].to(dev), controlnet_cond=inp["cond_bg"].to(dev), use_aug_text=True)
    r = common.metrics(out.mid_block_res_sample.float().cpu(), m_ref)
    print("branch with per-view prompts, mid residual vs oracle:", r)
    assert r["cos"] >= 0.999 and r["rel_l2"] <= 2e-2, r
    assert torch.equal(out.encoder_hidden_states_with_cam.float().cpu()[:, 1:78], enc_ref[:, 1:78].to(torch.bfloat16).float())
    with pytest.raises(ValueError, match="prompt embeddings"):
        net(inp["latents"].to(dev), 500, camera_param=inp["camera_param"].to(dev), bboxes_3d_data=common.to_dev(inp["boxes_bg"], dev),
            encoder_hidden_states=inp["prompt_embeds"][6:].to(dev), controlnet_cond=inp["cond_bg"].to(dev), use_aug_text=False)
    # the sampler's noise prediction with CFG
    den = DualDiffDenoiser(unet, nets, guidance_scale=2.0, use_cuda_graph=False)
    d = common.to_dev(inp, dev)
    den.prepare(d["latents"], d["prompt_embeds"], d["camera_param"], [d["boxes_bg"], d["boxes_fg"]], [d["cond_bg"], d["cond_fg"]],
                num_inference_steps=4)
    t0 = int(den.scheduler.timesteps[0])
    den.step(0)
    eps = den.eps_rows.float().cpu().reshape(2 * 6, h, w, 4).permute(0, 3, 1, 2)
    with torch.no_grad():
        ref = O.noise_prediction(sds["unet"], sds["bg"], sds["fg"], inp["latents"], t0, inp, 2.0, True)["eps_raw"]
    r = common.metrics(eps, ref)
    print("CFG noise prediction with per-view prompts vs oracle:", r)
    assert r["cos"] >= 0.999 and r["rel_l2"] <= 2e-2, r
