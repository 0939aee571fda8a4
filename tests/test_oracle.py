"""CPU tests (no GPU): the oracle against the golden vectors produced by the reference's own classes, shim
self-checks, the fused scheduler's coefficient tables against the oracle's UniPC/DDIM restatement, and — when
/root/reference is present (this container) — the oracle against the reference's classes executed live."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import common  # noqa: E402
from dualdiff_b200 import scheduler as SCH  # noqa: E402
from dualdiff_b200 import synthetic as S  # noqa: E402
from oracle import dualdiff_oracle as O  # noqa: E402

ROOT = common.ROOT


@pytest.fixture(scope="module")
def tiny():
    fix = torch.load(os.path.join(common.GOLDEN, "step_tiny.pt"))
    man = {k: {kk: tuple(vv) for kk, vv in v.items()} for k, v in fix["manifest"].items()}
    sds = {k: S.init_state_dict(man[k], common.SEEDS[k]) for k in man}
    c = fix["config"]
    inp = S.make_inputs(c["B"], c["h"], c["w"], seed=1, L_bg=c["L_bg"], L_fg=c["L_fg"])
    return fix, sds, inp, c


def test_oracle_reproduces_reference_golden(tiny):
    """golden = outputs of the reference's unmodified networks/*.py on the shim (oracle/make_golden.py)"""
    fix, sds, inp, c = tiny
    with torch.no_grad():
        out = O.noise_prediction(sds["unet"], sds["bg"], sds["fg"], inp["latents"], c["t"], inp, 2.0, True)
    for k in ("eps_raw", "eps", "mid"):
        rel = ((out[k] - fix[k]).abs().max() / fix[k].abs().max()).item()
        assert rel < 2e-5, (k, rel)   # fp32; the cross-view path is algebraically rearranged (2 b_o, single projection)
    dig = torch.stack([torch.stack([d.double().sum(), d.double().abs().sum(), (d.double() ** 2).sum()]).float() for d in out["down"]])
    assert torch.allclose(dig, fix["down_digest"], rtol=1e-4)


def test_oracle_rollout_matches_reference_rollout(tiny):
    fix, sds, inp, c = tiny
    sch = O.UniPC()
    sch.set_timesteps(4)
    assert torch.equal(sch.timesteps, fix["rollout4_timesteps"])
    lat = inp["latents"].clone()
    with torch.no_grad():
        for t in sch.timesteps:
            cur = dict(inp)
            cur["latents"] = lat
            lat, _ = O.denoise_step(sds["unet"], sds["bg"], sds["fg"], sch, lat, int(t), cur, 2.0, True)
    rel = ((lat - fix["rollout4_latents"]).abs().max() / fix["rollout4_latents"].abs().max()).item()
    assert rel < 1e-4, rel


def test_golden_full_fixture_is_consistent():
    fix = torch.load(os.path.join(common.GOLDEN, "step_full.pt"))
    assert fix["config"]["block_out"] == [320, 640, 1280, 1280] and fix["eps_raw"].shape == (12, 4, 28, 50)
    assert fix["oracle_vs_reference"]["eps_raw"] < 1e-4
    e_u, e_c = fix["eps_raw"].chunk(2)
    assert torch.allclose(e_u + 2.0 * (e_c - e_u), fix["eps"], atol=1e-5)


def test_product_modules_have_reference_state_dict_layout():
    """key names + shapes of the drop-in mirrors == those of the reference classes (digest recorded in the golden)"""
    fix = torch.load(os.path.join(common.GOLDEN, "step_full.pt"))
    unet, nets, _ = common.build_models(load=False)
    for name, m in (("unet", unet), ("bg", nets[0]), ("fg", nets[1])):
        assert S.manifest_digest(S.manifest_of(m)) == fix["manifest_digest"][name], name
    assert sum(p.numel() for p in unet.parameters()) == 921_522_884  # SDv1.5 859,520,964 + 16 x (attn4, norm4, connector)


def test_shim_sd15_parameter_count_and_primitives():
    sys.path.insert(0, os.path.join(ROOT, "oracle", "shim"))
    from diffusers import UNet2DConditionModel
    from diffusers.models.attention_processor import Attention
    with torch.device("meta"):
        u = UNet2DConditionModel(cross_attention_dim=768)
    assert sum(p.numel() for p in u.parameters()) == 859_520_964
    torch.manual_seed(0)
    attn = Attention(64, cross_attention_dim=48, heads=4, dim_head=16)
    x, ctx = torch.randn(2, 10, 64), torch.randn(2, 7, 48)
    q = attn.to_q(x).reshape(2, 10, 4, 16).transpose(1, 2)
    k = attn.to_k(ctx).reshape(2, 7, 4, 16).transpose(1, 2)
    v = attn.to_v(ctx).reshape(2, 7, 4, 16).transpose(1, 2)
    ref = attn.to_out[0](F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(2, 10, 64))
    assert torch.allclose(attn(x, encoder_hidden_states=ctx), ref, atol=1e-5)
    qh = torch.randn(3, 9, 8, 40); kh = torch.randn(3, 5, 8, 40); vh = torch.randn(3, 5, 8, 40)
    import xformers.ops
    o = xformers.ops.memory_efficient_attention(qh, kh, vh, scale=40 ** -0.5)
    ref = F.scaled_dot_product_attention(qh.transpose(1, 2), kh.transpose(1, 2), vh.transpose(1, 2)).transpose(1, 2)
    assert torch.allclose(o, ref, atol=1e-5)


@pytest.mark.parametrize("n", [4, 20, 25])
@pytest.mark.parametrize("kind", ["unipc", "ddim"])
def test_fused_scheduler_coefficients_equal_oracle_scheduler(n, kind):
    """dd_cfg_sched_step applies x' = b_xc*xc + b_x0*x0 + b_m0*m0 etc.; the tables must reproduce the multistep update"""
    if kind == "unipc":
        sch = O.UniPC(); ts = SCH.unipc_timesteps(n); coef = SCH.unipc_coefficients(ts)
    else:
        sch = O.DDIM(); ts = SCH.ddim_timesteps(n); coef = SCH.ddim_coefficients(ts)
    sch.set_timesteps(n)
    assert (np.asarray(sch.timesteps) == ts).all()
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 4, 8, 12, generator=g, dtype=torch.float64)
    xo = x.clone().float()
    last = torch.zeros_like(x); m0 = torch.zeros_like(x); m1 = torch.zeros_like(x)
    for i, t in enumerate(ts):
        e = torch.randn(2, 4, 8, 12, generator=g, dtype=torch.float64)
        xo = sch.step(e.float(), int(t), xo)
        c = coef[i]
        x0 = (x - c[1] * e) * c[2]
        xc = c[7] * x + c[3] * last + c[4] * m0 + c[5] * m1 + c[6] * x0
        xn = c[8] * xc + c[9] * x0 + c[10] * m0
        last, m1, m0, x = xc, m0, x0, xn
        assert ((x.float() - xo).abs().max() / xo.abs().max()).item() < 1e-5


def test_synthetic_weights_are_deterministic_and_live():
    a = S.init_tensor("down_blocks.0.resnets.0.conv1.weight", (8, 4, 3, 3), 0)
    b = S.init_tensor("down_blocks.0.resnets.0.conv1.weight", (8, 4, 3, 3), 0)
    assert torch.equal(a, b)
    for key in ("controlnet_down_blocks.3.weight", "controlnet_cond_embedding.conv_out.weight",
                "down_blocks.0.attentions.0.transformer_blocks.0.connector.weight"):
        assert S.init_tensor(key, (16, 16), 0).abs().max() > 0   # zero-init modules are re-randomised (non-vacuous parity)


@pytest.mark.skipif(not os.path.isdir("/root/reference/MD_txt_con_fusion"), reason="reference tree only exists in the build container")
def test_oracle_equals_reference_classes_live(tiny):
    from oracle import make_golden as MG
    from oracle import reference_model as RM
    fix, sds, inp, c = tiny
    nets = (RM.build_unet(c["block_out"]), RM.build_branch(False, c["block_out"]), RM.build_branch(True, c["block_out"]))
    for name, m in zip(("unet", "bg", "fg"), nets):
        m.load_state_dict(sds[name], strict=True)
    with torch.no_grad():
        ref = MG.reference_noise_prediction(nets, inp, c["t"], c["B"])
        out = O.noise_prediction(sds["unet"], sds["bg"], sds["fg"], inp["latents"], c["t"], inp, 2.0, True)
    assert torch.equal(out["enc"], ref["enc"]) and torch.equal(out["mid"], ref["mid"])
    assert ((out["eps_raw"] - ref["eps_raw"]).abs().max() / ref["eps_raw"].abs().max()).item() < 2e-5


def test_from_pretrained_save_pretrained_roundtrip(tmp_path):
    """diffusers-format checkpoint directories load into the drop-in mirrors (misc/test_utils.py:111-113,146-147)"""
    from dualdiff_b200.networks import BasicMultiviewTransformerBlock, BEVControlNetModel
    torch.manual_seed(0)
    cfg = dict(common.CONTROLNET_CONFIG)
    with torch.device("meta"):
        net = BEVControlNetModel(**cfg)
    net.adm_proj = None
    net.txt_con_fusionp = None
    sd = S.init_state_dict(S.manifest_of(net), 3)
    net.load_state_dict(sd, strict=True, assign=True)
    net.save_pretrained(str(tmp_path / "controlnet_bg_1"))
    back = BEVControlNetModel.from_pretrained(str(tmp_path / "controlnet_bg_1"), torch_dtype=torch.float16,
                                              low_cpu_mem_usage=False, device_map=None, ignore_mismatched_sizes=True)
    assert back.config.cross_attention_dim == 768 and not back.training
    bsd = back.state_dict()
    for k, v in sd.items():
        assert torch.equal(bsd[k], v.to(torch.float16)), k
    # keys that exist in the module but not in the checkpoint (adm_proj, txt_con_fusionp: "created by default,
    # deleted by the caller") keep their init, as with diffusers' loader
    assert any(k.startswith("adm_proj") for k in bsd)


def _sample_gaussian_data(kind, n, s2=0.25):
    """run the fused scheduler's coefficient tables on data ~ N(0, s2), for which the optimal noise prediction
    eps*(x, t) = sigma_t x / (alpha_t^2 s2 + sigma_t^2) and the probability-flow ODE solution are analytic"""
    _, alpha, sigma, _ = SCH.sd_schedule()
    if kind == "unipc":
        ts = SCH.unipc_timesteps(n); coef = SCH.unipc_coefficients(ts)
    else:
        ts = SCH.ddim_timesteps(n); coef = SCH.ddim_coefficients(ts)
    x = np.array([1.3, -0.4, 2.2])
    x_start = x.copy()
    last = m0 = m1 = np.zeros_like(x)
    for i, t in enumerate(ts):
        c = coef[i]
        e = sigma[t] * x / (alpha[t] ** 2 * s2 + sigma[t] ** 2)
        x0 = (x - c[1] * e) * c[2]
        xc = c[7] * x + c[3] * last + c[4] * m0 + c[5] * m1 + c[6] * x0
        xn = c[8] * xc + c[9] * x0 + c[10] * m0
        last, m1, m0, x = xc, m0, x0, xn
    t0 = ts[0]
    exact = x_start * np.sqrt((alpha[0] ** 2 * s2 + sigma[0] ** 2) / (alpha[t0] ** 2 * s2 + sigma[t0] ** 2))
    return float(np.abs(x - exact).max() / np.abs(exact).max())


def test_samplers_converge_to_the_analytic_ode_solution():
    """diffusers' UniPC / DDIM are not vendored (parity unpinned), so the restated update rules are also checked against
    mathematics: on Gaussian data both samplers must converge to the exact probability-flow solution as the step count
    grows, and the second-order UniPC must beat first-order DDIM at every step count"""
    err = {k: [_sample_gaussian_data(k, n) for n in (10, 20, 40, 80, 160)] for k in ("unipc", "ddim")}
    for k in err:
        assert all(b < 0.62 * a for a, b in zip(err[k], err[k][1:])), (k, err[k])     # halving the step at least ~halves the error
    assert all(u < d for u, d in zip(err["unipc"], err["ddim"])), err
    assert err["unipc"][-1] < 5e-3 and err["ddim"][-1] < 2e-2, err


def test_patch_weight_packing_reproduces_the_convolution():
    """host half of the conv_in path (engine.conv_in): [patch matrix] x [pack_conv3x3_patch weights]^T == F.conv2d"""
    from dualdiff_b200.packing import pack_conv3x3_patch
    g = torch.Generator().manual_seed(0)
    w, b = torch.randn(16, 4, 3, 3, generator=g), torch.randn(16, generator=g)
    x = torch.randn(3, 4, 5, 7, generator=g)
    wp = pack_conv3x3_patch(w)
    assert wp.shape == (16, 40) and wp.dtype == torch.bfloat16 and (wp[:, 36:] == 0).all()
    n, c, h, wd = x.shape
    cols = torch.nn.functional.unfold(x, 3, padding=1).reshape(n, c, 9, h * wd).permute(0, 3, 2, 1).reshape(n * h * wd, 9 * c)
    rows = x.new_zeros((n * h * wd, 40))
    rows[:, :36] = cols
    out = rows @ wp.float().T + b
    ref = torch.nn.functional.conv2d(x, w.to(torch.bfloat16).float(), b, padding=1).permute(0, 2, 3, 1).reshape(-1, 16)
    assert (out - ref).abs().max() < 1e-4


def _sfa_plus_case():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import make_golden_sfa_plus as MG
    from dualdiff_b200.networks.txt_con_fusion import txt_con_XFormersAttn_plus
    fix = torch.load(os.path.join(common.GOLDEN, "sfa_plus_small.pt"))
    with torch.device("meta"):
        m = txt_con_XFormersAttn_plus()
    sd = S.init_state_dict(S.manifest_of(m), fix["weight_seed"])
    cond, txt = MG.inputs(fix["input_seed"])
    assert tuple(cond.shape) == tuple(fix["shape"])
    return fix, sd, cond, txt


def test_sfa_plus_oracle_matches_reference_golden():
    """txt_con_XFormersAttn_plus (txt_con_fusion.py:184-337): the oracle restatement reproduces the output of the reference's
    own class (golden made by oracle/make_golden_sfa_plus.py)"""
    fix, sd, cond, txt = _sfa_plus_case()
    with torch.no_grad():
        out = O.sfa_plus({"txt_con_fusionp." + k: v for k, v in sd.items()}, cond, txt)
    assert ((out - fix["out"]).abs().max() / fix["out"].abs().max()).item() < 1e-5


@pytest.mark.skipif(not os.path.isdir("/root/reference/MD_txt_con_fusion"), reason="reference tree only exists in the build container")
def test_sfa_plus_oracle_matches_reference_class_live():
    from oracle import reference_model as RM
    RM._paths()
    from magicdrive.networks.txt_con_fusion import txt_con_XFormersAttn_plus as Ref
    fix, sd, cond, txt = _sfa_plus_case()
    ref = Ref()
    ref.load_state_dict(sd, strict=True)
    with torch.no_grad():
        want = ref(None, cond, encoder_hidden_states=txt)
        out = O.sfa_plus({"txt_con_fusionp." + k: v for k, v in sd.items()}, cond, txt)
    assert torch.equal(want, fix["out"])
    assert ((out - want).abs().max() / want.abs().max()).item() < 1e-5


@pytest.mark.skipif(not os.path.isdir("/root/reference/MD_txt_con_fusion"), reason="reference tree only exists in the build container")
@pytest.mark.parametrize("mode", ["add", "concat", "self"])
def test_oracle_block_matches_reference_class_for_every_neighboring_attn_type(mode):
    """BasicMultiviewTransformerBlock of the reference (networks/blocks.py, unmodified, on the shim) in the three
    neighboring_attn_type modes against oracle.transformer_block(attn_type=...)"""
    from oracle import reference_model as RM
    RM._paths()
    from magicdrive.networks.blocks import BasicMultiviewTransformerBlock as Ref
    ref = Ref(64, 8, 8, cross_attention_dim=768, neighboring_view_pair=common.NEIGHBORS, neighboring_attn_type=mode,
              zero_module_type="zero_linear").eval()
    sd = S.init_state_dict({k: tuple(v.shape) for k, v in ref.state_dict().items()}, seed=3)
    ref.load_state_dict(sd, strict=True)
    g = torch.Generator().manual_seed(1)
    x, enc = torch.randn(12, 10, 64, generator=g), torch.randn(12, 9, 768, generator=g)
    with torch.no_grad():
        want = ref(x, encoder_hidden_states=enc)
        got = O.transformer_block({"b." + k: v for k, v in sd.items()}, "b", x, enc, True, attn_type=mode)
    assert ((want - got).abs().max() / want.abs().max()).item() < 1e-6
