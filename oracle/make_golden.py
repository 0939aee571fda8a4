"""ORACLE / TEST INFRASTRUCTURE ONLY.  Generates tests/golden/*.pt by running the reference's OWN, unmodified
network classes (oracle/reference_model.py: /root/reference/MD_txt_con_fusion on top of oracle/shim) on the
seeded synthetic weights + inputs of dualdiff_b200/synthetic.py.  Run here (the GPU box has no /root/reference):

    python oracle/make_golden.py tiny      # seconds   -> tests/golden/step_tiny.pt
    python oracle/make_golden.py full      # minutes   -> tests/golden/step_full.pt  (config 1: B=1, 6 views, 28x50, CFG)
    python oracle/make_golden.py hd        # minutes   -> tests/golden/step_hd.pt    (config 4 geometry: B=1, 56x100, CFG)
    python oracle/make_golden.py rollout25 # ~10 min   -> tests/golden/rollout25_full.pt (config 2's sampler: 25 UniPC+CFG
                                           #               steps, B=1, 28x50, the reference classes in the loop)

The fixtures hold the reference outputs (noise prediction, mid residual, digests of the other tensors) and the
state-dict manifests digests, so a test can prove it regenerated identical weights before comparing.
"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dualdiff_b200 import synthetic as S  # noqa: E402
from oracle import dualdiff_oracle as O  # noqa: E402
from oracle import reference_model as RM  # noqa: E402
sys.path.insert(0, os.path.join(ROOT, "tests"))
from common import strided_sample  # noqa: E402

CONFIGS = {
    "tiny": dict(block_out=(320, 64, 64, 64), B=1, h=8, w=12, L_bg=5, L_fg=7, t=801),
    "full": dict(block_out=(320, 640, 1280, 1280), B=1, h=28, w=50, L_bg=28, L_fg=32, t=801),
    "hd": dict(block_out=(320, 640, 1280, 1280), B=1, h=56, w=100, L_bg=28, L_fg=32, t=801),
}
SAMPLE_N = 4096   # elements kept of every ControlNet residual (strided over the flattened NCHW tensor)
SEEDS = {"unet": 0, "bg": 1, "fg": 2}


def digest(t):
    t = t.double()
    return torch.stack([t.sum(), t.abs().sum(), (t * t).sum()]).float()


def reference_noise_prediction(nets, inp, t, B, guidance_scale=2.0):
    """pipeline_bev_controlnet.py:381-492 executed with the reference's own modules"""
    unet, bg, fg = nets
    lat = inp["latents"]
    lat_in = torch.cat([lat] * 2)
    tt = torch.full((2 * B,), t, dtype=torch.int64)
    kw = bg.add_uncond_to_kwargs(camera_param=inp["camera_param"],
                                 bboxes_3d_data=[inp["boxes_bg"], inp["boxes_fg"]], image=None, max_len=None)
    cam, boxes = kw["camera_param"], kw["bboxes_3d_data"]
    images = [torch.cat([inp["cond_bg"]] * 2), torch.cat([inp["cond_fg"]] * 2)]
    outs = []
    for i, net in enumerate((bg, fg)):
        outs.append(net(lat_in, tt, cam, encoder_hidden_states=inp["prompt_embeds"], controlnet_cond=images[i],
                        conditioning_scale=1.0, guess_mode=False, return_dict=False, bboxes_3d_data=boxes[i],
                        use_aug_text=False))
    down = [a + b for a, b in zip(outs[0][0], outs[1][0])]
    mid = outs[0][1] + outs[1][1]
    enc = outs[0][2]
    x = lat_in.reshape(-1, *lat_in.shape[2:])
    eps_raw = unet(x, torch.tensor(t), encoder_hidden_states=enc, down_block_additional_residuals=down,
                   mid_block_additional_residual=mid).sample
    e_u, e_c = eps_raw.chunk(2)
    eps = e_u + guidance_scale * (e_c - e_u)
    return dict(eps_raw=eps_raw, eps=eps, down=down, mid=mid, enc=enc, branch_down=[o[0] for o in outs],
                branch_mid=[o[1] for o in outs])


def main(which):
    c = CONFIGS[which]
    torch.manual_seed(0)
    t0 = time.time()
    nets = (RM.build_unet(c["block_out"]), RM.build_branch(False, c["block_out"]), RM.build_branch(True, c["block_out"]))
    sds, manifests = {}, {}
    for name, m in zip(("unet", "bg", "fg"), nets):
        man = S.manifest_of(m)
        sd = S.init_state_dict(man, SEEDS[name])
        m.load_state_dict(sd, strict=True)
        sds[name], manifests[name] = sd, man
    inp = S.make_inputs(c["B"], c["h"], c["w"], seed=1, L_bg=c["L_bg"], L_fg=c["L_fg"])
    print(f"[{which}] built reference modules in {time.time() - t0:.1f}s")
    t0 = time.time()
    with torch.no_grad():
        ref = reference_noise_prediction(nets, inp, c["t"], c["B"])
    t_ref = time.time() - t0
    t0 = time.time()
    with torch.no_grad():
        orc = O.noise_prediction(sds["unet"], sds["bg"], sds["fg"], inp["latents"], c["t"], inp, 2.0, True)
    t_orc = time.time() - t0
    rel = lambda a, b: ((a - b).abs().max() / b.abs().max()).item()
    print(f"[{which}] reference {t_ref:.1f}s oracle {t_orc:.1f}s | oracle vs reference: eps_raw {rel(orc['eps_raw'], ref['eps_raw']):.2e} "
          f"mid {rel(orc['mid'], ref['mid']):.2e} enc {rel(orc['enc'], ref['enc']):.2e}")
    fix = {
        "config": {k: (list(v) if isinstance(v, tuple) else v) for k, v in c.items()},
        "manifest_digest": {k: S.manifest_digest(v) for k, v in manifests.items()},
        "manifest": {k: {kk: list(vv) for kk, vv in v.items()} for k, v in manifests.items()},
        "eps_raw": ref["eps_raw"].clone(), "eps": ref["eps"].clone(), "mid": ref["mid"].clone(),
        "down_digest": torch.stack([digest(d) for d in ref["down"]]), "enc_digest": digest(ref["enc"]),
        # all 13 ControlNet residuals: strided samples of the 12 summed down residuals and of each
        # branch's own 12 + mid, so a wrong zero-conv scale of a low-energy skip cannot hide inside the eps tolerance
        "down_sample": [strided_sample(d, SAMPLE_N) for d in ref["down"]],
        "branch_down_sample": [[strided_sample(d, SAMPLE_N) for d in br] for br in ref["branch_down"]],
        "branch_mid": [strided_sample(m, 16384) for m in ref["branch_mid"]], "branch_mid_n": 16384,
        "sample_n": SAMPLE_N,
        "oracle_vs_reference": {"eps_raw": rel(orc["eps_raw"], ref["eps_raw"]), "mid": rel(orc["mid"], ref["mid"])},
        "generator": "oracle/make_golden.py (reference classes from /root/reference on oracle/shim)",
        "torch": str(torch.__version__),
    }
    if which == "tiny":
        # 4-step UniPC + CFG rollout with the reference modules in the loop (scheduler = oracle.UniPC: diffusers is
        # not vendored, so the scheduler arithmetic itself is "parity unpinned" — Appendix A.3 restated)
        sch = O.UniPC()
        sch.set_timesteps(4)
        lat = inp["latents"].clone()
        with torch.no_grad():
            for t in sch.timesteps:
                cur = dict(inp)
                cur["latents"] = lat
                r = reference_noise_prediction(nets, cur, int(t), c["B"])
                flat = lat.reshape(-1, *lat.shape[2:])
                lat = sch.step(r["eps"], int(t), flat).reshape(lat.shape)
        fix["rollout4_latents"] = lat.clone()
        fix["rollout4_timesteps"] = sch.timesteps.clone()
    if which in ("full", "hd"):
        del fix["manifest"]  # large; the digest is enough
    if which == "hd":        # 12 x 4 x 56 x 100 fp32 = 1 MB: keep the guided prediction and samples of the rest
        fix["eps_raw_sample"] = strided_sample(fix.pop("eps_raw"), 65536)
        fix["mid_sample"] = strided_sample(fix.pop("mid"), 65536)
    out = os.path.join(ROOT, "tests", "golden", f"step_{which}.pt")
    torch.save(fix, out)
    print(f"[{which}] wrote {out} ({os.path.getsize(out) / 1024:.0f} KiB)")


def rollout25(n_steps=25, keep=(1, 2, 5, 10, 15, 20, 25)):
    """BASELINE configs[1]'s sampler at B=1: n_steps UniPC(bh2, order 2) + CFG steps at 28x50 with the reference's own
    network classes producing every noise prediction (pipeline_bev_controlnet.py:378-512).  The scheduler arithmetic is
    oracle.UniPC (diffusers is not vendored: Appendix A.3 restated, parity unpinned against real diffusers)."""
    c = CONFIGS["full"]
    torch.manual_seed(0)
    nets = (RM.build_unet(c["block_out"]), RM.build_branch(False, c["block_out"]), RM.build_branch(True, c["block_out"]))
    manifests = {}
    for name, m in zip(("unet", "bg", "fg"), nets):
        man = S.manifest_of(m)
        m.load_state_dict(S.init_state_dict(man, SEEDS[name]), strict=True)
        manifests[name] = man
    inp = S.make_inputs(c["B"], c["h"], c["w"], seed=1, L_bg=c["L_bg"], L_fg=c["L_fg"])
    sch = O.UniPC()
    sch.set_timesteps(n_steps)
    lat = inp["latents"].clone()
    kept = {}
    t0 = time.time()
    with torch.no_grad():
        for i, t in enumerate(sch.timesteps):
            cur = dict(inp)
            cur["latents"] = lat
            r = reference_noise_prediction(nets, cur, int(t), c["B"])
            flat = lat.reshape(-1, *lat.shape[2:])
            lat = sch.step(r["eps"], int(t), flat).reshape(lat.shape)
            if i + 1 in keep:
                kept[i + 1] = lat.clone()
            print(f"[rollout25] step {i + 1}/{n_steps} t={int(t)} |lat| {lat.norm():.3f} ({time.time() - t0:.0f}s)", flush=True)
    fix = {
        "config": dict(B=c["B"], h=c["h"], w=c["w"], L_bg=c["L_bg"], L_fg=c["L_fg"], n_steps=n_steps, guidance_scale=2.0,
                       scheduler="UniPC(bh2, order 2, lower_order_final)"),
        "manifest_digest": {k: S.manifest_digest(v) for k, v in manifests.items()},
        "timesteps": sch.timesteps.clone(), "latents_at": kept, "final_latents": lat.clone(),
        "generator": "oracle/make_golden.py rollout25 (reference classes from /root/reference on oracle/shim)",
        "torch": str(torch.__version__),
    }
    out = os.path.join(ROOT, "tests", "golden", "rollout25_full.pt")
    torch.save(fix, out)
    print(f"[rollout25] wrote {out} ({os.path.getsize(out) / 1024:.0f} KiB)")


if __name__ == "__main__":
    for w in (sys.argv[1:] or ["tiny"]):
        rollout25() if w == "rollout25" else main(w)
