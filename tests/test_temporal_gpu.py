"""Temporal attention of the video configuration (BASELINE config 5) on the GPU, against oracle/dualdiff_oracle.py.
The reference has no temporal block: the oracle function IS the definition (parity unpinned, stated in DESIGN.md).
Tolerance: bf16 output, max-abs <= 2e-2 * max|ref|."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _rel(out, ref):
    return ((out.float().cpu() - ref).abs().max() / ref.abs().max().clamp_min(1e-6)).item()


@pytest.mark.parametrize("d,F,n_clip,V,T,split", [(40, 16, 1, 6, 70, 1), (80, 8, 2, 6, 35, 1), (160, 4, 2, 3, 28, 1),
                                                  (40, 16, 1, 6, 33, 4), (80, 6, 2, 2, 20, 3)])
def test_temporal_attention_kernel(d, F, n_clip, V, T, split):
    """split > 1: K/V are laid out as `split` rank blocks of F/split frames (the all-gather layout of FrameShard) and the
    queries are the frames of block `r`; every block's output must equal the unsharded result."""
    from dualdiff_b200 import ops
    from oracle.dualdiff_oracle import mha
    heads, C = 8, 8 * d
    g = torch.Generator().manual_seed(d + F)
    q, k, v = ((torch.randn(n_clip * F * V, T, C, generator=g) * 0.7).to(torch.bfloat16) for _ in range(3))

    def seq(t):
        return t.float().reshape(n_clip, F, V, T, C).permute(0, 2, 3, 1, 4).reshape(n_clip * V * T, F, C)
    ref = mha(seq(q), seq(k), seq(v), heads).reshape(n_clip, V, T, F, C).permute(0, 3, 1, 2, 4).reshape(n_clip * F * V, T, C)
    f_loc = F // split

    def block(t, r):   # frames [r*f_loc, (r+1)*f_loc) of every clip, image order (clip, frame, view)
        return t.reshape(n_clip, F, V, T, C)[:, r * f_loc:(r + 1) * f_loc].reshape(n_clip * f_loc * V * T, C)
    kg = torch.cat([block(k, r) for r in range(split)]).cuda()
    vg = torch.cat([block(v, r) for r in range(split)]).cuda()
    for r in range(split):
        out = ops.temporal_attention(block(q, r).cuda(), kg, vg, n_outer=n_clip, n_view=V, tokens=T, heads=heads, head_dim=d,
                                     frames_q=f_loc, frames_kv=F, frames_per_rank=f_loc, kv_rank_stride=n_clip * f_loc * V,
                                     q_hs=d, k_hs=d, v_hs=d)
        want = block(ref, r)
        assert _rel(out, want) < 2e-2, (r, _rel(out, want))
    if split > 1:
        # the token-sharded layout (FrameShard.to_token_shards): the QUERIES and the output are in rank blocks as well, all
        # frames in one launch
        qg = torch.cat([block(q, r) for r in range(split)]).cuda()
        out = ops.temporal_attention(qg, kg, vg, n_outer=n_clip, n_view=V, tokens=T, heads=heads, head_dim=d, frames_q=F,
                                     frames_kv=F, frames_per_rank=f_loc, kv_rank_stride=n_clip * f_loc * V,
                                     frames_q_per_rank=f_loc, q_rank_stride=n_clip * f_loc * V, q_hs=d, k_hs=d, v_hs=d)
        want = torch.cat([block(ref, r) for r in range(split)])
        assert _rel(out, want) < 2e-2, _rel(out, want)


def test_multiview_block_with_temporal_attention_matches_oracle():
    from dualdiff_b200 import synthetic as S
    from dualdiff_b200.networks import BasicMultiviewTransformerBlock
    from oracle import dualdiff_oracle as O
    nb = {0: [5, 1], 1: [0, 2], 2: [1, 3], 3: [2, 4], 4: [3, 5], 5: [4, 0]}
    F, n_clip, T = 4, 1, 96
    with torch.device("meta"):
        blk = BasicMultiviewTransformerBlock(320, 8, 40, cross_attention_dim=768, neighboring_view_pair=nb, temporal_frames=F)
    sd = S.init_state_dict(S.manifest_of(blk), seed=11)          # zero-initialised modules are re-randomised
    assert sd["attn_temp.to_out.0.weight"].abs().max() > 0
    blk.load_state_dict(sd, strict=True, assign=True)
    g = torch.Generator().manual_seed(5)
    n = n_clip * F * 6
    x = torch.randn(n, T, 320, generator=g)
    enc = torch.randn(n, 83, 768, generator=g)
    with torch.no_grad():
        ref = O.transformer_block({"b." + k: v for k, v in sd.items()}, "b", x, enc, True, n_frames=F)
        base = O.transformer_block({"b." + k: v for k, v in sd.items()}, "b", x, enc, True, n_frames=1)
    assert (ref - base).abs().max() > 1e-2 * ref.abs().max()     # the temporal step really contributes
    out = blk.to("cuda:0")(x.cuda(), encoder_hidden_states=enc.cuda()).float().cpu()
    cos = torch.nn.functional.cosine_similarity(out.flatten(), ref.flatten(), dim=0).item()
    rel = ((out - ref).norm() / ref.norm()).item()
    assert cos > 0.999 and rel < 2e-2, (cos, rel)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (run under gpurun --gpus 2)")
def test_frame_sharded_block_matches_single_gpu():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29537", os.path.join(ROOT, "tests", "run_frameshard.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=420)
    assert "FRAMESHARD OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]


def test_video_unet_forward_matches_oracle():
    """UNet2DConditionModelMultiview(temporal_frames=2): 1 clip x 2 frames x 6 views at 16x24 latents through the drop-in
    forward against oracle.unet_forward(n_frames=2) (fp32 CPU).  Tolerance: cosine >= 0.999, rel-L2 <= 2e-2."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import common
    from dualdiff_b200 import synthetic as S
    from dualdiff_b200.networks import UNet2DConditionModelMultiview
    from oracle import dualdiff_oracle as O
    F = 2
    with torch.device("meta"):
        unet = UNet2DConditionModelMultiview(cross_attention_dim=768, neighboring_view_pair=common.NEIGHBORS, temporal_frames=F)
    sd = S.init_state_dict(S.manifest_of(unet), seed=0)
    assert any(".attn_temp." in k for k in sd)
    unet.load_state_dict(sd, strict=True, assign=True)
    g = torch.Generator().manual_seed(9)
    n, h, w = F * 6, 16, 24
    x = torch.randn(n, 4, h, w, generator=g)
    enc = torch.randn(n, 83, 768, generator=g)
    with torch.no_grad():
        ref = O.unet_forward(sd, x, 500, enc, n_frames=F)
        base = O.unet_forward(sd, x, 500, enc, n_frames=1)
    assert (ref - base).abs().max() > 1e-3 * ref.abs().max()
    out = unet.to("cuda:0")(x.cuda(), 500, enc.cuda()).sample.float().cpu()
    m = common.metrics(out, ref)
    assert m["cos"] > 0.999 and m["rel_l2"] < 2e-2, m
