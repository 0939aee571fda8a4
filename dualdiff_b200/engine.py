"""The CUDA execution engine of the DualDiff denoising step.

`pack_*` turn a diffusers-keyed state dict into kernel-layout buffers once (bf16 GEMM operands, fp32
biases / norm parameters); `unet_forward` / `controlnet_forward` then run the reference's arithmetic
(networks/unet_2d_condition_multiview.py:327-527, networks/unet_addon_rawbox.py:794-1082,
networks/blocks.py:144-238) as a sequence of C-ABI kernel launches on the current stream — no torch math on
the hot path, no host synchronisation, CUDA-graph capturable.

Activations are bf16 channels-last "compact" rows [n_img*H*W, C]; only the inputs of 3x3 convolutions take
the zero-haloed padded-pixel layout (written directly by the GroupNorm+SiLU kernel).

Algebraic fusions (exact in real arithmetic; the oracle checks them):
  * cross-view attention: Q/K/V projected once per view (the reference projects every view twice, once per
    (view, neighbour) pair) and  connector(W_o(A_l + A_r) + 2 b_o) = (W_c W_o)(A_l + A_r) + (2 W_c b_o + b_c)
    is ONE GEMM with pre-multiplied weights (blocks.py:203-222);
  * timestep-invariant work is hoisted into `prepare_*`: text/box/camera tokens, the K/V projections of every
    text cross-attention, the condition embedding and Semantic Fusion Attention.
"""
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import torch

from . import ops
from .packing import pack_conv1x1, pack_conv3x3, pack_conv3x3_patch, pack_geglu, pack_linear

HEADS = 8
NEIGHBORS = {0: [5, 1], 1: [0, 2], 2: [1, 3], 3: [2, 4], 4: [3, 5], 5: [4, 0]}  # configs/dataset/Nuscenes.yaml:27-33 (default)


XVIEW_MODES = ("add", "concat", "self")      # neighboring_attn_type (networks/blocks.py:112-140)


def xview_plan(pairs, mode: str = "add"):
    """(pairs, K/V sources per view, concatenated softmax?, how often to_out's bias is counted) of a cross-view attention:
      add    -- one attention per (view, neighbour) pair, outputs summed AFTER to_out: n_nbr sources, bias n_nbr times
      concat -- the neighbours' tokens form one key sequence: n_nbr sources under one softmax, bias once
      self   -- every view attends to all views of its scene (blocks.py:134-137): n_cam sources under one softmax, bias once"""
    if mode not in XVIEW_MODES:
        raise NotImplementedError(f"Unknown type: {mode}")       # the reference's own message (blocks.py:139-140)
    pairs, n_nbr = check_view_pairs(pairs)
    if mode == "add":
        return pairs, n_nbr, False, n_nbr
    if mode == "concat":
        return pairs, n_nbr, True, 1
    if len(pairs) > 8:
        raise NotImplementedError("neighboring_attn_type='self' with more than 8 views (dd_attention takes up to 8 K/V sources)")
    return pairs, len(pairs), True, 1


def check_view_pairs(pairs):
    """the cross-view topology the kernels can run: every view lists the same number (1 or 2) of neighbour views, all inside
    the view range.  Returns (pairs with int keys, neighbours per view).  Anything else raises instead of silently computing
    the default ring (networks/blocks.py:106-121 builds the (view, neighbour) pairs from this table)."""
    if pairs is None:
        pairs = NEIGHBORS
    pairs = {int(k): [int(x) for x in v] for k, v in pairs.items()}
    n_cam = len(pairs)
    counts = {len(v) for v in pairs.values()}
    if sorted(pairs) != list(range(n_cam)) or len(counts) != 1 or next(iter(counts)) not in (1, 2) or \
            any(not 0 <= x < n_cam for v in pairs.values() for x in v):
        raise NotImplementedError(f"neighboring_view_pair {pairs}: the cross-view attention kernel needs views 0..n-1 with the "
                                  "same number (1 or 2) of neighbours each")
    return pairs, next(iter(counts))
BF = torch.bfloat16


def _dp(d):
    """Q/K head stride: head_dim 40 is zero-padded to 48 so the UMMA K extent is a multiple of 16"""
    return 48 if d == 40 else d


def pad_heads(w, d, dp):
    """[heads*d, in] -> [heads*dp, in] with zero rows appended per head"""
    if d == dp:
        return w
    h = w.shape[0] // d
    out = w.new_zeros((h, dp, w.shape[1]))
    out[:, :d] = w.reshape(h, d, w.shape[1])
    return out.reshape(h * dp, w.shape[1])


def pad_v_ones(w, d, dp):
    """value projection of head_dim-40 heads for the attention kernel's `v_ones` layout: [heads*d, in] -> weights
    [heads*dp, in] with zero rows appended per head, and a bias [heads*dp] that is 1.0 in column d of every head (0 elsewhere).
    The softmax denominator then falls out of P V as column d of the accumulator (csrc/dd_attention.cu, ONES)."""
    h = w.shape[0] // d
    wp = pad_heads(w, d, dp)
    b = torch.zeros(h, dp)
    b[:, d] = 1.0
    return wp, b.reshape(-1)


# ---------------------------------------------------------------------------------------------------
# packing
# ---------------------------------------------------------------------------------------------------
class Packer:
    def __init__(self, sd: Dict[str, torch.Tensor], device, n_nbr: int = 2):
        self.sd, self.dev, self.out = sd, device, {}
        self.n_nbr = n_nbr            # how often the cross-view attention counts to_out's bias (xview_plan: once per summed output)

    def f32(self, t):
        return t.detach().float().contiguous().to(self.dev)

    def put(self, name, t):
        self.out[name] = t.to(self.dev) if t.device != self.dev else t

    def conv3(self, p, pad_cin_to=None):
        w = self.sd[p + ".weight"].detach().float()
        if pad_cin_to is not None and w.shape[1] < pad_cin_to:
            wp = w.new_zeros((w.shape[0], pad_cin_to, 3, 3))
            wp[:, :w.shape[1]] = w
            w = wp
        self.put(p + ".w", pack_conv3x3(w))
        self.put(p + ".b", self.f32(self.sd[p + ".bias"]))

    def conv1(self, p):
        self.put(p + ".w", pack_conv1x1(self.sd[p + ".weight"].detach().float()))
        self.put(p + ".b", self.f32(self.sd[p + ".bias"]))

    def lin(self, p, bias=True):
        self.put(p + ".w", pack_linear(self.sd[p + ".weight"].detach().float()))
        if bias:
            self.put(p + ".b", self.f32(self.sd[p + ".bias"]))

    def lin32(self, p):
        self.put(p + ".w32", self.f32(self.sd[p + ".weight"]))
        self.put(p + ".b32", self.f32(self.sd[p + ".bias"]))

    def norm(self, p):
        self.put(p + ".g", self.f32(self.sd[p + ".weight"]))
        self.put(p + ".b", self.f32(self.sd[p + ".bias"]))

    def resnet(self, p, temb_list):
        self.norm(p + ".norm1"); self.conv3(p + ".conv1"); self.norm(p + ".norm2"); self.conv3(p + ".conv2")
        if (p + ".conv_shortcut.weight") in self.sd:
            self.conv1(p + ".conv_shortcut")
        temb_list.append(p)

    def attn_self(self, p, d):
        dp = _dp(d)
        f = lambda n: self.sd[f"{p}.{n}.weight"].detach().float()
        if d == 40:   # V heads on the 48-column stride with a column of ones (pad_v_ones): the projection gets a bias
            wv, bv = pad_v_ones(f("to_v"), d, dp)
            self.put(p + ".qkv.b", self.f32(torch.cat([torch.zeros(2 * HEADS * dp), bv])))
        else:
            wv = f("to_v")
        w = torch.cat([pad_heads(f("to_q"), d, dp), pad_heads(f("to_k"), d, dp), wv], 0)
        self.put(p + ".qkv.w", w.to(BF))

    def kv_cross(self, name, wk, wv, d):
        """[K | V] projection of a cross-attention's context tokens (text / camera / box tokens)"""
        dp = _dp(d)
        if d == 40:
            wv, bv = pad_v_ones(wv, d, dp)
            self.put(name + ".b", self.f32(torch.cat([torch.zeros(HEADS * dp), bv])))
        self.put(name + ".w", torch.cat([pad_heads(wk, d, dp), wv], 0).to(BF))

    def tblock(self, p, multiview):
        c = self.sd[p + ".norm1.weight"].shape[0]
        d = c // HEADS
        dp = _dp(d)
        f = lambda n: self.sd[f"{p}.{n}"].detach().float()
        for n in ("norm1", "norm2", "norm3"):
            self.norm(f"{p}.{n}")
        self.attn_self(p + ".attn1", d)
        self.lin(p + ".attn1.to_out.0")
        self.put(p + ".attn2.q.w", pad_heads(f("attn2.to_q.weight"), d, dp).to(BF))
        self.kv_cross(p + ".attn2.kv", f("attn2.to_k.weight"), f("attn2.to_v.weight"), d)
        self.lin(p + ".attn2.to_out.0")
        if multiview:
            self.norm(p + ".norm4")
            self.attn_self(p + ".attn4", d)
            wo, bo = f("attn4.to_out.0.weight").double(), f("attn4.to_out.0.bias").double()
            wc, bc = f("connector.weight").double(), f("connector.bias").double()
            self.put(p + ".attn4.oc.w", (wc @ wo).float().to(BF))          # connector o to_out, fused
            self.put(p + ".attn4.oc.b", self.f32(float(self.n_nbr) * (wc @ bo) + bc))   # bias once per neighbour (blocks.py:203-217)
        if (p + ".attn_temp.to_q.weight") in self.sd:   # temporal block of the video configuration (BASELINE config 5)
            self.norm(p + ".norm_temp")
            self.attn_self(p + ".attn_temp", d)
            self.lin(p + ".attn_temp.to_out.0")
        wg, bg = pack_geglu(f("ff.net.0.proj.weight"), f("ff.net.0.proj.bias"))
        self.put(p + ".ff.geglu.w", wg); self.put(p + ".ff.geglu.b", bg)
        self.lin(p + ".ff.net.2")

    def transformer2d(self, p, multiview):
        self.norm(p + ".norm"); self.conv1(p + ".proj_in"); self.conv1(p + ".proj_out")
        self.tblock(p + ".transformer_blocks.0", multiview)

    def encoder(self, multiview, temb_list):
        """conv_in, time embedding, 4 down blocks, mid block — shared by the UNet and the ControlNet branches"""
        # conv_in on the 4-channel latents: ONE K = 40 GEMM over an explicit 3x3 patch matrix (36 columns + zero tail)
        self.put("conv_in.wp", pack_conv3x3_patch(self.sd["conv_in.weight"].detach().float()))
        self.put("conv_in.b", self.f32(self.sd["conv_in.bias"]))
        self.lin32("time_embedding.linear_1"); self.lin32("time_embedding.linear_2")
        for i in range(4):
            for j in range(2):
                self.resnet(f"down_blocks.{i}.resnets.{j}", temb_list)
                if i < 3:
                    self.transformer2d(f"down_blocks.{i}.attentions.{j}", multiview)
            if i < 3:
                self.conv3(f"down_blocks.{i}.downsamplers.0.conv")
        self.resnet("mid_block.resnets.0", temb_list)
        self.transformer2d("mid_block.attentions.0", multiview)
        self.resnet("mid_block.resnets.1", temb_list)

    def temb(self, temb_list):
        ws = [self.sd[p + ".time_emb_proj.weight"].detach().float() for p in temb_list]
        bs = [self.sd[p + ".time_emb_proj.bias"].detach().float() for p in temb_list]
        self.put("temb_all.w", torch.cat(ws, 0).to(BF))
        self.put("temb_all.b", self.f32(torch.cat(bs, 0)))
        off, offs = 0, {}
        for p, w in zip(temb_list, ws):
            offs[p] = (off, w.shape[0])
            off += w.shape[0]
        self.out["temb_offsets"] = offs
        self.out["temb_total"] = off


def pack_unet(sd, device, neighboring_view_pair=None, neighboring_attn_type: str = "add"):
    pairs, n_src, concat, n_bias = xview_plan(neighboring_view_pair, neighboring_attn_type)
    pk = Packer(sd, device, n_bias)
    tl: List[str] = []
    pk.encoder(True, tl)
    for i in range(4):
        for j in range(3):
            pk.resnet(f"up_blocks.{i}.resnets.{j}", tl)
            if i > 0:
                pk.transformer2d(f"up_blocks.{i}.attentions.{j}", True)
        if i < 3:
            pk.conv3(f"up_blocks.{i}.upsamplers.0.conv")
    pk.norm("conv_norm_out")
    pk.conv3("conv_out")
    pk.temb(tl)
    pk.out["view_pairs"] = pairs
    pk.out["n_nbr"] = n_src
    pk.out["xview_mode"] = neighboring_attn_type
    return pk.out


def pack_sfa(pk: "Packer", p="txt_con_fusion"):
    """Semantic Fusion Attention (txt_con_fusion.py:27-33): 8 heads x 40, Q/K heads zero-padded to 48 columns"""
    f = lambda n: pk.sd[f"{p}.{n}.weight"].detach().float()
    pk.put(p + ".q.w", pad_heads(f("to_q"), 40, 48).to(BF))
    pk.kv_cross(p + ".kv", f("to_k"), f("to_v"), 40)
    pk.lin(p + ".to_out.0")


def pack_sfa_plus(pk: "Packer", p="txt_con_fusionp"):
    """txt_con_XFormersAttn_plus (txt_con_fusion.py:184-208): the three condition projections as ONE weight
    [q (heads padded to 48) | k (padded) | v] and the two text projections as one [k (padded) | v]"""
    f = lambda n: pk.sd[f"{p}.{n}.weight"].detach().float()
    wv, bv = pad_v_ones(f("to_v_occ"), 40, 48)
    pk.put(p + ".occ.w", torch.cat([pad_heads(f("to_q_occ"), 40, 48), pad_heads(f("to_k_occ"), 40, 48), wv], 0).to(BF))
    pk.put(p + ".occ.b", pk.f32(torch.cat([torch.zeros(16 * 48), bv])))
    pk.kv_cross(p + ".txt", f("to_k_txt"), f("to_v_txt"), 40)
    pk.lin(p + ".to_out.0")


def pack_cond_embedding(pk: "Packer", e="controlnet_cond_embedding"):
    """ControlNetConditioningEmbedding (map_embedder.py:81-112): conv_in (3 -> 16, channels padded to 8), 6 blocks, conv_out"""
    pk.conv3(e + ".conv_in", pad_cin_to=8)
    for i in range(6):
        pk.conv3(f"{e}.blocks.{i}")
    pk.conv3(e + ".conv_out")


def pack_controlnet(sd, device, use_occ_3d: bool, fusion: str = "sfa"):
    """fusion: "sfa" (txt_con_fusion, the dual-branch configs) or "sfa_plus" (txt_con_fusionp, occ_bg_fusionp.yaml)"""
    pk = Packer(sd, device)
    tl: List[str] = []
    pk.encoder(False, tl)
    pk.temb(tl)
    for i in range(12):
        pk.conv1(f"controlnet_down_blocks.{i}")
    pk.conv1("controlnet_mid_block")
    # embedders (fp32: camera intrinsics ~1e3 go through sin/cos(x * 2^k), SURVEY hard part 4)
    pk.lin32("cam2token")
    pk.put("uncond_cam", pk.f32(sd["uncond_cam.weight"]))
    for n in ("bbox_proj", "second_linear.0", "second_linear.2", "second_linear.4"):
        pk.lin32("bbox_embedder." + n)
    for n in ("_class_tokens", "null_class_feature", "null_pos_feature"):
        pk.put("bbox_embedder." + n, pk.f32(sd["bbox_embedder." + n]))
    if fusion == "sfa_plus":
        pack_sfa_plus(pk)
    else:
        pack_sfa(pk)
    if not use_occ_3d:
        pack_cond_embedding(pk)
    pk.out["use_occ_3d"] = use_occ_3d
    pk.out["fusion"] = fusion
    return pk.out


# ---------------------------------------------------------------------------------------------------
# forward building blocks
# ---------------------------------------------------------------------------------------------------
@dataclass
class Act:
    """compact channels-last activation: rows [n*H*W, C] bf16"""
    rows: torch.Tensor
    n: int
    H: int
    W: int

    @property
    def hw(self):
        return (self.H, self.W)

    @property
    def C(self):
        return self.rows.shape[1]


@dataclass
class StepCtx:
    n: int
    temb: torch.Tensor            # fp32 [m, temb_total]
    temb_rows_per_img_factor: int  # n // m
    text_kv: Dict[str, torch.Tensor] = field(default_factory=dict)   # attn2 prefix -> [n*Lk, 8*dp + C]
    lk: int = 0
    kv_map: Optional[torch.Tensor] = None
    n_nbr: int = 2                        # K/V sources per view of the cross-view attention (columns of kv_map)
    xview_concat: bool = False            # one softmax over the concatenated sources (neighboring_attn_type concat / self)
    view_shard: Optional[object] = None   # sharding.ViewShard when camera views are split across ranks
    n_outer: int = 0                      # scenes x CFG halves (sharded mode)
    n_frames: int = 1                     # video clips: local frames per clip; images are ordered [clip][frame][view]
    n_view: int = 6
    frame_shard: Optional[object] = None  # sharding.FrameShard when the frames of a clip are split across ranks


def time_embedding(P, t: torch.Tensor):
    """t fp32 [m] -> per-resnet projections fp32 [m, temb_total]  (unet_2d_condition_multiview.py:386-411 +
    every ResnetBlock2D.time_emb_proj(SiLU(emb)), batched into one tensor-core GEMM)"""
    c0 = P["time_embedding.linear_1.w32"].shape[1]
    s = ops.timestep_embedding(t, c0)
    h = ops.linear_f32(s, P["time_embedding.linear_1.w32"], P["time_embedding.linear_1.b32"], act=1)
    e16 = torch.empty((t.shape[0], P["time_embedding.linear_2.w32"].shape[0]), device=t.device, dtype=BF)
    ops.linear_f32(h, P["time_embedding.linear_2.w32"], P["time_embedding.linear_2.b32"], act=1, out16=e16,
                   want_f32=False)  # act=1: the SiLU every resnet applies to emb before its projection
    return ops.gemm(e16, P["temb_all.w"], bias=P["temb_all.b"], out_f32=True)


def resnet(P, p, x: Act, ctx: StepCtx, x2: Optional[torch.Tensor] = None) -> Act:
    n, hw = x.n, x.hw
    off, cout = P["temb_offsets"][p]
    rpi = ctx.temb_rows_per_img_factor * hw[0] * hw[1]
    g1 = ops.groupnorm(x.rows, P[p + ".norm1.g"], P[p + ".norm1.b"], n_img=n, hw=hw, x2=x2, eps=1e-5, silu=True,
                       padded_out=True)
    h = ops.gemm(g1, P[p + ".conv1.w"], bias=P[p + ".conv1.b"], taps=9, conv_hw=hw, n_img=n,
                 rowvec=ctx.temb[:, off:off + cout], rows_per_img=rpi)
    g2 = ops.groupnorm(h, P[p + ".norm2.g"], P[p + ".norm2.b"], n_img=n, hw=hw, eps=1e-5, silu=True, padded_out=True)
    if (p + ".conv_shortcut.w") in P:
        sc = ops.gemm(x.rows, P[p + ".conv_shortcut.w"], bias=P[p + ".conv_shortcut.b"], a2=x2)
    else:
        assert x2 is None
        sc = x.rows
    out = ops.gemm(g2, P[p + ".conv2.w"], bias=P[p + ".conv2.b"], taps=9, conv_hw=hw, n_img=n, res1=sc)
    return Act(out, n, x.H, x.W)


def text_kv(P, p_attn2, enc_rows):
    """K/V projection of the text/camera/box tokens for one attn2 layer (timestep-invariant)"""
    return ops.gemm(enc_rows, P[p_attn2 + ".kv.w"], bias=P.get(p_attn2 + ".kv.b"))


def transformer_block(P, p, h: torch.Tensor, n, T, ctx: StepCtx, multiview: bool):
    C = h.shape[1]
    d = C // HEADS
    dp = _dp(d)
    # 1. self-attention (blocks.py:163-172)
    ones = d == 40          # V heads padded to 48 columns with a column of ones (pad_v_ones): the kernel reads the denominator from O
    ln = ops.layernorm(h, P[p + ".norm1.g"], P[p + ".norm1.b"])
    qkv = ops.gemm(ln, P[p + ".attn1.qkv.w"], bias=P.get(p + ".attn1.qkv.b"))
    a = ops.attention(qkv, qkv, qkv, n_img=n, lq=T, lk=T, heads=HEADS, head_dim=d, q_col0=0, k_col0=HEADS * dp,
                      v_col0=2 * HEADS * dp, v_ones=ones)
    h = ops.gemm(a, P[p + ".attn1.to_out.0.w"], bias=P[p + ".attn1.to_out.0.b"], res1=h)
    # 2. text cross-attention (blocks.py:175-188); K/V come from the per-sample cache
    ln = ops.layernorm(h, P[p + ".norm2.g"], P[p + ".norm2.b"])
    q = ops.gemm(ln, P[p + ".attn2.q.w"])
    kv = ctx.text_kv[p + ".attn2"]
    a = ops.attention(q, kv, kv, n_img=n, lq=T, lk=ctx.lk, heads=HEADS, head_dim=d, k_col0=0, v_col0=HEADS * dp, v_ones=ones)
    h = ops.gemm(a, P[p + ".attn2.to_out.0.w"], bias=P[p + ".attn2.to_out.0.b"], res1=h)
    # 3. cross-view attention over the two ring neighbours (blocks.py:190-222)
    if multiview:
        if ctx.view_shard is None:
            ln = ops.layernorm(h, P[p + ".norm4.g"], P[p + ".norm4.b"])
            qkv = ops.gemm(ln, P[p + ".attn4.qkv.w"], bias=P.get(p + ".attn4.qkv.b"))
            a = ops.attention(qkv, qkv, qkv, n_img=n, lq=T, lk=T, heads=HEADS, head_dim=d, q_col0=0,
                              k_col0=HEADS * dp, v_col0=2 * HEADS * dp, kv_map=ctx.kv_map, n_src=ctx.n_nbr, v_ones=ones,
                              concat=ctx.xview_concat)
        else:
            # camera views sharded across ranks (the only exchange step of the path): the norm4 rows of the first / last local
            # view travel to the ring neighbours on a side stream WHILE the local views are projected; the K/V projection of
            # the two halo views (rows behind the local ones) follows when they have arrived
            vs = ctx.view_shard
            w_qkv, b_qkv = P[p + ".attn4.qkv.w"], P.get(p + ".attn4.qkv.b")
            wq, q_cols = w_qkv.shape[0], HEADS * dp
            n_kv = vs.kv_rows(ctx.n_outer)
            ln_ext = torch.empty((n_kv * T, C), device=h.device, dtype=BF)
            ops.layernorm(h, P[p + ".norm4.g"], P[p + ".norm4.b"], out=ln_ext[: n * T])
            vs.exchange_async(ln_ext, ctx.n_outer, T)
            buf = torch.empty((n_kv * T, wq), device=h.device, dtype=BF)
            ops.gemm(ln_ext[: n * T], w_qkv, bias=b_qkv, out=buf[: n * T])
            vs.exchange_wait(h.device)
            ops.gemm(ln_ext[n * T:], w_qkv[q_cols:], bias=None if b_qkv is None else b_qkv[q_cols:],
                     out=buf[n * T:, q_cols:])                                       # halo views: K and V columns only
            a = ops.attention(buf, buf, buf, n_img=n, n_kv_img=n_kv, lq=T, lk=T, heads=HEADS, head_dim=d, q_col0=0,
                              k_col0=HEADS * dp, v_col0=2 * HEADS * dp, kv_map=ctx.kv_map, n_src=ctx.n_nbr, v_ones=ones,
                              concat=ctx.xview_concat)
        h = ops.gemm(a, P[p + ".attn4.oc.w"], bias=P[p + ".attn4.oc.b"], res1=h)
    # 3b. temporal attention over the frames of the clip (no reference code: defined in csrc/dd_temporal.cu and
    #     oracle/dualdiff_oracle.py:temporal_attention); only packed for the video configuration
    if (p + ".attn_temp.qkv.w") in P and (ctx.n_frames > 1 or ctx.frame_shard is not None):
        h = temporal_step(P, p, h, n, T, ctx)
    # 4. GEGLU feed-forward (blocks.py:225-236)
    ln = ops.layernorm(h, P[p + ".norm3.g"], P[p + ".norm3.b"])
    ff = ops.gemm(ln, P[p + ".ff.geglu.w"], bias=P[p + ".ff.geglu.b"], geglu=True)
    return ops.gemm(ff, P[p + ".ff.net.2.w"], bias=P[p + ".ff.net.2.b"], res1=h)


def temporal_step(P, p, h, n, T, ctx: StepCtx):
    """h += to_out(MHA over frames) for every (clip, view, token); images are ordered [clip][frame][view]."""
    C = h.shape[1]
    d = C // HEADS
    dp = _dp(d)
    F_loc, V = ctx.n_frames, ctx.n_view
    n_clip = n // (F_loc * V)
    assert n_clip * F_loc * V == n, (n, F_loc, V)
    ln = ops.layernorm(h, P[p + ".norm_temp.g"], P[p + ".norm_temp.b"])
    w_qkv, b_qkv = P[p + ".attn_temp.qkv.w"], P.get(p + ".attn_temp.qkv.b")   # packed like the other self-attentions (V stride dp)
    kw = dict(n_outer=n_clip, n_view=V, tokens=T, heads=HEADS, head_dim=d, frames_q=F_loc)
    fs = ctx.frame_shard
    if fs is None or fs.world == 1:
        qkv = ops.gemm(ln, w_qkv, bias=b_qkv)
        a = ops.temporal_attention(qkv, qkv, qkv, q_col0=0, k_col0=HEADS * dp, v_col0=2 * HEADS * dp, v_hs=dp, **kw)
    else:
        # frames sharded over ranks: the exchange is a pair of all-to-alls around the attention (sharding.FrameShard).  The
        # LayerNorm rows travel (C columns, not the 3C' of the projections) from "my frames, all tokens" to "all frames, my
        # tokens"; the Q/K/V projection runs on the received rows -- the same number of rows as before, only other ones --
        # and the kernel addresses the [rank][clip][F_loc][view] blocks through its rank strides.
        t_me = len(fs.token_range(T))
        tok_rows = fs.to_token_shards(ln, n, T)
        if t_me > 0:
            qkv = ops.gemm(tok_rows, w_qkv, bias=b_qkv)
            F = F_loc * fs.world
            a_tok = ops.temporal_attention(qkv, qkv, qkv, n_outer=n_clip, n_view=V, tokens=t_me, heads=HEADS, head_dim=d,
                                           frames_q=F, frames_kv=F, frames_per_rank=F_loc, kv_rank_stride=n,
                                           frames_q_per_rank=F_loc, q_rank_stride=n, q_col0=0, k_col0=HEADS * dp,
                                           v_col0=2 * HEADS * dp, v_hs=dp)
        else:
            a_tok = torch.empty((0, C), device=h.device, dtype=BF)
        a = fs.from_token_shards(a_tok, n, T)
    return ops.gemm(a, P[p + ".attn_temp.to_out.0.w"], bias=P[p + ".attn_temp.to_out.0.b"], res1=h)


def transformer_2d(P, p, x: Act, ctx: StepCtx, multiview: bool) -> Act:
    n, hw = x.n, x.hw
    g = ops.groupnorm(x.rows, P[p + ".norm.g"], P[p + ".norm.b"], n_img=n, hw=hw, eps=1e-6, silu=False)
    h = ops.gemm(g, P[p + ".proj_in.w"], bias=P[p + ".proj_in.b"])
    h = transformer_block(P, p + ".transformer_blocks.0", h, n, hw[0] * hw[1], ctx, multiview)
    out = ops.gemm(h, P[p + ".proj_out.w"], bias=P[p + ".proj_out.b"], res1=x.rows)
    return Act(out, n, x.H, x.W)


def downsample(P, p, x: Act) -> Act:
    cols, (ho, wo) = ops.im2col_s2(x.rows, n_img=x.n, hw=x.hw)
    return Act(ops.gemm(cols, P[p + ".w"], bias=P[p + ".b"]), x.n, ho, wo)


def upsample(P, p, x: Act, hw2) -> Act:
    pad = ops.upsample_pad(x.rows, n_img=x.n, hw=x.hw, hw2=hw2)
    return Act(ops.gemm(pad, P[p + ".w"], bias=P[p + ".b"], taps=9, conv_hw=hw2, n_img=x.n), x.n, hw2[0], hw2[1])


def conv_in(P, latents, n_outer, n_view, H, W, res1=None) -> Act:
    """latents: fp32/bf16 NCHW storage of n_view images, logically repeated n_outer times (CFG halves)"""
    assert latents.is_contiguous()
    cols = ops.nchw_patches(latents, n_outer=n_outer, n_view=n_view, c=4, h=H, w=W, cp=P["conv_in.wp"].shape[1],
                            stride_outer=0 if n_outer > 1 and latents.shape[0] == n_view else n_view * 4 * H * W,
                            stride_view=4 * H * W, stride_c=H * W, stride_h=W)
    return Act(ops.gemm(cols, P["conv_in.wp"], bias=P["conv_in.b"], res1=res1), n_outer * n_view, H, W)


def down_path(P, x: Act, ctx: StepCtx, multiview: bool):
    skips = [x]
    for i in range(4):
        for j in range(2):
            x = resnet(P, f"down_blocks.{i}.resnets.{j}", x, ctx)
            if i < 3:
                x = transformer_2d(P, f"down_blocks.{i}.attentions.{j}", x, ctx, multiview)
            skips.append(x)
        if i < 3:
            x = downsample(P, f"down_blocks.{i}.downsamplers.0.conv", x)
            skips.append(x)
    return x, skips


def mid_block(P, x: Act, ctx: StepCtx, multiview: bool) -> Act:
    x = resnet(P, "mid_block.resnets.0", x, ctx)
    x = transformer_2d(P, "mid_block.attentions.0", x, ctx, multiview)
    return resnet(P, "mid_block.resnets.1", x, ctx)


ATTN2_LAYERS_ENC = [f"down_blocks.{i}.attentions.{j}.transformer_blocks.0.attn2" for i in range(3) for j in range(2)] + \
    ["mid_block.attentions.0.transformer_blocks.0.attn2"]
ATTN2_LAYERS_UNET = ATTN2_LAYERS_ENC + [f"up_blocks.{i}.attentions.{j}.transformer_blocks.0.attn2"
                                        for i in range(1, 4) for j in range(3)]


def make_kv_map(n, pairs, device, mode: str = "add"):
    """kv image of (query image, source slot); view index = image index mod n_cam (blocks.py:196-197).
    `pairs`: the model's neighboring_view_pair table (an int = that many views of the default ring, for callers without a model);
    mode "self": the sources of a view are all views of its scene"""
    if isinstance(pairs, int):
        assert pairs == len(NEIGHBORS), "an integer view count selects the default 6-view ring"
        pairs = NEIGHBORS
    n_cam = len(pairs)
    assert n % n_cam == 0
    if mode == "self":
        rows = [[(i // n_cam) * n_cam + v for v in range(n_cam)] for i in range(n)]
    else:
        rows = [[(i // n_cam) * n_cam + nb for nb in pairs[i % n_cam]] for i in range(n)]
    return torch.tensor(rows, dtype=torch.int32, device=device)


def prepare_text(P, layers, enc_rows):
    return {l: text_kv(P, l, enc_rows) for l in layers}


# ---------------------------------------------------------------------------------------------------
# UNet2DConditionModelMultiview.forward
# ---------------------------------------------------------------------------------------------------
def unet_forward(P, latents, n_outer, n_view, H, W, ctx: StepCtx, down_res: Optional[List[torch.Tensor]] = None,
                 mid_res: Optional[torch.Tensor] = None, down_res2: Optional[List[torch.Tensor]] = None,
                 mid_res2: Optional[torch.Tensor] = None, before_residuals=None) -> torch.Tensor:
    """returns eps as fp32 channels-last rows [n*H*W, 4].  down_res2 / mid_res2: residuals of a second branch added
    in the same pass (pipeline:422-429 sums the branches); before_residuals(): hook called right before the first
    use of the residuals (stream join when the branches run concurrently with the UNet encoder)."""
    x = conv_in(P, latents, n_outer, n_view, H, W)
    x, skips = down_path(P, x, ctx, True)
    x = mid_block(P, x, ctx, True)
    if before_residuals is not None:
        before_residuals()
    if down_res is not None:  # unet_2d_condition_multiview.py:464-473
        r2 = down_res2 if down_res2 is not None else [None] * len(down_res)
        skips = [Act(ops.add_bf16(s.rows, r, r_), s.n, s.H, s.W) for s, r, r_ in zip(skips, down_res, r2)]
    if mid_res is not None:
        x = Act(ops.add_bf16(x.rows, mid_res, mid_res2), x.n, x.H, x.W)
    for i in range(4):
        for j in range(3):
            s = skips.pop()
            x = resnet(P, f"up_blocks.{i}.resnets.{j}", x, ctx, x2=s.rows)
            if i > 0:
                x = transformer_2d(P, f"up_blocks.{i}.attentions.{j}", x, ctx, True)
        if i < 3:
            x = upsample(P, f"up_blocks.{i}.upsamplers.0.conv", x, skips[-1].hw)
    g = ops.groupnorm(x.rows, P["conv_norm_out.g"], P["conv_norm_out.b"], n_img=x.n, hw=x.hw, eps=1e-5, silu=True,
                      padded_out=True)
    return ops.gemm(g, P["conv_out.w"], bias=P["conv_out.b"], taps=9, conv_hw=x.hw, n_img=x.n, out_f32=True)


# ---------------------------------------------------------------------------------------------------
# BEVControlNetModel: hoisted (timestep-invariant) part and per-step part
# ---------------------------------------------------------------------------------------------------
def camera_tokens(P, camera_param):
    """(b, n_cam, 3, 7) fp32 -> (b*n_cam, 768) fp32   (unet_addon_rawbox.py:308-325,346-349)"""
    b, n_cam = camera_param.shape[:2]
    cols = camera_param.permute(0, 1, 3, 2).contiguous().float()      # (b, n, 7, 3): plumbing
    e = ops.fourier_embed(cols.reshape(-1, 3)).reshape(b * n_cam, -1)  # (b n) x (c d) = 189
    return ops.linear_f32(e, P["cam2token.w32"], P["cam2token.b32"])


def box_tokens(P, bboxes, classes, masks):
    """bboxes (R, L, n_pts, 3), classes (R, L), masks (R, L) -> (R*L, 768) fp32   (bbox_embedder.py:155-203)"""
    R, L = classes.shape
    nb = R * L
    dev = bboxes.device
    n_pts = bboxes.shape[-2]                  # 8 box corners / map-vector points (40 after `reinitialize()`, bbox_embedder.py:122-130)
    if P["bbox_embedder.bbox_proj.w32"].shape[1] != 27 * n_pts:
        raise ValueError(f"bbox_embedder expects {P['bbox_embedder.bbox_proj.w32'].shape[1] // 27} points per box, got {n_pts}")
    pos = torch.empty((nb, 27 * n_pts), device=dev, dtype=torch.float32)
    cat = torch.empty((nb, 768 + 768), device=dev, dtype=torch.float32)
    ops.box_features(bboxes.reshape(nb, n_pts, 3).float().contiguous(), classes.reshape(-1).contiguous(),
                     masks.reshape(-1), P["bbox_embedder._class_tokens"], P["bbox_embedder.null_pos_feature"],
                     P["bbox_embedder.null_class_feature"], pos, cat[:, 768:])
    ops.linear_f32(pos, P["bbox_embedder.bbox_proj.w32"], P["bbox_embedder.bbox_proj.b32"], act=1, out=cat[:, :768])
    h = ops.linear_f32(cat, P["bbox_embedder.second_linear.0.w32"], P["bbox_embedder.second_linear.0.b32"], act=1)
    h = ops.linear_f32(h, P["bbox_embedder.second_linear.2.w32"], P["bbox_embedder.second_linear.2.b32"], act=1)
    return ops.linear_f32(h, P["bbox_embedder.second_linear.4.w32"], P["bbox_embedder.second_linear.4.b32"])


def build_tokens(P, camera_param, text, bboxes_3d_data):
    """encoder_hidden_states_with_cam ++ box tokens: (b*n_cam, 1 + 77 + L, 768) bf16 (unet_addon_rawbox.py:832-896,
    1007,1066-1069).  Concats/expands here are pure data movement (torch), the arithmetic is in the kernels."""
    b, n_cam = camera_param.shape[:2]
    cam = camera_tokens(P, camera_param).reshape(b, n_cam, 1, 768)
    if text.shape[0] == b * n_cam and n_cam > 1:       # use_aug_text: one prompt per VIEW, '(b n) ... -> b n ...' (:351-352)
        txt = text.float().reshape(b, n_cam, text.shape[1], 768)
    elif text.shape[0] == b:                           # one prompt per scene, repeated over its views (:353-354)
        txt = text.float()[:, None].expand(b, n_cam, text.shape[1], 768)
    else:
        raise ValueError(f"text embeddings: {text.shape[0]} rows for {b} scenes x {n_cam} views (one per scene, or one per view "
                         "with use_aug_text)")
    if bboxes_3d_data is None:
        # the reference collate returns None when a batch has no visible box / map vector (dataset/utils.py:235-237) and
        # the branch then runs on [camera | text] tokens only (unet_addon_rawbox.py:892-895,1066-1069): zero box tokens
        return torch.cat([cam, txt], dim=2).reshape(b * n_cam, 78, 768).to(BF).contiguous()
    bb, cl, mk = bboxes_3d_data["bboxes"], bboxes_3d_data["classes"], bboxes_3d_data["masks"]
    n_box, L = bb.shape[1], bb.shape[2]
    if L == 0:
        return torch.cat([cam, txt], dim=2).reshape(b * n_cam, 78, 768).to(BF).contiguous()
    tok = box_tokens(P, bb.reshape(b * n_box, L, bb.shape[-2], 3), cl.reshape(b * n_box, L), mk.reshape(b * n_box, L))
    tok = tok.reshape(b, n_box, L, 768)
    if n_box != n_cam:
        tok = tok.expand(b, n_cam, L, 768)
    enc = torch.cat([cam, txt, tok], dim=2)                            # (b, n_cam, 78 + L, 768)
    return enc.reshape(b * n_cam, 78 + L, 768).to(BF).contiguous()


def cond_embedding(P, cond, n_cam=6) -> Act:
    """ControlNetConditioningEmbedding (map_embedder.py:114-138): (b, 3, Hc, 6*Wc) panorama -> (b*6, 320, Hc/8, Wc/8).
    Every conv runs on the tcgen05 GEMM (stride-1: padded implicit GEMM; stride-2: im2col), SiLU fused."""
    e = "controlnet_cond_embedding"
    b, c, Hc, Wt = cond.shape
    Wc = Wt // n_cam
    cond = cond.contiguous()
    pad = ops.nchw_to_padded(cond, n_outer=b, n_view=n_cam, c=c, h=Hc, w=Wc, cp=8, stride_outer=c * Hc * Wt,
                             stride_view=Wc, stride_c=Hc * Wt, stride_h=Wt)
    n = b * n_cam
    x = Act(ops.gemm(pad, P[e + ".conv_in.w"], bias=P[e + ".conv_in.b"], taps=9, conv_hw=(Hc, Wc), n_img=n, act=1), n, Hc, Wc)
    for i in range(6):
        p = f"{e}.blocks.{i}"
        if i % 2 == 0:
            pad = ops.pad_rows(x.rows, n_img=n, hw=x.hw)
            x = Act(ops.gemm(pad, P[p + ".w"], bias=P[p + ".b"], taps=9, conv_hw=x.hw, n_img=n, act=1), n, x.H, x.W)
        else:
            cols, (ho, wo) = ops.im2col_s2(x.rows, n_img=n, hw=x.hw)
            x = Act(ops.gemm(cols, P[p + ".w"], bias=P[p + ".b"], act=1), n, ho, wo)
    pad = ops.pad_rows(x.rows, n_img=n, hw=x.hw)
    return Act(ops.gemm(pad, P[e + ".conv_out.w"], bias=P[e + ".conv_out.b"], taps=9, conv_hw=x.hw, n_img=n), n, x.H, x.W)


def sfa_rows(P, cond_rows, txt_rows, n, T, L, p="txt_con_fusion") -> torch.Tensor:
    """cond + W_o MHA(W_q cond, W_k txt, W_v txt) + b_o on rows: cond_rows [n*T, 320], txt_rows [n*L, 768] (the L text tokens
    of every image, camera token already dropped) -> [n*T, 320]   (txt_con_fusion.py:110-177)"""
    q = ops.gemm(cond_rows, P[p + ".q.w"])
    kv = ops.gemm(txt_rows, P[p + ".kv.w"], bias=P[p + ".kv.b"])
    a = ops.attention(q, kv, kv, n_img=n, lq=T, lk=L, heads=8, head_dim=40, k_col0=0, v_col0=8 * 48, v_ones=True)
    return ops.gemm(a, P[p + ".to_out.0.w"], bias=P[p + ".to_out.0.b"], res1=cond_rows)


def sfa_plus_rows(P, cond_rows, txt_rows, n, T, L, p="txt_con_fusionp") -> torch.Tensor:
    """txt_con_XFormersAttn_plus on rows (txt_con_fusion.py:289-335): q' = MHA(W_q cond, W_kt txt, W_vt txt) gathers the text,
    then out = cond + W_o MHA(q', W_k cond, W_v cond) + b_o.  q' leaves the first attention as 8 heads of 40 columns and is
    read by the second one with a 40-column head stride: the 8 columns the 48-wide Q tile takes from the next head meet the
    zero padding of the keys."""
    proj = ops.gemm(cond_rows, P[p + ".occ.w"], bias=P[p + ".occ.b"])   # [n*T, 384 | 384 | 384] = q | k_occ | v_occ (+ ones)
    kv = ops.gemm(txt_rows, P[p + ".txt.w"], bias=P[p + ".txt.b"])      # [n*L, 384 | 384]
    q2 = ops.attention(proj, kv, kv, n_img=n, lq=T, lk=L, heads=8, head_dim=40, k_col0=0, v_col0=8 * 48, q_cols=8 * 48,
                       v_ones=True)
    a = ops.attention(q2, proj, proj, n_img=n, lq=T, lk=T, heads=8, head_dim=40, q_hs=40, k_col0=8 * 48, v_col0=16 * 48,
                      v_ones=True)
    return ops.gemm(a, P[p + ".to_out.0.w"], bias=P[p + ".to_out.0.b"], res1=cond_rows)


def sfa(P, cond: Act, enc_rows, lk_total, n) -> torch.Tensor:
    """Semantic Fusion Attention inside a branch.  enc_rows: [n*(78+L), 768]; the 77 text tokens are rows 1..77 of each image
    (camera token dropped, unet_addon_rawbox.py:977).  A strided window cannot be addressed as [n*77, ld]: the 77-token window
    is copied once (plumbing, timestep-invariant)."""
    txt = enc_rows.reshape(n, lk_total, enc_rows.shape[1])[:, 1:78].contiguous().reshape(n * 77, enc_rows.shape[1])
    if P.get("fusion", "sfa") == "sfa_plus":
        return sfa_plus_rows(P, cond.rows, txt, n, cond.H * cond.W, 77)
    return sfa_rows(P, cond.rows, txt, n, cond.H * cond.W, 77)


@dataclass
class BranchPrep:
    enc_rows: torch.Tensor      # [n*(78+L), 768] bf16
    lk: int
    cond: torch.Tensor          # [n*H*W, 320] bf16 — SFA-fused condition feature, added after conv_in (:990)
    text_kv: Dict[str, torch.Tensor]
    n: int


def controlnet_prepare(P, camera_param, text, bboxes_3d_data, controlnet_cond, H, W) -> BranchPrep:
    """everything of BEVControlNetModel.forward that does not depend on the timestep or the latents"""
    b, n_cam = camera_param.shape[:2]
    n = b * n_cam
    enc = build_tokens(P, camera_param, text, bboxes_3d_data)
    lk = enc.shape[1]
    enc_rows = enc.reshape(n * lk, 768)
    if P["use_occ_3d"]:
        assert controlnet_cond.shape[0] == n and controlnet_cond.shape[1] == 320
        cond = Act(ops.nchw_to_rows(controlnet_cond.contiguous()), n, H, W)   # ORS tensor (b*6, 320, h, w)
    else:
        cond = cond_embedding(P, controlnet_cond, n_cam)
        assert cond.hw == (H, W), (cond.hw, H, W)
    fused = sfa(P, cond, enc_rows, lk, n)
    return BranchPrep(enc_rows, lk, fused, prepare_text(P, ATTN2_LAYERS_ENC, enc_rows), n)


def controlnet_forward(P, prep: BranchPrep, latents, n_outer, n_view, H, W, t, acc: Optional[List[torch.Tensor]] = None,
                       conditioning_scale: float = 1.0):
    """per-step part of one branch.  acc: residuals of the previous branch to accumulate into (pipeline:422-429).
    Returns (12 down residual rows, mid residual rows)."""
    assert conditioning_scale == 1.0, "conditioning_scale != 1 is folded at pack time (not needed by the reference configs)"
    n = n_outer * n_view
    temb = time_embedding(P, t)
    ctx = StepCtx(n=n, temb=temb, temb_rows_per_img_factor=n // temb.shape[0], text_kv=prep.text_kv, lk=prep.lk)
    x = conv_in(P, latents, n_outer, n_view, H, W, res1=prep.cond)             # :965 + :990
    x, skips = down_path(P, x, ctx, False)
    x = mid_block(P, x, ctx, False)
    down = []
    for i, s in enumerate(skips):                                              # :1029-1039
        p = f"controlnet_down_blocks.{i}"
        down.append(ops.gemm(s.rows, P[p + ".w"], bias=P[p + ".b"], res1=None if acc is None else acc[i]))
    p = "controlnet_mid_block"
    mid = ops.gemm(x.rows, P[p + ".w"], bias=P[p + ".b"], res1=None if acc is None else acc[12])
    return down, mid
