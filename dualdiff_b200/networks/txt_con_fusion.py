"""Semantic Fusion Attention (SFA) (reference: networks/txt_con_fusion.py:18-181).
8-head cross-attention from the 320-channel condition feature map (queries) to the 77 text tokens (keys,
values) + output projection + residual.  Inside a branch the sampler reaches it through `engine.controlnet_prepare`;
called as a module (the reference's call site: unet_addon_rawbox.py:973-978) it runs the same `engine.sfa`."""
import torch
import torch.nn as nn


class txt_con_XFormersAttn(nn.Module):
    def __init__(self, con_dim=320, txt_dim=768, hidden_size=320):
        super().__init__()
        self.inner_dim = self.out_dim = hidden_size
        self.to_q = nn.Linear(con_dim, hidden_size, bias=False)
        self.to_k = nn.Linear(txt_dim, hidden_size, bias=False)
        self.to_v = nn.Linear(txt_dim, hidden_size, bias=False)
        self.to_out = nn.ModuleList([nn.Linear(hidden_size, hidden_size, bias=True), nn.Dropout(p=0.0)])
        self.heads = 8
        self.scale = (hidden_size // self.heads) ** -0.5
        self.rescale_output_factor = 1.0
        self.residual_connection = True
        self._packed = None

    def pack(self, device=None):
        from .. import engine
        device = torch.device(device) if device is not None else self.to_q.weight.device
        if device.type != "cuda":
            raise RuntimeError("dualdiff_b200 has no CPU path: move the module to a CUDA device (sm_100a) before use")
        pk = engine.Packer({"txt_con_fusion." + k: v for k, v in self.state_dict().items()}, device)
        engine.pack_sfa(pk)
        self._packed = pk.out
        return self

    def _apply(self, fn, *args, **kwargs):
        self._packed = None
        return super()._apply(fn, *args, **kwargs)

    def forward(self, attn=None, hidden_states=None, encoder_hidden_states=None, attention_mask=None, temb=None):
        """attention-processor signature of the reference (txt_con_fusion.py:42-49; `attn` is unused there too).
        hidden_states (n, 320, h, w) condition feature map; encoder_hidden_states (n, 77, 768) text tokens.
        Returns (n, 320, h, w) in the dtype of `hidden_states` (fp32 or bf16; a channels_last view of the kernel's rows)."""
        from .. import engine, ops
        if attention_mask is not None or temb is not None:
            raise NotImplementedError("attention_mask / temb are None on the reference path (txt_con_fusion.py:61-62)")
        if hidden_states.dim() != 4 or encoder_hidden_states is None:
            raise ValueError("txt_con_XFormersAttn: hidden_states must be (n, C, h, w) with encoder_hidden_states (n, L, 768)")
        if not hidden_states.is_cuda:
            raise RuntimeError("dualdiff_b200 has no CPU path: `hidden_states` must be a CUDA tensor")
        if self._packed is None:
            self.pack(hidden_states.device)
        n, c, h, w = hidden_states.shape
        x = hidden_states if hidden_states.dtype in (torch.float32, torch.bfloat16) else hidden_states.float()
        rows = ops.nchw_to_rows(x.contiguous())
        txt = encoder_hidden_states.to(torch.bfloat16).contiguous()
        L = txt.shape[1]
        out = engine.sfa_rows(self._packed, rows, txt.reshape(n * L, txt.shape[2]), n, h * w, L)
        res = out.reshape(n, h, w, c).permute(0, 3, 1, 2)
        return res if hidden_states.dtype == torch.bfloat16 else res.to(hidden_states.dtype)


class txt_con_XFormersAttn_plus(nn.Module):
    """the `_plus` variant (reference: networks/txt_con_fusion.py:184-337; config `use_txt_con_fusionp`,
    configs/exp/occ_bg_fusionp.yaml): the condition queries first gather the text tokens, and the result is the query of a
    self-attention over the condition feature map's own keys / values; output projection + residual.  Runs `engine.sfa_plus_rows`
    (two launches of the head_dim-40 attention kernel around three GEMMs)."""

    def __init__(self, con_dim=320, txt_dim=768, hidden_size=320):
        super().__init__()
        self.inner_dim = self.out_dim = hidden_size
        self.to_q_occ = nn.Linear(con_dim, hidden_size, bias=False)
        self.to_k_occ = nn.Linear(con_dim, hidden_size, bias=False)
        self.to_v_occ = nn.Linear(con_dim, hidden_size, bias=False)
        self.to_k_txt = nn.Linear(txt_dim, hidden_size, bias=False)
        self.to_v_txt = nn.Linear(txt_dim, hidden_size, bias=False)
        self.to_out = nn.ModuleList([nn.Linear(hidden_size, hidden_size, bias=True), nn.Dropout(p=0.0)])
        self.heads = 8
        self.scale = (hidden_size // self.heads) ** -0.5
        self.rescale_output_factor = 1.0
        self.residual_connection = True
        self._packed = None

    def pack(self, device=None):
        from .. import engine
        device = torch.device(device) if device is not None else self.to_q_occ.weight.device
        if device.type != "cuda":
            raise RuntimeError("dualdiff_b200 has no CPU path: move the module to a CUDA device (sm_100a) before use")
        pk = engine.Packer({"txt_con_fusionp." + k: v for k, v in self.state_dict().items()}, device)
        engine.pack_sfa_plus(pk)
        self._packed = pk.out
        return self

    def _apply(self, fn, *args, **kwargs):
        self._packed = None
        return super()._apply(fn, *args, **kwargs)

    def forward(self, attn=None, hidden_states=None, encoder_hidden_states=None, attention_mask=None, temb=None):
        """attention-processor signature of the reference (txt_con_fusion.py:211-218).  hidden_states (n, 320, h, w);
        encoder_hidden_states (n, L, 768).  Returns (n, 320, h, w) in the dtype of `hidden_states`."""
        from .. import engine, ops
        if attention_mask is not None or temb is not None:
            raise NotImplementedError("attention_mask / temb are None on the reference path (txt_con_fusion.py:311)")
        if hidden_states.dim() != 4 or encoder_hidden_states is None:
            raise ValueError("txt_con_XFormersAttn_plus: hidden_states must be (n, C, h, w) with encoder_hidden_states (n, L, 768)")
        if not hidden_states.is_cuda:
            raise RuntimeError("dualdiff_b200 has no CPU path: `hidden_states` must be a CUDA tensor")
        if self._packed is None:
            self.pack(hidden_states.device)
        n, c, h, w = hidden_states.shape
        x = hidden_states if hidden_states.dtype in (torch.float32, torch.bfloat16) else hidden_states.float()
        rows = ops.nchw_to_rows(x.contiguous())
        txt = encoder_hidden_states.to(torch.bfloat16).contiguous()
        L = txt.shape[1]
        out = engine.sfa_plus_rows(self._packed, rows, txt.reshape(n * L, txt.shape[2]), n, h * w, L)
        res = out.reshape(n, h, w, c).permute(0, 3, 1, 2)
        return res if hidden_states.dtype == torch.bfloat16 else res.to(hidden_states.dtype)
