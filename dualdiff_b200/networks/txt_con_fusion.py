"""Semantic Fusion Attention (SFA) — parameter layout (reference: networks/txt_con_fusion.py:18-181).
8-head cross-attention from the 320-channel condition feature map (queries) to the 77 text tokens (keys,
values) + output projection + residual.  Forward = dualdiff_b200.engine.sfa."""
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


class txt_con_XFormersAttn_plus(nn.Module):
    """parameter layout of the optional `_plus` variant (txt_con_fusion.py:184-208; config
    `use_txt_con_fusionp`, off in the dual-branch configs).  Kept so reference checkpoints load; its forward is
    not on the hot path and is not implemented."""

    def __init__(self, con_dim=320, txt_dim=768, hidden_size=320):
        super().__init__()
        self.to_q_occ = nn.Linear(con_dim, hidden_size, bias=False)
        self.to_k_occ = nn.Linear(con_dim, hidden_size, bias=False)
        self.to_v_occ = nn.Linear(con_dim, hidden_size, bias=False)
        self.to_k_txt = nn.Linear(txt_dim, hidden_size, bias=False)
        self.to_v_txt = nn.Linear(txt_dim, hidden_size, bias=False)
        self.to_out = nn.ModuleList([nn.Linear(hidden_size, hidden_size, bias=True), nn.Dropout(p=0.0)])
        self.heads = 8
