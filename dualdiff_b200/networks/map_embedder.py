"""ControlNetConditioningEmbedding — parameter layout (reference: networks/map_embedder.py:81-138).
(b, 3, H, 6W) occupancy-projection panorama -> 6 views -> 8 convs (SiLU after the first 7, three stride-2)
-> (b*6, 320, H/8, W/8).  Inside a branch the sampler reaches it through `engine.controlnet_prepare`; called as a module
(the reference's call site: unet_addon_rawbox.py:967-968) it runs the same `engine.cond_embedding`."""
import torch
import torch.nn as nn


class ControlNetConditioningEmbedding(nn.Module):
    def __init__(self, conditioning_embedding_channels, conditioning_channels=3,
                 block_out_channels=(16, 32, 96, 256), conditioning_size=None):
        super().__init__()
        self.conv_in = nn.Conv2d(conditioning_channels, block_out_channels[0], 3, padding=1)
        self.blocks = nn.ModuleList()
        for i in range(len(block_out_channels) - 1):
            cin, cout = block_out_channels[i], block_out_channels[i + 1]
            self.blocks.append(nn.Conv2d(cin, cin, 3, padding=1))
            self.blocks.append(nn.Conv2d(cin, cout, 3, padding=1, stride=2))
        self.conv_out = nn.Conv2d(block_out_channels[-1], conditioning_embedding_channels, 3, padding=1)
        for p in self.conv_out.parameters():  # zero_module (map_embedder.py:110-112)
            nn.init.zeros_(p)
        self._packed = None

    def pack(self, device=None):
        from .. import engine
        device = torch.device(device) if device is not None else self.conv_in.weight.device
        if device.type != "cuda":
            raise RuntimeError("dualdiff_b200 has no CPU path: move the module to a CUDA device (sm_100a) before use")
        pk = engine.Packer({"controlnet_cond_embedding." + k: v for k, v in self.state_dict().items()}, device)
        engine.pack_cond_embedding(pk)
        self._packed = pk.out
        return self

    def _apply(self, fn, *args, **kwargs):
        self._packed = None
        return super()._apply(fn, *args, **kwargs)

    def forward(self, conditioning):
        """(b, 3, H, 6*W) panorama in [0, 1] -> (b*6, 320, H/8, W/8), views split along the width (map_embedder.py:114-138).
        Returned as a channels_last bf16 view of the kernel's rows (fp32 input -> fp32 output)."""
        from .. import engine
        if not conditioning.is_cuda:
            raise RuntimeError("dualdiff_b200 has no CPU path: `conditioning` must be a CUDA tensor")
        if self._packed is None:
            self.pack(conditioning.device)
        x = conditioning if conditioning.dtype in (torch.float32, torch.bfloat16) else conditioning.float()
        act = engine.cond_embedding(self._packed, x)
        out = act.rows.reshape(act.n, act.H, act.W, act.C).permute(0, 3, 1, 2)
        return out if conditioning.dtype == torch.bfloat16 else out.to(conditioning.dtype)
