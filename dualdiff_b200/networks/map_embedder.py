"""ControlNetConditioningEmbedding — parameter layout (reference: networks/map_embedder.py:81-138).
(b, 3, H, 6W) occupancy-projection panorama -> 6 views -> 8 convs (SiLU after the first 7, three stride-2)
-> (b*6, 320, H/8, W/8).  Forward = dualdiff_b200.engine.cond_embedding."""
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
