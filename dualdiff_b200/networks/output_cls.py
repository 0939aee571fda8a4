"""Return containers (reference: networks/output_cls.py:9-13; diffusers UNet2DConditionOutput)."""
from dataclasses import dataclass, fields
from typing import List

import torch


class _Output:
    def to_tuple(self):
        return tuple(getattr(self, f.name) for f in fields(self))

    def __getitem__(self, k):
        return getattr(self, k) if isinstance(k, str) else self.to_tuple()[k]

    def __iter__(self):
        return iter(self.to_tuple())


@dataclass
class BEVControlNetOutput(_Output):
    down_block_res_samples: List[torch.Tensor]
    mid_block_res_sample: torch.Tensor
    encoder_hidden_states_with_cam: torch.Tensor


@dataclass
class UNet2DConditionOutput(_Output):
    sample: torch.Tensor
