"""Host logic of the ControlnetUnetWrapper mirror (runner/multiview_runner.py:30-132) with stub networks: branch order,
residual sum, tokens of branch 0, `(b n)` flattening and timestep repeat.  The module forwards it strings together are
GPU-tested in tests/test_step_gpu.py::test_dropin_modules_match_reference."""
import os
import sys
from types import SimpleNamespace

import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


class _Net(torch.nn.Module):
    def __init__(self, k):
        super().__init__()
        self.k, self.calls = k, []

    def forward(self, sample, timesteps, camera_param, encoder_hidden_states, encoder_hidden_states_uncond, controlnet_cond,
                return_dict, bboxes_3d_data, **kw):
        self.calls.append(dict(cond=controlnet_cond, boxes=bboxes_3d_data, kw=kw, t=timesteps))
        n = sample.shape[0] * sample.shape[1]
        down = [torch.full((n, 2, 3, 3), float(self.k * (j + 1))) for j in range(3)]
        return down, torch.full((n, 2, 1, 1), float(10 * self.k)), torch.full((n, 5, 8), float(self.k))


class _UNet(torch.nn.Module):
    def forward(self, x, t, encoder_hidden_states, down_block_additional_residuals, mid_block_additional_residual):
        self.seen = dict(x=x, t=t, enc=encoder_hidden_states, down=down_block_additional_residuals,
                         mid=mid_block_additional_residual)
        return SimpleNamespace(sample=x * 2)


def test_wrapper_composition():
    from dualdiff_b200.runner import ControlnetUnetWrapper
    nets, unet = [_Net(1), _Net(2)], _UNet()
    wrap = ControlnetUnetWrapper(nets, unet)
    lat = torch.randn(2, 6, 4, 3, 3)
    t = torch.tensor([7, 900])
    out = wrap(lat, t, torch.zeros(2, 6, 3, 7), torch.zeros(2, 77, 8), None, None, ["bg", "fg"],
               bboxes_3d_data=["boxes_bg", "boxes_fg"], use_aug_text=False)
    assert out.shape == lat.shape and torch.equal(out, lat * 2)
    assert nets[0].calls[0]["cond"] == "bg" and nets[1].calls[0]["boxes"] == "boxes_fg" and nets[1].calls[0]["kw"] == {"use_aug_text": False}
    s = unet.seen
    assert s["x"].shape == (12, 4, 3, 3) and s["t"].tolist() == [7] * 6 + [900] * 6
    assert [float(d[0, 0, 0, 0]) for d in s["down"]] == [3.0, 6.0, 9.0] and float(s["mid"][0, 0, 0, 0]) == 30.0
    assert float(s["enc"][0, 0, 0]) == 1.0                      # tokens of the first branch
    assert wrap.c_net1 is nets[0] and wrap.c_net2 is nets[1]
    with pytest.raises(NotImplementedError, match="rank 4"):
        wrap(lat.requires_grad_(), t, None, None, None, None, ["bg", "fg"], bboxes_3d_data=[None, None])
    with pytest.raises(NotImplementedError, match="dual-branch"):
        ControlnetUnetWrapper(_Net(1), unet)(lat.detach(), t, None, None, None, None, None, bboxes_3d_data=None)
