"""ControlnetUnetWrapper — the second caller of the denoising step (SURVEY.md §8b "Callers").

Reference: runner/multiview_runner.py:30-132.  The runner wraps [ControlNet-bg, ControlNet-fg] and the UNet into one module
whose forward maps (noisy latents, per-sample timesteps, camera, prompt states, condition inputs, boxes) to the noise
prediction: both branches, residual sum, tokens of branch 0, UNet over the `(b n)` batch.  This mirror keeps that call
signature and runs the forward on the CUDA mirrors (inference / validation-loss use: the kernels have no backward, so the
training step itself — §8f rank 4 — is not built and `requires_grad` inputs are refused).
"""
from typing import List, Optional, Union

import torch


class ControlnetUnetWrapper(torch.nn.Module):
    def __init__(self, controlnet, unet, controlnet_refer=None, weight_dtype=torch.float32, unet_in_fp16: bool = True):
        super().__init__()
        self.controlnet = controlnet
        if isinstance(controlnet, (list, tuple)):          # exposed so that `.to()` / state_dict reach the branches (:42-44)
            self.c_net1, self.c_net2 = controlnet[0], controlnet[1]
        self.unet = unet
        self.weight_dtype = weight_dtype
        self.unet_in_fp16 = unet_in_fp16

    @torch.no_grad()
    def forward(self, noisy_latents: torch.Tensor, timesteps: torch.Tensor, camera_param: torch.Tensor,
                encoder_hidden_states: torch.Tensor, encoder_hidden_states_uncond: Optional[torch.Tensor],
                controlnet_image, controlnet_image_occ: Union[torch.Tensor, List[torch.Tensor]], **kwargs):
        """noisy_latents (b, n_cam, 4, h, w); timesteps (b,) or (b, n_cam); -> model_pred (b, n_cam, 4, h, w)"""
        if noisy_latents.requires_grad:
            raise NotImplementedError("training (backward through the step) is not built: SURVEY §8f rank 4")
        if not isinstance(self.controlnet, (list, tuple)):
            raise NotImplementedError("the B200 path is the dual-branch configuration: controlnet = [bg branch, fg branch]")
        n_cam = noisy_latents.shape[1]
        kwargs = dict(kwargs)
        bboxes_3d_data = kwargs.pop("bboxes_3d_data")
        down = mid = states = None
        for i, net in enumerate(self.controlnet):                                        # :58-82
            d, m, s = net(noisy_latents, timesteps, camera_param=camera_param, encoder_hidden_states=encoder_hidden_states,
                          encoder_hidden_states_uncond=encoder_hidden_states_uncond, controlnet_cond=controlnet_image_occ[i],
                          return_dict=False, bboxes_3d_data=bboxes_3d_data[i], **kwargs)
            if i == 0:
                down, mid, states = list(d), m, s                                        # tokens from the first branch
            else:
                down = [a + b for a, b in zip(down, d)]
                mid = mid + m
        lat = noisy_latents.reshape(-1, *noisy_latents.shape[2:])                        # "b n ... -> (b n) ..."
        if timesteps.ndim == 1:
            timesteps = timesteps[:, None].expand(-1, n_cam)                             # "b -> (b n)"
        pred = self.unet(lat, timesteps.reshape(-1), encoder_hidden_states=states, down_block_additional_residuals=down,
                         mid_block_additional_residual=mid).sample
        return pred.reshape(-1, n_cam, *pred.shape[1:])
