"""ContinuousBBoxWithTextEmbedding — parameter layout (reference: networks/bbox_embedder.py:29-203).
3D-box / map-vector tokens: Fourier(8 corners) -> Linear(216,768)+SiLU -> cat CLIP class token ->
MLP(1536->512->512->768); masked boxes use the learned null features."""
import torch
import torch.nn as nn

from .embedder import get_embedder


class ContinuousBBoxWithTextEmbedding(nn.Module):
    def __init__(self, n_classes, class_token_dim=768, trainable_class_token=False, embedder_num_freq=4,
                 proj_dims=(768, 512, 512, 768), mode="cxyz", minmax_normalize=True, use_text_encoder_init=True,
                 **kwargs):
        super().__init__()
        if mode != "all-xyz" or minmax_normalize:
            raise NotImplementedError("dualdiff_b200 implements mode='all-xyz', minmax_normalize=False "
                                      "(sd-controlnet-seg/config.json:6-22)")
        self.mode, self.minmax_normalize, self.use_text_encoder_init = mode, minmax_normalize, use_text_encoder_init
        self.fourier_embedder = get_embedder(3, embedder_num_freq)
        self.bbox_proj = nn.Linear(self.fourier_embedder.out_dim * 8, proj_dims[0])
        self.second_linear = nn.Sequential(nn.Linear(proj_dims[0] + class_token_dim, proj_dims[1]), nn.SiLU(),
                                           nn.Linear(proj_dims[1], proj_dims[2]), nn.SiLU(),
                                           nn.Linear(proj_dims[2], proj_dims[3]))
        tokens = torch.randn(n_classes, class_token_dim)
        if trainable_class_token:
            self.register_parameter("_class_tokens", nn.Parameter(tokens))
        else:
            self.register_buffer("_class_tokens", tokens)
        self.null_class_feature = nn.Parameter(torch.zeros([class_token_dim]))
        self.null_pos_feature = nn.Parameter(torch.zeros([self.fourier_embedder.out_dim * 8]))

    @property
    def class_tokens(self):
        return self._class_tokens

    def reinitialize(self, output_num: int = 40):
        """re-create `bbox_proj` / `null_pos_feature` for map vectors of 40 points (reference bbox_embedder.py:122-130; called by
        misc/test_utils.py:116-121 before the branch's bbox_embedder weights are loaded a second time)"""
        proj_dim = self.bbox_proj.out_features
        dev, dt = self.bbox_proj.weight.device, self.bbox_proj.weight.dtype
        self.bbox_proj = nn.Linear(self.fourier_embedder.out_dim * output_num, proj_dim).to(device=dev, dtype=dt)
        self.null_pos_feature = nn.Parameter(torch.zeros([self.fourier_embedder.out_dim * output_num], device=dev, dtype=dt))

    def prepare(self, cfg, **kwargs):
        if self.use_text_encoder_init:
            self.set_category_token(kwargs["tokenizer"], kwargs["text_encoder"], cfg.dataset.object_classes)

    @torch.no_grad()
    def set_category_token(self, tokenizer, text_encoder, class_names):
        for idx, name in enumerate(class_names):
            ids = tokenizer([name], padding="do_not_pad", return_tensors="pt").input_ids.to(self._class_tokens.device)
            self._class_tokens[idx].copy_(text_encoder(ids).pooler_output[0])
