import torch


class ModelMixin(torch.nn.Module):
    _supports_gradient_checkpointing = False

    @property
    def dtype(self):
        for p in self.parameters():
            return p.dtype
        return torch.float32

    @property
    def device(self):
        for p in self.parameters():
            return p.device
        return torch.device("cpu")

    def enable_xformers_memory_efficient_attention(self, attention_op=None):
        return None
