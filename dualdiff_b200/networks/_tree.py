"""Parameter containers with diffusers-0.17.1 state-dict key names (SURVEY.md §8b "State-dict compatibility").

These classes only *hold* parameters (so `load_state_dict` of a reference/diffusers checkpoint works and
`state_dict()` round-trips); none of their `forward`s is ever used — compute happens in
:mod:`dualdiff_b200.engine` through the CUDA C ABI.  Structure restated from SURVEY.md Appendix A.1.
"""
import torch
import torch.nn as nn


class _NoForward(nn.Module):
    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError(f"{type(self).__name__} is a parameter container; the CUDA engine runs the math")


class FusedAttnProcessor:
    """The one attention implementation of this package: softmax(Q K^T * scale) V in the fused tcgen05 kernel
    (csrc/dd_attention.cu).  It stands where the reference installs diffusers' `XFormersAttnProcessor`
    (`enable_xformers_memory_efficient_attention`, misc/test_utils.py:164-165; protocol: box_adapter.py:33-40)."""

    def __repr__(self):
        return "FusedAttnProcessor(tcgen05)"


# stock processors whose arithmetic is exactly what the fused kernel computes: installing one changes nothing
_EQUIVALENT_PROCESSORS = ("AttnProcessor", "AttnProcessor2_0", "XFormersAttnProcessor", "FusedAttnProcessor")


def check_attn_processor(processor):
    """`set_attn_processor` / `set_processor` contract: None (default) or a processor with the stock arithmetic is accepted;
    anything else would be silently ignored by the fused kernel, so it raises (no silent dispatch)."""
    if processor is None or type(processor).__name__ in _EQUIVALENT_PROCESSORS:
        return
    raise NotImplementedError(
        f"attention processor {type(processor).__name__!r} is not supported: dualdiff_b200 runs every attention in its fused "
        "tcgen05 kernel and cannot call back into Python processors (e.g. box_adapter's Adapter_XFormersAttnProcessor, "
        "which the reference itself asserts incompatible with the dual branch, runner/multiview_runner.py:240)")


class Attention(_NoForward):
    """diffusers Attention parameter layout: to_q/to_k/to_v (no bias), to_out = [Linear(bias), Dropout]."""

    def __init__(self, query_dim, cross_attention_dim=None, heads=8, dim_head=64, bias=False):
        super().__init__()
        inner = heads * dim_head
        kv_dim = query_dim if cross_attention_dim is None else cross_attention_dim
        self.heads, self.dim_head, self.scale = heads, dim_head, dim_head ** -0.5
        self.to_q = nn.Linear(query_dim, inner, bias=bias)
        self.to_k = nn.Linear(kv_dim, inner, bias=bias)
        self.to_v = nn.Linear(kv_dim, inner, bias=bias)
        self.to_out = nn.ModuleList([nn.Linear(inner, query_dim), nn.Dropout(0.0)])
        self.processor = FusedAttnProcessor()

    def set_processor(self, processor):
        check_attn_processor(processor)
        self.processor = FusedAttnProcessor()


class GEGLU(_NoForward):
    def __init__(self, dim_in, dim_out):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out * 2)


class FeedForward(_NoForward):
    def __init__(self, dim, mult=4):
        super().__init__()
        self.net = nn.ModuleList([GEGLU(dim, dim * mult), nn.Dropout(0.0), nn.Linear(dim * mult, dim)])


class BasicTransformerBlock(_NoForward):
    def __init__(self, dim, num_attention_heads, attention_head_dim, cross_attention_dim=None):
        super().__init__()
        self.dim = dim
        self.attn1 = Attention(dim, None, num_attention_heads, attention_head_dim)
        self.ff = FeedForward(dim)
        self.attn2 = Attention(dim, cross_attention_dim, num_attention_heads, attention_head_dim)
        self.norm1 = nn.LayerNorm(dim)
        self.norm2 = nn.LayerNorm(dim)
        self.norm3 = nn.LayerNorm(dim)


class Transformer2DModel(_NoForward):
    def __init__(self, num_attention_heads, attention_head_dim, in_channels, cross_attention_dim, norm_num_groups=32,
                 block_cls=BasicTransformerBlock, block_kwargs=None):
        super().__init__()
        inner = num_attention_heads * attention_head_dim
        self.norm = nn.GroupNorm(norm_num_groups, in_channels, eps=1e-6)
        self.proj_in = nn.Conv2d(in_channels, inner, 1)
        self.transformer_blocks = nn.ModuleList([
            block_cls(inner, num_attention_heads, attention_head_dim, cross_attention_dim=cross_attention_dim,
                      **(block_kwargs or {}))])
        self.proj_out = nn.Conv2d(inner, in_channels, 1)


class ResnetBlock2D(_NoForward):
    def __init__(self, in_channels, out_channels, temb_channels, groups=32, eps=1e-5):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.norm1 = nn.GroupNorm(groups, in_channels, eps=eps)
        self.conv1 = nn.Conv2d(in_channels, out_channels, 3, padding=1)
        self.time_emb_proj = nn.Linear(temb_channels, out_channels)
        self.norm2 = nn.GroupNorm(groups, out_channels, eps=eps)
        self.conv2 = nn.Conv2d(out_channels, out_channels, 3, padding=1)
        self.conv_shortcut = nn.Conv2d(in_channels, out_channels, 1) if in_channels != out_channels else None


class Downsample2D(_NoForward):
    def __init__(self, channels, padding=1):
        super().__init__()
        self.conv = nn.Conv2d(channels, channels, 3, stride=2, padding=padding)


class Upsample2D(_NoForward):
    def __init__(self, channels):
        super().__init__()
        self.conv = nn.Conv2d(channels, channels, 3, padding=1)


class _Block(_NoForward):
    gradient_checkpointing = False


class CrossAttnDownBlock2D(_Block):
    has_cross_attention = True

    def __init__(self, in_channels, out_channels, temb_channels, num_layers, heads, cross_attention_dim,
                 add_downsample, groups=32, eps=1e-5, tf_kwargs=None):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(in_channels if i == 0 else out_channels, out_channels,
                                                    temb_channels, groups, eps) for i in range(num_layers)])
        self.attentions = nn.ModuleList([Transformer2DModel(heads, out_channels // heads, out_channels,
                                                            cross_attention_dim, groups, **(tf_kwargs or {}))
                                         for _ in range(num_layers)])
        self.downsamplers = nn.ModuleList([Downsample2D(out_channels)]) if add_downsample else None


class DownBlock2D(_Block):
    def __init__(self, in_channels, out_channels, temb_channels, num_layers, add_downsample, groups=32, eps=1e-5):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(in_channels if i == 0 else out_channels, out_channels,
                                                    temb_channels, groups, eps) for i in range(num_layers)])
        self.downsamplers = nn.ModuleList([Downsample2D(out_channels)]) if add_downsample else None


class UNetMidBlock2DCrossAttn(_Block):
    has_cross_attention = True

    def __init__(self, in_channels, temb_channels, heads, cross_attention_dim, groups=32, eps=1e-5, tf_kwargs=None):
        super().__init__()
        self.attentions = nn.ModuleList([Transformer2DModel(heads, in_channels // heads, in_channels,
                                                            cross_attention_dim, groups, **(tf_kwargs or {}))])
        self.resnets = nn.ModuleList([ResnetBlock2D(in_channels, in_channels, temb_channels, groups, eps)
                                      for _ in range(2)])


class CrossAttnUpBlock2D(_Block):
    has_cross_attention = True

    def __init__(self, in_channels, out_channels, prev_output_channel, temb_channels, num_layers, heads,
                 cross_attention_dim, add_upsample, groups=32, eps=1e-5, tf_kwargs=None):
        super().__init__()
        res = []
        for i in range(num_layers):
            skip = in_channels if i == num_layers - 1 else out_channels
            cin = prev_output_channel if i == 0 else out_channels
            res.append(ResnetBlock2D(cin + skip, out_channels, temb_channels, groups, eps))
        self.resnets = nn.ModuleList(res)
        self.attentions = nn.ModuleList([Transformer2DModel(heads, out_channels // heads, out_channels,
                                                            cross_attention_dim, groups, **(tf_kwargs or {}))
                                         for _ in range(num_layers)])
        self.upsamplers = nn.ModuleList([Upsample2D(out_channels)]) if add_upsample else None


class UpBlock2D(_Block):
    def __init__(self, in_channels, out_channels, prev_output_channel, temb_channels, num_layers, add_upsample,
                 groups=32, eps=1e-5):
        super().__init__()
        res = []
        for i in range(num_layers):
            skip = in_channels if i == num_layers - 1 else out_channels
            cin = prev_output_channel if i == 0 else out_channels
            res.append(ResnetBlock2D(cin + skip, out_channels, temb_channels, groups, eps))
        self.resnets = nn.ModuleList(res)
        self.upsamplers = nn.ModuleList([Upsample2D(out_channels)]) if add_upsample else None


class Timesteps(nn.Module):
    """weight-less (diffusers Timesteps); kept so `time_proj` exists on the models"""

    def __init__(self, num_channels, flip_sin_to_cos=True, downscale_freq_shift=0):
        super().__init__()
        self.num_channels = num_channels
        assert flip_sin_to_cos and downscale_freq_shift == 0, "dualdiff_b200 implements the SDv1.5 setting only"


class TimestepEmbedding(_NoForward):
    def __init__(self, in_channels, time_embed_dim):
        super().__init__()
        self.linear_1 = nn.Linear(in_channels, time_embed_dim)
        self.act = nn.SiLU()
        self.linear_2 = nn.Linear(time_embed_dim, time_embed_dim)


class AttrDict(dict):
    """`self.config` with attribute access and `**config` expansion (diffusers FrozenDict behaviour)"""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e


class ModelBase(nn.Module):
    """the slice of diffusers ModelMixin/ConfigMixin the reference's callers touch (SURVEY §8c attribute surface)"""

    config_name = "config.json"
    weights_names = ("diffusion_pytorch_model.safetensors", "diffusion_pytorch_model.bin")

    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path, torch_dtype=None, subfolder=None,
                        ignore_mismatched_sizes=False, **kwargs):
        """load a diffusers-format checkpoint directory (config.json + diffusion_pytorch_model.{safetensors,bin}) the
        way misc/test_utils.py:111-113,146-147 does.  Like diffusers' loader: keys missing from the checkpoint and
        unexpected keys are reported with a warning; a shape mismatch RAISES unless `ignore_mismatched_sizes=True` (the
        reference passes it for the ControlNet branches only, misc/test_utils.py:111-113), in which case the mismatched
        tensors keep their fresh initialisation and are listed in the warning.  Other kwargs of the diffusers loader
        (low_cpu_mem_usage, device_map, ...) are accepted and ignored."""
        import inspect
        import json
        import logging
        import os
        root = os.path.join(pretrained_model_name_or_path, subfolder) if subfolder else pretrained_model_name_or_path
        with open(os.path.join(root, cls.config_name)) as fh:
            cfg = {k: v for k, v in json.load(fh).items() if not k.startswith("_")}
        accepted = set(inspect.signature(cls.__init__).parameters) - {"self"}
        model = cls(**{k: v for k, v in cfg.items() if k in accepted})
        sd = None
        for name in cls.weights_names:
            path = os.path.join(root, name)
            if os.path.exists(path):
                if name.endswith(".safetensors"):
                    from safetensors.torch import load_file
                    sd = load_file(path)
                else:
                    sd = torch.load(path, map_location="cpu", weights_only=True)
                break
        if sd is None:
            raise FileNotFoundError(f"no {' / '.join(cls.weights_names)} under {root}")
        own = model.state_dict()
        mismatched = [(k, tuple(v.shape), tuple(own[k].shape)) for k, v in sd.items()
                      if k in own and tuple(own[k].shape) != tuple(v.shape)]
        if mismatched and not ignore_mismatched_sizes:
            lines = "\n".join(f"  {k}: checkpoint {a} vs model {b}" for k, a, b in mismatched[:20])
            raise RuntimeError(f"{cls.__name__}.from_pretrained({root}): size mismatch for {len(mismatched)} tensor(s); pass "
                               f"ignore_mismatched_sizes=True to keep the model's initialisation for them\n{lines}")
        bad = {k for k, _, _ in mismatched}
        res = model.load_state_dict({k: v for k, v in sd.items() if k not in bad}, strict=False)
        missing = [k for k in res.missing_keys if k not in bad]
        log = logging.getLogger(__name__)
        if missing:
            log.warning("%s.from_pretrained(%s): %d key(s) missing from the checkpoint keep their fresh initialisation: %s%s",
                        cls.__name__, root, len(missing), ", ".join(missing[:8]), " ..." if len(missing) > 8 else "")
        if res.unexpected_keys:
            log.warning("%s.from_pretrained(%s): %d unexpected key(s) in the checkpoint were ignored: %s%s", cls.__name__, root,
                        len(res.unexpected_keys), ", ".join(res.unexpected_keys[:8]), " ..." if len(res.unexpected_keys) > 8 else "")
        if mismatched:
            log.warning("%s.from_pretrained(%s): %d mismatched tensor(s) were NOT loaded (ignore_mismatched_sizes=True): %s",
                        cls.__name__, root, len(mismatched), ", ".join(k for k, _, _ in mismatched[:8]))
        model._load_report = dict(missing=missing, unexpected=list(res.unexpected_keys), mismatched=[k for k, _, _ in mismatched])
        if torch_dtype is not None:
            model = model.to(torch_dtype)
        return model.eval()

    def save_pretrained(self, save_directory, safe_serialization=True, **kwargs):
        import json
        import os
        os.makedirs(save_directory, exist_ok=True)
        cfg = {"_class_name": type(self).__name__, "_diffusers_version": "0.17.1"}
        cfg.update({k: (list(v) if isinstance(v, tuple) else v) for k, v in self.config.items()})
        with open(os.path.join(save_directory, self.config_name), "w") as fh:
            json.dump(cfg, fh, indent=2, default=str)
        sd = {k: v.detach().cpu().contiguous() for k, v in self.state_dict().items()}
        if safe_serialization:
            from safetensors.torch import save_file
            save_file(sd, os.path.join(save_directory, self.weights_names[0]))
        else:
            torch.save(sd, os.path.join(save_directory, self.weights_names[1]))

    # ---- the kernel-layout (bf16, packed) copy of the weights follows the parameters: whatever changes them drops it, the
    # next forward / prepare() re-packs, and DualDiffDenoiser re-captures its CUDA graph when the pack object changed ----
    def invalidate_pack(self):
        for m in self.modules():
            if getattr(m, "_packed", None) is not None:
                m._packed = None
            if hasattr(m, "_prep_cache"):
                m._prep_cache = None

    def load_state_dict(self, *args, **kwargs):
        out = super().load_state_dict(*args, **kwargs)
        self.invalidate_pack()
        return out

    def _apply(self, fn, *args, **kwargs):      # .to() / .cuda() / .half() ...
        out = super()._apply(fn, *args, **kwargs)
        self.invalidate_pack()
        return out

    def ensure_packed(self, device=None):
        """pack on first use; re-pack when load_state_dict / .to() dropped the copy or a parameter was modified in place"""
        if getattr(self, "_packed", None) is None or self.pack_is_stale():
            self.pack(device)
        return self._packed

    def pack_is_stale(self):
        """True when a parameter was modified in place after pack() (optimizer step, set_category_token, copy_)"""
        ver = getattr(self, "_packed_versions", None)
        return ver is not None and ver != self._param_versions()

    def _param_versions(self):
        return tuple((id(t), t._version) for t in list(self.parameters()) + list(self.buffers()))

    @property
    def dtype(self):
        return torch.bfloat16 if getattr(self, "_packed", None) is not None else next(self.parameters()).dtype

    @property
    def device(self):
        return next(self.parameters()).device

    def enable_xformers_memory_efficient_attention(self, *a, **k):
        return None  # accepted, no-op: attention is always the fused tcgen05 kernel (misc/test_utils.py:164-165)

    def enable_gradient_checkpointing(self, *a, **k):
        return None

    @property
    def attn_processors(self):
        out = {}
        for name, m in self.named_modules():
            if hasattr(m, "set_processor"):
                out[f"{name}.processor"] = m.processor
        return out

    def set_attn_processor(self, processor):
        """diffusers signature (unet_addon_rawbox.py:523-593).  Only the default / stock-arithmetic processors are accepted:
        the fused kernel cannot call a Python processor, and silently ignoring one would change the model's output."""
        mods = {f"{n}.processor": m for n, m in self.named_modules() if hasattr(m, "set_processor")}
        if isinstance(processor, dict):
            if len(processor) != len(mods):
                raise ValueError(
                    f"A dict of processors was passed, but the number of processors {len(processor)} does not match the"
                    f" number of attention layers: {len(mods)}. Please make sure to pass {len(mods)} processor classes.")
            for k, m in mods.items():
                m.set_processor(processor[k])
        else:
            for m in mods.values():
                m.set_processor(processor)

    def set_default_attn_processor(self):
        self.set_attn_processor(None)
