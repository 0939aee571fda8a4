"""Drop-in mirrors of the reference's network classes (same names / ctor kwargs / forward signatures /
state-dict keys as MD_txt_con_fusion/magicdrive/networks/*), executing on the CUDA engine."""
from .blocks import BasicMultiviewTransformerBlock  # noqa: F401
from .output_cls import BEVControlNetOutput, UNet2DConditionOutput  # noqa: F401
from .unet_2d_condition_multiview import UNet2DConditionModelMultiview  # noqa: F401
from .unet_addon_rawbox import BEVControlNetModel  # noqa: F401
from .occ3d_proj import OccupancyRay  # noqa: F401
from .vae import AutoencoderKLDecoder  # noqa: F401
from .clip_text import CLIPTextConfig, CLIPTextModel  # noqa: F401
