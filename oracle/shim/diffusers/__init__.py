"""ORACLE / TEST INFRASTRUCTURE ONLY — not part of the product path.

A pure-PyTorch restatement of the subset of HuggingFace diffusers **v0.17.1** that DualDiff's network
code imports (pin: MD_txt_con_fusion/sd-controlnet-seg/config.json:3; install path README.md:70-72; the
package itself is NOT vendored under /root/reference, so this restates its published semantics — see
SURVEY.md Appendix A).  With this directory on sys.path the reference's own
magicdrive/networks/*.py import unmodified and run on CPU in fp32.

Parity status: the reference ships no tests / golden vectors for this path ("parity unpinned" by the
reference); this shim is self-checked against torch primitives and the SDv1.5 parameter count
(859,520,964) in tests/test_oracle.py.
"""
from .models.unet_2d_condition import UNet2DConditionModel  # noqa: F401

__version__ = "0.17.1+oracle-shim"
