"""ORACLE / TEST INFRASTRUCTURE ONLY.  Builds the reference's OWN network classes, imported unmodified from
/root/reference/MD_txt_con_fusion on top of oracle/shim (diffusers 0.17.1 / xformers / accelerate stand-ins),
wired the way misc/test_utils.py:97-171 (build_pipe) wires the dual-branch model.  Only usable where
/root/reference exists (this container) — used by oracle/make_golden.py and tests that pin the oracle."""
import json
import os
import sys

import torch

REF_ROOT = "/root/reference/MD_txt_con_fusion"
SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shim")
NEIGHBORS = {0: [5, 1], 1: [0, 2], 2: [1, 3], 3: [2, 4], 4: [3, 5], 5: [4, 0]}


def available():
    return os.path.isdir(REF_ROOT)


def _paths():
    for p in (SHIM, REF_ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)


def controlnet_config(block_out=(320, 640, 1280, 1280)):
    cfg = json.load(open(os.path.join(REF_ROOT, "sd-controlnet-seg", "config.json")))
    cfg = {k: v for k, v in cfg.items() if not k.startswith("_")}
    cfg["block_out_channels"] = list(block_out)
    return cfg


def build_unet(block_out=(320, 640, 1280, 1280), device="cpu"):
    _paths()
    from magicdrive.networks.unet_2d_condition_multiview import UNet2DConditionModelMultiview
    with torch.device(device):
        m = UNet2DConditionModelMultiview(
            cross_attention_dim=768, block_out_channels=tuple(block_out), neighboring_view_pair=NEIGHBORS,
            neighboring_attn_type="add", zero_module_type="zero_linear", crossview_attn_type="basic")
    return m.eval()


def build_branch(use_occ_3d: bool, block_out=(320, 640, 1280, 1280), device="cpu"):
    """one BEVControlNetModel wired like misc/test_utils.py:123-136 (attributes set by hand after load)"""
    _paths()
    from magicdrive.networks.unet_addon_rawbox import BEVControlNetModel
    with torch.device(device):
        c = BEVControlNetModel(**controlnet_config(block_out))
    c.use_cam_in_temb = False
    c.use_box_adapter = False
    c.adm_proj = None
    c.use_txt_con_fusion = True
    c.use_txt_con_fusionp = False
    c.txt_con_fusionp = None
    c.use_occ_3d = use_occ_3d
    if use_occ_3d:
        c.controlnet_cond_embedding = None
    return c.eval()
