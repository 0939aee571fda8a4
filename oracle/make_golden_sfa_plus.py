"""Runs the REFERENCE's own txt_con_XFormersAttn_plus (networks/txt_con_fusion.py:184-337, unmodified, imported from
/root/reference on top of oracle/shim) on seeded weights and inputs and stores inputs + output as
tests/golden/sfa_plus_small.pt.
    python oracle/make_golden_sfa_plus.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle", "shim"))
sys.path.insert(0, "/root/reference/MD_txt_con_fusion")


def inputs(seed=0, n=2, h=6, w=10):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(n, 320, h, w, generator=g), torch.randn(n, 77, 768, generator=g)


def main():
    from magicdrive.networks.txt_con_fusion import txt_con_XFormersAttn_plus
    from dualdiff_b200 import synthetic as S
    m = txt_con_XFormersAttn_plus()
    sd = S.init_state_dict({k: tuple(v.shape) for k, v in m.state_dict().items()}, seed=5)
    m.load_state_dict(sd, strict=True)
    cond, txt = inputs()
    with torch.no_grad():
        out = m(None, cond, encoder_hidden_states=txt)
    path = os.path.join(ROOT, "tests", "golden", "sfa_plus_small.pt")
    torch.save({"weight_seed": 5, "input_seed": 0, "shape": tuple(cond.shape), "out": out}, path)   # inputs are regenerated from the seed
    print("wrote", path, tuple(out.shape), float(out.abs().mean()))


if __name__ == "__main__":
    main()
