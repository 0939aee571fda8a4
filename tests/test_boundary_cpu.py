"""Host-side contract of the drop-in boundary: attention
processors, checkpoint loading strictness, staleness of the packed weights, the cross-view topology, 40-point map vectors."""
import json
import logging
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import common  # noqa: E402


@pytest.fixture(scope="module")
def models():
    return common.build_models(load=False)      # meta tensors: structure only


def test_default_and_stock_processors_are_accepted_custom_ones_raise(models):
    from dualdiff_b200.networks._tree import FusedAttnProcessor
    unet, nets, _ = models
    procs = unet.attn_processors
    assert len(procs) == 16 * 3 and all(isinstance(p, FusedAttnProcessor) for p in procs.values())   # attn1/2/4 of 16 blocks
    assert len(nets[0].attn_processors) == 7 * 2
    unet.set_default_attn_processor()
    XFormersAttnProcessor = type("XFormersAttnProcessor", (), {})            # what enable_xformers... installs upstream
    unet.set_attn_processor(XFormersAttnProcessor())
    unet.set_attn_processor({k: XFormersAttnProcessor() for k in procs})
    Adapter = type("Adapter_XFormersAttnProcessor", (), {})                   # box_adapter.py:177 -- changes the arithmetic
    with pytest.raises(NotImplementedError, match="Adapter_XFormersAttnProcessor"):
        unet.set_attn_processor(Adapter())
    with pytest.raises(NotImplementedError):
        nets[1].set_attn_processor({k: Adapter() for k in nets[1].attn_processors})
    with pytest.raises(ValueError, match="number of processors"):
        unet.set_attn_processor({"x": None})
    assert all(isinstance(p, FusedAttnProcessor) for p in unet.attn_processors.values())


def test_cross_view_topology_comes_from_the_model_or_raises():
    from dualdiff_b200 import engine
    from dualdiff_b200.networks import BasicMultiviewTransformerBlock
    pairs, k = engine.check_view_pairs({"0": [5, 1], "1": [0, 2], "2": [1, 3], "3": [2, 4], "4": [3, 5], "5": [4, 0]})
    assert k == 2 and pairs[0] == [5, 1]
    assert engine.check_view_pairs({0: [1], 1: [2], 2: [0]})[1] == 1
    for bad in ({0: [1, 2], 1: [0]}, {0: [1, 2, 3], 1: [0, 2, 3], 2: [0, 1, 3], 3: [0, 1, 2]}, {0: [7], 1: [0]}, {1: [0], 2: [1]}):
        with pytest.raises(NotImplementedError):
            engine.check_view_pairs(bad)
    km = engine.make_kv_map(6, {0: [2, 1], 1: [0, 2], 2: [1, 0]}, "cpu")
    assert km.tolist() == [[2, 1], [0, 2], [1, 0], [5, 4], [3, 5], [4, 3]]
    with torch.device("meta"):
        blk = BasicMultiviewTransformerBlock(320, 8, 40, cross_attention_dim=768, neighboring_view_pair={0: [1], 1: [0]})
    assert blk.n_cam == 2


def _tiny_cls():
    from dualdiff_b200.networks import _tree

    class Tiny(_tree.ModelBase):
        def __init__(self, width=4):
            super().__init__()
            self.config = _tree.AttrDict(width=width)
            self.a = torch.nn.Linear(width, width)
            self.b = torch.nn.Linear(width, 2)
            self._packed = None

        def pack(self, device=None):
            self._packed = {k: v.clone() for k, v in self.state_dict().items()}
            self._packed_versions = self._param_versions()
            return self
    return Tiny


def test_from_pretrained_reports_missing_and_raises_on_mismatch(tmp_path, caplog):
    """ModelBase.from_pretrained (the loader of the UNet and of the ControlNet branches, misc/test_utils.py:111-113,146-147):
    strict like diffusers -- missing / unexpected keys are reported, a shape mismatch raises unless ignore_mismatched_sizes"""
    from safetensors.torch import load_file, save_file
    Tiny = _tiny_cls()
    m = Tiny()
    m.save_pretrained(tmp_path / "ok")
    again = Tiny.from_pretrained(str(tmp_path / "ok"))
    assert torch.equal(again.a.weight, m.a.weight) and again._load_report == dict(missing=[], unexpected=[], mismatched=[])
    t = load_file(str(tmp_path / "ok" / Tiny.weights_names[0]))
    t.pop("a.bias")
    t["not.a.key"] = torch.zeros(3)
    t["b.weight"] = torch.zeros(3, 4)
    os.makedirs(tmp_path / "bad")
    save_file(t, str(tmp_path / "bad" / Tiny.weights_names[0]))
    with open(tmp_path / "ok" / "config.json") as fh, open(tmp_path / "bad" / "config.json", "w") as out:
        out.write(fh.read())
    with pytest.raises(RuntimeError, match="size mismatch"):
        Tiny.from_pretrained(str(tmp_path / "bad"))
    with caplog.at_level(logging.WARNING):
        m2 = Tiny.from_pretrained(str(tmp_path / "bad"), ignore_mismatched_sizes=True)
    assert m2._load_report == dict(missing=["a.bias"], unexpected=["not.a.key"], mismatched=["b.weight"])
    assert torch.equal(m2.a.weight, m.a.weight)
    text = caplog.text
    assert "missing from the checkpoint" in text and "unexpected key" in text and "NOT loaded" in text


def test_packed_weights_follow_the_parameters():
    """load_state_dict / .to() drop the kernel-layout copy, an in-place update marks it stale"""
    m = _tiny_cls()()
    m.ensure_packed()
    first = m._packed
    assert m.ensure_packed() is first and not m.pack_is_stale()
    with torch.no_grad():
        m.a.weight.mul_(2.0)                      # in-place update (optimizer step, set_category_token, copy_)
    assert m.pack_is_stale()
    assert m.ensure_packed() is not first and torch.equal(m._packed["a.weight"], m.a.weight)
    m.load_state_dict(m.state_dict())
    assert m._packed is None
    m.ensure_packed()
    m.to(torch.float64)
    assert m._packed is None


def test_bbox_embedder_reinitialize_for_40_point_map_vectors():
    from dualdiff_b200.networks.bbox_embedder import ContinuousBBoxWithTextEmbedding
    e = ContinuousBBoxWithTextEmbedding(n_classes=3, mode="all-xyz", minmax_normalize=False, embedder_num_freq=4,
                                        proj_dims=[768, 512, 512, 768])
    assert e.bbox_proj.in_features == 27 * 8 and e.null_pos_feature.shape == (27 * 8,)
    e.reinitialize()
    assert e.bbox_proj.in_features == 27 * 40 and e.bbox_proj.out_features == 768 and e.null_pos_feature.shape == (27 * 40,)
    assert set(e.state_dict()) >= {"bbox_proj.weight", "bbox_proj.bias", "null_pos_feature", "_class_tokens"}
