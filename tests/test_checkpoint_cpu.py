"""Weight import: ssds_pytorch_b200.checkpoint against the reference's own
ssds/core/checkpoint.py — checkpoints WRITTEN by the reference's `save_checkpoints`, read back by both the
reference's `resume_checkpoint` (into a fresh reference model) and our `load_checkpoint`, results compared
tensor by tensor.  Needs /root/reference (authoring container); the self-contained cases below it do not.
"""
import os
import sys
import warnings
from collections import OrderedDict, defaultdict

import pytest
import torch

from ssds_pytorch_b200 import checkpoint as CK
from ssds_pytorch_b200 import synth

REF = os.environ.get("SSDS_REFERENCE", "/root/reference")
FL = [[3, 4, 5, "Conv:S"], [128, 256, 512, 256]]
CFG = {"MODEL": {"SSDS": "SSD", "NETS": "ResNet18", "IMAGE_SIZE": [96, 160], "NUM_CLASSES": 20, "FEATURE_LAYER": FL,
                 "SIZES": [[2.0, 2.828]] * 4, "ASPECT_RATIOS": [[1, 2, 0.5]] * 4}}


def _sd(seed):
    return synth.synthetic_state_dict("ResNet18", FL, [6] * 4, 20, seed=seed, style="test")


def test_index_module_prefix_scope_and_validation(tmp_path):
    sd = _sd(1)
    f1, f2 = str(tmp_path / "m_epoch_1.pth"), str(tmp_path / "m_epoch_7.pth")
    torch.save(sd, f1)
    torch.save({"state_dict": OrderedDict(("module." + k, v) for k, v in _sd(2).items())}, f2)   # DDP-style wrapper
    with open(tmp_path / "checkpoint_list.txt", "w") as f:
        f.write(f"epoch 1: {f1}\nepoch 7: {f2}\n")
    assert CK.find_previous_checkpoint(str(tmp_path)) == ([1, 7], [f1, f2])
    assert CK.find_previous_checkpoint(str(tmp_path / "nope")) is False
    got, rep = CK.load_checkpoint(f2)
    assert list(got) == list(sd) and rep["resumed"] == len(sd)
    assert all(torch.equal(got[k], v) for k, v in _sd(2).items())
    assert CK.check_state_dict(CFG, got)["ok"]
    # scope: only matching keys are taken, everything else keeps the base values (checkpoint.py:110-131)
    base = _sd(3)
    got, rep = CK.load_checkpoint(f2, "backbone.layer1,loc", base)
    for k in base:
        src = _sd(2)[k] if ("backbone.layer1" in k or "loc" in k) else base[k]
        assert torch.equal(got[k], src), k
    assert set(rep["unresumed"]) == {k for k in base if "backbone.layer1" not in k and "loc" not in k}
    # validation catches a foreign checkpoint before anything is packed
    bad = OrderedDict(sd)
    bad["backbone.conv1.weight"] = torch.zeros(32, 3, 7, 7)
    del bad["loc.0.weight"]
    bad["something.else"] = torch.zeros(1)
    chk = CK.check_state_dict(CFG, bad)
    assert not chk["ok"] and chk["missing"] == ["loc.0.weight"] and chk["unexpected"] == ["something.else"]
    assert chk["mismatched"] == [("backbone.conv1.weight", (32, 3, 7, 7), (64, 3, 7, 7))]
    with pytest.raises(FileNotFoundError):
        CK.load_checkpoint(str(tmp_path / "missing.pth"))
    assert CK.main(["--cfg", _write_cfg(tmp_path), "--checkpoint", str(tmp_path), "--dry-run"]) == 0


def _write_cfg(tmp_path):
    import yaml
    p = str(tmp_path / "cfg.yml")
    with open(p, "w") as f:
        yaml.safe_dump(CFG, f)
    return p


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "ssds")), reason="needs the reference tree")
def test_against_reference_save_and_resume(tmp_path, capsys):
    sys.path.insert(0, REF)
    warnings.filterwarnings("ignore")
    import torchvision as tv
    tv.models.resnet.model_urls = defaultdict(lambda: None)        # no download
    tv.models.densenet.model_urls = defaultdict(lambda: None)
    tv.models.mobilenet.model_urls = defaultdict(lambda: None)
    try:
        from ssds.core import checkpoint as RCK
        from ssds.core import config as rcfg
        from ssds.modeling import model_builder
    except Exception as e:                                         # reference imports that this image lacks
        pytest.skip(f"reference not importable here: {e}")
    m = rcfg.cfg.MODEL
    m.SSDS, m.NETS, m.IMAGE_SIZE, m.NUM_CLASSES = "SSD", "ResNet18", [96, 160], 20
    m.FEATURE_LAYER, m.SIZES, m.ASPECT_RATIOS = FL, CFG["MODEL"]["SIZES"], CFG["MODEL"]["ASPECT_RATIOS"]
    torch.manual_seed(0)
    trained = model_builder.create_model(m)
    with torch.no_grad():
        for p in trained.parameters():
            p.add_(torch.randn_like(p) * 0.01)
    RCK.save_checkpoints(torch.nn.DataParallel(trained), str(tmp_path), "ssd_resnet18_x", 3)     # `module.` keys
    epochs, files = CK.find_previous_checkpoint(str(tmp_path))
    assert (epochs, files) == RCK.find_previous_checkpoint(str(tmp_path))
    for scope in ("", "backbone,extras"):
        torch.manual_seed(1)
        fresh = model_builder.create_model(m)
        base = OrderedDict((k, v.clone()) for k, v in fresh.state_dict().items())
        RCK.resume_checkpoint(fresh, files[-1], scope)
        got, rep = CK.load_checkpoint(files[-1], scope, base)
        ref = fresh.state_dict()
        assert list(got) == list(ref)
        for k in ref:
            assert torch.equal(got[k], ref[k]), (scope, k)
        assert CK.check_state_dict(CFG, got)["ok"]
        if scope:
            assert rep["unresumed"] and all("backbone" not in k and "extras" not in k for k in rep["unresumed"])
    capsys.readouterr()
