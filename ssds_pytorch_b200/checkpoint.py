"""Weight import from the reference's checkpoints — host-side only.

Mirrors ssds/core/checkpoint.py of the reference for the engine's needs:
    find_previous_checkpoint(output_dir)                    checkpoint.py:38-56   (checkpoint_list.txt index)
    load_checkpoint(path, resume_scope, base_state_dict)    checkpoint.py:59-133  (resume_checkpoint:
        "state_dict" unwrapping, `module.` stripping :87-91, scope filter :110-117, keep-only-known keys :119,
        un-resumed keys keep the base (initial) values :125-131)
and adds what a weight-packing consumer needs on top: `check_state_dict` validates key names and shapes
against the layer table of the configured model (synth.model_shapes = the reference's state_dict()
order/shapes, pinned by tests/test_model_oracle_golden.py) BEFORE anything is packed for the GPU.

    python -m ssds_pytorch_b200.checkpoint --cfg experiments/x.yml --checkpoint exp/ssd_resnet50_epoch_100.pth
        [--scope backbone,extras] [--dry-run]

prints the import report; without --dry-run it also builds the detector on the current CUDA device
(BN folding + NHWC-bf16 packing happen in model.py) and runs one synthetic image through it.
"""
import argparse
import os
import re
import sys
from collections import OrderedDict

import torch

from . import synth
from .model import number_box_from_cfg
from .ssds import load_cfg


def find_previous_checkpoint(output_dir):
    """(epochs, files) listed in `checkpoint_list.txt`, or False (checkpoint.py:38-56)."""
    index = os.path.join(output_dir, "checkpoint_list.txt")
    if not os.path.exists(index):
        return False
    epochs, files = [], []
    entry = re.compile(r"epoch\s+(\d+):\s(.*)$")          # "epoch {n}: {path}" as save_checkpoints writes it
    with open(index) as f:
        for line in f:
            m = entry.search(line.rstrip("\n"))
            if m:
                epochs.append(int(m.group(1)))
                files.append(m.group(2))
    return epochs, files


def strip_module_prefix(state):
    """checkpoint.py:87-91: a (Distributed)DataParallel checkpoint carries `module.` on every key; the
    reference decides by looking at the FIRST key only and then drops the first dotted component of all."""
    if state and "module." in next(iter(state)):
        return OrderedDict((k.partition(".")[2], v) for k, v in state.items())
    return state


def load_checkpoint(path, resume_scope="", base_state_dict=None):
    """Returns (state_dict, report).  `base_state_dict` plays the role of `model.state_dict()` in the
    reference: only keys it contains are taken, and its values remain for everything not resumed.
    With base_state_dict=None every (scope-filtered) checkpoint entry is returned as is."""
    if not path or not os.path.isfile(path):
        raise FileNotFoundError(f"no checkpoint found at '{path}'")       # reference prints and returns False
    ckpt = torch.load(path, map_location=torch.device("cpu"), weights_only=True)
    if "state_dict" in ckpt:
        ckpt = ckpt["state_dict"]
    ckpt = strip_module_prefix(OrderedDict(ckpt))
    if resume_scope != "":
        scopes = resume_scope.split(",")
        ckpt = OrderedDict((k, v) for k, v in ckpt.items() if any(s in k for s in scopes))
    report = {"file": path, "scope": resume_scope, "in_checkpoint": len(ckpt)}
    if base_state_dict is None:
        report.update(resumed=len(ckpt), unresumed=[], ignored=[])
        return ckpt, report
    resumed = OrderedDict((k, v) for k, v in ckpt.items() if k in base_state_dict)
    out = OrderedDict(base_state_dict)
    out.update(resumed)
    report.update(resumed=len(resumed), unresumed=sorted(set(base_state_dict) - set(resumed)),
                  ignored=sorted(set(ckpt) - set(resumed)))
    return out, report


def expected_shapes(cfg):
    m = load_cfg(cfg)["MODEL"]
    return OrderedDict(synth.model_shapes(m["SSDS"], m["NETS"], m["FEATURE_LAYER"], number_box_from_cfg(m),
                                          m["NUM_CLASSES"]))


def check_state_dict(cfg, state_dict):
    """Key/shape validation against the configured model.  Returns a dict with `missing`, `unexpected`,
    `mismatched` [(key, got, expected)]; `num_batches_tracked` buffers are optional (BN folding ignores them)."""
    want = expected_shapes(cfg)
    have = {k: tuple(v.shape) for k, v in state_dict.items()}
    optional = lambda k: k.endswith("num_batches_tracked")          # noqa: E731
    missing = [k for k in want if k not in have and not optional(k)]
    unexpected = [k for k in have if k not in want and not optional(k)]
    mismatched = [(k, have[k], tuple(want[k])) for k in want if k in have and have[k] != tuple(want[k])]
    return {"missing": missing, "unexpected": unexpected, "mismatched": mismatched,
            "ok": not missing and not mismatched}


def detector_from_checkpoint(cfg, path, resume_scope="", init_seed=0, device=None, use_graph=True):
    """cfg (dict / yml path with the reference's keys) + checkpoint -> SSDDetector on the B200 path.
    Parameters outside `resume_scope` (or absent from the file) keep reference-style initial values, like
    a freshly constructed reference model would."""
    from .ssds import SSDDetector
    cfg = load_cfg(cfg)
    m = cfg["MODEL"]
    base = synth.synthetic_state_dict(m["NETS"], m["FEATURE_LAYER"], number_box_from_cfg(m), m["NUM_CLASSES"],
                                      seed=init_seed, style="init", ssds=m["SSDS"])
    sd, report = load_checkpoint(path, resume_scope, base)
    chk = check_state_dict(cfg, sd)
    if not chk["ok"]:
        raise ValueError(f"checkpoint does not fit the configured model: missing {chk['missing'][:5]}, "
                         f"mismatched {chk['mismatched'][:5]}")
    return SSDDetector(cfg, sd, device=device, use_graph=use_graph), report


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    ap.add_argument("--cfg", required=True, help="yml with the reference's config keys")
    ap.add_argument("--checkpoint", default="", help=".pth file, or an EXP_DIR holding checkpoint_list.txt")
    ap.add_argument("--scope", default="", help="cfg.TRAIN.RESUME_SCOPE semantics: comma-separated substrings")
    ap.add_argument("--dry-run", action="store_true", help="validate only; no CUDA needed")
    a = ap.parse_args(argv)
    path = a.checkpoint
    if os.path.isdir(path):
        prev = find_previous_checkpoint(path)
        if not prev or not prev[1]:
            print(f"no checkpoint_list.txt entries under {path}", file=sys.stderr)
            return 2
        path = prev[1][-1]
        print(f"latest checkpoint: epoch {prev[0][-1]}: {path}")
    sd, report = load_checkpoint(path, a.scope, None)
    chk = check_state_dict(a.cfg, sd)
    print(f"{report['in_checkpoint']} tensors in scope; missing {len(chk['missing'])}, unexpected "
          f"{len(chk['unexpected'])}, shape mismatches {len(chk['mismatched'])}")
    for k in chk["missing"][:10]:
        print("  missing   ", k)
    for k, got, exp in chk["mismatched"][:10]:
        print("  mismatched", k, got, "expected", exp)
    if a.dry_run:
        return 0 if (not chk["mismatched"]) else 1
    det, rep = detector_from_checkpoint(a.cfg, path, a.scope)
    print(f"resumed {rep['resumed']} tensors, {len(rep['unresumed'])} keep their initial values")
    import numpy as np
    h, w = det.image_size
    scores, boxes, classes = det(np.zeros((h, w, 3), np.uint8))
    print("detector built; one synthetic image ->", scores.shape, boxes.shape, classes.shape)
    return 0


if __name__ == "__main__":
    sys.exit(main())
