"""The `ssds._C` module the reference names (box.py:3-4) with the legacy ODTK call signatures of
box.py:419-421 / :483-485, plus the two switches those signatures lack (rescore, using_diou; both
default True like config.py:173-174).  Install as `ssds._C` with `ssds_pytorch_b200.install()`."""
import torch

from . import box as _box


def decode(cls_head, box_head, anchors, stride, threshold, top_n, rescore=True):
    """decode_cuda(all_cls_head.float(), all_box_head.float(), anchors.view(-1).tolist(), stride,
    threshold, top_n) — box.py:419-421.  `anchors` may be the flat python list or an [A,4] tensor."""
    if not isinstance(anchors, torch.Tensor):
        anchors = torch.tensor(anchors, dtype=torch.float32).view(-1, 4)
    return list(_box.decode(cls_head, box_head, stride, threshold, top_n, anchors, rescore))


def nms(scores, boxes, classes, nms, ndetections, using_diou=True):
    """nms_cuda(all_scores.float(), all_boxes.float(), all_classes.float(), nms, ndetections) —
    box.py:483-485."""
    return list(_box.nms(scores, boxes, classes, nms, ndetections, using_diou))
