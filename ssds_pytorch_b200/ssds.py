"""`SSDDetector` — the inference facade of the reference (ssds/ssds.py:14-68) on the B200 path.

    det = SSDDetector(cfg, state_dict)           # cfg: dict with the reference's yml keys
    scores, boxes, classes = det(imgs)           # numpy in -> numpy out, same contract

`cfg` uses the reference's config keys (ssds/core/config.py): MODEL.{SSDS,NETS,IMAGE_SIZE,
NUM_CLASSES,FEATURE_LAYER,SIZES,ASPECT_RATIOS}, POST_PROCESS.{SCORE_THRESHOLD,IOU_THRESHOLD,
MAX_DETECTIONS,MAX_DETECTIONS_PER_LEVEL,USE_DIOU,RESCORE_CENTER}, DATASET.PREPROC.{MEAN,STD}; a path to
a yml file is accepted too.  On the tcgen05 conv stack: SSD / SSDFPN / SSDBiFPN over ResNet18-152 and
RegNetX032, SSD over MobileNetV2, YOLOV3 / YOLOV4 over ResNet (model.ENGINES); FSSD/FCOS are out of scope (SURVEY 2
rows 9-11).
`ssds_pytorch_b200.checkpoint.detector_from_checkpoint` builds one from a reference `.pth`.

One process per GPU.  Under torch.distributed each rank runs its shard of the batch and
`gather_detections` all-gathers the fixed-size [B,D,6] detection block over NCCL — the only
collective on the path.
"""
import copy

import numpy as np
import torch

from .decoder import Decoder
from .model import engine_for, create_anchors, number_box_from_cfg

DEFAULTS = {   # the values of ssds/core/config.py:42-74,166-175,206-207 that this path reads
    "MODEL": {"SSDS": "SSD", "NETS": "ResNet50", "IMAGE_SIZE": [300, 300], "NUM_CLASSES": 21},
    "POST_PROCESS": {"SCORE_THRESHOLD": 0.01, "IOU_THRESHOLD": 0.6, "MAX_DETECTIONS": 100,
                     "MAX_DETECTIONS_PER_LEVEL": 300, "USE_DIOU": True, "RESCORE_CENTER": True},
    "DATASET": {"PREPROC": {"MEAN": 0, "STD": 255}},
}


def _merge(base, over):
    out = copy.deepcopy(base)
    for k, v in (over or {}).items():
        if isinstance(v, dict) and isinstance(out.get(k), dict):
            out[k] = _merge(out[k], v)
        else:
            out[k] = v
    return out


def load_cfg(cfg):
    if isinstance(cfg, str):
        import yaml
        with open(cfg) as f:
            cfg = yaml.safe_load(f)
    return _merge(DEFAULTS, cfg)


def create_decoder(pp):
    """reference model_builder.py:59-74."""
    return Decoder(pp["SCORE_THRESHOLD"], pp["IOU_THRESHOLD"], pp["MAX_DETECTIONS"],
                   pp["MAX_DETECTIONS_PER_LEVEL"], pp["RESCORE_CENTER"], pp["USE_DIOU"])


class SSDDetector(object):
    def __init__(self, cfg, state_dict, device=None, use_graph=True):
        cfg = load_cfg(cfg)
        m = cfg["MODEL"]
        engine = engine_for(m["SSDS"], m["NETS"])
        self.cfg = cfg
        self.device = (torch.device(device) if device is not None
                       else torch.device("cuda", torch.cuda.current_device()))
        self.mean = float(cfg["DATASET"]["PREPROC"]["MEAN"])
        self.std = float(cfg["DATASET"]["PREPROC"]["STD"])
        self.model = engine(state_dict, m["FEATURE_LAYER"], m["NUM_CLASSES"], number_box_from_cfg(m),
                            device=self.device, mean=self.mean, std=self.std).eval()
        self.image_size = tuple(m["IMAGE_SIZE"])
        self.num_classes = m["NUM_CLASSES"]
        self.anchors = create_anchors(m, self.model, m["IMAGE_SIZE"])
        self.decoder = create_decoder(cfg["POST_PROCESS"])
        self.use_graph = use_graph
        self._stage = {}
        self._copy_stream = None
        self._post_stream = None        # decode + NMS of step i overlap the backbone of step i+1
        self._post_done = None
        self._post_out = None

    # -------------------------------------------------------------- device-side entry points
    def detect_device(self, images, overlap=False, packed_out=None):
        """images already on the GPU: uint8 NHWC [B,H,W,3] (raw pixels) or fp32 NCHW [B,3,H,W] raw
        pixel values; normalisation (x-mean)/std (ssds.py:57) is fused into the first kernel.
        Returns device tensors (scores [B,D], boxes [B,D,4], classes [B,D]); with `packed_out` (a [B,D,6]
        fp32 tensor) the NMS kernel also fills (score, x1, y1, x2, y2, class) there.

        overlap=True runs decode + NMS on a side stream so that they overlap the conv stack of the next
        call (the memory-/latency-bound post-processing hides behind tensor-core work); the results are
        then produced on that stream — call `join()` before consuming them on the current stream (join makes
        the current stream wait for the side stream and hands the returned tensors over to it with
        `record_stream`, so the caching allocator cannot recycle them under a pending reader)."""
        if not overlap:
            loc, conf = self.model(images, use_graph=self.use_graph)
            return self.decoder(loc, conf, self.anchors, packed_out=packed_out)
        if self._post_stream is None:
            self._post_stream = torch.cuda.Stream(device=self.device)
        main = torch.cuda.current_stream()
        loc, conf = self.model(images, use_graph=self.use_graph, outputs_free=self._post_done)
        heads_done = torch.cuda.Event()
        heads_done.record(main)
        with torch.cuda.stream(self._post_stream):
            self._post_stream.wait_event(heads_done)
            out = self.decoder(loc, conf, self.anchors, packed_out=packed_out)
            self._post_done = torch.cuda.Event()
            self._post_done.record(self._post_stream)
        self._post_out = out if isinstance(out, (tuple, list)) else (out,)
        return out

    def join(self):
        """Make the current stream wait for the side-stream post-processing of the last call.  Tensors returned
        by `detect_device(overlap=True)` were allocated on the side stream: they are handed to the current
        stream here (`record_stream`) so that freeing them cannot let the allocator reuse their memory on
        the side stream while the current stream still reads them."""
        if self._post_done is not None:
            cur = torch.cuda.current_stream()
            cur.wait_event(self._post_done)
            for t in self._post_out or ():
                if isinstance(t, torch.Tensor) and t.is_cuda:
                    t.record_stream(cur)
            self._post_out = None

    def detect_host(self, imgs, out=None, slot=0, gather=False):
        """Host numpy/torch batch -> pinned staging -> H2D -> detect -> [all-gather] -> packed
        [B,D,6] on the host (score, x1, y1, x2, y2, class).  The H2D copy runs on a side stream into
        one of two device slots, so the copy of step i+1 overlaps the kernels of step i when callers
        alternate `slot`.  One small D2H copy into the returned pinned tensor, issued on the post-processing
        side stream: the result is valid only after `det.join()` followed by a synchronize of the current
        stream (what `__call__` does); syncing the current stream alone is NOT enough.
        Unpinned input is first copied into the slot's pinned staging buffer; that host copy waits for the
        slot's previous H2D transfer, so pipelined callers cannot overwrite bytes the GPU has not read yet."""
        t = torch.as_tensor(imgs)
        key = (tuple(t.shape), t.dtype, slot)
        st = self._stage.get(key)
        if st is None:
            D = self.decoder.top_n
            st = {"pin_in": torch.empty(t.shape, dtype=t.dtype).pin_memory(),
                  "dev_in": torch.empty(t.shape, dtype=t.dtype, device=self.device),
                  "dev_out": torch.empty((t.shape[0], D, 6), dtype=torch.float32, device=self.device),
                  "pin_out": None, "copied": torch.cuda.Event(), "consumed": torch.cuda.Event(),
                  "pin_used": False}
            st["consumed"].record()
            self._stage[key] = st
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=self.device)
        if not t.is_pinned():
            if st["pin_used"]:
                st["copied"].synchronize()        # the previous async H2D copy out of pin_in must have finished
            st["pin_in"].copy_(t)
            st["pin_used"] = True
            t = st["pin_in"]
        cur = torch.cuda.current_stream()
        self._copy_stream.wait_event(st["consumed"])      # the slot's previous user has read it
        with torch.cuda.stream(self._copy_stream):
            st["dev_in"].copy_(t, non_blocking=True)
            st["copied"].record()
        cur.wait_event(st["copied"])
        # the NMS kernel writes the packed [B,D,6] block itself (no packing launches on the post stream)
        self.detect_device(st["dev_in"], overlap=True, packed_out=st["dev_out"])
        st["consumed"].record()          # (the plan copied dev_in into its own buffer first)
        with torch.cuda.stream(self._post_stream):      # all-gather and D2H ride with decode/NMS
            o = st["dev_out"]
            if gather:
                o = gather_detections(o)
            if out is None:
                if st["pin_out"] is None or st["pin_out"].shape != o.shape:
                    st["pin_out"] = torch.empty(o.shape, dtype=torch.float32).pin_memory()
                out = st["pin_out"]
            out.copy_(o, non_blocking=True)
            self._post_done = torch.cuda.Event()
            self._post_done.record(self._post_stream)
        return out

    # -------------------------------------------------------------- reference-compatible call
    def __call__(self, imgs):
        """reference ssds/ssds.py:41-68: imgs np.ndarray [H,W,3], [3,H,W], [N,H,W,3] or [N,3,H,W]."""
        imgs = np.asarray(imgs)
        pick1st = False
        if len(imgs.shape) == 3:
            imgs = imgs[None, ...]
            pick1st = True
        if len(imgs.shape) != 4:
            raise AssertionError("image dims has to be 3 or 4")
        if imgs.dtype == np.uint8 and imgs.shape[3] == 3:
            batch = imgs                                    # raw HWC bytes: 4x fewer H2D bytes
        else:
            if imgs.shape[3] == 3:
                imgs = imgs.transpose(0, 3, 1, 2)
            batch = np.ascontiguousarray(imgs, dtype=np.float32)
        with torch.cuda.device(self.device):
            det = self.detect_host(batch)
            self.join()
            torch.cuda.current_stream().synchronize()
        det = det.numpy()
        out_scores, out_boxes, out_classes = det[:, :, 0].copy(), det[:, :, 1:5].copy(), det[:, :, 5].copy()
        if pick1st:
            return out_scores[0], out_boxes[0].astype(int), out_classes[0].astype(int)
        return out_scores, out_boxes.astype(int), out_classes.astype(int)


def gather_detections(det_local, group=None, n_items=None):
    """All-gather the per-rank detection block [B_local, D, 6] -> [B_global, D, 6] (NCCL over
    NVLink on the GPU box, gloo in the CPU tests).  Ranks hold contiguous shards of the batch.

    `all_gather_into_tensor` needs the same shard size on every rank: pass `n_items` (the global batch)
    when it is not divisible by the world size — every rank then pads its `shard_batch` shard with zero rows
    to ceil(n_items / world) before the collective and the padding is trimmed afterwards.  Without
    `n_items` the shards must be equal (checked with a cheap all-reduce only in debug mode, so stated here)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return det_local
    world = dist.get_world_size(group)
    per = det_local.shape[0]
    if n_items is not None:
        per = (n_items + world - 1) // world
        if det_local.shape[0] > per:
            raise ValueError(f"gather_detections: local shard of {det_local.shape[0]} rows exceeds ceil({n_items}/{world})")
        if det_local.shape[0] < per:
            pad = torch.zeros((per - det_local.shape[0],) + tuple(det_local.shape[1:]), dtype=det_local.dtype,
                              device=det_local.device)
            det_local = torch.cat([det_local, pad], 0)
    out = torch.empty((world * per,) + tuple(det_local.shape[1:]), dtype=det_local.dtype, device=det_local.device)
    dist.all_gather_into_tensor(out, det_local.contiguous(), group=group)
    if n_items is not None and world * per != n_items:
        # rank r holds rows [r*per, min((r+1)*per, n_items)): drop each rank's padding
        keep = [out[r * per:r * per + max(0, min(per, n_items - r * per))] for r in range(world)]
        out = torch.cat(keep, 0)
    return out


def shard_batch(n_items, rank, world):
    """Contiguous shard [lo, hi) of a global batch for this rank."""
    per = (n_items + world - 1) // world
    lo = min(rank * per, n_items)
    return lo, min(lo + per, n_items)
