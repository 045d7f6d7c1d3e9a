"""`Decoder` — same constructor and call contract as reference ssds/modeling/layers/decoder.py:15-49,
backed by the fused all-level decode and the batched NMS kernel (3 launches per batch instead of
B x (L x ~20 + ~1500) eager kernels)."""
from .box import decode_levels, nms


class Decoder(object):
    def __init__(self, conf_threshold, nms_threshold, top_n, top_n_per_level, rescore, use_diou):
        self.conf_threshold = conf_threshold
        self.nms_threshold = nms_threshold
        self.top_n = top_n
        self.top_n_per_level = top_n_per_level
        self.rescore = rescore
        self.use_diou = use_diou

    def __call__(self, loc, conf, anchors, return_indices=False, packed_out=None):
        """loc/conf: tuples of per-level maps; anchors: OrderedDict{stride: base_anchors[A,4]}.
        Returns (scores [B,top_n], boxes [B,top_n,4] ltrb, classes [B,top_n]) zero padded.
        packed_out (extension): a [B,top_n,6] tensor the NMS kernel also fills with (score, box, class);
        `packed_out=True` returns only that block."""
        decoded = decode_levels(conf, loc, list(anchors.items()), self.conf_threshold,
                                self.top_n_per_level, self.rescore)
        return nms(*decoded, self.nms_threshold, self.top_n, using_diou=self.use_diou,
                   return_indices=return_indices, packed_out=packed_out)
