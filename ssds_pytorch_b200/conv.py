"""Host side of the tcgen05 conv stack: weight packing (BN folding, one-time) and thin launchers.

Layout contract (see csrc/conv_igemm.cu): activations NHWC bf16; weights bf16 [Cout, KH*KW, Cin]
with the eval-mode BatchNorm scale folded in; bias fp32 [Cout] = folded BN shift or conv bias.
"""
import ctypes as C

import torch

from . import _lib
from ._lib import lib, check, ptr, stream_ptr


def fold_bn(weight, bn=None, bias=None):
    """weight [Cout,Cin,KH,KW] fp32 (+ BatchNorm2d eval statistics) -> (folded weight, bias) fp32.
    y = gamma * (conv(x) - mean) / sqrt(var + eps) + beta  (torch BatchNorm2d, eval)."""
    w = weight.detach().float()
    cout = w.shape[0]
    b = bias.detach().float() if bias is not None else torch.zeros(cout)
    if bn is not None:
        gamma, beta, mean, var, eps = bn
        scale = gamma.detach().float() / torch.sqrt(var.detach().float() + eps)
        w = w * scale.view(-1, 1, 1, 1)
        b = (b - mean.detach().float()) * scale + beta.detach().float()
    return w, b


def pack_weight(w_folded, cin_pad=None):
    """[Cout,Cin,KH,KW] fp32 -> bf16 [Cout, KH*KW, Cin_pad] (K-major: tap-major, channel-minor)."""
    cout, cin, kh, kw = w_folded.shape
    cin_pad = cin_pad or cin
    out = torch.zeros((cout, kh * kw, cin_pad), dtype=torch.float32)
    out[:, :, :cin] = w_folded.permute(0, 2, 3, 1).reshape(cout, kh * kw, cin)
    return out.to(torch.bfloat16).contiguous()


def pack_stem_weight_s2d(w_folded, pad=None):
    """Stride-2 RGB stem [Cout,3,k,k] (7x7/p3 ResNet, 3x3/p1 MobileNet) -> the equivalent 4x4/s1 kernel
    on the 2x2 space-to-depth image: bf16 [Cout, 16 taps, 16 ch], channel (a*2+b)*3+c.
    Output row ho reads original rows 2ho-p+r = s2d row ho-2+kh', sub-row a  =>  r = 2(kh'-2)+a+p
    (zero weight where r falls outside the kernel)."""
    cout, _, k, _ = w_folded.shape
    pad = (k - 1) // 2 if pad is None else pad
    assert w_folded.shape[1] == 3 and k <= 7 and pad <= 3
    out = torch.zeros((cout, 4, 4, 16), dtype=torch.float32)
    for khp in range(4):
        for a in range(2):
            r = 2 * (khp - 2) + a + pad
            if r < 0 or r >= k:
                continue
            for kwp in range(4):
                for b in range(2):
                    s = 2 * (kwp - 2) + b + pad
                    if s < 0 or s >= k:
                        continue
                    for c in range(3):
                        out[:, khp, kwp, (a * 2 + b) * 3 + c] = w_folded[:, c, r, s]
    return out.reshape(cout, 16, 16).to(torch.bfloat16).contiguous()


def conv2d(x, w, bias, KH, KW, stride, pad, relu=False, residual=None, out=None, Ho=0, Wo=0,
           x_kind=0, x_width=None, chunk=0):
    """x NHWC bf16 [N,H,W,Cin] -> NHWC bf16 [N,Ho,Wo,Cout]; w bf16 [Cout, KH*KW, Cin].
    x_kind=1: x is the left-padded s2d image [N,H,row_px,16] (logical width x_width), windowed stem."""
    N, H, W, Cin = x.shape
    row_px = 0
    if x_width is not None:
        row_px, W = W, x_width
    Cout = Cin if chunk else w.shape[0]
    ho = Ho or (H + 2 * pad - KH) // stride + 1
    wo = Wo or (W + 2 * pad - KW) // stride + 1
    if out is None:
        out = torch.empty((N, ho, wo, Cout), dtype=torch.bfloat16, device=x.device)
    d = _lib.ConvDesc(N=N, H=H, W=W, Cin=Cin, Cout=Cout, KH=KH, KW=KW, stride=stride, pad=pad,
                      Ho=ho, Wo=wo, x_cstride=x.stride(2), out_cstride=out.stride(2),
                      res_cstride=residual.stride(2) if residual is not None else 0,
                      x_row_pixels=row_px, chunk=chunk, x_kind=x_kind,
                      w_rows=w.shape[0], relu=int(relu), out_mode=0, n_loc=0, sigmoid=0)
    with torch.cuda.device(x.device):
        check(lib.ssdsb_conv2d_bf16(C.byref(d), ptr(x), ptr(w), ptr(bias), ptr(residual), ptr(out),
                                    None, stream_ptr()), "conv2d")
    return out


def last_launch():
    """{block_n, block_k, ways, b_resident, stages, grid, groups, ghost} of this thread's last conv2d."""
    out = (C.c_int * 8)()
    check(lib.ssdsb_conv_last_launch(out), "conv_last_launch")
    return dict(zip(("block_n", "block_k", "ways", "b_resident", "stages", "grid", "groups", "ghost"), list(out)))


def sm_count():
    return torch.cuda.get_device_properties(torch.cuda.current_device()).multi_processor_count


def conv1x1_pair(x, w1, bias1, relu1, residual, w2, bias2, relu2, out1=None, out2=None):
    """y1 = act1(conv1x1(x, w1) + bias1 [+ residual]); y2 = act2(conv1x1(y1, w2) + bias2) in ONE launch
    (a Bottleneck's conv3 + the next block's conv1; y1 is re-read from L2, not HBM).  Dense NHWC bf16;
    w1 [Cmid, 1, Cin] / w2 [Cout2, 1, Cmid] as `pack_weight` makes them.  Returns (y1, y2)."""
    N, H, W, Cin = x.shape
    Cmid, Cout2 = w1.shape[0], w2.shape[0]
    assert x.is_contiguous() and w1.numel() == Cmid * Cin and w2.numel() == Cout2 * Cmid
    if out1 is None:
        out1 = torch.empty((N, H, W, Cmid), dtype=torch.bfloat16, device=x.device)
    if out2 is None:
        out2 = torch.empty((N, H, W, Cout2), dtype=torch.bfloat16, device=x.device)
    assert out1.is_contiguous() and out2.is_contiguous() and (residual is None or residual.is_contiguous())
    with torch.cuda.device(x.device):
        check(lib.ssdsb_conv1x1_pair_bf16(N, H, W, Cin, Cmid, Cout2, int(relu1), int(relu2), ptr(x), ptr(w1),
                                          ptr(bias1), ptr(residual), ptr(out1), ptr(w2), ptr(bias2),
                                          ptr(out2), stream_ptr()), "conv1x1_pair")
    return out1, out2


def conv2d_head(x, w, bias, n_loc, sigmoid, KH=3, KW=3, stride=1, pad=1, loc=None, conf=None):
    """Fused multibox head: x NHWC bf16 -> (loc fp32 NCHW [N,n_loc,H,W], conf fp32 NCHW [N,Cout-n_loc,H,W]).
    w rows [0,n_loc) are the loc conv, [n_loc,Cout) the conf conv (ssd.py:100-103)."""
    N, H, W, Cin = x.shape
    Cout = w.shape[0]
    ho = (H + 2 * pad - KH) // stride + 1
    wo = (W + 2 * pad - KW) // stride + 1
    if loc is None:
        loc = torch.empty((N, n_loc, ho, wo), dtype=torch.float32, device=x.device)
    if conf is None:
        conf = torch.empty((N, Cout - n_loc, ho, wo), dtype=torch.float32, device=x.device)
    d = _lib.ConvDesc(N=N, H=H, W=W, Cin=Cin, Cout=Cout, KH=KH, KW=KW, stride=stride, pad=pad,
                      Ho=ho, Wo=wo, x_cstride=x.stride(2), out_cstride=0, res_cstride=0,
                      x_row_pixels=0, chunk=0, x_kind=0,
                      w_rows=w.shape[0], relu=0, out_mode=1, n_loc=n_loc, sigmoid=int(sigmoid))
    with torch.cuda.device(x.device):
        check(lib.ssdsb_conv2d_bf16(C.byref(d), ptr(x), ptr(w), ptr(bias), None, ptr(loc), ptr(conf),
                                    stream_ptr()), "conv2d_head")
    return loc, conf


STEM_LEFT_PAD = 2      # zero pixels before each s2d row (the windowed stem reads [w-2, w+2))
STEM_ROW_EXTRA = 4     # row_px = W/2 + 4 (2 left + >= 1 right, kept even)


def pack_image_s2d(images, mean=0.0, std=1.0, out=None, padded=False):
    """images: fp32 NCHW [N,3,H,W] or uint8 NHWC [N,H,W,3] (CUDA) -> bf16 [N,H/2,W/2,16]
    (padded=True: [N,H/2,W/2+4,16] with the image at columns [2, 2+W/2) and zero padding)."""
    if images.dtype == torch.uint8:
        N, H, W, _ = images.shape
        fmt = 1
    else:
        images = images.float()
        N, _, H, W = images.shape
        fmt = 0
    images = images.contiguous()
    if out is None:
        if padded:
            out = torch.zeros((N, H // 2, W // 2 + STEM_ROW_EXTRA, 16), dtype=torch.bfloat16,
                              device=images.device)
        else:
            out = torch.empty((N, H // 2, W // 2, 16), dtype=torch.bfloat16, device=images.device)
    row_px = out.shape[2]
    left = STEM_LEFT_PAD if row_px != W // 2 else 0
    with torch.cuda.device(images.device):
        check(lib.ssdsb_pack_image_s2d(ptr(images), fmt, N, H, W, float(mean), float(std), row_px, left,
                                       ptr(out), stream_ptr()), "pack_image_s2d")
    return out


def maxpool3x3s2(x, out=None):
    """NHWC bf16 max pooling 3x3 / stride 2 / pad 1 (resnet.py:45)."""
    N, H, W, Cc = x.shape
    ho, wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    if out is None:
        out = torch.empty((N, ho, wo, Cc), dtype=torch.bfloat16, device=x.device)
    with torch.cuda.device(x.device):
        check(lib.ssdsb_maxpool3x3s2_nhwc_bf16(ptr(x), N, H, W, Cc, ptr(out), stream_ptr()), "maxpool")
    return out


def upsample2x_add(coarse, fine):
    """fine += nearest-2x-upsampled coarse (FPN top-down merge, fpn.py:80-87); NHWC bf16, in place."""
    N, H, W, Cc = fine.shape
    assert coarse.shape == (N, H // 2, W // 2, Cc), (coarse.shape, fine.shape)
    with torch.cuda.device(fine.device):
        check(lib.ssdsb_upsample2x_add_nhwc_bf16(ptr(coarse), ptr(fine), N, H, W, Cc, stream_ptr()),
              "upsample2x_add")
    return fine


def upsample2x_concat(fine, coarse, out=None):
    """cat(fine, nearest-2x-upsampled coarse) along channels (YOLOv3 top-down merge, yolo.py:70-72); NHWC bf16."""
    N, H, W, Cf = fine.shape
    Cc = coarse.shape[3]
    assert coarse.shape == (N, H // 2, W // 2, Cc) and fine.is_contiguous() and coarse.is_contiguous()
    if out is None:
        out = torch.empty((N, H, W, Cf + Cc), dtype=torch.bfloat16, device=fine.device)
    with torch.cuda.device(fine.device):
        check(lib.ssdsb_upsample2x_concat_nhwc_bf16(ptr(fine), ptr(coarse), N, H, W, Cf, Cc, ptr(out), stream_ptr()),
              "upsample2x_concat")
    return out


def pack_dw_weight(w_folded, c_pad=None):
    """depthwise [C,1,3,3] fp32 -> bf16 [9, C_pad] (tap-major)."""
    c = w_folded.shape[0]
    c_pad = c_pad or c
    out = torch.zeros((9, c_pad), dtype=torch.float32)
    out[:, :c] = w_folded.reshape(c, 9).t()
    return out.to(torch.bfloat16).contiguous()


def maxpool5x5s1(x, out):
    """5x5 / stride 1 / pad 2 max-pool, NHWC bf16; x and out may be channel slices of wider buffers (SPP concat)."""
    N, H, W, Cc = x.shape
    assert out.shape == x.shape and x.stride(3) == 1 and out.stride(3) == 1
    with torch.cuda.device(x.device):
        check(lib.ssdsb_maxpool5x5s1_nhwc_bf16(ptr(x), N, H, W, Cc, x.stride(2), ptr(out), out.stride(2), stream_ptr()),
              "maxpool5x5s1")
    return out


def dwconv3x3(x, w, bias, stride=1, relu=2, out=None):
    """NHWC bf16 depthwise 3x3/pad 1 + bias + activation (0 none, 1 ReLU, 2 ReLU6)."""
    N, H, W, Cc = x.shape
    ho, wo = (H - 1) // stride + 1, (W - 1) // stride + 1
    if out is None:
        out = torch.empty((N, ho, wo, Cc), dtype=torch.bfloat16, device=x.device)
    with torch.cuda.device(x.device):
        check(lib.ssdsb_dwconv3x3_nhwc_bf16(ptr(x), ptr(w), ptr(bias), N, H, W, Cc, stride, int(relu),
                                            ptr(out), stream_ptr()), "dwconv3x3")
    return out


def mbconv(x, w_exp, b_exp, w_dw, b_dw, w_proj, b_proj, stride=1, residual=False, relu=(2, 2, 0), out=None):
    """Fused MobileNetV2 inverted residual in one launch (ssdsb_mbconv_bf16):
    [1x1 expand + act] -> depthwise 3x3 (stride) + act -> 1x1 project [+ x].  x NHWC bf16 [N,H,W,Cin];
    w_exp bf16 [hid, 1, Cin] (pack_weight) or None (no expand layer: hid == Cin), w_dw bf16 [9, hid]
    (pack_dw_weight), w_proj bf16 [Cout, 1, hid].  Bit-identical to conv2d -> dwconv3x3 -> conv2d.
    Raises NotImplementedError where the kernel has no configuration (Cout > 256): use the separate launches."""
    N, H, W, Cin = x.shape
    hid = w_dw.shape[1]
    Cout = w_proj.shape[0]
    ho, wo = (H - 1) // stride + 1, (W - 1) // stride + 1
    if out is None:
        out = torch.empty((N, ho, wo, Cout), dtype=torch.bfloat16, device=x.device)
    d = _lib.MbconvDesc(N=N, H=H, W=W, Cin=Cin, hid=hid, Cout=Cout, stride=stride, residual=int(bool(residual)),
                        relu_expand=int(relu[0]), relu_dw=int(relu[1]), relu_project=int(relu[2]),
                        w_exp_rows=w_exp.shape[0] if w_exp is not None else 0, w_proj_rows=w_proj.shape[0],
                        out_cstride=out.stride(2))
    with torch.cuda.device(x.device):
        rc = lib.ssdsb_mbconv_bf16(C.byref(d), ptr(x), ptr(w_exp), ptr(b_exp), ptr(w_dw), ptr(b_dw), ptr(w_proj),
                                   ptr(b_proj), ptr(out), stream_ptr())
    check(rc, "mbconv")          # SSDSB_ERR_UNSUPPORTED -> NotImplementedError
    return out


def mbconv_last_launch():
    """{hc, tile_w, tile_h, pm, chunks, x_buffers, staging, dw_segments, dw_rows, grid, smem, block_n}"""
    out = (C.c_int * 12)()
    check(lib.ssdsb_mbconv_last_launch(out), "mbconv_last_launch")
    return dict(zip(("hc", "tile_w", "tile_h", "pm", "chunks", "x_buffers", "staging", "dw_segments", "dw_rows",
                     "grid", "smem", "block_n"), list(out)))


MB_BARRIERS = ("WORK", "we_empty", "hp_empty", "wp_empty", "a_full", "wp_full", "we_full", "eacc_empty", "slot_full",
               "pacc_full", "eacc_full", "hp_full", "a_empty", "x_empty", "pacc_empty", "x_full")
MB_WARPS = ("tma", "mma", "alloc", "store", "cvt0", "cvt1", "cvt2", "cvt3") + tuple(f"dw{i}" for i in range(8))


def mbconv_profile():
    """wait profile of CTA 0 of the last mbconv launch (needs SSDSB_MB_PROF=1 in the environment at launch):
    {"cycles", "chunks", "waits": {warp: {barrier: cycles}}}"""
    out = (C.c_ulonglong * 322)()
    check(lib.ssdsb_mbconv_profile(out), "mbconv_profile")
    waits = {}
    for w, wn in enumerate(MB_WARPS):
        row = {MB_BARRIERS[i]: int(out[w * 16 + i]) for i in range(0, 16) if out[w * 16 + i]}
        if row:
            waits[wn] = row
    events = [[int(out[258 + g * 8 + k]) for k in range(8)] for g in range(8)]
    return {"cycles": int(out[256]), "chunks": int(out[257]), "waits": waits, "events": events}


def pack_grouped_weight(w_folded, chunk, c_pad):
    """Grouped conv [C, gw, KH, KW] (groups = C / gw) -> block-diagonal chunk slabs for the chunked
    igemm: bf16 [(c_pad/chunk)*128, KH*KW, chunk]; slab s covers channels [s*chunk, (s+1)*chunk), its
    row r < chunk is output channel s*chunk + r, which only sees the gw inputs of its own group."""
    c, gw, kh, kw = w_folded.shape
    assert chunk % gw == 0 and c_pad % chunk == 0 and c_pad >= c
    n_chunks = c_pad // chunk
    out = torch.zeros((n_chunks, 128, kh * kw, chunk), dtype=torch.float32)
    wk = w_folded.permute(0, 2, 3, 1).reshape(c, kh * kw, gw)      # [C, taps, gw]
    for o in range(c):
        s_, r = divmod(o, chunk)
        g_local = (o // gw) - s_ * (chunk // gw)                    # group index inside the chunk
        out[s_, r, :, g_local * gw:(g_local + 1) * gw] = wk[o]
    return out.reshape(n_chunks * 128, kh * kw, chunk).to(torch.bfloat16).contiguous()


def bifpn_fuse(a, b, w0, w1, c=None, w2=0.0, mode=0, out=None):
    """BiFPN weighted fusion (bifpn.py:41-62): mode 0 out = w0*a + w1*up2(b); mode 1 out = w0*a +
    w1*maxpool2(b) [+ w2*c].  NHWC bf16."""
    N, H, W, Cc = a.shape
    if out is None:
        out = torch.empty_like(a)
    with torch.cuda.device(a.device):
        check(lib.ssdsb_bifpn_fuse_nhwc_bf16(ptr(a), ptr(b), ptr(c), int(mode), float(w0), float(w1),
                                             float(w2), N, H, W, Cc, ptr(out), stream_ptr()), "bifpn_fuse")
    return out
