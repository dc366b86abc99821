"""Oracle: the KITTI wavelet depth decoders, restated functionally on torch-CPU.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Follows
KITTI/networks/decoders/depth_decoder.py:72-168 (dense) and :171-428 (sparse).
Parameters are taken as a plain ``state_dict`` with the reference's key names
(``decoder.<k>...`` from the ModuleList at depth_decoder.py:122, order
upconv(4,0), upconv(4,1), waveconv(4,0), (4,1), (4,-1), upconv(3,0), ...), so a
reference checkpoint, a reference module's ``state_dict()`` or
``wavelet_monodepth_b200``'s modules can all feed it.
"""
import torch
import torch.nn.functional as F

from . import haar
from . import sparse_ops as sp

NUM_CH_DEC = (16, 32, 64, 128, 256)          # depth_decoder.py:82
_LEVEL_BASE = {4: 0, 3: 5, 2: 9, 1: 13}


def slot(i, name):
    """Index into the reference ModuleList for level i; name in upconv0|upconv1|ll|pos|neg."""
    base = _LEVEL_BASE[i]
    if name == "upconv0":
        return base
    if name == "upconv1":
        return base + 1
    if name == "ll":
        assert i == 4
        return base + 2
    off = 3 if i == 4 else 2
    return base + off + (0 if name == "pos" else 1)


def _block(params, k):
    return params["decoder.%d.conv.conv.weight" % k], params["decoder.%d.conv.conv.bias" % k]


def _head(params, k):
    return (params["decoder.%d.0.conv.weight" % k], params["decoder.%d.0.conv.bias" % k],
            params["decoder.%d.2.conv.weight" % k], params["decoder.%d.2.conv.bias" % k])


def _conv3_reflect(x, w, b):
    # layers.py:146-161 (Conv3x3, use_refl=True): ReflectionPad2d(1) + 3x3 conv
    return F.conv2d(F.pad(x, (1, 1, 1, 1), mode="reflect"), w, b)


def _conv_block(x, w, b):
    # layers.py:120-143 ConvBlock: Conv3x3 + ELU (norm = Identity)
    return F.elu(_conv3_reflect(x, w, b))


def _dense_head(x, w1, b1, w2, b2):
    # depth_decoder.py:104-120: Conv1x1 -> LeakyReLU(0.1) -> Conv3x3(reflect)
    return _conv3_reflect(F.leaky_relu(F.conv2d(x, w1, b1), 0.1), w2, b2)


def _up2(x):
    return F.interpolate(x, scale_factor=2, mode="nearest")      # layers.py:233-236


def _idwt(yl, yh):
    return haar.DWTInverse("haar", "zero")((yl, [yh]))


def dense_forward(params, feats):
    """DepthWaveProgressiveDecoder.forward (depth_decoder.py:138-168)."""
    out = {}
    x = feats[-1]
    yl = None
    for i in range(4, 0, -1):
        x = _conv_block(x, *_block(params, slot(i, "upconv0")))
        x = torch.cat([_up2(x), feats[i - 1]], 1)
        x = _conv_block(x, *_block(params, slot(i, "upconv1")))
        if i == 4:
            yl = (2 ** i) * torch.sigmoid(_dense_head(x, *_head(params, slot(i, "ll"))))
        pos = torch.sigmoid(_dense_head(x, *_head(params, slot(i, "pos"))))
        neg = torch.sigmoid(_dense_head(x, *_head(params, slot(i, "neg"))))
        yh = (2 ** (i - 1)) * pos.unsqueeze(1) - (2 ** (i - 1)) * neg.unsqueeze(1)   # :133-135
        out[("wavelets", i - 1, "LL")] = yl
        out[("wavelets", i - 1, "LH")] = yh[:, :, 0]
        out[("wavelets", i - 1, "HL")] = yh[:, :, 1]
        out[("wavelets", i - 1, "HH")] = yh[:, :, 2]
        yl = _idwt(yl, yh)
        out[("disp", i - 1)] = torch.clamp(yl / 2 ** (i - 1), 0, 1)
    return out


def level_masks(yl, yh, thresh_ratio, like=None):
    """The six per-level pixel sets (depth_decoder.py:305-319).

    Returns dict S0 (float 0/1, low-res), S1 lowres, S2 upconv0 (low-res bool),
    S3 upsample, S4 upconv1, S5 wavelet (hi-res bool).  ``yl is None`` = first
    level (all ones, shaped like ``like``).
    """
    if yl is None:
        s0 = torch.ones_like(like[:, 0:1])
    else:
        thresh = (yl.max() - yl.min()) * thresh_ratio
        s0 = (torch.abs(yh).max(2)[0] > thresh).float()
    u = _up2(s0)
    return {
        "S0": s0,
        "S1": F.max_pool2d(s0, 3, 1, 1).bool(),
        "S2": F.max_pool2d(s0, 5, 1, 2).bool(),
        "S3": F.max_pool2d(u, 5, 1, 2).bool(),
        "S4": F.max_pool2d(u, 3, 1, 1).bool(),
        "S5": u.bool(),
    }


def _dense_conv_ops(x, w):
    # depth_decoder.py:386-387: bias counted once per output channel
    return (1 + 9 * x.shape[1] * x.shape[2] * x.shape[3]) * w.shape[0]


def _dense_head_ops(x, w1, w2):
    # depth_decoder.py:247-266
    hw = x.shape[2] * x.shape[3]
    return (1 + w1.shape[1] * hw) * w1.shape[0] + (1 + 9 * w2.shape[1] * hw) * w2.shape[0]


def sparse_forward(params, feats, thresh_ratio=0.05, sparse_scales=(0, 1, 2, 3)):
    """SparseDepthWaveProgressiveDecoder.forward (depth_decoder.py:292-428), batch 1."""
    out = {}
    x = feats[-1]
    assert x.shape[0] == 1, "works with single input only"          # :297
    total_ops = 0
    yl = yh = None
    xvals = xchn = prev_idxmap = None
    for i in range(4, 0, -1):
        s = level_masks(yl if i != 4 else None, yh, thresh_ratio, like=x)
        lo_h, lo_w = s["S0"].shape[2:]
        ops = (3 * lo_h * lo_w if i != 4 else 0) + 25 * lo_h * lo_w + 100 * lo_h * lo_w   # :310,322-323
        for name, key in (("lowres_mask", "S1"), ("upconv0_mask", "S2"), ("upsample_mask", "S3"),
                          ("upconv1_mask", "S4"), ("wavelet_mask", "S5")):
            out[(name, i - 1)] = s[key].clone()

        if i in sparse_scales:
            assert yl is not None
            map1, o1 = sp.index_map(s["S1"])
            map2, o2 = sp.index_map(s["S2"])
            map3, o3 = sp.index_map(s["S3"])
            map4, o4 = sp.index_map(s["S4"])
            ops += o1 + o2 + o3 + o4
            if i == max(sparse_scales):
                xchn = x.shape[1]
                xvals = x[s["S1"].expand(-1, xchn, -1, -1)]                       # :346-348
            else:
                xvals = sp.select(xvals, xchn, prev_idxmap, s["S1"], pad=True)     # :350
            w, b = _block(params, slot(i, "upconv0"))
            xvals, xchn, o = sp.conv3x3(w, b, xvals, map1, s["S2"], nonlin=F.elu, make_result=False)
            ops += o
            xvals, xchn = sp.upsample_concat(xvals, xchn, map2, feats[i - 1], s["S3"], make_result=False)
            w, b = _block(params, slot(i, "upconv1"))
            xvals, xchn, o = sp.conv3x3(w, b, xvals, map3, s["S4"], nonlin=F.elu, make_result=False)
            ops += o
            pos, o = sp.head3x3(*_head(params, slot(i, "pos")), xvals, map4, s["S5"], torch.sigmoid)
            ops += o
            neg, o = sp.head3x3(*_head(params, slot(i, "neg")), xvals, map4, s["S5"], torch.sigmoid)
            ops += o
            yh = ((2 ** (i - 1)) * (pos - neg)).unsqueeze(1)                       # :288
            prev_idxmap = map4
        else:
            w, b = _block(params, slot(i, "upconv0"))
            ops += _dense_conv_ops(x, w)
            x = _conv_block(x, w, b)
            ux = torch.cat([_up2(x), feats[i - 1]], 1)
            w, b = _block(params, slot(i, "upconv1"))
            ops += _dense_conv_ops(ux, w)
            ux = _conv_block(ux, w, b)
            if i == 4:
                hd = _head(params, slot(i, "ll"))
                ops += _dense_head_ops(ux, hd[0], hd[2])
                yl = (2 ** i) * torch.sigmoid(_dense_head(ux, *hd))
            hp = _head(params, slot(i, "pos"))
            hn = _head(params, slot(i, "neg"))
            ops += _dense_head_ops(ux, hn[0], hn[2]) + _dense_head_ops(ux, hp[0], hp[2])
            pos = torch.sigmoid(_dense_head(ux, *hp))
            neg = torch.sigmoid(_dense_head(ux, *hn))
            yh = ((2 ** (i - 1)) * (pos - neg) * s["S5"]).unsqueeze(1)             # :271-272
            x = ux

        out[("wavelets", i - 1, "LL")] = yl
        out[("wavelets", i - 1, "LH")] = yh[:, :, 0]
        out[("wavelets", i - 1, "HL")] = yh[:, :, 1]
        out[("wavelets", i - 1, "HH")] = yh[:, :, 2]
        yl = _idwt(yl, yh)
        ops += 4 * yl.shape[2] * yl.shape[3]                                        # :373,417
        out[("disp", i - 1)] = torch.clamp(yl / 2 ** (i - 1), 0, 1)
        total_ops += ops
        out[("total_ops", i - 1)] = ops
    out["total_ops"] = total_ops
    return out
