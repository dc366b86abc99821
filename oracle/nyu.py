"""Oracle: the NYUv2 DenseDepth-style wavelet decoders, restated functionally on torch-CPU.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Follows
NYUv2/networks/decoders/densedepth_decoder.py:92-148 (DecoderWave) and :224-409
(SparseDecoderWave); conv blocks from NYUv2/networks/layers.py:11-32,57-67.
Parameters are a plain ``state_dict`` with the reference's key names
(``conv2.conv.weight``, ``up1.convA.conv.weight``, ``wave1_ll.conv.weight`` ...).
Depthwise variants (dw_waveconv / dw_upconv) are not part of the hot path and
are not restated.
"""
import torch
import torch.nn.functional as F

from . import haar
from . import sparse_ops as sp

_PAD = {"reflection": "reflect", "replicate": "replicate"}


def _conv3(x, w, b, padding):
    # NYUv2/networks/layers.py:11-32 Conv3x3: explicit pad layer then unpadded conv
    if padding in _PAD:
        x = F.pad(x, (1, 1, 1, 1), mode=_PAD[padding])
    else:
        x = F.pad(x, (1, 1, 1, 1))
    return F.conv2d(x, w, b)


def _p(params, name):
    return params[name + ".conv.weight"], params[name + ".conv.bias"]


def _up_block(params, name, x, skip):
    # layers.py:57-67 UpSampleBlock: nearest x2, concat, convA (reflection), LeakyReLU(0.2)
    x = torch.cat([F.interpolate(x, scale_factor=2, mode="nearest"), skip], 1)
    return F.leaky_relu(_conv3(x, *_p(params, name + ".convA"), "reflection"), 0.2)


def _idwt(ll, h):
    return haar.DWTInverse("haar", "zero")((ll, [h]))


def dense_forward(params, blocks):
    """DecoderWave.forward (densedepth_decoder.py:117-148)."""
    out = {}
    d0 = _conv3(blocks[-1], *_p(params, "conv2"), "replicate")
    d1 = _up_block(params, "up1", d0, blocks[-2])
    ll = (2 ** 3) * _conv3(d1, *_p(params, "wave1_ll"), "replicate")
    out[("disp", 3)] = ll / (2 ** 3)
    h = (2 ** 2) * _conv3(d1, *_p(params, "wave1"), "zero").unsqueeze(1)
    out[("wavelets", 2, "LL")] = ll
    for k, band in enumerate(("LH", "HL", "HH")):
        out[("wavelets", 2, band)] = h[:, :, k]
    ll = _idwt(ll, h)
    out[("disp", 2)] = ll / (2 ** 2)

    d2 = _up_block(params, "up2", d1, blocks[-3])
    h = (2 ** 1) * _conv3(d2, *_p(params, "wave2"), "zero").unsqueeze(1)
    for k, band in enumerate(("LH", "HL", "HH")):
        out[("wavelets", 1, band)] = h[:, :, k]
    ll = _idwt(ll, h)
    out[("disp", 1)] = ll / (2 ** 1)

    d3 = _up_block(params, "up3", d2, blocks[-4])
    h = _conv3(d3, *_p(params, "wave3"), "zero").unsqueeze(1)
    for k, band in enumerate(("LH", "HL", "HH")):
        out[("wavelets", 0, band)] = h[:, :, k]
    ll = _idwt(ll, h)
    out[("disp", 0)] = ll
    return out


def level_masks(ll, h, thresh_ratio):
    """densedepth_decoder.py:316-322 / :363-368.  S0 mask, S2 up_mask (low-res), S3 conva, S4 wave, S5 wavelet (float)."""
    thresh = (ll.max() - ll.min()) * thresh_ratio
    s0 = (torch.abs(h).max(2)[0] > thresh).float()
    u = F.interpolate(s0, scale_factor=2, mode="nearest")
    return {
        "S0": s0,
        "S2": F.max_pool2d(s0, 5, 1, 2).bool(),
        "S3": F.max_pool2d(u, 5, 1, 2).bool(),
        "S4": F.max_pool2d(u, 3, 1, 1).bool(),
        "S5": u,
    }


def _sparse_block(params, up_name, wave_name, scale, ll, h, skip, thresh_ratio, carried, double_count):
    """One sparse level (densedepth_decoder.py:314-359 / :361-406).

    carried = ("dense", x_d1) for the first block (features taken from the dense
    map at up_mask, :339) or ("sparse", xvals, xchn, prev_idxmap) for the second
    (sparse_select with pad, :386).
    """
    s = level_masks(ll, h, thresh_ratio)
    lh, lw = s["S0"].shape[2:]
    ops = 3 * lh * lw + 25 * lh * lw + 100 * lh * lw                # :318,324-325
    _, o5 = sp.index_map(s["S5"])
    map3, o3 = sp.index_map(s["S3"])
    map4, o4 = sp.index_map(s["S4"])
    map2, o2 = sp.index_map(s["S2"])
    ops += o5 + o3 + o4 + o2
    if double_count:
        ops += o4                                                     # :381-382 wave_mask indexed twice
    if carried[0] == "dense":
        xd = carried[1]
        xchn = xd.shape[1]
        xvals = xd[s["S2"].expand(-1, xchn, -1, -1)]
    else:
        _, pv, pc, pmap = carried
        xchn = pc
        xvals = sp.select(pv, pc, pmap, s["S2"], pad=True)
    xvals, xchn = sp.upsample_concat(xvals, xchn, map2, skip, s["S3"], make_result=False)
    w, b = _p(params, up_name + ".convA")
    xvals, xchn, o = sp.conv3x3(w, b, xvals, map3, s["S4"], nonlin=lambda t: F.leaky_relu(t, 0.2),
                                padding="reflect", make_result=False)
    ops += o
    w, b = _p(params, wave_name)
    hd, o = sp.conv3x3(w, b, xvals, map4, s["S5"], nonlin=None, padding="constant", make_result=True)
    ops += o
    h_new = scale * hd.unsqueeze(1)
    ll_new = _idwt(ll, s["S5"].unsqueeze(2) * h_new)                  # :357,404
    ops += ll_new.shape[2] * ll_new.shape[3]
    return s, h_new, ll_new, ops, (xvals, xchn, map4)


def sparse_forward(params, blocks, thresh_ratio=0.1):
    """SparseDecoderWave.forward (densedepth_decoder.py:271-409), batch 1."""
    out = {}
    total = 0
    xb = blocks[-1]
    w, b = _p(params, "conv2")
    total += (1 + 9 * xb.shape[1]) * xb.shape[2] * xb.shape[3] * w.shape[0]     # :276-278
    d0 = _conv3(xb, w, b, "replicate")
    d1 = _up_block(params, "up1", d0, blocks[-2])
    chn = d0.shape[1] + blocks[-2].shape[1]
    total += (1 + 9 * chn) * d1.shape[2] * d1.shape[3] * d1.shape[1]            # :284-286
    ll = (2 ** 3) * _conv3(d1, *_p(params, "wave1_ll"), "replicate")
    out[("disp", 3)] = ll / (2 ** 3)
    h = ((2 ** 2) * _conv3(d1, *_p(params, "wave1"), "zero")).unsqueeze(1)
    total += (1 + 9 * d1.shape[1]) * d1.shape[2] * d1.shape[3] * 4              # :296-298
    out[("wavelet_mask", 2)] = torch.ones_like(h[:, 0])
    out[("wavelets", 2, "LL")] = ll
    for k, band in enumerate(("LH", "HL", "HH")):
        out[("wavelets", 2, band)] = h[:, :, k]
    ll = _idwt(ll, h)
    total += ll.shape[2] * ll.shape[3]
    out[("disp", 2)] = ll / (2 ** 2)

    s, h, ll, ops, carry = _sparse_block(params, "up2", "wave2", 2 ** 1, ll, h, blocks[-3], thresh_ratio,
                                         ("dense", d1), False)
    total += ops
    out[("wavelet_mask", 1)] = s["S5"]
    for k, band in enumerate(("LH", "HL", "HH")):
        out[("wavelets", 1, band)] = h[:, :, k]
    out[("disp", 1)] = ll / (2 ** 1)

    s, h, ll, ops, _ = _sparse_block(params, "up3", "wave3", 1, ll, h, blocks[-4], thresh_ratio,
                                     ("sparse",) + carry, True)
    total += ops
    out[("wavelet_mask", 0)] = s["S5"]
    for k, band in enumerate(("LH", "HL", "HH")):
        out[("wavelets", 0, band)] = h[:, :, k]
    out[("disp", 0)] = ll / (2 ** 0)
    out["total_ops"] = total
    return out
