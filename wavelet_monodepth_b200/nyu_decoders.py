"""NYUv2 DenseDepth-style wavelet decoders with the reference's contract, on libwmd.

Mirrors NYUv2/networks/layers.py:11-79 (``Conv3x3``, ``upsample``, ``UpSampleBlock``, ``depthwise``,
``pointwise``) and NYUv2/networks/decoders/densedepth_decoder.py:92-148 (``DecoderWave``), :224-409
(``SparseDecoderWave``).  State-dict names (``conv2.conv.weight``, ``up1.convA.conv.weight``,
``wave1_ll.conv.weight`` ..., ``iwt.*`` / ``iwt_LL.*`` buffers), constructor and forward signatures and the
output-dict keys are the reference's.  The functional ``sparse_*`` ops of NYUv2/networks/layers.py:82-223
are the KITTI ones minus the 1x1 branch; they are re-exported from ``kitti_layers`` (whose
``sparse_conv3x3`` accepts this file's ``Conv3x3`` as well).

As for KITTI: inference runs natively and batched in the pixel-major row layout; grad-enabled calls of
``DecoderWave`` take the differentiable cuDNN + native-IDWT path (NYUv2/train.py:293-327 trains through it).
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import opcount, ops
from .opsfuture import OpsFuture
from ._lib import ACT_LRELU, ACT_NONE, PAD_REFLECT, PAD_REPLICATE, PAD_ZERO, WmdError
from .kitti_decoders import _PackCache, _need_cuda, _pm
from .kitti_layers import (make_result, mask2idxmap, mask2yx, sparse_conv3x3, sparse_select,  # noqa: F401
                           sparse_upsample)
from .wavelets import IDWT

_PAD_CODE = {"reflection": PAD_REFLECT, "replicate": PAD_REPLICATE, "zero": PAD_ZERO}


def depthwise(in_channels, kernel_size):
    """[NYUv2/networks/layers.py:70-75]"""
    return nn.Sequential(
        nn.Conv2d(in_channels, in_channels, kernel_size, stride=1, padding=0, bias=False, groups=in_channels),
        nn.ReLU(inplace=True),
    )


def pointwise(in_channels, out_channels):
    """[NYUv2/networks/layers.py:78-79]"""
    return nn.Conv2d(in_channels, out_channels, 1, 1, 0, bias=False)


class Conv3x3(nn.Module):
    """Pad (reflection / replicate / zero) and convolve.  [NYUv2/networks/layers.py:11-32]"""

    def __init__(self, in_channels, out_channels, padding="zero", stride=1, is_depthwise=False):
        super().__init__()
        if padding == "reflection":
            self.pad = nn.ReflectionPad2d(1)
        elif padding == "replicate":
            self.pad = nn.ReplicationPad2d(1)
        else:
            self.pad = nn.ZeroPad2d(1)
        self.padding = padding if padding in ("reflection", "replicate") else "zero"
        self.is_depthwise = bool(is_depthwise)
        if is_depthwise:
            self.conv = nn.Sequential(depthwise(int(in_channels), kernel_size=3),
                                      pointwise(int(in_channels), int(out_channels)))
        else:
            self.conv = nn.Conv2d(int(in_channels), int(out_channels), 3, stride=stride, padding=0)

    def forward(self, x):
        return self.conv(self.pad(x))


def upsample(x):
    """[NYUv2/networks/layers.py:35-36]"""
    return F.interpolate(x, scale_factor=2, mode="nearest")


class UpSampleBlock(nn.Sequential):
    """nearest x2, concat skip, convA, LeakyReLU(0.2).  [NYUv2/networks/layers.py:57-67]"""

    def __init__(self, skip_input, output_features, padding="zero", is_depthwise=False):
        super().__init__()
        self.convA = Conv3x3(skip_input, output_features, padding=padding, is_depthwise=is_depthwise)
        self.leakyreluA = nn.LeakyReLU(0.2)
        self.upsample = nn.Upsample(scale_factor=2, mode="nearest")

    def forward(self, x, concat_with):
        return self.leakyreluA(self.convA(torch.cat([self.upsample(x), concat_with], dim=1)))


class _NyuWaveBase(nn.Module):
    def _build(self, enc_features, decoder_width, dw_waveconv=False, dw_upconv=False):
        features = int(enc_features[-1] * decoder_width)
        self.features = features
        self.enc_features = list(enc_features)
        wave_pad = "zero"
        padding = "reflection"
        self.iwt = IDWT(wave="haar", mode=wave_pad)
        self.iwt_LL = IDWT(wave="haar", mode="zero")
        self.conv2 = Conv3x3(enc_features[-1], features, padding="replicate")
        self.up1 = UpSampleBlock(skip_input=features // 1 + enc_features[-2], output_features=features // 2,
                                 padding=padding, is_depthwise=dw_upconv)
        self.wave1_ll = Conv3x3(features // 2, 1, padding="replicate")
        self.wave1 = Conv3x3(features // 2, 3, padding=wave_pad, is_depthwise=dw_waveconv)
        self.up2 = UpSampleBlock(skip_input=features // 2 + enc_features[-3], output_features=features // 4,
                                 padding=padding, is_depthwise=dw_upconv)
        self.wave2 = Conv3x3(features // 4, 3, padding=wave_pad, is_depthwise=dw_waveconv)
        self.up3 = UpSampleBlock(skip_input=features // 4 + enc_features[-4], output_features=features // 8,
                                 padding=padding, is_depthwise=dw_upconv)
        self.wave3 = Conv3x3(features // 8, 3, padding=wave_pad, is_depthwise=dw_waveconv)
        self._depthwise = bool(dw_waveconv or dw_upconv)
        # optional consumer epilogue of ("disp", 0), off by default: (div, lo, hi) adds ("depth", 0) =
        # clamp(("disp", 0) / div, lo, hi) - NYUv2/utils.py:219,229 uses (100, 0.4, 10) - fused into the last IDWT
        self.depth_epilogue = None
        self._packs = _PackCache()
        self.register_load_state_dict_post_hook(lambda module, incompatible: module.invalidate_packs())

    def invalidate_packs(self):
        """Drop the packed weight copies (see kitti_decoders._PackCache)."""
        self._packs.invalidate()

    def _apply(self, fn, *args, **kwargs):
        if hasattr(self, "_packs"):
            self._packs.invalidate()
        return super()._apply(fn, *args, **kwargs)

    def _gemm(self, name, layer, c1=0):
        conv = layer.conv
        return self._packs.get(("gemm", name, ops.default_conv_kind()), [conv.weight],
                               lambda: ops.pack_weight(conv.weight, c1)), conv.bias.detach()

    def _head(self, name, layer):
        conv = layer.conv
        return self._packs.get(("head", name), [conv.weight], lambda: ops.pack_head_weight(conv.weight)), conv.bias.detach()

    @torch.no_grad()
    def _native_forward(self, blocks, thresh_ratio, sparse):
        """conv2/up1/wave1 dense, then the up2/wave2 and up3/wave3 levels dense or on active lists.

        Returns (outputs, counts): counts = int32 device tensor (2, 2, N+1), row offsets of S4 / S5 of the two sparse
        blocks (None on the dense path)."""
        _need_cuda(blocks)
        if self._depthwise:
            raise NotImplementedError("depthwise-separable variants only run on the differentiable cuDNN path")
        dev = blocks[-1].device
        if dev.index is not None and dev.index != torch.cuda.current_device():
            with torch.cuda.device(dev):       # libwmd launches on the current device
                return self._native_forward_on_device(blocks, thresh_ratio, sparse)
        return self._native_forward_on_device(blocks, thresh_ratio, sparse)

    def _native_forward_on_device(self, blocks, thresh_ratio, sparse):
        out = {}
        xb = blocks[-1]
        n, _, h, w = xb.shape
        f = self.features
        counts = []
        wp, b = self._gemm("conv2", self.conv2)
        d0 = ops.conv_rows(ops.nchw_to_rows(xb), xb.shape[1], wp, b, f, n, h, w, pad=PAD_REPLICATE, act=ACT_NONE)
        skip = blocks[-2]
        if tuple(skip.shape[2:]) != (2 * h, 2 * w):
            raise WmdError("skip block has shape %s, expected spatial %s" % (tuple(skip.shape), (2 * h, 2 * w)))
        wp, b = self._gemm("up1", self.up1.convA, skip.shape[1])
        d1 = ops.conv_rows(d0, f, wp, b, f // 2, n, 2 * h, 2 * w, pad=PAD_REFLECT, act=ACT_LRELU, act_param=0.2,
                           shift0=1, x1=ops.nchw_to_rows(skip), c1=skip.shape[1])
        h, w = 2 * h, 2 * w
        wl, bl = self._head("wave1_ll", self.wave1_ll)
        raw = ops.head_conv3x3(d1, f // 2, 0, wl, bl, n, h, w, 1, scale=1.0, act=ACT_NONE, pad=PAD_REPLICATE)
        ll = raw * float(2 ** 3)             # exact power-of-two scaling of a (N,1,H/16,W/16) map
        out[("disp", 3)] = raw               # == ll / 2**3 (densedepth_decoder.py:123)
        wh, bh = self._head("wave1", self.wave1)
        hcoef = ops.head_conv3x3(d1, f // 2, 0, wh, bh, n, h, w, 3, scale=float(2 ** 2), act=ACT_NONE, pad=PAD_ZERO)
        if sparse:
            # the reference builds this one as ones_like(h[:, 0]) with h already (N,1,3,H,W): a 3-channel map
            # (densedepth_decoder.py:301-303); kept as is
            out[("wavelet_mask", 2)] = torch.ones((n, 3, h, w), dtype=torch.float32, device=xb.device)
        out[("wavelets", 2, "LL")] = ll
        for k, band in enumerate(("LH", "HL", "HH")):
            out[("wavelets", 2, band)] = hcoef[:, k:k + 1]
        ll, disp = ops.idwt_haar(ll, hcoef.unsqueeze(1), disp_scale=1.0 / 2 ** 2, clamp01=False)
        out[("disp", 2)] = disp

        x_rows, x_c, prev_map = d1, f // 2, None
        for s, (up, wave, scale) in enumerate(((self.up2, self.wave2, 2.0), (self.up3, self.wave3, 1.0))):
            name = "up%d" % (s + 2)
            skip = blocks[-3 - s]
            cs = skip.shape[1]
            cout = up.convA.conv.weight.shape[0]
            wp, b = self._gemm(name, up.convA, cs)
            wh, bh = self._head("wave%d" % (s + 2), wave)
            if tuple(skip.shape[2:]) != (2 * h, 2 * w):
                raise WmdError("skip block has shape %s, expected spatial %s" % (tuple(skip.shape), (2 * h, 2 * w)))
            if sparse:
                thresh = ops.range_thresh(ll, thresh_ratio)
                masks = ops.level_masks(hcoef, thresh, want=("S2", "S3", "S4", "S5"))
                gmap = ops.gate_map(masks["S2"], prev_map)
                map4, pix4, off4 = ops.compact(masks["S4"])
                _, pix5, off5 = ops.compact(masks["S5"], want_idxmap=False)
                counts.append((off4, off5))
                out[("wavelet_mask", 1 - s)] = masks["S5"].to(torch.float32)
                xa = ops.conv_rows(x_rows, x_c, wp, b, cout, n, 2 * h, 2 * w, pad=PAD_REFLECT, act=ACT_LRELU,
                                   act_param=0.2, map0=gmap, shift0=1, x1=ops.nchw_to_rows(skip), c1=cs,
                                   gate=masks["S3"], pixels=pix4, count=off4[n:],
                                   m_in0=_pm(lambda: (gmap >= 0).sum()), m_in1=_pm(lambda: masks["S3"].sum()))
                hcoef = ops.head_conv3x3(xa, cout, 0, wh, bh, n, 2 * h, 2 * w, 3, scale=scale, act=ACT_NONE,
                                         pad=PAD_ZERO, idxmap=map4, pixels=pix5, count=off5[n:])
                prev_map = map4
            else:
                xa = ops.conv_rows(x_rows, x_c, wp, b, cout, n, 2 * h, 2 * w, pad=PAD_REFLECT, act=ACT_LRELU,
                                   act_param=0.2, shift0=1, x1=ops.nchw_to_rows(skip), c1=cs)
                hcoef = ops.head_conv3x3(xa, cout, 0, wh, bh, n, 2 * h, 2 * w, 3, scale=scale, act=ACT_NONE,
                                         pad=PAD_ZERO)
            for k, band in enumerate(("LH", "HL", "HH")):
                out[("wavelets", 1 - s, band)] = hcoef[:, k:k + 1]
            if s == 0:
                ll, disp = ops.idwt_haar(ll, hcoef.unsqueeze(1), disp_scale=0.5, clamp01=False)
                out[("disp", 1)] = disp
            elif self.depth_epilogue is not None:
                ll, depth = ops.idwt_haar(ll, hcoef.unsqueeze(1), epilogue=("div_clamp",) + tuple(self.depth_epilogue))
                out[("disp", 0)], out[("depth", 0)] = ll, depth
            else:
                ll = ops.idwt_haar(ll, hcoef.unsqueeze(1))
                out[("disp", 0)] = ll
            x_rows, x_c = xa, cout
            h, w = 2 * h, 2 * w
        return out, (torch.stack([torch.stack(c) for c in counts]) if counts else None)


class DecoderWave(_NyuWaveBase):
    """Dense wavelet decoder.  [densedepth_decoder.py:92-148]"""

    def __init__(self, enc_features=[96, 96, 192, 384, 2208], decoder_width=0.5, dw_waveconv=False, dw_upconv=False):
        super().__init__()
        self._build(enc_features, decoder_width, dw_waveconv, dw_upconv)

    def _autograd_forward(self, x_blocks):
        outputs = {}
        x_d0 = self.conv2(x_blocks[-1])
        x_d1 = self.up1(x_d0, x_blocks[-2])
        ll = (2 ** 3) * self.wave1_ll(x_d1)
        outputs[("disp", 3)] = ll / (2 ** 3)
        h = (2 ** 2) * self.wave1(x_d1).unsqueeze(1)
        outputs[("wavelets", 2, "LL")] = ll
        for k, band in enumerate(("LH", "HL", "HH")):
            outputs[("wavelets", 2, band)] = h[:, :, k]
        ll = self.iwt((ll, list([h])))
        outputs[("disp", 2)] = ll / (2 ** 2)
        x_d2 = self.up2(x_d1, x_blocks[-3])
        h = (2 ** 1) * self.wave2(x_d2).unsqueeze(1)
        for k, band in enumerate(("LH", "HL", "HH")):
            outputs[("wavelets", 1, band)] = h[:, :, k]
        ll = self.iwt((ll, list([h])))
        outputs[("disp", 1)] = ll / (2 ** 1)
        x_d3 = self.up3(x_d2, x_blocks[-4])
        h = self.wave3(x_d3).unsqueeze(1)
        for k, band in enumerate(("LH", "HL", "HH")):
            outputs[("wavelets", 0, band)] = h[:, :, k]
        ll = self.iwt((ll, list([h])))
        outputs[("disp", 0)] = ll
        return outputs

    def forward(self, x_blocks):
        _need_cuda(x_blocks)
        needs_grad = torch.is_grad_enabled() and (
            any(p.requires_grad for p in self.parameters()) or any(f.requires_grad for f in x_blocks))
        if needs_grad or self._depthwise:
            return self._autograd_forward(x_blocks)
        out, _ = self._native_forward(x_blocks, 0.0, sparse=False)
        return out


class SparseDecoderWave(_NyuWaveBase):
    """Threshold-gated sparse decoder (levels 1 and 0 sparse), batched.  [densedepth_decoder.py:224-409]"""

    def __init__(self, enc_features=[96, 96, 192, 384, 2208], decoder_width=0.5):
        super().__init__()
        self._build(enc_features, decoder_width)
        self.sparse_padding = "reflect"
        self.sparse_wave_pad = "constant"
        self.leakyreluA = nn.LeakyReLU(0.2)
        self.maxpool3 = nn.MaxPool2d(3, stride=1, padding=1)
        self.maxpool5 = nn.MaxPool2d(5, stride=1, padding=2)
        self.maxpool7 = nn.MaxPool2d(7, stride=1, padding=3)
        self.count_ops = True

    def forward(self, x_blocks, thresh_ratio=0.1):
        out, counts = self._native_forward(x_blocks, float(thresh_ratio), sparse=True)
        if self.count_ops:
            fut = self.ops_future(counts, x_blocks)
            if self.count_ops == "async":
                out["total_ops"] = fut                       # OpsFuture: nothing waits (see opsfuture.py)
            else:
                out.update(fut.result())
        return out

    def ops_future(self, counts, x_blocks):
        """OpsFuture of one forward: enqueues the count read-back on the current stream (no host wait)."""
        n, cin, h, w = (int(v) for v in x_blocks[-1].shape)
        f = self.features
        c2, c3, c4 = (int(x_blocks[k].shape[1]) for k in (-2, -3, -4))

        def finish(host):                                    # host: (2, 2, N+1) offsets of S4 / S5 per sparse block
            per_sample = []
            for b in range(n):
                v = opcount.nyu_dense_part_ops(cin, h, w, f, c2)
                m4, m5 = (int(host[0][k][b + 1] - host[0][k][b]) for k in range(2))
                v += opcount.nyu_sparse_block_ops(2 * h, 2 * w, f // 2 + c3, f // 4, m4, m5, False)
                m4, m5 = (int(host[1][k][b + 1] - host[1][k][b]) for k in range(2))
                v += opcount.nyu_sparse_block_ops(4 * h, 4 * w, f // 4 + c4, f // 8, m4, m5, True)
                per_sample.append(v)
            res = {"total_ops": sum(per_sample)}
            if n > 1:
                res["total_ops_per_sample"] = per_sample
            return res

        return OpsFuture(counts, finish)


# ------------------------------------------------------------------------------------------------------------------
# API surface outside the hot path (SURVEY 8b lists them as constructible from NYUv2/model.py:47-64): the DenseDepth
# baseline decoders and the 224-pixel wavelet variant.  They run on the differentiable cuDNN path (+ the native IDWT
# through its autograd function); no native gather-GEMM engine is built for them.
# ------------------------------------------------------------------------------------------------------------------
class _BaselineDecoder(nn.Module):
    """conv2 -> four UpSampleBlocks -> [x2 + conv5 + LeakyReLU(0.2)] -> conv3; zero padding everywhere."""

    def _build(self, enc_features, decoder_width, is_depthwise, extra_stage):
        f = int(enc_features[-1] * decoder_width)
        self.conv2 = Conv3x3(enc_features[-1], f, padding="zero")
        for k in range(1, 5):
            setattr(self, "up%d" % k, UpSampleBlock(skip_input=f // 2 ** (k - 1) + enc_features[-1 - k],
                                                    output_features=f // 2 ** k, padding="zero",
                                                    is_depthwise=is_depthwise))
        last = f // 16
        if extra_stage:
            self.conv5 = nn.Sequential(Conv3x3(f // 16, f // 32, is_depthwise=is_depthwise), nn.LeakyReLU(0.2))
            last = f // 32
        if is_depthwise:
            self.conv3 = Conv3x3(last, 1, is_depthwise=True)
        else:
            self.conv3 = nn.Conv2d(last, 1, kernel_size=3, stride=1, padding=1, padding_mode="zeros")
        if extra_stage:
            self.upsample = nn.Upsample(scale_factor=2, mode="nearest")
        self._extra_stage = extra_stage

    def forward(self, features):
        blocks = tuple(features)
        if len(blocks) != 5:
            raise ValueError("expected the five encoder blocks, fine to coarse")
        x = self.conv2(blocks[4])
        for k in range(1, 5):
            x = getattr(self, "up%d" % k)(x, blocks[4 - k])
        if self._extra_stage:
            x = self.conv5(self.upsample(x))
        return {("disp", 0): self.conv3(x)}


class Decoder(_BaselineDecoder):
    """DenseDepth baseline decoder (no wavelets).  [densedepth_decoder.py:15-47]"""

    def __init__(self, enc_features=[96, 96, 192, 384, 2208], decoder_width=0.5, is_depthwise=False):
        super().__init__()
        self._build(enc_features, decoder_width, is_depthwise, extra_stage=False)


class Decoder224(_BaselineDecoder):
    """Baseline decoder for 224-pixel inputs: one more x2 + conv stage.  [densedepth_decoder.py:50-89]"""

    def __init__(self, enc_features=[96, 96, 192, 384, 2208], decoder_width=0.5, is_depthwise=False):
        super().__init__()
        self._build(enc_features, decoder_width, is_depthwise, extra_stage=True)


class DecoderWave224(nn.Module):
    """Four-level wavelet decoder for 224-pixel inputs.  [densedepth_decoder.py:151-221]

    Same state-dict names as the reference (conv2, up1..up4, wave1_ll, wave1..wave4, iwt / iwt_LL buffers).  Keeps the
    reference's quirk of FLOOR-dividing ``("disp", 1)`` (:212)."""

    def __init__(self, enc_features=[96, 96, 192, 384, 2208], decoder_width=0.5, dw_waveconv=False, dw_upconv=False):
        super().__init__()
        f = int(enc_features[-1] * decoder_width)
        self.iwt = IDWT(wave="haar", mode="zero")
        self.iwt_LL = IDWT(wave="haar", mode="zero")
        self.conv2 = Conv3x3(enc_features[-1], f, padding="replicate")
        self.up1 = UpSampleBlock(skip_input=f + enc_features[-2], output_features=f // 2, padding="reflection",
                                 is_depthwise=dw_upconv)
        self.wave1_ll = Conv3x3(f // 2, 1, padding="replicate")
        self.wave1 = Conv3x3(f // 2, 3, padding="zero", is_depthwise=dw_waveconv)
        for k in range(2, 5):
            setattr(self, "up%d" % k, UpSampleBlock(skip_input=f // 2 ** (k - 1) + enc_features[-1 - k],
                                                    output_features=f // 2 ** k, padding="reflection",
                                                    is_depthwise=dw_upconv))
            setattr(self, "wave%d" % k, Conv3x3(f // 2 ** k, 3, padding="zero", is_depthwise=dw_waveconv))
        self.sigmoid = nn.Sigmoid()

    def forward(self, x_blocks):
        _need_cuda(x_blocks)
        out = {}
        x = self.up1(self.conv2(x_blocks[-1]), x_blocks[-2])
        ll = (2 ** 4) * self.wave1_ll(x)
        out[("wavelets", 3, "LL")] = ll
        for k in range(1, 5):                                # level k emits scale 4 - k
            s = 4 - k
            if k > 1:
                x = getattr(self, "up%d" % k)(x, x_blocks[-1 - k])
            hcoef = (2 ** s) * getattr(self, "wave%d" % k)(x).unsqueeze(1)
            for j, band in enumerate(("LH", "HL", "HH")):
                out[("wavelets", s, band)] = hcoef[:, :, j]
            ll = self.iwt((ll, [hcoef]))
            out[("disp", s)] = ll // 2 if s == 1 else ll / (2 ** s)
        return out
