"""Oracle: Haar DWT / IDWT as ``pytorch_wavelets`` (v1.3.0) computes them.

TEST INFRASTRUCTURE (see oracle/__init__.py).  The package is an un-vendored
dependency of the reference (README.md:58-65; imported at
KITTI/networks/decoders/depth_decoder.py:15, NYUv2/networks/decoders/
densedepth_decoder.py:10, NYUv2/train.py:21) and is not present in the build
container, so this file restates its *published* algorithm:

* synthesis (``DWTInverse`` / alias ``IDWT``): per level, column pass then row
  pass, each pass being two grouped stride-2 transposed convolutions with the
  2-tap reconstruction filters (lo = [s, s], hi = [s, -s], s = 1/sqrt(2)) added
  together;
* analysis (``DWTForward`` / alias ``DWT``): per level, row pass then column pass,
  each a grouped stride-2 correlation with the *reversed* decomposition filters
  (lo = [s, s], hi = [s, -s] after reversal), band order (LL | LH, HL, HH).

For Haar and even sizes the boundary ``mode`` is irrelevant (zero samples of
padding are needed); that is the only regime the reference exercises
(depth_decoder.py:85,182; densedepth_decoder.py:99,101,234,236; NYUv2/train.py:258).
Odd sizes are handled for ``mode='zero'`` only (one trailing zero sample).

The reference's own closed form of the same synthesis is
``SparseDepthWaveProgressiveDecoder.my_iwt_once`` (depth_decoder.py:225-239);
``closed_form_idwt`` below restates it and tests check both agree to ~1e-6.

The module classes keep the dependency's constructor signature, buffer names
(``g0_col, g1_col, g0_row, g1_row`` / ``h0_col, h1_col, h0_row, h1_row``) and
call convention so that the *unmodified reference* can be imported against this
file (see oracle/pin_against_reference.py).
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

_S = 1.0 / math.sqrt(2.0)
_SUPPORTED = ("haar", "db1")


def _taps(kind):
    """Return (lo, hi) 2-tap filters as python floats.

    'rec': synthesis taps; 'dec': analysis taps already reversed, i.e. in the
    order a cross-correlation (torch conv2d) applies them.
    """
    if kind == "rec":
        return (_S, _S), (_S, -_S)
    # pywt dec_hi = [-s, s]; reversed for correlation -> [s, -s]
    return (_S, _S), (_S, -_S)


def _filt(vals, along):
    t = torch.tensor(vals, dtype=torch.float32)
    return t.reshape(1, 1, 2, 1) if along == "col" else t.reshape(1, 1, 1, 2)


def _synth_1d(lo, hi, g_lo, g_hi, dim):
    """One synthesis pass along ``dim`` (2 = height / 'col', 3 = width / 'row')."""
    c = lo.shape[1]
    stride = (2, 1) if dim == 2 else (1, 2)
    w_lo = g_lo.to(lo.dtype).repeat(c, 1, 1, 1)
    w_hi = g_hi.to(lo.dtype).repeat(c, 1, 1, 1)
    return (F.conv_transpose2d(lo, w_lo, stride=stride, groups=c)
            + F.conv_transpose2d(hi, w_hi, stride=stride, groups=c))


def _analysis_1d(x, h_lo, h_hi, dim, mode):
    """One analysis pass along ``dim``; returns channels interleaved [lo, hi] per input channel."""
    c = x.shape[1]
    n = x.shape[dim]
    if n % 2 == 1:
        if mode != "zero":
            raise NotImplementedError("odd sizes only restated for mode='zero'")
        pad = (0, 0, 0, 1) if dim == 2 else (0, 1, 0, 0)
        x = F.pad(x, pad)
    stride = (2, 1) if dim == 2 else (1, 2)
    w = torch.cat([h_lo, h_hi], 0).to(x.dtype).repeat(c, 1, 1, 1)
    return F.conv2d(x, w, stride=stride, groups=c)


class DWTInverse(nn.Module):
    """2-D inverse DWT, ``forward((yl, yh_list)) -> y``; ``yh[j]`` is (N, C, 3, H_j, W_j), finest first."""

    def __init__(self, wave="db1", mode="zero"):
        super().__init__()
        if wave not in _SUPPORTED:
            raise NotImplementedError("oracle restates the Haar wavelet only")
        lo, hi = _taps("rec")
        self.register_buffer("g0_col", _filt(lo, "col"))
        self.register_buffer("g1_col", _filt(hi, "col"))
        self.register_buffer("g0_row", _filt(lo, "row"))
        self.register_buffer("g1_row", _filt(hi, "row"))
        self.mode = mode

    def forward(self, coeffs):
        yl, yh = coeffs
        ll = yl
        for h in yh[::-1]:
            if h is None:
                h = torch.zeros(ll.shape[0], ll.shape[1], 3, ll.shape[-2], ll.shape[-1],
                                dtype=ll.dtype, device=ll.device)
            if ll.shape[-2] > h.shape[-2]:
                ll = ll[..., :-1, :]
            if ll.shape[-1] > h.shape[-1]:
                ll = ll[..., :-1]
            lh, hl, hh = torch.unbind(h, dim=2)
            lo = _synth_1d(ll, lh, self.g0_col, self.g1_col, 2)
            hi = _synth_1d(hl, hh, self.g0_col, self.g1_col, 2)
            ll = _synth_1d(lo, hi, self.g0_row, self.g1_row, 3)
        return ll


class DWTForward(nn.Module):
    """2-D forward DWT, ``forward(x) -> (yl, [yh_1 (finest) ... yh_J])``."""

    def __init__(self, J=1, wave="db1", mode="zero"):
        super().__init__()
        if wave not in _SUPPORTED:
            raise NotImplementedError("oracle restates the Haar wavelet only")
        lo, hi = _taps("dec")
        self.register_buffer("h0_col", _filt(lo, "col"))
        self.register_buffer("h1_col", _filt(hi, "col"))
        self.register_buffer("h0_row", _filt(lo, "row"))
        self.register_buffer("h1_row", _filt(hi, "row"))
        self.J = J
        self.mode = mode

    def forward(self, x):
        yh = []
        ll = x
        for _ in range(self.J):
            n, c = ll.shape[:2]
            rows = _analysis_1d(ll, self.h0_row, self.h1_row, 3, self.mode)
            both = _analysis_1d(rows, self.h0_col, self.h1_col, 2, self.mode)
            both = both.reshape(n, c, 4, both.shape[-2], both.shape[-1])
            ll = both[:, :, 0].contiguous()
            yh.append(both[:, :, 1:].contiguous())
        return ll, yh


# aliases the reference imports (`from pytorch_wavelets import IDWT`, `DWT`)
IDWT = DWTInverse
DWT = DWTForward


def closed_form_idwt(yl, yh):
    """The reference's own one-level closed form (depth_decoder.py:225-239).

    yl (N,1,H,W), yh (N,1,3,H,W) -> (N,1,2H,2W); out[2i+a, 2j+b] =
    1/2 (ll + (-1)^a lh + (-1)^b hl + (-1)^(a+b) hh).
    """
    half_l = yl / 2
    half_h = yh / 2
    lh, hl, hh = half_h[:, :, 0], half_h[:, :, 1], half_h[:, :, 2]
    quad = torch.cat([lh + hl + hh, lh - hl - hh, -lh + hl - hh, -lh - hl + hh], 1)
    return F.pixel_shuffle(half_l.expand(-1, 4, -1, -1) + quad, 2)
