"""Drop-in for the ``pytorch_wavelets`` API surface the reference uses, on libwmd kernels.

Reference call sites: ``from pytorch_wavelets import IDWT`` (KITTI/networks/decoders/
depth_decoder.py:15; NYUv2/networks/decoders/densedepth_decoder.py:10) and
``from pytorch_wavelets import DWT`` (NYUv2/train.py:21).  Constructor signatures,
call conventions (``IDWT((yl, [yh]))``, ``DWT(x) -> (yl, [yh_1..yh_J])``, ``yh_j`` of
shape (N,C,3,H_j,W_j), bands LH,HL,HH, finest first), buffer names in the
state dict (``g0_col`` ... / ``h0_col`` ...) and differentiability are kept, so
``sys.modules['pytorch_wavelets'] = wavelet_monodepth_b200.wavelets`` makes the
unmodified reference run on these kernels.

Only what the reference exercises is implemented natively: the Haar wavelet,
even sizes (for which the boundary ``mode`` is irrelevant).  Anything else raises.
Autograd: the transforms are orthonormal, so IDWT.backward = DWT and vice versa
(KITTI/trainer.py:208-212 and NYUv2/train.py:327 back-propagate through the IDWT).
"""
import math

import torch
import torch.nn as nn

from . import ops

_S = 1.0 / math.sqrt(2.0)


def _check_wave(wave):
    if wave not in ("haar", "db1"):
        raise NotImplementedError("libwmd implements the Haar wavelet only (the reference uses wave='haar')")


class _IDWTFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, ll, hf):
        return ops.idwt_haar(ll, hf)

    @staticmethod
    def backward(ctx, grad):
        g_ll, g_hf = ops.dwt_haar(grad)
        return g_ll, g_hf


class _DWTFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        ll, hf = ops.dwt_haar(x)
        return ll, hf

    @staticmethod
    def backward(ctx, g_ll, g_hf):
        return ops.idwt_haar(g_ll, g_hf)


def idwt_level(ll, hf):
    """One differentiable synthesis level; ll (N,C,H,W), hf (N,C,3,H,W)."""
    return _IDWTFn.apply(ll, hf)


def dwt_level(x):
    """One differentiable analysis level."""
    return _DWTFn.apply(x)


def _col(vals):
    return torch.tensor(vals, dtype=torch.float32).reshape(1, 1, 2, 1)


def _row(vals):
    return torch.tensor(vals, dtype=torch.float32).reshape(1, 1, 1, 2)


class DWTInverse(nn.Module):
    """2-D inverse DWT.  ``forward((yl, yh))``: yl (N,C,H,W), yh list of (N,C,3,H_j,W_j), finest first."""

    def __init__(self, wave="db1", mode="zero"):
        super().__init__()
        _check_wave(wave)
        # non-trainable taps kept only for state-dict compatibility with checkpoints of the reference
        self.register_buffer("g0_col", _col((_S, _S)))
        self.register_buffer("g1_col", _col((_S, -_S)))
        self.register_buffer("g0_row", _row((_S, _S)))
        self.register_buffer("g1_row", _row((_S, -_S)))
        self.mode = mode

    def forward(self, coeffs):
        yl, yh = coeffs
        ll = yl
        for h in yh[::-1]:
            if h is None:
                h = torch.zeros(ll.shape[0], ll.shape[1], 3, ll.shape[-2], ll.shape[-1], dtype=ll.dtype,
                                device=ll.device)
            if ll.shape[-2] > h.shape[-2]:
                ll = ll[..., :-1, :]
            if ll.shape[-1] > h.shape[-1]:
                ll = ll[..., :-1]
            ll = idwt_level(ll, h)
        return ll


class DWTForward(nn.Module):
    """2-D forward DWT with J levels.  ``forward(x) -> (yl, [yh_1 (finest) ... yh_J])``."""

    def __init__(self, J=1, wave="db1", mode="zero"):
        super().__init__()
        _check_wave(wave)
        self.register_buffer("h0_col", _col((_S, _S)))
        self.register_buffer("h1_col", _col((_S, -_S)))
        self.register_buffer("h0_row", _row((_S, _S)))
        self.register_buffer("h1_row", _row((_S, -_S)))
        self.J = J
        self.mode = mode

    def forward(self, x):
        yh = []
        ll = x
        for _ in range(self.J):
            if ll.shape[-2] % 2 or ll.shape[-1] % 2:
                raise NotImplementedError("libwmd DWT needs even sizes at every level (got %s)" % (tuple(ll.shape),))
            ll, h = dwt_level(ll)
            yh.append(h)
        return ll, yh


IDWT = DWTInverse
DWT = DWTForward
