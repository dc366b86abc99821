"""Oracle: the reference's batch-1 "sparse convolution" op library, restated on torch-CPU.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Follows KITTI/layers.py:335-508
(NYUv2/networks/layers.py:82-223 is a near-copy without the 1x1 branch).

Wire format between ops (KITTI/layers.py:358,402,440,458,496):
  xvals   1-D fp32, length C*M, channel-major: element (c, m) at c*M + m, m
          enumerating active pixels in row-major (y, x) order;
  xidxmap (1,1,H,W) int64, -1 where inactive else m;
  masks   (1,1,H,W) bool or 0/1 float.

The formulation here is coordinate-based (compute each tap's source pixel, apply
the border rule, look the row up in the index map) instead of the reference's
"pad the index map, slice nine shifted boolean masks" construction
(layers.py:444-453); both enumerate outputs row-major and read input pixel
(y+ky-1, x+kx-1) for tap (ky,kx), so the gathered (C*9, M_out) matrix - and hence
the matmul result - is identical.
"""
import torch
import torch.nn.functional as F


def _check_single(mask):
    # layers.py:372-373,383-384: batch 1, one channel
    assert mask.shape[0] == 1 and mask.shape[1] == 1, "sparse ops are batch-1, single-channel masks"


def active_coords(mask):
    """(2, M) int64 rows/cols of active pixels, row-major.  [layers.py:371-379 mask2yx]"""
    _check_single(mask)
    yx = torch.nonzero(mask[0, 0] > 0.5 if mask.dtype != torch.bool else mask[0, 0], as_tuple=False)
    return yx.t().contiguous()


def index_map(mask):
    """idxmap (1,1,H,W) int64 (-1 inactive, else running index) and op count H*W.  [layers.py:382-389 mask2idxmap]"""
    _check_single(mask)
    b = (mask > 0.5) if mask.dtype != torch.bool else mask
    flat = b.reshape(-1)
    run = torch.cumsum(flat.to(torch.int64), 0) - 1
    idx = torch.where(flat, run, torch.full_like(run, -1))
    return idx.reshape(mask.shape), mask.shape[2] * mask.shape[3]


def scatter_dense(xvals, xchn, mask):
    """Dense (1,C,H,W) with xvals at active pixels, zero elsewhere.  [layers.py:365-368 make_result]"""
    h, w = mask.shape[2:]
    yx = active_coords(mask)
    out = torch.zeros(xchn, h * w, dtype=xvals.dtype)
    out[:, yx[0] * w + yx[1]] = xvals.reshape(xchn, -1)
    return out.reshape(1, xchn, h, w)


def _rows_with_zero_column(xvals, xchn):
    m = xvals.numel() // xchn
    return torch.cat([torch.zeros(xchn, 1, dtype=xvals.dtype), xvals.reshape(xchn, m)], 1)


def select(xvals, xchn, xidxmap, ymask, ufactor=1, pad=False):
    """Re-index sparse features onto the active set of ``ymask``.  [layers.py:337-362 sparse_select]

    ufactor=2 reads the half-resolution source pixel (y//2, x//2).  pad=True maps
    misses to a zero column; without pad every target must be active in xidxmap.
    """
    xh, xw = xidxmap.shape[2:]
    assert xh * ufactor == ymask.shape[2] and xw * ufactor == ymask.shape[3]
    yx = active_coords(ymask)
    if ufactor == 2:
        yx = yx // 2
    rows = xidxmap[0, 0][yx[0], yx[1]]
    if pad:
        table = _rows_with_zero_column(xvals, xchn)
        rows = rows + 1
    else:
        assert bool((rows >= 0).all()), "select without pad hit an inactive source pixel"
        table = xvals.reshape(xchn, -1)
    return table[:, rows].reshape(-1)


def conv1x1(weight, bias, xvals, nonlin):
    """Per-active-pixel 1x1 conv.  [layers.py:392-406 sparse_conv1x1]  Returns (vals (Cout,M), Cout, ops)."""
    ochn, ichn = weight.shape[:2]
    m = xvals.numel() // ichn
    out = nonlin(weight.reshape(ochn, ichn) @ xvals.reshape(ichn, m) + bias.reshape(ochn, 1))
    return out, ochn, m * ichn * ochn + m * ochn


def _tap_rows(xidxmap, mask, padding):
    """(9, M_out) int64 rows into the zero-column-prefixed table (0 = zero column)."""
    h, w = mask.shape[2:]
    yx = active_coords(mask)
    table = xidxmap[0, 0] + 1
    rows = []
    for ky in range(3):
        for kx in range(3):
            qy = yx[0] + (ky - 1)
            qx = yx[1] + (kx - 1)
            if padding == "reflect":
                qy = torch.where(qy < 0, -qy, qy)
                qy = torch.where(qy >= h, 2 * (h - 1) - qy, qy)
                qx = torch.where(qx < 0, -qx, qx)
                qx = torch.where(qx >= w, 2 * (w - 1) - qx, qx)
                r = table[qy, qx]
            elif padding == "replicate":
                r = table[qy.clamp(0, h - 1), qx.clamp(0, w - 1)]
            elif padding == "constant":
                inside = (qy >= 0) & (qy < h) & (qx >= 0) & (qx < w)
                r = torch.where(inside, table[qy.clamp(0, h - 1), qx.clamp(0, w - 1)],
                                torch.zeros_like(qy))
            else:
                raise ValueError(padding)
            rows.append(r)
    return torch.stack(rows, 0)


def conv3x3(weight, bias, xvals, xidxmap, mask, nonlin=None, padding="reflect", make_result=True):
    """Sparse 3x3 conv on raw (weight (Cout,Cin,3,3), bias) tensors.  [layers.py:409-480 sparse_conv3x3]

    Inputs live on the active set of ``xidxmap``; outputs are produced at the
    active pixels of ``mask``.  A tap whose source pixel is not active in
    ``xidxmap`` contributes zero (layers.py:439-444).  Border handling pads the
    *index map* with ``padding`` in {'reflect','replicate','constant'} (:444).
    Returns (dense (1,Cout,H,W), ops) if make_result else (flat (Cout*M,), Cout, ops).
    """
    ochn, ichn = weight.shape[:2]
    table = _rows_with_zero_column(xvals, ichn)
    rows = _tap_rows(xidxmap, mask, padding)                    # (9, M)
    m_out = rows.shape[1]
    vals = table[:, rows.reshape(-1)]                           # (Cin, 9*M)
    ops = vals.numel()
    vals = vals.reshape(ichn * 9, m_out)                        # row c*9 + tap, as weight.reshape(Cout,-1)
    out = weight.reshape(ochn, ichn * 9) @ vals + bias.reshape(ochn, 1)
    ops += (1 + 9 * ichn) * m_out * ochn
    if nonlin is not None:
        out = nonlin(out)
    if make_result:
        return scatter_dense(out.reshape(-1), ochn, mask), ops
    return out.reshape(-1), ochn, ops


def head3x3(w1, b1, w2, b2, xvals, xidxmap, mask, nonlin):
    """1x1 -> LeakyReLU(0.1) on every active input, then sparse 3x3 (reflect) -> nonlin, dense result.

    This is sparse_conv3x3's nn.Sequential branch (layers.py:426-431) as the
    KITTI coefficient heads use it (depth_decoder.py:276-290).
    """
    mid, ichn, ops1 = conv1x1(w1, b1, xvals, lambda t: F.leaky_relu(t, 0.1))
    dense, ops2 = conv3x3(w2, b2, mid.reshape(-1), xidxmap, mask, nonlin=nonlin,
                          padding="reflect", make_result=True)
    return dense, ops1 + ops2


def upsample_concat(xvals, xchn, xidxmap, skip, mask, make_result=True):
    """Nearest x2 upsample of sparse features + channel-concat of a dense skip map, at ``mask``.

    [layers.py:483-508 sparse_upsample]  For each active hi-res pixel (y,x): the
    low-res feature at (y//2, x//2) (must be active in xidxmap) followed by
    skip[:, y, x].
    """
    yx = active_coords(mask)
    rows = xidxmap[0, 0][yx[0] // 2, yx[1] // 2]
    assert bool((rows >= 0).all()), "upsample hit an inactive low-res pixel"
    lo = xvals.reshape(xchn, -1)[:, rows]
    sk = skip[0][:, yx[0], yx[1]]
    vals = torch.cat([lo.reshape(-1), sk.reshape(-1)], 0)
    ochn = xchn + skip.shape[1]
    if make_result:
        return scatter_dense(vals, ochn, mask)
    return vals, ochn
