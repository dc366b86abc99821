"""The reference's analytic FLOP counter (`total_ops`), as closed-form host arithmetic.

The reference accumulates Python ints while it runs (KITTI/networks/decoders/
depth_decoder.py:246-266,299-427; KITTI/layers.py:388,405,462,469; NYUv2/networks/
decoders/densedepth_decoder.py:276-408; NYUv2/networks/layers.py:177,184).  The only
data-dependent inputs are the active-pixel counts of the compacted sets, which
the CUDA path reads back once per forward; everything else is shape arithmetic.
`total_ops` is a known-answer test (17 473 692 295 for KITTI ResNet50 1024x320,
33 463 546 800 for NYU DenseNet161 640x480 with all masks full), so the formulas
are reproduced term by term, quirks included.
"""


def sparse_conv3x3_ops(cin, cout, m_out):
    # layers.py:462 (gathered elements) + :469 (matmul with bias)
    return cin * 9 * m_out + (1 + 9 * cin) * m_out * cout


def sparse_conv1x1_ops(cin, cout, m):
    # layers.py:405
    return m * cin * cout + m * cout


def dense_conv3x3_ops(cin, cout, h, w):
    # depth_decoder.py:386-387,396-397: bias counted once per output channel, not per pixel
    return (1 + 9 * cin * h * w) * cout


def dense_head_ops(cin, cmid, cout, h, w):
    # depth_decoder.py:247-266: Conv1x1(cin->cmid) then Conv3x3(cmid->cout)
    return (1 + cin * h * w) * cmid + (1 + 9 * cmid * h * w) * cout


def kitti_level_ops(i, h, w, cin0, c, cskip, sparse, m2=None, m4=None, m5=None):
    """Ops of KITTI level i on low-res grid h x w (one sample).

    cin0: channels entering upconv(i,0); c: num_ch_dec[i]; cskip: encoder skip channels.
    sparse levels need the active counts m2=|upconv0_mask|, m4=|upconv1_mask|, m5=|wavelet_mask|.
    """
    ops = (3 * h * w if i != 4 else 0) + 25 * h * w + 100 * h * w      # :310,322-323
    if sparse:
        ops += h * w + h * w + 4 * h * w + 4 * h * w                      # four mask2idxmap, :333-340
        ops += sparse_conv3x3_ops(cin0, c, m2)
        ops += sparse_conv3x3_ops(c + cskip, c, m4)
        ops += 2 * (sparse_conv1x1_ops(c, c, m4) + sparse_conv3x3_ops(c, 3, m5))
    else:
        ops += dense_conv3x3_ops(cin0, c, h, w)
        ops += dense_conv3x3_ops(c + cskip, c, 2 * h, 2 * w)
        if i == 4:
            ops += dense_head_ops(c, c // 4, 1, 2 * h, 2 * w)
        ops += 2 * dense_head_ops(c, c, 3, 2 * h, 2 * w)
    ops += 4 * (4 * h) * (4 * w)                                            # IDWT, :373,417
    return ops


def nyu_dense_part_ops(c_in, h, w, features, c_skip):
    """conv2 + up1 + wave1/wave1_ll + first IDWT of SparseDecoderWave (densedepth_decoder.py:276-311)."""
    ops = (1 + 9 * c_in) * h * w * features                               # conv2
    ops += (1 + 9 * (features + c_skip)) * (2 * h) * (2 * w) * (features // 2)   # up1
    ops += (1 + 9 * (features // 2)) * (2 * h) * (2 * w) * 4              # wave1 + wave1_ll
    ops += (4 * h) * (4 * w)                                              # IDWT counts 1 per output pixel
    return ops


def nyu_sparse_block_ops(h, w, cin, cout, m4, m5, second):
    """One sparse NYU block on low-res grid h x w (densedepth_decoder.py:316-359 / :363-406).

    cin: channels entering convA (carried + skip), cout: convA outputs, m4=|wave_mask|, m5=|wavelet_mask|.
    """
    ops = 3 * h * w + 25 * h * w + 100 * h * w
    ops += 4 * h * w + 4 * h * w + 4 * h * w + h * w                      # wavelet, conva, wave, up index maps
    if second:
        ops += 4 * h * w                                                   # wave_mask indexed twice, :381-382
    ops += sparse_conv3x3_ops(cin, cout, m4)
    ops += sparse_conv3x3_ops(cout, 3, m5)
    ops += (4 * h) * (4 * w)
    return ops
