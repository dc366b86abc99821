"""Deterministic synthetic weights and encoder features (no datasets / checkpoints offline).

Everything is drawn from ``numpy.random.RandomState`` (the frozen legacy MT19937
stream), so the build container, the GPU box, the golden-fixture generator, the
tests and both bench arms all see bit-identical tensors for a given seed.

* ``random_state_dict`` fills conv weights/biases with PyTorch's default Conv2d
  scale, U(-1/sqrt(fan_in), 1/sqrt(fan_in)) (what the reference gets from
  ``nn.Conv2d`` at KITTI/layers.py:153, NYUv2/networks/layers.py:27), with an
  optional per-key gain (used to make the coefficient heads produce
  non-degenerate threshold masks: with the default scale every |yh| is either all
  above or all below ``thresh_ratio * range(LL)``, SURVEY 8d "sparsity caveat").
* ``blocky_features`` makes encoder-like feature maps: non-negative, piecewise
  constant on a coarse grid (so detail coefficients concentrate on block edges,
  like depth discontinuities) plus low-amplitude texture.
"""
from collections import OrderedDict

import numpy as np
import torch

# encoder output shapes of the BASELINE configs, fine -> coarse
RESNET18_CH = (64, 64, 128, 256, 512)
RESNET50_CH = (64, 256, 512, 1024, 2048)
DENSENET161_CH = (96, 96, 192, 384, 2208)

# The synthetic workload bench.py measures and the full-size parity tests check (one definition for both):
# seeded PyTorch-default init with high-pass coefficient heads x4 and blocky features, cell 16 px.
KITTI_HEAD_KEYS = ["decoder.%d.2.conv." % k for k in (3, 4, 7, 8, 11, 12, 15, 16)]   # +/- coefficient heads' 3x3 stage
BENCH_SYNTH = dict(param_seed=7, feat_seed=1000, cell=16, texture=0.01, head_gain=4.0)


def bench_kitti_params(module_or_shapes):
    """Bench weights: loads them into a module (returns the state dict) or builds the dict from {name: shape}."""
    gains = {k: BENCH_SYNTH["head_gain"] for k in KITTI_HEAD_KEYS}
    if isinstance(module_or_shapes, dict):
        return random_state_dict(module_or_shapes, BENCH_SYNTH["param_seed"], gains, KITTI_HEAD_KEYS)
    return load_random(module_or_shapes, seed=BENCH_SYNTH["param_seed"], gains=gains, highpass=KITTI_HEAD_KEYS)


def bench_kitti_features(n, height, width, ch, first_sample=0, pin=False):
    """Bench features of samples first_sample .. first_sample+n-1 (sample k of any batch is the same tensor)."""
    return blocky_features(kitti_feature_shapes(n, height, width, ch), seed=BENCH_SYNTH["feat_seed"] + first_sample,
                           cell=BENCH_SYNTH["cell"], texture=BENCH_SYNTH["texture"], pin=pin)


def random_state_dict(shapes, seed, gains=None, highpass=None):
    """shapes: ordered {name: shape}; names ending in '.weight' / '.bias' are filled, others skipped.

    gains: optional {substring: factor}; every filled tensor whose name contains
    the substring is multiplied by factor (first match wins).
    highpass: optional list of substrings; matching 3x3 weights get their spatial
    mean removed per (out,in) pair and matching biases are zeroed, i.e. the conv
    becomes a bank of zero-sum (edge / detail) filters.  Applied to the
    coefficient heads this makes |yh| vanish on flat regions and peak at feature
    discontinuities - the structure trained detail heads have (README.md:19-47)
    and the only way random weights yield clustered, sparse threshold masks.
    """
    rs = np.random.RandomState(seed)
    out = OrderedDict()
    bound = 1.0
    for name, shape in shapes.items():
        shape = tuple(int(v) for v in shape)
        if name.endswith(".weight") and len(shape) == 4:
            fan_in = shape[1] * shape[2] * shape[3]
            bound = 1.0 / np.sqrt(fan_in)
        elif not name.endswith(".bias"):
            continue
        arr = rs.uniform(-bound, bound, size=shape).astype(np.float32)
        if highpass and any(sub in name for sub in highpass):
            if arr.ndim == 4 and arr.shape[2] == 3:
                arr = (arr - arr.mean(axis=(2, 3), keepdims=True)).astype(np.float32)
            elif arr.ndim == 1:
                arr = np.zeros_like(arr)
        if gains:
            for sub, g in gains.items():
                if sub in name:
                    arr = (arr * np.float32(g)).astype(np.float32)
                    break
        out[name] = torch.from_numpy(arr)
    return out


def module_shapes(module):
    """Ordered {name: shape} of a module's conv parameters (buffers such as the IDWT taps are skipped)."""
    return OrderedDict((k, tuple(v.shape)) for k, v in module.state_dict().items()
                       if k.endswith(".weight") or k.endswith(".bias"))


def load_random(module, seed, gains=None, highpass=None):
    """Fill ``module``'s conv parameters in place from ``random_state_dict``; returns the dict used."""
    sd = random_state_dict(module_shapes(module), seed, gains, highpass)
    missing = module.load_state_dict(sd, strict=False)
    assert not missing.unexpected_keys
    return sd


def blocky_features(shapes, seed, cell=8, texture=0.05, pin=False):
    """List of fp32 CPU tensors with the given (N,C,H,W) shapes.

    value = U[0,1) constant over ``cell x cell`` blocks (offset per channel) +
    texture * U[0,1) per pixel.  Sample n of every map uses stream seed+n so a
    rank's shard equals the corresponding slice of the global batch.
    """
    feats = []
    for shape in shapes:
        n, c, h, w = (int(v) for v in shape)
        t = torch.empty((n, c, h, w), dtype=torch.float32)
        if pin and torch.cuda.is_available():
            t = t.pin_memory()
        feats.append(t)
    n = int(shapes[0][0])
    for b in range(n):
        rs = np.random.RandomState(seed + b)
        for t in feats:
            _, c, h, w = t.shape
            gh, gw = -(-h // cell), -(-w // cell)
            coarse = rs.uniform(0.0, 1.0, size=(c, gh, gw)).astype(np.float32)
            img = np.repeat(np.repeat(coarse, cell, axis=1), cell, axis=2)[:, :h, :w]
            img = img + np.float32(texture) * rs.uniform(0.0, 1.0, size=(c, h, w)).astype(np.float32)
            t[b] = torch.from_numpy(np.ascontiguousarray(img))
    return feats


def kitti_feature_shapes(n, height, width, num_ch_enc):
    """Encoder pyramid shapes for an (height x width) image: strides 2,4,8,16,32 (resnet_encoder.py:87-98)."""
    return [(n, int(c), height // (2 << k), width // (2 << k)) for k, c in enumerate(num_ch_enc)]


def nyu_feature_shapes(n, height, width, enc_features):
    """DenseNet/ResNet block shapes used by the NYUv2 decoders (densenet_encoder.py:26-33): strides 2,4,8,16,32."""
    return [(n, int(c), height // (2 << k), width // (2 << k)) for k, c in enumerate(enc_features)]
