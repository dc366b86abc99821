"""CUDA-graph replay of the sparse decoder forward (serving mode).

The native forward (`_WaveDecoderBase._native_forward`) is ~70 launches of libwmd kernels with static shapes: every
buffer is sized by capacity, the active counts stay on the device, and the only host read is the one at the very end
for ``total_ops``.  Capturing it once removes the per-launch host cost (Python + ctypes + allocator) from every later
step; the kernels, their order and their results are exactly the eager ones.

A graph is bound to the tensors it was captured with: a producer (the encoder) has to write its features into those
same tensors, which is how a CUDA-graphed encoder behaves anyway.
"""
import torch

from . import _lib
from .kitti_decoders import SparseDepthWaveProgressiveDecoder, WmdError


class GraphedSparseDecoder:
    """decoder(features, thresh_ratio) captured into one CUDA graph.

    >>> g = GraphedSparseDecoder(decoder, features, 0.05)
    >>> out = g.replay()            # same dict as decoder(features, 0.05); tensors are reused by the next replay
    """

    def __init__(self, decoder, features, thresh_ratio=0.05, sparse_scales=(0, 1, 2, 3), warmup=2):
        if not isinstance(decoder, SparseDepthWaveProgressiveDecoder):
            raise WmdError("GraphedSparseDecoder wraps a SparseDepthWaveProgressiveDecoder")
        self.decoder = decoder
        self.features = list(features)
        self.thresh_ratio = float(thresh_ratio)
        self.sparse_levels = tuple(i for i in range(1, 4) if i in sparse_scales)
        if any((i + 1) in self.sparse_levels and i not in self.sparse_levels for i in range(1, 4)):
            raise NotImplementedError("a dense level below a sparse level is not defined by the reference either")
        for _ in range(max(1, warmup)):                     # packs the weights, sizes the scratch buffers
            decoder._native_forward(self.features, self.thresh_ratio, self.sparse_levels, with_masks=True)
        torch.cuda.synchronize()
        l0 = _lib.launch_count()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self._out, self._counts = decoder._native_forward(self.features, self.thresh_ratio, self.sparse_levels,
                                                              with_masks=True)
        self.launches = _lib.launch_count() - l0            # libwmd kernels per replay

    def bound_to(self, features):
        """True if `features` are the tensors this graph reads."""
        return len(features) == len(self.features) and all(
            a.data_ptr() == b.data_ptr() and a.shape == b.shape for a, b in zip(features, self.features))

    def replay(self):
        self.graph.replay()
        out = dict(self._out)
        if self.decoder.count_ops:
            self.decoder._add_total_ops(out, self._counts, self.features)   # the one host read, as in eager mode
        return out
