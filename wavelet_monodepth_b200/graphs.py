"""CUDA-graph replay of the sparse decoders' forward (serving mode).

The native forward (`_WaveDecoderBase._native_forward`, `_NyuWaveBase._native_forward`) is a few dozen launches of
libwmd kernels with static shapes: every buffer is sized by capacity and the active counts stay on the device.
Capturing it once removes the per-launch host cost (Python + ctypes + allocator) from every later step; the kernels,
their order and their results are exactly the eager ones.

Nothing in a replay waits for the device: the count tensor behind ``total_ops`` is read back by one non-blocking copy
into pinned memory enqueued after the graph (opsfuture.OpsFuture) and is evaluated on demand - with
``decoder.count_ops == "async"`` the caller gets the future, with ``True`` the ints (one wait per replay).

A graph is bound to the tensors it was captured with: a producer (the encoder) has to write its features into those
same tensors, which is how a CUDA-graphed encoder behaves anyway.  Scratch buffers the capture used are kept alive by
ops._scratch (buffers are retired, never freed) and, being allocated on the capture stream, are not shared with eager
calls on other streams.
"""
import gc

import torch

from . import _lib
from .kitti_decoders import SparseDepthWaveProgressiveDecoder, WmdError
from .nyu_decoders import SparseDecoderWave


class GraphedSparseDecoder:
    """decoder(features, thresh_ratio) captured into one CUDA graph.

    >>> g = GraphedSparseDecoder(decoder, features, 0.05)
    >>> out = g.replay()            # same dict as decoder(features, 0.05); tensors are reused by the next replay

    decoder: ``SparseDepthWaveProgressiveDecoder`` (KITTI; ``sparse_scales`` as in its forward) or
    ``SparseDecoderWave`` (NYUv2)."""

    def __init__(self, decoder, features, thresh_ratio=0.05, sparse_scales=(0, 1, 2, 3), warmup=2):
        self.decoder = decoder
        self.features = list(features)
        self.thresh_ratio = float(thresh_ratio)
        if isinstance(decoder, SparseDepthWaveProgressiveDecoder):
            self.sparse_levels = decoder._sparse_levels(sparse_scales)
        elif isinstance(decoder, SparseDecoderWave):
            self.sparse_levels = None
        else:
            raise WmdError("GraphedSparseDecoder wraps a SparseDepthWaveProgressiveDecoder or a SparseDecoderWave")
        self.device = next(f.device for f in self.features if f.is_cuda)
        with torch.cuda.device(self.device):
            for _ in range(max(1, warmup)):                 # packs the weights, sizes the scratch buffers
                self._run()
            # Destroying a CUDA graph is not permitted while ANY stream of the process is capturing (global capture
            # mode): it would invalidate this capture.  Make sure no dead graph object is waiting for the cycle
            # collector (this class itself holds no reference cycles, so `del graph` frees it at once).
            gc.collect()
            torch.cuda.synchronize()
            l0 = _lib.launch_count()
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self._out, self._counts = self._run()
            self.launches = _lib.launch_count() - l0        # libwmd kernels per replay

    def _run(self):
        if self.sparse_levels is None:
            return self.decoder._native_forward(self.features, self.thresh_ratio, sparse=True)
        return self.decoder._native_forward(self.features, self.thresh_ratio, self.sparse_levels, with_masks=True)

    def _future(self):
        if self.sparse_levels is None:
            return self.decoder.ops_future(self._counts, self.features)
        return self.decoder.ops_future(self._counts, self.features, self.sparse_levels)

    def bound_to(self, features):
        """True if `features` are the tensors this graph reads."""
        return len(features) == len(self.features) and all(
            a.data_ptr() == b.data_ptr() and a.shape == b.shape for a, b in zip(features, self.features))

    def replay(self):
        with torch.cuda.device(self.device):
            self.graph.replay()
            out = dict(self._out)
            mode = self.decoder.count_ops
            if mode:
                fut = self._future()                        # async copy of the counts + event; no host wait
                if mode == "async":
                    out["total_ops"] = fut
                else:
                    out.update(fut.result())
        return out
