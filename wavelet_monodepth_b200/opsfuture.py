"""Asynchronous read-back of the active-pixel counts behind ``total_ops``.

The reference's analytic op counter (depth_decoder.py:299-427) needs the sizes of the compacted active sets, which the
CUDA path keeps on the device.  Reading them with ``.cpu()`` at the end of every forward is a host synchronisation
per step: the host cannot enqueue step k+1 (or the all-gather of step k) before step k has drained.  ``OpsFuture``
instead enqueues ONE non-blocking device->pinned-host copy of the stacked count tensor plus an event on the forward's
stream and evaluates the closed-form counter (opcount.py) only when somebody asks for the number.

    decoder.count_ops = "async"
    out = decoder(features, 0.05)        # returns without waiting for the device
    fut = out["total_ops"]               # OpsFuture
    ...                                  # enqueue more work
    fut.result()["total_ops"]            # int; waits for that forward's event only

With ``count_ops = True`` (the default, reference-compatible) the decoders call ``result()`` right away and store plain
Python ints under the reference's keys.
"""
import numpy as np
import torch

_RING = 4


class _PinnedRing:
    """A few pinned host buffers per (device, shape): page-locking memory is far too slow to do per forward."""

    def __init__(self):
        self.slots = {}

    def take(self, device, shape, owner):
        key = (str(device), tuple(shape))
        ring = self.slots.setdefault(key, {"bufs": [], "owners": [], "next": 0})
        if len(ring["bufs"]) < _RING:
            ring["bufs"].append(torch.empty(tuple(shape), dtype=torch.int32).pin_memory())
            ring["owners"].append(None)
            k = len(ring["bufs"]) - 1
        else:
            k = ring["next"]
            ring["next"] = (k + 1) % _RING
            prev = ring["owners"][k]
            if prev is not None:
                prev._finalise()                        # an old forward: its event has long completed
        ring["owners"][k] = owner
        return ring["bufs"][k]


_ring = _PinnedRing()


class OpsFuture:
    """``total_ops`` (and its per-level / per-sample breakdown) of one forward, evaluated on demand."""

    def __init__(self, counts_dev, finish):
        """counts_dev: int32 CUDA tensor of active counts (or None when every level is dense);
        finish(np.int64 array or None) -> dict of the reference's total_ops entries."""
        self._finish = finish
        self._value = None
        self._host = None
        self._np = None
        self._event = None
        if counts_dev is not None:
            dev = counts_dev.device
            self._host = _ring.take(dev, counts_dev.shape, self)
            self._host.copy_(counts_dev, non_blocking=True)
            self._event = torch.cuda.Event()
            self._event.record(torch.cuda.current_stream(dev))

    def done(self):
        return self._event is None or self._event.query()

    def _finalise(self):
        """Take the counts out of the shared pinned buffer (waits for this forward's copy only)."""
        if self._host is not None:
            self._event.synchronize()
            self._np = self._host.numpy().astype(np.int64)       # astype copies: the ring slot can be reused
            self._host = None

    def result(self):
        if self._value is None:
            self._finalise()
            self._value = self._finish(self._np)
        return self._value

    def __int__(self):
        return int(self.result()["total_ops"])

    __index__ = __int__

    def __repr__(self):
        return "OpsFuture(%s)" % (self._value["total_ops"] if self._value is not None else "pending")
