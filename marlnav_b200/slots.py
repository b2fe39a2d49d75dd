"""Reusable output slots for ``Env.step``.

``Env.step`` must hand out FRESH tensors every step: the reference's rollout buffer keeps a reference to
every step's rewards (/root/reference/marlnav/models.py:121).  But allocating them and cutting the six
``Observations`` views costs more host time than the kernel launch at small batches.  A slot is one
buffer with all its views (and its launch arguments) prebuilt; it is handed out again only when nobody
outside holds any of its tensors:

* every Python object of the slot is back at its resting reference count (a caller keeping ``rew`` or
  ``obs.target_angle`` itself), AND
* the buffer's storage is shared by no tensor beyond the slot's own (a caller keeping a slice of a
  slice: a new tensor on the same storage), AND
* the slot was last used on the same CUDA stream (like torch's caching allocator, which only
  recycles a block on the stream it was allocated on).

Busy slots are skipped and a new one is made, so a caller that keeps T steps simply grows the ring to T
slots (capped by count and bytes; past the cap, or on a torch without the storage use-count hook, or
without the GIL -- reference counts are exact only with it --, every step allocates).
"""
import sys
from collections import deque

import torch

_storage_use_count = getattr(torch._C, '_storage_Use_Count', None)     # tensors sharing a storage
if not getattr(sys, '_is_gil_enabled', lambda: True)():
    _storage_use_count = None


def _refcounts(objs, _getref=sys.getrefcount):
    """Reference counts of a slot's tensors; baseline and check go through this one function so
    that the temporaries of the counting itself cancel."""
    return [_getref(o) for o in objs]


class OutputSlots:
    """Ring of output slots.  ``make()`` returns a dict with at least ``objs`` (every Python object
    handed to the caller: tensors and the namedtuple holding them) and ``storage`` (their common
    ``UntypedStorage``); the pool adds its bookkeeping keys to it."""

    MAX_SLOTS = 4096
    MAX_BYTES = 2 << 30
    TRIES = 3            # a caller may be holding some slots for long: look at a few before allocating

    def __init__(self, make):
        self._make = make
        self._ring = deque()
        self._bytes = 0

    def __len__(self):
        return len(self._ring)

    def _new(self, stream):
        slot = self._make()
        slot['stream'] = stream
        slot['nbytes'] = slot['storage'].nbytes()
        slot['rest'] = _refcounts(slot['objs'])
        slot['shared'] = _storage_use_count(slot['storage']._cdata) if _storage_use_count else -1
        return slot

    def is_free(self, slot, stream):
        if slot['stream'] != stream or not _storage_use_count or \
                _storage_use_count(slot['storage']._cdata) != slot['shared']:
            return False
        return _refcounts(slot['objs']) == slot['rest']

    def take(self, stream):
        ring = self._ring
        for _ in range(min(len(ring), self.TRIES)):
            slot = ring[0]
            ring.rotate(-1)
            if self.is_free(slot, stream):
                return slot
        slot = self._new(stream)
        if _storage_use_count is None:                 # nothing could ever be recycled: do not hoard
            return slot
        if len(ring) < self.MAX_SLOTS and self._bytes + slot['nbytes'] <= self.MAX_BYTES:
            ring.append(slot)
            self._bytes += slot['nbytes']
        return slot
