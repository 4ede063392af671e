"""Peer output: the owner rank's result buffers mapped into every rank of one node (CUDA IPC over NVLink / NVSwitch).

The reference fills one HaplotypeLikelihoodArray per region on one host (haplotype_likelihood_array.cpp:64-103). With the reads
of a region (or the regions of a job) sharded over the GPUs of a node, the [H, R] values have to end up in ONE place. Instead of
a gather collective after the fact, every rank's epilogue kernel stores its values straight into the owner's HBM: the owner
allocates a ring of result slots (phmm_device_alloc), exports it (phmm_ipc_export), the other ranks map it (phmm_ipc_open) and
pass windows of it as ``out`` to ``PairHMMEngine.populate`` (phmm_populate_ld). What remains of the "gather" is one barrier per
step, which tells the owner that every rank's stores of that step have landed (a writer's stores are complete when its
populate call has returned).

Ring protocol (``n_buf`` slots): step k writes slot k % n_buf; every rank calls ``publish(k)`` after its populate k returned; the
owner may read slot k once ``publish(k)``'s event has fired and must be done with it before it calls ``publish(k + 1)``; a rank
makes its engine wait for the event of ``publish(k - n_buf + 1)`` (``wait_slot``) before populate k overwrites the slot.
"""
import ctypes as C

import numpy as np

from . import _lib


class PeerView:
    """A [rows, cols] float64 window (row stride ``ld`` elements) at a raw device address: the duck type ``PairHMMEngine.populate``
    needs of ``out`` (shape, dtype, stride, data_ptr). No torch tensor is made of a peer address, so that a writer process never
    touches the owner's device through torch."""

    def __init__(self, ptr, rows, cols, ld):
        import torch
        self._ptr, self.shape, self._ld, self.dtype = int(ptr), (int(rows), int(cols)), int(ld), torch.float64

    def data_ptr(self):
        return self._ptr

    def stride(self, dim):
        return self._ld if dim == 0 else 1


class _CudaArray:
    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"data": (int(ptr), False), "shape": tuple(int(x) for x in shape), "typestr": typestr,
                                         "strides": None, "version": 3}


class PeerBuffer:
    """``nbytes`` of device memory on rank ``owner``'s GPU, addressable from every rank of the (single-node) group."""

    def __init__(self, nbytes, device, rank, world, owner=0, group=None):
        import torch.distributed as dist
        self._lib = _lib.load()
        self.nbytes, self.device, self.rank, self.world, self.owner = int(nbytes), int(device), int(rank), int(world), int(owner)
        self.ptr, self._mapped = 0, False
        handle, err = None, None
        if rank == owner:
            p = C.c_void_p()
            rc = self._lib.phmm_device_alloc(self.device, self.nbytes, C.byref(p))
            if rc != _lib.PHMM_OK:
                err = "phmm_device_alloc(%d bytes) failed: %d" % (self.nbytes, rc)
            else:
                self.ptr = int(p.value)
                buf = C.create_string_buffer(64)
                if self._lib.phmm_ipc_export(self.ptr, buf) != _lib.PHMM_OK:
                    err = "phmm_ipc_export failed"
                else:
                    handle = buf.raw
        if world > 1:                       # the owner's failure reaches every rank: nobody is left waiting in the broadcast
            box = [handle, err]
            dist.broadcast_object_list(box, src=owner, group=group)
            handle, err = box
        if err is not None:
            self.close()
            raise RuntimeError(err)
        if rank != owner:
            p = C.c_void_p()
            rc = self._lib.phmm_ipc_open(self.device, handle, C.byref(p))
            if rc != _lib.PHMM_OK:
                raise RuntimeError("phmm_ipc_open failed (%d): no peer access to the owner's GPU?" % rc)
            self.ptr, self._mapped = int(p.value), True

    def view(self, byte_offset, rows, cols, ld=None):
        """A [rows, cols] float64 window starting ``byte_offset`` bytes into the buffer."""
        ld = cols if ld is None else ld
        assert byte_offset % 8 == 0 and byte_offset + 8 * ((rows - 1) * ld + cols) <= self.nbytes, "window outside the peer buffer"
        return PeerView(self.ptr + byte_offset, rows, cols, ld)

    def owner_tensor(self, byte_offset, shape):
        """The owner's own torch view of part of the buffer (float64)."""
        import torch
        assert self.rank == self.owner
        n = int(np.prod(shape))
        assert byte_offset + 8 * n <= self.nbytes
        return torch.as_tensor(_CudaArray(self.ptr + byte_offset, shape, "<f8"), device="cuda:%d" % self.device)

    def close(self):
        if self.ptr:
            (self._lib.phmm_ipc_close if self._mapped else self._lib.phmm_device_free)(self.ptr)
            self.ptr = 0


class PeerRing:
    """A ring of ``n_buf`` result slots of ``slot_bytes`` on the owner, plus the per-step barrier (see the module docstring)."""

    def __init__(self, slot_bytes, device, rank, world, n_buf=3, owner=0, group=None):
        import torch
        self.slot_bytes = (int(slot_bytes) + 255) // 256 * 256
        self.n_buf, self.rank, self.world, self.owner, self.group = int(n_buf), rank, world, owner, group
        self.buf = PeerBuffer(self.slot_bytes * self.n_buf, device, rank, world, owner, group)
        self._token = torch.zeros(1, dtype=torch.int32, device="cuda:%d" % device)
        self._events = {}

    def window(self, k, byte_offset, rows, cols, ld=None):
        """This step's window: ``byte_offset`` inside slot k % n_buf."""
        return self.buf.view((k % self.n_buf) * self.slot_bytes + byte_offset, rows, cols, ld)

    def owner_slot(self, k, shape, byte_offset=0):
        return self.buf.owner_tensor((k % self.n_buf) * self.slot_bytes + byte_offset, shape)

    def publish(self, k):
        """Call after populate k has returned on this rank. Enqueues the barrier on torch's current stream and returns the event
        that fires when every rank's step-k values are in the owner's slot."""
        import torch
        import torch.distributed as dist
        if self.world > 1:
            dist.all_reduce(self._token, group=self.group)
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        self._events[k] = ev
        self._events.pop(k - 2 * self.n_buf, None)
        return ev

    def wait_slot(self, engine, k):
        """Before populate k: make the engine's stream wait until the slot's previous contents (step k - n_buf) are released."""
        ev = self._events.get(k - self.n_buf + 1)
        if ev is not None:
            engine.wait_event(ev)

    def close(self):
        self.buf.close()
