"""Device plumbing: PyTorch supplies memory and streams, nothing else.

Every array on the hot path is a contiguous fp64 / int64 ``torch.Tensor`` on a CUDA
device; kernels come from libsmcb.so through ctypes (``_lib``).  There is no CPU
fallback: without a CUDA device these helpers raise.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib

_ctxs = {}


class Context:
    """One libsmcb context (workspace + Philox key) per CUDA device."""

    def __init__(self, device, seed=0):
        self.device = torch.device(device)
        self.lib = _lib.load()
        h = C.c_void_p()
        _lib.check(self.lib.smcb_create(C.byref(h), self.device.index or 0, C.c_uint64(seed)))
        self.handle = h

    def bind_stream(self):
        s = torch.cuda.current_stream(self.device).cuda_stream
        _lib.check(self.lib.smcb_set_stream(self.handle, C.c_void_p(s)))

    def seed(self, seed):
        _lib.check(self.lib.smcb_seed(self.handle, C.c_uint64(int(seed) & (2 ** 64 - 1))))

    @property
    def launches(self):
        return int(self.lib.smcb_launch_count(self.handle))


def require_cuda():
    if not torch.cuda.is_available():
        raise _lib.SmcbError(
            "particles_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")


def context(device=None):
    require_cuda()
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device())
    device = torch.device(device)
    if device.index is None:
        device = torch.device("cuda", torch.cuda.current_device())
    key = device.index
    if key not in _ctxs:
        with torch.cuda.device(device):
            _ctxs[key] = Context(device)
    ctx = _ctxs[key]
    ctx.bind_stream()
    return ctx


def seed(s):
    """Re-key the device generator (the role numpy.random.seed plays in the reference)."""
    context().seed(s)


def as_device(a, dtype=torch.float64, device=None):
    """numpy array / scalar list / tensor -> contiguous CUDA tensor of `dtype`."""
    require_cuda()
    if isinstance(a, torch.Tensor):
        t = a
        if not t.is_cuda:
            t = t.cuda() if device is None else t.to(device)
        if t.dtype != dtype:
            t = t.to(dtype)
        return t.contiguous()
    arr = np.ascontiguousarray(np.asarray(a), dtype={torch.float64: np.float64,
                                                     torch.int64: np.int64}[dtype])
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else device
    return torch.from_numpy(arr).to(dev)


def ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def empty(n, dtype=torch.float64, device=None, like=None):
    if like is not None:
        device = like.device
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device())
    return torch.empty(n, dtype=dtype, device=device)


class _RawDeviceBuffer:
    """Minimal ``__cuda_array_interface__`` carrier: lets torch view device memory it did not allocate
    (peer-mapped arenas from ``smcb_p2p_alloc``).  The owner keeps the allocation alive."""

    def __init__(self, ptr, shape, typestr="<f8", owner=None):
        self.owner = owner          # keeps the allocation's owner alive as long as a tensor views it
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 2, "strides": None}


def tensor_from_ptr(ptr, shape, dtype=torch.float64, owner=None):
    """A torch tensor over raw device memory (no copy); `owner` is referenced by the tensor's base object so
    that the allocation outlives every view handed out."""
    typestr = {torch.float64: "<f8", torch.int64: "<i8"}[dtype]
    return torch.as_tensor(_RawDeviceBuffer(ptr, shape, typestr, owner), device=context().device)
