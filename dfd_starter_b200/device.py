"""Per-device context of the C-ABI library + small torch plumbing helpers
(torch is used for device memory and streams only)."""
import ctypes as C

import torch

from . import _lib

_contexts = {}


class Context(object):
    """One `dfd_ctx` per CUDA device (include/dfd_b200.h: one context per device)."""

    def __init__(self, device_index):
        if not torch.cuda.is_available():
            raise _lib.DfdError("dfd_starter_b200 needs a CUDA (sm_100a) device; there is no CPU fallback")
        self.lib = _lib.load()
        self.device_index = int(device_index)
        self.device = torch.device("cuda", self.device_index)
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.dfd_ctx_create(self.device_index, C.byref(h)), "dfd_ctx_create")
        self.handle = h
        self.sm_count = self.lib.dfd_ctx_sm_count(h)

    @property
    def stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def launch_count(self):
        return int(self.lib.dfd_ctx_launch_count(self.handle))

    def host_stage(self, src, dst):
        """Small transfer between a PINNED host tensor and a device tensor (either direction) with `dfd_host_stage`: a
        kernel accessing the pinned buffer through its device alias.  Stream-ordered like `dst.copy_(src,
        non_blocking=True)`, but it neither queues on a DMA copy engine behind a large upload nor blocks the host."""
        nbytes = src.numel() * src.element_size()
        assert dst.numel() * dst.element_size() == nbytes and (src.is_cuda or src.is_pinned()) and (dst.is_cuda or dst.is_pinned())
        _lib.check(self.lib.dfd_host_stage(self.handle, C.c_void_p(src.data_ptr()), C.c_void_p(dst.data_ptr()), nbytes,
                                           self.stream), "dfd_host_stage")
        return dst

    def zeros_bytes(self, nbytes):
        """Zero-filled, 256-byte aligned scratch (the library's counters self-reset)."""
        return torch.zeros(max(int(nbytes), 256) + 256, dtype=torch.uint8, device=self.device)


def get_context(device=None):
    if device is None:
        idx = torch.cuda.current_device() if torch.cuda.is_available() else 0
    elif isinstance(device, int):
        idx = device
    else:
        d = torch.device(device)
        idx = d.index if d.index is not None else (torch.cuda.current_device() if torch.cuda.is_available() else 0)
    if idx not in _contexts:
        _contexts[idx] = Context(idx)
    return _contexts[idx]


def ptr(t):
    """Device (or pinned-host) pointer of a tensor as c_void_p; None -> NULL."""
    if t is None:
        return C.c_void_p(0)
    return C.c_void_p(t.data_ptr())


def aligned_ptr(t, align=256):
    a = t.data_ptr()
    return C.c_void_p((a + align - 1) // align * align)
