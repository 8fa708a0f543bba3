"""torch-side plumbing shared by the op wrappers: device placement, current stream, scratch."""
import threading

import torch

from . import _cabi

_workspaces = {}


def require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError("detmatch_b200 ops run on a CUDA device only (no CPU fallback) and "
                           "torch.cuda.is_available() is False")


def to_device(t, device=None):
    """Returns (tensor on a CUDA device, original device).  CPU inputs are uploaded to the
    current CUDA device, computed there and the results are brought back by the caller."""
    if t.is_cuda:
        return t, t.device
    require_cuda()
    dev = device if device is not None else torch.device("cuda", torch.cuda.current_device())
    return t.to(dev, non_blocking=False), t.device


def stream_ptr(device):
    return torch.cuda.current_stream(device).cuda_stream


def workspace(device, nbytes):
    """Scratch owned by torch's caching allocator, cached per (device, stream, host thread) and
    grown on demand.  256-byte aligned (the caching allocator aligns to 512).  The thread is part of
    the key because ctypes releases the GIL during the native call: two Python threads enqueueing
    multi-kernel sequences on the same stream must not share prepared-box / bitmap scratch."""
    key = (device.index, torch.cuda.current_stream(device).cuda_stream, threading.get_ident())
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)
        _workspaces[key] = ws
    return ws


def ptr(t):
    return _cabi.c_void_p(t.data_ptr())
