"""Helpers for the -m gpu parity tests: device buffers come from torch (plumbing only), every computation goes
through the C ABI of libhevcasm_b200.so with raw device pointers."""
import ctypes as C

import numpy as np


def torch_cuda():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch


def to_dev(arr):
    torch = torch_cuda()
    return torch.from_numpy(np.ascontiguousarray(arr)).cuda()


def dev_zeros(shape, dtype):
    torch = torch_cuda()
    td = {np.uint8: torch.uint8, np.int16: torch.int16, np.int32: torch.int32}[dtype]
    return torch.zeros(shape, dtype=td, device="cuda")


def dev_full(shape, dtype, value):
    t = dev_zeros(shape, dtype)
    t.fill_(value)
    return t


def dptr(t, offset_elems=0):
    return C.c_void_p(t.data_ptr() + offset_elems * t.element_size())


def to_host(t):
    torch = torch_cuda()
    torch.cuda.synchronize()
    return t.cpu().numpy()
