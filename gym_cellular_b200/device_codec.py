"""Batched mixed-radix codec on the device: the reference's `generalized_cellular2tabular` /
`generalized_tabular2cellular` (gym_cellular/envs/utils/generalized_space_transformations.py:1-12, 15-23)
with their per-cell space list -- any minimum and any length per cell -- over a whole batch.

    encode_mixed(cells, spaces)  int8 [n_cells, n] device tensor -> int64 [n]   (cell 0 least significant)
    decode_mixed(index, spaces)  int [n] device tensor -> int8 [n_cells, n]

`spaces` is what the reference passes as `space_set`: one sequence (usually a `range`) per cell; digit c is
`cells[c] - min(spaces[c])`, radix c is `len(spaces[c])`.  The product of the lengths must fit 32 bits (the
reference works on unbounded Python ints; the host functions of `gym_cellular_b200.codec` keep that).
"""
import ctypes as C

import numpy as np
import torch

from . import _lib


def _space_arrays(spaces):
    radix = np.ascontiguousarray([len(s) for s in spaces], np.int32)
    mins = np.ascontiguousarray([min(s) for s in spaces], np.int32)
    return radix, mins


def _round_up(n, m):
    return (n + m - 1) // m * m


def _stream(device):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def encode_mixed(cells, spaces):
    L = _lib.load()
    cells = torch.as_tensor(cells)
    if not cells.is_cuda:
        raise RuntimeError("encode_mixed needs a CUDA tensor: there is no CPU fallback (use gym_cellular_b200.codec on the host)")
    radix, mins = _space_arrays(spaces)
    n_cells, n = cells.shape
    if n_cells != len(radix):
        raise ValueError("one space per cell")
    ld = _round_up(n, 16)
    buf = torch.zeros(n_cells, ld, dtype=torch.int8, device=cells.device)
    buf[:, :n] = cells.to(torch.int8)
    out = torch.empty(ld, dtype=torch.int32, device=cells.device)
    _lib.check(L.gc_encode_mixed(cells.device.index, n, ld, n_cells, radix.ctypes.data, mins.ctypes.data,
                                 C.c_void_p(buf.data_ptr()), C.c_void_p(out.data_ptr()), _stream(cells.device)))
    return out[:n].to(torch.int64) & 0xFFFFFFFF


def decode_mixed(index, spaces):
    L = _lib.load()
    index = torch.as_tensor(index)
    if not index.is_cuda:
        raise RuntimeError("decode_mixed needs a CUDA tensor: there is no CPU fallback (use gym_cellular_b200.codec on the host)")
    radix, mins = _space_arrays(spaces)
    n = index.shape[0]
    ld = _round_up(n, 16)
    idx = torch.zeros(ld, dtype=torch.int32, device=index.device)
    idx[:n] = (index.to(torch.int64) & 0xFFFFFFFF).to(torch.int32) if index.dtype != torch.int32 else index
    out = torch.empty(len(radix), ld, dtype=torch.int8, device=index.device)
    _lib.check(L.gc_decode_mixed(index.device.index, n, ld, len(radix), radix.ctypes.data, mins.ctypes.data,
                                 C.c_void_p(idx.data_ptr()), C.c_void_p(out.data_ptr()), _stream(index.device)))
    return out[:, :n]
