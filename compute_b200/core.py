"""Queue + dtype plumbing shared by the algorithm mirror."""
from __future__ import annotations

import numpy as np
import torch

DTYPE_NAMES = ["char", "uchar", "short", "ushort", "int", "uint", "long", "ulong", "float", "double"]
OP_NAMES = ["plus", "multiplies", "min", "max", "bit_and", "bit_or", "bit_xor", "minus", "divides"]

_TORCH_CODES = {
    torch.int8: 0, torch.uint8: 1, torch.int16: 2, torch.uint16: 3, torch.int32: 4, torch.uint32: 5,
    torch.int64: 6, torch.uint64: 7, torch.float32: 8, torch.float64: 9,
}
_NP_CODES = {
    np.dtype(np.int8): 0, np.dtype(np.uint8): 1, np.dtype(np.int16): 2, np.dtype(np.uint16): 3,
    np.dtype(np.int32): 4, np.dtype(np.uint32): 5, np.dtype(np.int64): 6, np.dtype(np.uint64): 7,
    np.dtype(np.float32): 8, np.dtype(np.float64): 9,
}
NP_OF_CODE = {v: k for k, v in _NP_CODES.items()}
TORCH_OF_CODE = {v: k for k, v in _TORCH_CODES.items()}


def dtype_code(dt) -> int:
    """types/fundamental.hpp:30-39 scalar -> bcb_dtype."""
    if isinstance(dt, int):
        return dt
    if isinstance(dt, str):
        return DTYPE_NAMES.index(dt)
    if isinstance(dt, torch.dtype):
        return _TORCH_CODES[dt]
    return _NP_CODES[np.dtype(dt)]


def op_code(op) -> int:
    """functional/operator.hpp:73-96 functor -> bcb_op."""
    return op if isinstance(op, int) else OP_NAMES.index(op)


class command_queue:
    """In-order queue = one CUDA stream (command_queue.hpp:78-162).  Algorithms enqueue and return;
    finish() waits (command_queue.hpp:1564-1572)."""

    def __init__(self, stream: "torch.cuda.Stream | None" = None, device: "int | None" = None):
        if not torch.cuda.is_available():
            raise RuntimeError("compute_b200 needs a CUDA device (no CPU fallback)")
        if device is not None:
            torch.cuda.set_device(device)
        self.device = torch.cuda.current_device()
        self.stream = stream

    @property
    def handle(self) -> int:
        s = self.stream if self.stream is not None else torch.cuda.current_stream(self.device)
        return s.cuda_stream

    def finish(self) -> None:
        from ._capi import check, lib
        check(lib().bcb_stream_synchronize(self.handle))


_default_queues = {}


def default_queue() -> command_queue:
    """system::default_queue() (system.hpp:181-184): a queue on the current device's current stream."""
    dev = torch.cuda.current_device()
    q = _default_queues.get(dev)
    if q is None:
        q = _default_queues[dev] = command_queue()
    return q
