"""torch views over the device mirrors owned by a TracerContext (plumbing only: torch is used for device
memory, streams and torch.distributed, never for the arithmetic of the path)."""
from __future__ import annotations

import numpy as np
import torch


class _CAI:
    def __init__(self, ptr: int, shape, dtype: np.dtype):
        self.__cuda_array_interface__ = {
            "shape": tuple(int(s) for s in shape),
            "typestr": np.dtype(dtype).str,
            "data": (int(ptr), False),
            "version": 2,
            "strides": None,
        }


def device_view(ptr: int, shape, dtype, device: int = 0) -> torch.Tensor:
    return torch.as_tensor(_CAI(ptr, shape, dtype), device=f"cuda:{device}")


def field_shape(ctx, field: str, nq: int | None = None):
    n, npz, nt = ctx.n, ctx.npz, ctx.nt
    nq = ctx.nq_max if nq is None else nq
    return {
        "q": (nt, nq, npz, n + 6, n + 6),
        "dp1": (nt, npz, n + 6, n + 6),
        "delp": (nt, npz, n + 6, n + 6),
        "cx": (nt, npz, n + 6, n + 1),
        "cy": (nt, npz, n + 1, n + 6),
        "mfx": (nt, npz, n, n + 1),
        "mfy": (nt, npz, n + 1, n),
        "pe": (nt, n + 2, npz + 1, n + 2),
    }[field]


def field_view(ctx, field: str, nq: int | None = None, device: int = 0) -> torch.Tensor:
    """View of a context field (for 'q': ping-pong buffer 0 laid out for `nq` tracers)."""
    return device_view(ctx.device_ptr(field), field_shape(ctx, field, nq), ctx.dtype, device)
