"""Multi-GPU plumbing: one process per GPU, scene replicated, pixels split by interleaved rows, and ONE
sum-reduce of the framebuffer to rank 0 at frame end.  The path has no other exchange
step, so torch.distributed (NCCL over NVLink on the GPU box, gloo in the CPU tests) is used as-is."""
import numpy as np
import torch
import torch.distributed as dist


def owned_rows(grid_h, rank, world):
    """Grid rows sampled by `rank`: r % world == rank (matches lys_context_set_partition / local_to_pixel)."""
    return np.arange(rank, grid_h, world)


def local_pixel_count(grid_h, grid_w, rank, world):
    return ((grid_h - rank + world - 1) // world) * grid_w if grid_h > rank else 0


def reduce_framebuffer(buf, dst=0):
    """Sum-reduce a [h][w][3] f32 framebuffer to `dst`.  Every pixel is non-zero on exactly one rank, so the
    sum is exact and the result equals the single-GPU image bit-for-bit."""
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.reduce(buf, dst=dst, op=dist.ReduceOp.SUM)
    return buf


class DeviceArray:
    """Zero-copy view of device memory owned by libtracer, exposed through __cuda_array_interface__."""

    def __init__(self, ptr, shape, owner=None):
        self._owner = owner
        self.__cuda_array_interface__ = {'shape': tuple(shape), 'typestr': '<f4', 'data': (int(ptr), False), 'version': 2}


def as_torch(ptr, shape, device, owner=None):
    return torch.as_tensor(DeviceArray(ptr, shape, owner), device=device)
