"""Multi-GPU plumbing: one process per GPU, scene replicated, and ONE sum-reduce of the framebuffer to rank 0 at
frame end.  The path has no other exchange step, so torch.distributed (NCCL over NVLink on
the GPU box, gloo in the CPU tests) is used as-is.  Two ways to split a frame:

  rows    interleaved pixel rows (lys_context_set_partition): every rank runs ALL passes for its rows, the other
          pixels stay zero, the sum is exact and the image is bit-identical to the single-GPU one.
  passes  contiguous pass ranges (lys_state_advance_rng): every rank renders the whole frame for its share of the
          passes; the per-rank running averages are combined with weights proportional to the passes they hold.
          Per-GPU work per pass is that of the full frame (best occupancy); results agree with the single-GPU image
          to rounding, not bitwise (the running average is order dependent, integrator.fut:180-192)."""
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


def pass_ranges(total_passes, world):
    """Contiguous split of `total_passes` sample passes over `world` ranks: [(first_pass, count)], counts differ by <= 1.
    Pass k of a frame uses the frame rng advanced k times (integrator.fut:116)."""
    base, extra = divmod(int(total_passes), int(world))
    out, first = [], 0
    for r in range(world):
        n = base + (1 if r < extra else 0)
        out.append((first, n))
        first += n
    return out


def averaged_passes(count):
    """Passes that a `sample_n_frames(count)` image is the mean of: the reference's accumulation scales the first frame
    by (n-1)/n with the OLD n_frames = 1, i.e. drops it (integrator.fut:183-186, lib.fut:67-74), so count - 1 for
    count >= 2 and 1 for count <= 1."""
    return max(int(count) - 1, 1)


def pass_weight(my_count, counts):
    """Weight of this rank's `sample_n_frames(my_count)` image in a pass-split frame whose ranks hold `counts` passes."""
    total = sum(averaged_passes(c) for c in counts if c > 0)
    return averaged_passes(my_count) / total if my_count > 0 else 0.0


def merge_pass_split(buf, my_count, counts, dst=0, weighted=False):
    """Combine per-rank images of a pass-split frame: buf (this rank's `sample_n_frames(my_count)` result) is scaled by its
    share of the averaged passes and sum-reduced to `dst`.  `counts` = the pass counts of all ranks.  weighted=True: the
    library already applied the weight inside its last accumulate kernel (State.sample_n_frames_device(..., weight=w)), so
    only the reduce is left -- no extra pass over the framebuffer."""
    if not weighted:
        buf.mul_(pass_weight(my_count, counts))
    return reduce_framebuffer(buf, dst=dst)


def context_stream(ctx, device):
    """torch view of the stream the context launches on.  The library renders on its own non-blocking stream, which is NOT
    ordered with torch's current stream: every torch / NCCL operation on a library buffer must be issued on this stream (or
    after ctx.sync())."""
    return torch.cuda.ExternalStream(ctx.stream, device=device)


def render_frame(ctx, state, total_passes, mode, rank, world, device, stream=None):
    """One frame of `total_passes` sample passes over `world` ranks (mode 'rows' or 'passes'); returns
    (array handle, torch view of the [h][w][3] f32 device image).  After the call the view holds the full image on
    rank 0.  The caller frees the handle with state.free_f32_3d.  For 'rows' the context must have been given
    ctx.set_partition(rank, world) before the state's passes run.
    Ordering: the reduce is issued on the context's own stream (stream=None builds the torch view of it), i.e. behind the
    passes that produce the image; freeing the handle afterwards is safe because the pool reuses blocks on that same stream."""
    if stream is None:
        stream = context_stream(ctx, device) if torch.device(device).type == 'cuda' else None
    if mode == 'rows':
        hnd, ptr, shape, _ = state.sample_n_frames_device(total_passes, want_stats=False)
        view = as_torch(ptr, shape, device)
        with torch.cuda.stream(stream) if stream is not None else _nullcontext():
            reduce_framebuffer(view, dst=0)
        return hnd, view
    if mode != 'passes':
        raise ValueError("mode must be 'rows' or 'passes'")
    ranges = pass_ranges(total_passes, world)
    first, count = ranges[rank]
    counts = [c for _, c in ranges]
    s = state.advance_rng(first) if first else state
    hnd, ptr, shape, _ = s.sample_n_frames_device(max(count, 1), want_stats=False, weight=pass_weight(count, counts))
    if s is not state:
        s.free()
    view = as_torch(ptr, shape, device)
    with torch.cuda.stream(stream) if stream is not None else _nullcontext():
        merge_pass_split(view, count, counts, dst=0, weighted=True)
    return hnd, view


class _nullcontext:
    def __enter__(self):
        return None

    def __exit__(self, *a):
        return False


class DeviceArray:
    """Zero-copy view of device memory owned by libtracer, exposed through __cuda_array_interface__."""

    def __init__(self, ptr, shape, owner=None):
        self._owner = owner
        self.__cuda_array_interface__ = {'shape': tuple(shape), 'typestr': '<f4', 'data': (int(ptr), False), 'version': 2}


def as_torch(ptr, shape, device, owner=None):
    return torch.as_tensor(DeviceArray(ptr, shape, owner), device=device)
