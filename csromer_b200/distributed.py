"""Pixel sharding across the GPUs of one node (one process per GPU, ``torch.distributed``).

Lines of sight are independent (the reference itself forks one process per pixel, ``README.md:309-315``), so the
hot path shards with NO data-path collective: rank r owns a contiguous slab of pixels, holds its own copy of the
plan (twiddle tables) and iterates alone.  Collectives are used for exactly two things (BASELINE.json north_star):
the global convergence norm (one small all-reduce) and gathering the output cube.  The reference computes the
per-LOS norm ``e = sum|x - x_old| / len(x)`` and discards it (``optimization/methods/fista.py:89``).

Works with the ``nccl`` backend on GPUs and with ``gloo`` on CPU tensors (used by the CPU tests).
"""
import numpy as np
import torch
import torch.distributed as dist


def world_info(group=None):
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def shard_bounds(n_items, world, rank):
    """Contiguous, balanced slab [lo, hi) of ``n_items`` for ``rank``; the first ``n_items % world`` ranks get one
    extra item.  Ranks beyond ``n_items`` get an empty slab."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("bad rank %d / world %d" % (rank, world))
    base, extra = divmod(int(n_items), world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_pixels(array, world=None, rank=None, group=None):
    """Slice the trailing pixel axes of a frequency/phi-first array ``(L, *pix)`` for this rank: pixels are
    flattened in C order and cut into contiguous slabs (whole image rows when the row count divides evenly).
    Returns ``(local (L, P_local), (lo, hi))``."""
    if world is None or rank is None:
        rank, world = world_info(group)
    lead = array.shape[0]
    flat = array.reshape(lead, -1)
    lo, hi = shard_bounds(flat.shape[1], world, rank)
    return flat[:, lo:hi], (lo, hi)


def global_convergence_norm(delta_sum, n_los, group=None, device=None):
    """mean over ALL lines of sight of ``sum|x_new - x_old| / len(x)``.  ``delta_sum``: this rank's sum of the
    per-LOS values (tensor or float), ``n_los``: this rank's line-of-sight count.  One 2-float all-reduce."""
    dev = device or (delta_sum.device if isinstance(delta_sum, torch.Tensor) else torch.device("cpu"))
    buf = torch.zeros(2, dtype=torch.float64, device=dev)
    buf[0] = float(delta_sum)
    buf[1] = float(n_los)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    return float(buf[0] / torch.clamp(buf[1], min=1.0))


def max_over_ranks(value, device=None, group=None):
    """max of a scalar over ranks (used for device-timed measurements)"""
    buf = torch.tensor([float(value)], dtype=torch.float64, device=device or "cpu")
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(buf, op=dist.ReduceOp.MAX, group=group)
    return float(buf[0])


def gather_pixels(local, n_total, group=None, dst=None):
    """Inverse of ``shard_pixels``: every rank passes its ``(L, P_local)`` tensor; the full ``(L, n_total)`` tensor
    is returned on every rank (``dst=None``, all-gather) or only on rank ``dst`` (others get None).  Slabs may have
    different sizes, so they are padded to the largest one for the collective."""
    rank, world = world_info(group)
    if world == 1:
        return local
    sizes = [shard_bounds(n_total, world, r) for r in range(world)]
    widest = max(hi - lo for lo, hi in sizes)
    lead = local.shape[0]
    padded = torch.zeros((lead, widest), dtype=local.dtype, device=local.device)
    padded[:, :local.shape[1]] = local
    if dst is None:
        parts = [torch.empty_like(padded) for _ in range(world)]
        dist.all_gather(parts, padded, group=group)
    else:
        parts = [torch.empty_like(padded) for _ in range(world)] if rank == dst else None
        dist.gather(padded, parts, dst=dst, group=group)
        if rank != dst:
            return None
    return torch.cat([p[:, :hi - lo] for p, (lo, hi) in zip(parts, sizes)], dim=1)


def reconstruct_cube(make_problem, run_slab, cube, pixel_chunk=16384, group=None, gather=True):
    """Driver for a whole cube on N ranks (the batched replacement of the README's joblib loop, README.md:282-377).

    ``cube``        complex ``(m, Y, X)`` (or ``(m, P)``) host array / tensor, identical on every rank
    ``make_problem``  callable(data_slab (m, P_chunk)) -> anything ``run_slab`` understands (e.g. builds the Dataset,
                    Parameter, Chi2/L1/FISTA objects for that slab)
    ``run_slab``    callable(problem) -> (x (L, P_chunk) tensor, delta_sum float)
    Every rank walks its own slab in chunks of ``pixel_chunk`` lines of sight; the outputs are gathered (optional)
    and the global convergence norm is all-reduced once at the end.  Returns (x_full or x_local, norm, (lo, hi))."""
    rank, world = world_info(group)
    local, (lo, hi) = shard_pixels(cube, world, rank)
    outs, delta_total = [], 0.0
    for start in range(0, hi - lo, pixel_chunk):
        x, dsum = run_slab(make_problem(local[:, start:start + pixel_chunk]))
        outs.append(x)
        delta_total += float(dsum)
    if outs:
        x_local = torch.cat([torch.as_tensor(o) for o in outs], dim=1)
    else:
        x_local = torch.zeros((0, 0))
    norm = global_convergence_norm(delta_total, hi - lo, group, device=x_local.device if outs else None)
    n_total = int(np.prod(cube.shape[1:]))
    if gather and world > 1:
        lead = torch.tensor([x_local.shape[0]], device=x_local.device)
        dist.all_reduce(lead, op=dist.ReduceOp.MAX, group=group)
        if x_local.shape[0] == 0:  # a rank with an empty slab still has to join the collective
            x_local = torch.zeros((int(lead), 0), dtype=torch.float32, device=lead.device)
        return gather_pixels(x_local, n_total, group), norm, (lo, hi)
    return x_local, norm, (lo, hi)
