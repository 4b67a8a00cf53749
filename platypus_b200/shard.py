"""
Multi-GPU driver for the window path.

Windows are independent units (each callVariantsInWindow touches only its own haplotypes and
reads; the reference deals regions round-robin to processes, src/python/runner.py:470-474, and
heap-merges their VCFs, runner.py:301-352).  Here every rank takes one contiguous block of
windows, so the merged result is a plain concatenation in window order, and the only collective of
the path is ONE all-gather of the per-window genotype-likelihood blocks.

`compute` is any callable WindowBatch -> dict with at least "gl" [W_shard, nInd, Gmax]; in production
it is the CUDA engine, the CPU tests pass a stand-in.  Nothing here computes likelihoods.
"""
from typing import Callable, Dict, Sequence

import numpy as np
import torch
import torch.distributed as dist

from .batch import WindowBatch, shard_bounds


def gather_blocks(local: torch.Tensor, counts: Sequence[int]) -> torch.Tensor:
    """All-gather per-window blocks whose first dimension differs per rank (counts[r] windows on
    rank r).  One collective: shards are padded to the largest shard."""
    world = dist.get_world_size()
    assert len(counts) == world and local.shape[0] == counts[dist.get_rank()]
    wmax = max(counts)
    pad = torch.zeros((wmax,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[:local.shape[0]] = local
    out = torch.empty((world * wmax,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, pad)
    out = out.view((world, wmax) + tuple(local.shape[1:]))
    return torch.cat([out[r, :counts[r]] for r in range(world)], dim=0)


def run_sharded(batch: WindowBatch, compute: Callable[[WindowBatch], Dict[str, np.ndarray]],
                keys: Sequence[str] = ("gl",), device: str = "cpu") -> Dict[str, np.ndarray]:
    """Shard `batch` by contiguous window blocks over the ranks of the default process group, run
    `compute` on the local shard and all-gather the requested per-window outputs."""
    world, rank = dist.get_world_size(), dist.get_rank()
    bounds = shard_bounds(batch.n_windows, world)
    counts = [bounds[r + 1] - bounds[r] for r in range(world)]
    shard = batch.slice_windows(bounds[rank], bounds[rank + 1])
    res = compute(shard) if shard.n_windows else {}
    out = {}
    for k in keys:
        if shard.n_windows:
            loc = torch.as_tensor(np.ascontiguousarray(res[k]), device=device)
            tail = tuple(loc.shape[1:])
            meta = torch.tensor(list(tail) + [0] * (4 - len(tail)), dtype=torch.int64, device=device)
        else:
            loc, meta = None, torch.zeros(4, dtype=torch.int64, device=device)
        # ranks without windows learn the block shape from the others
        dist.all_reduce(meta, op=dist.ReduceOp.MAX)
        if loc is None:
            tail = tuple(int(x) for x in meta.tolist() if x > 0)
            loc = torch.zeros((0,) + tail, dtype=torch.float64, device=device)
        out[k] = gather_blocks(loc, counts).cpu().numpy()
    return out


def run_select_sharded(ref_batch: WindowBatch, variants, compute, device: str = "cpu") -> Dict[str, np.ndarray]:
    """The haplotype selection loop over the ranks: windows are independent here too
    (getFilteredHaplotypes sees one window's variants and reads), so every rank runs `compute(ref_shard, variant_shard)`
    -> {"n_sel", "sel_mask", "sel_score", "n_scored"} on its contiguous block of windows and the per-window results are
    all-gathered in window order (masks travel as int64 bit patterns)."""
    world, rank = dist.get_world_size(), dist.get_rank()
    bounds = shard_bounds(ref_batch.n_windows, world)
    counts = [bounds[r + 1] - bounds[r] for r in range(world)]
    lo, hi = bounds[rank], bounds[rank + 1]
    res = compute(ref_batch.slice_windows(lo, hi), variants.slice_windows(lo, hi)) if hi > lo else None
    width = torch.tensor([res["sel_mask"].shape[1] if res is not None else 0], dtype=torch.int64, device=device)
    dist.all_reduce(width, op=dist.ReduceOp.MAX)      # ranks may have sized max_sel differently
    ms = int(width.item())

    def block(name, dtype, wide):
        shape = (hi - lo, ms) if wide else (hi - lo,)
        a = np.zeros(shape, dtype)
        if res is not None:
            src = res[name]
            if wide:
                a[:, :src.shape[1]] = src
            else:
                a[:] = src
        return a
    out = {}
    for name, dtype, wide in (("n_sel", np.int32, False), ("n_scored", np.int32, False), ("sel_mask", np.uint64, True),
                              ("sel_score", np.float64, True)):
        a = block(name, dtype, wide)
        t = torch.as_tensor(a.view(np.int64) if dtype == np.uint64 else a, device=device)
        g = gather_blocks(t, counts).cpu().numpy()
        out[name] = g.view(np.uint64) if dtype == np.uint64 else g
    out["max_sel"] = ms
    return out
