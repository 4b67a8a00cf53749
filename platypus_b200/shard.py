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


_DTYPE_CODES = {np.dtype(np.float64): 1, np.dtype(np.int32): 2, np.dtype(np.int64): 3, np.dtype(np.uint64): 4,
                np.dtype(np.float32): 5, np.dtype(np.uint8): 6}
_CODE_DTYPES = {v: k for k, v in _DTYPE_CODES.items()}


def run_sharded(batch: WindowBatch, compute: Callable[[WindowBatch], Dict[str, np.ndarray]],
                keys: Sequence[str] = ("gl",), device: str = "cpu") -> Dict[str, np.ndarray]:
    """Shard `batch` by contiguous window blocks over the ranks of the default process group, run
    `compute` on the local shard and all-gather the requested per-window outputs.

    Ranks may size their blocks differently - with max_haps left to the engine every rank strides its outputs by the
    largest haplotype count of ITS windows - so the per-window block shape is agreed on first (element-wise maximum over
    the ranks, one small all-reduce per key) and every rank zero-pads its block to it; unused genotype / haplotype /
    variant entries are zero in the engine's own layout, so padding changes nothing.  Ranks without windows learn shape
    and dtype from the others."""
    world, rank = dist.get_world_size(), dist.get_rank()
    bounds = shard_bounds(batch.n_windows, world)
    counts = [bounds[r + 1] - bounds[r] for r in range(world)]
    shard = batch.slice_windows(bounds[rank], bounds[rank + 1])
    res = compute(shard) if shard.n_windows else {}
    out = {}
    for k in keys:
        meta = torch.zeros(6, dtype=torch.int64, device=device)     # ndim of the tail, up to 4 tail dims, dtype code
        loc = None
        if shard.n_windows:
            loc = np.ascontiguousarray(res[k])
            tail = loc.shape[1:]
            assert len(tail) <= 4 and loc.dtype in _DTYPE_CODES, (k, loc.shape, loc.dtype)
            meta[0] = len(tail)
            for i, d in enumerate(tail):
                meta[1 + i] = d
            meta[5] = _DTYPE_CODES[loc.dtype]
        dist.all_reduce(meta, op=dist.ReduceOp.MAX)
        m = [int(x) for x in meta.tolist()]
        tail, dtype = tuple(m[1:1 + m[0]]), _CODE_DTYPES.get(m[5], np.dtype(np.float64))
        block = np.zeros((shard.n_windows,) + tail, dtype)
        if loc is not None:
            assert loc.dtype == dtype and loc.ndim == 1 + len(tail), "ranks disagree on dtype / rank of %r" % k
            block[tuple(slice(0, d) for d in loc.shape)] = loc
        as_i64 = dtype == np.dtype(np.uint64)                       # torch has no uint64 collectives
        t = torch.as_tensor(block.view(np.int64) if as_i64 else block, device=device)
        g = gather_blocks(t, counts).cpu().numpy()
        out[k] = g.view(np.uint64) if as_i64 else g
    return out


class DeviceShard:
    """One rank's shard of the windows resident on its GPU, stepped with the path's one collective overlapped
    (what `bench.py --gpus N` times).  step() launches the whole path on the engine's stream and then
    all-gathers the shard's genotype-likelihood block on a side stream behind an event, so the next step's kernels start
    while this step's block travels; the likelihoods alternate between two buffers and step k+2 waits for the gather
    that read its buffer.  The reference's analogue is the per-process VCF merge of src/python/runner.py:470-504."""

    def __init__(self, engine, batch: WindowBatch, stream, opt=None, gather: bool = True):
        self.eng, self.opt, self.stream = engine, opt, stream
        self.world = dist.get_world_size() if (gather and dist.is_initialized()) else 1
        self.rank = dist.get_rank() if self.world > 1 else 0
        dev = torch.device("cuda", engine.device)
        W, nI, Hm = batch.n_windows, batch.n_individuals, batch.max_haps()
        Gm, V = Hm * (Hm + 1) // 2, max(batch.max_variants, 1)
        if self.world > 1:   # every rank must stride by the same haplotype count for one fixed-size gather
            t = torch.tensor([Hm, W], dtype=torch.int64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            assert int(t[1]) == W, "DeviceShard gathers equal-sized shards (use run_sharded for ragged ones)"
            Hm = int(t[0])
            Gm = Hm * (Hm + 1) // 2
        self.shape = (W, nI, Hm, Gm, V)
        with torch.cuda.stream(stream):
            self.handle = engine.upload(batch)
            f64, i32 = torch.float64, torch.int32
            self.out = {"gl": torch.zeros((W, nI, Gm), dtype=f64, device=dev),
                        "gl_log_max": torch.zeros((W, nI), dtype=f64, device=dev),
                        "gof": torch.zeros((W, Gm, nI), dtype=f64, device=dev),
                        "hap_like": torch.zeros((W, nI, Hm), dtype=f64, device=dev),
                        "freq": torch.zeros((W, Hm), dtype=f64, device=dev),
                        "em_post": torch.zeros((W, nI, Gm), dtype=f64, device=dev),
                        "call": torch.zeros((W, nI), dtype=i32, device=dev),
                        "var_phred": torch.zeros((W, V), dtype=f64, device=dev),
                        "em_iters": torch.zeros((W,), dtype=i32, device=dev)}
            self.ll = torch.zeros((int(batch.ll_offsets()[-1]),), dtype=f64, device=dev)
            self.gl_bufs = [self.out["gl"]] + ([torch.zeros_like(self.out["gl"])] if self.world > 1 else [])
            self.gl_all = torch.zeros((self.world, W, nI, Gm), dtype=f64, device=dev) if self.world > 1 else None
        ptrs = {k: v.data_ptr() for k, v in self.out.items()}
        ptrs["max_haps"] = Hm
        self.ptr_sets = [dict(ptrs, gl=g.data_ptr()) for g in self.gl_bufs]
        self.side = torch.cuda.Stream(device=dev) if self.world > 1 else None
        self.computed = [torch.cuda.Event() for _ in self.gl_bufs]
        self.gathered = [torch.cuda.Event() for _ in self.gl_bufs]
        self.n_steps = 0

    def step(self):
        k = self.n_steps % len(self.gl_bufs)
        if self.world > 1 and self.n_steps >= len(self.gl_bufs):
            self.stream.wait_event(self.gathered[k])
        self.eng.run_device(self.handle, self.ptr_sets[k], ll_ptr=self.ll.data_ptr(), opt=self.opt)
        if self.world > 1:
            self.computed[k].record(self.stream)
            self.side.wait_event(self.computed[k])
            with torch.cuda.stream(self.side):
                dist.all_gather_into_tensor(self.gl_all, self.gl_bufs[k])
                self.gathered[k].record(self.side)
        self.n_steps += 1

    def join(self):
        """The engine's stream waits for the gathers issued so far (call before timing ends / reading gl_all)."""
        if self.world > 1:
            self.stream.wait_stream(self.side)

    def last_gl(self):
        return self.gl_bufs[(self.n_steps - 1) % len(self.gl_bufs)]

    def close(self):
        if self.handle is not None:
            self.eng.free(self.handle)
            self.handle = None


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
