"""world_size-2 gloo test of the multi-GPU host logic (window sharding + the single all-gather of
genotype-likelihood blocks).  The GPU engine cannot run here, so the per-shard compute is
the oracle - which is exactly what the gathered result is checked against on a single process."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_windows, out_path):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    from oracle import oracle as O
    from platypus_b200 import shard, synth
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        batch = synth.make_batch(n_windows, n_haps=4, n_reads=6, read_len=60, hap_len=130)

        def compute(b):
            arrs, _, _, _ = O.population_run(b, max_haps=4)
            return arrs

        got = shard.run_sharded(batch, compute, keys=("gl", "freq", "var_phred"))
        if rank == 0:
            np.savez(out_path, **got)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_windows", [7, 1])
def test_sharded_gather_equals_single_process(tmp_path, n_windows, oracle):
    from platypus_b200 import synth
    out = str(tmp_path / "gathered.npz")
    mp.spawn(_worker, args=(2, _free_port(), n_windows, out), nprocs=2, join=True)
    got = np.load(out)
    batch = synth.make_batch(n_windows, n_haps=4, n_reads=6, read_len=60, hap_len=130)
    want, _, _, _ = oracle.population_run(batch, max_haps=4)
    assert np.array_equal(got["gl"], want["gl"])
    assert np.array_equal(got["freq"], want["freq"])
    nv = want["var_phred"].shape[1]
    assert np.array_equal(got["var_phred"][:, :nv], want["var_phred"])


def _ragged_batch(n_windows):
    """Windows with 2..5 haplotypes (what the selection loop leaves), fewest first, so the ranks' local maxima differ."""
    from platypus_b200 import synth
    from platypus_b200.batch import concat_batches
    parts = [synth.make_batch(1, n_haps=2 + (4 * w) // max(1, n_windows), n_reads=6, read_len=60, hap_len=130, window_offset=w)
             for w in range(n_windows)]
    return concat_batches(parts)


def _ragged_worker(rank, world, port, n_windows, out_path):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    from oracle import oracle as O
    from platypus_b200 import shard
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        batch = _ragged_batch(n_windows)

        def compute(b):   # like Engine.population_run(b): strides by the largest haplotype count of THIS shard
            arrs, _, _, _ = O.population_run(b)
            return arrs

        got = shard.run_sharded(batch, compute, keys=("gl", "gof", "freq", "call", "em_iters"))
        if rank == 0:
            np.savez(out_path, **got)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_windows", [9, 1])
def test_sharded_gather_ragged_haplotype_counts(tmp_path, n_windows, oracle):
    """Ranks whose windows have different haplotype counts stride their blocks differently; run_sharded pads to the
    common shape, keeps integer outputs integer, and an empty rank (1 window on 2 ranks) adopts shape and dtype."""
    out = str(tmp_path / "ragged.npz")
    mp.spawn(_ragged_worker, args=(2, _free_port(), n_windows, out), nprocs=2, join=True)
    got = np.load(out)
    batch = _ragged_batch(n_windows)
    want, _, _, _ = oracle.population_run(batch)
    for k in ("gl", "gof", "freq", "call", "em_iters"):
        assert got[k].dtype == want[k].dtype and np.array_equal(got[k], want[k]), k


def test_shard_bounds_cover_all_windows():
    from platypus_b200.batch import shard_bounds
    for n in (0, 1, 5, 8, 10001):
        for w in (1, 2, 4, 8):
            b = shard_bounds(n, w)
            assert b[0] == 0 and b[-1] == n and all(0 <= b[i + 1] - b[i] <= (n + w - 1) // w for i in range(w))


def _select_compute(b, v):
    """Stand-in for Engine.select_haplotypes on a shard: the Python oracle, packed like PlbSelectOut."""
    from oracle import select_oracle as S
    ms = 16
    out = {"n_sel": np.zeros(b.n_windows, np.int32), "n_scored": np.zeros(b.n_windows, np.int32),
           "sel_mask": np.zeros((b.n_windows, ms), np.uint64), "sel_score": np.full((b.n_windows, ms), np.nan)}
    for w in range(b.n_windows):
        tr = []
        got = S.select_haplotypes(S.window_from_batch(b, v, w), 9, 9, 8, 1, 30, trace=tr)
        out["n_sel"][w] = len(got)
        out["n_scored"][w] = sum(len(r) for r in tr)
        for j, (s_, sc) in enumerate(got):
            out["sel_mask"][w, j] = sum(1 << i for i in s_)
            out["sel_score"][w, j] = np.nan if sc is None else sc
    return out


def _select_worker(rank, world, port, n_windows, out_path):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    from platypus_b200 import shard, synth
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        b, v = synth.make_select_batch(n_windows, n_vars=5, n_reads=10, read_len=60, hap_len=160)
        got = shard.run_select_sharded(b, v, _select_compute)
        if rank == 0:
            np.savez(out_path, **got)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_windows", [5, 1])
def test_sharded_selection_equals_single_process(tmp_path, n_windows, oracle):
    """N1 over two ranks (window blocks + all-gather of masks / scores) = the same windows on one process."""
    from platypus_b200 import synth
    out = str(tmp_path / "sel.npz")
    mp.spawn(_select_worker, args=(2, _free_port(), n_windows, out), nprocs=2, join=True)
    got = np.load(out)
    b, v = synth.make_select_batch(n_windows, n_vars=5, n_reads=10, read_len=60, hap_len=160)
    want = _select_compute(b, v)
    assert np.array_equal(got["n_sel"], want["n_sel"]) and np.all(want["n_sel"] == 8)
    assert np.array_equal(got["n_scored"], want["n_scored"])
    assert np.array_equal(got["sel_mask"], want["sel_mask"])
    assert np.array_equal(np.nan_to_num(got["sel_score"]), np.nan_to_num(want["sel_score"]))
