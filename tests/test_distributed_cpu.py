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


def test_shard_bounds_cover_all_windows():
    from platypus_b200.batch import shard_bounds
    for n in (0, 1, 5, 8, 10001):
        for w in (1, 2, 4, 8):
            b = shard_bounds(n, w)
            assert b[0] == 0 and b[-1] == n and all(0 <= b[i + 1] - b[i] <= (n + w - 1) // w for i in range(w))
