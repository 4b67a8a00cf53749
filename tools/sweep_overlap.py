#!/usr/bin/env python
"""Sweep of the chunk pipeline knobs on the headline workload (one process, one GPU): resident CTAs per SM of the two
persistent kernels (PLB_DP_OCC / PLB_ANCHOR_OCC) x number of pipelined chunks, device-resident (PLB_DEVICE_CHUNKS) and
through the host entry points (PLB_PIPE_CHUNKS).  Prints one line per setting."""
import itertools
import os
import sys
import time

os.environ.setdefault("OMP_WAIT_POLICY", "PASSIVE")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

from platypus_b200 import synth
from platypus_b200.engine import Engine
from platypus_b200.shard import DeviceShard


def pin(a):
    return torch.from_numpy(np.ascontiguousarray(a)).pin_memory()


def main():
    W = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
    what = sys.argv[2] if len(sys.argv) > 2 else "device,host"
    batch = synth.make_batch_parallel(W)
    cells = synth.algorithmic_cells(batch)
    torch.cuda.set_device(0)
    stream = torch.cuda.Stream()
    eng = Engine(0, stream=stream.cuda_stream)
    ref = None
    if "device" in what:
        for dch, docc, aocc, chain in itertools.product((1, 2, 3, 4, 6), (3, 2), (4, 2, 1), (1, 0)):
            if dch == 1 and (chain == 0 or aocc != 4 or docc != 3):
                continue
            if docc == 3 and aocc != 4 and dch > 1:
                continue
            os.environ["PLB_DEVICE_CHUNKS"] = str(dch)
            os.environ["PLB_DP_OCC"] = str(docc)
            os.environ["PLB_ANCHOR_OCC"] = str(aocc)
            if chain:
                os.environ["PLB_ANCHOR_CHAIN"] = "1"
            else:
                os.environ.pop("PLB_ANCHOR_CHAIN", None)
            ds = DeviceShard(eng, batch, stream, gather=False)
            with torch.cuda.stream(stream):
                for _ in range(3):
                    ds.step()
                stream.synchronize()
                eng.set_timing(True)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                for _ in range(10):
                    ds.step()
                e1.record(stream)
                stream.synchronize()
            ms = e0.elapsed_time(e1) / 10
            kt, _ = eng.kernel_times()
            eng.set_timing(False)
            gl = ds.out["gl"].cpu().numpy()
            if ref is None:
                ref = gl
            ok = np.array_equal(ref, gl)
            print("device chunks %d dp_occ %d anchor_occ %d chain %d: %.3f ms  %.0f GCUPS  k_anchor %.2f k_dp %.2f  same=%s" %
                  (dch, docc, aocc, chain, ms, cells / ms / 1e6, kt["k_anchor"], kt["k_dp"], ok), flush=True)
            ds.close()
    if "host" in what:
        p = batch.pack()
        kw = {f: (pin(getattr(p, f)).numpy() if isinstance(getattr(p, f), np.ndarray) else getattr(p, f))
              for f in p.__dataclass_fields__ if f != "_keep"}
        hb = type(p)(**kw)
        outs = []
        for j in range(2):
            o = eng.alloc_population_out(batch)
            outs.append({k: (pin(v).numpy() if isinstance(v, np.ndarray) else v) for k, v in o.items()})
        os.environ.pop("PLB_DEVICE_CHUNKS", None)
        for pch, docc, aocc, chain in itertools.product((1, 2, 3, 4), (3, 2), (4, 2, 1), (1, 0)):
            if docc == 3 and aocc != 4:
                continue
            if pch == 1 and chain == 0:
                continue
            os.environ["PLB_PIPE_CHUNKS"] = str(pch)
            os.environ["PLB_DP_OCC"] = str(docc)
            os.environ["PLB_ANCHOR_OCC"] = str(aocc)
            if chain:
                os.environ["PLB_ANCHOR_CHAIN"] = "1"
            else:
                os.environ.pop("PLB_ANCHOR_CHAIN", None)

            def run(n):
                jobs = []
                for i in range(n):
                    jobs.append(eng.population_submit(hb, out=outs[i % 2]))
                    if len(jobs) == 2:
                        eng.population_wait(jobs.pop(0))
                while jobs:
                    eng.population_wait(jobs.pop(0))
            run(3)
            t0 = time.perf_counter()
            run(12)
            ms = (time.perf_counter() - t0) / 12 * 1e3
            t0 = time.perf_counter()
            eng.population_run(hb, out=outs[0])
            single = (time.perf_counter() - t0) * 1e3
            ok = ref is None or np.array_equal(ref, outs[0]["gl"])
            print("host pipe_chunks %d dp_occ %d anchor_occ %d chain %d: %.3f ms/job  %.0f GCUPS  single call %.2f ms  same=%s" %
                  (pch, docc, aocc, chain, ms, cells / ms / 1e6, single, ok), flush=True)


if __name__ == "__main__":
    main()
