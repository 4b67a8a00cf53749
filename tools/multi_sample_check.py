"""Multi-sample sanity check on a GPU box: kernel times of the device-resident path for batches with many
individuals (config-5-like shapes, scaled down).  usage: python tools/multi_sample_check.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from platypus_b200 import synth  # noqa: E402
from platypus_b200.engine import Engine  # noqa: E402

dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(device=dev)
eng = Engine(0, stream=stream.cuda_stream)
for (W, nI, R) in ((10000, 1, 64), (400, 100, 16), (100, 1000, 8), (50, 2000, 8)):
    b = synth.make_batch(W, n_individuals=nI, n_reads=R)
    Hm = b.max_haps()
    Gm = Hm * (Hm + 1) // 2
    V = max(b.max_variants, 1)
    with torch.cuda.stream(stream):
        h = eng.upload(b)
        f64 = dict(dtype=torch.float64, device=dev)
        out = {"gl": torch.zeros((W, nI, Gm), **f64), "gl_log_max": torch.zeros((W, nI), **f64),
               "gof": torch.zeros((W, Gm, nI), **f64), "hap_like": torch.zeros((W, nI, Hm), **f64),
               "freq": torch.zeros((W, Hm), **f64), "em_post": torch.zeros((W, nI, Gm), **f64),
               "call": torch.zeros((W, nI), dtype=torch.int32, device=dev), "var_phred": torch.zeros((W, V), **f64),
               "em_iters": torch.zeros((W,), dtype=torch.int32, device=dev)}
        ptrs = {k: v.data_ptr() for k, v in out.items()}
        ptrs["max_haps"] = Hm
        ll = torch.zeros((int(b.ll_offsets()[-1]),), **f64)
        for _ in range(2):
            eng.run_device(h, ptrs, ll_ptr=ll.data_ptr())
        eng.set_timing(True)
        for _ in range(3):
            eng.run_device(h, ptrs, ll_ptr=ll.data_ptr())
        kt, n = eng.kernel_times()
        eng.set_timing(False)
        st = eng.last_stats()
        tot = sum(kt.values())
        print("W=%d nInd=%d reads/ind=%d: %d pairs, %.2f ms, %.0f GCUPS, em_iters max %d  %s" %
              (W, nI, R, st["n_pairs"], tot, st["cells"] / tot / 1e6, int(out["em_iters"].max()),
               {k: round(v, 3) for k, v in kt.items()}))
        eng.free(h)
