"""Quick timing of the haplotype selection loop (N1) on the synth-select-v1 workload; prints stage times."""
import json
import sys
import time

sys.path.insert(0, ".")
from platypus_b200 import synth  # noqa: E402
from platypus_b200.engine import Engine  # noqa: E402

args = [a for a in sys.argv[1:] if not a.startswith("--")]
W = int(args[0]) if len(args) > 0 else 10000
reps = int(args[1]) if len(args) > 1 else 3
t0 = time.time()
batch, vset = synth.make_select_batch_parallel(W)
print("generated %d windows in %.1f s" % (W, time.time() - t0), flush=True)
if "--pin" in sys.argv:   # page-locked inputs, as bench.py's e2e leg uses
    import torch
    keep = []
    for name in ("hap_seq", "read_seq", "read_qual", "read_pos", "read_end", "read_mapq", "read_qcfail", "read_seq_off",
                 "slot_read", "wi_slot_off", "hap_seq_off"):
        t = torch.from_numpy(getattr(batch, name)).pin_memory()
        keep.append(t)
        setattr(batch, name, t.numpy())
eng = Engine(0)
eng.set_timing(True)
for i in range(reps):
    t0 = time.perf_counter()
    out = eng.select_haplotypes(batch, vset)
    dt = time.perf_counter() - t0
    st = eng.select_stats()
    st["wall_ms"] = dt * 1e3
    st["gcups_wall"] = st["cells"] / dt / 1e9
    st["gcups_score_kernels"] = st["cells"] / ((st["score_ms"] + st["ref_pass_ms"]) * 1e-3) / 1e9
    st["kernel_ms_mean_per_launch"] = eng.kernel_times()
    print(json.dumps(st), flush=True)
print("n_sel", out["n_sel"][:4], "n_scored", out["n_scored"][:4])
