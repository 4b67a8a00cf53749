#!/usr/bin/env python
"""
bench.py — headline benchmark of the read-vs-haplotype likelihood path (BASELINE.json).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Metric: pair-HMM GCUPS = 10^9 algorithmic band cells per second, 16 * readLen cells per scored
(read, haplotype) pair.  Workload at every N: BASELINE config 2 per GPU (weak
scaling) — synth-v1, 10,000 windows x 8 haplotypes x 64 reads, 150 bp reads x 250 bp haplotypes.
A "step" is one pass of the whole path (anchor voting, band alignment, LL, genotype likelihoods,
EM, posteriors) over that batch.

  value     device-resident pass: inputs already in HBM, timed with CUDA events on the launch stream
  e2e       the same pass through the C-ABI host entry point plb_population_run_host with pinned HOST
            buffers: H2D of all inputs and D2H of all population outputs inside the timed region
  roofline  k_dp (the dominant kernel): algorithmic bytes per launch / mean launch duration measured
            with CUDA events inside the timed region, against MEASURED_PEAKS.json hbm_gbs
  cpu_baseline  the REFERENCE'S OWN CODE on the host cores: its Haplotype / DiploidGenotype / Population classes
            (chaplotype.pyx, cgenotype.pyx, cpopulation.pyx, calign.pyx, align.c with traceback, built for Python 3 into
            oracle/_ref by oracle/build.py) over the whole workload, one process per core with windows dealt to them
            the way runner.py:470-485 deals regions to its processes (kind "reference").  If those modules are missing
            it falls back to the oracle's C restatement with the reference's align.c, then to the pure restatement.

--impl reference times that CPU path alone (the reference has no GPU implementation).
Nothing here reads /root/reference at run time.
"""
import argparse
import json
import os

# with several ranks on the box, the library's host threads (tile planning) sleep between calls instead of spinning next
# to the other ranks' threads; libgomp reads this when it is loaded, i.e. before torch / numpy are imported.  A single
# rank keeps libgomp's default (spin briefly, then sleep): the selection loop has many short parallel regions per call.
if int(os.environ.get("LOCAL_WORLD_SIZE", "1")) > 1:
    os.environ.setdefault("OMP_WAIT_POLICY", "PASSIVE")
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

WINDOWS_PER_GPU = 10000
N_HAPS, N_READS, READ_LEN, HAP_LEN = 8, 64, 150, 250
CPU_SAMPLE_WINDOWS = 10000   # the whole config-2 workload: ~0.8 s on 16 threads (~13 s of CPU work) per step
METRIC = "pair_hmm_gcups"
UNIT = "GCUPS"


MODE = "default"


def workload_config(n_gpus, windows):
    cfg = _workload_config(n_gpus, windows)
    if MODE != "default":
        cfg["mode"] = {"flank": "--calculateFlankScore=1", "hla": "--HLATyping=1"}[MODE] + " (every alignment on the scalar path; not the headline)"
    return cfg


def _workload_config(n_gpus, windows):
    if CONFIG == 4:
        return {"workload": "synth-v1 config4 (chr22-sized: ~500k windows over 8 GPUs): %d windows x %d haplotypes x %d reads per GPU, "
                            "%d bp reads x %d bp haplotypes; all-gather of gl[W][36] f64 (%.0f MB over %d ranks)"
                            % (windows, N_HAPS, N_READS, READ_LEN, HAP_LEN, windows * n_gpus * 36 * 8 / 1e6, n_gpus),
                "windows_per_gpu": windows, "n_gpus": n_gpus, "haplotypes": N_HAPS, "reads": N_READS, "read_len": READ_LEN,
                "hap_len": HAP_LEN, "individuals": 1, "cells_per_pair": "16*readLen (algorithmic band cells, SURVEY 8d)",
                "l2": "inputs of one step (~1.4 GB per GPU) exceed the 126 MB L2; no explicit flush",
                "parallelism": "contiguous blocks of windows per GPU, one process per GPU, one NCCL all-gather per step"}
    if RAGGED:
        return {"workload": "synth-v1 config3 shapes: %d windows x %d haplotypes x %d reads per GPU, reads 100-250 bp, "
                            "haplotypes 200-500 bp (profiling configuration, not the headline)" % (windows, N_HAPS, N_READS),
                "windows_per_gpu": windows, "n_gpus": n_gpus}
    return {
        "workload": "synth-v1 config2: %d windows x %d haplotypes x %d reads per GPU, %d bp reads x %d bp haplotypes"
                    % (windows, N_HAPS, N_READS, READ_LEN, HAP_LEN),
        "windows_per_gpu": windows, "n_gpus": n_gpus, "haplotypes": N_HAPS, "reads": N_READS,
        "read_len": READ_LEN, "hap_len": HAP_LEN, "individuals": 1,
        "cells_per_pair": "16*readLen (algorithmic band cells, SURVEY 8d)",
        "l2": "inputs+outputs of one step (~310 MB per GPU) exceed the 126 MB L2; no explicit flush",
        "parallelism": "windows sharded over %d GPU(s), one process per GPU%s" %
                       (n_gpus, ", NCCL all-gather of genotype likelihoods" if n_gpus > 1 else ""),
    }


OPT = None       # --mode flank / hla: PlbOptions with the run-time mode switched on
CONFIG = 2
RAGGED = False   # --config 3: read length U{100..250}, haplotype length U[max(200, Lmax+16), 500]


def make_workload(rank, windows):
    from platypus_b200 import synth
    t0 = time.time()
    kw = dict(read_len_range=(100, 250), hap_len_range=(200, 500)) if RAGGED else dict(read_len=READ_LEN, hap_len=HAP_LEN)
    procs = max(1, min(32, (os.cpu_count() or 1) // max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1")))))
    b = synth.make_batch_parallel(windows, window_offset=rank * windows, n_procs=procs, n_haps=N_HAPS, n_reads=N_READS, **kw)
    return b, time.time() - t0


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
                power.append(float(f[3]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


_REF_ARGS = None    # per-window argument tuples of l3_ref_wrap.population_seq, inherited by the forked workers


def _window_args(batch, n_windows):
    """Windows of a single-individual batch as the Python objects the reference's classes take."""
    b = batch
    out = []
    for w in range(min(n_windows, b.n_windows)):
        h0, h1 = int(b.win_hap_off[w]), int(b.win_hap_off[w + 1])
        haps = [b.hap_seq[b.hap_seq_off[h]:b.hap_seq_off[h + 1]].tobytes() for h in range(h0, h1)]
        s0, s1 = int(b.wi_slot_off[w]), int(b.wi_slot_off[w + 1])
        ng, nb = int(b.wi_n_good[w]), int(b.wi_n_bad[w])
        reads = []
        for s_ in range(s0, s1):
            r = int(b.slot_read[s_])
            o0, o1 = int(b.read_seq_off[r]), int(b.read_seq_off[r + 1])
            reads.append((b.read_seq[o0:o1].tobytes(), b.read_qual[o0:o1].tobytes(), int(b.read_pos[r]), int(b.read_end[r]),
                          int(b.read_mapq[r]), 512 if b.read_qcfail[r] else 0))
        out.append((haps, int(b.win_start[w]), int(b.win_end[w]), int(b.hap_start[w]),
                    [(reads[:ng], reads[ng:ng + nb], reads[ng + nb:])]))
    return out


def _ref_worker(span):
    from oracle import oracle as O
    W = O.ref_l3()
    acc = 0.0
    for w in range(span[0], span[1]):
        acc += W.population_seq(*_REF_ARGS[w])["gl_log_max"][0]
    return acc


def true_reference_run(batch, n_windows, procs, steps=1, warmup=0):
    """Times the reference's own classes (oracle/_ref/l3_ref_wrap) over the first n_windows windows with `procs`
    processes.  Returns (gcups, seconds per step, cells, windows) or None when the modules are not there or the
    workload cannot be expressed (several individuals, flank not of the form min(2*rlen, 500))."""
    global _REF_ARGS
    import multiprocessing as mp
    from oracle import oracle as O
    from platypus_b200 import synth
    if O.ref_l3() is None or batch.n_individuals != 1:
        return None
    sub = batch.slice_windows(0, min(n_windows, batch.n_windows))
    try:
        _REF_ARGS = _window_args(sub, sub.n_windows)
        O.ref_l3().population_seq(*_REF_ARGS[0])
    except Exception:
        return None
    nw = len(_REF_ARGS)
    cells = synth.algorithmic_cells(sub)
    procs = max(1, min(procs, nw))
    n_chunks = procs * 8     # round-robin-sized pieces: the pool balances them
    spans = [(nw * i // n_chunks, nw * (i + 1) // n_chunks) for i in range(n_chunks)]
    spans = [sp for sp in spans if sp[1] > sp[0]]
    if procs == 1:
        for _ in range(warmup):
            _ref_worker((0, min(nw, 16)))
        t0 = time.perf_counter()
        for _ in range(max(1, steps)):
            _ref_worker((0, nw))
        dt = (time.perf_counter() - t0) / max(1, steps)
        return cells / dt / 1e9, dt, cells, nw
    with mp.get_context("fork").Pool(procs) as pool:
        pool.map(_ref_worker, [(i, i + 1) for i in range(min(nw, procs))])     # workers up, modules imported
        for _ in range(warmup):
            pool.map(_ref_worker, spans)
        t0 = time.perf_counter()
        for _ in range(max(1, steps)):
            pool.map(_ref_worker, spans)
        dt = (time.perf_counter() - t0) / max(1, steps)
    return cells / dt / 1e9, dt, cells, nw


def cpu_reference_run(batch, n_windows, threads, steps=1, warmup=0):
    """CPU path over the first n_windows windows.  Preferred: the reference's own classes (true_reference_run);
    otherwise the oracle (anchoring/LL/GL/EM restated in C) with the reference's align.c when oracle/_ref has it.
    Returns (gcups, kind, seconds per step, cells, windows, description)."""
    from oracle import oracle as O
    from platypus_b200 import synth
    if not RAGGED:
        r = true_reference_run(batch, n_windows, threads, steps, warmup)
        if r is not None:
            return r[0], "reference", r[1], r[2], r[3], ("the reference's own Haplotype / DiploidGenotype / Population classes "
                                                        "(chaplotype.pyx, cgenotype.pyx, cpopulation.pyx, calign.pyx, align.c with "
                                                        "traceback; oracle/_ref), %d process%s" % (threads, "" if threads == 1 else "es"))
    sub = batch.slice_windows(0, min(n_windows, batch.n_windows))
    kind = "reference" if O.use_reference_kernel(True, traceback=True) else "port"
    cells = synth.algorithmic_cells(sub)
    for _ in range(warmup):
        O.population_run(sub, n_threads=threads, want_ll=False)
    t0 = time.perf_counter()
    for _ in range(max(1, steps)):
        O.population_run(sub, n_threads=threads, want_ll=False)
    dt = (time.perf_counter() - t0) / max(1, steps)
    O.use_reference_kernel(False)
    desc = ("oracle restatement of anchoring / LL / GL / EM in C with %s, %d OpenMP threads" %
            ("the unmodified reference align.c (with traceback) for every band alignment" if kind == "reference"
             else "its own band alignment", threads))
    return cells / dt / 1e9, kind, dt, cells, sub.n_windows, desc


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    windows = args.windows
    batch, _ = make_workload(0, min(windows, CPU_SAMPLE_WINDOWS))
    cores = args.cpu_procs or (os.cpu_count() or 1)
    gcups, kind, dt, cells, nw, desc = cpu_reference_run(batch, args.cpu_windows or CPU_SAMPLE_WINDOWS, cores,
                                                         steps=args.steps, warmup=args.warmup)
    sample = "first %d windows of the workload per step (%d pairs, %.3g cells); %s" % (nw, nw * N_HAPS * N_READS, cells, desc)
    line = {
        "impl": "reference", "metric": METRIC, "value": gcups, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "int16 scores / f64 likelihoods", "data": "synthetic (synth-v1, seed 20261017)",
        "config": workload_config(args.gpus, windows),
        "cpu_baseline": {"value": gcups, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": gcups, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---- --stage select: the haplotype selection loop, the step before the window model ----------------

SEL_VARS = 8
_SEL_ARGS = None


def select_config(windows):
    return {"workload": "synth-select-v1: %d windows x %d candidate variants x %d reads, %d bp reads x %d bp reference segment; "
                        "default options (maxHaplotypes 50, coverageSamplingLevel 30): 163 trial haplotypes per window in 8 "
                        "rounds, every 6th read sampled" % (windows, SEL_VARS, N_READS, READ_LEN, HAP_LEN),
            "stage": "select (getFilteredHaplotypes; not the headline)", "windows_per_gpu": windows, "n_gpus": 1,
            "cells_per_pair": "16*readLen per (sampled read, trial haplotype) pair + one reference-haplotype pass",
            "l2": "per-round working set (reads + trial haplotypes + LL) exceeds the 126 MB L2; no explicit flush"}


def _select_window_args(batch, vset, n_windows):
    """Windows of the select workload as the arguments of l3_ref_wrap.select_haplotypes (coordinates moved next to the
    origin; the reference segment becomes the in-memory genome)."""
    out = []
    for w in range(min(n_windows, batch.n_windows)):
        hs = int(batch.hap_start[w])
        shift = hs - 8
        ref = batch.hap_seq[batch.hap_seq_off[w]:batch.hap_seq_off[w + 1]].tobytes()
        genome = b"N" * 8 + ref + b"N"
        vs = []
        for i in range(int(vset.win_var_off[w]), int(vset.win_var_off[w + 1])):
            p, nr = int(vset.var_pos[i]) - shift, int(vset.var_n_removed[i])
            add = vset.var_added[int(vset.var_added_off[i]):int(vset.var_added_off[i + 1])].tobytes()
            rem = genome[p:p + nr] if len(add) == nr else genome[p + 1:p + 1 + nr]
            vs.append((p, rem, add, int(vset.var_n_support[i])))
        per_ind = []
        for i in range(batch.n_individuals):
            wi = w * batch.n_individuals + i
            s0 = int(batch.wi_slot_off[wi])
            reads = []
            for s_ in range(s0, s0 + int(batch.wi_n_good[wi])):
                r = int(batch.slot_read[s_])
                o0, o1 = int(batch.read_seq_off[r]), int(batch.read_seq_off[r + 1])
                reads.append((batch.read_seq[o0:o1].tobytes(), batch.read_qual[o0:o1].tobytes(), int(batch.read_pos[r]) - shift,
                              int(batch.read_end[r]) - shift, int(batch.read_mapq[r]), 0))
            per_ind.append(reads)
        flank = int(batch.win_start[w]) - hs
        out.append((genome, int(batch.win_start[w]) - shift, int(batch.win_end[w]) - shift, vs, per_ind, flank // 2))
    return out


def _sel_worker(span):
    from oracle import oracle as O
    W = O.ref_l3()
    n = 0
    for w in range(span[0], span[1]):
        n += len(W.select_haplotypes(*_SEL_ARGS[w])["selected"])
    return n


_SEL_PORT = None


def _sel_port_worker(span):
    from oracle import select_oracle as S
    b, v = _SEL_PORT
    n = 0
    for w in range(span[0], span[1]):
        n += len(S.select_haplotypes(S.window_from_batch(b, v, w)))
    return n


def select_port_run(batch, vset, n_windows, procs, steps=1, warmup=0):
    """Fallback CPU arm when oracle/_ref is absent: the oracle's restatement of the loop (oracle/select_oracle.py over the C
    oracle's per-read scoring), one process per core.  Returns (seconds per step, windows)."""
    global _SEL_PORT
    import multiprocessing as mp
    nw = min(n_windows, batch.n_windows)
    _SEL_PORT = (batch, vset)
    procs = max(1, min(procs, nw))
    spans = [(nw * i // (procs * 4), nw * (i + 1) // (procs * 4)) for i in range(procs * 4)]
    spans = [sp for sp in spans if sp[1] > sp[0]]
    with mp.get_context("fork").Pool(procs) as pool:
        pool.map(_sel_port_worker, [(i, i + 1) for i in range(min(nw, procs))])
        for _ in range(warmup):
            pool.map(_sel_port_worker, spans)
        t0 = time.perf_counter()
        for _ in range(max(1, steps)):
            pool.map(_sel_port_worker, spans)
        dt = (time.perf_counter() - t0) / max(1, steps)
    return dt, nw


def select_reference_run(batch, vset, n_windows, procs, steps=1, warmup=0):
    """The reference's own getFilteredHaplotypes / computeBestScoreForGenotype / Haplotype classes (oracle/_ref/n1_ref +
    l3_ref_wrap) over the first n_windows windows, one process per core.  Returns (seconds per step, windows) or None."""
    global _SEL_ARGS
    import multiprocessing as mp
    from oracle import oracle as O
    if O.ref_l3() is None or not hasattr(O.ref_l3(), "select_haplotypes"):
        return None
    _SEL_ARGS = _select_window_args(batch, vset, n_windows)
    nw = len(_SEL_ARGS)
    procs = max(1, min(procs, nw))
    spans = [(nw * i // (procs * 4), nw * (i + 1) // (procs * 4)) for i in range(procs * 4)]
    spans = [sp for sp in spans if sp[1] > sp[0]]
    with mp.get_context("fork").Pool(procs) as pool:
        pool.map(_sel_worker, [(i, i + 1) for i in range(min(nw, procs))])
        for _ in range(warmup):
            pool.map(_sel_worker, spans)
        t0 = time.perf_counter()
        for _ in range(max(1, steps)):
            pool.map(_sel_worker, spans)
        dt = (time.perf_counter() - t0) / max(1, steps)
    return dt, nw


def run_select(args):
    """bench.py --stage select: plb_select_haplotypes_host on the synth-select-v1 workload (one GPU)."""
    import torch
    from platypus_b200 import synth
    from platypus_b200.engine import Engine
    W = args.windows
    batch, vset = synth.make_select_batch_parallel(W, n_vars=SEL_VARS, n_reads=N_READS, read_len=READ_LEN, hap_len=HAP_LEN)
    if args.impl == "reference":
        cores = args.cpu_procs or (os.cpu_count() or 1)
        nw = args.cpu_windows or min(W, 40 * cores)
        r = select_reference_run(batch, vset, nw, cores, steps=args.steps, warmup=min(args.warmup, 1))
        kind = "reference"
        if r is None:   # oracle/_ref not built: the oracle's restatement of the loop
            kind = "port"
            r = select_port_run(batch, vset, min(nw, 8 * cores), cores, steps=args.steps, warmup=min(args.warmup, 1))
        dt, nw = r
        cells = SEL_CELLS_PER_WINDOW * nw
        g = cells / dt / 1e9
        sample = ("first %d windows per step; %s, %d processes; cells counted as for "
                  "the GPU arm (the reference re-scores the reference haplotype for every trial: not credited)" %
                  (nw, "the reference's own getFilteredHaplotypes / computeBestScoreForGenotype / Haplotype classes (variantFilter.pyx, "
                       "chaplotype.pyx, calign.pyx, align.c; oracle/_ref)" if kind == "reference" else
                       "the oracle's restatement of the loop (oracle/select_oracle.py + the C oracle's per-read scoring)", cores))
        print(json.dumps({"impl": "reference", "metric": METRIC, "value": g, "unit": UNIT, "n_gpus": 1, "steps": args.steps,
                          "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
                          "vs_baseline": None, "dtype": "int16 scores / f64 likelihoods", "data": "synthetic (synth-select-v1)",
                          "config": select_config(W), "windows_per_s": nw / dt,
                          "cpu_baseline": {"value": g, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
                          "e2e": {"value": g, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}),
              flush=True)
        return
    keep = []
    for name in ("hap_seq", "read_seq", "read_qual", "read_pos", "read_end", "read_mapq", "read_qcfail", "read_seq_off",
                 "slot_read", "wi_slot_off", "wi_n_good", "hap_seq_off", "win_start", "win_end", "hap_start"):
        t = torch.from_numpy(np.ascontiguousarray(getattr(batch, name))).pin_memory()
        keep.append(t)
        setattr(batch, name, t.numpy())
    torch.cuda.set_device(0)
    eng = Engine(0)
    eng.set_timing(True)
    out = None
    for _ in range(max(3, args.warmup)):
        out = eng.select_haplotypes(batch, vset, max_sel=64, out=out)
    sampler = ClockSampler(0)
    sampler.start()
    l0 = eng.launch_count
    eng.set_timing(True)
    wall, dev = [], []
    for _ in range(args.steps):
        t0 = time.perf_counter()
        out = eng.select_haplotypes(batch, vset, max_sel=64, out=out)
        wall.append(time.perf_counter() - t0)
        st = eng.select_stats()
        dev.append((st["ref_pass_ms"] + st["build_ms"] + st["score_ms"] + st["reduce_ms"]) * 1e-3)
    launches = eng.launch_count - l0
    clocks = sampler.stop()
    kt, n_runs = eng.kernel_times()
    st = eng.select_stats()
    cells = st["cells"]
    t_wall, t_dev = sum(wall) / len(wall), sum(dev) / len(dev)
    assert abs(cells / W - SEL_CELLS_PER_WINDOW) < 1e-6 * SEL_CELLS_PER_WINDOW, (cells / W, SEL_CELLS_PER_WINDOW)
    # algorithmic bytes of the scoring rounds (SURVEY 8d layout): per pair-producing launch the sampled reads, the trial
    # haplotypes and one f64 log-likelihood per pair
    reads_per_w = st["n_pairs"] / (st["n_trials"] + W)           # sampled reads per window
    rd_bytes = (READ_LEN + 3) // 4 + READ_LEN + 8
    alg_bytes = (SEL_VARS + 1) * W * reads_per_w * rd_bytes + (st["n_trials"] + W) * ((HAP_LEN + 3) // 4 + 8) + 8 * st["n_pairs"]
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peak, peak_src = 6650.0, "B200_PROFILING.md fallback"
    if os.path.exists(peaks_path):
        try:
            peak = float(json.load(open(peaks_path))["hbm_gbs"])
            peak_src = "MEASURED_PEAKS.json hbm_gbs"
        except Exception:
            pass
    dp_ms_total = kt["k_dp"] * st["rounds"]                      # k_dp time per step: mean launch x round launches of a step
    ach = alg_bytes / (dp_ms_total * 1e-3) / 1e9 if dp_ms_total > 0 else None
    line = {
        "metric": METRIC, "value": cells / t_dev / 1e9, "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": max(3, args.warmup),
        "ms_per_step": t_dev * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "int16 scores / f64 likelihoods", "data": "synthetic (synth-select-v1)", "config": select_config(W),
        "value_definition": "cells / GPU-busy time of a call (reference pass + haplotype construction + scoring kernels + score "
                            "reduction, CUDA events); reads are uploaded once per call, the rounds run from HBM",
        "e2e": {"value": cells / t_wall / 1e9, "unit": UNIT, "ms_per_step": t_wall * 1e3, "ms_each_step": [x * 1e3 for x in wall],
                # what the call copies: the compact pool of sampled reads (bases, qualities, 18 B of fields each), the reference
                # segments, the variant table, and per trial haplotype its mask, sequence offset and window map (tile lists are
                # read zero-copy from pinned memory and not counted)
                "h2d_bytes_per_step": int(W * reads_per_w * (2 * READ_LEN + 18) + W * (HAP_LEN + 24) + len(vset.var_pos) * 16
                                          + int(vset.var_added_off[-1]) + 24 * st["n_trials"]),
                "d2h_bytes_per_step": int(8 * st["n_trials"]),
                "host_buffer_bytes": int(batch.input_nbytes()),
                "windows_per_s": W / t_wall,
                "api": "plb_select_haplotypes_host (pinned host buffers in, variant masks + scores out, all rounds inside)"},
        "select_stats": st, "kernel_ms_mean_per_launch": kt,
        "roofline": {"bound": "hbm", "kernel": "k_dp (all rounds of a step)", "achieved": ach, "peak": peak, "unit": "GB/s",
                     "frac": ach / peak if ach else None, "traffic": None, "peak_source": peak_src,
                     "algorithmic_bytes_per_step": alg_bytes, "kernel_ms_per_step": dp_ms_total,
                     "note": "integer-issue bound like the headline path (DESIGN.md 4)"},
        "gpu_launches": launches, "clocks": clocks,
        "check": {"n_sel_min": int(out["n_sel"].min()), "n_sel_max": int(out["n_sel"].max()), "n_scored": int(out["n_scored"][0])},
    }
    cores = os.cpu_count() or 1
    r = select_reference_run(batch, vset, min(W, 40 * cores), cores, steps=1, warmup=0) if not args.no_cpu else None
    if r is not None:
        dt, nw = r
        line["cpu_baseline"] = {"value": SEL_CELLS_PER_WINDOW * nw / dt / 1e9, "unit": UNIT, "cores": cores, "kind": "reference",
                                "windows_per_s": nw / dt,
                                "sample": "first %d windows, the reference's own getFilteredHaplotypes (oracle/_ref/n1_ref), %d processes" % (nw, cores)}
    print(json.dumps(line), flush=True)


# cells per window of synth-select-v1 with the default options: 163 trial haplotypes + the reference haplotype, 11 sampled
# reads (every 6th of 64) of 150 bp, 16 cells per base
SEL_CELLS_PER_WINDOW = 164 * 11 * 16 * 150


def pin(a):
    import torch
    if a is None:
        return None
    t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    return t


def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from platypus_b200 import _abi, synth
    from platypus_b200.engine import Engine

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    windows = args.windows
    batch, gen_s = make_workload(rank, windows)
    cells = synth.algorithmic_cells(batch)
    alg_bytes = synth.algorithmic_bytes(batch)
    W, nI, Hm = batch.n_windows, batch.n_individuals, batch.max_haps()
    Gm = Hm * (Hm + 1) // 2
    V = max(batch.max_variants, 1)

    stream = torch.cuda.Stream(device=dev)
    eng = Engine(local_rank, stream=stream.cuda_stream)
    eng_check = Engine(local_rank) if (world > 1 and rank == 0) else None     # own context for the gather check
    # The shard resident on this GPU and the path's one collective (all-gather of the per-window genotype likelihoods on a
    # side stream, overlapped with the next step): platypus_b200.shard.DeviceShard - the multi-GPU API the package ships.
    from platypus_b200.shard import DeviceShard
    ds = DeviceShard(eng, batch, stream, opt=OPT)
    handle, out, ll, gl_all, side = ds.handle, ds.out, ds.ll, ds.gl_all, ds.side
    step, join_side = ds.step, ds.join
    with torch.cuda.stream(stream):
        for _ in range(max(args.warmup, 3)):
            step()
        join_side()
        stream.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()
        launches0 = eng.launch_count
        eng.set_timing(True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(args.steps):
            step()
        join_side()          # the timed region ends when the last gather has landed
        e1.record(stream)
        stream.synchronize()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms_total = e0.elapsed_time(e1)
        gather_check = None
        if world > 1 and rank == 0:
            # the gathered tensor against a recomputation here of sampled windows of OTHER ranks' blocks (their inputs are
            # regenerated from the window ids; the engine's host path recomputes them on this GPU)
            rng = np.random.default_rng(4)
            n_ok = 0
            for r in sorted(set([1, world - 1, world // 2])):
                for k in rng.choice(windows, 4, replace=False):
                    one = synth.make_batch(1, window_offset=r * windows + int(k), n_haps=N_HAPS, n_reads=N_READS,
                                           **(dict(read_len_range=(100, 250), hap_len_range=(200, 500)) if RAGGED else
                                              dict(read_len=READ_LEN, hap_len=HAP_LEN)))
                    again = eng_check.population_run(one, opt=OPT, max_haps=Hm)["gl"][0]
                    assert np.array_equal(gl_all[r, int(k)].cpu().numpy(), again), ("gathered block differs", r, int(k))
                    n_ok += 1
            gather_check = {"windows_recomputed": n_ok, "bytes_gathered_per_step": int(gl_all.numel() * 8), "equal": True}
        launches = eng.launch_count - launches0 + (args.steps if world > 1 else 0)
        ktimes, n_timed = eng.kernel_times()
        eng.set_timing(False)
        clocks = sampler.stop() if rank == 0 else None
        stats = eng.last_stats()
        if MODE == "hla":   # clipped reads count their clipped length
            cells = stats["cells"]
        assert stats["cells"] == cells, (stats, cells)

        # ---- e2e: the C-ABI host entry points, pinned HOST buffers, every copy inside the timed region ----
        # A region loop keeps two batches in flight (plb_population_submit / plb_population_wait): batch k+1 travels
        # over PCIe while batch k computes.  Inputs are the staged form of the reads (2-bit bases + 8-bit qualities,
        # batch.pack(): what the N3 staging step produces from BAM nibbles); --ascii sends the byte-per-base arrays.
        if world > 1:
            assert torch.equal(gl_all[rank], ds.last_gl())   # the gathered block holds this rank's block
        src = batch if args.ascii else batch.pack()
        pinned = {}
        hb = type(batch)(**{f: (pin(getattr(src, f)).numpy() if isinstance(getattr(src, f), np.ndarray) else getattr(src, f))
                            for f in batch.__dataclass_fields__ if f != "_keep"})
        host_outs = []
        for j in range(2):
            ho = {"max_haps": Hm}
            for k, v in out.items():
                pinned[(k, j)] = torch.zeros(v.shape, dtype=v.dtype).pin_memory()
                ho[k] = pinned[(k, j)].numpy()
            host_outs.append(ho)
        h2d = hb.input_nbytes()
        d2h = sum(int(v.numel() * v.element_size()) for (k, j), v in pinned.items() if j == 0)
        gl_dev = [torch.zeros_like(out["gl"]) for _ in range(2)]
        e2e_gathered = [torch.cuda.Event() for _ in range(2)]

        if world > 1:   # the likelihood block stays on the GPU for the collective (output pointers may be device pointers)
            for j in range(2):
                host_outs[j]["gl"] = gl_dev[j]

        def finish(job, j, n_done):
            eng.population_wait(job)
            if world > 1:   # the step's genotype likelihoods join the other ranks' (side stream; next job keeps running)
                with torch.cuda.stream(side):
                    pinned[("gl", j)].copy_(gl_dev[j], non_blocking=True)     # ... and reaches the host like every output
                    dist.all_gather_into_tensor(gl_all, gl_dev[j])
                    e2e_gathered[j].record(side)

        def e2e_run(n):
            jobs, each, t_prev, n_done = [], [], time.perf_counter(), 0
            for i in range(n):
                if world > 1 and i >= 2:
                    e2e_gathered[i % 2].synchronize()      # the job writes gl_dev[i % 2]: its previous gather must be through
                jobs.append((eng.population_submit(hb, out=host_outs[i % 2], opt=OPT), i % 2))
                if len(jobs) == 2:
                    finish(*jobs.pop(0), n_done)
                    n_done += 1
                    t = time.perf_counter()
                    each.append(round((t - t_prev) * 1e3, 3))
                    t_prev = t
            while jobs:
                finish(*jobs.pop(0), n_done)
                n_done += 1
                t = time.perf_counter()
                each.append(round((t - t_prev) * 1e3, 3))
                t_prev = t
            if world > 1:
                side.synchronize()
            return each

        e2e_run(3)
        single = []
        for _ in range(3):   # one call alone (no second batch in flight): the latency of plb_population_run_host
            t1 = time.perf_counter()
            eng.population_run(hb, out=host_outs[0], opt=OPT)
            single.append(round((time.perf_counter() - t1) * 1e3, 3))
        if world > 1:
            dist.barrier()
        e2e_steps = max(4, min(args.steps, 20))
        t0 = time.perf_counter()
        e2e_each = e2e_run(e2e_steps)
        e2e_s = (time.perf_counter() - t0) / e2e_steps
        # parity guard: the host path and the device path must agree bit for bit
        if world > 1:
            side.synchronize()
        for j in range(2):
            assert np.array_equal(pinned[("gl", j)].numpy(), out["gl"].cpu().numpy())
        if world > 1:
            assert torch.equal(gl_all[rank], out["gl"])

    ms_step = ms_total / args.steps
    t_max = torch.tensor([ms_step, e2e_s * 1e3], dtype=torch.float64, device=dev)
    tot = torch.tensor([float(cells), float(launches)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_max, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    ms_step, e2e_ms = float(t_max[0]), float(t_max[1])
    total_cells, total_launches = float(tot[0]), int(tot[1])
    if rank != 0:
        eng.free(handle)
        return
    value = total_cells / (ms_step * 1e-3) / 1e9
    e2e_value = total_cells / (e2e_ms * 1e-3) / 1e9

    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured)"
    else:
        peak, peak_src = 6650.0, "B200_PROFILING.md fallback"
    traffic = None
    tp = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get("k_dp_dram_bytes_per_launch")
        except Exception:
            traffic = None
    kdp_ms = ktimes["k_dp"]
    achieved = alg_bytes / (kdp_ms * 1e-3) / 1e9 if kdp_ms > 0 else 0.0
    roofline = {"bound": "hbm", "kernel": "k_dp", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": kdp_ms, "launches_averaged": n_timed,
                "kernel_ms_all": ktimes,
                "note": "integer-issue bound by construction (0.014 B/cell); see DESIGN.md for the issue-rate roofline"}

    cpu = None
    if world == 1 and not args.no_cpu:
        # The CPU baseline runs in a fresh process (this one holds a CUDA context; the baseline forks workers).
        cores = os.cpu_count() or 1
        try:
            cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "1", "--warmup", "0",
                   "--windows", str(args.windows), "--config", str(args.config)]
            r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
            ref = json.loads(r.stdout.strip().splitlines()[-1])
            cpu = dict(ref["cpu_baseline"])
            r1 = subprocess.run(cmd + ["--cpu-procs", "1", "--cpu-windows", "250"], capture_output=True, text=True, timeout=900)
            one = json.loads(r1.stdout.strip().splitlines()[-1])
            cpu["single_process_value"] = one["value"]
            cpu["sample"] += "; single process: %.3f GCUPS on 250 windows" % one["value"]
        except Exception as e:   # never lose the GPU line over the baseline
            cpu = {"value": None, "unit": UNIT, "cores": cores, "kind": "unavailable", "sample": "cpu baseline failed: %r" % (e,)}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "int16 scores / f64 likelihoods", "data": "synthetic (synth-v1, seed 20261017)",
        "config": workload_config(world, windows),
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "ms_per_step": e2e_ms, "ms_each_step": e2e_each, "steps": e2e_steps,
                "single_call_ms": single, "input_format": "ascii" if args.ascii else "2-bit bases + %d-bit quality codes (batch.pack)" % (hb.qual_bits or 8),
                "api": "plb_population_submit / plb_population_wait, two batches in flight (pinned host buffers)%s"
                       % ("; per step the likelihood block is all-gathered over NCCL" if world > 1 else "")},
        "gpu_launches": total_launches,
        "roofline": roofline,
        "cpu_baseline": cpu,
        "stats": stats, "gen_seconds": gen_s,
    }
    if gather_check:
        line["gather_check"] = gather_check
    print(json.dumps(line), flush=True)
    eng.free(handle)


# ---- --config 5: multi-sample mode (BASELINE config 5) ------------------------------------------------------------------

C5_WINDOWS, C5_IND, C5_READS = 30000, 2000, 50     # the stated 1/1000 subsample of ~30 M windows; 2000 samples x ~50 reads


def c5_template(n_windows):
    """A batch of the config-5 SHAPE (offsets and counts only; the bytes are generated on the device)."""
    from platypus_b200.batch import WindowBatch
    W, nI, R, H = n_windows, C5_IND, C5_READS, N_HAPS
    n_reads = W * nI * R
    mv = 2 * (H - 1)
    return WindowBatch(
        n_windows=W, n_individuals=nI, win_hap_off=np.arange(W + 1, dtype=np.int32) * H,
        win_start=np.zeros(W, np.int32), win_end=np.full(W, 50, np.int32), hap_start=np.zeros(W, np.int32),
        hap_seq_off=np.arange(W * H + 1, dtype=np.int64) * HAP_LEN, hap_seq=np.full(W * H * HAP_LEN + 1, 65, np.uint8),
        wi_slot_off=np.arange(W * nI + 1, dtype=np.int64) * R, wi_n_good=np.full(W * nI, R, np.int32),
        wi_n_bad=np.zeros(W * nI, np.int32), slot_read=np.arange(n_reads, dtype=np.int32),
        read_seq_off=np.arange(n_reads + 1, dtype=np.int64) * READ_LEN, read_seq=np.full(n_reads * READ_LEN + 1, 65, np.uint8),
        read_qual=np.zeros(n_reads * READ_LEN + 1, np.uint8), read_pos=np.zeros(n_reads, np.int32),
        read_end=np.full(n_reads, READ_LEN, np.int32), read_mapq=np.zeros(n_reads, np.uint8),
        read_qcfail=np.ones(n_reads, np.uint8),     # the template itself is never scored (every read "fails QC")
        max_variants=mv, win_n_var=np.zeros(W, np.int32), hap_var_mask=np.zeros(W * H, np.uint64),
        var_prior=np.zeros((W, mv), np.float64))


def run_config5(args, rank, world, local_rank):
    """BASELINE config 5: ~30 M windows x 2000 samples is ~6e16 cells and ~1 PB of inputs, so - as SURVEY 8d lays down - a
    fixed 1/1000 subsample (30,000 windows, the same ones at every N: strong scaling) is generated ON THE DEVICE chunk by
    chunk (plb_synth_fill_device, "synth-v1d"), scored and reduced to genotype calls / variant posteriors / haplotype
    frequencies on the device, and only those are gathered (one all-gather at the end).  A step = one chunk of windows."""
    import torch
    import torch.distributed as dist
    from platypus_b200 import _abi
    from platypus_b200.engine import Engine
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    total = args.c5_windows
    chunk = args.c5_chunk
    lo, hi = total * rank // world, total * (rank + 1) // world
    nI, R, H = C5_IND, C5_READS, N_HAPS
    Gm, V = H * (H + 1) // 2, 2 * (H - 1)
    stream = torch.cuda.Stream(device=dev)
    eng = Engine(local_rank, stream=stream.cuda_stream)
    tmpl = c5_template(chunk)
    cells_per_window = nI * R * H * 16 * READ_LEN
    f64, i32 = torch.float64, torch.int32
    n_mine = hi - lo
    n_pad = (total + world - 1) // world
    with torch.cuda.stream(stream):
        handle = eng.upload(tmpl)
        out = {"gl": torch.zeros((chunk, nI, Gm), dtype=f64, device=dev), "gl_log_max": torch.zeros((chunk, nI), dtype=f64, device=dev),
               "freq": torch.zeros((chunk, H), dtype=f64, device=dev), "em_post": torch.zeros((chunk, nI, Gm), dtype=f64, device=dev),
               "call": torch.zeros((chunk, nI), dtype=i32, device=dev), "var_phred": torch.zeros((chunk, V), dtype=f64, device=dev),
               "em_iters": torch.zeros((chunk,), dtype=i32, device=dev)}
        ptrs = {k: v.data_ptr() for k, v in out.items()}
        ptrs["max_haps"] = H
        # per-rank results: what is gathered at the end (calls as bytes: G = 36 < 256)
        calls = torch.zeros((n_pad, nI), dtype=torch.uint8, device=dev)
        phred = torch.zeros((n_pad, V), dtype=f64, device=dev)
        freq = torch.zeros((n_pad, H), dtype=f64, device=dev)

        gen_events = []

        def do_chunk(first, n_live, store_at):
            if store_at is not None:
                g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                g0.record(stream)
            eng.synth_fill(handle, first)
            if store_at is not None:
                g1.record(stream)
                gen_events.append((g0, g1))
            eng.run_device(handle, ptrs, opt=OPT)
            if store_at is not None and n_live > 0:
                calls[store_at:store_at + n_live] = (out["call"][:n_live] + 1).to(torch.uint8)      # 0 = no call
                phred[store_at:store_at + n_live] = out["var_phred"][:n_live]
                freq[store_at:store_at + n_live] = out["freq"][:n_live]

        for k in range(max(3, args.warmup)):
            do_chunk(lo + k * chunk, 0, None)
        stream.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()
        eng.set_timing(True)
        l0 = eng.launch_count
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        n_chunks = 0
        for first in range(lo, hi, chunk):
            do_chunk(first, min(chunk, hi - first), first - lo)
            n_chunks += 1
        if world > 1:   # the one collective: calls, posteriors and frequencies of every window to every rank
            g_calls = torch.zeros((world, n_pad, nI), dtype=torch.uint8, device=dev)
            g_phred = torch.zeros((world, n_pad, V), dtype=f64, device=dev)
            g_freq = torch.zeros((world, n_pad, H), dtype=f64, device=dev)
            dist.all_gather_into_tensor(g_calls, calls)
            dist.all_gather_into_tensor(g_phred, phred)
            dist.all_gather_into_tensor(g_freq, freq)
        e1.record(stream)
        stream.synchronize()
        torch.cuda.synchronize()
        ms_total = e0.elapsed_time(e1)
        ms_gen = sum(a.elapsed_time(b_) for a, b_ in gen_events)
        launches = eng.launch_count - l0
        ktimes, n_timed = eng.kernel_times()
        eng.set_timing(False)
        clocks = sampler.stop() if rank == 0 else None
        stats = eng.last_stats()

        # parity: regenerate sampled windows, download their inputs and run the CPU oracle on them
        check = None
        if rank == 0 and not args.no_cpu:
            from oracle import oracle as O
            rng = np.random.default_rng(5)
            picks = sorted(int(x) for x in rng.choice(max(1, n_mine), min(2, n_mine), replace=False))
            n_checked = 0
            for k in picks:
                first = lo + (k // chunk) * chunk
                eng.synth_fill(handle, first)
                eng.run_device(handle, ptrs, opt=OPT)
                stream.synchronize()
                host = eng.download(handle, tmpl)
                one = host.slice_windows(k % chunk, k % chunk + 1)
                want, _, _, _ = O.population_run(one, OPT, n_threads=os.cpu_count() or 1, max_haps=H)
                got_call = out["call"][k % chunk].cpu().numpy()
                assert np.array_equal(got_call, want["call"][0]), "config 5: genotype calls differ from the oracle"
                assert np.array_equal(out["var_phred"][k % chunk].cpu().numpy()[:want["var_phred"].shape[1]], want["var_phred"][0])
                np.testing.assert_allclose(out["freq"][k % chunk].cpu().numpy(), want["freq"][0], rtol=1e-9)
                assert np.array_equal(calls[k].cpu().numpy(), (got_call + 1).astype(np.uint8))
                n_checked += 1
            check = {"windows_vs_oracle": n_checked, "individuals": nI, "calls_equal": True, "var_phred_equal": True}

    t = torch.tensor([ms_total, ms_total - ms_gen], dtype=torch.float64, device=dev)
    tot = torch.tensor([float(n_mine) * cells_per_window, float(launches)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    if rank != 0:
        eng.free(handle)
        return
    ms, ms_score = float(t[0]), float(t[1])
    cells = float(tot[0])
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peak = float(json.load(open(peaks_path))["hbm_gbs"]) if os.path.exists(peaks_path) else 6650.0
    # algorithmic bytes of one chunk (SURVEY 8d layout): reads, haplotypes, per-read LL, GL
    alg = chunk * (nI * R * ((READ_LEN + 3) // 4 + READ_LEN + 8) + H * ((HAP_LEN + 3) // 4 + 8) + 8 * H * nI * R + 8 * nI * Gm)
    ach = alg / (ktimes["k_dp"] * 1e-3) / 1e9 if ktimes["k_dp"] > 0 else None
    line = {
        # value: a chunk's inputs are resident when its scoring starts (generation time excluded, max over ranks);
        # e2e: everything - generation, scoring, window model, final gather
        "metric": METRIC, "value": cells / (ms_score * 1e-3) / 1e9, "unit": UNIT, "n_gpus": world, "steps": n_chunks,
        "warmup": max(3, args.warmup), "ms_per_step": ms_score / max(1, n_chunks), "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "int16 scores / f64 likelihoods", "data": "synthetic, generated on the device (synth-v1d, seed 20261017)",
        "config": {"workload": "config5 multi-sample: fixed 1/1000 subsample = %d windows x %d individuals x %d reads x %d haplotypes, "
                               "%d bp reads x %d bp haplotypes; the same windows at every N, %d windows per chunk; calls / posteriors / "
                               "frequencies reduced on the device and gathered once" % (total, nI, R, H, READ_LEN, HAP_LEN, chunk),
                   "windows_total": total, "windows_per_gpu": n_mine, "individuals": nI, "reads_per_individual": R, "n_gpus": world,
                   "extrapolation": "30 M windows = 1000 x this job: %.1f GPU-hours at this rate" % (1000 * ms * 1e-3 * world / 3600.0),
                   "l2": "one chunk's inputs (~%.1f GB) exceed the 126 MB L2; no explicit flush" % (chunk * nI * R * 2 * READ_LEN / 1e9)},
        "total_ms": ms, "scoring_ms": ms_score, "cells": cells, "clocks": clocks, "gpu_launches": int(tot[1]),
        "e2e": {"value": cells / (ms * 1e-3) / 1e9, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                "note": "inputs are generated on the device (they do not fit PCIe at this scale); the timed region includes generation, "
                        "scoring, the window model and the final gather"},
        "roofline": {"bound": "hbm", "kernel": "k_dp", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak if ach else None,
                     "traffic": None, "algorithmic_bytes_per_launch": alg, "kernel_ms": ktimes["k_dp"], "kernel_ms_all": ktimes,
                     "launches_averaged": n_timed},
        "gathered_bytes": int(world * n_pad * (nI + 8 * V + 8 * H)) if world > 1 else 0,
        "oracle_check": check, "stats_last_chunk": stats,
    }
    print(json.dumps(line), flush=True)
    eng.free(handle)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--windows", type=int, default=0, help="windows per GPU (default: 10,000 = config 2; 62,500 for --config 4)")
    ap.add_argument("--cpu-procs", type=int, default=0, help="--impl reference: processes (default: all host cores)")
    ap.add_argument("--cpu-windows", type=int, default=0, help="--impl reference: windows per step (default: the workload)")
    ap.add_argument("--mode", default="default", choices=["default", "flank", "hla"],
                    help="run-time mode of the path: --calculateFlankScore=1 / --HLATyping=1 (not the headline)")
    ap.add_argument("--config", type=int, default=2, choices=[2, 3, 4, 5],
                    help="2 = BASELINE config 2 (headline); 3 = ragged read/haplotype lengths (profiling only); 4 = 62,500 "
                         "windows per GPU (chr22-sized over 8 GPUs) with the gathered likelihoods checked; 5 = multi-sample "
                         "mode, 2000 individuals, device-generated windows, strong scaling")
    ap.add_argument("--stage", default="path", choices=["path", "select"],
                    help="path = the headline likelihood path; select = the haplotype selection loop before it (N1, one GPU)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline leg")
    ap.add_argument("--c5-windows", type=int, default=C5_WINDOWS, help="--config 5: windows of the whole job (all GPUs together)")
    ap.add_argument("--c5-chunk", type=int, default=148, help="--config 5: windows generated and scored per step")
    ap.add_argument("--ascii", action="store_true", help="e2e leg: send byte-per-base sequences instead of the packed form")
    args = ap.parse_args()
    if args.stage == "select":
        args.windows = args.windows or WINDOWS_PER_GPU
        if int(os.environ.get("RANK", "0")) == 0:
            run_select(args)
        return
    global RAGGED, OPT, MODE, CONFIG
    CONFIG = args.config
    if not args.windows:
        args.windows = 62500 if args.config == 4 else WINDOWS_PER_GPU
    RAGGED = args.config == 3
    MODE = args.mode
    if args.mode != "default":
        from platypus_b200 import _abi
        OPT = _abi.PlbOptions.default(**({"calc_flank_score": 1} if args.mode == "flank" else {"use_mapq_cap": 1}))

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        if args.config == 5:
            run_config5(args, rank, world, local_rank)
        else:
            run_ours(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
