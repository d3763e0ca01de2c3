#!/usr/bin/env python
"""bench.py -- subspace-detector throughput (template*samples / s) on N B200s.

    python bench.py --gpus N --steps K --warmup W            # our arm
    python bench.py --impl reference --gpus N --steps K --warmup W   # CPU reference arm

Headline workload (weak scaling, one process per GPU): every GPU runs the detector of ONE station of
BASELINE.json configs[3] -- 3 channels x 100 Hz x 30 days as 720 chunks of 3720 s
(L = 1 116 000 multiplexed samples, T = 369 001 lags per chunk) against 256 subspaces of
rank 1..8 (R = 1152 basis vectors, n = 9000).  At N = 8 the job IS configs[3].
One step = one pass of the hot path (K0 prep -> K1 tcgen05 projection+normalisation -> K3
max / histogram / candidate compaction / LTA) over all 720 chunks of the station -- batches are
enqueued back to back, results accumulate on the device and are fetched once per step -- plus the
end-of-step gather of trigger candidates and histogram all-reduce when N > 1.

`value`  : whole-job template*samples/s with the chunks already resident in HBM.
`e2e`    : the same through the C ABI with HOST (pinned) buffers: H2D of every chunk and
           D2H of MaxDS / candidates / histograms inside the timed region.
`roofline`: K1, tensor-bound: algorithmic flops (2*n*R per lag) / its CUDA-event duration.
`cpu_baseline`: the reference's own `_SSDetex._MPXDS` (detect.py:559-578, unmodified copy under
           oracle/_ref, made by oracle/make_ref.py) on the host cores, bounded sample; reported
           beside, not the optimisation target.
Sub-objects `cfg1`, `ccx`, `fas`: the other BASELINE.json configs ([1] rank-3 day, [2] CCX of 4096
events, [4] FAS sweep of 1000 x 256), each with its own value / e2e / roofline / cpu_baseline.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SR = 100.0
NC = 3
NS = 3000                 # 30 s template at 100 Hz
N_MUX = NC * NS           # 9000
LS = 372000               # 3720 s chunk (conDatDuration 3600 + conBuff 120, getdata.py:299)
LS_FAS = 360000           # 3600 s null segment (BASELINE configs[4])
CHUNKS_PER_STATION = 720  # 30 days
NSUB = 256
T_PER_CHUNK = LS - NS + 1
T_FAS = LS_FAS - NS + 1
METRIC = "subspace_detector_template_samples_per_sec"
UNIT = "template*samples/s"
CCX_EVENTS, CCX_NS = 4096, 1000      # BASELINE configs[2]: 4096 events x 3 ch x 10 s x 100 Hz

DTYPES = {
    "tcgen05": "fp16x3-split (fp32-equivalent), f64 window energy",
    "tcgen05_x8": "fp16 hi*hi + e4m3 x e5m2 cross terms, fp32 accumulate, f64 window energy",
    "tcgen05_auto": "per chunk: fp16x3-split, or fp16 hi*hi + e4m3 x e5m2 cross terms where the error "
                    "model admits them; fp32 accumulate, f64 window energy",
}


def ranks_list(nsub):
    return [(i % 8) + 1 for i in range(nsub)]


# ----------------------------------------------------------------------------- clocks
class ClockSampler(object):
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
                pw.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"),
                                 f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w": float(np.median(pw)) if pw else None, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------- CPU arms
# Workers of a multiprocessing pool (spawned: the GPU arm's parent holds a CUDA context).  A task
# names a bounded sample of the workload by seed; the worker builds the arrays once (untimed, cached),
# prepares what the reference prepares ONCE per station outside its hot loop -- ssFD = fft of the
# reversed basis (detect.py:371; fas.py:171), MPfd of every event (construct.py:669-676) -- and times
# only what the reference repeats per chunk / per pair.
_W = {}


def _ref():
    """The unmodified reference through the compatibility shim, or None (then the oracle port runs)."""
    if "ref" not in _W:
        try:
            from oracle import ref_shim
            _W["ref"] = ref_shim.RefFunctions() if ref_shim.available() else None
        except Exception:
            _W["ref"] = None
    return _W["ref"]


def _sample_chunk(ls):
    from detex_b200 import synth
    return synth.multiplex(synth.bandpassed_noise(np.random.default_rng([4004, ls]), ls, sr=SR, nchan=NC))


def _sample_basis(i):
    from detex_b200 import synth
    return synth.random_basis(np.random.default_rng([4004, 7, i]), N_MUX, ranks_list(i + 1)[i])


def _det_task(args):
    """Detection (`fas` = 0: detect.py:255-269, one data FFT per chunk then `_MPXDS` per subspace) or
    FAS (`fas` = 1: fas.py:110-111, `_MPXSSCorr` re-transforms the data for every subspace) of ONE
    chunk against subspaces [lo, hi).  Returns (DS values produced, seconds inside the hot loop)."""
    import scipy.fftpack
    ls, lo, hi, is_fas = args
    key = ("det", ls, lo, hi)
    R = _ref()
    if key not in _W:
        chunk = _sample_chunk(ls)
        bases = [_sample_basis(i) for i in range(lo, hi)]
        reqlen = int(len(chunk) + N_MUX)
        nfft = 2 ** reqlen.bit_length()
        ssfd = [np.array([scipy.fftpack.fft(u[::-1], n=nfft) for u in U]) for U in bases] if R else None
        _W[key] = (chunk, bases, reqlen, nfft, ssfd)
    chunk, bases, reqlen, nfft, ssfd = _W[key]
    t0 = time.perf_counter()
    tot = 0
    if R is None:
        from oracle import detex_oracle as orc
        for U in bases:
            tot += len(orc.mpx_ds_fft(chunk, U, NC))
    elif is_fas:
        for U, fd in zip(bases, ssfd):
            tot += len(R.fas._MPXSSCorr(chunk, reqlen, U, fd, NC))
    else:
        MPconFD = scipy.fftpack.fft(chunk, n=nfft)                      # detect.py:255-256
        for U, fd in zip(bases, ssfd):
            tot += len(R._ssd._MPXDS(chunk, reqlen, U, fd, NC, MPconFD))
    return tot, time.perf_counter() - t0


def _ccx_task(args):
    """Rows [lo, hi) of `_makeDFcclags`' pair loop (construct.py:380-393) over `nev` events."""
    import scipy.fftpack
    nev, lo, hi = args
    key = ("ccx", nev)
    R = _ref()
    if key not in _W:
        from detex_b200 import synth
        X = synth.event_families(3003, max(1, nev // 64), 64, CCX_NS, NC, max_shift=100)[:nev]
        nfft = 2 ** int(2 * X.shape[1]).bit_length()
        fd = [scipy.fftpack.fft(x, n=nfft) for x in X] if R else None
        _W[key] = (X, fd)
    X, fd = _W[key]
    chans = ["C%d" % i for i in range(NC)]
    t0 = time.perf_counter()
    npairs = 0
    if R is None:
        from oracle import detex_oracle as orc
        for b in range(lo, hi):
            for c in range(b + 1, nev):
                orc.ccx2(X[b], X[c], NC)
                npairs += 1
    else:
        for b in range(lo, hi):
            for c in range(b + 1, nev):
                R.construct._CCX2(fd[b], fd[c], X[b], X[c], chans, chans)
                npairs += 1
    return npairs, time.perf_counter() - t0


def _kind():
    return "reference" if _ref() is not None else "port"


def cpu_workers():
    """Processes for the CPU arms: all host cores, bounded by memory (a detection task holds the
    FFTs of 36 basis vectors of 2^21 points plus the reference's r x N_fft temporaries, ~3.5 GB)."""
    cores = os.cpu_count() or 1
    try:
        import psutil
        cores = max(1, min(cores, int(psutil.virtual_memory().available / 4.5e9)))
    except Exception:
        pass
    return cores


def make_pool(workers):
    import multiprocessing as mp
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
    os.environ.setdefault("MKL_NUM_THREADS", "1")
    return mp.get_context("spawn").Pool(workers)


def cpu_detection(pool, workers, ls, is_fas, reps=1):
    """One bounded sample: 1 chunk x (8 x workers, at most 256) subspaces in groups of 8 (ranks 1..8).
    Returns (values / s over all workers, wall s, values / s of ONE process = as shipped, sample text)."""
    ngroups = max(1, min(NSUB // 8, workers))
    tasks = [(ls, 8 * g, 8 * g + 8, int(is_fas)) for g in range(ngroups)]
    pool.map(_det_task, tasks)                        # builds and caches the arrays and ssFD (untimed)
    t0 = time.perf_counter()
    tot = 0
    for _ in range(reps):
        tot += sum(r[0] for r in pool.map(_det_task, tasks))
    wall = time.perf_counter() - t0
    one = pool.apply(_det_task, (tasks[0],))          # one process alone on the machine
    what = "fas._MPXSSCorr (fas.py:120-134)" if is_fas else "_SSDetex._MPXDS (detect.py:559-578), data FFT once per chunk"
    sample = ("1 chunk (%d s x 3 ch x 100 Hz) x %d subspaces (ranks 1-8, n=9000) per step; %s %s; basis FFTs "
              "prepared once outside the timed loop as detect.py:371 does; Pool(%d)"
              % (ls // 100, 8 * ngroups, _kind(), what, workers))
    return tot / wall, wall / reps, one[0] / one[1], sample, 8 * ngroups


def cpu_ccx(pool, workers, nev=256):
    """`_CCX2` over all pairs of `nev` events, row blocks with equal pair counts over the workers."""
    from detex_b200 import parallel
    blocks = [b for b in parallel.ccx_row_blocks(nev, max(1, min(workers, nev // 4))) if b[1] > b[0]]
    tasks = [(nev, b0, b1) for b0, b1 in blocks]
    pool.map(_ccx_task, [(nev, 0, 0)] * len(tasks))   # builds and caches X and MPfd (untimed)
    t0 = time.perf_counter()
    res = pool.map(_ccx_task, tasks)
    wall = time.perf_counter() - t0
    pairs = sum(r[0] for r in res)
    one = pool.apply(_ccx_task, ((nev, 0, 8),))
    return pairs / wall, wall, one[0] / one[1], pairs


def reference_arm(args):
    """--impl reference: the reference's own CPU implementation of the headline path on all host cores,
    each step a bounded sample of the GPU arm's workload (same shapes, same metric)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    workers = cpu_workers()
    with make_pool(workers) as pool:
        ngroups = max(1, min(NSUB // 8, workers))
        tasks = [(LS, 8 * g, 8 * g + 8, 0) for g in range(ngroups)]
        pool.map(_det_task, tasks)                    # untimed: arrays, ssFD (detect.py:371), imports
        for _ in range(args.warmup):
            pool.map(_det_task, tasks)
        t0 = time.perf_counter()
        tot = 0
        for _ in range(args.steps):
            tot += sum(r[0] for r in pool.map(_det_task, tasks))
        el = time.perf_counter() - t0
        one = pool.apply(_det_task, (tasks[0],))
        kind = pool.apply(_kind)
    value = tot / el
    nsub = 8 * ngroups
    sample = ("1 chunk (3720 s x 3 ch x 100 Hz) x %d subspaces (ranks 1-8, n=9000) per step; %s _SSDetex._MPXDS "
              "(detect.py:559-578) with ssFD prepared once (detect.py:371) and MPconFD once per chunk and worker "
              "(detect.py:255-256); Pool(%d)" % (nsub, kind, workers))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * el / max(1, args.steps),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": "BASELINE configs[3] shapes on the host CPU: chunks of 3720 s x 3 ch x 100 Hz "
                               "(L = 1 116 000), subspaces of rank 1-8, n = 9000; bounded sample per step",
                   "subspaces": NSUB, "subspaces_per_step": nsub, "chunks_per_step": 1, "n": N_MUX,
                   "lags_per_chunk": T_PER_CHUNK, "input_dtype": "f64", "engine": "scipy.fftpack + pandas rolling"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": workers, "kind": kind, "sample": sample,
                         "one_process_value": one[0] / one[1],
                         "note": "one_process_value = the reference as shipped (1 process, 1 thread, "
                                 "subspace.py:1843-1845 rejects multiprocess)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------ GPU arm
def make_station_data(torch, dev, nchunks, seed):
    """[nchunks, L] float64 multiplexed band-passed noise, generated on the GPU (plumbing)."""
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    L = LS * NC
    out = torch.empty((nchunks, L), dtype=torch.float64, device=dev)
    freqs = torch.fft.rfftfreq(LS, d=1.0 / SR).to(dev)
    # 2nd-order Butterworth band-pass magnitude (zero phase), 1-10 Hz
    w = freqs.clamp_min(1e-6)
    mag = 1.0 / torch.sqrt(1 + ((w * w - 1.0 * 10.0) / (w * (10.0 - 1.0))) ** 4)
    for i in range(nchunks):
        x = torch.randn((NC, LS), generator=g, device=dev, dtype=torch.float32)
        X = torch.fft.rfft(x, dim=1) * mag[None, :]
        y = torch.fft.irfft(X, n=LS, dim=1)
        y = y / y.std(dim=1, keepdim=True)
        out[i] = y.t().contiguous().reshape(-1).to(torch.float64)
    return out


def measured_peaks():
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return json.load(open(pk))
    except Exception:
        return {}


def gpu_arm(args):
    import torch
    import torch.distributed as dist
    from detex_b200 import fas as dfas
    from detex_b200 import parallel, synth
    from detex_b200.engine import Engine

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    stream = torch.cuda.current_stream()
    eng = Engine(local, stream=stream.cuda_stream)
    eng.set_fused(args.fused)
    sections = set(args.sections.split(","))
    peaks = measured_peaks()

    ranks = ranks_list(args.nsub)
    rng = np.random.default_rng(4004)          # same bases on every rank (one detector set)
    bases = [synth.random_basis(rng, N_MUX, r) for r in ranks]
    thr = [0.25] * args.nsub
    eng.set_bases(0, bases, NC, thresholds=thr)

    data = make_station_data(torch, dev, args.chunks, seed=1000 + rank)     # resident in HBM
    # one planted event per day
    for day in range(max(1, args.chunks // 24)):
        ci = min(args.chunks - 1, day * 24 + 7)
        s = (17 * day + rank) % args.nsub
        tem = torch.from_numpy(np.ones(ranks[s]) @ bases[s]).to(dev)
        t0 = 1000 + 97 * day
        data[ci, t0 * NC:t0 * NC + N_MUX] += 6.0 * np.sqrt(N_MUX) / float(tem.norm()) * tem
    torch.cuda.synchronize()
    L = LS * NC
    offs_all = np.arange(args.chunks, dtype=np.int64) * L
    lens_all = np.full(args.chunks, L, dtype=np.int64)
    nbatch = (args.chunks + args.batch - 1) // args.batch
    flops_per_chunk = 2.0 * N_MUX * sum(ranks) * T_PER_CHUNK

    k1_ms = []

    def step_resident(engine=args.engine, set_id=0, offs=offs_all, lens=lens_all, hist_range=(0.0, 1.0),
                      want_fas=False, lta=int(5 * SR), batch=args.batch, collect=True):
        """One pass over the station with the chunks resident in HBM: the batches are enqueued back to
        back, the results accumulate on the device and come back once at the end."""
        n = len(offs)
        eng.accumulate_begin(n)
        for lo in range(0, n, batch):
            hi = min(n, lo + batch)
            eng.attach_device_chunks(data.data_ptr(), offs[lo:hi], lens[lo:hi])
            eng.detect_run(set_id, engine=engine, kblk=args.kblk, hist_range=hist_range, lta_window=lta,
                           want_fas=want_fas)
        c = eng.candidates()
        hist = eng.hist(set_id, reset=True)
        if collect:
            k1_ms.extend(zip(eng.k1_ms_history().tolist(), [min(n, lo + batch) - lo for lo in range(0, n, batch)]))
        eng.accumulate_end()
        if world > 1:                      # the only exchange: trigger lists + histograms
            c = parallel.gather_records(c)
            hist = parallel.allreduce_sum(hist)
        return c, hist

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, nsteps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(stream)
        out = None
        for _ in range(nsteps):
            out = fn()
        e1.record(stream)
        barrier()
        wall = time.perf_counter() - t0
        ms = e0.elapsed_time(e1)
        own = max(ms / 1e3, wall)
        if world > 1:
            t = torch.tensor([ms, wall * 1e3], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms, wall = float(t[0]), float(t[1]) / 1e3
        return ms, wall, out, own

    def all_ranks(obj):
        if world == 1:
            return [obj]
        out = [None] * world
        dist.all_gather_object(out, obj)
        return out

    line = {"metric": METRIC, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": DTYPES[args.engine],
            "data": "synthetic"}

    # =============================================================== headline: configs[3] shard
    if "main" in sections:
        for _ in range(args.warmup):
            step_resident(collect=False)
        k1_ms.clear()
        launches0 = eng.launch_count()
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
        ms, wall, (cands, hist), own_s = timed(step_resident, args.steps)
        clocks = sampler.stop() if rank == 0 else None
        launches = (eng.launch_count() - launches0) // max(1, args.steps)
        ts_per_step = float(args.chunks) * T_PER_CHUNK * args.nsub * world
        # device time (events) and wall time agree to <1 %: the host only enqueues; report the
        # slower of the two so the result D2H at the end of the step is inside the number
        step_s = max(ms / 1e3, wall) / args.steps
        value = ts_per_step / step_s

        # K1 roofline (this rank's launches in the timed region)
        tot_ms = sum(m for m, _ in k1_ms)
        tot_chunks = sum(n for _, n in k1_ms)
        k1_avg_ms = tot_ms / len(k1_ms)
        achieved = flops_per_chunk * tot_chunks / (tot_ms * 1e-3) / 1e12
        peak, peak_src = 1590.0, "fallback (B200_PROFILING.md)"
        if "bf16_tflops_sustained" in peaks:
            peak = float(peaks["bf16_tflops_sustained"])
            peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside a long step)"
        traffic = None
        tj = os.path.join(ROOT, "profiles", "k1_traffic.json")
        if os.path.exists(tj):
            try:
                traffic = json.load(open(tj)).get("dram_bytes_per_launch")
            except Exception:
                pass
        per_rank = all_ranks({"rank": rank, "k1_ms_min": float(min(m for m, _ in k1_ms)),
                              "k1_ms_median": float(np.median([m for m, _ in k1_ms])),
                              "k1_ms_max": float(max(m for m, _ in k1_ms)), "k1_ms_total": float(tot_ms),
                              "step_s": own_s / args.steps,
                              "k1_share_of_own_step": tot_ms / 1e3 / own_s})
        roofline = {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                    "traffic": traffic,
                    "traffic_note": "dram bytes of one 48-chunk K1 launch (ncu); algorithmic bytes of that launch = "
                                    "19.0e9 with the dense DS rows it writes (18.2e9) + inputs 0.7e9",
                    "kernel": "k1_kernel (tcgen05 Hankel projection + normalisation)",
                    "peak_source": peak_src, "k1_ms_per_launch": k1_avg_ms,
                    "k1_share_of_step": tot_ms / 1e3 / (step_s * args.steps),
                    "note": "algorithmic flops = 2*n*R per lag; fp32-equivalent precision costs 3 fp16 MMAs per "
                            "product, so frac is bounded by 1/3"}
        roofline["issued_tflops"] = achieved * 3.0 * (3 * 3008.0 / N_MUX)
        line.update({"value": value, "ms_per_step": 1e3 * step_s, "config": workload_config(args),
                     "gpu_launches": int(launches), "roofline": roofline, "clocks": clocks, "per_rank": per_rank,
                     "candidates_per_step": int(len(cands)), "hist_total": int(hist.sum())})

        # ------------------------------------------------- parity spot check at full size (untimed)
        # the planted chunk of day 0 against the float64 closed form on the device, first 8 subspaces
        npar = min(8, args.nsub)
        eng.set_bases(1, bases[:npar], NC)
        ds_default = {}

        def parity_check(engine):
            ci = min(args.chunks - 1, 7)
            eng.attach_device_chunks(data.data_ptr(), offs_all[ci:ci + 1], lens_all[ci:ci + 1])
            eng.detect_run(1, engine=engine, kblk=args.kblk, keep_ds64=True)
            perr, pmax, ndiff = 0.0, 0.0, 0
            for si in range(npar):
                d64 = eng.get_ds64(0, si)
                ds = eng.get_ds(0, si)
                perr = max(perr, float(np.abs(ds - d64).max()))
                pmax = max(pmax, float(d64.max()))
                if engine == "tcgen05":
                    ds_default[si] = ds
                elif si in ds_default:
                    ndiff += int((ds != ds_default[si]).sum())
            assert perr < 1e-5, "detection statistic out of tolerance against the float64 closed form: %g" % perr
            out = {"max_abs_err_vs_fp64": perr, "tol": 1e-5, "max_ds": pmax, "chunk": ci, "subspaces": npar,
                   "lags": T_PER_CHUNK, "x8_mode": int(eng.chunk_modes()[0])}
            if engine != "tcgen05":
                # the maximum error sits on the planted peak (DS ~ 0.98), where the hi*hi truncation bias
                # dominates and both engines round to the same float32; elsewhere the values differ
                out["ds_values_differing_from_default_engine"] = ndiff
            return out

        line["parity_check"] = parity_check("tcgen05")

        # -------------------- the opt-in adaptive-precision engine on the same resident data (reported
        # beside the headline, which stays on the worst-case-bounded default engine)
        if args.engine == "tcgen05" and not args.no_alt:
            x8n = 0
            for lo in range(0, args.chunks, args.batch):   # untimed pass: which chunks the model admits
                hi = min(args.chunks, lo + args.batch)
                eng.attach_device_chunks(data.data_ptr(), offs_all[lo:hi], lens_all[lo:hi])
                eng.detect_run(0, engine="tcgen05_auto", kblk=args.kblk)
                x8n += int(eng.chunk_modes().sum())
            eng.hist(0, reset=True)
            k1_ms.clear()
            nalt = max(1, min(args.steps, 3))
            ms_a, wall_a, (cands_a, hist_a), _ = timed(lambda: step_resident("tcgen05_auto"), nalt)
            line["adaptive_engine"] = {
                "engine": "tcgen05_auto", "dtype": DTYPES["tcgen05_auto"], "steps": nalt,
                "value": ts_per_step / (max(ms_a / 1e3, wall_a) / nalt), "unit": UNIT,
                "x8_chunk_fraction": x8n / float(args.chunks),
                "k1_ms_per_launch": sum(m for m, _ in k1_ms) / len(k1_ms),
                "candidates_per_step": int(len(cands_a)),
                "hist_bins_moved_vs_default": int(np.abs(hist_a - hist).sum() // 2),
                "parity_check": parity_check("tcgen05_auto")}

        # ---------------------------------------------------------------- end-to-end (host buffers)
        host = torch.empty((args.chunks, L), dtype=torch.float64, pin_memory=True)
        host.copy_(data)
        torch.cuda.synchronize()
        hnp = host.numpy()
        d2h = [0]

        def step_e2e():
            eng.accumulate_begin(args.chunks)
            for b in range(nbatch):
                lo, hi = b * args.batch, min(args.chunks, (b + 1) * args.batch)
                eng.load_chunks([hnp[i] for i in range(lo, hi)])                  # H2D (pinned), async
                eng.detect_run(0, engine=args.engine, kblk=args.kblk, lta_window=int(5 * SR))
            mx, fl = eng.rowstats()                                                # D2H
            c = eng.candidates()                                                   # D2H
            hist = eng.hist(0, reset=True)                                         # D2H
            eng.accumulate_end()
            d2h[0] += mx.nbytes + fl.nbytes + c.nbytes + 8 + hist.nbytes
            if world > 1:
                c = parallel.gather_records(c)
                hist = parallel.allreduce_sum(hist)
            return c, hist

        step_e2e()
        d2h[0] = 0
        ms_e, wall_e, (cands_e, hist_e), _ = timed(step_e2e, args.steps)
        e2e_val = ts_per_step / (max(ms_e / 1e3, wall_e) / args.steps)
        assert len(cands_e) == len(cands) and np.array_equal(hist_e, hist), "resident and host paths disagree"
        line["e2e"] = {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": int(args.chunks) * L * 8,
                       "d2h_bytes_per_step": int(d2h[0] // max(1, args.steps))}
    else:
        hnp = None

    pool = None
    workers = cpu_workers()
    if rank == 0 and world == 1 and not args.no_cpu:
        pool = make_pool(workers)
    if pool is not None and "main" in sections:
        v, dt, one, sample, _ = cpu_detection(pool, workers, LS, False)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": workers, "kind": pool.apply(_kind),
                                "sample": sample + " (%.1f s)" % dt, "one_process_value": one}

    # ======================================================== configs[1]: rank-3 subspace, one day
    if "cfg1" in sections:
        nday = min(24, args.chunks)
        U3 = synth.random_basis(np.random.default_rng(2002), N_MUX, 3)
        eng.set_bases(2, [U3], NC, thresholds=[0.25])

        def step_cfg1():
            eng.attach_device_chunks(data.data_ptr(), offs_all[:nday], lens_all[:nday])
            eng.detect_run(2, engine="tcgen05", kblk=args.kblk, lta_window=int(5 * SR))
            c = eng.candidates()
            return c, eng.hist(2, reset=True)

        for _ in range(3):
            step_cfg1()
        nrep = 10
        ms1, wall1, (c1_res, h1_res), _ = timed(step_cfg1, nrep)
        k1 = eng.k1_ms()
        ts1 = nday * T_PER_CHUNK * world
        step1 = max(ms1 / 1e3, wall1) / nrep
        other_ms = max(1e-6, step1 * 1e3 - k1)
        hbm_peak = float(peaks.get("hbm_gbs", 6549.1))
        # K0 + K3 + LTA: raw chunk read (8 B x Nc per sample, three K0 passes share it through L2), split
        # planes + mu/invE written and read once, the dense DS row written by K1 and read once by K3
        hbm_bytes = nday * (LS * NC * 8 + 2 * (2 * NC * LS * 2 + 8 * T_PER_CHUNK) + 4 * T_PER_CHUNK)
        cfg1 = {"workload": "BASELINE configs[1]: 1 station x 3 ch x 100 Hz x 1 day (%d chunks of 3720 s), 1 subspace "
                            "of rank 3, n = 9000, one batch; per GPU" % nday,
                "value": ts1 / step1, "unit": UNIT, "ms_per_step": 1e3 * step1, "steps": nrep, "k1_ms": k1,
                "roofline": {"bound": "tensor", "achieved": 2.0 * N_MUX * 3 * T_PER_CHUNK * nday / (k1 * 1e-3) / 1e12,
                             "peak": float(peaks.get("bf16_tflops", 1639.1)), "unit": "TFLOP/s",
                             "note": "one 16-slot basis block holds the 3 vectors: 13/16 of the MMA rows are padding "
                                     "and 3 MMAs per product -> bounded by 3/16/3 = 0.0625 of peak; kernel-baseline "
                                     "configuration, latency / fill bound by design"},
                "hbm_part": {"kernels": "k0_stats + k0_norm + k0_split + k3_fast + lta (everything but K1)",
                             "ms": other_ms, "algorithmic_bytes": hbm_bytes,
                             "achieved_gbs": hbm_bytes / (other_ms * 1e-3) / 1e9, "peak_gbs": hbm_peak,
                             "frac": hbm_bytes / (other_ms * 1e-3) / 1e9 / hbm_peak}}
        cfg1["roofline"]["frac"] = cfg1["roofline"]["achieved"] / cfg1["roofline"]["peak"]
        if hnp is not None:
            def step_cfg1_e2e():
                # the day in three batches: the H2D of a batch overlaps the projection of the one before
                # (same calls as the headline's end-to-end step)
                eng.accumulate_begin(nday)
                nb1 = max(1, (nday + 2) // 3)
                for lo in range(0, nday, nb1):
                    eng.load_chunks([hnp[i] for i in range(lo, min(nday, lo + nb1))])
                    eng.detect_run(2, engine="tcgen05", kblk=args.kblk, lta_window=int(5 * SR))
                eng.rowstats()
                c = eng.candidates()
                h = eng.hist(2, reset=True)
                eng.accumulate_end()
                return c, h
            step_cfg1_e2e()
            ms1e, wall1e, (c1_e, h1_e), _ = timed(step_cfg1_e2e, nrep)
            assert len(c1_e) == len(c1_res) and np.array_equal(h1_e, h1_res), "cfg1: resident and host paths disagree"
            cfg1["e2e"] = {"value": ts1 / (max(ms1e / 1e3, wall1e) / nrep), "unit": UNIT,
                           "h2d_bytes_per_step": nday * L * 8, "d2h_bytes_per_step": nday * 8 + 400 * 8}
        if pool is not None:
            tasks = [(LS, 0, 3, 0)]          # subspaces 0..2 have ranks 1, 2, 3: 6 vectors ~ 2 rank-3 subspaces
            pool.map(_det_task, tasks)
            r = pool.apply(_det_task, ((LS, 0, 3, 0),))
            cfg1["cpu_baseline"] = {"value": r[0] / r[1], "unit": UNIT, "cores": 1, "kind": pool.apply(_kind),
                                    "sample": "1 chunk x 3 subspaces of rank 1, 2, 3 (same cost per value as 2 "
                                              "rank-3 subspaces), one process: the reference has one subspace to "
                                              "parallelise over here (%.1f s)" % r[1]}
        # the function-level drop-ins called once per (chunk, subspace) / per pair,
        # the way the reference's own loops would call them: latency paths, measured so that nobody has to guess
        from detex_b200 import construct as dconstruct
        from detex_b200 import detect as ddetect
        chunk_h = data[0].cpu().numpy()
        pair = synth.event_families(3003, 1, 2, CCX_NS, NC, max_shift=100)
        ddetect._MPXDS(chunk_h, 0, U3, None, NC, None, engine=eng)
        dconstruct._CCX2(None, None, pair[0], pair[1], "ENZ", "ENZ", engine=eng)
        t0 = time.perf_counter()
        for _ in range(5):
            ddetect._MPXDS(chunk_h, 0, U3, None, NC, None, engine=eng)
        t_mpx = (time.perf_counter() - t0) / 5
        t0 = time.perf_counter()
        for _ in range(20):
            dconstruct._CCX2(None, None, pair[0], pair[1], "ENZ", "ENZ", engine=eng)
        t_ccx = (time.perf_counter() - t0) / 20
        cfg1["dropin_latency"] = {
            "_MPXDS_ms_per_call": 1e3 * t_mpx, "_MPXDS_ts_per_s": T_PER_CHUNK / t_mpx,
            "_CCX2_ms_per_pair": 1e3 * t_ccx,
            "note": "detect._MPXDS(one 3720 s chunk, one rank-3 subspace) and construct._CCX2(one pair, n = 3000) through "
                    "the reference's per-call signatures: basis upload + 8.9 MB H2D + K0/K1/K3 + 1.5 MB D2H per call; the "
                    "batched seams above are what the throughput numbers use"}
        line["cfg1"] = cfg1

    # ===================================================================== configs[4]: FAS sweep
    if "fas" in sections:
        nfas = max(world, int(round(args.fas_chunks * args.chunks / float(CHUNKS_PER_STATION))))
        nloc = parallel.shard_range(nfas, rank, world)
        nloc = nloc[1] - nloc[0]
        # null segments: 3600 s windows cut from the resident station noise, clear of the planted events
        # (window [120 s, 3720 s] of chunk j, then window [90 s, 3690 s] of chunk j - 720)
        jj = np.arange(nloc, dtype=np.int64)
        offs_f = (jj % args.chunks) * L + np.where(jj < args.chunks, 12000, 9000) * NC
        lens_f = np.full(nloc, LS_FAS * NC, dtype=np.int64)
        fas_flops = 2.0 * N_MUX * sum(ranks) * T_FAS

        def fas_finish():
            st = eng.fas(0, reset=True)
            if world > 1:
                st = parallel.allreduce_sum(st)
            return st

        def step_fas():
            k1_ms.clear()
            c, hist = step_resident(args.engine, 0, offs_f, lens_f, hist_range=(-.01, 1.0), want_fas=True, lta=0)
            return hist, fas_finish()

        eng.set_bases(0, bases, NC)               # FAS: no thresholds (they are what it calibrates)
        step_fas()
        nrep = 2
        msf, wallf, (hist_f, st_f), _ = timed(step_fas, nrep)
        stepf = max(msf / 1e3, wallf) / nrep
        tsf = float(nfas) * T_FAS * args.nsub
        totk = sum(m for m, _ in k1_ms)
        fasd = {"workload": "BASELINE configs[4]: FAS null-space sweep, %d one-hour segments (3600 s x 3 ch x 100 Hz, "
                            "L = 1 080 000) x %d subspaces -> 400-bin histograms on linspace(-.01, 1, 401) + beta-fit "
                            "sufficient statistics; segments sharded over the GPUs (strong scaling)" % (nfas, args.nsub),
                "value": tsf / stepf, "unit": UNIT, "ms_per_step": 1e3 * stepf, "steps": nrep, "scaling": "strong",
                "segments": nfas, "hist_total": int(hist_f.sum()),
                "roofline": {"bound": "tensor", "achieved": fas_flops * nloc / (totk * 1e-3) / 1e12,
                             "peak": float(peaks.get("bf16_tflops_sustained", 1354.8)), "unit": "TFLOP/s",
                             "kernel": "k1_kernel (this rank's launches of the last step)"}}
        fasd["roofline"]["frac"] = fasd["roofline"]["achieved"] / fasd["roofline"]["peak"]
        if rank == 0:
            t0 = time.perf_counter()
            fits = [dfas.beta_fit_from_stats(*st_f[s]) for s in range(args.nsub)]
            fasd["beta_fit_s"] = time.perf_counter() - t0
            fasd["beta_a_rank1"], fasd["beta_b_rank1"] = float(fits[0][0]), float(fits[0][1])
            fasd["beta_a_rank8"], fasd["beta_b_rank8"] = float(fits[7 % args.nsub][0]), float(fits[7 % args.nsub][1])
        if hnp is not None:
            def step_fas_e2e():
                eng.accumulate_begin(nloc)
                for lo in range(0, nloc, args.batch):
                    hi = min(nloc, lo + args.batch)
                    eng.load_chunks([hnp[int(o // L)][int(o % L):int(o % L) + LS_FAS * NC] for o in offs_f[lo:hi]])
                    eng.detect_run(0, engine=args.engine, kblk=args.kblk, hist_range=(-.01, 1.0), want_fas=True)
                hist = eng.hist(0, reset=True)
                eng.accumulate_end()
                st = fas_finish()
                if world > 1:
                    hist = parallel.allreduce_sum(hist)
                return hist, st
            step_fas_e2e()
            msfe, wallfe, (hist_fe, _), _ = timed(step_fas_e2e, nrep)
            assert np.array_equal(hist_fe, hist_f), "FAS: resident and host paths disagree"
            fasd["e2e"] = {"value": tsf / (max(msfe / 1e3, wallfe) / nrep), "unit": UNIT,
                           "h2d_bytes_per_step": int(nloc) * LS_FAS * NC * 8,
                           "d2h_bytes_per_step": args.nsub * (400 + 5) * 8}
        if pool is not None:
            v, dt, one, sample, _ = cpu_detection(pool, workers, LS_FAS, True)
            fasd["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": workers, "kind": pool.apply(_kind),
                                    "sample": sample + " (%.1f s)" % dt, "one_process_value": one}
        eng.set_bases(0, bases, NC, thresholds=thr)
        line["fas"] = fasd

    # ================================================================== configs[2]: CCX matrix
    if "ccx" in sections:
        del data
        torch.cuda.empty_cache()
        N = args.ccx_events
        X = synth.event_families(3003, max(1, N // 64), 64, CCX_NS, NC, max_shift=100)[:N]
        n = X.shape[1]
        npair = N * (N - 1) // 2
        nlag = CCX_NS + 1
        Xp = eng.pinned_empty(X.shape, np.float64)
        Xp[:] = X
        out = (eng.pinned_empty((npair,), np.float64), eng.pinned_empty((npair,), np.int32),
               eng.pinned_empty((npair,), np.float64))

        if args.ccx_ds_gib > 0:
            eng.set_ccx_batch(512, args.ccx_ds_gib << 30)

        def step_ccx(root=None):
            return parallel.ccx_sharded(eng, Xp, NC, engine="tcgen05", out=out, root=root)

        step_ccx()
        nrep = 3

        def wall_of(root):
            t_c, res = [], None
            for _ in range(nrep):
                barrier()
                t0 = time.perf_counter()
                r = step_ccx(root)
                barrier()
                t_c.append(time.perf_counter() - t0)
                res = r if r is not None else res
            w = float(np.median(t_c))
            if world > 1:
                t = torch.tensor([w], device=dev, dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                w = float(t[0])
            return w, res

        wall_c, (cc, lag, sub) = wall_of(None)          # every rank ends up with the whole matrix on its host
        wall_root = wall_of(0)[0] if world > 1 else wall_c   # only rank 0 (the one that clusters) fetches it
        wall_host = wall_c
        hbuf = None
        if world > 1:
            # no gather: every GPU writes its rows into ONE page-locked host matrix all ranks have mapped
            try:
                hbuf = parallel.CcxHostBuffer(eng, N)
            except RuntimeError:          # /dev/shm too small on this box (all ranks agree): NCCL gather only
                hbuf = None
                wall_host = min(wall_c, wall_root)
        if hbuf is not None:

            def step_host():
                return parallel.ccx_sharded(eng, Xp, NC, engine="tcgen05", host=hbuf)

            step_host()
            t_h = []
            for _ in range(nrep):
                barrier()
                t0 = time.perf_counter()
                res_h = step_host()
                barrier()
                t_h.append(time.perf_counter() - t0)
            assert np.array_equal(res_h[0], cc) and np.array_equal(res_h[1], lag) and np.array_equal(res_h[2], sub), \
                "CCX: shared-host and gathered results disagree"
            t = torch.tensor([float(np.median(t_h))], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            wall_host = float(t[0])
            barrier()
            hbuf.close()
        # device-resident variant: waveforms already in HBM, results left in HBM (no PCIe, no pack)
        dX = torch.from_numpy(X).to(dev)
        slot_rows, nmax = parallel.ccx_slot_rows(N, world)
        mine = slot_rows.reshape(world, nmax)[rank]
        mine = mine[mine >= 0]
        d_cc = torch.zeros((nmax, N), dtype=torch.float64, device=dev)
        d_lag = torch.zeros((nmax, N), dtype=torch.int32, device=dev)
        d_sub = torch.zeros((nmax, N), dtype=torch.float64, device=dev)

        def step_ccx_dev():
            eng.ccx_device(None, NC, mine, d_cc.data_ptr(), d_lag.data_ptr(), d_sub.data_ptr(), engine="tcgen05",
                           x_device_ptr=dX.data_ptr(), shape=(N, n))
            eng.sync()

        step_ccx_dev()
        msd, walld, _, _ = timed(step_ccx_dev, nrep)
        step_d = max(msd / 1e3, walld) / nrep
        k1c = eng.k1_ms_history()              # the K1 launches (signal batches) of the last call
        k1_tot = float(np.sum(k1c))
        # this rank's share of the pair*lag work against its own K1 time
        my_pairs = float((N - 1 - mine.astype(np.int64)).sum())
        ccxd = {"workload": "BASELINE configs[2]: pairwise CCX of %d events x 3 ch x 10 s x 100 Hz (n = %d, %d lags, "
                            "%d pairs), template rows dealt over the GPUs" % (N, n, nlag, npair),
                "scaling": "strong", "steps": nrep,
                "value": npair * nlag / step_d, "unit": "pair*lags/s", "pairs_per_s": npair / step_d,
                "ms_per_step": 1e3 * step_d,
                "value_note": "waveforms resident in HBM, results left in HBM (dense dealt rows); no gather, no pack",
                "e2e": {"value": npair * nlag / wall_host, "unit": "pair*lags/s", "pairs_per_s": npair / wall_host,
                        "ms_per_step": 1e3 * wall_host, "h2d_bytes_per_step": int(X.nbytes) // world,
                        "d2h_bytes_per_step": int(npair * 20) // world,
                        "ms_per_step_nccl_gather_every_rank": 1e3 * wall_c,
                        "ms_per_step_nccl_gather_rank0_only": 1e3 * wall_root,
                        "note": "host X (pinned) in, SciPy-condensed cc f64 / lag i32 / subsamp f64 on the host out. "
                                "N > 1: every rank uploads 1/N of X (all-gathered over NVLink) and writes the rows it "
                                "computed straight into ONE page-locked host matrix all ranks have mapped "
                                "(parallel.CcxHostBuffer, each GPU over its own PCIe link, no collective); bytes are "
                                "per rank.  The two NCCL variants: dense blocks all-gathered, then packed and copied "
                                "to the host by every rank / by rank 0 only"},
                "roofline": {"bound": "tensor", "kernel": "k1_kernel<128,1> (this rank's launches)",
                             "achieved": 2.0 * n * my_pairs * nlag / (k1_tot * 1e-3) / 1e12,
                             "peak": float(peaks.get("bf16_tflops", 1639.1)), "unit": "TFLOP/s",
                             "k1_ms_per_call": k1_tot, "launches_per_call": int(len(k1c)),
                             "note": "algorithmic flops = 2*n per pair*lag; 3 fp16 MMAs per product -> bounded by 1/3; "
                                     "burst peak (launches of a few ms)"},
                "max_cc": float(np.max(cc)), "gpu_ms_other_than_k1": 1e3 * step_d - k1_tot}
        ccxd["roofline"]["frac"] = ccxd["roofline"]["achieved"] / ccxd["roofline"]["peak"]
        if pool is not None:
            nev = 256
            v, dt, one, pairs = cpu_ccx(pool, workers, nev)
            ccxd["cpu_baseline"] = {"value": v * nlag, "unit": "pair*lags/s", "pairs_per_s": v, "cores": workers,
                                    "kind": pool.apply(_kind), "one_process_pairs_per_s": one,
                                    "sample": "construct._CCX2 (construct.py:425-466) on all %d pairs of %d events "
                                              "(MPfd prepared once per event as construct.py:669-676), Pool(%d), "
                                              "%.1f s; the cost per pair does not depend on N, so the whole matrix "
                                              "extrapolates quadratically: x %.0f" % (pairs, nev, workers, dt,
                                                                                    npair / float(pairs))}
        line["ccx"] = ccxd

    if pool is not None:
        pool.close()
        pool.join()
    if rank == 0:
        print(json.dumps(line))
    eng.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def workload_config(args):
    return {
        "workload": "BASELINE configs[3] shard: 1 station x 3 ch x 100 Hz x %d chunks of 3720 s x %d subspaces "
                    "(rank 1-8, n=9000) per GPU; N=8 is configs[3]" % (args.chunks, args.nsub),
        "chunks_per_gpu": args.chunks, "subspaces": args.nsub, "basis_vectors": sum(ranks_list(args.nsub)),
        "n": N_MUX, "lags_per_chunk": T_PER_CHUNK, "batch_chunks": args.batch,
        "l2": "inputs (%.1f GB/GPU) larger than L2" % (args.chunks * LS * NC * 8 / 1e9),
        "input_dtype": "f64", "kblk": args.kblk or 3, "engine": args.engine, "fused_epilogue": bool(args.fused),
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--chunks", type=int, default=CHUNKS_PER_STATION, help="chunks per GPU (720 = 30 days)")
    ap.add_argument("--nsub", type=int, default=NSUB)
    ap.add_argument("--batch", type=int, default=48, help="chunks per detect_run (DS buffer = batch*S*T*4 B)")
    ap.add_argument("--kblk", type=int, default=0, help="64-tap stages per TMEM accumulation (0 = library default 3)")
    ap.add_argument("--fused", action="store_true", help="fused mode: K1 does the row reductions, DS is never written")
    ap.add_argument("--engine", default="tcgen05", choices=["tcgen05", "tcgen05_x8", "tcgen05_auto"])
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-alt", action="store_true", help="skip the extra pass with the adaptive-precision engine")
    ap.add_argument("--sections", default="main,cfg1,fas,ccx",
                    help="main = configs[3] headline, cfg1 = configs[1], fas = configs[4], ccx = configs[2]")
    ap.add_argument("--fas-chunks", type=int, default=1000, help="null segments of the FAS sweep (at 720 chunks)")
    ap.add_argument("--ccx-events", type=int, default=CCX_EVENTS)
    ap.add_argument("--ccx-ds-gib", type=int, default=0,
                    help="correlation-series buffer of one CCX signal batch in GiB (0 = library default 16)")
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_arm(args)
    return gpu_arm(args)


if __name__ == "__main__":
    sys.exit(main())
