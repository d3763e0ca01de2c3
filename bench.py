#!/usr/bin/env python
"""bench.py -- subspace-detector throughput (template*samples / s) on N B200s.

    python bench.py --gpus N --steps K --warmup W            # our arm
    python bench.py --impl reference --gpus N --steps K --warmup W   # CPU reference arm

Workload (weak scaling, one process per GPU): every GPU runs the detector of ONE station of
BASELINE.json configs[3] -- 3 channels x 100 Hz x 30 days as 720 chunks of 3720 s
(L = 1 116 000 multiplexed samples, T = 369 001 lags per chunk) against 256 subspaces of
rank 1..8 (R = 1152 basis vectors, n = 9000).  At N = 8 the job IS configs[3].
One step = one pass of the hot path (K0 prep -> K1 tcgen05 projection+normalisation -> K3
max / histogram / candidate compaction / LTA) over all 720 chunks of the station, plus the
end-of-step gather of trigger candidates and histogram all-reduce when N > 1.

`value`  : whole-job template*samples/s with the chunks already resident in HBM.
`e2e`    : the same through the C ABI with HOST (pinned) buffers: H2D of every chunk and
           D2H of MaxDS / candidates inside the timed region.
`roofline`: K1, tensor-bound: algorithmic flops (2*n*R per lag) / its CUDA-event duration.
`cpu_baseline`: the oracle port of the reference's FFT algorithm on the host cores, bounded
           sample, reported beside (not the optimisation target).
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
CHUNKS_PER_STATION = 720  # 30 days
NSUB = 256
T_PER_CHUNK = LS - NS + 1
METRIC = "subspace_detector_template_samples_per_sec"
UNIT = "template*samples/s"


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
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"),
                                 f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------- CPU baseline
def _cpu_one(args):
    from oracle import detex_oracle as orc
    chunk, U = args
    return len(orc.mpx_ds_fft(chunk, U, NC))


def cpu_sample(nsub_sample, seed=4004):
    """One bounded sample of the workload for the CPU arm: 1 chunk x the first
    `nsub_sample` subspaces (ranks cycling 1..8)."""
    from detex_b200 import synth
    rng = np.random.default_rng(seed)
    chunk = synth.multiplex(synth.bandpassed_noise(rng, LS, sr=SR, nchan=NC))
    bases = [synth.random_basis(rng, N_MUX, r) for r in ranks_list(nsub_sample)]
    return chunk, bases


def cpu_run(chunk, bases, cores, pool):
    t0 = time.perf_counter()
    tot = sum(pool.map(_cpu_one, [(chunk, U) for U in bases]))
    dt = time.perf_counter() - t0
    return tot / dt, dt


def reference_arm(args):
    """--impl reference: the reference's CPU algorithm (oracle port; the reference is pure
    Python and cannot travel to the GPU box) on all host cores, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import multiprocessing as mp
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    cores = os.cpu_count() or 1
    nsub = max(16, min(NSUB, 8 * cores))
    chunk, bases = cpu_sample(nsub)
    with mp.get_context("fork").Pool(cores) as pool:
        for _ in range(args.warmup):
            cpu_run(chunk, bases[:max(8, cores)], cores, pool)
        t0 = time.perf_counter()
        tot = 0
        for _ in range(args.steps):
            v, dt = cpu_run(chunk, bases, cores, pool)
            tot += len(bases) * T_PER_CHUNK
        el = time.perf_counter() - t0
    value = tot / el
    sample = "1 chunk (3720 s x 3 ch x 100 Hz) x %d subspaces (ranks 1-8, n=9000) per step" % nsub
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * el / max(1, args.steps),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": workload_config(args),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))
    return 0


def workload_config(args):
    return {
        "workload": "BASELINE configs[3] shard: 1 station x 3 ch x 100 Hz x %d chunks of 3720 s x %d subspaces "
                    "(rank 1-8, n=9000) per GPU; N=8 is configs[3]" % (args.chunks, args.nsub),
        "chunks_per_gpu": args.chunks, "subspaces": args.nsub, "basis_vectors": sum(ranks_list(args.nsub)),
        "n": N_MUX, "lags_per_chunk": T_PER_CHUNK, "batch_chunks": args.batch,
        "l2": "inputs (%.1f GB/GPU) larger than L2" % (args.chunks * LS * NC * 8 / 1e9),
        "input_dtype": "f64", "kblk": args.kblk, "engine": args.engine,
    }


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


def gpu_arm(args):
    import torch
    import torch.distributed as dist
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
    x8_chunks = [0]

    def step_resident(engine=args.engine):
        """One pass over the station with the chunks resident in HBM."""
        cands = []
        for b in range(nbatch):
            lo, hi = b * args.batch, min(args.chunks, (b + 1) * args.batch)
            eng.attach_device_chunks(data.data_ptr(), offs_all[lo:hi], lens_all[lo:hi])
            eng.detect_run(0, engine=engine, kblk=args.kblk, lta_window=int(5 * SR))
            c = eng.candidates()
            c["row"] += lo * args.nsub
            cands.append(c)
            k1_ms.append((eng.k1_ms(), hi - lo))
            if engine != "tcgen05":
                x8_chunks[0] += int(eng.chunk_modes().sum())
        c = np.concatenate(cands)
        hist = eng.hist(0, reset=True)
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
        if world > 1:
            t = torch.tensor([ms, wall * 1e3], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms, wall = float(t[0]), float(t[1]) / 1e3
        return ms, wall, out

    for _ in range(args.warmup):
        step_resident()
    k1_ms.clear()
    x8_chunks[0] = 0
    launches0 = eng.launch_count()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms, wall, (cands, hist) = timed(step_resident, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    launches = (eng.launch_count() - launches0) // max(1, args.steps)
    ts_per_step = float(args.chunks) * T_PER_CHUNK * args.nsub * world
    # device time (events) and wall time agree to <1 %: the host only enqueues; report the
    # slower of the two so the candidate D2H at the end of each batch is inside the number
    step_s = max(ms / 1e3, wall) / args.steps
    value = ts_per_step / step_s

    # K1 roofline (rank 0's launches in the timed region)
    tot_ms = sum(m for m, _ in k1_ms)
    tot_chunks = sum(n for _, n in k1_ms)
    k1_avg_ms = tot_ms / len(k1_ms)
    achieved = flops_per_chunk * tot_chunks / (tot_ms * 1e-3) / 1e12
    peak, peak_src = 1590.0, "fallback (B200_PROFILING.md)"
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        try:
            peak = float(json.load(open(pk))["bf16_tflops_sustained"])
            peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside a long step)"
        except Exception:
            pass
    traffic = None
    tj = os.path.join(ROOT, "profiles", "k1_traffic.json")
    if os.path.exists(tj):
        try:
            traffic = json.load(open(tj)).get("dram_bytes_per_launch")
        except Exception:
            pass
    roofline = {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                "traffic": traffic, "kernel": "k1_kernel (tcgen05 Hankel projection + normalisation)",
                "peak_source": peak_src, "k1_ms_per_launch": k1_avg_ms, "k1_share_of_step": tot_ms / 1e3 / (step_s * args.steps),
                "note": "algorithmic flops = 2*n*R per lag; fp32-equivalent precision costs 3 fp16 MMAs per "
                        "product, so frac is bounded by 1/3"}
    # tensor-pipe slots issued per algorithmic product: 3 fp16 MMAs, or 2 (fp16 + one 8-bit MMA of
    # twice the K) for the chunks that ran with 8-bit cross terms; K padded 3000 -> 3008 per channel
    x8_frac = x8_chunks[0] / float(args.chunks * args.steps)
    roofline["issued_tflops"] = achieved * (3.0 - x8_frac) * (3 * 3008.0 / N_MUX)
    if args.engine != "tcgen05":
        roofline["note"] = ("algorithmic flops = 2*n*R per lag; %.0f %% of the chunks ran with 8-bit cross terms "
                            "(2 tensor-pipe slots per product, frac bounded by 1/2), the rest with fp16 cross "
                            "terms (3 slots, 1/3)" % (100 * x8_frac))

    # ------------------------------------------------- parity spot check at full size (untimed)
    # the planted chunk of day 0 against the float64 closed form on the device, first 8 subspaces
    npar = min(8, args.nsub)
    eng.set_bases(1, bases[:npar], NC)

    def parity_check(engine):
        ci = min(args.chunks - 1, 7)
        eng.attach_device_chunks(data.data_ptr(), offs_all[ci:ci + 1], lens_all[ci:ci + 1])
        eng.detect_run(1, engine=engine, kblk=args.kblk, keep_ds64=True)
        perr, pmax = 0.0, 0.0
        for si in range(npar):
            d64 = eng.get_ds64(0, si)
            perr = max(perr, float(np.abs(eng.get_ds(0, si) - d64).max()))
            pmax = max(pmax, float(d64.max()))
        assert perr < 1e-5, "detection statistic out of tolerance against the float64 closed form: %g" % perr
        return {"max_abs_err_vs_fp64": perr, "tol": 1e-5, "max_ds": pmax, "chunk": ci, "subspaces": npar,
                "lags": T_PER_CHUNK, "x8_mode": int(eng.chunk_modes()[0])}

    parity = parity_check(args.engine)

    # -------------------- the opt-in adaptive-precision engine on the same resident data (reported
    # beside the headline, which stays on the worst-case-bounded default engine)
    alt = None
    if args.engine == "tcgen05" and not args.no_alt:
        x8_chunks[0] = 0
        k1_ms.clear()
        step_resident("tcgen05_auto")
        x8_chunks[0] = 0
        k1_ms.clear()
        ms_a, wall_a, (cands_a, hist_a) = timed(lambda: step_resident("tcgen05_auto"), args.steps)
        alt = {"engine": "tcgen05_auto", "dtype": DTYPES["tcgen05_auto"],
               "value": ts_per_step / (max(ms_a / 1e3, wall_a) / args.steps), "unit": UNIT,
               "x8_chunk_fraction": x8_chunks[0] / float(args.chunks * args.steps),
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
        cands = []
        for b in range(nbatch):
            lo, hi = b * args.batch, min(args.chunks, (b + 1) * args.batch)
            eng.load_chunks([hnp[i] for i in range(lo, hi)])                  # H2D (pinned)
            eng.detect_run(0, engine=args.engine, kblk=args.kblk, lta_window=int(5 * SR))
            mx, fl = eng.rowstats()                                            # D2H
            c = eng.candidates()                                               # D2H
            c["row"] += lo * args.nsub
            cands.append(c)
            d2h[0] += mx.nbytes + fl.nbytes + c.nbytes + 8
        c = np.concatenate(cands)
        hist = eng.hist(0, reset=True)
        d2h[0] += hist.nbytes
        if world > 1:
            c = parallel.gather_records(c)
            hist = parallel.allreduce_sum(hist)
        return c, hist

    step_e2e()
    d2h[0] = 0
    ms_e, wall_e, (cands_e, hist_e) = timed(step_e2e, args.steps)
    e2e_val = ts_per_step / (max(ms_e / 1e3, wall_e) / args.steps)
    assert len(cands_e) == len(cands) and np.array_equal(hist_e, hist), "resident and host paths disagree"

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * step_s, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": DTYPES[args.engine], "data": "synthetic",
        "config": workload_config(args),
        "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": int(args.chunks) * L * 8,
                "d2h_bytes_per_step": int(d2h[0] // max(1, args.steps))},
        "gpu_launches": int(launches), "roofline": roofline, "clocks": clocks,
        "candidates_per_step": int(len(cands)), "hist_total": int(hist.sum()), "parity_check": parity,
    }
    if args.engine != "tcgen05":
        line["x8_chunk_fraction"] = x8_frac
    if alt is not None:
        line["adaptive_engine"] = alt
    if rank == 0 and world == 1 and not args.no_cpu:
        import multiprocessing as mp
        os.environ.setdefault("OMP_NUM_THREADS", "1")
        cores = os.cpu_count() or 1
        nsub = max(16, min(NSUB, 8 * cores))
        chunk, cb = cpu_sample(nsub)
        with mp.get_context("spawn").Pool(cores) as pool:   # spawn: the parent holds a CUDA context
            cpu_run(chunk, cb[:cores], cores, pool)          # warm the workers (imports)
            v, dt = cpu_run(chunk, cb, cores, pool)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": "1 chunk x %d subspaces (%.1f s); oracle port of the reference's "
                                          "FFT algorithm, multiprocessing.Pool(%d)" % (nsub, dt, cores)}
    if rank == 0:
        print(json.dumps(line))
    eng.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--chunks", type=int, default=CHUNKS_PER_STATION, help="chunks per GPU (720 = 30 days)")
    ap.add_argument("--nsub", type=int, default=NSUB)
    ap.add_argument("--batch", type=int, default=48, help="chunks per detect_run (DS buffer = batch*S*T*4 B)")
    ap.add_argument("--kblk", type=int, default=2)
    ap.add_argument("--engine", default="tcgen05", choices=["tcgen05", "tcgen05_x8", "tcgen05_auto"])
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-alt", action="store_true", help="skip the extra pass with the adaptive-precision engine")
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_arm(args)
    return gpu_arm(args)


if __name__ == "__main__":
    sys.exit(main())
