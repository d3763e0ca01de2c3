"""Timings of the other BASELINE.json configs on one B200 (parity of these shapes is in tests/):
  configs[1]  1 station x 3 ch x 100 Hz x 1 day (24 chunks), 1 subspace of rank 3
  configs[2]  CCX matrix of 4096 events x 3 ch x 10 s (one GPU does all row blocks here)
  configs[4]  FAS sweep: 1000 null chunks of 3600 s x 256 subspaces -> histograms + beta fit
Writes one JSON line per config.  Run: python experiments/config_runs.py [--fas-chunks N]
"""
import argparse
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import bench  # noqa: E402  (data generator + constants)
from detex_b200 import fas, synth  # noqa: E402
from detex_b200.engine import Engine  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--fas-chunks", type=int, default=1000)
    ap.add_argument("--ccx-events", type=int, default=4096)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    eng = Engine(0, stream=torch.cuda.current_stream().cuda_stream)
    rng = np.random.default_rng(2002)

    # ---------------------------------------------------------------- configs[1]
    data = bench.make_station_data(torch, dev, 24, seed=2002)
    U = synth.random_basis(rng, bench.N_MUX, 3)
    eng.set_bases(1, [U], bench.NC, thresholds=[0.25])
    L = bench.LS * bench.NC
    offs, lens = np.arange(24, dtype=np.int64) * L, np.full(24, L, dtype=np.int64)
    for rep in range(4):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        eng.attach_device_chunks(data.data_ptr(), offs, lens)
        eng.detect_run(1, lta_window=500)
        eng.candidates()
        dt = time.perf_counter() - t0
    ts = 24 * bench.T_PER_CHUNK
    print(json.dumps({"config": "configs[1]: 1 station x 3ch x 100Hz x 1 day, 1 subspace rank 3", "seconds": dt,
                      "k1_ms": eng.k1_ms(), "template_samples_per_s": ts / dt,
                      "note": "13 of 16 vector slots are padding (one 16-slot block): latency/fill bound by design"}))
    del data

    # ---------------------------------------------------------------- configs[2]
    N = args.ccx_events
    X = synth.event_families(3003, max(1, N // 64), 64, 1000, 3, max_shift=100)[:N]
    for rep in range(2):
        t0 = time.perf_counter()
        cc, lag, sub = eng.ccx(X, 3, engine="tcgen05")
        dt = time.perf_counter() - t0
    pairs = N * (N - 1) // 2
    print(json.dumps({"config": "configs[2]: CCX %d events x 3ch x 10 s x 100 Hz, 1001 lags" % N, "seconds": dt,
                      "pairs_per_s": pairs / dt, "pair_lags_per_s": pairs * 1001 / dt,
                      "useful_tflops": pairs * 1001 * 2 * 3000 / dt / 1e12, "max_cc": float(np.nanmax(cc))}))

    # ---------------------------------------------------------------- configs[4]
    nch, B = args.fas_chunks, 40
    ranks = bench.ranks_list(256)
    bases = [synth.random_basis(rng, bench.N_MUX, r) for r in ranks]
    eng.set_bases(5, bases, bench.NC)
    LS5 = 360000                                  # 3600 s null chunks
    L5 = LS5 * bench.NC
    old = bench.LS
    bench.LS = LS5
    eng.hist(5, reset=True); eng.fas(5, reset=True)
    torch.cuda.synchronize()
    tgen = 0.0
    t0 = time.perf_counter()
    for b0 in range(0, nch, B):
        nb = min(B, nch - b0)
        tg = time.perf_counter()
        data = bench.make_station_data(torch, dev, nb, seed=5005 + b0)   # surrogate noise (plumbing, excluded)
        torch.cuda.synchronize()
        tgen += time.perf_counter() - tg
        eng.attach_device_chunks(data.data_ptr(), np.arange(nb, dtype=np.int64) * L5, np.full(nb, L5, dtype=np.int64))
        eng.detect_run(5, hist_range=(-.01, 1.0), want_fas=True)
        eng.sync()
    hist = eng.hist(5, reset=True)
    st = eng.fas(5, reset=True)
    fits = [fas.beta_fit_from_stats(*st[s]) for s in range(256)]
    dt = time.perf_counter() - t0 - tgen
    bench.LS = old
    T5 = LS5 - bench.NS + 1
    print(json.dumps({"config": "configs[4]: FAS %d null chunks of 3600 s x 256 subspaces" % nch, "seconds": dt,
                      "template_samples_per_s": nch * T5 * 256 / dt, "hist_total": int(hist.sum()),
                      "beta_a_rank1": fits[0][0], "beta_b_rank1": fits[0][1], "beta_a_rank8": fits[7][0],
                      "beta_b_rank8": fits[7][1]}))
    eng.close()


if __name__ == "__main__":
    main()
