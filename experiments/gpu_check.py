"""First-contact GPU check: parity of every kernel against the oracle + first timings.
Run on the GPU box:  python experiments/gpu_check.py
"""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from detex_b200 import synth  # noqa: E402
from detex_b200.engine import Engine  # noqa: E402
from oracle import detex_oracle as orc  # noqa: E402


def section(s):
    print("\n==== " + s, flush=True)


def main():
    eng = Engine(0)
    # ------------------------------------------------------------ small parity
    section("small parity (Nc=3, ns=300, ranks 1,3,5,8,2)")
    Nc, ns, Ls = 3, 300, 9000
    ranks = [1, 3, 5, 8, 2]
    chunks, bases, truth = synth.detection_case(11, 3, Ls, ns, Nc, ranks, planted=3)
    chunks[1] = chunks[1][: (Ls - 37) * Nc]  # ragged chunk
    chunks[2] = chunks[2] + 1000.0            # big DC level
    thr = [0.3] * len(ranks)
    eng.set_bases(0, bases, Nc, thresholds=thr)
    eng.load_chunks(chunks)
    ref = [[orc.mpx_ds_direct(c, U, Nc) for U in bases] for c in chunks]
    for engine, kblk in (("fp64", 0), ("tcgen05", 1), ("tcgen05", 2), ("tcgen05", 4)):
        eng.detect_run(0, engine=engine, kblk=kblk, keep_ds64=(engine == "fp64"), lta_window=50)
        worst = 0.0
        for ci in range(len(chunks)):
            for si in range(len(bases)):
                ds = eng.get_ds(ci, si)
                assert ds.shape == ref[ci][si].shape, (ds.shape, ref[ci][si].shape)
                worst = max(worst, np.abs(ds - ref[ci][si]).max())
        print("engine=%s kblk=%d max|DS-oracle|=%.3e  maxDS=%.3f" % (engine, kblk, worst, max(r.max() for rr in ref for r in rr)))
        if engine == "fp64":
            w64 = max(np.abs(eng.get_ds64(ci, si) - ref[ci][si]).max() for ci in range(3) for si in range(len(bases)))
            print("   fp64 engine double output max err %.3e" % w64)
    mx, fl = eng.rowstats()
    refmax = np.array([[r.max() for r in rr] for rr in ref])
    print("rowmax err", np.abs(mx - refmax).max(), "flags", fl.ravel())
    h = eng.hist(0, reset=True)
    # 4 runs accumulated
    href = np.array([sum(orc.ds_histogram(ref[ci][si]) for ci in range(3)) for si in range(len(bases))])
    print("hist total", h.sum(), "expected", 4 * href.sum(), "L1 diff vs 4*ref", np.abs(h - 4 * href).sum())
    cand = eng.candidates()
    nref = sum(int((ref[ci][si] >= thr[si]).sum()) for ci in range(3) for si in range(len(bases)) if ref[ci][si].max() > thr[si])
    print("candidates", len(cand), "expected", nref)
    if len(cand):
        c = cand[0]
        ci, si = divmod(int(c["row"]), len(bases))
        lt = orc._replace_nan_with_mean(orc._rolling_mean_centered(np.abs(ref[ci][si]), 50))
        print("   cand0", c, "ref ds", ref[ci][si][c["t"]], "ref lta", lt[c["t"]])

    # ------------------------------------------------------------ FAS stats
    section("FAS stats")
    eng.detect_run(0, engine="tcgen05", kblk=1, hist_range=(-0.01, 1.0), want_fas=True)
    f = eng.fas(0, reset=True)
    for si in range(len(bases)):
        dss = np.concatenate([ref[ci][si] for ci in range(3)])
        st = orc.beta_sufficient_stats(dss)
        print("   s%d gpu" % si, f[si], "rel err", np.abs((f[si] - np.array(st)) / np.array(st)).max())
    hf = eng.hist(0, reset=True)
    hfr = np.array([np.histogram(np.concatenate([ref[ci][si] for ci in range(3)]), bins=orc.FAS_BINS)[0] for si in range(len(bases))])
    print("   FAS hist L1 diff", np.abs(hf - hfr).sum())

    # ------------------------------------------------------------ CCX
    section("CCX parity")
    X = synth.event_families(5, 3, 4, 200, 3, max_shift=20)
    cc, lag, sub = eng.ccx(X, 3)
    rcc, rlag, rsub = orc.make_cclags(X, 3, fft=False)
    N = X.shape[0]
    iu = np.triu_indices(N, 1)
    print("cc err", np.abs(cc[iu] - rcc[iu[0], iu[1] - 1]).max(), "lag mismatches", int((lag[iu] != rlag[iu[0], iu[1] - 1]).sum()),
          "subsamp err", np.nanmax(np.abs(sub[iu] - rsub[iu[0], iu[1] - 1])))

    # ------------------------------------------------------------ cfg2-like timing + full-size parity
    section("cfg2 shape: 3ch x 100Hz, chunk 3720 s, rank-3, n=9000; 4 chunks")
    Nc, ns, Ls = 3, 3000, 372000
    chunks, bases, _ = synth.detection_case(2002, 4, Ls, ns, Nc, [3], planted=2)
    eng.set_bases(1, bases, Nc, thresholds=[0.25])
    eng.load_chunks(chunks)
    for kblk in (1, 2, 4):
        eng.detect_run(1, engine="tcgen05", kblk=kblk, keep_ds64=(kblk == 1))
        eng.sync()
        ms = eng.k1_ms()
        if kblk == 1:
            ds64 = [eng.get_ds64(ci, 0) for ci in range(4)]
            t0 = time.time()
            o = orc.mpx_ds_fft(chunks[0], bases[0], Nc)
            tor = time.time() - t0
            print("   fp64 GPU engine vs oracle(fft) chunk0: %.3e (oracle %.2fs)" % (np.abs(ds64[0] - o).max(), tor))
        err = max(np.abs(eng.get_ds(ci, 0) - ds64[ci]).max() for ci in range(4))
        T = eng.num_lags(0)
        print("   kblk=%d K1 %.3f ms  -> %.3e template-samples/s ; max|DS - fp64| = %.3e (maxDS %.3f)" % (
            kblk, ms, 4 * T / (ms * 1e-3), err, max(d.max() for d in ds64)))

    # ------------------------------------------------------------ cfg4-like shard timing
    section("cfg4 shape: 256 subspaces ranks 1..8 (R=1152), 4 chunks")
    ranks = [(i % 8) + 1 for i in range(256)]
    rng = np.random.default_rng(4004)
    n = ns * Nc
    t0 = time.time()
    bases = [synth.random_basis(rng, n, r) for r in ranks]
    print("   bases built in %.1fs" % (time.time() - t0))
    eng.set_bases(2, bases, Nc, thresholds=[0.25] * 256)
    eng.load_chunks(chunks)
    for kblk in (1, 2, 4):
        eng.detect_run(2, engine="tcgen05", kblk=kblk)
        eng.sync()
        ms = eng.k1_ms()
        T = eng.num_lags(0)
        ts = 4 * T * 256
        flops = 2.0 * n * sum(ranks) * 4 * T
        print("   kblk=%d K1 %.1f ms -> %.3e template-samples/s, useful %.1f TFLOP/s" % (kblk, ms, ts / (ms * 1e-3), flops / (ms * 1e-3) / 1e12))
    # spot parity against the oracle on a few subspaces of chunk 0
    for si in (0, 7, 100, 255):
        o = orc.mpx_ds_fft(chunks[0], bases[si], Nc)
        print("   subspace %d rank %d: max|DS-oracle| = %.3e" % (si, ranks[si], np.abs(eng.get_ds(0, si) - o).max()))
    print("\nALL DONE")


if __name__ == "__main__":
    main()
