"""Parity of the CUDA detection path (through the C ABI) against the oracle and the
reference-generated golden vectors.  Tolerance: |DS - reference| <= 1e-5 absolute
(BASELINE.json north_star); integer work (histogram, candidate sets, lags) exact on the
statistic the GPU produced."""
import numpy as np
import pytest

from detex_b200 import synth
from detex_b200.engine import ShortChunk
from oracle import detex_oracle as orc

pytestmark = pytest.mark.gpu
TOL = 1e-5


def _golden_case(g, case):
    Nc = int(g[case + "_Nc"])
    chunks = [g["%s_chunk%d" % (case, i)] for i in range(int(g[case + "_nchunks"]))]
    bases = [g["%s_U%d" % (case, i)] for i in range(int(g[case + "_nbases"]))]
    return Nc, chunks, bases


@pytest.mark.parametrize("case", ["nc3", "nc1", "nc2odd"])
@pytest.mark.parametrize("kblk", [1, 2])
def test_tcgen05_matches_reference_golden(engine, ds_golden, case, kblk):
    Nc, chunks, bases = _golden_case(ds_golden, case)
    engine.set_bases(10, bases, Nc)
    engine.load_chunks(chunks)
    engine.detect_run(10, engine="tcgen05", kblk=kblk)
    for ci in range(len(chunks)):
        for si in range(len(bases)):
            ref = ds_golden["%s_DS_%d_%d" % (case, ci, si)]
            ds = engine.get_ds(ci, si)
            assert ds.shape == ref.shape
            assert np.abs(ds - ref).max() < TOL, (case, ci, si)


@pytest.mark.parametrize("case", ["nc3", "nc2odd"])
def test_fp64_engine_matches_reference_golden(engine, ds_golden, case):
    Nc, chunks, bases = _golden_case(ds_golden, case)
    engine.set_bases(11, bases, Nc)
    engine.load_chunks(chunks)
    engine.detect_run(11, engine="fp64", keep_ds64=True)
    for ci in range(len(chunks)):
        for si in range(len(bases)):
            ref = ds_golden["%s_DS_%d_%d" % (case, ci, si)]
            tol = 1e-9 if ci == 1 else 1e-11  # chunk 1 carries a DC level: reference FFT round-off
            assert np.abs(engine.get_ds64(ci, si) - ref).max() < tol


def test_singleton_nonzero_mean_template(engine, ds_golden):
    g = ds_golden
    engine.set_bases(12, [g["single_U"]], 3)
    engine.load_chunks([g["single_chunk"]])
    engine.detect_run(12)
    assert np.abs(engine.get_ds(0, 0) - g["single_DS"]).max() < TOL


def test_float32_input(engine, ds_golden):
    Nc, chunks, bases = _golden_case(ds_golden, "nc1")
    engine.set_bases(13, bases, Nc)
    engine.load_chunks([c.astype(np.float32) for c in chunks])
    engine.detect_run(13)
    for si, U in enumerate(bases):
        ref = orc.mpx_ds_direct(chunks[0].astype(np.float32).astype(np.float64), U, Nc)
        assert np.abs(engine.get_ds(0, si) - ref).max() < TOL


def test_rank16_packing_and_ragged_chunks(engine):
    Nc, ns, Ls = 3, 200, 5000
    ranks = [16, 1, 9, 7, 8, 8, 3, 2, 5, 11, 4]
    chunks, bases, _ = synth.detection_case(21, 3, Ls, ns, Nc, ranks, planted=3)
    chunks[1] = chunks[1][:(Ls - 411) * Nc + 2]  # ragged, not a multiple of Nc
    chunks[2] = chunks[2][:2100 * Nc]            # fewer lags than one tile
    engine.set_bases(14, bases, Nc)
    engine.load_chunks(chunks)
    engine.detect_run(14)
    for ci, c in enumerate(chunks):
        c = c[:len(c) // Nc * Nc]
        for si, U in enumerate(bases):
            ref = orc.mpx_ds_direct(c, U, Nc)
            ds = engine.get_ds(ci, si)
            assert ds.shape == ref.shape
            assert np.abs(ds - ref).max() < TOL, (ci, si)


def test_long_template_spans_several_k_segments(engine):
    Nc, ns, Ls = 1, 3500, 9000     # ns + 7 > 3072 -> two K segments
    chunks, bases, _ = synth.detection_case(22, 1, Ls, ns, Nc, [2, 3], planted=1)
    engine.set_bases(15, bases, Nc)
    engine.load_chunks(chunks)
    engine.detect_run(15)
    for si, U in enumerate(bases):
        assert np.abs(engine.get_ds(0, si) - orc.mpx_ds_direct(chunks[0], U, Nc)).max() < TOL


def test_short_chunk_is_rejected(engine):
    chunks, bases, _ = synth.detection_case(23, 1, 1000, 300, 3, [2])
    engine.set_bases(16, bases, 3)
    engine.load_chunks([chunks[0][:900]])        # L == n
    with pytest.raises(ShortChunk):
        engine.detect_run(16)
    engine.load_chunks([chunks[0][:900 + 3 * 8]])  # 9 lags < 10 (detect.py:270)
    with pytest.raises(ShortChunk):
        engine.detect_run(16)
    assert orc.chunk_ds(chunks[0][:900], bases[0], 3) is None


def test_large_dynamic_range_spike(engine):
    """An earthquake-sized transient (1e5 x noise) must not destroy the statistic of the
    windows around it (fp16 split with per-chunk power-of-two scaling)."""
    Nc, ns, Ls = 3, 300, 6000
    chunks, bases, _ = synth.detection_case(24, 1, Ls, ns, Nc, [3, 5], planted=2)
    x = chunks[0]
    x[9000:9030] += 1e5 * np.hanning(30)
    engine.set_bases(17, bases, Nc)
    engine.load_chunks([x])
    engine.detect_run(17)
    for si, U in enumerate(bases):
        assert np.abs(engine.get_ds(0, si) - orc.mpx_ds_direct(x, U, Nc)).max() < TOL


def test_constant_run_gives_inf_that_is_zeroed(engine):
    """A window of constant data has zero energy: the reference's statistic there is sum(if1^2)/0 =
    +inf (detect.py:577), MaxDS = inf > 1.1 so the infs are zeroed (detect.py:275-281) and the rest of
    the chunk is used as usual.  (Before round 2 the exactly-zero projection made these windows NaN
    and the whole row was lost.)  Reference golden: tests/test_gpu_scale.py::test_zero_filled_gap..."""
    Nc, ns, Ls = 1, 100, 3000
    _, bases, _ = synth.detection_case(25, 1, Ls, ns, Nc, [2])
    # integer-valued samples with an exactly-zero sum: centring is exact in every implementation
    x = np.random.default_rng(25).integers(-50, 51, size=Ls).astype(np.float64)
    x[1000:1400] = 0.0
    x[0] -= x.sum()
    assert x.sum() == 0.0
    engine.set_bases(18, bases, Nc, thresholds=[0.2])
    engine.hist(18, reset=True)
    engine.load_chunks([x])
    engine.detect_run(18, lta_window=50)
    mx, fl = engine.rowstats()
    ref = orc.mpx_ds_fft(x, bases[0], Nc)                 # the reference's algorithm: inf in the run
    bad = ~np.isfinite(ref)
    assert bad.sum() == 301          # (+inf, or NaN where the FFT round-off happens to be exactly 0)
    ds = engine.get_ds(0, 0)
    assert np.array_equal(np.isinf(ds), bad) and not np.isnan(ds).any()
    assert np.abs(ds[~bad] - ref[~bad]).max() < TOL
    assert fl[0, 0] == 2 and abs(mx[0, 0] - ref[~bad].max()) < TOL
    assert engine.hist(18, reset=True).sum() == len(ds)
    cand = engine.candidates()
    assert np.isfinite(cand["ds"]).all() and not bad[cand["t"]].any()


def test_rowstats_histogram_candidates_lta(engine):
    Nc, ns, Ls = 3, 300, 9000
    ranks = [1, 3, 5, 8, 2]
    chunks, bases, _ = synth.detection_case(11, 3, Ls, ns, Nc, ranks, planted=3)
    thr = [0.3, 0.25, 0.4, 0.5, 0.2]
    engine.set_bases(19, bases, Nc, thresholds=thr)
    engine.hist(19, reset=True)
    engine.load_chunks(chunks)
    W = 50
    engine.detect_run(19, lta_window=W)
    mx, fl = engine.rowstats()
    cand = engine.candidates()
    hist = engine.hist(19, reset=True)
    S = len(bases)
    exp_hist = np.zeros((S, 400), dtype=np.int64)
    exp_cand = set()
    for ci in range(3):
        for si in range(S):
            ds = engine.get_ds(ci, si)
            assert mx[ci, si] == ds.max() and fl[ci, si] == 0           # exact on the GPU's own DS
            exp_hist[si] += np.histogram(ds.astype(np.float64), bins=orc.HIST_BINS)[0]
            if ds.max() > np.float32(thr[si]):
                for t in np.nonzero(ds >= np.float32(thr[si]))[0]:
                    exp_cand.add((ci * S + si, int(t)))
            assert abs(mx[ci, si] - orc.mpx_ds_direct(chunks[ci], bases[si], Nc).max()) < TOL
    assert np.array_equal(hist, exp_hist)
    assert set((int(c["row"]), int(c["t"])) for c in cand) == exp_cand and len(cand) == len(exp_cand)
    for c in cand[:50]:
        ci, si = divmod(int(c["row"]), S)
        ds = engine.get_ds(ci, si).astype(np.float64)
        lta = orc._replace_nan_with_mean(orc._rolling_mean_centered(np.abs(ds), W))
        assert abs(c["lta"] - lta[c["t"]]) < 1e-6
        assert c["ds"] == np.float32(ds[c["t"]])
        sl = engine.get_stalta(ci, si, W)
        assert np.abs(sl - orc.sta_lta(ds, W, 0)).max() < 1e-3 * np.abs(sl).max()


@pytest.mark.parametrize("Wsta,W", [(7, 50), (20, 21), (64, 500)])
def test_trigger_sta_window(engine, Wsta, W):
    """triggerSTATime != 0 (`_getStaLtaArray`, detect.py:501-515): STA is a centred rolling mean of
    |DS| with the same edge rule as the LTA; sparse (candidates) and dense (CorDF.STALTA) paths."""
    Nc, ns, Ls = 3, 200, 5000
    chunks, bases, _ = synth.detection_case(29, 2, Ls, ns, Nc, [2, 4], planted=3)
    engine.set_bases(23, bases, Nc, thresholds=[0.2, 0.25])
    engine.load_chunks(chunks)
    engine.set_trigger_sta(Wsta)
    try:
        engine.detect_run(23, lta_window=W)
        cand = engine.candidates()
        assert len(cand) > 0
        S = len(bases)
        seen = set()
        for c in cand:
            ci, si = divmod(int(c["row"]), S)
            ds = engine.get_ds(ci, si).astype(np.float64)
            ref = orc.sta_lta(ds, W, Wsta)
            assert abs(abs(float(c["ds"])) / float(c["lta"]) - ref[c["t"]]) < 1e-5 * max(1.0, abs(ref[c["t"]]))
            if (ci, si) not in seen:
                seen.add((ci, si))
                sl = engine.get_stalta(ci, si, W).astype(np.float64)
                assert np.abs(sl - ref).max() < 1e-5 * np.abs(ref).max()
    finally:
        engine.set_trigger_sta(0)
    # back to the default: STA = |DS|
    engine.detect_run(23, lta_window=W)
    c = engine.candidates()[0]
    ci, si = divmod(int(c["row"]), len(bases))
    ds = engine.get_ds(ci, si).astype(np.float64)
    assert abs(abs(float(c["ds"])) / float(c["lta"]) - orc.sta_lta(ds, W, 0)[c["t"]]) < 1e-5


def test_full_size_chunk_properties(engine):
    """BASELINE config 2 shape (3 ch x 100 Hz x 3720 s, rank 3, n = 9000): tcgen05 against the
    independent float64 evaluation on the device, plus shift/scale invariance of DS."""
    Nc, ns, Ls = 3, 3000, 372000
    chunks, bases, truth = synth.detection_case(2002, 2, Ls, ns, Nc, [3], planted=2)
    engine.set_bases(20, bases, Nc, thresholds=[0.25])
    engine.load_chunks(chunks)
    engine.detect_run(20, keep_ds64=True)
    ds = [engine.get_ds(ci, 0) for ci in range(2)]
    for ci in range(2):
        d64 = engine.get_ds64(ci, 0)
        assert d64.shape == (Ls - ns + 1,)
        assert np.abs(ds[ci] - d64).max() < TOL
        assert d64.max() <= 1.0 + 1e-12 and d64.min() >= 0.0
    o = orc.mpx_ds_fft(chunks[0], bases[0], Nc)
    assert np.abs(engine.get_ds64(0, 0) - o).max() < 1e-10
    for (ci, si, t) in truth:
        assert ds[ci][t] > 0.25                                     # planted events are found
    engine.load_chunks([c * 37.5 - 1234.5 for c in chunks])          # DS(a x + b) == DS(x)
    engine.detect_run(20)
    for ci in range(2):
        assert np.abs(engine.get_ds(ci, 0) - ds[ci]).max() < 2 * TOL


def test_rank_above_16_is_split_and_accumulated(engine):
    Nc, ns, Ls = 3, 200, 5000
    ranks = [37, 3, 16, 17, 1]
    chunks, bases, _ = synth.detection_case(26, 2, Ls, ns, Nc, ranks, planted=2)
    engine.set_bases(21, bases, Nc)
    engine.load_chunks(chunks)
    for eng_name in ("tcgen05", "fp64"):
        engine.detect_run(21, engine=eng_name)
        for ci, c in enumerate(chunks):
            for si, U in enumerate(bases):
                assert np.abs(engine.get_ds(ci, si) - orc.mpx_ds_direct(c, U, Nc)).max() < TOL, (eng_name, ci, si)


def test_short_chunks_use_the_1024_lag_tile(engine):
    """Chunks with <= 1024 lags run the N = 128 variant of the Hankel GEMM (tiles of 1024 lags)."""
    Nc, ns = 3, 300
    ranks = [2, 5, 8]
    chunks, bases, _ = synth.detection_case(27, 3, 1200, ns, Nc, ranks, planted=1)
    chunks[1] = chunks[1][:(ns + 40) * Nc]      # 41 lags
    chunks[2] = chunks[2][:(ns + 1023) * Nc]    # exactly 1024 lags
    engine.set_bases(22, bases, Nc, thresholds=[0.3] * 3)
    engine.load_chunks(chunks)
    engine.detect_run(22, lta_window=20)
    for ci, c in enumerate(chunks):
        for si, U in enumerate(bases):
            ref = orc.mpx_ds_direct(c, U, Nc)
            ds = engine.get_ds(ci, si)
            assert ds.shape == ref.shape and np.abs(ds - ref).max() < TOL


def test_item_order_does_not_change_results(engine, monkeypatch):
    """K1's work-item order (chunk groups x superblocks of basis blocks, chosen by a DRAM-traffic model
    in project_run) is a pure reordering: every forced order gives bit-identical DS rows, histograms
    and candidates."""
    rng = np.random.default_rng(77)
    Nc, ns = 2, 96
    ranks = [16, 5, 11, 16, 3, 8, 8, 16, 1, 2, 13, 16, 7, 9, 16, 4]      # ~11 blocks of 16 slots
    bases = [synth.random_basis(rng, ns * Nc, r) for r in ranks]
    chunks = [synth.multiplex(synth.bandpassed_noise(rng, 5000 + 37 * i, nchan=Nc)) for i in range(5)]
    thr = [0.2] * len(ranks)
    engine.set_bases(31, bases, Nc, thresholds=thr)
    ref = None
    for group, sup in ((None, None), (1, 1), (5, 1), (2, 4), (5, 4), (5, 3)):
        if group is None:
            monkeypatch.delenv("DTX_K1_GROUP", raising=False)
            monkeypatch.delenv("DTX_K1_SUPER", raising=False)
        else:
            monkeypatch.setenv("DTX_K1_GROUP", str(group))
            monkeypatch.setenv("DTX_K1_SUPER", str(sup))
        engine.hist(31, reset=True)
        engine.load_chunks(chunks)
        engine.detect_run(31, lta_window=50)
        ds = [engine.get_ds(ci, si) for ci in range(len(chunks)) for si in (0, 3, 8, 15)]
        cand = np.sort(engine.candidates(), order=["row", "t"])
        out = (ds, engine.hist(31, reset=True), cand, engine.rowstats()[0])
        if ref is None:
            ref = out
            assert np.abs(ds[0] - orc.mpx_ds_direct(chunks[0], bases[0], Nc)).max() < TOL
        else:
            assert all(np.array_equal(a, b) for a, b in zip(out[0], ref[0]))
            assert np.array_equal(out[1], ref[1]) and np.array_equal(out[2], ref[2])
            assert np.array_equal(out[3], ref[3])


def test_cta_pair_kernel_matches_default(engine):
    """K1 with CTA pairs (`tcgen05.mma.cta_group::2`, M = 256: two basis blocks per MMA, the tile's B rows held
    half by each CTA; opt-in through DTX_K1_CG2=1, see profiles/r02_cta_pairs.md): same statistic as the default
    single-CTA kernel, incl. an odd number of basis blocks (zero padding block) and ragged chunks."""
    import os
    Nc, ns, Ls = 3, 300, 9000
    ranks = [16, 1, 9, 7, 8, 8, 3, 2, 5, 11, 4, 16, 6]            # 5 basis blocks (odd)
    chunks, bases, _ = synth.detection_case(27, 3, Ls, ns, Nc, ranks, planted=3)
    chunks[1] = chunks[1][:(Ls - 411) * Nc]
    engine.set_bases(19, bases, Nc, thresholds=[0.3] * len(ranks))
    out = {}
    for cg2 in ("0", "1"):
        old = os.environ.get("DTX_K1_CG2")
        os.environ["DTX_K1_CG2"] = cg2
        try:
            engine.hist(19, reset=True)
            engine.load_chunks(chunks)
            engine.detect_run(19, lta_window=50)
            out[cg2] = ([engine.get_ds(ci, si).copy() for ci in range(3) for si in range(len(ranks))],
                        np.sort(engine.candidates(), order=["row", "t"]), engine.hist(19, reset=True))
        finally:
            if old is None:
                os.environ.pop("DTX_K1_CG2", None)
            else:
                os.environ["DTX_K1_CG2"] = old
    for a, b in zip(out["0"][0], out["1"][0]):
        assert np.abs(a - b).max() < 1e-6
    assert np.array_equal(out["0"][1]["t"], out["1"][1]["t"]) and np.array_equal(out["0"][1]["row"], out["1"][1]["row"])
    assert np.abs(out["0"][2] - out["1"][2]).sum() <= 4
    for si, U in enumerate(bases):
        assert np.abs(out["1"][0][si] - orc.mpx_ds_direct(chunks[0], U, Nc)).max() < TOL
