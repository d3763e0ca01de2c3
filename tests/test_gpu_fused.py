"""Fused mode (dtx_set_fused): K1's own read-out produces MaxDS / flags / histograms / candidates / FAS
sums and the dense statistic is never written (reference loop detect.py:177-190).
Everything integer must equal the unfused path bit for bit; the candidates' LTA denominators come from a
float64 re-evaluation of the windows and agree to float rounding."""
import numpy as np
import pytest

from detex_b200 import synth
from detex_b200.engine import DtxError
from oracle import detex_oracle as orc

pytestmark = pytest.mark.gpu


def _run(engine, sid, fused, chunks, lta=50, want_fas=False, hist_range=(0.0, 1.0), core=None, engine_name="tcgen05"):
    engine.hist(sid, reset=True)
    engine.fas(sid, reset=True)
    engine.set_fused(fused)
    try:
        engine.load_chunks(chunks)
        if core is not None:
            engine.set_core_lags(*core)
        engine.detect_run(sid, engine=engine_name, lta_window=lta, want_fas=want_fas, hist_range=hist_range)
        mx, fl = engine.rowstats()
        cand = engine.candidates()
        if fused:
            with pytest.raises(DtxError):
                engine.get_ds(0, 0)
    finally:
        engine.set_fused(False)
    return mx, fl, np.sort(cand, order=["row", "t"]), engine.hist(sid, reset=True), engine.fas(sid, reset=True)


def _same(a, b, lta_rtol=2e-5):
    mx0, fl0, c0, h0, f0 = a
    mx1, fl1, c1, h1, f1 = b
    assert np.array_equal(mx0, mx1, equal_nan=True) and np.array_equal(fl0, fl1)
    assert np.array_equal(h0, h1)
    assert len(c0) == len(c1) and np.array_equal(c0["row"], c1["row"]) and np.array_equal(c0["t"], c1["t"])
    assert np.array_equal(c0["ds"], c1["ds"])
    ok = np.isfinite(c0["lta"]) & (c0["lta"] != 0)            # lta_window = 0 leaves the field at 0
    assert np.array_equal(np.isfinite(c0["lta"]), np.isfinite(c1["lta"]))
    if ok.any():
        assert np.abs(c1["lta"][ok] / c0["lta"][ok] - 1).max() < lta_rtol
    assert np.allclose(f0, f1, rtol=1e-6, atol=1e-9)


def test_fused_equals_unfused(engine):
    Nc, ns, Ls = 3, 300, 9000
    ranks = [1, 3, 5, 8, 2, 16, 4]
    chunks, bases, _ = synth.detection_case(81, 4, Ls, ns, Nc, ranks, planted=4)
    chunks[1] = chunks[1][:(Ls - 777) * Nc]                      # ragged: last tile partly empty
    engine.set_bases(80, bases, Nc, thresholds=[0.3] * len(ranks))
    a = _run(engine, 80, False, chunks, want_fas=True)
    b = _run(engine, 80, True, chunks, want_fas=True)
    assert len(a[2]) > 10 and a[3].sum() == sum((len(c) // Nc - ns + 1) for c in chunks) * len(ranks)
    _same(a, b)
    # FAS histogram range and the 8-bit cross-term engine go through the same read-out
    _same(_run(engine, 80, False, chunks, lta=0, want_fas=True, hist_range=(-.01, 1.0)),
          _run(engine, 80, True, chunks, lta=0, want_fas=True, hist_range=(-.01, 1.0)))
    _same(_run(engine, 80, False, chunks, engine_name="tcgen05_x8"), _run(engine, 80, True, chunks, engine_name="tcgen05_x8"))
    # the values the fused run counted are the oracle's
    mx = b[0]
    for ci, c in enumerate(chunks):
        for si, U in enumerate(bases):
            assert abs(mx[ci, si] - orc.mpx_ds_direct(c, U, Nc).max()) < 1e-5


def test_fused_with_gap_nan_chunk_and_core_lags(engine, gap_golden):
    g = gap_golden
    x, Nc = g["gap_chunk"], int(g["gap_Nc"])
    bases = [g["gap_U0"], g["gap_U1"]]
    bad = x.copy()
    bad[5000] = np.nan                                           # a non-finite sample: every row of the chunk is dropped
    chunks = [x, bad, x[: 4000 * Nc].copy()]
    engine.set_bases(81, bases, Nc, thresholds=[0.5, 0.5])
    a = _run(engine, 81, False, chunks)
    b = _run(engine, 81, True, chunks)
    assert (a[1][0] == 2).all() and (a[1][1] & 1).all() and np.isnan(a[0][1]).all()
    _same(a, b)
    T = [len(c) // Nc - 300 + 1 for c in chunks]
    core = ([400, 0, 1000], [T[0] - 100, T[1], 2500])            # halo lags only feed the LTA windows
    _same(_run(engine, 81, False, chunks, core=core), _run(engine, 81, True, chunks, core=core))


def test_fused_accumulates_over_batches(engine):
    Nc, ns, Ls = 3, 200, 7000
    chunks, bases, _ = synth.detection_case(83, 5, Ls, ns, Nc, [2, 4, 6], planted=5)
    engine.set_bases(82, bases, Nc, thresholds=[0.3] * 3)
    ref = _run(engine, 82, False, chunks)
    engine.hist(82, reset=True)
    engine.set_fused(True)
    try:
        engine.accumulate_begin(len(chunks))
        for lo in (0, 2, 4):
            engine.load_chunks(chunks[lo:lo + 2])
            engine.detect_run(82, lta_window=50)
        mx, fl = engine.rowstats()
        cand = np.sort(engine.candidates(), order=["row", "t"])
        engine.accumulate_end()
    finally:
        engine.set_fused(False)
    _same(ref, (mx, fl, cand, engine.hist(82, reset=True), ref[4]))


def test_fused_falls_back_for_rank_above_16(engine):
    chunks, bases, _ = synth.detection_case(84, 1, 5000, 150, 3, [20, 3], planted=1)
    engine.set_bases(83, bases, 3, thresholds=[0.3, 0.3])
    engine.set_fused(True)
    try:
        engine.load_chunks(chunks)
        engine.detect_run(83, lta_window=50)
        assert engine.get_ds(0, 0).shape == (5000 - 150 + 1,)     # pieces must be summed first: dense path
    finally:
        engine.set_fused(False)
