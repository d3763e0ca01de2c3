"""The opt-in 8-bit cross-term engines ("tcgen05_x8": forced, "tcgen05_auto": per chunk).

hi*hi runs in fp16 and both cross terms in ONE e4m3 x e5m2 MMA, so the error is statistical:
it averages down with the number of taps that carry energy.  These tests pin (i) the forced
engine on diffuse (noise-like) data and templates against the float64 closed form, (ii) that
the adaptive engine sends exactly the chunks its fourth-moment model admits to the 8-bit mode
and the others (spikes, short / concentrated templates) to the fp16 cross terms, where the
result is bit-identical to the default engine, and (iii) the 1e-5 tolerance in every case the
adaptive engine accepts."""
import numpy as np
import pytest

from detex_b200 import synth
from oracle import detex_oracle as orc

pytestmark = pytest.mark.gpu
TOL = 1e-5
NC, NS, LS = 3, 2000, 30000


def _case(seed, nchunks=3, ranks=(3, 5, 1)):
    chunks, bases, _ = synth.detection_case(seed, nchunks, LS, NS, NC, list(ranks), planted=2)
    return chunks, bases


def _max_err(engine, chunks, bases):
    return max(float(np.abs(engine.get_ds(ci, si) - orc.mpx_ds_direct(c, U, NC)).max())
               for ci, c in enumerate(chunks) for si, U in enumerate(bases))


def test_forced_x8_on_diffuse_data(engine):
    chunks, bases = _case(41, nchunks=2)
    engine.set_bases(40, bases, NC)
    engine.load_chunks(chunks)
    engine.detect_run(40, engine="tcgen05_x8")
    assert engine.chunk_modes().tolist() == [1, 1]
    assert _max_err(engine, chunks, bases) < TOL


def test_auto_picks_per_chunk_and_falls_back_bit_exact(engine):
    chunks, bases = _case(42, nchunks=3)
    chunks[1][9000:9030] += 1e5 * np.hanning(30)          # an earthquake-sized transient
    engine.set_bases(41, bases, NC)
    engine.load_chunks(chunks)
    engine.detect_run(41, engine="tcgen05")
    ref1 = [engine.get_ds(1, si).copy() for si in range(len(bases))]
    engine.detect_run(41, engine="tcgen05_auto")
    assert engine.chunk_modes().tolist() == [1, 0, 1]
    for si in range(len(bases)):                           # the rejected chunk ran the default path
        assert np.array_equal(engine.get_ds(1, si), ref1[si], equal_nan=True)
    assert _max_err(engine, chunks, bases) < TOL


def test_auto_rejects_concentrated_templates(engine):
    """Energy in ~0.3 % of the taps: the 8-bit errors do not average down (forced x8 measures
    ~1e-5 here), so the basis' fourth-moment concentration must keep every chunk on fp16."""
    chunks, _ = _case(43, nchunks=2)
    rng = np.random.default_rng(43)
    n = NS * NC
    A = rng.standard_normal((n, 3)) * np.exp(-np.arange(n) / (0.003 * n))[:, None]
    U = np.linalg.qr(A)[0].T.copy()
    chunks[0][3000 * NC:3000 * NC + n] += 8.0 * np.sqrt(n) * U[0]     # a match of that template
    engine.set_bases(42, [U], NC)
    engine.load_chunks(chunks)
    engine.detect_run(42, engine="tcgen05_auto")
    assert engine.chunk_modes().tolist() == [0, 0]
    assert _max_err(engine, chunks, [U]) < TOL


def test_auto_tolerance_zero_is_the_default_engine(engine):
    chunks, bases = _case(44, nchunks=2)
    engine.set_bases(43, bases, NC)
    engine.load_chunks(chunks)
    engine.detect_run(43, engine="tcgen05")
    ref = [[engine.get_ds(ci, si).copy() for si in range(len(bases))] for ci in range(2)]
    engine.set_x8_tolerance(0.0)
    try:
        engine.detect_run(43, engine="tcgen05_auto")
        assert engine.chunk_modes().tolist() == [0, 0]
        for ci in range(2):
            for si in range(len(bases)):
                assert np.array_equal(engine.get_ds(ci, si), ref[ci][si])
    finally:
        engine.set_x8_tolerance(2e-6)


def test_auto_nan_chunk_stays_on_fp16(engine):
    chunks, bases = _case(45, nchunks=2)
    chunks[0][5000] = np.nan
    engine.set_bases(44, bases, NC)
    engine.load_chunks(chunks)
    engine.detect_run(44, engine="tcgen05_auto")
    assert engine.chunk_modes().tolist() == [0, 1]


def test_x8_triggers_match_default_engine(engine):
    """Same trigger list (rows, lags) from both engines when no statistic sits within
    tolerance of the threshold."""
    chunks, bases = _case(46, nchunks=2)
    thr = [0.3] * len(bases)
    engine.set_bases(45, bases, NC, thresholds=thr)
    engine.load_chunks(chunks)
    out = {}
    for e in ("tcgen05", "tcgen05_auto"):
        engine.hist(45, reset=True)
        engine.detect_run(45, engine=e, lta_window=500)
        c = engine.candidates()
        out[e] = c[np.lexsort((c["t"], c["row"]))]
    a, b = out["tcgen05"], out["tcgen05_auto"]
    near = np.abs(a["ds"] - 0.3) < 2 * TOL
    assert not near.any()
    assert np.array_equal(a["row"], b["row"]) and np.array_equal(a["t"], b["t"])
    assert np.abs(a["ds"] - b["ds"]).max() < TOL
