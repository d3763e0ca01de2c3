"""Host-side mirrors (`detect.SSDetex`, `fas.initFAS`, `construct._makeDFcclags`, `preprocess`) driven
by the oracle-backed engine stand-in: batching, chunk skipping, grouping by basis length, greedy
picks, table columns and error behaviour, on a box without a GPU.  The GPU suite runs the same
mirrors on the CUDA engine."""
import numpy as np
import pandas as pd
import pytest

from detex_b200 import construct, detect, fas, preprocess, synth
from oracle import detex_oracle as orc
from oracle_engine import OracleEngine


def test_corDat_batches_skips_short_chunks_and_groups_by_length():
    Nc, sr = 3, 50.0
    rng = np.random.default_rng(5)
    chunks, bases, _ = synth.detection_case(41, 5, 4000, 150, Nc, [2, 3], planted=3, sr=sr)
    other = synth.random_basis(rng, 100 * Nc, 1)               # a second basis length -> second basis set
    names = ["SS0", "SS1", "SS2"]
    ssTD = dict(zip(names, bases + [other]))
    thr = dict(zip(names, [0.3, 0.3, 0.2]))
    offs = {n: [0.5, 1.0, 2.0] for n in names}
    chunks.insert(2, chunks[0][:300])                          # shorter than the template: skipped with a warning
    starts = [1.0e9 + 100.0 * i for i in range(len(chunks))]
    det = detect.SSDetex(ssTD, thr, offs, Nc, sta="TST", engine=OracleEngine(), set_id=3, triggerLTATime=1)
    df, hist = det.corDat(chunks, sr, starts, batch=4)
    exp = []
    for ci, c in enumerate(chunks):
        if ci == 2:
            continue
        for name in names:
            ds = orc.mpx_ds_direct(c, ssTD[name], Nc).astype(np.float32).astype(np.float64)
            if ds.max() > np.float32(thr[name]):
                sl = orc.sta_lta(ds, 1 * sr, 0)
                for r in orc.greedy_triggers(ds, float(np.float32(thr[name])), sr, starts[ci], offs[name], stalta=sl):
                    exp.append((name, r["STMP"], r["MSTAMPmin"], r["MSTAMPmax"]))
    got = sorted(zip(df.Name, df.STMP, df.MSTAMPmin, df.MSTAMPmax))
    assert got == sorted(exp) and len(got) > 0
    assert list(df.columns) == detect.SAR_COLS and set(df.Sta) == {"TST"}
    for name in names:
        n = ssTD[name].shape[1] // Nc
        assert hist[name].sum() == sum(len(c) // Nc - n + 1 for i, c in enumerate(chunks) if i != 2)
    cor = det.getRA(chunks[0], sr, starts[0], File="f0")
    assert list(cor.columns) == detect.CORDF_COLS and list(cor.index) == names
    assert cor.at["SS2", "File"] == "f0" and cor.at["SS2", "Nc"] == Nc
    assert det.getRA(chunks[2], sr, starts[2]) is None


def test_kill_switch_and_ds_above_one_filter():
    Nc, sr = 1, 10.0
    eng = OracleEngine()
    U = synth.random_basis(np.random.default_rng(1), 20, 1)
    det = detect.SSDetex({"SS0": U}, {"SS0": 1e-9}, {"SS0": [0.0]}, Nc, engine=eng, triggerLTATime=1)
    x = np.random.default_rng(2).standard_normal(3000)
    df, _, _ = det.run_chunks([x], sr, [0.0])
    # every pick zeroes +-20 s = +-200 samples: ~3000 / 200 picks, far below the 4000 kill switch
    assert 5 < len(df) < 30 and (np.diff(np.sort(df.STMP.values)) >= 20.0 - 1e-9).all()
    with pytest.raises(ValueError):
        detect.SSDetex({}, {}, {}, 1, engine=eng)
    with pytest.raises(ValueError):
        detect.SSDetex({"SS0": U}, {"SS0": .5}, {"SS0": [0.0]}, 1, engine=eng, triggerSTATime=-1)


def test_initFAS_and_screen_with_stand_in():
    Nc, sr = 3, 40.0
    rng = np.random.default_rng(9)
    null = [synth.multiplex(synth.bandpassed_noise(rng, 3000, sr=sr, nchan=Nc)) for _ in range(5)]
    null[1][300 * Nc:305 * Nc] += 80.0                          # a transient: fails the STA/LTA screen
    bases = [synth.random_basis(rng, 120 * Nc, r) for r in (1, 3)]
    eng = OracleEngine()
    passes = fas.screen_chunks(null, Nc, sr, STATime=0.5, LTATime=5, staltalimit=8.0, engine=eng, batch=2)
    assert passes == [True, False, True, True, True]
    kept = fas.select_null_chunks(passes, 3)
    assert kept == [0, 2, 3]
    res = fas.initFAS(bases, [null[i] for i in kept], Nc, engine=eng, batch=2)
    for U, r in zip(bases, res):
        ref = orc.fas_stats([orc.mpx_ds_direct(null[i], U, Nc).astype(np.float32).astype(np.float64) for i in kept])
        assert np.array_equal(r["hist"], ref["hist"])
        assert np.allclose(r["betadist"][:2], ref["betadist"][:2], rtol=1e-6)
        assert abs(r["nnlf"] - ref["nnlf"]) < 1e-6 * abs(ref["nnlf"])
    # numBins (fas.py:31): any bin count up to the device limit; the beta fit does not depend on it
    r101 = fas.initFAS(bases, [null[i] for i in kept], Nc, numBins=101, engine=eng, batch=2)
    for U, r, r0 in zip(bases, r101, res):
        ds = np.concatenate([orc.mpx_ds_direct(null[i], U, Nc).astype(np.float32).astype(np.float64) for i in kept])
        assert len(r["bins"]) == 101 and np.array_equal(r["hist"], np.histogram(ds, bins=np.linspace(-.01, 1, 101))[0])
        assert r["betadist"] == r0["betadist"]
    assert eng.hist_bins == 400                                     # restored for detection (detect.py:80)
    with pytest.raises(ValueError):
        fas.initFAS(bases, null, Nc, numBins=5000, engine=eng)


def test_makeDFcclags_frames_and_errors():
    X = synth.event_families(4, 2, 3, 120, 2, sr=50.0, max_shift=8)
    evs = ["e%d" % i for i in range(len(X))]
    row = pd.Series({"MPtd": dict(zip(evs, X)), "MPfd": {e: None for e in evs}, "Channels": {e: ["Z", "N"] for e in evs}})
    cc, lag, sub = construct._makeDFcclags(evs, row, engine=OracleEngine())
    assert list(cc.index) == list(range(5)) and list(cc.columns) == list(range(1, 6))
    assert np.isnan(cc.values.astype(float)[np.tril_indices(5, -1)]).all()
    rcc, rlag, _ = orc.make_cclags(X, 2)
    m = ~np.isnan(rcc)
    assert np.array_equal(cc.values.astype(float)[m], rcc[m]) and np.array_equal(lag.values.astype(float)[m], rlag[m])
    row.Channels["e1"] = ["Z"]
    with pytest.raises(Exception, match="Channels"):
        construct._makeDFcclags(evs, row, engine=OracleEngine())
    row.Channels["e1"] = ["Z", "N"]
    row.MPtd["e2"] = X[2][:-2]
    with pytest.raises(Exception, match="Lengths"):
        construct._makeDFcclags(evs, row, engine=OracleEngine())


def test_bandpass_design_matches_oracle_filter():
    rng = np.random.default_rng(11)
    tr = [[rng.standard_normal(900) + 3.0 for _ in range(3)]]
    got = preprocess.applyFilter_multiplex(tr, 40.0, (1, 10, 2, True), engine=OracleEngine())[0]
    assert np.abs(got - orc.apply_filter(tr[0], 40.0, (1, 10, 2, True))).max() < 1e-12
    got = preprocess.applyFilter_multiplex(tr, 40.0, (1, 4, 2, False), decimate=4, engine=OracleEngine())[0]
    ref = orc.apply_filter(tr[0], 40.0, (1, 4, 2, False), decimate=4)
    assert got.shape == ref.shape == (225 * 3,) and np.abs(got - ref).max() < 1e-12
    with pytest.raises(ArithmeticError):                           # ObsPy refuses automatic design above 16
        preprocess.applyFilter(tr, 40.0, None, decimate=17, engine=OracleEngine())
    with pytest.warns(UserWarning):                                # ObsPy: high corner at Nyquist -> high-pass
        preprocess.bandpass_sos(1.0, 20.0, 40.0, 2)
    with pytest.raises(ValueError):
        preprocess.bandpass_sos(30.0, 35.0, 40.0, 2)


def test_long_array_segments_equal_single_chunk():
    """Time-segment sharding with a halo (`SSDetex.run_long_array`): cutting a long array into
    overlapping segments whose core lags partition the array gives the triggers, MaxDS and histograms
    of the uncut array processed as one chunk (host logic on the oracle-backed engine)."""
    Nc, ns, Ls, sr = 3, 100, 9000, 100.0
    chunks, bases, _ = synth.detection_case(93, 1, Ls, ns, Nc, [2, 3], planted=4)
    x = chunks[0]
    names = ["SS0", "SS1"]
    mk = lambda: detect.SSDetex(dict(zip(names, bases)), {n: 0.3 for n in names}, {n: [0.0, 1.5] for n in names}, Nc,
                                engine=OracleEngine(), triggerLTATime=0.5)
    one = mk()
    ref, mref, _ = one.run_chunks([x], sr, [1000.0])
    for seg in (2000, 4096, 1 << 18):
        det = mk()
        got, mx = det.run_long_array(x, sr, 1000.0, seg_lags=seg, batch=2)
        segs = det.long_array_segments(Ls, ns, seg, 52)
        assert sum(hi - lo for _, _, lo, hi in segs) == Ls - ns + 1
        assert len(got) == len(ref) > 0
        assert np.array_equal(got.STMP.values, ref.STMP.values) and list(got.Name) == list(ref.Name)
        assert np.abs(got.DS.values - ref.DS.values).max() < 1e-6
        assert np.abs(got.DS_STALTA.values - ref.DS_STALTA.values).max() < 1e-4 * ref.DS_STALTA.values.max()
        for name in names:
            assert abs(mx[name] - mref[0][name]) < 1e-6
            assert np.abs(det.histdic[name] - one.histdic[name]).sum() <= 4
    with pytest.raises(ValueError):
        mk().long_array_segments(Ls, ns, 1001, 52)
