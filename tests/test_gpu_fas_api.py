"""FAS statistics and the reference-shaped Python API on the GPU against the oracle."""
import numpy as np
import pytest

from detex_b200 import detect, fas, synth
from oracle import detex_oracle as orc

pytestmark = pytest.mark.gpu


def test_initFAS_matches_oracle(engine):
    Nc, ns, Ls = 3, 300, 8000
    chunks, bases, _ = synth.detection_case(31, 5, Ls, ns, Nc, [1, 4, 8])
    res = fas.initFAS(bases, chunks, Nc, engine=engine, batch=2)
    for si, U in enumerate(bases):
        ref = orc.fas_stats([orc.mpx_ds_direct(c, U, Nc) for c in chunks])
        assert np.array_equal(res[si]["bins"], ref["bins"])
        assert res[si]["hist"].sum() == ref["hist"].sum()
        # values within the 1e-5 tolerance of a bin edge may sit in the neighbouring bin: 0.8 % of the 38 505
        # values per subspace lie that close to an edge; the float32 statistic (error ~2e-6) moves a handful
        assert np.abs(res[si]["hist"] - ref["hist"]).sum() <= 16
        a, b = res[si]["betadist"][:2]
        ra, rb = ref["betadist"][:2]
        assert abs(a - ra) < 1e-4 * ra and abs(b - rb) < 1e-4 * rb
        assert abs(res[si]["nnlf"] - ref["nnlf"]) < 1e-5 * abs(ref["nnlf"])
        assert abs(orc.threshold_from_beta(a, b) - orc.threshold_from_beta(ra, rb)) < 1e-5


@pytest.mark.parametrize("numBins", [101, 257, 1001])
def test_initFAS_numBins(engine, numBins):
    """fas._initFAS(numBins=...) (fas.py:31, 79): the device histogram on np.linspace(-.01, 1, numBins)
    equals np.histogram of the GPU's own DS values, for any bin count; detection stays at 400 bins."""
    Nc, ns = 3, 150
    rng = np.random.default_rng(71)
    null = [synth.multiplex(synth.bandpassed_noise(rng, 6000 + 11 * i, nchan=Nc)) for i in range(3)]
    bases = [synth.random_basis(rng, ns * Nc, r) for r in (1, 4, 2)]
    res = fas.initFAS(bases, null, Nc, numBins=numBins, engine=engine, set_id=905)
    ref = fas.initFAS(bases, null, Nc, engine=engine, set_id=905)
    engine.set_bases(906, bases, Nc)
    engine.load_chunks(null)
    engine.detect_run(906)
    edges = np.linspace(-.01, 1, numBins)
    for si, r in enumerate(res):
        ds = np.concatenate([engine.get_ds(ci, si).astype(np.float64) for ci in range(len(null))])
        assert r["hist"].shape == (numBins - 1,) and np.array_equal(r["hist"], np.histogram(ds, bins=edges)[0])
        assert np.array_equal(r["bins"], edges)
        # the sums behind the beta fit are grouped differently (the quad fast path follows the bin runs)
        assert np.allclose(r["betadist"][:2], ref[si]["betadist"][:2], rtol=1e-6) and ref[si]["hist"].shape == (400,)
    assert engine.hist(906, reset=True).shape == (3, 400)
    with pytest.raises(ValueError):
        fas.initFAS(bases, null, Nc, numBins=1, engine=engine)


def test_mpxds_dropins(engine):
    chunks, bases, _ = synth.detection_case(32, 1, 5000, 200, 3, [3])
    ref = orc.mpx_ds_fft(chunks[0], bases[0], 3)
    a = detect._MPXDS(chunks[0], 0, bases[0], None, 3, None, engine=engine)
    b = fas._MPXSSCorr(chunks[0], 0, bases[0], None, 3, engine=engine)
    assert a.dtype == np.float64 and a.shape == ref.shape
    assert np.abs(a - ref).max() < 1e-5 and np.abs(b - ref).max() < 1e-5


def test_corDat_triggers_match_reference_loop(engine):
    Nc, ns, Ls, sr = 3, 300, 12000, 100.0
    ranks = [2, 3, 5]
    chunks, bases, truth = synth.detection_case(33, 4, Ls, ns, Nc, ranks, planted=4)
    names = ["SS%d" % i for i in range(len(bases))]
    ssTD = dict(zip(names, bases))
    thr = dict(zip(names, [0.3, 0.3, 0.35]))
    offs = {n: [1.0, 2.0, 4.5] for n in names}
    starts = [1.0e9 + 3600.0 * i for i in range(len(chunks))]
    det = detect.SSDetex(ssTD, thr, offs, Nc, sta="TST", engine=engine, set_id=7)
    df, hist = det.corDat(chunks, sr, starts, batch=3)
    exp = []
    for ci, c in enumerate(chunks):
        for name in names:
            ds = orc.mpx_ds_direct(c, ssTD[name], Nc)
            if not orc.eval_trig_con(ds.max(), thr[name]):
                continue
            sl = orc.sta_lta(ds, 5 * sr, 0)
            for r in orc.greedy_triggers(ds, thr[name], sr, starts[ci], offs[name], stalta=sl):
                exp.append((name, r))
    assert len(df) == len(exp) and len(exp) >= len(truth) // 2
    got = sorted(zip(df.Name, df.STMP, df.DS, df.DS_STALTA, df.MSTAMPmin, df.MSTAMPmax))
    want = sorted((n, r["STMP"], r["DS"], r["DS_STALTA"], r["MSTAMPmin"], r["MSTAMPmax"]) for n, r in exp)
    for g, w in zip(got, want):
        assert g[0] == w[0] and g[1] == w[1]                 # trigger times bit-exact
        assert abs(g[2] - w[2]) < 1e-5                       # DS within tolerance
        assert abs(g[3] - w[3]) < 1e-3 * abs(w[3])
        assert g[4] == w[4] and g[5] == w[5]
    assert list(df.columns) == detect.SAR_COLS
    tot = sum(h.sum() for h in hist.values())
    assert tot == sum((Ls - ns + 1) for _ in chunks) * len(names)
    cor = det.getRA(chunks[0], sr, starts[0])
    assert list(cor.columns) == detect.CORDF_COLS and list(cor.index) == names
    ds0 = orc.mpx_ds_direct(chunks[0], ssTD["SS1"], Nc)
    assert np.abs(cor.SSdetect["SS1"] - ds0).max() < 1e-5
    assert abs(cor.MaxDS["SS1"] - ds0.max()) < 1e-5
    st = orc.sta_lta(ds0, 5 * sr, 0)
    assert np.abs(cor.STALTA["SS1"] - st).max() < 2e-3 * st.max()


def test_corDat_with_trigger_sta_time(engine):
    """SSDetex(triggerSTATime=0.1): DS_STALTA column against the reference loop restated in the oracle."""
    Nc, ns, Ls, sr = 3, 200, 8000, 100.0
    chunks, bases, _ = synth.detection_case(35, 2, Ls, ns, Nc, [2, 3], planted=3)
    names = ["SS0", "SS1"]
    ssTD = dict(zip(names, bases))
    thr = dict(zip(names, [0.3, 0.3]))
    offs = {n: [0.5, 1.0] for n in names}
    starts = [5.0e8, 5.0e8 + 3600.0]
    det = detect.SSDetex(ssTD, thr, offs, Nc, sta="TST", engine=engine, set_id=8, triggerLTATime=2,
                         triggerSTATime=0.1)
    try:
        df, _ = det.corDat(chunks, sr, starts)
    finally:
        engine.set_trigger_sta(0)
    exp = {}
    for ci, c in enumerate(chunks):
        for name in names:
            ds = orc.mpx_ds_direct(c, ssTD[name], Nc)
            if not orc.eval_trig_con(ds.max(), thr[name]):
                continue
            sl = orc.sta_lta(ds, 2 * sr, 0.1 * sr)
            for r in orc.greedy_triggers(ds, thr[name], sr, starts[ci], offs[name], stalta=sl):
                exp[(name, r["STMP"])] = r["DS_STALTA"]
    assert len(df) == len(exp) > 0
    for name, stmp, sl in zip(df.Name, df.STMP, df.DS_STALTA):
        assert abs(sl - exp[(name, stmp)]) < 1e-3 * abs(exp[(name, stmp)])


def test_stalta_screen_matches_oracle(engine):
    """fas._checkSTALTA on the GPU vs the restated ObsPy classic_sta_lta (parity of that
    third-party function itself is unpinned: ObsPy is not installable)."""
    rng = np.random.default_rng(41)
    Nc, Ls, sr = 3, 20000, 100.0
    chunks = []
    for i in range(6):
        ch = synth.bandpassed_noise(rng, Ls - 7 * i, nchan=Nc)
        if i % 2:
            ch[2, 5000 + 100 * i: 5300 + 100 * i] *= 12.0      # a transient on Z
        chunks.append(synth.multiplex(ch))
    engine.load_chunks(chunks)
    mx = engine.sta_lta_max(Nc, 2, int(0.5 * sr), int(5 * sr))
    ref = [orc.classic_sta_lta(c[2::Nc], 0.5 * sr, 5 * sr).max() for c in chunks]
    assert np.allclose(mx, ref, rtol=1e-6)
    passes = fas.screen_chunks(chunks, Nc, sr, engine=engine, batch=4)
    assert passes == [orc.check_stalta(c[2::Nc], sr, 0.5, 5, 8.0) for c in chunks]
    assert passes == [True, False, True, False, True, False]
    assert fas.select_null_chunks(passes, 2) == orc.select_null_chunks(passes, 2) == [0, 2]
    assert fas.select_null_chunks([False] * 7 + [True], 3) == [0, 1, 2]      # <= 25 % pass: screen dropped


def test_est_mags_matches_reference_golden(engine):
    """N1: per-detection ProEnMag / Mag / SNR on the GPU against `_estMag` of the reference."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "mag_golden.npz"))
    x, Nc, U, ewf, mags = g["x"], int(g["Nc"]), g["U"], g["ewf"], g["mags"]
    trigs = g["trigs"].astype(np.int32)
    engine.set_bases(40, [U], Nc, thresholds=[0.5])
    engine.load_chunks([x, x[: len(x) - 30]])
    engine.detect_run(40)
    wfu = (ewf @ U.T) @ U
    engine.set_events(40, 0, ewf, mags, wfu_var=np.var(wfu, axis=1))
    out = engine.est_mags(40, np.zeros(3, np.int32), np.zeros(3, np.int32), trigs)
    assert np.allclose(out, g["sub_out"], rtol=0, atol=1e-9)
    out1 = engine.est_mags(40, np.ones(2, np.int32), np.zeros(2, np.int32), trigs[:2])   # second chunk
    assert np.allclose(out1, g["sub_out"][:2], rtol=0, atol=1e-9)
    engine.set_events(40, 0, ewf, np.full(7, -99.0), wfu_var=np.var(wfu, axis=1))
    o = engine.est_mags(40, [0], [0], [1500])[0]
    assert np.isnan(o[0]) and np.isnan(o[1]) and abs(o[2] - g["nomag_out"][2]) < 1e-9
    # singleton
    single = g["single"]
    us = (single / np.linalg.norm(single))[None, :]
    engine.set_bases(41, [us], Nc, thresholds=[0.5])
    engine.load_chunks([x])
    engine.detect_run(41)
    engine.set_events(41, 0, single[None, :], [1.7], is_single=True)
    outs = engine.est_mags(41, np.zeros(3, np.int32), np.zeros(3, np.int32), trigs)
    assert np.allclose(outs, g["single_out"], rtol=0, atol=1e-9)


def test_corDat_fills_magnitude_columns(engine):
    rng = np.random.default_rng(51)
    Nc, ns, Ls, sr = 3, 200, 9000, 100.0
    n = Nc * ns
    fam = synth.wavelet_basis(rng, ns, Nc, 2)
    ewf = np.array([rng.standard_normal(2) @ fam * 100 + rng.standard_normal(n) for _ in range(5)])
    U = orc.svd_basis(ewf, select_value=0.9)["U"]
    mags = np.array([1.0, 1.5, 2.0, 0.5, 1.2])
    x = synth.multiplex(synth.bandpassed_noise(rng, Ls, nchan=Nc))
    x[4000 * Nc:4000 * Nc + n] += 0.2 * ewf[2]
    det = detect.SSDetex({"SS0": U}, {"SS0": 0.4}, {"SS0": [1.0, 2.0]}, Nc, sta="TST", engine=engine, set_id=9,
                         ewf={"SS0": ewf}, mags={"SS0": mags})
    df, _ = det.corDat([x], sr, [0.0])
    assert len(df) == 1 and abs(df.STMP[0] - 40.0) < 0.05
    ref = orc.est_mag(int(round(df.STMP[0] * sr)), x, Nc, U, ewf, mags, True)
    assert abs(df.ProEnMag[0] - ref[0]) < 1e-9 and abs(df.Mag[0] - ref[1]) < 1e-9 and abs(df.SNR[0] - ref[2]) < 1e-9


def test_preprocess_matches_scipy_restatement(engine):
    """N2: detrend + zero-phase Butterworth band-pass + multiplex on the device against the
    SciPy restatement of ObsPy's bandpass (oracle.apply_filter)."""
    from detex_b200 import preprocess
    rng = np.random.default_rng(61)
    sr = 100.0
    traces = []
    for i in range(3):
        n0 = 30000 - 17 * i
        ch = [np.cumsum(rng.standard_normal(n0 - 5 * c)) * 0.05 + 40.0 * (c + 1) + 3e-3 * np.arange(n0 - 5 * c)
              for c in range(3)]                      # random walk + offset + ramp, ragged channel lengths
        traces.append(ch)
    for filt in ([1, 10, 2, True], [2, 8, 4, False], None):
        got = preprocess.applyFilter_multiplex(traces, sr, filt, engine=engine)
        for g, ch in zip(got, traces):
            ref = orc.apply_filter(ch, sr, filt)
            assert g.shape == ref.shape
            assert np.abs(g - ref).max() < 1e-9 * max(1.0, np.abs(ref).max())
    # st.decimate(factor) first (ObsPy: Chebyshev-II low-pass, then every factor-th sample)
    for factor, filt in ((2, [1, 10, 2, True]), (5, [1, 8, 4, False]), (4, None)):
        got = preprocess.applyFilter_multiplex(traces, sr, filt, decimate=factor, engine=engine)
        for g, ch in zip(got, traces):
            ref = orc.apply_filter(ch, sr, filt, decimate=factor)
            assert g.shape == ref.shape
            assert np.abs(g - ref).max() < 1e-9 * max(1.0, np.abs(ref).max())
    # and straight into the detector: same triggers as filtering on the host first
    chunks, bases, _ = synth.detection_case(62, 2, 12000, 300, 3, [3, 4], planted=2)
    raw = [[c[k::3] + 25.0 + 1e-3 * np.arange(len(c) // 3) for k in range(3)] for c in chunks]
    names = ["SS0", "SS1"]
    det = detect.SSDetex(dict(zip(names, bases)), {n: 0.3 for n in names}, {n: [0.0, 1.0] for n in names}, 3,
                         engine=engine, set_id=11)
    df_raw, _, _ = det.run_raw_chunks(raw, sr, [0.0, 3600.0], filt=[1, 10, 2, True])
    host = [orc.apply_filter(ch, sr, [1, 10, 2, True]) for ch in raw]
    df_host, _, _ = det.run_chunks(host, sr, [0.0, 3600.0])
    assert len(df_raw) == len(df_host)
    assert np.array_equal(df_raw.STMP.values, df_host.STMP.values)
    assert np.abs(df_raw.DS.values - df_host.DS.values).max() < 1e-5
