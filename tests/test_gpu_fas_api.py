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
        assert np.abs(res[si]["hist"] - ref["hist"]).sum() <= 6     # values within 1e-5 of a bin edge
        a, b = res[si]["betadist"][:2]
        ra, rb = ref["betadist"][:2]
        assert abs(a - ra) < 1e-4 * ra and abs(b - rb) < 1e-4 * rb
        assert abs(res[si]["nnlf"] - ref["nnlf"]) < 1e-5 * abs(ref["nnlf"])
        assert abs(orc.threshold_from_beta(a, b) - orc.threshold_from_beta(ra, rb)) < 1e-5


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
