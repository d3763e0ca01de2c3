"""BASELINE configs[0] stand-in: the reference's whole function sequence
createCluster -> updateReqCC -> SVD -> getFAS -> thresholds -> detex on a synthetic,
Case1-SHAPED data set (2 stations x 3 channels x 40 Hz; event families; hour-like chunks with
planted family members), GPU path against the oracle doing the same sequence on the CPU.
The bundled Case1 needs IRIS downloads (tests/test_cases/test_case1.py:40) and ships no
waveforms, so the shapes are kept and the sizes reduced to what the oracle finishes in seconds.
"""
import numpy as np
import pytest
from scipy.cluster.hierarchy import fcluster

from detex_b200 import construct, detect, fas, subspace, synth
from oracle import detex_oracle as orc

pytestmark = pytest.mark.gpu

SR, NC = 40.0, 3
NS_EVENT = 800        # 20 s event windows
NS_TRIM = 400         # 10 s templates after trimming
LS = 24000            # 600 s chunks
CCREQ = 0.5


def _align_and_trim(X, lag_to_first):
    """Test-side stand-in for createSubSpace's alignment (out of scope): shift every member
    by its CCX lag against the first member, then keep NS_TRIM samples per channel."""
    out = []
    for x, lag in zip(X, lag_to_first):
        s = int(lag) // NC
        ch = x.reshape(-1, NC).T
        start = 200 + s
        out.append(synth.multiplex(ch[:, start:start + NS_TRIM]))
    return np.array(out)


def _station(seed):
    rng = np.random.default_rng(seed)
    events = synth.event_families(seed, 3, 6, NS_EVENT, NC, sr=SR, max_shift=40, noise=0.35)
    chunks = [synth.multiplex(synth.bandpassed_noise(rng, LS, sr=SR, nchan=NC)) for _ in range(5)]
    null = [synth.multiplex(synth.bandpassed_noise(rng, LS, sr=SR, nchan=NC)) for _ in range(4)]
    planted = []
    for ci, (ev, t) in enumerate([(1, 5000), (8, 11000), (14, 3000), (3, 17000)]):
        x = events[ev] * (6.0 / events[ev].std())
        chunks[ci][t * NC:t * NC + len(x)] += x
        planted.append((ci, ev // 6, t))
    return events, chunks, null, planted


@pytest.mark.parametrize("seed", [1001, 1002])
def test_case1_shaped_sequence(engine, seed):
    events, chunks, null, planted = _station(seed)
    N = len(events)
    # ---- createCluster: CCX matrix -> linkage (construct.py:139-157)
    import pandas as pd
    evs = ["ev%02d" % i for i in range(N)]
    row = pd.Series({"MPtd": dict(zip(evs, events)), "MPfd": {e: None for e in evs},
                     "Channels": {e: ["BHE", "BHN", "BHZ"] for e in evs}})
    DFcc, DFlag, DFsub = construct._makeDFcclags(evs, row, engine=engine)
    rcc, rlag, rsub = orc.make_cclags(events, NC)
    m = ~np.isnan(rcc)
    assert np.abs(DFcc.values.astype(float)[m] - rcc[m]).max() < 1e-10
    assert np.array_equal(DFlag.values.astype(float)[m], rlag[m])
    link = construct.cluster_link(DFcc)
    assert np.allclose(link, orc.cluster_link(rcc), atol=1e-9)
    # ---- updateReqCC: cut the dendrogram (subspace.py:385-395)
    T = fcluster(link, 1 - CCREQ, criterion="distance")
    fams = {}
    for i, c in enumerate(T):
        fams.setdefault(c, []).append(i)
    clusters = sorted([v for v in fams.values() if len(v) >= 2], key=lambda v: v[0])
    assert len(clusters) == 3 and sorted(len(c) for c in clusters) == [6, 6, 6]
    # ---- SVD (subspace.py:875-905) on aligned + trimmed members
    lags_full = np.zeros((N, N))
    iu = np.triu_indices(N, 1)
    lags_full[iu] = rlag[iu[0], iu[1] - 1]
    ssTD, thr, offs, names = {}, {}, {}, []
    for k, members in enumerate(clusters):
        W = _align_and_trim(events[members], [0] + [lags_full[members[0], j] for j in members[1:]])
        a = subspace.svd_basis(W, selectCriteria=2, selectValue=0.9)
        b = orc.svd_basis(W, select_criteria=2, select_value=0.9)
        assert a["NumBasis"] == b["ndim"] and np.allclose(a["U"], b["U"])
        name = "SS%d" % k
        names.append(name)
        ssTD[name] = a["U"]
        offs[name] = [1.0, 2.0, 3.0]
    # ---- getFAS + thresholds (fas.py:23-86; subspace.py:1027-1047)
    res = fas.initFAS([ssTD[nm] for nm in names], null, NC, engine=engine)
    for nm, r in zip(names, res):
        ref = orc.fas_stats([orc.mpx_ds_direct(c, ssTD[nm], NC) for c in null])
        a, b = r["betadist"][:2]
        assert abs(a - ref["betadist"][0]) < 1e-4 * a and abs(b - ref["betadist"][1]) < 1e-4 * b
        thr[nm] = subspace.threshold_from_fas(r, Pf=1e-12)
        assert abs(thr[nm] - orc.threshold_from_beta(*ref["betadist"][:2], Pf=1e-12)) < 1e-5
        assert 0.02 < thr[nm] < 0.6
    # ---- detex (detect.py:137-218)
    starts = [3600.0 * i for i in range(len(chunks))]
    det = detect.SSDetex(ssTD, thr, offs, NC, sta="M17A", engine=engine, set_id=20 + seed % 7)
    df, hist = det.corDat(chunks, SR, starts)
    exp = []
    for ci, c in enumerate(chunks):
        for nm in names:
            ds = orc.mpx_ds_direct(c, ssTD[nm], NC)
            if orc.eval_trig_con(ds.max(), thr[nm]):
                for r in orc.greedy_triggers(ds, thr[nm], SR, starts[ci], offs[nm], stalta=orc.sta_lta(ds, 5 * SR, 0)):
                    exp.append((nm, r["STMP"], r["DS"]))
    got = sorted(zip(df.Name, df.STMP, df.DS))
    exp = sorted(exp)
    assert len(got) == len(exp) >= len(planted)
    for g, w in zip(got, exp):
        assert g[0] == w[0] and g[1] == w[1] and abs(g[2] - w[2]) < 1e-5
    # every planted family member is detected by its own family's subspace, near where it was put
    for ci, fam, t in planted:
        hit = df[(df.STMP > starts[ci] + t / SR - 10) & (df.STMP < starts[ci] + t / SR + 15)]
        assert len(hit) >= 1 and hit.DS.max() > 0.3
