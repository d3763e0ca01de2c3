"""The reference's user-facing sequence (createCluster -> updateReqCC -> createSubSpace ->
attachPickTimes -> SVD -> detex -> SQLite tables) through `detex_b200.workflow`.

CPU (`-m "not gpu"`): the host logic driven by the oracle-backed engine stand-in
(tests/oracle_engine.py) -- containers, dendrogram cut, alignment, trims, FAS bookkeeping, tables.
GPU: the same sequence on the CUDA engine, compared with the stand-in's output row by row.
"""
import numpy as np
import pandas as pd
import pytest

from detex_b200 import results, synth, workflow
from oracle_engine import OracleEngine

CCREQ = 0.55


def _run(eng, tmp_path, seed=77, decimate=None, filt=(1, 10, 2, True)):
    case = synth.workflow_case(seed)
    fetcher = workflow.ArrayFetcher(case["events"], case["continuous"], sr=case["sr"], conDatDuration=280,
                                    conBuff=20, seed=5)
    cl = workflow.createCluster(CCreq=CCREQ, fetch_arg=fetcher, filt=list(filt), stationKey=case["stakey"],
                                templateKey=case["temkey"], trim=[2, 18], saveclust=True, decimate=decimate,
                                fileName=str(tmp_path / "clust.pkl"), engine=eng)
    ss = workflow.createSubSpace(Pf=1e-10, clust=cl, engine=eng)
    ss.attachPickTimes(case["picks"], defaultDuration=8)
    ss.SVD(selectCriteria=2, selectValue=0.9, conDatNum=3)
    db = str(tmp_path / "SubSpace.db")
    found = ss.detex(subspaceDB=db, useSingles=True, estimateMags=True)
    return case, cl, ss, db, found


def _check_structure(case, cl, ss, db):
    assert len(cl) == 2 and repr(cl).startswith("SSClusterStream with 2 stations")
    for sta in ("TA.M17A", "TA.M18A"):
        c = cl[sta]
        assert c is cl[sta.split(".")[1]]
        assert sorted(len(x) for x in c.clusts) == [5, 5, 5] and len(c.singles) == 2
        names = list(case["temkey"].NAME)
        for k in c.clusts:                                   # families are consecutive blocks of 5 events
            idx = sorted(names.index(e) for e in k)
            assert idx[-1] - idx[0] == 4 and idx[0] % 5 == 0
        sp = ss.subspaces[sta]
        assert list(sp.Name) == ["SS0", "SS1", "SS2"]
        for _, row in sp.iterrows():
            assert row.SVDdefined and 1 <= row.NumBasis <= 5 and len(row.UsedSVDKeys) == row.NumBasis
            assert row.SampleTrims["Starttime"] % 3 == 0 and row.SampleTrims["Endtime"] % 3 == 0
            n = row.SampleTrims["Endtime"] - row.SampleTrims["Starttime"]
            assert all(len(v) == n for v in row.SVD.values())
            assert 0.05 < row.Threshold < 0.9 and len(row.Offsets) == 3
            assert set(row.FAS.keys()) == {"bins", "hist", "betadist", "nnlf"}
        sg = ss.singles[sta]
        assert list(sg.Name) == ["SG0", "SG1"] and all(0.05 < t < 0.95 for t in sg.Threshold)
    # the pickled ClusterStream loads and can be cut again
    import pickle
    cl2 = pickle.load(open(cl.filename, "rb"))
    cl2.updateReqCC(0.999)
    assert all(len(c.clusts) == 0 and len(c.singles) == 17 for c in cl2.clusters)
    # tables (subspace.py:1883-1902)
    ssdf = results.loadSQLite(db, "ss_df")
    info = results.loadSQLite(db, "ss_info")
    hist = results.loadSQLite(db, "ss_hist")
    filt = results.loadSQLite(db, "filt_params")
    assert list(ssdf.columns) == ['DS', 'DS_STALTA', 'STMP', 'Name', 'Sta', 'MSTAMPmin', 'MSTAMPmax', 'Mag', 'SNR',
                                  'ProEnMag']
    assert len(info) == 6 and set(info.Sta) == {"TA.M17A", "TA.M18A"}
    assert len(hist) == 7 and list(filt.iloc[0]) == [1, 10, 2, 1]
    assert results.loadSQLite(db, "sg_info") is not None
    # every planted family member is found by a subspace on its station, near where it was put
    t0 = float(case["stakey"].STARTTIME.iloc[0])
    for sta, c, fam, tsec in case["planted"]:
        hit = ssdf[(ssdf.Sta == sta) & (np.abs(ssdf.STMP - (t0 + c * case["chunk_seconds"] + tsec)) < 12)]
        assert len(hit) >= 1 and hit.DS.max() > 0.3, (sta, c, fam)
    assert np.isfinite(ssdf.Mag).all() and np.isfinite(ssdf.SNR).all()
    return ssdf


def test_workflow_host_logic_with_oracle_engine(tmp_path):
    case, cl, ss, db, found = _run(OracleEngine(), tmp_path)
    ssdf = _check_structure(case, cl, ss, db)
    assert sum(v for (sta, issub), v in found.items() if issub) == len(ssdf)
    # ---- detResults (results.py:22-173): both stations must see an event; planted members become Dets
    t0 = float(case["stakey"].STARTTIME.iloc[0])
    veri = pd.DataFrame([{"TIME": t0 + c * case["chunk_seconds"] + tsec - 3.0, "LAT": 40.0, "LON": -110.0, "MAG": 1.0,
                          "DEPTH": 5.0, "NAME": "known%d" % c}
                         for sta, c, fam, tsec in case["planted"] if sta == "TA.M17A"][:2])
    res = results.detResults(ssDB=db, templateKey=case["temkey"], stationKey=case["stakey"], requiredNumStations=2,
                             ss_associateBuffer=8, sg_associateBuffer=8, veriFile=veri, veriBuffer=30)
    assert "new detections" in repr(res) and len(res.Autos) == 0          # training events are a month earlier
    assert len(res.Dets) >= 4 and (res.Dets.NumStations == 2).all()
    for c in range(4):                                                      # chunks 0..3 hold a planted event per station
        m = (res.Dets.MSTAMPmin < t0 + (c + 1) * case["chunk_seconds"]) & (res.Dets.MSTAMPmax > t0 + c * case["chunk_seconds"])
        assert m.any()
    assert res.NumVerified == 2 and set(res.Vers.VerName) == {"known0", "known1"}
    assert res.Dets.Verified.sum() == 2
    one = results.detResults(ssDB=db, templateKey=case["temkey"], stationKey=case["stakey"], requiredNumStations=3)
    assert len(one.Dets) == 0 and one.NumVerified == 'N/A'
    pf = results.detResults(ssDB=db, templateKey=case["temkey"], stationKey=case["stakey"], requiredNumStations=1,
                            Pf=1e-4)
    lo = results.detResults(ssDB=db, templateKey=case["temkey"], stationKey=case["stakey"], requiredNumStations=1)
    assert 0 < len(pf.Dets) <= len(lo.Dets)
    with pytest.raises(Exception):
        results.detResults(ssDB=db, templateKey=case["temkey"], stationKey=case["stakey"], associateReq=1)


def test_subspace_options_without_fas(tmp_path):
    """SVD with a fixed threshold / selectCriteria 3 and 4 (no FAS run), validateClusters, updateReqCC by
    station, createSubSpace from the pickled ClusterStream (subspace.py:738-773, 1015-1054; construct.py:236-243)."""
    case = synth.workflow_case(79, nchunks=2)
    eng = OracleEngine()
    f = workflow.ArrayFetcher(case["events"], case["continuous"], sr=case["sr"], seed=1)
    path = str(tmp_path / "clust.pkl")
    cl = workflow.createCluster(CCreq=CCREQ, fetch_arg=f, stationKey=case["stakey"], templateKey=case["temkey"],
                                trim=[2, 18], fileName=path, engine=eng)
    cl.updateReqCC({"M17A": 0.6, "TA.M18A": 0.5})
    assert cl["M17A"].ccReq == 0.6 and cl[1].ccReq == 0.5
    ss = workflow.createSubSpace(clust=path, engine=eng)             # un-pickles the ClusterStream
    assert sorted(ss.subspaces) == ["TA.M17A", "TA.M18A"]
    ss.attachPickTimes(case["picks"], defaultDuration=8, function="mean")
    before = {sta: [list(r.Events) for _, r in ss.subspaces[sta].iterrows()] for sta in ss.subspaces}
    ss.validateClusters()                                            # well-aligned families: nothing is dropped
    assert before == {sta: [list(r.Events) for _, r in ss.subspaces[sta].iterrows()] for sta in ss.subspaces}
    # a member replaced by noise fails the check against every later member and is removed
    row = ss.subspaces["TA.M17A"].iloc[0]
    ev0 = row.Events[0]
    row.AlignedTD[ev0] = np.random.default_rng(0).standard_normal(len(row.AlignedTD[ev0]))
    ss.validateClusters()
    assert ev0 not in ss.subspaces["TA.M17A"].iloc[0].Events and len(ss.subspaces["TA.M17A"].iloc[0].Events) == 4
    ss.SVD(selectCriteria=4, selectValue=1, threshold=0.33, useSingles=False)
    for sta in ss.subspaces:
        assert (ss.subspaces[sta].NumBasis == 2).all() and (ss.subspaces[sta].Threshold == 0.33).all()
    ss.SVD(selectCriteria=3, selectValue=0.8, useSingles=False)
    for sta in ss.subspaces:
        for _, r in ss.subspaces[sta].iterrows():
            assert abs(r.Threshold - 0.8 * r.FracEnergy["Minimum"][r.NumBasis]) < 1e-12
            assert r.FracEnergy["Average"][r.NumBasis] >= 0.8
    ss.SVD(selectCriteria=4, selectValue=0, threshold=0.4)           # singles get the manual threshold too
    assert all((ss.singles[sta].Threshold == 0.4).all() for sta in ss.singles)
    found = ss.detex(subspaceDB=str(tmp_path / "a.db"), useSingles=True, estimateMags=False, fillZeros=True)
    df = results.loadSQLite(str(tmp_path / "a.db"), "ss_df")
    assert len(df) == sum(v for (sta, sub), v in found.items() if sub)
    assert (df.DS_STALTA == 0).all() and df.Mag.isna().all()          # fillZeros / estimateMags=False (detect.py:414-432)


def test_duplicate_events_do_not_break_the_alignment(tmp_path):
    """Two catalogue entries with identical waveforms give equal correlation coefficients; the reference
    perturbs them (`_ensureUnique`, construct.py:814-835) instead of failing in the dendrogram walk."""
    case = synth.workflow_case(80, stations=("TA.M17A",), nfam=1, per_fam=4, nsingles=0, nchunks=1)
    ev = case["events"]["TA.M17A"]
    names = list(case["temkey"].NAME)
    ev[names[1]] = ([t.copy() for t in ev[names[0]][0]], ev[names[1]][1])       # event 1 := event 0
    eng = OracleEngine()
    f = workflow.ArrayFetcher(case["events"], case["continuous"], sr=case["sr"])
    cl = workflow.createCluster(CCreq=0.4, fetch_arg=f, stationKey=case["stakey"], templateKey=case["temkey"],
                                trim=[2, 18], saveclust=False, engine=eng)
    row = cl.trdf.iloc[0]
    cc = row.CCs.values.astype(float)
    # cc(0,2) == cc(1,2): duplicate coefficients.  (cc(0,1) itself is round-off dependent in the reference:
    # a lag-0 value of 1 + 2e-16 trips its |res| > 1 rule, construct.py:455-459, and the next-best lag wins.)
    assert cc[0, 1] == cc[1, 1] and cc[0, 2] == cc[1, 2]
    ss = workflow.createSubSpace(clust=cl, engine=eng)
    r = ss.subspaces["TA.M17A"].iloc[0]
    assert r.Events == names and len(set(len(v) for v in r.AlignedTD.values())) == 1
    assert np.array_equal(r.AlignedTD[names[0]], r.AlignedTD[names[1]])


def test_workflow_with_decimation(tmp_path):
    """createCluster(decimate=2): 40 Hz traces are low-passed and decimated to 20 Hz before detrend /
    band-pass (construct.py:1014-1015); lags, trims and trigger times then live on the 20 Hz grid."""
    case, cl, ss, db, found = _run(OracleEngine(), tmp_path, decimate=2, filt=(1, 8, 2, True))
    for sta in ("TA.M17A", "TA.M18A"):
        assert sorted(len(x) for x in cl[sta].clusts) == [5, 5, 5]
        row = cl.trdf[cl.trdf.Station == sta].iloc[0]
        assert all(st["sampling_rate"] == 20.0 for st in row.Stats.values())
        for _, r in ss.subspaces[sta].iterrows():
            assert r.SampleTrims["Endtime"] - r.SampleTrims["Starttime"] == 8 * 20 * 3   # 8 s x 20 Hz x 3 ch
    ssdf = results.loadSQLite(db, "ss_df")
    t0 = float(case["stakey"].STARTTIME.iloc[0])
    for sta, c, fam, tsec in case["planted"]:
        hit = ssdf[(ssdf.Sta == sta) & (np.abs(ssdf.STMP - (t0 + c * case["chunk_seconds"] + tsec)) < 12)]
        assert len(hit) >= 1 and hit.DS.max() > 0.3
    assert np.allclose((ssdf.STMP.values - t0) * 20.0, np.round((ssdf.STMP.values - t0) * 20.0), atol=1e-5)


def test_cluster_cut_matches_fcluster():
    """Cluster.updateReqCC (subspace.py:305-346) against scipy's flat clusters on random linkages."""
    from scipy.cluster.hierarchy import fcluster, linkage
    rng = np.random.default_rng(3)
    for N in (2, 5, 12, 40):
        cx = rng.uniform(0.05, 0.95, N * (N - 1) // 2)
        link = linkage(cx)
        key = ["e%02d" % i for i in range(N)]
        for ccreq in (0.2, 0.5, 0.8):
            c = workflow.Cluster(None, "TA.X", None, key, link, ccreq, None, None, None, None)
            T = fcluster(link, 1 - ccreq, criterion="distance")
            want = {}
            for e, t in zip(key, T):
                want.setdefault(t, []).append(e)
            want = sorted(v for v in want.values() if len(v) > 1)
            assert sorted(c.clusts) == want
            assert sorted(c.singles) == sorted(e for e in key if not any(e in v for v in want))


def test_workflow_input_errors(tmp_path):
    case = synth.workflow_case(78, nfam=1, per_fam=2, nsingles=0, nchunks=1)
    f = workflow.ArrayFetcher(case["events"], case["continuous"], sr=case["sr"])
    with pytest.raises(TypeError):
        workflow.createCluster(fetch_arg="EventWaveForms", stationKey=case["stakey"], templateKey=case["temkey"])
    with pytest.raises(Exception):
        workflow.createCluster(fetch_arg=f, filt=[1, 10], stationKey=case["stakey"], templateKey=case["temkey"],
                               engine=OracleEngine())
    cl = workflow.createCluster(CCreq=0.3, fetch_arg=f, stationKey=case["stakey"], templateKey=case["temkey"],
                                trim=[2, 18], saveclust=False, engine=OracleEngine())
    with pytest.raises(Exception):
        cl.updateReqCC(1.5)
    ss = workflow.createSubSpace(clust=cl, engine=OracleEngine())
    with pytest.raises(Exception):                      # SVD not called yet (subspace.py:1858-1861)
        ss.detex(subspaceDB=str(tmp_path / "x.db"))
    with pytest.raises(Exception):                      # trigCon != 0 is rejected by the reference too
        ss.detex(subspaceDB=str(tmp_path / "x.db"), trigCon=1)
    with pytest.raises(ValueError):
        ss.SVD(selectCriteria=2, selectValue=1.5)


@pytest.mark.gpu
def test_workflow_on_gpu_matches_oracle_engine(engine, tmp_path):
    (tmp_path / "gpu").mkdir()
    (tmp_path / "cpu").mkdir()
    case, cl, ss, db, _ = _run(engine, tmp_path / "gpu")
    g = _check_structure(case, cl, ss, db)
    _, cl0, ss0, db0, _ = _run(OracleEngine(), tmp_path / "cpu")
    for sta in ("TA.M17A", "TA.M18A"):
        assert cl[sta].clusts == cl0[sta].clusts and cl[sta].singles == cl0[sta].singles
        a = cl.trdf[cl.trdf.Station == sta].iloc[0]
        b = cl0.trdf[cl0.trdf.Station == sta].iloc[0]
        m = ~np.isnan(b.CCs.values.astype(float))
        assert np.abs(a.CCs.values.astype(float)[m] - b.CCs.values.astype(float)[m]).max() < 1e-9
        assert np.array_equal(a.Lags.values.astype(float)[m], b.Lags.values.astype(float)[m])
        for (_, r), (_, r0) in zip(ss.subspaces[sta].iterrows(), ss0.subspaces[sta].iterrows()):
            assert r.Events == r0.Events and r.SampleTrims == r0.SampleTrims and r.NumBasis == r0.NumBasis
            assert abs(r.Threshold - r0.Threshold) < 2e-4
        for (_, r), (_, r0) in zip(ss.singles[sta].iterrows(), ss0.singles[sta].iterrows()):
            assert abs(r.Threshold - r0.Threshold) < 2e-4
    w = results.loadSQLite(db0, "ss_df")
    # thresholds differ in the 5th digit (float32 DS sums), so compare the detections well above them
    gs = g[g.DS > 0.3].sort_values(["Sta", "Name", "STMP"]).reset_index(drop=True)
    ws = w[w.DS > 0.3].sort_values(["Sta", "Name", "STMP"]).reset_index(drop=True)
    assert len(gs) == len(ws) > 0
    assert (gs.Name == ws.Name).all() and (gs.STMP == ws.STMP).all()
    assert np.abs(gs.DS - ws.DS).max() < 1e-5
    assert np.abs(gs.Mag - ws.Mag).max() < 1e-3 and np.abs(gs.SNR - ws.SNR).max() < 1e-3 * ws.SNR.abs().max()


def test_array_fetcher_generators():
    """getTemData / getConData stand-ins (getdata.py:351, 455): key order, time window, seeded sampling."""
    case = synth.workflow_case(81, nfam=1, per_fam=3, nsingles=1, nchunks=5)
    f = workflow.ArrayFetcher(case["events"], case["continuous"], sr=case["sr"], seed=3)
    got = [name for _, _, name in f.getTemData(case["temkey"], case["stakey"].iloc[:1])]
    assert got == list(case["temkey"].NAME)
    sub = case["temkey"].iloc[[2, 0]]
    assert [n for _, _, n in f.getTemData(sub, case["stakey"].iloc[:1])] == list(sub.NAME)
    t0 = float(case["stakey"].STARTTIME.iloc[0])
    sk = case["stakey"].iloc[1:2]
    starts = [s for _, s in f.getConData(sk)]
    assert starts == [t0 + 300.0 * i for i in range(5)]
    assert [s for _, s in f.getConData(sk, utcstart=t0 + 300.0, utcend=t0 + 900.0)] == [t0 + 300.0, t0 + 600.0]
    a = [s for _, s in f.getConData(sk, randSamps=3)]
    b = [s for _, s in workflow.ArrayFetcher(case["events"], case["continuous"], sr=case["sr"], seed=3).getConData(sk, randSamps=3)]
    assert len(a) == 3 and a == b and set(a) <= set(starts)
    assert len([1 for _ in f.getConData(sk, randSamps=50)]) == 5            # asks for more than there is
    assert workflow._timestamp("2010-01-01T00-00-00") == workflow._timestamp("2010-01-01 00:00:00") == 1262304000.0


# ------------------------------------------------------------------ the reference's fetcher protocol (ObsPy Streams)
class _FakeDataFetcher(object):
    """Yields what `detex.getdata.DataFetcher` yields (getdata.py:351-453, 455-539): ObsPy-Stream-like lists of
    traces with `.data` and `.stats`; traces out of channel order, one channel starting 3 samples early and
    another ending 5 samples late, as real archives deliver them."""

    def __init__(self, case, sr):
        self.case, self.sr = case, sr
        self.conDatDuration, self.conBuff = 280, 20

    def _stream(self, traces, start):
        import types
        rng = np.random.default_rng(int(start) % 1000)
        chans = ["BHE", "BHN", "BHZ"]
        out = []
        for k, (c, x) in enumerate(zip(chans, traces)):
            pre, post = (3, 0) if k == 1 else ((0, 5) if k == 2 else (0, 0))
            data = np.concatenate([rng.normal(size=pre), x, rng.normal(size=post)])
            st = types.SimpleNamespace(channel=c, sampling_rate=self.sr, npts=len(data), station="M17A",
                                       starttime=types.SimpleNamespace(timestamp=start - pre / self.sr))
            out.append(types.SimpleNamespace(data=data, stats=st))
        return [out[2], out[0], out[1]]

    def getTemData(self, temkey, stakey, tb4=None, taft=None, returnName=True, phases=None, **kw):
        inner = workflow.ArrayFetcher(self.case["events"], self.case["continuous"], sr=self.sr)
        for traces, start, name in inner.getTemData(temkey, stakey):
            yield self._stream(traces, start), name
        yield [], "empty-stream"                               # the reference logs and skips these

    def getConData(self, stakey, utcstart=None, utcend=None, randSamps=None, **kw):
        inner = workflow.ArrayFetcher(self.case["events"], self.case["continuous"], sr=self.sr, seed=5)
        for traces, start in inner.getConData(stakey, utcstart=utcstart, utcend=utcend, randSamps=randSamps):
            yield self._stream(traces, start)


def test_stream_fetcher_adapter_equals_array_fetcher(tmp_path):
    """`workflow.StreamFetcher` in front of a reference-style fetcher: sorted by channel, cut to the common window,
    empty streams skipped -- the cluster tables equal those from the arrays themselves."""
    case = synth.workflow_case(77)
    eng = OracleEngine()
    kw = dict(CCreq=CCREQ, filt=[1, 10, 2, True], stationKey=case["stakey"], templateKey=case["temkey"], trim=[2, 18],
              saveclust=False, engine=eng)
    ref = workflow.createCluster(fetch_arg=workflow.ArrayFetcher(case["events"], case["continuous"], sr=case["sr"]), **kw)
    sf = workflow.StreamFetcher(_FakeDataFetcher(case, case["sr"]))
    got = workflow.createCluster(fetch_arg=sf, **kw)
    assert sf.sr == case["sr"] and sf.channels == ["BHE", "BHN", "BHZ"]
    assert len(got) == len(ref)
    for a, b in zip(ref.clusters, got.clusters):
        assert a.station == b.station and a.clusts == b.clusts and a.singles == b.singles
        assert np.array_equal(np.asarray(a.link), np.asarray(b.link))
    # continuous data through the same adapter
    chunks = list(sf.getConData(case["stakey"].iloc[:1]))
    want = list(workflow.ArrayFetcher(case["events"], case["continuous"], sr=case["sr"], seed=5)
                .getConData(case["stakey"].iloc[:1]))
    assert len(chunks) == len(want) > 0
    for (tr, t0), (wtr, w0) in zip(chunks, want):
        assert t0 == w0 and all(np.array_equal(x, y) for x, y in zip(tr, wtr))


def test_reference_built_clusterstream_is_accepted(tmp_path):
    """`createSubSpace(clust=<the reference's ClusterStream>)`: the reference keeps its constructor arguments as
    attributes (subspace.py:52-59) and its Clusters carry `ccReq` (subspace.py:304-316); the object is re-cut from
    the stored linkage.  The stand-in below has exactly those attributes and a reference-style fetcher."""
    import types
    case = synth.workflow_case(77)
    eng = OracleEngine()
    fetcher = workflow.ArrayFetcher(case["events"], case["continuous"], sr=case["sr"], conDatDuration=280, conBuff=20, seed=5)
    ours = workflow.createCluster(CCreq=CCREQ, fetch_arg=fetcher, filt=[1, 10, 2, True], stationKey=case["stakey"],
                                  templateKey=case["temkey"], trim=[2, 18], saveclust=False, engine=eng)
    ours["TA.M18A"].updateReqCC(0.7)
    ref_like = types.SimpleNamespace(
        trdf=ours.trdf, temkey=ours.temkey, stakey=ours.stakey, fetcher=_FakeDataFetcher(case, case["sr"]),
        eventList=ours.eventList, ccReq=None, filt=ours.filt, decimate=ours.decimate, trim=ours.trim,
        fileName="clust.pkl", filename="clust.pkl", eventsOnAllStations=ours.eventsOnAllStations,
        enforceOrigin=ours.enforceOrigin, stalist=list(ours.stalist),
        clusters=[types.SimpleNamespace(ccReq=c.ccReq, station=c.station) for c in ours.clusters])
    conv = workflow.ClusterStream.from_reference(ref_like)
    for a, b in zip(ours.clusters, conv.clusters):
        assert a.ccReq == b.ccReq and a.clusts == b.clusts and a.singles == b.singles
    ss_ref = workflow.createSubSpace(Pf=1e-10, clust=ours, engine=eng)
    ss_new = workflow.createSubSpace(Pf=1e-10, clust=ref_like, engine=eng)      # events re-read through the Stream fetcher
    for sta in ("TA.M17A", "TA.M18A"):
        a, b = ss_ref.subspaces[sta], ss_new.subspaces[sta]
        assert list(a.Name) == list(b.Name) and [sorted(x) for x in a.Events] == [sorted(x) for x in b.Events]
        for (_, ra), (_, rb) in zip(a.iterrows(), b.iterrows()):
            for ev in ra.Events:
                assert np.allclose(ra.AlignedTD[ev], rb.AlignedTD[ev], rtol=0, atol=1e-9)


def test_public_signatures_match_the_reference():
    """Argument names (and the defaults the reference gives them) of the user-facing calls: a script written for
    Detex passes the same keywords.  Extra keywords here: the engine handles and batch sizes."""
    import inspect
    from oracle import ref_shim
    if not ref_shim.available():
        pytest.skip("reference tree not present")
    ref_shim.load()
    import detex.construct as rc
    import detex.subspace as rs
    pairs = [(rc.createCluster, workflow.createCluster), (rc.createSubSpace, workflow.createSubSpace),
             (rs.SubSpace.SVD, workflow.SubSpace.SVD), (rs.SubSpace.getFAS, workflow.SubSpace.getFAS),
             (rs.SubSpace.detex, workflow.SubSpace.detex), (rs.SubSpace.attachPickTimes, workflow.SubSpace.attachPickTimes),
             (rs.SubSpace.setSinglesThresholds, workflow.SubSpace.setSinglesThresholds),
             (rs.SubSpace.validateClusters, workflow.SubSpace.validateClusters),
             (rs.Cluster.updateReqCC, workflow.Cluster.updateReqCC),
             (rs.ClusterStream.updateReqCC, workflow.ClusterStream.updateReqCC)]
    import detex.results as rr
    import detex.util as ru
    pairs += [(rr.detResults, results.detResults), (ru.saveSQLite, results.saveSQLite),
              (ru.loadSQLite, results.loadSQLite), (ru.readKey, workflow.readKey)]
    for ref, ours in pairs:
        pa, pb = inspect.signature(ref).parameters, inspect.signature(ours).parameters
        assert [k for k in pa if k not in pb] == [], ref.__qualname__
        fixed = [k for k in pa if pa[k].kind is not inspect.Parameter.VAR_KEYWORD]
        assert list(pb)[:len(fixed)] == fixed, ref.__qualname__                      # same order: positional calls work
        for k in pa:
            assert pa[k].default == pb[k].default or pa[k].default is inspect._empty, (ref.__qualname__, k)
        assert set(pb) - set(pa) <= {"engine", "ccx_engine", "batch", "utcstart", "utcend"}


def test_containers_pickles_and_printouts(tmp_path, capsys):
    """The non-plotting conveniences of the reference's containers (subspace.py:203-205, 693-707, 1998-2037) and the
    loaders of util.py:934-969: indexing, iteration, printAtr / printOffsets, write + load round trips."""
    from detex_b200 import util
    case, cl, ss, db, found = _run(OracleEngine(), tmp_path)
    c = cl["TA.M17A"]
    assert [sorted(x) for x in c] == [sorted(x) for x in c.clusts] and len(c) == len(c.clusts)
    cl.printAtr()
    out = capsys.readouterr().out
    assert "TA.M17A Cluster" in out and "Required Cross Correlation Coeficient = %.3f" % CCREQ in out
    assert len(ss) == 2 and ss[0] is ss.subspaces[ss.ssStations[0]] and ss["M18A"] is ss["TA.M18A"]
    with pytest.raises(Exception):
        ss["XX.NOPE"]
    ss.printOffsets()
    out = capsys.readouterr().out
    assert out.count("range=") == sum(len(v) for v in ss.subspaces.values())
    # pickles: the ClusterStream written by createCluster and the SubSpace (without its engine handle)
    cl2 = util.loadClusters(cl.filename)
    assert isinstance(cl2, workflow.ClusterStream) and cl2["TA.M17A"].clusts == c.clusts
    p = str(tmp_path / "subspace.pkl")
    ss.write(p)
    ss2 = util.loadSubSpace(p)
    assert ss2._engine is None and list(ss2.subspaces["TA.M17A"].Name) == list(ss.subspaces["TA.M17A"].Name)
    for (_, a), (_, b) in zip(ss.subspaces["TA.M17A"].iterrows(), ss2.subspaces["TA.M17A"].iterrows()):
        assert a.Threshold == b.Threshold and all(np.array_equal(a.SVD[k], b.SVD[k]) for k in a.SVD)
    with pytest.raises(Exception, match="not a SubSpaceStream"):
        util.loadSubSpace(cl.filename)
