"""Live pins of the workflow / results host code against the UNMODIFIED reference under the py3 shim
(only where /root/reference exists; the pieces below are the ones that run under the shim)."""
import warnings

import numpy as np
import pandas as pd
import pytest

from detex_b200 import results as ours, subspace as osub, synth, workflow
from oracle import ref_shim

pytestmark = pytest.mark.skipif(not ref_shim.available(), reason="reference tree not present")


@pytest.fixture(scope="module")
def detex():
    ref_shim.load()
    import detex
    return detex


def _row():
    stats = {'a': {'Nc': 3, 'sampling_rate': 40.0, 'starttime': 100.0, 'origintime': 98.0},
             'b': {'Nc': 3, 'sampling_rate': 40.0, 'starttime': 200.3, 'origintime': 197.0},
             'c': {'Nc': 3, 'sampling_rate': 40.0, 'starttime': 300.0, 'origintime': 299.0}}
    return pd.Series({'Events': ['a', 'b', 'c'], 'Stats': stats, 'Station': 'TA.X', 'Name': 'SS0',
                      'AlignedTD': {k: np.zeros(2400) for k in 'abc'}, 'SampleTrims': {}})


@pytest.mark.parametrize("fun", [np.median, np.mean, np.min, np.max])
@pytest.mark.parametrize("duration", [8, 30, None])
def test_getSampTrim_and_offsets(detex, fun, duration):
    """subspace.py:1547-1637."""
    pk = pd.DataFrame({'TimeStamp': [105.0, 109.5, 205.55, 212.0, 299.0, 304.0], 'Station': 'TA.X',
                       'Event': ['a', 'a', 'b', 'b', 'c', 'c'], 'Phase': ['P', 'S'] * 3})
    ss = object.__new__(detex.subspace.SubSpace)
    r1, r2 = _row(), _row()
    eves, st, Nc, Sr = ss._getStats(r1)
    DF = pd.DataFrame({'Stats': [r1.Stats]})
    want = ss._getSampTrim(eves, st, Nc, Sr, pk, duration, fun, 'TA.X', 0, DF, r1)
    got = workflow.SubSpace._getSampTrim(r2, pk, duration, fun, 'TA.X')
    assert got == want
    for k in 'abc':                                        # start times / offsets written back to Stats
        assert r2.Stats[k]['offset'] == r1.Stats[k]['offset'] and r2.Stats[k]['Starttime'] == r1.Stats[k]['Starttime']
    for offs in ([1.0, 1.1, 0.9, 50.0], [2.0], [0.5, 0.5, 0.5], [3.0, 1.0]):
        a = ss._getOffsets(np.array(offs))
        b = workflow.SubSpace._getOffsets(offs)
        assert [float(x) for x in a] == [float(x) for x in b]


def test_pick_beyond_waveform_is_skipped(detex):
    pk = pd.DataFrame({'TimeStamp': [100.0 + 30.0], 'Station': 'TA.X', 'Event': ['a'], 'Phase': ['P']})
    ss = object.__new__(detex.subspace.SubSpace)
    r1, r2 = _row(), _row()
    eves, st, Nc, Sr = ss._getStats(r1)
    assert ss._getSampTrim(eves, st, Nc, Sr, pk, 8, np.median, 'TA.X', 0, pd.DataFrame({'Stats': [r1.Stats]}), r1) is None
    assert workflow.SubSpace._getSampTrim(r2, pk, 8, np.median, 'TA.X') is None


def test_threshold_grid_search_and_pfkey(detex):
    """subspace.py:1110-1140, results.py:176-229."""
    ss = object.__new__(detex.subspace.SubSpace)
    for a, b, pf in ((2.0, 800.0, 1e-12), (0.5, 4000.0, 1e-9), (4.0, 4500.0, 1e-6)):
        x, p = ss._approxThld(a, b, 'TA.X', pd.Series({'Name': 'SS0'}), pf, 1000, 3, None)
        x2, p2 = osub._approxThld(a, b, pf, 1000, 3, None)
        assert x == x2 and p == p2
        assert detex.results._approximateThreshold(a, b, pf, 1000, 3) == ours._approximateThreshold(a, b, pf, 1000, 3)
    info = pd.DataFrame({'Name': ['SS0', 'SS1'], 'Sta': ['TA.A', 'TA.B'], 'beta1': [2.0, 0.5], 'beta2': [800.0, 4000.0]})
    want, none = detex.results._makePfKey(info, None, 1e-8)
    got = ours.makePfKey(info, 1e-8)
    assert none is None and ours.makePfKey(None, 1e-8) is None and ours.makePfKey(info, False) is None
    assert list(got.columns) == list(want.columns) and np.array_equal(got.DS.values, want.DS.values)
    assert list(got.Sta) == list(want.Sta) and list(got.betadist) == list(want.betadist)


def test_verify_events(tmp_path):
    """results.py:232-293.  The reference's `_verifyEvents` does not run under pandas 3 (`Series & list`,
    results.py:255), so the behaviour is pinned by its stated rules: a catalogue event verifies the
    not-yet-verified detection with the highest DSav whose origin window +- veriBuffer/2 contains it;
    auto-detections are only searched when no new detection matches."""
    det, temkey = synth.detection_table(5, nev=30)
    D, A = ours.associateDetections(ours.deleteDetDups(det, 1.0), 2, 1.0, temkey)
    assert len(D) > 4 and len(A) > 0
    veri = pd.DataFrame({'TIME': [float(D.MSTAMPmin.iloc[0]) + 0.1, float(D.MSTAMPmin.iloc[3]) + 0.2,
                                  float(A.MSTAMPmin.iloc[0]) + 0.1, 5.0],
                         'LAT': 1.0, 'LON': 2.0, 'MAG': [1., 2., 3., 4.], 'DEPTH': 5.0, 'NAME': ['v0', 'v1', 'v2', 'v3'],
                         'EXTRA': [7, 8, 9, 10]})
    path = str(tmp_path / "veri.csv")
    veri.to_csv(path, index=False)
    D2, A2 = D.copy(), A.copy()
    got = ours.verifyEvents(D2, A2, path, 1, True)
    assert list(np.nonzero(D2.Verified.values)[0]) == [0, 3] and list(np.nonzero(A2.Verified.values)[0]) == [0]
    assert list(got.VerName) == ['v0', 'v1', 'v2'] and list(got.EXTRA) == [7, 8, 9] and list(got.VerMag) == [1., 2., 3.]
    assert list(got.Event) == [D.Event.iloc[0], D.Event.iloc[3], A.Event.iloc[0]] and 'Verified' not in got.columns
    again = ours.verifyEvents(D2, A2, veri, 1, False)        # everything that matches is already verified
    assert len(again) == 0
    assert ours.verifyEvents(D2, A2, None) is None
    with pytest.raises(Exception):
        ours.verifyEvents(D2, A2, veri.drop(columns=['MAG']))
