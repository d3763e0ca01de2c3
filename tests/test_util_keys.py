"""`detex_b200.util.readKey` against `detex.util.readKey` (util.py:564-627): same rows kept, same columns and
dtypes, station / network codes as strings, same errors; the row ORDER is documented to differ (the reference
sorts by the iteration order of a Python-2 set)."""
import glob
import os

import numpy as np
import pandas as pd
import pytest

from detex_b200 import util


def _write(tmp_path, name, text):
    p = tmp_path / name
    p.write_text(text)
    return str(p)


TEMKEY = """TIME,NAME,LAT,LON,MAG,DEPTH,COMMENT
2014-03-02T10-11-12.50,2014-03-02T10-11-12.50,41.10,-111.20,1.9,7.5,b
2014-03-01T01-02-03.00,2014-03-01T01-02-03.00,41.00,-111.00,2.1,5.0,a
2014-03-03T20-21-22.25,2014-03-03T20-21-22.25,41.20,-111.40,0.7,9.1,c
"""
STAKEY = """NETWORK,STATION,STARTTIME,ENDTIME,LAT,LON,ELEVATION,CHANNELS
TA,M18A,2014-03-01T00-00-00,2014-03-05T00-00-00,41.4,-110.1,2100,BHE-BHN-BHZ
7A,1234,2014-03-01T00-00-00,2014-03-05T00-00-00,41.2,-110.6,1900,BHE-BHN-BHZ
"""
PICKS = """,TimeStamp,Station,Event,Phase
0,1393635725.5,TA.M18A,2014-03-01T01-02-03.00,S
1,1393635723.0,TA.M18A,2014-03-01T01-02-03.00,P
"""


def test_read_key_files(tmp_path):
    tem = util.readKey(_write(tmp_path, "TemplateKey.csv", TEMKEY), "template")
    assert list(tem.NAME) == sorted(tem.NAME) and list(tem.index) == [0, 1, 2] and "COMMENT" in tem.columns
    sta = util.readKey(_write(tmp_path, "StationKey.csv", STAKEY), "station")
    assert list(sta.STATION) == ["1234", "M18A"] and all(isinstance(x, str) for x in sta.STATION)
    assert list(sta.NETWORK) == ["7A", "TA"]
    pk = util.readKey(_write(tmp_path, "PhasePicks.csv", PICKS), "phases")
    assert list(pk.Phase) == ["P", "S"]
    # a DataFrame passes through the same checks
    again = util.readKey(tem, "template")
    assert again.equals(tem)


def test_read_key_errors(tmp_path):
    with pytest.raises(Exception, match="does not exists"):
        util.readKey(str(tmp_path / "nope.csv"), "template")
    with pytest.raises(Exception, match="Required columns"):
        util.readKey(pd.DataFrame({"TIME": [1], "NAME": ["a"]}), "template")
    with pytest.raises(Exception, match="unsported key type"):
        util.readKey(pd.DataFrame(), "events")
    with pytest.raises(Exception, match="not understood"):
        util.readKey(42, "station")
    # rows with an empty required field are dropped (util.py:614-617)
    df = pd.DataFrame({"TimeStamp": [1.0, 2.0], "Event": ["a", ""], "Station": ["N.S", "N.S"], "Phase": ["P", "S"]})
    assert len(util.readKey(df, "phases")) == 1


def test_against_the_reference_on_its_own_key_files():
    from oracle import ref_shim
    root = os.path.join(ref_shim.REF_ROOT, "tests")
    files = [(f, "station") for f in glob.glob(os.path.join(root, "**", "StationKey*.csv"), recursive=True)]
    files += [(f, "template") for f in glob.glob(os.path.join(root, "**", "TemplateKey*.csv"), recursive=True)]
    files += [(f, "phases") for f in glob.glob(os.path.join(root, "**", "*Picks*.csv"), recursive=True)]
    if not ref_shim.available() or not files:
        pytest.skip("reference tree with its test data not present")
    d = ref_shim.load()
    import detex.util as rutil
    checked = 0
    for path, kind in files:
        try:
            want = rutil.readKey(path, kind)
        except Exception:
            with pytest.raises(Exception):
                util.readKey(path, kind)
            continue
        got = util.readKey(path, kind)
        assert list(got.columns) == list(want.columns) and len(got) == len(want)
        cols = list(util.REQ_COLUMNS[kind])
        a = got.sort_values(by=cols).reset_index(drop=True)
        b = want.sort_values(by=cols).reset_index(drop=True)
        for c in got.columns:
            if a[c].dtype.kind == "f":
                assert np.allclose(a[c].to_numpy(), b[c].to_numpy(), equal_nan=True)
            else:
                assert list(a[c].astype(str)) == list(b[c].astype(str))
        checked += 1
    assert checked >= 3
