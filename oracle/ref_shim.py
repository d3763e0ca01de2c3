"""ref_shim.py -- TEST INFRASTRUCTURE ONLY (never imported by the product path).

Imports the unmodified reference package from /root/reference under Python 3.12
so its pure-array hot-path functions can be run to (a) validate the oracle
restatement in `oracle/detex_oracle.py` and (b) generate the golden vectors
committed under `tests/golden/` (see `tests/golden/make_golden.py`).

The reference is Python 2.7 / pandas 0.17 / numpy 1.x / scipy 0.18 era code and
needs obspy, matplotlib, PyQt4 ... which are not installed.  The shim stubs the
missing modules and restores the removed aliases; it does not touch the
arithmetic.  `/root/reference` only exists in the build container, so anything
using this module must skip when `available()` is False.
"""
import builtins
import importlib
import os
import sys
import types
from unittest import mock

import numpy as np
import pandas as pd
import scipy
import scipy.fftpack

# the reference tree itself (build container), else the unmodified copy oracle/make_ref.py put under
# oracle/_ref (git-ignored; it travels to the GPU box so that bench.py can time the real reference)
_HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.environ.get("DETEX_REFERENCE", "/root/reference")
if not os.path.isdir(os.path.join(REF_ROOT, "detex")) and os.path.isdir(os.path.join(_HERE, "_ref", "detex")):
    REF_ROOT = os.path.join(_HERE, "_ref")


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "detex"))


def _rolling(fn):
    def f(x, n, center=False, **kw):
        r = pd.Series(np.asarray(x)).rolling(int(n), center=center)
        return getattr(r, fn)().to_numpy(copy=True)  # writable: reference does `b *= n`
    return f


class _UTCDateTime(object):
    """Stand-in for obspy.UTCDateTime on the results path (results.py:421, 478-479): `.timestamp`
    from epoch seconds or an ISO string, `str()` as ObsPy prints it."""

    def __init__(self, x):
        import datetime
        if isinstance(x, (int, float, np.integer, np.floating)):
            self.timestamp = float(x)
        else:
            d = datetime.datetime.fromisoformat(str(x).replace("Z", ""))
            self.timestamp = d.replace(tzinfo=datetime.timezone.utc).timestamp()

    def __str__(self):
        import datetime
        d = datetime.datetime(1970, 1, 1) + datetime.timedelta(microseconds=int(round(self.timestamp * 1e6)))
        return d.strftime("%Y-%m-%dT%H:%M:%S.%fZ")


_loaded = None


def load():
    """Return the reference `detex` package (imported once)."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)
    stubs = [
        "obspy", "obspy.clients", "obspy.clients.fdsn", "obspy.clients.neic",
        "obspy.clients.earthworm", "obspy.signal", "obspy.signal.trigger", "obspy.core",
        "obspy.core.event", "obspy.core.util", "obspy.core.util.attribdict",
        "obspy.core.utcdatetime", "obspy.geodetics",
        "matplotlib", "matplotlib.pyplot", "matplotlib.figure", "matplotlib.transforms",
        "matplotlib.backends", "matplotlib.backends.backend_qt4agg", "matplotlib.colors",
        "matplotlib.patches", "matplotlib.dates", "matplotlib.cm",
        "mpl_toolkits", "mpl_toolkits.basemap",
        "PyQt4", "PyQt4.QtGui", "PyQt4.QtCore", "simplekml", "glob2", "pathlib2",
    ]
    for name in stubs:
        if name not in sys.modules:
            m = mock.MagicMock(name=name)
            m.__path__ = []
            m.__name__ = name
            m.__spec__ = None
            sys.modules[name] = m
    # removed pandas / numpy / scipy aliases the reference still uses
    pd.rolling_mean = _rolling("mean")
    pd.rolling_var = _rolling("var")
    pd.rolling_std = _rolling("std")
    for alias in ("NaN", "NAN", "Nan"):
        if not hasattr(np, alias):
            setattr(np, alias, np.nan)
    if not hasattr(np, "float"):
        np.float = float
    if not hasattr(np, "int"):
        np.int = int
    if not hasattr(np.lib, "pad"):
        np.lib.pad = np.pad
    scipy.real = np.real
    scipy.dot = np.dot
    builtins.reload = importlib.reload
    builtins.xrange = range
    builtins.unicode = str
    import collections
    import collections.abc
    if not hasattr(collections, "Iterable"):
        collections.Iterable = collections.abc.Iterable
    if not hasattr(pd.DataFrame, "iteritems"):       # removed alias of .items() (util.py:925, results.py:519)
        pd.DataFrame.iteritems = pd.DataFrame.items
        pd.Series.iteritems = pd.Series.items
    # the only ObsPy object the results path needs: epoch seconds <-> ISO string
    sys.modules["obspy"].UTCDateTime = _UTCDateTime
    sys.modules["obspy"].core.UTCDateTime = _UTCDateTime
    sys.modules["obspy.core"].UTCDateTime = _UTCDateTime
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import detex  # noqa: E402
    _loaded = detex
    return detex


class RefFunctions(object):
    """Bound handles to the reference's hot-path callables (file:line cited)."""

    def __init__(self):
        d = load()
        import detex.construct as construct
        import detex.detect as detect
        import detex.fas as fas
        self.detex = d
        self.construct = construct
        self.detect = detect
        self.fas = fas
        self._ssd = object.__new__(detect._SSDetex)  # bypass the do-everything __init__

    # detex/detect.py:559-578
    def MPXDS(self, MPcon, U, Nc):
        n = U.shape[1]
        reqlen = int(len(MPcon) + n)
        nfft = 2 ** reqlen.bit_length()
        ssFD = np.array([scipy.fftpack.fft(x[::-1], n=nfft) for x in U])  # detect.py:371
        MPconFD = scipy.fftpack.fft(MPcon, n=nfft)                         # detect.py:256
        return self._ssd._MPXDS(MPcon, reqlen, U, ssFD, Nc, MPconFD)

    # detex/fas.py:120-134
    def MPXSSCorr(self, MPcon, U, Nc):
        n = U.shape[1]
        reqlen = int(len(MPcon) + n)
        nfft = 2 ** reqlen.bit_length()
        ssFD = np.array([scipy.fftpack.fft(x[::-1], n=nfft) for x in U])  # fas.py:171
        return self.fas._MPXSSCorr(MPcon, reqlen, U, ssFD, Nc)

    # detex/construct.py:425-466 (+ :669-676 for the FFT prep)
    def CCX2(self, x1, x2, Nc):
        n = len(x1)
        nfft = 2 ** int(2 * n).bit_length()
        f1 = scipy.fftpack.fft(x1, n=nfft)
        f2 = scipy.fftpack.fft(x2, n=nfft)
        chans = ["C%d" % i for i in range(Nc)]
        return self.construct._CCX2(f1, f2, x1, x2, chans, chans)

    # detex/construct.py:369-394
    def makeDFcclags(self, X, Nc):
        n = X.shape[1]
        nfft = 2 ** int(2 * n).bit_length()
        evs = ["ev%04d" % i for i in range(X.shape[0])]
        chans = ["C%d" % i for i in range(Nc)]
        row = pd.Series({
            "MPtd": {e: X[i] for i, e in enumerate(evs)},
            "MPfd": {e: scipy.fftpack.fft(X[i], n=nfft) for i, e in enumerate(evs)},
            "Channels": {e: chans for e in evs},
        })
        return self.construct._makeDFcclags(evs, row)

    # detex/construct.py:397-422
    def subSamp(self, Ceval, ind):
        return self.construct._subSamp(Ceval, ind)

    # detex/detect.py:501-524
    def getStaLtaArray(self, C, LTA, STA):
        return self._ssd._getStaLtaArray(np.array(C, dtype=float), LTA, STA)

    # detex/detect.py:545-557
    def downPlay(self, C, sr, dpv=0, buff=20):
        return self._ssd._downPlayArrayAroundMax(C, sr, dpv, buff)

    # detex/detect.py:390-445 with estimateMags=False, trigCon=0
    def CreateCoeffArray(self, DS, stalta, sr, start, thr, offsets, name="SS0", sta="STA"):
        s = self._ssd
        s.trigCon = 0
        s.fillZeros = False
        s.estimateMags = False
        cs = pd.Series({"SSdetect": DS, "STALTA": stalta, "SampRate": sr, "TimeStamp": start,
                        "Nc": 1})
        return s._CreateCoeffArray(cs, name, {name: thr}, sta, {name: offsets}, {name: None},
                                   {name: None}, None, {name: None}, None, {name: None},
                                   {name: None})

    # detex/detect.py:447-499
    def estMag(self, trigIndex, MPcon, Nc, U, ewf, mags, issubspace=True):
        s = self._ssd
        s.issubspace = issubspace
        U = np.atleast_2d(U)
        UtU = np.dot(np.transpose(U), U)                 # detect.py:367
        WFU = np.dot(np.atleast_2d(ewf), UtU)            # detect.py:381
        cs = pd.Series({"Nc": Nc})
        evs = ["e%d" % i for i in range(len(mags))]
        return s._estMag(trigIndex, cs, MPcon, np.asarray(mags, dtype=float), evs, WFU, UtU, np.atleast_2d(ewf), 0.0,
                         0.0, "SS0", "STA")

    # detex/subspace.py:875-905 (the body of SubSpace.SVD for one subspace row): `_trimGroups` :921-943,
    # scipy.linalg.svd :890, svdDict :892-893, `_getFracEnergy` :968-997, `_getUsedBasis` :999-1013 --
    # the reference's own methods, unmodified.  They never touch `self`; the only Python-2 idiom in
    # them is `d.keys().sort()`, which the inputs satisfy with a dict whose keys() is a list.
    def svdSelect(self, aligned, start, end, selectCriteria, selectValue, normalize=False):
        import detex.subspace as subspace
        import scipy.linalg

        class Py2Dict(dict):
            def keys(self):
                return list(dict.keys(self))

        SS = subspace.SubSpace
        evs = ["ev%03d" % i for i in range(len(aligned))]
        row = types.SimpleNamespace(Events=list(evs), Name="SS0",
                                    AlignedTD={e: np.asarray(aligned[i], dtype=float) for i, e in enumerate(evs)},
                                    SampleTrims=Py2Dict(Starttime=int(start), Endtime=int(end)))
        keys = sorted(row.Events)
        arr, basisLength = SS._trimGroups(None, 0, row, keys, "STA")          # subspace.py:879
        if normalize:
            arr = np.array([x / np.linalg.norm(x) for x in arr])               # :887-888
        U, s, Vh = scipy.linalg.svd(np.transpose(arr), full_matrices=False)   # :889-890
        svdDict = Py2Dict()
        for einum, eival in enumerate(s):                                      # :892-893
            svdDict[eival] = U[:, einum]
        frac = SS._getFracEnergy(None, 0, row, svdDict, U)                     # :896
        used = SS._getUsedBasis(None, 0, row, svdDict, frac, selectCriteria, selectValue)   # :898-899
        return dict(basisLength=basisLength, s=s, Ufull=U, frac_avg=np.asarray(frac['Average'], dtype=float),
                    frac_min=np.asarray(frac['Minimum'], dtype=float), used_keys=np.asarray(used, dtype=float),
                    U=np.array([svdDict[k] for k in used]).reshape(len(used), basisLength))

    # detex/construct.py:928-987
    def multiplex(self, chans):
        st = [types.SimpleNamespace(data=np.asarray(c)) for c in chans]
        return self.construct.multiplex(st, Nc=len(chans))

    # detex/construct.py:469-483
    def fast_normcorr(self, t, s):
        return self.construct.fast_normcorr(t, s)

    # detex/construct.py:272-281, 710-786, 486-503: linkage -> per-event alignment delays.
    # `_getDelays` itself stores a list into a DataFrame cell through `.loc` (construct.py:725),
    # which today's pandas refuses, so its bookkeeping lines 711-728 are replayed here with the
    # same content and the arithmetic -- `_traceEventDendro`, `_updateLags`, `_getDow`, `_getAcr`,
    # `_makeCC2LagMap`, `_getClustDict`, `_ensureUnique`, `_flatNoNan`, `_alignTD` -- is the
    # reference's own, unmodified.
    def getDelays(self, DFcc, DFlag):
        from scipy.cluster.hierarchy import linkage
        C = self.construct
        cxdf = 1.0000001 - DFcc
        cx = C._flatNoNan(cxdf)
        cx, cxdf = C._ensureUnique(cx, cxdf)
        lags = C._flatNoNan(DFlag)
        link = linkage(cx)
        CCtoLag = C._makeCC2LagMap(cx, lags)
        N = len(link)
        linkup = np.append(link, np.arange(N + 1, 2 * N + 1).reshape(N, 1), 1)
        clustDict = C._getClustDict(linkup, len(linkup))
        rows = []
        for r in linkup:
            tempdf = cxdf[cxdf == r[2]].dropna(how='all').dropna(axis=1)
            rows.append(dict(i1=r[0], i2=r[1], cc=r[2], num=r[3], clust=r[4],
                             II=clustDict[int(r[0])].tolist() + clustDict[int(r[1])].tolist(),
                             ev1=tempdf.index[0], ev2=tempdf.columns[0]))
        dflink = pd.DataFrame(rows).astype(object)
        delays = C._traceEventDendro(dflink, cx, lags, CCtoLag, clustDict, clustDict.iloc[-1])
        return link, np.asarray(delays.values, dtype=np.int64)

    # detex/construct.py:283-286, 486-503
    def alignTD(self, delays, X):
        evs = ["ev%04d" % i for i in range(len(X))]
        delayNP = -1 * np.min(delays)
        delayDF = pd.DataFrame(np.asarray(delays) + delayNP, columns=['SampleDelays'])
        delayDF['Events'] = [evs[x] for x in delayDF.index]
        srow = pd.Series({"MPtd": {e: X[i] for i, e in enumerate(evs)}, "Station": "STA"})
        al = self.construct._alignTD(delayDF, srow)
        return np.array([al[e] for e in evs])

    # detex/util.py:870-893 (pandas_dbms.write_frame / get_schema)
    def saveSQLite(self, DF, db, table):
        import detex.util
        return detex.util.saveSQLite(DF, db, table)

    # detex/results.py:371-401 (reads the table back through detex.util.loadSQLite)
    def deleteDetDups(self, db, associateBuffer, table="ss_df", trigCon=0, trigParameter=0.0):
        import detex.results
        return detex.results._deleteDetDups(db, trigCon, trigParameter, associateBuffer, None, None, None, table)

    # detex/results.py:404-466, associateReq = 0
    def associateDetections(self, ssdf, requiredNumStations, associateBuffer, temkey, exceptionalThreshold=None):
        import detex.results
        return detex.results._associateDetections(ssdf.copy(), 0, requiredNumStations, associateBuffer, None,
                                                  temkey.copy(), exceptionalThreshold)
