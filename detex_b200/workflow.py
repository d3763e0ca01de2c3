"""workflow.py -- the reference's user-facing sequence on top of the CUDA engine:

    createCluster -> ClusterStream / Cluster.updateReqCC -> createSubSpace ->
    SubSpace.attachPickTimes / SVD / getFAS / detex -> results tables

with the reference's names, argument meaning, containers and error behaviour
(reference detex/construct.py:25-301, detex/subspace.py:46-420, 715-1140, 1443-1995,
detex/fas.py:23-117, detex/detect.py:27-218).  Everything that touches samples runs through
`Engine`: pre-processing (K8), pairwise CCX (K4), zero-lag validation, FAS screen (K6) and
statistics (K1 + K5), detection (K0/K1/K3), magnitudes (K7).  What stays on the host is what the
reference does in pandas / SciPy on small tables: linkage, dendrogram cut, alignment delays, SVD of
an (events x n) matrix, the beta fit from five sums, SQLite rows.

Out of scope: ObsPy streams and the DataFetcher back ends (directories,
FDSN, NEIC, Earthworm).  `ArrayFetcher` is the in-memory stand-in with the two generators the
path calls (`getTemData`, `getConData`; getdata.py:351, 455): it hands out raw per-channel arrays
and start times, which is all the hot path ever reads from a Stream.

There is no CPU fallback here either: every numeric step calls the engine.
"""
import logging
import os
import pickle
import random
import re

import numpy as np
import pandas as pd
from scipy.cluster.hierarchy import fcluster, linkage

from . import construct, fas as _fas, preprocess, results, subspace as _subspace
from .detect import HIST_BINS, SSDetex, default_engine

log = logging.getLogger("detex_b200")
from .util import readKey  # noqa: E402


def _error(msg, e=Exception):
    """`detex.log(..., level='error')` raises (detex/__init__.py:131-138)."""
    log.error(msg)
    raise e(msg)


def _timestamp(t):
    """Epoch seconds from whatever a key holds: a number, or a UTC date string -- including the
    reference's own 'YYYY-MM-DDTHH-MM-SS' spelling of template names / times (util.py:574-627)."""
    if isinstance(t, (int, float, np.integer, np.floating)):
        return float(t)
    txt = str(t)
    m = re.match(r'^(\d{4}-\d{2}-\d{2})[T ](\d{2})-(\d{2})-(\d{2}(?:\.\d+)?)$', txt)
    if m:
        txt = '%s %s:%s:%s' % m.groups()
    ts = pd.Timestamp(txt)
    if ts.tzinfo is None:
        ts = ts.tz_localize('UTC')
    return ts.timestamp()


# =====================================================================================
# data source
# =====================================================================================
class ArrayFetcher(object):
    """In-memory stand-in for `detex.getdata.DataFetcher` (getdata.py:244-609).

    events     : {'NET.STA': {event name: (list of channel arrays, starttime epoch s)}}
                 raw (unfiltered) windows already cut around the origin time, channels in sorted
                 channel-name order (what `st.sort()` gives, construct.py:998)
    continuous : {'NET.STA': [(list of channel arrays, starttime epoch s), ...]} chronological chunks
                 of conDatDuration + conBuff seconds (getdata.py:299, 517-518)
    The channels of one entry share ONE start time (a single epoch per chunk is carried): the
    common-window trim of `_applyFilter` (construct.py:1019-1024) then reduces to cutting every
    channel to the shortest one, which the device pre-processing does before detrend and filter.
    """
    method = 'array'

    def __init__(self, events=None, continuous=None, sr=100.0, channels=('BHE', 'BHN', 'BHZ'),
                 conDatDuration=3600, conBuff=120, seed=None):
        self.events = events or {}
        self.continuous = continuous or {}
        self.sr = float(sr)
        self.channels = list(channels)
        self.conDatDuration = conDatDuration
        self.conBuff = conBuff
        self._rng = random.Random(seed)   # the reference's random.sample is unseeded (getdata.py:895)

    def getTemData(self, temkey, stakey, tb4=None, taft=None, returnName=True, phases=None):
        for _, srow in stakey.iterrows():
            sta = '%s.%s' % (srow.NETWORK, srow.STATION)
            have = self.events.get(sta, {})
            for name in temkey.NAME:
                if name in have:
                    traces, start = have[name]
                    yield traces, float(start), name

    def getConData(self, stakey, utcstart=None, utcend=None, randSamps=None):
        for _, srow in stakey.iterrows():
            sta = '%s.%s' % (srow.NETWORK, srow.STATION)
            chunks = self.continuous.get(sta, [])
            t1 = -np.inf if utcstart is None else _timestamp(utcstart)
            t2 = np.inf if utcend is None else _timestamp(utcend)
            idx = [i for i, (_, start) in enumerate(chunks) if t1 <= start < t2]
            if randSamps is not None:
                idx = self._rng.sample(idx, min(int(randSamps), len(idx)))
            for i in idx:
                traces, start = chunks[i]
                yield traces, float(start)


class StreamFetcher(object):
    """Adapter: the reference's own `detex.getdata.DataFetcher` (getdata.py:244-609) -- or anything else whose
    `getTemData` / `getConData` yield ObsPy Streams the way it does (getdata.py:351-453, 455-539) -- in front of
    the array protocol of this module, so `createCluster` / `createSubSpace` / `SubSpace.getFAS / detex` take the
    fetcher a Detex user already has.  Only duck-typed trace attributes are touched (`tr.data`,
    `tr.stats.channel / sampling_rate / starttime.timestamp / npts`), no ObsPy import.

    Per stream: traces sorted by channel name (`st.sort()`, construct.py:998) and cut to their common window
    [latest start, earliest end] (construct.py:1019-1024) -- the one start time per entry the array protocol
    carries; streams that are empty, have fewer traces than the first good one, or whose channels do not
    overlap are skipped with a warning, as `_applyFilter` returns an empty Stream for them
    (construct.py:1003-1022).  With `decimate` and channels that start at different samples the reference
    decimates before it trims; here the trim comes first (edge samples of such a stream may differ)."""
    method = 'stream'

    def __init__(self, fetcher, conDatDuration=None, conBuff=None):
        self.fetcher = fetcher
        self.sr = None               # learnt from the first stream
        self.channels = None
        self.conDatDuration = conDatDuration if conDatDuration is not None else getattr(fetcher, 'conDatDuration', 3600)
        self.conBuff = conBuff if conBuff is not None else getattr(fetcher, 'conBuff', 120)

    def _arrays(self, st):
        if st is None:
            return None
        trs = sorted(list(st), key=lambda tr: str(tr.stats.channel))
        if len(trs) < 1 or (self.channels is not None and len(trs) != len(self.channels)):
            log.warning('stream with %d traces skipped', len(trs))
            return None
        sr = float(trs[0].stats.sampling_rate)
        if any(float(tr.stats.sampling_rate) != sr for tr in trs) or (self.sr is not None and sr != self.sr):
            log.warning('stream with mixed sampling rates skipped')
            return None
        starts = [float(tr.stats.starttime.timestamp) for tr in trs]
        ends = [b + (len(tr.data) - 1) / sr for tr, b in zip(trs, starts)]
        t0, t1 = max(starts), min(ends)
        if t0 > t1:
            log.warning('channels do not overlap, stream skipped')
            return None
        npts = int(round((t1 - t0) * sr)) + 1
        out = []
        for tr, b in zip(trs, starts):
            i0 = int(round((t0 - b) * sr))
            out.append(np.asarray(tr.data[i0:i0 + npts], dtype=np.float64))
        npts = min(len(x) for x in out)
        if self.sr is None:
            self.sr = sr
            self.channels = [str(tr.stats.channel) for tr in trs]
        return [x[:npts] for x in out], t0

    def getTemData(self, temkey, stakey, tb4=None, taft=None, returnName=True, phases=None):
        for st, name in self.fetcher.getTemData(temkey, stakey, tb4, taft, returnName=True, phases=phases):
            got = self._arrays(st)
            if got is not None:
                yield got[0], got[1], name

    def getConData(self, stakey, utcstart=None, utcend=None, randSamps=None):
        for st in self.fetcher.getConData(stakey, utcstart=utcstart, utcend=utcend, randSamps=randSamps):
            got = self._arrays(st)
            if got is not None:
                yield got


def _filter_multiplex(traces_list, sr, filt, decimate, engine):
    """`_applyFilter` + `multiplex` (construct.py:990-1030, 928-987) of a batch on the device;
    returns the multiplexed float64 arrays."""
    preprocess.applyFilter(traces_list, sr, filt, decimate=decimate, engine=engine)
    return [engine.get_chunk(i) for i in range(len(traces_list))]


# =====================================================================================
# createCluster
# =====================================================================================
def _checkClusterInputs(filt, dtype, trim, decimate):
    """construct.py:304-324."""
    if filt is not None and len(filt) != 4:
        _error('filt must be a list of length 4')
    if dtype not in ('double', 'single'):
        _error("dtype must be 'double' or 'single'")
    if len(trim) != 2:
        _error('trim must be a list of length 2')
    if decimate is not None and not isinstance(decimate, int):
        _error('decimate must be an int or None')


def _loadEvents(fetcher, filt, trim, stakey, temkey, decimate, dtype, engine, batch=64):
    """`_loadEvents` / `_loadStream` / `_getTimeDomainWFs` / `_testStreamLengths`
    (construct.py:615-698, 852-925).  MPfd (the FFT copies, construct.py:669-676) is not built: the
    GPU path correlates in the time domain."""
    rows = []
    for _, srow in stakey.iterrows():
        station = '%s.%s' % (srow.NETWORK, srow.STATION)
        csta = stakey[stakey.STATION == srow.STATION]
        names, raw, starts = [], [], []
        for traces, start, name in fetcher.getTemData(temkey, csta, trim[0], trim[1]):
            if traces is None or len(traces) < 1:
                continue
            names.append(name)
            raw.append([np.asarray(t) for t in traces])
            starts.append(start)
        MPtd, stats, chans, allzeros, lens = {}, {}, {}, [], {}
        for i in range(0, len(names), batch):
            mp = _filter_multiplex(raw[i:i + batch], fetcher.sr, filt, decimate, engine)
            for name, x, tr, st in zip(names[i:i + batch], mp, raw[i:i + batch], starts[i:i + batch]):
                if dtype == 'single':
                    x = x.astype(np.float32)
                Nc = len(tr)
                MPtd[name] = x
                chans[name] = list(fetcher.channels[:Nc])
                stats[name] = {'processing': ['detrend(linear)', 'filter(bandpass)'],
                               'sampling_rate': fetcher.sr / (decimate or 1), 'starttime': st, 'Nc': Nc}
                lens[name] = sum(len(t) for t in tr)
                if any(not np.any(t) for t in tr):
                    allzeros.append(name)
        if lens:
            mlen = np.median(list(lens.values()))
            for key in [k for k in MPtd if lens[k] < mlen * .2]:      # construct.py:899-906
                log.warning('%s is fractured or missing data, removing', key)
                MPtd.pop(key); chans.pop(key); stats.pop(key)
        for key in set(allzeros):                                     # construct.py:908-913
            log.warning('%s has at least one channel that is all zeros, deleting', key)
            MPtd.pop(key, None); chans.pop(key, None); stats.pop(key, None)
        if len(MPtd) < 2:                                             # construct.py:915-919
            log.warning('Less than 2 events survived preprocessing for station %s', station)
            continue
        events = sorted(MPtd.keys())
        # _testStreamLengths (construct.py:679-698): common length, drop the out-of-tolerance ones
        ln = np.array([len(MPtd[k]) for k in events])
        le = int(np.min(ln[ln > np.median(ln) * .9]))
        for key in [k for k in events if len(MPtd[k]) < le]:
            log.warning('%s on %s is out of length tolerance, removing', key, station)
            MPtd.pop(key)
        events = [k for k in events if k in MPtd]
        for k in events:
            MPtd[k] = MPtd[k][:le]
        rows.append({'Station': station, 'Events': events, 'MPtd': MPtd, 'MPfd': {k: None for k in events},
                     'Channels': {k: chans[k] for k in events}, 'Stats': {k: stats[k] for k in events},
                     'Link': None, 'Clust': None, 'Lags': None, 'Subsamp': None, 'CCs': None,
                     'numEvents': len(events)})
    cols = ['Events', 'MPtd', 'MPfd', 'Channels', 'Stats', 'Link', 'Clust', 'Lags', 'Subsamp', 'CCs',
            'numEvents', 'Station']
    TRDF = pd.DataFrame(rows, columns=cols).astype(object)
    if len(TRDF):
        TRDF = TRDF.sort_values(by='Station').reset_index(drop=True)
    return TRDF


def createCluster(CCreq=0.5, fetch_arg='EventWaveForms', filt=[1, 10, 2, True], stationKey='StationKey.csv',
                  templateKey='TemplateKey.csv',
                  trim=[10, 120], saveclust=True, fileName='clust.pkl', decimate=None, dtype='double',
                  eventsOnAllStations=False, enforceOrigin=False, fillZeros=False, phases=None,
                  engine=None, ccx_engine="tcgen05"):
    """`detex.createCluster` (construct.py:25-171).  `fetch_arg` is an ArrayFetcher, a StreamFetcher, or a
    reference-style DataFetcher (anything with getTemData / getConData yielding ObsPy Streams: wrapped in a
    StreamFetcher); the keys are DataFrames (STATION/NETWORK/... and NAME/TIME/MAG/..., util.py:574-627)."""
    if not isinstance(fetch_arg, (ArrayFetcher, StreamFetcher)):
        if hasattr(fetch_arg, 'getTemData') and hasattr(fetch_arg, 'getConData'):
            fetch_arg = StreamFetcher(fetch_arg)
        else:
            _error('fetch_arg must be an ArrayFetcher, a StreamFetcher or a DataFetcher-like object with '
                   'getTemData / getConData (directory names and client strings are resolved by the reference\'s '
                   'own DataFetcher)', TypeError)
    if enforceOrigin or fillZeros or phases is not None:
        raise NotImplementedError('enforceOrigin / fillZeros / phases act on ObsPy streams inside the '
                                  'DataFetcher; hand ArrayFetcher windows that already reflect them')
    eng = engine or default_engine()
    # paths or DataFrames, as the reference takes them (construct.py:104-106, util.readKey)
    stakey = readKey(stationKey, 'station')
    temkey = readKey(templateKey, 'template')
    _checkClusterInputs(filt, dtype, trim, decimate)
    fetcher = fetch_arg
    TRDF = _loadEvents(fetcher, filt, trim, stakey, temkey, decimate, dtype, eng)
    if len(TRDF) < 1:
        _error('No events survived pre-processing, check DataFetcher and event quality')
    if eventsOnAllStations:
        eventList = sorted(set.intersection(*[set(x) for x in TRDF.Events]))
        if len(eventList) < 2:
            _error('less than 2 events in population have required stations')
    for ind, row in TRDF.iterrows():
        if not eventsOnAllStations:
            eventList = row.Events
        if len(row.Events) < 2:
            log.warning('Less than 2 valid events on station %s', row.Station)
            continue
        DFcc, DFlag, DFsubsamp = construct._makeDFcclags(eventList, row, engine=eng, kernel=ccx_engine)
        TRDF.at[ind, 'Lags'] = DFlag
        TRDF.at[ind, 'CCs'] = DFcc
        TRDF.at[ind, 'Subsamp'] = DFsubsamp
        TRDF.at[ind, 'Link'] = construct.cluster_link(DFcc)          # construct.py:152-157
    trdf = TRDF[['Station', 'Link', 'CCs', 'Lags', 'Subsamp', 'Events', 'Stats']]
    eventListAll = sorted(set.union(*[set(x) for x in TRDF.Events]))
    clust = ClusterStream(trdf, temkey, stakey, fetcher, eventListAll, CCreq, filt, decimate, trim, fileName,
                          eventsOnAllStations, enforceOrigin)
    clust._TRDF = TRDF     # the event waveforms, so createSubSpace does not have to filter them again
    if saveclust:
        clust.write()
    return clust


class ClusterStream(object):
    """`detex.subspace.ClusterStream` (subspace.py:46-287) without the plotting / hypoDD writers."""

    def __init__(self, trdf, temkey, stakey, fetcher, eventList, ccReq, filt, decimate, trim, fileName,
                 eventsOnAllStations, enforceOrigin):
        self.trdf, self.temkey, self.stakey, self.fetcher = trdf, temkey, stakey, fetcher
        self.eventList, self.filt, self.decimate, self.trim = eventList, filt, decimate, trim
        self.eventsOnAllStations, self.enforceOrigin = eventsOnAllStations, enforceOrigin
        self.ccReq = None
        self.stalist = trdf.Station.values.tolist()
        self.stalist2 = [x.split('.')[1] for x in self.stalist]
        self.filename = fileName
        self.clusters = []
        for _, row in trdf.iterrows():
            evlist = eventList if eventsOnAllStations else row.Events
            self.clusters.append(Cluster(self, row.Station, temkey, evlist, row.Link, ccReq, filt, decimate,
                                         trim, row.CCs))

    @classmethod
    def from_reference(cls, obj, fetcher=None):
        """A ClusterStream the REFERENCE built (`detex.createCluster`, or unpickled from its `clust.pkl` with Detex
        importable) as one of ours.  The reference's constructor keeps all its arguments as attributes
        (subspace.py:52-59: trdf, temkey, stakey, fetcher, eventList, filt, decimate, trim, fileName,
        eventsOnAllStations, enforceOrigin) and its `trdf` carries the same Station / Link / CCs / Lags / Subsamp /
        Events / Stats columns `createCluster` fills here (construct.py:139-168), so the dendrograms are re-cut
        from the stored linkage at each station's stored `ccReq` (subspace.py:304-345) -- no waveform is touched.
        `fetcher`: replacement for the pickled DataFetcher (default: the object's own, wrapped in a
        StreamFetcher when first used)."""
        f = fetcher if fetcher is not None else getattr(obj, 'fetcher', None)
        new = cls(obj.trdf, obj.temkey, obj.stakey, f, obj.eventList, 1.0, obj.filt, obj.decimate, obj.trim,
                  getattr(obj, 'filename', getattr(obj, 'fileName', 'clust.pkl')), obj.eventsOnAllStations,
                  obj.enforceOrigin)
        for c_new, c_old in zip(new.clusters, obj.clusters):
            c_new.updateReqCC(float(c_old.ccReq))
        return new

    def updateReqCC(self, reqCC):
        """subspace.py:108-147: a float for every station, or a {station: float} dict."""
        if isinstance(reqCC, (float, int)):
            if reqCC < 0 or reqCC > 1:
                _error('reqCC must be between 0 and 1')
            for cl in self.clusters:
                cl.updateReqCC(reqCC)
        elif isinstance(reqCC, dict):
            for key, val in reqCC.items():
                self[key].updateReqCC(val)
        else:
            _error('reqCC must be a number or a dict')

    def printAtr(self):
        for cl in self.clusters:
            cl.printAtr()

    def write(self):
        with open(self.filename, 'wb') as f:
            pickle.dump(self, f)

    def __getitem__(self, key):
        if isinstance(key, (int, np.integer)):
            return self.clusters[key]
        if isinstance(key, str):
            if len(key.split('.')) == 1:
                return self.clusters[self.stalist2.index(key)]
            return self.clusters[self.stalist.index(key)]
        _error('indexer must either be an int or str of sta.net or sta you passed %s' % key)

    def __len__(self):
        return len(self.clusters)

    def __repr__(self):
        return 'SSClusterStream with %d stations ' % len(self.stalist)


class Cluster(object):
    """`detex.subspace.Cluster` (subspace.py:290-413): dendrogram cut of one station."""

    def __init__(self, clustStream, station, temkey, eventList, link, ccReq, filt, decimate, trim, DFcc):
        self.link, self.DFcc, self.station, self.temkey = link, DFcc, station, temkey
        self.key = list(eventList)
        self.trim, self.decimate = trim, decimate
        self.updateReqCC(ccReq)

    def updateReqCC(self, newccReq):
        """subspace.py:305-346.  The reference builds every intermediate cluster of the linkage as a
        Python list and keeps the maximal ones below the cut; those are the flat clusters of
        `fcluster(link, 1 - ccReq, 'distance')` with at least two members, ordered by the height of
        their top merge (highest first), members in ascending event order (`list(set(...))` of small
        ints, subspace.py:337)."""
        if newccReq < 0. or newccReq > 1.:
            _error('Parameter ccReq must be between 0 and 1')
        self.ccReq = newccReq
        link = np.asarray(self.link, dtype=np.float64)
        N = len(link)
        members = {i: [i] for i in range(N + 1)}
        top = {}                                   # cluster id -> (height, member list) of maximal merges
        parent_of = {}
        for a in range(N):
            i1, i2 = int(link[a, 0]), int(link[a, 1])
            members[N + 1 + a] = members[i1] + members[i2]
            if link[a, 2] <= 1 - self.ccReq:
                top[N + 1 + a] = (link[a, 2], members[N + 1 + a])
                top.pop(i1, None)
                top.pop(i2, None)
                parent_of[i1] = parent_of[i2] = N + 1 + a
        # highest link first; ties keep linkage order (stable sort of descending disSim, subspace.py:322)
        order = sorted(top.keys(), key=lambda k: (-top[k][0], k))
        # subspace.py:337: clustEvents = list(set(...)) -> ascending event indices
        self.clusts = [[self.key[y] for y in sorted(set(top[k][1]))] for k in order]
        clustset = set(y for x in self.clusts for y in x)
        self.singles = [k for k in self.key if k not in clustset]
        self.clustcount = int(np.sum([len(x) for x in self.clusts])) if self.clusts else 0
        self.flat = fcluster(link[:, 0:4], 1 - self.ccReq, criterion='distance') if N else np.array([1])

    def __getitem__(self, index):
        return self.clusts[index]

    def __iter__(self):                  # subspace.py:703-704
        return iter(self.clusts)

    def printAtr(self):                  # subspace.py:693-698
        print('%s Cluster' % self.station)
        print('%d Events cluster out of %d' % (self.clustcount, len(self.singles) + self.clustcount))
        print('Total number of clusters = %d' % len(self.clusts))
        print('Required Cross Correlation Coeficient = %.3f' % self.ccReq)

    def __len__(self):
        return len(self.clusts)

    def __repr__(self):
        return 'Cluster object of station %s with %d event clusters and %d singletons' % (
            self.station, len(self.clusts), len(self.singles))


# =====================================================================================
# createSubSpace
# =====================================================================================
def _sub_frames(DFcc, DFlag, all_events, events):
    """`_getInfoFromClust` (construct.py:304-320): the cluster's own (n-1)x(n-1) CC / lag frames."""
    odi = [all_events.index(e) for e in events]
    cc = np.asarray(DFcc, dtype=np.float64)
    lg = np.asarray(DFlag, dtype=np.float64)
    n = len(odi)
    scc = np.full((n - 1, n - 1), np.nan)
    slag = np.full((n - 1, n - 1), np.nan)
    for a in range(n - 1):
        for b in range(a + 1, n):
            scc[a, b - 1] = cc[odi[a], odi[b] - 1]      # rows 0..N-2, columns 1..N-1 (construct.py:373-376)
            slag[a, b - 1] = lg[odi[a], odi[b] - 1]
    return scc, slag


def _ensureUnique(DFcc, seed=0):
    """`_ensureUnique` (construct.py:814-835): the dendrogram walk looks lags up by coefficient, so equal
    coefficients (duplicate catalogue entries, say) are nudged apart by < 1e-5 -- the reference lowers
    the dissimilarity by `abs(.00001 * np.random.rand())`, unseeded; here the draw is seeded."""
    cc = np.array(DFcc, dtype=np.float64)
    iu = np.triu_indices(cc.shape[0])
    rng = np.random.default_rng(seed)
    for _ in range(11):
        v = cc[iu]
        _, first = np.unique(v, return_index=True)
        dup = np.setdiff1d(np.arange(len(v)), first)
        if len(dup) == 0:
            return cc
        log.warning('Duplicates found in correlation coefficients, perturbing slightly to get unique values')
        v[dup] += np.abs(.00001 * rng.random(len(dup)))       # cx - d  <=>  cc + d
        cc[iu] = v
    _error('cannot make Coeficients unique, killing program')


def createSubSpace(Pf=10 ** -12, clust='clust.pkl', minEvents=2, dtype='double', conDatFetcher=None, engine=None):
    """`detex.createSubSpace` (construct.py:177-301): one row per (station, cluster) with the
    aligned waveforms, offsets and statistics the detector needs; singles per station."""
    if isinstance(clust, str):
        with open(clust, 'rb') as f:
            cl = pickle.load(f)
        if not isinstance(cl, ClusterStream):
            cl = ClusterStream.from_reference(cl)    # a clust.pkl the reference wrote (Detex importable)
    elif isinstance(clust, ClusterStream):
        cl = clust
    elif all(hasattr(clust, a) for a in ('trdf', 'clusters', 'temkey', 'stakey', 'eventList')):
        cl = ClusterStream.from_reference(clust)     # the reference's own ClusterStream object
    else:
        _error('Invalid clust type, must be a path or ClusterStream instance.', ValueError)
    eng = engine or default_engine()
    temkey, stakey = cl.temkey, cl.stakey
    cfetcher = conDatFetcher if conDatFetcher is not None else cl.fetcher
    if not isinstance(cfetcher, (ArrayFetcher, StreamFetcher)) and hasattr(cfetcher, 'getConData'):
        cfetcher = StreamFetcher(cfetcher)        # the reference's DataFetcher (ObsPy Streams)
    TRDF = getattr(cl, '_TRDF', None)
    if TRDF is None:
        lf = cl.fetcher
        if not isinstance(lf, (ArrayFetcher, StreamFetcher)) and hasattr(lf, 'getTemData'):
            lf = StreamFetcher(lf)
        TRDF = _loadEvents(lf, cl.filt, cl.trim, stakey, temkey, cl.decimate, dtype, eng)
    origin = {r.NAME: _timestamp(r.TIME) for _, r in temkey.iterrows()}
    mags = {r.NAME: r.MAG for _, r in temkey.iterrows()}
    ssDict, singDic = {}, {}
    for _, row in TRDF.iterrows():
        c = cl[row.Station]
        cll = cl.trdf[cl.trdf.Station == row.Station].iloc[0]
        recs = []
        for num, evelist in enumerate(c.clusts):
            evelist = sorted(evelist)
            if len(evelist) < minEvents:                 # keep numbering, construct.py:596-598
                continue
            DFcc, DFlag = _sub_frames(cll.CCs, cll.Lags, list(cll.Events), evelist)
            DFcc = _ensureUnique(DFcc)                                          # construct.py:271
            link, delays = construct.get_delays(DFcc, DFlag)                    # construct.py:272-281
            aligned, sample_delays = construct.alignTD(delays, [row.MPtd[e] for e in evelist])
            stats = construct.update_start_times([row.Stats[e] for e in evelist], sample_delays,
                                                 [origin[e] for e in evelist], [mags[e] for e in evelist])
            offsets = [s['offset'] for s in stats]
            recs.append({'Name': 'SS%d' % num, 'Station': row.Station, 'Events': evelist,
                         'numEvents': len(evelist),
                         'AlignedTD': dict(zip(evelist, aligned)), 'SVD': None, 'UsedSVDKeys': None,
                         'FracEnergy': None, 'SVDdefined': False, 'SampleTrims': {}, 'Threshold': np.nan,
                         'SigDimRep': None, 'FAS': None, 'NumBasis': 0,
                         'Offsets': [np.min(offsets), np.median(offsets), np.max(offsets)],
                         'Stats': dict(zip(evelist, stats)),
                         'Channels': {e: row.Channels[e] for e in evelist}, '_index': num})
        if recs:
            df = pd.DataFrame(recs).astype(object)
            df.index = [r['_index'] for r in recs]
            ssDict[row.Station] = df.drop(columns=['_index'])
        else:
            log.warning('No events grouped into subspaces on %s', row.Station)
        srecs = []
        for sn, ev in enumerate(c.singles):                                      # construct.py:525-559
            if ev not in row.MPtd:
                continue
            st = dict(row.Stats[ev])
            st.update(origintime=origin[ev], offset=st['starttime'] - origin[ev], magnitude=mags[ev])
            srecs.append({'Name': 'SG%d' % sn, 'Station': row.Station, 'Events': [ev],
                          'MPtd': {ev: row.MPtd[ev]}, 'Stats': {ev: st}, 'Channels': {ev: row.Channels[ev]},
                          'Offsets': [st['offset']] * 3, 'SampleTrims': {}, 'FAS': None, 'Threshold': np.nan})
        if srecs:
            singDic[row.Station] = pd.DataFrame(srecs).astype(object)
    return SubSpace(singDic, ssDict, cl, dtype, Pf, cfetcher, engine=eng)


# =====================================================================================
# SubSpace
# =====================================================================================
class SubSpace(object):
    """`detex.subspace.SubSpace` (subspace.py:715-2037) without plotting and the GUI picker."""

    def __init__(self, singlesDict, subSpaceDict, cl, dtype, Pf, cfetcher, engine=None):
        self.cfetcher, self.clusters = cfetcher, cl
        self.subspaces, self.singles = subSpaceDict, singlesDict
        self.singletons = singlesDict
        self.dtype, self.Pf = dtype, Pf
        self.ssStations = list(self.subspaces.keys())
        self.singStations = list(self.singles.keys())
        self.Stations = sorted(set(self.ssStations) | set(self.singStations))
        self._engine = engine
        self.histSubSpaces = self.histSingles = None

    @property
    def engine(self):
        if self._engine is None:
            self._engine = default_engine()
        return self._engine

    def __getstate__(self):              # the CUDA context does not pickle: a loaded SubSpace opens its own
        d = dict(self.__dict__)
        d['_engine'] = None
        return d

    # ------------------------------------------------------------ containers / MISC (subspace.py:1998-2037)
    def __getitem__(self, key):
        if isinstance(key, (int, np.integer)):
            return self.subspaces[self.ssStations[key]]
        if isinstance(key, str):
            parts = key.split('.')
            if len(parts) == 2 and key in self.subspaces:
                return self.subspaces[key]
            if len(parts) == 1:
                hits = [k for k in self.ssStations if k.split('.')[1] == key]
                if hits:
                    return self.subspaces[hits[0]]
            _error('%s is not a station in this cluster object' % key)
        _error('%s must either be a int or str of station name' % key)

    def __len__(self):
        return len(self.subspaces)

    def write(self, filename='subspace.pkl'):
        with open(filename, 'wb') as f:
            pickle.dump(self, f)

    def printOffsets(self):
        for station in self.ssStations:
            for _, row in self.subspaces[station].iterrows():
                print('%s, %s, min=%3f, max=%3f, range=%3f' % (row.Station, row.Name, row.Offsets[0], row.Offsets[2],
                                                            row.Offsets[2] - row.Offsets[0]))

    # ------------------------------------------------------------ validateClusters
    def validateClusters(self):
        """subspace.py:738-773; zero-lag coefficients from `corr0_kernel`."""
        for sta, subs in self.subspaces.items():
            ccreq = self.clusters[sta].ccReq
            for clustNum, row in subs.iterrows():
                if 'Starttime' in row.SampleTrims and 'Endtime' in row.SampleTrims:
                    start, stop = row.SampleTrims['Starttime'], row.SampleTrims['Endtime']
                else:
                    start, stop = 0, -1
                if len(row.Events) < 2:
                    continue
                arr = np.array([row.AlignedTD[e][start:stop] for e in row.Events])
                for i in _subspace.validate_cluster(arr, ccreq, engine=self.engine):
                    ev = row.Events[i]
                    log.info('%s fails validation check or is ill-aligned on station %s, removing', ev, sta)
                    row.AlignedTD.pop(ev, None)
                bad = [e for e in row.Events if e not in row.AlignedTD]
                for e in bad:
                    row.Events.remove(e)

    # ------------------------------------------------------------ pick times
    @staticmethod
    def _getOffsets(offsets, m=25.):
        """subspace.py:1623-1637."""
        offsets = np.asarray(offsets, dtype=float)
        if len(offsets) == 1:
            return [offsets[0], offsets[0], offsets[0]]
        d = np.abs(offsets - np.median(offsets))
        mdev = np.median(d)
        offs = offsets[d / mdev < m] if mdev else offsets
        return [np.min(offs), np.median(offs), np.max(offs)]

    def _updateOffsets(self):
        """subspace.py:1443-1459."""
        for frames in (self.subspaces, self.singles):
            for sta in frames:
                for num, row in frames[sta].iterrows():
                    frames[sta].at[num, 'Offsets'] = self._getOffsets([row.Stats[x]['offset'] for x in row.Stats])

    def attachPickTimes(self, pksFile='PhasePicks.csv', function='median', defaultDuration=30):
        """subspace.py:1461-1545: phase picks (TimeStamp, Station, Event, Phase) -> SampleTrims."""
        pks = pd.read_csv(pksFile) if isinstance(pksFile, str) else pksFile
        for col in ('TimeStamp', 'Station', 'Event'):
            if col not in pks.columns:
                _error('%s is a required column of the pick file' % col)
        if 'Phase' in pks.columns:
            pks = readKey(pks, 'phases')          # subspace.py:1484: empty rows dropped, sorted
        fun = {'mean': np.mean, 'max': np.max, 'min': np.min, 'median': np.median}.get(function)
        if fun is None:
            _error('function %s not supported, options are: mean, median, min, max' % function)
        for cl in self.clusters.clusters:
            sta = cl.station
            for frames in (self.singles, self.subspaces):
                if sta not in frames:
                    continue
                for ind, row in frames[sta].iterrows():
                    if len(row.SampleTrims) > 0:
                        continue
                    pk = pks[pks.Event.isin(row.Events) & (pks.Station == sta)]
                    if len(pk) > 0:
                        trims = self._getSampTrim(row, pk, defaultDuration, fun, sta)
                        if isinstance(trims, dict):
                            frames[sta].at[ind, 'SampleTrims'] = trims
                self._updateOffsets()

    @staticmethod
    def _getSampTrim(row, pk, defaultDuration, fun, sta):
        """subspace.py:1547-1603."""
        eves = row.Events
        Nc = row.Stats[eves[0]]['Nc']
        Sr = np.round(row.Stats[eves[0]]['sampling_rate'])
        startsamps, stopsamps, secduration = [], [], []
        for ev in eves:
            p = pk[pk.Event == ev]
            if len(p) < 1:
                continue
            st0 = row.Stats[ev]['starttime']
            start = p.TimeStamp.min()
            startsampsEve = (start - st0) * (Nc * Sr)
            wf = row.MPtd[ev] if 'MPtd' in row.index else row.AlignedTD[ev]
            if len(wf) < startsampsEve:
                log.warning('Start samples for %s on %s exceeds avaliable data, skipping attaching pick', ev, sta)
                return None
            if startsampsEve < 0:
                startsampsEve, start = 0, st0
            if defaultDuration:
                stop = start + defaultDuration
                secduration.append(defaultDuration)
            else:
                stop = p.TimeStamp.max()
                secduration.append(stop - start)
            assert stop > start and stop > st0
            startsamps.append(startsampsEve)
            stopsamps.append((stop - st0) * (Nc * Sr))
            row.Stats[ev]['Starttime'] = start
            row.Stats[ev]['offset'] = start - row.Stats[ev]['origintime']
        if not startsamps:
            return None
        sSamps, eSamps = int(fun(startsamps)), int(fun(stopsamps))
        return {'Starttime': sSamps - sSamps % Nc, 'Endtime': eSamps - eSamps % Nc,
                'DurationSeconds': int(fun(secduration))}

    # ------------------------------------------------------------ SVD
    def _trimmed(self, row, keys):
        """`_trimGroups` without the demeaning (subspace.py:921-943)."""
        if 'Starttime' in row.SampleTrims and 'Endtime' in row.SampleTrims:
            stim, etim = max(row.SampleTrims['Starttime'], 0), row.SampleTrims['Endtime']
            return np.vstack([row.AlignedTD[x][stim:etim] for x in keys])
        log.warning('No trim times for %s and station %s, try running attachPickTimes', row.Name, row.Station)
        return np.vstack([row.AlignedTD[x] for x in keys])

    def SVD(self, selectCriteria=2, selectValue=0.9, conDatNum=100, threshold=None, normalize=False,
            useSingles=True, validateWaveforms=True, backupThreshold=None, **kwargs):
        """subspace.py:786-912 (validateWaveforms is accepted and unused, as in the reference)."""
        if selectCriteria in (1, 2, 3):
            if selectValue > 1 or selectValue < 0:
                _error('When selectCriteria==%d selectValue must be a float between 0 and 1' % selectCriteria,
                       ValueError)
        elif selectCriteria == 4:
            if selectValue < 0 or not isinstance(selectValue, int):
                _error('When selectCriteria==4 selectValue must be an integer greater than 0', ValueError)
        else:
            _error('selectCriteria of %s is not supported' % selectCriteria)
        if threshold is not None and (not isinstance(threshold, (int, float)) or threshold < 0):
            _error('Unsupported type for threshold, must be None or float', ValueError)
        for station in self.ssStations:
            for ind, row in self.subspaces[station].iterrows():
                keys = sorted(row.Events)
                W = self._trimmed(row, keys)
                if W.shape[1] == 0:
                    log.warning('subspace %d on %s is failing alignment and trimming, deleting it', ind, station)
                    sp = self.subspaces[station]
                    self.subspaces[station] = sp[sp.index != ind]
                    continue
                r = _subspace.svd_basis(W, selectCriteria, selectValue, normalize)
                Ufull = r['Ufull']
                svdDict = {s: Ufull[:, i] for i, s in enumerate(r['s'])}
                used = sorted(svdDict.keys(), reverse=True)[:r['NumBasis']]
                frac = dict(zip(keys, r['cum']))
                frac.update(r['FracEnergy'])
                sp = self.subspaces[station]
                sp.at[ind, 'SVD'] = svdDict
                sp.at[ind, 'FracEnergy'] = frac
                sp.at[ind, 'UsedSVDKeys'] = used
                sp.at[ind, 'SVDdefined'] = True
                sp.at[ind, 'NumBasis'] = len(used)
        if len(self.ssStations) > 0:
            self._setThresholds(selectCriteria, selectValue, conDatNum, threshold, backupThreshold, kwargs)
        if len(self.singStations) > 0 and useSingles:
            self.setSinglesThresholds(conDatNum=conDatNum, threshold=threshold, backupThreshold=backupThreshold,
                                      **kwargs)

    def _setThresholds(self, selectCriteria, selectValue, conDatNum, threshold, backupThreshold, kwargs):
        """subspace.py:1015-1054."""
        if threshold is not None and threshold > 0:
            for station in self.ssStations:
                for ind in self.subspaces[station].index:
                    self.subspaces[station].at[ind, 'Threshold'] = threshold
        elif selectCriteria == 1:
            _error('selectCriteria 1 currently not supported', ValueError)
        elif selectCriteria in (2, 4):
            self.getFAS(conDatNum, **kwargs)
            for station in self.ssStations:
                for ind, row in self.subspaces[station].iterrows():
                    th = _subspace.threshold_from_fas(row.FAS, self.Pf, backupThreshold)
                    self.subspaces[station].at[ind, 'Threshold'] = th
        elif selectCriteria == 3:
            for station in self.ssStations:
                for ind, row in self.subspaces[station].iterrows():
                    th = row.FracEnergy['Minimum'][row.NumBasis] * selectValue
                    self.subspaces[station].at[ind, 'Threshold'] = th

    def setSinglesThresholds(self, conDatNum=50, recalc=False, threshold=None, backupThreshold=None, **kwargs):
        """subspace.py:1056-1108: singles without pick times are dropped."""
        for sta in self.singStations:
            sing = self.singles[sta]
            sing = sing[[len(x) > 0 for x in sing.SampleTrims]].reset_index(drop=True)
            sing['Name'] = ['SG%d' % x for x in range(len(sing))]
            self.singles[sta] = sing
        if threshold is None:
            self.getFAS(conDatNum, useSingles=True, useSubSpaces=False, recalc=recalc, **kwargs)
        for sta in self.singStations:
            for ind, row in self.singles[sta].iterrows():
                if threshold:
                    th = threshold
                else:
                    th = _subspace.threshold_from_fas(row.FAS[0], self.Pf, backupThreshold)
                self.singles[sta].at[ind, 'Threshold'] = th

    # ------------------------------------------------------------ bases
    def _bases(self, sta, issubspace):
        """What `_loadMPSubSpace` collects (detect.py:319-388, fas.py:137-172), without the FFT copies
        and without the n x n `UtU`.  Returns (names, ssTD, offsets, mags, ewf, Nc, sr)."""
        DF = self.subspaces[sta] if issubspace else self.singles[sta]
        names, ssTD, offsets, mags, ewf = [], {}, {}, {}, {}
        Nc = sr = None
        for _, row in DF.iterrows():
            events = row.Events
            if issubspace:
                if not isinstance(row.UsedSVDKeys, list):
                    _error('SVD not defined, run SVD on subspace stream class before calling false alarm '
                           'statistic class')
                if len(row.UsedSVDKeys) == 0:
                    continue
                U = np.array([row.SVD[x] for x in row.UsedSVDKeys])
                W = self._trimmed(row, events)
            else:
                if not row.SampleTrims:
                    continue
                mptd = list(row.MPtd.values())[0]
                upr = mptd[row.SampleTrims['Starttime']:row.SampleTrims['Endtime']]
                U = _subspace.single_basis(upr)
                W = np.array([upr])
            names.append(row.Name)
            ssTD[row.Name], offsets[row.Name], ewf[row.Name] = U, row.Offsets, W
            mags[row.Name] = np.array([row.Stats[x]['magnitude'] for x in events], dtype=np.float64)
            st0 = row.Stats[events[0]]
            Nc, sr = st0['Nc'], st0['sampling_rate']
        return names, ssTD, offsets, mags, ewf, Nc, sr

    def _stakey(self, sta):
        return self.clusters.stakey[self.clusters.stakey.STATION == sta.split('.')[1]]

    # ------------------------------------------------------------ getFAS
    def getFAS(self, conDatNum, LTATime=5, STATime=0.5, staltalimit=8.0, useSubSpaces=True, useSingles=False,
               numBins=401, recalc=False, utcstart=None, utcend=None, **kwargs):
        """subspace.py:1652-1743 + fas._initFAS / _getDSVect (fas.py:23-117).  One draw of
        4 x conDatNum random chunks per station serves all of its subspaces (the reference draws
        again for every subspace, unseeded; fas.py:35, 93-94)."""
        jobs = []
        if useSubSpaces:
            self._updateOffsets()
            for sta in self.subspaces:
                done = [isinstance(f, dict) for f in self.subspaces[sta]['FAS']]
                if len(done) and all(done) and not recalc:
                    log.info('FAS for station %s already calculated, to recalculate pass True to recalc', sta)
                    continue
                jobs.append((sta, True))
        if useSingles:
            for sta in self.singles:
                # the reference tests `isinstance(fas1, dict)` on what is a 1-element LIST for singles
                # (subspace.py:1722-1727), so it always recalculates; the evident intent is kept here
                done = [isinstance(f, list) for f in self.singles[sta]['FAS']]
                if len(done) and all(done) and not recalc:
                    continue
                jobs.append((sta, False))
        for sta, issub in jobs:
            names, ssTD, _, _, _, Nc, sr = self._bases(sta, issub)
            if not names:
                continue
            stakey = self._stakey(sta)
            u1 = stakey.iloc[0].STARTTIME if utcstart is None else utcstart
            u2 = stakey.iloc[0].ENDTIME if utcend is None else utcend
            raw = [tr for tr, _ in self.cfetcher.getConData(stakey, utcstart=u1, utcend=u2,
                                                            randSamps=conDatNum * 4)]
            if len(raw) == 0:
                _error('Could not get any data for %s' % sta)
            chunks = []
            for i in range(0, len(raw), 32):
                chunks.extend(_filter_multiplex(raw[i:i + 32], self.cfetcher.sr, self.clusters.filt,
                                                self.clusters.decimate, self.engine))
            passes = _fas.screen_chunks(chunks, Nc, sr, STATime=STATime, LTATime=LTATime,
                                        staltalimit=staltalimit, engine=self.engine)
            kept = _fas.select_null_chunks(passes, conDatNum)
            if len(kept) != conDatNum:
                log.warning('%d samps not avaliable, using all avaliable', conDatNum)
            groups = {}
            for nm in names:
                groups.setdefault(ssTD[nm].shape[1], []).append(nm)
            res = {}
            for n, nms in groups.items():
                out = _fas.initFAS([ssTD[nm] for nm in nms], [chunks[i] for i in kept], Nc, numBins=numBins,
                                   engine=self.engine)
                res.update(dict(zip(nms, out)))
            DF = self.subspaces[sta] if issub else self.singles[sta]
            for ind, row in DF.iterrows():
                if row.Name in res:
                    DF.at[ind, 'FAS'] = res[row.Name] if issub else [res[row.Name]]

    # ------------------------------------------------------------ detex
    def detex(self, utcStart=None, utcEnd=None, subspaceDB='SubSpace.db', trigCon=0, triggerLTATime=5,
              triggerSTATime=0, multiprocess=False, delOldCorrs=True, calcHist=True, useSubSpaces=True,
              useSingles=False, estimateMags=True, classifyEvents=None, eventCorFile='EventCors',
              utcSaves=None, fillZeros=False, batch=16):
        """subspace.py:1745-1902 + `_SSDetex` (detect.py:27-218): run every station's subspaces /
        singles over the continuous data and write ss_df / sg_df, *_info, *_hist, filt_params."""
        if multiprocess or trigCon != 0:
            _error('multiprocessing and trigcon other than 0 not supported')
        if classifyEvents is not None or utcSaves is not None:
            # detect.py:52-56, 63-72, 88-110: event classification never fills its `eventCorList` in the reference
            # and utcSaves pickles raw debugging windows; neither is on the path -- refuse instead of ignoring
            raise NotImplementedError('classifyEvents / utcSaves are not supported')
        if os.path.exists(subspaceDB) and delOldCorrs:
            os.remove(subspaceDB)
        out = {}
        for issub, use in ((True, useSubSpaces), (False, useSingles)):
            if not use:
                continue
            frames = self.subspaces if issub else self.singles
            if issub and not all(all(frames[sta].SVDdefined) for sta in frames):
                _error('call SVD before running subspace detectors')
            if not issub:
                self.setSinglesThresholds()
            hist = {'Bins': HIST_BINS}
            for si, sta in enumerate(sorted(frames.keys())):
                names, ssTD, offsets, mags, ewf, Nc, sr = self._bases(sta, issub)
                if not names:
                    continue
                DF = frames[sta]
                thr = {r.Name: float(r.Threshold) for _, r in DF.iterrows() if r.Name in ssTD}
                det = SSDetex(ssTD, thr, offsets, Nc, sta=sta, engine=self.engine,   # 'NET.STA' (detect.py:82-86)
                              set_id=(40 if issub else 60) + si, triggerLTATime=triggerLTATime,
                              triggerSTATime=triggerSTATime, fillZeros=fillZeros, calcHist=calcHist,
                              ewf=ewf if estimateMags else None, mags=mags if estimateMags else None,
                              issubspace=issub)
                stakey = self._stakey(sta)
                u1 = stakey.iloc[0].STARTTIME if utcStart is None else utcStart
                u2 = stakey.iloc[0].ENDTIME if utcEnd is None else utcEnd
                raw, starts, found = [], [], 0
                def flush():
                    if not raw:
                        return 0
                    Sar, _, _ = det.run_raw_chunks(raw, self.cfetcher.sr, starts, filt=self.clusters.filt,
                                                   decimate=self.clusters.decimate)
                    if len(Sar):
                        results.saveSQLite(Sar, subspaceDB, 'ss_df' if issub else 'sg_df')
                    del raw[:], starts[:]
                    return len(Sar)
                for traces, start in self.cfetcher.getConData(stakey, utcstart=u1, utcend=u2):
                    raw.append(traces)
                    starts.append(start)
                    if len(raw) >= batch:
                        found += flush()
                found += flush()
                hist[sta] = det.histdic
                out[(sta, issub)] = found
            if issub:
                self.histSubSpaces = hist
            else:
                self.histSingles = hist
        if useSubSpaces or useSingles:
            ssinfo, sginfo = self._getInfoDF()

            def hframe(h):   # `_getHistograms` (subspace.py:1956-1995)
                return results.hist_frame({k: v for k, v in h.items() if k != 'Bins'}, bins=h['Bins'])
            results.write_run(subspaceDB, issubspace=True, info=ssinfo if useSubSpaces else None,
                              hist=hframe(self.histSubSpaces) if useSubSpaces else None, filt=self.clusters.filt)
            if useSingles:
                results.write_run(subspaceDB, issubspace=False, info=sginfo, hist=hframe(self.histSingles),
                                  filt=None)
        return out

    def _getInfoDF(self):
        """subspace.py:1904-1954."""
        ss, sg = [], []
        for sta in self.Stations:
            if sta in self.subspaces:
                for _, r in self.subspaces[sta].iterrows():
                    b = r.FAS['betadist'] if isinstance(r.FAS, dict) and len(r.FAS) > 1 else (np.nan, np.nan)
                    ss.append([r.Name, r.Station, ','.join(r.Events), r.Threshold, r.NumBasis, b[0], b[1]])
            if sta in self.singles:
                for _, r in self.singles[sta].iterrows():
                    ok = isinstance(r.FAS, list) and len(r.FAS[0]) > 1
                    b = r.FAS[0]['betadist'] if ok else (np.nan, np.nan)
                    sg.append([r.Name, r.Station, ','.join(r.Events), r.Threshold, b[0], b[1]])
        ssinfo = pd.DataFrame(ss, columns=['Name', 'Sta', 'Events', 'Threshold', 'NumBasisUsed', 'beta1',
                                           'beta2']) if ss else None
        sginfo = pd.DataFrame(sg, columns=['Name', 'Sta', 'Events', 'Threshold', 'beta1', 'beta2']) if sg else None
        return ssinfo, sginfo
