"""results.py -- N4: the detector's tables on disk and the association step
that consumes them, host side (pandas / sqlite3; no GPU work -- these are sparse trigger lists).

  saveSQLite / loadSQLite        detex/util.py:870-931, detex/pandas_dbms.py:64-130, 214-251
                                 (same CREATE TABLE text and column affinities, so a database
                                 written here opens in the reference's `detResults` and vice versa)
  info_frame / hist_frame /      detex/subspace.py:1883-1995 (`ss_info`, `sg_info`, `ss_hist`,
  filt_frame / write_run         `sg_hist`, `filt_params`), detex/detect.py:144, 206-211 (`ss_df`, `sg_df`)
  select_detections              detex/results.py:318-368 (`_buildSQL`, DS / DS_STALTA gate)
  deleteDetDups                  detex/results.py:371-401, vectorised (one lexsort, no groupby)
  associateDetections            detex/results.py:404-515 (`associateReq = 0` branch, the reference
                                 default), vectorised group statistics

Detection-table columns (detect.py:397-398): DS, DS_STALTA, STMP, Name, Sta, MSTAMPmin, MSTAMPmax,
Mag, SNR, ProEnMag.
"""
import datetime
import json
import os
import sqlite3

import numpy as np
import pandas as pd

DET_COLS = ['DS', 'DS_STALTA', 'STMP', 'Name', 'Sta', 'MSTAMPmin', 'MSTAMPmax', 'Mag', 'SNR', 'ProEnMag']
EVENT_COLS = ['Event', 'DSav', 'DSmax', 'NumStations', 'DS_STALTA', 'MSTAMPmin', 'MSTAMPmax', 'Mag',
              'ProEnMag', 'Verified', 'Dets']


# ------------------------------------------------------------------ SQLite wire format
def _schema(frame, name):
    """`pandas_dbms.get_schema(frame, name, 'sqlite')` (pandas_dbms.py:214-251): integer, bool and
    float columns are NUMBER, datetimes TIMESTAMP, everything else VARCHAR2; same text layout."""
    cols = []
    for k in frame.columns:
        ser = frame[k]
        if pd.api.types.is_datetime64_any_dtype(ser):
            t = 'TIMESTAMP'
        elif pd.api.types.is_bool_dtype(ser) or pd.api.types.is_numeric_dtype(ser):
            t = 'NUMBER'
        else:
            t = 'VARCHAR2'
        cols.append('%s %s' % (str(k).replace(' ', '_').strip(), t))
    return """CREATE TABLE %(name)s (
                      %(columns)s
                    );""" % {'name': name, 'columns': ',\n  '.join(cols)}


def _table_exists(con, name):
    q = "SELECT name FROM sqlite_master WHERE type='table' AND name=?"
    return len(con.execute(q, (name,)).fetchall()) > 0


def saveSQLite(DF, CorDB, Tablename, silent=True):
    """`detex.util.saveSQLite` (util.py:870-893): create the table on first use, append after."""
    with sqlite3.connect(CorDB, detect_types=sqlite3.PARSE_DECLTYPES) as con:
        if not _table_exists(con, Tablename):
            con.execute(_schema(DF, Tablename))
        rows = [tuple(v.item() if isinstance(v, np.generic) else v for v in r)
                for r in DF.itertuples(index=False, name=None)]
        con.executemany('INSERT INTO %s VALUES (%s)' % (Tablename, ','.join(['?'] * len(DF.columns))), rows)


def loadSQLite(corDB, tableName, sql=None, readExcpetion=False, silent=True, convertNumeric=True):
    """`detex.util.loadSQLite` (util.py:896-931): None when the database / table is missing."""
    if sql is None:
        sql = 'SELECT %s FROM %s' % ('*', tableName)
    if not os.path.exists(corDB):
        return None
    with sqlite3.connect(corDB, detect_types=sqlite3.PARSE_DECLTYPES) as con:
        try:
            df = pd.read_sql(sql, con)
        except Exception:
            return None
    if convertNumeric:
        for col in df.columns:
            try:
                df[col] = pd.to_numeric(df[col])
            except (ValueError, TypeError):
                pass
    return df


def info_frame(rows, issubspace=True):
    """`SubSpace._getInfoDF` (subspace.py:1906-1952).  rows: iterable of dicts with Name, Station,
    Events (list), Threshold, NumBasis (subspaces only) and FAS (`{'betadist': (a, b, 0, 1), ...}`,
    or a 1-element list of it for singles)."""
    cols = (['Name', 'Sta', 'Events', 'Threshold', 'NumBasisUsed', 'beta1', 'beta2'] if issubspace
            else ['Name', 'Sta', 'Events', 'Threshold', 'beta1', 'beta2'])
    out = []
    for r in rows:
        fas = r.get('FAS')
        if isinstance(fas, list):
            fas = fas[0] if fas else None
        if isinstance(fas, dict) and len(fas.keys()) > 1:
            b1, b2 = fas['betadist'][0], fas['betadist'][1]
        else:
            b1, b2 = np.nan, np.nan
        rec = [r['Name'], r['Station'], ','.join(r['Events']), r['Threshold']]
        if issubspace:
            rec.append(r['NumBasis'])
        out.append(rec + [b1, b2])
    return pd.DataFrame(out, columns=cols) if out else None


def hist_frame(hist, bins=None):
    """`SubSpace._getHistograms` (subspace.py:1954-1995): hist = {sta: {name: int[400]}}; first row
    holds the bin edges, each value is a JSON list."""
    if bins is None:
        bins = np.linspace(0, 1, 401)
    rows = [['Bins', 'Bins', json.dumps(np.asarray(bins).tolist())]]
    for sta in hist:
        for name in hist[sta]:
            rows.append([name, sta, json.dumps(np.asarray(hist[sta][name]).tolist())])
    return pd.DataFrame(rows, columns=['Name', 'Sta', 'Value'])


def filt_frame(filt):
    """`filt_params` table (subspace.py:1883-1886)."""
    return pd.DataFrame([list(filt)], columns=['FREQMIN', 'FREQMAX', 'CORNERS', 'ZEROPHASE'], index=[0])


def write_run(subspaceDB, detections=None, issubspace=True, info=None, hist=None, filt=None):
    """Write one detection run the way `SubSpace.detex` leaves it (subspace.py:1869-1904,
    detect.py:144-211): detections -> ss_df / sg_df in blocks of 500 rows, then filt_params,
    ss_info / sg_info, ss_hist / sg_hist."""
    pre = 'ss' if issubspace else 'sg'
    if detections is not None and len(detections) > 0:
        det = detections[DET_COLS]
        for lo in range(0, len(det), 500):
            saveSQLite(det.iloc[lo:lo + 500], subspaceDB, pre + '_df')
    if filt is not None:
        saveSQLite(filt_frame(filt), subspaceDB, 'filt_params')
    if info is not None:
        saveSQLite(info, subspaceDB, pre + '_info')
    if hist is not None:
        saveSQLite(hist, subspaceDB, pre + '_hist')


def select_detections(ssDB, tableName='ss_df', trigCon=0, trigParameter=0.0, stations=None, starttime=None,
                      endtime=None):
    """The SELECTs of `_buildSQL` without a PfKey (results.py:349-367): `DS >= trigParameter`
    (trigCon 0) or `DS_STALTA >= trigParameter` (trigCon 1), MSTAMPmin inside [starttime, endtime].
    (`_deleteDetDups` passes starttime / stations to `_buildSQL` in swapped positions,
    results.py:378-379, so the reference effectively never filters on either when stations is
    None; the filters here do what the signature says.)"""
    if not starttime or not endtime:
        starttime, endtime = 0.0, 4500 * 3600 * 24 * 365.25
    cond = 'DS' if trigCon == 0 else 'DS_STALTA'
    parts = []
    for sta in (stations if isinstance(stations, (list, tuple)) else ['*']):
        where = '%s >= %s AND MSTAMPmin>=%f AND MSTAMPmin<=%f' % (cond, trigParameter, starttime, endtime)
        if sta != '*':
            where = 'Sta="%s" AND ' % sta + where
        df = loadSQLite(ssDB, tableName, sql='SELECT * FROM %s WHERE %s' % (tableName, where))
        if isinstance(df, pd.DataFrame):
            parts.append(df)
    if not parts:
        return None
    return pd.concat(parts, ignore_index=True)


# ------------------------------------------------------------------ duplicate removal
def _group_last(keys_sorted):
    """Index of the last element of every run of equal keys in a sorted key array."""
    n = len(keys_sorted)
    if n == 0:
        return np.zeros(0, dtype=np.int64)
    return np.nonzero(np.r_[keys_sorted[1:] != keys_sorted[:-1], True])[0]


def deleteDetDups(ssdf, associateBuffer):
    """Array part of `_deleteDetDups` (results.py:388-399): per station, detections whose
    [MSTAMPmin, MSTAMPmax] windows chain within `associateBuffer` seconds are one group (`Gnum`);
    keep the row with the highest DS of each group.  Returns a new frame ordered by Gnum with the
    `Gnum` column, index reset -- the reference's output."""
    if ssdf is None or len(ssdf) == 0:
        return None
    sta = ssdf['Sta'].to_numpy()
    tmin = ssdf['MSTAMPmin'].to_numpy(dtype=np.float64)
    tmax = ssdf['MSTAMPmax'].to_numpy(dtype=np.float64)
    ds = ssdf['DS'].to_numpy(dtype=np.float64)
    sta_code = np.unique(sta, return_inverse=True)[1]
    o = np.lexsort((tmin, sta_code))                       # sort_values(['Sta', 'MSTAMPmin'])
    brk = np.ones(len(o), dtype=bool)                      # row 0: Sta != NaN -> new group
    brk[1:] = ((tmin[o][1:] - associateBuffer) > tmax[o][:-1]) | (sta_code[o][1:] != sta_code[o][:-1])
    gnum = np.cumsum(brk)
    o2 = np.lexsort((ds[o], gnum))                         # sort_values(['Gnum', 'DS'])
    keep = o2[_group_last(gnum[o2])]                       # drop_duplicates('Gnum', keep='last')
    out = ssdf.iloc[o[keep]].copy()
    out['Gnum'] = gnum[keep]
    return out.reset_index(drop=True)


# ------------------------------------------------------------------ association
def _event_name(ts):
    """`str(obspy.UTCDateTime(ts)).replace(':', '-').split('.')[0]` (results.py:478-479)."""
    us = int(round(ts * 1e6))
    d = datetime.datetime(1970, 1, 1) + datetime.timedelta(microseconds=us)
    return d.strftime('%Y-%m-%dT%H-%M-%S')


def _segment_nanmedian(v, st, cnt):
    """np.nanmedian of every segment v[st[k] : st[k] + cnt[k]] (NaN for all-NaN segments,
    `_getMagnitudes`, results.py:505-514) without a Python loop: values sorted inside their segment
    with NaN last, then the middle one / two of the finite ones."""
    v = np.asarray(v, dtype=np.float64)
    seg = np.repeat(np.arange(len(st)), cnt)
    o = np.lexsort((v, seg))                       # NaN sorts last inside each segment
    vs = v[o]
    m = np.add.reduceat((~np.isnan(v)).astype(np.int64), st)
    lo = st + np.maximum(m - 1, 0) // 2
    hi = st + m // 2
    med = 0.5 * (vs[lo] + vs[np.minimum(hi, len(vs) - 1)])
    return np.where(m > 0, med, np.nan)


def associateDetections(ssdf, requiredNumStations, associateBuffer, temkey, exceptionalThreshold=None,
                        with_dets=True):
    """`_associateDetections` for `associateReq = 0` (results.py:404-466 `else` branch, 476-515).

    ssdf: output of `deleteDetDups`; temkey: frame with NAME and STMP (epoch seconds of TIME --
    the reference fills STMP with obspy.UTCDateTime(TIME).timestamp, results.py:421).  Detections
    whose windows chain within `associateBuffer` across stations form an event; an event needs
    `requiredNumStations` distinct stations, or a DS at / above `exceptionalThreshold` (float, or
    {sta: value} with the DS <= 1.01 guard of `_check_if_exceptional`); one row per station is
    kept (highest DS); an event overlapping a template origin (+- buffer) goes to the auto table
    under that template's NAME.  Returns [detTable, autoTable] with the reference's columns; group
    statistics are computed with segmented NumPy reductions instead of a pandas groupby loop
    (2e5 detections: ~4 s, of which ~3 s build the per-event `Dets` frames; `with_dets=False` leaves
    that column None)."""
    empty = pd.DataFrame(columns=EVENT_COLS)
    if ssdf is None or len(ssdf) == 0:
        return [empty, empty.copy()]
    df = ssdf.iloc[np.argsort(ssdf['MSTAMPmin'].to_numpy(dtype=np.float64), kind='stable')].reset_index(drop=True)
    tmin = df['MSTAMPmin'].to_numpy(dtype=np.float64)
    tmax = df['MSTAMPmax'].to_numpy(dtype=np.float64)
    ds = df['DS'].to_numpy(dtype=np.float64)
    sta_code = np.unique(df['Sta'].to_numpy(), return_inverse=True)[1]
    n = len(df)
    brk = np.zeros(n, dtype=bool)
    brk[1:] = (tmin[1:] - associateBuffer) > tmax[:-1]
    gid = np.cumsum(brk)                                    # 0-based group id, rows already in time order
    starts = np.nonzero(np.r_[True, brk[1:]])[0]
    ngroups = len(starts)
    # distinct stations per group
    o = np.lexsort((sta_code, gid))
    first_of_pair = np.r_[True, (gid[o][1:] != gid[o][:-1]) | (sta_code[o][1:] != sta_code[o][:-1])]
    nsta = np.bincount(gid[o][first_of_pair], minlength=ngroups)
    ok = nsta >= requiredNumStations
    if isinstance(exceptionalThreshold, float):
        ok |= np.maximum.reduceat(ds, starts) >= exceptionalThreshold
    elif isinstance(exceptionalThreshold, dict):
        lim = np.array([exceptionalThreshold.get(s, 100) for s in df['Sta'].to_numpy()], dtype=np.float64)
        exc = (ds >= lim) & (ds <= 1.01)
        ok |= np.bincount(gid[exc], minlength=ngroups) > 0
    # one row per (group, station): the highest DS, then back to time order inside the group
    o = np.lexsort((ds, sta_code, gid))
    last = np.r_[(gid[o][1:] != gid[o][:-1]) | (sta_code[o][1:] != sta_code[o][:-1]), True]
    keep = np.sort(o[last & ok[gid[o]]])
    if len(keep) == 0:
        return [empty, empty.copy()]
    g = df.iloc[keep].reset_index(drop=True)
    gid_k, tmin_k, tmax_k, ds_k = gid[keep], tmin[keep], tmax[keep], ds[keep]
    st = np.nonzero(np.r_[True, gid_k[1:] != gid_k[:-1]])[0]
    cnt = np.diff(np.r_[st, len(keep)])
    stalta = g['DS_STALTA'].to_numpy(dtype=np.float64)
    dsav = np.add.reduceat(ds_k, st) / cnt
    dsmax = np.maximum.reduceat(ds_k, st)
    stav = np.add.reduceat(stalta, st) / cnt
    gmin = np.minimum.reduceat(tmin_k, st)
    gmax = np.maximum.reduceat(tmax_k, st)
    tmean = 0.5 * (np.add.reduceat(tmin_k, st) / cnt + np.add.reduceat(tmax_k, st) / cnt)
    # auto detections: a template origin within the buffer of any row's window; the name is the
    # first such template (temkey order) of the LAST matching row (results.py:487-493)
    tk_t = np.asarray(temkey['STMP'], dtype=np.float64) if temkey is not None and len(temkey) else np.zeros(0)
    tk_name = list(temkey['NAME']) if len(tk_t) else []
    first_tem = np.full(len(keep), -1, dtype=np.int64)
    if len(tk_t):
        for lo in range(0, len(keep), 65536):
            hi = min(len(keep), lo + 65536)
            hit = ((tk_t[None, :] + associateBuffer > tmin_k[lo:hi, None]) &
                   (tk_t[None, :] - associateBuffer < tmax_k[lo:hi, None]))
            anyhit = hit.any(axis=1)
            first_tem[lo:hi] = np.where(anyhit, hit.argmax(axis=1), -1)
    mag = _segment_nanmedian(g['Mag'].to_numpy(dtype=np.float64), st, cnt)
    pem = _segment_nanmedian(g['ProEnMag'].to_numpy(dtype=np.float64), st, cnt)
    # template of the last matching row of each group (-1: none)
    pos = np.where(first_tem >= 0, np.arange(len(keep)), -1)
    last_hit = np.maximum.reduceat(pos, st)
    tem_of_group = np.where(last_hit >= 0, first_tem[np.maximum(last_hit, 0)], -1)
    det_rows, auto_rows = [], []
    for k in range(len(st)):
        sub = g.iloc[st[k]:st[k] + cnt[k]] if with_dets else None
        rec = [None, dsav[k], dsmax[k], int(cnt[k]), stav[k], gmin[k], gmax[k], mag[k], pem[k], False, sub]
        if tem_of_group[k] >= 0:
            rec[0] = tk_name[int(tem_of_group[k])]
            auto_rows.append(rec)
        else:
            rec[0] = _event_name(tmean[k])
            det_rows.append(rec)
    return [pd.DataFrame(det_rows, columns=EVENT_COLS), pd.DataFrame(auto_rows, columns=EVENT_COLS)]


# ------------------------------------------------------------------ detResults
def _approximateThreshold(beta_a, beta_b, target, numintervals=1000, numloops=3):
    """results.py:209-229: forward grid search where `beta.isf` breaks down."""
    import scipy.stats
    startVal, stopVal = 0, 1
    for _ in range(numloops):
        Xs = np.linspace(startVal, stopVal, numintervals)
        pfs = scipy.stats.beta.sf(Xs, beta_a, beta_b)
        minind = int(np.abs(pfs - target).argmin())
        bestPf, bestX = pfs[minind], Xs[minind]
        if minind == 0 or minind == numintervals - 1:
            raise ValueError('Grind search failing, set threshold manually')
        startVal, stopVal = Xs[minind - 1], Xs[minind + 1]
    return bestX, bestPf


def makePfKey(info, Pf):
    """`_makePfKey` for one info table (results.py:176-206): per (Sta, Name) the DS value whose
    false-alarm probability under the fitted beta distribution is Pf."""
    import scipy.stats
    if not Pf or not isinstance(info, pd.DataFrame):
        return None
    rows = []
    for _, row in info.iterrows():
        TH = scipy.stats.beta.isf(Pf, row.beta1, row.beta2, 0, 1)
        if TH > .94:
            TH, _ = _approximateThreshold(row.beta1, row.beta2, Pf, 1000, 3)
        rows.append([row.Sta, row.Name, TH, [row.beta1, row.beta2, 0, 1]])
    return pd.DataFrame(rows, columns=['Sta', 'Name', 'DS', 'betadist'])


def verifyEvents(Dets, Autos, veriFile, veriBuffer=1, includeAllVeriColumns=True):
    """`_verifyEvents` (results.py:232-293): mark the detection (highest DSav) whose origin window
    +- veriBuffer / 2 contains a known event; returns the frame of verified detections with the
    Ver* columns.  veriFile: DataFrame or csv path with TIME, LAT, LON, MAG, DEPTH, NAME."""
    if veriFile is None:
        return None
    ver = pd.read_csv(veriFile) if isinstance(veriFile, str) else veriFile.copy()
    req = ['TIME', 'LAT', 'LON', 'MAG', 'DEPTH', 'NAME']
    if not set(req).issubset(ver.columns):
        raise Exception('veriFile does not have the required columns, it needs TIME,LAT,LON,MAG,DEPTH,NAME')
    from .workflow import _timestamp
    ver['STMP'] = [_timestamp(x) for x in ver['TIME']]
    extra = [c for c in ver.columns if c not in ('TIME', 'LAT', 'LON', 'MAG', 'ProEnMag', 'DEPTH', 'NAME')]
    out = []
    for _, v in ver.iterrows():
        for table in (Dets, Autos):
            if table is None or len(table) == 0:
                continue
            m = ((table.MSTAMPmin - veriBuffer / 2.0 < v.STMP) & (table.MSTAMPmax + veriBuffer / 2.0 > v.STMP)
                 & ~table.Verified.astype(bool))
            tem = table[m]
            if len(tem) == 0:
                continue
            tru = tem[tem.DSav == tem.DSav.max()].iloc[:1].copy()
            table.loc[tru.index[0], 'Verified'] = True
            if includeAllVeriColumns:
                for col in extra:
                    if col not in tru.columns:
                        tru[col] = v[col]
            tru['VerMag'], tru['VerLat'], tru['VerLon'] = v.MAG, v.LAT, v.LON
            tru['VerDepth'], tru['VerName'] = v.DEPTH, v.NAME
            out.append(tru)
            break                                       # Autos are only searched when Dets had no match
    if not out:
        return pd.DataFrame()
    return pd.concat(out, ignore_index=True).drop(columns=['Verified'])


class SSResults(object):
    """`detex.results.SSResults` (results.py:588-601) without `writeDetections` (it fetches and writes
    ObsPy waveforms: out of scope)."""

    def __init__(self, Dets, Autos, Vers, ss_info, ss_filt, temkey, stakey, templateKey=None, fetcher=None):
        self.Autos, self.Dets, self.Vers = Autos, Dets, Vers
        self.NumVerified = len(Vers) if isinstance(Vers, pd.DataFrame) else 'N/A'
        self.info, self.filt = ss_info, ss_filt
        self.StationKey, self.TemplateKey, self.TemKeyPath, self.fetcher = stakey, temkey, templateKey, fetcher

    def __repr__(self):
        return ('SSResults instance with %d autodections and %d new detections, %s are verified'
                % (len(self.Autos), len(self.Dets), self.NumVerified))


def detResults(trigCon=0, trigParameter=0, associateReq=0, ss_associateBuffer=1, sg_associateBuffer=2.5,
               requiredNumStations=4, veriBuffer=1, ssDB='SubSpace.db', templateKey='TemplateKey.csv',
               stationKey='StationKey.csv', veriFile=None, includeAllVeriColumns=True, reduceDets=True, Pf=False,
               stations=None, starttime=None, endtime=None, fetch='ContinuousWaveForms', exceptionalThreshold=None):
    """`detex.results.detResults` (results.py:22-173): load the detections a run left in `ssDB`, drop
    per-station duplicates, associate across stations, split off the auto-detections of the training
    events, verify against a catalogue.  The keys are DataFrames (or csv paths)."""
    import os
    if not os.path.exists(ssDB):
        raise Exception('%s does not exist' % ssDB)
    if trigCon not in (0, 1):
        raise Exception('trigCon must be 0 or 1')
    if associateReq != 0:
        raise Exception('associateReq values other than 0 not yet supported')     # results.py:120-122
    from .util import readKey
    temkey = readKey(templateKey, 'template').copy()                              # results.py:121-122
    stakey = readKey(stationKey, 'station')
    from .workflow import _timestamp
    temkey['STMP'] = [_timestamp(t) for t in temkey['TIME']]                      # results.py:421
    ss_info, sg_info = loadSQLite(ssDB, 'ss_info'), loadSQLite(ssDB, 'sg_info')
    filt = loadSQLite(ssDB, 'filt_params')
    frames = []
    for table, buf, info in (('ss_df', ss_associateBuffer, ss_info), ('sg_df', sg_associateBuffer, sg_info)):
        if reduceDets:
            df = select_detections(ssDB, table, trigCon, trigParameter, stations, starttime, endtime)
            key = makePfKey(info, Pf)
            if df is not None and key is not None:       # `_buildSQL` with a PfKey: per-detector DS floor
                floor_of = {(r.Sta, r.Name): r.DS for _, r in key.iterrows()}   # Sta is NET.STA in both tables
                floor = np.array([floor_of.get((s, n), np.inf) for s, n in zip(df.Sta, df.Name)])
                df = df[df.DS.to_numpy() >= floor]
            df = deleteDetDups(df, buf)
        else:
            if Pf:
                raise Exception('When using the Pf parameter reduceDets must be True')
            df = loadSQLite(ssDB, table)
        if df is not None and len(df):
            frames.append(df)
    if not frames:
        raise Exception('No detections found that meet given criteria')
    df = pd.concat(frames, ignore_index=True)
    if isinstance(stations, (list, tuple)):
        df = df[df.Sta.isin(stations)]
    Dets, Autos = associateDetections(df, requiredNumStations, ss_associateBuffer, temkey, exceptionalThreshold)
    Vers = verifyEvents(Dets, Autos, veriFile, veriBuffer, includeAllVeriColumns)
    return SSResults(Dets, Autos, Vers, ss_info, filt, temkey, stakey, templateKey if isinstance(templateKey, str) else None,
                     fetch)
