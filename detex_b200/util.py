"""util.py -- the key files either side of the path: `detex.util.readKey` (util.py:564-627).

Detex is driven by three CSV tables: the template key (one row per event: TIME, NAME, LAT, LON, MAG, DEPTH), the
station key (NETWORK, STATION, STARTTIME, ENDTIME, LAT, LON, ELEVATION, CHANNELS) and the phase picks
(TimeStamp, Event, Station, Phase).  `createCluster`, `createSubSpace` and `SubSpace.attachPickTimes` take a
path or a DataFrame for each, as the reference does.
"""
import logging
import os

import pandas as pd

log = logging.getLogger("detex_b200")

# util.py:564-571
REQ_COLUMNS = {
    'template': ('TIME', 'NAME', 'LAT', 'LON', 'MAG', 'DEPTH'),
    'station': ('NETWORK', 'STATION', 'STARTTIME', 'ENDTIME', 'LAT', 'LON', 'ELEVATION', 'CHANNELS'),
    'phases': ('TimeStamp', 'Event', 'Station', 'Phase'),
}


def _error(msg):
    log.error(msg)
    raise Exception(msg)          # detex.log(level='error') raises a plain Exception


def readKey(dfkey, key_type='template'):
    """`detex.util.readKey`: read a key CSV (or take the DataFrame), check the required columns, drop the rows
    with an empty required field, sort, reset the index; station / network codes become strings (a station
    called 1234 stays '1234').  Errors raise `Exception` like `detex.log(level='error')`.

    One deliberate difference: the reference sorts by `list(req_columns[key_type])`, the iteration order of a
    Python-2 `set` of column names -- an arbitrary but fixed priority under Python 2, a per-process random one
    under Python 3.  Here the priority is the documented column order above (events by TIME, stations by
    NETWORK then STATION, picks by TimeStamp); the rows kept are the same."""
    if key_type not in REQ_COLUMNS:
        _error('unsported key type, supported types are %s' % (list(REQ_COLUMNS),))
    if isinstance(dfkey, str):
        if not os.path.exists(dfkey):
            _error('%s does not exists, check path' % dfkey)
        df = pd.read_csv(dfkey)
    elif isinstance(dfkey, pd.DataFrame):
        df = dfkey
    else:
        _error('Data type of dfkey not understood')
    req = list(REQ_COLUMNS[key_type])
    if not set(req).issubset(df.columns):
        _error('Required columns not in %s, required columns for %s key are %s' % (list(df.columns), key_type, req))
    keep = [all(x != '' for x in row) for row in df.loc[:, req].itertuples(index=False)]
    df = df[keep].sort_values(by=req).reset_index(drop=True)
    if key_type == 'station':
        df['STATION'] = [str(x) for x in df['STATION']]
        df['NETWORK'] = [str(x) for x in df['NETWORK']]
    return df


def loadClusters(filename='clust.pkl'):
    """`detex.util.loadClusters` (util.py:934-950): the pickled ClusterStream -- one written here, or one the
    reference wrote (Detex importable), which is converted (`ClusterStream.from_reference`)."""
    from . import workflow
    cl = pd.read_pickle(filename)
    if isinstance(cl, workflow.ClusterStream):
        return cl
    if all(hasattr(cl, a) for a in ('trdf', 'clusters', 'temkey', 'stakey', 'eventList')):
        return workflow.ClusterStream.from_reference(cl)
    _error('%s is not a ClusterStream instance' % filename)


def loadSubSpace(filename='subspace.pkl'):
    """`detex.util.loadSubSpace` (util.py:953-969): a SubSpace written by `SubSpace.write` (it opens its own CUDA
    context when first used)."""
    from . import workflow
    ss = pd.read_pickle(filename)
    if not isinstance(ss, workflow.SubSpace):
        _error('%s is not a SubSpaceStream instance' % filename)
    return ss
