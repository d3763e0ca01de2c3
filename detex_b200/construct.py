"""construct.py -- host-side mirror of the reference's clustering hot path
(`detex/construct.py`) on top of the CUDA engine.

  _CCX2(mpfd1, mpfd2, mptd1, mptd2, Nc1, Nc2)        construct.py:425-466
  _makeDFcclags(eventList, row)                      construct.py:369-394
  cluster_link(DFcc)                                 construct.py:152-157 (linkage stays on CPU:
                                                     SciPy single linkage is an O(N^2) MST)
"""
import numpy as np
import pandas as pd
from scipy.cluster.hierarchy import linkage

from .detect import default_engine


def _CCX2(mpfd1, mpfd2, mptd1, mptd2, Nc1, Nc2, engine=None):
    """Drop-in for `_CCX2`: (maxcc, sampleLag, subsamp) of one event pair.  The frequency
    domain operands are accepted and ignored.  Raises like the reference
    (`detex.log(level='error')`) on unequal channel counts / lengths (construct.py:430-436)."""
    if len(Nc1) != len(Nc2):
        raise Exception('Number of Channels not equal, cannot perform correlation')
    if len(mptd1) != len(mptd2):
        raise Exception('Lengths not equal on multiplexed data, cannot correlate')
    eng = engine or default_engine()
    X = np.vstack([np.asarray(mptd1, dtype=np.float64), np.asarray(mptd2, dtype=np.float64)])
    cc, lag, sub = eng.ccx(X, len(Nc1))
    return float(cc[0, 1]), int(lag[0, 1]), float(sub[0, 1])


def ccx_matrix(X, Nc, engine=None, row_begin=0, row_end=None):
    """Upper-triangular CCX block for waveforms X[N][n]: dense (rows, N) arrays."""
    eng = engine or default_engine()
    return eng.ccx(X, Nc, row_begin=row_begin, row_end=row_end)


def _makeDFcclags(eventList, row, engine=None):
    """Drop-in for `_makeDFcclags`: returns (DFcc, DFlag, DFsubsamp) with index 0..N-2,
    columns 1..N-1 and NaN below the diagonal, as the reference builds them."""
    N = len(eventList)
    chans = [row.loc['Channels'][e] for e in eventList]
    if any(len(c) != len(chans[0]) for c in chans):
        raise Exception('Number of Channels not equal, cannot perform correlation')
    lens = set(len(row.loc['MPtd'][e]) for e in eventList)
    if len(lens) != 1:
        raise Exception('Lengths not equal on multiplexed data, cannot correlate')
    X = np.array([np.asarray(row.loc['MPtd'][e], dtype=np.float64) for e in eventList])
    cc, lag, sub = ccx_matrix(X, len(chans[0]), engine=engine, row_end=N - 1)
    cols = np.arange(1, N)
    idx = np.arange(0, N - 1)
    mask = cols[None, :] > idx[:, None]

    def frame(a):
        v = np.where(mask, a[:, 1:].astype(np.float64), np.nan)
        return pd.DataFrame(v, index=idx, columns=cols).astype(object)

    return frame(cc), frame(lag), frame(sub)


def _flatNoNan(df):
    v = np.asarray(df, dtype=np.float64).flatten()
    return v[~np.isnan(v)]


def cluster_link(DFcc):
    """createCluster tail (construct.py:152-157)."""
    cx = _flatNoNan(1.0000001 - np.asarray(DFcc, dtype=np.float64))
    return linkage(cx)
