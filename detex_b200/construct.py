"""construct.py -- host-side mirror of the reference's clustering hot path
(`detex/construct.py`) on top of the CUDA engine.

  _CCX2(mpfd1, mpfd2, mptd1, mptd2, Nc1, Nc2)        construct.py:425-466
  _makeDFcclags(eventList, row)                      construct.py:369-394
  cluster_link(DFcc)                                 construct.py:152-157 (linkage stays on CPU:
                                                     SciPy single linkage is an O(N^2) MST)
  get_delays(DFcc, DFlag) / alignTD(delays, MPtd)    construct.py:272-286, 486-503, 710-812
                                                     (dendrogram-walk alignment, O(N^2) instead of
                                                     the reference's O(N^3) pandas loops)
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


def ccx_matrix(X, Nc, engine=None, row_begin=0, row_end=None, kernel="fp64"):
    """Upper-triangular CCX block for waveforms X[N][n]: dense (rows, N) arrays.
    kernel: "tcgen05" (Hankel GEMM + float64 re-scoring of the near-maximal lags) or "fp64"."""
    eng = engine or default_engine()
    return eng.ccx(X, Nc, row_begin=row_begin, row_end=row_end, engine=kernel)


def _makeDFcclags(eventList, row, engine=None, kernel="tcgen05"):
    """Drop-in for `_makeDFcclags`: returns (DFcc, DFlag, DFsubsamp) with index 0..N-2,
    columns 1..N-1 and NaN below the diagonal, as the reference builds them.  Both kernels give
    float64-accurate coefficients and identical lags (tests/test_gpu_ccx.py); the tensor-core one
    is ~10x faster at N = 4096."""
    N = len(eventList)
    chans = [row.loc['Channels'][e] for e in eventList]
    if any(len(c) != len(chans[0]) for c in chans):
        raise Exception('Number of Channels not equal, cannot perform correlation')
    lens = set(len(row.loc['MPtd'][e]) for e in eventList)
    if len(lens) != 1:
        raise Exception('Lengths not equal on multiplexed data, cannot correlate')
    X = np.array([np.asarray(row.loc['MPtd'][e], dtype=np.float64) for e in eventList])
    cc, lag, sub = ccx_matrix(X, len(chans[0]), engine=engine, row_end=N - 1, kernel=kernel)
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


def get_delays(DFcc, DFlag):
    """`_getDelays` + `_traceEventDendro` (construct.py:710-761) without the condensed lag vector.

    The reference walks the single linkage in merge order; for the pair (ev1 < ev2) whose distance
    caused the merge it reads that pair's *current* lag, delays every event of the merged cluster
    that does not contain ev1 by it, and rewrites all affected entries of the N(N-1)/2 lag vector
    (`_updateLags`, construct.py:764-793: pairs (b, j > b) += lag, pairs (a < b, b) -= lag).  Those
    updates are exactly `lag(i, j) = lag0(i, j) + d[i] - d[j]` with d the per-event delay so far,
    so only d is kept: one vectorised add per merge, O(N^2) in total (N = 4096: seconds; the
    reference's per-merge DataFrame scans and Python loops are O(N^3)).

    DFcc, DFlag: (N-1) x (N-1) frames / arrays as `_makeDFcclags` returns them.  Returns
    (link, delays) with delays the int64 `lagSeries` of construct.py:742 in event order.
    Duplicate coefficients raise (the reference perturbs them with unseeded random numbers,
    construct.py:814-835)."""
    cc = np.asarray(DFcc, dtype=np.float64)
    lagm = np.asarray(DFlag, dtype=np.float64)
    N = cc.shape[0]                                   # events - 1
    iu = np.triu_indices(N)                           # (row b, column c-1) with c-1 >= b
    cx = 1.0000001 - cc[iu]
    if np.isnan(cx).any():
        raise ValueError("get_delays: NaN in the upper triangle of DFcc")
    order = np.argsort(cx, kind="stable")
    if N > 1 and (np.diff(cx[order]) == 0).any():
        raise ValueError("get_delays: duplicate correlation coefficients (construct.py:814-835)")
    lag0 = lagm[iu]
    link = linkage(cx)
    members = [None] * (2 * N + 1)
    for i in range(N + 1):
        members[i] = np.array([i], dtype=np.int64)
    owner = np.arange(N + 1, dtype=np.int64)          # current cluster id of every event
    d = np.zeros(N + 1, dtype=np.int64)
    sorted_cx = cx[order]
    for a in range(N):
        i1, i2 = int(link[a, 0]), int(link[a, 1])
        p = int(order[np.searchsorted(sorted_cx, link[a, 2])])
        ev1, ev2 = int(iu[0][p]), int(iu[1][p]) + 1
        cl22 = members[i2] if owner[ev1] == i1 else members[i1]
        cur = int(np.round(lag0[p] + d[ev1] - d[ev2]))
        d[cl22] += cur
        new = N + 1 + a
        members[new] = np.concatenate([members[i1], members[i2]])
        owner[members[new]] = new
        members[i1] = members[i2] = None
    return link, d


def alignTD(delays, MPtd):
    """`_alignTD` (construct.py:486-503) on the shifted delays of construct.py:283-284:
    returns (aligned [N][len], SampleDelays).  Raises like the reference's
    `detex.log(level='error')` when nothing is left."""
    d = np.asarray(delays, dtype=np.int64)
    d = d - int(d.min())
    length = len(MPtd[0]) - int(d.max())
    if length <= 0:
        raise Exception('Alignment of multiplexed stream failing, try raising ccreq or widenning '
                        'trim window')
    return np.array([np.asarray(x)[k:k + length] for x, k in zip(MPtd, d)]), d


def update_start_times(stats, sample_delays, origin_times, magnitudes):
    """`_updateStartTimes` (construct.py:346-366): start time after the alignment trim, predicted
    origin-to-window offset.  stats: list of dicts {Nc, sampling_rate, starttime, ...}."""
    out = []
    for st, k, ot, mag in zip(stats, sample_delays, origin_times, magnitudes):
        st = dict(st)
        new = st['starttime'] + k / (st['sampling_rate'] * st['Nc'])
        st.update(starttime=new, origintime=ot, magnitude=mag, offset=new - ot)
        out.append(st)
    return out
