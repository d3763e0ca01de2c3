"""detex_oracle.py -- CPU (NumPy float64) restatement of Detex's hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under `detex_b200/` may import this module;
only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu-baseline /
`--impl reference` legs do, and only as the checker / the CPU arm.

Parity status: **pinned against the reference itself** -- every function below is
checked (tests/test_oracle_vs_reference.py, runs where /root/reference exists)
against the unmodified reference function imported through `oracle/ref_shim.py`,
and against golden vectors generated from the reference by
`tests/golden/make_golden.py` (committed, so the check also runs on the GPU box).
The reference's own test-suite pins no numerical result on this path
(SURVEY.md section 8c), so those vectors are the pin.

Two flavours are kept for the heavy functions:
  * `*_fft`    : follows the reference's ALGORITHM line by line (FFT correlation
                 + rolling statistics).  This is the "port" timed as the CPU baseline.
  * `*_direct` : the closed form, evaluated with explicit dot products in float64.
                 Independent of FFT round-off; used to check the CUDA path on
                 small shapes.

All citations are file:line under /root/reference.
"""
from __future__ import annotations

import numpy as np
import scipy.fft
import scipy.stats
import scipy.linalg
from scipy.cluster.hierarchy import linkage

# --------------------------------------------------------------------------
# rolling statistics (pandas 0.17 pd.rolling_mean / rolling_var / rolling_std,
# ddof=1, as used at detect.py:567-568, fas.py:126-127, construct.py:446-447)
# --------------------------------------------------------------------------


def _window_sums(x, n):
    """S1[t] = sum x[t:t+n], S2[t] = sum x[t:t+n]**2 for t = 0..len(x)-n.  float64 prefix sums
    of the (already centred, see callers) data: the prefix of ~1e6 unit-variance samples is
    exact to ~1e-10, i.e. 1e-14 of a window's energy."""
    x = np.asarray(x, dtype=np.float64)
    c1 = np.concatenate(([0.0], np.cumsum(x)))
    c2 = np.concatenate(([0.0], np.cumsum(np.square(x))))
    return c1[n:] - c1[:-n], c2[n:] - c2[:-n]


def rolling_mean(x, n):
    """pd.rolling_mean(x, n)[n-1:]"""
    x = np.asarray(x, dtype=np.float64)
    m = x.mean()
    s1, _ = _window_sums(x - m, n)
    return s1 / n + m


def rolling_var(x, n):
    """pd.rolling_var(x, n)[n-1:]  (ddof = 1)"""
    x = np.asarray(x, dtype=np.float64)
    # centre on the global mean first: variance is shift invariant and this keeps
    # S2 - S1^2/n well conditioned when the DC level is large.
    s1, s2 = _window_sums(x - x.mean(), n)
    return np.maximum(s2 - s1 * s1 / n, 0.0) / (n - 1.0)


def rolling_std(x, n):
    return np.sqrt(rolling_var(x, n))


# --------------------------------------------------------------------------
# a1  multiplex  (construct.py:928-987)
# --------------------------------------------------------------------------


def multiplex(chans):
    """Interleave channels [c0[0], c1[0], .., c0[1], ..]; channels are trimmed to the
    shortest (construct.py:964-978)."""
    if len(chans) == 1:
        return np.asarray(chans[0], dtype=np.float64)
    m = min(len(c) for c in chans)
    C = np.vstack([np.asarray(c, dtype=np.float64)[:m] for c in chans])
    return C.flatten(order="F")


# --------------------------------------------------------------------------
# a3  detection statistic  (detect.py:559-578 == fas.py:120-134)
# --------------------------------------------------------------------------


def mpx_ds_fft(MPcon, U, Nc):
    """Port of `_SSDetex._MPXDS`: FFT correlation of every basis vector with the
    multiplexed chunk, mean correction, square-sum over the basis, divide by the
    running window power, keep every Nc-th lag.

    MPcon : (L,) multiplexed chunk;  U : (r, n) basis rows;  returns (T,), T=(L-n)//Nc+1.
    """
    MPcon = np.asarray(MPcon, dtype=np.float64)
    U = np.atleast_2d(np.asarray(U, dtype=np.float64))
    n = U.shape[1]
    L = len(MPcon)
    reqlen = int(L + n)                                   # detect.py:368
    nfft = 2 ** reqlen.bit_length()                       # detect.py:369 / :255
    ssFD = scipy.fft.fft(U[:, ::-1], n=nfft, axis=1)      # detect.py:371
    MPconFD = scipy.fft.fft(MPcon, n=nfft)                # detect.py:256
    a = rolling_mean(MPcon, n)                            # detect.py:567
    b = rolling_var(MPcon, n) * n                         # detect.py:568-569
    sum_ss = U.sum(axis=1)                                # detect.py:570
    av_norm = a[None, :] * sum_ss[:, None]                # detect.py:571-573
    m1 = ssFD * MPconFD[None, :]                          # detect.py:574
    if1 = np.real(scipy.fft.ifft(m1, axis=1))[:, n - 1:L] - av_norm   # detect.py:576
    with np.errstate(divide="ignore", invalid="ignore"):
        result = np.sum(np.square(if1), axis=0) / b       # detect.py:577
    return result[::Nc]                                   # detect.py:578


def mpx_ds_direct(MPcon, U, Nc, block=2048):
    """Closed form of `_MPXDS` (verified 5e-16 against the reference):
    DS[t] = ((n-1)/n) * sum_k (u_k . (w_t - mean w_t))^2 / ||w_t - mean w_t||^2,
    w_t = MPcon[t*Nc : t*Nc + n].  Explicit float64 dot products, no FFT."""
    MPcon = np.asarray(MPcon, dtype=np.float64)
    U = np.atleast_2d(np.asarray(U, dtype=np.float64))
    n = U.shape[1]
    L = len(MPcon)
    T = (L - n) // Nc + 1
    xc = MPcon - MPcon.mean()
    W = np.lib.stride_tricks.sliding_window_view(xc, n)[::Nc]
    out = np.empty(T)
    for t0 in range(0, T, block):
        w = W[t0:t0 + block]
        mu = w.mean(axis=1, keepdims=True)
        wc = w - mu
        num = np.square(wc @ U.T).sum(axis=1)
        den = np.square(wc).sum(axis=1)
        with np.errstate(divide="ignore", invalid="ignore"):
            out[t0:t0 + block] = ((n - 1.0) / n) * num / den
    return out


# --------------------------------------------------------------------------
# a4  per-chunk bookkeeping of _getRA  (detect.py:262-281)
# --------------------------------------------------------------------------


def chunk_ds(MPcon, U, Nc, fft=True):
    """DS vector + MaxDS with the reference's inf-zeroing rule (detect.py:275-281).
    Returns None if the chunk must be skipped (detect.py:262-274)."""
    U = np.atleast_2d(U)
    if len(MPcon) <= max(U.shape):
        return None
    ssd = mpx_ds_fft(MPcon, U, Nc) if fft else mpx_ds_direct(MPcon, U, Nc)
    if len(ssd) < 10:
        return None
    mx = ssd.max()
    if mx > 1.1:
        ssd[np.isinf(ssd)] = 0
        mx = ssd.max()
    return ssd, mx


# --------------------------------------------------------------------------
# a5  STA/LTA of the detection statistic  (detect.py:501-524)
# --------------------------------------------------------------------------


def _rolling_mean_centered(x, W):
    """pd.rolling_mean(x, W, center=True): window [i-W+1+off, i+off], off=(W-1)//2."""
    W = int(W)
    x = np.asarray(x, dtype=np.float64)
    out = np.full(len(x), np.nan)
    if len(x) < W:
        return out
    s1, _ = _window_sums(x, W)
    off = (W - 1) // 2
    # trailing mean ends at index j = W-1 .. len-1 ; centred value sits at j - off
    out[W - 1 - off: len(x) - off] = s1 / W
    return out


def _replace_nan_with_mean(arg):
    """detect.py:517-524 (note: leading NaNs take the value at first+1)."""
    ind = np.where(~np.isnan(arg))[0]
    first, last = ind[0], ind[-1]
    arg[:first] = arg[first + 1]
    arg[last + 1:] = arg[last]
    return arg


def sta_lta(C, LTA, STA):
    """`_getStaLtaArray` (detect.py:501-515)."""
    C = np.asarray(C, dtype=np.float64)
    if STA == 0:
        st = np.abs(C)
    else:
        st = _replace_nan_with_mean(_rolling_mean_centered(np.abs(C), STA))
    lt = _replace_nan_with_mean(_rolling_mean_centered(np.abs(C), LTA))
    with np.errstate(divide="ignore", invalid="ignore"):
        return st / lt


# --------------------------------------------------------------------------
# a6-a8  histogram, trigger test, greedy trigger picking
# --------------------------------------------------------------------------

HIST_BINS = np.linspace(0, 1, num=401)          # detect.py:80
FAS_BINS = np.linspace(-.01, 1, num=401)        # fas.py:31


def ds_histogram(DS, bins=HIST_BINS):
    """np.histogram(row.SSdetect, bins=self.hist['Bins'])  (detect.py:178-181)."""
    return np.histogram(DS, bins=bins)[0]


def downplay(C, sr, dpv=0, buff=20):
    """`_downPlayArrayAroundMax` (detect.py:545-557) -- in place, returns C."""
    index = C.argmax()
    if index < buff * sr + 1:
        C[0:int(index + buff * sr)] = dpv
    elif index > len(C) - buff * sr:
        C[int(index - sr * buff):] = dpv
    else:
        C[int(index - sr * buff):int(sr * buff + index)] = dpv
    return C


def greedy_triggers(DS, thr, sr, start, offsets, stalta=None, buff=20, max_count=4000):
    """`_CreateCoeffArray` with trigCon=0, estimateMags=False (detect.py:390-445).
    Returns a list of dict rows (DS, DS_STALTA, STMP, MSTAMPmin, MSTAMPmax, index)."""
    Ceval = np.array(DS, dtype=np.float64, copy=True)
    rows = []
    count = 0
    minof, maxof = np.min(offsets), np.max(offsets)
    while Ceval.max() >= thr:                               # detect.py:410
        trig = int(Ceval.argmax())
        coef = float(DS[trig])
        times = float(trig) / sr + start                    # detect.py:413
        sl = float(stalta[trig]) if stalta is not None else 0.0
        downplay(Ceval, sr, 0, buff)                        # detect.py:421
        if count > max_count:                               # detect.py:433-436
            raise Exception("over 4000 events found in single data block")
        rows.append(dict(DS=coef, DS_STALTA=sl, STMP=times, MSTAMPmin=times - maxof,
                         MSTAMPmax=times - minof, index=trig))
        count += 1
    return rows


def eval_trig_con(maxDS, thr):
    """`_evalTrigCon` trigCon=0 (detect.py:526-543): strict >."""
    return maxDS > thr


# --------------------------------------------------------------------------
# a10-a12  pairwise CCX  (construct.py:369-466)
# --------------------------------------------------------------------------


def sub_samp(Ceval, ind):
    """`_subSamp` cosine-fit interpolation (construct.py:397-422).  Mirrors the
    reference's quirk of returning the integer index when |tau| > .5."""
    if ind == 0 or ind == len(Ceval) - 1:
        return 0.0
    cb4, caf, cn = Ceval[ind - 1], Ceval[ind + 1], Ceval[ind]
    with np.errstate(divide="ignore", invalid="ignore"):
        alpha = np.arccos((cb4 + caf) / (2 * cn))
        alsi = np.sin(alpha)
        tau = -(np.arctan((cb4 - caf) / (2 * cn * alsi)) / alpha)
    if abs(tau) > .5:
        return ind
    return tau


def _ccx_finish(result, trunc, Nc, n):
    """Shared tail of `_CCX2` (construct.py:453-466)."""
    if np.all(np.isnan(result)):
        return 0.0, 0.0, 0.0
    maxcc = np.nanmax(result)
    mincc = np.nanmin(result)
    maxind = int(np.nanargmax(result))
    if maxcc > 1. or mincc < -1.:
        result = result.copy()
        with np.errstate(invalid="ignore"):
            result[(result > 1) | (result < -1)] = 0
        maxcc = np.nanmax(result)
        maxind = int(np.nanargmax(result))
    subsamp = sub_samp(result, maxind)
    return maxcc, (maxind + 1 + trunc) * Nc - n, subsamp


def ccx_lag_series_fft(x1, x2, Nc):
    """The normalised, stride-Nc, truncated correlation series of `_CCX2`
    (construct.py:438-453), FFT flavour.  Returns (result, trunc)."""
    x1 = np.asarray(x1, dtype=np.float64)
    x2 = np.asarray(x2, dtype=np.float64)
    n = len(x1)
    trunc = n // (2 * Nc) - 1
    nfft = 2 ** int(2 * n).bit_length()                    # construct.py:673
    f1 = scipy.fft.fft(x1, n=nfft)
    f2 = scipy.fft.fft(x2, n=nfft)
    pad = np.pad(x2, (n - 1, n - 1))
    a = rolling_mean(pad, n)
    b = rolling_std(pad, n) * np.sqrt((n - 1.0) / n)
    c = np.real(scipy.fft.ifft(np.conj(f1) * f2))
    c1 = np.concatenate([c[-(n - 1):], c[:n]])
    with np.errstate(divide="ignore", invalid="ignore"):
        result = ((c1 - x1.sum() * a) / (n * b * np.std(x1)))[Nc - 1::Nc]
    return result[trunc:-trunc], trunc


def ccx_lag_series_direct(x1, x2, Nc):
    """Closed form: res[m] = Pearson(x1, zero-padded x2[k : k+n]) at lag
    k = (m + trunc + 1)*Nc - n  (SURVEY.md 8a), explicit float64 dot products."""
    x1 = np.asarray(x1, dtype=np.float64)
    x2 = np.asarray(x2, dtype=np.float64)
    n = len(x1)
    trunc = n // (2 * Nc) - 1
    pad = np.pad(x2, (n - 1, n - 1))
    nl = (2 * n - 1 - (Nc - 1) + Nc - 1) // Nc             # len(range(Nc-1, 2n-1, Nc))
    idx = np.arange(Nc - 1, 2 * n - 1, Nc)[trunc:nl - trunc]
    x1c = x1 - x1.mean()
    sd1 = np.std(x1)
    res = np.empty(len(idx))
    for m, i in enumerate(idx):
        w = pad[i:i + n]
        sdw = np.std(w)
        with np.errstate(divide="ignore", invalid="ignore"):
            res[m] = np.dot(x1c, w - w.mean()) / (n * sdw * sd1)
    return res, trunc


def ccx2(x1, x2, Nc, fft=True):
    """`_CCX2` (construct.py:425-466): (maxcc, sampleLag, subsamp)."""
    n = len(x1)
    res, trunc = (ccx_lag_series_fft if fft else ccx_lag_series_direct)(x1, x2, Nc)
    return _ccx_finish(res, trunc, Nc, n)


def make_cclags(X, Nc, fft=True):
    """`_makeDFcclags` (construct.py:369-394) on an (N, n) array of multiplexed
    waveforms.  Returns three (N-1, N-1) float arrays (row b = 0..N-2, column c-1
    for c = 1..N-1), NaN below the diagonal -- the values of DFcc, DFlag, DFsubsamp."""
    N = X.shape[0]
    cc = np.full((N - 1, N - 1), np.nan)
    lag = np.full((N - 1, N - 1), np.nan)
    ss = np.full((N - 1, N - 1), np.nan)
    for b in range(N - 1):
        for c in range(b + 1, N):
            m, l, s = ccx2(X[b], X[c], Nc, fft=fft)
            cc[b, c - 1], lag[b, c - 1], ss[b, c - 1] = m, l, s
    return cc, lag, ss


def flat_no_nan(cc):
    """`_flatNoNan(1.0000001 - DFcc)` (construct.py:153-155): SciPy condensed order."""
    v = (1.0000001 - cc).flatten()
    return v[~np.isnan(v)]


def cluster_link(cc):
    """createCluster tail (construct.py:152-157): single linkage on 1.0000001 - CC."""
    return linkage(flat_no_nan(cc))


# --------------------------------------------------------------------------
# N3  alignment from the dendrogram  (construct.py:272-286, 486-503, 710-812)
# --------------------------------------------------------------------------


def _triangular(n):
    return n * (n + 1) // 2


def get_delays(cc, lag):
    """`_getDelays` + `_traceEventDendro` (construct.py:710-761), literally: walk the single
    linkage of `1.0000001 - cc` in merge order; the pair (ev1 < ev2) whose distance made the merge
    gives `currentLag = round(lag of that pair, as updated so far)`; every event b of the merged
    cluster that does NOT contain ev1 is delayed by it, and the condensed lag vector is updated
    with `_updateLags` (construct.py:764-793): pairs (b, j > b) += currentLag, pairs (a < b, b)
    -= currentLag.  cc, lag: (N-1) x (N-1) arrays laid out as DFcc / DFlag (row b, column c-1,
    NaN below the diagonal).  Coefficients must be unique (the reference perturbs duplicates with
    unseeded random numbers, construct.py:814-835; that branch cannot be pinned).
    Returns (link, delays[int64 per event])."""
    cc = np.asarray(cc, dtype=np.float64)
    cx = flat_no_nan(cc)
    if len(np.unique(cx)) != len(cx):
        raise ValueError("duplicate correlation coefficients (construct.py:814-835)")
    lagm = np.asarray(lag, dtype=np.float64).flatten()
    lags = lagm[~np.isnan(lagm)].copy()
    link = linkage(cx)
    N = len(link)                      # events - 1
    members = {i: [i] for i in range(N + 1)}
    for a in range(N):                 # _getClustDict, construct.py:799-811
        members[N + 1 + a] = members[int(link[a, 0])] + members[int(link[a, 1])]
    delays = np.zeros(N + 1, dtype=np.int64)
    for a in range(N):
        p = int(np.where(cx == link[a, 2])[0][0])
        ev1 = 0
        while p >= _triangular(N) - _triangular(N - (ev1 + 1)):
            ev1 += 1
        i1, i2 = int(link[a, 0]), int(link[a, 1])
        cl22 = members[i2] if ev1 in members[i1] else members[i1]
        cur = int(np.round(lags[p]))
        for b in cl22:
            delays[b] += cur
            acr0 = _triangular(N) - _triangular(N - b)            # _getAcr
            for k in range(N - b):
                lags[acr0 + k] += cur
            for k in range(b):                                    # _getDow
                lags[_triangular(N - 1) - 1 + b - _triangular(N - 1 - k)] -= cur
    return link, delays


def align_td(delays, X):
    """`_alignTD` (construct.py:486-503) after `delays - min(delays)` (construct.py:283-284):
    drop the first SampleDelays samples of each multiplexed waveform, keep the common length."""
    d = np.asarray(delays, dtype=np.int64) - int(np.min(delays))
    length = len(X[0]) - int(d.max())
    return np.array([np.asarray(x)[k:][:length] for x, k in zip(X, d)])


# --------------------------------------------------------------------------
# a14-a15  FAS statistics and thresholds  (fas.py:74-84, subspace.py:1027-1047,1110-1140)
# --------------------------------------------------------------------------


def fas_stats(ds_list, num_bins=401):
    """Tail of `_initFAS` (fas.py:74-84) on a list of DS vectors."""
    dss = np.concatenate([np.asarray(d, dtype=np.float64) for d in ds_list])
    bins = np.linspace(-.01, 1, num=num_bins)
    hist = np.histogram(dss, bins=bins)[0]
    betaparams = scipy.stats.beta.fit(dss, floc=0, fscale=1)
    nnlf = scipy.stats.beta.nnlf(betaparams, dss)
    return dict(bins=bins, hist=hist, betadist=betaparams, nnlf=nnlf)


def beta_sufficient_stats(dss):
    """(N, sum x, sum x^2, sum log x, sum log1p(-x)) -- what the GPU FAS epilogue returns."""
    dss = np.asarray(dss, dtype=np.float64)
    return (len(dss), dss.sum(), np.square(dss).sum(), np.log(dss).sum(), np.log1p(-dss).sum())


def approx_thld(beta_a, beta_b, target, numint=1000, numloops=3, backup=None):
    """`_approxThld` forward grid search (subspace.py:1110-1140)."""
    start, stop = 0, 1
    for _ in range(numloops):
        Xs = np.linspace(start, stop, numint)
        pfs = scipy.stats.beta.sf(Xs, beta_a, beta_b)
        resid = np.abs(pfs - target)
        i = int(resid.argmin())
        if i == 0 or i == numint - 1:
            if backup is None:
                raise ValueError("grid search for threshold failing")
            return backup, target
        bestX, bestPf = Xs[i], pfs[i]
        start, stop = Xs[i - 1], Xs[i + 1]
    return bestX, bestPf


def threshold_from_beta(beta_a, beta_b, Pf=1e-12, backup=None):
    """subspace.py:1034-1047."""
    th = scipy.stats.beta.isf(Pf, beta_a, beta_b, 0, 1)
    if th > .9:
        th, _ = approx_thld(beta_a, beta_b, Pf, 1000, 3, backup)
    return th


# --------------------------------------------------------------------------
# a16  SVD / basis selection  (subspace.py:875-905, 921-943, 968-1013)
# --------------------------------------------------------------------------


def svd_basis(aligned, select_criteria=2, select_value=0.9, normalize=False):
    """aligned : (events, n) trimmed aligned waveforms (NOT demeaned).
    Returns dict(U=(r,n) used basis rows in descending singular value, s, frac_avg,
    frac_min, ndim) following SubSpace.SVD."""
    aligned = np.asarray(aligned, dtype=np.float64)
    arr = aligned - aligned.mean(axis=1, keepdims=True)        # subspace.py:932-933
    if normalize:
        arr = arr / np.linalg.norm(arr, axis=1, keepdims=True)
    U, s, Vh = scipy.linalg.svd(arr.T, full_matrices=False)     # subspace.py:890
    # fractional energy capture of the (un-demeaned) waveforms (subspace.py:968-997)
    fr = []
    for w in aligned:
        rep = np.square(U.T @ w / np.linalg.norm(w))
        fr.append(np.concatenate(([0.0], np.cumsum(rep))))
    fr = np.array(fr)
    avg = fr.mean(axis=0)
    mn = fr.min(axis=0)
    if select_criteria in (1, 2, 3):
        avg[-1] = 1.0                                           # subspace.py:1008
        ndim = int(np.argmax(avg >= select_value))
    else:
        ndim = int(select_value) + 1
    return dict(U=U[:, :ndim].T.copy(), s=s, frac_avg=avg, frac_min=mn, ndim=ndim)


def single_basis(x):
    """Singleton template: unit-norm, NOT demeaned (detect.py:356-357, fas.py:144)."""
    x = np.asarray(x, dtype=np.float64)
    return (x / np.linalg.norm(x))[None, :]


# --------------------------------------------------------------------------
# a14 (screen)  fas._checkSTALTA  (fas.py:175-205)
# The arithmetic is ObsPy's classic_sta_lta (obspy 1.0.2, pinned in the reference's
# environment.txt, NOT vendored): restated from obspy/signal/trigger.py::classic_sta_lta_py
# == obspy/signal/src/stalta.c.  ObsPy is not installable here, so THIS function's parity is
# unpinned; everything downstream of the screen is pinned.
# --------------------------------------------------------------------------


def classic_sta_lta(a, nsta, nlta):
    a = np.asarray(a, dtype=np.float64)
    nsta, nlta = int(nsta), int(nlta)
    sta = np.cumsum(a ** 2)
    lta = sta.copy()
    sta[nsta:] = sta[nsta:] - sta[:-nsta]
    sta /= nsta
    lta[nlta:] = lta[nlta:] - lta[:-nlta]
    lta /= nlta
    sta[:nlta - 1] = 0
    dtiny = np.finfo(0.0).tiny
    lta[lta < dtiny] = dtiny
    return sta / lta


def check_stalta(z, sr, STATime, LTATime, limit):
    """`_checkSTALTA` on the screening channel (fas.py:191-197)."""
    if limit is None:
        return True
    cft = classic_sta_lta(z, STATime * sr, LTATime * sr)
    return bool(np.max(cft) <= limit)


def select_null_chunks(passes, conDatNum):
    """Bookkeeping of `_getDSVect` / `_initFAS` (fas.py:56-71, 96-112): walk the candidate
    chunks in order, keep those that pass the screen until conDatNum are kept; if <= 25 %
    pass, drop the screen.  Returns the kept indices."""
    def walk(flags):
        kept, count = [], 0
        for i, ok in enumerate(flags):
            count += 1
            if not ok:
                continue
            if len(kept) >= conDatNum:
                break
            kept.append(i)
        return kept, count
    kept, count = walk(passes)
    if count and float(len(kept)) / count <= .25:
        kept, count = walk([True] * len(passes))
    return kept


# --------------------------------------------------------------------------
# N1  per-detection magnitude / SNR  (detect.py:447-499, 637-664; construct.py:469-483)
# --------------------------------------------------------------------------


def fast_normcorr(t, s):
    """`construct.fast_normcorr` (construct.py:469-483)."""
    t = np.asarray(t, dtype=np.float64)
    s = np.asarray(s, dtype=np.float64)
    if len(t) > len(s):
        t, s = s, t
    n = len(t)
    nt = (t - np.mean(t)) / (np.std(t) * n)
    sum_nt = nt.sum()
    a = rolling_mean(s, n)
    b = rolling_std(s, n) * np.sqrt((n - 1.0) / n)
    c = np.convolve(nt[::-1], s, mode="valid")
    return (c - sum_nt * a) / b


def est_mag(trigIndex, MPcon, Nc, U, ewf, mags, issubspace=True):
    """`_SSDetex._estMag` (detect.py:447-499).  U: (r, n) basis; ewf: (events, n) trimmed
    waveforms (single: the trimmed template); returns (ProEnMag, Mag, SNR)."""
    MPcon = np.asarray(MPcon, dtype=np.float64)
    U = np.atleast_2d(np.asarray(U, dtype=np.float64))
    ewf = np.atleast_2d(np.asarray(ewf, dtype=np.float64))
    mags = np.asarray(mags, dtype=np.float64)
    WFU = (ewf @ U.T) @ U                                    # detect.py:381 without the n x n UtU
    WFlen = WFU.shape[1]
    ConDat = MPcon[trigIndex * Nc: trigIndex * Nc + WFlen]
    if issubspace:
        ssCon = U.T @ (U @ ConDat)                           # detect.py:460
        proEn = np.var(ssCon) / np.var(WFU, axis=1)          # :462
    if trigIndex * Nc > 5 * WFlen:                           # :465-470
        pe = MPcon[trigIndex * Nc - 5 * WFlen: trigIndex * Nc]
    else:
        pe = MPcon[trigIndex * Nc: trigIndex * Nc + WFlen + 6 * WFlen]
    rstd = rolling_std(pe, WFlen) if len(pe) >= WFlen else np.array([])
    baseNoise = np.median(rstd) if len(rstd) else np.nan
    SNR = np.std(ConDat) / baseNoise
    touse = mags > -15
    if issubspace:
        if not np.any(touse):
            return np.nan, np.nan, SNR
        cors = np.array([fast_normcorr(x, ConDat)[0] for x in ewf])
        den = np.sum(np.square(cors[touse]))
        pe_mag = sum((mags[i] + np.log10(np.sqrt(proEn[i]))) * cors[i] ** 2 for i in range(len(mags)) if mags[i] > -15) / den
        st_mag = sum((mags[i] + np.log10(np.std(ConDat) / np.std(ewf[i]))) * cors[i] ** 2
                     for i in range(len(mags)) if mags[i] > -15) / den
        return pe_mag, st_mag, SNR
    if np.isnan(mags[0]) or mags[0] < -15:
        return np.nan, np.nan, SNR
    d1 = np.dot(ConDat, WFU[0])
    d2 = np.dot(WFU[0], WFU[0])
    return mags[0] + d1 / d2, mags[0] + np.log10(np.std(ConDat) / np.std(WFU[0])), SNR


# --------------------------------------------------------------------------
# N2  pre-processing: array part of construct._applyFilter (construct.py:1017-1029)
# ObsPy 1.0.2 (not vendored): Trace.detrend('linear') -> scipy.signal.detrend(type='linear');
# Trace.filter('bandpass') -> obspy/signal/filter.py::bandpass (iirfilter + zpk2sos + sosfilt,
# zerophase = second sosfilt over the reversed trace).  ObsPy is not installable here, so the
# ObsPy-side wiring is restated from its published source (parity of this step unpinned); the
# arithmetic itself is SciPy's, which IS run here.
# --------------------------------------------------------------------------


def apply_filter(chans, sr, filt=(1, 10, 2, True), decimate=None):
    """chans: list of 1-D channel arrays (sorted order).  Returns the multiplexed chunk.
    decimate: `st.decimate(factor)` first (construct.py:1014-1015) = ObsPy 1.0.2 Trace.decimate:
    filter('lowpass_cheby_2', freq=0.5*sr/factor, maxorder=12) then data[::factor]
    (obspy/core/trace.py, obspy/signal/filter.py::lowpass_cheby_2; restated, unpinned)."""
    import scipy.signal
    out = []
    # channels are trimmed to their common window AFTER the optional decimation and BEFORE detrend /
    # filter (st.trim(startTrim, endTrim), construct.py:1019-1024); with a shared start time that is
    # the first min-length samples of every channel
    fac = int(decimate) if decimate and int(decimate) > 1 else 1
    common = min(-(-len(x) // fac) for x in chans)
    for x in chans:
        y = np.asarray(x, dtype=np.float64)
        fs = sr
        if decimate and int(decimate) > 1:
            nyq = 0.5 * sr
            ws = (0.5 * sr / float(decimate)) / nyq
            wp, order = ws, 1e99
            while order > 12:
                wp = wp * 0.99
                order, wn = scipy.signal.cheb2ord(wp, ws, 1, 96, analog=0)
            z, p, k = scipy.signal.cheby2(order, 96, wn, btype="low", analog=0, output="zpk")
            y = scipy.signal.sosfilt(scipy.signal.zpk2sos(z, p, k), y)[::int(decimate)]
            fs = sr / float(decimate)
        y = y[:common]
        y = scipy.signal.detrend(y, type="linear")
        if filt is not None:
            fe = 0.5 * fs
            z, p, k = scipy.signal.iirfilter(filt[2], [filt[0] / fe, filt[1] / fe], btype="band", ftype="butter",
                                             output="zpk")
            sos = scipy.signal.zpk2sos(z, p, k)
            y = scipy.signal.sosfilt(sos, y)
            if filt[3]:
                y = scipy.signal.sosfilt(sos, y[::-1])[::-1]
        out.append(y)
    return multiplex(out)
