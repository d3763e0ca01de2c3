"""fas.py -- host-side mirror of the reference's false-alarm-statistic path
(`detex/fas.py`) on top of the CUDA engine.

  _MPXSSCorr(MPcon, reqlen, ssArrayTD, ssArrayFD, Nc)          fas.py:120-134
  initFAS(bases, chunks, Nc, numBins)                           fas.py:23-86 (array part)

The GPU computes, per subspace and over all null chunks at once: the DS vectors, their
histogram on linspace(-.01, 1, numBins) and the sufficient statistics
(N, sum x, sum x^2, sum log x, sum log1p(-x)).  `scipy.stats.beta.fit(dss, floc=0, fscale=1)`
(fas.py:81) only needs those: its fixed-loc/scale branch solves the two digamma equations
with a method-of-moments start; `beta_fit_from_stats` restates exactly that, and
`beta.nnlf` (fas.py:84) is a closed form of the same sums.
"""
import numpy as np
import scipy.optimize
import scipy.special as sc

from .detect import default_engine


def _MPXSSCorr(MPcon, reqlen, ssArrayTD, ssArrayFD, Nc, engine=None, kernel="tcgen05"):
    """Drop-in for `fas._MPXSSCorr` (fas.py:120-134); FFT operands are ignored."""
    eng = engine or default_engine()
    eng.set_bases(-2, [np.atleast_2d(np.asarray(ssArrayTD, dtype=np.float64))], int(Nc))
    eng.load_chunks([np.asarray(MPcon)])
    eng.detect_run(-2, engine=kernel)
    return eng.get_ds(0, 0).astype(np.float64)


def beta_fit_from_stats(N, sx, sxx, slog, slog1m):
    """scipy.stats.beta.fit(data, floc=0, fscale=1) from the data's sufficient statistics
    (scipy/stats/_continuous_distns.py, beta_gen.fit, fixed loc and scale branch)."""
    if not all(np.isfinite(v) for v in (N, sx, sxx, slog, slog1m)) or N < 2:
        raise RuntimeError("beta fit failed: non-finite sufficient statistics (DS outside (0, 1)?)")
    xbar = sx / N
    var = sxx / N - xbar * xbar
    fac = xbar * (1 - xbar) / var - 1
    a0, b0 = xbar * fac, (1 - xbar) * fac

    def func(x):
        a, b = x
        return [slog - N * (-sc.psi(a + b) + sc.psi(a)), slog1m - N * (-sc.psi(a + b) + sc.psi(b))]

    theta, info, ier, mesg = scipy.optimize.fsolve(func, [a0, b0], full_output=True)
    if ier != 1:
        raise RuntimeError("beta fit failed: " + mesg)
    a, b = theta
    if not (a > 0 and b > 0):
        raise RuntimeError("beta fit failed: non-positive shape")
    return float(a), float(b), 0, 1


def beta_nnlf_from_stats(a, b, N, slog, slog1m):
    """scipy.stats.beta.nnlf((a, b, 0, 1), data) = -sum(logpdf)."""
    return -((a - 1) * slog + (b - 1) * slog1m - N * sc.betaln(a, b))


def screen_chunks(chunks, Nc, sr, zchan=None, STATime=0.5, LTATime=5, staltalimit=8.0, engine=None, batch=32):
    """`_checkSTALTA` for every candidate chunk (fas.py:175-205): True where the classic
    STA/LTA of the screening channel (Z = last channel in ObsPy's sorted order by default)
    never exceeds `staltalimit`.  staltalimit=None passes everything."""
    if staltalimit is None:
        return [True] * len(chunks)
    eng = engine or default_engine()
    zchan = Nc - 1 if zchan is None else zchan
    out = []
    for i in range(0, len(chunks), batch):
        eng.load_chunks(chunks[i:i + batch])
        mx = eng.sta_lta_max(Nc, zchan, int(STATime * sr), int(LTATime * sr))
        out.extend(bool(m <= staltalimit) for m in mx)
    return out


def select_null_chunks(passes, conDatNum):
    """Chunk bookkeeping of `_getDSVect` / `_initFAS` (fas.py:56-71, 96-112): keep screened
    chunks in order until conDatNum are kept; if <= 25 % pass, drop the screen."""
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


def initFAS(bases, chunks, Nc, numBins=401, engine=None, set_id=900, kernel="tcgen05", batch=16):
    """Array part of `_initFAS` (fas.py:23-86).  bases: list of (r, n) arrays (one per
    subspace / single, same n); chunks: the null-space multiplexed chunks that passed the
    STA/LTA screen.  Returns a list of {'bins','hist','betadist','nnlf'} dicts."""
    if not 2 <= int(numBins) <= 1025:
        raise ValueError("numBins must be between 2 and 1025 (the device histogram holds up to 1024 bins)")
    eng = engine or default_engine()
    eng.set_bases(set_id, bases, Nc)
    eng.set_hist_bins(int(numBins) - 1)                    # np.linspace(-.01, 1, numBins) has numBins - 1 bins
    try:
        eng.hist(set_id, reset=True)
        eng.fas(set_id, reset=True)
        for i in range(0, len(chunks), batch):
            eng.load_chunks(chunks[i:i + batch])
            eng.detect_run(set_id, engine=kernel, hist_range=(-.01, 1.0), want_fas=True)
        hist = eng.hist(set_id, reset=True)
        st = eng.fas(set_id, reset=True)
    finally:
        eng.set_hist_bins(400)                             # detection histograms: 401 edges (detect.py:80)
    bins = np.linspace(-.01, 1, num=numBins)
    out = []
    for s in range(len(bases)):
        N, sx, sxx, slog, slog1m = st[s]
        a, b, loc, scale = beta_fit_from_stats(N, sx, sxx, slog, slog1m)
        out.append({'bins': bins, 'hist': hist[s], 'betadist': (a, b, loc, scale),
                    'nnlf': beta_nnlf_from_stats(a, b, N, slog, slog1m)})
    return out
