"""detect.py -- host-side mirror of the reference's detection hot path
(`detex/detect.py`, class `_SSDetex`) on top of the CUDA engine.

Same names, argument meaning, return objects and error behaviour as the reference's
private callables, so a Detex maintainer can swap them in:

  _MPXDS(MPcon, reqlen, ssTD, ssFD, Nc, MPconFD)     detect.py:559-578
  getRA(...)                                          detect.py:220-296 (array part)
  _CreateCoeffArray(...)                              detect.py:390-445
  _downPlayArrayAroundMax / _evalTrigCon              detect.py:545-557 / 526-543
  corDat(...)                                         detect.py:137-218 (chunk loop)

All arithmetic on the path (projection, window energy, normalisation, max, histogram,
threshold compaction, LTA) runs in the CUDA library.  The greedy pick over the compacted
candidates is the reference's sequential loop, executed on the sparse list.
There is no CPU fallback: without the extension / an sm_100 GPU these raise.
"""
import logging

import numpy as np
import pandas as pd

from .engine import DtxError, Engine, ShortChunk

log = logging.getLogger("detex_b200")

CORDF_COLS = ['SSdetect', 'STALTA', 'TimeStamp', 'SampRate', 'MaxDS', 'MaxSTALTA', 'Nc', 'File']
SAR_COLS = ['DS', 'DS_STALTA', 'STMP', 'Name', 'Sta', 'MSTAMPmin', 'MSTAMPmax', 'Mag', 'SNR',
            'ProEnMag']
HIST_BINS = np.linspace(0, 1, num=401)  # detect.py:80

_default_engine = None


def default_engine():
    global _default_engine
    if _default_engine is None:
        _default_engine = Engine(0)
    return _default_engine


def _MPXDS(MPcon, reqlen, ssTD, ssFD, Nc, MPconFD=None, engine=None, kernel="tcgen05"):
    """Drop-in for `_SSDetex._MPXDS` (detect.py:559-578): detection statistic of ONE
    subspace on ONE multiplexed chunk.  `reqlen`, `ssFD`, `MPconFD` (the FFT operands) are
    accepted for signature compatibility and ignored.  Returns float64 ndarray of length
    (len(MPcon) - n)//Nc + 1."""
    eng = engine or default_engine()
    U = np.atleast_2d(np.asarray(ssTD, dtype=np.float64))
    eng.set_bases(-1, [U], int(Nc))
    eng.load_chunks([np.asarray(MPcon)])
    eng.detect_run(-1, engine=kernel)
    return eng.get_ds(0, 0).astype(np.float64)


def _downPlayArrayAroundMax(index, length, sr, buff=20):
    """Index range zeroed around a pick (detect.py:545-557), as (lo, hi) half-open."""
    if index < buff * sr + 1:
        return 0, int(index + buff * sr)
    elif index > length - buff * sr:
        return int(index - sr * buff), length
    else:
        return int(index - sr * buff), int(sr * buff + index)


def greedy_pick(t, ds, length, sr, buff=20):
    """The `while Ceval.max() >= threshold` loop (detect.py:410-421) on the compacted
    candidate list (t ascending lag indices, ds their values, all >= threshold).
    Returns the positions (into t) of the picks in pick order."""
    t = np.asarray(t)
    ds = np.asarray(ds)
    alive = np.ones(len(t), dtype=bool)
    order = np.lexsort((t, -ds))  # value descending, first occurrence on ties (argmax rule)
    picks = []
    for k in order:
        if not alive[k]:
            continue
        picks.append(k)
        lo, hi = _downPlayArrayAroundMax(int(t[k]), length, sr, buff)
        i0, i1 = np.searchsorted(t, lo, side="left"), np.searchsorted(t, hi, side="left")
        alive[i0:i1] = False
    return picks


def _evalTrigCon(maxDS, threshold):
    """detect.py:526-543, trigCon == 0."""
    return bool(maxDS > threshold)


# candidate records with 64-bit lags (a month of 100 Hz data has 2.6e8 lags; segment lags are int32)
LONG_CAND_DTYPE = np.dtype([("row", np.int64), ("t", np.int64), ("ds", np.float32), ("lta", np.float32)])


class SSDetex(object):
    """Batched detector for the subspaces (or singles) of one station.

    Mirrors the state `_SSDetex._corDat` builds with `_loadMPSubSpace` (detect.py:149-150):
    ssTD {name: U}, thresholds {name: float}, offsets {name: [..]}.
    """

    def __init__(self, ssTD, threshold, offsets, Nc, sta="", engine=None, set_id=0,
                 triggerLTATime=5, triggerSTATime=0, fillZeros=False, calcHist=True,
                 kernel="tcgen05", kblk=0, ewf=None, mags=None, issubspace=True, estimateMags=None):
        self.names = sorted(ssTD.keys())
        if not self.names:
            raise ValueError("no subspaces")
        self.ssTD = {k: np.atleast_2d(np.asarray(ssTD[k], dtype=np.float64)) for k in self.names}
        self.threshold = dict(threshold)
        self.offsets = dict(offsets)
        self.Nc = int(Nc)
        self.sta = sta
        self.engine = engine or default_engine()
        self.kernel = kernel
        self.kblk = kblk
        self.triggerLTATime = triggerLTATime
        self.triggerSTATime = triggerSTATime
        if triggerSTATime < 0:
            raise ValueError("triggerSTATime must be >= 0")
        self.fillZeros = fillZeros
        self.calcHist = calcHist
        # group by basis length n: one basis set per distinct n (SampleTrims differ per subspace)
        self.groups = {}
        for name in self.names:
            self.groups.setdefault(self.ssTD[name].shape[1], []).append(name)
        self.set_ids = {}
        for gi, (n, names) in enumerate(sorted(self.groups.items())):
            sid = set_id * 1000 + gi
            self.engine.set_bases(sid, [self.ssTD[k] for k in names], self.Nc,
                                  thresholds=[self.threshold[k] for k in names])
            self.set_ids[n] = sid
        self.histdic = {na: np.zeros(len(HIST_BINS) - 1, dtype=np.int64) for na in self.names}
        # magnitude / SNR estimation (_estMag, detect.py:447-499): needs the event waveforms and
        # magnitudes _loadMPSubSpace keeps per subspace (ewf, mags; detect.py:346-381)
        self.estimateMags = (ewf is not None and mags is not None) if estimateMags is None else estimateMags
        self.issubspace = issubspace
        if self.estimateMags:
            if ewf is None or mags is None:
                raise ValueError("estimateMags needs ewf and mags")
            for n, names in self.groups.items():
                for si, name in enumerate(names):
                    W = np.atleast_2d(np.asarray(ewf[name], dtype=np.float64))
                    U = self.ssTD[name]
                    wfu = (W @ U.T) @ U                      # WFU = WFs . UtU (detect.py:381)
                    self.engine.set_events(self.set_ids[n], si, W if issubspace else wfu[:1], mags[name],
                                           wfu_var=np.var(wfu, axis=1), is_single=not issubspace)

    # ------------------------------------------------------------------ core
    def run_chunks(self, chunks, sr, starts, keep_ds=False):
        """Detection on a batch of multiplexed chunks.

        Returns (ss_df rows as DataFrame with the reference's columns, per-chunk dicts
        {name: MaxDS}, and optionally the dense DS arrays {(chunk, name): ndarray}).
        Chunks the reference would skip (detect.py:262-274) are dropped with a warning."""
        return self._run(chunks, [len(c) for c in chunks], sr, starts, keep_ds, raw_filt=None)

    def run_raw_chunks(self, traces, sr, starts, filt=(1, 10, 2, True), keep_ds=False, decimate=None):
        """As run_chunks, but from RAW per-channel traces (list of chunks, each a list of channel
        arrays in sorted order): decimation, linear detrend, band-pass and multiplex run on the device
        too (`_applyFilter` + `multiplex`, detect.py:231-241), so the samples cross PCIe once.
        `sr` is the rate of the raw traces; lags and trigger times use sr / decimate."""
        f = int(decimate) if decimate else 1
        lens = [min(-(-len(t) // f) for t in ch) * self.Nc for ch in traces]
        return self._run(traces, lens, sr / float(f), starts, keep_ds, raw_filt=(filt, decimate, sr))

    def _run(self, chunks, lens, sr, starts, keep_ds, raw_filt):
        good = []
        for i, c in enumerate(chunks):
            L = lens[i] // self.Nc * self.Nc
            nmax = max(self.groups.keys())
            if L <= nmax or (L - nmax) // self.Nc + 1 < 10:
                log.warning("current data block on %s starting %s is shorter than template, skipping",
                            self.sta, starts[i])
                continue
            good.append(i)
        rows = []
        maxds = [dict() for _ in chunks]
        dense = {}
        if not good:
            return pd.DataFrame(columns=SAR_COLS), maxds, dense
        eng = self.engine
        if raw_filt is None:
            eng.load_chunks([chunks[i] for i in good])
        else:
            from . import preprocess
            preprocess.applyFilter([chunks[i] for i in good], raw_filt[2], raw_filt[0], decimate=raw_filt[1],
                                   engine=eng)
        W = int(self.triggerLTATime * sr)
        eng.set_trigger_sta(int(self.triggerSTATime * sr))   # detect.py:285-287
        for n, names in sorted(self.groups.items()):
            sid = self.set_ids[n]
            eng.detect_run(sid, engine=self.kernel, kblk=self.kblk, hist_range=(0.0, 1.0),
                           lta_window=0 if self.fillZeros else W)
            mx, fl = eng.rowstats()
            cand = eng.candidates()
            S = len(names)
            for gi, ci in enumerate(good):
                maxds[ci].update(zip(names, (float(v) for v in mx[gi, :S])))
                if keep_ds:
                    for si, name in enumerate(names):
                        dense[(ci, name)] = eng.get_ds(gi, si).astype(np.float64)
            # candidates grouped by (chunk, subspace) row once; only rows that triggered have any
            cand = cand[np.lexsort((cand["t"], cand["row"]))]
            urows, first = np.unique(cand["row"], return_index=True)
            last = np.r_[first[1:], len(cand)]
            for r, a, b in zip(urows, first, last):
                gi, si = divmod(int(r), S)
                ci, name = good[gi], names[si]
                if not _evalTrigCon(mx[gi, si], self.threshold[name]):
                    continue
                T = eng.num_lags(gi)
                sel = cand[a:b]
                picks = greedy_pick(sel["t"], sel["ds"], T, sr)
                if len(picks) > 4000:  # kill switch, detect.py:433-436
                    raise Exception('over 4000 events found in single data block on %s for %s'
                                    % (self.sta, name))
                minof, maxof = np.min(self.offsets[name]), np.max(self.offsets[name])
                if self.estimateMags and len(picks):
                    tt = np.asarray(sel["t"][picks], dtype=np.int32)
                    mg = eng.est_mags(sid, np.full(len(tt), gi), np.full(len(tt), si), tt)
                for pi, k in enumerate(picks):
                    coef = float(sel["ds"][k])
                    times = float(sel["t"][k]) / sr + starts[ci]      # detect.py:413
                    if self.fillZeros:
                        sl = 0.0
                    else:
                        # |DS| / lta == STA / LTA at the trigger (detect.py:501-515); a chunk shorter
                        # than a window has no STA/LTA array in the reference -> 0.0 (detect.py:416-419)
                        # (a window of less than one sample -- triggerLTATime * sr < 1 -- has none either)
                        den = float(sel["lta"][k])
                        sl = abs(coef) / den if (np.isfinite(den) and den > 0.0) else 0.0
                    pe_mag, st_mag, snr = (mg[pi] if self.estimateMags else (np.nan, np.nan, np.nan))
                    rows.append([coef, sl, times, name, self.sta, times - maxof, times - minof,
                                 st_mag, snr, pe_mag])       # Mag = stMag, ProEnMag = peMag (detect.py:428,442)
            if self.calcHist:
                h = eng.hist(sid, reset=True)
                for si, name in enumerate(names):
                    self.histdic[name] = self.histdic[name] + h[si]
        Sar = pd.DataFrame(rows, columns=SAR_COLS)
        if len(Sar) and (Sar.DS > 1.05).any():  # detect.py:199-204
            log.warning("DS values above 1 found in sar on %s, removing values above 1", self.sta)
            Sar = Sar[Sar.DS <= 1.05].reset_index(drop=True)
        return Sar, maxds, dense

    # ------------------------------------------------------- time-segment sharding with a halo
    @staticmethod
    def long_array_segments(Ls, ns, seg_lags, halo):
        """Cut the T = Ls - ns + 1 lags of a long array (Ls samples per channel) into segments of
        `seg_lags` core lags plus `halo` lags on either inner side (both multiples of 4).  Returns a list
        of (first lag, number of lags incl. halos, core lo, core hi) -- lo / hi relative to the segment."""
        if seg_lags % 4 or halo % 4 or seg_lags < 4:
            raise ValueError("seg_lags and halo must be multiples of 4")
        T = Ls - ns + 1
        segs = []
        for g0 in range(0, T, seg_lags):
            g1 = min(T, g0 + seg_lags)
            a = max(0, g0 - halo)
            b = min(T, g1 + halo)
            segs.append((a, b - a, g0 - a, g1 - a))
        return segs

    def run_long_array(self, x, sr, start, seg_lags=1 << 18, batch=16, shard=None):
        """Detection on ONE long pre-processed multiplexed array, sharded by time segment with a halo
        ("long-array mode"; BASELINE.json north_star: "Continuous data shards naturally by
        time segment, with a template-length halo").  Every segment is a chunk of `seg_lags` core lags
        whose samples extend ns - 1 further (the template-length halo a lag needs) and whose lags extend
        an LTA window further on both inner sides (so triggers next to a cut see the LTA of the uncut
        array); only core lags are counted, so histograms, MaxDS and the candidate set are those of the
        whole array processed as a single chunk, and the greedy +-20 s pick runs once over the merged
        list.  shard = (rank, world) runs a contiguous range of the segments on this rank and merges the
        candidate lists / histograms / maxima over the ranks (one gather at the end, no data-path
        collective).  Returns (Sar DataFrame, {name: MaxDS}, histograms are added to self.histdic)."""
        if self.estimateMags:
            raise ValueError("run_long_array: magnitude estimates work per chunk; use run_chunks")
        if len(self.groups) != 1:
            raise ValueError("run_long_array: all subspaces must share one basis length")
        n, names = next(iter(self.groups.items()))
        Nc, ns = self.Nc, n // self.Nc
        x = np.asarray(x)
        Ls = len(x) // Nc
        T = Ls - ns + 1
        if T < 10:
            return pd.DataFrame(columns=SAR_COLS), {}
        W = int(self.triggerLTATime * sr)
        Wsta = int(self.triggerSTATime * sr)
        halo = -(-max(W, Wsta, 4) // 4) * 4
        segs = self.long_array_segments(Ls, ns, int(seg_lags), halo)
        lo_s, hi_s = 0, len(segs)
        if shard is not None:
            from . import parallel
            lo_s, hi_s = parallel.shard_range(len(segs), shard[0], shard[1])
        mine = segs[lo_s:hi_s]
        eng, sid, S = self.engine, self.set_ids[n], len(names)
        eng.set_trigger_sta(Wsta)
        cands, mx = [], np.full(S, -np.inf)
        nanrow = np.zeros(S, dtype=bool)
        for b0 in range(0, len(mine), batch):
            part = mine[b0:b0 + batch]
            eng.load_chunks([x[a * Nc:(a + nl + ns - 1) * Nc] for a, nl, _, _ in part])
            eng.set_core_lags([p[2] for p in part], [p[3] for p in part])
            eng.detect_run(sid, engine=self.kernel, kblk=self.kblk, hist_range=(0.0, 1.0),
                           lta_window=0 if self.fillZeros else W)
            m, fl = eng.rowstats()
            c = eng.candidates()
            c = c.astype(LONG_CAND_DTYPE)
            c["t"] += np.array([p[0] for p in part], dtype=np.int64)[c["row"] // S]   # segment lag -> array lag
            c["row"] %= S
            cands.append(c)
            nanrow |= (fl & 1).any(axis=0)
            mx = np.fmax(mx, np.where(np.isnan(m), -np.inf, m).max(axis=0))
        cand = np.concatenate(cands) if cands else np.zeros(0, dtype=LONG_CAND_DTYPE)
        hist = eng.hist(sid, reset=True)
        if shard is not None and shard[1] > 1:
            from . import parallel
            cand = parallel.gather_records(cand)
            hist = parallel.allreduce_sum(hist)
            mx = parallel.allreduce_max(mx)
            nanrow = parallel.allreduce_max(nanrow.astype(np.float64)) > 0
        if self.calcHist:
            for si, name in enumerate(names):
                if not nanrow[si]:                      # np.histogram raises on NaN: row skipped (detect.py:182-185)
                    self.histdic[name] = self.histdic[name] + hist[si]
        maxds = {name: (float("nan") if nanrow[si] else float(mx[si])) for si, name in enumerate(names)}
        rows = []
        cand = cand[np.lexsort((cand["t"], cand["row"]))]
        for si, name in enumerate(names):
            sel = cand[cand["row"] == si]
            if nanrow[si] or not len(sel) or not _evalTrigCon(mx[si], self.threshold[name]):
                continue
            picks = greedy_pick(sel["t"], sel["ds"], T, sr)
            if len(picks) > 4000:  # kill switch, detect.py:433-436
                raise Exception('over 4000 events found in single data block on %s for %s' % (self.sta, name))
            minof, maxof = np.min(self.offsets[name]), np.max(self.offsets[name])
            for k in picks:
                coef = float(sel["ds"][k])
                times = float(sel["t"][k]) / sr + start
                den = float(sel["lta"][k])
                sl = 0.0 if self.fillZeros else (abs(coef) / den if (np.isfinite(den) and den > 0.0) else 0.0)
                rows.append([coef, sl, times, name, self.sta, times - maxof, times - minof, np.nan, np.nan, np.nan])
        Sar = pd.DataFrame(rows, columns=SAR_COLS)
        if len(Sar) and (Sar.DS > 1.05).any():  # detect.py:199-204
            Sar = Sar[Sar.DS <= 1.05].reset_index(drop=True)
        return Sar, maxds

    def getRA(self, chunk, sr, start, File=None):
        """CorDF of one chunk (array part of `_getRA`, detect.py:225-296): index = sorted
        names, columns as the reference.  Returns None if the chunk must be skipped."""
        L = len(chunk) // self.Nc * self.Nc
        nmax = max(self.groups.keys())
        if L <= nmax or (L - nmax) // self.Nc + 1 < 10:
            return None
        eng = self.engine
        eng.load_chunks([chunk])
        CorDF = pd.DataFrame(index=self.names, columns=CORDF_COLS, dtype=object)
        W = int(self.triggerLTATime * sr)
        eng.set_trigger_sta(int(self.triggerSTATime * sr))
        for n, names in sorted(self.groups.items()):
            eng.detect_run(self.set_ids[n], engine=self.kernel, kblk=self.kblk)
            mx, _ = eng.rowstats()
            eng.hist(self.set_ids[n], reset=True)  # getRA does not feed the station histogram
            for si, name in enumerate(names):
                CorDF.at[name, 'SSdetect'] = eng.get_ds(0, si).astype(np.float64)
                CorDF.at[name, 'MaxDS'] = float(mx[0, si])
                CorDF.at[name, 'Nc'] = self.Nc
                CorDF.at[name, 'SampRate'] = sr
                CorDF.at[name, 'TimeStamp'] = start
                CorDF.at[name, 'File'] = File
                if not self.fillZeros:
                    sl = eng.get_stalta(0, si, W).astype(np.float64)
                    CorDF.at[name, 'STALTA'] = sl
                    CorDF.at[name, 'MaxSTALTA'] = np.max(sl)
        return CorDF

    def corDat(self, chunks, sr, starts, batch=16):
        """Chunk loop of `_corDat` (detect.py:157-212): returns (ss_df, histdic)."""
        frames = []
        for i in range(0, len(chunks), batch):
            Sar, _, _ = self.run_chunks(chunks[i:i + batch], sr, starts[i:i + batch])
            if len(Sar) > 300 * max(1, len(chunks[i:i + batch])):
                log.warning("over 300 events found in single data block on %s", self.sta)
            if len(Sar):
                frames.append(Sar)
        DF = pd.concat(frames, ignore_index=True) if frames else pd.DataFrame(columns=SAR_COLS)
        return DF, self.histdic


def _CreateCoeffArray(corSeries, name, threshold, sta, offsets, sr=None, start=None, buff=20):
    """Mirror of `_SSDetex._CreateCoeffArray` (detect.py:390-445) for a CorDF row holding a
    dense detection statistic (trigCon=0, estimateMags=False): greedy picks as a DataFrame
    with the reference's columns.  Pure bookkeeping over an array the GPU produced."""
    DS = np.asarray(corSeries.SSdetect)
    sr = corSeries.SampRate if sr is None else sr
    start = corSeries.TimeStamp if start is None else start
    thr = threshold[name]
    idx = np.nonzero(DS >= thr)[0]
    picks = greedy_pick(idx, DS[idx], len(DS), sr, buff)
    minof, maxof = np.min(offsets[name]), np.max(offsets[name])
    rows = []
    for k in picks:
        i = int(idx[k])
        times = float(i) / sr + start
        try:
            sl = float(corSeries.STALTA[i])
        except TypeError:
            sl = 0.0
        rows.append([float(DS[i]), sl, times, name, sta, times - maxof, times - minof, np.nan, np.nan,
                     np.nan])
    if len(rows) > 4001:
        raise Exception('over 4000 events found in single data block on %s for %s' % (sta, name))
    return pd.DataFrame(rows, columns=SAR_COLS)
