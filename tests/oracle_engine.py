"""TEST INFRASTRUCTURE ONLY: an `Engine` look-alike whose numbers come from the CPU oracle.

It lets the `-m "not gpu"` suite drive the host-side mirrors (`detex_b200.detect / fas /
construct / workflow`: batching, bookkeeping, table schemas, greedy picks, error behaviour) on a
box without a GPU, and lets the GPU suite compare the SAME host code run on the real engine
against it.  Nothing under `detex_b200/` imports this file; the product has no CPU path.
"""
import numpy as np
import scipy.signal

from detex_b200._lib import CAND_DTYPE
from detex_b200.engine import ShortChunk
from oracle import detex_oracle as orc


class OracleEngine(object):
    def __init__(self):
        self.sets = {}
        self.chunks = []
        self.sta_window = 0
        self.hist_bins = 400
        self._events = {}
        self.nchunks = 0

    # ---- inputs
    def set_bases(self, set_id, bases, Nc, thresholds=None):
        bases = [np.atleast_2d(np.asarray(b, dtype=np.float64)) for b in bases]
        assert len(set(b.shape[1] for b in bases)) == 1
        self.sets[set_id] = dict(bases=bases, Nc=int(Nc), thr=None if thresholds is None else list(thresholds),
                                 hist=np.zeros((len(bases), 1024), dtype=np.int64),
                                 fas=np.zeros((len(bases), 5)))

    def load_chunks(self, chunks):
        self.chunks = [np.asarray(c, dtype=np.float64) for c in chunks]
        self.nchunks = len(chunks)
        self._core = None

    def set_core_lags(self, lo=None, hi=None):
        self._core = None if lo is None else (list(lo), list(hi))

    def preprocess_chunks(self, traces, sos, zerophase=True, detrend=True, dec_sos=None, factor=1):
        out = []
        for ch in traces:
            ys = []
            common = min(-(-len(x) // factor) for x in ch)   # common window before detrend / filter
            for x in ch:
                y = np.asarray(x, dtype=np.float64)
                if factor > 1:
                    y = scipy.signal.sosfilt(dec_sos, y)[::factor]
                y = y[:common]
                if detrend:
                    y = scipy.signal.detrend(y, type="linear")
                if len(sos):
                    y = scipy.signal.sosfilt(sos, y)
                    if zerophase:
                        y = scipy.signal.sosfilt(sos, y[::-1])[::-1]
                ys.append(y)
            out.append(orc.multiplex(ys))
        self.load_chunks(out)
        return [len(c) for c in out]

    def get_chunk(self, i):
        return self.chunks[i].copy()

    def set_trigger_sta(self, W):
        self.sta_window = int(W)

    def set_hist_bins(self, nbins):
        self.hist_bins = int(nbins)

    # ---- run
    def detect_run(self, set_id, engine="tcgen05", kblk=0, hist_range=(0.0, 1.0), lta_window=0, want_fas=False,
                   keep_ds64=False):
        st = self.sets[set_id]
        Nc, n = st["Nc"], st["bases"][0].shape[1]
        for c in self.chunks:
            L = len(c) // Nc * Nc
            if L <= n or (L - n) // Nc + 1 < 10:
                raise ShortChunk(4, "chunk not longer than the template")
        S = len(st["bases"])
        self._run = dict(set=set_id, S=S)
        self._ds = {}
        mx = np.zeros((len(self.chunks), S), dtype=np.float32)
        fl = np.zeros((len(self.chunks), S), dtype=np.int32)
        cands = []
        bins = np.linspace(hist_range[0], hist_range[1], self.hist_bins + 1)
        for ci, c in enumerate(self.chunks):
            for si, U in enumerate(st["bases"]):
                ds = orc.mpx_ds_direct(c, U, Nc).astype(np.float32)
                # core lags (time-segment sharding): only these count; the LTA reads the whole row
                c_lo, c_hi = (0, len(ds)) if getattr(self, "_core", None) is None else (self._core[0][ci], self._core[1][ci])
                core = ds[c_lo:c_hi]
                m = np.nanmax(core) if not np.isnan(core).any() else np.nan
                if m > 1.1:                                  # detect.py:275-281
                    ds[np.isinf(ds)] = 0
                    m = core.max()
                    fl[ci, si] |= 2
                self._ds[(ci, si)] = ds
                mx[ci, si] = m
                if np.isnan(core).any():
                    fl[ci, si] |= 1
                    continue
                st["hist"][si, :self.hist_bins] += np.histogram(core.astype(np.float64), bins=bins)[0]
                if want_fas:
                    x = core.astype(np.float64)
                    st["fas"][si] += [len(x), x.sum(), (x * x).sum(), np.log(x).sum(), np.log1p(-x).sum()]
                if st["thr"] is not None and m > np.float32(st["thr"][si]):
                    idx = c_lo + np.nonzero(core >= np.float32(st["thr"][si]))[0]
                    lta = np.zeros(len(idx), dtype=np.float32)
                    if lta_window > 0 and len(ds) >= lta_window:
                        sl = orc.sta_lta(ds.astype(np.float64), lta_window, self.sta_window)
                        lta = (np.abs(ds[idx].astype(np.float64)) / sl[idx]).astype(np.float32)
                    elif lta_window > 0:
                        lta[:] = np.nan
                    for t, l in zip(idx, lta):
                        cands.append((ci * S + si, int(t), ds[t], l))
        self._mx, self._fl = mx, fl
        self._cand = np.array(cands, dtype=CAND_DTYPE) if cands else np.zeros(0, dtype=CAND_DTYPE)

    def sync(self):
        pass

    def num_lags(self, chunk):
        st = self.sets[self._run["set"]]
        return (len(self.chunks[chunk]) // st["Nc"] * st["Nc"] - st["bases"][0].shape[1]) // st["Nc"] + 1

    def get_ds(self, chunk, subspace):
        return self._ds[(chunk, subspace)].copy()

    def get_stalta(self, chunk, subspace, W):
        ds = self._ds[(chunk, subspace)].astype(np.float64)
        if len(ds) < W or len(ds) < self.sta_window:
            return np.full(len(ds), np.nan, dtype=np.float32)
        return orc.sta_lta(ds, W, self.sta_window).astype(np.float32)

    def rowstats(self):
        return self._mx.copy(), self._fl.copy()

    def candidates(self, cap=1 << 20):
        return self._cand.copy()

    def hist(self, set_id, reset=False):
        h = self.sets[set_id]["hist"][:, :self.hist_bins].copy()
        if reset:
            self.sets[set_id]["hist"][:] = 0
        return h

    def fas(self, set_id, reset=False):
        f = self.sets[set_id]["fas"].copy()
        if reset:
            self.sets[set_id]["fas"][:] = 0
        return f

    # ---- FAS screen, magnitudes
    def sta_lta_max(self, Nc, chan, nsta, nlta):
        out = []
        for c in self.chunks:
            z = c[chan::Nc]
            out.append(np.max(orc.classic_sta_lta(z, nsta, nlta)))
        return np.array(out, dtype=np.float32)

    def set_events(self, set_id, subspace, ewf, mags, wfu_var=None, is_single=False):
        self._events[(set_id, subspace)] = (np.atleast_2d(ewf), np.asarray(mags, dtype=np.float64), is_single)

    def est_mags(self, set_id, chunk, subspace, t):
        st = self.sets[set_id]
        out = []
        for ci, si, ti in zip(np.atleast_1d(chunk), np.atleast_1d(subspace), np.atleast_1d(t)):
            ewf, mags, single = self._events[(set_id, int(si))]
            out.append(orc.est_mag(int(ti), self.chunks[int(ci)], st["Nc"], st["bases"][int(si)], ewf, mags,
                                   issubspace=not single))
        return np.array(out, dtype=np.float64)

    # ---- CCX
    def corr_zero_lag(self, X):
        X = np.asarray(X, dtype=np.float64)
        N = len(X)
        out = np.eye(N)
        for i in range(N):
            for j in range(i + 1, N):
                out[i, j] = out[j, i] = float(np.max(orc.fast_normcorr(X[i], X[j])))
        return out

    def ccx(self, X, Nc, row_begin=0, row_end=None, engine="fp64"):
        X = np.asarray(X, dtype=np.float64)
        N = len(X)
        row_end = N if row_end is None else row_end
        cc = np.zeros((row_end - row_begin, N))
        lag = np.zeros((row_end - row_begin, N), dtype=np.int32)
        sub = np.zeros((row_end - row_begin, N))
        for b in range(row_begin, row_end):
            for c in range(b + 1, N):
                cc[b - row_begin, c], lag[b - row_begin, c], sub[b - row_begin, c] = orc.ccx2(X[b], X[c], Nc)
        return cc, lag, sub

    # the device-resident CCX calls of parallel.ccx_sharded; "device" addresses are CPU tensors here
    @staticmethod
    def _view(ptr, shape, dtype):
        import ctypes
        n = int(np.prod(shape)) * np.dtype(dtype).itemsize
        return np.frombuffer((ctypes.c_char * n).from_address(int(ptr)), dtype=dtype).reshape(shape)

    def ccx_device(self, X, Nc, rows, d_cc, d_lag, d_sub, engine="tcgen05"):
        X = np.asarray(X, dtype=np.float64)
        N = len(X)
        cc, lag, sub = (self._view(d_cc, (len(rows), N), np.float64), self._view(d_lag, (len(rows), N), np.int32),
                        self._view(d_sub, (len(rows), N), np.float64))
        for r, b in enumerate(rows):
            for c in range(int(b) + 1, N):
                cc[r, c], lag[r, c], sub[r, c] = orc.ccx2(X[int(b)], X[c], Nc)

    def ccx_pack(self, d_cc, d_lag, d_sub, slot_rows, N, out=None):
        from detex_b200 import parallel
        ns = len(slot_rows)
        return parallel.pack_condensed(self._view(d_cc, (ns, N), np.float64), self._view(d_lag, (ns, N), np.int32),
                                       self._view(d_sub, (ns, N), np.float64), slot_rows, N)

    def ccx_pack_rows(self, d_cc, d_lag, d_sub, rows, N, out):
        cc, lag, sub = (self._view(d_cc, (len(rows), N), np.float64), self._view(d_lag, (len(rows), N), np.int32),
                        self._view(d_sub, (len(rows), N), np.float64))
        for r, b in enumerate(rows):
            b = int(b)
            o = b * N - b * (b + 1) // 2
            out[0][o:o + N - 1 - b] = cc[r, b + 1:]
            out[1][o:o + N - 1 - b] = lag[r, b + 1:]
            out[2][o:o + N - 1 - b] = sub[r, b + 1:]

    def ccx_condensed(self, X, Nc, engine="tcgen05", out=None):
        cc, lag, sub = self.ccx(X, Nc, 0, len(X) - 1)
        iu = np.triu_indices(len(X), 1)
        return cc[iu], lag[iu], sub[iu]

    def close(self):
        pass
