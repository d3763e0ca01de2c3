"""engine.py -- thin object wrapper over the C ABI (one context = one GPU = one stream).

Everything numerical happens in the CUDA library; this class only marshals NumPy arrays
to pointers and maps status codes to exceptions the way the reference maps problems to
`detex.log(level='error')` (raise) or to a skipped chunk (ShortChunk).
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import CAND_DTYPE, DtxError, ENGINE_FP64, ENGINE_TCGEN05, HIST_BINS


class ShortChunk(DtxError):
    """Chunk not longer than the template / fewer than 10 lags (reference: skip the chunk,
    detex/detect.py:262-274)."""


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class Engine(object):
    def __init__(self, device=0, stream=None):
        self._L = _lib.load()
        h = C.c_void_p()
        rc = self._L.dtx_create(int(device), C.c_void_p(stream) if stream else None, C.byref(h))
        if rc != 0:
            raise DtxError(rc, "dtx_create failed (needs an sm_100 GPU; there is no CPU fallback)")
        self._h = h
        self.device = device
        self._S = {}
        self._keep = None
        self.nchunks = 0

    def close(self):
        if getattr(self, "_h", None):
            self._L.dtx_destroy(self._h)
            self._h = None
            for ptr in getattr(self, "_pinned", []):
                self._L.dtx_host_free(ptr)
            self._pinned = []

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc == 0:
            return
        msg = self._L.dtx_last_error(self._h).decode()
        if rc == _lib.DTX_ERR_SHORT_CHUNK:
            raise ShortChunk(rc, msg)
        raise DtxError(rc, msg)

    # ------------------------------------------------------------------ bases
    def set_bases(self, set_id, bases, Nc, thresholds=None):
        """bases: list of (r_i, n) float64 arrays (rows = basis vectors, multiplexed order)."""
        bases = [np.atleast_2d(np.asarray(b, dtype=np.float64)) for b in bases]
        n = bases[0].shape[1]
        for b in bases:
            if b.shape[1] != n:
                raise DtxError(2, "all bases of a set must share the same length n")
        U = np.ascontiguousarray(np.vstack(bases))
        off = np.zeros(len(bases) + 1, dtype=np.int32)
        off[1:] = np.cumsum([b.shape[0] for b in bases])
        thr = None
        if thresholds is not None:
            thr = np.ascontiguousarray(np.asarray(thresholds, dtype=np.float64))
            assert thr.shape == (len(bases),)
        self._check(self._L.dtx_set_bases(self._h, int(set_id), _ptr(U), _ptr(off), len(bases), int(n),
                                          int(Nc), _ptr(thr) if thr is not None else None))
        self._S[set_id] = len(bases)

    # ----------------------------------------------------------------- chunks
    def load_chunks(self, chunks):
        """chunks: list of 1-D multiplexed arrays (float64 or float32), host memory."""
        dt = np.float32 if all(np.asarray(c).dtype == np.float32 for c in chunks) else np.float64
        arrs = [np.ascontiguousarray(np.asarray(c, dtype=dt)) for c in chunks]
        ptrs = (C.c_void_p * len(arrs))(*[a.ctypes.data for a in arrs])
        L = np.array([a.shape[0] for a in arrs], dtype=np.int64)
        self._check(self._L.dtx_load_chunks(self._h, len(arrs), ptrs, _ptr(L),
                                            _lib.DTX_F32 if dt == np.float32 else _lib.DTX_F64))
        self._keep = arrs  # async H2D: keep the host arrays alive until the next sync
        self.nchunks = len(arrs)

    def preprocess_chunks(self, traces, sos, zerophase=True, detrend=True, dec_sos=None, factor=1):
        """traces: list (chunks) of lists (channels, ObsPy sort order) of 1-D host arrays.  Optional
        decimation (forward low-pass `dec_sos`, every factor-th sample), detrend, SOS filter and
        multiplex on the device; the result becomes the loaded batch."""
        Nc = len(traces[0])
        f32 = all(np.asarray(t).dtype == np.float32 for ch in traces for t in ch)
        dt = np.float32 if f32 else np.float64
        arrs = [np.ascontiguousarray(np.asarray(t, dtype=dt)) for ch in traces for t in ch]
        if any(len(ch) != Nc for ch in traces):
            raise DtxError(2, "every chunk needs the same number of channels")
        ptrs = (C.c_void_p * len(arrs))(*[a.ctypes.data for a in arrs])
        lens = np.array([a.shape[0] for a in arrs], dtype=np.int64)
        sos = np.ascontiguousarray(np.atleast_2d(np.asarray(sos, dtype=np.float64)))
        factor = int(factor)
        if factor > 1:
            dsos = np.ascontiguousarray(np.atleast_2d(np.asarray(dec_sos, dtype=np.float64)))
            self._check(self._L.dtx_preprocess_chunks_dec(
                self._h, len(traces), Nc, ptrs, _ptr(lens), _lib.DTX_F32 if f32 else _lib.DTX_F64, _ptr(sos),
                sos.shape[0], int(bool(zerophase)), int(bool(detrend)), _ptr(dsos), dsos.shape[0], factor))
        else:
            self._check(self._L.dtx_preprocess_chunks(self._h, len(traces), Nc, ptrs, _ptr(lens),
                                                      _lib.DTX_F32 if f32 else _lib.DTX_F64, _ptr(sos), sos.shape[0],
                                                      int(bool(zerophase)), int(bool(detrend))))
        self.nchunks = len(traces)
        return [int(min(-(-len(t) // factor) for t in ch)) * Nc for ch in traces]

    def get_chunk(self, chunk):
        L = C.c_int64()
        probe = np.empty(1, dtype=np.float64)
        rc = self._L.dtx_get_chunk(self._h, int(chunk), _ptr(probe), 0, C.byref(L))
        if rc not in (0, _lib.DTX_ERR_CAPACITY):
            self._check(rc)
        out = np.empty(L.value, dtype=np.float64)
        self._check(self._L.dtx_get_chunk(self._h, int(chunk), _ptr(out), out.size, C.byref(L)))
        return out

    def attach_device_chunks(self, base_ptr, elem_offsets, lengths, f32=False):
        off = np.ascontiguousarray(np.asarray(elem_offsets, dtype=np.int64))
        L = np.ascontiguousarray(np.asarray(lengths, dtype=np.int64))
        self._check(self._L.dtx_attach_device_chunks(self._h, len(L), C.c_void_p(int(base_ptr)), _ptr(off),
                                                     _ptr(L), _lib.DTX_F32 if f32 else _lib.DTX_F64))
        self.nchunks = len(L)

    def set_core_lags(self, lo=None, hi=None):
        """Core lag range [lo[i], hi[i]) of every loaded chunk (time-segment sharding with a halo): only
        these lags count for MaxDS / histograms / candidates / FAS sums of later runs.  None = all lags."""
        if lo is None:
            self._check(self._L.dtx_set_core_lags(self._h, None, None))
            return
        lo = np.ascontiguousarray(np.asarray(lo, dtype=np.int64))
        hi = np.ascontiguousarray(np.asarray(hi, dtype=np.int64))
        assert len(lo) == len(hi) == self.nchunks
        self._check(self._L.dtx_set_core_lags(self._h, _ptr(lo), _ptr(hi)))

    # -------------------------------------------------------------------- run
    def detect_run(self, set_id, engine="tcgen05", kblk=0, hist_range=(0.0, 1.0), lta_window=0,
                   want_fas=False, keep_ds64=False):
        if engine not in _lib.ENGINES:
            raise ValueError("engine must be one of %s" % sorted(_lib.ENGINES))
        eng = _lib.ENGINES[engine]
        self._check(self._L.dtx_detect_run(self._h, int(set_id), eng, int(kblk), float(hist_range[0]),
                                           float(hist_range[1]), int(lta_window), int(bool(want_fas)),
                                           int(bool(keep_ds64))))
        self._run_S = self._S[set_id]
        if getattr(self, "_accumulating", False):
            self._acc_total += self.nchunks

    def sync(self):
        self._check(self._L.dtx_sync(self._h))

    def num_lags(self, chunk):
        T = C.c_int64()
        self._check(self._L.dtx_num_lags(self._h, int(chunk), C.byref(T)))
        return T.value

    def get_ds(self, chunk, subspace):
        out = np.empty(self.num_lags(chunk), dtype=np.float32)
        self._check(self._L.dtx_get_ds(self._h, int(chunk), int(subspace), _ptr(out), out.size))
        return out

    def get_ds64(self, chunk, subspace):
        out = np.empty(self.num_lags(chunk), dtype=np.float64)
        self._check(self._L.dtx_get_ds64(self._h, int(chunk), int(subspace), _ptr(out), out.size))
        return out

    def get_stalta(self, chunk, subspace, W):
        out = np.empty(self.num_lags(chunk), dtype=np.float32)
        self._check(self._L.dtx_get_stalta(self._h, int(chunk), int(subspace), int(W), _ptr(out), out.size))
        return out

    def set_trigger_sta(self, sta_window):
        """triggerSTATime in samples (detect.py:282-288): 0 = the reference default STA = |DS|."""
        self._check(self._L.dtx_set_trigger_sta(self._h, int(sta_window)))

    def set_fused(self, on=True):
        """Fused mode: later detect_run calls never write the dense statistic (no get_ds / get_stalta);
        maxima, histograms, candidates and FAS sums come out of the projection kernel itself."""
        self._check(self._L.dtx_set_fused(self._h, int(bool(on))))

    def set_x8_tolerance(self, eps):
        """Adaptive engine ("tcgen05_auto"): admitted rms error of a normalised projection."""
        self._check(self._L.dtx_set_x8_tolerance(self._h, float(eps)))

    def chunk_modes(self):
        """Per chunk of the last run: 1 = ran with 8-bit cross terms."""
        m = np.empty(self.nchunks, dtype=np.int32)
        self._check(self._L.dtx_get_chunk_modes(self._h, _ptr(m)))
        return m

    def rowstats(self):
        nch = self._acc_total if getattr(self, "_accumulating", False) else self.nchunks
        n = nch * self._run_S
        mx = np.empty(n, dtype=np.float32)
        fl = np.empty(n, dtype=np.int32)
        self._check(self._L.dtx_get_rowstats(self._h, _ptr(mx), _ptr(fl), n))
        return mx.reshape(nch, self._run_S), fl.reshape(nch, self._run_S)

    def set_hist_bins(self, nbins):
        """Bins of the device histograms for later runs (numBins - 1 of fas._initFAS; default 400)."""
        self._check(self._L.dtx_set_hist_bins(self._h, int(nbins)))
        self._hist_bins = int(nbins)

    def hist(self, set_id, reset=False):
        S = self._S[set_id]
        nb = getattr(self, "_hist_bins", HIST_BINS)
        h = np.empty(S * nb, dtype=np.uint64)
        self._check(self._L.dtx_get_hist(self._h, int(set_id), _ptr(h), h.size, int(reset)))
        return h.reshape(S, nb).astype(np.int64)

    def fas(self, set_id, reset=False):
        S = self._S[set_id]
        f = np.empty(S * 5, dtype=np.float64)
        self._check(self._L.dtx_get_fas(self._h, int(set_id), _ptr(f), f.size, int(reset)))
        return f.reshape(S, 5)

    def candidates(self, cap=1 << 20):
        n = C.c_int64()
        rc = self._L.dtx_get_candidates(self._h, None, 0, C.byref(n))      # count only
        if rc == _lib.DTX_ERR_CAPACITY or n.value > cap:
            # mirrors the reference's kill switch for runaway trigger counts (detect.py:433-436)
            raise DtxError(_lib.DTX_ERR_CAPACITY, "more than %d candidate lags above threshold" % cap)
        self._check(rc)
        out = np.empty(n.value, dtype=CAND_DTYPE)
        if n.value:
            self._check(self._L.dtx_get_candidates(self._h, _ptr(out), n.value, C.byref(n)))
        return out

    def accumulate_begin(self, total_chunks):
        """Results of the following detect_run calls (the batches of one station) accumulate on the
        device; candidates() / rowstats() then return everything since this call, with rows numbered
        by the chunk's position in the whole sequence.  No host synchronisation between batches."""
        self._check(self._L.dtx_accumulate_begin(self._h, int(total_chunks)))
        self._acc_total = 0
        self._accumulating = True

    def accumulate_end(self):
        self._check(self._L.dtx_accumulate_end(self._h))
        self._accumulating = False

    def k1_ms(self):
        ms = C.c_float()
        self._check(self._L.dtx_last_k1_ms(self._h, C.byref(ms)))
        return ms.value

    def k1_ms_history(self):
        """K1 durations (ms) of every run since accumulate_begin / the last call; waits for them."""
        n = C.c_int64()
        self._check(self._L.dtx_k1_ms_history(self._h, None, 0, C.byref(n)))
        out = np.empty(n.value, dtype=np.float32)
        if n.value:
            self._check(self._L.dtx_k1_ms_history(self._h, _ptr(out), out.size, C.byref(n)))
        return out

    def sta_lta_max(self, Nc, chan, nsta, nlta):
        """max of ObsPy's classic STA/LTA of channel `chan` for every loaded chunk (fas.py:175-205)."""
        out = np.empty(self.nchunks, dtype=np.float32)
        self._check(self._L.dtx_sta_lta_max(self._h, int(Nc), int(chan), int(nsta), int(nlta), _ptr(out), out.size))
        return out

    def set_events(self, set_id, subspace, ewf, mags, wfu_var=None, is_single=False):
        ewf = np.ascontiguousarray(np.atleast_2d(np.asarray(ewf, dtype=np.float64)))
        mags = np.ascontiguousarray(np.asarray(mags, dtype=np.float64))
        wv = None if wfu_var is None else np.ascontiguousarray(np.asarray(wfu_var, dtype=np.float64))
        self._check(self._L.dtx_set_events(self._h, int(set_id), int(subspace), ewf.shape[0], _ptr(ewf), _ptr(mags),
                                           _ptr(wv) if wv is not None else None, int(bool(is_single))))

    def est_mags(self, set_id, chunk, subspace, t):
        """(ntrig, 3) array of ProEnMag, Mag, SNR for triggers of the current batch (_estMag)."""
        chunk = np.ascontiguousarray(np.asarray(chunk, dtype=np.int32))
        subspace = np.ascontiguousarray(np.asarray(subspace, dtype=np.int32))
        t = np.ascontiguousarray(np.asarray(t, dtype=np.int32))
        out = np.empty((len(t), 3), dtype=np.float64)
        self._check(self._L.dtx_est_mags(self._h, int(set_id), len(t), _ptr(chunk), _ptr(subspace), _ptr(t), _ptr(out)))
        return out

    def launch_count(self):
        n = C.c_int64()
        self._check(self._L.dtx_launch_count(self._h, C.byref(n)))
        return n.value

    def corr_zero_lag(self, X):
        X = np.ascontiguousarray(np.asarray(X, dtype=np.float64))
        N, n = X.shape
        out = np.empty((N, N), dtype=np.float64)
        self._check(self._L.dtx_corr_zero_lag(self._h, _ptr(X), N, n, _ptr(out)))
        return out

    # -------------------------------------------------------------------- ccx
    def ccx(self, X, Nc, row_begin=0, row_end=None, engine="fp64"):
        """Dense rows [row_begin, row_end) of the pair matrix: cc, lag, subsamp as (rows, N) arrays
        (entries with c <= b are zero)."""
        X = np.asarray(X)
        dt = np.float32 if X.dtype == np.float32 else np.float64
        X = np.ascontiguousarray(X, dtype=dt)
        N, n = X.shape
        row_end = N if row_end is None else row_end
        rows = row_end - row_begin
        cc = np.zeros((rows, N), dtype=np.float64)
        lag = np.zeros((rows, N), dtype=np.int32)
        sub = np.zeros((rows, N), dtype=np.float64)
        eng = ENGINE_TCGEN05 if engine == "tcgen05" else ENGINE_FP64
        self._check(self._L.dtx_ccx(self._h, _ptr(X), _lib.DTX_F32 if dt == np.float32 else _lib.DTX_F64,
                                    N, n, int(Nc), int(row_begin), int(row_end), eng, _ptr(cc), _ptr(lag),
                                    _ptr(sub)))
        return cc, lag, sub

    def set_ccx_batch(self, max_signals=512, ds_bytes=16 << 30):
        """Limits of one tensor-core CCX batch (tests lower them to force several batches)."""
        self._check(self._L.dtx_set_ccx_batch(self._h, int(max_signals), int(ds_bytes)))

    def set_ccx_passes(self, passes=1):
        """MMAs per K step of the tensor-core CCX series: 1 = hi*hi screening (default), 3 = fp16x3."""
        self._check(self._L.dtx_set_ccx_passes(self._h, int(passes)))

    def pinned_empty(self, shape, dtype):
        """NumPy array in page-locked host memory (staging buffer of the end-to-end paths); it stays
        valid until the engine is closed."""
        dtype = np.dtype(dtype)
        nbytes = int(np.prod(shape)) * dtype.itemsize
        ptr = C.c_void_p()
        if self._L.dtx_host_alloc(C.byref(ptr), max(1, nbytes)) != 0:
            raise DtxError(1, "dtx_host_alloc failed")
        self._pinned = getattr(self, "_pinned", [])
        self._pinned.append(ptr)
        buf = (C.c_char * max(1, nbytes)).from_address(ptr.value)
        return np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    def host_register(self, arr):
        """Page-lock (and map for the device) the memory of a NumPy array the caller owns, e.g. a view of a
        POSIX shared-memory segment; `host_unregister` before the memory goes away."""
        if self._L.dtx_host_register(_ptr(arr), int(arr.nbytes)) != 0:
            raise DtxError(1, "dtx_host_register failed")

    def host_unregister(self, arr):
        self._L.dtx_host_unregister(_ptr(arr))

    def ccx_pack_rows(self, d_cc, d_lag, d_sub, rows, N, out):
        """This GPU's dense device slots (slot r = event rows[r]) -> their places in the page-locked condensed
        host arrays `out` = (cc, lag, subsamp) of the whole matrix (`pinned_empty` / `host_register`)."""
        rows = np.ascontiguousarray(np.asarray(rows, dtype=np.int32))
        cc, lag, sub = out
        self._check(self._L.dtx_ccx_pack_rows(self._h, C.c_void_p(int(d_cc)), C.c_void_p(int(d_lag)),
                                              C.c_void_p(int(d_sub)), _ptr(rows), len(rows), int(N), _ptr(cc), _ptr(lag),
                                              _ptr(sub)))

    def ccx_condensed(self, X, Nc, engine="tcgen05", out=None):
        """All pairs b < c in SciPy's condensed order: (cc float64, lag int32, subsamp float64), each
        N (N-1)/2 long.  `out` = three preallocated (e.g. pinned) arrays to fill."""
        X = np.asarray(X)
        dt = np.float32 if X.dtype == np.float32 else np.float64
        X = np.ascontiguousarray(X, dtype=dt)
        N, n = X.shape
        npair = N * (N - 1) // 2
        if out is None:
            out = (np.empty(npair, np.float64), np.empty(npair, np.int32), np.empty(npair, np.float64))
        cc, lag, sub = out
        assert cc.size == npair and lag.size == npair and sub.size == npair
        eng = ENGINE_TCGEN05 if engine == "tcgen05" else ENGINE_FP64
        self._check(self._L.dtx_ccx_condensed(self._h, _ptr(X), _lib.DTX_F32 if dt == np.float32 else _lib.DTX_F64,
                                              N, n, int(Nc), eng, _ptr(cc), _ptr(lag), _ptr(sub)))
        return cc, lag, sub

    def ccx_device(self, X, Nc, rows, d_cc, d_lag, d_sub, engine="tcgen05", x_device_ptr=None, shape=None):
        """Template rows `rows` (ascending event indices) of the pair matrix, results left in the
        caller's device arrays (addresses; dense [len(rows)][N] float64 / int32 / float64).  X is a
        host array, or (x_device_ptr, shape) for waveforms already in HBM."""
        rows = np.ascontiguousarray(np.asarray(rows, dtype=np.int32))
        eng = ENGINE_TCGEN05 if engine == "tcgen05" else ENGINE_FP64
        if x_device_ptr is not None:
            N, n = shape
            dt = np.dtype(X) if X is not None else np.dtype(np.float64)
            xp, ondev = C.c_void_p(int(x_device_ptr)), 1
        else:
            X = np.asarray(X)
            dt = np.dtype(np.float32) if X.dtype == np.float32 else np.dtype(np.float64)
            X = np.ascontiguousarray(X, dtype=dt)
            N, n = X.shape
            xp, ondev = _ptr(X), 0
        self._check(self._L.dtx_ccx_device(self._h, xp, ondev, _lib.DTX_F32 if dt == np.float32 else _lib.DTX_F64,
                                           N, n, int(Nc), _ptr(rows), len(rows), eng, C.c_void_p(int(d_cc)),
                                           C.c_void_p(int(d_lag)), C.c_void_p(int(d_sub))))
        self._keep = X

    def ccx_pack(self, d_cc, d_lag, d_sub, slot_rows, N, out=None):
        """Dense device slots (slot s = event slot_rows[s], -1 = padding) -> condensed host arrays."""
        slot_rows = np.ascontiguousarray(np.asarray(slot_rows, dtype=np.int32))
        npair = N * (N - 1) // 2
        if out is None:
            out = (np.empty(npair, np.float64), np.empty(npair, np.int32), np.empty(npair, np.float64))
        cc, lag, sub = out
        self._check(self._L.dtx_ccx_pack(self._h, C.c_void_p(int(d_cc)), C.c_void_p(int(d_lag)), C.c_void_p(int(d_sub)),
                                         _ptr(slot_rows), len(slot_rows), int(N), _ptr(cc), _ptr(lag), _ptr(sub)))
        return cc, lag, sub
