/* detex_b200.h -- C ABI of the B200-native Detex hot path.
 *
 * Plain C, plain pointers and sizes, int status codes; no exceptions and no torch types
 * cross this boundary.  The library is CUDA-only (sm_100a): dtx_create() fails on any
 * other device and there is no CPU fallback.
 *
 * The reference (d-chambers/Detex, pure Python) has no FFI; the entry points below
 * replace its private hot-path callables.  Each declaration cites the reference
 * interface it stands in for (file:line under the reference tree).  INTEGRATION.md shows
 * the ctypes binding a Detex maintainer would add.
 *
 * Threading: one host thread per context; all work of a context is issued on one CUDA
 * stream (the caller's, if given to dtx_create).  Calls are asynchronous until a
 * dtx_get_* / dtx_sync call.
 */
#ifndef DETEX_B200_H
#define DETEX_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct dtx_ctx dtx_ctx;

enum {
    DTX_OK = 0,
    DTX_ERR_CUDA = 1,        /* a CUDA call failed; see dtx_last_error */
    DTX_ERR_ARG = 2,         /* invalid argument / shape (reference: detex.log(level='error') raises) */
    DTX_ERR_DEVICE = 3,      /* not an sm_100 device */
    DTX_ERR_SHORT_CHUNK = 4, /* chunk shorter than the template or < 10 lags (detect.py:262-274: skip) */
    DTX_ERR_STATE = 5,       /* call order violated (no chunks loaded, unknown basis set, ...) */
    DTX_ERR_CAPACITY = 6     /* caller-provided output too small */
};

enum { DTX_F64 = 0, DTX_F32 = 1 };

/* engine selection for dtx_detect_run / dtx_ccx_run */
enum {
    DTX_ENGINE_TCGEN05 = 0, /* production: Hankel-tiled tcgen05 GEMM, split fp16 x3 */
    DTX_ENGINE_FP64 = 1,    /* validation: float64 CUDA-core closed form */
    DTX_ENGINE_TCGEN05_X8 = 2, /* detection only, opt-in: hi*hi in fp16, both cross terms in ONE
                                  e4m3 x e5m2 MMA (2 MMAs per K step instead of 3).  Its error is
                                  statistical, not worst-case bounded */
    DTX_ENGINE_TCGEN05_AUTO = 3 /* detection only, opt-in: per chunk, 8-bit cross terms where the
                                   fourth-moment error model admits them (dtx_set_x8_tolerance),
                                   fp16 cross terms everywhere else */
};

/* One candidate trigger: a lag whose statistic is >= the subspace threshold
 * (the set `while Ceval.max() >= threshold` iterates over, detect.py:410). */
typedef struct dtx_cand {
    int32_t row;  /* chunk * S + subspace */
    int32_t t;    /* lag index (sample of the chunk, channel-aligned) */
    float ds;     /* detection statistic */
    float lta;    /* denominator of DS_STALTA: |ds| / lta = STA / LTA at t (`_getStaLtaArray`,
                     detect.py:501-524).  With the default STA = |DS| it is the centred rolling mean of
                     |DS| over the LTA window; with dtx_set_trigger_sta it is LTA * |ds| / STA */
} dtx_cand;

int dtx_version(void);

/* Context ---------------------------------------------------------------------------- */
/* `stream` is a cudaStream_t (or NULL for a private stream). */
int dtx_create(int device, void* stream, dtx_ctx** out);
void dtx_destroy(dtx_ctx* ctx);
const char* dtx_last_error(const dtx_ctx* ctx);
int dtx_sync(dtx_ctx* ctx);

/* Basis upload -------------------------------------------------------------------------
 * Replaces _SSDetex._loadMPSubSpace (detect.py:319-388) and fas._loadMPSubSpace /
 * _loadMPSingles (fas.py:137-172): U holds the rows `row.SVD[k] for k in row.UsedSVDKeys`
 * of S subspaces (or the unit-norm singleton templates) back to back, each of length n
 * multiplexed samples (n % Nc == 0); subspace s owns rows rank_off[s]..rank_off[s+1]-1
 * (any rank >= 1; ranks above 16 are split into 16-vector pieces whose DS contributions
 * are accumulated).  thresholds[s] (may be NULL) is row.Threshold.  The library copies. */
int dtx_set_bases(dtx_ctx* ctx, int set_id, const double* U, const int32_t* rank_off, int S, int n,
                  int Nc, const double* thresholds);

/* Continuous data ----------------------------------------------------------------------
 * Chunks are the multiplexed arrays `MPcon` produced by construct.multiplex
 * (construct.py:928-987) inside _getRA (detect.py:241) / _getDSVect (fas.py:110).
 * dtx_load_chunks copies host arrays to the device (the H2D leg of the end-to-end path);
 * dtx_attach_device_chunks uses arrays already resident in HBM. L[i] is the multiplexed
 * length; lengths are trimmed to a multiple of Nc at run time. */
int dtx_load_chunks(dtx_ctx* ctx, int nchunks, const void* const* host_ptrs, const int64_t* L,
                    int dtype);
int dtx_attach_device_chunks(dtx_ctx* ctx, int nchunks, const void* dev_base,
                             const int64_t* elem_offsets, const int64_t* L, int dtype);

/* Time-segment sharding with a halo ("long-array mode") -------------------------------
 * A long pre-processed array can be cut into segments that overlap: each segment carries, besides the
 * template-length tail every chunk needs, a halo of extra lags on either side so that the LTA windows
 * of triggers near a cut see the same samples as in the uncut array.  lo[i] / hi[i] (lags of chunk i,
 * lo % 4 == 0) delimit the CORE of each loaded chunk: MaxDS, histograms, candidates and FAS sums of
 * later runs only count lags lo <= t < hi, while candidate lags stay relative to the chunk start and the
 * LTA / STA windows read the halo.  With cores that partition the lags of the long array the results are
 * those of the uncut array.  NULL pointers (or loading new chunks) restore "every lag counts".  The
 * reference analogue is the conBuff overlap of consecutive chunks (getdata.py:509-518), whose duplicate
 * detections results._deleteDetDups (results.py:393-397) removes afterwards. */
int dtx_set_core_lags(dtx_ctx* ctx, const int64_t* lo, const int64_t* hi);

/* On-device pre-processing (next row N2) ----------------------------------------------------
 * Replaces the array part of construct._applyFilter (construct.py:1017-1029) + multiplex
 * (construct.py:928-987): every trace is linearly detrended (scipy.signal.detrend, which is what
 * ObsPy 1.0.2's Trace.detrend('linear') calls), filtered with the given second-order sections
 * exactly as obspy/signal/filter.py::bandpass does (sosfilt from a zero state; zerophase = a
 * second pass over the reversed trace), channels are trimmed to the shortest one and
 * interleaved.  The multiplexed float64 chunks become the context's loaded chunks (as after
 * dtx_load_chunks).  chan_ptrs[chunk*Nc + c] / chan_len[...] are host arrays; sos is
 * [nsos][6] = b0 b1 b2 a0 a1 a2 (scipy layout), designed by the caller. */
int dtx_preprocess_chunks(dtx_ctx* ctx, int nchunks, int Nc, const void* const* chan_ptrs,
                          const int64_t* chan_len, int dtype, const double* sos, int nsos, int zerophase,
                          int detrend);
/* As above with `st.decimate(factor)` first (construct.py:1014-1015; ObsPy 1.0.2 Trace.decimate:
 * forward-only low-pass `dec_sos` [ndec][6] -- lowpass_cheby_2 designed by the caller -- then
 * data[::factor]); detrend and band-pass then act on the decimated traces.  factor = 1 skips it. */
int dtx_preprocess_chunks_dec(dtx_ctx* ctx, int nchunks, int Nc, const void* const* chan_ptrs,
                              const int64_t* chan_len, int dtype, const double* sos, int nsos, int zerophase,
                              int detrend, const double* dec_sos, int ndec, int factor);
/* the multiplexed chunk as the device holds it (MPcon of detect.py:241), float64 chunks only */
int dtx_get_chunk(dtx_ctx* ctx, int chunk, double* out, int64_t count, int64_t* L);

/* Detection statistic ------------------------------------------------------------------
 * Replaces the `for ind,row in CorDF.iterrows(): _MPXDS(...)` loop of _SSDetex._getRA
 * (detect.py:259-281) and fas._MPXSSCorr (fas.py:120-134) for every loaded chunk and every
 * subspace of the set, then the per-row reductions of _corDat (detect.py:177-190):
 * MaxDS with inf zeroing, np.histogram on linspace(hist_lo, hist_hi, 401) accumulated
 * per subspace, candidate compaction against the thresholds, LTA of |DS| over
 * `lta_window` samples at the candidates, and (want_fas) the beta-fit sufficient
 * statistics of fas._initFAS (fas.py:74-84).
 * kblk: 64-tap chunks accumulated in TMEM between drains (0 = default 3). */
int dtx_detect_run(dtx_ctx* ctx, int set_id, int engine, int kblk, double hist_lo, double hist_hi,
                   int lta_window, int want_fas, int keep_ds64);

/* Accumulating the batches of one station ------------------------------------------------
 * _corDat (detect.py:154-212) walks the chunks of a station one after the other and only looks at
 * the results at the end of each chunk; nothing there needs the host between chunks.  Between
 * dtx_accumulate_begin(total_chunks) and dtx_accumulate_end, every dtx_detect_run APPENDS: candidate
 * rows and the rowmax / rowflags entries are numbered by the chunk's position in the whole sequence
 * (row = chunk_index * S + subspace, chunk_index counted over all runs since begin), histograms and
 * FAS sums keep accumulating as always, and no call synchronises the stream until the caller
 * fetches (dtx_get_candidates / dtx_get_rowstats return everything since begin).  All runs must use
 * the same basis set; total_chunks bounds the sum of the batch sizes.  dtx_get_ds / dtx_get_stalta
 * / dtx_est_mags keep addressing the chunks of the LAST run by their index inside that run. */
int dtx_accumulate_begin(dtx_ctx* ctx, int64_t total_chunks);
int dtx_accumulate_end(dtx_ctx* ctx);

/* Fused mode ---------------------------------------------------------------------------------
 * on != 0: later dtx_detect_run calls with a tensor-core engine (and keep_ds64 == 0, ranks <= 16) never
 * write the dense detection statistic: MaxDS, flags, histograms, candidates and FAS sums come out of
 * the projection kernel's own read-out (the reductions of _corDat detect.py:177-190 on values still in
 * registers), and the LTA windows of the few candidates are re-evaluated from the float64 closed form.
 * Same MaxDS / histograms / candidate set, bit for bit; the candidates' `lta` differs from the
 * unfused run's by float rounding (~1e-6 relative).  No DS buffer (18 GB per 48-chunk batch of
 * configs[3]) is needed, so a whole station fits one batch; dtx_get_ds / dtx_get_stalta report
 * DTX_ERR_STATE after a fused run. */
int dtx_set_fused(dtx_ctx* ctx, int on);

/* DTX_ENGINE_TCGEN05_AUTO: admitted rms error of a normalised projection (u.w)/(|u||w|) under the
 * random-rounding model (default 2e-6: the 8-bit terms then add 2 sqrt(DS) eps <= 4e-6 per model
 * standard deviation; measured maxima are 0.2-0.6 of that, profiles/r01_x8_engine.md);
 * 0 turns the 8-bit mode off for every chunk. */
int dtx_set_x8_tolerance(dtx_ctx* ctx, double eps);

/* Results (each call synchronises the stream) ------------------------------------------ */
/* modes[chunk] of the last dtx_detect_run: 1 = the chunk ran with 8-bit cross terms */
int dtx_get_chunk_modes(dtx_ctx* ctx, int32_t* modes);
int dtx_num_lags(dtx_ctx* ctx, int chunk, int64_t* T);
/* dense DS row (CorDF.SSdetect[name], detect.py:268) */
int dtx_get_ds(dtx_ctx* ctx, int chunk, int subspace, float* out, int64_t count);
int dtx_get_ds64(dtx_ctx* ctx, int chunk, int subspace, double* out, int64_t count);
/* triggerSTATime of _SSDetex (detect.py:38, 282-288) in samples: 0 (reference default) makes the
 * short-term average |DS| itself; > 0 a centred rolling mean of |DS| with the same edge rule as
 * the LTA.  Applies to the candidates of later dtx_detect_run calls and to dtx_get_stalta. */
int dtx_set_trigger_sta(dtx_ctx* ctx, int sta_window);
/* dense STA / LTA (CorDF.STALTA, _getStaLtaArray detect.py:501-515), W = LTA window in samples;
 * filled with NaN when the row is shorter than a window */
int dtx_get_stalta(dtx_ctx* ctx, int chunk, int subspace, int W, float* out, int64_t count);
/* maxds[chunk*S+s] = CorDF.MaxDS (detect.py:275-281); flags bit0 = row has NaN, bit1 = infs zeroed */
int dtx_get_rowstats(dtx_ctx* ctx, float* maxds, int32_t* flags, int64_t count);
/* Number of histogram bins of later dtx_detect_run calls = numBins - 1 of fas._initFAS (fas.py:31,79;
 * `np.histogram(dss, bins=np.linspace(lo, hi, numBins))`), 1..1024.  Default 400 (detect.py:80 fixes
 * 401 edges for detection).  Reset a set's histogram (dtx_get_hist with reset) before changing it. */
int dtx_set_hist_bins(dtx_ctx* ctx, int nbins);
/* hist[s*nbins+b] accumulated over every run since the last reset (histdic, detect.py:146,181);
 * reset before switching the histogram range of a set (detection [0,1] vs FAS [-.01,1]) */
int dtx_get_hist(dtx_ctx* ctx, int set_id, uint64_t* hist, int64_t count, int reset);
/* fas[s*5 + {0:N, 1:sum x, 2:sum x^2, 3:sum log x, 4:sum log1p(-x)}] */
int dtx_get_fas(dtx_ctx* ctx, int set_id, double* fas, int64_t count, int reset);
/* candidates of the last run (or of every run since dtx_accumulate_begin); *n receives the number
 * produced (may exceed cap: truncated, DTX_ERR_CAPACITY).  out == NULL only queries the count. */
int dtx_get_candidates(dtx_ctx* ctx, dtx_cand* out, int64_t cap, int64_t* n);

/* STA/LTA screen of the loaded chunks -----------------------------------------------------
 * Replaces fas._checkSTALTA (fas.py:175-205): out[i] = max over the chunk of ObsPy's
 * classic_sta_lta (obspy 1.0.2, obspy/signal/src/stalta.c) of channel `chan` (0-based position
 * in the multiplexed order; Detex screens the Z component) with nsta / nlta samples.  A chunk
 * passes the FAS screen when out[i] <= staltalimit. */
int dtx_sta_lta_max(dtx_ctx* ctx, int Nc, int chan, int nsta, int nlta, float* out, int64_t count);

/* Per-detection magnitude / SNR (next row N1) ------------------------------------------------
 * Replaces _SSDetex._estMag with _estPEMag / _estSTDMag (detect.py:447-499, 637-664) and
 * construct.fast_normcorr (construct.py:469-483).  dtx_set_events uploads, per subspace, what
 * _loadMPSubSpace keeps for it (detect.py:346-381): the trimmed aligned event waveforms
 * ewf[nev][n], their magnitudes, and var(WFU_i) of the events projected into the subspace
 * (singles: nev = 1, ewf = the trimmed template, wfu_var ignored).  dtx_est_mags evaluates the
 * triggers (chunk, subspace, lag) of the current batch; out[i] = {ProEnMag, Mag, SNR}. */
int dtx_set_events(dtx_ctx* ctx, int set_id, int subspace, int nev, const double* ewf, const double* mags,
                   const double* wfu_var, int is_single);
int dtx_est_mags(dtx_ctx* ctx, int set_id, int ntrig, const int32_t* chunk, const int32_t* subspace,
                 const int32_t* t, double* out);

/* Timing aid for bench.py: device milliseconds of the dominant kernel (K1) in the last
 * dtx_detect_run, measured with CUDA events on the context's stream. */
int dtx_last_k1_ms(dtx_ctx* ctx, float* ms);
/* The K1 durations of every run since dtx_accumulate_begin (or the last call of this function), in
 * run order; ms == NULL only queries the count.  Waits for those runs to finish. */
int dtx_k1_ms_history(dtx_ctx* ctx, float* ms, int64_t cap, int64_t* n);
/* Number of CUDA kernels this context has launched since creation (bench.py's gpu_launches). */
int dtx_launch_count(dtx_ctx* ctx, int64_t* n);

/* Pairwise CCX ---------------------------------------------------------------------------
 * Replaces construct._makeDFcclags / _CCX2 / _subSamp (construct.py:369-466) for the N
 * multiplexed event waveforms X[N][n] of one station: for every pair b < c in
 * [row_begin, row_end) x (b, N) the maximum normalised cross-correlation, its lag in
 * multiplexed samples and the cosine-fit sub-sample shift.  Outputs are dense row-major
 * [(row_end-row_begin)][N] arrays (entries with c <= b untouched). */
int dtx_ccx(dtx_ctx* ctx, const void* X, int dtype, int N, int n, int Nc, int row_begin, int row_end,
            int engine, double* cc, int32_t* lag, double* subsamp);
/* The same for an arbitrary ascending list of template rows b (the share of one GPU when the
 * matrix is dealt over several) with the results LEFT ON THE DEVICE: d_cc / d_lag /
 * d_sub are caller-owned device arrays, dense [nrows][N], slot r holding event rows[r] (entries
 * with c <= rows[r] are zero).  X is a host array, or a device array if x_on_device != 0.  The
 * call is asynchronous on the context's stream except for one small read-back (the list of
 * degenerate pairs the float64 kernel re-does).  No device memory is allocated once the context's
 * CCX workspace has reached the problem size. */
int dtx_ccx_device(dtx_ctx* ctx, const void* X, int x_on_device, int dtype, int N, int n, int Nc,
                   const int32_t* rows, int nrows, int engine, double* d_cc, int32_t* d_lag, double* d_sub);
/* Dense device slots (e.g. the all-gathered shares of several GPUs, [nslots][N], slot s holding event
 * slot_rows[s], -1 = padding; every b in [0, N-1) must appear) -> host arrays in SciPy's condensed
 * order: pair (b, c), b < c, at index b*N - b*(b+1)/2 + (c-b-1), N*(N-1)/2 entries each.  That is
 * the order `_flatNoNan(1.0000001 - DFcc)` feeds to linkage (construct.py:152-156). */
int dtx_ccx_pack(dtx_ctx* ctx, const double* d_cc, const int32_t* d_lag, const double* d_sub,
                 const int32_t* slot_rows, int nslots, int N, double* cc, int32_t* lag, double* subsamp);
/* (multi-GPU, no reference counterpart) The template rows ONE GPU computed with dtx_ccx_device (dense device slots
 * r = 0 .. nrows-1 holding events rows[r]) written straight to their places in the host's condensed arrays of the
 * whole matrix, over that GPU's own PCIe link.  cc / lag / subsamp must be page-locked memory the device can
 * address (dtx_host_alloc, or dtx_host_register of POSIX shared memory mapped by every rank of a box): the 8 GPUs
 * then fill one matrix in parallel -- createCluster's per-station CCX (construct.py:139-157) without funnelling
 * N(N-1)/2 x 20 bytes through one link.  Synchronises the context's stream before returning. */
int dtx_ccx_pack_rows(dtx_ctx* ctx, const double* d_cc, const int32_t* d_lag, const double* d_sub,
                      const int32_t* rows, int nrows, int N, double* cc, int32_t* lag, double* subsamp);
/* dtx_ccx_device over all rows + dtx_ccx_pack: host X in, condensed host cc / lag / subsamp out. */
int dtx_ccx_condensed(dtx_ctx* ctx, const void* X, int dtype, int N, int n, int Nc, int engine, double* cc,
                      int32_t* lag, double* subsamp);
/* MMAs per K step of the tensor-core CCX series: 1 (default) = fp16 hi*hi only -- the float32 series only
 * LOCATES the maximum, every lag within a rigorous error band of it is re-scored in float64, so cc / lag /
 * subsamp are unchanged; 3 = the fp16x3 series of the detection path. */
int dtx_set_ccx_passes(dtx_ctx* ctx, int passes);
/* Limits of one tensor-core CCX batch: at most max_signals padded events per K1 launch and at most
 * ds_bytes of correlation series (defaults 512 and 16 GiB -- 4096 events then take 8 launches; same-box 4 / 8 / 16 GiB:
 * 89.2 / 85.6 / 83.2 ms; tests lower them to force several batches). */
int dtx_set_ccx_batch(dtx_ctx* ctx, int max_signals, int64_t ds_bytes);

/* Page-locked host memory for the staging buffers of the end-to-end paths (H2D / D2H at PCIe rate
 * instead of through the driver's bounce buffer). */
int dtx_host_alloc(void** out, int64_t bytes);
int dtx_host_free(void* p);
/* Page-lock (and map for the device) host memory the caller owns, e.g. a shared-memory segment several ranks have
 * mapped; undo with dtx_host_unregister before the memory is released. */
int dtx_host_register(void* p, int64_t bytes);
int dtx_host_unregister(void* p);

/* Zero-lag Pearson matrix of N equal-length waveforms (next row N3, validateClusters,
 * subspace.py:738-773: fast_normcorr of every pair of aligned, trimmed cluster members).
 * out is the dense symmetric [N][N] matrix. */
int dtx_corr_zero_lag(dtx_ctx* ctx, const double* X, int N, int n, double* out);

#ifdef __cplusplus
}
#endif
#endif /* DETEX_B200_H */
