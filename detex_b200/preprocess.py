"""preprocess.py -- host-side mirror of the array part of `construct._applyFilter`
(reference detex/construct.py:990-1030) + `multiplex` (:928-987) on the CUDA engine.

The filter DESIGN (a handful of scalars) is done here exactly the way ObsPy 1.0.2 does it
(obspy/signal/filter.py::bandpass -> scipy.signal.iirfilter + zpk2sos); detrending,
filtering and multiplexing of the samples run on the device (k8_preproc.cu).
Decimation (`Trace.decimate`, a Chebyshev low-pass + stride) is not implemented: the
reference default is decimate=None.
"""
import warnings

import numpy as np
import scipy.signal

from .detect import default_engine


def bandpass_sos(freqmin, freqmax, df, corners=4):
    """Second-order sections of ObsPy's `bandpass` (falls back to a high-pass when freqmax is
    at or above Nyquist, as ObsPy does)."""
    fe = 0.5 * df
    low = freqmin / fe
    high = freqmax / fe
    if high - 1.0 > -1e-6:
        warnings.warn("Selected high corner frequency is at or above Nyquist. Applying a high-pass instead.")
        z, p, k = scipy.signal.iirfilter(corners, low, btype='highpass', ftype='butter', output='zpk')
        return scipy.signal.zpk2sos(z, p, k)
    if low > 1:
        raise ValueError("Selected low corner frequency is above Nyquist.")
    z, p, k = scipy.signal.iirfilter(corners, [low, high], btype='band', ftype='butter', output='zpk')
    return scipy.signal.zpk2sos(z, p, k)


def applyFilter(traces, sr, filt=(1, 10, 2, True), decimate=None, engine=None):
    """Detrend + band-pass + multiplex a batch of chunks on the device.

    traces: list (chunks) of lists (channels in ObsPy's sorted order) of 1-D arrays.
    filt  : [freqmin, freqmax, corners, zerophase] as Detex's `filt` (construct.py:25-38), or
            None for detrend only.
    The multiplexed chunks stay loaded in the engine (ready for detect_run); returns their
    multiplexed lengths."""
    if decimate:
        raise NotImplementedError("decimate is not supported on the device path")
    eng = engine or default_engine()
    if filt is None:
        sos, zp = np.zeros((0, 6)), False
    else:
        sos, zp = bandpass_sos(filt[0], filt[1], sr, corners=filt[2]), bool(filt[3])
    return eng.preprocess_chunks(traces, sos, zerophase=zp, detrend=True)


def applyFilter_multiplex(traces, sr, filt=(1, 10, 2, True), engine=None):
    """As applyFilter, and fetch the multiplexed arrays (`MPcon` of detect.py:241)."""
    eng = engine or default_engine()
    applyFilter(traces, sr, filt, engine=eng)
    return [eng.get_chunk(i) for i in range(len(traces))]
