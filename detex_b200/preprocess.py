"""preprocess.py -- host-side mirror of the array part of `construct._applyFilter`
(reference detex/construct.py:990-1030) + `multiplex` (:928-987) on the CUDA engine.

The filter DESIGN (a handful of scalars) is done here exactly the way ObsPy 1.0.2 does it
(obspy/signal/filter.py::bandpass -> scipy.signal.iirfilter + zpk2sos); detrending,
filtering and multiplexing of the samples run on the device (k8_preproc.cu).
Decimation (`st.decimate(factor)`, construct.py:1014-1015) follows ObsPy 1.0.2's Trace.decimate:
`lowpass_cheby_2` at 0.5 * sr / factor (order <= 12, 96 dB stop band, forward only) and every
factor-th sample, BEFORE detrend and band-pass; the band-pass is designed for the new rate.
"""
import warnings

import numpy as np
import scipy.signal

from .detect import default_engine


def bandpass_sos(freqmin, freqmax, df, corners=4):
    """Second-order sections of ObsPy's `bandpass` (falls back to a high-pass when freqmax is
    at or above Nyquist, as ObsPy does).

    ObsPy itself cannot be installed here, so this is pinned on the text of the function it restates
    -- obspy 1.0.2 (the version `environment.txt` of the reference pins), obspy/signal/filter.py::bandpass:

        fe = 0.5 * df
        low = freqmin / fe
        high = freqmax / fe
        # raise for some bad scenarios
        if high - 1.0 > -1e-6:
            ... warnings.warn(msg)
            return highpass(data, freq=freqmin, df=df, corners=corners, zerophase=zerophase)
        if low > 1:
            raise ValueError("Selected low corner frequency is above Nyquist.")
        z, p, k = iirfilter(corners, [low, high], btype='band', ftype='butter', output='zpk')
        sos = zpk2sos(z, p, k)
        if zerophase:
            firstpass = sosfilt(sos, data)
            return sosfilt(sos, firstpass[::-1])[::-1]
        else:
            return sosfilt(sos, data)

    and on known answers that do not come from SciPy (tests/test_host_logic.py::test_bandpass_design_kat):
    the closed-form bilinear-transform coefficients of the first-order band-pass and the -3 dB points
    of every order sitting exactly on freqmin / freqmax."""
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


def lowpass_cheby2_sos(freq, df, maxorder=12):
    """Second-order sections of ObsPy's `lowpass_cheby_2` (obspy/signal/filter.py): the pass-band
    edge is lowered in 1 % steps until cheb2ord needs no more than `maxorder` poles."""
    nyquist = df * 0.5
    rp, rs, order = 1, 96, 1e99
    ws = freq / nyquist
    wp = ws
    if ws > 1:
        ws = 1.0
        warnings.warn("Selected corner frequency is above Nyquist. Setting Nyquist as high corner.")
    while True:
        if order <= maxorder:
            break
        wp = wp * 0.99
        order, wn = scipy.signal.cheb2ord(wp, ws, rp, rs, analog=0)
    z, p, k = scipy.signal.cheby2(order, rs, wn, btype='low', analog=0, output='zpk')
    return scipy.signal.zpk2sos(z, p, k)


def applyFilter(traces, sr, filt=(1, 10, 2, True), decimate=None, engine=None):
    """Decimate + detrend + band-pass + multiplex a batch of chunks on the device.

    traces: list (chunks) of lists (channels in ObsPy's sorted order) of 1-D arrays.
    filt  : [freqmin, freqmax, corners, zerophase] as Detex's `filt` (construct.py:25-38), or
            None for detrend only.
    decimate : None or int factor (<= 16, ObsPy's limit for its automatic filter design).
    The multiplexed chunks stay loaded in the engine (ready for detect_run); returns their
    multiplexed lengths."""
    eng = engine or default_engine()
    factor, dec_sos = 1, None
    if decimate:
        factor = int(decimate)
        if factor > 16:
            raise ArithmeticError("Automatic filter design is unstable for decimation factors above 16. "
                                  "Manual decimation is necessary.")
        if factor > 1:
            dec_sos = lowpass_cheby2_sos(sr * 0.5 / float(factor), sr, maxorder=12)
    sr_out = sr / float(factor)
    if filt is None:
        sos, zp = np.zeros((0, 6)), False
    else:
        sos, zp = bandpass_sos(filt[0], filt[1], sr_out, corners=filt[2]), bool(filt[3])
    return eng.preprocess_chunks(traces, sos, zerophase=zp, detrend=True, dec_sos=dec_sos, factor=factor)


def applyFilter_multiplex(traces, sr, filt=(1, 10, 2, True), decimate=None, engine=None):
    """As applyFilter, and fetch the multiplexed arrays (`MPcon` of detect.py:241)."""
    eng = engine or default_engine()
    applyFilter(traces, sr, filt, decimate=decimate, engine=eng)
    return [eng.get_chunk(i) for i in range(len(traces))]
