"""Host-side logic of the package (no GPU): greedy pick on compacted candidates, beta fit
from sufficient statistics, SVD basis selection, shard arithmetic."""
import numpy as np
import scipy.stats

import pytest

from detex_b200 import construct, detect, fas, parallel, subspace, synth
from oracle import detex_oracle as orc


def test_greedy_pick_sparse_equals_dense_reference_loop(trig_golden):
    g = trig_golden
    ds = g["syn_DS"]
    thr = 0.4
    idx = np.nonzero(ds >= thr)[0]
    picks = detect.greedy_pick(idx, ds[idx], len(ds), 100.0)
    got = ds[idx[picks]]
    assert np.array_equal(got, g["syn_trig_DS"])
    assert np.array_equal(idx[picks] / 100.0 + 1.0e9, g["syn_trig_STMP"])


def test_greedy_pick_random_against_oracle():
    rng = np.random.default_rng(5)
    for trial in range(20):
        T = int(rng.integers(3000, 30000))
        ds = rng.uniform(0, 0.3, size=T)
        for _ in range(int(rng.integers(0, 12))):
            ds[int(rng.integers(0, T))] = rng.uniform(0.3, 1.0)
        sr = float(rng.choice([20.0, 40.0, 100.0]))
        rows = orc.greedy_triggers(ds, 0.25, sr, 0.0, [0.0])
        idx = np.nonzero(ds >= 0.25)[0]
        picks = detect.greedy_pick(idx, ds[idx], T, sr)
        assert [int(idx[k]) for k in picks] == [r["index"] for r in rows]


def test_createcoeffarray_columns(trig_golden):
    import pandas as pd
    g = trig_golden
    cs = pd.Series({"SSdetect": g["syn_DS"], "STALTA": g["syn_stalta"], "SampRate": 100.0,
                    "TimeStamp": 1.0e9, "Nc": 1})
    df = detect._CreateCoeffArray(cs, "SS0", {"SS0": 0.4}, "STA", {"SS0": [1.0, 2.5, 4.0]})
    assert list(df.columns) == ['DS', 'DS_STALTA', 'STMP', 'Name', 'Sta', 'MSTAMPmin', 'MSTAMPmax', 'Mag',
                                'SNR', 'ProEnMag']
    assert np.array_equal(df.DS.values, g["syn_trig_DS"])
    assert np.allclose(df.DS_STALTA.values, g["syn_trig_STALTA"], rtol=0, atol=1e-12)
    assert np.array_equal(df.MSTAMPmin.values, g["syn_trig_MSTAMPmin"])


def test_beta_fit_from_stats_matches_scipy():
    rng = np.random.default_rng(1)
    for a0, b0 in ((1.5, 4000.0), (3.0, 2500.0), (8.0, 900.0)):
        x = rng.beta(a0, b0, size=100000)
        st = orc.beta_sufficient_stats(x)
        a, b, loc, scale = fas.beta_fit_from_stats(*st)
        ra, rb, _, _ = scipy.stats.beta.fit(x, floc=0, fscale=1)
        assert abs(a - ra) < 1e-8 * ra and abs(b - rb) < 1e-8 * rb
        nn = fas.beta_nnlf_from_stats(a, b, st[0], st[3], st[4])
        assert abs(nn - scipy.stats.beta.nnlf((ra, rb, 0, 1), x)) < 1e-6 * abs(nn)


def test_svd_basis_matches_oracle():
    rng = np.random.default_rng(2)
    base = synth.wavelet_basis(rng, 200, 3, 3)
    ev = np.array([rng.standard_normal(3) @ base + 0.05 * rng.standard_normal(600) for _ in range(9)])
    a = subspace.svd_basis(ev)
    b = orc.svd_basis(ev)
    assert a["NumBasis"] == b["ndim"] and a["NumBasis"] >= 3
    assert np.allclose(a["U"], b["U"])
    assert np.allclose(a["U"] @ a["U"].T, np.eye(a["NumBasis"]), atol=1e-12)
    th = subspace.threshold_from_fas({"betadist": (2.0, 3000.0, 0, 1)}, Pf=1e-12)
    assert abs(th - orc.threshold_from_beta(2.0, 3000.0, 1e-12)) < 1e-15


def test_shard_range_partitions():
    for n in (0, 1, 7, 720, 5760):
        for w in (1, 2, 3, 8):
            r = [parallel.shard_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[i][1] == r[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1


def test_ccx_row_blocks_balanced():
    for N in (5, 81, 4096):
        for w in (1, 2, 4, 8):
            blocks = parallel.ccx_row_blocks(N, w)
            assert blocks[0][0] == 0 and blocks[-1][1] == N - 1
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(w - 1))
            pairs = [sum(N - 1 - b for b in range(b0, b1)) for b0, b1 in blocks]
            assert sum(pairs) == N * (N - 1) // 2
            if N == 4096:
                assert max(pairs) / (sum(pairs) / w) < 1.01


# ---- N3: alignment from the dendrogram (construct.py:272-286, 486-503, 710-812)

@pytest.mark.parametrize("case", ["fam", "rand2", "rand3", "rand24"])
def test_get_delays_matches_reference_golden(align_golden, case):
    g = align_golden
    link, delays = construct.get_delays(g[case + "_cc"], g[case + "_lag"])
    assert np.array_equal(link, g[case + "_link"])
    assert np.array_equal(delays, g[case + "_delays"])
    aligned, sd = construct.alignTD(delays, g[case + "_X"])
    assert np.array_equal(aligned, g[case + "_aligned"])
    assert sd.min() == 0 and np.array_equal(sd, delays - delays.min())


def _random_cc_lag(rng, N, consistent):
    true = rng.integers(-50, 50, N)
    iu = np.triu_indices(N - 1)
    cc = np.full((N - 1, N - 1), np.nan)
    lag = np.full((N - 1, N - 1), np.nan)
    cc[iu] = rng.permutation(np.linspace(0.1, 0.99, len(iu[0])))
    lag[iu] = (true[iu[1] + 1] - true[iu[0]]) * 3
    if not consistent:
        lag[iu] += 3 * rng.integers(-3, 4, len(iu[0]))
    return cc, lag, true


def test_get_delays_random_against_oracle():
    rng = np.random.default_rng(11)
    for N in (2, 3, 4, 7, 16, 41):
        for consistent in (True, False):
            cc, lag, _ = _random_cc_lag(rng, N, consistent)
            l1, d1 = orc.get_delays(cc, lag)
            l2, d2 = construct.get_delays(cc, lag)
            assert np.array_equal(l1, l2) and np.array_equal(d1, d2)


def test_get_delays_recovers_consistent_shifts_at_scale():
    """With mutually consistent lags the walk must return the true shifts (up to a constant),
    whatever the merge order; N = 1500 runs in about a second (the reference is O(N^3))."""
    rng = np.random.default_rng(12)
    cc, lag, true = _random_cc_lag(rng, 1500, True)
    _, d = construct.get_delays(cc, lag)
    # positive lag(i, j) = event j is delayed relative to i -> j is trimmed by more samples
    assert np.array_equal(d - d.min(), 3 * (true - true.min()))


def test_get_delays_rejects_duplicates_and_alignTD_raises_when_empty():
    cc, lag, _ = _random_cc_lag(np.random.default_rng(13), 5, True)
    cc[0, 1] = cc[0, 0]
    with pytest.raises(ValueError):
        construct.get_delays(cc, lag)
    with pytest.raises(Exception):
        construct.alignTD(np.array([0, 50, 120]), np.zeros((3, 100)))


def test_update_start_times():
    st = [dict(Nc=3, sampling_rate=100.0, starttime=1000.0), dict(Nc=3, sampling_rate=100.0, starttime=2000.0)]
    out = construct.update_start_times(st, [0, 30], [990.0, 1995.0], [1.5, 2.0])
    assert out[0]["starttime"] == 1000.0 and out[0]["offset"] == 10.0
    assert out[1]["starttime"] == 2000.0 + 30 / 300.0 and out[1]["magnitude"] == 2.0
    assert abs(out[1]["offset"] - (5.0 + 0.1)) < 1e-12
    assert st[1]["starttime"] == 2000.0          # inputs untouched


def test_bandpass_design_kat():
    """N2: known answers for the Butterworth band-pass design `preprocess.bandpass_sos` hands to the
    device filter (ObsPy 1.0.2 `bandpass`: iirfilter(corners, [low, high], 'band', 'butter') -> zpk2sos).
    (i) corners = 1 has a closed form: with W1 = tan(pi f1 / fs), W2 = tan(pi f2 / fs), BW = W2 - W1,
    W0^2 = W1 W2, the bilinear transform of BW s / (s^2 + BW s + W0^2) is
        b = [BW, 0, -BW] / D,  a = [1, 2 (W0^2 - 1) / D, (1 - BW + W0^2) / D],  D = 1 + BW + W0^2.
    (ii) a Butterworth band-pass of any order is exactly -3 dB (|H|^2 = 1/2) at both corner frequencies
    and 0 dB at the (pre-warped) geometric centre; zero-phase = two passes squares the response."""
    import scipy.signal
    from detex_b200 import preprocess
    f1, f2, fs = 1.0, 10.0, 100.0
    sos = preprocess.bandpass_sos(f1, f2, fs, corners=1)
    W1, W2 = np.tan(np.pi * f1 / fs), np.tan(np.pi * f2 / fs)
    BW, W02 = W2 - W1, W1 * W2
    D = 1 + BW + W02
    assert sos.shape == (1, 6)
    assert np.allclose(sos[0], [BW / D, 0.0, -BW / D, 1.0, 2 * (W02 - 1) / D, (1 - BW + W02) / D], rtol=0, atol=1e-14)
    for corners, (lo, hi, rate) in ((2, (1.0, 10.0, 100.0)), (4, (2.0, 8.0, 40.0)), (3, (0.5, 4.0, 20.0))):
        sos = preprocess.bandpass_sos(lo, hi, rate, corners=corners)
        assert sos.shape == (corners, 6)
        fc = rate / np.pi * np.arctan(np.sqrt(np.tan(np.pi * lo / rate) * np.tan(np.pi * hi / rate)))
        w, h = scipy.signal.sosfreqz(sos, worN=np.array([lo, hi, fc]) * 2 * np.pi / rate)
        assert np.allclose(np.abs(h[:2]) ** 2, 0.5, atol=1e-12) and abs(abs(h[2]) - 1.0) < 1e-12
    with pytest.raises(ValueError):
        preprocess.bandpass_sos(30.0, 35.0, 40.0, 2)
