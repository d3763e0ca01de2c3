"""Oracle (oracle/detex_oracle.py) against the golden vectors generated from the unmodified
reference (tests/golden/make_golden.py).  CPU only; this is the oracle's pin."""
import os

import numpy as np
import pytest

from oracle import detex_oracle as orc

TOL = 1e-12


@pytest.mark.parametrize("case", ["nc3", "nc1", "nc2odd"])
@pytest.mark.parametrize("flavour", ["fft", "direct"])
def test_mpx_ds_matches_reference(ds_golden, case, flavour):
    g = ds_golden
    Nc = int(g[case + "_Nc"])
    fn = orc.mpx_ds_fft if flavour == "fft" else orc.mpx_ds_direct
    for ci in range(int(g[case + "_nchunks"])):
        for si in range(int(g[case + "_nbases"])):
            ds = fn(g["%s_chunk%d" % (case, ci)], g["%s_U%d" % (case, si)], Nc)
            ref = g["%s_DS_%d_%d" % (case, ci, si)]
            assert ds.shape == ref.shape
            # DC-shifted chunk: the reference's own FFT round-off is ~1e-10 there
            tol = 1e-9 if (case == "nc3" and ci == 1) else TOL
            assert np.abs(ds - ref).max() < tol


def test_singleton_template_not_demeaned(ds_golden):
    g = ds_golden
    for fn in (orc.mpx_ds_fft, orc.mpx_ds_direct):
        ds = fn(g["single_chunk"], g["single_U"], 3)
        assert np.abs(ds - g["single_DS"]).max() < TOL
    assert g["single_DS"].max() > 0.5  # the template was cut from the chunk itself


def test_n_minus_one_over_n_factor(ds_golden):
    """The (n-1)/n factor from rolling_var's ddof=1 is real."""
    g = ds_golden
    x, U = g["single_chunk"], g["single_U"]
    n = U.shape[1]
    ds = orc.mpx_ds_direct(x, U, 3)
    t = int(np.argmax(ds))
    w = x[t * 3:t * 3 + n]
    wc = w - w.mean()
    plain = np.square(U @ wc).sum() / np.square(wc).sum()
    assert abs(ds[t] - plain * (n - 1.0) / n) < 1e-12
    assert abs(ds[t] - plain) > 1e-5


@pytest.mark.parametrize("lta,sta", [(50.0, 0), (500.0, 0), (50.0, 7.0)])
def test_sta_lta(ds_golden, trig_golden, lta, sta):
    ds = ds_golden["nc3_DS_0_1"]
    out = orc.sta_lta(ds, lta, sta)
    assert np.abs(out - trig_golden["stalta_%g_%g" % (lta, sta)]).max() < 1e-10


def test_greedy_triggers(trig_golden):
    g = trig_golden
    rows = orc.greedy_triggers(g["syn_DS"], 0.4, 100.0, 1.0e9, [1.0, 2.5, 4.0], stalta=g["syn_stalta"])
    assert len(rows) == len(g["syn_trig_DS"])
    assert np.array_equal([r["DS"] for r in rows], g["syn_trig_DS"])
    assert np.array_equal([r["STMP"] for r in rows], g["syn_trig_STMP"])
    assert np.allclose([r["DS_STALTA"] for r in rows], g["syn_trig_STALTA"], rtol=0, atol=1e-12)
    assert np.array_equal([r["MSTAMPmin"] for r in rows], g["syn_trig_MSTAMPmin"])
    assert np.array_equal([r["MSTAMPmax"] for r in rows], g["syn_trig_MSTAMPmax"])


def test_greedy_is_not_local_max():
    """A point suppressed by a larger neighbour does not suppress its own neighbours
    (peaks .9@25 s, .8@32.5 s, .7@47.5 s -> picks at 25 s and 47.5 s)."""
    ds = np.zeros(10000)
    ds[2500], ds[3250], ds[4750] = .9, .8, .7
    rows = orc.greedy_triggers(ds, 0.5, 100.0, 0.0, [0.0])
    assert [r["index"] for r in rows] == [2500, 4750]


@pytest.mark.parametrize("case", ["nc3", "nc1"])
@pytest.mark.parametrize("fft", [True, False])
def test_ccx_matches_reference(ccx_golden, case, fft):
    g = ccx_golden
    cc, lag, sub = orc.make_cclags(g[case + "_X"], int(g[case + "_Nc"]), fft=fft)
    m = ~np.isnan(g[case + "_cc"])
    assert np.array_equal(np.isnan(cc), ~m)
    assert np.abs(cc[m] - g[case + "_cc"][m]).max() < 1e-12
    assert np.array_equal(lag[m], g[case + "_lag"][m])
    assert np.nanmax(np.abs(sub[m] - g[case + "_sub"][m])) < 1e-9


def test_ccx_edge_cases(ccx_golden):
    g = ccx_golden
    cc, lag = g["nc3_cc"], g["nc3_lag"]
    # identical pair (events 2 and 7): cc = 1 at lag 0
    assert abs(cc[2, 6] - 1.0) < 1e-12 and lag[2, 6] == 0
    # all-zero event 5: reference returns (0, 0, 0)
    assert cc[5, 5] == 0.0 and cc[0, 4] == 0.0 and lag[0, 4] == 0.0


def test_multiplex():
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "multiplex_golden.npz"))
    assert np.array_equal(orc.multiplex([g["c0"], g["c1"], g["c2"]]), g["out"])


def test_fas_stats_consistent():
    rng = np.random.default_rng(3)
    dss = [rng.beta(3.0, 2000.0, size=5000) for _ in range(4)]
    f = orc.fas_stats(dss)
    assert f["hist"].sum() == 20000
    a, b = f["betadist"][:2]
    assert abs(a - 3.0) < 0.2 and abs(b - 2000.0) < 150
    th = orc.threshold_from_beta(a, b, Pf=1e-12)
    assert 0 < th < 0.1


def test_est_mag_matches_reference():
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "mag_golden.npz"))
    x, Nc, U, ewf, mags = g["x"], int(g["Nc"]), g["U"], g["ewf"], g["mags"]
    for t, ref in zip(g["trigs"], g["sub_out"]):
        assert np.allclose(orc.est_mag(int(t), x, Nc, U, ewf, mags, True), ref, rtol=0, atol=1e-12)
    single = g["single"]
    us = (single / np.linalg.norm(single))[None, :]
    for t, ref in zip(g["trigs"], g["single_out"]):
        assert np.allclose(orc.est_mag(int(t), x, Nc, us, single[None, :], np.array([1.7]), False), ref, rtol=0, atol=1e-12)
    out = orc.est_mag(1500, x, Nc, U, ewf, np.full(7, -99.0), True)
    assert np.isnan(out[0]) and np.isnan(out[1]) and abs(out[2] - g["nomag_out"][2]) < 1e-12


@pytest.mark.parametrize("case", ["fam", "rand2", "rand3", "rand24"])
def test_alignment_delays_match_reference(align_golden, case):
    """`_getDelays` / `_traceEventDendro` / `_alignTD` (construct.py:486-503, 710-812): integer
    work, exact."""
    g = align_golden
    link, delays = orc.get_delays(g[case + "_cc"], g[case + "_lag"])
    assert np.array_equal(link, g[case + "_link"])
    assert np.array_equal(delays, g[case + "_delays"])
    assert np.array_equal(orc.align_td(delays, g[case + "_X"]), g[case + "_aligned"])


def _same_basis(U, ref):
    """Singular vectors are defined up to sign: rows agree up to a factor of +-1."""
    assert U.shape == ref.shape
    for a, b in zip(U, ref):
        assert min(np.abs(a - b).max(), np.abs(a + b).max()) < 1e-10


def test_svd_selection_matches_reference():
    """a16: SubSpace.SVD's per-subspace body run by the REFERENCE (`_trimGroups`, scipy.linalg.svd,
    `_getFracEnergy`, `_getUsedBasis`, subspace.py:875-1013; golden from make_golden.svd_case) against
    the oracle restatement and the product's `subspace.svd_basis`, every selectCriteria."""
    from detex_b200 import subspace
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "svd_golden.npz"))
    W = g["svd_X"][:, int(g["svd_start"]):int(g["svd_end"])]              # the trimmed aligned waveforms
    for k, (crit, val, nm) in enumerate(g["svd_cases"]):
        crit, nm = int(crit), bool(nm)
        val = int(val) if crit == 4 else float(val)
        o = orc.svd_basis(W, select_criteria=crit, select_value=val, normalize=nm)
        p = subspace.svd_basis(W, selectCriteria=crit, selectValue=val, normalize=nm)
        ref_U, ref_s = g["svd%d_U" % k], g["svd%d_s" % k]
        assert o["ndim"] == p["NumBasis"] == len(ref_U) == len(g["svd%d_used_keys" % k])
        assert np.allclose(o["s"], ref_s, rtol=1e-12) and np.allclose(p["s"], ref_s, rtol=1e-12)
        assert np.array_equal(g["svd%d_used_keys" % k], ref_s[:len(ref_U)])   # UsedSVDKeys = top singular values
        for got_avg, got_min in ((o["frac_avg"], o["frac_min"]), (p["FracEnergy"]["Average"], p["FracEnergy"]["Minimum"])):
            assert np.abs(got_avg - g["svd%d_frac_avg" % k]).max() < 1e-12
            assert np.abs(got_min - g["svd%d_frac_min" % k]).max() < 1e-12
        _same_basis(o["U"], ref_U)
        _same_basis(p["U"], ref_U)
