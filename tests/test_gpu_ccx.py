"""Parity of the CUDA CCX path against the golden vectors / oracle: CC within 1e-5
(float64 engine: 1e-10), lags bit-exact."""
import numpy as np
import pytest

from detex_b200 import construct, synth
from oracle import detex_oracle as orc

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", ["nc3", "nc1"])
def test_ccx_matches_reference_golden(engine, ccx_golden, case):
    g = ccx_golden
    X, Nc = g[case + "_X"], int(g[case + "_Nc"])
    cc, lag, sub = engine.ccx(X, Nc)
    N = X.shape[0]
    iu = np.triu_indices(N, 1)
    rcc, rlag, rsub = g[case + "_cc"][iu[0], iu[1] - 1], g[case + "_lag"][iu[0], iu[1] - 1], g[case + "_sub"][iu[0], iu[1] - 1]
    assert np.abs(cc[iu] - rcc).max() < 1e-10
    assert np.array_equal(lag[iu].astype(float), rlag)
    assert np.nanmax(np.abs(sub[iu] - rsub)) < 1e-7


def test_ccx_row_blocks_equal_full(engine, ccx_golden):
    X = ccx_golden["nc3_X"]
    cc, lag, sub = engine.ccx(X, 3)
    for b0, b1 in ((0, 4), (4, 9), (9, 11)):
        c2, l2, s2 = engine.ccx(X, 3, row_begin=b0, row_end=b1)
        for r in range(b0, b1):
            assert np.array_equal(c2[r - b0, r + 1:], cc[r, r + 1:])
            assert np.array_equal(l2[r - b0, r + 1:], lag[r, r + 1:])


def test_makeDFcclags_frames(engine, ccx_golden):
    import pandas as pd
    g = ccx_golden
    X = g["nc3_X"]
    evs = ["ev%02d" % i for i in range(len(X))]
    row = pd.Series({"MPtd": {e: X[i] for i, e in enumerate(evs)}, "MPfd": {e: None for e in evs},
                     "Channels": {e: ["E", "N", "Z"] for e in evs}})
    DFcc, DFlag, DFsub = construct._makeDFcclags(evs, row, engine=engine)
    assert list(DFcc.columns) == list(range(1, len(X))) and list(DFcc.index) == list(range(len(X) - 1))
    a = DFcc.values.astype(float)
    assert np.array_equal(np.isnan(a), np.isnan(g["nc3_cc"]))
    m = ~np.isnan(a)
    assert np.abs(a[m] - g["nc3_cc"][m]).max() < 1e-10
    assert np.array_equal(DFlag.values.astype(float)[m], g["nc3_lag"][m])
    link = construct.cluster_link(DFcc)
    assert np.allclose(link, orc.cluster_link(g["nc3_cc"]), atol=1e-9)
    cc1 = construct._CCX2(None, None, X[0], X[1], "ENZ", "ENZ", engine=engine)
    assert abs(cc1[0] - g["nc3_cc"][0, 0]) < 1e-10 and cc1[1] == g["nc3_lag"][0, 0]
    with pytest.raises(Exception):
        construct._CCX2(None, None, X[0], X[1][:-3], "ENZ", "ENZ", engine=engine)


def test_ccx_config3_shape_subset(engine):
    """config 3 waveform shape (3 ch x 10 s x 100 Hz, n = 3000, 1001 lags), 24 events."""
    X = synth.event_families(3003, 4, 6, 1000, 3, max_shift=100)
    cc, lag, sub = engine.ccx(X, 3)
    rcc, rlag, rsub = orc.make_cclags(X, 3, fft=True)
    iu = np.triu_indices(len(X), 1)
    assert np.abs(cc[iu] - rcc[iu[0], iu[1] - 1]).max() < 1e-10
    assert np.array_equal(lag[iu].astype(float), rlag[iu[0], iu[1] - 1])
    assert cc[0, 1] > 0.4 and abs(lag[0, 1]) <= 600 and lag[0, 1] % 3 == 0


@pytest.mark.parametrize("case", ["nc3", "nc1"])
def test_ccx_tensor_engine_matches_reference_golden(engine, ccx_golden, case):
    """tcgen05 engine: float32 correlation series from the Hankel GEMM, arg-max neighbourhood
    re-scored in float64 -> same accuracy as the float64 engine, lags bit-exact."""
    g = ccx_golden
    X, Nc = g[case + "_X"], int(g[case + "_Nc"])
    cc, lag, sub = engine.ccx(X, Nc, engine="tcgen05")
    N = X.shape[0]
    iu = np.triu_indices(N, 1)
    rcc, rlag, rsub = g[case + "_cc"][iu[0], iu[1] - 1], g[case + "_lag"][iu[0], iu[1] - 1], g[case + "_sub"][iu[0], iu[1] - 1]
    assert np.abs(cc[iu] - rcc).max() < 1e-10
    assert np.array_equal(lag[iu].astype(float), rlag)
    assert np.nanmax(np.abs(sub[iu] - rsub)) < 1e-7


def test_ccx_tensor_engine_config3_shape_and_row_blocks(engine):
    X = synth.event_families(3003, 6, 8, 1000, 3, max_shift=100)      # 48 events, n = 3000, 1001 lags
    cc, lag, sub = engine.ccx(X, 3, engine="tcgen05")
    c64, l64, s64 = engine.ccx(X, 3, engine="fp64")
    iu = np.triu_indices(len(X), 1)
    assert np.abs(cc[iu] - c64[iu]).max() < 1e-12
    assert np.array_equal(lag[iu], l64[iu])
    assert np.nanmax(np.abs(sub[iu] - s64[iu])) < 1e-9
    for b0, b1 in ((0, 17), (17, 40), (40, 47)):
        c2, l2, s2 = engine.ccx(X, 3, row_begin=b0, row_end=b1, engine="tcgen05")
        for r in range(b0, b1):
            assert np.array_equal(c2[r - b0, r + 1:], cc[r, r + 1:])
            assert np.array_equal(l2[r - b0, r + 1:], lag[r, r + 1:])


def test_validate_cluster_zero_lag_cc(engine):
    """N3 (part): validateClusters' pairwise zero-lag fast_normcorr on the GPU."""
    from detex_b200 import subspace
    rng = np.random.default_rng(71)
    base = synth.multiplex(synth.bandpassed_noise(rng, 400, nchan=3))
    W = np.array([base + s * synth.multiplex(synth.bandpassed_noise(rng, 400, nchan=3)) for s in (0.2, 0.3, 3.0, 0.25, 4.0)])
    cc = engine.corr_zero_lag(W)
    for i in range(5):
        for j in range(5):
            assert abs(cc[i, j] - orc.fast_normcorr(W[i], W[j])[0]) < 1e-12
    bad = subspace.validate_cluster(W, 0.5, engine=engine)
    exp = [i for i in range(4) if max(orc.fast_normcorr(W[i], W[j])[0] for j in range(i + 1, 5)) < 0.5]
    assert bad == exp and 2 in bad


def test_pack_rows_fills_one_host_matrix_from_dealt_rows(engine):
    """The multi-GPU result path on one GPU: the template rows dealt to 3 "ranks" are computed one rank at a time with
    the results left in HBM (dtx_ccx_device) and written straight into ONE page-locked condensed host matrix
    (dtx_ccx_pack_rows); the matrix equals dtx_ccx_condensed's and the oracle's."""
    import torch
    from detex_b200 import parallel
    from detex_b200.engine import DtxError
    X = synth.event_families(515, 3, 7, 100, 3, max_shift=12)          # 21 events, n = 300
    N = len(X)
    npair = N * (N - 1) // 2
    out = (engine.pinned_empty((npair,), np.float64), engine.pinned_empty((npair,), np.int32),
           engine.pinned_empty((npair,), np.float64))
    out[0][:] = np.nan
    out[1][:] = -99999
    for rows in parallel.ccx_deal_rows(N, 3):
        d_cc = torch.zeros((len(rows), N), dtype=torch.float64, device="cuda")
        d_lag = torch.zeros((len(rows), N), dtype=torch.int32, device="cuda")
        d_sub = torch.zeros((len(rows), N), dtype=torch.float64, device="cuda")
        engine.ccx_device(X, 3, rows, d_cc.data_ptr(), d_lag.data_ptr(), d_sub.data_ptr(), engine="tcgen05")
        engine.ccx_pack_rows(d_cc.data_ptr(), d_lag.data_ptr(), d_sub.data_ptr(), rows, N, out)
    cc, lag, sub = engine.ccx_condensed(X, 3, engine="tcgen05")
    assert np.array_equal(out[0], cc) and np.array_equal(out[1], lag) and np.array_equal(out[2], sub)
    rcc, rlag, rsub = orc.make_cclags(X, 3)
    iu = np.triu_indices(N, 1)
    assert np.abs(out[0] - rcc[iu[0], iu[1] - 1]).max() < 1e-10
    assert np.array_equal(out[1].astype(float), rlag[iu[0], iu[1] - 1])
    # pageable outputs are refused (the kernel itself writes them), loudly
    bad = (np.empty(npair), np.empty(npair, np.int32), np.empty(npair))
    with pytest.raises(DtxError):
        engine.ccx_pack_rows(d_cc.data_ptr(), d_lag.data_ptr(), d_sub.data_ptr(), rows, N, bad)


def test_long_events_take_the_fallback_paths(engine):
    """The only workload the reference publishes a timing for (130 s x 100 Hz x 1 channel events):
    n = 13 000 -> 13 001 lags (several 2048-lag tiles: no DUAL items) and 104 KB per float64 waveform (two do not
    fit beside a ring in shared memory: tiled re-scoring with one resident signal).  Tensor engine = float64 engine
    = oracle."""
    X = synth.event_families(130, 2, 3, 13000, 1, max_shift=300)         # 6 events
    cc, lag, sub = engine.ccx(X, 1, engine="tcgen05")
    c64, l64, s64 = engine.ccx(X, 1, engine="fp64")
    iu = np.triu_indices(len(X), 1)
    assert np.array_equal(lag[iu], l64[iu])
    assert np.abs(cc[iu] - c64[iu]).max() < 1e-12
    assert np.nanmax(np.abs(sub[iu] - s64[iu])) < 1e-7
    for b, c in ((0, 1), (1, 4), (3, 5)):
        rcc, rlag, rsub = orc.ccx2(X[b], X[c], 1)
        assert abs(cc[b, c] - rcc) < 1e-10 and lag[b, c] == rlag
