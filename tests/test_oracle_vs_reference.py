"""Oracle against the LIVE reference functions (imported through oracle/ref_shim.py) on
fresh seeded inputs.  Only runs where /root/reference exists (the build container); the
committed golden vectors cover the same ground elsewhere."""
import numpy as np
import pytest

from detex_b200 import synth
from oracle import detex_oracle as orc
from oracle import ref_shim

pytestmark = pytest.mark.skipif(not ref_shim.available(), reason="reference tree not present")


@pytest.fixture(scope="module")
def R():
    return ref_shim.RefFunctions()


@pytest.mark.parametrize("seed,Nc,ns,Ls,ranks", [(1, 3, 200, 4000, [2, 7]), (2, 1, 128, 3000, [1]),
                                                 (3, 2, 251, 2999, [16])])
def test_mpxds(R, seed, Nc, ns, Ls, ranks):
    chunks, bases, _ = synth.detection_case(seed, 1, Ls, ns, Nc, ranks, planted=1)
    for U in bases:
        ref = R.MPXDS(chunks[0], U, Nc)
        assert np.abs(orc.mpx_ds_fft(chunks[0], U, Nc) - ref).max() < 1e-12
        assert np.abs(orc.mpx_ds_direct(chunks[0], U, Nc) - ref).max() < 1e-12
        assert np.abs(R.MPXSSCorr(chunks[0], U, Nc) - ref).max() == 0.0


def test_ccx(R):
    X = synth.event_families(9, 2, 3, 150, 3, max_shift=10)
    rcc, rlag, rsub = R.makeDFcclags(X, 3)
    cc, lag, sub = orc.make_cclags(X, 3)
    m = ~np.isnan(cc)
    assert np.abs(cc[m] - rcc.values.astype(float)[m]).max() < 1e-12
    assert np.array_equal(lag[m], rlag.values.astype(float)[m])
    assert np.abs(sub[m] - rsub.values.astype(float)[m]).max() < 1e-9
    for b in range(2):
        a = R.CCX2(X[b], X[b + 3], 3)
        o = orc.ccx2(X[b], X[b + 3], 3)
        assert abs(a[0] - o[0]) < 1e-12 and a[1] == o[1] and abs(a[2] - o[2]) < 1e-9


def test_subsamp_quirk(R):
    c = np.array([0.1, 0.9, 0.2, 0.95, 0.1])
    for ind in (0, 1, 3, 4):
        a, b = R.subSamp(c, ind), orc.sub_samp(c, ind)
        assert (np.isnan(a) and np.isnan(b)) or a == b


def test_downplay_and_triggers(R):
    rng = np.random.default_rng(4)
    ds = rng.uniform(0, .2, size=9000)
    ds[[10, 2100, 2150, 4400, 8990]] = [.5, .9, .8, .7, .6]
    st = R.getStaLtaArray(ds, 250.0, 0)
    assert np.abs(orc.sta_lta(ds, 250.0, 0) - st).max() < 1e-12
    df = R.CreateCoeffArray(ds, st, 50.0, 123.0, 0.3, [0.5, 1.5])
    rows = orc.greedy_triggers(ds, 0.3, 50.0, 123.0, [0.5, 1.5], stalta=st)
    assert np.array_equal(df.DS.values.astype(float), [r["DS"] for r in rows])
    assert np.array_equal(df.STMP.values.astype(float), [r["STMP"] for r in rows])


def test_multiplex(R):
    ch = [np.arange(5.0), np.arange(5.0) * 2, np.arange(6.0) * 3]
    assert np.array_equal(R.multiplex(ch), orc.multiplex(ch))
    assert np.array_equal(synth.multiplex([c[:5] for c in ch]), orc.multiplex(ch))


@pytest.mark.parametrize("seed,N", [(31, 2), (32, 6), (33, 19)])
def test_alignment_delays(R, seed, N):
    """construct.py:710-812 through the reference's own `_traceEventDendro` / `_alignTD`."""
    import pandas as pd
    rng = np.random.default_rng(seed)
    iu = np.triu_indices(N - 1)
    cc = np.full((N - 1, N - 1), np.nan)
    lag = np.full((N - 1, N - 1), np.nan)
    cc[iu] = rng.permutation(np.linspace(0.2, 0.95, len(iu[0])))
    lag[iu] = 3 * rng.integers(-60, 60, len(iu[0]))
    idx, cols = range(N - 1), range(1, N)
    link, delays = R.getDelays(pd.DataFrame(cc, index=idx, columns=cols), pd.DataFrame(lag, index=idx, columns=cols))
    l2, d2 = orc.get_delays(cc, lag)
    assert np.array_equal(link, l2) and np.array_equal(delays, d2)
    X = rng.standard_normal((N, 1500))
    assert np.array_equal(R.alignTD(delays, X), orc.align_td(d2, X))


@pytest.mark.parametrize("crit,val,nm", [(1, 0.7, False), (2, 0.9, True), (3, 0.8, False), (4, 1, False)])
def test_svd_selection_live(R, crit, val, nm):
    """a16 against the live reference: `_trimGroups` / svd / `_getFracEnergy` / `_getUsedBasis`
    (subspace.py:875-1013), unmodified, on a fresh cluster."""
    from detex_b200 import subspace
    X = synth.event_families(900 + crit, 1, 5, 150, 3, max_shift=5, noise=0.6) - 0.02
    ref = R.svdSelect(X, 12, 432, crit, val, normalize=nm)
    W = X[:, 12:432]
    for got in (orc.svd_basis(W, select_criteria=crit, select_value=val, normalize=nm),
                subspace.svd_basis(W, selectCriteria=crit, selectValue=val, normalize=nm)):
        U = got["U"]
        assert U.shape == ref["U"].shape and np.allclose(got["s"], ref["s"], rtol=1e-12)
        for a, b in zip(U, ref["U"]):
            assert min(np.abs(a - b).max(), np.abs(a + b).max()) < 1e-10
    assert np.abs(orc.svd_basis(W, crit, val, nm)["frac_avg"] - ref["frac_avg"]).max() < 1e-12
