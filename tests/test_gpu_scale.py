"""Parity at the BASELINE.json scales, against the ORACLE (not against another engine of this repo):

  * CCX through several signal batches and >= 38 basis blocks (configs[2] code path: N = 600 with a
    lowered batch cap against `orc.make_cclags` on every pair; N = 4096 / n = 3000 with 2000 random pairs
    against `orc.ccx2` and the rest against the float64 engine);
  * detection at configs[3]'s shape -- S = 256 subspaces, R = 1152 basis vectors (72 basis blocks),
    n = 9000, two full 3720 s chunks -- on subspaces from the first / middle / last basis blocks against
    `orc.mpx_ds_fft`, for the default work-item order and the plain (superblock 1) order;
  * the 8-bit cross-term engine really runs (its DS differs from the default engine's) and stays in
    tolerance;
  * FAS at configs[4]'s chunk shape (L = 1 080 000) against `orc.fas_stats`;
  * rank > 16 subspaces are bit-reproducible run to run.

Tolerances: |DS - reference| <= 1e-5, CC 1e-10, lags bit-exact (BASELINE.json north_star)."""
import os

import numpy as np
import pytest

from detex_b200 import fas, synth
from oracle import detex_oracle as orc

pytestmark = pytest.mark.gpu
TOL = 1e-5
NC, NS, LS = 3, 3000, 372000


# ------------------------------------------------------------------------------------------ CCX
def test_ccx_tensor_multibatch_n600_vs_oracle_all_pairs(engine):
    """600 events (38 basis blocks of 16 templates), batch cap lowered to 200 signals -> 3 signal
    batches with per-signal basis-block trimming; every one of the 179 700 pairs against the oracle."""
    X = synth.event_families(6001, 20, 30, 200, 3, max_shift=40)          # n = 600, 201 lags
    assert X.shape == (600, 600)
    engine.set_ccx_batch(200, 4 << 30)
    try:
        cc, lag, sub = engine.ccx_condensed(X, 3, engine="tcgen05")
    finally:
        engine.set_ccx_batch()
    rcc, rlag, rsub = orc.make_cclags(X, 3, fft=True)
    iu = np.triu_indices(len(X), 1)
    rcc, rlag, rsub = rcc[iu[0], iu[1] - 1], rlag[iu[0], iu[1] - 1], rsub[iu[0], iu[1] - 1]
    assert np.abs(cc - rcc).max() < 1e-10
    assert np.array_equal(lag.astype(float), rlag)
    assert np.nanmax(np.abs(sub - rsub)) < 1e-6
    # dense rows of a middle block through the same multi-batch path
    engine.set_ccx_batch(128, 4 << 30)
    try:
        c2, l2, _ = engine.ccx(X, 3, row_begin=250, row_end=330, engine="tcgen05")
    finally:
        engine.set_ccx_batch()
    for b in range(250, 330):
        o = b * 600 - b * (b + 1) // 2
        assert np.array_equal(c2[b - 250, b + 1:], cc[o:o + 600 - b - 1])
        assert np.array_equal(l2[b - 250, b + 1:], lag[o:o + 600 - b - 1])


def test_ccx_config2_4096_events(engine):
    """BASELINE configs[2]: 4096 events x 3 ch x 10 s x 100 Hz (n = 3000, 1001 lags, 256 basis blocks,
    16 signal batches)."""
    X = synth.event_families(3003, 64, 64, 1000, 3, max_shift=100)
    N = len(X)
    assert X.shape == (4096, 3000)
    cc, lag, sub = engine.ccx_condensed(X, 3, engine="tcgen05")
    rng = np.random.default_rng(5)
    b = rng.integers(0, N - 1, size=2000)
    c = np.array([rng.integers(bi + 1, N) for bi in b])
    idx = b * N - b * (b + 1) // 2 + (c - b - 1)
    worst, worst_sub = 0.0, 0.0
    for bi, ci, k in zip(b, c, idx):
        m, l, s = orc.ccx2(X[bi], X[ci], 3)
        worst = max(worst, abs(cc[k] - m))
        assert lag[k] == l, (bi, ci)
        if np.isfinite(s):
            worst_sub = max(worst_sub, abs(sub[k] - s))
    assert worst < 1e-10 and worst_sub < 1e-6
    c64, l64, s64 = engine.ccx_condensed(X, 3, engine="fp64")
    assert np.abs(cc - c64).max() < 1e-12
    assert np.array_equal(lag, l64)
    assert np.nanmax(np.abs(sub - s64)) < 1e-8
    # within-family pairs correlate, lags are multiples of Nc inside the search range
    assert cc[0] > 0.3 and np.all(lag % 3 == 0) and np.abs(lag).max() <= 1503


# ------------------------------------------------------------------------------------ detection
@pytest.fixture(scope="module")
def cfg3_case():
    rng = np.random.default_rng(4004)
    ranks = [(i % 8) + 1 for i in range(256)]
    bases = [synth.random_basis(rng, NC * NS, r) for r in ranks]
    chunks = [synth.multiplex(synth.bandpassed_noise(rng, LS, nchan=NC)) for _ in range(2)]
    # planted events: a member of subspace 5 in chunk 0, of subspace 250 in chunk 1
    for ci, s, t0 in ((0, 5, 70000), (1, 250, 301234)):
        tem = np.ones(ranks[s]) @ bases[s]
        chunks[ci] = synth.plant(chunks[ci], tem, t0, NC, 6.0 * np.sqrt(NC * NS) / np.linalg.norm(tem))
    return ranks, bases, chunks


_REF = {}


def _oracle_ds(chunks, bases, ci, s):
    """orc.mpx_ds_fft of (chunk ci, subspace s), computed once per session (1-2 s each)."""
    if (ci, s) not in _REF:
        _REF[(ci, s)] = orc.mpx_ds_fft(chunks[ci], bases[s], NC)
    return _REF[(ci, s)]


@pytest.mark.parametrize("order", ["default", "super1"])
def test_detection_config3_shape_vs_oracle(engine, cfg3_case, order):
    ranks, bases, chunks = cfg3_case
    assert sum(ranks) == 1152
    engine.set_bases(70, bases, NC, thresholds=[0.25] * 256)
    engine.load_chunks(chunks)
    old = os.environ.get("DTX_K1_SUPER")
    if order == "super1":
        os.environ["DTX_K1_SUPER"] = "1"
    try:
        engine.detect_run(70, engine="tcgen05", lta_window=500)
    finally:
        if old is None:
            os.environ.pop("DTX_K1_SUPER", None)
        else:
            os.environ["DTX_K1_SUPER"] = old
    # subspaces of every rank; the rank-8 ones are packed first (first basis blocks), the rank-1 ones last
    subs = [0, 7, 5, 15, 64, 100, 121, 127, 128, 133, 190, 201, 248, 250, 252, 255]
    worst = 0.0
    for ci in (0, 1):
        for s in subs:
            ref = _oracle_ds(chunks, bases, ci, s)
            ds = engine.get_ds(ci, s)
            assert ds.shape == ref.shape == (LS - NS + 1,)
            worst = max(worst, float(np.abs(ds - ref).max()))
    assert worst < TOL, worst
    mx, fl = engine.rowstats()
    assert mx[0, 5] > 0.9 and mx[1, 250] > 0.9 and not fl.any()
    cand = engine.candidates()
    rows = set(cand["row"].tolist())
    assert 0 * 256 + 5 in rows and 1 * 256 + 250 in rows
    best = cand[cand["row"] == 5]
    assert best["t"][np.argmax(best["ds"])] == 70000


def test_x8_engine_is_not_vacuous(engine, cfg3_case):
    """On an admitted chunk the adaptive engine's DS differs from the default engine's (the 8-bit MMA
    really ran) and both are within tolerance of the oracle."""
    ranks, bases, chunks = cfg3_case
    subs = [5, 100, 255]
    engine.set_bases(71, [bases[s] for s in subs], NC)
    engine.load_chunks(chunks[:1])
    engine.detect_run(71, engine="tcgen05")
    d0 = [engine.get_ds(0, i).copy() for i in range(len(subs))]
    engine.detect_run(71, engine="tcgen05_auto")
    assert engine.chunk_modes().tolist() == [1]
    d8 = [engine.get_ds(0, i).copy() for i in range(len(subs))]
    for i, s in enumerate(subs):
        ref = _oracle_ds(chunks, bases, 0, s)
        assert not np.array_equal(d0[i], d8[i])
        assert (d0[i] != d8[i]).mean() > 0.5            # not a corner: most values move by an ulp or more
        assert np.abs(d0[i] - ref).max() < TOL and np.abs(d8[i] - ref).max() < TOL


# ------------------------------------------------------------------------------------------ FAS
def test_fas_config4_chunk_shape_vs_oracle(engine):
    """configs[4] chunk shape: 3600 s x 3 ch x 100 Hz (L = 1 080 000, T = 357 001), n = 9000."""
    rng = np.random.default_rng(5005)
    Ls = 360000
    ranks = [1, 3, 5, 8]
    bases = [synth.random_basis(rng, NC * NS, r) for r in ranks]
    chunks = [synth.multiplex(synth.bandpassed_noise(rng, Ls, nchan=NC)) for _ in range(2)]
    assert len(chunks[0]) == 1080000
    res = fas.initFAS(bases, chunks, NC, engine=engine, set_id=907, batch=1)
    for si, U in enumerate(bases):
        ref = orc.fas_stats([orc.mpx_ds_fft(c, U, NC) for c in chunks])
        assert res[si]["hist"].sum() == ref["hist"].sum() == 2 * (Ls - NS + 1)
        # a value within 1e-5 of a bin edge may sit in the neighbouring bin
        assert np.abs(res[si]["hist"] - ref["hist"]).sum() <= 2 * max(8, int(2e-5 / (1.01 / 400) * ref["hist"].sum()))
        a, b = res[si]["betadist"][:2]
        ra, rb = ref["betadist"][:2]
        assert abs(a - ra) < 1e-4 * ra and abs(b - rb) < 1e-4 * rb
        assert abs(res[si]["nnlf"] - ref["nnlf"]) < 1e-5 * abs(ref["nnlf"])
        assert abs(orc.threshold_from_beta(a, b) - orc.threshold_from_beta(ra, rb)) < 1e-5


# --------------------------------------------------------------------------- determinism, gaps
def test_rank_above_16_is_bit_reproducible(engine):
    """Pieces of a rank > 16 subspace are summed in a fixed order: two runs give identical bits."""
    Nc, ns, Ls = 3, 200, 20000
    chunks, bases, _ = synth.detection_case(61, 2, Ls, ns, Nc, [40, 3, 17, 16], planted=2)
    engine.set_bases(72, bases, Nc, thresholds=[0.3] * 4)
    engine.load_chunks(chunks)
    runs = []
    for _ in range(3):
        engine.detect_run(72, lta_window=50)
        runs.append(([engine.get_ds(ci, si).copy() for ci in range(2) for si in range(4)], engine.candidates()))
    for ds, cand in runs[1:]:
        assert all(np.array_equal(a, b) for a, b in zip(ds, runs[0][0]))
        assert np.array_equal(np.sort(cand, order=["row", "t"]), np.sort(runs[0][1], order=["row", "t"]))
    for si, U in enumerate(bases):
        assert np.abs(runs[0][0][si] - orc.mpx_ds_direct(chunks[0], U, Nc)).max() < TOL


def test_zero_filled_gap_matches_reference_golden(engine, gap_golden):
    """A zero-filled gap longer than the template (fillZeros=True): the reference's statistic is +inf
    on the windows inside the gap and is zeroed by the MaxDS > 1.1 rule (detect.py:275-281); the rest
    of the chunk keeps its detections.  Golden vector from the reference's own FFT path."""
    g = gap_golden
    x, Nc = g["gap_chunk"], int(g["gap_Nc"])
    bases = [g["gap_U0"], g["gap_U1"]]
    engine.set_bases(73, bases, Nc, thresholds=[0.5, 0.5])
    engine.hist(73, reset=True)
    engine.load_chunks([x])
    engine.detect_run(73, lta_window=50)
    mx, fl = engine.rowstats()
    for si in range(2):
        ref = g["gap_DS%d" % si]
        ds = engine.get_ds(0, si)
        inf_ref = np.isinf(ref)
        assert inf_ref.sum() > 300
        assert np.array_equal(np.isinf(ds), inf_ref) and not np.isnan(ds).any()
        assert np.abs(ds[~inf_ref] - ref[~inf_ref]).max() < TOL
        zeroed = np.where(inf_ref, 0.0, ref)
        assert fl[0, si] == 2 and abs(mx[0, si] - zeroed.max()) < TOL
    hist = engine.hist(73, reset=True)
    assert hist.sum() == 2 * len(g["gap_DS0"])               # the gap's windows count as DS = 0
    cand = engine.candidates()
    assert len(cand) > 0 and set(cand["row"].tolist()) == {1}   # the planted event of subspace 1 survives
    assert np.isfinite(cand["ds"]).all() and np.isfinite(cand["lta"]).all()


def test_long_array_time_segments_with_halo(engine):
    """Time-segment sharding with a halo on the GPU: a 3-hour array cut into 7 overlapping segments
    gives the triggers (bit-exact times), maxima and histograms of the same array run as ONE chunk."""
    from detex_b200 import detect
    Nc, ns, Ls, sr = 3, 300, 160000, 100.0
    chunks, bases, _ = synth.detection_case(94, 1, Ls, ns, Nc, [2, 4, 7], planted=6)
    x = chunks[0]
    names = ["SS0", "SS1", "SS2"]
    mk = lambda sid: detect.SSDetex(dict(zip(names, bases)), {n: 0.3 for n in names}, {n: [0.0, 2.0] for n in names},
                                    Nc, engine=engine, set_id=sid)
    one = mk(21)
    ref, mref, _ = one.run_chunks([x], sr, [5.0e8])
    det = mk(22)
    got, mx = det.run_long_array(x, sr, 5.0e8, seg_lags=24576, batch=3)
    assert len(det.long_array_segments(Ls, ns, 24576, 500)) == 7
    assert len(ref) >= 4 and len(got) == len(ref)
    assert np.array_equal(got.STMP.values, ref.STMP.values) and list(got.Name) == list(ref.Name)
    assert np.abs(got.DS.values - ref.DS.values).max() < 2e-6     # per-segment centring / scaling of the fp16 split
    assert np.abs(got.DS_STALTA.values / ref.DS_STALTA.values - 1).max() < 1e-4
    for name in names:
        assert abs(mx[name] - mref[0][name]) < 2e-6
        assert det.histdic[name].sum() == one.histdic[name].sum() == Ls - ns + 1
        assert np.abs(det.histdic[name] - one.histdic[name]).sum() <= 8
    # and against the oracle on the uncut array
    for si, name in enumerate(names):
        ds = orc.mpx_ds_direct(x, bases[si], Nc)
        sel = got[got.Name == name]
        t = np.rint((sel.STMP.values - 5.0e8) * sr).astype(int)
        assert np.abs(sel.DS.values - ds[t]).max() < TOL


def test_ccx_screening_series_equals_three_pass_series(engine):
    """The tensor-core CCX series only locates the maximum: the default one-MMA (hi*hi) screening series
    and the fp16x3 series give the same cc / lag / subsamp (float64 re-scoring of the same lags) -- also
    for waveforms with a DC offset, where the rounding of the operands is amplified and the candidate
    band widens accordingly."""
    X = synth.event_families(6002, 8, 20, 500, 3, max_shift=60)           # 160 events, n = 1500
    X[7] = X[3]                                                           # an identical pair (cc = 1)
    X[11] = 0.0                                                           # a zeroed-out waveform
    rng = np.random.default_rng(9)
    Xoff = synth.event_families(6003, 2, 12, 300, 3, max_shift=30) + 4.0 * rng.standard_normal((24, 1))
    for Xc, check_oracle in ((X, False), (Xoff, True)):
        res = {}
        for passes in (1, 3):
            engine.set_ccx_passes(passes)
            try:
                res[passes] = engine.ccx_condensed(Xc, 3, engine="tcgen05")
            finally:
                engine.set_ccx_passes(1)
        assert np.abs(res[1][0] - res[3][0]).max() < 1e-12
        assert np.array_equal(res[1][1], res[3][1])
        assert np.nanmax(np.abs(res[1][2] - res[3][2])) < 1e-8
        c64, l64, _ = engine.ccx_condensed(Xc, 3, engine="fp64")
        assert np.abs(res[1][0] - c64).max() < 1e-12 and np.array_equal(res[1][1], l64)
        if check_oracle:
            rcc, rlag, _ = orc.make_cclags(Xc, 3, fft=False)
            iu = np.triu_indices(len(Xc), 1)
            assert np.abs(res[1][0] - rcc[iu[0], iu[1] - 1]).max() < 1e-9
            assert np.array_equal(res[1][1].astype(float), rlag[iu[0], iu[1] - 1])
