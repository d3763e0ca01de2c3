"""`detex_b200.dropin.install`: the reference-side binding of INTEGRATION.md section 1.

CPU part: binding / restoring mechanics on a stand-in package (no compute).
GPU part: the UNMODIFIED reference (the copy `oracle/make_ref.py` puts under `oracle/_ref`, which
travels to the GPU box; `/root/reference` in the build container) runs its OWN loops -- the pair loop of
`construct._makeDFcclags` (construct.py:369-394), `_SSDetex._MPXDS` as a bound method, `fas._MPXSSCorr` --
with the three callables rebound to the CUDA path, and must return what it returns without the rebinding.
"""
import sys
import types

import numpy as np
import pytest

from detex_b200 import dropin, synth


def _fake_package(name="fakedetex"):
    pkg = types.ModuleType(name)
    pkg.__path__ = []
    det = types.ModuleType(name + ".detect")
    fas = types.ModuleType(name + ".fas")
    con = types.ModuleType(name + ".construct")

    class _SSDetex(object):
        def _MPXDS(self, MPcon, reqlen, ssTD, ssFD, Nc, MPconFD):
            return "ref-mpxds"

    det._SSDetex = _SSDetex
    fas._MPXSSCorr = lambda *a: "ref-sscorr"
    con._CCX2 = lambda *a: "ref-ccx2"
    con._makeDFcclags = lambda *a: "ref-cclags"
    for m in (pkg, det, fas, con):
        sys.modules[m.__name__] = m
    pkg.detect, pkg.fas, pkg.construct = det, fas, con
    return pkg


def test_install_rebinds_and_uninstall_restores():
    pkg = _fake_package()
    try:
        orig = (pkg.detect._SSDetex._MPXDS, pkg.fas._MPXSSCorr, pkg.construct._CCX2, pkg.construct._makeDFcclags)
        saved = dropin.install(pkg)
        assert saved["_MPXDS"] is orig[0] and saved["_CCX2"] is orig[2]
        assert pkg.detect._SSDetex._MPXDS is not orig[0]
        assert pkg.fas._MPXSSCorr is not orig[1]
        assert pkg.construct._CCX2 is not orig[2]
        assert pkg.construct._makeDFcclags is orig[3]              # the pair loop stays the reference's own
        dropin.install(pkg, batched=True)                           # re-install replaces, does not stack
        assert pkg.construct._makeDFcclags is not orig[3]
        dropin.uninstall(pkg)
        now = (pkg.detect._SSDetex._MPXDS, pkg.fas._MPXSSCorr, pkg.construct._CCX2, pkg.construct._makeDFcclags)
        assert all(a is b for a, b in zip(orig, now))
        assert pkg.detect._SSDetex()._MPXDS(0, 0, 0, 0, 0, 0) == "ref-mpxds"
        dropin.uninstall(pkg)                                       # idempotent
    finally:
        for k in [k for k in sys.modules if k.startswith("fakedetex")]:
            del sys.modules[k]


def test_rebound_signatures_are_the_references():
    import inspect
    pkg = _fake_package()
    try:
        dropin.install(pkg, batched=True)
        # detect.py:559, fas.py:120, construct.py:425, construct.py:369
        assert list(inspect.signature(pkg.detect._SSDetex._MPXDS).parameters) == \
            ["self", "MPcon", "reqlen", "ssTD", "ssFD", "Nc", "MPconFD"]
        assert list(inspect.signature(pkg.fas._MPXSSCorr).parameters) == \
            ["MPcon", "reqlen", "ssArrayTD", "ssArrayFD", "Nc"]
        assert list(inspect.signature(pkg.construct._CCX2).parameters) == \
            ["mpfd1", "mpfd2", "mptd1", "mptd2", "Nc1", "Nc2"]
        assert list(inspect.signature(pkg.construct._makeDFcclags).parameters) == ["eventList", "row"]
    finally:
        for k in [k for k in sys.modules if k.startswith("fakedetex")]:
            del sys.modules[k]


# ------------------------------------------------------------------------------------------ GPU part
def _ref():
    from oracle import ref_shim
    if not ref_shim.available():
        pytest.skip("no reference tree (neither /root/reference nor oracle/_ref)")
    return ref_shim.RefFunctions()


def _frames_equal(a, b, tol):
    va, vb = a.to_numpy(dtype=float), b.to_numpy(dtype=float)
    assert va.shape == vb.shape and list(a.index) == list(b.index) and list(a.columns) == list(b.columns)
    assert np.array_equal(np.isnan(va), np.isnan(vb))
    m = ~np.isnan(va)
    return np.abs(va[m] - vb[m]).max() <= tol if m.any() else True


@pytest.mark.gpu
def test_reference_pair_loop_runs_on_the_gpu(engine):
    ref = _ref()
    X = synth.event_families(4242, 3, 4, 200, 3, max_shift=15)          # 12 events, n = 600
    want = ref.makeDFcclags(X, 3)                                        # unmodified reference, FFT path
    dropin.install(ref.detex, engine=engine)
    try:
        assert ref.construct._CCX2.__module__ == "detex_b200.dropin"
        got = ref.makeDFcclags(X, 3)                                     # the reference's loop, our _CCX2 per pair
        dropin.install(ref.detex, engine=engine, batched=True)
        got_b = ref.construct._makeDFcclags(*_cclags_args(X, 3))         # one call for the whole matrix
    finally:
        dropin.uninstall(ref.detex)
    assert ref.construct._CCX2.__module__ == "detex.construct"
    for g in (got, got_b):
        assert _frames_equal(g[0], want[0], 1e-10)                       # DFcc
        assert _frames_equal(g[1], want[1], 0.0)                         # DFlag: bit-exact
        assert _frames_equal(g[2], want[2], 1e-7)                        # DFsubsamp


def _cclags_args(X, Nc):
    import pandas as pd
    evs = ["ev%04d" % i for i in range(X.shape[0])]
    chans = ["C%d" % i for i in range(Nc)]
    row = pd.Series({"MPtd": {e: X[i] for i, e in enumerate(evs)}, "MPfd": {e: None for e in evs},
                     "Channels": {e: chans for e in evs}})
    return evs, row


@pytest.mark.gpu
def test_reference_mpxds_and_sscorr_run_on_the_gpu(engine):
    ref = _ref()
    rng = np.random.default_rng(77)
    Nc, ns, Ls = 3, 120, 6000
    U = synth.random_basis(rng, Nc * ns, 4)
    x = synth.multiplex(synth.bandpassed_noise(rng, Ls, sr=100.0, nchan=Nc))
    x[3000:3000 + Nc * ns] += 12.0 * U[1] * np.abs(x).max()
    want_d = ref.MPXDS(x, U, Nc)
    want_f = ref.MPXSSCorr(x, U, Nc)
    dropin.install(ref.detex, engine=engine)
    try:
        got_d = ref.MPXDS(x, U, Nc)          # _SSDetex._MPXDS as a bound method of the reference's own class
        got_f = ref.MPXSSCorr(x, U, Nc)
    finally:
        dropin.uninstall(ref.detex)
    assert got_d.dtype == np.float64 and got_d.shape == want_d.shape
    assert np.abs(got_d - want_d).max() < 1e-5 and want_d.max() > 0.5
    assert got_f.shape == want_f.shape and np.abs(got_f - want_f).max() < 1e-5
    assert int(np.argmax(got_d)) == int(np.argmax(want_d))
