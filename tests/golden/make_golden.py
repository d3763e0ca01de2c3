"""Generate the golden vectors under tests/golden/ from the UNMODIFIED reference.

Run in the build container (where /root/reference exists):
    python tests/golden/make_golden.py
The reference functions are imported through oracle/ref_shim.py and fed seeded synthetic
inputs; inputs and the reference's outputs are stored together so that the oracle and the
CUDA path can be checked against them anywhere (the GPU box has no /root/reference).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from detex_b200 import synth  # noqa: E402
from oracle import ref_shim  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def gap_case(R):
    """A zero-filled gap longer than the template (fillZeros=True merges, construct.py:1004-1005):
    the reference's own FFT path gives +inf on the windows inside the gap (sum(if1^2)/0,
    detect.py:577), which _getRA zeroes when MaxDS > 1.1 (detect.py:275-281)."""
    chunks, bases, _ = synth.detection_case(seed=105, nchunks=1, Ls=6000, ns=300, Nc=3, ranks=[2, 4], planted=2)
    x = chunks[0].copy()
    x[3000 * 3:3000 * 3 + 3 * 700] = 0.0
    out = dict(gap_chunk=x, gap_Nc=3, gap_U0=bases[0], gap_U1=bases[1])
    for si, U in enumerate(bases):
        ds = R.MPXDS(x, U, 3)
        assert np.isinf(ds).sum() == 401 and not np.isnan(ds).any()
        out["gap_DS%d" % si] = ds
    np.savez_compressed(os.path.join(OUT, "gap_golden.npz"), **out)


def svd_case(R):
    """SubSpace.SVD's per-subspace body (subspace.py:875-905, 921-943, 968-1013) run by the reference
    itself: aligned waveforms -> trimmed, demeaned stack -> SVD -> fractional energy -> used basis, for
    every selectCriteria."""
    X = synth.event_families(301, 1, 7, 260, 3, max_shift=6, noise=0.8) + 0.05     # one cluster, n = 780
    out = dict(svd_X=X, svd_start=30, svd_end=750)
    cases = [(1, 0.5, False), (2, 0.9, False), (2, 0.6, True), (3, 0.95, False), (4, 2, False), (2, 1.0, False),
             (2, 0.0, False)]
    out["svd_cases"] = np.array([[c, v, float(nm)] for c, v, nm in cases])
    for k, (crit, val, nm) in enumerate(cases):
        r = R.svdSelect(X, 30, 750, crit, val, normalize=nm)
        assert r["basisLength"] == 720
        for name in ("s", "frac_avg", "frac_min", "used_keys", "U"):
            out["svd%d_%s" % (k, name)] = r[name]
    np.savez_compressed(os.path.join(OUT, "svd_golden.npz"), **out)


def main():
    R = ref_shim.RefFunctions()
    gap_case(R)
    svd_case(R)
    if len(sys.argv) > 1 and sys.argv[1] in ("gap", "svd"):
        return
    # ---- detection statistic (_MPXDS, detect.py:559-578; _MPXSSCorr, fas.py:120-134)
    cases = {}
    specs = [
        ("nc3", dict(seed=101, nchunks=2, Ls=6000, ns=300, Nc=3, ranks=[1, 3, 5, 8], planted=2)),
        ("nc1", dict(seed=102, nchunks=1, Ls=8000, ns=250, Nc=1, ranks=[2, 4], planted=2)),
        ("nc2odd", dict(seed=103, nchunks=1, Ls=5003, ns=333, Nc=2, ranks=[3, 16], planted=1)),
    ]
    for name, kw in specs:
        chunks, bases, truth = synth.detection_case(**kw)
        if name == "nc3":
            chunks[1] = chunks[1] + 250.0      # DC level (fillZeros-like conditioning case)
        cases[name + "_Nc"] = kw["Nc"]
        cases[name + "_nchunks"] = len(chunks)
        cases[name + "_nbases"] = len(bases)
        for ci, c in enumerate(chunks):
            cases["%s_chunk%d" % (name, ci)] = c
        for si, U in enumerate(bases):
            cases["%s_U%d" % (name, si)] = U
            for ci, c in enumerate(chunks):
                ds = R.MPXDS(c, U, kw["Nc"])
                ds2 = R.MPXSSCorr(c, U, kw["Nc"])
                assert np.abs(ds - ds2).max() == 0.0
                cases["%s_DS_%d_%d" % (name, ci, si)] = ds
    # singleton: unit-norm, non-demeaned template (detect.py:356-357)
    rng = np.random.default_rng(104)
    x = synth.multiplex(synth.bandpassed_noise(rng, 5000, nchan=3))
    tem = x[3000:3000 + 900].copy() + 0.3
    U = (tem / np.linalg.norm(tem))[None, :]
    cases["single_chunk"] = x
    cases["single_U"] = U
    cases["single_DS"] = R.MPXDS(x, U, 3)
    np.savez_compressed(os.path.join(OUT, "ds_golden.npz"), **cases)

    # ---- STA/LTA + greedy trigger picking (detect.py:390-445, 501-557)
    ds = cases["nc3_DS_0_1"]
    trig = {}
    for lta, sta in ((50.0, 0), (500.0, 0), (50.0, 7.0)):
        trig["stalta_%g_%g" % (lta, sta)] = R.getStaLtaArray(ds, lta, sta)
    rng = np.random.default_rng(7)
    syn = rng.uniform(0, 0.2, size=20000)
    for pos, val in ((5, .8), (1500, .9), (2300, .85), (3600, .7), (3601, .7), (9000, .95), (10990, .6),
                     (11050, .9), (19990, .75), (14000, .5), (14000 + 1999, .45), (14000 + 2000, .44)):
        syn[pos] = val
    stl = R.getStaLtaArray(syn, 500.0, 0)
    df = R.CreateCoeffArray(syn, stl, 100.0, 1.0e9, 0.4, [1.0, 2.5, 4.0])
    trig["syn_DS"] = syn
    trig["syn_stalta"] = stl
    trig["syn_trig_DS"] = df.DS.values.astype(float)
    trig["syn_trig_STMP"] = df.STMP.values.astype(float)
    trig["syn_trig_STALTA"] = df.DS_STALTA.values.astype(float)
    trig["syn_trig_MSTAMPmin"] = df.MSTAMPmin.values.astype(float)
    trig["syn_trig_MSTAMPmax"] = df.MSTAMPmax.values.astype(float)
    np.savez_compressed(os.path.join(OUT, "trigger_golden.npz"), **trig)

    # ---- CCX (_makeDFcclags / _CCX2 / _subSamp, construct.py:369-466)
    ccx = {}
    for name, (seed, nfam, per, ns, Nc) in {"nc3": (201, 3, 4, 200, 3), "nc1": (202, 2, 3, 301, 1)}.items():
        X = synth.event_families(seed, nfam, per, ns, Nc, max_shift=20)
        if name == "nc3":
            X[5] = 0.0          # an all-zero waveform (zeroed-out event, construct.py:457-461)
            X[7] = X[2]         # identical pair
        cc, lag, sub = R.makeDFcclags(X, Nc)
        ccx[name + "_X"] = X
        ccx[name + "_Nc"] = Nc
        ccx[name + "_cc"] = cc.values.astype(float)
        ccx[name + "_lag"] = lag.values.astype(float)
        ccx[name + "_sub"] = sub.values.astype(float)
    np.savez_compressed(os.path.join(OUT, "ccx_golden.npz"), **ccx)

    # ---- alignment from the dendrogram (_getDelays/_traceEventDendro/_alignTD,
    #      construct.py:272-286, 486-503, 710-812): reference CCX output -> reference delays
    import pandas as pd
    al = {}
    X = synth.event_families(203, 3, 5, 240, 3, max_shift=25)
    cc, lag, _ = R.makeDFcclags(X, 3)
    link, delays = R.getDelays(cc.astype(float), lag.astype(float))
    al.update(fam_X=X, fam_cc=cc.values.astype(float), fam_lag=lag.values.astype(float), fam_link=link,
              fam_delays=delays, fam_aligned=R.alignTD(delays, X))
    rng = np.random.default_rng(204)
    for name, N in (("rand2", 2), ("rand3", 3), ("rand24", 24)):
        true = rng.integers(-40, 40, N)
        iu = np.triu_indices(N - 1)
        ccm = np.full((N - 1, N - 1), np.nan)
        lgm = np.full((N - 1, N - 1), np.nan)
        ccm[iu] = rng.permutation(np.linspace(0.2, 0.97, len(iu[0])))
        noise = 3 * rng.integers(-2, 3, len(iu[0])) * (rng.random(len(iu[0])) < 0.4)
        lgm[iu] = (true[iu[1] + 1] - true[iu[0]]) * 3 + noise      # inconsistent lags on purpose
        idx, cols = range(N - 1), range(1, N)
        link, delays = R.getDelays(pd.DataFrame(ccm, index=idx, columns=cols),
                                   pd.DataFrame(lgm, index=idx, columns=cols))
        Xr = rng.standard_normal((N, 900))
        al.update({name + "_cc": ccm, name + "_lag": lgm, name + "_link": link, name + "_delays": delays,
                   name + "_X": Xr, name + "_aligned": R.alignTD(delays, Xr)})
    np.savez_compressed(os.path.join(OUT, "align_golden.npz"), **al)

    # ---- results tables (N4): SQLite wire format, duplicate removal, association
    #      (util.py:870-931, pandas_dbms.py:214-251, results.py:371-515) run by the reference itself
    import sqlite3
    import tempfile
    from detex_b200.synth import detection_table
    res = {}
    det, temkey = detection_table(401)
    with tempfile.TemporaryDirectory() as td:
        db = os.path.join(td, "ref.db")
        R.saveSQLite(det, db, "ss_df")
        R.saveSQLite(det.iloc[:7], db, "ss_df")          # append path
        con = sqlite3.connect(db)
        res["schema"] = np.array(con.execute("select sql from sqlite_master").fetchall()[0][0])
        res["nrows"] = np.array(con.execute("select count(*) from ss_df").fetchall()[0][0])
        con.close()
        dd = R.deleteDetDups(db, 1.0)
    num = [c for c in dd.columns if c not in ("Name", "Sta")]
    res["dedup_num"] = dd[num].to_numpy(dtype=float)
    res["dedup_cols"] = np.array(num)
    res["dedup_name"] = dd["Name"].to_numpy(dtype=str)
    res["dedup_sta"] = dd["Sta"].to_numpy(dtype=str)
    evnum = ["DSav", "DSmax", "NumStations", "DS_STALTA", "MSTAMPmin", "MSTAMPmax", "Mag", "ProEnMag"]
    for tag, (req, exc) in {"req2": (2, None), "req3": (3, None), "req3_exc": (3, 0.9),
                            "req3_excdict": (3, {"TA.M17A": 0.8}), "req1": (1, None)}.items():
        dt, at = R.associateDetections(dd, req, 1.0, temkey, exc)
        for kind, t in (("det", dt), ("auto", at)):
            res["%s_%s_num" % (tag, kind)] = t[evnum].to_numpy(dtype=float).reshape(len(t), len(evnum))
            res["%s_%s_event" % (tag, kind)] = t["Event"].to_numpy(dtype=str)
            res["%s_%s_nd" % (tag, kind)] = np.array([len(g) for g in t["Dets"]], dtype=np.int64)
            res["%s_%s_dets_stmp" % (tag, kind)] = (np.concatenate([g["STMP"].to_numpy(dtype=float) for g in t["Dets"]])
                                                    if len(t) else np.zeros(0))
    np.savez_compressed(os.path.join(OUT, "results_golden.npz"), **res)

    # ---- magnitude / SNR estimates (_estMag, detect.py:447-499)
    from oracle import detex_oracle as orc
    rng = np.random.default_rng(301)
    Nc, ns = 3, 120
    n = Nc * ns
    fam = synth.wavelet_basis(rng, ns, Nc, 3)
    ewf = np.array([rng.standard_normal(3) @ fam * rng.uniform(50, 500) + 5.0 * rng.standard_normal(n) for _ in range(7)])
    Umag = orc.svd_basis(ewf, select_value=0.9)["U"]
    mags = np.array([1.1, 2.3, -20.0, 0.7, 1.9, 2.8, 1.4])
    x = synth.multiplex(synth.bandpassed_noise(rng, 3000, nchan=Nc)) * 20.0
    trigs = [100, 1500, 2700]                         # early (post-event noise), middle, late
    for t in trigs:
        x[t * Nc:t * Nc + n] += 0.8 * ewf[t % 7]
    single = ewf[1][30:330].copy()
    mag = {"x": x, "Nc": Nc, "U": Umag, "ewf": ewf, "mags": mags, "trigs": np.array(trigs), "single": single}
    mag["sub_out"] = np.array([R.estMag(t, x, Nc, Umag, ewf, mags, True) for t in trigs], dtype=float)
    us = (single / np.linalg.norm(single))[None, :]
    mag["single_out"] = np.array([R.estMag(t, x, Nc, us, single[None, :], np.array([1.7]), False) for t in trigs], dtype=float)
    mag["nomag_out"] = np.array(R.estMag(1500, x, Nc, Umag, ewf, np.full(7, -99.0), True), dtype=float)
    np.savez_compressed(os.path.join(OUT, "mag_golden.npz"), **mag)

    # ---- multiplex (construct.py:928-987)
    chans = [np.arange(6.0), np.arange(6.0) + 10, np.arange(7.0) + 20]
    np.savez_compressed(os.path.join(OUT, "multiplex_golden.npz"), c0=chans[0], c1=chans[1], c2=chans[2],
                        out=R.multiplex(chans))
    for f in sorted(os.listdir(OUT)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(OUT, f)) // 1024, "KiB")


if __name__ == "__main__":
    main()
