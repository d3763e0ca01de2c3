"""The one workload the reference publishes a timing for: `createCluster()` on the tutorial's
222-event key -- 213 / 214 events used, 2 stations x 1 channel (EHZ), trim 10 + 120 s at 100 Hz
(n = 13 000 samples, 13 001 lags per pair, ~22.7 k pairs per station): wall 2 min 36 s / 2 min 24 s /
136.88 s on an unnamed desktop, file I/O and ObsPy filtering included
(docs/tutorial/Advanced/Clustering/clustering.md:751-775, 823-844, 1411-1417).

Same shapes on synthetic data through `detex_b200.workflow.createCluster` (device band-pass, CCX on
the tensor-core engine with float64 re-scoring, SciPy linkage); data already in memory, so the
reference's file reads are not part of our time.  Prints one JSON line.
    python experiments/createcluster_published_shape.py
"""
import json
import sys
import time

import numpy as np
import pandas as pd

sys.path.insert(0, ".")
from detex_b200 import workflow  # noqa: E402
from detex_b200.engine import Engine  # noqa: E402


def main():
    rng = np.random.default_rng(2140)
    sr, ns, nev = 100.0, 13000, 214
    stations = ["TA.M17A", "TA.M18A"]
    names = ["ev%03d" % i for i in range(nev)]
    temkey = pd.DataFrame({"NAME": names, "TIME": 1.2e9 + 3600.0 * np.arange(nev), "LAT": 40.0, "LON": -110.0,
                           "DEPTH": 5.0, "MAG": 1.5})
    stakey = pd.DataFrame({"NETWORK": ["TA", "TA"], "STATION": ["M17A", "M18A"], "STARTTIME": 1.3e9,
                           "ENDTIME": 1.3e9 + 86400, "LAT": 40.5, "LON": -110.5, "ELEVATION": 1500.0,
                           "CHANNELS": "EHZ"})
    events = {}
    for sta in stations:
        fams = [rng.standard_normal(4000) * np.hanning(4000) for _ in range(8)]
        events[sta] = {}
        for i, nm in enumerate(names):
            x = 0.3 * rng.standard_normal(ns) + 5.0
            s0 = 2000 + int(rng.integers(-100, 101))
            x[s0:s0 + 4000] += fams[i % 8]
            events[sta][nm] = ([x], 1.2e9 + 3600.0 * i - 10.0)
    fetcher = workflow.ArrayFetcher(events, {}, sr=sr, channels=("EHZ",))
    eng = Engine(0)
    out = {}
    for rep in range(2):
        t0 = time.perf_counter()
        cl = workflow.createCluster(CCreq=0.5, fetch_arg=fetcher, filt=[1, 10, 2, True], stationKey=stakey,
                                    templateKey=temkey, trim=[10, 120], saveclust=False, engine=eng)
        out["seconds_rep%d" % rep] = time.perf_counter() - t0
    # spot check of the tensor-core CCX against the float64 kernel on one station
    row = cl._TRDF.iloc[0]
    X = np.array([row.MPtd[e] for e in row.Events[:24]])
    c64, l64, _ = eng.ccx(X, 1, engine="fp64")
    ctc, ltc, _ = eng.ccx(X, 1, engine="tcgen05")
    iu = np.triu_indices(len(X), 1)
    out.update(config="createCluster, %d events x 2 stations x 1 channel x 130 s @ 100 Hz (n = %d)" % (nev, ns),
               pairs_per_station=nev * (nev - 1) // 2, lags_per_pair=2 * ns - 1 - 2 * (ns // 2 - 1),
               clusters=[len(c.clusts) for c in cl.clusters], singles=[len(c.singles) for c in cl.clusters],
               max_cc_diff_vs_fp64=float(np.abs(ctc[iu] - c64[iu]).max()),
               lags_equal=bool(np.array_equal(ltc[iu], l64[iu])),
               published_reference_seconds=[156.0, 144.0, 136.88],
               published_source="docs/tutorial/Advanced/Clustering/clustering.md:751-775, 823-844, 1411-1417 "
                                "(unnamed desktop, includes file I/O)")
    print(json.dumps(out))
    eng.close()


if __name__ == "__main__":
    main()
