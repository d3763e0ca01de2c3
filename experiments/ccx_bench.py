"""CCX throughput on one GPU: BASELINE configs[2] shape (3 ch x 10 s x 100 Hz, 1001 lags)."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from detex_b200 import synth
from detex_b200.engine import Engine

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
eng = Engine(0)
X = synth.event_families(3003, max(1, N // 64), 64, 1000, 3, max_shift=100)[:N]
print("events", X.shape)
for engine in ("tcgen05", "fp64"):
    if engine == "fp64" and N > 2048:
        continue
    for rep in range(2):
        t0 = time.time()
        cc, lag, sub = eng.ccx(X, 3, engine=engine)
        dt = time.time() - t0
    pairs = N * (N - 1) // 2
    print("%s: %.3f s  %.3e pairs/s  %.3e pair-lags/s  useful %.1f TFLOP/s" % (
        engine, dt, pairs / dt, pairs * 1001 / dt, pairs * 1001 * 2 * 3000 / dt / 1e12))
    if engine == "tcgen05":
        keep = (cc.copy(), lag.copy())
    else:
        iu = np.triu_indices(N, 1)
        print("   engines agree: cc %.2e, lag mismatches %d" % (np.abs(cc[iu] - keep[0][iu]).max(), int((lag[iu] != keep[1][iu]).sum())))
