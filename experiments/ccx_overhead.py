"""Where does the CCX call spend its wall time?  Launch counts and wall time for several N."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from detex_b200 import synth
from detex_b200.engine import Engine

eng = Engine(0)
for N in (256, 1024, 2048, 4096):
    X = synth.event_families(3003, max(1, N // 64), 64, 1000, 3, max_shift=100)[:N]
    for rep in range(3):
        l0 = eng.launch_count()
        t0 = time.perf_counter()
        cc, lag, sub = eng.ccx(X, 3, engine="tcgen05")
        dt = time.perf_counter() - t0
        print("N %5d rep %d  %.3f s  launches %d" % (N, rep, dt, eng.launch_count() - l0), flush=True)
    t0 = time.perf_counter()
    a = np.zeros((N, N)); b = np.zeros((N, N)); c = np.zeros((N, N), dtype=np.int32)
    a[:] = 1; b[:] = 1; c[:] = 1
    print("   host alloc+touch of the three outputs: %.3f s" % (time.perf_counter() - t0))
