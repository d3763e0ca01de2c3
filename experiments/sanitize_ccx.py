"""compute-sanitizer target: smoke() plus a CCX call whose pairs carry several candidate lags (noisy, weakly
correlated events: the one-pass multi-candidate path of the ring kernel) and an odd number of signals (a self-paired
DUAL item)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g  # noqa: E402
from detex_b200 import synth  # noqa: E402
from detex_b200.engine import Engine  # noqa: E402

g.smoke()
eng = Engine(0)
X = synth.event_families(11, 4, 6, 120, 3, max_shift=20, noise=1.5)        # 24 events, n = 360, noisy
cc, lag, sub = eng.ccx_condensed(X, 3, engine="tcgen05")
c64, l64, s64 = eng.ccx(X, 3, engine="fp64")
iu = np.triu_indices(len(X), 1)
assert np.array_equal(lag, l64[iu]) and np.abs(cc - c64[iu]).max() < 1e-12
eng.close()
print("sanitize target ok")
