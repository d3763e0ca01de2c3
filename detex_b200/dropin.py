"""The reference-side binding of INTEGRATION.md section 1, executable.

`install(detex)` rebinds the three private hot-path callables of an already imported
reference package -- `detex.detect._SSDetex._MPXDS` (detect.py:559-578),
`detex.fas._MPXSSCorr` (fas.py:120-134), `detex.construct._CCX2` (construct.py:425-466) --
to the GPU mirrors of this package, and, with `batched=True`, also
`detex.construct._makeDFcclags` (construct.py:369-394) so that `createCluster` computes
the whole CCX matrix in one call instead of one pair at a time.  Everything else of the
reference (its classes, loops, tables, pickles) keeps running unchanged on top of them.
`uninstall(detex)` restores the originals.

The reference package is the caller's: this module never imports it.  There is no CPU
fallback -- the rebound functions raise if the CUDA library is missing.
"""
import numpy as np

from . import construct as _construct
from . import detect as _detect
from . import fas as _fas

_SAVED = "_detex_b200_saved"


def install(detex, engine=None, batched=False, kernel="tcgen05"):
    """Rebind the hot-path callables of the reference package `detex`.  Returns the dict of
    the replaced originals (also kept on the package for `uninstall`)."""
    import importlib
    rdet = importlib.import_module(detex.__name__ + ".detect")
    rfas = importlib.import_module(detex.__name__ + ".fas")
    rcon = importlib.import_module(detex.__name__ + ".construct")
    if getattr(detex, _SAVED, None) is not None:
        uninstall(detex)
    saved = {"_MPXDS": rdet._SSDetex._MPXDS, "_MPXSSCorr": rfas._MPXSSCorr, "_CCX2": rcon._CCX2,
             "_makeDFcclags": rcon._makeDFcclags}

    def _MPXDS(self, MPcon, reqlen, ssTD, ssFD, Nc, MPconFD):
        return _detect._MPXDS(MPcon, reqlen, ssTD, ssFD, Nc, MPconFD, engine=engine, kernel=kernel)

    def _MPXSSCorr(MPcon, reqlen, ssArrayTD, ssArrayFD, Nc):
        return _fas._MPXSSCorr(MPcon, reqlen, ssArrayTD, ssArrayFD, Nc, engine=engine, kernel=kernel)

    def _CCX2(mpfd1, mpfd2, mptd1, mptd2, Nc1, Nc2):
        return _construct._CCX2(mpfd1, mpfd2, np.asarray(mptd1), np.asarray(mptd2), Nc1, Nc2, engine=engine)

    def _makeDFcclags(eventList, row):
        return _construct._makeDFcclags(eventList, row, engine=engine, kernel=kernel)

    rdet._SSDetex._MPXDS = _MPXDS
    rfas._MPXSSCorr = _MPXSSCorr
    rcon._CCX2 = _CCX2
    if batched:
        rcon._makeDFcclags = _makeDFcclags
    setattr(detex, _SAVED, saved)
    return saved


def uninstall(detex):
    import importlib
    saved = getattr(detex, _SAVED, None)
    if saved is None:
        return
    rdet = importlib.import_module(detex.__name__ + ".detect")
    rfas = importlib.import_module(detex.__name__ + ".fas")
    rcon = importlib.import_module(detex.__name__ + ".construct")
    rdet._SSDetex._MPXDS = saved["_MPXDS"]
    rfas._MPXSSCorr = saved["_MPXSSCorr"]
    rcon._CCX2 = saved["_CCX2"]
    rcon._makeDFcclags = saved["_makeDFcclags"]
    setattr(detex, _SAVED, None)
