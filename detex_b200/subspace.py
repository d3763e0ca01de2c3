"""subspace.py -- the arithmetic of `SubSpace.SVD` (reference detex/subspace.py:875-1054)
that PRODUCES the basis the GPU path consumes.  The matrices are tiny (events x n), so
this stays on LAPACK; it is kept here so the package offers the
reference's sequence createCluster -> SVD -> getFAS -> detex end to end.
"""
import numpy as np
import scipy.linalg
import scipy.stats


def svd_basis(aligned, selectCriteria=2, selectValue=0.9, normalize=False):
    """aligned: (events, n) trimmed aligned waveforms.  Returns dict(U (r,n), s, FracEnergy
    {'Average','Minimum'}, NumBasis) following subspace.py:875-905, 968-1013."""
    aligned = np.asarray(aligned, dtype=np.float64)
    arr = aligned - aligned.mean(axis=1, keepdims=True)          # _trimGroups :932
    if normalize:
        arr = np.array([x / np.linalg.norm(x) for x in arr])
    U, s, Vh = scipy.linalg.svd(arr.T, full_matrices=False)       # :890
    cum = []
    for w in aligned:                                             # _getFracEnergy :968-997
        rep = np.insert(np.square(np.dot(U.T, w) / scipy.linalg.norm(w)), 0, 0)
        cum.append(np.cumsum(rep))
    avg = np.average(cum, axis=0)
    mn = np.min(cum, axis=0)
    if selectCriteria in (1, 2, 3):                               # _getUsedBasis :999-1013
        avg[-1] = 1.00
        ndim = int(np.argmax(avg >= selectValue))
    elif selectCriteria == 4:
        ndim = int(selectValue) + 1
    else:
        raise Exception('selectCriteria of %s is not supported' % selectCriteria)
    return dict(U=np.ascontiguousarray(U[:, :ndim].T), s=s, Ufull=U, cum=cum,
                FracEnergy={'Average': avg, 'Minimum': mn}, NumBasis=ndim)


def validate_cluster(aligned, ccreq, engine=None):
    """`SubSpace.validateClusters` for one cluster (subspace.py:738-773): aligned is the
    (events, n) array of aligned + trimmed members in `row.Events` order.  Returns the indices
    of the events that fail (max CC with every LATER member < ccreq), in the order the
    reference would remove them.  The pairwise zero-lag coefficients come from the GPU."""
    from .detect import default_engine
    eng = engine or default_engine()
    cc = eng.corr_zero_lag(aligned)
    N = cc.shape[0]
    return [i for i in range(N - 1) if np.max(cc[i, i + 1:]) < ccreq]


def single_basis(mptd):
    """Singleton template: unit norm, not demeaned (detect.py:356-357; fas.py:144)."""
    x = np.asarray(mptd, dtype=np.float64)
    return (x / np.linalg.norm(x))[None, :]


def _approxThld(beta_a, beta_b, target, numint=1000, numloops=3, backupThreshold=None):
    """subspace.py:1110-1140."""
    startVal, stopVal = 0, 1
    for _ in range(numloops):
        Xs = np.linspace(startVal, stopVal, numint)
        pfs = scipy.stats.beta.sf(Xs, beta_a, beta_b)
        minind = int(np.abs(pfs - target).argmin())
        if minind == 0 or minind == numint - 1:
            if backupThreshold is None:
                raise ValueError('Grid search for threshold failing, set it manually')
            return backupThreshold, target
        bestPf, bestX = pfs[minind], Xs[minind]
        startVal, stopVal = Xs[minind - 1], Xs[minind + 1]
    return bestX, bestPf


def threshold_from_fas(fas, Pf=1e-12, backupThreshold=None):
    """subspace.py:1031-1047: threshold from the fitted beta distribution."""
    beta_a, beta_b = fas['betadist'][0:2]
    th = scipy.stats.beta.isf(Pf, beta_a, beta_b, 0, 1)
    if th > .9:
        th, _ = _approxThld(beta_a, beta_b, Pf, 1000, 3, backupThreshold)
    return float(th)
