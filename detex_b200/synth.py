"""Seeded synthetic inputs with the shapes of BASELINE.json's configs.

Host-side NumPy only; used by tests/, bench.py and __graft_entry__.smoke().  No network,
no waveform files: the reference bundles none either.
"""
import numpy as np
import scipy.signal


def bandpassed_noise(rng, nsamp, sr=100.0, band=(1.0, 10.0), nchan=3):
    """(nchan, nsamp) unit-variance Gaussian noise, zero-phase Butterworth band-passed --
    the spectral shape `_applyFilter` (reference detex/construct.py:990-1030) leaves."""
    x = rng.standard_normal((nchan, nsamp + 2000))
    sos = scipy.signal.butter(2, band, btype="band", fs=sr, output="sos")
    y = scipy.signal.sosfiltfilt(sos, x, axis=1)[:, 1000:1000 + nsamp]
    return y / y.std(axis=1, keepdims=True)


def multiplex(chans):
    """construct.multiplex layout: [c0[0], c1[0], c2[0], c0[1], ...]."""
    C = np.asarray(chans)
    return np.ascontiguousarray(C.T).reshape(-1)


def random_basis(rng, n, r, zero_mean=False):
    """(r, n) orthonormal rows (QR of a Gaussian matrix), non-zero mean on purpose."""
    q, _ = np.linalg.qr(rng.standard_normal((n, r)))
    U = q.T.copy()
    if zero_mean:
        U -= U.mean(axis=1, keepdims=True)
        q, _ = np.linalg.qr(U.T)
        U = q.T.copy()
    return U


def wavelet_basis(rng, ns, Nc, r, sr=100.0):
    """Basis whose span contains band-limited wavelets (so planted events score high)."""
    n = ns * Nc
    W = bandpassed_noise(rng, ns, sr=sr, nchan=Nc * r).reshape(r, Nc, ns)
    taper = np.hanning(ns)
    W = W * taper[None, None, :]
    rows = np.array([multiplex(w) for w in W])
    q, _ = np.linalg.qr(rows.T)
    return q.T.copy().reshape(r, n)


def plant(chunk_mux, template_mux, t, Nc, amp):
    """Add amp * template at channel-aligned lag t of a multiplexed chunk (in place)."""
    n = len(template_mux)
    chunk_mux[t * Nc: t * Nc + n] += amp * template_mux
    return chunk_mux


def detection_case(seed, nchunks, Ls, ns, Nc, ranks, planted=0, sr=100.0, dtype=np.float64):
    """Chunks + bases for a detection run.  ranks: list of subspace ranks."""
    rng = np.random.default_rng(seed)
    n = ns * Nc
    bases = [random_basis(rng, n, r) if i % 2 else wavelet_basis(rng, ns, Nc, r, sr)
             for i, r in enumerate(ranks)]
    chunks = []
    truth = []
    for c in range(nchunks):
        x = multiplex(bandpassed_noise(rng, Ls, sr=sr, nchan=Nc))
        for _ in range(planted):
            s = int(rng.integers(0, len(bases)))
            t = int(rng.integers(0, Ls - ns))
            coef = rng.standard_normal(bases[s].shape[0])
            tem = coef @ bases[s]
            amp = float(rng.uniform(3.0, 8.0)) * np.sqrt(n) / np.linalg.norm(tem)
            plant(x, tem, t, Nc, amp)
            truth.append((c, s, t))
        chunks.append(x.astype(dtype))
    return chunks, bases, truth


def event_families(seed, nfam, per_fam, ns, Nc, sr=100.0, max_shift=100, noise=0.5):
    """(nfam*per_fam, ns*Nc) multiplexed event waveforms: family wavelet shifted by a few
    samples plus noise (config 3 shape: CC within family ~0.6-0.95)."""
    rng = np.random.default_rng(seed)
    out = []
    for f in range(nfam):
        base = bandpassed_noise(rng, ns + 2 * max_shift, sr=sr, nchan=Nc)
        env = np.exp(-0.5 * ((np.arange(ns + 2 * max_shift) - (ns / 2 + max_shift)) / (ns / 6.0)) ** 2)
        base = base * env[None, :]
        for m in range(per_fam):
            sh = int(rng.integers(-max_shift, max_shift + 1))
            w = base[:, max_shift + sh: max_shift + sh + ns]
            w = w + noise * w.std() * bandpassed_noise(rng, ns, sr=sr, nchan=Nc)
            out.append(multiplex(w))
    return np.array(out)


def detection_table(seed, nev=80, stas=("TA.M17A", "TA.M18A", "TA.N17A"), t0=1.4e9, n_templates=6):
    """A shuffled multi-station detections table (columns of detect.py:397-398) plus a template
    key whose origins coincide with some of the events: input of the N4 results path."""
    import pandas as pd
    rng = np.random.default_rng(seed)
    rows, times = [], []
    for e in range(nev):
        te = t0 + e * 400.0 + float(rng.uniform(0, 50))
        times.append(te)
        for sta in stas:
            if rng.random() < 0.7:
                for _ in range(int(rng.integers(1, 3))):          # sometimes two detectors fire
                    t = te + float(rng.uniform(-1.5, 1.5))
                    rows.append([float(rng.uniform(0.3, 0.95)), float(rng.uniform(3, 9)), t + 3.0,
                                 "SS%d" % int(rng.integers(0, 3)), sta, t, t + float(rng.uniform(1, 4)),
                                 float(rng.choice([np.nan, 1.5, 2.0, 2.7])), float(rng.uniform(1, 5)),
                                 float(rng.choice([np.nan, 1.1, 1.9]))])
    cols = ['DS', 'DS_STALTA', 'STMP', 'Name', 'Sta', 'MSTAMPmin', 'MSTAMPmax', 'Mag', 'SNR', 'ProEnMag']
    det = pd.DataFrame(rows, columns=cols)
    det = det.iloc[rng.permutation(len(det))].reset_index(drop=True)
    pick = np.sort(rng.choice(nev, size=n_templates, replace=False))
    temkey = pd.DataFrame({"NAME": ["ev%02d" % i for i in range(n_templates)][::-1],   # not in time order
                           "TIME": [times[k] + 1.0 for k in pick], "MAG": np.linspace(1, 3, n_templates)})
    temkey["STMP"] = temkey["TIME"].astype(float)
    return det, temkey


def workflow_case(seed, stations=("TA.M17A", "TA.M18A"), nfam=3, per_fam=5, nsingles=2, sr=40.0, Nc=3,
                  ev_seconds=20.0, nchunks=6, chunk_seconds=300.0, t0=1.3e9, pick_after=5.0):
    """A small Case1-shaped network for the whole createCluster -> ... -> detex sequence
    (BASELINE configs[0] stand-in; the reference's Case1 needs IRIS downloads): RAW per-channel
    event windows and continuous chunks (coloured noise with an offset and a trend, so that the
    detrend + band-pass step matters), families of similar events, planted family members.

    Returns dict(events, continuous, temkey, stakey, picks, planted, sr, Nc) in the shapes
    `workflow.ArrayFetcher` / `createCluster` take."""
    import pandas as pd
    rng = np.random.default_rng(seed)
    ns = int(ev_seconds * sr)
    Ls = int(chunk_seconds * sr)
    nev = nfam * per_fam + nsingles
    names = ["2010-01-%02dT%02d-00-00" % (1 + i // 20, i % 20) for i in range(nev)]
    origins = [t0 - 86400.0 * 30 + 3600.0 * i for i in range(nev)]
    temkey = pd.DataFrame({"NAME": names, "TIME": origins, "LAT": 40.0 + 0.01 * np.arange(nev),
                           "LON": -110.0 + 0.01 * np.arange(nev), "DEPTH": 5.0,
                           "MAG": np.round(rng.uniform(1.0, 3.0, nev), 2)})
    stakey = pd.DataFrame({"NETWORK": [s.split('.')[0] for s in stations],
                           "STATION": [s.split('.')[1] for s in stations],
                           "STARTTIME": t0, "ENDTIME": t0 + nchunks * chunk_seconds,
                           "LAT": 40.5, "LON": -110.5, "ELEVATION": 1500.0, "CHANNELS": "BHE-BHN-BHZ"})
    events, continuous, picks, planted = {}, {}, [], []
    # an event is seen by every station at (nearly) the same time: one origin per chunk, small moveout
    plant_t = [int(rng.integers(int(20 * sr), Ls - int(40 * sr))) for _ in range(nchunks)]

    def raw_noise(nsamp, amp):
        x = amp * rng.standard_normal((Nc, nsamp))
        x += 3.0 * amp * rng.standard_normal((Nc, 1))                       # offset
        x += amp * rng.standard_normal((Nc, 1)) * np.linspace(-1, 1, nsamp)  # trend
        return x

    for sta in stations:
        fams = []
        for f in range(nfam):
            w = bandpassed_noise(rng, ns // 2, sr=sr, band=(1.5, 8.0), nchan=Nc)
            fams.append(w * np.hanning(ns // 2)[None, :])
        events[sta] = {}
        for i, name in enumerate(names):
            fam = i // per_fam if i < nfam * per_fam else -1
            x = raw_noise(ns, 0.15)
            if fam >= 0:
                sh = int(rng.integers(-int(0.5 * sr), int(0.5 * sr) + 1))
                a = float(rng.uniform(0.8, 1.5))
                s0 = ns // 4 + sh
                x[:, s0:s0 + ns // 2] += a * fams[fam]
            else:
                w = bandpassed_noise(rng, ns // 2, sr=sr, band=(1.5, 8.0), nchan=Nc) * np.hanning(ns // 2)[None, :]
                x[:, ns // 4: ns // 4 + ns // 2] += w
            start = origins[i] - 2.0
            events[sta][name] = ([x[c].copy() for c in range(Nc)], start)
            picks.append({"TimeStamp": start + pick_after, "Station": sta, "Event": name, "Phase": "P"})
        continuous[sta] = []
        for c in range(nchunks):
            x = raw_noise(Ls, 0.15)
            if c < nfam + 1:
                fam = c % nfam
                t = plant_t[c] + int(rng.integers(0, int(1.0 * sr)))
                x[:, t:t + ns // 2] += float(rng.uniform(0.8, 1.2)) * fams[fam]
                planted.append((sta, c, fam, t / sr))
            continuous[sta].append(([x[k].copy() for k in range(Nc)], t0 + c * chunk_seconds))
    return dict(events=events, continuous=continuous, temkey=temkey, stakey=stakey,
                picks=pd.DataFrame(picks), planted=planted, sr=sr, Nc=Nc, chunk_seconds=chunk_seconds)
