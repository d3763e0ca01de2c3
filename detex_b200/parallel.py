"""parallel.py -- multi-GPU sharding of the hot path (one process per GPU).

The path shards by independent units: detection / FAS by continuous-data
chunk, CCX by row block.  There is NO data-path collective; NCCL (or gloo on CPU for the
tests) is used once at the end to gather the variable-length trigger lists, sum the
per-subspace histograms / FAS sufficient statistics and gather the CCX row blocks.
"""
import numpy as np
import torch
import torch.distributed as dist


def shard_range(nunits, rank, world):
    """Contiguous, balanced [lo, hi) of `nunits` independent units for `rank`."""
    base, rem = divmod(nunits, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def ccx_row_blocks(N, world):
    """Row ranges [b0, b1) of the upper-triangular pair matrix with ~equal pair counts.
    Row b owns N-1-b pairs."""
    total = N * (N - 1) // 2
    bounds = [0]
    acc = 0
    b = 0
    for r in range(1, world):
        target = total * r / float(world)
        while b < N - 1 and acc + (N - 1 - b) <= target:
            acc += N - 1 - b
            b += 1
        bounds.append(b)
    bounds.append(N - 1)
    return [(bounds[i], bounds[i + 1]) for i in range(world)]


def ccx_deal_rows(N, world):
    """Template rows b = 0..N-2 of the upper-triangular pair matrix dealt to `world` ranks back and
    forth (rank k gets the rows with b mod 2W in {k, 2W-1-k}): row b owns N-1-b pairs, so each such
    couple owns the same number and every rank ends up with the same number of rows (within 2) AND
    of pairs (within ~2N/W), whatever N.  Returns a list of ascending int32 arrays."""
    b = np.arange(N - 1, dtype=np.int64)
    j = b % (2 * world)
    owner = np.where(j < world, j, 2 * world - 1 - j)
    return [b[owner == k].astype(np.int32) for k in range(world)]


def ccx_slot_rows(N, world):
    """Row held by every slot of the all-gathered dense buffer: ranks' row lists padded with -1 to
    the longest one, back to back.  Returns (slot_rows int32 [world * nmax], nmax)."""
    deal = ccx_deal_rows(N, world)
    nmax = max(len(r) for r in deal)
    slots = np.full((world, nmax), -1, dtype=np.int32)
    for k, r in enumerate(deal):
        slots[k, :len(r)] = r
    return slots.reshape(-1), nmax


def pack_condensed(cc, lag, sub, slot_rows, N):
    """Host statement of dtx_ccx_pack (tests / CPU plumbing): dense slots [nslots][N] -> condensed."""
    slot_of = np.full(N, -1, dtype=np.int64)
    for s, b in enumerate(slot_rows):
        if b >= 0:
            slot_of[b] = s
    iu = np.triu_indices(N, 1)
    src = slot_of[iu[0]]
    return cc[src, iu[1]], lag[src, iu[1]], sub[src, iu[1]]


class CcxHostBuffer(object):
    """Condensed (cc float64, lag int32, subsamp float64) arrays of ONE N-event pair matrix in host memory that
    every rank of the box has mapped (a file under /dev/shm, unlinked as soon as all ranks hold it) and page-locked
    for its GPU: `ccx_sharded(..., host=buf)` lets each GPU write the rows it computed straight to their places
    over its own PCIe link -- no gather, no 20 B x N(N-1)/2 funnel through one link -- and every rank (the one that
    runs `linkage`, construct.py:152-157, included) reads the whole matrix from `buf.cc / buf.lag / buf.sub`.
    Create once per problem size (page-locking costs ~0.1 s per GB), reuse for every call, `close()` at the end."""

    def __init__(self, eng, N, shm_dir="/dev/shm"):
        import os
        self.eng, self.N = eng, int(N)
        npair = self.N * (self.N - 1) // 2
        off_lag = 8 * npair
        off_sub = off_lag + 8 * ((4 * npair + 7) // 8)
        total = max(8, off_sub + 8 * npair)
        world = _world()
        self._registered = False
        if world == 1:
            self._base = eng.pinned_empty((total,), np.uint8) if hasattr(eng, "pinned_empty") else np.zeros(total, np.uint8)
        else:
            rank = dist.get_rank()
            path = [None]
            if rank == 0:
                try:        # a tmpfs smaller than the matrix would only fail later, with SIGBUS on first touch
                    st = os.statvfs(shm_dir)
                    if st.f_bavail * st.f_frsize > total + (64 << 20):
                        path[0] = os.path.join(shm_dir, "detex_b200_ccx_%d_%x" % (os.getpid(), id(self)))
                        self._base = np.memmap(path[0], dtype=np.uint8, mode="w+", shape=(total,))
                except OSError:
                    path[0] = None
            dist.broadcast_object_list(path, src=0)
            if path[0] is None:             # every rank raises: callers fall back to the NCCL gather
                raise RuntimeError("CcxHostBuffer: %s cannot hold %d bytes" % (shm_dir, total))
            if rank != 0:
                self._base = np.memmap(path[0], dtype=np.uint8, mode="r+", shape=(total,))
            dist.barrier()
            if rank == 0:
                os.unlink(path[0])          # the mappings keep it alive; nothing to clean up after a crash
            ok = 1
            if hasattr(eng, "host_register"):
                try:
                    eng.host_register(self._base)
                    self._registered = True
                except Exception:
                    ok = 0
            # all ranks succeed or all give up (a rank that raised alone would leave the others in a barrier)
            flag = torch.tensor([ok], dtype=torch.int32, device=_dev())
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            if int(flag[0]) == 0:
                self.close()
                raise RuntimeError("CcxHostBuffer: page-locking the shared segment failed on a rank")
        self.cc = self._base[:8 * npair].view(np.float64)
        self.lag = self._base[off_lag:off_lag + 4 * npair].view(np.int32)
        self.sub = self._base[off_sub:off_sub + 8 * npair].view(np.float64)

    def arrays(self):
        return self.cc, self.lag, self.sub

    def close(self):
        if getattr(self, "_registered", False):
            self.eng.host_unregister(self._base)
            self._registered = False
        self.cc = self.lag = self.sub = self._base = None


def ccx_sharded(eng, X, Nc, engine="tcgen05", out=None, root=None, host=None):
    """The whole CCX matrix of one station over all ranks (BASELINE configs[2]): every
    rank computes its dealt rows with the results left in HBM.  Then either
      * ONE all-gather of the equal-sized dense blocks over NCCL / NVLink, after which the upper triangle is
        packed into SciPy's condensed order and copied to the host -- on every rank (root=None), or only on
        rank `root` (the one that runs `linkage`, construct.py:152-157; the other ranks return None); or
      * host = CcxHostBuffer: no collective at all -- each GPU writes its rows to their places in the shared,
        page-locked condensed arrays over its own PCIe link (`dtx_ccx_pack_rows`), one barrier, and every
        rank returns views of the same host memory.
    On NCCL each rank uploads only its 1/world slice of X and the slices are all-gathered over NVLink.
    Returns (cc, lag, subsamp) condensed; identical on every rank and to the single-GPU result."""
    world = _world()
    if world == 1:
        return eng.ccx_condensed(X, Nc, engine=engine, out=host.arrays() if host is not None else out)
    X = np.asarray(X)
    N = X.shape[0]
    rank = dist.get_rank()
    dev = _dev()
    slot_rows, nmax = ccx_slot_rows(N, world)
    mine = slot_rows.reshape(world, nmax)[rank]
    mine = mine[mine >= 0]
    d_cc = torch.zeros((nmax, N), dtype=torch.float64, device=dev)
    d_lag = torch.zeros((nmax, N), dtype=torch.int32, device=dev)
    d_sub = torch.zeros((nmax, N), dtype=torch.float64, device=dev)
    xg = None
    if dev.type == "cuda" and X.dtype == np.float64:
        # 1/world of the waveforms over this GPU's PCIe link, the rest over NVLink
        per = (N + world - 1) // world
        n = X.shape[1]
        xg = torch.empty((world * per, n), dtype=torch.float64, device=dev)
        part = torch.zeros((per, n), dtype=torch.float64, device=dev)
        lo, hi = min(N, rank * per), min(N, (rank + 1) * per)
        if hi > lo:
            part[:hi - lo].copy_(torch.from_numpy(X[lo:hi]))
        dist.all_gather_into_tensor(xg, part)
        torch.cuda.current_stream().synchronize()      # the engine runs on its own stream
    if len(mine):
        if xg is not None:
            eng.ccx_device(None, Nc, mine, d_cc.data_ptr(), d_lag.data_ptr(), d_sub.data_ptr(), engine=engine,
                           x_device_ptr=xg.data_ptr(), shape=(N, X.shape[1]))
        else:
            eng.ccx_device(X, Nc, mine, d_cc.data_ptr(), d_lag.data_ptr(), d_sub.data_ptr(), engine=engine)
    if host is not None:
        eng.ccx_pack_rows(d_cc.data_ptr(), d_lag.data_ptr(), d_sub.data_ptr(), mine, N, host.arrays())   # synchronises
        del xg
        dist.barrier()
        return host.arrays()
    if dev.type == "cuda":
        eng.sync()                                     # the engine may run on another stream than the collective
    del xg
    g_cc = torch.empty((world * nmax, N), dtype=torch.float64, device=dev)
    g_lag = torch.empty((world * nmax, N), dtype=torch.int32, device=dev)
    g_sub = torch.empty((world * nmax, N), dtype=torch.float64, device=dev)
    dist.all_gather_into_tensor(g_cc, d_cc)
    dist.all_gather_into_tensor(g_lag, d_lag)
    dist.all_gather_into_tensor(g_sub, d_sub)
    if dev.type == "cuda":
        torch.cuda.current_stream().synchronize()      # the engine may run on another stream
    res = None
    if root is None or root == rank:
        res = eng.ccx_pack(g_cc.data_ptr(), g_lag.data_ptr(), g_sub.data_ptr(), slot_rows, N, out=out)
    del g_cc, g_lag, g_sub
    return res


def _dev():
    return torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")


def _world():
    return dist.get_world_size() if dist.is_initialized() else 1


def allreduce_sum(arr):
    """Sum a NumPy array (int64 histogram / float64 statistics) over ranks."""
    if _world() == 1:
        return arr
    t = torch.from_numpy(np.ascontiguousarray(arr)).to(_dev())
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy()


def allreduce_max(arr):
    if _world() == 1:
        return arr
    t = torch.from_numpy(np.ascontiguousarray(arr)).to(_dev())
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.cpu().numpy()


def gather_records(rec):
    """All-gather a 1-D structured NumPy array of variable length (trigger / candidate
    records): counts first, then the padded byte payload.  Returns the concatenation in
    rank order on every rank."""
    world = _world()
    if world == 1:
        return rec
    dev = _dev()
    n = torch.tensor([len(rec)], dtype=torch.int64, device=dev)
    counts = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(counts, n)
    counts = [int(c.item()) for c in counts]
    mx = max(counts)
    item = rec.dtype.itemsize
    buf = np.zeros(mx * item, dtype=np.uint8)
    buf[:len(rec) * item] = np.frombuffer(np.ascontiguousarray(rec).tobytes(), dtype=np.uint8)
    t = torch.from_numpy(buf).to(dev)
    outs = [torch.zeros(mx * item, dtype=torch.uint8, device=dev) for _ in range(world)]
    if mx > 0:
        dist.all_gather(outs, t)
    parts = [np.frombuffer(o.cpu().numpy().tobytes()[:c * item], dtype=rec.dtype) for o, c in zip(outs, counts)]
    return np.concatenate(parts) if parts else rec


def gather_row_blocks(block, blocks, N):
    """All-gather CCX row blocks (rows_i, N) into the full (N-1, N) matrix on every rank."""
    world = _world()
    if world == 1:
        return block
    dev = _dev()
    mx = max(b1 - b0 for b0, b1 in blocks)
    pad = np.zeros((mx, N), dtype=block.dtype)
    pad[:block.shape[0]] = block
    t = torch.from_numpy(pad).to(dev)
    outs = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(outs, t)
    return np.concatenate([o.cpu().numpy()[:b1 - b0] for o, (b0, b1) in zip(outs, blocks)], axis=0)
