"""CPU model check of the multi-GPU step's SCHEDULE (slab.slab_step): no GPU, no kernels.

`slab_step` only sequences library calls, stream/event dependencies, flag signals/waits and
collectives.  Here every one of those is replaced by a recorder (a fake `torch.cuda`, fake ranks, a
fake communicator), the per-rank op queues of P ranks x 2 streams x several steps are then executed
"concurrently" by a randomised scheduler with the semantics of the real things:

* ops of one stream run in order; `wait_event` blocks until the event was recorded;
* `signal(slot)` stores this rank's epoch into every rank's flag word, `wait(slot)` blocks until all P
  words of the slot carry the waiter's epoch (pm_slab_peer_signal / pm_slab_peer_wait);
* a collective completes when every participant has reached it, and one rank's collectives complete
  in the order the host issued them (one NCCL communicator).

The run must drain (no deadlock) under every interleaving tried, and the data hazards of the
distributed transform must be ordered in all of them:

  H1  a rank's z pass of chunk c runs after EVERY rank has delivered its block of chunk c;
  H2  a rank fetches chunk c back only after EVERY rank has finished its z pass of chunk c;
  H3  no rank stores step n+1's chunk c into the z-pass arrays before EVERY rank has fetched step n's.

and, with the ghost planes pushed through peer memory (ghosts="peer", experimental):

  H4  ghost_add runs after rank-1 delivered this step's density ghost plane;
  H5  the gather runs after both neighbours delivered this step's potential ghost planes;
  H6  a neighbour's density ghost (and its two upper potential planes) land in a rank's phi buffer only
      after that rank's gather of the previous step has read it;
  H7  rank-1's potential plane lands on phi plane 0 only after this step's ghost_add has read the density
      ghost from there.

and, with the migration through peer memory (migrate="peer"):

  H8  a rank unpacks step n's arrivals only after EVERY rank has stored its step-n records;
  H9  no rank stores step n+1's records into a receive buffer before its owner unpacked step n's.
"""
import contextlib
import itertools
import random
import types

import pytest

from cosmological_particle_mesh_simulation_b200 import slab

HALF = slab.PEER_SLOTS_HALF


# ------------------------------------------------------------------------------------------------
# recorders
# ------------------------------------------------------------------------------------------------
class Stream:
    def __init__(self, name):
        self.name, self.ops = name, []

    def wait_event(self, e):
        self.ops.append(("wait_event", e))


class Event:
    def __init__(self, **_):
        self.done = False

    def record(self, stream=None):
        (stream or FakeCuda.state.current).ops.append(("record", self))


class FakeCuda:
    state = None
    Event = Event

    @staticmethod
    def current_stream(*_):
        return FakeCuda.state.current

    @staticmethod
    @contextlib.contextmanager
    def stream(s):
        prev, FakeCuda.state.current = FakeCuda.state.current, s
        try:
            yield
        finally:
            FakeCuda.state.current = prev


class RankRecorder:
    """Stands in for slab.SlabRank: every compute entry point becomes an op on the current stream."""
    KERNELS = ["deposit", "ghost_add", "fft_rows_forward", "fft_y_forward", "fft_z", "fft_y_inverse",
               "fft_rows_inverse", "fft_y_forward_local", "fft_push", "fft_pull", "fft_y_inverse_local",
               "fft_y_forward_push", "fft_y_inverse_pull", "gather", "migrate_pack", "migrate_unpack",
               "ghost_push_rho", "ghost_wait_rho", "ghost_push_phi", "ghost_wait_phi", "signal", "wait"]
    SLOT_MIG_COUNTS, SLOT_MIG_DATA = 19, 20          # PM_SLOT_MIG_* of csrc/pm_internal.cuh

    def __init__(self, rank, nranks, n_cells=512):
        self.rank, self.nranks, self.n_cells, self.peers_ready, self.ghosts_ready = rank, nranks, n_cells, True, True
        self.mig_ready = True
        self.total_particles = None
        self.buf = {k: (k, rank) for k in slab.BUF}
        self.main, self.side = Stream("main"), Stream("side")
        self.current = self.main
        self.issued = itertools.count()
        for k in self.KERNELS:
            setattr(self, k, self._recorder(k))

    def _recorder(self, name):
        def call(*args):
            c = args[0] if args and name.startswith("fft_") and name not in ("fft_rows_forward", "fft_rows_inverse") else None
            kind = name if name in ("signal", "wait") else "kernel"
            self.current.ops.append((kind, name, args[0] if kind != "kernel" else c))
        return call

    def chunk(self, name, c, C):
        return (name, self.rank, c)

    # migration through peer memory: stores + flag word / flag wait (+ the host read), as pm_migrate.cu does
    def migrate_counts_push(self):
        self.current.ops.append(("signal", "signal", self.SLOT_MIG_COUNTS))

    def migrate_counts_read(self):
        import numpy as np
        self.current.ops.append(("wait", "wait", self.SLOT_MIG_COUNTS))
        m = np.zeros((self.nranks, self.nranks + 3), dtype=np.int64)
        m[:, self.nranks + 2] = 1 << 30
        return m

    def migrate_push(self, counts, offsets):
        self.current.ops.append(("kernel", "migrate_push", None))
        self.current.ops.append(("signal", "signal", self.SLOT_MIG_DATA))

    def migrate_wait(self):
        self.current.ops.append(("wait", "wait", self.SLOT_MIG_DATA))

    # capacity bookkeeping the NCCL migration path consults before its all-to-all-v
    np_capacity = 1 << 30

    def leave_capacity(self):
        return 1 << 30

    def entries(self):
        return 0


class CommRecorder:
    """Stands in for slab.DistComm (one rank per process, collectives on the current stream)."""

    def __init__(self, rank_rec, two_streams):
        self.r = rank_rec
        self.nranks = rank_rec.nranks
        self.side_stream = rank_rec.side if two_streams else None

    def _coll(self, kind, peers):
        r = self.r
        r.current.ops.append(("collective", kind, frozenset(peers), next(r.issued)))

    def shift(self, send, recv, direction):
        P, me = self.nranks, self.r.rank
        self._coll("shift%+d" % direction, {me, (me + direction) % P, (me - direction) % P})

    def all_to_all(self, send, recv):
        self._coll("a2a", range(self.nranks))

    def exchange_count_tensors(self, counts):
        self._coll("counts", range(self.nranks))
        z = [0] * self.nranks
        return [z], [z]

    def all_to_all_v(self, send, sc, recv, rc):
        self._coll("a2av", range(self.nranks))

    def agree_any(self, flag):
        self._coll("agree", range(self.nranks))      # one more collective every rank enters
        return bool(flag)


def record_program(P, steps, transport, chunks, two_streams, monkeypatch, ghosts="nccl", migrate="nccl"):
    """Run slab_step `steps` times for each rank against the recorders -> per-rank op queues."""
    monkeypatch.setattr(slab, "torch", types.SimpleNamespace(cuda=FakeCuda))
    monkeypatch.setattr(slab.rt, "lib", lambda: types.SimpleNamespace(pm_slab_set_rho_mean=lambda *a: 0))
    cfg = types.SimpleNamespace(N_CELLS=512, N_PARTS=256, H0=0.68, OMEGA_LAMBDA0=0.69, OMEGA_K0=0.0, OMEGA_M0=0.31)
    ranks = []
    for r in range(P):
        rec = RankRecorder(r, P)
        FakeCuda.state = rec
        comm = CommRecorder(rec, two_streams)
        for s in range(steps):
            slab.slab_step([rec], comm, 0.1 + 0.01 * s, 0.01, mass=8.0, cfg=cfg, chunks=chunks, transport=transport,
                           ghosts=ghosts, migrate=migrate)
        ranks.append(rec)
    return ranks


# ------------------------------------------------------------------------------------------------
# the executor
# ------------------------------------------------------------------------------------------------
class Hazard(AssertionError):
    pass


def execute(ranks, transport, chunks, rng):
    P = len(ranks)
    pc = {(r.rank, s.name): 0 for r in ranks for s in (r.main, r.side)}
    queues = {(r.rank, s.name): s.ops for r in ranks for s in (r.main, r.side)}
    NSLOT = 24
    flags = [[[0] * P for _ in range(NSLOT)] for _ in range(P)]          # flags[owner][slot][writer]
    sig_epoch = [[0] * NSLOT for _ in range(P)]
    wait_epoch = {}                                                       # (rank, stream, pc) -> epoch
    wait_count = [[0] * NSLOT for _ in range(P)]
    mig_pushed, unpacked = {}, [0] * P                                    # step -> ranks that stored their records
    arrived = {}                                                          # collective key -> set of ranks at it
    coll_done = [0] * P                                                   # collectives completed per rank (issue order)
    step = [0] * P                                                        # deposits executed
    delivered, z_done, fetched = {}, {}, {}                               # (step, c) -> set of ranks
    gflag = {}                                                            # (owner, kind) -> pushes received
    gwaited = {}                                                          # (rank, stream, pc) -> epochs awaited
    gcount = [[0, 0] for _ in range(P)]                                   # ghost waits issued per rank (rho, phi)
    added, gathered = [0] * P, [0] * P                                    # ghost_add / gather executed per rank

    def mark(table, key, r):
        table.setdefault(key, set()).add(r)

    def full(table, key):
        return len(table.get(key, ())) == P

    def runnable(r, sname):
        q, i = queues[(r, sname)], pc[(r, sname)]
        if i >= len(q):
            return False
        op = q[i]
        if op[0] == "wait_event":
            return op[1].done
        if op[0] == "wait":
            slot = op[2]
            key = (r, sname, i)
            if key not in wait_epoch:
                wait_count[r][slot] += 1
                wait_epoch[key] = wait_count[r][slot]
            return all(flags[r][slot][s] >= wait_epoch[key] for s in range(P))
        if op[0] == "kernel" and op[1] in ("ghost_wait_rho", "ghost_wait_phi"):
            which = 0 if op[1] == "ghost_wait_rho" else 1
            key = (r, sname, i)
            if key not in gwaited:
                gcount[r][which] += 1
                gwaited[key] = gcount[r][which]
            kinds = ("rho",) if which == 0 else ("phi_up", "phi_dn")
            return all(gflag.get((r, k), 0) >= gwaited[key] for k in kinds)
        if op[0] == "collective":
            _, kind, peers, seq = op
            if seq != coll_done[r]:
                return False                       # an earlier collective of this rank is still pending
            arrived.setdefault(seq, {})[r] = kind
            return all(p in arrived[seq] for p in peers)
        return True

    def run(r, sname):
        q, i = queues[(r, sname)], pc[(r, sname)]
        op = q[i]
        if op[0] == "record":
            op[1].done = True
        elif op[0] == "signal":
            slot = op[2]
            sig_epoch[r][slot] += 1
            for owner in range(P):
                flags[owner][slot][r] = sig_epoch[r][slot]
        elif op[0] == "collective":
            coll_done[r] += 1
        elif op[0] == "kernel":
            name, c = op[1], op[2]
            n = step[r]
            up, dn = (r + 1) % P, (r - 1) % P
            if name == "deposit":
                step[r] += 1
            elif name == "ghost_push_rho":
                if gathered[up] < n - 1:
                    raise Hazard(f"H6: rank {r} writes step {n}'s density ghost into rank {up} before its gather of step {n - 1}")
                gflag[(up, "rho")] = gflag.get((up, "rho"), 0) + 1
            elif name == "ghost_add":
                if any(o[1] == "ghost_push_rho" for o in q) and gflag.get((r, "rho"), 0) < n:
                    raise Hazard(f"H4: rank {r} adds the density ghost of step {n} before it arrived")
                added[r] += 1
            elif name == "ghost_push_phi":
                if added[up] < n:
                    raise Hazard(f"H7: rank {r} overwrites rank {up}'s phi plane 0 before its ghost_add of step {n}")
                if gathered[dn] < n - 1:
                    raise Hazard(f"H6: rank {r} writes potential planes into rank {dn} before its gather of step {n - 1}")
                gflag[(up, "phi_up")] = gflag.get((up, "phi_up"), 0) + 1
                gflag[(dn, "phi_dn")] = gflag.get((dn, "phi_dn"), 0) + 1
            elif name == "gather":
                if any(o[1] == "ghost_push_phi" for o in q) and min(gflag.get((r, "phi_up"), 0), gflag.get((r, "phi_dn"), 0)) < n:
                    raise Hazard(f"H5: rank {r} gathers step {n} before its potential ghost planes arrived")
                gathered[r] += 1
            elif name == "migrate_push":
                if any(unpacked[d] < n - 1 for d in range(P)):
                    raise Hazard(f"H9: rank {r} stores step {n}'s records before every owner unpacked step {n - 1}'s")
                mark(mig_pushed, n, r)
            elif name == "migrate_unpack":
                if any(o[1] == "migrate_push" for o in q) and not full(mig_pushed, n):
                    raise Hazard(f"H8: rank {r} unpacks step {n} before every rank stored its records")
                unpacked[r] += 1
            elif name in ("fft_push", "fft_y_forward_push"):
                if n > 1 and not full(fetched, (n - 1, c)):
                    raise Hazard(f"H3: rank {r} stores step {n} chunk {c} before everyone fetched step {n - 1}'s")
                mark(delivered, (n, c), r)
            elif name == "fft_z":
                if transport != "nccl" and not full(delivered, (n, c)):
                    raise Hazard(f"H1: rank {r} runs z pass of step {n} chunk {c} before every block arrived")
                mark(z_done, (n, c), r)
            elif name in ("fft_pull", "fft_y_inverse_pull"):
                if not full(z_done, (n, c)):
                    raise Hazard(f"H2: rank {r} fetches step {n} chunk {c} before every z pass finished")
                mark(fetched, (n, c), r)
        pc[(r, sname)] = i + 1

    keys = list(queues)
    total = sum(len(q) for q in queues.values())
    executed = 0
    while executed < total:
        ready = [k for k in keys if runnable(*k)]
        if not ready:
            heads = {k: queues[k][pc[k]][:3] for k in keys if pc[k] < len(queues[k])}
            raise AssertionError(f"deadlock after {executed}/{total} ops; heads: {heads}")
        run(*rng.choice(ready))
        executed += 1
    return executed


@pytest.mark.parametrize("transport,chunks,two_streams", [
    ("fused", 1, True), ("fused", 2, True), ("fused", 4, False),
    ("fused2", 1, True), ("fused2", 2, True), ("fused2", 4, True), ("fused2", 2, False),
    ("peer", 1, True), ("peer", 2, True), ("peer", 4, True), ("peer", 2, False),
    ("nccl", 1, True), ("nccl", 4, True), ("nccl", 2, False)])
@pytest.mark.parametrize("P", [2, 4])
def test_schedule_drains_and_orders_its_hazards(monkeypatch, P, transport, chunks, two_streams):
    ranks = record_program(P, steps=3, transport=transport, chunks=chunks, two_streams=two_streams,
                           monkeypatch=monkeypatch)
    for r in ranks:        # every rank issued the same program (signal/wait pair up by call count)
        assert [op[:3] for op in r.main.ops if op[0] != "collective" and op[0] not in ("record", "wait_event")] == \
               [op[:3] for op in ranks[0].main.ops if op[0] != "collective" and op[0] not in ("record", "wait_event")]
    n_ops = sum(len(s.ops) for r in ranks for s in (r.main, r.side))
    for seed in range(25):
        for r in ranks:    # events are one-shot per run
            for s in (r.main, r.side):
                for op in s.ops:
                    if op[0] == "record":
                        op[1].done = False
        assert execute(ranks, transport, chunks, random.Random(seed)) == n_ops


@pytest.mark.parametrize("transport,chunks,two_streams", [("fused", 1, True), ("fused2", 2, True), ("peer", 2, True)])
@pytest.mark.parametrize("P", [2, 3, 4])
def test_ghost_planes_through_peer_memory_are_ordered(monkeypatch, P, transport, chunks, two_streams):
    ranks = record_program(P, steps=3, transport=transport, chunks=chunks, two_streams=two_streams,
                           monkeypatch=monkeypatch, ghosts="peer")
    assert not any(op[0] == "collective" and op[1].startswith("shift") for r in ranks for op in r.main.ops)
    n_ops = sum(len(s.ops) for r in ranks for s in (r.main, r.side))
    for seed in range(25):
        for r in ranks:
            for s in (r.main, r.side):
                for op in s.ops:
                    if op[0] == "record":
                        op[1].done = False
        assert execute(ranks, transport, chunks, random.Random(seed)) == n_ops


@pytest.mark.parametrize("transport,chunks,two_streams,ghosts", [("fused", 1, True, "peer"), ("fused2", 2, True, "peer"),
                                                                 ("nccl", 2, True, "nccl"), ("peer", 2, False, "peer")])
@pytest.mark.parametrize("P", [2, 3, 4])
def test_migration_through_peer_memory_is_ordered(monkeypatch, P, transport, chunks, two_streams, ghosts):
    ranks = record_program(P, steps=3, transport=transport, chunks=chunks, two_streams=two_streams,
                           monkeypatch=monkeypatch, ghosts=ghosts, migrate="peer")
    assert not any(op[0] == "collective" and op[1] in ("counts", "a2av") for r in ranks for op in r.main.ops)
    n_ops = sum(len(s.ops) for r in ranks for s in (r.main, r.side))
    for seed in range(25):
        for r in ranks:
            for s in (r.main, r.side):
                for op in s.ops:
                    if op[0] == "record":
                        op[1].done = False
        assert execute(ranks, transport, chunks, random.Random(seed)) == n_ops


def test_the_model_catches_unordered_migration(monkeypatch):
    """Drop the waits on the migration flags and H8 must fire for some interleaving."""
    ranks = record_program(2, steps=2, transport="fused", chunks=1, two_streams=False, monkeypatch=monkeypatch,
                           ghosts="peer", migrate="peer")
    for r in ranks:
        r.main.ops = [op for op in r.main.ops if not (op[0] == "wait" and op[2] == RankRecorder.SLOT_MIG_DATA)]
    fired = set()
    for seed in range(60):
        try:
            execute(ranks, "fused", 1, random.Random(seed))
        except Hazard as e:
            fired.add(str(e)[:2])
    assert "H8" in fired


def test_the_model_catches_an_unordered_ghost_push(monkeypatch):
    """Without the FFT's flag barriers between them, rank-1's potential plane could land on phi plane 0
    before ghost_add has read the density ghost from it (H7): drop every flag wait of the transform and
    the model must say so."""
    ranks = record_program(2, steps=2, transport="fused", chunks=1, two_streams=False, monkeypatch=monkeypatch,
                           ghosts="peer")
    for r in ranks:
        r.main.ops = [op for op in r.main.ops if op[0] != "wait" and op[1] != "ghost_wait_rho"]
    fired = set()
    for seed in range(60):
        try:
            execute(ranks, "fused", 1, random.Random(seed))
        except Hazard as e:
            fired.add(str(e)[:2])
    assert fired & {"H4", "H7", "H1"}


def test_the_model_catches_a_missing_barrier(monkeypatch):
    """Sanity of the checker itself: drop the 'chunk pushed' waits and H1 must fire for some interleaving."""
    ranks = record_program(2, steps=2, transport="fused", chunks=1, two_streams=False, monkeypatch=monkeypatch)
    for r in ranks:
        r.main.ops = [op for op in r.main.ops if not (op[0] == "wait" and op[2] < HALF)]
    fired = 0
    for seed in range(40):
        try:
            execute(ranks, "fused", 1, random.Random(seed))
        except Hazard as e:
            assert "H1" in str(e)
            fired += 1
    assert fired > 0
