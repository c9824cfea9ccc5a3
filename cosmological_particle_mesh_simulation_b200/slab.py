"""Slab-decomposed particle-mesh step for 2/4/8 GPUs of one box.

The reference is single-process; nothing here has a reference counterpart except the physics,
which must equal the single-GPU step.  Layout of the work:

  * rank r owns mesh planes [r*Nc/P, (r+1)*Nc/P) along array axis 0 (the reference's z =
    positions[2], density.py:14,21,37) and the particles whose z cell lies in them
    (`slab_of_particles`: slab = (int(floor(z)) mod Nc) // (Nc/P), bit-exact incl. SURVEY Q4);
  * all computation is in libpmstep.so (pm_slab_* entry points, csrc/pm_slab.cu);
  * this module only sequences those calls and moves the plan's buffers between ranks through a
    `Comm`: `DistComm` = torch.distributed (NCCL on GPUs; send/recv for the ghost planes,
    all_to_all_single around the pack/unpack transposes, all_to_all_single with split sizes for
    migrating particles), `LocalComm` = all P ranks inside one process on one GPU, exchanges by
    tensor copies -- the single-GPU rank loop the tests use to check every slab kernel against the
    plain single-GPU step.
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

try:
    from . import _runtime as rt
    from .cosmology import f
except ImportError:  # flat layout
    import _runtime as rt
    from cosmology import f

BUF = dict(RHO=0, RHO_GHOST_SEND=1, RHO_GHOST_RECV=2, FFT_SEND_MAIN=3, FFT_SEND_SIDE=4,
           FFT_RECV_MAIN=5, FFT_RECV_SIDE=6, PHI=7, PHI_LO_SEND=8, PHI_HI_SEND=9, PHI_LO_RECV=10,
           PHI_HI_RECV=11, MIG_SEND=12, MIG_RECV=13, LEAVE_COUNTS=14, PEER_FLAGS=15)
PEER_SLOTS_HALF = 8     # flag slots 0..7: "chunk c pushed", 8..15: "z pass of chunk c done"


class _RawCudaBuffer:
    """Zero-copy view of plan-owned device memory for torch (via __cuda_array_interface__)."""

    def __init__(self, ptr, nbytes, owner):
        self._owner = owner  # keep the plan alive
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False),
                                         "version": 2, "strides": None}


def slab_of_particles(pos_z, n_cells: int, nranks: int):
    """Owner rank of each particle from its z coordinate (NumPy array or tensor, float32):
    (int(floor(z)) mod Nc) // (Nc/P) -- the cell index rule of density.py:21."""
    nzl = n_cells // nranks
    if isinstance(pos_z, torch.Tensor):
        c = torch.remainder(torch.floor(pos_z).to(torch.int64), n_cells)
        return (c // nzl).to(torch.int32)
    c = np.floor(pos_z).astype(np.int64) % n_cells
    return (c // nzl).astype(np.int32)


class SlabRank:
    """One rank's plan and the views of its exchange buffers."""

    def __init__(self, n_cells: int, np_capacity: int, device: int, rank: int, nranks: int):
        self.n_cells, self.rank, self.nranks, self.device = int(n_cells), int(rank), int(nranks), int(device)
        self.nzl = self.n_cells // self.nranks
        self.np_capacity = int(np_capacity)
        h = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            rt.check(rt.lib().pm_plan_create_slab(ctypes.byref(h), self.n_cells, self.np_capacity,
                                                  self.device, self.rank, self.nranks),
                     f"pm_plan_create_slab(n_cells={n_cells}, np={np_capacity}, rank={rank}/{nranks})")
            rt.install_reference_tables(h, self.n_cells)     # the reference's own sin^2 table: G bit for bit
        self.handle = h
        self.total_particles = None      # particles of ALL ranks (make_ranks / make_rank_from_local set it)
        self.buf = {}
        n, nzl, nyl, hh, P = self.n_cells, self.nzl, self.nzl, self.n_cells // 2, self.nranks
        shapes = dict(RHO=(torch.float32, (nzl, n, n)), RHO_GHOST_SEND=(torch.float32, (n, n)),
                      RHO_GHOST_RECV=(torch.float32, (n, n)),
                      FFT_SEND_MAIN=(torch.float32, (P, nzl * nyl * hh * 2)),
                      FFT_SEND_SIDE=(torch.float32, (P, nzl * nyl * 2)),
                      FFT_RECV_MAIN=(torch.float32, (P, nzl * nyl * hh * 2)),
                      FFT_RECV_SIDE=(torch.float32, (P, nzl * nyl * 2)),
                      PHI=(torch.float32, (nzl, n, n)), PHI_LO_SEND=(torch.float32, (2, n, n)),
                      PHI_HI_SEND=(torch.float32, (n, n)), PHI_LO_RECV=(torch.float32, (n, n)),
                      PHI_HI_RECV=(torch.float32, (2, n, n)), MIG_SEND=(torch.float32, (-1, 7)),
                      MIG_RECV=(torch.float32, (-1, 7)), LEAVE_COUNTS=(torch.int32, (P,)),
                      PEER_FLAGS=(torch.int32, (-1, 16)))
        self.peers_ready = False
        self.ghosts_ready = False
        self.mig_ready = False
        for name, which in BUF.items():
            ptr, nbytes = ctypes.c_void_p(), ctypes.c_size_t()
            rt.check(rt.lib().pm_slab_buffer(self.handle, which, ctypes.byref(ptr), ctypes.byref(nbytes)),
                     "pm_slab_buffer")
            raw = torch.as_tensor(_RawCudaBuffer(ptr.value, nbytes.value, self), device=f"cuda:{self.device}")
            dtype, shape = shapes[name]
            self.buf[name] = raw.view(dtype).view(*shape)

    def close(self):
        if getattr(self, "handle", None):
            self.buf = {}
            rt.lib().pm_plan_destroy(self.handle)
            self.handle = None

    def _call(self, fn, *args):
        with torch.cuda.device(self.device):
            rt.check(getattr(rt.lib(), fn)(self.handle, *args, rt.stream_ptr(self.device)), fn)

    # -- state -------------------------------------------------------------------------------
    def load(self, pos, vel, ids):
        rt.check_dev_f32(pos, name="positions")
        rt.check_dev_f32(vel, tuple(pos.shape), "velocities")
        ids = ids.to(torch.int32).contiguous()
        self._call("pm_slab_load", pos.data_ptr(), vel.data_ptr(), ids.data_ptr(), pos.shape[1])

    @property
    def count(self):
        return int(rt.lib().pm_slab_count(self.handle))

    def entries(self):
        """Slots in use (live particles + entries that left and are dropped by the next sort)."""
        return int(rt.lib().pm_slab_entries(self.handle))

    def export(self):
        """(pos[3,n], vel[3,n], ids[n]) of the live particles, storage order."""
        n = int(rt.lib().pm_slab_entries(self.handle))
        dev = f"cuda:{self.device}"
        pos = torch.empty((3, n), dtype=torch.float32, device=dev)
        vel = torch.empty_like(pos)
        ids = torch.empty(n, dtype=torch.int32, device=dev)
        live = torch.empty(n, dtype=torch.int32, device=dev)
        self._call("pm_slab_export", pos.data_ptr(), vel.data_ptr(), ids.data_ptr(), live.data_ptr())
        keep = live != 0
        return pos[:, keep], vel[:, keep], ids[keep]

    # -- compute phases ------------------------------------------------------------------------
    def deposit(self, mass):
        self._call("pm_slab_deposit", float(mass))

    def ghost_add(self):
        self._call("pm_slab_ghost_add")

    def fft_rows_forward(self):
        self._call("pm_slab_fft_rows_forward")

    def fft_y_forward(self, c, C):
        self._call("pm_slab_fft_y_forward", int(c), int(C))

    def fft_z(self, c, C, a, omega_m0):
        self._call("pm_slab_fft_z", int(c), int(C), float(a), float(omega_m0))

    def fft_y_inverse(self, c, C):
        self._call("pm_slab_fft_y_inverse", int(c), int(C))

    def fft_rows_inverse(self):
        self._call("pm_slab_fft_rows_inverse")

    # -- peer-memory transposes (pm_slab_peer_* / pm_slab_fft_push / pm_slab_fft_pull) -------------
    def peer_export(self):
        """(64-byte CUDA IPC handle of the plan's workspace, offset of the z-pass array, offset of
        the flag words) -- what the other ranks need to map this rank's buffers."""
        h = (ctypes.c_ubyte * 64)()
        o1, o2 = ctypes.c_uint64(), ctypes.c_uint64()
        rt.check(rt.lib().pm_slab_peer_export(self.handle, h, ctypes.byref(o1), ctypes.byref(o2)),
                 "pm_slab_peer_export")
        return bytes(h), int(o1.value), int(o2.value)

    def peer_import(self, peer, handle, recv_offset, flags_offset):
        buf = (ctypes.c_ubyte * 64).from_buffer_copy(handle) if handle is not None else None
        with torch.cuda.device(self.device):
            rt.check(rt.lib().pm_slab_peer_import(self.handle, int(peer), buf, int(recv_offset), int(flags_offset)),
                     f"pm_slab_peer_import(peer={peer})")

    def peer_set(self, peer, recv_ptr, flags_ptr):
        rt.check(rt.lib().pm_slab_peer_set(self.handle, int(peer), ctypes.c_void_p(recv_ptr),
                                           ctypes.c_void_p(flags_ptr)), "pm_slab_peer_set")

    def peer_timeouts(self):
        n = ctypes.c_uint32()
        rt.check(rt.lib().pm_slab_peer_timeouts(self.handle, ctypes.byref(n)), "pm_slab_peer_timeouts")
        return int(n.value)

    def peer_release(self):
        self.ghosts_ready = False
        self.mig_ready = False
        if getattr(self, "handle", None):
            with torch.cuda.device(self.device):
                torch.cuda.synchronize()
                rt.check(rt.lib().pm_slab_peer_release(self.handle), "pm_slab_peer_release")
        self.peers_ready = False

    def signal(self, slot):
        self._call("pm_slab_peer_signal", int(slot))

    def wait(self, slot):
        self._call("pm_slab_peer_wait", int(slot))

    def fft_y_forward_local(self, c, C):
        self._call("pm_slab_fft_y_forward_local", int(c), int(C))

    def fft_push(self, c, C):
        self._call("pm_slab_fft_push", int(c), int(C))

    def fft_pull(self, c, C):
        self._call("pm_slab_fft_pull", int(c), int(C))

    # ghost planes through peer memory (experimental; pm_slab_ghost_*)
    def ghost_push_rho(self):
        self._call("pm_slab_ghost_push_rho")

    def ghost_wait_rho(self):
        self._call("pm_slab_ghost_wait_rho")

    def ghost_push_phi(self):
        self._call("pm_slab_ghost_push_phi")

    def ghost_wait_phi(self):
        self._call("pm_slab_ghost_wait_phi")

    def fft_y_forward_push(self, c, C):
        self._call("pm_slab_fft_y_forward_push", int(c), int(C))

    def fft_y_inverse_pull(self, c, C):
        self._call("pm_slab_fft_y_inverse_pull", int(c), int(C))

    def fft_y_inverse_local(self, c, C):
        self._call("pm_slab_fft_y_inverse_local", int(c), int(C))

    def chunk(self, name, c, C):
        """Chunk c of C of an FFT_*_MAIN buffer as a (P, .) view (one row per peer rank)."""
        flat = self.buf[name].view(-1)
        n = flat.numel() // C
        return flat[c * n:(c + 1) * n].view(self.nranks, -1)

    def gather(self, a, f_a1, da):
        self._call("pm_slab_gather", float(a), float(f_a1), float(da))

    def leave_capacity(self):
        """Records per destination rank and step the leave lists hold (csrc/pm_api.cu, compute_layout)."""
        npad = (self.np_capacity + 3) // 4 * 4
        return max(npad // 32, 65536)

    def migrate_pack(self, counts):
        arr = (ctypes.c_int64 * self.nranks)(*[int(c) for c in counts])
        self._call("pm_slab_migrate_pack", arr)

    # migration through peer memory (csrc/pm_migrate.cu)
    def migrate_counts_push(self):
        self._call("pm_slab_migrate_counts_push")

    def migrate_counts_read(self):
        """The count matrix every rank holds after the exchange: int64 array [P, P + 3]; row s = rank s's
        leavers per destination, then its flag-wait timeouts, its leave-list overflow flag and its free
        particle capacity.  Synchronises this rank's stream (the one host read of the step)."""
        P = self.nranks
        m = np.zeros((P, P + 3), dtype=np.uint32)
        with torch.cuda.device(self.device):
            rt.check(rt.lib().pm_slab_migrate_counts_read(self.handle, m.ctypes.data, rt.stream_ptr(self.device)),
                     "pm_slab_migrate_counts_read")
        return m.astype(np.int64)

    def migrate_push(self, send_counts, dest_offsets):
        P = self.nranks
        a = (ctypes.c_int64 * P)(*[int(c) for c in send_counts])
        b = (ctypes.c_int64 * P)(*[int(c) for c in dest_offsets])
        self._call("pm_slab_migrate_push", a, b)

    def migrate_wait(self):
        self._call("pm_slab_migrate_wait")

    def migrate_unpack(self, n_arrive, n_leave):
        self._call("pm_slab_migrate_unpack", int(n_arrive), int(n_leave))


# -------------------------------------------------------------------------------------------------
# communication back ends
# -------------------------------------------------------------------------------------------------
class LocalComm:
    """All P ranks live in this process (one GPU): exchanges are copies.  Test harness."""

    def __init__(self, nranks):
        self.nranks = nranks
        self.local_ranks = list(range(nranks))

    def shift(self, send, recv, direction):
        P = self.nranks
        for i in range(P):
            recv[(i + direction) % P].copy_(send[i])

    def all_to_all(self, send, recv):
        P = self.nranks
        staged = [s.clone() for s in send]  # send and recv may alias across calls
        for i in range(P):
            for j in range(P):
                recv[j][i].copy_(staged[i][j])

    def exchange_counts(self, counts):
        P = self.nranks
        return [[int(counts[src][dst]) for src in range(P)] for dst in range(P)]

    def exchange_count_tensors(self, counts):
        """counts: per local rank an integer tensor [P] (leavers per destination).  Returns
        (send_counts, recv_counts) as lists of Python ints per local rank."""
        send = [c.tolist() for c in counts]
        return send, self.exchange_counts(send)

    def agree_any(self, flag):
        """True on every rank if `flag` is true on any (all ranks live here: nothing to exchange)."""
        return bool(flag)

    def all_to_all_v(self, send, send_counts, recv, recv_counts):
        P = self.nranks
        off_s = [np.concatenate([[0], np.cumsum(c)]) for c in send_counts]
        off_r = [np.concatenate([[0], np.cumsum(c)]) for c in recv_counts]
        for src in range(P):
            for dst in range(P):
                n = int(send_counts[src][dst])
                if n:
                    recv[dst][off_r[dst][src]:off_r[dst][src] + n].copy_(
                        send[src][off_s[src][dst]:off_s[src][dst] + n])


class DistComm:
    """One rank per process over torch.distributed (NCCL on GPUs, gloo on CPU for the tests).
    With a backend that cannot move CUDA tensors (gloo) the exchanges are staged through host
    memory -- slow, but it lets two processes SHARE one GPU, which is how the CUDA IPC mappings of
    the peer transport are tested on a one-GPU box (tests/test_slab_gpu.py)."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.nranks = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.local_ranks = [self.rank]
        nccl = dist.get_backend(group) == "nccl"
        self._count_device = "cuda" if nccl else "cpu"
        # collectives of the FFT pipeline run here so they overlap the compute stream
        self.side_stream = torch.cuda.Stream() if (nccl and self.nranks > 1) else None
        self._staged = not nccl

    def _host(self, t):
        return t.cpu() if (self._staged and t.is_cuda) else t

    def shift(self, send, recv, direction):
        P, r, dist = self.nranks, self.rank, self.dist
        if P == 1:
            recv[0].copy_(send[0])
            return
        if self._staged and send[0].is_cuda:
            hs, hr = send[0].cpu(), torch.empty(recv[0].shape, dtype=recv[0].dtype)
            self.shift([hs], [hr], direction)
            recv[0].copy_(hr)
            return
        ops = [dist.P2POp(dist.isend, send[0], (r + direction) % P, self.group),
               dist.P2POp(dist.irecv, recv[0], (r - direction) % P, self.group)]
        for req in dist.batch_isend_irecv(ops):
            req.wait()

    def all_to_all(self, send, recv):
        if self.nranks == 1:
            recv[0].copy_(send[0])
            return
        if self._staged and send[0].is_cuda:
            hs, hr = send[0].cpu(), torch.empty(recv[0].shape, dtype=recv[0].dtype)
            self.dist.all_to_all_single(hr.view(-1), hs.view(-1), group=self.group)
            recv[0].copy_(hr)
            return
        self.dist.all_to_all_single(recv[0].view(-1), send[0].view(-1), group=self.group)

    def exchange_counts(self, counts):
        c = torch.tensor([int(x) for x in counts[0]], dtype=torch.int64, device=self._count_device)
        out = torch.empty_like(c)
        if self.nranks == 1:
            out.copy_(c)
        else:
            self.dist.all_to_all_single(out, c, group=self.group)
        return [out.tolist()]

    _count_device = "cpu"

    def agree_any(self, flag):
        """True on every rank if `flag` is true on any rank (one small all-reduce)."""
        if self.nranks == 1:
            return bool(flag)
        t = torch.tensor([1 if flag else 0], dtype=torch.int32, device=self._count_device)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX, group=self.group)
        return bool(int(t.item()))

    def exchange_count_tensors(self, counts):
        """As LocalComm.exchange_count_tensors, with ONE device->host read per step: the counts
        are exchanged on the device (all-to-all of P integers) and both vectors come back together."""
        c = self._host(counts[0])
        if self.nranks == 1:
            v = c.tolist()
            return [v], [v]
        both = torch.empty((2, self.nranks), dtype=c.dtype, device=c.device)
        both[0].copy_(c)
        self.dist.all_to_all_single(both[1], both[0], group=self.group)
        v = both.tolist()
        return [v[0]], [v[1]]

    def all_to_all_v(self, send, send_counts, recv, recv_counts):
        sc, rc = [int(x) for x in send_counts[0]], [int(x) for x in recv_counts[0]]
        ns, nr = sum(sc), sum(rc)
        if self.nranks == 1:
            recv[0][:nr].copy_(send[0][:ns])
            return
        width = send[0].shape[1]
        if self._staged and send[0].is_cuda:
            hs, hr = send[0][:ns].cpu(), torch.empty((nr, width), dtype=recv[0].dtype)
            self.dist.all_to_all_single(hr.view(-1), hs.view(-1), output_split_sizes=[c * width for c in rc],
                                        input_split_sizes=[c * width for c in sc], group=self.group)
            recv[0][:nr].copy_(hr)
            return
        self.dist.all_to_all_single(recv[0][:nr].reshape(-1), send[0][:ns].reshape(-1),
                                    output_split_sizes=[c * width for c in rc],
                                    input_split_sizes=[c * width for c in sc], group=self.group)


def setup_peers(ranks, comm):
    """Give every rank the addresses of everybody's z-pass array and flag words so the FFT
    transposes can go through peer memory (slab_step(transport="peer")).  LocalComm: plain pointers.
    DistComm: CUDA IPC handles exchanged with all_gather_object, then a flag round trip through the
    mapped memory as a handshake; all ranks agree on the outcome.  Returns True when the peer path is
    usable (and marks the ranks), False otherwise -- the NCCL path needs no set-up."""
    P = comm.nranks
    if P > 16:
        return False
    if isinstance(comm, LocalComm):
        for r in ranks:
            for s in ranks:
                r.peer_set(s.rank, s.buf["FFT_RECV_MAIN"].data_ptr(), s.buf["PEER_FLAGS"].data_ptr())
        for r in ranks:
            r.signal(PEER_SLOTS_HALF - 1)
        for r in ranks:
            r.wait(PEER_SLOTS_HALF - 1)
        torch.cuda.synchronize()
        ok = all(r.peer_timeouts() == 0 for r in ranks)
        for r in ranks:
            r.peers_ready = ok
        return ok
    dist, r = comm.dist, ranks[0]
    ok = 1
    try:
        mine = r.peer_export()
    except Exception:
        mine, ok = None, 0
    table = [None] * P
    dist.all_gather_object(table, mine, group=comm.group)
    if ok and all(t is not None for t in table):
        try:
            for s, (h, o1, o2) in enumerate(table):
                r.peer_import(s, None if s == r.rank else h, o1, o2)
        except Exception:
            ok = 0
    else:
        ok = 0
    dev = r.buf["RHO"].device if dist.get_backend(comm.group) == "nccl" else "cpu"
    flag = torch.tensor([ok], dtype=torch.int32, device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=comm.group)
    if int(flag.item()) == 0:
        return False
    # handshake through the mapped memory (slot 7 is never used by the pipeline: at most 4 chunks)
    r.signal(PEER_SLOTS_HALF - 1)
    r.wait(PEER_SLOTS_HALF - 1)
    torch.cuda.synchronize()
    flag = torch.tensor([1 if r.peer_timeouts() == 0 else 0], dtype=torch.int32, device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=comm.group)
    r.peers_ready = bool(int(flag.item()))
    return r.peers_ready


def setup_ghost_peers(ranks, comm):
    """After setup_peers(): also publish where every rank's phi buffer, migration receive buffer and
    count matrix live, so that slab_step can push the density / potential ghost planes into the
    neighbours' memory (ghosts="peer") and migrate particles through peer memory (migrate="peer",
    csrc/pm_migrate.cu) instead of NCCL send/recv and all-to-all-v.  Returns True when usable."""
    if not all(r.peers_ready for r in ranks):
        return False
    vp = ctypes.c_void_p
    if isinstance(comm, LocalComm):
        ptrs = {}
        for s in ranks:
            a, b, c = vp(), vp(), vp()
            rt.check(rt.lib().pm_slab_aux_buffers(s.handle, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c)),
                     "pm_slab_aux_buffers")
            ptrs[s.rank] = (a, b, c)
        for r in ranks:
            for s in ranks:
                rt.check(rt.lib().pm_slab_peer_aux_set(r.handle, s.rank, *ptrs[s.rank]), "pm_slab_peer_aux_set")
        for r in ranks:
            r.ghosts_ready = r.mig_ready = True
        return True
    dist, r = comm.dist, ranks[0]
    off = (ctypes.c_uint64 * 3)()
    rt.check(rt.lib().pm_slab_peer_aux_export(r.handle, off), "pm_slab_peer_aux_export")
    table = [None] * comm.nranks
    dist.all_gather_object(table, [int(off[0]), int(off[1]), int(off[2])], group=comm.group)
    ok = 1
    try:
        for s, o in enumerate(table):
            rt.check(rt.lib().pm_slab_peer_aux_import(r.handle, s, (ctypes.c_uint64 * 3)(*o)), "pm_slab_peer_aux_import")
    except Exception:
        ok = 0
    dev = r.buf["RHO"].device if dist.get_backend(comm.group) == "nccl" else "cpu"
    flag = torch.tensor([ok], dtype=torch.int32, device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=comm.group)
    r.ghosts_ready = r.mig_ready = bool(int(flag.item()))
    return r.ghosts_ready


def release_peers(ranks, comm):
    """Unmap the other ranks' buffers on every rank, then a barrier: call before closing the ranks
    of a DistComm run (memory must not be freed while another process still has it mapped)."""
    for r in ranks:
        r.peer_release()
    if isinstance(comm, DistComm) and comm.nranks > 1:
        comm.dist.barrier(group=comm.group)


# -------------------------------------------------------------------------------------------------
# the step
# -------------------------------------------------------------------------------------------------
class PhaseTimer:
    """CUDA-event timing of the phases of slab_step on the compute stream (bench.py)."""

    def __init__(self):
        self.records = []

    def begin_step(self):
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        self.cur = [("", e)]

    def mark(self, name):
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        self.cur.append((name, e))

    def end_step(self):
        self.records.append(self.cur)

    def mean_ms(self):
        torch.cuda.synchronize()
        out = {}
        for rec in self.records:
            for (_, e0), (name, e1) in zip(rec[:-1], rec[1:]):
                out[name] = out.get(name, 0.0) + e0.elapsed_time(e1) / len(self.records)
        return out


def default_chunks(n_cells, nranks=1, transport="nccl"):
    """kx chunks of the distributed FFT pipeline, by the size of one rank's half spectrum: deep
    pipelines pay off when a chunk's all-to-all is long against a kernel launch (measured on
    8 B200: 67 MB/rank is faster unchunked, 268 MB/rank and up is faster pipelined)."""
    if transport == "fused":
        # the transfer already overlaps the transform inside the y-pass kernels; chunking only adds
        # launches and flag waits (measured on 2 B200: 512^3 mesh 0.87 / 0.94 / 1.09 ms for 1 / 2 / 4 chunks)
        return 1
    if transport == "fused2":
        return 2 if (n_cells // 2) % 2 == 0 and ((n_cells // 2) // 2) % (8 if n_cells >= 1024 else 16) == 0 else 1
    tile = 8 if n_cells >= 1024 else 16
    mb = 4.0 * n_cells ** 3 / max(nranks, 1) / 2 ** 20
    want = 4 if mb >= 256 else 2 if mb >= 128 else 1   # 8 chunks measured no better than 4 at 4.3 GB/rank
    for c in (4, 2):
        if c <= want and (n_cells // 2) % c == 0 and ((n_cells // 2) // c) % tile == 0:
            return c
    return 1


class SlabExchangeError(RuntimeError):
    """A peer-memory exchange failed on SOME rank (flag wait timed out, leave list or particle capacity
    exceeded).  Every rank reads the same count matrix, so every rank raises this in the same step."""


def slab_step(ranks, comm, a, da, mass=None, cfg=None, timer=None, chunks=None, transport=None, ghosts=None,
              migrate=None):
    """One body of the loop src/pmesh.py:60-61 across the slabs.  `ranks`: the SlabRank objects
    of comm.local_ranks (one for DistComm, all P for LocalComm).  The distributed FFT runs as a
    pipeline of `chunks` kx chunks: with DistComm on GPUs the transposes go to a second stream
    and overlap the y and z passes of the neighbouring chunks.  transport: "fused" = the y passes
    themselves store into / load from the other ranks' z-pass arrays over NVLink (one stream, no
    copy kernel; needs setup_peers), "peer" = separate push / pull copy kernels on the second
    stream, "nccl" = pack -> all-to-all -> unpack; default: "fused" when the ranks are set up for it.
    ghosts: "peer" = the ghost planes are stored into the neighbours' memory + one flag word (default once
    setup_ghost_peers succeeded), "nccl" = send/recv.  migrate: "peer" = count matrix and particle records
    through peer memory, one host read per step, errors agreed by all ranks (same default), "nccl" =
    all-to-all of the counts + all-to-all-v of the records."""
    cfg = cfg or rt.config()
    if timer:
        timer.begin_step()
    mark = timer.mark if timer else (lambda name: None)
    if mass is None:
        mass = (cfg.N_CELLS / cfg.N_PARTS) ** 3          # src/pmesh.py:28
    f_a1 = f(a + da, [cfg.H0, cfg.OMEGA_LAMBDA0, cfg.OMEGA_K0])   # src/integrate.py:12
    B = lambda name: [r.buf[name] for r in ranks]    # noqa: E731
    if transport is None:
        transport = "fused" if all(r.peers_ready for r in ranks) else "nccl"
    if transport not in ("fused", "fused2", "peer", "nccl"):
        raise ValueError(f"unknown transport {transport!r}")
    if transport != "nccl" and not all(r.peers_ready for r in ranks):
        raise RuntimeError(f"transport={transport!r} needs slab.setup_peers(ranks, comm) first")
    C = chunks or default_chunks(ranks[0].n_cells, ranks[0].nranks, transport)
    CH = lambda name, c: [r.chunk(name, c, C) for r in ranks]    # noqa: E731

    if ghosts is None:
        ghosts = "peer" if all(r.ghosts_ready for r in ranks) else "nccl"
    if migrate is None:
        migrate = "peer" if all(r.mig_ready for r in ranks) else "nccl"
    for r in ranks:
        r.deposit(mass)
    mark("deposit")
    if ghosts not in ("nccl", "peer"):
        raise ValueError(f"unknown ghosts mode {ghosts!r}")
    if ghosts == "peer" and not all(r.ghosts_ready for r in ranks):
        raise RuntimeError("ghosts='peer' needs slab.setup_ghost_peers(ranks, comm) first")
    if ghosts == "peer":
        for r in ranks:
            r.ghost_push_rho()
        for r in ranks:
            r.ghost_wait_rho()
    else:
        comm.shift(B("RHO_GHOST_SEND"), B("RHO_GHOST_RECV"), +1)
    for r in ranks:
        r.ghost_add()
    mark("rho_ghost")

    # ---- distributed FFT, pipelined over kx chunks ----
    side = getattr(comm, "side_stream", None)
    main = torch.cuda.current_stream() if side is not None else None

    def on_comm(after, fn):
        """Run the collective `fn` after compute event `after`; returns the event it completes at."""
        if side is None:
            fn()
            return None
        side.wait_event(after)
        with torch.cuda.stream(side):
            fn()
            done = torch.cuda.Event()
            done.record(side)
        return done

    def ev():
        if side is None:
            return None
        e = torch.cuda.Event()
        e.record(main)
        return e

    if transport != "nccl" and C > PEER_SLOTS_HALF - 1:
        raise ValueError("at most %d chunks with the peer-memory transports" % (PEER_SLOTS_HALF - 1))

    for r in ranks:
        if r.total_particles is not None:       # <rho> of the whole mesh: the transform runs on rho - <rho>
            rt.check(rt.lib().pm_slab_set_rho_mean(r.handle, float(mass) * r.total_particles / float(r.n_cells) ** 3),
                     "pm_slab_set_rho_mean")
        r.fft_rows_forward()
    if transport == "fused":
        # one stream: chunk c's stores drain over NVLink while chunk c+1 is transformed; the flag
        # waits sit right before the first consumer of the data
        for c in range(C):
            for r in ranks:
                r.fft_y_forward_push(c, C)
            for r in ranks:
                r.signal(c)
        for c in range(C):
            for r in ranks:
                r.wait(c)
            for r in ranks:
                r.fft_z(c, C, a, cfg.OMEGA_M0)
            for r in ranks:
                r.signal(PEER_SLOTS_HALF + c)
        for c in range(C):
            for r in ranks:
                r.wait(PEER_SLOTS_HALF + c)
            for r in ranks:
                r.fft_y_inverse_pull(c, C)
    arrived = []
    for c in range(C if transport == "peer" else 0):
        for r in ranks:
            r.fft_y_forward_local(c, C)

        def push(c=c):
            for r in ranks:
                r.fft_push(c, C)
            for r in ranks:
                r.signal(c)
        arrived.append(on_comm(ev(), push))
    zdone = []
    for c in range(C if transport == "peer" else 0):
        if arrived[c] is not None:
            main.wait_event(arrived[c])       # my own block is in place
        for r in ranks:
            r.wait(c)                          # ... and so is everybody else's
        for r in ranks:
            r.fft_z(c, C, a, cfg.OMEGA_M0)
        for r in ranks:
            r.signal(PEER_SLOTS_HALF + c)      # "my z pass of chunk c is done: pull your planes"
        zdone.append(ev())
    pulled = []
    for c in range(C if transport == "peer" else 0):
        def pull(c=c):
            for r in ranks:
                r.wait(PEER_SLOTS_HALF + c)
            for r in ranks:
                r.fft_pull(c, C)
        pulled.append(on_comm(zdone[c], pull))
    for c in range(C if transport == "peer" else 0):
        if pulled[c] is not None:
            main.wait_event(pulled[c])
        for r in ranks:
            r.fft_y_inverse_local(c, C)
    if transport == "fused2":
        # EXPERIMENTAL (not yet measured on hardware): the fused y passes -- NVLink-bound when most of
        # the data is remote -- on the second stream, the z passes -- HBM-bound -- on the first, so
        # that chunk c's z pass overlaps chunk c+1's push and chunk c's pull overlaps chunk c+1's z
        # pass.  Without a second stream this is the order of "fused".  Schedule checked by
        # tests/test_slab_schedule_sim.py.
        def pushes():
            for c in range(C):
                for r in ranks:
                    r.fft_y_forward_push(c, C)
                for r in ranks:
                    r.signal(c)
        on_comm(ev(), pushes)
        zdone = []
        for c in range(C):
            for r in ranks:
                r.wait(c)
            for r in ranks:
                r.fft_z(c, C, a, cfg.OMEGA_M0)
            for r in ranks:
                r.signal(PEER_SLOTS_HALF + c)
            zdone.append(ev())

        def pulls():
            for c in range(C):
                for r in ranks:
                    r.wait(PEER_SLOTS_HALF + c)
                for r in ranks:
                    r.fft_y_inverse_pull(c, C)
        back = on_comm(zdone[0], pulls)
        if back is not None:
            main.wait_event(back)
    C_nccl = C if transport == "nccl" else 0
    arrived = []
    for c in range(C_nccl):
        for r in ranks:
            r.fft_y_forward(c, C)

        def fwd(c=c):
            comm.all_to_all(CH("FFT_SEND_MAIN", c), CH("FFT_RECV_MAIN", c))
            if c == 0:
                comm.all_to_all(B("FFT_SEND_SIDE"), B("FFT_RECV_SIDE"))
        arrived.append(on_comm(ev(), fwd))
    returned = []
    for c in range(C_nccl):
        if arrived[c] is not None:
            main.wait_event(arrived[c])
        for r in ranks:
            r.fft_z(c, C, a, cfg.OMEGA_M0)

        def bwd(c=c):
            comm.all_to_all(CH("FFT_RECV_MAIN", c), CH("FFT_SEND_MAIN", c))
            if c == 0:
                comm.all_to_all(B("FFT_RECV_SIDE"), B("FFT_SEND_SIDE"))
        returned.append(on_comm(ev(), bwd))
    for c in range(C_nccl):
        if returned[c] is not None:
            main.wait_event(returned[c])
        for r in ranks:
            r.fft_y_inverse(c, C)
    for r in ranks:
        r.fft_rows_inverse()
    mark("fft_distributed")

    if ghosts == "peer":
        for r in ranks:
            r.ghost_push_phi()
        for r in ranks:
            r.ghost_wait_phi()
    else:
        comm.shift(B("PHI_HI_SEND"), B("PHI_LO_RECV"), +1)    # my last plane is rank+1's plane z0-1
        comm.shift(B("PHI_LO_SEND"), B("PHI_HI_RECV"), -1)    # my first two planes close rank-1's stencil
    mark("phi_ghost")
    for r in ranks:
        r.gather(a, f_a1, da)
    mark("gather")
    if migrate not in ("nccl", "peer"):
        raise ValueError(f"unknown migrate mode {migrate!r}")
    if migrate == "peer":
        if not all(r.mig_ready for r in ranks):
            raise RuntimeError("migrate='peer' needs slab.setup_ghost_peers(ranks, comm) first")
        # every rank stores its row of the count matrix into every rank, then ONE host read: all ranks
        # hold the same matrix, which sizes the messages and carries every rank's error state
        for r in ranks:
            r.migrate_counts_push()
        mats = [r.migrate_counts_read() for r in ranks]
        P = ranks[0].nranks
        m = mats[0]
        cnt, tmo, over, room = m[:, :P], m[:, P], m[:, P + 1], m[:, P + 2]
        arrive = cnt.sum(axis=0)
        if tmo.any() or over.any() or (arrive > room).any():
            raise SlabExchangeError(
                "slab step failed on rank(s): flag-wait timeouts %s, leave-list overflow %s, particle capacity "
                "exceeded %s (every rank raises this in the same step)"
                % (np.nonzero(tmo)[0].tolist(), np.nonzero(over)[0].tolist(), np.nonzero(arrive > room)[0].tolist()))
        offs = np.cumsum(cnt, axis=0) - cnt          # offs[s, d]: records of ranks below s towards d
        for r in ranks:
            r.migrate_push(cnt[r.rank], offs[r.rank])
        for r in ranks:
            r.migrate_wait()
        for r in ranks:
            r.migrate_unpack(int(arrive[r.rank]), int(cnt[r.rank].sum()))
    else:
        # the leave counts are exchanged on the device; one small device->host read per step brings both
        # count vectors back (they size the messages)
        send_counts, recv_counts = comm.exchange_count_tensors([r.buf["LEAVE_COUNTS"] for r in ranks])
        # An overflow is local knowledge (a leave list beyond its capacity, arrivals beyond the free room); a rank
        # that raised alone would leave the others inside the all-to-all-v.  One flag all-reduce makes it common.
        bad = [r.rank for r, sc, rc in zip(ranks, send_counts, recv_counts)
               if max(sc) > r.leave_capacity() or r.entries() + sum(rc) > r.np_capacity]
        if comm.agree_any(bool(bad)):
            raise SlabExchangeError("slab step failed: leave-list or particle capacity exceeded on some rank "
                                    "(here: %s); every rank raises this in the same step" % (bad or "none"))
        for r, sc in zip(ranks, send_counts):
            r.migrate_pack(sc)
        comm.all_to_all_v(B("MIG_SEND"), send_counts, B("MIG_RECV"), recv_counts)
        for r, sc, rc in zip(ranks, send_counts, recv_counts):
            r.migrate_unpack(sum(rc), sum(sc))
    mark("migrate")
    if timer:
        timer.end_step()


def make_ranks(n_cells, pos, vel, comm, device=None, slack=1.25, ids=None):
    """Create the SlabRank(s) of this process and load their share of the particles.
    pos/vel: the FULL particle set (CUDA tensors, original order) -- every process passes the
    same arrays (tests, bench); ids default to the original particle index."""
    P = comm.nranks
    dev = torch.cuda.current_device() if device is None else device
    npart = pos.shape[1]
    owner = slab_of_particles(pos[2], n_cells, P)
    if ids is None:
        ids = torch.arange(npart, dtype=torch.int32, device=pos.device)
    cap = int(npart / P * slack) + 4096
    out = []
    for r in comm.local_ranks:
        sel = owner == r
        sr = SlabRank(n_cells, max(cap, int(sel.sum().item()) + 4096), dev, r, P)
        sr.load(pos[:, sel].contiguous(), vel[:, sel].contiguous(), ids[sel].contiguous())
        sr.total_particles = int(npart)
        out.append(sr)
    return out


def make_rank_from_local(n_cells, pos_l, vel_l, ids_l, rank, nranks, device=None, capacity=None, total_particles=None):
    """A SlabRank loaded with particles the caller already knows belong to `rank` (e.g. initial
    conditions generated per slab).  total_particles: the particle count over ALL ranks (fixes the
    mean density the transform subtracts); None = not subtracted."""
    dev = torch.cuda.current_device() if device is None else device
    n = pos_l.shape[1]
    cap = int(capacity) if capacity else int(n * 1.25) + 4096
    sr = SlabRank(n_cells, max(cap, n + 4096), dev, rank, nranks)
    sr.load(pos_l.contiguous(), vel_l.contiguous(), ids_l.contiguous())
    sr.total_particles = None if total_particles is None else int(total_particles)
    return sr


def collect(ranks, comm, npart):
    """Gather every rank's particles back into original particle order (tests / snapshots).
    LocalComm only needs concatenation; DistComm uses all_gather of padded buffers."""
    parts = [r.export() for r in ranks]
    dev = parts[0][0].device
    pos = torch.empty((3, npart), dtype=torch.float32, device=dev)
    vel = torch.empty_like(pos)
    if isinstance(comm, LocalComm):
        for p, v, i in parts:
            pos[:, i.long()] = p
            vel[:, i.long()] = v
        return pos, vel
    dist = comm.dist
    p, v, i = parts[0]
    n = torch.tensor([p.shape[1]], dtype=torch.int64, device=dev)
    counts = [torch.zeros_like(n) for _ in range(comm.nranks)]
    dist.all_gather(counts, n, group=comm.group)
    nmax = int(max(c.item() for c in counts))
    pad = torch.zeros((7, nmax), dtype=torch.float32, device=dev)
    pad[0:3, :p.shape[1]], pad[3:6, :p.shape[1]] = p, v
    pad[6, :p.shape[1]] = i.view(torch.float32)
    bufs = [torch.empty_like(pad) for _ in range(comm.nranks)]
    dist.all_gather(bufs, pad, group=comm.group)
    for b, c in zip(bufs, counts):
        k = int(c.item())
        idx = b[6, :k].contiguous().view(torch.int32).long()
        pos[:, idx] = b[0:3, :k]
        vel[:, idx] = b[3:6, :k]
    return pos, vel
