"""world_size-2 gloo test (CPU) of the exchange patterns the slab step uses: ring shift of the
ghost planes, all-to-all of the packed spectrum, count exchange and all-to-all-v of migrating
particle records -- slab.DistComm against slab.LocalComm on the same data."""
import os
import subprocess
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import sys, torch, torch.distributed as dist
sys.path.insert(0, %r)
from cosmological_particle_mesh_simulation_b200 import slab
dist.init_process_group("gloo")
r, P = dist.get_rank(), dist.get_world_size()
comm = slab.DistComm()
assert comm.local_ranks == [r] and comm.nranks == P
g = torch.Generator().manual_seed(0)
# what every rank would hold, generated identically everywhere
planes = [torch.rand(4, 4, generator=g) for _ in range(P)]
chunks = [torch.rand(P, 6, generator=g) for _ in range(P)]
counts = [[(3 * s + d) %% 4 for d in range(P)] for s in range(P)]
recs = [torch.rand(sum(counts[s]), 7, generator=g) for s in range(P)]
# reference: the in-process rank loop
loc = slab.LocalComm(P)
up = [torch.empty(4, 4) for _ in range(P)]; loc.shift(planes, up, +1)
dn = [torch.empty(4, 4) for _ in range(P)]; loc.shift(planes, dn, -1)
a2a = [torch.empty(P, 6) for _ in range(P)]; loc.all_to_all(chunks, a2a)
rc = loc.exchange_counts(counts)
rv = [torch.zeros(16, 7) for _ in range(P)]; loc.all_to_all_v(recs, counts, rv, rc)
# distributed
x = torch.empty(4, 4); comm.shift([planes[r]], [x], +1); assert torch.equal(x, up[r])
x = torch.empty(4, 4); comm.shift([planes[r]], [x], -1); assert torch.equal(x, dn[r])
y = torch.empty(P, 6); comm.all_to_all([chunks[r]], [y]); assert torch.equal(y, a2a[r])
c = comm.exchange_counts([counts[r]]); assert c[0] == rc[r], (c, rc)
sc2, rc2 = comm.exchange_count_tensors([torch.tensor(counts[r], dtype=torch.int32)])   # one read per step
assert sc2[0] == list(counts[r]) and rc2[0] == rc[r], (sc2, rc2)
ls, lr = loc.exchange_count_tensors([torch.tensor(c_, dtype=torch.int32) for c_ in counts])
assert ls == [list(c_) for c_ in counts] and lr == rc
z = torch.zeros(16, 7); comm.all_to_all_v([recs[r]], [counts[r]], [z], c); assert torch.equal(z, rv[r])
# an error seen by one rank becomes every rank's (the NCCL migration path agrees on overflows this way)
assert comm.agree_any(r == 1) is True and comm.agree_any(False) is False and loc.agree_any(True) is True
assert slab.slab_of_particles(torch.tensor([0.0, 15.9, 16.0, 32.0]), 32, 2).tolist() == [0, 0, 1, 0]
dist.barrier(); dist.destroy_process_group()
sys.stdout.write("rank" + str(r) + "-ok\n"); sys.stdout.flush()
'''


def test_dist_comm_matches_local_comm_world2(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % REPO)
    env = dict(os.environ, OMP_NUM_THREADS="1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29533", str(script)],
                         capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    assert out.stdout.count("-ok") == 2, out.stdout
