"""Initial conditions generated slab by slab (slab_ic.py over pm_ic_slab_* of include/pmstep.h)
against the package's single-GPU generator (gaussian_random_field() + zeldovich(), itself held to
the IC oracle and the reference's golden outputs by tests/test_gpu_ic.py): P = 1, 2, 4 ranks on
ONE GPU (slab.LocalComm), union of the slabs = the single-GPU particle set.  The arithmetic is
float64 with float32 outputs and only the order of the partial transforms differs, so all but a
handful of values are bit-identical; the bound is 1e-6 like the other IC tests."""
import importlib
import types

import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import oracle_ic as IC  # noqa: E402


@pytest.fixture()
def pm():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import cosmological_particle_mesh_simulation_b200 as pm
    yield pm
    pm.set_config(None)
    pm.release_plans()


def _ns(cfg, seed=38):
    d = dict(cfg.__dict__)
    d.update(RANDOM_SEED=seed, STEPS=100, A_END=1.0)
    return types.SimpleNamespace(**d)


def _single_gpu_ic(pm):
    G = importlib.import_module("cosmological_particle_mesh_simulation_b200.gaussian_random_field")
    Z = importlib.import_module("cosmological_particle_mesh_simulation_b200.zeldovich")
    rho = G.gaussian_random_field()
    pos, vel = Z.zeldovich(rho)
    return rho, pos, vel


@pytest.mark.parametrize("P", [1, 2, 4])
@pytest.mark.parametrize("n_parts,n_cells,power", [(32, 64, 1.0), (64, 128, 1.0), (16, 32, -1.5)])
def test_union_of_slabs_is_the_single_gpu_particle_set(pm, P, n_parts, n_cells, power):
    cfg = IC.ICConfig(N_PARTS=n_parts, N_CELLS=n_cells, POWER=power)
    pm.set_config(_ns(cfg))
    rho, pos, vel = _single_gpu_ic(pm)
    S = pm.slab_ic
    comm = pm.slab.LocalComm(P)
    ops = [S.DeviceOps(pm.config(), torch.cuda.current_device(), 38) for _ in range(P)]

    nl = n_parts // P
    # the field, held as columns
    dens = S.slab_gaussian_random_field(comm, ops)
    for r in range(P):
        want = rho[:, :, r * nl:(r + 1) * nl]
        assert dens[r].shape == want.shape and dens[r].dtype == torch.float32
        assert float((dens[r].double() - want.double()).norm() / want.double().norm()) <= 1e-6
    same = sum(int((dens[r] == rho[:, :, r * nl:(r + 1) * nl]).sum()) for r in range(P))
    assert same >= 0.999 * n_parts ** 3

    # the particles, before routing: rank r holds lattice columns [r nl, (r+1) nl) in (i0, i1, i2 local) order
    parts = S.slab_zeldovich(dens, comm, ops)
    i0, i1, i2 = torch.meshgrid(torch.arange(n_parts), torch.arange(n_parts), torch.arange(nl), indexing="ij")
    for r, (p, v, ids) in enumerate(parts):
        want_ids = ((i0 * n_parts + i1) * n_parts + (i2 + r * nl)).reshape(-1).to(torch.int32).cuda()
        assert torch.equal(ids, want_ids)

    # routed: every particle exactly once, on its owner, equal to the single-GPU generator's
    routed = S.route_to_owners(parts, comm, n_cells)
    seen = torch.zeros(n_parts ** 3, dtype=torch.int64, device="cuda")
    ident = 0
    for r, (p, v, ids) in enumerate(routed):
        i = ids.long()
        seen[i] += 1
        assert bool((pm.slab.slab_of_particles(p[2], n_cells, P) == r).all())
        d = torch.remainder(p.double() - pos[:, i].double() + n_cells / 2, n_cells) - n_cells / 2
        assert float(d.norm() / pos.double().norm()) <= 1e-6
        assert float(d.abs().max()) <= 2e-5 * n_cells / 64 + 1e-5
        assert float((v.double() - vel[:, i].double()).norm() / vel.double().norm()) <= 1e-6
        ident += int((p == pos[:, i]).sum()) + int((v == vel[:, i]).sum())
        assert float(p.min()) >= 0.0 and float(p.max()) <= n_cells
    assert bool((seen == 1).all())
    assert ident >= 0.999 * 6 * n_parts ** 3
    # a second run gives the same bits (counter-based noise and jitter, fixed transposes)
    again = S.slab_initial_conditions(comm, cfg=pm.config(), seed=38)
    for (p, v, i), (p2, v2, i2_) in zip(routed, again):
        assert torch.equal(p, p2) and torch.equal(v, v2) and torch.equal(i, i2_)


def test_noise_range_is_a_slice_of_the_single_gpu_noise(pm):
    cfg = IC.ICConfig(N_PARTS=16, N_CELLS=32)
    pm.set_config(_ns(cfg))
    from cosmological_particle_mesh_simulation_b200 import _runtime as rt
    G = importlib.import_module("cosmological_particle_mesh_simulation_b200.gaussian_random_field")
    f1, f2 = G.gaussian_random_numbers()
    e0, n = 5 * 256 + 3, 1000
    a = torch.empty(n, dtype=torch.float32, device="cuda")
    b = torch.empty_like(a)
    rt.check(rt.lib().pm_ic_noise_range(a.data_ptr(), b.data_ptr(), e0, n, 38, rt.stream_ptr(0)), "pm_ic_noise_range")
    torch.cuda.synchronize()
    assert torch.equal(a, f1.flatten()[e0:e0 + n]) and torch.equal(b, f2.flatten()[e0:e0 + n])


@pytest.mark.parametrize("axes,dims", [(0, (1, 2)), (1, (0,)), (2, (0, 1)), (3, (2,))])
def test_partial_transforms_against_torch_fft(pm, axes, dims):
    """pm_ic_slab_fft on a non-cubic array (the advanced cuFFT layouts of the strided cases)."""
    from cosmological_particle_mesh_simulation_b200 import _runtime as rt
    g = torch.Generator(device="cuda").manual_seed(4)
    z = torch.randn((6, 8, 10), dtype=torch.complex128, device="cuda", generator=g)
    for inverse in (0, 1):
        got = z.clone()
        rt.check(rt.lib().pm_ic_slab_fft(got.data_ptr(), 6, 8, 10, axes, inverse, rt.stream_ptr(0)), "pm_ic_slab_fft")
        want = torch.fft.ifftn(z, dim=dims, norm="forward") if inverse else torch.fft.fftn(z, dim=dims)
        assert float((got - want).abs().max()) <= 1e-12 * float(want.abs().max())


def test_slab_ic_feeds_the_slab_step(pm):
    """make_ranks_from_ic -> slab_step on P = 2 ranks against the single-GPU step on the single-GPU
    initial conditions."""
    n_parts, n_cells, P = 32, 64, 2
    cfg = IC.ICConfig(N_PARTS=n_parts, N_CELLS=n_cells)
    ns = _ns(cfg)
    pm.set_config(ns)
    _, pos, vel = _single_gpu_ic(pm)
    comm = pm.slab.LocalComm(P)
    ranks = pm.slab_ic.make_ranks_from_ic(comm)
    npart = n_parts ** 3
    assert sum(r.count for r in ranks) == npart
    mass = (n_cells / n_parts) ** 3
    a, da = 0.01, 0.0099
    for _ in range(3):
        pm.step(pos, vel, a, da, mass=mass)
        pm.slab.slab_step(ranks, comm, a, da, mass=mass, cfg=ns)
        a += da
    got_p, got_v = pm.slab.collect(ranks, comm, npart)
    d = torch.remainder(got_p.double() - pos.double() + n_cells / 2, n_cells) - n_cells / 2
    assert float(d.norm() / pos.double().norm()) <= 1e-5
    assert float((got_v.double() - vel.double()).norm() / vel.double().norm()) <= 1e-5
    for r in ranks:
        r.close()


def test_full_slab_run_from_slab_ic_matches_single_gpu_power_spectrum(pm):
    """BASELINE configs[0] sizes (64^3 on 128^3, 99 iterations) end to end on P = 4 slabs: initial
    conditions generated per slab, every step through slab_step (deposit ghosts, distributed FFT,
    phi ghosts, migration), against the single-GPU run from the single-GPU initial conditions:
    particle count conserved, every particle on its owner, P(k) of the final density within 0.1 %
    (the north_star's acceptance figure)."""
    n_parts, n_cells, P = 64, 128, 4
    cfg = IC.ICConfig(N_PARTS=n_parts, N_CELLS=n_cells)
    ns = _ns(cfg)
    pm.set_config(ns)
    _, pos, vel = _single_gpu_ic(pm)
    comm = pm.slab.LocalComm(P)
    ranks = pm.slab_ic.make_ranks_from_ic(comm, slack=1.5)
    npart = n_parts ** 3
    mass = (n_cells / n_parts) ** 3
    state = pm.ResidentParticles(pos, vel)
    nsteps = 0
    for a, da in pm.loop_scale_factors(ns):
        state.step(a, da)
        pm.slab.slab_step(ranks, comm, a, da, mass=mass, cfg=ns)
        nsteps += 1
    assert nsteps == 99
    assert sum(r.count for r in ranks) == npart
    state.store(pos, vel)
    got_p, got_v = pm.slab.collect(ranks, comm, npart)
    for r in ranks:
        p, _, _ = r.export()
        assert bool((pm.slab.slab_of_particles(p[2], n_cells, P) == r.rank).all())
    _, p1 = pm.analysis.power_spectrum(pm.density(pos, mass).clone())
    _, ps = pm.analysis.power_spectrum(pm.density(got_p, mass).clone())
    assert float((ps / p1 - 1.0).abs().max()) <= 1e-3
    for r in ranks:
        r.close()
