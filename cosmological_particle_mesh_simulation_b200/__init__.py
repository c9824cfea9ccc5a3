"""B200-native particle-mesh timestep: a drop-in for the per-step loop of
grkooij/Cosmological-Particle-Mesh-Simulation (reference: src/pmesh.py:56-63).

The module names and call signatures mirror the reference's flat source tree:

    density.density(positions, mass)                       src/density.py:8
    fourier_utils.fourier_grid()                           src/fourier_utils.py:5
    potential.potential(density, fgrid, a)                 src/potential.py:7
    integrate.advance_time(density, positions, velocities, fgrid, a, da)   src/integrate.py:9
    integrate.integrate(positions, velocities, a_val, f_a1, da, potentials) src/integrate.py:16
    cosmology.f(a, cosmology)                              src/cosmology.py:20
    configure_me.*                                         src/configure_me.py:7-40
    gaussian_random_field.gaussian_random_field()          src/gaussian_random_field.py:9   (initial conditions)
    zeldovich.zeldovich(density)                           src/zeldovich.py:10
    save_data.save_file(rho, positions, velocities, step, a), from_file(step)   src/save_data.py:7,29
    plot_helper.plot_step / plot_grf / plot_projection / project               src/plot_helper.py:12-72
    pmesh.simulator()                                      src/pmesh.py:18  (the whole driver)

All arithmetic runs in libpmstep.so (hand-written sm_100a CUDA + cuFFT) through the C ABI of
include/pmstep.h.  CUDA tensors in -> CUDA tensors out (state stays in HBM); NumPy arrays in ->
NumPy arrays out (uploaded, computed on the GPU, downloaded).  There is no CPU fallback.
"""
from . import configure_me, cosmology  # noqa: F401
from ._runtime import PMStepError, set_config, config, launch_count, release_plans, set_poisson_options, poisson_options  # noqa: F401
from .density import density  # noqa: F401
from .fourier_utils import fourier_grid, FourierGrid  # noqa: F401
from .potential import potential  # noqa: F401
from .integrate import advance_time, integrate  # noqa: F401
from .cosmology import f  # noqa: F401
from .pmesh import step, step_host, simulator, run, run_slabs, loop_scale_factors, ResidentParticles  # noqa: F401
from . import slab, slab_ic, analysis, _session  # noqa: F401,E402
from ._session import forget as forget_resident, set_enabled as set_resident_dropin, sync as sync_particles  # noqa: F401,E402
from .gaussian_random_field import gaussian_random_field  # noqa: F401,E402
from .zeldovich import zeldovich  # noqa: F401,E402
from .save_data import save_file, from_file  # noqa: F401,E402
from . import save_data, plot_helper  # noqa: F401,E402
