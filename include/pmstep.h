/*
 * pmstep.h -- C ABI of libpmstep.so: a B200 (sm_100a) particle-mesh timestep.
 *
 * This is the drop-in boundary for the per-timestep loop of
 * grkooij/Cosmological-Particle-Mesh-Simulation (reference: src/pmesh.py:56-63).  The reference
 * has no FFI layer of its own (it is Python calling numba-jitted functions and pyFFTW), so each
 * entry point below replaces one reference *callable* and cites it; INTEGRATION.md shows the
 * ctypes stub a maintainer of the reference would add.
 *
 * Conventions
 *   - plain C, no C++/torch types; every pointer is a raw address, every size explicit;
 *   - "_d" pointers are device memory of the plan's device, "_h" pointers are host memory;
 *   - particle arrays are SoA float32[3][np], row 0 = x (zeldovich.py:12-13); meshes are
 *     float32[Nc][Nc][Nc] indexed [z][y][x], x contiguous (density.py:11,37);
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream); all device
 *     entry points are asynchronous on that stream and never synchronise, the *_host entry
 *     points return after their results are in host memory;
 *   - return value: 0 on success, a positive cudaError_t, or one of the negative PM_ERR_* codes;
 *     nothing throws, nothing prints;
 *   - no allocation after pm_plan_create(): all scratch lives in the plan's workspace, whose
 *     size pm_plan_workspace_bytes() reports in advance;
 *   - there is no CPU fallback: without a CUDA device every entry point fails with
 *     PM_ERR_NO_DEVICE or the cudaError_t it hit.
 */
#ifndef PMSTEP_H
#define PMSTEP_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* The library is built with -fvisibility=hidden; only these entry points are exported. */
#if defined(__GNUC__)
#define PM_API __attribute__((visibility("default")))
#else
#define PM_API
#endif

#define PM_OK 0
#define PM_ERR_INVALID (-1)     /* bad argument (NULL pointer, size out of range, np > capacity) */
#define PM_ERR_UNSUPPORTED (-2) /* configuration outside what this build handles            */
#define PM_ERR_NOMEM (-3)       /* workspace allocation failed                              */
#define PM_ERR_CUFFT (-4)       /* cuFFT returned an error (pm_last_cufft_status())         */
#define PM_ERR_NO_DEVICE (-5)   /* no usable CUDA device                                    */

typedef struct pm_plan pm_plan; /* opaque */
typedef void *pm_stream_t;      /* cudaStream_t */

/* Library identification; counts kernels launched by this library in this process (bench.py's
 * gpu_launches). */
PM_API const char *pm_version(void);
PM_API const char *pm_error_string(int code);
PM_API int pm_last_cufft_status(void);
PM_API uint64_t pm_launch_count(void);

/*
 * Plan = everything fourier_grid() stands for in the reference (src/fourier_utils.py:5-16,
 * called once at src/pmesh.py:54) plus the scratch the reference re-allocates every step
 * (potential.py:19,26; integrate.py:33,61-62): cuFFT R2C/C2R plans, the sin^2(pi i/Nc) table of
 * the Green's function, sort buffers, row offsets, the half-spectrum and one mesh.
 *
 *   n_cells      N_CELLS  (configure_me.py:8), 4 <= n_cells and n_cells^3 < 2^32
 *   np_capacity  largest particle count later calls may pass (N_PARTS^3, configure_me.py:7)
 *   device       CUDA device ordinal, or -1 for the current device
 */
PM_API size_t pm_plan_workspace_bytes(int n_cells, int64_t np_capacity);
PM_API int pm_plan_create(pm_plan **plan, int n_cells, int64_t np_capacity, int device);
PM_API int pm_plan_destroy(pm_plan *plan);
PM_API int pm_plan_n_cells(const pm_plan *plan);
/* Poisson backend: 0 = hand-written sm_100a FFT with the Green's function fused into the z pass
 * (default for power-of-two meshes 32..1024), 1 = cuFFT R2C/C2R around a separate Green's kernel
 * (any mesh size; also selectable with the environment variable PM_FFT_BACKEND=cufft), 2 = DIAGNOSTIC:
 * the reference's own transform precision -- float64 D2Z/Z2D through cuFFT between the float32 density
 * and the float32 potential (src/potential.py:12-29 works in complex128); single-GPU meshes up to 256^3,
 * buffers allocated on first use.  It lets a test change the transform precision and nothing else. */
PM_API int pm_plan_set_fft_backend(pm_plan *plan, int backend);
/* The Green's function of src/fourier_utils.py:5-16 is 1/((s[z] + s[y]) + s[x]) with the Nc-entry
 * table s[i] = sin^2(k_i/2), k_i = float32(2*pi*fftfreq(Nc)[i]), all in float32.  A plan computes s in
 * float64 and rounds once (within 2 ulp of the reference's NumPy float32 sine, which no C library
 * reproduces bit for bit).  A caller that has the reference's own values -- the Python layer evaluates
 * the reference's three NumPy expressions -- installs them here (sin2_h: n_cells floats, host memory);
 * G is then the reference's table bit for bit (correctly rounded float32 sums and reciprocal). */
PM_API int pm_plan_set_sin2_table(pm_plan *plan, const float *sin2_h);
/* Options of the Poisson solve that the reference does not have (BASELINE north_star (2)), both 0
 * by default = the reference's scheme (src/potential.py:12-15, src/integrate.py:84-91) and the parity mode:
 *   deconvolve       0, 1 or 2: phi_k is divided by W(k)^deconvolve, W(k) = prod_i [sin(k_i/2)/(k_i/2)]^2 the
 *                    window of the cloud-in-cell assignment (1: the deposit's smoothing, 2: the force
 *                    interpolation's too)
 *   kspace_gradient  1: accelerations by spectral differentiation, -i k_d phi_k transformed into three force
 *                    meshes that pm_gather_kick_drift / the step entry points interpolate with the CIC
 *                    weights, instead of central differences of phi at the eight corners
 * With an option on, the solve uses the library transforms around fused k-space kernels (one pass applies
 * Green's function, the constants and the deconvolution; one pass per axis forms the gradient spectrum), the
 * gather is the one-thread-per-particle kernel, and the step is not graph-replayed.  Single-GPU plans only
 * (PM_ERR_UNSUPPORTED on slab plans); the force meshes (3 x 4 Nc^3 bytes + one spectrum) are allocated by this
 * call, not during a step. */
PM_API int pm_plan_set_poisson_options(pm_plan *plan, int deconvolve, int kspace_gradient);
PM_API int pm_plan_poisson_options(const pm_plan *plan, int *deconvolve, int *kspace_gradient);
PM_API int pm_plan_fft_backend(const pm_plan *plan);
/* Hand-written FFT, meshes 256..1024: fuse != 0 runs the row pass and the y pass of each
 * direction in one persistent launch that keeps the intermediate plane in L2 (pm_fft.cu,
 * k_fft_plane; also PM_FFT_FUSE=1); 0 (default: measured 4 % faster at 512^3, the passes are
 * bound on the SM side, not by HBM) runs them as two launches.  lag > 0 sets the distance in
 * planes between the two passes (default 12, PM_FFT_LAG).  Results are identical either way.
 * pm_plan_fft_sync_errors: 1 if a wait inside the fused launch ever gave up (a bug), else 0;
 * synchronises the device. */
PM_API int pm_plan_set_fft_fuse(pm_plan *plan, int fuse, int lag);
/* Hand-written FFT, meshes 256..1024, kernel selection (same mathematics, results agree to float32
 * rounding; 1, 3 and 4 are bit-identical to each other):
 *   0  three-stage radix-8 kernels for every pass (also PM_FFT_V2=0)
 *   1  register-resident two-stage transforms (one shared-memory exchange per 1-D transform,
 *      pm_fft2.cuh) for the row and y passes, radix-8 kernel for the fused z + Green + z^-1 pass
 *   2  two-stage kernels for every pass (also PM_FFT_ZMIX=0)
 *   3  (default) as 1, with the y passes run by the persistent cp.async-pipelined kernel of
 *      pm_fft3.cuh, ring of 2 tiles per SM (PM_FFT_V3=2; PM_FFT_V3=0 gives 1)
 *   4  as 3 with a ring of 3 tiles (PM_FFT_V3=3) */
PM_API int pm_plan_set_fft_variant(pm_plan *plan, int two_stage);
PM_API int pm_plan_fft_sync_errors(pm_plan *plan);
/* Resident gather + kick + drift (pm_step_resident): tiled != 0 stages the potential through shared
 * memory (a CTA owns a block of particle rows and marches along z with a ring of phi slabs fed by
 * cp.async; csrc/pm_gather_tiled.cuh) on meshes of 128, 256 or 512 cells; 0 runs one thread per
 * particle with scattered loads (also PM_GATHER_TILED=0).  Default on: 0.50 ms against 0.55 ms at
 * 256^3 particles on 512^3 cells.  Bit-identical results. */
PM_API int pm_plan_set_gather_tiled(pm_plan *plan, int tiled);
/* Diagnostics of the resident gather under particle skew (bench.py reports them for the clustered
 * loads of BASELINE configs[4]).  pm_plan_gather_tile: the tile shape of the tiled kernel -- mesh
 * rows per CTA and the staged-particle capacity of one (z, row-block) step; a block holding more
 * particles than the capacity handles the excess in further rounds.  pm_plan_block_stats: over the
 * row table of the LAST sort (synchronises the stream), for blocks of `rows_per_block` consecutive
 * mesh rows: out[0] = blocks, out[1] = blocks with more than `cap` particles, out[2] = particles
 * beyond `cap` summed over those blocks, out[3] = particles of the fullest block.  No reference
 * counterpart (the reference never sorts, src/density.py:17). */
PM_API int pm_plan_gather_tile(const pm_plan *plan, int *rows_per_block, int *cap);
PM_API int pm_plan_block_stats(pm_plan *plan, int rows_per_block, int cap, int64_t *out4, pm_stream_t stream);
/* How the resident paths (pm_step_resident, pm_slab_deposit) order the particle list by cell key
 * (the order fixes the deposit's summation tree and the locality of deposit and gather; the
 * reference scatters in particle-index order, src/density.py:17, and has no counterpart):
 *   PM_SORT_AUTO  re-sort only the entries whose cell key changed since the previous step and merge
 *                 them into the still-sorted rest; falls back to PM_SORT_FULL when there is no
 *                 previous order or more than 40 % of the entries moved.  The mover count and that
 *                 decision stay on the device: no host synchronisation.
 *   PM_SORT_FULL  stable radix sort of every entry (also: environment variable PM_SORT=full).
 * Both give the same order bit for bit (own stable radix sort, csrc/pm_sort.cu).  pm_plan_sort_stats
 * reports what the last sort did (mode: PM_SORT_FULL or PM_SORT_INCREMENTAL); it reads the device and
 * therefore synchronises. */
#define PM_SORT_AUTO 0
#define PM_SORT_FULL 1
#define PM_SORT_INCREMENTAL 2
#define PM_SORT_ERROR 3        /* pm_plan_sort_stats only: a grid barrier of the radix sort timed out */
PM_API int pm_plan_set_sort_mode(pm_plan *plan, int mode);
PM_API int pm_plan_sort_stats(const pm_plan *plan, int64_t *entries, int64_t *movers, int *mode);
PM_API int64_t pm_plan_np_capacity(const pm_plan *plan);

/*
 * fourier_grid() made explicit (src/fourier_utils.py:5-16): writes the float32[Nc^3] table
 * 1/(sin^2(kz/2)+sin^2(ky/2)+sin^2(kx/2)) with the DC entry set to 0 (the reference leaves it
 * uninitialised).  Only for inspection/parity -- the solver never materialises this table.
 */
PM_API int pm_fourier_grid(pm_plan *plan, float *fgrid_d, pm_stream_t stream);

/*
 * Cell key of every particle: key = (z_c*Nc + y_c)*Nc + x_c with c = int(floor(pos)) mod Nc
 * (src/density.py:19-21,37).  keys_d: uint32[np].
 */
PM_API int pm_cell_keys(pm_plan *plan, const float *pos_d, int64_t np, uint32_t *keys_d,
                 pm_stream_t stream);

/*
 * Stable radix sort of the particles by cell key.  keys_sorted_d, order_d: uint32[np];
 * order_d[j] = original index of the j-th particle in cell order.  Either output may be NULL.
 */
PM_API int pm_sort_by_cell(pm_plan *plan, const float *pos_d, int64_t np, uint32_t *keys_sorted_d,
                    uint32_t *order_d, pm_stream_t stream);

/*
 * density(positions, mass) (src/density.py:7-48; called at src/pmesh.py:60): cloud-in-cell
 * deposit.  keys -> sort -> deterministic warp-segmented scatter; every cell of rho_d is written
 * exactly once (no memset, no atomics).  pos_d is not modified.
 */
PM_API int pm_deposit_cic(pm_plan *plan, const float *pos_d, int64_t np, double mass, float *rho_d,
                   pm_stream_t stream);

/*
 * potential(density, fgrid, a) (src/potential.py:7-29; called at src/integrate.py:11):
 * phi = IFFT( -3*omega_m0/(8a) * G(k) * FFT(rho) ), G from pm_fourier_grid, inverse normalised
 * by 1/Nc^3.  cuFFT R2C -> fused Green's kernel on the half spectrum -> cuFFT C2R, float32.
 * rho_d is not modified; phi_d may alias rho_d.
 */
PM_API int pm_poisson(pm_plan *plan, const float *rho_d, double a, double omega_m0, float *phi_d,
               pm_stream_t stream);

/*
 * Matter power spectrum of a density mesh (SURVEY 8f row f3 -- the reference has no estimator;
 * the acceptance check "P(k) within 0.1 %" needs one).  Reuses the forward half of the Poisson
 * transform: psum_d[b] = sum of w*|rho_k|^2 over the modes with round(|k|) == b (integer frequency
 * units), pcnt_d[b] = sum of w, w = 2 for the modes whose Hermitian partner is not stored.
 * P(b) = psum/pcnt / (mean(rho)^2 * Nc^6) is the spectrum of the density contrast.  Power-of-two
 * meshes, single-GPU plans.  The plan's spectrum scratch is overwritten; rho_d is not.
 */
PM_API int pm_power_spectrum(pm_plan *plan, const float *rho_d, int nbins, double *psum_d,
                             double *pcnt_d, pm_stream_t stream);

/*
 * integrate(positions, velocities, a_val, f_a1, da, potentials) (src/integrate.py:15-97):
 * CIC force gather (central difference of phi at the 8 corners), kick and drift with periodic
 * wrap, fused in one kernel; pos_d and vel_d are updated in place.  f_a1 is the host scalar
 * f(a+da, ...) of src/integrate.py:12 / src/cosmology.py:20-27.  acc_d (optional, float32[3][np])
 * receives g_p of src/integrate.py:92.
 */
PM_API int pm_gather_kick_drift(pm_plan *plan, float *pos_d, float *vel_d, int64_t np,
                         const float *phi_d, double a_val, double f_a1, double da, float *acc_d,
                         pm_stream_t stream);

/*
 * One body of the loop src/pmesh.py:60-61 on device-resident state:
 *   rho = density(pos, mass); pos, vel = advance_time(rho, pos, vel, fgrid, a, da)
 * rho_d (optional) receives the pre-step density (what the reference keeps for save/plot,
 * src/pmesh.py:67-73); when NULL the deposit lands in the plan's own mesh.
 */
PM_API int pm_step(pm_plan *plan, float *pos_d, float *vel_d, int64_t np, double mass, double a,
            double da, double f_a1, double omega_m0, float *rho_d, pm_stream_t stream);

/*
 * Resident particle state -- the fast way to run many steps.  The reference never permutes its
 * particle arrays (snapshots are in original order, src/save_data.py:19-24), but a step is far
 * cheaper when particles sit in HBM in cell order: the deposit and the force gather then stream.
 * pm_particles_load copies the caller's arrays (original order) into the plan; pm_step_resident
 * advances that state by one loop body (src/pmesh.py:60-61), keeping it sorted by the cell of the
 * previous positions and carrying each particle's original index; pm_particles_store scatters the
 * state back to the caller's arrays in ORIGINAL particle order.  load -> n x step -> store gives
 * the same particles as n x pm_step.  pm_particles_order returns the original index of the
 * particle in each storage slot (the sort order the parity tests check).
 */
PM_API int pm_particles_load(pm_plan *plan, const float *pos_d, const float *vel_d, int64_t np,
                             pm_stream_t stream);
PM_API int pm_step_resident(pm_plan *plan, double mass, double a, double da, double f_a1,
                            double omega_m0, float *rho_d, pm_stream_t stream);
PM_API int pm_particles_store(pm_plan *plan, float *pos_d, float *vel_d, pm_stream_t stream);
PM_API int pm_particles_order(pm_plan *plan, uint32_t *ids_d, pm_stream_t stream);
PM_API int64_t pm_particles_count(const pm_plan *plan);

/*
 * Slab-decomposed step for 2/4/8 GPUs of one box (no reference counterpart: the reference is
 * single-process; design in SURVEY.md 8e / DESIGN.md 6).  One plan per rank; rank r owns mesh planes
 * [r*Nc/P, (r+1)*Nc/P) along array axis 0 and the particles whose z cell lies in them.  The entry
 * points only compute; the caller moves the buffers pm_slab_buffer() names between ranks (NCCL
 * send/recv for ghost planes, NCCL all-to-all around the hand-written pack/unpack transposes of
 * the distributed FFT, all-to-all-v for migrating particles).  Order of calls: see csrc/pm_slab.cu.
 * Requires a power-of-two mesh with Nc/P a multiple of 16.
 */
#define PM_BUF_RHO 0             /* float[nzl][Nc][Nc]   density of the owned planes              */
#define PM_BUF_RHO_GHOST_SEND 1  /* float[Nc][Nc]        plane nzl of the deposit -> rank+1        */
#define PM_BUF_RHO_GHOST_RECV 2  /* float[Nc][Nc]        <- rank-1, added by pm_slab_ghost_add (aliases
                                    PM_BUF_PHI_LO_RECV: the phi buffer is dead at that point of the step) */
#define PM_BUF_FFT_SEND_MAIN 3   /* float2[P][nzl][nyl][Nc/2]  packed spectrum, chunk s -> rank s  */
#define PM_BUF_FFT_SEND_SIDE 4   /* float2[P][nzl][nyl]        packed Nyquist plane                */
#define PM_BUF_FFT_RECV_MAIN 5   /* float2[Nc][nyl][Nc/2]      transposed spectrum (z pass)        */
#define PM_BUF_FFT_RECV_SIDE 6   /* float2[Nc][nyl]                                               */
#define PM_BUF_PHI 7             /* float[nzl][Nc][Nc]   potential of the owned planes            */
#define PM_BUF_PHI_LO_SEND 8     /* float[2][Nc][Nc]     first two owned planes -> rank-1          */
#define PM_BUF_PHI_HI_SEND 9     /* float[Nc][Nc]        last owned plane -> rank+1                */
#define PM_BUF_PHI_LO_RECV 10    /* float[Nc][Nc]        ghost plane z0-1 <- rank-1's HI_SEND      */
#define PM_BUF_PHI_HI_RECV 11    /* float[2][Nc][Nc]     ghost planes z0+nzl, +1 <- rank+1's LO_SEND */
#define PM_BUF_MIG_SEND 12       /* float[.][7]          (x,y,z,vx,vy,vz,id) of leavers, by dest   */
#define PM_BUF_MIG_RECV 13       /* float[.][7]          arrivals                                  */
#define PM_BUF_LEAVE_COUNTS 14   /* uint32[P]            leavers per destination (after gather)    */
#define PM_BUF_PEER_FLAGS 15     /* uint32[25][16]       flag words of the peer-memory exchanges   */
PM_API int pm_plan_create_slab(pm_plan **plan, int n_cells, int64_t np_capacity, int device, int rank,
                               int nranks);
PM_API int pm_slab_buffer(pm_plan *plan, int which, void **ptr, size_t *bytes);
PM_API int pm_slab_load(pm_plan *plan, const float *pos_d, const float *vel_d, const uint32_t *ids_d,
                        int64_t np, pm_stream_t stream);
PM_API int64_t pm_slab_count(const pm_plan *plan);   /* live particles owned by this rank */
PM_API int64_t pm_slab_entries(const pm_plan *plan); /* storage entries incl. departed ones */
PM_API int pm_slab_deposit(pm_plan *plan, double mass, pm_stream_t stream);
PM_API int pm_slab_ghost_add(pm_plan *plan, pm_stream_t stream);
/* The distributed transform in C chunks of kx columns (chunk c occupies bytes [c, c+1)/C of the
 * FFT_*_MAIN buffers; the Nyquist plane FFT_*_SIDE travels with chunk 0): the all-to-all of one
 * chunk overlaps the y and z passes of its neighbours.  (Nc/2)/C must be a multiple of the column
 * tile width (16; 8 for Nc >= 1024). */
PM_API int pm_slab_fft_rows_forward(pm_plan *plan, pm_stream_t stream);
/* Mean of the density over the WHOLE mesh (all ranks' particles * mass / Nc^3).  The forward
 * transform runs on rho - mean: the Green's factor zeroes the DC mode anyway, and in
 * float32 the DC's rounding noise would otherwise leak into the lowest-k modes that 1/k^2
 * amplifies.  Single-GPU entry points know the mean themselves; a slab rank must be told. */
PM_API int pm_slab_set_rho_mean(pm_plan *plan, double mean);
PM_API int pm_slab_fft_y_forward(pm_plan *plan, int chunk, int nchunks, pm_stream_t stream);
PM_API int pm_slab_fft_z(pm_plan *plan, int chunk, int nchunks, double a, double omega_m0,
                         pm_stream_t stream);
PM_API int pm_slab_fft_y_inverse(pm_plan *plan, int chunk, int nchunks, pm_stream_t stream);
PM_API int pm_slab_fft_rows_inverse(pm_plan *plan, pm_stream_t stream);
/* Peer-memory transposes: instead of pack -> NCCL all-to-all -> unpack, the copy kernels store into
 * (push, forward leg) and load from (pull, way back) the z-pass arrays of the other ranks over
 * NVLink, and flag words replace the collective as the barrier.  Each rank publishes its workspace
 * once (pm_slab_peer_export -> 64-byte CUDA IPC handle + two offsets, exchanged by the caller) and
 * maps everybody else's (pm_slab_peer_import; its own entry needs no handle); plans that live in one
 * process use pm_slab_peer_set with the addresses of PM_BUF_FFT_RECV_MAIN / PM_BUF_PEER_FLAGS.
 * Slots 0..7: "chunk c pushed", 8..15: "z pass of chunk c done".  signal/wait pair up by call
 * count, so every rank must make the same calls every step.  A wait gives up after 2 s and counts
 * a timeout (pm_slab_peer_timeouts) instead of hanging the device.  Order of calls: csrc/pm_slab.cu. */
PM_API int pm_slab_peer_export(pm_plan *plan, void *handle64, uint64_t *recv_offset, uint64_t *flags_offset);
PM_API int pm_slab_peer_import(pm_plan *plan, int peer, const void *handle64, uint64_t recv_offset,
                               uint64_t flags_offset);
PM_API int pm_slab_peer_set(pm_plan *plan, int peer, void *recv_main, void *flags);
PM_API int pm_slab_peer_signal(pm_plan *plan, int slot, pm_stream_t stream);
PM_API int pm_slab_peer_wait(pm_plan *plan, int slot, pm_stream_t stream);
PM_API int pm_slab_peer_timeouts(pm_plan *plan, uint32_t *timeouts);
PM_API int pm_slab_peer_release(pm_plan *plan);   /* unmap the peers (all ranks, then a barrier, then destroy) */
PM_API int pm_slab_fft_y_forward_local(pm_plan *plan, int chunk, int nchunks, pm_stream_t stream);
PM_API int pm_slab_fft_push(pm_plan *plan, int chunk, int nchunks, pm_stream_t stream);
PM_API int pm_slab_fft_pull(pm_plan *plan, int chunk, int nchunks, pm_stream_t stream);
PM_API int pm_slab_fft_y_inverse_local(pm_plan *plan, int chunk, int nchunks, pm_stream_t stream);
/* ... or fused: the forward y pass stores its result straight into the peers' z-pass arrays and
 * the inverse y pass loads its input from them -- no copy kernel at all (same signal/wait calls). */
PM_API int pm_slab_fft_y_forward_push(pm_plan *plan, int chunk, int nchunks, pm_stream_t stream);
PM_API int pm_slab_fft_y_inverse_pull(pm_plan *plan, int chunk, int nchunks, pm_stream_t stream);
/* EXPERIMENTAL (not yet validated on hardware): ghost planes through the same peer mappings instead of
 * NCCL send/recv.  push = copy the plane(s) into the neighbour's phi buffer + one flag word to that
 * neighbour; wait = poll that one word.  Set-up: pm_slab_peer_ghost_export -> exchange ->
 * pm_slab_peer_ghost_import (after pm_slab_peer_import), or pm_slab_peer_ghost_set with the address
 * of PM_BUF_PHI_LO_RECV for plans in one process.  Per step: deposit, push_rho, wait_rho, ghost_add,
 * transform, push_phi, wait_phi, gather. */
PM_API int pm_slab_peer_ghost_export(pm_plan *plan, uint64_t *mesh2_offset);
PM_API int pm_slab_peer_ghost_import(pm_plan *plan, int peer, uint64_t mesh2_offset);
PM_API int pm_slab_peer_ghost_set(pm_plan *plan, int peer, void *phi_lo_recv);
PM_API int pm_slab_ghost_push_rho(pm_plan *plan, pm_stream_t stream);
PM_API int pm_slab_ghost_wait_rho(pm_plan *plan, pm_stream_t stream);
PM_API int pm_slab_ghost_push_phi(pm_plan *plan, pm_stream_t stream);
PM_API int pm_slab_ghost_wait_phi(pm_plan *plan, pm_stream_t stream);
PM_API int pm_slab_gather(pm_plan *plan, double a, double f_a1, double da, pm_stream_t stream);
/* Ghost planes and particle migration through peer memory (csrc/pm_migrate.cu; no reference counterpart,
 * SURVEY 8e).  After pm_slab_peer_import / pm_slab_peer_set, pm_slab_peer_aux_export / _import / _set publish
 * three more buffers of every rank: its potential buffer (ghost planes), its migration receive buffer and
 * its count matrix (offsets[3] inside the exported workspace, or plain pointers from pm_slab_aux_buffers for
 * plans inside one process).  Then, per step and in this order on every rank:
 *   pm_slab_migrate_counts_push   my leavers per destination + my error state -> every rank's matrix, flag
 *   pm_slab_migrate_counts_read   wait for all rows; copies the matrix to the host (the step's one host read):
 *                                 matrix_h[nranks][nranks + 3], row s = rank s's leavers per destination, its
 *                                 flag-wait timeouts, its leave-list overflow flag, its free particle capacity
 *   pm_slab_migrate_push          records of my leavers (ascending slot order) into the destinations' receive
 *                                 buffers at dest_offsets[d] (= leavers of lower ranks towards d), flag
 *   pm_slab_migrate_wait          wait for every sender; then pm_slab_migrate_unpack as with the NCCL route */
PM_API int pm_slab_peer_aux_export(pm_plan *plan, uint64_t *offsets3);
PM_API int pm_slab_peer_aux_import(pm_plan *plan, int peer, const uint64_t *offsets3);
PM_API int pm_slab_peer_aux_set(pm_plan *plan, int peer, void *phi_buffer, void *mig_recv, void *count_matrix);
PM_API int pm_slab_aux_buffers(pm_plan *plan, void **phi_buffer, void **mig_recv, void **count_matrix);
PM_API int pm_slab_migrate_counts_push(pm_plan *plan, pm_stream_t stream);
PM_API int pm_slab_migrate_counts_read(pm_plan *plan, uint32_t *matrix_h, pm_stream_t stream);
PM_API int pm_slab_migrate_push(pm_plan *plan, const int64_t *send_counts, const int64_t *dest_offsets,
                                pm_stream_t stream);
PM_API int pm_slab_migrate_wait(pm_plan *plan, pm_stream_t stream);
PM_API int pm_slab_migrate_pack(pm_plan *plan, const int64_t *counts_h, pm_stream_t stream);
PM_API int pm_slab_migrate_unpack(pm_plan *plan, int64_t n_arrive, int64_t n_leave, pm_stream_t stream);
/* live_d[s] = 0 for entries whose particle has left; rows are dense [3][pm_slab_entries()] */
PM_API int pm_slab_export(pm_plan *plan, float *pos_d, float *vel_d, uint32_t *ids_d, uint32_t *live_d,
                          pm_stream_t stream);

/* pm_step_resident replays its steady-state launch sequence as a CUDA graph (two graphs, one per buffer-set
 * parity; only a one-thread parameter kernel is re-parameterised per step) when the step is graphable: own
 * FFT, tiled gather (128/256/512 meshes), incremental sort, not being profiled, a non-default stream.
 * pm_plan_set_graph(plan, 0) (or PM_GRAPH=0) keeps it eager; pm_plan_graph_replays counts the replays. */
PM_API int pm_plan_set_graph(pm_plan *plan, int on);
PM_API int pm_plan_graph_replays(const pm_plan *plan);
/* Work list of the last tiled gather (diagnostic; synchronises): the CTAs' shares are (row block, z chunk)
 * columns filed whole ("light") or, where a column holds more than twice the mean, in pieces ("heavy",
 * dispatched first) -- csrc/pm_particles.cu, k_gather_items.  overflow != 0 would mean the list outgrew its
 * buffer (it cannot, by the bound it is sized with).  PM_GATHER_ITEMS=0 turns the list off. */
PM_API int pm_plan_gather_items(pm_plan *plan, int64_t *heavy, int64_t *light, int *overflow);
/* The two halves of pm_step_resident as separate calls, for callers that drive the reference's loop
 * body statement by statement (src/pmesh.py:60-61):
 *     rho = density(positions, mass)                          -> pm_resident_deposit
 *     positions, velocities = advance_time(rho, ..., a, da)   -> pm_resident_advance
 * pm_resident_deposit orders the resident state by cell (incrementally) and deposits it into rho_d;
 * calling it again before the state changes deposits again without re-sorting.  pm_resident_advance
 * solves for the potential of rho_d (any density mesh; rho_mean = its mean if the caller knows it --
 * total mass / Nc^3 for a density this library deposited -- or NaN to have it measured) and applies
 * gather + kick + drift to the resident state.  pm_particles_store then writes the state back in the
 * caller's original particle order.  Reference: src/density.py:7-48, src/integrate.py:9-13. */
PM_API int pm_resident_deposit(pm_plan *plan, double mass, float *rho_d, pm_stream_t stream);
PM_API int pm_resident_advance(pm_plan *plan, const float *rho_d, double rho_mean, double a, double da,
                               double f_a1, double omega_m0, pm_stream_t stream);
/*
 * The same loop body for a caller that keeps its state in host memory like the reference does
 * (NumPy arrays): uploads pos_h/vel_h, runs the step, downloads the updated pos_h/vel_h IN PLACE (and
 * rho_h when not NULL).  Pinned host buffers make the copies asynchronous and overlapped;
 * pageable ones work but are slower.  Returns after the results are in host memory.
 * The results are those of pm_step bit for bit.  Inside, the step is arranged around the two PCIe
 * directions: positions are uploaded first; sort, deposit and Poisson solve run
 * under the velocity upload; on 128^3 / 256^3 / 512^3 meshes the gather proper then stores every
 * particle's three stencil sums at its original index and kick + drift run in the caller's order,
 * range by range behind the arriving velocities, each range downloaded while the next is pushed --
 * so parts of vel_h are overwritten with results while later parts are still being read.
 * PM_HOST_SPLIT=0 (environment) selects the older route: fused gather, un-permute, download.
 * The call runs on the plan's own streams: it starts with a device synchronisation (work the caller
 * enqueued on other streams with this plan must be finished before the workspace is reused) and ends
 * with the results in host memory.
 */
PM_API int pm_step_host(pm_plan *plan, float *pos_h, float *vel_h, int64_t np, double mass, double a,
                 double da, double f_a1, double omega_m0, float *rho_h);
/* Diagnostic (host arithmetic only, no device needed): the particle range [*i0, *i1) that pm_step_host
 * uploads, pushes and downloads as range k of its *n_ranges ranges for np particles. */
PM_API int pm_step_host_range(int64_t np, int k, int64_t *i0, int64_t *i1, int *n_ranges);

/*
 * Page-locking of caller-owned host arrays (the reference keeps its state in NumPy arrays that live for
 * the whole run, src/pmesh.py:39-52: registering them once makes every later copy a direct DMA at the
 * PCIe rate instead of a staged pageable copy).  pm_host_register returns PM_OK, or PM_ERR_UNSUPPORTED
 * when the range cannot be registered (overlaps a registered range, unsupported memory ...) -- the
 * caller then simply uses the array as pageable memory; no CUDA error state is left behind.  The range
 * must be unregistered before the memory is freed.
 */
PM_API int pm_host_register(void *ptr, size_t bytes);
PM_API int pm_host_unregister(void *ptr);

/*
 * Per-stage device timing of pm_step (bench.py's live roofline).  pm_plan_profile_begin arms a
 * ring of CUDA events for up to max_steps calls of pm_step (0 disarms it); each armed call records
 * an event on the caller's stream between stages -- no synchronisation, no extra kernels.
 * pm_plan_profile_read synchronises on the last event and writes ms[step][PM_NUM_STAGES]
 * (row-major) for the *n_steps calls recorded since begin.
 */
#define PM_STAGE_KEYS 0    /* cell keys                         (own kernel)            */
#define PM_STAGE_SORT 1    /* radix sort by cell                (cub)                   */
#define PM_STAGE_ROWS 2    /* row offsets                       (own kernel)            */
#define PM_STAGE_DEPOSIT 3 /* warp-segmented CIC scatter        (own kernel)            */
#define PM_STAGE_R2C 4     /* forward FFT                       (cuFFT)                 */
#define PM_STAGE_GREEN 5   /* Green's function on half spectrum (own kernel)            */
#define PM_STAGE_C2R 6     /* inverse FFT                       (cuFFT)                 */
#define PM_STAGE_GATHER 7  /* force gather + kick + drift       (own kernel)            */
#define PM_NUM_STAGES 8
PM_API int pm_plan_profile_begin(pm_plan *plan, int max_steps);
PM_API int pm_plan_profile_read(pm_plan *plan, float *ms, int *n_steps);

/* ------------------------------------------------------------------------------------------------
 * Initial conditions (not part of the per-step path).  Replaces
 * gaussian_random_field() (src/gaussian_random_field.py:9-29) and zeldovich(density)
 * (src/zeldovich.py:10-22), called once from src/pmesh.py:40,42.  float64 / complex128 arithmetic
 * like the reference (cuFFT Z2Z for its two pyFFTW inverse transforms and its np.fft.fftn), float32
 * where the reference stores float32.  The two things the reference draws irreproducibly
 * are explicit arrays here, so a caller can also supply its own:
 *   pm_ic_noise   f1, f2: float32[n] standard normals by the reference's polar Box-Muller
 *                 (gaussian_random_field.py:31-63) from Philox4x32-10(seed, element index);
 *   pm_ic_jitter  float64[3][n] uniform(-2, 2) (zeldovich.py:89-91), same generator.
 * Entries the reference leaves uninitialised at k = 0 are 0.  All pointers are device pointers;
 * work_d needs pm_ic_workspace_bytes(n_parts) bytes.  cuFFT allocates its own plan work area inside
 * the calls (one-time setup code, unlike the per-step path).  The calls synchronise the stream.
 * ---------------------------------------------------------------------------------------------- */
typedef struct pm_ic_params {
    int n_parts, n_cells;     /* configure_me.N_PARTS, N_CELLS */
    double box_size;          /* BOX_SIZE [Mpc/h] */
    double power;             /* POWER */
    int lcdm_transfer;        /* LCDM_TRANSFER_FUNCTION */
    double omega_m0, omega_b0, omega_k0, omega_lambda0, h0, a_init;
} pm_ic_params;
PM_API size_t pm_ic_workspace_bytes(int n_parts);
PM_API int pm_ic_noise(float *f1_d, float *f2_d, int64_t n, uint64_t seed, pm_stream_t stream);
PM_API int pm_ic_jitter(double *jitter_d, int64_t n, uint64_t seed, pm_stream_t stream);
/* p_d: float64[n_parts^3], the grid power_spectrum() returns (gaussian_random_field.py:91-123) */
PM_API int pm_ic_power_spectrum(const pm_ic_params *prm, double *p_d, void *work_d, size_t work_bytes,
                                pm_stream_t stream);
/* density_d: float32[n_parts^3] = (ifftn(sqrt(p D^2) (f1 + i f2)).real).astype(float32) */
PM_API int pm_ic_gaussian_random_field(const pm_ic_params *prm, const float *f1_d, const float *f2_d,
                                       float *density_d, void *work_d, size_t work_bytes, pm_stream_t stream);
/* pos_d, vel_d: float32[3][n_parts^3] in the reference's particle order; jitter_d: float64[3][n_parts^3] */
PM_API int pm_ic_zeldovich(const pm_ic_params *prm, const float *density_d, const double *jitter_d,
                           float *pos_d, float *vel_d, void *work_d, size_t work_bytes, pm_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Initial conditions, one slab per rank (configs[2]-[3]: the 1024^3 lattice does not have to exist on
 * any one GPU).  The same generator as above, cut into the pieces a slab-decomposed run needs; the
 * host side (slab_ic.py) strings them together with two kinds of all-to-all transposes.  Replaces the
 * same reference functions (src/gaussian_random_field.py:9-29, src/zeldovich.py:10-100).
 *
 * Decomposition: k-space fields are held as planes [i0_lo, i0_lo + n0l) of the slowest array axis
 * ([n0l][n][n] complex128); real-space fields as columns [i2_lo, i2_lo + n2l) of the fastest axis
 * ([n][n][n2l]) -- the lattice coordinate that becomes positions[2], the axis the mesh slabs are cut
 * along (slab.py), so nearly every particle is generated on the rank that owns it.  A 3-D transform is
 * PM_IC_AXES_12 on k-space planes, transpose, PM_IC_AXIS_0 on real-space columns (inverse), or
 * PM_IC_AXES_01, transpose, PM_IC_AXIS_2 (forward).  Noise and jitter are functions of the GLOBAL
 * element / particle index, the power-spectrum normalisation is summed over the whole grid in the
 * single-GPU generator's order on every rank: the union of the slabs is the single-GPU particle set
 * up to the rounding of the differently ordered transforms.
 * ---------------------------------------------------------------------------------------------- */
#define PM_IC_AXES_12 0   /* 2-D transform over the two fastest axes, one per index of the slowest   */
#define PM_IC_AXIS_0 1    /* 1-D transform along the slowest axis (stride d1*d2)                     */
#define PM_IC_AXES_01 2   /* 2-D transform over the two slowest axes (stride d2), one per fastest    */
#define PM_IC_AXIS_2 3    /* 1-D transform along the fastest axis                                    */
PM_API size_t pm_ic_slab_workspace_bytes(void);
/* f1, f2: the elements [e0, e0 + n) of the fields pm_ic_noise(seed) fills */
PM_API int pm_ic_noise_range(float *f1_d, float *f2_d, int64_t e0, int64_t n, uint64_t seed, pm_stream_t stream);
/* zk_d: complex128 [n0l][n][n] = sqrt(p D^2) (f1 + i f2) on planes i0_lo .. i0_lo + n0l - 1; f1/f2: those planes */
PM_API int pm_ic_slab_rho_k(const pm_ic_params *prm, const float *f1_d, const float *f2_d, int i0_lo, int n0l,
                            void *zk_d, void *work_d, size_t work_bytes, pm_stream_t stream);
/* in-place complex128 transform of a contiguous [d0][d1][d2] array over `axes` (cuFFT Z2Z, unnormalised) */
PM_API int pm_ic_slab_fft(void *z_d, int d0, int d1, int d2, int axes, int inverse, pm_stream_t stream);
/* out[i] = float32(re(z[i]) * scale);   z[i] = in[i] + 0i */
PM_API int pm_ic_slab_real_f32(const void *z_d, int64_t count, double scale, float *out_d, pm_stream_t stream);
PM_API int pm_ic_slab_from_f32(const float *in_d, int64_t count, void *z_d, pm_stream_t stream);
/* out = (-i l_dir) (rho_k / -k^2) (N_CELLS / N_PARTS) on planes i0_lo .. (zeldovich.py:24-38, 56-69) */
PM_API int pm_ic_slab_displacement_k(const pm_ic_params *prm, int dir, const void *rho_k_d, int i0_lo, int n0l,
                                     void *out_d, pm_stream_t stream);
/* z_d: complex128 [n][n][n2l], the unnormalised inverse transform of the above on columns i2_lo ..;
 * pos_d, vel_d: float32[n*n*n2l] (row `dir` of the caller's [3][..] arrays), ids_d (optional): the
 * single-GPU particle index (i0 n + i1) n + i2; jitter_d (optional): float64[n*n*n2l] in local order,
 * default the pm_ic_jitter(seed) stream of the particle's global index */
PM_API int pm_ic_slab_particles(const pm_ic_params *prm, int dir, const void *z_d, int i2_lo, int n2l, uint64_t seed,
                                const double *jitter_d, float *pos_d, float *vel_d, int32_t *ids_d, pm_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* PMSTEP_H */
