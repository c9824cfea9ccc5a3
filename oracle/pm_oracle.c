/*
 * pm_oracle.c -- CPU restatement of the reference's particle-mesh step.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this
 * file.  The CUDA path in cosmological_particle_mesh_simulation_b200/ never calls it.
 *
 * Parity status: PINNED.  The reference ships no tests or golden vectors, so the
 * pin is the reference's own code executed in the build container: oracle/make_golden.py
 * imports /root/reference/src/{density,potential,integrate,fourier_utils,cosmology}.py
 * unmodified (numba, 1 thread, pyfftw replaced by a scipy.fft shim) and writes
 * tests/golden/ *.npz; tests/test_oracle_golden.py demands this file reproduce those outputs
 * bit-for-bit (density, positions, velocities) given the same inputs.
 *
 * Every function follows the reference's structure AND its temporaries (cell-centre table,
 * weight table, per-direction copies), because the threaded build of this file is also the
 * "port" CPU baseline timed by bench.py and must keep the reference's memory behaviour.
 *
 * Build: see oracle/Makefile.  -ffp-contract=off is REQUIRED (numba/LLVM does not fuse
 * multiply-add without fast-math, so neither may we).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* Python's integer %: result has the sign of the divisor (density.py:19, integrate.py:20,65).
 * numba compiles N_CELLS in as a constant, so for a power-of-two mesh its % is a single AND;
 * the same shortcut here (two's-complement AND == Python % for a power-of-two divisor) keeps the
 * timing build comparable.  Values are identical either way. */
static inline int64_t pymod_i64(int64_t a, int64_t n)
{
    if ((n & (n - 1)) == 0) return a & (n - 1);
    int64_t r = a % n;
    return (r < 0) ? r + n : r;
}

/* Python/numba float %: fmod, then shift a negative remainder up by the (positive) divisor
 * (integrate.py:95). */
static inline double pymod_f64(double a, double n)
{
    double r = fmod(a, n);
    if (r != 0.0) {
        if (r < 0.0) r += n;
    } else {
        r = copysign(0.0, n);
    }
    return r;
}

int pmo_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/*
 * density(positions, mass)                                     reference: src/density.py:7-48
 *
 * positions: float32[3, Np] C-contiguous, row 0 = x (density.py:12-14)
 * grid     : float32[Nc, Nc, Nc] indexed [z, y, x]; zeroed here (density.py:11)
 * nthreads : 1 -> particle-index order, the deterministic result parity is judged against
 *            (SURVEY Q9).  >1 -> the prange of density.py:17; the reference's "+=" is a racy
 *            read-modify-write, here it is an atomic so the timing build has defined behaviour.
 */
void pmo_density(const float *positions, int64_t np, int nc, double mass, float *grid, int nthreads)
{
    const int64_t n = nc;
    const float *px = positions, *py = positions + np, *pz = positions + 2 * np;
    memset(grid, 0, sizeof(float) * (size_t)(n * n * n));

#define PMO_DEPOSIT_BODY(ADD)                                                         \
    /* density.py:19-21  int32(floor(x)) % N_CELLS */                                 \
    int64_t x_c = pymod_i64((int64_t)(int32_t)floorf(px[i]), n);                      \
    int64_t y_c = pymod_i64((int64_t)(int32_t)floorf(py[i]), n);                      \
    int64_t z_c = pymod_i64((int64_t)(int32_t)floorf(pz[i]), n);                      \
    /* density.py:24-30  float32 - int64 promotes to float64 */                       \
    double d_x = (double)px[i] - (double)x_c;                                         \
    double d_y = (double)py[i] - (double)y_c;                                         \
    double d_z = (double)pz[i] - (double)z_c;                                         \
    double t_x = 1 - d_x, t_y = 1 - d_y, t_z = 1 - d_z;                               \
    /* density.py:33-35 */                                                            \
    int64_t X = pymod_i64(x_c + 1, n), Y = pymod_i64(y_c + 1, n), Z = pymod_i64(z_c + 1, n); \
    /* density.py:37-47, products left to right in float64, "+=" rounds to float32 */ \
    ADD(grid[(z_c * n + y_c) * n + x_c], mass * t_x * t_y * t_z);                     \
    ADD(grid[(z_c * n + y_c) * n + X], mass * d_x * t_y * t_z);                       \
    ADD(grid[(z_c * n + Y) * n + x_c], mass * t_x * d_y * t_z);                       \
    ADD(grid[(Z * n + y_c) * n + x_c], mass * t_x * t_y * d_z);                       \
    ADD(grid[(z_c * n + Y) * n + X], mass * d_x * d_y * t_z);                         \
    ADD(grid[(Z * n + Y) * n + x_c], mass * t_x * d_y * d_z);                         \
    ADD(grid[(Z * n + y_c) * n + X], mass * d_x * t_y * d_z);                         \
    ADD(grid[(Z * n + Y) * n + X], mass * d_x * d_y * d_z);

    if (nthreads <= 1) {
#define ADD_SERIAL(cell, v) (cell) = (float)((double)(cell) + (v))
        for (int64_t i = 0; i < np; ++i) {
            PMO_DEPOSIT_BODY(ADD_SERIAL)
        }
    } else {
#ifdef _OPENMP
#pragma omp parallel for num_threads(nthreads) schedule(static)
#endif
        for (int64_t i = 0; i < np; ++i) {
#define ADD_ATOMIC(cell, v)                      \
    {                                            \
        float v32_ = (float)(v);                 \
        _Pragma("omp atomic")(cell) += v32_;     \
    }
            PMO_DEPOSIT_BODY(ADD_ATOMIC)
        }
    }
}

/*
 * potential_k(density_k, fgrid, a)                             reference: src/potential.py:12-15
 *   -3*OMEGA_M0/8/a * fgrid * density_k
 * rho_k: complex128[M] as interleaved (re, im), multiplied in place; fgrid: float32[M].
 * The scalar is folded left to right in float64, the float32 table is promoted to float64,
 * and a real*complex product scales both parts.
 */
void pmo_potential_k(double *rho_k, const float *fgrid, int64_t m, double omega_m0, double a,
                     int nthreads)
{
    const double c = -3 * omega_m0 / 8 / a;
#ifdef _OPENMP
#pragma omp parallel for num_threads(nthreads > 1 ? nthreads : 1) schedule(static)
#endif
    for (int64_t i = 0; i < m; ++i) {
        double s = c * (double)fgrid[i];
        rho_k[2 * i] = s * rho_k[2 * i];
        rho_k[2 * i + 1] = s * rho_k[2 * i + 1];
    }
}

/*
 * weights(cell_centers, positions)                             reference: src/integrate.py:27-52
 * t: float32[8, Np]; order [ttt, dtt, tdt, ttd, ddt, dtd, tdd, ddd] in (x, y, z)  
 */
static void pmo_weights(const int64_t *cc, const float *positions, int64_t np, float *t,
                        int nthreads)
{
#ifdef _OPENMP
#pragma omp parallel for num_threads(nthreads > 1 ? nthreads : 1) schedule(static)
#endif
    for (int64_t i = 0; i < np; ++i) {
        /* integrate.py:36-42, float32 - int64 -> float64 */
        double d_x = (double)positions[i] - (double)cc[i];
        double d_y = (double)positions[np + i] - (double)cc[np + i];
        double d_z = (double)positions[2 * np + i] - (double)cc[2 * np + i];
        double t_x = 1 - d_x, t_y = 1 - d_y, t_z = 1 - d_z;
        /* integrate.py:44-51, stored as float32 */
        t[0 * np + i] = (float)(t_x * t_y * t_z);
        t[1 * np + i] = (float)(d_x * t_y * t_z);
        t[2 * np + i] = (float)(t_x * d_y * t_z);
        t[3 * np + i] = (float)(t_x * t_y * d_z);
        t[4 * np + i] = (float)(d_x * d_y * t_z);
        t[5 * np + i] = (float)(d_x * t_y * d_z);
        t[6 * np + i] = (float)(t_x * d_y * d_z);
        t[7 * np + i] = (float)(d_x * d_y * d_z);
    }
}

/* numba wraps a negative index once (integrate.py:64 produces -1). */
static inline int64_t wrapneg(int64_t i, int64_t n) { return i < 0 ? i + n : i; }

/*
 * sweep_one_direction(...)                                     reference: src/integrate.py:54-97
 * acc (optional, float64[Np]) receives g_p of integrate.py:92 -- not a reference output, exposed
 * so the GPU kernel's accelerations can be compared (north_star tolerance list).
 */
static void pmo_sweep_one_direction(const int64_t *cc, float *pos_d, float *vel_d, const float *t,
                                    const float *phi, int64_t np, int nc, double da, double f_a1,
                                    double a_val, int direction, double *acc, int nthreads)
{
    const int64_t n = nc;
    /* integrate.py:61-65: two full copies of the cell table, one row rewritten in each */
    int64_t *cc_n = (int64_t *)malloc(sizeof(int64_t) * 3 * (size_t)np);
    int64_t *cc_p = (int64_t *)malloc(sizeof(int64_t) * 3 * (size_t)np);
#ifdef _OPENMP
#pragma omp parallel for num_threads(nthreads > 1 ? nthreads : 1) schedule(static)
#endif
    for (int64_t i = 0; i < 3 * np; ++i) {
        cc_n[i] = cc[i];
        cc_p[i] = cc[i];
    }
#ifdef _OPENMP
#pragma omp parallel for num_threads(nthreads > 1 ? nthreads : 1) schedule(static)
#endif
    for (int64_t i = 0; i < np; ++i) {
        cc_n[direction * np + i] = cc[direction * np + i] - 1;
        cc_p[direction * np + i] = pymod_i64(cc[direction * np + i] + 1, n);
    }

#define PHI(zz, yy, xx) phi[(wrapneg(zz, n) * n + wrapneg(yy, n)) * n + wrapneg(xx, n)]
#ifdef _OPENMP
#pragma omp parallel for num_threads(nthreads > 1 ? nthreads : 1) schedule(static)
#endif
    for (int64_t i = 0; i < np; ++i) {
        /* integrate.py:69-82 */
        int64_t x = cc_p[i], y = cc_p[np + i], z = cc_p[2 * np + i];
        int64_t x2 = cc_n[i], y2 = cc_n[np + i], z2 = cc_n[2 * np + i];
        int64_t X = pymod_i64(x + 1, n), Y = pymod_i64(y + 1, n), Z = pymod_i64(z + 1, n);
        int64_t X2 = pymod_i64(x2 + 1, n), Y2 = pymod_i64(y2 + 1, n), Z2 = pymod_i64(z2 + 1, n);
        /* integrate.py:84-91, float32 */
        float g = -PHI(z, y, x) + PHI(z2, y2, x2);
        float g_x = -PHI(z, y, X) + PHI(z2, y2, X2);
        float g_y = -PHI(z, Y, x) + PHI(z2, Y2, x2);
        float g_z = -PHI(Z, y, x) + PHI(Z2, y2, x2);
        float g_xy = -PHI(z, Y, X) + PHI(z2, Y2, X2);
        float g_xz = -PHI(Z, y, X) + PHI(Z2, y2, X2);
        float g_yz = -PHI(Z, Y, x) + PHI(Z2, Y2, x2);
        float g_xyz = -PHI(Z, Y, X) + PHI(Z2, Y2, X2);
        /* integrate.py:92: float32 products and sums left to right, "/2." promotes to float64 */
        float s = g * t[0 * np + i];
        s = s + g_x * t[1 * np + i];
        s = s + g_y * t[2 * np + i];
        s = s + g_z * t[3 * np + i];
        s = s + g_xy * t[4 * np + i];
        s = s + g_xz * t[5 * np + i];
        s = s + g_yz * t[6 * np + i];
        s = s + g_xyz * t[7 * np + i];
        double g_p = (double)s / 2.;
        if (acc) acc[i] = g_p;
        /* integrate.py:94: float32 += float64 evaluates in float64, stores float32 */
        vel_d[i] = (float)((double)vel_d[i] + da * f_a1 * g_p);
        /* integrate.py:95: uses the stored float32 velocity */
        double xnew = (double)pos_d[i] + da * (double)vel_d[i] / ((a_val + da) * (a_val + da)) * f_a1;
        pos_d[i] = (float)pymod_f64(xnew, (double)n);
    }
#undef PHI
    free(cc_n);
    free(cc_p);
}

/*
 * integrate(positions, velocities, a_val, f_a1, da, potentials)   reference: src/integrate.py:15-25
 * In place.  acc: optional float64[3, Np].
 */
void pmo_integrate(float *positions, float *velocities, int64_t np, int nc, double a_val,
                   double f_a1, double da, const float *phi, double *acc, int nthreads)
{
    /* integrate.py:20  floor(positions).astype(int64) % N_CELLS */
    int64_t *cc = (int64_t *)malloc(sizeof(int64_t) * 3 * (size_t)np);
    float *t = (float *)malloc(sizeof(float) * 8 * (size_t)np);
#ifdef _OPENMP
#pragma omp parallel for num_threads(nthreads > 1 ? nthreads : 1) schedule(static)
#endif
    for (int64_t i = 0; i < 3 * np; ++i) cc[i] = pymod_i64((int64_t)floorf(positions[i]), nc);
    pmo_weights(cc, positions, np, t, nthreads); /* integrate.py:21 */
    for (int direction = 0; direction < 3; ++direction) /* integrate.py:23-24 */
        pmo_sweep_one_direction(cc, positions + direction * np, velocities + direction * np, t, phi,
                                np, nc, da, f_a1, a_val, direction,
                                acc ? acc + direction * np : NULL, nthreads);
    free(t);
    free(cc);
}

/* Stable cell key of a particle, the quantity the CUDA deposit sorts by (not a reference
 * function; derived from density.py:19-21,37: key = (z_c*Nc + y_c)*Nc + x_c).  SURVEY Q12. */
void pmo_cell_keys(const float *positions, int64_t np, int nc, int64_t *keys)
{
    const int64_t n = nc;
    for (int64_t i = 0; i < np; ++i) {
        int64_t x_c = pymod_i64((int64_t)(int32_t)floorf(positions[i]), n);
        int64_t y_c = pymod_i64((int64_t)(int32_t)floorf(positions[np + i]), n);
        int64_t z_c = pymod_i64((int64_t)(int32_t)floorf(positions[2 * np + i]), n);
        keys[i] = (z_c * n + y_c) * n + x_c;
    }
}
