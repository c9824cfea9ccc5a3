// pm_ic.cu -- initial conditions on the device: the Gaussian random field of
// src/gaussian_random_field.py:9-123 and the Zel'dovich displacements of src/zeldovich.py:10-100.
//
// The reference builds its noise with a seeded-but-racy numba loop and its jitter with 3*Np calls of
// the unseeded stdlib random.uniform: neither is reproducible, and the Python loop alone
// is 3.2e9 interpreter iterations at 1024^3.  Here both come from a counter-based generator
// (Philox4x32-10 keyed by the seed, counter = element index), so every element can be produced
// independently and a run is reproducible bit for bit.  Everything else follows the reference's
// float64 / complex128 arithmetic (cuFFT Z2Z for its two pyFFTW inverse transforms and its
// np.fft.fftn), with the results stored as float32 exactly where the reference stores float32.
// Entries the reference leaves uninitialised at k = 0 (np.power / np.divide with where= and no out=)
// are 0.
#include <cufft.h>

#include <stdlib.h>
#include <string.h>

#include "pm_internal.cuh"

#define PM_ARGS(cond)                       \
    do {                                    \
        if (!(cond)) return PM_ERR_INVALID; \
    } while (0)
#define PM_TRY(expr)                  \
    do {                              \
        int rc_ = (expr);             \
        if (rc_ != PM_OK) return rc_; \
    } while (0)

namespace {

// ---- Philox4x32-10 (Salmon et al. 2011) -----------------------------------------------------------
struct Philox {
    uint32_t c[4];
};
__device__ __forceinline__ Philox philox(uint64_t ctr, uint32_t sub, uint32_t stream, uint64_t seed)
{
    uint32_t c0 = (uint32_t)ctr, c1 = (uint32_t)(ctr >> 32), c2 = sub, c3 = stream;
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    Philox p;
    p.c[0] = c0; p.c[1] = c1; p.c[2] = c2; p.c[3] = c3;
    return p;
}
// 32 random bits -> double in [0, 1) / 53 bits -> double in [0, 1)
__device__ __forceinline__ double u01_32(uint32_t a) { return (double)a * (1.0 / 4294967296.0); }
__device__ __forceinline__ double u01_53(uint32_t a, uint32_t b)
{
    const uint64_t m = ((uint64_t)(a >> 5) << 26) | (uint64_t)(b >> 6);   // 27 + 26 bits, as CPython's random()
    return (double)m * (1.0 / 9007199254740992.0);
}

// gaussian_random_field.py:31-63: pairs (u, v) uniform in (-1, 1), kept when 0 < u^2+v^2 < 1, stored
// as float32; polar Box-Muller in float32.  Element e draws attempts (e, 0), (e, 1), ... until one is
// accepted: two candidate pairs per Philox call.
// The arrays hold the elements [e0, e0 + n) of the field (e0 = 0: the whole field; a slab of planes is
// a contiguous element range, so every rank of a slab run draws exactly its part of the same field).
__global__ void __launch_bounds__(256) k_ic_noise(float *__restrict__ f1, float *__restrict__ f2, int64_t e0, int64_t n,
                                                  uint64_t seed)
{
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    float u = 0.f, v = 0.f, s = 0.f;
    for (uint32_t sub = 0;; ++sub) {
        const Philox p = philox((uint64_t)(e0 + e), sub, 0u, seed);
        u = (float)(2.0 * u01_32(p.c[0]) - 1.0);
        v = (float)(2.0 * u01_32(p.c[1]) - 1.0);
        s = u * u + v * v;
        if (s > 0.f && s < 1.f) break;
        u = (float)(2.0 * u01_32(p.c[2]) - 1.0);
        v = (float)(2.0 * u01_32(p.c[3]) - 1.0);
        s = u * u + v * v;
        if (s > 0.f && s < 1.f) break;
    }
    const float g = sqrtf(-2.0f * logf(s) / s);
    f1[e] = u * g;
    f2[e] = v * g;
}

// zeldovich.py:89-91: random.uniform(-2., 2.) = -2 + 4*random(), one draw per particle and direction
__global__ void __launch_bounds__(256) k_ic_jitter(double *__restrict__ jit, int64_t n, uint64_t seed)
{
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    const Philox p = philox((uint64_t)e, 0u, 1u, seed);
    const Philox q = philox((uint64_t)e, 1u, 1u, seed);
    jit[e] = -2.0 + 4.0 * u01_53(p.c[0], p.c[1]);
    jit[n + e] = -2.0 + 4.0 * u01_53(p.c[2], p.c[3]);
    jit[2 * n + e] = -2.0 + 4.0 * u01_53(q.c[0], q.c[1]);
}

struct IcGeom {
    int n;          // N_PARTS
    double scale;   // 2 pi N / BOX_SIZE  (gaussian_random_field.py:79, zeldovich.py:26)
};
// scale * fftfreq(n)[i]
__device__ __forceinline__ double ic_freq(int i, const IcGeom g)
{
    const int m = (i < (g.n + 1) / 2) ? i : i - g.n;
    return g.scale * ((double)m / (double)g.n);
}

struct IcPower {
    double power, gamma;   // POWER; Gamma of the transfer function (gaussian_random_field.py:68)
    int lcdm;
};
// k^POWER * T(k) (POWER >= 0 with the transfer function), k^POWER otherwise; 0 at k = 0
// (gaussian_random_field.py:65-74, 105-121)
__device__ __forceinline__ double ic_pk_term(double k, const IcPower w)
{
    if (k == 0.0) return 0.0;
    if (w.power < 0.0) return 1.0 / pow(k, -w.power);
    double t = pow(k, w.power);
    if (w.lcdm) {
        const double q = k / w.gamma;
        const double q2 = 16.1 * q, q3 = 5.46 * q, q4 = 6.71 * q;
        const double factor1 = sqrt(1.0 + 3.89 * q + q2 * q2 + q3 * q3 * q3 + (q4 * q4) * (q4 * q4));
        const double l = log(1.0 + 2.34 * q), d = 2.34 * q;
        t *= (l * l) / (d * d) / factor1;
    }
    return t;
}
__device__ __forceinline__ double ic_kabs(int64_t idx, const IcGeom g)
{
    const int n = g.n;
    const int i2 = (int)(idx % n), i1 = (int)((idx / n) % n), i0 = (int)(idx / ((int64_t)n * n));
    const double lz = ic_freq(i0, g), ly = ic_freq(i1, g), lx = ic_freq(i2, g);
    return sqrt(lx * lx + ly * ly + lz * lz);   // gaussian_random_field.py:87
}

// sum over the grid of the term above: fixed grid and fixed tree, so the normalisation is reproducible
constexpr int kSumBlocks = 1024;
__global__ void __launch_bounds__(256) k_ic_power_partial(int64_t n3, IcGeom g, IcPower w, double *__restrict__ part)
{
    __shared__ double s_w[8];
    double acc = 0.0;
    for (int64_t idx = (int64_t)blockIdx.x * 256 + threadIdx.x; idx < n3; idx += (int64_t)kSumBlocks * 256)
        acc += ic_pk_term(ic_kabs(idx, g), w);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, d);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int k = 0; k < 8; ++k) t += s_w[k];
        part[blockIdx.x] = t;
    }
}
__global__ void k_ic_power_total(double *__restrict__ part)
{
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        double t = 0.0;
        for (int k = 0; k < kSumBlocks; ++k) t += part[k];
        part[kSumBlocks] = t;
    }
}

// rho_k = sqrt(p D^2) f1 + i sqrt(p D^2) f2, p = A k^POWER T(k), A = 64 H0^2 Npix^2 / sum
// (gaussian_random_field.py:14-24, 100-112)
// The arrays hold the `cnt` grid elements that start at global element idx0 (0 and n^3: the whole grid).
__global__ void __launch_bounds__(256) k_ic_rhok(int64_t idx0, int64_t cnt, IcGeom g, IcPower w, double sigma2_npix2,
                                                 double growth, const double *__restrict__ total,
                                                 const float *__restrict__ f1, const float *__restrict__ f2,
                                                 double2 *__restrict__ rho_k)
{
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= cnt) return;
    const double A = sigma2_npix2 / *total;
    const double p = A * ic_pk_term(ic_kabs(idx0 + idx, g), w);
    const double amp = sqrt(p * (growth * growth));
    rho_k[idx] = make_double2(amp * (double)f1[idx], amp * (double)f2[idx]);
}

// the power spectrum grid itself (analysis / tests)
__global__ void __launch_bounds__(256) k_ic_power_grid(int64_t n3, IcGeom g, IcPower w, double sigma2_npix2,
                                                       const double *__restrict__ total, double *__restrict__ p)
{
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n3) return;
    p[idx] = (sigma2_npix2 / *total) * ic_pk_term(ic_kabs(idx, g), w);
}

// (ifftn(...).real).astype(float32): cuFFT's inverse is unnormalised
__global__ void __launch_bounds__(256) k_ic_real_f32(int64_t n3, const double2 *__restrict__ z, double inv_n3,
                                                     float *__restrict__ out)
{
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < n3) out[idx] = (float)(z[idx].x * inv_n3);
}

__global__ void __launch_bounds__(256) k_ic_f32_to_z(int64_t n3, const float *__restrict__ in, double2 *__restrict__ z)
{
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < n3) z[idx] = make_double2((double)in[idx], 0.0);
}

// zeldovich.py:24-38 and 56-69: phi_k = rho_k / -(k^2) (0 at k = 0);
// d_k = ((-i l_dir) * phi_k) * (N_CELLS / N_PARTS), l_dir along array axis `dir`
__global__ void __launch_bounds__(256) k_ic_dfk(int64_t idx0, int64_t cnt, IcGeom g, int dir, double resolution,
                                                const double2 *__restrict__ rho_k, double2 *__restrict__ out)
{
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= cnt) return;
    const int n = g.n;
    const int64_t gi = idx0 + idx;   // global grid element; the arrays are indexed by idx
    const int i2 = (int)(gi % n), i1 = (int)((gi / n) % n), i0 = (int)(gi / ((int64_t)n * n));
    const double lz = ic_freq(i0, g), ly = ic_freq(i1, g), lx = ic_freq(i2, g);
    const double del_sq = -(lx * lx + ly * ly + lz * lz);
    double2 phi = make_double2(0.0, 0.0);
    if (del_sq != 0.0) {
        const double2 r = rho_k[idx];
        phi = make_double2(r.x / del_sq, r.y / del_sq);
    }
    const double l = dir == 0 ? lz : (dir == 1 ? ly : lx);   // meshgrid(..., indexing='ij')[dir]: axis `dir`
    // (0 - i l)(x + i y) = l y - i l x
    out[idx] = make_double2((l * phi.y) * resolution, (-l * phi.x) * resolution);
}

// zeldovich.py:71-100: positions = (lattice + D*disp + jitter) % N_CELLS, velocities = vfac*disp,
// disp = ifftn(d_k).real * (N_CELLS/BOX_SIZE); both stored float32 (zeldovich.py:12-13, 21)
__global__ void __launch_bounds__(256) k_ic_particles(int64_t n3, int n, int dir, const double2 *__restrict__ z,
                                                      double inv_n3, double force_resolution, double step, double growth,
                                                      double vfac, double n_cells, const double *__restrict__ jitter,
                                                      float *__restrict__ pos, float *__restrict__ vel)
{
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n3) return;
    const int i2 = (int)(idx % n), i1 = (int)((idx / n) % n), i0 = (int)(idx / ((int64_t)n * n));
    const int id = dir == 0 ? i0 : (dir == 1 ? i1 : i2);
    const double disp = (z[idx].x * inv_n3) * force_resolution;
    double x = ((double)id * step + 0.5) + growth * disp;   // np.linspace(0, Nc - res, n)[id] + 0.5
    x += jitter[idx];
    double r = fmod(x, n_cells);                            // Python/NumPy %: sign of the divisor
    if (r != 0.0 && r < 0.0) r += n_cells;
    pos[idx] = (float)r;
    vel[idx] = (float)(vfac * disp);
}

// The same for ONE slab of the lattice along its fastest index: z holds the real-space displacement field
// of the lattice points (i0, i1, i2_lo + j), j < n2l, as [i0][i1][j].  Particle (i0, i1, i2) keeps the global
// id (i0 n + i1) n + i2 of the single-GPU generator and draws the same jitter (k_ic_jitter: element = id),
// so the union over the slabs IS the single-GPU particle set.  `jitter` (optional): float64[n n n2l] in the
// local order, for callers that bring their own.
__global__ void __launch_bounds__(256) k_ic_particles_slab(int64_t cnt, int n, int i2_lo, int n2l, int dir,
                                                           const double2 *__restrict__ z, double inv_n3,
                                                           double force_resolution, double step, double growth,
                                                           double vfac, double n_cells, uint64_t seed,
                                                           const double *__restrict__ jitter, float *__restrict__ pos,
                                                           float *__restrict__ vel, int32_t *__restrict__ ids)
{
    const int64_t loc = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (loc >= cnt) return;
    const int j = (int)(loc % n2l);
    const int64_t row = loc / n2l;   // i0 n + i1
    const int i1 = (int)(row % n), i0 = (int)(row / n), i2 = i2_lo + j;
    const int64_t gid = row * n + i2;
    const int id = dir == 0 ? i0 : (dir == 1 ? i1 : i2);
    double jit;
    if (jitter) {
        jit = jitter[loc];
    } else {
        const Philox p = philox((uint64_t)gid, dir == 2 ? 1u : 0u, 1u, seed);
        jit = -2.0 + 4.0 * (dir == 1 ? u01_53(p.c[2], p.c[3]) : u01_53(p.c[0], p.c[1]));
    }
    const double disp = (z[loc].x * inv_n3) * force_resolution;
    double x = ((double)id * step + 0.5) + growth * disp;
    x += jit;
    double r = fmod(x, n_cells);
    if (r != 0.0 && r < 0.0) r += n_cells;
    pos[loc] = (float)r;
    vel[loc] = (float)(vfac * disp);
    if (ids) ids[loc] = (int32_t)gid;
}

struct IcHost {
    IcGeom g;
    IcPower w;
    int64_t n3;
    double growth, sigma2_npix2;
};

double ic_Dt(double a, double om, double ol) { return 5.0 / 2.0 / om / (pow(om, 4.0 / 7.0) - ol + (1.0 + om / 2.0) * (1.0 + ol / 70.0)) * a; }

int ic_setup(const pm_ic_params *q, IcHost *h)
{
    if (!q || q->n_parts < 2 || q->n_cells < 2 || !(q->box_size > 0.0) || !(q->a_init > 0.0)) return PM_ERR_INVALID;
    const double two_pi = 6.283185307179586476925286766559;
    h->g.n = q->n_parts;
    h->g.scale = two_pi * q->n_parts / q->box_size;
    h->w.power = q->power;
    h->w.lcdm = (q->lcdm_transfer && q->power >= 0.0) ? 1 : 0;
    h->w.gamma = q->omega_m0 * q->h0 * exp(-q->omega_b0 - q->omega_b0 / q->omega_m0);
    h->n3 = (int64_t)q->n_parts * q->n_parts * q->n_parts;
    h->growth = ic_Dt(q->a_init, q->omega_m0, q->omega_lambda0);
    const double npix = (double)h->n3;
    h->sigma2_npix2 = 64.0 * q->h0 * q->h0 * npix * npix;
    return PM_OK;
}

// host scalars of zeldovich.py in the reference's order of operations
struct IcScalars {
    double resolution, force_resolution, step, vfac;
};
IcScalars ic_scalars(const pm_ic_params *q, const IcHost &h)
{
    IcScalars c;
    c.resolution = (double)q->n_cells / (double)q->n_parts;                              // zeldovich.py:58
    c.force_resolution = (double)q->n_cells / q->box_size;                               // zeldovich.py:46
    c.step = ((double)q->n_cells - c.resolution) / (double)(q->n_parts - 1);             // np.linspace step, zeldovich.py:79
    const double h0a = sqrt(q->h0 * q->h0 * (q->omega_m0 / (q->a_init * q->a_init * q->a_init) +
                                             q->omega_k0 / (q->a_init * q->a_init) + q->omega_lambda0));   // cosmology.py:18
    const double f0 = 1.0 / sqrt((q->omega_m0 + q->omega_k0 * q->a_init +
                                  q->omega_lambda0 * (q->a_init * q->a_init * q->a_init)) / q->a_init);   // cosmology.py:27
    c.vfac = q->a_init * f0 * h0a * h.growth;                                            // zeldovich.py:100
    return c;
}

inline unsigned ic_blocks(int64_t n) { return (unsigned)((n + 255) / 256); }

int ic_fft(double2 *z, int n, int direction, cudaStream_t st)
{
    cufftHandle plan;
    if (cufftPlan3d(&plan, n, n, n, CUFFT_Z2Z) != CUFFT_SUCCESS) return PM_ERR_CUFFT;
    int rc = PM_OK;
    if (cufftSetStream(plan, st) != CUFFT_SUCCESS ||
        cufftExecZ2Z(plan, reinterpret_cast<cufftDoubleComplex *>(z), reinterpret_cast<cufftDoubleComplex *>(z),
                     direction) != CUFFT_SUCCESS)
        rc = PM_ERR_CUFFT;
    cudaStreamSynchronize(st);   // the plan's work area must outlive the transform
    cufftDestroy(plan);
    return rc;
}

}  // namespace

extern "C" {

size_t pm_ic_workspace_bytes(int n_parts)
{
    if (n_parts < 2) return 0;
    const size_t n3 = (size_t)n_parts * n_parts * n_parts;
    return 2 * n3 * sizeof(double2) + (kSumBlocks + 8) * sizeof(double);
}

int pm_ic_noise(float *f1_d, float *f2_d, int64_t n, uint64_t seed, pm_stream_t stream)
{
    PM_ARGS(n >= 0 && (n == 0 || (f1_d && f2_d)));
    if (n == 0) return PM_OK;
    cudaStream_t st = pm_cu(stream);
    PM_LAUNCH(k_ic_noise, ic_blocks(n), 256, 0, st, f1_d, f2_d, (int64_t)0, n, seed);
    PM_CHECK_LAUNCH();
    return PM_OK;
}

int pm_ic_jitter(double *jitter_d, int64_t n, uint64_t seed, pm_stream_t stream)
{
    PM_ARGS(n >= 0 && (n == 0 || jitter_d));
    if (n == 0) return PM_OK;
    cudaStream_t st = pm_cu(stream);
    PM_LAUNCH(k_ic_jitter, ic_blocks(n), 256, 0, st, jitter_d, n, seed);
    PM_CHECK_LAUNCH();
    return PM_OK;
}

int pm_ic_power_spectrum(const pm_ic_params *q, double *p_d, void *work_d, size_t work_bytes, pm_stream_t stream)
{
    IcHost h;
    PM_TRY(ic_setup(q, &h));
    PM_ARGS(p_d && work_d && work_bytes >= pm_ic_workspace_bytes(q->n_parts));
    cudaStream_t st = pm_cu(stream);
    double *part = reinterpret_cast<double *>(static_cast<char *>(work_d) + 2 * (size_t)h.n3 * sizeof(double2));
    PM_LAUNCH(k_ic_power_partial, kSumBlocks, 256, 0, st, h.n3, h.g, h.w, part);
    PM_LAUNCH(k_ic_power_total, 1, 32, 0, st, part);
    PM_LAUNCH(k_ic_power_grid, ic_blocks(h.n3), 256, 0, st, h.n3, h.g, h.w, h.sigma2_npix2,
              (const double *)(part + kSumBlocks), p_d);
    PM_CHECK_LAUNCH();
    return PM_OK;
}

int pm_ic_gaussian_random_field(const pm_ic_params *q, const float *f1_d, const float *f2_d, float *density_d,
                                void *work_d, size_t work_bytes, pm_stream_t stream)
{
    IcHost h;
    PM_TRY(ic_setup(q, &h));
    PM_ARGS(f1_d && f2_d && density_d && work_d && work_bytes >= pm_ic_workspace_bytes(q->n_parts));
    cudaStream_t st = pm_cu(stream);
    double2 *z = static_cast<double2 *>(work_d);
    double *part = reinterpret_cast<double *>(static_cast<char *>(work_d) + 2 * (size_t)h.n3 * sizeof(double2));
    PM_LAUNCH(k_ic_power_partial, kSumBlocks, 256, 0, st, h.n3, h.g, h.w, part);
    PM_LAUNCH(k_ic_power_total, 1, 32, 0, st, part);
    PM_LAUNCH(k_ic_rhok, ic_blocks(h.n3), 256, 0, st, (int64_t)0, h.n3, h.g, h.w, h.sigma2_npix2, h.growth,
              (const double *)(part + kSumBlocks), f1_d, f2_d, z);
    PM_CHECK_LAUNCH();
    PM_TRY(ic_fft(z, q->n_parts, CUFFT_INVERSE, st));
    PM_LAUNCH(k_ic_real_f32, ic_blocks(h.n3), 256, 0, st, h.n3, (const double2 *)z, 1.0 / (double)h.n3, density_d);
    PM_CHECK_LAUNCH();
    return PM_OK;
}

int pm_ic_zeldovich(const pm_ic_params *q, const float *density_d, const double *jitter_d, float *pos_d, float *vel_d,
                    void *work_d, size_t work_bytes, pm_stream_t stream)
{
    IcHost h;
    PM_TRY(ic_setup(q, &h));
    PM_ARGS(density_d && jitter_d && pos_d && vel_d && work_d && work_bytes >= pm_ic_workspace_bytes(q->n_parts));
    cudaStream_t st = pm_cu(stream);
    double2 *rho_k = static_cast<double2 *>(work_d);
    double2 *z = rho_k + h.n3;
    const int n = q->n_parts;
    PM_LAUNCH(k_ic_f32_to_z, ic_blocks(h.n3), 256, 0, st, h.n3, density_d, rho_k);
    PM_CHECK_LAUNCH();
    PM_TRY(ic_fft(rho_k, n, CUFFT_FORWARD, st));   // np.fft.fftn(density), zeldovich.py:17
    const IcScalars c = ic_scalars(q, h);
    const double resolution = c.resolution, force_resolution = c.force_resolution, step = c.step, vfac = c.vfac;
    for (int dir = 0; dir < 3; ++dir) {
        PM_LAUNCH(k_ic_dfk, ic_blocks(h.n3), 256, 0, st, (int64_t)0, h.n3, h.g, dir, resolution, (const double2 *)rho_k, z);
        PM_CHECK_LAUNCH();
        PM_TRY(ic_fft(z, n, CUFFT_INVERSE, st));
        PM_LAUNCH(k_ic_particles, ic_blocks(h.n3), 256, 0, st, h.n3, n, dir, (const double2 *)z, 1.0 / (double)h.n3,
                  force_resolution, step, h.growth, vfac, (double)q->n_cells, jitter_d + (size_t)dir * h.n3,
                  pos_d + (size_t)dir * h.n3, vel_d + (size_t)dir * h.n3);
        PM_CHECK_LAUNCH();
    }
    return PM_OK;
}

/* ---- per-slab initial conditions (include/pmstep.h, "Initial conditions, one slab per rank") ---- */

size_t pm_ic_slab_workspace_bytes(void) { return (kSumBlocks + 8) * sizeof(double); }

int pm_ic_noise_range(float *f1_d, float *f2_d, int64_t e0, int64_t n, uint64_t seed, pm_stream_t stream)
{
    PM_ARGS(e0 >= 0 && n >= 0 && (n == 0 || (f1_d && f2_d)));
    if (n == 0) return PM_OK;
    cudaStream_t st = pm_cu(stream);
    PM_LAUNCH(k_ic_noise, ic_blocks(n), 256, 0, st, f1_d, f2_d, e0, n, seed);
    PM_CHECK_LAUNCH();
    return PM_OK;
}

int pm_ic_slab_rho_k(const pm_ic_params *q, const float *f1_d, const float *f2_d, int i0_lo, int n0l, void *zk_d,
                     void *work_d, size_t work_bytes, pm_stream_t stream)
{
    IcHost h;
    PM_TRY(ic_setup(q, &h));
    PM_ARGS(f1_d && f2_d && zk_d && work_d && work_bytes >= pm_ic_slab_workspace_bytes());
    PM_ARGS(i0_lo >= 0 && n0l >= 1 && i0_lo + n0l <= q->n_parts);
    cudaStream_t st = pm_cu(stream);
    double *part = static_cast<double *>(work_d);
    const int64_t plane = (int64_t)q->n_parts * q->n_parts, cnt = plane * n0l;
    // the normalisation is the sum over the WHOLE grid, in the single-GPU generator's fixed tree: every
    // rank computes it (it is analytic in k) and gets the same bits
    PM_LAUNCH(k_ic_power_partial, kSumBlocks, 256, 0, st, h.n3, h.g, h.w, part);
    PM_LAUNCH(k_ic_power_total, 1, 32, 0, st, part);
    PM_LAUNCH(k_ic_rhok, ic_blocks(cnt), 256, 0, st, plane * i0_lo, cnt, h.g, h.w, h.sigma2_npix2, h.growth,
              (const double *)(part + kSumBlocks), f1_d, f2_d, static_cast<double2 *>(zk_d));
    PM_CHECK_LAUNCH();
    return PM_OK;
}

int pm_ic_slab_fft(void *z_d, int d0, int d1, int d2, int axes, int inverse, pm_stream_t stream)
{
    PM_ARGS(z_d && d0 >= 1 && d1 >= 1 && d2 >= 1 && axes >= PM_IC_AXES_12 && axes <= PM_IC_AXIS_2);
    PM_ARGS((int64_t)d0 * d1 * d2 < ((int64_t)1 << 31));
    cudaStream_t st = pm_cu(stream);
    cufftHandle plan;
    cufftResult cr;
    if (axes == PM_IC_AXES_12) {          // 2-D over the two fastest axes, one transform per d0
        int n[2] = {d1, d2};
        cr = cufftPlanMany(&plan, 2, n, nullptr, 1, d1 * d2, nullptr, 1, d1 * d2, CUFFT_Z2Z, d0);
    } else if (axes == PM_IC_AXIS_0) {    // 1-D along the slowest axis, stride d1 d2, one transform per (i1, i2)
        int n[1] = {d0}, emb[1] = {d0};
        cr = cufftPlanMany(&plan, 1, n, emb, d1 * d2, 1, emb, d1 * d2, 1, CUFFT_Z2Z, d1 * d2);
    } else if (axes == PM_IC_AXES_01) {   // 2-D over the two slowest axes, stride d2, one transform per i2
        int n[2] = {d0, d1}, emb[2] = {d0, d1};
        cr = cufftPlanMany(&plan, 2, n, emb, d2, 1, emb, d2, 1, CUFFT_Z2Z, d2);
    } else {                              // 1-D along the fastest axis
        int n[1] = {d2};
        cr = cufftPlanMany(&plan, 1, n, nullptr, 1, d2, nullptr, 1, d2, CUFFT_Z2Z, d0 * d1);
    }
    if (cr != CUFFT_SUCCESS) return PM_ERR_CUFFT;
    int rc = PM_OK;
    cufftDoubleComplex *z = static_cast<cufftDoubleComplex *>(z_d);
    if (cufftSetStream(plan, st) != CUFFT_SUCCESS ||
        cufftExecZ2Z(plan, z, z, inverse ? CUFFT_INVERSE : CUFFT_FORWARD) != CUFFT_SUCCESS)
        rc = PM_ERR_CUFFT;
    cudaStreamSynchronize(st);   // the plan's work area must outlive the transform
    cufftDestroy(plan);
    return rc;
}

int pm_ic_slab_real_f32(const void *z_d, int64_t count, double scale, float *out_d, pm_stream_t stream)
{
    PM_ARGS(count >= 0 && (count == 0 || (z_d && out_d)));
    if (count == 0) return PM_OK;
    PM_LAUNCH(k_ic_real_f32, ic_blocks(count), 256, 0, pm_cu(stream), count, static_cast<const double2 *>(z_d), scale,
              out_d);
    PM_CHECK_LAUNCH();
    return PM_OK;
}

int pm_ic_slab_from_f32(const float *in_d, int64_t count, void *z_d, pm_stream_t stream)
{
    PM_ARGS(count >= 0 && (count == 0 || (z_d && in_d)));
    if (count == 0) return PM_OK;
    PM_LAUNCH(k_ic_f32_to_z, ic_blocks(count), 256, 0, pm_cu(stream), count, in_d, static_cast<double2 *>(z_d));
    PM_CHECK_LAUNCH();
    return PM_OK;
}

int pm_ic_slab_displacement_k(const pm_ic_params *q, int dir, const void *rho_k_d, int i0_lo, int n0l, void *out_d,
                              pm_stream_t stream)
{
    IcHost h;
    PM_TRY(ic_setup(q, &h));
    PM_ARGS(rho_k_d && out_d && dir >= 0 && dir < 3 && i0_lo >= 0 && n0l >= 1 && i0_lo + n0l <= q->n_parts);
    const int64_t plane = (int64_t)q->n_parts * q->n_parts, cnt = plane * n0l;
    const double resolution = (double)q->n_cells / (double)q->n_parts;   // zeldovich.py:58
    PM_LAUNCH(k_ic_dfk, ic_blocks(cnt), 256, 0, pm_cu(stream), plane * i0_lo, cnt, h.g, dir, resolution,
              static_cast<const double2 *>(rho_k_d), static_cast<double2 *>(out_d));
    PM_CHECK_LAUNCH();
    return PM_OK;
}

int pm_ic_slab_particles(const pm_ic_params *q, int dir, const void *z_d, int i2_lo, int n2l, uint64_t seed,
                         const double *jitter_d, float *pos_d, float *vel_d, int32_t *ids_d, pm_stream_t stream)
{
    IcHost h;
    PM_TRY(ic_setup(q, &h));
    PM_ARGS(z_d && pos_d && vel_d && dir >= 0 && dir < 3 && i2_lo >= 0 && n2l >= 1 && i2_lo + n2l <= q->n_parts);
    PM_ARGS(h.n3 < ((int64_t)1 << 31));   // int32 particle ids, as everywhere in the slab path
    const int n = q->n_parts;
    const int64_t cnt = (int64_t)n * n * n2l;
    const IcScalars c = ic_scalars(q, h);
    PM_LAUNCH(k_ic_particles_slab, ic_blocks(cnt), 256, 0, pm_cu(stream), cnt, n, i2_lo, n2l, dir,
              static_cast<const double2 *>(z_d), 1.0 / (double)h.n3, c.force_resolution, c.step, h.growth, c.vfac,
              (double)q->n_cells, seed, jitter_d, pos_d, vel_d, ids_d);
    PM_CHECK_LAUNCH();
    return PM_OK;
}

}  // extern "C"
