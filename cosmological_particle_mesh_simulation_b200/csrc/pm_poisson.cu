// pm_poisson.cu -- Poisson solve of the PM step: cuFFT R2C -> fused Green's kernel -> cuFFT C2R.
//
// Reference: src/potential.py:7-29 (complex128 c2c transforms of the real density, a separate
// elementwise multiply by the float32[Nc^3] table of src/fourier_utils.py:5-16, real part kept).
// Here the density is real, so only the Nc*Nc*(Nc/2+1) half spectrum exists, the table is never
// materialised (three float32 sin^2 values per mode from an Nc-entry table), and the constant
// -3*Omega_m/(8a), the 1/Nc^3 of the inverse transform (pyFFTW normalises, cuFFT does not) and
// the zeroed DC mode are folded into the same pass over the spectrum.
#include <math.h>

#include "pm_internal.cuh"

// sin^2(k_i/2), k_i = 2*pi*fftfreq(Nc)[i]   (src/fourier_utils.py:8-15).  sin^2 is even and
// pi-periodic, so sin^2(pi*i/Nc) covers the negative frequencies too.  Evaluated in float64 on
// the host and rounded once to float32 (the reference rounds k to float32 first; the difference
// is below 2e-7 relative per mode).
int pm_k_sin2_table(pm_plan *p)
{
    const int nc = p->nc;
    float *h = (float *)malloc(sizeof(float) * nc);
    if (!h) return PM_ERR_NOMEM;
    for (int i = 0; i < nc; ++i) {
        double s = sin(M_PI * (double)i / (double)nc);
        h[i] = (float)(s * s);
    }
    cudaError_t e = cudaMemcpy(p->sin2, h, sizeof(float) * nc, cudaMemcpyHostToDevice);
    free(h);
    return (int)e;
}

// G(z,y,x) = 1 / ((s[z] + s[y]) + s[x]) in float32, 0 at DC -- the table fourier_grid() returns.
__device__ __forceinline__ float pm_green(float sz, float sy, float sx)
{
    const float ksq = __fadd_rn(__fadd_rn(sz, sy), sx);
    return ksq != 0.0f ? __fdiv_rn(1.0f, ksq) : 0.0f;
}

__global__ void __launch_bounds__(256) k_fourier_grid(const float *__restrict__ sin2, int nc,
                                                      float *__restrict__ fgrid)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y, z = blockIdx.z;
    if (x >= nc) return;
    fgrid[((size_t)z * nc + y) * nc + x] = pm_green(sin2[z], sin2[y], sin2[x]);
}

int pm_k_fourier_grid(pm_plan *p, float *fgrid, cudaStream_t st)
{
    dim3 grid((p->nc + 255) / 256, p->nc, p->nc);
    PM_LAUNCH(k_fourier_grid, grid, 256, 0, st, p->sin2, p->nc, fgrid);
    PM_CHECK_LAUNCH();
    return PM_OK;
}

// phi_k = (-3*Omega_m/8/a / Nc^3) * G(k) * rho_k on the half spectrum [z][y][x <= Nc/2], in place.
// One CTA per (z, y) row; the x-row of sin^2 sits in shared memory; float2 (8-byte) accesses.
__global__ void __launch_bounds__(256) k_green_multiply(float2 *__restrict__ spec,
                                                        const float *__restrict__ sin2, int nc,
                                                        int nxh, double scale)
{
    extern __shared__ float s_sin2[];
    for (int i = threadIdx.x; i < nxh; i += blockDim.x) s_sin2[i] = sin2[i];
    __syncthreads();
    const int rows = nc * nc;
    for (int r = blockIdx.x; r < rows; r += gridDim.x) {
        const int z = r / nc, y = r - z * nc;
        const float sz = sin2[z], sy = sin2[y];
        float2 *row = spec + (size_t)r * nxh;
        for (int x = threadIdx.x; x < nxh; x += blockDim.x) {
            const double g = scale * (double)pm_green(sz, sy, s_sin2[x]);
            float2 c = row[x];
            c.x = (float)(g * (double)c.x);
            c.y = (float)(g * (double)c.y);
            row[x] = c;
        }
    }
}

// ---- options: CIC-window deconvolution and spectral gradient (north_star (2)) -----------
// The reference has neither: its potential is G(k) rho_k (src/potential.py:12-15) and its accelerations are
// central differences of phi at the eight corners (src/integrate.py:84-91).  Both are options here, off by
// default; with them on the solve runs the library transforms around ONE pass over the half spectrum that
// applies Green's function, the constant and 1/W(k)^p together, and -- for the spectral gradient -- three
// more passes that write 2 * (-i k_d) phi_k for one axis each into a second spectrum, transformed into the
// plan's three force meshes.  W(k) = prod_i [sin(k_i/2) / (k_i/2)]^2 is the CIC assignment window: p = 1
// undoes the deposit's smoothing, p = 2 the interpolation's as well.
int pm_k_poisson_tables(pm_plan *p)
{
    const int nc = p->nc;
    float *h = (float *)malloc(sizeof(float) * 2 * nc);
    if (!h) return PM_ERR_NOMEM;
    for (int i = 0; i < nc; ++i) {
        const int fi = i <= (nc - 1) / 2 ? i : i - nc;                 // fftfreq numerator
        const double k = 2.0 * M_PI * (double)fi / (double)nc;
        const bool nyq = (nc % 2 == 0 && i == nc / 2);
        double w = 1.0;
        if (fi != 0) {
            const double s = sin(0.5 * k) / (0.5 * k);
            w = s * s;
        }
        h[i] = (float)pow(w, -(double)p->deconv);
        h[nc + i] = nyq ? 0.0f : (float)k;                            // the derivative of the Nyquist mode is dropped
    }
    cudaError_t e = cudaMemcpy(p->dec_tab, h, sizeof(float) * nc, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(p->k_tab, h + nc, sizeof(float) * nc, cudaMemcpyHostToDevice);
    free(h);
    return (int)e;
}

__global__ void __launch_bounds__(256) k_green_multiply_dec(float2 *__restrict__ spec, const float *__restrict__ sin2,
                                                            const float *__restrict__ dec, int nc, int nxh, double scale)
{
    const int rows = nc * nc;
    for (int r = blockIdx.x; r < rows; r += gridDim.x) {
        const int z = r / nc, y = r - z * nc;
        const float sz = sin2[z], sy = sin2[y];
        const double dzy = (double)dec[z] * (double)dec[y];
        float2 *row = spec + (size_t)r * nxh;
        for (int x = threadIdx.x; x < nxh; x += blockDim.x) {
            const double g = scale * (double)pm_green(sz, sy, sin2[x]) * (dzy * (double)dec[x]);
            float2 c = row[x];
            c.x = (float)(g * (double)c.x);
            c.y = (float)(g * (double)c.y);
            row[x] = c;
        }
    }
}

// out = 2 * (-i k_axis) * in on the half spectrum [z][y][x <= nc/2]
__global__ void __launch_bounds__(256) k_grad_spec(const float2 *__restrict__ in, float2 *__restrict__ out,
                                                   const float *__restrict__ ktab, int nc, int nxh, int axis)
{
    const int rows = nc * nc;
    for (int r = blockIdx.x; r < rows; r += gridDim.x) {
        const int z = r / nc, y = r - z * nc;
        const float2 *src = in + (size_t)r * nxh;
        float2 *dst = out + (size_t)r * nxh;
        const float kzy = axis == 2 ? ktab[z] : ktab[y];
        for (int x = threadIdx.x; x < nxh; x += blockDim.x) {
            const float k2 = 2.0f * (axis == 0 ? ktab[x] : kzy);
            const float2 c = src[x];
            dst[x] = make_float2(k2 * c.y, -k2 * c.x);                  // (a + ib)(-i k2) = k2 b - i k2 a
        }
    }
}

// ---- <rho>: the constant the forward transform subtracts (pm_internal.cuh, rho_mean_d) -----------
__global__ void k_set_mean(float *dst, float v) { *dst = v; }

// Deterministic mean of an arbitrary mesh: 1024 fixed blocks write float64 partial sums, one block
// adds them in index order (no atomics: the same input gives the same bits).
__global__ void __launch_bounds__(256) k_mean_partial(const float *__restrict__ rho, size_t n,
                                                      double *__restrict__ part)
{
    __shared__ double s_w[8];
    const size_t per = (n + gridDim.x - 1) / gridDim.x;
    const size_t lo = per * blockIdx.x, hi = lo + per < n ? lo + per : n;
    double s = 0.0;
    for (size_t i = lo + threadIdx.x; i < hi; i += 256) s += (double)rho[i];
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += s_w[w];
        part[blockIdx.x] = t;
    }
}

__global__ void __launch_bounds__(32) k_mean_final(const double *__restrict__ part, int nparts,
                                                   double total_cells, float *__restrict__ mean)
{
    if (threadIdx.x != 0) return;
    double t = 0.0;
    for (int i = 0; i < nparts; ++i) t += part[i];
    *mean = (float)(t / total_cells);
}

// Sets p->rho_mean_d from p->rho_mean_hint, or from the mesh itself when the hint is NaN.  `n` cells
// of this rank, `total_cells` the divisor (a slab rank with an unknown mean would need an all-reduce:
// slab callers always give the hint).
int pm_k_rho_mean(pm_plan *p, const float *rho, size_t n, double total_cells, cudaStream_t st)
{
    if (!isnan(p->rho_mean_hint)) {
        PM_LAUNCH(k_set_mean, 1, 1, 0, st, p->rho_mean_d, (float)p->rho_mean_hint);
    } else {
        PM_LAUNCH(k_mean_partial, 1024, 256, 0, st, rho, n, p->mean_ws);
        PM_LAUNCH(k_mean_final, 1, 32, 0, st, (const double *)p->mean_ws, 1024, total_cells, p->rho_mean_d);
    }
    PM_CHECK_LAUNCH();
    return PM_OK;
}

__global__ void __launch_bounds__(256) k_sub_mean(const float4 *__restrict__ in, float4 *__restrict__ out,
                                                  size_t n4, const float *__restrict__ mean_ptr)
{
    const float m = __ldg(mean_ptr);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        float4 v = in[i];
        v.x -= m; v.y -= m; v.z -= m; v.w -= m;
        out[i] = v;
    }
}

// ---- diagnostic backend 2: the reference's transform precision ------------------------------------
// src/potential.py:12-29 transforms the float32 density in complex128, multiplies by
// (-3*Omega_m/8/a) * fgrid (float64 scalar x float32 table) and keeps the real part as float32.  The same
// here with cuFFT D2Z/Z2D: everything between the float32 density and the float32 potential is float64,
// the Green's table is the float32 one of fourier_grid().  Not a production path (no fusion, 3x the
// memory): it exists so that a test can change ONLY the transform precision and show which part of a
// parity residue is float32 transform noise (tests/test_gpu_parity.py, the Q4 "spike" fixtures).
__global__ void __launch_bounds__(256) k_f2d(const float *__restrict__ in, double *__restrict__ out, size_t n)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        out[i] = (double)in[i];
}

__global__ void __launch_bounds__(256) k_d2f(const double *__restrict__ in, float *__restrict__ out, size_t n)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        out[i] = (float)in[i];
}

__global__ void __launch_bounds__(256) k_green_multiply_f64(double2 *__restrict__ spec, const float *__restrict__ sin2,
                                                            int nc, int nxh, double c_pot, double inv_m)
{
    const int rows = nc * nc;
    for (int r = blockIdx.x; r < rows; r += gridDim.x) {
        const int z = r / nc, y = r - z * nc;
        const float sz = sin2[z], sy = sin2[y];
        double2 *row = spec + (size_t)r * nxh;
        for (int x = threadIdx.x; x < nxh; x += blockDim.x) {
            // (-3*Omega_m/8/a * fgrid) * density_k, then the inverse transform's 1/Nc^3 (pyFFTW normalises)
            const double g = __dmul_rn(c_pot, (double)pm_green(sz, sy, sin2[x]));
            double2 c = row[x];
            c.x = __dmul_rn(__dmul_rn(g, c.x), inv_m);
            c.y = __dmul_rn(__dmul_rn(g, c.y), inv_m);
            row[x] = c;
        }
    }
}

static int pm_k_poisson_f64(pm_plan *p, const float *rho, double a, double omega_m0, float *phi, cudaStream_t st)
{
    const int nc = p->nc, nxh = nc / 2 + 1;
    const size_t cells = (size_t)nc * nc * nc;
    if (!p->f64_ready) {
        PM_CUDA(cudaMalloc(&p->f64_mesh, cells * sizeof(double)));
        PM_CUDA(cudaMalloc(&p->f64_spec, (size_t)nc * nc * nxh * 2 * sizeof(double)));
        PM_CUFFT(cufftPlan3d(&p->d2z, nc, nc, nc, CUFFT_D2Z));
        PM_CUFFT(cufftPlan3d(&p->z2d, nc, nc, nc, CUFFT_Z2D));
        p->f64_ready = true;
    }
    PM_CUFFT(cufftSetStream(p->d2z, st));
    PM_CUFFT(cufftSetStream(p->z2d, st));
    PM_LAUNCH(k_f2d, p->sm_count * 8, 256, 0, st, rho, p->f64_mesh, cells);
    PM_CUFFT(cufftExecD2Z(p->d2z, p->f64_mesh, reinterpret_cast<cufftDoubleComplex *>(p->f64_spec)));
    pm_prof_mark(p, PM_STAGE_R2C + 1, st);
    const int rows = nc * nc;
    PM_LAUNCH(k_green_multiply_f64, rows < p->sm_count * 8 ? rows : p->sm_count * 8, 256, 0, st,
              reinterpret_cast<double2 *>(p->f64_spec), p->sin2, nc, nxh, -3 * omega_m0 / 8 / a,
              1.0 / ((double)nc * nc * nc));
    pm_prof_mark(p, PM_STAGE_GREEN + 1, st);
    PM_CUFFT(cufftExecZ2D(p->z2d, reinterpret_cast<cufftDoubleComplex *>(p->f64_spec), p->f64_mesh));
    PM_LAUNCH(k_d2f, p->sm_count * 8, 256, 0, st, (const double *)p->f64_mesh, phi, cells);
    PM_CHECK_LAUNCH();
    pm_prof_mark(p, PM_STAGE_C2R + 1, st);
    return PM_OK;
}

int pm_k_poisson(pm_plan *p, const float *rho, double a, double omega_m0, float *phi,
                 cudaStream_t st)
{
    const size_t cells = (size_t)p->nc * p->nc * p->nc;
    if (p->fft_f64 && !(p->deconv || p->kgrad)) return pm_k_poisson_f64(p, rho, a, omega_m0, phi, st);
    {
        const int rc = pm_k_rho_mean(p, rho, cells, (double)cells, st);
        if (rc != PM_OK) return rc;
    }
    if (p->own_fft && !(p->deconv || p->kgrad)) return pm_k_poisson_own(p, rho, a, omega_m0, phi, st);
    if (!p->have_fft) return PM_ERR_UNSUPPORTED;      // (slab plans: options are refused at set time)
    const int nc = p->nc, nxh = nc / 2 + 1;
    PM_CUFFT(cufftSetStream(p->r2c, st));
    PM_CUFFT(cufftSetStream(p->c2r, st));
    // library path (non-power-of-two meshes, A/B checks): rho - <rho> staged in the output mesh
    const float *src = rho;
    if (cells % 4 == 0 && (uintptr_t)rho % 16 == 0 && (uintptr_t)phi % 16 == 0) {
        PM_LAUNCH(k_sub_mean, p->sm_count * 8, 256, 0, st, reinterpret_cast<const float4 *>(rho),
                  reinterpret_cast<float4 *>(phi), cells / 4, (const float *)p->rho_mean_d);
        src = phi;
    }
    PM_CUFFT(cufftExecR2C(p->r2c, const_cast<float *>(src), reinterpret_cast<cufftComplex *>(p->spec)));
    pm_prof_mark(p, PM_STAGE_R2C + 1, st);
    const double m = (double)nc * (double)nc * (double)nc;
    const double scale = -3 * omega_m0 / 8 / a / m;  // potential.py:15, plus the IFFT's 1/Nc^3
    int rows = nc * nc;
    int grid = rows < p->sm_count * 8 ? rows : p->sm_count * 8;
    if (p->deconv || p->kgrad) {
        PM_LAUNCH(k_green_multiply_dec, grid, 256, 0, st, p->spec, p->sin2, p->dec_tab, nc, nxh, scale);
        if (p->kgrad) {
            // before the potential's own inverse transform: a multi-dimensional C2R may overwrite its input
            for (int d = 0; d < 3; ++d) {
                PM_LAUNCH(k_grad_spec, grid, 256, 0, st, (const float2 *)p->spec, p->spec2, (const float *)p->k_tab, nc, nxh, d);
                PM_CUFFT(cufftExecC2R(p->c2r, reinterpret_cast<cufftComplex *>(p->spec2), p->fmesh + (size_t)d * cells));
            }
        }
    } else {
        PM_LAUNCH(k_green_multiply, grid, 256, nxh * sizeof(float), st, p->spec, p->sin2, nc, nxh, scale);
    }
    PM_CHECK_LAUNCH();
    pm_prof_mark(p, PM_STAGE_GREEN + 1, st);
    PM_CUFFT(cufftExecC2R(p->c2r, reinterpret_cast<cufftComplex *>(p->spec), phi));
    pm_prof_mark(p, PM_STAGE_C2R + 1, st);
    return PM_OK;
}
