// pm_fft.cu -- hand-written 3-D real FFT Poisson solve for power-of-two meshes, sm_100a.
//
// Replaces, for N_CELLS in {32 ... 2048}, the cuFFT R2C -> Green -> cuFFT C2R sequence of
// pm_poisson.cu (which stays as the general-size path).  Reference semantics are unchanged:
// phi = IFFT(-3*Omega_m/(8a) * G(k) * FFT(rho)) (src/potential.py:7-29, src/fourier_utils.py:5-16).
//
// Five passes, each reading and writing the 4*Nc^3-byte array exactly once (40*Nc^3 bytes per
// solve; cuFFT + a separate Green pass moves 56*Nc^3):
//   1. rows  R2C : N reals -> N/2 complex per x-row ("packed": slot 0 holds (DC, Nyquist))
//   2. cols  y   : forward complex FFT along y, 16 adjacent kx columns per CTA
//   3. cols  z   : forward FFT along z, multiply by the Green's function, inverse FFT along z,
//                  all inside one CTA -- the spectrum never returns to HBM in between
//   4. cols  y   : inverse
//   5. rows  C2R : N/2 complex -> N reals
// The Nyquist-in-x plane (kx = N/2) lives in a small side array [z][y]; it is split off the
// packed slot in pass 2 and merged back in pass 4.
//
// Every 1-D transform is an in-place decimation-in-frequency FFT (radix 8/8/8 for 512) whose
// output is left in digit-reversed order; the inverse is the matching decimation-in-time FFT
// that consumes digit-reversed input.  Nothing is ever un-permuted: the Green's kernel looks
// sin^2 up in a table stored in the same digit-reversed order (p->sin2rev).  In-place
// butterflies own disjoint element sets, so one __syncthreads per radix stage suffices and a
// single shared-memory tile [N][17] float2 serves the whole pass; the first and last stage of
// a pass exchange data with global memory directly from registers.
#include <math.h>

#include <type_traits>

#include "pm_internal.cuh"

namespace {

#ifndef PM_FFT_COLS
#define PM_FFT_COLS 16
#endif
#ifndef PM_FFT_CTHREADS
#define PM_FFT_CTHREADS 256
#endif
// adjacent columns (contiguous in memory) per column tile: 16 (128-byte segments) while three
// tiles fit an SM, 8 for the 1024-point transforms (a 16-wide tile would be 147 KB: one CTA/SM)
template <int N>
constexpr int kColsCN = (N >= 1024) ? 8 : PM_FFT_COLS;
template <int N>
constexpr int kThrC = (N >= 256) ? PM_FFT_CTHREADS : 256;   // threads of a column-pass CTA
constexpr int kCols = 16;          // x-rows per CTA of the row passes
constexpr int kPitch = kCols + 1;  // +1: the split-off Nyquist column of the packed slot; also
                                   // keeps transposing row accesses conflict-free
#ifndef PM_FFT_THREADS
#define PM_FFT_THREADS 256
#endif
// threads per CTA: PM_FFT_THREADS for the large transforms, 256 for the small ones (whose tiles
// hold fewer butterflies than that many threads)
template <int N>
constexpr int kThr = (N >= 2048) ? 512 : ((N >= 256) ? PM_FFT_THREADS : 256);
// Column kernels: the CTAs/SM the register budget targets (3 tiles of 72 KB fit an SM at Nc = 512).
#ifndef PM_FFT_MINB
#define PM_FFT_MINB 3
#endif

// ---- compile-time radix plan: as many 8s as divide n, then one 4 or 2 -------------------------
__host__ __device__ constexpr int fft_radix(int n, int stage)
{
    for (int t = 0;; ++t) {
        const int r = (n % 8 == 0) ? 8 : n;
        if (t == stage) return r;
        n /= r;
    }
}
__host__ __device__ constexpr int fft_stages(int n)
{
    int s = 0;
    while (n > 1) {
        n /= (n % 8 == 0) ? 8 : n;
        ++s;
    }
    return s;
}
__host__ __device__ constexpr int fft_len(int n, int stage)  // sub-transform length entering `stage`
{
    for (int t = 0; t < stage; ++t) n /= fft_radix(n, 0);
    return n;
}
__host__ __device__ constexpr int ilog2(int n) { return n <= 1 ? 0 : 1 + ilog2(n / 2); }

// position of natural index k after the DIF passes (mixed-radix digit reversal)
template <int N>
__host__ __device__ inline int digit_rev(int k)
{
    int pos = 0, len = N, n = N;
#pragma unroll
    for (int s = 0; s < fft_stages(N); ++s) {
        const int r = (n % 8 == 0) ? 8 : n;
        len /= r;
        pos += (k % r) * len;
        k /= r;
        n /= r;
    }
    return pos;
}

// inverse of digit_rev: natural index of the element sitting at position pos
template <int N>
__host__ __device__ inline int digit_unrev(int pos)
{
    int k = 0, len = N, n = N, w = 1;
#pragma unroll
    for (int s = 0; s < fft_stages(N); ++s) {
        const int r = (n % 8 == 0) ? 8 : n;
        len /= r;
        k += ((pos / len) % r) * w;
        w *= r;
        n /= r;
    }
    return k;
}

// ---- complex helpers ----------------------------------------------------------------------------
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 cmul(float2 a, float2 b)
{
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ float2 cmulc(float2 a, float2 b)  // a * conj(b)
{
    return make_float2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y);
}

// S = -1: forward kernel exp(-2 pi i jk/R);  S = +1: inverse.  Natural order in and out.
template <int S>
__device__ __forceinline__ void dft2(float2 &a, float2 &b)
{
    const float2 t = a;
    a = cadd(t, b);
    b = csub(t, b);
}
template <int S>
__device__ __forceinline__ void dft4(float2 &v0, float2 &v1, float2 &v2, float2 &v3)
{
    const float2 s = cadd(v0, v2), d = csub(v0, v2), t = cadd(v1, v3), u = csub(v1, v3);
    v0 = cadd(s, t);
    v2 = csub(s, t);
    v1 = make_float2(d.x - S * u.y, d.y + S * u.x);  // d + S*i*u
    v3 = make_float2(d.x + S * u.y, d.y - S * u.x);  // d - S*i*u
}
template <int S>
__device__ __forceinline__ void dft8(float2 (&v)[8])
{
    dft4<S>(v[0], v[2], v[4], v[6]);
    dft4<S>(v[1], v[3], v[5], v[7]);
    const float c = 0.70710678118654752440f;
    const float2 o0 = v[1];
    const float2 o1 = make_float2(c * (v[3].x - S * v[3].y), c * (S * v[3].x + v[3].y));
    const float2 o2 = make_float2(-S * v[5].y, S * v[5].x);
    const float2 o3 = make_float2(c * (-v[7].x - S * v[7].y), c * (S * v[7].x - v[7].y));
    const float2 e0 = v[0], e1 = v[2], e2 = v[4], e3 = v[6];
    v[0] = cadd(e0, o0); v[4] = csub(e0, o0);
    v[1] = cadd(e1, o1); v[5] = csub(e1, o1);
    v[2] = cadd(e2, o2); v[6] = csub(e2, o2);
    v[3] = cadd(e3, o3); v[7] = csub(e3, o3);
}
template <int R, int S>
__device__ __forceinline__ void dft(float2 (&v)[R])
{
    if constexpr (R == 8) dft8<S>(v);
    else if constexpr (R == 4) dft4<S>(v[0], v[1], v[2], v[3]);
    else dft2<S>(v[0], v[1]);
}

// One butterfly of stage ST of an N-point transform on the element set
//   pos(r) = b*L + i + r*(L/R),   u = b*(L/R) + i  in [0, N/R)
// FWD: DIF  (DFT_R, then twiddle exp(-2 pi i * i*q / L));  !FWD: the exact inverse (DIT).
// `tw` is the table exp(-2 pi i m / TWN), TWN a multiple of N (rows use the N-point table for
// their N/2-point transforms).  ld(pos) / st(pos, value) do the addressing.
template <int N, int ST, bool FWD, int TWN, class Ld, class St, class Mid>
__device__ __forceinline__ void butterfly(int u, const float2 *tw, Ld ld, St st, Mid mid)
{
    constexpr int R = fft_radix(N, ST);
    constexpr int L = fft_len(N, ST);
    constexpr int SUB = L / R;
    const int b = u / SUB, i = u % SUB;  // SUB is a power of two
    const int pos0 = b * L + i;
    float2 v[R];
#pragma unroll
    for (int r = 0; r < R; ++r) v[r] = ld(pos0 + r * SUB);
    if constexpr (FWD) {
        dft<R, -1>(v);
        if constexpr (SUB > 1) {
#pragma unroll
            for (int q = 1; q < R; ++q) v[q] = cmul(v[q], tw[i * q * (TWN / L)]);
        }
    } else {
        if constexpr (SUB > 1) {
#pragma unroll
            for (int q = 1; q < R; ++q) v[q] = cmulc(v[q], tw[i * q * (TWN / L)]);
        }
        dft<R, +1>(v);
    }
    mid(v, pos0);
#pragma unroll
    for (int r = 0; r < R; ++r) st(pos0 + r * SUB, v[r]);
}

struct NoMid {
    template <class V>
    __device__ __forceinline__ void operator()(V &, int) const {}
};

// G(k) of src/fourier_utils.py:15-16 in float32, summed as (s_axis0 + s_axis1) + s_axis2, 0 at DC
__device__ __forceinline__ float green_f32(float sz, float sy, float sx)
{
    const float ksq = (sz + sy) + sx;
#ifdef PM_GREEN_IEEE_DIV
    return ksq != 0.0f ? 1.0f / ksq : 0.0f;
#else
#ifdef PM_GREEN_EXACT_RCP
    // correctly rounded 1/ksq = np.divide(1, k_squared) of fourier_utils.py:16 bit for bit: measured +0.05 ms
    // on the 0.31 ms fused z pass of the 512^3 mesh, and invisible in every parity figure
    return ksq != 0.0f ? __frcp_rn(ksq) : 0.0f;
#else
    // 1/ksq by MUFU.RCP plus one Newton step: within 1 ulp of the reference's correctly rounded float32
    // quotient.  That ulp is far below the float32 transform noise of this path -- the float64 diagnostic
    // backend (pm_poisson.cu), which uses the exact quotient, shows which is which
    // (tests/test_gpu_parity.py, spike fixtures): with float32 transforms the exact reciprocal changed no
    // parity figure (profiles/r02_notes.md).
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(ksq));
    r = fmaf(r, fmaf(-ksq, r, 1.0f), r);
    return ksq != 0.0f ? r : 0.0f;
#endif
#endif
}

enum ColMode { COL_FWD = 0, COL_INV = 1, COL_FUSED = 2 };

struct ColArgs {
    float2 *main;      // [N][N][N/2] spectrum, digit-reversed along transformed axes
    float2 *side;      // [N][N] Nyquist-in-x plane
    const float2 *tw;  // exp(-2 pi i m / N)
    const float *sin2, *sin2rev;   // natural / digit-reversed sin^2(pi i/N)
    const float *sin2y;            // in the order the y pass left the y axis in (v1: sin2rev, v2: sin2)
    float scale;       // -3*Omega_m/(8a)/N^3
    const float *scale_ptr;   // non-null: the factor is read from device memory instead (CUDA-graph replays, pm_api.cu)
    int axis;          // 1: along y (tiles = z x kx-tile), 0: along z (tiles = y x kx-tile, + side)
    // z pass geometry: the array is [N z][nyl][hw] holding y positions [y0, y0+nyl) -- the whole
    // y range on one GPU, this rank's share after the all-to-all transpose in slab mode
    int nyl, y0;
    // kx chunking (slab pipeline): this launch covers kx tiles [kt0, kt0+tpr); the z-pass array
    // holds only those columns, hw = tpr*16 float2 per row (y pass: the full-width array, hw = N/2)
    int tpr, kt0, hw;
    int side_tiles;    // z pass: also process the Nyquist plane tiles (first chunk only)
};

// Two adjacent columns per thread: every tile access is a 16-byte LDS/STS/LDG/STG and the two
// butterflies share their twiddles -- half the memory instructions per point of a scalar version.
template <int N, int ST, bool FWD, class Ld4, class St4, class Mid>
__device__ __forceinline__ void butterfly2(int u, const float2 *tw, Ld4 ld, St4 st, Mid mid)
{
    constexpr int R = fft_radix(N, ST);
    constexpr int L = fft_len(N, ST);
    constexpr int SUB = L / R;
    const int b = u / SUB, i = u % SUB;
    const int pos0 = b * L + i;
    float2 va[R], vb[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const float4 q = ld(pos0 + r * SUB);
        va[r] = make_float2(q.x, q.y);
        vb[r] = make_float2(q.z, q.w);
    }
    if constexpr (FWD) {
        dft<R, -1>(va);
        dft<R, -1>(vb);
        if constexpr (SUB > 1) {
#pragma unroll
            for (int q = 1; q < R; ++q) {
                const float2 w = tw[i * q * (N / L)];
                va[q] = cmul(va[q], w);
                vb[q] = cmul(vb[q], w);
            }
        }
    } else {
        if constexpr (SUB > 1) {
#pragma unroll
            for (int q = 1; q < R; ++q) {
                const float2 w = tw[i * q * (N / L)];
                va[q] = cmulc(va[q], w);
                vb[q] = cmulc(vb[q], w);
            }
        }
        dft<R, +1>(va);
        dft<R, +1>(vb);
    }
    mid(va, vb, pos0);
#pragma unroll
    for (int r = 0; r < R; ++r) st(pos0 + r * SUB, make_float4(va[r].x, va[r].y, vb[r].x, vb[r].y));
}

// Column FFT pass.  A tile is N points (stride `gs` float2 apart) x 16 adjacent columns, held in
// shared memory as [N][16] float2 (128-byte rows: a quarter-warp's 16-byte accesses cover one row,
// conflict-free), followed by the split-off Nyquist column [N] and the twiddle table [N].
// CG: global loads bypass L1 (ld.global.cg).  The plane kernels below read data that another CTA
// of the same launch wrote moments ago; L1 is not coherent across SMs, L2 is.
template <bool CG, class T>
__device__ __forceinline__ T pm_ld(const T *p)
{
    if constexpr (CG) return __ldcg(p);
    else return *p;
}

// Peer-memory y passes (slab decomposition): the forward y pass stores its result straight into
// the z-pass arrays of the ranks that own the y positions (position pos -> rank pos / nyl, over
// NVLink when that is another GPU), the inverse y pass loads its input from there -- the
// all-to-all transposes of the distributed transform happen in the store / load phase of the
// butterflies, with no pack kernel, staging buffer or collective.
struct PeerArgs {
    float2 *p[PM_PEER_MAX];   // every rank's z-pass array C[z][nyl][hc] (+ side plane [z][nyl])
    size_t main_off;          // this kx chunk's offset inside that array
    size_t side_off;          // offset of the Nyquist plane
    int rank, nzl, nyl_shift, hc;
};

// One tile of a column pass.  s_tile: [N][cols] float2, s_x: [N] (split-off Nyquist column),
// s_tw: [N] twiddles, already loaded and visible.
template <int N, int MODE, bool CG, bool PEER = false>
__device__ __forceinline__ void fft_cols_tile(const ColArgs &a, const int t, float2 *s_tile,
                                              float2 *s_x, const float2 *s_tw,
                                              const PeerArgs *pa = nullptr)
{
    constexpr int H = N / 2;
    constexpr int S = fft_stages(N);
    constexpr int CP = kColsCN<N> / 2;       // column pairs
    const int tid = threadIdx.x;

    // ---- which tile ----
    float2 *g;            // first element of the tile
    size_t gs;            // stride between successive points
    bool extra = false;   // y pass, kx-tile 0: also carries the split-off Nyquist column
    float2 *gx = nullptr; // its global home: side[z][.]
    float sy_fixed = 0.f; // Green's: sin^2 term that is constant over the tile
    int col0 = 0;         // first column index (kx for main tiles, y position for side tiles)
    bool side_tile = false;
    const int TPR = a.tpr;
    size_t peer_row = 0;   // PEER: row of (this rank's plane z, y position 0) in a peer's array
    int peer_col = 0;      //       first column of the tile inside the chunk
    if (a.axis == 1) {
        const int z = t / TPR, kt = a.kt0 + t % TPR;
        g = a.main + (size_t)z * N * H + kt * kColsCN<N>;
        gs = H;
        if (kt == 0) {
            extra = true;
            gx = a.side + (size_t)z * N;
        }
        if constexpr (PEER) {
            peer_row = ((size_t)pa->rank * pa->nzl + z) << pa->nyl_shift;
            peer_col = (t % TPR) * kColsCN<N>;
        }
    } else {
        if (t < a.nyl * TPR) {
            const int yl = t / TPR, ktl = t % TPR;
            g = a.main + (size_t)yl * a.hw + ktl * kColsCN<N>;
            gs = (size_t)a.nyl * a.hw;
            col0 = (a.kt0 + ktl) * kColsCN<N>;
            if (MODE == COL_FUSED) sy_fixed = __ldg(a.sin2y + a.y0 + yl);
        } else {
            side_tile = true;
            const int yt = t - a.nyl * TPR;
            g = a.side + yt * kColsCN<N>;
            gs = a.nyl;
            col0 = a.y0 + yt * kColsCN<N>;
        }
    }

    auto sm4 = [&](int pos, int cp) -> float4 & {
        return *reinterpret_cast<float4 *>(s_tile + pos * kColsCN<N> + 2 * cp);
    };
    // PEER: where y position `pos` of this tile lives on the rank that owns it
    auto peer_main = [&](int pos) -> float2 * {
        const int s = pos >> pa->nyl_shift, yl = pos & ((1 << pa->nyl_shift) - 1);
        return pa->p[s] + pa->main_off + (peer_row + yl) * pa->hc + peer_col;
    };
    auto peer_side = [&](int pos) -> float2 * {
        const int s = pos >> pa->nyl_shift, yl = pos & ((1 << pa->nyl_shift) - 1);
        return pa->p[s] + pa->side_off + peer_row + yl;
    };

    // ---- generic stage runner over the 8 column pairs (+ the extra column) ----
    // src/dst: 0 = shared tile, 1 = global.
    auto run_stage = [&](auto st_tag, auto fwd_tag, auto src_tag, auto dst_tag) {
        constexpr int ST = decltype(st_tag)::value;
        constexpr bool FWD = decltype(fwd_tag)::value;
        constexpr int src = decltype(src_tag)::value, dst = decltype(dst_tag)::value;
        constexpr int R = fft_radix(N, ST);
        constexpr int NB = N / R;
        constexpr int ITERS = (NB * CP + kThrC<N> - 1) / kThrC<N>;
#pragma unroll
        for (int it = 0; it < ITERS; ++it) {
            const int w = it * kThrC<N> + tid;
            const int cp = w % CP, u = w / CP;
            if (w < NB * CP) {
                auto ld = [&](int pos) -> float4 {
                    if (src == 0) return sm4(pos, cp);
                    if constexpr (PEER && !FWD)   // inverse y pass: pull from the owner of this y position
                        return *reinterpret_cast<const float4 *>(peer_main(pos) + 2 * cp);
                    float4 q = pm_ld<CG>(reinterpret_cast<const float4 *>(g + (size_t)pos * gs + 2 * cp));
                    if (extra && cp == 0 && FWD) q.y = 0.0f;  // packed slot: real part = DC column
                    return q;
                };
                auto st = [&](int pos, float4 q) {
                    if (dst == 0 || (extra && cp == 0 && !FWD)) sm4(pos, cp) = q;  // inverse: merged later
                    else if constexpr (PEER && FWD)   // forward y pass: push to the owner of this y position
                        *reinterpret_cast<float4 *>(peer_main(pos) + 2 * cp) = q;
                    else *reinterpret_cast<float4 *>(g + (size_t)pos * gs + 2 * cp) = q;
                };
                butterfly2<N, ST, FWD>(u, s_tw, ld, st, [](auto &, auto &, int) {});
            }
        }
        if (extra) {
            for (int u = tid; u < NB; u += kThrC<N>) {
                auto ld = [&](int pos) -> float2 {
                    if (src == 0) return s_x[pos];
                    if (FWD) return make_float2(pm_ld<CG>(g + (size_t)pos * gs).y, 0.0f);  // Nyquist part of the packed slot
                    if constexpr (PEER) return *peer_side(pos);
                    return pm_ld<CG>(gx + pos);
                };
                auto st = [&](int pos, float2 v) {
                    if (dst == 0 || !FWD) s_x[pos] = v;
                    else if constexpr (PEER) *peer_side(pos) = v;
                    else gx[pos] = v;
                };
                butterfly<N, ST, FWD, N>(u, s_tw, ld, st, NoMid());
            }
        }
    };

#define PM_ST(k) std::integral_constant<int, (k)>()
#define PM_T std::true_type()
#define PM_F std::false_type()

    if constexpr (MODE == COL_FWD) {
        // stage 0 from global, stages 1..S-2 in shared memory, stage S-1 to global
        run_stage(PM_ST(0), PM_T, PM_ST(1), PM_ST(0));
        __syncthreads();
        if constexpr (S >= 3) { run_stage(PM_ST(1), PM_T, PM_ST(0), PM_ST(0)); __syncthreads(); }
        if constexpr (S >= 4) { run_stage(PM_ST(2), PM_T, PM_ST(0), PM_ST(0)); __syncthreads(); }
        run_stage(PM_ST(S - 1), PM_T, PM_ST(0), PM_ST(1));
    } else if constexpr (MODE == COL_INV) {
        run_stage(PM_ST(S - 1), PM_F, PM_ST(1), PM_ST(0));
        __syncthreads();
        if constexpr (S >= 4) { run_stage(PM_ST(2), PM_F, PM_ST(0), PM_ST(0)); __syncthreads(); }
        if constexpr (S >= 3) { run_stage(PM_ST(1), PM_F, PM_ST(0), PM_ST(0)); __syncthreads(); }
        run_stage(PM_ST(0), PM_F, PM_ST(0), PM_ST(1));
        if (extra) {
            // merge: packed slot = (Re DC column, Re Nyquist column); both are real up to rounding.
            // Column 1 shared the 16-byte accesses of column 0, so it is flushed here too.
            __syncthreads();
            for (int pos = tid; pos < N; pos += kThrC<N>) {
                const float4 q = sm4(pos, 0);
                *reinterpret_cast<float4 *>(g + (size_t)pos * gs) = make_float4(q.x, s_x[pos].x, q.z, q.w);
            }
        }
    } else {
        // forward along z, Green's function, inverse along z -- one trip through HBM
        run_stage(PM_ST(0), PM_T, PM_ST(1), PM_ST(0));
        __syncthreads();
        if constexpr (S >= 3) { run_stage(PM_ST(1), PM_T, PM_ST(0), PM_ST(0)); __syncthreads(); }
        if constexpr (S >= 4) { run_stage(PM_ST(2), PM_T, PM_ST(0), PM_ST(0)); __syncthreads(); }
        {
            // last forward stage, multiply, first inverse stage: same R consecutive points
            constexpr int ST = S - 1;
            constexpr int R = fft_radix(N, ST);
            constexpr int NB = N / R;
            constexpr int ITERS = (NB * CP + kThrC<N> - 1) / kThrC<N>;
#pragma unroll
            for (int it = 0; it < ITERS; ++it) {
                const int w = it * kThrC<N> + tid;
                const int cp = w % CP, pos0 = (w / CP) * R;
                if (w < NB * CP) {
                    float2 va[R], vb[R];
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        const float4 q = sm4(pos0 + r, cp);
                        va[r] = make_float2(q.x, q.y);
                        vb[r] = make_float2(q.z, q.w);
                    }
                    dft<R, -1>(va);
                    dft<R, -1>(vb);
                    const int c = col0 + 2 * cp;
                    const float sya = side_tile ? __ldg(a.sin2y + c) : sy_fixed;
                    const float syb = side_tile ? __ldg(a.sin2y + c + 1) : sy_fixed;
                    const float sxa = side_tile ? __ldg(a.sin2 + H) : __ldg(a.sin2 + c);
                    const float sxb = side_tile ? sxa : __ldg(a.sin2 + c + 1);
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        const float sz = __ldg(a.sin2rev + pos0 + r);
                        const float ga = a.scale * green_f32(sz, sya, sxa);
                        const float gb = a.scale * green_f32(sz, syb, sxb);
                        va[r].x *= ga; va[r].y *= ga;
                        vb[r].x *= gb; vb[r].y *= gb;
                    }
                    dft<R, +1>(va);
                    dft<R, +1>(vb);
#pragma unroll
                    for (int r = 0; r < R; ++r)
                        sm4(pos0 + r, cp) = make_float4(va[r].x, va[r].y, vb[r].x, vb[r].y);
                }
            }
        }
        __syncthreads();
        if constexpr (S >= 4) { run_stage(PM_ST(2), PM_F, PM_ST(0), PM_ST(0)); __syncthreads(); }
        if constexpr (S >= 3) { run_stage(PM_ST(1), PM_F, PM_ST(0), PM_ST(0)); __syncthreads(); }
        run_stage(PM_ST(0), PM_F, PM_ST(0), PM_ST(1));
    }
}

// Column FFT pass, one tile per CTA.  Twiddles live in shared memory: with the carveout these
// tiles need, L1 is too small to keep a __ldg table resident against the streaming tile traffic.
template <int N, int MODE>
__global__ void __launch_bounds__(kThrC<N>, PM_FFT_MINB) k_fft_cols(ColArgs a)
{
    if (MODE == COL_FUSED && a.scale_ptr) a.scale = __ldg(a.scale_ptr);
    extern __shared__ float2 s_dyn[];
    float2 *s_tile = s_dyn;
    float2 *s_x = s_tile + N * kColsCN<N>;   // extra column
    float2 *s_tw = s_x + N;
    for (int m = threadIdx.x; m < N; m += kThrC<N>) s_tw[m] = a.tw[m];
    __syncthreads();
    fft_cols_tile<N, MODE, false>(a, blockIdx.x, s_tile, s_x, s_tw);
}

// The y passes of the slab path with the transposes fused in (PeerArgs above).
template <int N, int MODE>
__global__ void __launch_bounds__(kThrC<N>, PM_FFT_MINB) k_fft_cols_peer(ColArgs a, PeerArgs pa)
{
    static_assert(MODE == COL_FWD || MODE == COL_INV, "y passes only");
    extern __shared__ float2 s_dyn[];
    float2 *s_tile = s_dyn;
    float2 *s_x = s_tile + N * kColsCN<N>;
    float2 *s_tw = s_x + N;
    for (int m = threadIdx.x; m < N; m += kThrC<N>) s_tw[m] = a.tw[m];
    __syncthreads();
    fft_cols_tile<N, MODE, false, true>(a, blockIdx.x, s_tile, s_x, s_tw, &pa);
}

// Row pass: 16 x-rows per CTA, each an (N/2)-point complex FFT of z_j = x_2j + i x_2j+1 plus the
// real-transform split (forward) or merge (inverse).  Global accesses run along the row
// (coalesced), the butterflies run across the 16 rows (conflict-free): the tile is [N/2][17].
// The split/merge pairs k with N/2-k, so it wants natural order: the last forward stage (and the
// first inverse stage) is done out of place -- read everything, barrier, write to the natural
// positions -- which keeps every shared-memory access of the kernel free of bank conflicts.
// One 16-row tile of a row pass.  s_tile: [N/2][17] float2; s_tw: [N] twiddles in shared memory
// (they only need to be visible after the first barrier; `tw` is the same table in global memory).
template <int N, bool FWD, bool CG>
__device__ __forceinline__ void fft_rows_tile(const float2 *__restrict__ in, float2 *__restrict__ out,
                                              const float2 *__restrict__ tw, const size_t row0,
                                              float2 *s_tile, const float2 *s_tw, const float mean = 0.f)
{
    constexpr int H = N / 2;
    constexpr int S = fft_stages(H);
    constexpr int RL = fft_radix(H, S - 1);   // radix of the last DIF stage (sub-length 1)
    constexpr int NBL = H / RL;
    constexpr int ITL = (NBL * kCols + kThr<N> - 1) / kThr<N>;
    constexpr int LD_IT = kCols * H / kThr<N>;   // tile elements per thread
    const int tid = threadIdx.x;
    auto sm = [&](int pos, int c) -> float2 & { return s_tile[pos * kPitch + c]; };

    auto run_stage = [&](auto st_tag, auto fwd_tag) {
        constexpr int ST = decltype(st_tag)::value;
        constexpr bool F = decltype(fwd_tag)::value;
        constexpr int NB = H / fft_radix(H, ST);
        constexpr int ITERS = (NB * kCols + kThr<N> - 1) / kThr<N>;
#pragma unroll
        for (int it = 0; it < ITERS; ++it) {
            const int w = it * kThr<N> + tid;
            const int c = w % kCols, u = w / kCols;
            if (w < NB * kCols) {
                auto ld = [&](int pos) -> float2 { return sm(pos, c); };
                auto st = [&](int pos, float2 v) { sm(pos, c) = v; };
                butterfly<H, ST, F, N>(u, s_tw, ld, st, NoMid());
            }
        }
        __syncthreads();
    };

    if constexpr (FWD) {
        {
            float2 buf[LD_IT];
#pragma unroll
            for (int it = 0; it < LD_IT; ++it) {
                const int idx = it * kThr<N> + tid;
                buf[it] = in[(row0 + idx / H) * H + idx % H];   // rho: written by an earlier launch
                buf[it].x -= mean;   // rho - <rho>: the DC mode is zeroed by the Green's factor anyway, and
                buf[it].y -= mean;   // without it no float32 rounding of the huge DC lineage leaks into low k
            }
#pragma unroll
            for (int it = 0; it < LD_IT; ++it) {
                const int idx = it * kThr<N> + tid;
                sm(idx % H, idx / H) = buf[it];
            }
        }
        __syncthreads();
        if constexpr (S >= 2) run_stage(PM_ST(0), PM_T);
        if constexpr (S >= 3) run_stage(PM_ST(1), PM_T);
        if constexpr (S >= 4) run_stage(PM_ST(2), PM_T);
        {
            float2 v[ITL][RL];
#pragma unroll
            for (int it = 0; it < ITL; ++it) {
                const int w = it * kThr<N> + tid;
                const int c = w % kCols, pos0 = (w / kCols) * RL;
                if (w < NBL * kCols) {
#pragma unroll
                    for (int r = 0; r < RL; ++r) v[it][r] = sm(pos0 + r, c);
                    dft<RL, -1>(v[it]);
                }
            }
            __syncthreads();
#pragma unroll
            for (int it = 0; it < ITL; ++it) {
                const int w = it * kThr<N> + tid;
                const int c = w % kCols, pos0 = (w / kCols) * RL;
                if (w < NBL * kCols) {
                    const int k0 = digit_unrev<H>(pos0);
#pragma unroll
                    for (int r = 0; r < RL; ++r) sm(k0 + r * NBL, c) = v[it][r];
                }
            }
            __syncthreads();
        }
#pragma unroll 4
        for (int it = 0; it < LD_IT; ++it) {
            const int idx = it * kThr<N> + tid;
            const int b = idx / H, k = idx % H;
            float2 X;
            const float2 Zk = sm(k, b);
            if (k == 0) {
                X = make_float2(Zk.x + Zk.y, Zk.x - Zk.y);  // packed (DC, Nyquist)
            } else {
                const float2 Zm = sm(H - k, b);
                const float2 A = make_float2(Zk.x + Zm.x, Zk.y - Zm.y);  // Zk + conj(Zm)
                const float2 B = make_float2(Zk.x - Zm.x, Zk.y + Zm.y);  // Zk - conj(Zm)
                const float2 t = cmul(s_tw[k], B);
                X = make_float2(0.5f * (A.x + t.y), 0.5f * (A.y - t.x)); // (A - i t) / 2
            }
            out[(row0 + b) * H + k] = X;
        }
    } else {
#pragma unroll 4
        for (int it = 0; it < LD_IT; ++it) {
            const int idx = it * kThr<N> + tid;
            const int b = idx / H, k = idx % H;
            const float2 A = pm_ld<CG>(in + (row0 + b) * H + k);
            float2 Z;
            if (k == 0) {
                Z = make_float2(A.x + A.y, A.x - A.y);  // (DC + Nyq) + i (DC - Nyq)
            } else {
                const float2 Bm = pm_ld<CG>(in + (row0 + b) * H + (H - k));
                const float2 P = make_float2(A.x + Bm.x, A.y - Bm.y);   // A + conj(B)
                const float2 Q = make_float2(A.x - Bm.x, A.y + Bm.y);   // A - conj(B)
                const float2 t = cmulc(Q, __ldg(tw + k));               // conj(w^k) * Q (before the first barrier)
                Z = make_float2(P.x - t.y, P.y + t.x);                  // P + i t
            }
            sm(k, b) = Z;
        }
        __syncthreads();
        {
            float2 v[ITL][RL];
#pragma unroll
            for (int it = 0; it < ITL; ++it) {
                const int w = it * kThr<N> + tid;
                const int c = w % kCols, pos0 = (w / kCols) * RL;
                if (w < NBL * kCols) {
                    const int k0 = digit_unrev<H>(pos0);
#pragma unroll
                    for (int r = 0; r < RL; ++r) v[it][r] = sm(k0 + r * NBL, c);
                    dft<RL, +1>(v[it]);
                }
            }
            __syncthreads();
#pragma unroll
            for (int it = 0; it < ITL; ++it) {
                const int w = it * kThr<N> + tid;
                const int c = w % kCols, pos0 = (w / kCols) * RL;
                if (w < NBL * kCols) {
#pragma unroll
                    for (int r = 0; r < RL; ++r) sm(pos0 + r, c) = v[it][r];
                }
            }
            __syncthreads();
        }
        if constexpr (S >= 4) run_stage(PM_ST(2), PM_F);
        if constexpr (S >= 3) run_stage(PM_ST(1), PM_F);
        if constexpr (S >= 2) run_stage(PM_ST(0), PM_F);
#pragma unroll 4
        for (int it = 0; it < LD_IT; ++it) {
            const int idx = it * kThr<N> + tid;
            out[(row0 + idx / H) * H + idx % H] = sm(idx % H, idx / H);
        }
    }
}

template <int N, bool FWD>
__global__ void __launch_bounds__(kThr<N>) k_fft_rows(const float2 *__restrict__ in,
                                                       float2 *__restrict__ out,
                                                       const float2 *__restrict__ tw,
                                                       const float *__restrict__ mean_ptr)
{
    extern __shared__ float2 s_dyn[];
    float2 *s_tw = s_dyn + (N / 2) * kPitch;   // N-entry twiddle table (visible after the first barrier)
    for (int m = threadIdx.x; m < N; m += kThr<N>) s_tw[m] = tw[m];
    const float mean = (FWD && mean_ptr) ? __ldg(mean_ptr) : 0.f;   // forward: transform rho - <rho>
    fft_rows_tile<N, FWD, false>(in, out, tw, (size_t)blockIdx.x * kCols, s_dyn, s_tw, mean);
}

#include "pm_fft2.cuh"
#include "pm_fft3.cuh"

template <int N, int MODE>
__global__ void __launch_bounds__(kThr2, 2) k_fft2_cols(ColArgs a)
{
    if (MODE == COL_FUSED && a.scale_ptr) a.scale = __ldg(a.scale_ptr);
    extern __shared__ float2 s_dyn[];
    float2 *s_tw = s_dyn, *s_tile = s_dyn + N;
    for (int m = threadIdx.x; m < N; m += kThr2) s_tw[m] = a.tw[m];
    __syncthreads();
    fft2_cols_tile<N, MODE, false>(a, blockIdx.x, s_tile, s_tw);
}

template <int N, bool FWD>
__global__ void __launch_bounds__(kThr2, N >= 2048 ? 1 : 2) k_fft2_rows(const float2 *__restrict__ in,
                                                        float2 *__restrict__ out,
                                                        const float2 *__restrict__ tw,
                                                        const float *__restrict__ mean_ptr)
{
    extern __shared__ float2 s_dyn[];
    float2 *s_tw = s_dyn, *s_tile = s_dyn + N;
    for (int m = threadIdx.x; m < N; m += kThr2) s_tw[m] = tw[m];
    __syncthreads();
    const float mean = (FWD && mean_ptr) ? __ldg(mean_ptr) : 0.f;   // forward: transform rho - <rho>
    fft2_rows_tile<N, FWD, false>(in, out, (size_t)blockIdx.x * kRowsPT2<N>, s_tile, s_tw, mean);
}

// ---- x and y passes of one direction in ONE persistent launch ------------------------------------
// The row pass and the y pass of a mesh plane touch the same 4*N^2 bytes.  Run as two launches
// over the whole mesh, the spectrum makes a round trip through HBM between them (the mesh is
// several times the 126 MB L2).  k_fft_plane hands out work items through a ticket counter in an
// order that keeps the consumer of a plane `lag` planes behind its producer:
//     forward:  slot q = [row tiles of plane q][y tiles of plane q - lag]
//     inverse:  slot q = [y tiles of plane q][row tiles of plane q - lag]
// so the intermediate plane is still in L2 when it is read back and each direction moves 8*N^3
// bytes through HBM instead of 16*N^3.  A consumer item waits (acquire) on its plane's counter of
// finished producer items; producers never wait and tickets are taken in order, so every wait is
// on an item some running CTA already holds: no deadlock, whatever the number of resident CTAs.
// Consumers read the intermediate with ld.global.cg (L1 is not coherent across SMs).
struct PlaneArgs {
    ColArgs ca;             // the y pass (axis = 1, all kx tiles)
    const float2 *rows_in;  // forward: rho viewed as float2 rows; inverse: the spectrum
    float2 *rows_out;       // forward: the spectrum; inverse: phi viewed as float2 rows
    unsigned *ticket;       // [0] ticket counter, [1 + z] finished producer items of plane z
    unsigned *err;          // set if a wait gave up (must never happen; checked by the tests)
    int nplanes, lag;
    const float *mean_ptr;  // forward: <rho>, subtracted while the rows are loaded (nullptr: 0)
};

__device__ __forceinline__ unsigned pm_ld_acquire(const unsigned *p)
{
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

template <int N, bool FWD, bool V2>
__global__ void __launch_bounds__(kThrC<N>, V2 ? 2 : PM_FFT_MINB) k_fft_plane(PlaneArgs pa)
{
    static_assert(kThr<N> == kThrC<N> && kThrC<N> == kThr2, "row and column tiles share the CTA");
    extern __shared__ float2 s_dyn[];
    __shared__ int s_ticket;
    float2 *s_tw = s_dyn;                      // [N]
    float2 *s_tile = s_dyn + N;                // column tile [N][cols] + [N], or row tile [N/2][17]
    float2 *s_x = s_tile + N * kColsCN<N>;
    static_assert(N * kColsCN<N> + N >= (N / 2) * kPitch, "row tile fits the column tile's space");
    static_assert(!V2 || N * kColsCN<N> >= kRowsPT2<N> * RowFac<N>::RA * (RowFac<N>::RB + 1), "two-stage row tile fits");
    for (int m = threadIdx.x; m < N; m += kThrC<N>) s_tw[m] = pa.ca.tw[m];
    constexpr int RT = V2 ? kRowsPT2<N> : kCols;   // x-rows per row tile
    constexpr int RPP = N / RT;                // row tiles per plane
    const int TPP = pa.ca.tpr;                 // y tiles per plane
    const int nprod = FWD ? RPP : TPP;         // producer items per plane (first in a slot)
    const int per = RPP + TPP;
    const int total = (pa.nplanes + pa.lag) * per;
    unsigned *done = pa.ticket + 1;
    for (;;) {
        __syncthreads();                       // the previous item is finished with shared memory
        if (threadIdx.x == 0) s_ticket = (int)atomicAdd(pa.ticket, 1u);
        __syncthreads();
        const int t = s_ticket;
        if (t >= total) break;
        const int q = t / per, r = t - q * per;
        const bool producer = r < nprod;
        const int z = producer ? q : q - pa.lag;
        if (z < 0 || z >= pa.nplanes) continue;
        const int idx = producer ? r : r - nprod;
        if (!producer) {
            if (threadIdx.x == 0) {
                int spins = 0;
                while (pm_ld_acquire(done + z) < (unsigned)nprod) {
                    if (++spins > (1 << 20) || pm_ld_acquire(pa.err) != 0u) {   // never hang the GPU
                        atomicExch(pa.err, 1u);
                        break;
                    }
                    __nanosleep(200);
                }
                __threadfence();
            }
            __syncthreads();
        }
        if (FWD == producer) {
            const size_t row0 = (size_t)z * N + (size_t)idx * RT;
            const float mean = (FWD && pa.mean_ptr) ? __ldg(pa.mean_ptr) : 0.f;
            if constexpr (V2) fft2_rows_tile<N, FWD, !FWD>(pa.rows_in, pa.rows_out, row0, s_tile, s_tw, mean);
            else fft_rows_tile<N, FWD, !FWD>(pa.rows_in, pa.rows_out, pa.ca.tw, row0, s_tile, s_tw, mean);
        } else {
            if constexpr (V2) fft2_cols_tile<N, FWD ? COL_FWD : COL_INV, FWD>(pa.ca, z * TPP + idx, s_tile, s_tw);
            else fft_cols_tile<N, FWD ? COL_FWD : COL_INV, FWD>(pa.ca, z * TPP + idx, s_tile, s_x, s_tw);
        }
        if (producer) {
            __syncthreads();                   // every thread's stores are issued
            if (threadIdx.x == 0) {
                __threadfence();               // ... and visible device-wide before the count moves
                atomicAdd(done + z, 1u);
            }
        }
    }
}

template <int N>
constexpr bool kPlaneFused = (N >= 256 && N <= 1024);

// The two-stage path (pm_fft2.cuh): natural order on every axis.
template <int N>
int poisson_launch_v2(pm_plan *p, const float *rho, double a, double omega_m0, float *phi, cudaStream_t st)
{
    if constexpr (!kHasV2<N>) {
        return PM_ERR_UNSUPPORTED;
    } else {
        constexpr int H = N / 2, C = kColsCN<N>;
        constexpr int RT = kRowsPT2<N>;
        const size_t smem_cols = ((size_t)N + (size_t)N * C) * sizeof(float2);
        const size_t smem_rows = ((size_t)N + (size_t)RT * RowFac<N>::RA * (RowFac<N>::RB + 1)) * sizeof(float2);
        auto rows_fwd = k_fft2_rows<N, true>;
        auto rows_inv = k_fft2_rows<N, false>;
        auto cols_fwd = k_fft2_cols<N, COL_FWD>;
        auto cols_inv = k_fft2_cols<N, COL_INV>;
        auto cols_fused = k_fft2_cols<N, COL_FUSED>;
        auto plane_fwd = k_fft_plane<N, true, true>;
        auto plane_inv = k_fft_plane<N, false, true>;
        auto cols_fused_v1 = k_fft_cols<N, COL_FUSED>;
        auto cols3_fwd2 = k_fft3_cols<N, COL_FWD, 2, kThr2>;
        auto cols3_inv2 = k_fft3_cols<N, COL_INV, 2, kThr2>;
        auto cols3_fwd3 = k_fft3_cols<N, COL_FWD, 3, kThr2>;
        auto cols3_inv3 = k_fft3_cols<N, COL_INV, 3, kThr2>;
        constexpr size_t smem3_2 = kSmem3<N, 2>, smem3_3 = kSmem3<N, 3>;
        constexpr bool ring3_fits = smem3_3 <= 227 * 1024;
        const size_t smem_cols_v1 = ((size_t)N * kColsCN<N> + 2 * N) * sizeof(float2);
        static int plane_per_sm_dev[64];
        int &plane_per_sm = plane_per_sm_dev[p->device & 63];
        PM_ONCE_PER_DEVICE_BEGIN(p->device)
            PM_CUDA(cudaFuncSetAttribute(rows_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_rows));
            PM_CUDA(cudaFuncSetAttribute(rows_inv, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_rows));
            PM_CUDA(cudaFuncSetAttribute(cols_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_cols));
            PM_CUDA(cudaFuncSetAttribute(cols_inv, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_cols));
            PM_CUDA(cudaFuncSetAttribute(cols_fused, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_cols));
            PM_CUDA(cudaFuncSetAttribute(plane_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_cols));
            PM_CUDA(cudaFuncSetAttribute(plane_inv, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_cols));
            PM_CUDA(cudaFuncSetAttribute(cols_fwd, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
            PM_CUDA(cudaFuncSetAttribute(cols_inv, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
            PM_CUDA(cudaFuncSetAttribute(cols_fused, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
            PM_CUDA(cudaFuncSetAttribute(cols3_fwd2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem3_2));
            PM_CUDA(cudaFuncSetAttribute(cols3_inv2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem3_2));
            if (ring3_fits) {
                PM_CUDA(cudaFuncSetAttribute(cols3_fwd3, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem3_3));
                PM_CUDA(cudaFuncSetAttribute(cols3_inv3, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem3_3));
            }
            PM_CUDA(cudaFuncSetAttribute(cols_fused_v1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_cols_v1));
            PM_CUDA(cudaFuncSetAttribute(cols_fused_v1, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
            PM_CUDA(cudaFuncSetAttribute(plane_fwd, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
            PM_CUDA(cudaFuncSetAttribute(plane_inv, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
            PM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&plane_per_sm, plane_fwd, kThr2, smem_cols));
            if (plane_per_sm < 1) plane_per_sm = 1;
        PM_ONCE_PER_DEVICE_END()
        ColArgs ca;
        ca.main = p->spec;
        ca.side = p->spec + (size_t)N * N * H;
        ca.tw = p->tw;
        ca.sin2 = p->sin2;
        ca.sin2rev = p->sin2;   // natural order everywhere on this path
        ca.sin2y = p->sin2;
        const double m = (double)N * N * N;
        ca.scale = (float)(-3 * omega_m0 / 8 / a / m);
        ca.scale_ptr = p->graph_params ? &p->graph_params->green_scale : nullptr;
        ca.nyl = N;
        ca.y0 = 0;
        ca.tpr = H / C;
        ca.kt0 = 0;
        ca.hw = H;
        ca.side_tiles = 1;
        const int row_ctas = N * N / RT;
        const int tiles = N * (H / C);
        const int grid3 = tiles < p->sm_count ? tiles : p->sm_count;   // one persistent CTA per SM
        const bool fused = p->fft_fuse && p->fft_sync;
        PlaneArgs pa;
        if (fused) {
            PM_CUDA(cudaMemsetAsync(p->fft_sync + 1, 0, sizeof(unsigned) * 2 * (N + 1), st));
            pa.ca = ca;
            pa.ca.axis = 1;
            pa.err = p->fft_sync;
            pa.nplanes = N;
            pa.lag = p->fft_lag;
            pa.mean_ptr = p->rho_mean_d;
            pa.rows_in = reinterpret_cast<const float2 *>(rho);
            pa.rows_out = ca.main;
            pa.ticket = p->fft_sync + 1;
            PM_LAUNCH(plane_fwd, plane_per_sm * p->sm_count, kThr2, smem_cols, st, pa);
        } else {
            PM_LAUNCH(rows_fwd, row_ctas, kThr2, smem_rows, st, reinterpret_cast<const float2 *>(rho), ca.main,
                      (const float2 *)p->tw, (const float *)p->rho_mean_d);
            ca.axis = 1;
            if (p->fft_v3 == 3 && ring3_fits) PM_LAUNCH(cols3_fwd3, grid3, kThr2, smem3_3, st, ca, tiles);
            else if (p->fft_v3) PM_LAUNCH(cols3_fwd2, grid3, kThr2, smem3_2, st, ca, tiles);
            else PM_LAUNCH(cols_fwd, tiles, kThr2, smem_cols, st, ca);
        }
        pm_prof_mark(p, PM_STAGE_R2C + 1, st);
        ca.axis = 0;
        if (p->fft_zmix) {
            // z pass + Green + inverse z pass by the radix-8.8.8 kernel: 80 registers and three
            // CTAs per SM beat the two-stage version's 128 registers here (measured 0.35 vs 0.48 ms
            // at 512^3).  Its z order is internal to the pass; y stays natural (sin2y).
            ColArgs cz = ca;
            cz.sin2rev = p->sin2rev;
            PM_LAUNCH(cols_fused_v1, tiles + N / C, kThrC<N>, smem_cols_v1, st, cz);
        } else {
            PM_LAUNCH(cols_fused, tiles + N / C, kThr2, smem_cols, st, ca);
        }
        pm_prof_mark(p, PM_STAGE_GREEN + 1, st);
        if (fused) {
            pa.rows_in = ca.main;
            pa.rows_out = reinterpret_cast<float2 *>(phi);
            pa.ticket = p->fft_sync + 1 + (N + 1);
            PM_LAUNCH(plane_inv, plane_per_sm * p->sm_count, kThr2, smem_cols, st, pa);
        } else {
            ca.axis = 1;
            if (p->fft_v3 == 3 && ring3_fits) PM_LAUNCH(cols3_inv3, grid3, kThr2, smem3_3, st, ca, tiles);
            else if (p->fft_v3) PM_LAUNCH(cols3_inv2, grid3, kThr2, smem3_2, st, ca, tiles);
            else PM_LAUNCH(cols_inv, tiles, kThr2, smem_cols, st, ca);
            PM_LAUNCH(rows_inv, row_ctas, kThr2, smem_rows, st, reinterpret_cast<const float2 *>(ca.main),
                      reinterpret_cast<float2 *>(phi), (const float2 *)p->tw, (const float *)nullptr);
        }
        pm_prof_mark(p, PM_STAGE_C2R + 1, st);
        PM_CHECK_LAUNCH();
        return PM_OK;
    }
}

template <int N>
int poisson_launch(pm_plan *p, const float *rho, double a, double omega_m0, float *phi, cudaStream_t st)
{
    if constexpr (kHasV2<N>) {
        if (p->fft_v2) return poisson_launch_v2<N>(p, rho, a, omega_m0, phi, st);
    }
    constexpr int H = N / 2;
    const size_t smem_cols = ((size_t)N * kColsCN<N> + 2 * N) * sizeof(float2);
    const size_t smem_rows = ((size_t)H * kPitch + N) * sizeof(float2);
    PM_ONCE_PER_DEVICE_BEGIN(p->device)
        PM_CUDA(cudaFuncSetAttribute(k_fft_cols<N, COL_FWD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_cols));
        PM_CUDA(cudaFuncSetAttribute(k_fft_cols<N, COL_INV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_cols));
        PM_CUDA(cudaFuncSetAttribute(k_fft_cols<N, COL_FUSED>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_cols));
        PM_CUDA(cudaFuncSetAttribute(k_fft_rows<N, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_rows));
        PM_CUDA(cudaFuncSetAttribute(k_fft_rows<N, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_rows));
        // ask for the largest shared-memory carveout so the occupancy the tiles were sized for holds
        PM_CUDA(cudaFuncSetAttribute(k_fft_cols<N, COL_FWD>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        PM_CUDA(cudaFuncSetAttribute(k_fft_cols<N, COL_INV>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        PM_CUDA(cudaFuncSetAttribute(k_fft_cols<N, COL_FUSED>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        PM_CUDA(cudaFuncSetAttribute(k_fft_rows<N, true>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        PM_CUDA(cudaFuncSetAttribute(k_fft_rows<N, false>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    PM_ONCE_PER_DEVICE_END()
    ColArgs ca;
    ca.main = p->spec;
    ca.side = p->spec + (size_t)N * N * H;
    ca.tw = p->tw;
    ca.sin2 = p->sin2;
    ca.sin2rev = p->sin2rev;
    ca.sin2y = p->sin2rev;
    const double m = (double)N * N * N;
    ca.scale = (float)(-3 * omega_m0 / 8 / a / m);
    ca.scale_ptr = p->graph_params ? &p->graph_params->green_scale : nullptr;
    ca.nyl = N;
    ca.y0 = 0;
    ca.tpr = H / kColsCN<N>;
    ca.kt0 = 0;
    ca.hw = H;
    ca.side_tiles = 1;
    const int row_ctas = N * N / kCols;
    const int tiles = N * (H / kColsCN<N>);

    auto rows_fwd = k_fft_rows<N, true>;
    auto rows_inv = k_fft_rows<N, false>;
    auto cols_fwd = k_fft_cols<N, COL_FWD>;
    auto cols_inv = k_fft_cols<N, COL_INV>;
    auto cols_fused = k_fft_cols<N, COL_FUSED>;
    bool fused = false;
    if constexpr (kPlaneFused<N>) fused = p->fft_fuse && p->fft_sync;
    PlaneArgs pa;
    int plane_grid = 0;
    if constexpr (kPlaneFused<N>) {
        if (fused) {
            PM_ONCE_PER_DEVICE_BEGIN(p->device)
                PM_CUDA(cudaFuncSetAttribute(k_fft_plane<N, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_cols));
                PM_CUDA(cudaFuncSetAttribute(k_fft_plane<N, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_cols));
                PM_CUDA(cudaFuncSetAttribute(k_fft_plane<N, true, false>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
                PM_CUDA(cudaFuncSetAttribute(k_fft_plane<N, false, false>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
            PM_ONCE_PER_DEVICE_END()
            // one CTA per resident slot: the ticket loop hands every CTA its share of the items
            int per_sm = 0;
            PM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_fft_plane<N, true, false>, kThrC<N>, smem_cols));
            if (per_sm < 1) per_sm = 1;
            plane_grid = per_sm * p->sm_count;
            PM_CUDA(cudaMemsetAsync(p->fft_sync + 1, 0, sizeof(unsigned) * 2 * (N + 1), st));
            pa.ca = ca;
            pa.ca.axis = 1;
            pa.err = p->fft_sync;
            pa.nplanes = N;
            pa.lag = p->fft_lag;
            pa.mean_ptr = p->rho_mean_d;
        }
    }
    if (fused) {
        if constexpr (kPlaneFused<N>) {
            pa.rows_in = reinterpret_cast<const float2 *>(rho);
            pa.rows_out = ca.main;
            pa.ticket = p->fft_sync + 1;
            auto plane_fwd = k_fft_plane<N, true, false>;
            PM_LAUNCH(plane_fwd, plane_grid, kThrC<N>, smem_cols, st, pa);
        }
    } else {
        PM_LAUNCH(rows_fwd, row_ctas, kThr<N>, smem_rows, st, reinterpret_cast<const float2 *>(rho),
                  ca.main, (const float2 *)p->tw, (const float *)p->rho_mean_d);
        ca.axis = 1;
        PM_LAUNCH(cols_fwd, tiles, kThrC<N>, smem_cols, st, ca);
    }
    pm_prof_mark(p, PM_STAGE_R2C + 1, st);
    ca.axis = 0;
    PM_LAUNCH(cols_fused, tiles + N / kColsCN<N>, kThrC<N>, smem_cols, st, ca);
    pm_prof_mark(p, PM_STAGE_GREEN + 1, st);
    if (fused) {
        if constexpr (kPlaneFused<N>) {
            pa.rows_in = ca.main;
            pa.rows_out = reinterpret_cast<float2 *>(phi);
            pa.ticket = p->fft_sync + 1 + (N + 1);
            auto plane_inv = k_fft_plane<N, false, false>;
            PM_LAUNCH(plane_inv, plane_grid, kThrC<N>, smem_cols, st, pa);
        }
    } else {
        ca.axis = 1;
        PM_LAUNCH(cols_inv, tiles, kThrC<N>, smem_cols, st, ca);
        PM_LAUNCH(rows_inv, row_ctas, kThr<N>, smem_rows, st, reinterpret_cast<const float2 *>(ca.main),
                  reinterpret_cast<float2 *>(phi), (const float2 *)p->tw, (const float *)nullptr);
    }
    pm_prof_mark(p, PM_STAGE_C2R + 1, st);
    PM_CHECK_LAUNCH();
    return PM_OK;
}

// ---- slab decomposition: the same five passes with an all-to-all between 2|3 and 3|4 ----------
// Chunked slab pipeline.  The kx range is cut into C chunks of hc = (N/2)/C columns; chunk c of the
// packed send buffer is B_c[s][zl][yl][kc] at offset c*nzl*N*hc, and after the all-to-all the
// matching chunk of the receive buffer is C_c[z][yl][kc] -- so the y pass + pack of chunk c+1 and the
// z pass of chunk c-1 run while chunk c is on the wire.  The Nyquist plane travels with chunk 0.
template <bool UNPACK>
__global__ void __launch_bounds__(256) k_slab_pack_chunk(const float2 *__restrict__ full,
                                                         float2 *__restrict__ packed, int nzl, int n,
                                                         int h, int k0, int hc, int nyl)
{
    // full: [nzl][n][h] (this rank's planes, all y); packed: [P][nzl][nyl][hc]
    const size_t total = (size_t)nzl * n * hc;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (size_t)gridDim.x * blockDim.x) {
        const int kc = (int)(i % hc);
        const size_t r = i / hc;
        const int y = (int)(r % n), zl = (int)(r / n);
        const int s = y / nyl, yl = y - s * nyl;
        const size_t jf = ((size_t)zl * n + y) * h + k0 + kc;
        const size_t jp = (((size_t)s * nzl + zl) * nyl + yl) * hc + kc;
        if (UNPACK) const_cast<float2 *>(full)[jf] = packed[jp];
        else packed[jp] = full[jf];
    }
}

template <int N>
ColArgs slab_args(pm_plan *p)
{
    ColArgs ca;
    ca.main = p->spec;
    ca.side = p->spec + (size_t)p->nzl * N * (N / 2);
    ca.tw = p->tw; ca.sin2 = p->sin2; ca.sin2rev = p->sin2rev; ca.sin2y = p->sin2rev;
    ca.scale = 0.f; ca.scale_ptr = nullptr; ca.axis = 1;
    ca.nyl = N / p->nranks; ca.y0 = p->rank * ca.nyl;
    ca.tpr = (N / 2) / kColsCN<N>; ca.kt0 = 0; ca.hw = N / 2; ca.side_tiles = 1;
    return ca;
}

// Row passes of a slab: the two-stage register-resident kernels (pm_fft2.cuh) where they exist (256..2048
// points; same input, output layout and packing of the Nyquist term as k_fft_rows: natural kx), else the
// radix-8 kernels.  At 1024 points the former stream at 5 TB/s, the latter at 2.7.
template <int N>
int slab_rows_fwd(pm_plan *p, const float *rho, cudaStream_t st)
{
    if constexpr (kHasRows2<N>) {
        if (p->fft_v2) {
            constexpr int RT = kRowsPT2<N>;
            const size_t smem2 = ((size_t)N + (size_t)RT * RowFac<N>::RA * (RowFac<N>::RB + 1)) * sizeof(float2);
            auto rows2 = k_fft2_rows<N, true>;
            PM_CUDA(cudaFuncSetAttribute(rows2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
            PM_LAUNCH(rows2, p->nzl * N / RT, kThr2, smem2, st, reinterpret_cast<const float2 *>(rho), p->spec,
                      (const float2 *)p->tw, (const float *)p->rho_mean_d);
            PM_CHECK_LAUNCH();
            return PM_OK;
        }
    }
    const size_t smem_rows = ((size_t)(N / 2) * kPitch + N) * sizeof(float2);
    auto rows_fwd = k_fft_rows<N, true>;
    PM_CUDA(cudaFuncSetAttribute(rows_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_rows));
    PM_LAUNCH(rows_fwd, p->nzl * N / kCols, kThr<N>, smem_rows, st,
              reinterpret_cast<const float2 *>(rho), p->spec, (const float2 *)p->tw, (const float *)p->rho_mean_d);
    PM_CHECK_LAUNCH();
    return PM_OK;
}

template <int N>
int slab_y_fwd_pack(pm_plan *p, int c, int C, float2 *send_main_c, float2 *send_side, cudaStream_t st)
{
    constexpr int H = N / 2;
    const int nzl = p->nzl, nyl = N / p->nranks, hc = H / C;
    const size_t smem_cols = ((size_t)N * kColsCN<N> + 2 * N) * sizeof(float2);
    auto cols_fwd = k_fft_cols<N, COL_FWD>;
    PM_CUDA(cudaFuncSetAttribute(cols_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_cols));
    ColArgs ca = slab_args<N>(p);
    ca.tpr = hc / kColsCN<N>;
    ca.kt0 = c * ca.tpr;
    PM_LAUNCH(cols_fwd, nzl * ca.tpr, kThrC<N>, smem_cols, st, ca);
    const int grid = p->sm_count * 8;
    PM_LAUNCH(k_slab_pack_chunk<false>, grid, 256, 0, st, (const float2 *)ca.main, send_main_c, nzl, N,
              H, c * hc, hc, nyl);
    if (c == 0)
        PM_LAUNCH(k_slab_pack_chunk<false>, grid, 256, 0, st, (const float2 *)ca.side, send_side, nzl,
                  N, 1, 0, 1, nyl);
    PM_CHECK_LAUNCH();
    return PM_OK;
}

template <int N>
int slab_z_chunk(pm_plan *p, int c, int C, float2 *main_t_c, float2 *side_t, double a, double omega_m0,
                 cudaStream_t st)
{
    constexpr int H = N / 2;
    const int hc = H / C;
    const size_t smem_cols = ((size_t)N * kColsCN<N> + 2 * N) * sizeof(float2);
    auto cols_fused = k_fft_cols<N, COL_FUSED>;
    PM_CUDA(cudaFuncSetAttribute(cols_fused, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_cols));
    ColArgs ca = slab_args<N>(p);
    ca.main = main_t_c; ca.side = side_t;
    const double m = (double)N * N * N;
    ca.scale = (float)(-3 * omega_m0 / 8 / a / m);
    ca.axis = 0;
    ca.tpr = hc / kColsCN<N>; ca.kt0 = c * ca.tpr; ca.hw = hc;
    ca.side_tiles = (c == 0);
    const int tiles = ca.nyl * ca.tpr + (c == 0 ? ca.nyl / kColsCN<N> : 0);
    PM_LAUNCH(cols_fused, tiles, kThrC<N>, smem_cols, st, ca);
    PM_CHECK_LAUNCH();
    return PM_OK;
}

template <int N>
int slab_unpack_y_inv(pm_plan *p, int c, int C, const float2 *back_main_c, const float2 *back_side,
                      cudaStream_t st)
{
    constexpr int H = N / 2;
    const int nzl = p->nzl, nyl = N / p->nranks, hc = H / C;
    const size_t smem_cols = ((size_t)N * kColsCN<N> + 2 * N) * sizeof(float2);
    auto cols_inv = k_fft_cols<N, COL_INV>;
    PM_CUDA(cudaFuncSetAttribute(cols_inv, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_cols));
    ColArgs ca = slab_args<N>(p);
    ca.tpr = hc / kColsCN<N>;
    ca.kt0 = c * ca.tpr;
    const int grid = p->sm_count * 8;
    PM_LAUNCH(k_slab_pack_chunk<true>, grid, 256, 0, st, (const float2 *)ca.main,
              const_cast<float2 *>(back_main_c), nzl, N, H, c * hc, hc, nyl);
    if (c == 0)
        PM_LAUNCH(k_slab_pack_chunk<true>, grid, 256, 0, st, (const float2 *)ca.side,
                  const_cast<float2 *>(back_side), nzl, N, 1, 0, 1, nyl);
    PM_LAUNCH(cols_inv, nzl * ca.tpr, kThrC<N>, smem_cols, st, ca);
    PM_CHECK_LAUNCH();
    return PM_OK;
}

template <int N>
int slab_rows_inv(pm_plan *p, float *phi, cudaStream_t st)
{
    if constexpr (kHasRows2<N>) {
        if (p->fft_v2) {
            constexpr int RT = kRowsPT2<N>;
            const size_t smem2 = ((size_t)N + (size_t)RT * RowFac<N>::RA * (RowFac<N>::RB + 1)) * sizeof(float2);
            auto rows2 = k_fft2_rows<N, false>;
            PM_CUDA(cudaFuncSetAttribute(rows2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
            PM_LAUNCH(rows2, p->nzl * N / RT, kThr2, smem2, st, reinterpret_cast<const float2 *>(p->spec),
                      reinterpret_cast<float2 *>(phi), (const float2 *)p->tw, (const float *)nullptr);
            PM_CHECK_LAUNCH();
            return PM_OK;
        }
    }
    const size_t smem_rows = ((size_t)(N / 2) * kPitch + N) * sizeof(float2);
    auto rows_inv = k_fft_rows<N, false>;
    PM_CUDA(cudaFuncSetAttribute(rows_inv, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_rows));
    PM_LAUNCH(rows_inv, p->nzl * N / kCols, kThr<N>, smem_rows, st,
              reinterpret_cast<const float2 *>(p->spec), reinterpret_cast<float2 *>(phi),
              (const float2 *)p->tw, (const float *)nullptr);
    PM_CHECK_LAUNCH();
    return PM_OK;
}

// ---- peer-memory transposes (pm_slab_fft_push / pm_slab_fft_pull) ---------------------------------
// The all-to-all of the distributed transform done by the copy kernels themselves.  Rank r's
// z-pass array (tbuf[1]) chunk c is C_c[z][yl][kc]; the planes z in [s*nzl, (s+1)*nzl) of it belong
// to rank s's slab.  PUSH (forward leg): rank r stores its y-transformed planes into block [r] of
// every rank's array -- 16-byte stores, 2*hc*8 bytes contiguous per (zl, y) row, over NVLink for
// s != r.  PULL (way back): rank r loads block [r] of every rank's array into its spectrum.
struct PeerPtrs {
    float2 *p[PM_PEER_MAX];
};

// VEC = 2: two kx per thread (16-byte accesses; hc even), VEC = 1: the Nyquist plane (hc = 1)
template <bool PULL, int VEC>
__global__ void __launch_bounds__(256) k_slab_peer_copy(float2 *__restrict__ full, PeerPtrs peers,
                                                        size_t peer_off, int rank, int nzl, int n,
                                                        int h, int k0, int hc, int nyl)
{
    // full: [nzl][n][h] local spectrum; peers.p[s] + peer_off: [P*nzl][nyl][hc] on rank s
    const int hv = hc / VEC;
    const size_t total = (size_t)nzl * n * hv;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (size_t)gridDim.x * blockDim.x) {
        const int kc = (int)(i % hv) * VEC;
        const size_t r = i / hv;
        const int y = (int)(r % n), zl = (int)(r / n);
        const int s = y / nyl, yl = y - s * nyl;
        float2 *loc = full + ((size_t)zl * n + y) * h + k0 + kc;
        float2 *rem = peers.p[s] + peer_off + (((size_t)rank * nzl + zl) * nyl + yl) * hc + kc;
        if (VEC == 2) {
            if (PULL) *reinterpret_cast<float4 *>(loc) = *reinterpret_cast<const float4 *>(rem);
            else *reinterpret_cast<float4 *>(rem) = *reinterpret_cast<const float4 *>(loc);
        } else {
            if (PULL) *loc = *rem;
            else *rem = *loc;
        }
    }
}

// The same transposes by the COPY ENGINES (transport "peer", PM_PEER_DMA != 0): per destination rank one
// strided 3-D copy (rows of hc*8 bytes, nyl rows per plane, nzl planes) and, with chunk 0, one 2-D copy of
// the Nyquist plane.  The copy kernels above need SMs and shared nothing with the one-CTA-per-SM column
// kernels of the wide meshes, so "overlapping" them only interleaved them; DMA transfers run beside the y
// and z passes of the neighbouring chunks for real (NVSwitch gives one GPU its full 900 GB/s towards any
// single peer, so the per-peer copies simply follow each other on the stream).
template <bool PULL>
int slab_peer_copy_dma(pm_plan *p, int c, int C, cudaStream_t st)
{
    const int N = p->nc, H = N / 2;
    const int nzl = p->nzl, nyl = N / p->nranks, hc = H / C;
    const size_t main_n = (size_t)nzl * N * H;
    float2 *spec_main = p->spec, *spec_side = p->spec + main_n;
    for (int k = 0; k < p->nranks; ++k) {
        const int s = (p->rank + k) % p->nranks;          // start with the own block, then round the ring
        float2 *rem_main = p->peer_recv[s] + main_n / C * c;      // [P*nzl][nyl][hc] on rank s
        cudaMemcpy3DParms q;
        memset(&q, 0, sizeof(q));
        cudaPitchedPtr loc = make_cudaPitchedPtr(spec_main, (size_t)H * sizeof(float2), (size_t)H * sizeof(float2), (size_t)N);
        cudaPitchedPtr rem = make_cudaPitchedPtr(rem_main, (size_t)hc * sizeof(float2), (size_t)hc * sizeof(float2), (size_t)nyl);
        const cudaPos loc_pos = make_cudaPos((size_t)c * hc * sizeof(float2), (size_t)s * nyl, 0);
        const cudaPos rem_pos = make_cudaPos(0, 0, (size_t)p->rank * nzl);
        q.srcPtr = PULL ? rem : loc; q.srcPos = PULL ? rem_pos : loc_pos;
        q.dstPtr = PULL ? loc : rem; q.dstPos = PULL ? loc_pos : rem_pos;
        q.extent = make_cudaExtent((size_t)hc * sizeof(float2), (size_t)nyl, (size_t)nzl);
        q.kind = cudaMemcpyDefault;
        PM_CUDA(cudaMemcpy3DAsync(&q, st));
        if (c == 0) {
            // Nyquist plane: local [nzl][N], remote [P*nzl][nyl] behind the main array
            float2 *l2 = spec_side + (size_t)s * nyl;
            float2 *r2 = p->peer_recv[s] + main_n + (size_t)p->rank * nzl * nyl;
            if (PULL)
                PM_CUDA(cudaMemcpy2DAsync(l2, (size_t)N * sizeof(float2), r2, (size_t)nyl * sizeof(float2),
                                          (size_t)nyl * sizeof(float2), (size_t)nzl, cudaMemcpyDefault, st));
            else
                PM_CUDA(cudaMemcpy2DAsync(r2, (size_t)nyl * sizeof(float2), l2, (size_t)N * sizeof(float2),
                                          (size_t)nyl * sizeof(float2), (size_t)nzl, cudaMemcpyDefault, st));
        }
    }
    return PM_OK;
}

template <bool PULL>
int slab_peer_copy(pm_plan *p, int c, int C, cudaStream_t st)
{
    const int N = p->nc, H = N / 2;
    const int nzl = p->nzl, nyl = N / p->nranks, hc = H / C;
    if (p->peer_dma) return slab_peer_copy_dma<PULL>(p, c, C, st);
    if (hc % 2) return PM_ERR_UNSUPPORTED;
    PeerPtrs pp;
    for (int s = 0; s < PM_PEER_MAX; ++s) pp.p[s] = s < p->nranks ? p->peer_recv[s] : nullptr;
    const size_t main_n = (size_t)nzl * N * H;
    float2 *spec_main = p->spec, *spec_side = p->spec + main_n;
    const int grid = p->sm_count * 8;
    auto copy2 = k_slab_peer_copy<PULL, 2>;
    auto copy1 = k_slab_peer_copy<PULL, 1>;
    PM_LAUNCH(copy2, grid, 256, 0, st, spec_main, pp, main_n / C * c, p->rank, nzl, N,
              H, c * hc, hc, nyl);
    if (c == 0)
        PM_LAUNCH(copy1, p->sm_count * 2, 256, 0, st, spec_side, pp, main_n, p->rank, nzl,
                  N, 1, 0, 1, nyl);
    PM_CHECK_LAUNCH();
    return PM_OK;
}

// Flag words: rank r's word [slot][s] is written by rank s.  Signal: one thread per peer stores this
// step's epoch (the kernel boundary before it has already made the pushed data visible; the fence
// orders it for the system scope anyway).  Wait: one thread per peer polls its word; after ~2 s it
// gives up and counts a timeout instead of hanging the device.
struct PeerFlagPtrs {
    uint32_t *p[PM_PEER_MAX];
};

// How long a flag wait spins before it gives up, counts a timeout (reported by pm_slab_peer_timeouts and, with
// the peer-memory migration, raised as an error by every rank in the same step) and lets the stream go on.
// 10 s by default -- a peer whose host stalls (GC pause, synchronous snapshot I/O) must not turn into a
// silently wrong step; PM_PEER_TIMEOUT_MS overrides it at plan creation (pm_peer_timeout_init).
__device__ unsigned long long g_peer_timeout_ns = 10000000000ull;

__global__ void k_peer_signal_impl(PeerFlagPtrs peers, int nranks, int rank, int slot, uint32_t epoch)
{
    const int s = threadIdx.x;
    if (s >= nranks) return;
    __threadfence_system();
    volatile uint32_t *w = peers.p[s] + (size_t)slot * PM_PEER_MAX + rank;
    *w = epoch;
}

__global__ void k_peer_wait_impl(uint32_t *flags, int nranks, int slot, uint32_t epoch)
{
    const int s = threadIdx.x;
    if (s < nranks) {
        const volatile uint32_t *w = flags + (size_t)slot * PM_PEER_MAX + s;
        unsigned long long t0 = 0, now = 0;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        while ((int)(*w - epoch) < 0) {
            __nanosleep(100);
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            if (now - t0 > g_peer_timeout_ns) {
                atomicAdd(flags + (size_t)PM_PEER_SLOTS * PM_PEER_MAX, 1u);
                break;
            }
        }
    }
    __threadfence_system();
}

// Neighbour-only variants for the ghost planes: one flag word to one rank, wait for one rank.
__global__ void k_peer_signal_one(uint32_t *word, uint32_t epoch)
{
    __threadfence_system();
    *reinterpret_cast<volatile uint32_t *>(word) = epoch;
}

__global__ void k_peer_wait_one(uint32_t *flags, int slot, int src, uint32_t epoch)
{
    const volatile uint32_t *w = flags + (size_t)slot * PM_PEER_MAX + src;
    unsigned long long t0 = 0, now = 0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    while ((int)(*w - epoch) < 0) {
        __nanosleep(100);
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
        if (now - t0 > g_peer_timeout_ns) {
            atomicAdd(flags + (size_t)PM_PEER_SLOTS * PM_PEER_MAX, 1u);
            break;
        }
    }
    __threadfence_system();
}

__global__ void __launch_bounds__(256) k_peer_put(float4 *__restrict__ dst, const float4 *__restrict__ src, size_t n4)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x)
        dst[i] = src[i];
}

template <int N>
int slab_y_peer(pm_plan *p, int c, int C, bool fwd, cudaStream_t st)
{
    constexpr int H = N / 2;
    const int nzl = p->nzl, nyl = N / p->nranks, hc = H / C;
    const size_t smem_cols = ((size_t)N * kColsCN<N> + 2 * N) * sizeof(float2);
    ColArgs ca = slab_args<N>(p);
    ca.tpr = hc / kColsCN<N>;
    ca.kt0 = c * ca.tpr;
    PeerArgs pa;
    for (int s = 0; s < PM_PEER_MAX; ++s) pa.p[s] = s < p->nranks ? p->peer_recv[s] : nullptr;
    const size_t main_n = (size_t)nzl * N * H;
    pa.main_off = main_n / C * c;
    pa.side_off = main_n;
    pa.rank = p->rank;
    pa.nzl = nzl;
    pa.hc = hc;
    pa.nyl_shift = 0;
    while ((1 << pa.nyl_shift) < nyl) ++pa.nyl_shift;
    if ((1 << pa.nyl_shift) != nyl) return PM_ERR_UNSUPPORTED;
    if (fwd) {
        auto k = k_fft_cols_peer<N, COL_FWD>;
        PM_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_cols));
        PM_LAUNCH(k, nzl * ca.tpr, kThrC<N>, smem_cols, st, ca, pa);
    } else {
        auto k = k_fft_cols_peer<N, COL_INV>;
        PM_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_cols));
        PM_LAUNCH(k, nzl * ca.tpr, kThrC<N>, smem_cols, st, ca, pa);
    }
    PM_CHECK_LAUNCH();
    return PM_OK;
}

template <int N>
int slab_y_only(pm_plan *p, int c, int C, bool fwd, cudaStream_t st)
{
    constexpr int H = N / 2;
    const int nzl = p->nzl, hc = H / C;
    const size_t smem_cols = ((size_t)N * kColsCN<N> + 2 * N) * sizeof(float2);
    ColArgs ca = slab_args<N>(p);
    ca.tpr = hc / kColsCN<N>;
    ca.kt0 = c * ca.tpr;
    if (fwd) {
        auto cols_fwd = k_fft_cols<N, COL_FWD>;
        PM_CUDA(cudaFuncSetAttribute(cols_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_cols));
        PM_LAUNCH(cols_fwd, nzl * ca.tpr, kThrC<N>, smem_cols, st, ca);
    } else {
        auto cols_inv = k_fft_cols<N, COL_INV>;
        PM_CUDA(cudaFuncSetAttribute(cols_inv, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_cols));
        PM_LAUNCH(cols_inv, nzl * ca.tpr, kThrC<N>, smem_cols, st, ca);
    }
    PM_CHECK_LAUNCH();
    return PM_OK;
}

// ---- matter power spectrum (the reference has no estimator) --------------------
// After rows R2C + y forward + z forward the array holds rho_k at (zpos, ypos, kx) with zpos/ypos
// digit-reversed.  Each mode adds w*|rho_k|^2 to the spherical bin round(|k|) (integer frequency
// units, edges at half-integers), w = 2 for 0 < kx < N/2 (the Hermitian half that is not stored),
// 1 for kx = 0 and kx = N/2.  Block-level float64 partial sums, then one atomicAdd per bin.
template <int N>
__global__ void __launch_bounds__(256) k_pk_bin(const float2 *__restrict__ main,
                                                const float2 *__restrict__ side, int nbins,
                                                double *__restrict__ psum, double *__restrict__ pcnt)
{
    extern __shared__ double s_bins[];   // [2][nbins]
    constexpr int H = N / 2;
    for (int i = threadIdx.x; i < 2 * nbins; i += blockDim.x) s_bins[i] = 0.0;
    __syncthreads();
    const size_t total = (size_t)N * N * (H + 1);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (size_t)gridDim.x * blockDim.x) {
        const int kx = (int)(i % (H + 1));
        const size_t r = i / (H + 1);
        const int ypos = (int)(r % N), zpos = (int)(r / N);
        const float2 v = (kx < H) ? main[((size_t)zpos * N + ypos) * H + kx] : side[(size_t)zpos * N + ypos];
        int kz = digit_unrev<N>(zpos), ky = digit_unrev<N>(ypos);
        if (kz > N / 2) kz -= N;
        if (ky > N / 2) ky -= N;
        const double kk = sqrt((double)kz * kz + (double)ky * ky + (double)kx * kx);
        const int b = (int)floor(kk + 0.5);
        if (b >= 1 && b < nbins) {
            const double w = (kx > 0 && kx < H) ? 2.0 : 1.0;
            atomicAdd(&s_bins[b], w * ((double)v.x * v.x + (double)v.y * v.y));
            atomicAdd(&s_bins[nbins + b], w);
        }
    }
    __syncthreads();
    for (int b = threadIdx.x; b < nbins; b += blockDim.x) {
        if (s_bins[nbins + b] != 0.0) {
            atomicAdd(psum + b, s_bins[b]);
            atomicAdd(pcnt + b, s_bins[nbins + b]);
        }
    }
}

template <int N>
int pk_launch(pm_plan *p, const float *rho, int nbins, double *psum, double *pcnt, cudaStream_t st)
{
    constexpr int H = N / 2;
    const size_t smem_cols = ((size_t)N * kColsCN<N> + 2 * N) * sizeof(float2);
    const size_t smem_rows = ((size_t)H * kPitch + N) * sizeof(float2);
    auto rows_fwd = k_fft_rows<N, true>;
    auto cols_fwd = k_fft_cols<N, COL_FWD>;
    PM_CUDA(cudaFuncSetAttribute(rows_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_rows));
    PM_CUDA(cudaFuncSetAttribute(cols_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_cols));
    ColArgs ca;
    ca.main = p->spec;
    ca.side = p->spec + (size_t)N * N * H;
    ca.tw = p->tw; ca.sin2 = p->sin2; ca.sin2rev = p->sin2rev; ca.sin2y = p->sin2rev;
    ca.scale = 0.f; ca.scale_ptr = nullptr; ca.nyl = N; ca.y0 = 0;
    ca.tpr = H / kColsCN<N>; ca.kt0 = 0; ca.hw = H; ca.side_tiles = 1;
    PM_CUDA(cudaMemsetAsync(psum, 0, sizeof(double) * nbins, st));
    PM_CUDA(cudaMemsetAsync(pcnt, 0, sizeof(double) * nbins, st));
    PM_LAUNCH(rows_fwd, N * N / kCols, kThr<N>, smem_rows, st, reinterpret_cast<const float2 *>(rho),
              ca.main, (const float2 *)p->tw, (const float *)nullptr);   // P(k): the estimator wants the DC too
    ca.axis = 1;
    PM_LAUNCH(cols_fwd, N * ca.tpr, kThrC<N>, smem_cols, st, ca);
    ca.axis = 0;
    PM_LAUNCH(cols_fwd, N * ca.tpr + N / kColsCN<N>, kThrC<N>, smem_cols, st, ca);
    auto bin = k_pk_bin<N>;
    PM_LAUNCH(bin, p->sm_count * 4, 256, 2 * nbins * sizeof(double), st, (const float2 *)ca.main,
              (const float2 *)ca.side, nbins, psum, pcnt);
    PM_CHECK_LAUNCH();
    return PM_OK;
}

}  // namespace

int pm_peer_timeout_init()
{
    const char *e = getenv("PM_PEER_TIMEOUT_MS");
    if (!e || !*e) return PM_OK;
    const double ms = atof(e);
    if (!(ms >= 1.0) || ms > 3.6e6) return PM_ERR_INVALID;
    const unsigned long long ns = (unsigned long long)(ms * 1e6);
    return (int)cudaMemcpyToSymbol(g_peer_timeout_ns, &ns, sizeof(ns));
}

bool pm_fft_supported(int nc)
{
    return nc == 32 || nc == 64 || nc == 128 || nc == 256 || nc == 512 || nc == 1024 || nc == 2048;
}

// exp(-2 pi i m / N) in float64 -> float32, and the digit-reversed sin^2 table of the Green's kernel
int pm_k_fft_tables(pm_plan *p)
{
    const int n = p->nc;
    float2 *tw = (float2 *)malloc(sizeof(float2) * n);
    float *sr = (float *)malloc(sizeof(float) * n);
    if (!tw || !sr) return PM_ERR_NOMEM;
    for (int m = 0; m < n; ++m) {
        const double ang = -2.0 * M_PI * (double)m / (double)n;
        tw[m] = make_float2((float)cos(ang), (float)sin(ang));
    }
    for (int k = 0; k < n; ++k) {
        int pos = 0;
        switch (n) {
            case 32: pos = digit_rev<32>(k); break;
            case 64: pos = digit_rev<64>(k); break;
            case 128: pos = digit_rev<128>(k); break;
            case 256: pos = digit_rev<256>(k); break;
            case 512: pos = digit_rev<512>(k); break;
            case 1024: pos = digit_rev<1024>(k); break;
            case 2048: pos = digit_rev<2048>(k); break;
        }
        const double s = sin(M_PI * (double)k / (double)n);
        sr[pos] = (float)(s * s);
    }
    cudaError_t e = cudaMemcpy(p->tw, tw, sizeof(float2) * n, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(p->sin2rev, sr, sizeof(float) * n, cudaMemcpyHostToDevice);
    free(tw);
    free(sr);
    return (int)e;
}

// The digit-reversed copy of a caller-supplied sin^2 table (pm_plan_set_sin2_table).
int pm_k_sin2rev_install(pm_plan *p, const float *sin2_h)
{
    const int n = p->nc;
    if (!pm_fft_supported(n)) return PM_OK;
    float *sr = (float *)malloc(sizeof(float) * n);
    if (!sr) return PM_ERR_NOMEM;
    for (int k = 0; k < n; ++k) {
        int pos = 0;
        switch (n) {
            case 32: pos = digit_rev<32>(k); break;
            case 64: pos = digit_rev<64>(k); break;
            case 128: pos = digit_rev<128>(k); break;
            case 256: pos = digit_rev<256>(k); break;
            case 512: pos = digit_rev<512>(k); break;
            case 1024: pos = digit_rev<1024>(k); break;
            case 2048: pos = digit_rev<2048>(k); break;
        }
        sr[pos] = sin2_h[k];
    }
    const cudaError_t e = cudaMemcpy(p->sin2rev, sr, sizeof(float) * n, cudaMemcpyHostToDevice);
    free(sr);
    return (int)e;
}

int pm_k_poisson_own(pm_plan *p, const float *rho, double a, double omega_m0, float *phi, cudaStream_t st)
{
    switch (p->nc) {
        case 32: return poisson_launch<32>(p, rho, a, omega_m0, phi, st);
        case 64: return poisson_launch<64>(p, rho, a, omega_m0, phi, st);
        case 128: return poisson_launch<128>(p, rho, a, omega_m0, phi, st);
        case 256: return poisson_launch<256>(p, rho, a, omega_m0, phi, st);
        case 512: return poisson_launch<512>(p, rho, a, omega_m0, phi, st);
        case 1024: return poisson_launch<1024>(p, rho, a, omega_m0, phi, st);
        case 2048: return poisson_launch<2048>(p, rho, a, omega_m0, phi, st);
    }
    return PM_ERR_UNSUPPORTED;
}

#define PM_FFT_DISPATCH(fn, ...)                              \
    switch (p->nc) {                                          \
        case 32: return fn<32>(__VA_ARGS__);                  \
        case 64: return fn<64>(__VA_ARGS__);                  \
        case 128: return fn<128>(__VA_ARGS__);                \
        case 256: return fn<256>(__VA_ARGS__);                \
        case 512: return fn<512>(__VA_ARGS__);                \
        case 1024: return fn<1024>(__VA_ARGS__);              \
        case 2048: return fn<2048>(__VA_ARGS__);              \
    }                                                         \
    return PM_ERR_UNSUPPORTED

int pm_k_fft_slab_rows_fwd(pm_plan *p, const float *rho, cudaStream_t st)
{
    PM_FFT_DISPATCH(slab_rows_fwd, p, rho, st);
}

int pm_k_fft_slab_y_fwd_pack(pm_plan *p, int c, int C, float2 *send_main_c, float2 *send_side,
                             cudaStream_t st)
{
    PM_FFT_DISPATCH(slab_y_fwd_pack, p, c, C, send_main_c, send_side, st);
}

int pm_k_fft_slab_z_chunk(pm_plan *p, int c, int C, float2 *main_t_c, float2 *side_t, double a,
                          double omega_m0, cudaStream_t st)
{
    PM_FFT_DISPATCH(slab_z_chunk, p, c, C, main_t_c, side_t, a, omega_m0, st);
}

int pm_k_fft_slab_unpack_y_inv(pm_plan *p, int c, int C, const float2 *back_main_c,
                               const float2 *back_side, cudaStream_t st)
{
    PM_FFT_DISPATCH(slab_unpack_y_inv, p, c, C, back_main_c, back_side, st);
}

int pm_k_fft_slab_rows_inv(pm_plan *p, float *phi, cudaStream_t st)
{
    PM_FFT_DISPATCH(slab_rows_inv, p, phi, st);
}

int pm_k_fft_slab_y_fwd(pm_plan *p, int c, int C, cudaStream_t st)
{
    PM_FFT_DISPATCH(slab_y_only, p, c, C, true, st);
}

int pm_k_fft_slab_y_inv(pm_plan *p, int c, int C, cudaStream_t st)
{
    PM_FFT_DISPATCH(slab_y_only, p, c, C, false, st);
}

int pm_k_fft_slab_y_fwd_push(pm_plan *p, int c, int C, cudaStream_t st)
{
    PM_FFT_DISPATCH(slab_y_peer, p, c, C, true, st);
}

int pm_k_fft_slab_y_inv_pull(pm_plan *p, int c, int C, cudaStream_t st)
{
    PM_FFT_DISPATCH(slab_y_peer, p, c, C, false, st);
}

int pm_k_fft_slab_push(pm_plan *p, int c, int C, cudaStream_t st) { return slab_peer_copy<false>(p, c, C, st); }
int pm_k_fft_slab_pull(pm_plan *p, int c, int C, cudaStream_t st) { return slab_peer_copy<true>(p, c, C, st); }

int pm_k_peer_signal(pm_plan *p, int slot, uint32_t epoch, cudaStream_t st)
{
    PeerFlagPtrs pf;
    for (int s = 0; s < PM_PEER_MAX; ++s) pf.p[s] = s < p->nranks ? p->peer_flag_of[s] : nullptr;
    PM_LAUNCH(k_peer_signal_impl, 1, 32, 0, st, pf, p->nranks, p->rank, slot, epoch);
    PM_CHECK_LAUNCH();
    return PM_OK;
}

int pm_k_peer_wait(pm_plan *p, int slot, uint32_t epoch, cudaStream_t st)
{
    PM_LAUNCH(k_peer_wait_impl, 1, 32, 0, st, p->peer_flags, p->nranks, slot, epoch);
    PM_CHECK_LAUNCH();
    return PM_OK;
}

int pm_k_peer_put(pm_plan *p, float *dst, const float *src, size_t nfloat, cudaStream_t st)
{
    const size_t n4 = nfloat / 4;      // planes: nc*nc floats, nc a multiple of 16
    const size_t want = (n4 + 255) / 256;
    const int grid = (int)(want < (size_t)p->sm_count * 4 ? want : (size_t)p->sm_count * 4);
    PM_LAUNCH(k_peer_put, grid, 256, 0, st, reinterpret_cast<float4 *>(dst), reinterpret_cast<const float4 *>(src), n4);
    PM_CHECK_LAUNCH();
    return PM_OK;
}

int pm_k_peer_signal_to(pm_plan *p, int slot, int dst_rank, uint32_t epoch, cudaStream_t st)
{
    uint32_t *word = p->peer_flag_of[dst_rank] + (size_t)slot * PM_PEER_MAX + p->rank;
    PM_LAUNCH(k_peer_signal_one, 1, 1, 0, st, word, epoch);
    PM_CHECK_LAUNCH();
    return PM_OK;
}

int pm_k_peer_wait_from(pm_plan *p, int slot, int src_rank, uint32_t epoch, cudaStream_t st)
{
    PM_LAUNCH(k_peer_wait_one, 1, 1, 0, st, p->peer_flags, slot, src_rank, epoch);
    PM_CHECK_LAUNCH();
    return PM_OK;
}

int pm_fft_cols_per_tile(int nc) { return nc >= 1024 ? 8 : PM_FFT_COLS; }

int pm_k_power_spectrum(pm_plan *p, const float *rho, int nbins, double *psum, double *pcnt,
                        cudaStream_t st)
{
    PM_FFT_DISPATCH(pk_launch, p, rho, nbins, psum, pcnt, st);
}
