// pm_fft2.cuh -- two-stage, register-resident 1-D transforms for the 256/512/1024 meshes.
// Included by pm_fft.cu inside its anonymous namespace (uses its complex helpers, dft8, ColArgs,
// green_f32, pm_ld).
//
// The 8.8.8 kernels of pm_fft.cu cross shared memory twice per 512-point transform and
// synchronise after every radix-8 stage; the row pass is bound by shared-memory wavefronts and the
// column passes by their load -> sync -> stage -> sync -> stage -> store phase structure.  Here a
// transform of N = RA*RB points is two in-register DFTs with ONE shared-memory exchange between
// them (Cooley-Tukey, n = i + RB*r, k = q + RA*p):
//     stage A   for each i: DFT_RA over r, times W_N^(i*q)           thread (i, column)
//     exchange  (q, i) tile in shared memory, one barrier
//     stage B   for each q: DFT_RB over i -> X[q + RA*p]              thread (q, column)
// Each thread issues its RA (or RB) global loads back to back before the first butterfly, so the
// memory pipeline sees 16-32 independent 128-byte-coalesced requests per thread.  Output order is
// NATURAL on every axis (stores pick their row freely), so this path needs no digit-reversed tables.
// The inverse runs the two stages backwards with conjugated twiddles; the z pass does
// A -> B -> Green -> B^-1 -> A^-1 in one kernel (the spectrum crosses HBM once).
#pragma once

// exp(-2 pi i k / 32), k = 0..15 (cos, sin tables)
__device__ __forceinline__ constexpr float pm_c32(int k)
{
    constexpr float c[17] = {1.0f, 0.98078528040323044913f, 0.92387953251128675613f, 0.83146961230254523708f,
                             0.70710678118654752440f, 0.55557023301960222474f, 0.38268343236508977173f,
                             0.19509032201612826785f, 0.0f, -0.19509032201612826785f, -0.38268343236508977173f,
                             -0.55557023301960222474f, -0.70710678118654752440f, -0.83146961230254523708f,
                             -0.92387953251128675613f, -0.98078528040323044913f, -1.0f};
    return c[k];
}
__device__ __forceinline__ constexpr float pm_s32(int k) { return k <= 8 ? pm_c32(8 - k) : pm_c32(k - 8); }

// v * exp(S * 2 pi i * K / 32): compile-time twiddle with the trivial cases folded
template <int S, int K>
__device__ __forceinline__ float2 pm_tw32(float2 v)
{
    if constexpr (K == 0) return v;
    else if constexpr (K == 8) return make_float2(-S * v.y, S * v.x);   // * (S i)
    else {
        constexpr float c = pm_c32(K), s = S * pm_s32(K);
        return make_float2(v.x * c - v.y * s, v.x * s + v.y * c);
    }
}

template <int S>
__device__ __forceinline__ void dft16(float2 (&v)[16])
{
    float2 a[8], b[8];
#define PM_H16(r)                                      \
    a[r] = cadd(v[r], v[r + 8]);                       \
    b[r] = pm_tw32<S, 2 * r>(csub(v[r], v[r + 8]));
    PM_H16(0) PM_H16(1) PM_H16(2) PM_H16(3) PM_H16(4) PM_H16(5) PM_H16(6) PM_H16(7)
#undef PM_H16
    dft8<S>(a);
    dft8<S>(b);
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        v[2 * q] = a[q];
        v[2 * q + 1] = b[q];
    }
}

template <int S>
__device__ __forceinline__ void dft32(float2 (&v)[32])
{
    float2 a[16], b[16];
#define PM_H32(r)                                      \
    a[r] = cadd(v[r], v[r + 16]);                      \
    b[r] = pm_tw32<S, r>(csub(v[r], v[r + 16]));
    PM_H32(0) PM_H32(1) PM_H32(2) PM_H32(3) PM_H32(4) PM_H32(5) PM_H32(6) PM_H32(7)
    PM_H32(8) PM_H32(9) PM_H32(10) PM_H32(11) PM_H32(12) PM_H32(13) PM_H32(14) PM_H32(15)
#undef PM_H32
    dft16<S>(a);
    dft16<S>(b);
#pragma unroll
    for (int q = 0; q < 16; ++q) {
        v[2 * q] = a[q];
        v[2 * q + 1] = b[q];
    }
}

// natural order in and out, V[q] = sum_r v[r] exp(S 2 pi i r q / R)
template <int R, int S>
__device__ __forceinline__ void dftr(float2 (&v)[R])
{
    static_assert(R == 8 || R == 16 || R == 32, "register DFT sizes");
    if constexpr (R == 8) dft8<S>(v);
    else if constexpr (R == 16) dft16<S>(v);
    else dft32<S>(v);
}

constexpr int kThr2 = 256;   // threads of every two-stage CTA

// N = RA * RB for the column transforms; C columns per tile (the same tile width as pm_fft.cu)
template <int N> struct ColFac;
template <> struct ColFac<256>  { static constexpr int RA = 16, RB = 16; };
template <> struct ColFac<512>  { static constexpr int RA = 32, RB = 16; };
template <> struct ColFac<1024> { static constexpr int RA = 32, RB = 32; };
// N/2 = RA * RB for the packed real row transforms; RT = 256 / RA rows per tile
template <int N> struct RowFac;
template <> struct RowFac<256>  { static constexpr int RA = 16, RB = 8; };
template <> struct RowFac<512>  { static constexpr int RA = 16, RB = 16; };
template <> struct RowFac<1024> { static constexpr int RA = 32, RB = 16; };
template <> struct RowFac<2048> { static constexpr int RA = 32, RB = 32; };   // rows only (slab path): 32 x 32 = 1024 complex points
template <int N>
constexpr int kRowsPT2 = kThr2 / RowFac<N>::RA;
template <int N>
constexpr bool kHasV2 = (N == 256 || N == 512 || N == 1024);
// the row kernel alone also exists for 2048 points (two 32-point register DFTs; one CTA of 256 threads per SM
// pair of row tiles would need > 128 registers in the inverse, so it runs at one CTA per 255-register budget)
template <int N>
constexpr bool kHasRows2 = kHasV2<N> || N == 2048;

// ---- column transform of a tile: N points x CC columns ----------------------------------------
// ld(pos, c) / st(pos, c, v) address the global tile; green(k, c) is the real factor applied to
// mode k of column c (COL_FUSED only).  Shared tile: [(q*RB + i)][CC] float2.
template <int N, int MODE, int CC, class Ld, class St, class Green>
__device__ __forceinline__ void fft2_col_pass(float2 *s_tile, const float2 *s_tw, Ld ld, St st, Green green)
{
    constexpr int RA = ColFac<N>::RA, RB = ColFac<N>::RB;
    const int tid = threadIdx.x;
    if constexpr (MODE == COL_FWD || MODE == COL_FUSED) {
        // stage A: items (i, c)
#pragma unroll 1
        for (int item = tid; item < RB * CC; item += kThr2) {
            const int i = item / CC, c = item % CC;
            float2 v[RA];
#pragma unroll
            for (int r = 0; r < RA; ++r) v[r] = ld(i + RB * r, c);
            dftr<RA, -1>(v);
#pragma unroll
            for (int q = 1; q < RA; ++q) v[q] = cmul(v[q], s_tw[i * q]);
#pragma unroll
            for (int q = 0; q < RA; ++q) s_tile[(q * RB + i) * CC + c] = v[q];
        }
        __syncthreads();
        // stage B: items (q, c)
#pragma unroll 1
        for (int item = tid; item < RA * CC; item += kThr2) {
            const int q = item / CC, c = item % CC;
            float2 v[RB];
#pragma unroll
            for (int i = 0; i < RB; ++i) v[i] = s_tile[(q * RB + i) * CC + c];
            dftr<RB, -1>(v);
            if constexpr (MODE == COL_FWD) {
#pragma unroll
                for (int p = 0; p < RB; ++p) st(q + RA * p, c, v[p]);
            } else {
#pragma unroll
                for (int p = 0; p < RB; ++p) {
                    const float g = green(q + RA * p, c);
                    v[p].x *= g;
                    v[p].y *= g;
                }
                dftr<RB, +1>(v);   // inverse over p -> u_q[i]
#pragma unroll
                for (int i = 1; i < RB; ++i) v[i] = cmulc(v[i], s_tw[i * q]);
#pragma unroll
                for (int i = 0; i < RB; ++i) s_tile[(q * RB + i) * CC + c] = v[i];   // own slots
            }
        }
    }
    if constexpr (MODE == COL_INV) {
        // stage B^-1: items (q, c), straight from global
#pragma unroll 1
        for (int item = tid; item < RA * CC; item += kThr2) {
            const int q = item / CC, c = item % CC;
            float2 v[RB];
#pragma unroll
            for (int p = 0; p < RB; ++p) v[p] = ld(q + RA * p, c);
            dftr<RB, +1>(v);
#pragma unroll
            for (int i = 1; i < RB; ++i) v[i] = cmulc(v[i], s_tw[i * q]);
#pragma unroll
            for (int i = 0; i < RB; ++i) s_tile[(q * RB + i) * CC + c] = v[i];
        }
    }
    if constexpr (MODE == COL_INV || MODE == COL_FUSED) {
        __syncthreads();
        // stage A^-1: items (i, c)
#pragma unroll 1
        for (int item = tid; item < RB * CC; item += kThr2) {
            const int i = item / CC, c = item % CC;
            float2 v[RA];
#pragma unroll
            for (int q = 0; q < RA; ++q) v[q] = s_tile[(q * RB + i) * CC + c];
            dftr<RA, +1>(v);
#pragma unroll
            for (int r = 0; r < RA; ++r) st(i + RB * r, c, v[r]);
        }
    }
}

struct NoGreen {
    __device__ __forceinline__ float operator()(int, int) const { return 1.0f; }
};

// One column tile, same tile numbering and ColArgs as fft_cols_tile; NATURAL order along the
// transformed axis (the Green's factor reads a.sin2 for all three axes).
template <int N, int MODE, bool CG>
__device__ __forceinline__ void fft2_cols_tile(const ColArgs &a, const int t, float2 *s_tile,
                                               const float2 *s_tw)
{
    constexpr int H = N / 2;
    constexpr int C = kColsCN<N>;
    const int TPR = a.tpr;
    if (a.axis == 1) {
        // y pass: plane z, kx tile kt; tile 0 also carries the Nyquist-in-x column split off the
        // packed slot (forward) / merged back into it (inverse)
        const int z = t / TPR, kt = a.kt0 + t % TPR;
        float2 *g = a.main + (size_t)z * N * H + kt * C;
        const bool extra = (kt == 0);
        auto ld = [&](int pos, int c) -> float2 {
            float2 v = pm_ld<CG>(g + (size_t)pos * H + c);
            if (MODE == COL_FWD && extra && c == 0) v.y = 0.0f;   // packed slot: real part = DC column
            return v;
        };
        auto st = [&](int pos, int c, float2 v) { g[(size_t)pos * H + c] = v; };
        float2 *gx = a.side + (size_t)z * N;
        if (MODE == COL_FWD && extra) {
            // the Nyquist parts of the packed slots, before the main pass overwrites them
            auto ldx = [&](int pos, int) -> float2 {
                return make_float2(pm_ld<CG>(g + (size_t)pos * H).y, 0.0f);
            };
            auto stx = [&](int pos, int, float2 v) { gx[pos] = v; };
            fft2_col_pass<N, MODE, 1>(s_tile, s_tw, ldx, stx, NoGreen());
            __syncthreads();
        }
        fft2_col_pass<N, MODE, C>(s_tile, s_tw, ld, st, NoGreen());
        if (MODE == COL_INV && extra) {
            __syncthreads();   // the main pass stored column 0 (its .y is rounding noise)
            auto ldx = [&](int pos, int) -> float2 { return pm_ld<CG>(gx + pos); };
            auto stx = [&](int pos, int, float2 v) {
                reinterpret_cast<float *>(g + (size_t)pos * H)[1] = v.x;   // packed slot .y = Re Nyquist
            };
            fft2_col_pass<N, MODE, 1>(s_tile, s_tw, ldx, stx, NoGreen());
        }
    } else {
        // z pass (COL_FUSED, or COL_FWD for spectra): [N z][nyl][hw] main tiles, then side tiles
        float2 *g;
        size_t gs;
        bool side_tile = false;
        int col0, yl = 0;
        if (t < a.nyl * TPR) {
            yl = t / TPR;
            const int ktl = t % TPR;
            g = a.main + (size_t)yl * a.hw + ktl * C;
            gs = (size_t)a.nyl * a.hw;
            col0 = (a.kt0 + ktl) * C;
        } else {
            side_tile = true;
            const int yt = t - a.nyl * TPR;
            g = a.side + yt * C;
            gs = a.nyl;
            col0 = a.y0 + yt * C;
        }
        auto ld = [&](int pos, int c) -> float2 { return pm_ld<CG>(g + (size_t)pos * gs + c); };
        auto st = [&](int pos, int c, float2 v) { g[(size_t)pos * gs + c] = v; };
        const float sy_fixed = side_tile ? 0.0f : __ldg(a.sin2 + a.y0 + yl);
        const float sx_side = __ldg(a.sin2 + H);
        auto green = [&](int k, int c) -> float {
            const float sz = __ldg(a.sin2 + k);
            const float sy = side_tile ? __ldg(a.sin2 + col0 + c) : sy_fixed;
            const float sx = side_tile ? sx_side : __ldg(a.sin2 + col0 + c);
            return a.scale * green_f32(sz, sy, sx);
        };
        fft2_col_pass<N, MODE, C>(s_tile, s_tw, ld, st, green);
    }
}

// ---- row transform: RT x-rows per tile, packed real FFT of N points as H = N/2 complex ----------
// forward:  z[n] = x[2n] + i x[2n+1]  ->  Z = DFT_H(z)  ->  X[k] = (A - i W_N^k B)/2,
//           A = Z[k] + conj(Z[H-k]), B = Z[k] - conj(Z[H-k]);  slot 0 = (DC, Nyquist)
// Thread (row, q) of stage B holds Z[q + RA*p]; the partner Z[H-k] = Z[(RA-q) + RA*(RB-1-p)] sits in
// lane RA-q of the same row at register RB-1-p, so the split is a register shuffle, not a trip
// through shared memory (q = 0 pairs with itself at register RB-p).
// Shared tile: [(row*RA + q)][RB + 1] float2 (the pad keeps stage B's stride-RB reads conflict-free).
template <int N, bool FWD, bool CG>
__device__ __forceinline__ void fft2_rows_tile(const float2 *__restrict__ in, float2 *__restrict__ out,
                                               const size_t row0, float2 *s_tile, const float2 *s_tw,
                                               const float mean = 0.f)
{
    constexpr int H = N / 2;
    constexpr int RA = RowFac<N>::RA, RB = RowFac<N>::RB, RT = kRowsPT2<N>, P = RB + 1;
    static_assert(RA * RB == H && RA * RT == kThr2, "row factorisation");
    const int tid = threadIdx.x;
    const unsigned full = 0xffffffffu;
    // stage-B identity of this thread
    const int rowb = tid / RA, q = tid % RA;
    const int partner = (threadIdx.x & 31 & ~(RA - 1)) | ((RA - q) & (RA - 1));
    if constexpr (FWD) {
#pragma unroll 1
        for (int item = tid; item < RB * RT; item += kThr2) {
            const int row = item / RB, i = item % RB;
            const float2 *src = in + (row0 + row) * H;
            float2 v[RA];
#pragma unroll
            for (int r = 0; r < RA; ++r) {
                v[r] = src[i + RB * r];   // rho: written by an earlier launch
                v[r].x -= mean;           // rho - <rho> (see fft_rows_tile)
                v[r].y -= mean;
            }
            dftr<RA, -1>(v);
#pragma unroll
            for (int qq = 1; qq < RA; ++qq) v[qq] = cmul(v[qq], s_tw[2 * i * qq]);   // W_H^(i q)
#pragma unroll
            for (int qq = 0; qq < RA; ++qq) s_tile[(row * RA + qq) * P + i] = v[qq];
        }
        __syncthreads();
        float2 v[RB];
#pragma unroll
        for (int i = 0; i < RB; ++i) v[i] = s_tile[(rowb * RA + q) * P + i];
        dftr<RB, -1>(v);   // v[p] = Z[q + RA p]
        float2 *dst = out + (row0 + rowb) * H;
#pragma unroll
        for (int p = 0; p < RB; ++p) {
            // partner Z[H - k]
            float2 zm;
            zm.x = __shfl_sync(full, v[RB - 1 - p].x, partner);
            zm.y = __shfl_sync(full, v[RB - 1 - p].y, partner);
            if (q == 0) zm = v[(RB - p) % RB];
            const int k = q + RA * p;
            const float2 zk = v[p];
            float2 X;
            if (k == 0) {
                X = make_float2(zk.x + zk.y, zk.x - zk.y);   // packed (DC, Nyquist)
            } else {
                const float2 A = make_float2(zk.x + zm.x, zk.y - zm.y);
                const float2 B = make_float2(zk.x - zm.x, zk.y + zm.y);
                const float2 t = cmul(s_tw[k], B);
                X = make_float2(0.5f * (A.x + t.y), 0.5f * (A.y - t.x));
            }
            dst[k] = X;
        }
    } else {
        const float2 *src = in + (row0 + rowb) * H;
        float2 v[RB];
#pragma unroll
        for (int p = 0; p < RB; ++p) v[p] = pm_ld<CG>(src + q + RA * p);
        float2 z[RB];
#pragma unroll
        for (int p = 0; p < RB; ++p) {
            float2 bm;
            bm.x = __shfl_sync(full, v[RB - 1 - p].x, partner);
            bm.y = __shfl_sync(full, v[RB - 1 - p].y, partner);
            if (q == 0) bm = v[(RB - p) % RB];
            const int k = q + RA * p;
            const float2 A = v[p];
            if (k == 0) {
                z[p] = make_float2(A.x + A.y, A.x - A.y);   // (DC + Nyq) + i (DC - Nyq)
            } else {
                const float2 Pp = make_float2(A.x + bm.x, A.y - bm.y);   // A + conj(B)
                const float2 Q = make_float2(A.x - bm.x, A.y + bm.y);    // A - conj(B)
                const float2 t = cmulc(Q, s_tw[k]);                      // conj(w^k) * Q
                z[p] = make_float2(Pp.x - t.y, Pp.y + t.x);              // P + i t
            }
        }
        dftr<RB, +1>(z);   // inverse over p -> u_q[i]
#pragma unroll
        for (int i = 1; i < RB; ++i) z[i] = cmulc(z[i], s_tw[2 * i * q]);
#pragma unroll
        for (int i = 0; i < RB; ++i) s_tile[(rowb * RA + q) * P + i] = z[i];
        __syncthreads();
#pragma unroll 1
        for (int item = tid; item < RB * RT; item += kThr2) {
            const int row = item / RB, i = item % RB;
            float2 w[RA];
#pragma unroll
            for (int qq = 0; qq < RA; ++qq) w[qq] = s_tile[(row * RA + qq) * P + i];
            dftr<RA, +1>(w);
            float2 *dst = out + (row0 + row) * H;
#pragma unroll
            for (int r = 0; r < RA; ++r) dst[i + RB * r] = w[r];
        }
    }
}
