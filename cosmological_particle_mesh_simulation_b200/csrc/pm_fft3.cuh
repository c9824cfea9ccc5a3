// pm_fft3.cuh -- software-pipelined column passes: the two-stage transforms of pm_fft2.cuh fed by
// an asynchronous-copy ring.  Included by pm_fft.cu after pm_fft2.cuh.
//
// The column kernels of pm_fft.cu / pm_fft2.cuh are phase-structured: a CTA loads its tile, works
// on it, stores it, and only the other one or two CTAs resident on the SM cover its load latency.
// ncu (profiles/r01_notes.md) shows them waiting on global loads (long scoreboard, LG throttle)
// at half of the DRAM rate a copy with the same access pattern reaches.  Here ONE persistent CTA
// per SM walks its tiles with a ring of NBUF shared-memory buffers:
//     cp.async (16 bytes per thread per request, L2 only) tile k+NBUF-1  ->  buf[(k+NBUF-1) % NBUF]
//     wait for tile k, transform it IN PLACE in buf[k % NBUF], store the results from registers
// so NBUF-1 whole tiles (64 KB each at 512 points x 16 columns) are in flight per SM while the
// butterflies run, and the loads hold no registers.  The transforms are those of pm_fft2.cuh
// (same factorisation, twiddles and operation order: results are bit-identical to it).
//
// In-place slot maps (row = point index inside the [N][C] tile):
//   forward   stage A item (i, c) reads rows i + RB*r and writes V_q to row q*RB + i -- the same
//             row set; stage B item (q, c) reads rows q*RB + i and stores X[q + RA*p] to global.
//   inverse   stage B^-1 item (q, c) reads rows q + RA*p and writes u_i to row q + RA*i -- the
//             same row set; stage A^-1 item (i, c) reads rows q + RA*i and stores x[i + RB*r].
// A row is C float2 = 128 bytes = all 32 banks, and a warp covers two rows: no bank conflicts.
#pragma once

__device__ __forceinline__ void pm_cp_async16(void *smem, const void *gmem)
{
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void pm_cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int K>
__device__ __forceinline__ void pm_cp_async_wait()
{
    asm volatile("cp.async.wait_group %0;" ::"n"(K) : "memory");
}

// geometry of y-pass tile t: plane z, kx tile kt
template <int N>
struct Tile3 {
    float2 *g;      // first element
    size_t gs;      // stride between successive points (float2)
    bool extra;     // kx tile 0: carries the Nyquist-in-x column in the .y of its packed slot
    float2 *gx;     // that column's home: side[z][.]
};

template <int N>
__device__ __forceinline__ Tile3<N> fft3_tile(const ColArgs &a, int t)
{
    constexpr int H = N / 2, C = kColsCN<N>;
    Tile3<N> T;
    const int z = t / a.tpr, kt = a.kt0 + t % a.tpr;
    T.g = a.main + (size_t)z * N * H + kt * C;
    T.gs = H;
    T.extra = (kt == 0);
    T.gx = a.side + (size_t)z * N;
    return T;
}

// all threads: request tile T into buf ([N][C] float2, then the [N] extra column), 16 bytes per request
template <int N, int MODE, int NT>
__device__ __forceinline__ void fft3_request(const Tile3<N> &T, float2 *buf)
{
    constexpr int C = kColsCN<N>;
    constexpr int PARTS = C / 2;              // 16-byte pieces per row
    constexpr int TOTAL = N * PARTS;
    static_assert(TOTAL % NT == 0, "whole rounds");
#pragma unroll
    for (int k = 0; k < TOTAL / NT; ++k) {
        const int idx = k * NT + threadIdx.x;
        const int row = idx / PARTS, part = idx % PARTS;
        pm_cp_async16(buf + row * C + 2 * part, T.g + (size_t)row * T.gs + 2 * part);
    }
    if (MODE == COL_INV && T.extra) {
        // the Nyquist-in-x column of this plane (contiguous), for the inverse of the packed slot
        for (int idx = threadIdx.x; idx < N / 2; idx += NT) pm_cp_async16(buf + N * C + 2 * idx, T.gx + 2 * idx);
    }
}

template <int N, int MODE, int NT>
__device__ __forceinline__ void fft3_transform(const Tile3<N> &T, float2 *b, const float2 *s_tw, float2 *s_x,
                                               float *s_y)
{
    constexpr int C = kColsCN<N>;
    constexpr int RA = ColFac<N>::RA, RB = ColFac<N>::RB;
    static_assert(NT == kThr2, "the one-column helper passes are written for kThr2 threads");
    const int tid = threadIdx.x;
    if constexpr (MODE == COL_FWD) {
        if (T.extra) {
            // the Nyquist parts of the packed slots -> their own column (side plane)
            auto ldx = [&](int pos, int) -> float2 { return make_float2(b[pos * C].y, 0.0f); };
            auto stx = [&](int pos, int, float2 v) { T.gx[pos] = v; };
            fft2_col_pass<N, COL_FWD, 1>(s_x, s_tw, ldx, stx, NoGreen());
            __syncthreads();
        }
#pragma unroll 1
        for (int item = tid; item < RB * C; item += NT) {
            const int i = item / C, c = item % C;
            float2 v[RA];
#pragma unroll
            for (int r = 0; r < RA; ++r) v[r] = b[(i + RB * r) * C + c];
            if (T.extra && c == 0) {
#pragma unroll
                for (int r = 0; r < RA; ++r) v[r].y = 0.0f;   // packed slot: real part = DC column
            }
            dftr<RA, -1>(v);
#pragma unroll
            for (int q = 1; q < RA; ++q) v[q] = cmul(v[q], s_tw[i * q]);
#pragma unroll
            for (int q = 0; q < RA; ++q) b[(q * RB + i) * C + c] = v[q];
        }
        __syncthreads();
#pragma unroll 1
        for (int item = tid; item < RA * C; item += NT) {
            const int q = item / C, c = item % C;
            float2 v[RB];
#pragma unroll
            for (int i = 0; i < RB; ++i) v[i] = b[(q * RB + i) * C + c];
            dftr<RB, -1>(v);
#pragma unroll
            for (int p = 0; p < RB; ++p) T.g[(size_t)(q + RA * p) * T.gs + c] = v[p];
        }
    } else {
        static_assert(MODE == COL_INV, "y passes only");
        if (T.extra) {
            // inverse of the Nyquist column first; its real part becomes the .y of the packed slots
            const float2 *xc = b + N * C;
            auto ldx = [&](int pos, int) -> float2 { return xc[pos]; };
            auto stx = [&](int pos, int, float2 v) { s_y[pos] = v.x; };
            fft2_col_pass<N, COL_INV, 1>(s_x, s_tw, ldx, stx, NoGreen());
            __syncthreads();
        }
#pragma unroll 1
        for (int item = tid; item < RA * C; item += NT) {
            const int q = item / C, c = item % C;
            float2 v[RB];
#pragma unroll
            for (int p = 0; p < RB; ++p) v[p] = b[(q + RA * p) * C + c];
            dftr<RB, +1>(v);
#pragma unroll
            for (int i = 1; i < RB; ++i) v[i] = cmulc(v[i], s_tw[i * q]);
#pragma unroll
            for (int i = 0; i < RB; ++i) b[(q + RA * i) * C + c] = v[i];
        }
        __syncthreads();
#pragma unroll 1
        for (int item = tid; item < RB * C; item += NT) {
            const int i = item / C, c = item % C;
            float2 v[RA];
#pragma unroll
            for (int q = 0; q < RA; ++q) v[q] = b[(q + RA * i) * C + c];
            dftr<RA, +1>(v);
            if (T.extra && c == 0) {
#pragma unroll
                for (int r = 0; r < RA; ++r) v[r].y = s_y[i + RB * r];
            }
#pragma unroll
            for (int r = 0; r < RA; ++r) T.g[(size_t)(i + RB * r) * T.gs + c] = v[r];
        }
    }
}

// shared memory: s_tw[N] | NBUF x (tile[N][C] + extra column [N]) | s_x[N] | s_y[N] (floats)
template <int N>
constexpr int kSlot3 = N * kColsCN<N> + N;
template <int N, int NBUF>
constexpr size_t kSmem3 = ((size_t)N + (size_t)NBUF * kSlot3<N> + N) * sizeof(float2) + (size_t)N * sizeof(float);

template <int N, int MODE, int NBUF, int NT>
__global__ void __launch_bounds__(NT, 1) k_fft3_cols(ColArgs a, int ntiles)
{
    extern __shared__ float2 s_dyn[];
    float2 *s_tw = s_dyn;
    float2 *s_buf = s_dyn + N;
    float2 *s_x = s_buf + (size_t)NBUF * kSlot3<N>;
    float *s_y = reinterpret_cast<float *>(s_x + N);
    for (int m = threadIdx.x; m < N; m += NT) s_tw[m] = a.tw[m];
    const int t0 = blockIdx.x, dt = gridDim.x;
    // prologue: NBUF-1 tiles on their way
#pragma unroll
    for (int s = 0; s < NBUF - 1; ++s) {
        const int t = t0 + s * dt;
        if (t < ntiles) fft3_request<N, MODE, NT>(fft3_tile<N>(a, t), s_buf + (size_t)s * kSlot3<N>);
        pm_cp_async_commit();
    }
    int k = 0;
    for (int t = t0; t < ntiles; t += dt, ++k) {
        // buf[(k-1) % NBUF] was released by the barrier that ended iteration k-1
        const int tn = t + (NBUF - 1) * dt;
        if (tn < ntiles)
            fft3_request<N, MODE, NT>(fft3_tile<N>(a, tn), s_buf + (size_t)((k + NBUF - 1) % NBUF) * kSlot3<N>);
        pm_cp_async_commit();            // one group per iteration, empty at the tail
        pm_cp_async_wait<NBUF - 1>();    // this thread's requests for tile t have landed ...
        __syncthreads();                 // ... and everybody else's
        fft3_transform<N, MODE, NT>(fft3_tile<N>(a, t), s_buf + (size_t)(k % NBUF) * kSlot3<N>, s_tw, s_x, s_y);
        __syncthreads();
    }
}
