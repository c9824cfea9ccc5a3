// pm_sort.cu -- ordering the particle list by cell key (src/density.py:19-21,37: key = (z_c*Nc + y_c)*Nc + x_c).
//
// Two ways to produce (keys_sorted, order_sorted), with bit-identical results:
//
//   full         stable LSD radix sort of (key, slot) -- cub::DeviceRadixSort, 4 passes for 27-bit keys.
//
//   incremental  the resident state is stored in the cell order of the PREVIOUS step, and a particle
//                drifts much less than a cell per step, so most entries still carry the key they
//                were sorted by ("stayers": an already sorted subsequence) and only the "movers"
//                need sorting.  Per step:
//                  1. k_mover_count      movers per 2048-entry tile            (read keys, old keys)
//                  2. k_mover_scan       exclusive scan of the tile counts; total -> host (4 bytes)
//                  3. k_mover_partition  stayers -> A, movers -> B as 64-bit (key << 32 | slot)
//                  4. cub radix sort of B on the key bits (stable: slots stay ascending per key)
//                  5. k_merge_splits     merge-path split of every 2048-entry output tile
//                  6. k_merge_tiles      merge A and B by (key, slot) -> keys_sorted, order_sorted,
//                                        and the row table of the deposit (row_start) on the way
//                Because both inputs are ordered by the composite (key, slot), the merge reproduces
//                exactly the order a stable sort of all entries gives.  The host reads the mover
//                count (one 4-byte copy) to size the sort of B; if more than 40 % of the entries
//                moved (or there is no previous order) it runs the full sort instead.
//
// There is no reference counterpart: the reference scatters in particle-index order
// (src/density.py:17) and never sorts.  The order only fixes the float32 summation tree of the
// deposit and the memory locality of the deposit and gather kernels.
#include <cub/device/device_radix_sort.cuh>

#include <stdlib.h>
#include <string.h>

#include "pm_internal.cuh"

namespace {

constexpr int kTile = PM_SORT_TILE;   // entries per CTA
constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kRounds = kTile / kThreads;   // 32-entry rounds per warp
static_assert(kTile == kThreads * kRounds, "tile shape");

// Entry i is a mover when its key changed since the previous sort or it has no previous key
// (slab mode: arrivals appended behind the n_old entries the previous sort placed).
__device__ __forceinline__ bool pm_is_mover(uint32_t knew, uint32_t kold, int64_t i, int64_t n_old)
{
    return i >= n_old || knew != kold;
}

// Each warp owns 256 consecutive entries of the tile and visits them 32 at a time, so ranks
// inside the warp come from ballots and every global access is a contiguous 128-byte run.
__global__ void __launch_bounds__(kThreads) k_mover_count(const uint32_t *__restrict__ keys,
                                                          const uint32_t *__restrict__ keys_old,
                                                          int64_t n, int64_t n_old,
                                                          uint32_t *__restrict__ tile_cnt)
{
    __shared__ uint32_t s_w[kWarps];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t base = (int64_t)blockIdx.x * kTile + warp * (kRounds * 32);
    uint32_t cnt = 0;
#pragma unroll
    for (int k = 0; k < kRounds; ++k) {
        const int64_t i = base + k * 32 + lane;
        bool m = false;
        if (i < n) {
            const uint32_t kn = keys[i];
            const uint32_t ko = (i < n_old) ? keys_old[i] : 0u;
            m = pm_is_mover(kn, ko, i, n_old);
        }
        cnt += __popc(__ballot_sync(0xffffffffu, m));
    }
    if (lane == 0) s_w[warp] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) t += s_w[w];
        tile_cnt[blockIdx.x] = t;
    }
}

// In-place exclusive scan of cnt[0..nt); cnt[nt] = total.  One CTA; each thread owns a contiguous
// chunk of tiles.
__global__ void __launch_bounds__(1024) k_mover_scan(uint32_t *__restrict__ cnt, int nt)
{
    __shared__ uint32_t s_w[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int per = (nt + 1023) / 1024;
    const int lo = min(tid * per, nt), hi = min(lo + per, nt);
    uint32_t sum = 0;
    for (int t = lo; t < hi; ++t) sum += cnt[t];
    uint32_t inc = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t o = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += o;
    }
    if (lane == 31) s_w[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = s_w[lane];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t o = __shfl_up_sync(0xffffffffu, w, d);
            if (lane >= d) w += o;
        }
        s_w[lane] = w;   // inclusive over warps
    }
    __syncthreads();
    uint32_t run = inc - sum + (warp ? s_w[warp - 1] : 0u);   // exclusive prefix of this thread's chunk
    for (int t = lo; t < hi; ++t) {
        const uint32_t c = cnt[t];
        cnt[t] = run;
        run += c;
    }
    if (tid == 1023) cnt[nt] = s_w[31];
}

__global__ void __launch_bounds__(kThreads) k_mover_partition(const uint32_t *__restrict__ keys,
                                                              const uint32_t *__restrict__ keys_old,
                                                              int64_t n, int64_t n_old,
                                                              const uint32_t *__restrict__ tile_off,
                                                              uint64_t *__restrict__ a64,
                                                              uint64_t *__restrict__ b64)
{
    __shared__ uint32_t s_w[kWarps];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t lt = (1u << lane) - 1u;
    const int64_t base = (int64_t)blockIdx.x * kTile + warp * (kRounds * 32);
    uint32_t kn[kRounds];
    uint32_t mball[kRounds];
    uint32_t cnt = 0;
#pragma unroll
    for (int k = 0; k < kRounds; ++k) {
        const int64_t i = base + k * 32 + lane;
        bool m = false;
        kn[k] = 0;
        if (i < n) {
            kn[k] = keys[i];
            const uint32_t ko = (i < n_old) ? keys_old[i] : 0u;
            m = pm_is_mover(kn[k], ko, i, n_old);
        }
        mball[k] = __ballot_sync(0xffffffffu, m);
        cnt += __popc(mball[k]);
    }
    if (lane == 0) s_w[warp] = cnt;
    __syncthreads();
    uint32_t before = tile_off[blockIdx.x];   // movers ahead of this warp's first entry
    for (int w = 0; w < warp; ++w) before += s_w[w];
#pragma unroll
    for (int k = 0; k < kRounds; ++k) {
        const int64_t i = base + k * 32 + lane;
        const uint32_t ahead = before + __popc(mball[k] & lt);   // movers ahead of entry i
        if (i < n) {
            const uint64_t v = ((uint64_t)kn[k] << 32) | (uint64_t)(uint32_t)i;
            if ((mball[k] >> lane) & 1u) b64[ahead] = v;
            else a64[i - ahead] = v;
        }
        before += __popc(mball[k]);
    }
}

// Merge path: number of A elements among the first d outputs of merge(A, B); no ties (the slot
// part makes every composite unique).
__device__ __forceinline__ uint32_t pm_merge_split(const uint64_t *__restrict__ a, uint32_t na,
                                                   const uint64_t *__restrict__ b, uint32_t nb,
                                                   uint32_t d)
{
    uint32_t lo = d > nb ? d - nb : 0u, hi = d < na ? d : na;
    while (lo < hi) {
        const uint32_t mid = lo + ((hi - lo) >> 1);
        if (a[mid] < b[d - 1 - mid]) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}

__global__ void __launch_bounds__(256) k_merge_splits(const uint64_t *__restrict__ a,
                                                      const uint64_t *__restrict__ b,
                                                      uint32_t nb, uint32_t n, int nt,
                                                      uint32_t *__restrict__ split)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t > nt) return;
    const uint32_t na = n - nb;
    const uint64_t d64 = (uint64_t)t * kTile;
    const uint32_t d = d64 < n ? (uint32_t)d64 : n;
    split[t] = pm_merge_split(a, na, b, nb, d);
}

// row_start[r] = first sorted entry whose mesh row segment (key / xseg) is >= r, r in [0, nrows];
// entries past the last real row (PM_KEY_DEAD of a slab) clamp to nrows, so row_start[nrows] is the
// number of live particles.  A boundary j (between sorted entries j-1 and j) fills every row that
// starts there, empty rows included.
__device__ __forceinline__ int64_t pm_rowseg(uint32_t key, int xseg, int shift, int64_t nrows)
{
    const int64_t r = shift >= 0 ? (int64_t)(key >> shift) : (int64_t)(key / (uint32_t)xseg);
    return r > nrows ? nrows : r;
}

struct RowArgs {
    uint32_t *row_start;
    int xseg, shift;
    int64_t nrows;
};

// shared-memory index of element j of a 64-bit tile: one pad slot per 16 elements, so that threads
// walking 8 consecutive elements each (stride 64 bytes) spread over all banks
__device__ __forceinline__ uint32_t pm_pad(uint32_t j) { return j + (j >> 4); }
constexpr int kTilePad = kTile + kTile / 16;

__global__ void __launch_bounds__(kThreads) k_merge_tiles(const uint64_t *__restrict__ a,
                                                          const uint64_t *__restrict__ b,
                                                          const uint32_t *__restrict__ split,
                                                          uint32_t n, uint32_t *__restrict__ keys_sorted,
                                                          uint32_t *__restrict__ order_sorted, RowArgs ra)
{
    __shared__ uint64_t s_in[kTilePad];
    __shared__ uint64_t s_out[kTilePad];
    __shared__ int64_t s_prev_row;
    const int tid = threadIdx.x;
    const uint64_t d064 = (uint64_t)blockIdx.x * kTile;
    const uint32_t d0 = (uint32_t)d064;
    const uint32_t d1 = (d064 + kTile < n) ? d0 + kTile : n;
    const uint32_t a0 = split[blockIdx.x], a1 = split[blockIdx.x + 1];
    const uint32_t b0 = d0 - a0, b1 = d1 - a1;
    const uint32_t ca = a1 - a0, cb = b1 - b0, total = ca + cb;   // total == d1 - d0
    for (uint32_t j = tid; j < total; j += kThreads)
        s_in[pm_pad(j)] = (j < ca) ? a[a0 + j] : b[b0 + (j - ca)];
    if (tid == 0) {
        // row of the last entry of the previous tile (-1 before the first entry)
        int64_t pr = -1;
        if (d0 > 0) {
            const uint64_t la = a0 > 0 ? a[a0 - 1] : 0ull, lb = b0 > 0 ? b[b0 - 1] : 0ull;
            const uint64_t last = la > lb ? la : lb;
            pr = pm_rowseg((uint32_t)(last >> 32), ra.xseg, ra.shift, ra.nrows);
        }
        s_prev_row = pr;
    }
    __syncthreads();
    auto sa = [&](uint32_t i) { return s_in[pm_pad(i)]; };
    auto sb = [&](uint32_t i) { return s_in[pm_pad(ca + i)]; };
    const uint32_t ld = min((uint32_t)tid * kRounds, total);
    uint32_t la;
    {
        uint32_t lo = ld > cb ? ld - cb : 0u, hi = ld < ca ? ld : ca;
        while (lo < hi) {
            const uint32_t mid = lo + ((hi - lo) >> 1);
            if (sa(mid) < sb(ld - 1 - mid)) lo = mid + 1;
            else hi = mid;
        }
        la = lo;
    }
    uint32_t lb = ld - la;
#pragma unroll
    for (int k = 0; k < kRounds; ++k) {
        if (ld + k < total) {
            const bool take_a = (lb >= cb) || (la < ca && sa(la) < sb(lb));
            s_out[pm_pad(ld + k)] = take_a ? sa(la) : sb(lb);
            la += take_a ? 1u : 0u;
            lb += take_a ? 0u : 1u;
        }
    }
    __syncthreads();
    for (uint32_t j = tid; j < total; j += kThreads) {
        const uint64_t v = s_out[pm_pad(j)];
        const uint32_t key = (uint32_t)(v >> 32);
        keys_sorted[d0 + j] = key;
        order_sorted[d0 + j] = (uint32_t)v;
        // boundary d0 + j
        const int64_t prev = (j == 0) ? s_prev_row
                                      : pm_rowseg((uint32_t)(s_out[pm_pad(j - 1)] >> 32), ra.xseg, ra.shift, ra.nrows);
        const int64_t cur = pm_rowseg(key, ra.xseg, ra.shift, ra.nrows);
        for (int64_t r = prev + 1; r <= cur; ++r) ra.row_start[r] = d0 + j;
    }
    if (d1 == n && tid == 0) {
        // boundary n: everything after the last entry's row
        const int64_t prev = total > 0 ? pm_rowseg((uint32_t)(s_out[pm_pad(total - 1)] >> 32), ra.xseg, ra.shift, ra.nrows)
                                       : s_prev_row;
        for (int64_t r = prev + 1; r <= ra.nrows; ++r) ra.row_start[r] = n;
    }
}

// The same table from an already sorted key array (full-sort path).  One thread per four
// boundaries j in [0, np].
__global__ void __launch_bounds__(256) k_row_offsets(const uint32_t *__restrict__ keys_sorted,
                                                     int64_t np, int xseg, int shift, int64_t nrows,
                                                     uint32_t *__restrict__ row_start)
{
    const int64_t j0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (j0 > np) return;
    uint32_t k[4];
    if (j0 + 4 <= np) {
        const uint4 v = *reinterpret_cast<const uint4 *>(keys_sorted + j0);
        k[0] = v.x; k[1] = v.y; k[2] = v.z; k[3] = v.w;
    } else {
#pragma unroll
        for (int t = 0; t < 4; ++t) k[t] = (j0 + t < np) ? keys_sorted[j0 + t] : 0u;
    }
    int64_t prev = (j0 == 0) ? -1 : pm_rowseg(keys_sorted[j0 - 1], xseg, shift, nrows);
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        const int64_t j = j0 + t;
        if (j > np) break;
        const int64_t cur = (j == np) ? nrows : pm_rowseg(k[t], xseg, shift, nrows);
        for (int64_t r = prev + 1; r <= cur; ++r) row_start[r] = (uint32_t)j;
        prev = cur;
    }
}

RowArgs row_args(pm_plan *p)
{
    RowArgs ra;
    ra.row_start = p->row_start;
    ra.xseg = p->nc / p->dep_nseg;
    ra.shift = -1;
    if ((ra.xseg & (ra.xseg - 1)) == 0) {
        ra.shift = 0;
        while ((1 << ra.shift) < ra.xseg) ++ra.shift;
    }
    ra.nrows = (int64_t)p->nzl * p->nc * p->dep_nseg;
    return ra;
}

int full_sort(pm_plan *p, int64_t n, cudaStream_t st)
{
    size_t bytes = p->cub_bytes;
    PM_CUDA(cub::DeviceRadixSort::SortPairs(p->cub_tmp, bytes, (const uint32_t *)p->keys,
                                            p->keys_sorted, (const uint32_t *)p->iota,
                                            p->order_sorted, n, 0, p->key_bits, st));
    p->rows_valid = false;
    p->sort_last_mode = PM_SORT_FULL;
    p->sort_last_n = n;
    p->sort_last_movers = n;
    return PM_OK;
}

}  // namespace

int64_t pm_sort_tiles(int64_t np) { return (np + kTile - 1) / kTile; }

int64_t pm_sort_mover_capacity(int64_t np)
{
    // movers the incremental path accepts before it falls back to the full sort
    return np / 5 * 2 + kTile;
}

size_t pm_sort_temp_bytes(int64_t np, int key_bits)
{
    size_t pairs = 0, movers = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, pairs, (const uint32_t *)nullptr, (uint32_t *)nullptr,
                                    (const uint32_t *)nullptr, (uint32_t *)nullptr, np, 0, key_bits);
    cub::DeviceRadixSort::SortKeys(nullptr, movers, (const uint64_t *)nullptr, (uint64_t *)nullptr,
                                   pm_sort_mover_capacity(np), 32, 32 + key_bits);
    return pairs > movers ? pairs : movers;
}

// Stable sort of (key, slot) over p->keys[0..np) into p->keys_sorted / p->order_sorted.
// `n_old` > 0: the first n_old entries are stored in the order of the previous sort and
// p->keys_sorted[0..n_old) still holds the keys they were sorted by -> incremental path.
int pm_k_sort(pm_plan *p, int64_t np, int64_t n_old, cudaStream_t st)
{
    p->rows_valid = false;
    if (np == 0) return PM_OK;
    if (n_old > np) n_old = np;
    if (p->sort_mode == PM_SORT_FULL || n_old <= 0 || !p->inc_a) {
        p->inc_counted = false;
        return full_sort(p, np, st);
    }

    const int nt = (int)pm_sort_tiles(np);
    // the resident gather counts the movers of each tile while it writes the keys (inc_counted)
    if (!(p->inc_counted && n_old == np))
        PM_LAUNCH(k_mover_count, nt, kThreads, 0, st, (const uint32_t *)p->keys,
                  (const uint32_t *)p->keys_sorted, np, n_old, p->inc_tile);
    p->inc_counted = false;
    PM_LAUNCH(k_mover_scan, 1, 1024, 0, st, p->inc_tile, nt);
    PM_CHECK_LAUNCH();
    PM_CUDA(cudaMemcpyAsync(p->h_word, p->inc_tile + nt, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    PM_CUDA(cudaStreamSynchronize(st));
    const int64_t nb = (int64_t)p->h_word[0];
    if (nb > p->inc_bcap - kTile) return full_sort(p, np, st);

    PM_LAUNCH(k_mover_partition, nt, kThreads, 0, st, (const uint32_t *)p->keys,
              (const uint32_t *)p->keys_sorted, np, n_old, (const uint32_t *)p->inc_tile, p->inc_a,
              p->inc_b);
    PM_CHECK_LAUNCH();
    const uint64_t *bs = p->inc_b;
    if (nb > 1) {
        size_t bytes = p->cub_bytes;
        PM_CUDA(cub::DeviceRadixSort::SortKeys(p->cub_tmp, bytes, (const uint64_t *)p->inc_b, p->inc_bs,
                                               nb, 32, 32 + p->key_bits, st));
        bs = p->inc_bs;
    }
    PM_LAUNCH(k_merge_splits, (nt + 1 + 255) / 256, 256, 0, st, (const uint64_t *)p->inc_a, bs,
              (uint32_t)nb, (uint32_t)np, nt, p->inc_split);
    PM_LAUNCH(k_merge_tiles, nt, kThreads, 0, st, (const uint64_t *)p->inc_a, bs,
              (const uint32_t *)p->inc_split, (uint32_t)np, p->keys_sorted, p->order_sorted, row_args(p));
    PM_CHECK_LAUNCH();
    p->rows_valid = true;   // the merge wrote row_start as well
    p->sort_last_mode = PM_SORT_INCREMENTAL;
    p->sort_last_n = np;
    p->sort_last_movers = nb;
    return PM_OK;
}

// Row table of the sorted list (no-op when the merge of the incremental sort already wrote it).
int pm_k_row_offsets(pm_plan *p, int64_t np, cudaStream_t st)
{
    if (p->rows_valid && np > 0) return PM_OK;
    const RowArgs ra = row_args(p);
    const int64_t nthr = np / 4 + 1;
    PM_LAUNCH(k_row_offsets, (unsigned)((nthr + 255) / 256), 256, 0, st, (const uint32_t *)p->keys_sorted, np,
              ra.xseg, ra.shift, ra.nrows, p->row_start);
    PM_CHECK_LAUNCH();
    p->rows_valid = true;
    return PM_OK;
}

// ascending sort of a short list of 32-bit values (migration leave lists)
int pm_k_sort_u32(pm_plan *p, const uint32_t *in, uint32_t *out, int64_t count, cudaStream_t st)
{
    size_t bytes = p->cub_bytes;
    PM_CUDA(cub::DeviceRadixSort::SortKeys(p->cub_tmp, bytes, in, out, count, 0, 32, st));
    return PM_OK;
}
