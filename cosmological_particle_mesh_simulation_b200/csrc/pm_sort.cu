// pm_sort.cu -- ordering the particle list by cell key (src/density.py:19-21,37: key = (z_c*Nc + y_c)*Nc + x_c).
//
// Two ways to produce (keys_sorted, order_sorted), with bit-identical results:
//
//   full         stable LSD radix sort of (key << 32 | slot), 8-bit digits -- k_radix_sort below, ONE
//                persistent cooperative launch for all passes (per pass: per-CTA digit histograms of a
//                contiguous range of tiles, grid barrier, column sums -> the CTA's first output slot
//                of every digit, stable in-tile ranking by warp match, scatter, grid barrier).
//
//   incremental  the resident state is stored in the cell order of the PREVIOUS step, and a particle
//                drifts much less than a cell per step, so most entries still carry the key they
//                were sorted by ("stayers": an already sorted subsequence) and only the "movers"
//                need sorting.  Per step:
//                  1. k_mover_count      movers per 2048-entry tile (skipped: the resident gather counts them)
//                  2. k_mover_scan       exclusive scan of the tile counts; total and the mode -> SortCtl
//                  3. k_mover_partition  stayers -> A, movers -> B as 64-bit (key << 32 | slot)
//                  4. k_radix_sort       B on the key bits (stable: slots stay ascending per key)
//                  5. k_merge_splits     merge-path split of every 2048-entry output tile
//                  6. k_merge_tiles      merge A and B by (key, slot) -> keys_sorted, order_sorted,
//                                        and the row table of the deposit (row_start) on the way
//                Because both inputs are ordered by the composite (key, slot), the merge reproduces
//                exactly the order a stable sort of all entries gives.
//
// No host round trip: the mover count and the decision between the two ways (more than 40 % movers ->
// full) live in device memory (SortCtl); every kernel of both ways is launched every step with its
// worst-case grid and returns at once when the other way was chosen, so the whole step is a fixed
// launch sequence that a CUDA graph can hold.  pm_plan_sort_stats reads SortCtl back on demand.
//
// There is no reference counterpart: the reference scatters in particle-index order
// (src/density.py:17) and never sorts.  The order only fixes the memory locality of the deposit and
// gather kernels (and, for k_deposit_rows, its float32 summation tree).
#include <stdlib.h>
#include <string.h>

#include "pm_internal.cuh"

namespace {

// Device-resident control block of one sort (p->sort_ctl).
struct SortCtl {
    uint32_t nb;        // movers of this step (incremental way)
    uint32_t full;      // 1: this step runs the full radix sort, 0: the incremental way
    uint32_t n;         // entries sorted
    uint32_t err;       // a grid barrier of k_radix_sort gave up (never on a healthy device; sticky)
    uint32_t bar;       // arrival counter of those barriers (zeroed before every launch)
    uint32_t pad[3];
};

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
__global__ void k_sort_ctl_init(SortCtl *ctl, uint32_t n, uint32_t force_full)
{
    ctl->nb = force_full ? n : 0u;
    ctl->full = force_full;
    ctl->n = n;
    ctl->bar = 0u;
}

__global__ void k_sort_bar_reset(SortCtl *ctl) { ctl->bar = 0u; }

__global__ void __launch_bounds__(1024) k_mover_scan(uint32_t *__restrict__ cnt, int nt, SortCtl *ctl, uint32_t cap)
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
    if (tid == 1023) {
        cnt[nt] = s_w[31];
        ctl->nb = s_w[31];
        if (s_w[31] > cap) ctl->full = 1u;     // too many movers: the full sort is cheaper (and B would overflow)
    }
}

__global__ void __launch_bounds__(kThreads) k_mover_partition(const uint32_t *__restrict__ keys,
                                                              const uint32_t *__restrict__ keys_old,
                                                              int64_t n, int64_t n_old,
                                                              const uint32_t *__restrict__ tile_off,
                                                              uint64_t *__restrict__ a64,
                                                              uint64_t *__restrict__ b64,
                                                              const SortCtl *__restrict__ ctl)
{
    __shared__ uint32_t s_w[kWarps];
    if (ctl->full) return;
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
                                                      const SortCtl *__restrict__ ctl, uint32_t n, int nt,
                                                      uint32_t *__restrict__ split)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t > nt || ctl->full) return;
    const uint32_t nb = ctl->nb;
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

// One output tile of the merge.  97 % of the entries are stayers, so instead of a generic two-way merge
// (a merge-path search and a serial compare-and-take loop per thread: 151 instructions per entry, issue-bound
// at 91 us) the tile's FEW movers are placed and the stayers flow around them: each mover j of the tile
// (already sorted) finds by binary search how many of the tile's stayers precede it -> its output slot, one
// bit per slot; a popcount prefix over the 2048-bit map then tells every output slot whether it takes a mover
// (and which: movers keep their order) or the stayer of index slot - movers before it.  Same result as the
// merge, bit for bit (no ties: the slot half makes every composite unique).
__global__ void __launch_bounds__(kThreads) k_merge_tiles(const uint64_t *__restrict__ a,
                                                          const uint64_t *__restrict__ b,
                                                          const uint32_t *__restrict__ split,
                                                          uint32_t n, uint32_t *__restrict__ keys_sorted,
                                                          uint32_t *__restrict__ order_sorted, RowArgs ra,
                                                          const SortCtl *__restrict__ ctl)
{
    __shared__ uint64_t s_in[kTilePad];
    if (ctl->full) return;
    static_assert(kTile == 2048, "the prefix below handles 64 map words with one warp");
    __shared__ uint32_t s_key[kTile];            // merged keys (the row table looks at neighbours)
    __shared__ uint32_t s_bits[kTile / 32];      // output slots taken by movers
    __shared__ uint32_t s_pre[kTile / 32];       // movers in the words before
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
    if (tid < kTile / 32) s_bits[tid] = 0u;
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
    // movers -> output slots
    for (uint32_t j = tid; j < cb; j += kThreads) {
        const uint64_t v = sb(j);
        uint32_t lo = 0, hi = ca;                 // stayers of the tile below v
        while (lo < hi) {
            const uint32_t mid = lo + ((hi - lo) >> 1);
            if (sa(mid) < v) lo = mid + 1;
            else hi = mid;
        }
        const uint32_t slot = lo + j;
        atomicOr(&s_bits[slot >> 5], 1u << (slot & 31));
    }
    __syncthreads();
    if (tid < 32) {
        // exclusive prefix of the per-word mover counts (64 words: two per lane)
        const uint32_t c0 = __popc(s_bits[2 * tid]), c1 = __popc(s_bits[2 * tid + 1]);
        uint32_t inc = c0 + c1;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t u = __shfl_up_sync(0xffffffffu, inc, o);
            if (tid >= o) inc += u;
        }
        s_pre[2 * tid] = inc - c0 - c1;
        s_pre[2 * tid + 1] = inc - c1;
    }
    __syncthreads();
    for (uint32_t j = tid; j < total; j += kThreads) {
        const uint32_t w = s_bits[j >> 5], bit = 1u << (j & 31);
        const uint32_t r = s_pre[j >> 5] + __popc(w & (bit - 1u));       // movers before slot j
        const uint64_t v = (w & bit) ? sb(r) : sa(j - r);
        const uint32_t key = (uint32_t)(v >> 32);
        s_key[j] = key;
        keys_sorted[d0 + j] = key;
        order_sorted[d0 + j] = (uint32_t)v;
    }
    __syncthreads();
    for (uint32_t j = tid; j < total; j += kThreads) {
        // boundary d0 + j
        const int64_t prev = (j == 0) ? s_prev_row : pm_rowseg(s_key[j - 1], ra.xseg, ra.shift, ra.nrows);
        const int64_t cur = pm_rowseg(s_key[j], ra.xseg, ra.shift, ra.nrows);
        for (int64_t r = prev + 1; r <= cur; ++r) ra.row_start[r] = d0 + j;
    }
    if (d1 == n && tid == 0) {
        // boundary n: everything after the last entry's row
        const int64_t prev = total > 0 ? pm_rowseg(s_key[total - 1], ra.xseg, ra.shift, ra.nrows) : s_prev_row;
        for (int64_t r = prev + 1; r <= ra.nrows; ++r) ra.row_start[r] = n;
    }
}

// The same table from an already sorted key array (full-sort path).  One thread per four
// boundaries j in [0, np].
__global__ void __launch_bounds__(256) k_row_offsets(const uint32_t *__restrict__ keys_sorted,
                                                     int64_t np, int xseg, int shift, int64_t nrows,
                                                     uint32_t *__restrict__ row_start,
                                                     const SortCtl *__restrict__ ctl)
{
    if (ctl && !ctl->full) return;     // the merge of the incremental way wrote the table already
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

// ------------------------------------------------------------------------------------------------
// k_radix_sort: stable LSD radix sort of 64-bit items on an 8-bit digit per pass, every pass inside
// ONE persistent cooperative launch.  Which of two jobs it runs is read from device memory
// (SortCtl::full), as is the item count, so the host never has to know either.
// ------------------------------------------------------------------------------------------------
struct RadixJob {
    const uint32_t *in_keys32;    // non-null: pass 0 synthesises item i = (in_keys32[i] << 32) | i
    uint64_t *buf_a, *buf_b;      // ping-pong buffers; without in_keys32 the input is in buf_a
    uint32_t *out_hi, *out_lo;    // non-null out_hi: the LAST pass writes the two halves here instead of a buffer
    const uint32_t *n_ptr;        // item count in device memory (nullptr: n_fixed)
    uint32_t n_fixed;
    int bit0, passes;             // pass p sorts on bits [bit0 + 8p, bit0 + 8p + 8)
};

struct RadixArgs {
    RadixJob job[2];
    const uint32_t *select;       // job index in device memory (nullptr: 0)
    uint32_t *hist;               // [gridDim.x][256]
    uint32_t *bar, *err;
};

constexpr int kRadixItems = kTile / kThreads;   // items per thread and tile

// The first `nact` CTAs of the (cooperative, hence co-resident) grid meet here; the others left at
// kernel entry.  `bar` was zeroed before the launch and counts arrivals; barrier number k (1-based) is
// passed when it reaches k * nact.  A wait that sees no progress for ~1 s gives up and raises `err`
// instead of hanging the device.
__device__ __forceinline__ void pm_grid_barrier(uint32_t *bar, uint32_t *err, uint32_t &round, uint32_t nact)
{
    __syncthreads();
    if (threadIdx.x == 0) {
        ++round;
        const uint32_t target = round * nact;
        __threadfence();
        atomicAdd(bar, 1u);
        uint32_t spins = 0;
        for (;;) {
            uint32_t v;
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory");
            if ((int32_t)(v - target) >= 0) break;
            if (++spins > (1u << 22)) {
                atomicExch(err, 1u);
                break;
            }
            __nanosleep(100);
        }
        __threadfence();
    }
    __syncthreads();
}

// DIG-bit digits (8 or 9): NB bins, NB / kThreads bins per thread.
template <int DIG>
__global__ void __launch_bounds__(kThreads, 3) k_radix_sort(RadixArgs A)
{
    constexpr int NB = 1 << DIG, DPT = NB / kThreads;    // bins; bins per thread
    static_assert(DPT >= 1 && DPT * kThreads == NB && kTile * 2 >= 8 * NB, "digit width");
    __shared__ uint64_t s_items[kTile];        // the tile in digit order, written out in contiguous runs
    __shared__ uint32_t s_cnt[kWarps][NB];     // digit counts per warp -> tile-local first slot per (warp, digit)
    __shared__ uint32_t s_base[NB];            // next global output slot of every digit for this CTA
    __shared__ uint32_t s_goff[NB];            // global slot of tile-local index i with digit d = s_goff[d] + i
    __shared__ uint32_t s_scan[NB];
    const RadixJob J = A.job[A.select ? (*A.select ? 1 : 0) : 0];
    const uint32_t n = J.n_ptr ? *J.n_ptr : J.n_fixed;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned full = 0xffffffffu;
    const uint32_t c = blockIdx.x;
    const uint32_t ntiles = (n + kTile - 1) / kTile;
    // contiguous ranges of `per` tiles (at least two: halves the CTAs in the barriers and column sums of
    // a short list); only the CTAs that own tiles stay
    uint32_t per = (ntiles + gridDim.x - 1) / gridDim.x;
    per = per < 2u ? 2u : per;
    const uint32_t nact = (ntiles + per - 1) / per;
    if (c >= nact) return;
    const uint32_t t0 = c * per, t1 = min(t0 + per, ntiles);
    const uint32_t mask = (uint32_t)NB - 1u;
    uint32_t round = 0;
    uint64_t *cur = J.buf_a;                   // where the items are (once they exist as 64-bit items)
    for (int p = 0; p < J.passes; ++p) {
        const bool synth = (p == 0 && J.in_keys32 != nullptr);
        const bool split = (p == J.passes - 1 && J.out_hi != nullptr);
        const uint64_t *src = cur;
        uint64_t *dst = synth ? J.buf_a : (cur == J.buf_a ? J.buf_b : J.buf_a);
        const int shift = J.bit0 + DIG * p;
        auto load = [&](uint32_t i) -> uint64_t {
            // ld.cg: the buffers are rewritten by other SMs between passes of this one launch and L1 is not
            // coherent across SMs -- a cached line from two passes ago would be stale
            return synth ? (((uint64_t)J.in_keys32[i] << 32) | (uint64_t)i) : __ldcg(src + i);
        };
        // ---- histogram of this CTA's tiles ----
#pragma unroll
        for (int q = 0; q < DPT; ++q) s_scan[tid + q * kThreads] = 0u;
        __syncthreads();
        // All loads of a tile first, then the digits, and no branch around a load (the index is clamped instead):
        // with `i < n ? load(i) : ...` inside the digit loop every load sat in its own reconvergence region, one
        // DRAM latency after the other (ncu source page of the 16.7 M-entry sort, call LL: the shift behind each
        // load held 27 % of the kernel's samples).  Measured effect (call MM): none to speak of -- full sort 1.11 ->
        // 1.10 ms, movers' sort unchanged -- so the kernel is bound elsewhere (24 warps per SM, issue 20 %: the
        // match -> shared-memory read-modify-write -> shuffle chain of the ranking phase); kept because it is the
        // better code, recorded because it was not the fix.
        auto hist_tiles = [&](auto ld) {
            for (uint32_t t = t0; t < t1; ++t) {
                const uint32_t base = t * kTile;
                uint64_t v[kRadixItems];
#pragma unroll
                for (int k = 0; k < kRadixItems; ++k) {
                    const uint32_t i = base + k * kThreads + tid;
                    v[k] = ld(i < n ? i : n - 1u);
                }
#pragma unroll
                for (int k = 0; k < kRadixItems; ++k) {
                    const uint32_t i = base + k * kThreads + tid;
                    const uint32_t d = i < n ? ((uint32_t)(v[k] >> shift) & mask) : (uint32_t)NB + lane;   // idle lanes match nobody
                    // one shared-memory atomic per distinct digit of the warp: the high digits of a nearly
                    // sorted list are all the same, which would serialise 32 ways otherwise
                    const unsigned peers = __match_any_sync(full, d);
                    if (i < n && lane == __ffs(peers) - 1) atomicAdd(&s_scan[d], (uint32_t)__popc(peers));
                }
            }
        };
        if (synth) hist_tiles([&](uint32_t i) -> uint64_t { return ((uint64_t)J.in_keys32[i] << 32) | (uint64_t)i; });
        else hist_tiles([&](uint32_t i) -> uint64_t { return __ldcg(src + i); });
        __syncthreads();
#pragma unroll
        for (int q = 0; q < DPT; ++q) A.hist[(size_t)c * NB + tid + q * kThreads] = s_scan[tid + q * kThreads];
        pm_grid_barrier(A.bar, A.err, round, nact);
        // ---- first output slot of every digit for this CTA: digits below, then CTAs below ----
        // Column scan of hist[0..nact)[NB], shared out over the CTAs: CTA c scans the columns (digits)
        // [c*cpc, c*cpc+cpc) over all rows -- one warp per column, 32 rows per shuffle scan -- writing the
        // exclusive prefixes to pref[][] and the column totals to tot[]; after one more grid barrier every
        // CTA needs only its own row of pref and tot.  (Each CTA summing all rows itself made nact^2
        // row reads per pass: half of this kernel's time.)
        {
            uint32_t *pref = A.hist + (size_t)PM_SORT_MAX_GRID * NB;       // [nact][NB]
            uint32_t *tot = pref + (size_t)PM_SORT_MAX_GRID * NB;          // [NB]
            const uint32_t cpc = ((uint32_t)NB + nact - 1) / nact;
            for (uint32_t q = warp; q < cpc; q += kWarps) {
                const uint32_t d = c * cpc + q;
                if (d >= (uint32_t)NB) break;
                uint32_t carry = 0;
                for (uint32_t r0 = 0; r0 < nact; r0 += 32) {
                    const uint32_t r = r0 + lane;
                    const uint32_t v = r < nact ? __ldcg(A.hist + (size_t)r * NB + d) : 0u;
                    uint32_t inc = v;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const uint32_t u = __shfl_up_sync(full, inc, o);
                        if (lane >= o) inc += u;
                    }
                    if (r < nact) pref[(size_t)r * NB + d] = carry + inc - v;
                    carry += __shfl_sync(full, inc, 31);
                }
                if (lane == 0) tot[d] = carry;
            }
            pm_grid_barrier(A.bar, A.err, round, nact);
            // thread t owns the DPT CONSECUTIVE digits t*DPT .. t*DPT+DPT-1 (so the scan below is in digit order)
            uint32_t total[DPT], prefix[DPT], tsum = 0;
#pragma unroll
            for (int q = 0; q < DPT; ++q) {
                const int d = tid * DPT + q;
                total[q] = __ldcg(tot + d);
                prefix[q] = __ldcg(pref + (size_t)c * NB + d);
                tsum += total[q];
            }
            uint32_t inc = tsum;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t o = __shfl_up_sync(full, inc, d);
                if (lane >= d) inc += o;
            }
            if (lane == 31) s_scan[warp] = inc;            // s_scan was last read before the grid barriers
            __syncthreads();
            uint32_t run = inc - tsum;
            for (int w = 0; w < warp; ++w) run += s_scan[w];
#pragma unroll
            for (int q = 0; q < DPT; ++q) {
                s_base[tid * DPT + q] = run + prefix[q];
                run += total[q];
            }
            __syncthreads();
        }
        // ---- rank, stage in digit order, write out in runs; tile by tile, in order ----
        for (uint32_t t = t0; t < t1; ++t) {
            for (int d = tid; d < kWarps * NB; d += kThreads) (&s_cnt[0][0])[d] = 0u;
            __syncthreads();
            // warp w owns entries [w * 256, w * 256 + 256) of the tile, 32 consecutive ones per round:
            // (warp, round, lane) is the input order, which the ranks below preserve (stability)
            const uint32_t tbase = t * kTile, wbase = tbase + warp * (kRadixItems * 32);
            const uint32_t tcount = min((uint32_t)kTile, n - tbase);
            uint64_t item[kRadixItems];
            uint32_t rank[kRadixItems];
#pragma unroll
            for (int k = 0; k < kRadixItems; ++k) {
                const uint32_t i = wbase + k * 32 + lane;
                item[k] = i < n ? load(i) : 0ull;
            }
#pragma unroll
            for (int k = 0; k < kRadixItems; ++k) {
                const uint32_t i = wbase + k * 32 + lane;
                const bool valid = i < n;
                const uint32_t d = valid ? ((uint32_t)(item[k] >> shift) & mask) : (uint32_t)NB + lane;   // idle lanes match nobody
                const unsigned peers = __match_any_sync(full, d);
                const int leader = __ffs(peers) - 1;
                uint32_t old = 0;
                if (valid && lane == leader) {
                    old = s_cnt[warp][d];
                    s_cnt[warp][d] = old + __popc(peers);
                }
                old = __shfl_sync(full, old, leader);
                rank[k] = old + __popc(peers & ((1u << lane) - 1u));
                __syncwarp();
            }
            __syncthreads();
            // digits tid*DPT ..: the tile's counts, their first tile-local slots (exclusive scan over the
            // digits), the (warp, digit) first slots, and the global slot of each digit's run
            {
                uint32_t tot[DPT], tsum = 0;
#pragma unroll
                for (int q = 0; q < DPT; ++q) {
                    tot[q] = 0;
#pragma unroll
                    for (int w = 0; w < kWarps; ++w) tot[q] += s_cnt[w][tid * DPT + q];
                    tsum += tot[q];
                }
                uint32_t inc = tsum;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const uint32_t o = __shfl_up_sync(full, inc, d);
                    if (lane >= d) inc += o;
                }
                if (lane == 31) s_scan[warp] = inc;
                __syncthreads();
                uint32_t tstart = inc - tsum;
                for (int w = 0; w < warp; ++w) tstart += s_scan[w];
#pragma unroll
                for (int q = 0; q < DPT; ++q) {
                    const int d = tid * DPT + q;
                    uint32_t run = tstart;
#pragma unroll
                    for (int w = 0; w < kWarps; ++w) {
                        const uint32_t v = s_cnt[w][d];
                        s_cnt[w][d] = run;
                        run += v;
                    }
                    s_goff[d] = s_base[d] - tstart;        // modular arithmetic: only the sum with i is used
                    s_base[d] += tot[q];
                    tstart += tot[q];
                }
            }
            __syncthreads();
#pragma unroll
            for (int k = 0; k < kRadixItems; ++k) {
                const uint32_t i = wbase + k * 32 + lane;
                if (i < n) s_items[s_cnt[warp][(uint32_t)(item[k] >> shift) & mask] + rank[k]] = item[k];
            }
            __syncthreads();
            for (uint32_t i = tid; i < tcount; i += kThreads) {
                const uint64_t v = s_items[i];
                const uint32_t o = s_goff[(uint32_t)(v >> shift) & mask] + i;
                if (split) {
                    J.out_hi[o] = (uint32_t)(v >> 32);
                    if (J.out_lo) J.out_lo[o] = (uint32_t)v;
                } else {
                    dst[o] = v;
                }
            }
            __syncthreads();
        }
        cur = dst;
        if (p + 1 < J.passes) pm_grid_barrier(A.bar, A.err, round, nact);
    }
}

// digit width for a key of `bits` bits: 9 where that saves a pass (27-bit keys of a 512^3 mesh: 3 passes)
int radix_digit(int bits) { return (bits > 0 && (bits + 8) / 9 < (bits + 7) / 8) ? 9 : 8; }
int radix_passes(int bits, int dig) { return bits <= 0 ? 1 : (bits + dig - 1) / dig; }

int radix_grid(pm_plan *p)
{
    // co-resident CTAs of k_radix_sort on this device (cooperative launch), at most three per SM
    static int per_sm_dev[64];
    int &per_sm = per_sm_dev[p->device & 63];
    if (per_sm == 0) {
        int v8 = 0, v9 = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v8, k_radix_sort<8>, kThreads, 0) != cudaSuccess || v8 < 1) v8 = 1;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v9, k_radix_sort<9>, kThreads, 0) != cudaSuccess || v9 < 1) v9 = 1;
        const int v = v8 < v9 ? v8 : v9;
        per_sm = v > 3 ? 3 : v;
    }
    int g = per_sm * p->sm_count;
    return g > PM_SORT_MAX_GRID ? PM_SORT_MAX_GRID : g;
}

int launch_radix(pm_plan *p, RadixArgs &A, int dig, cudaStream_t st)
{
    A.hist = p->sort_hist;
    SortCtl *ctl = reinterpret_cast<SortCtl *>(p->sort_ctl);
    A.bar = &ctl->bar;
    A.err = &ctl->err;
    void *args[] = {&A};
    const void *fn = dig == 9 ? (const void *)k_radix_sort<9> : (const void *)k_radix_sort<8>;
    PM_CUDA(cudaLaunchCooperativeKernel(fn, dim3(radix_grid(p)), dim3(kThreads), args, 0, st));
    ++g_pm_launches;
    return PM_OK;
}


}  // namespace

int64_t pm_sort_tiles(int64_t np) { return (np + kTile - 1) / kTile; }

int64_t pm_sort_mover_capacity(int64_t np)
{
    // movers the incremental path accepts before it falls back to the full sort
    return np / 5 * 2 + kTile;
}

// Stable sort of (key, slot) over p->keys[0..np) into p->keys_sorted / p->order_sorted, and the row
// table of the deposit.  `n_old` > 0: the first n_old entries are stored in the order of the previous
// sort and p->keys_sorted[0..n_old) still holds the keys they were sorted by -> incremental way, unless
// the device finds more than 40 % movers.  No host synchronisation.
int pm_k_sort(pm_plan *p, int64_t np, int64_t n_old, cudaStream_t st)
{
    p->rows_valid = false;
    if (np == 0) return PM_OK;
    if (n_old > np) n_old = np;
    SortCtl *ctl = reinterpret_cast<SortCtl *>(p->sort_ctl);
    const bool force_full = (p->sort_mode == PM_SORT_FULL || n_old <= 0 || !p->inc_a);
    const int nt = (int)pm_sort_tiles(np);
    PM_LAUNCH(k_sort_ctl_init, 1, 1, 0, st, ctl, (uint32_t)np, force_full ? 1u : 0u);
    if (!force_full) {
        // the resident gather counts the movers of each tile while it writes the keys (inc_counted)
        if (!(p->inc_counted && n_old == np))
            PM_LAUNCH(k_mover_count, nt, kThreads, 0, st, (const uint32_t *)p->keys,
                      (const uint32_t *)p->keys_sorted, np, n_old, p->inc_tile);
        PM_LAUNCH(k_mover_scan, 1, 1024, 0, st, p->inc_tile, nt, ctl, (uint32_t)(p->inc_bcap - kTile));
        PM_LAUNCH(k_mover_partition, nt, kThreads, 0, st, (const uint32_t *)p->keys,
                  (const uint32_t *)p->keys_sorted, np, n_old, (const uint32_t *)p->inc_tile, p->inc_a,
                  p->inc_b, (const SortCtl *)ctl);
        PM_CHECK_LAUNCH();
    }
    p->inc_counted = false;
    const int dig = radix_digit(p->key_bits), passes = radix_passes(p->key_bits, dig);
    RadixArgs A;
    // job 0: the movers B, 64-bit items already, sorted on the key half
    A.job[0].in_keys32 = nullptr;
    A.job[0].buf_a = p->inc_b; A.job[0].buf_b = p->inc_bs;
    A.job[0].out_hi = nullptr; A.job[0].out_lo = nullptr;
    A.job[0].n_ptr = &ctl->nb; A.job[0].n_fixed = 0;
    A.job[0].bit0 = 32; A.job[0].passes = passes;
    // job 1: every entry, synthesised from the key array, written as (keys_sorted, order_sorted)
    A.job[1].in_keys32 = p->keys;
    A.job[1].buf_a = p->inc_a; A.job[1].buf_b = p->sort_tmp;
    A.job[1].out_hi = p->keys_sorted; A.job[1].out_lo = p->order_sorted;
    A.job[1].n_ptr = nullptr; A.job[1].n_fixed = (uint32_t)np;
    A.job[1].bit0 = 32; A.job[1].passes = passes;
    A.select = &ctl->full;
    int dig_launch = dig;
    if (p->sort_rows_only && force_full) {
        // The host-buffer step (pm_step_host, split route) sorts particles it will never see again: the tile deposit
        // and the row-block gather only need them grouped by mesh ROW (the row table), and their results do not
        // depend on the order inside a row (exact integer accumulation; per-particle gather).  Sorting on the row
        // bits alone -- stable, so a row keeps the caller's order -- saves one pass of three on 128^3..512^3 meshes.
        const RowArgs ra = row_args(p);
        int xbits = 0;
        while ((1 << xbits) < p->nc) ++xbits;
        if ((1 << xbits) == p->nc && ra.xseg == p->nc && p->key_bits > xbits) {
            const int rbits = p->key_bits - xbits;
            dig_launch = radix_digit(rbits);
            A.job[1].bit0 = 32 + xbits;
            A.job[1].passes = radix_passes(rbits, dig_launch);
        }
    }
    {
        const int rc = launch_radix(p, A, dig_launch, st);
        if (rc != PM_OK) return rc;
    }
    if (!force_full) {
        // sorted movers: passes ping-pong between inc_b and inc_bs, starting in inc_b
        const uint64_t *bs = (passes & 1) ? p->inc_bs : p->inc_b;
        PM_LAUNCH(k_merge_splits, (nt + 1 + 255) / 256, 256, 0, st, (const uint64_t *)p->inc_a, bs,
                  (const SortCtl *)ctl, (uint32_t)np, nt, p->inc_split);
        PM_LAUNCH(k_merge_tiles, nt, kThreads, 0, st, (const uint64_t *)p->inc_a, bs,
                  (const uint32_t *)p->inc_split, (uint32_t)np, p->keys_sorted, p->order_sorted, row_args(p),
                  (const SortCtl *)ctl);
    }
    {
        const RowArgs ra = row_args(p);
        const int64_t nthr = np / 4 + 1;
        PM_LAUNCH(k_row_offsets, (unsigned)((nthr + 255) / 256), 256, 0, st, (const uint32_t *)p->keys_sorted, np,
                  ra.xseg, ra.shift, ra.nrows, p->row_start, force_full ? (const SortCtl *)nullptr : (const SortCtl *)ctl);
    }
    PM_CHECK_LAUNCH();
    p->rows_valid = true;
    p->sort_last_n = np;
    return PM_OK;
}

// What the last pm_k_sort did, read back from the device (synchronises).
int pm_k_sort_stats(pm_plan *p, int64_t *entries, int64_t *movers, int *mode)
{
    SortCtl h;
    memset(&h, 0, sizeof(h));
    PM_CUDA(cudaDeviceSynchronize());
    PM_CUDA(cudaMemcpy(&h, p->sort_ctl, sizeof(h), cudaMemcpyDeviceToHost));
    if (entries) *entries = p->sort_last_n;
    if (movers) *movers = h.nb;
    if (mode) *mode = h.err ? PM_SORT_ERROR : (p->sort_last_n == 0 ? 0 : (h.full ? PM_SORT_FULL : PM_SORT_INCREMENTAL));
    return PM_OK;
}

// Row table of the sorted list (no-op: pm_k_sort writes it).
int pm_k_row_offsets(pm_plan *p, int64_t np, cudaStream_t st)
{
    if (p->rows_valid && np > 0) return PM_OK;
    const RowArgs ra = row_args(p);
    const int64_t nthr = np / 4 + 1;
    PM_LAUNCH(k_row_offsets, (unsigned)((nthr + 255) / 256), 256, 0, st, (const uint32_t *)p->keys_sorted, np,
              ra.xseg, ra.shift, ra.nrows, p->row_start, (const SortCtl *)nullptr);
    PM_CHECK_LAUNCH();
    p->rows_valid = true;
    return PM_OK;
}

// ascending sort of a short list of 32-bit values (migration leave lists): the same kernel, the
// values in the key half of synthesised items
int pm_k_sort_u32(pm_plan *p, const uint32_t *in, uint32_t *out, int64_t count, cudaStream_t st)
{
    if (count <= 0) return PM_OK;
    SortCtl *ctl = reinterpret_cast<SortCtl *>(p->sort_ctl);
    PM_LAUNCH(k_sort_bar_reset, 1, 1, 0, st, ctl);     // leaves the statistics of the last particle sort alone
    RadixArgs A;
    A.job[0].in_keys32 = in;
    A.job[0].buf_a = p->inc_a; A.job[0].buf_b = p->sort_tmp;
    A.job[0].out_hi = out; A.job[0].out_lo = nullptr;
    A.job[0].n_ptr = nullptr; A.job[0].n_fixed = (uint32_t)count;
    int bits = 1;
    while (bits < 32 && ((int64_t)1 << bits) <= p->np_cap) ++bits;     // the values are storage slots < np_cap
    const int dig = radix_digit(bits);
    A.job[0].bit0 = 32; A.job[0].passes = radix_passes(bits, dig);
    A.job[1] = A.job[0];
    A.select = nullptr;
    return launch_radix(p, A, dig, st);
}
