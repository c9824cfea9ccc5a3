// pm_deposit_tiles.cuh -- CIC deposit (src/density.py:7-48) by output tiles with exact fixed-point
// accumulation.  Included by pm_particles.cu (uses pm_cell).
//
// k_deposit_rows (pm_particles.cu) lets one warp own one output mesh row and visit the four source
// rows that feed it: every particle is loaded, binned and weighted FOUR times, by warps whose 32
// lanes are a third empty (a mesh row of the 256^3/512^3 run holds ~64 particles), and ncu shows the
// kernel bound by instruction issue, not by HBM (profiles/r01_notes.md: 323 M warp instructions,
// 0.28 of the HBM peak).  Here a CTA owns an output TILE of ZB planes x YB rows x the whole x extent:
//   * its source particles are the rows y0-1 .. y0+YB-1 of the planes Z0-1 .. Z0+ZB-1: per plane ONE
//     contiguous run of the cell-sorted list (two when y0-1 wraps), so a particle is visited
//     (1+1/ZB)(1+1/YB) times (1.69 for 2 x 8) by warps that take 32 consecutive entries at a time;
//   * each visit forms the particle's up to eight contributions exactly as the reference does
//     (mass*{t,d}_x*{t,d}_y*{t,d}_z, left to right, float64: density.py:24-47) and adds those that
//     fall inside the tile to shared-memory accumulators;
//   * the accumulators are 64-bit FIXED-POINT integers (2^-24 of a mass unit; held as two 32-bit
//     words because sm_100a has no native 64-bit shared-memory atomic add -- the low word is added
//     with ATOMS.ADD, whose returned old value tells whether a carry goes to the high word).  Integer
//     addition is associative, so the result does not depend on the order in which warps, CTAs or
//     work items arrive: bit-reproducible without ownership rules, and more accurate than the
//     reference's own float32 running sum (each cell is the exact sum of its contributions, each
//     rounded to 2^-24, rounded ONCE to float32; the reference rounds every partial sum).
//   * every mesh cell is written exactly once, zeros included (no memset pass).
// Skew (BASELINE configs[4]): a tile whose source runs hold more than kDepHeavy particles is not
// processed by its own CTA; the CTA cuts the concatenated runs into work items of at most kDepItem particles
// and queues them.  A second, persistent launch drains the queue -- any CTA takes any item, accumulates it
// in shared memory and adds the non-zero cells to the tile's 64-bit scratch slot in global memory
// (REDG.ADD.64, exact again); the CTA that finishes a tile's last item converts its scratch slot to float32
// mesh rows and zeroes it for the next step.  With a near-uniform load the second launch finds an empty queue.
#pragma once

constexpr uint32_t kDepHeavy = 16384;   // source particles above which a tile is split into work items (sweep on the z = 0
                                        // snapshot, profiles/r02_notes.md: 8192 -> 0.70 ms, 16384 -> 0.63, 32768 -> 0.69, never -> 0.84)
constexpr uint32_t kDepItem = 8192;     // particles per work item
constexpr int kDepThreads = 384;     // 12 warps share a tile: three CTAs of 64 KB (and 48 registers) per SM = 36 warps
constexpr int kDepMaxSlots = PM_DEP_MAX_SLOTS;   // heavy tiles per step that get a scratch slot (the rest run unsplit)
constexpr int kDepMaxItems = 16384;

struct DepItem {
    uint32_t slot, tile, beg, cnt;
};

struct DepositTileArgs {
    const float *px, *py, *pz;
    const uint32_t *order, *row_start;
    float *rho;
    int nc, nseg;             // mesh size; row_start is indexed by (row * nseg)
    int nz_out;               // output planes: nc, or nzl + 1 for a slab
    int z0, nzl, slab;
    int tiles_y, tiles_z;     // tile grid
    double smass, inv_scale;         // mass * 2^k (fixed-point units per unit weight) and 2^-k
    int fast_ok;                     // 0 <= mass * 2^k < 2^32: ordinary contributions fit 32 bits (pm_dep_accumulate)
    unsigned long long *scratch;     // [kDepMaxSlots][ZB*YB*nc] int64 sums of the heavy tiles (all zero between steps)
    uint32_t *ctl;                   // [0] slots used, [1] items queued, [2] items taken, [3] tiles that ran unsplit for lack of slots
    uint32_t *slot_tile;             // [kDepMaxSlots] tile of each slot
    uint32_t *slot_items;            // [kDepMaxSlots] work items queued for it
    uint32_t *slot_done;             // [kDepMaxSlots] work items finished (zeroed with ctl); the CTA that finishes the last one converts the slot
    DepItem *items;                  // [kDepMaxItems]
    uint32_t heavy, item;            // thresholds (kDepHeavy, kDepItem unless PM_DEP_HEAVY / PM_DEP_ITEM say otherwise)
};

// (hi:lo) += v, v a signed 64-bit fixed-point value.  Exact whatever the interleaving: the low words
// add modulo 2^32 and every wrap-around is seen by exactly one adder (its returned old value), which
// forwards it to the high word.
__device__ __forceinline__ void pm_fx_add(uint32_t *lo, uint32_t *hi, long long v)
{
    const uint32_t vlo = (uint32_t)v;
    uint32_t vhi = (uint32_t)((unsigned long long)v >> 32);
    if (vlo) {
        const uint32_t old = atomicAdd(lo, vlo);
        vhi += ((uint32_t)(old + vlo) < old) ? 1u : 0u;
    }
    if (vhi) atomicAdd(hi, vhi);
}

// the same for 0 <= v < 2^32 (an ordinary contribution: weight in [0, 1], mass * 2^k < 2^32)
__device__ __forceinline__ void pm_fx_add32(uint32_t *lo, uint32_t *hi, uint32_t v)
{
    const uint32_t old = atomicAdd(lo, v);
    if ((uint32_t)(old + v) < old) atomicAdd(hi, 1u);
}

__device__ __forceinline__ float pm_fx_to_float(uint32_t lo, uint32_t hi, double inv_scale)
{
    if (hi == 0u) return __fmul_rn(__uint2float_rn(lo), (float)inv_scale);   // one rounding; the scale is a power of two
    const long long v = (long long)(((unsigned long long)hi << 32) | lo);
    return (float)__dmul_rn((double)v, inv_scale);   // int64 -> float64 may round at 2^53 units: far below float32's ulp
}

// The contributions of the sorted entries [beg, end) that fall into tile (Z0, y0) are added to the
// shared-memory accumulators.  All threads of the CTA take part; 32 consecutive entries per warp.
//
// A.smass = mass * 2^k: the products are the reference's ((mass*w_x)*w_y)*w_z scaled by a power of
// two, which commutes with every rounding.  Ordinary batch (A.fast_ok and every |d| < 1.5 in the
// warp): each product is in [0, 2^32) fixed-point units and its integer value is the low word of
// (product + 1.5*2^52) -- one DADD, no conversion instruction; one ATOMS.ADD plus a carry test per
// corner.  Otherwise (a particle at x == N_CELLS, SURVEY Q4, has weights of size Nc and of either
// sign; or an unusual mass) the whole warp takes the general 64-bit route.
// The entries are given as up to NR runs [rb[k], re[k]) of the sorted list, walked as ONE sequence: a
// warp's batches stream through all of them with a single pipeline prologue and one ragged batch.
// NCT: the mesh width as a compile-time constant (0: read A.nc).  With the usual power-of-two meshes every row
// offset, wrap test and tile-row product folds into immediates and shifts -- the runtime-nc version executed
// 14 constant loads and ~60 integer multiply-adds per 32-particle batch on addressing alone (ncu source page,
// profiles/r02_notes.md).
template <int ZB, int YB, int NR, int NCT>
__device__ __forceinline__ void pm_dep_accumulate(const DepositTileArgs &A, const uint32_t *rb, const uint32_t *re, int nruns,
                                                  int Z0, int y0, uint32_t *s_lo, uint32_t *s_hi,
                                                  uint32_t vbeg = 0u, uint32_t vcnt = 0xffffffffu)
{
    // cumulative run lengths in registers; virtual index v -> sorted-list index
    uint32_t cum[NR + 1], rbeg[NR];
    cum[0] = 0;
#pragma unroll
    for (int k = 0; k < NR; ++k) {
        const uint32_t len = k < nruns ? re[k] - rb[k] : 0u;
        rbeg[k] = k < nruns ? rb[k] : 0u;
        cum[k + 1] = cum[k] + len;
    }
    // [vbeg, vbeg + vcnt) of the concatenated runs: the whole sequence for a tile's own CTA, a slice of it for a
    // work item of a crowded tile
    const uint32_t beg = vbeg < cum[NR] ? vbeg : cum[NR];
    const uint32_t end = vcnt < cum[NR] - beg ? beg + vcnt : cum[NR];
    auto list_index = [&](uint32_t v) -> uint32_t {
        uint32_t j = rbeg[0] + v;
#pragma unroll
        for (int k = 1; k < NR; ++k)
            if (v >= cum[k]) j = rbeg[k] + (v - cum[k]);
        return j;
    };
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = kDepThreads / 32;
    const int nc = NCT ? NCT : A.nc;
    const double magic = 6755399441055744.0;   // 1.5 * 2^52: low mantissa word = the integer part of the addend
    // two-deep software pipeline over a warp's batches: the permutation entry of batch b+2 and the position
    // of batch b+1 are in flight while batch b is computed (the order -> position chain of dependent loads
    // was what the warps waited for)
    const uint32_t step = nwarps * 32;
    uint32_t base = beg + warp * 32;
    uint32_t i_nxt = 0, i_nx2 = 0;
    float xn = 0.f, yn = 0.f, zn = 0.f;
    if (base + lane < end) i_nxt = A.order[list_index(base + lane)];
    if (base + step + lane < end) i_nx2 = A.order[list_index(base + step + lane)];
    if (base + lane < end) { xn = A.px[i_nxt]; yn = A.py[i_nxt]; zn = A.pz[i_nxt]; }
    for (; base < end; base += step) {
        const uint32_t j = base + lane;
        const bool valid = j < end;
        const float x = xn, y = yn, z = zn;
        {
            const uint32_t i1 = i_nx2;                                   // batch b+1: its entry is here by now
            if (j + 2 * step < end) i_nx2 = A.order[list_index(j + 2 * step)];   // batch b+2
            if (j + step < end) { xn = A.px[i1]; yn = A.py[i1]; zn = A.pz[i1]; }
        }
        uint32_t cell = 0xffffffffu - (uint32_t)lane;   // distinct dummy cells: idle lanes never join a run
        int xc = 0, x1 = 0;
        int off[4] = {0, 0, 0, 0};      // accumulator offset of tile row (cz, cy), -1: outside the tile
        double c[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) c[k] = 0.0;
        bool ordinary = true;
        if (valid) {
            xc = pm_cell(x, nc);
            const int yc = pm_cell(y, nc), zc = pm_cell(z, nc);
            x1 = xc + 1 == nc ? 0 : xc + 1;
            cell = ((uint32_t)(zc - A.z0) * nc + yc) * nc + xc;   // rank-local key: fits 32 bits (pm_api.cu)
            int ry = yc - y0;
            if (ry > YB) ry -= nc;                      // row nc-1 feeding tile row 0
            int rz = zc - A.z0 - Z0;
            if (!A.slab && rz > ZB) rz -= nc;           // plane nc-1 feeding tile plane 0 (periodic mesh only)
#pragma unroll
            for (int cz = 0; cz < 2; ++cz)
#pragma unroll
                for (int cy = 0; cy < 2; ++cy) {
                    const int pz_ = rz + cz, py_ = ry + cy;
                    off[cz * 2 + cy] = ((unsigned)pz_ < (unsigned)ZB && (unsigned)py_ < (unsigned)YB) ? (pz_ * YB + py_) * nc : -1;
                }
            // weights and products as density.py:24-47: float64, mass * w_x * w_y * w_z left to right
            const double d_x = (double)x - (double)xc, d_y = (double)y - (double)yc, d_z = (double)z - (double)zc;
            const double t_x = 1.0 - d_x, t_y = 1.0 - d_y, t_z = 1.0 - d_z;
            const double mx[2] = {__dmul_rn(A.smass, t_x), __dmul_rn(A.smass, d_x)};
            const double wy[2] = {t_y, d_y}, wz[2] = {t_z, d_z};
#pragma unroll
            for (int cz = 0; cz < 2; ++cz)
#pragma unroll
                for (int cy = 0; cy < 2; ++cy)
#pragma unroll
                    for (int cx = 0; cx < 2; ++cx) c[(cz * 2 + cy) * 2 + cx] = __dmul_rn(__dmul_rn(mx[cx], wy[cy]), wz[cz]);
            ordinary = fabs(d_x) < 1.5 && fabs(d_y) < 1.5 && fabs(d_z) < 1.5;
        }
        // Entries of one cell are consecutive (sorted list).  Where eight or more of them sit in the warp
        // (inside a halo), add the run up with a segmented scan first and let its last lane do the atomics
        // -- thousands of same-address atomics would serialise otherwise; shorter runs are cheaper left to
        // the hardware's conflict replay.  Integer sums: exact, hence the same result either way.
        const uint32_t prv = __shfl_up_sync(full, cell, 1);
        const unsigned heads = __ballot_sync(full, lane == 0 || prv != cell);
        const unsigned cont = ~heads;                       // lanes that continue their predecessor's run
        unsigned c7 = cont & (cont << 1);
        c7 &= c7 << 2;
        c7 &= (c7 << 3) & (cont << 6);                      // lanes whose 7 predecessors are in their run
        const bool runs = c7 != 0u;
        if (A.fast_ok && __all_sync(full, ordinary)) {
            uint32_t v[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) v[k] = (uint32_t)__double2loint(__dadd_rn(c[k], magic));
            if (!runs) {
                if (valid) {
#pragma unroll
                    for (int r = 0; r < 4; ++r)
                        if (off[r] >= 0) {
                            pm_fx_add32(s_lo + off[r] + xc, s_hi + off[r] + xc, v[2 * r]);
                            pm_fx_add32(s_lo + off[r] + x1, s_hi + off[r] + x1, v[2 * r + 1]);
                        }
                }
            } else {
                unsigned long long w[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) w[k] = v[k];
                unsigned m = cont;          // lanes whose d predecessors all continue their run
#pragma unroll 1
                for (int d = 1; m != 0u; d <<= 1) {
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const unsigned long long o = __shfl_up_sync(full, w[k], d);
                        if ((m >> lane) & 1u) w[k] += o;
                    }
                    m &= (m << d);
                }
                const bool tail = valid && (lane == 31 || ((heads >> (lane + 1)) & 1u));
                if (tail) {
#pragma unroll
                    for (int r = 0; r < 4; ++r)
                        if (off[r] >= 0) {
                            pm_fx_add(s_lo + off[r] + xc, s_hi + off[r] + xc, (long long)w[2 * r]);
                            pm_fx_add(s_lo + off[r] + x1, s_hi + off[r] + x1, (long long)w[2 * r + 1]);
                        }
                }
            }
        } else if (valid) {
            // general route, rare: full-range conversion, signed 64-bit adds, no aggregation
#pragma unroll
            for (int r = 0; r < 4; ++r)
                if (off[r] >= 0) {
                    pm_fx_add(s_lo + off[r] + xc, s_hi + off[r] + xc, __double2ll_rn(c[2 * r]));
                    pm_fx_add(s_lo + off[r] + x1, s_hi + off[r] + x1, __double2ll_rn(c[2 * r + 1]));
                }
        }
    }
}

// Source runs of tile (Z0, y0): for each source plane (tile planes and the one below), the rows
// y0-1 .. y0+YB-1 of the sorted list; run[k] = {beg, end}, at most 2 * (ZB + 1) runs.
template <int ZB, int YB>
__device__ __forceinline__ int pm_dep_runs(const DepositTileArgs &A, int Z0, int y0, uint32_t (&rb)[2 * (ZB + 1)],
                                           uint32_t (&re)[2 * (ZB + 1)])
{
    int n = 0;
    const int nc = A.nc, nseg = A.nseg;
    const int zsrc_max = A.slab ? A.nzl : nc;       // source planes that exist on this rank
#pragma unroll
    for (int dz = -1; dz < ZB; ++dz) {
        int zs = Z0 + dz;                            // plane index local to the rank
        if (zs < 0) {
            if (A.slab) continue;                    // rank-1's share arrives through the ghost plane
            zs += nc;
        }
        if (zs >= zsrc_max) continue;                // slab: plane nzl is output only
        const uint32_t rowbase = (uint32_t)zs * nc;
        if (y0 > 0) {
            rb[n] = A.row_start[(size_t)(rowbase + y0 - 1) * nseg];
            re[n] = A.row_start[(size_t)(rowbase + y0 + YB - 1) * nseg + nseg];
            ++n;
        } else {
            rb[n] = A.row_start[(size_t)(rowbase + nc - 1) * nseg];
            re[n] = A.row_start[(size_t)(rowbase + nc - 1) * nseg + nseg];
            ++n;
            rb[n] = A.row_start[(size_t)rowbase * nseg];
            re[n] = A.row_start[(size_t)(rowbase + YB - 1) * nseg + nseg];
            ++n;
        }
    }
    return n;
}

template <int ZB, int YB>
__device__ __forceinline__ void pm_dep_zero(uint32_t *s_lo, int words)
{
    uint4 *p4 = reinterpret_cast<uint4 *>(s_lo);
    for (int i = threadIdx.x; i < words / 4; i += kDepThreads) p4[i] = make_uint4(0u, 0u, 0u, 0u);
}

// Launch 1: one CTA per tile.
template <int ZB, int YB, int NCT>
__global__ void __launch_bounds__(kDepThreads) k_deposit_tiles(DepositTileArgs A)
{
    extern __shared__ uint4 s_dep4[];
    const int nc = NCT ? NCT : A.nc, cells = ZB * YB * nc;
    uint32_t *s_lo = reinterpret_cast<uint32_t *>(s_dep4), *s_hi = s_lo + cells;
    __shared__ uint32_t s_rb[2 * (ZB + 1)], s_re[2 * (ZB + 1)];
    __shared__ int s_mode;               // mode 0: accumulate here; 1: queued as work items
    __shared__ uint32_t s_slot;
    const int ty = blockIdx.x, tz = blockIdx.y;
    const int Z0 = tz * ZB, y0 = ty * YB;
    const uint32_t tile = (uint32_t)tz * A.tiles_y + ty;

    // the tile's source runs: slot k = 2 * (dz + 1) + half, looked up by 2 * (ZB + 1) threads at once
    constexpr int NR = 2 * (ZB + 1);
    if (threadIdx.x < NR) {
        const int k = threadIdx.x, dz = k / 2 - 1, half = k & 1;
        uint32_t b = 0, e = 0;
        int zs = Z0 + dz;
        bool have = true;
        if (zs < 0) {
            if (A.slab) have = false;                     // rank-1's share arrives through the ghost plane
            else zs += nc;
        }
        if (have && zs >= (A.slab ? A.nzl : nc)) have = false;    // slab: plane nzl is output only
        if (have) {
            const uint32_t rowbase = (uint32_t)zs * nc;
            const int nseg = A.nseg;
            if (y0 > 0) {
                if (half == 0) {
                    b = A.row_start[(size_t)(rowbase + y0 - 1) * nseg];
                    e = A.row_start[(size_t)(rowbase + y0 + YB - 1) * nseg + nseg];
                }
            } else if (half == 0) {
                b = A.row_start[(size_t)(rowbase + nc - 1) * nseg];
                e = A.row_start[(size_t)(rowbase + nc - 1) * nseg + nseg];
            } else {
                b = A.row_start[(size_t)rowbase * nseg];
                e = A.row_start[(size_t)(rowbase + YB - 1) * nseg + nseg];
            }
        }
        s_rb[k] = b;
        s_re[k] = e;
    }
    pm_dep_zero<ZB, YB>(s_lo, 2 * cells);                 // independent of the lookups above
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t total = 0, nitems = 0;
        for (int k = 0; k < NR; ++k) {
            total += s_re[k] - s_rb[k];
        }
        nitems = (total + A.item - 1) / A.item;             // slices of the concatenated runs, not of each run
        int mode = 0;
        if (total > A.heavy && A.scratch) {
            const uint32_t slot = atomicAdd(A.ctl + 0, 1u);
            uint32_t first = 0;
            bool ok = slot < (uint32_t)kDepMaxSlots;
            if (ok) {
                first = atomicAdd(A.ctl + 1, nitems);
                ok = first + nitems <= (uint32_t)kDepMaxItems;
                if (!ok) atomicSub(A.ctl + 1, nitems);   // nobody reads ctl[1] before the next launch
            }
            if (ok) {
                A.slot_tile[slot] = tile;
                A.slot_items[slot] = nitems;
                uint32_t w = first;
                for (uint32_t b = 0; b < total; b += A.item) {
                    DepItem it;
                    it.slot = slot; it.tile = tile; it.beg = b;          // virtual index into the tile's runs
                    it.cnt = total - b < A.item ? total - b : A.item;
                    A.items[w++] = it;
                }
                mode = 1;
            } else {
                atomicAdd(A.ctl + 3, 1u);                // statistics: ran unsplit
            }
        }
        s_mode = mode;
    }
    __syncthreads();
    if (s_mode == 1) return;                             // launches 2 and 3 produce this tile

    pm_dep_accumulate<ZB, YB, NR, NCT>(A, s_rb, s_re, NR, Z0, y0, s_lo, s_hi);
    __syncthreads();
    // write-out: every cell of the tile once, 16-byte stores
    const int quads_row = nc / 4, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll 1
    for (int row = warp; row < ZB * YB; row += kDepThreads / 32) {
        const int pz_ = row / YB, py_ = row % YB;      // compile-time divisors
        if (Z0 + pz_ >= A.nz_out) continue;
        const uint4 *lo4 = reinterpret_cast<const uint4 *>(s_lo + row * nc), *hi4 = reinterpret_cast<const uint4 *>(s_hi + row * nc);
        float4 *out4 = reinterpret_cast<float4 *>(A.rho + ((size_t)(Z0 + pz_) * nc + (y0 + py_)) * nc);
        for (int xq = lane; xq < quads_row; xq += 32) {
            const uint4 lo = lo4[xq], hi = hi4[xq];
            float4 o;
            if ((hi.x | hi.y | hi.z | hi.w) == 0u) {      // the usual case: four sums below 2^32 units
                const float is = (float)A.inv_scale;
                o.x = __fmul_rn(__uint2float_rn(lo.x), is);
                o.y = __fmul_rn(__uint2float_rn(lo.y), is);
                o.z = __fmul_rn(__uint2float_rn(lo.z), is);
                o.w = __fmul_rn(__uint2float_rn(lo.w), is);
            } else {
                o.x = pm_fx_to_float(lo.x, hi.x, A.inv_scale);
                o.y = pm_fx_to_float(lo.y, hi.y, A.inv_scale);
                o.z = pm_fx_to_float(lo.z, hi.z, A.inv_scale);
                o.w = pm_fx_to_float(lo.w, hi.w, A.inv_scale);
            }
            out4[xq] = o;
        }
    }
}

// Launch 2: persistent CTAs drain the work-item queue of the heavy tiles.
template <int ZB, int YB, int NCT>
__global__ void __launch_bounds__(kDepThreads) k_deposit_items(DepositTileArgs A)
{
    extern __shared__ uint4 s_dep4[];
    const int nc = NCT ? NCT : A.nc, cells = ZB * YB * nc;
    uint32_t *s_lo = reinterpret_cast<uint32_t *>(s_dep4), *s_hi = s_lo + cells;
    __shared__ uint32_t s_item;
    const uint32_t nitems = A.ctl[1];
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) s_item = atomicAdd(A.ctl + 2, 1u);
        __syncthreads();
        const uint32_t k = s_item;
        if (k >= nitems) return;
        const DepItem it = A.items[k];
        const int tz = it.tile / A.tiles_y, ty = it.tile - tz * A.tiles_y;
        pm_dep_zero<ZB, YB>(s_lo, 2 * cells);
        __syncthreads();
        {
            uint32_t rb[2 * (ZB + 1)], re[2 * (ZB + 1)];
            const int nr = pm_dep_runs<ZB, YB>(A, tz * ZB, ty * YB, rb, re);     // the order the tile's own CTA queued them in
            pm_dep_accumulate<ZB, YB, 2 * (ZB + 1), NCT>(A, rb, re, nr, tz * ZB, ty * YB, s_lo, s_hi, it.beg, it.cnt);
        }
        __syncthreads();
        unsigned long long *dst = A.scratch + (size_t)it.slot * cells;
        for (int c = threadIdx.x; c < cells; c += kDepThreads) {
            const uint32_t lo = s_lo[c], hi = s_hi[c];
            if (lo | hi) atomicAdd(dst + c, ((unsigned long long)hi << 32) | lo);
        }
        // The CTA that finishes a slot's LAST item turns the 64-bit sums into the tile's float32 mesh rows and
        // zeroes the slot for the next step (a separate pass over all slots did this before: 146 us on the
        // z = 0 snapshot).  Every CTA fences its adds before it counts itself done.
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) s_item = (atomicAdd(A.slot_done + it.slot, 1u) + 1u == A.slot_items[it.slot]) ? 1u : 0u;
        __syncthreads();
        if (s_item) {
            __threadfence();
            const int Z0 = tz * ZB, y0 = ty * YB;
            // eight independent loads in flight per thread (one load -> store chain per iteration made this
            // latency-bound: ~30 us per slot)
            constexpr int U = 8;
            for (int c0 = threadIdx.x; c0 < cells; c0 += U * kDepThreads) {
                unsigned long long v[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int c = c0 + u * kDepThreads;
                    v[u] = c < cells ? __ldcg(dst + c) : 0ull;
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int c = c0 + u * kDepThreads;
                    if (c < cells) {
                        const int row = c / nc, x = c - row * nc;
                        const int pz_ = row / YB, py_ = row % YB;
                        dst[c] = 0ull;
                        if (Z0 + pz_ < A.nz_out)
                            A.rho[((size_t)(Z0 + pz_) * nc + (y0 + py_)) * nc + x] =
                                pm_fx_to_float((uint32_t)v[u], (uint32_t)(v[u] >> 32), A.inv_scale);
                    }
                }
            }
        }
    }
}
