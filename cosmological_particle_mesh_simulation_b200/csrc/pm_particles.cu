// pm_particles.cu -- particle-side kernels of the PM step: cell keys, sort, row offsets,
// deterministic CIC deposit, fused force gather + kick + drift.
//
// Compiled with --fmad=false: the reference's numba/LLVM code never contracts a*b+c, and the
// gather/kick/drift arithmetic below is written to be bit-identical to it given the same phi
// (reference: src/integrate.py:27-97).  The explicit __f*_rn / __d*_rn intrinsics make the
// rounding points visible (and keep them if the flag is ever lost).
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "pm_internal.cuh"

// --------------------------------------------------------------------------------------------
// int(floor(x)) mod Nc with Python's sign convention (src/density.py:19-21, integrate.py:20).
// Positions live in [0, Nc] (integrate.py:95 wraps them; the float32 store can round up to
// exactly Nc -- SURVEY Q4 -- which must map to cell 0 while the offset d = x - cell stays Nc).
// --------------------------------------------------------------------------------------------
__device__ __forceinline__ int pm_cell(float x, int nc)
{
    int c = (int)floorf(x);
    if ((unsigned)c >= (unsigned)nc) {
        if (c == nc) {
            c = 0;
        } else {
            c %= nc;
            if (c < 0) c += nc;
        }
    }
    return c;
}

// --------------------------------------------------------------------------------------------
// Cell keys: key = (z_c*Nc + y_c)*Nc + x_c            (src/density.py:19-21,37)
// VEC: one thread takes 4 consecutive particles of each SoA row with 128-bit loads.
// --------------------------------------------------------------------------------------------
// Slab mode: the key uses the plane index local to this rank's slab [z0, z0+nzl); a particle
// outside the slab gets PM_KEY_DEAD, which sorts behind every real key.
#define PM_KEY_DEAD 0xffffffffu
__device__ __forceinline__ uint32_t pm_key(float x, float y, float z, int nc, int z0, int nzl)
{
    const int zl = pm_cell(z, nc) - z0;
    if ((unsigned)zl >= (unsigned)nzl) return PM_KEY_DEAD;
    return ((uint32_t)zl * nc + pm_cell(y, nc)) * nc + pm_cell(x, nc);
}

template <bool VEC>
__global__ void __launch_bounds__(256) k_cell_keys(const float *__restrict__ px,
                                                   const float *__restrict__ py,
                                                   const float *__restrict__ pz, int64_t np, int nc,
                                                   int z0, int nzl, uint32_t *__restrict__ keys,
                                                   uint32_t *__restrict__ order)
{
    // px == nullptr: keys of the mesh ROW only (x cell 0) -- the host-buffer step sorts by row and computes them
    // while the x coordinates are still on their way over PCIe
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (VEC) {
        int64_t i = t * 4;
        if (i >= np) return;
        float4 x = px ? *reinterpret_cast<const float4 *>(px + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        float4 y = *reinterpret_cast<const float4 *>(py + i);
        float4 z = *reinterpret_cast<const float4 *>(pz + i);
        uint4 k;
        k.x = pm_key(x.x, y.x, z.x, nc, z0, nzl);
        k.y = pm_key(x.y, y.y, z.y, nc, z0, nzl);
        k.z = pm_key(x.z, y.z, z.z, nc, z0, nzl);
        k.w = pm_key(x.w, y.w, z.w, nc, z0, nzl);
        *reinterpret_cast<uint4 *>(keys + i) = k;
        if (order) {
            uint32_t b = (uint32_t)i;
            *reinterpret_cast<uint4 *>(order + i) = make_uint4(b, b + 1, b + 2, b + 3);
        }
    } else {
        if (t >= np) return;
        keys[t] = pm_key(px ? px[t] : 0.f, py[t], pz[t], nc, z0, nzl);
        if (order) order[t] = (uint32_t)t;
    }
}

int pm_k_cell_keys(pm_plan *p, const float *pos, int64_t np, int64_t stride, uint32_t *keys,
                   uint32_t *order, cudaStream_t st, bool rows_only)
{
    if (np == 0) return PM_OK;
    const float *px = rows_only ? nullptr : pos, *py = pos + stride, *pz = pos + 2 * stride;
    bool vec = (np % 4 == 0) && (stride % 4 == 0) && ((uintptr_t)pos % 16 == 0) && ((uintptr_t)keys % 16 == 0) &&
               (!order || (uintptr_t)order % 16 == 0);
    if (vec) {
        int64_t nthr = np / 4;
        PM_LAUNCH(k_cell_keys<true>, (unsigned)((nthr + 255) / 256), 256, 0, st, px, py, pz, np,
                  p->nc, p->z0, p->nzl, keys, order);
    } else {
        PM_LAUNCH(k_cell_keys<false>, (unsigned)((np + 255) / 256), 256, 0, st, px, py, pz, np,
                  p->nc, p->z0, p->nzl, keys, order);
    }
    PM_CHECK_LAUNCH();
    return PM_OK;
}

// --------------------------------------------------------------------------------------------
// CIC deposit (src/density.py:7-48), deterministic and atomic-free.
//
// One warp owns one output mesh row (Z, y) held in shared memory (pm_acc_t, see below).  The row receives
// mass from the particles of four source rows, visited in a fixed order:
//     (Z, y) * t_z t_y,   (Z, y-1) * t_z d_y,   (Z-1, y) * d_z t_y,   (Z-1, y-1) * d_z d_y
// Each source row is a contiguous, x-sorted run of the cell-sorted particle list.  The warp walks
// it 32 particles at a time: every lane forms its two x-contributions exactly as the reference
// does (mass*{t,d}_x*{t,d}_y*{t,d}_z left to right in float64, density.py:37-47), a warp-segmented
// inclusive scan adds up lanes that share a cell, and the last lane of each cell run adds the run
// totals to bins x_c and x_c+1 in two conflict-free phases.  The summation tree is fixed by the
// sort order, so the result is bit-reproducible run to run; it differs from the reference's
// sequential float32 "+=" only by the order of the float32 additions.
// Every cell of rho is written exactly once, zeros included -- no memset pass.
// --------------------------------------------------------------------------------------------
// Accumulator type of the warp scan and of the shared-memory row.  float (default): every
// contribution is still the reference's float64 product, rounded once to float32 and summed in
// float32 like the reference's own grid (density.py:11,37) -- measured 9 % faster than double
// (half the shared-memory traffic and bank conflicts, half the shuffles).  -DPM_DEPOSIT_ACC=double
// keeps float64 running sums.
#ifndef PM_DEPOSIT_ACC
#define PM_DEPOSIT_ACC float
#endif
typedef PM_DEPOSIT_ACC pm_acc_t;
// 16-byte zero-fill and write-out of the shared-memory rows (float accumulators only)
#ifndef PM_DEPOSIT_VEC4
#define PM_DEPOSIT_VEC4 (sizeof(pm_acc_t) == 4)
#endif
__host__ __device__ __forceinline__ int pm_deposit_pitch(int xseg) { return (xseg + 1 + 3) & ~3; }

template <int RY>
__global__ void __launch_bounds__(RY * 32) k_deposit_rows(const float *__restrict__ px,
                                                          const float *__restrict__ py,
                                                          const float *__restrict__ pz,
                                                          const uint32_t *__restrict__ order,
                                                          const uint32_t *__restrict__ row_start,
                                                          int nc, int nseg, double mass,
                                                          float *__restrict__ rho, int z0, int nzl,
                                                          int slab)
{
    // slab != 0: this rank owns planes [z0, z0+nzl) and writes nzl+1 output planes; plane nzl is
    // the ghost that holds the d_z share of the top plane's particles (it belongs to rank+1), and
    // output plane 0 lacks the d_z share of rank-1's top plane until pm_k_ghost_add.  No z wrap.
    //
    // Wide meshes: a row is cut into nseg segments of xseg = nc/nseg cells and a warp owns one
    // segment of one output row, so shared memory per warp stays (xseg+1) doubles whatever nc is
    // (nseg = nc/512 for nc = 1024, 2048).  row_start is indexed by (row*nseg + segment).  A
    // particle in the last cell of a segment spills its d_x share into bin xseg, which the owner
    // of the next segment (same CTA) folds into its bin 0 after the barrier -- for nseg = 1 that
    // is the periodic wrap of the row onto itself.
    extern __shared__ pm_acc_t s_rows[];
    const unsigned full = 0xffffffffu;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int xseg = nc / nseg;
    const int rslot = warp / nseg, seg = warp - rslot * nseg;
    const int xs = seg * xseg;
    const int Z = blockIdx.y;
    const int y = blockIdx.x * (RY / nseg) + rslot;
    const bool active = y < nc;  // warp-uniform
    // row pitch: xseg bins + the spill bin, padded so every row starts on a 16-byte boundary
    const int pitch = pm_deposit_pitch(xseg);
    pm_acc_t *row = s_rows + (size_t)warp * pitch;
    if (PM_DEPOSIT_VEC4 && (xseg & 3) == 0) {
        float4 *row4 = reinterpret_cast<float4 *>(row);
        for (int x = lane; x < pitch / 4; x += 32) row4[x] = make_float4(0.f, 0.f, 0.f, 0.f);
    } else {
        for (int x = lane; x <= xseg; x += 32) row[x] = 0;
    }
    __syncwarp();

    const int Zm = slab ? Z - 1 : ((Z == 0) ? nc - 1 : Z - 1);
    const int ym = (y == 0) ? nc - 1 : y - 1;
#pragma unroll 1
    for (int src = 0; src < 4 && active; ++src) {
        const int zsl = (src & 2) ? Zm : Z;    // source plane, slab-local
        if (slab && (zsl < 0 || zsl >= nzl)) continue;
        const int zs = z0 + zsl;               // global cell coordinate of that plane
        const int ys = (src & 1) ? ym : y;
        const uint32_t r = ((uint32_t)zsl * nc + ys) * nseg + seg;
        const uint32_t beg = row_start[r], end = row_start[r + 1];
        for (uint32_t base = beg; base < end; base += 32) {
            const uint32_t j = base + lane;
            const bool valid = j < end;
            int xc = -1 - lane;  // distinct dummy cells: idle lanes never join a run
            pm_acc_t v0 = 0, v1 = 0;
            if (valid) {
                const uint32_t i = order[j];
                const float x = px[i], yy = py[i], zz = pz[i];
                xc = pm_cell(x, nc);
                const double d_x = (double)x - (double)xc;      // density.py:24-26 (float64)
                const double d_y = (double)yy - (double)ys;     // particle's y cell is ys by sort
                const double d_z = (double)zz - (double)zs;
                const double t_x = 1.0 - d_x;
                const double w_y = (src & 1) ? d_y : 1.0 - d_y;
                const double w_z = (src & 2) ? d_z : 1.0 - d_z;
                v0 = (pm_acc_t)__dmul_rn(__dmul_rn(__dmul_rn(mass, t_x), w_y), w_z);
                v1 = (pm_acc_t)__dmul_rn(__dmul_rn(__dmul_rn(mass, d_x), w_y), w_z);
            }
            // Warp-segmented inclusive scan over runs of equal x cell.  The list is sorted, so a
            // run is a stretch of lanes that are not "heads"; the ballot of heads tells the whole
            // warp how long the longest run is, and the scan stops after ceil(log2(longest)) steps
            // (0 or 1 for a near-uniform particle load, 5 inside a halo).  m holds, for step d,
            // the lanes whose d predecessors all continue their run (m_2d = m_d & (m_d << d)).
            const int prv = __shfl_up_sync(full, xc, 1);
            const unsigned heads = __ballot_sync(full, lane == 0 || prv != xc);
            unsigned m = ~heads;
#pragma unroll 1
            for (int d = 1; m != 0; d <<= 1) {
                const pm_acc_t o0 = __shfl_up_sync(full, v0, d);
                const pm_acc_t o1 = __shfl_up_sync(full, v1, d);
                if ((m >> lane) & 1u) {
                    v0 += o0;
                    v1 += o1;
                }
                m &= (m << d);
            }
            const bool tail = valid && (lane == 31 || ((heads >> (lane + 1)) & 1u));
            const int lb = xc - xs;   // bin inside the segment; lb + 1 may be the spill bin xseg
            if (tail) row[lb] += v0;
            __syncwarp();
            if (tail) row[lb + 1] += v1;
            __syncwarp();
        }
    }
    __syncthreads();
    if (!active) return;
    const int pseg = (seg == 0) ? nseg - 1 : seg - 1;
    const pm_acc_t spill = s_rows[(size_t)(rslot * nseg + pseg) * pitch + xseg];
    float *out = rho + ((size_t)Z * nc + y) * nc + xs;
    if (PM_DEPOSIT_VEC4 && (xseg & 3) == 0 && (nc & 3) == 0) {
        // 16-byte row reads and mesh stores: a quarter of the instructions of the scalar loop
        const float4 *row4 = reinterpret_cast<const float4 *>(row);
        float4 *out4 = reinterpret_cast<float4 *>(out);
        for (int x = lane; x < xseg / 4; x += 32) {
            float4 v = row4[x];
            if (x == 0) v.x += spill;
            out4[x] = v;
        }
    } else {
        for (int x = lane; x < xseg; x += 32) out[x] = (float)(x == 0 ? row[0] + spill : row[x]);
    }
}

// warps (output row segments) per CTA: 4 measured best at 512^3 (0.406 ms; 8: 0.429, 16: 0.460);
// a CTA must hold all segments of a row, so meshes cut into more than 4 segments use 8
#ifndef PM_DEPOSIT_RY
#define PM_DEPOSIT_RY 4
#endif

int pm_deposit_segments(int nc)
{
    // segments per mesh row in the deposit: 512-cell segments for wide power-of-two meshes.
    // PM_DEPOSIT_XSEG=<cells> overrides the segment length (tests exercise the segmented path on
    // small meshes with it).
    if (const char *e = getenv("PM_DEPOSIT_XSEG")) {
        const int xseg = atoi(e);
        if (xseg > 0 && nc % xseg == 0) {
            const int s = nc / xseg;
            if (s == 1 || s == 2 || s == 4 || s == 8) return s;
        }
    }
    if (nc >= 1024 && nc % 512 == 0) {
        const int s = nc / 512;
        if (s == 2 || s == 4 || s == 8) return s;
    }
    return 1;
}

template <int RY>
static int pm_launch_deposit_ry(pm_plan *p, const float *pos, int64_t stride, double mass, float *rho,
                                int nz_out, int slab, cudaStream_t st)
{
    const int nc = p->nc, nseg = p->dep_nseg;
    const int rows_per_cta = RY / nseg;
    const size_t smem = (size_t)RY * pm_deposit_pitch(nc / nseg) * sizeof(pm_acc_t);
    static size_t smem_set_dev[64];   // per device ordinal (function attributes are per device)
    size_t &smem_set = smem_set_dev[p->device & 63];
    if (smem > smem_set) {
        PM_CUDA(cudaFuncSetAttribute(k_deposit_rows<RY>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        smem_set = smem;
    }
    dim3 grid((nc + rows_per_cta - 1) / rows_per_cta, nz_out);
    PM_LAUNCH(k_deposit_rows<RY>, grid, RY * 32, smem, st, pos, pos + stride, pos + 2 * stride,
              p->order_sorted, p->row_start, nc, nseg, mass, rho, p->z0, p->nzl, slab);
    PM_CHECK_LAUNCH();
    return PM_OK;
}

#include "pm_deposit_tiles.cuh"

template <int ZB, int YB, int NCT>
static int pm_launch_deposit_tiles_zy(pm_plan *p, const float *pos, int64_t stride, double mass, float *rho,
                                      int nz_out, int slab, cudaStream_t st)
{
    const int nc = p->nc;
    const size_t smem = (size_t)2 * ZB * YB * nc * sizeof(uint32_t);
    auto k1 = k_deposit_tiles<ZB, YB, NCT>;
    auto k2 = k_deposit_items<ZB, YB, NCT>;
    PM_ONCE_PER_DEVICE_BEGIN(p->device)
        // the largest tile any mesh gives this instantiation (8192 cells), not this call's
        PM_CUDA(cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 8192 * (int)sizeof(uint32_t)));
        PM_CUDA(cudaFuncSetAttribute(k2, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 8192 * (int)sizeof(uint32_t)));
        PM_CUDA(cudaFuncSetAttribute(k1, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        PM_CUDA(cudaFuncSetAttribute(k2, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    PM_ONCE_PER_DEVICE_END()
    DepositTileArgs A;
    A.px = pos; A.py = pos + stride; A.pz = pos + 2 * stride;
    A.order = p->order_sorted; A.row_start = p->row_start; A.rho = rho;
    A.nc = nc; A.nseg = p->dep_nseg; A.nz_out = nz_out;
    A.z0 = slab ? p->z0 : 0; A.nzl = p->nzl; A.slab = slab;
    A.tiles_y = nc / YB; A.tiles_z = (nz_out + ZB - 1) / ZB;
    // fixed-point unit 2^-k of a mass unit: k = 24 unless |mass| * Nc^3 (every axis at x == N_CELLS) would
    // not leave 2^8 such terms of headroom in 63 bits
    int k = 24;
    {
        const double big = fabs(mass) * (double)nc * nc * nc;
        if (big > 0.0 && isfinite(big)) {
            const int need = (int)ceil(log2(big));
            if (55 - need < k) k = 55 - need;
        }
    }
    A.smass = ldexp(mass, k); A.inv_scale = ldexp(1.0, -k);
    A.fast_ok = (A.smass >= 0.0 && A.smass < 4294967296.0) ? 1 : 0;
    A.scratch = p->dep_scratch; A.ctl = p->dep_ctl; A.items = (DepItem *)p->dep_items;
    {
        static const uint32_t heavy = getenv("PM_DEP_HEAVY") ? (uint32_t)atoi(getenv("PM_DEP_HEAVY")) : kDepHeavy;
        static const uint32_t item = getenv("PM_DEP_ITEM") ? (uint32_t)atoi(getenv("PM_DEP_ITEM")) : kDepItem;
        A.heavy = heavy < 256 ? 256 : heavy;
        A.item = item < 256 ? 256 : item;
    }
    A.slot_done = p->dep_ctl + 16; A.slot_tile = p->dep_slot_tile; A.slot_items = p->dep_slot_tile + PM_DEP_MAX_SLOTS;
    PM_CUDA(cudaMemsetAsync(p->dep_ctl, 0, (16 + PM_DEP_MAX_SLOTS) * sizeof(uint32_t), st));   // counters + slot_done
    dim3 grid(A.tiles_y, A.tiles_z);
    PM_LAUNCH(k1, grid, kDepThreads, smem, st, A);
    PM_LAUNCH(k2, p->sm_count * 3, kDepThreads, smem, st, A);
    PM_CHECK_LAUNCH();
    return PM_OK;
}

// tile shape by mesh width: at most 8192 cells (64 KB of accumulators) per tile
static int pm_launch_deposit_tiles(pm_plan *p, const float *pos, int64_t stride, double mass, float *rho,
                                   int nz_out, int slab, cudaStream_t st)
{
    const int nc = p->nc;
    if (nc % 16 || nc < 16) return PM_ERR_UNSUPPORTED;
    switch (nc) {       // the usual meshes: width folded into the code (pm_deposit_tiles.cuh, NCT)
    case 128: return pm_launch_deposit_tiles_zy<2, 8, 128>(p, pos, stride, mass, rho, nz_out, slab, st);
    case 256: return pm_launch_deposit_tiles_zy<2, 8, 256>(p, pos, stride, mass, rho, nz_out, slab, st);
    case 512: return pm_launch_deposit_tiles_zy<2, 8, 512>(p, pos, stride, mass, rho, nz_out, slab, st);
    case 1024: return pm_launch_deposit_tiles_zy<2, 4, 1024>(p, pos, stride, mass, rho, nz_out, slab, st);
    case 2048: return pm_launch_deposit_tiles_zy<2, 2, 2048>(p, pos, stride, mass, rho, nz_out, slab, st);
    default: break;
    }
    if (nc <= 512) return pm_launch_deposit_tiles_zy<2, 8, 0>(p, pos, stride, mass, rho, nz_out, slab, st);
    if (nc <= 1024) return pm_launch_deposit_tiles_zy<2, 4, 0>(p, pos, stride, mass, rho, nz_out, slab, st);
    if (nc <= 2048) return pm_launch_deposit_tiles_zy<2, 2, 0>(p, pos, stride, mass, rho, nz_out, slab, st);
    return PM_ERR_UNSUPPORTED;
}

static int pm_launch_deposit(pm_plan *p, const float *pos, int64_t stride, double mass, float *rho,
                             int nz_out, int slab, cudaStream_t st)
{
    if (p->deposit_tiles) {
        const int rc = pm_launch_deposit_tiles(p, pos, stride, mass, rho, nz_out, slab, st);
        if (rc != PM_ERR_UNSUPPORTED) return rc;
    }
    if (p->dep_nseg <= PM_DEPOSIT_RY && PM_DEPOSIT_RY % p->dep_nseg == 0)
        return pm_launch_deposit_ry<PM_DEPOSIT_RY>(p, pos, stride, mass, rho, nz_out, slab, st);
    return pm_launch_deposit_ry<8>(p, pos, stride, mass, rho, nz_out, slab, st);
}

int pm_k_deposit(pm_plan *p, const float *pos, int64_t stride, double mass, float *rho,
                 cudaStream_t st)
{
    return pm_launch_deposit(p, pos, stride, mass, rho, p->nc, 0, st);
}

// Slab variant on the resident state: nzl+1 output planes into rho[(nzl+1)][nc][nc].
int pm_k_deposit_slab(pm_plan *p, const float *pos, double mass, float *rho, cudaStream_t st)
{
    return pm_launch_deposit(p, pos, p->rstride, mass, rho, p->nzl + 1, 1, st);
}

// --------------------------------------------------------------------------------------------
// Fused force gather + kick + drift (src/integrate.py:15-97), one thread per particle.
//
// Per particle: cell and CIC weights from the PRE-step position for all three directions
// (integrate.py:20-24), 32 distinct phi loads (the 8 corners and their +-1 neighbours
// along each axis), then per direction
//     g_p  = (sum_c g_c * t_c) / 2          float32 products/sums left to right (integrate.py:92)
//     v   += da*f_a1*g_p                    float64, stored float32          (integrate.py:94)
//     x    = (x + da*v/(a+da)^2*f_a1) % Nc  float64 with the stored v, stored float32 (:95)
// The weight table t[8,Np], the int64 cell table and its six per-direction copies that the
// reference materialises never exist here.
// --------------------------------------------------------------------------------------------
__device__ __forceinline__ double pm_pymod(double a, double n)
{
    // Python float %: fmod then lift negatives by n.  Fast exact paths for the usual ranges.
    if (a >= 0.0) {
        if (a < n) return a;
        if (a < 2.0 * n) return a - n;  // exact (Sterbenz)
        return fmod(a, n);
    }
    double r = (a >= -n) ? a : fmod(a, n);
    return (r < 0.0) ? __dadd_rn(r, n) : fabs(r);
}

template <int DIR>
__device__ __forceinline__ float pm_gp(const float (&v)[4][4][4], const float (&t)[8])
{
    // corner (oz,oy,ox): g = -phi[c+o+e_DIR] + phi[c+o-e_DIR]        (integrate.py:84-91)
#define PM_G(oz, oy, ox)                                                                   \
    __fsub_rn(v[1 + oz - (DIR == 2)][1 + oy - (DIR == 1)][1 + ox - (DIR == 0)],            \
              v[1 + oz + (DIR == 2)][1 + oy + (DIR == 1)][1 + ox + (DIR == 0)])
    float s = __fmul_rn(PM_G(0, 0, 0), t[0]);           // g    * ttt
    s = __fadd_rn(s, __fmul_rn(PM_G(0, 0, 1), t[1]));   // g_x  * dtt
    s = __fadd_rn(s, __fmul_rn(PM_G(0, 1, 0), t[2]));   // g_y  * tdt
    s = __fadd_rn(s, __fmul_rn(PM_G(1, 0, 0), t[3]));   // g_z  * ttd
    s = __fadd_rn(s, __fmul_rn(PM_G(0, 1, 1), t[4]));   // g_xy * ddt
    s = __fadd_rn(s, __fmul_rn(PM_G(1, 0, 1), t[5]));   // g_xz * dtd
    s = __fadd_rn(s, __fmul_rn(PM_G(1, 1, 0), t[6]));   // g_yz * tdd
    s = __fadd_rn(s, __fmul_rn(PM_G(1, 1, 1), t[7]));   // g_xyz* ddd
#undef PM_G
    return s;
}

// x / b for a divisor that is the same for every particle of a launch: with r = RN(1/b) formed on
// the host, q = RN(x*r) is within an ulp of the quotient, e = x - b*q is exact in one FMA, and
// RN(q + e*r) is the correctly rounded x/b (Markstein's final division step) -- three FP64
// instructions instead of the ~35 of the IEEE division sequence, same bits.  Valid away from
// overflow/underflow (pm_div_rcp checks the ranges on the host and returns 0 to ask for __ddiv_rn);
// the CPU test suite replays the sequence in exact rational arithmetic (constant-divisor test).
__device__ __forceinline__ double pm_div_const(double x, double b, double r)
{
    if (r == 0.0) return __ddiv_rn(x, b);   // launch-uniform
    const double q = __dmul_rn(x, r);
    const double e = __fma_rn(-b, q, x);
    return __fma_rn(e, r, q);
}

static double pm_div_rcp(double b, double da)
{
    const double ab = fabs(b), ad = fabs(da);
    if (!(ab > 1e-60 && ab < 1e60) || !(ad < 1e60) || (ad != 0.0 && ad < 1e-60)) return 0.0;
    return 1.0 / b;
}

__device__ __forceinline__ void pm_push(float &x, float &vel, float s, double k_kick, double da,
                                        double aa, double raa, double f_a1, int nc, float *acc_out)
{
    const double g_p = (double)s / 2.0;
    if (acc_out) *acc_out = (float)g_p;
    vel = (float)__dadd_rn((double)vel, __dmul_rn(k_kick, g_p));
    const double step = __dmul_rn(pm_div_const(__dmul_rn(da, (double)vel), aa, raa), f_a1);
    x = (float)pm_pymod(__dadd_rn((double)x, step), (double)nc);
}

// PERM = false: stateless call, particle i updated in place (pos_in == pos_out).
// PERM = true : resident state.  Slot s of the NEW cell order takes particle j = perm[s] of the
//               current buffers (a near-coalesced gather: the current order is last step's cell
//               order), and writes position, velocity, id and the cell key of the NEW position to
//               slot s of the other buffer set -- the permutation costs no pass of its own.
// SLAB        : (implies PERM) phi is this rank's (nzl+3)-plane buffer whose plane 0 is global
//               plane z0-1, so z needs no wrap; a particle whose new z cell belongs to another
//               rank gets PM_KEY_DEAD and its slot is appended to that rank's leave list.
struct SlabArgs {
    int z0, nzl, rank;
    const uint32_t *np_valid;  // live particle count (row_start[nzl*nc]) -- no host round trip
    uint32_t *leave_cnt, *leave_slot;
    int64_t leave_cap;
};

// 128-thread CTAs at 64 registers (8 per SM): measured 0.549 ms at 256^3/512^3 against 0.566 ms
// for 256-thread CTAs; 72 registers (3 x 256 threads per SM) 0.686 ms
#ifndef PM_GATHER_THREADS
#define PM_GATHER_THREADS 128
#endif
// how the resident gather learns the key a slot was filed under when it counts the movers:
// 0 = keep it in a register from the cell computation (measured 0.584 ms), 1 = reload
// keys_sorted[i] at the end (0.606 ms: a dependent load at the end of every warp's life)
#ifndef PM_GATHER_KOLD_LOAD
#define PM_GATHER_KOLD_LOAD 0
#endif
#ifndef PM_GATHER_MINB
#define PM_GATHER_MINB 8
#endif
template <bool PERM, bool SLAB, bool KGRAD = false>
__global__ void __launch_bounds__(PM_GATHER_THREADS, PM_GATHER_MINB) k_gather_kick_drift(
    const float *pos_in, const float *vel_in,   // may alias pos_out/vel_out when !PERM
    const uint32_t *__restrict__ id_in, const uint32_t *__restrict__ perm,
    float *pos_out, float *vel_out, uint32_t *__restrict__ id_out,
    uint32_t *__restrict__ keys_out, int64_t np, int64_t sin, int64_t sout,
    const float *__restrict__ phi, int nc, double k_kick, double da, double aa, double raa,
    double f_a1, float *__restrict__ acc, SlabArgs sl, uint32_t *__restrict__ mover_cnt,
    const uint32_t *__restrict__ keys_old, const float *__restrict__ fmesh = nullptr)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= np) return;
    if (SLAB && i >= (int64_t)*sl.np_valid) return;
    const int64_t j = PERM ? (int64_t)perm[i] : i;
    float x = pos_in[j], y = pos_in[sin + j], z = pos_in[2 * sin + j];
    float vx = vel_in[j], vy = vel_in[sin + j], vz = vel_in[2 * sin + j];

    const int xc = pm_cell(x, nc), yc = pm_cell(y, nc), zc = pm_cell(z, nc);
#if !PM_GATHER_KOLD_LOAD
    const uint32_t kold = ((uint32_t)zc * nc + yc) * nc + xc;   // the key slot i was filed under (!SLAB)
#endif
    // weights (integrate.py:36-51): float64 products left to right, stored float32
    const double d_x = (double)x - (double)xc, d_y = (double)y - (double)yc,
                 d_z = (double)z - (double)zc;
    const double t_x = 1.0 - d_x, t_y = 1.0 - d_y, t_z = 1.0 - d_z;
    float t[8];
    t[0] = (float)__dmul_rn(__dmul_rn(t_x, t_y), t_z);
    t[1] = (float)__dmul_rn(__dmul_rn(d_x, t_y), t_z);
    t[2] = (float)__dmul_rn(__dmul_rn(t_x, d_y), t_z);
    t[3] = (float)__dmul_rn(__dmul_rn(t_x, t_y), d_z);
    t[4] = (float)__dmul_rn(__dmul_rn(d_x, d_y), t_z);
    t[5] = (float)__dmul_rn(__dmul_rn(d_x, t_y), d_z);
    t[6] = (float)__dmul_rn(__dmul_rn(t_x, d_y), d_z);
    t[7] = (float)__dmul_rn(__dmul_rn(d_x, d_y), d_z);

    // periodic neighbour indices c-1, c, c+1, c+2 (integrate.py:64-65,77-82)
    uint32_t xo[4], yo[4], zo[4];
    {
        const int n = nc;
        int a1 = xc + 1 == n ? 0 : xc + 1;
        xo[0] = xc == 0 ? n - 1 : xc - 1; xo[1] = xc; xo[2] = a1; xo[3] = a1 + 1 == n ? 0 : a1 + 1;
        int b1 = yc + 1 == n ? 0 : yc + 1;
        yo[0] = yc == 0 ? n - 1 : yc - 1; yo[1] = yc; yo[2] = b1; yo[3] = b1 + 1 == n ? 0 : b1 + 1;
        if (SLAB) {
            const int p0 = zc - sl.z0;   // buffer plane of global plane zc-1
            zo[0] = p0; zo[1] = p0 + 1; zo[2] = p0 + 2; zo[3] = p0 + 3;
        } else {
            int c1 = zc + 1 == n ? 0 : zc + 1;
            zo[0] = zc == 0 ? n - 1 : zc - 1; zo[1] = zc; zo[2] = c1; zo[3] = c1 + 1 == n ? 0 : c1 + 1;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            yo[k] *= (uint32_t)n;
            zo[k] *= (uint32_t)n * (uint32_t)n;
        }
    }
    float sx, sy, sz;
    if (KGRAD) {
        // option kspace_gradient (pm_plan_set_poisson_options): fmesh[d] holds 2 * (-dphi/dx_d) from the
        // spectral derivative; interpolate it with the same eight weights, in the order of integrate.py:92
        const size_t cells = (size_t)nc * nc * nc;
        const uint32_t cidx[8] = {zo[1] + yo[1] + xo[1], zo[1] + yo[1] + xo[2], zo[1] + yo[2] + xo[1], zo[2] + yo[1] + xo[1],
                                  zo[1] + yo[2] + xo[2], zo[2] + yo[1] + xo[2], zo[2] + yo[2] + xo[1], zo[2] + yo[2] + xo[2]};
        float s3[3];
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            const float *fm = fmesh + d * cells;
            float s = __fmul_rn(__ldg(fm + cidx[0]), t[0]);
#pragma unroll
            for (int c = 1; c < 8; ++c) s = __fadd_rn(s, __fmul_rn(__ldg(fm + cidx[c]), t[c]));
            s3[d] = s;
        }
        sx = s3[0]; sy = s3[1]; sz = s3[2];
    } else {
    // the 32 cells with at most one "outer" (0 or 3) coordinate
    float v[4][4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b)
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int outer = (a == 0 || a == 3) + (b == 0 || b == 3) + (c == 0 || c == 3);
                v[a][b][c] = (outer <= 1) ? __ldg(phi + (zo[a] + yo[b] + xo[c])) : 0.0f;
            }

    sx = pm_gp<0>(v, t); sy = pm_gp<1>(v, t); sz = pm_gp<2>(v, t);
    }
    pm_push(x, vx, sx, k_kick, da, aa, raa, f_a1, nc, acc ? acc + i : nullptr);
    pm_push(y, vy, sy, k_kick, da, aa, raa, f_a1, nc, acc ? acc + np + i : nullptr);
    pm_push(z, vz, sz, k_kick, da, aa, raa, f_a1, nc, acc ? acc + 2 * np + i : nullptr);

    pos_out[i] = x; pos_out[sout + i] = y; pos_out[2 * sout + i] = z;
    vel_out[i] = vx; vel_out[sout + i] = vy; vel_out[2 * sout + i] = vz;
    if (PERM) {
        id_out[i] = id_in[j];
        if (SLAB) {
            const int dest = pm_cell(z, nc) / sl.nzl;
            uint32_t key = PM_KEY_DEAD;
            if (dest == sl.rank) {
                key = pm_key(x, y, z, nc, sl.z0, sl.nzl);
            } else {
                const uint32_t k = atomicAdd(sl.leave_cnt + dest, 1u);
                if ((int64_t)k < sl.leave_cap) sl.leave_slot[(int64_t)dest * sl.leave_cap + k] = (uint32_t)i;
            }
            keys_out[i] = key;
        } else {
            const uint32_t knew = pm_key(x, y, z, nc, 0, nc);
            keys_out[i] = knew;
            if (mover_cnt) {
                // The incremental sort's first pass, for free: slot i was filed under the key of the
                // pre-step position, so it is a "mover" iff the key changed.  A warp's 32 slots lie
                // in one sort tile; integer atomics keep the counts exact in any order.
#if PM_GATHER_KOLD_LOAD
                const uint32_t kold = keys_old[i];   // the key slot i was filed under
#endif
                const unsigned act = __activemask();
                const unsigned mv = __ballot_sync(act, knew != kold);
                if (mv && (threadIdx.x & 31) == __ffs(act) - 1) atomicAdd(mover_cnt + i / PM_SORT_TILE, __popc(mv));
            }
        }
    }
}

int pm_k_gather_kick_drift(pm_plan *p, float *pos, float *vel, int64_t np, const float *phi,
                           double a_val, double f_a1, double da, float *acc, cudaStream_t st)
{
    if (np == 0) return PM_OK;
    // host scalars evaluated exactly as integrate.py:94-95 does: da*f_a1, (a_val+da)**2
    const double k_kick = da * f_a1;
    const double aa = (a_val + da) * (a_val + da);
    auto kern = p->kgrad ? k_gather_kick_drift<false, false, true> : k_gather_kick_drift<false, false, false>;
    PM_LAUNCH(kern, (unsigned)((np + PM_GATHER_THREADS - 1) / PM_GATHER_THREADS), PM_GATHER_THREADS, 0, st, pos, vel,
              (const uint32_t *)nullptr, (const uint32_t *)nullptr, pos, vel, (uint32_t *)nullptr,
              (uint32_t *)nullptr, np, np, np, phi, p->nc, k_kick, da, aa, pm_div_rcp(aa, da), f_a1, acc, SlabArgs(), (uint32_t *)nullptr, (const uint32_t *)nullptr,
              (const float *)p->fmesh);
    PM_CHECK_LAUNCH();
    return PM_OK;
}

#include "pm_gather_tiled.cuh"
#include "pm_gather_ws.cuh"

#ifndef PM_GT_YB
#define PM_GT_YB 4      // particle rows per CTA
#endif
#ifndef PM_GT_NT
#define PM_GT_NT 384    // threads: one round covers the ~64*YB particles of a step with margin; 384 measured faster than 320, 288
#endif
#ifndef PM_GT_MINB
#define PM_GT_MINB 2    // CTAs per SM the ring is sized for
#endif
#ifndef PM_GT_CAP
#define PM_GT_CAP 384   // staged particles per step
#endif

// Tiled variant (pm_gather_tiled.cuh) when the mesh and the plan allow it; returns PM_ERR_UNSUPPORTED
// otherwise so that the caller falls back to the one-thread-per-particle kernel.
// warp-specialised variant (pm_gather_ws.cuh): consumer warps, phi ring depth, staging ring depth
#ifndef PM_GW_CW
#define PM_GW_CW 8
#endif
#ifndef PM_GW_R
#define PM_GW_R 5
#endif
#ifndef PM_GW_S
#define PM_GW_S 3
#endif
#ifndef PM_GW_PP
#define PM_GW_PP 1      // particles per consumer lane and batch (pm_gather_ws.cuh)
#endif

// ---- work list of the warp-specialised gather ------------------------------------------------------
// With a fixed grid every CTA marches zc planes of one row block whatever they hold.  On an evolved
// particle distribution (BASELINE configs[4]) a few (row block, chunk) columns inside haloes hold 10-100
// times the mean and their CTAs finish long after the rest of the device has drained: 1.2 ms instead of
// 0.53 ms on the z = 0 snapshot (profiles/r02_m_bench_main_evolved.json).  k_gather_items, one thread per
// base chunk, reads the chunk's plane counts from the row table and files either the chunk as it is
// ("light", taken from the back of the list) or -- when it holds more than T particles -- pieces of at
// most ~T particles ("heavy", from the front, so that they are dispatched first): runs of planes, and
// for a single plane above T, ranges of H = T/2 of its particles.  T = twice the mean chunk load (at
// least 4096), so a near-uniform load files every chunk unchanged.  The list order depends on atomics;
// the result does not (every particle belongs to exactly one item, all cross-item sums are integer).
static uint32_t pm_gather_item_threshold(int nc, int64_t np, int zc0)
{
    const int64_t chunks = (int64_t)(nc / PM_GT_YB) * (nc / zc0);
    static const double factor = getenv("PM_GATHER_T_FACTOR") ? atof(getenv("PM_GATHER_T_FACTOR")) : 2.0;
    const int64_t t = chunks > 0 ? (int64_t)((factor < 1.0 ? 1.0 : factor) * (double)(np / chunks)) : np;
    return (uint32_t)(t < 4096 ? 4096 : (t > 0x7fffffff ? 0x7fffffff : t));
}

static int pm_gather_base_zc(int nc, int sm_count)
{
    int zc = 32;
    while (zc > 1 && (nc % zc || (int64_t)(nc / PM_GT_YB) * (nc / zc) < 8LL * sm_count)) zc /= 2;
    return zc;
}

// List length bound for `chunks` base chunks and threshold T.  A chunk is filed whole (1 item) or in pieces:
// run pieces cut by "acc + c > T" pair up with their successor to more than T particles, so a chunk holding n
// particles yields at most 2n/T + 1 of them (the +1 is the closing piece); one flush in front of every plane
// above T (<= n/T); range pieces of H = T/2 particles, the last of a plane shorter (<= 2n/T + n/T).  Per chunk
// <= 1 + 6n/T, in total <= chunks + 6 np/T.
static int64_t pm_gather_item_bound(int64_t chunks, int64_t np, uint32_t T) { return chunks + 6 * (np / T) + 64; }

int pm_gather_item_capacity(int nc, int64_t np)
{
    if (nc % PM_GT_YB) return 64;
    // room for any chunk size down to one plane and the smallest threshold
    int64_t cap = pm_gather_item_bound((int64_t)(nc / PM_GT_YB) * nc, np, 4096);
    if (cap > (1 << 22)) cap = 1 << 22;
    return (int)cap;
}

template <int NC, int YB>
__global__ void __launch_bounds__(128) k_gather_items(const uint32_t *__restrict__ row_start, int zc0, uint32_t T,
                                                      GatherItem *__restrict__ items, int cap, uint32_t *__restrict__ ctl)
{
    constexpr int NYB = NC / YB;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= NYB * (NC / zc0)) return;
    const int yb = idx % NYB, zs = (idx / NYB) * zc0, y0 = yb * YB;
    const uint32_t H = T / 2;
    auto emit = [&](bool heavy, int z, int nz, uint32_t b, uint32_t e) {
        const uint32_t k = atomicAdd(ctl + (heavy ? 0 : 1), 1u);
        const uint32_t other = ctl[heavy ? 1 : 0];
        if (k + other + 1u >= (uint32_t)cap) {          // cannot happen with pm_gather_item_capacity; never write out of bounds
            atomicExch(ctl + 2, 1u);
            return;
        }
        GatherItem it;
        it.yb_zs = (uint32_t)yb | ((uint32_t)z << 16);
        it.zc = (uint32_t)nz;
        it.beg = b;
        it.end = e;
        items[heavy ? k : (uint32_t)cap - 1u - k] = it;
    };
    const uint32_t cb = row_start[(uint32_t)zs * NC + y0];
    const uint32_t ce = row_start[(uint32_t)(zs + zc0 - 1) * NC + y0 + YB];
    // (the planes of a chunk are separate runs of the sorted list, so ce - cb is an upper bound that also counts
    // the other row blocks in between -- it is only used to skip the per-plane scan of an obviously light chunk)
    uint32_t total = 0;
    if (ce - cb > T) {
        for (int i = 0; i < zc0; ++i) {
            const uint32_t r = (uint32_t)(zs + i) * NC + y0;
            total += row_start[r + YB] - row_start[r];
        }
    }
    if (total <= T) {
        emit(false, zs, zc0, 0u, 0u);
        return;
    }
    int z_first = zs;
    uint32_t acc = 0;
    for (int i = 0; i < zc0; ++i) {
        const uint32_t r = (uint32_t)(zs + i) * NC + y0;
        const uint32_t b = row_start[r], e = row_start[r + YB], c = e - b;
        if (c > T) {
            if (zs + i > z_first) emit(true, z_first, zs + i - z_first, 0u, 0u);
            for (uint32_t q = b; q < e; q += H) emit(true, zs + i, 1, q, e - q < H ? e : q + H);
            z_first = zs + i + 1;
            acc = 0;
        } else if (acc + c > T) {
            emit(true, z_first, zs + i - z_first, 0u, 0u);
            z_first = zs + i;
            acc = c;
        } else {
            acc += c;
        }
    }
    if (zs + zc0 > z_first) emit(true, z_first, zs + zc0 - z_first, 0u, 0u);
}

// s_out != nullptr: the sums-only variant of the warp-specialised kernel (see k_gather_ws, SONLY): s_out is
// an array of float4 records indexed by the particle's original index.
template <int NC>
static int pm_launch_gather_tiled(pm_plan *p, const float *phi, double k_kick, double da, double aa, double f_a1,
                                  uint32_t *cnt, cudaStream_t st, float *s_out = nullptr, int64_t s_stride = 0)
{
    if (s_out && !p->gather_ws) return PM_ERR_UNSUPPORTED;
    int zc = 32;
    if (const char *e = getenv("PM_GATHER_ZC")) {      // planes per CTA (tuning: shorter marches balance a clustered load)
        const int v = atoi(e);
        if (v >= 1 && v <= 32) zc = v;
    }
    while (zc > 1 && (NC % zc || (int64_t)(NC / PM_GT_YB) * (NC / zc) < 8LL * p->sm_count)) zc /= 2;
    constexpr size_t smem = kGtSmem<NC, PM_GT_YB, PM_GT_CAP>;
    static_assert(smem <= 227 * 1024 / PM_GT_MINB, "PM_GT_MINB CTAs per SM");
    auto kern = k_gather_tiled<NC, PM_GT_YB, PM_GT_NT, PM_GT_CAP, PM_GT_MINB>;
    using WsSmem = pmws::Smem<NC, PM_GT_YB, PM_GT_CAP, PM_GW_R, PM_GW_S>;
    static_assert(WsSmem::total <= 227 * 1024 / PM_GT_MINB, "PM_GT_MINB CTAs per SM (warp-specialised gather)");
    auto kern_ws = k_gather_ws<NC, PM_GT_YB, PM_GW_CW, PM_GT_CAP, PM_GW_R, PM_GW_S, PM_GT_MINB, PM_GW_PP>;
    auto kern_ws_s = k_gather_ws<NC, PM_GT_YB, PM_GW_CW, PM_GT_CAP, PM_GW_R, PM_GW_S, PM_GT_MINB, 1, true>;
    PM_ONCE_PER_DEVICE_BEGIN(p->device)
        PM_CUDA(cudaFuncSetAttribute(kern_ws_s, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WsSmem::total));
        PM_CUDA(cudaFuncSetAttribute(kern_ws_s, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        PM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        PM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        PM_CUDA(cudaFuncSetAttribute(kern_ws, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WsSmem::total));
        PM_CUDA(cudaFuncSetAttribute(kern_ws, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    PM_ONCE_PER_DEVICE_END()
    const int c = p->rcur, o = c ^ 1;
    GatherTiledArgs A;
    A.px = p->rpos[c]; A.py = A.px + p->rstride; A.pz = A.py + p->rstride;
    A.vx = p->rvel[c]; A.vy = A.vx + p->rstride; A.vz = A.vy + p->rstride;
    A.id_in = p->rid[c];
    A.perm = p->order_sorted; A.row_start = p->row_start;
    A.pos_out = p->rpos[o]; A.vel_out = p->rvel[o]; A.id_out = p->rid[o];
    A.keys_out = p->keys; A.mover_cnt = cnt; A.phi = phi;
    A.sout = p->rstride;
    A.zc = zc;
    A.k_kick = k_kick; A.da = da; A.aa = aa; A.raa = pm_div_rcp(aa, da); A.f_a1 = f_a1;
    A.sp = p->graph_params;
    A.items = nullptr; A.item_ctl = nullptr; A.item_cap = 0;
    if (s_out) {
        A.pos_out = s_out; A.sout = s_stride;
        A.vel_out = nullptr; A.id_out = nullptr; A.keys_out = nullptr; A.mover_cnt = nullptr;
    }
    dim3 grid(NC / PM_GT_YB, NC / zc);
    if (p->gather_ws && p->gather_items && p->gat_items && NC <= 65535) {
        const int nchunks = (NC / PM_GT_YB) * (NC / zc);
        const uint32_t T = pm_gather_item_threshold(NC, p->rnp, zc);
        // worst-case grid for a fixed launch sequence (CUDA graph): CTAs beyond the list return at once
        int64_t gmax = pm_gather_item_bound(nchunks, p->rnp, T);
        if (gmax > p->gat_cap) gmax = p->gat_cap;
        PM_CUDA(cudaMemsetAsync(p->gat_ctl, 0, 4 * sizeof(uint32_t), st));
        auto kitems = k_gather_items<NC, PM_GT_YB>;
        PM_LAUNCH(kitems, (nchunks + 127) / 128, 128, 0, st, (const uint32_t *)p->row_start, zc, T, (GatherItem *)p->gat_items,
                  (int)gmax, p->gat_ctl);
        A.items = (const GatherItem *)p->gat_items; A.item_ctl = p->gat_ctl; A.item_cap = (int)gmax;
        grid = dim3((unsigned)gmax, 1);
    }
    if (s_out) PM_LAUNCH(kern_ws_s, grid, (PM_GW_CW + 1) * 32, WsSmem::total, st, A, p->fft_sync + 0);
    else if (p->gather_ws) PM_LAUNCH(kern_ws, grid, (PM_GW_CW + 1) * 32, WsSmem::total, st, A, p->fft_sync + 0);
    else PM_LAUNCH(kern, grid, PM_GT_NT, smem, st, A);
    PM_CHECK_LAUNCH();
    return PM_OK;
}

// the scalars the gather kernels take, as pm_k_gather_kick_drift_resident derives them from (a, da, f_a1)
void pm_gather_step_scalars(double a_val, double f_a1, double da, PmStepParams *out)
{
    out->k_kick = da * f_a1;
    out->aa = (a_val + da) * (a_val + da);
    out->da = da;
    out->raa = pm_div_rcp(out->aa, da);
    out->f_a1 = f_a1;
}

bool pm_gather_graphable(const pm_plan *p)
{
    return p->gather_tiled && p->dep_nseg == 1 && (p->nc == 128 || p->nc == 256 || p->nc == 512);
}

static int pm_try_gather_tiled(pm_plan *p, const float *phi, double k_kick, double da, double aa, double f_a1,
                               uint32_t *cnt, cudaStream_t st, float *s_out = nullptr, int64_t s_stride = 0)
{
    if (!p->gather_tiled || p->dep_nseg != 1 || !p->rows_valid || p->kgrad) return PM_ERR_UNSUPPORTED;
    switch (p->nc) {
    case 128: return pm_launch_gather_tiled<128>(p, phi, k_kick, da, aa, f_a1, cnt, st, s_out, s_stride);
    case 256: return pm_launch_gather_tiled<256>(p, phi, k_kick, da, aa, f_a1, cnt, st, s_out, s_stride);
    case 512: return pm_launch_gather_tiled<512>(p, phi, k_kick, da, aa, f_a1, cnt, st, s_out, s_stride);
    default: return PM_ERR_UNSUPPORTED;
    }
}

// ---- the host-buffer step's split gather (pm_step_host) ------------------------------------------------
// The stencil sums of every particle of the resident (cell-ordered) set rcur, stored at the particle's original
// index in the idle half-spectrum buffer as one (s_x, s_y, s_z, -) record per particle.  PM_ERR_UNSUPPORTED when this plan has no warp-specialised
// gather (mesh sizes other than 128/256/512, spectral-gradient option) or the buffer is too small: the caller
// then takes the fused route.
bool pm_gather_sums_ok(const pm_plan *p)
{
    const size_t spec_bytes = (size_t)p->nc * p->nc * (p->nc / 2 + 1) * sizeof(float2);
    return !p->slab && p->spec && p->gather_tiled && p->gather_ws && p->dep_nseg == 1 && !p->kgrad &&
           (p->nc == 128 || p->nc == 256 || p->nc == 512) && p->rnp > 0 && (size_t)p->rnp * 16 <= spec_bytes;
}

int pm_k_gather_sums(pm_plan *p, const float *phi, cudaStream_t st)
{
    if (!pm_gather_sums_ok(p)) return PM_ERR_UNSUPPORTED;
    return pm_try_gather_tiled(p, phi, 0.0, 0.0, 1.0, 0.0, nullptr, st, reinterpret_cast<float *>(p->spec), p->rnp);
}

// Particles [i0, i1) of the CALLER's order: kick and drift (pm_push, the very function the fused kernels call)
// from the uploaded positions / velocities (stride sin) and the stencil-sum records, into dense [3][np] rows.
__global__ void __launch_bounds__(256) k_push_rows(const float *__restrict__ pos_in, const float *__restrict__ vel_in,
                                                   int64_t sin, const float4 *__restrict__ sums, int64_t i0, int64_t i1,
                                                   int64_t np, int nc, double k_kick, double da, double aa, double raa,
                                                   double f_a1, float *__restrict__ pos_out, float *__restrict__ vel_out)
{
    const int64_t i = i0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= i1) return;
    float x = pos_in[i], y = pos_in[sin + i], z = pos_in[2 * sin + i];
    float vx = vel_in[i], vy = vel_in[sin + i], vz = vel_in[2 * sin + i];
    const float4 s4 = sums[i];
    const float sx = s4.x, sy = s4.y, sz = s4.z;
    pm_push(x, vx, sx, k_kick, da, aa, raa, f_a1, nc, nullptr);
    pm_push(y, vy, sy, k_kick, da, aa, raa, f_a1, nc, nullptr);
    pm_push(z, vz, sz, k_kick, da, aa, raa, f_a1, nc, nullptr);
    pos_out[i] = x; pos_out[np + i] = y; pos_out[2 * np + i] = z;
    vel_out[i] = vx; vel_out[np + i] = vy; vel_out[2 * np + i] = vz;
}

int pm_k_push_rows(pm_plan *p, const float *pos_in, const float *vel_in, int64_t i0, int64_t i1, double a_val, double f_a1,
                   double da, float *pos_out, float *vel_out, cudaStream_t st)
{
    if (i1 <= i0) return PM_OK;
    PmStepParams v;
    pm_gather_step_scalars(a_val, f_a1, da, &v);
    PM_LAUNCH(k_push_rows, (unsigned)((i1 - i0 + 255) / 256), 256, 0, st, pos_in, vel_in, (int64_t)p->rstride,
              reinterpret_cast<const float4 *>(p->spec), i0, i1, p->rnp, p->nc, v.k_kick, v.da, v.aa, v.raa, v.f_a1, pos_out,
              vel_out);
    PM_CHECK_LAUNCH();
    return PM_OK;
}

// ---- diagnostics: particles per block of mesh rows, from the row table of the last sort ----------
__global__ void __launch_bounds__(256) k_block_stats(const uint32_t *__restrict__ row_start, int64_t nblocks,
                                                     int rows_per_block, int nseg, uint32_t cap,
                                                     unsigned long long *__restrict__ out)
{
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nblocks) return;
    const int64_t r0 = b * rows_per_block * nseg, r1 = r0 + (int64_t)rows_per_block * nseg;
    const uint32_t n = row_start[r1] - row_start[r0];
    if (n > cap) {
        atomicAdd(out + 1, 1ull);
        atomicAdd(out + 2, (unsigned long long)(n - cap));
    }
    atomicMax(out + 3, (unsigned long long)n);
}

int pm_k_block_stats(pm_plan *p, int rows_per_block, int cap, int64_t *out4, cudaStream_t st)
{
    if (!p->rows_valid) return PM_ERR_INVALID;
    const int64_t nrows = (int64_t)p->nzl * p->nc;
    if (rows_per_block < 1 || nrows % rows_per_block) return PM_ERR_INVALID;
    const int64_t nblocks = nrows / rows_per_block;
    unsigned long long *d = reinterpret_cast<unsigned long long *>(p->diag);
    PM_CUDA(cudaMemsetAsync(d, 0, 4 * sizeof(unsigned long long), st));
    PM_LAUNCH(k_block_stats, (unsigned)((nblocks + 255) / 256), 256, 0, st, (const uint32_t *)p->row_start, nblocks,
              rows_per_block, p->dep_nseg, (uint32_t)cap, d);
    PM_CHECK_LAUNCH();
    unsigned long long h[4];
    PM_CUDA(cudaMemcpyAsync(h, d, sizeof(h), cudaMemcpyDeviceToHost, st));
    PM_CUDA(cudaStreamSynchronize(st));
    out4[0] = nblocks; out4[1] = (int64_t)h[1]; out4[2] = (int64_t)h[2]; out4[3] = (int64_t)h[3];
    return PM_OK;
}

void pm_gather_tile_shape(int *rows_per_block, int *cap)
{
    *rows_per_block = PM_GT_YB;
    *cap = PM_GT_CAP;
}

// Resident variant: reads buffer set `cur` through the new cell order, writes set `cur^1` and the
// next step's keys (p->keys).
int pm_k_gather_kick_drift_resident(pm_plan *p, const float *phi, double a_val, double f_a1,
                                    double da, cudaStream_t st)
{
    const int64_t np = p->rnp;
    if (np == 0) return PM_OK;
    const double k_kick = da * f_a1;
    const double aa = (a_val + da) * (a_val + da);
    const int c = p->rcur, o = c ^ 1;
    // movers per sort tile, counted on the way (pm_sort.cu then skips its counting pass)
    static const bool count_here = !(getenv("PM_GATHER_COUNT") && strcmp(getenv("PM_GATHER_COUNT"), "0") == 0);
    uint32_t *cnt = (count_here && p->inc_a && p->sort_mode != PM_SORT_FULL) ? p->inc_tile : nullptr;
    if (cnt) PM_CUDA(cudaMemsetAsync(cnt, 0, sizeof(uint32_t) * (pm_sort_tiles(np) + 1), st));
    {
        const int rc = pm_try_gather_tiled(p, phi, k_kick, da, aa, f_a1, cnt, st);
        if (rc == PM_OK) {
            p->inc_counted = (cnt != nullptr);
            return PM_OK;
        }
        if (rc != PM_ERR_UNSUPPORTED) return rc;
    }
    auto kern = p->kgrad ? k_gather_kick_drift<true, false, true> : k_gather_kick_drift<true, false, false>;
    PM_LAUNCH(kern, (unsigned)((np + PM_GATHER_THREADS - 1) / PM_GATHER_THREADS), PM_GATHER_THREADS, 0, st, p->rpos[c],
              p->rvel[c], p->rid[c], p->order_sorted, p->rpos[o], p->rvel[o], p->rid[o], p->keys, np,
              p->rstride, p->rstride, phi, p->nc, k_kick, da, aa, pm_div_rcp(aa, da), f_a1, (float *)nullptr, SlabArgs(), cnt, (const uint32_t *)p->keys_sorted,
              (const float *)p->fmesh);
    PM_CHECK_LAUNCH();
    p->inc_counted = (cnt != nullptr);
    return PM_OK;
}

// Slab variant: launched over every entry of the current set (dead ones sort last and are cut
// off on the device by np_valid), clears and fills the leave lists.
int pm_k_gather_kick_drift_slab(pm_plan *p, const float *phi, double a_val, double f_a1, double da,
                                cudaStream_t st)
{
    const int64_t np = p->rtotal;
    PM_CUDA(cudaMemsetAsync(p->leave_cnt, 0, sizeof(uint32_t) * p->nranks, st));
    if (np == 0) return PM_OK;
    const double k_kick = da * f_a1;
    const double aa = (a_val + da) * (a_val + da);
    const int c = p->rcur, o = c ^ 1;
    SlabArgs sl;
    sl.z0 = p->z0; sl.nzl = p->nzl; sl.rank = p->rank;
    sl.np_valid = p->row_start + (size_t)p->nzl * p->nc * p->dep_nseg;
    sl.leave_cnt = p->leave_cnt; sl.leave_slot = p->leave_slot; sl.leave_cap = p->leave_cap;
    auto kern = k_gather_kick_drift<true, true>;
    PM_LAUNCH(kern, (unsigned)((np + PM_GATHER_THREADS - 1) / PM_GATHER_THREADS), PM_GATHER_THREADS, 0, st, p->rpos[c], p->rvel[c], p->rid[c],
              p->order_sorted, p->rpos[o], p->rvel[o], p->rid[o], p->keys, np, p->rstride, p->rstride,
              phi, p->nc, k_kick, da, aa, pm_div_rcp(aa, da), f_a1, (float *)nullptr, sl, (uint32_t *)nullptr, (const uint32_t *)nullptr,
              (const float *)nullptr);
    PM_CHECK_LAUNCH();
    return PM_OK;
}

// out[id[s]] = in[s] for the six particle rows: back to the caller's original particle order.
__global__ void __launch_bounds__(256) k_unpermute(const float *__restrict__ pos_in,
                                                   const float *__restrict__ vel_in,
                                                   const uint32_t *__restrict__ id, int64_t np,
                                                   int64_t sin, float *__restrict__ pos_out,
                                                   float *__restrict__ vel_out)
{
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= np) return;
    const int64_t i = id[s];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        pos_out[d * np + i] = pos_in[d * sin + s];
        vel_out[d * np + i] = vel_in[d * sin + s];
    }
}

// The same in two passes whose every DRAM access is a full 32-byte sector.  The direct scatter above
// writes six isolated 4-byte words per particle: six read-modify-write sectors (1.05 ms for 256^3
// particles).  Here pass one scatters ONE sector per particle -- (x, y, z, vx, vy, vz, -, -) to slot id[s] of
// an array-of-structures scratch -- and pass two streams that scratch into the caller's six rows:
// 0.3 ms.  The scratch is the half-spectrum buffer, idle between two Poisson solves.
__global__ void __launch_bounds__(256) k_unpermute_aos(const float *__restrict__ pos_in, const float *__restrict__ vel_in,
                                                       const uint32_t *__restrict__ id, int64_t np, int64_t sin,
                                                       float4 *__restrict__ aos)
{
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= np) return;
    const int64_t i = id[s];
    const float4 a = make_float4(pos_in[s], pos_in[sin + s], pos_in[2 * sin + s], vel_in[s]);
    const float4 b = make_float4(vel_in[sin + s], vel_in[2 * sin + s], 0.0f, 0.0f);
    aos[2 * i] = a;
    aos[2 * i + 1] = b;
}

__global__ void __launch_bounds__(256) k_aos_to_rows(const float4 *__restrict__ aos, int64_t i0, int64_t i1, int64_t np,
                                                     float *__restrict__ pos_out, float *__restrict__ vel_out)
{
    const int64_t i = i0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= i1) return;
    const float4 a = aos[2 * i], b = aos[2 * i + 1];
    pos_out[i] = a.x; pos_out[np + i] = a.y; pos_out[2 * np + i] = a.z;
    vel_out[i] = a.w; vel_out[np + i] = b.x; vel_out[2 * np + i] = b.y;
}

// out[d][s] = in[d][order[s]] for the three rows of a particle array (both with the plan's row stride)
__global__ void __launch_bounds__(256) k_reorder_rows(const float *__restrict__ in, const uint32_t *__restrict__ order,
                                                      int64_t np, int64_t stride, float *__restrict__ out)
{
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= np) return;
    const int64_t j = order[s];
    out[s] = in[j];
    out[stride + s] = in[stride + j];
    out[2 * stride + s] = in[2 * stride + j];
}

int pm_k_reorder_rows(pm_plan *p, const float *in, const uint32_t *order, int64_t np, float *out, cudaStream_t st)
{
    if (np == 0) return PM_OK;
    PM_LAUNCH(k_reorder_rows, (unsigned)((np + 255) / 256), 256, 0, st, in, order, np, p->rstride, out);
    PM_CHECK_LAUNCH();
    return PM_OK;
}

// the two passes separately (pm_step_host streams pass two in chunks, each followed by its download)
bool pm_unpermute_aos_ok(const pm_plan *p, const float *pos_out, const float *vel_out)
{
    const int64_t np = p->rnp;
    const size_t spec_bytes = (size_t)p->nc * p->nc * (p->nc / 2 + 1) * sizeof(float2);
    const void *lo = pos_out < vel_out ? (const void *)pos_out : (const void *)vel_out;
    const void *hi = pos_out < vel_out ? (const void *)(vel_out + 3 * np) : (const void *)(pos_out + 3 * np);
    const bool apart = (const char *)hi <= (const char *)p->spec || (const char *)lo >= (const char *)p->spec + spec_bytes;
    return !p->slab && p->spec && np > 0 && (size_t)np * 32 <= spec_bytes && apart;
}

int pm_k_unpermute_scatter_aos(pm_plan *p, cudaStream_t st)
{
    const int64_t np = p->rnp;
    const int c = p->rcur;
    PM_LAUNCH(k_unpermute_aos, (unsigned)((np + 255) / 256), 256, 0, st, p->rpos[c], p->rvel[c], p->rid[c], np,
              p->rstride, reinterpret_cast<float4 *>(p->spec));
    PM_CHECK_LAUNCH();
    return PM_OK;
}

int pm_k_aos_rows_range(pm_plan *p, int64_t i0, int64_t i1, float *pos_out, float *vel_out, cudaStream_t st)
{
    if (i1 <= i0) return PM_OK;
    PM_LAUNCH(k_aos_to_rows, (unsigned)((i1 - i0 + 255) / 256), 256, 0, st, reinterpret_cast<const float4 *>(p->spec), i0, i1,
              p->rnp, pos_out, vel_out);
    PM_CHECK_LAUNCH();
    return PM_OK;
}

int pm_k_unpermute(pm_plan *p, float *pos_out, float *vel_out, cudaStream_t st)
{
    const int64_t np = p->rnp;
    if (np == 0) return PM_OK;
    const int c = p->rcur;
    if (pm_unpermute_aos_ok(p, pos_out, vel_out)) {
        const int rc = pm_k_unpermute_scatter_aos(p, st);
        return rc != PM_OK ? rc : pm_k_aos_rows_range(p, 0, np, pos_out, vel_out, st);
    }
    PM_LAUNCH(k_unpermute, (unsigned)((np + 255) / 256), 256, 0, st, p->rpos[c], p->rvel[c],
              p->rid[c], np, p->rstride, pos_out, vel_out);
    PM_CHECK_LAUNCH();
    return PM_OK;
}

__global__ void __launch_bounds__(256) k_iota(uint32_t *out, int64_t n)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (uint32_t)i;
}

int pm_k_iota(uint32_t *out, int64_t n, cudaStream_t st)
{
    if (n == 0) return PM_OK;
    PM_LAUNCH(k_iota, (unsigned)((n + 255) / 256), 256, 0, st, out, n);
    PM_CHECK_LAUNCH();
    return PM_OK;
}

// --------------------------------------------------------------------------------------------
// Slab mode: ghost-plane add and particle migration (no reference counterpart).
// --------------------------------------------------------------------------------------------
// rho plane 0 += the ghost plane received from rank-1 (its d_z share of our bottom plane).
__global__ void __launch_bounds__(256) k_ghost_add(float4 *__restrict__ plane,
                                                   const float4 *__restrict__ ghost, size_t n4)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4;
         i += (size_t)gridDim.x * blockDim.x) {
        float4 a = plane[i];
        const float4 b = ghost[i];
        a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
        plane[i] = a;
    }
}

int pm_k_ghost_add(pm_plan *p, float *plane, const float *ghost, cudaStream_t st)
{
    const size_t n4 = (size_t)p->nc * p->nc / 4;
    PM_LAUNCH(k_ghost_add, p->sm_count * 4, 256, 0, st, reinterpret_cast<float4 *>(plane),
              reinterpret_cast<const float4 *>(ghost), n4);
    PM_CHECK_LAUNCH();
    return PM_OK;
}

// Leavers -> 7-float records (x, y, z, vx, vy, vz, id bits), in ascending slot order per
// destination so that the receiver's storage order (and with it every later summation order)
// does not depend on the atomics that built the leave lists.
__global__ void __launch_bounds__(256) k_migrate_pack(const float *__restrict__ pos,
                                                      const float *__restrict__ vel,
                                                      const uint32_t *__restrict__ id, int64_t stride,
                                                      const uint32_t *__restrict__ slots, int64_t n,
                                                      float *__restrict__ rec)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t s = slots[i];
    float *r = rec + i * 7;
    r[0] = pos[s]; r[1] = pos[stride + s]; r[2] = pos[2 * stride + s];
    r[3] = vel[s]; r[4] = vel[stride + s]; r[5] = vel[2 * stride + s];
    r[6] = __uint_as_float(id[s]);
}

int pm_k_migrate_pack(pm_plan *p, int dest, int64_t count, int64_t rec_offset, cudaStream_t st)
{
    if (count == 0) return PM_OK;
    uint32_t *sorted = p->leave_sorted;
    {
        const int rc = pm_k_sort_u32(p, p->leave_slot + (size_t)dest * p->leave_cap, sorted, count, st);
        if (rc != PM_OK) return rc;
    }
    const int c = p->rcur;
    PM_LAUNCH(k_migrate_pack, (unsigned)((count + 255) / 256), 256, 0, st, p->rpos[c], p->rvel[c],
              p->rid[c], p->rstride, sorted, count, p->mig_send + rec_offset * 7);
    PM_CHECK_LAUNCH();
    return PM_OK;
}

// Arrivals are appended behind the current entries; their keys are formed here so the next
// sort files them into cell order together with everything else.
__global__ void __launch_bounds__(256) k_migrate_unpack(const float *__restrict__ rec, int64_t n,
                                                        float *__restrict__ pos,
                                                        float *__restrict__ vel,
                                                        uint32_t *__restrict__ id,
                                                        uint32_t *__restrict__ keys, int64_t stride,
                                                        int64_t first, int nc, int z0, int nzl)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float *r = rec + i * 7;
    const int64_t s = first + i;
    pos[s] = r[0]; pos[stride + s] = r[1]; pos[2 * stride + s] = r[2];
    vel[s] = r[3]; vel[stride + s] = r[4]; vel[2 * stride + s] = r[5];
    id[s] = __float_as_uint(r[6]);
    keys[s] = pm_key(r[0], r[1], r[2], nc, z0, nzl);
}

int pm_k_migrate_unpack(pm_plan *p, int64_t n_arrive, cudaStream_t st)
{
    if (n_arrive == 0) return PM_OK;
    const int c = p->rcur;
    PM_LAUNCH(k_migrate_unpack, (unsigned)((n_arrive + 255) / 256), 256, 0, st, p->mig_recv, n_arrive,
              p->rpos[c], p->rvel[c], p->rid[c], p->keys, p->rstride, p->rtotal, p->nc, p->z0, p->nzl);
    PM_CHECK_LAUNCH();
    return PM_OK;
}

// live particles of the current set, storage order, dense [3][n] rows
__global__ void __launch_bounds__(256) k_export(const float *__restrict__ pos,
                                                const float *__restrict__ vel,
                                                const uint32_t *__restrict__ id,
                                                const uint32_t *__restrict__ keys, int64_t stride,
                                                int64_t n, float *__restrict__ pos_out,
                                                float *__restrict__ vel_out,
                                                uint32_t *__restrict__ id_out,
                                                uint32_t *__restrict__ live_out)
{
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        pos_out[d * n + s] = pos[d * stride + s];
        vel_out[d * n + s] = vel[d * stride + s];
    }
    id_out[s] = id[s];
    live_out[s] = keys ? (keys[s] != PM_KEY_DEAD) : 1u;
}

int pm_k_export(pm_plan *p, float *pos_out, float *vel_out, uint32_t *id_out, uint32_t *live_out,
                cudaStream_t st)
{
    const int64_t n = p->rtotal;
    if (n == 0) return PM_OK;
    const int c = p->rcur;
    PM_LAUNCH(k_export, (unsigned)((n + 255) / 256), 256, 0, st, p->rpos[c], p->rvel[c], p->rid[c],
              p->rkeys_valid ? p->keys : (const uint32_t *)nullptr, p->rstride, n, pos_out, vel_out,
              id_out, live_out);
    PM_CHECK_LAUNCH();
    return PM_OK;
}
