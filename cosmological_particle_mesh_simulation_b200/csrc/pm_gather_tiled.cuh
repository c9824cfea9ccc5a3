// pm_gather_tiled.cuh -- force gather + kick + drift of the RESIDENT (cell-ordered) particle list
// with the potential staged through shared memory.  Included by pm_particles.cu (uses pm_cell,
// pm_key, pm_gp, pm_push).  Same arithmetic and rounding points as k_gather_kick_drift
// (src/integrate.py:15-97); only the way phi reaches the registers differs.
//
// Why: k_gather_kick_drift issues 32 scattered 4-byte loads of phi per particle.  In cell order a
// warp's 32 particles span ~256 cells of one mesh row, so every one of those warp loads touches ~8
// cache lines: ~270 L1 wavefronts per warp, and ncu shows the kernel bound by exactly that
// (profiles/r01_notes.md).  The lines themselves are shared by neighbouring particle rows and by
// the four planes a particle reads, so here a CTA owns a block of YB particle rows (z, y0..y0+YB-1)
// and marches along z:
//     ring of 5 phi slabs in shared memory, slab = rows y0-1 .. y0+YB+1 of one plane (cp.async,
//     16 bytes per request, straight from L2); step z reads planes z-1..z+2 while plane z+3 lands
//     particle inputs of step z+2 (pos, vel, id through the sort permutation) and the permutation
//     entries of step z+3 land in two small staging rings at the same time (4-byte cp.async), so
//     the dependent chain perm -> particle -> phi of the one-thread-per-particle kernel is gone:
//     every global access of a step was issued one or two steps earlier.
// Each plane of a CTA's column is fetched once (44-56 bytes of L2 traffic per particle instead of
// 32 scattered sectors) and the 32 phi reads per particle are shared-memory reads.
//
// Requirements (else the caller uses k_gather_kick_drift): resident non-slab state, one deposit
// segment per row, a mesh size the kernel is instantiated for (128, 256, 512), zc | nc, zc <= 32.
// A (z, y-block) with more than CAP particles (deep inside a halo) handles the excess through
// plain global loads of the particle data (phi still comes from the ring).
#pragma once

__device__ __forceinline__ void pm_cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void pm_cp_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// One CTA's share of the gather: zc planes starting at plane zs of row block yb; a crowded single plane
// (zc == 1) may be cut into particle ranges [beg, end) of the sorted list (end > beg), else beg = end = 0.
struct GatherItem {
    uint32_t yb_zs;             // yb | zs << 16
    uint32_t zc;
    uint32_t beg, end;
};

struct GatherTiledArgs {
    const float *px, *py, *pz, *vx, *vy, *vz;   // current buffer set, one pointer per SoA row
    const uint32_t *id_in, *perm, *row_start;
    float *pos_out, *vel_out;
    uint32_t *id_out, *keys_out, *mover_cnt;
    const float *phi;
    int64_t sout;
    int zc;                     // planes per CTA
    // k_gather_ws only.  non-null: the CTAs work through a list instead of the (row block, z chunk) grid
    // (k_gather_items, pm_particles.cu): item i < ctl[0] is items[i] (pieces of crowded chunks, taken
    // first), the ctl[1] others are items[item_cap - 1 - (i - ctl[0])]
    const struct GatherItem *items;
    const uint32_t *item_ctl;
    int item_cap;
    double k_kick, da, aa, raa, f_a1;
    const PmStepParams *sp;     // non-null: the five scalars above are read from device memory (graph replays)
};

__device__ __forceinline__ void pm_gather_step_params(GatherTiledArgs &A)
{
    if (A.sp) {
        A.k_kick = A.sp->k_kick; A.da = A.sp->da; A.aa = A.sp->aa; A.raa = A.sp->raa; A.f_a1 = A.sp->f_a1;
    }
}

constexpr int kGtRing = 5;

template <int NC, int YB, int CAP>
constexpr size_t kGtSmem = ((size_t)kGtRing * (YB + 3) * NC + 3 * 7 * CAP + 4 * CAP + 2 * 32) * sizeof(float);

// NC is a template parameter: every shared-memory offset of the 32 phi reads, the wrap tests and
// the row/quad split of the slab copies fold into immediates (the first, runtime-nc version of
// this kernel executed twice the instructions of k_gather_kick_drift and was issue-bound).
template <int NC, int YB, int NT, int CAP, int MINB>
__global__ void __launch_bounds__(NT, MINB) k_gather_tiled(GatherTiledArgs A)
{
    pm_gather_step_params(A);
    constexpr int SR = YB + 3;          // rows of a slab: y0-1 .. y0+YB+1
    constexpr int SLAB = SR * NC;       // floats
    constexpr int QUADS = NC / 4;
    static_assert(NC % 4 == 0 && NC % YB == 0, "mesh shape");
    extern __shared__ float4 s_raw4[];
    float *ring = reinterpret_cast<float *>(s_raw4);                        // [5][SR][NC]
    float *stage = ring + kGtRing * SLAB;                                   // [3][7][CAP]
    uint32_t *pst = reinterpret_cast<uint32_t *>(stage + 3 * 7 * CAP);      // [4][CAP]
    uint32_t *s_beg = pst + 4 * CAP, *s_end = s_beg + 32;                   // particle range of each step
    const unsigned u_ring = (unsigned)__cvta_generic_to_shared(ring);
    const unsigned u_stage = (unsigned)__cvta_generic_to_shared(stage);
    const unsigned u_pst = (unsigned)__cvta_generic_to_shared(pst);
    const int tid = threadIdx.x;
    const int y0 = blockIdx.x * YB;
    const int zc = A.zc;
    const int zs = blockIdx.y * zc;

    for (int i = tid; i < zc; i += NT) {
        const uint32_t r = (uint32_t)(zs + i) * NC + y0;
        s_beg[i] = A.row_start[r];
        s_end[i] = A.row_start[r + YB];
    }
    __syncthreads();

    auto cp16 = [](unsigned sa, const void *g) { asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(g) : "memory"); };
    auto cp4 = [](unsigned sa, const void *g) { asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(sa), "l"(g) : "memory"); };

    // plane zp (zs-1 .. zs+zc+1, periodic) into ring slot `slot`
    auto issue_phi = [&](int zp, int slot) {
        const int zw = zp < 0 ? zp + NC : (zp >= NC ? zp - NC : zp);
        const unsigned dst = u_ring + (unsigned)slot * (SLAB * 4);
        for (int idx = tid; idx < SR * QUADS; idx += NT) {
            const int row = idx / QUADS, q = idx % QUADS;
            int y = y0 - 1 + row;
            y = y < 0 ? y + NC : (y >= NC ? y - NC : y);
            cp16(dst + (unsigned)(row * NC + 4 * q) * 4, A.phi + (uint32_t)((zw * NC + y) * NC + 4 * q));
        }
    };
    // k = step index inside the chunk
    auto staged = [&](int k) -> int {
        const uint32_t n = s_end[k] - s_beg[k];
        return n < (uint32_t)CAP ? (int)n : CAP;
    };
    auto issue_perm = [&](int k, int pslot) {
        if (k >= zc) return;
        const int n = staged(k);
        const uint32_t *src = A.perm + s_beg[k];
        const unsigned dst = u_pst + (unsigned)pslot * (CAP * 4);
        for (int i = tid; i < n; i += NT) cp4(dst + i * 4, src + i);
    };
    auto issue_particles = [&](int k, int pslot, int sslot) {
        if (k >= zc) return;
        const int n = staged(k);
        const uint32_t *pj = pst + pslot * CAP;
        const unsigned dst = u_stage + (unsigned)sslot * (7 * CAP * 4);
        for (int i = tid; i < n; i += NT) {
            const uint32_t j = pj[i];
            const unsigned d = dst + i * 4;
            cp4(d + 0 * CAP * 4, A.px + j);
            cp4(d + 1 * CAP * 4, A.py + j);
            cp4(d + 2 * CAP * 4, A.pz + j);
            cp4(d + 3 * CAP * 4, A.vx + j);
            cp4(d + 4 * CAP * 4, A.vy + j);
            cp4(d + 5 * CAP * 4, A.vz + j);
            cp4(d + 6 * CAP * 4, A.id_in + j);
        }
    };

    // one particle: phi from the ring (plane z-1 in slot s0), outputs to slot p of the other buffer set
    auto update = [&](int s0, uint32_t p, float x, float y, float z, float vx, float vy, float vz, uint32_t id) {
        const int xc = pm_cell(x, NC), yc = pm_cell(y, NC), zcell = pm_cell(z, NC);
        const uint32_t kold = ((uint32_t)zcell * NC + yc) * NC + xc;
        const double d_x = (double)x - (double)xc, d_y = (double)y - (double)yc, d_z = (double)z - (double)zcell;
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
        // x neighbours c-1, c, c+1, c+2 (periodic); y rows of the slab; z planes of the ring
        int xo[4];
        {
            const int a1 = xc + 1 == NC ? 0 : xc + 1;
            xo[0] = xc == 0 ? NC - 1 : xc - 1; xo[1] = xc; xo[2] = a1; xo[3] = a1 + 1 == NC ? 0 : a1 + 1;
        }
        int ry = yc - y0;                       // 0 .. YB-1 by the sort order = slab row of y_c - 1
        ry = ry < 0 ? 0 : (ry > YB - 1 ? YB - 1 : ry);
        const float *row0 = ring + ry * NC;
        float v[4][4][4];
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            int slot = s0 + a;
            slot = slot >= kGtRing ? slot - kGtRing : slot;
            const float *pl = row0 + slot * SLAB;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const float *pc = pl + xo[c];
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    const int outer = (a == 0 || a == 3) + (b == 0 || b == 3) + (c == 0 || c == 3);
                    v[a][b][c] = (outer <= 1) ? pc[b * NC] : 0.0f;
                }
            }
        }
        const float sx = pm_gp<0>(v, t), sy = pm_gp<1>(v, t), sz = pm_gp<2>(v, t);
        pm_push(x, vx, sx, A.k_kick, A.da, A.aa, A.raa, A.f_a1, NC, nullptr);
        pm_push(y, vy, sy, A.k_kick, A.da, A.aa, A.raa, A.f_a1, NC, nullptr);
        pm_push(z, vz, sz, A.k_kick, A.da, A.aa, A.raa, A.f_a1, NC, nullptr);
        A.pos_out[p] = x; A.pos_out[A.sout + p] = y; A.pos_out[2 * A.sout + p] = z;
        A.vel_out[p] = vx; A.vel_out[A.sout + p] = vy; A.vel_out[2 * A.sout + p] = vz;
        A.id_out[p] = id;
        const uint32_t knew = pm_key(x, y, z, NC, 0, NC);
        A.keys_out[p] = knew;
        if (A.mover_cnt) {
            // movers per sort tile (pm_sort.cu); a warp's slots are consecutive but may straddle a tile
            const unsigned act = __activemask();
            const uint32_t tile = p / PM_SORT_TILE;
            const unsigned same = __match_any_sync(act, tile);
            const unsigned mv = __ballot_sync(act, knew != kold) & same;
            if (mv && (int)(threadIdx.x & 31) == __ffs(same) - 1) atomicAdd(A.mover_cnt + tile, __popc(mv));
        }
    };

    // Prefetch distances: phi one step (L2 hits, 16-byte requests), particle inputs two steps (scattered
    // 4-byte DRAM requests), permutation entries three steps.  Each step commits two groups, phi and
    // permutation first: cp.async groups retire in order, so wait_group 1 at the end of step k leaves
    // only the particle requests of step k+2 in flight.
    // prologue: planes zs-1 .. zs+2 in slots 0..3, permutation entries of steps 0..2, particles of 0, 1
    issue_phi(zs - 1, 0); issue_phi(zs, 1); issue_phi(zs + 1, 2); issue_phi(zs + 2, 3);
    issue_perm(0, 0); issue_perm(1, 1); issue_perm(2, 2);
    pm_cp_commit(); pm_cp_wait_all();
    __syncthreads();
    issue_particles(0, 0, 0); issue_particles(1, 1, 1);
    pm_cp_commit(); pm_cp_wait_all();
    __syncthreads();

    int s0 = 0;    // ring slot of plane z-1
    int p3 = 0;    // k % 3 (staging slot of step k)
    int p4 = 0;    // k % 4 (permutation slot of step k)
    for (int k = 0; k < zc; ++k) {
        // requests for later steps; every buffer they write was released by the barrier below
        const int snew = s0 + 4 >= kGtRing ? s0 + 4 - kGtRing : s0 + 4;
        if (k + 1 < zc) issue_phi(zs + k + 3, snew);
        issue_perm(k + 3, (p4 + 3) & 3);   // read by issue_particles of the NEXT step: must be in the first group
        pm_cp_commit();
        issue_particles(k + 2, (p4 + 2) & 3, p3 + 2 >= 3 ? p3 - 1 : p3 + 2);
        pm_cp_commit();

        const uint32_t beg = s_beg[k], end = s_end[k];
        const int n = staged(k);
        const float *st = stage + p3 * 7 * CAP;
        for (int i = tid; i < n; i += NT)
            update(s0, beg + i, st[i], st[CAP + i], st[2 * CAP + i], st[3 * CAP + i], st[4 * CAP + i], st[5 * CAP + i],
                   __float_as_uint(st[6 * CAP + i]));
        for (uint32_t p = beg + CAP + tid; p < end; p += NT) {   // overflow of a crowded block
            const uint32_t j = A.perm[p];
            update(s0, p, A.px[j], A.py[j], A.pz[j], A.vx[j], A.vy[j], A.vz[j], A.id_in[j]);
        }
        asm volatile("cp.async.wait_group 1;" ::: "memory");
        __syncthreads();
        s0 = s0 + 1 >= kGtRing ? 0 : s0 + 1;
        p3 = p3 + 1 >= 3 ? 0 : p3 + 1;
        p4 = (p4 + 1) & 3;
    }
}
