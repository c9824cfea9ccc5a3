// pm_gather_ws.cuh -- force gather + kick + drift of the resident (cell-ordered) particle list,
// warp-specialised: one producer warp feeds shared memory, consumer warps do the arithmetic, and
// mbarriers -- not CTA-wide barriers -- order them.  Included by pm_particles.cu after
// pm_gather_tiled.cuh (same GatherTiledArgs; same arithmetic and rounding points as
// k_gather_kick_drift / src/integrate.py:15-97: pm_cell, pm_key, pm_gp, pm_push).
//
// What k_gather_tiled left on the table (ncu, profiles/r01_notes.md): 384 threads per CTA for the
// ~256 particles of a step (a third of the lanes idle), one __syncthreads per z step at 16 busy
// warps per SM, 4-byte cp.async requests issued by the compute threads themselves, and -- under
// clustering (BASELINE configs[4]) -- every particle beyond the staging capacity of a step read
// through a dependent perm -> particle chain of plain global loads (1.09 ms instead of 0.50 ms on the
// z = 0 snapshot, profiles/r02_a_bench_evolved_baseline.json).  Here:
//   * a CTA still owns YB particle rows (z, y0..y0+YB-1) and marches along z through a ring of R phi
//     slabs (rows y0-1 .. y0+YB+1 of one plane = ONE contiguous (YB+3)*NC*4-byte run of the mesh, two
//     runs where y wraps), but a slab now arrives by ONE bulk copy (cp.async.bulk, the TMA engine;
//     UBLKCP in SASS) that a single producer lane issues against an mbarrier (complete_tx);
//   * a z step is cut into SUB-STEPS of at most CAP particles, each with its own staging buffer
//     (ring of S): a crowded (z, row-block) is simply more sub-steps through the same pipeline;
//   * the producer warp reads the sort permutation and issues the 4-byte cp.async gathers of a
//     sub-step's particle rows (position, velocity, id through the permutation), then
//     cp.async.mbarrier.arrive: the "full" barrier of the staging buffer completes when the data is in;
//   * consumer warp w takes the 32-particle batches w, w+CW, ... of a sub-step -- full lanes whatever
//     the count -- and releases the staging buffer / the oldest phi slab by mbarrier.arrive; warps
//     drift apart by up to the ring depths instead of meeting at a barrier every step.
// Every wait is a bounded spin: a protocol error raises the plan's error word instead of hanging.
#pragma once

namespace pmws {

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// completes one arrival on `bar` when all cp.async of the calling thread issued so far have landed
__device__ __forceinline__ void mbar_arrive_on_cp_async(uint64_t *bar)
{
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, unsigned parity)
{
    unsigned ok;
    asm volatile(
        "{\n\t.reg .pred P;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, P;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// bounded wait; false (and *err = 1) if the phase never completes
__device__ __forceinline__ bool mbar_wait(uint64_t *bar, unsigned parity, unsigned *err)
{
    for (unsigned spins = 0; spins < (1u << 24); ++spins)
        if (mbar_try_wait(bar, parity)) return true;
    atomicExch(err, 1u);
    return false;
}
// one contiguous run global -> shared through the bulk-copy (TMA) engine; 16-byte aligned, size % 16 == 0
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void cp4(unsigned dst, const void *src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}

template <int NC, int YB, int CAP, int R, int S>
struct Smem {
    static constexpr int SR = YB + 3;
    static constexpr int SLAB = SR * NC;                 // floats per phi slab
    static constexpr size_t ring = (size_t)R * SLAB * 4;
    static constexpr size_t stage = (size_t)S * 7 * CAP * 4;
    static constexpr size_t bars = (size_t)(2 * S + 2 * R) * 8;
    static constexpr size_t ranges = 2 * 32 * 4;
    static constexpr size_t total = ring + stage + bars + ranges;
};

}  // namespace pmws

// CW consumer warps + 1 producer warp.  Grid: (NC / YB, NC / zc), zc <= 32 planes per CTA.
// PP: particles per consumer lane and batch.  PP = 2 runs two particles' dependent chains (shared-memory phi
// reads -> float32 differences -> float64 kick and drift) side by side in one thread, phase by phase, so that
// the instruction scheduler can interleave them: the kernel is bound by latency at ~18 warps per SM, not by
// issue slots (43 %) or DRAM (46 %).
// SONLY (the host-buffer step, pm_step_host): the kernel stops after the gather proper -- the three stencil
// sums s_x, s_y, s_z of a particle, which is everything pm_push needs from the mesh -- and stores them at the
// particle's ORIGINAL index (one float4 record: ((float4 *)pos_out)[id]); velocities are neither read for arithmetic nor written,
// so the kernel can run while they are still on their way over PCIe, and the push itself is done in the
// caller's order afterwards (k_push_rows, same pm_push).
template <int NC, int YB, int CW, int CAP, int R, int S, int MINB, int PP = 1, bool SONLY = false>
__global__ void __launch_bounds__((CW + 1) * 32, MINB) k_gather_ws(GatherTiledArgs A, unsigned *err)
{
    static_assert(!SONLY || PP == 1, "sums-only gather: one particle per lane");
    pm_gather_step_params(A);
    using namespace pmws;
    using L = Smem<NC, YB, CAP, R, S>;
    constexpr int SR = L::SR, SLAB = L::SLAB;
    static_assert(NC % 4 == 0 && NC % YB == 0 && R >= 5 && S >= 2, "shape");
    extern __shared__ float4 s_raw4[];
    float *ring = reinterpret_cast<float *>(s_raw4);                              // [R][SR][NC]
    float *stage = ring + R * SLAB;                                               // [S][7][CAP]
    uint64_t *bars = reinterpret_cast<uint64_t *>(stage + S * 7 * CAP);
    uint64_t *full_p = bars, *empty_p = bars + S, *full_phi = bars + 2 * S, *empty_phi = bars + 2 * S + R;
    uint32_t *s_beg = reinterpret_cast<uint32_t *>(bars + 2 * S + 2 * R), *s_end = s_beg + 32;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int y0 = blockIdx.x * YB;
    int zc = A.zc;
    int zs = blockIdx.y * zc;
    uint32_t ov_beg = 0, ov_end = 0;
    if (A.items) {
        // work list: the same march over zc planes, but (row block, first plane, planes) come from an item, and
        // an item of one crowded plane covers only a range of that plane's particles.  One CTA per list slot:
        // the grid is the list's worst-case length and CTAs beyond its end leave here.  (A variant whose CTAs
        // loop over items drawn from a counter -- one CTA per base chunk, no empty CTAs -- measured slower:
        // +5 registers and the per-item barrier re-initialisation cost more than the ~8000 empty CTAs.)
        const uint32_t nh = A.item_ctl[0], nl = A.item_ctl[1], i = blockIdx.x;
        if (i >= nh + nl) return;
        const GatherItem it = A.items[i < nh ? i : (uint32_t)A.item_cap - 1u - (i - nh)];
        y0 = (int)(it.yb_zs & 0xffffu) * YB;
        zs = (int)(it.yb_zs >> 16);
        zc = (int)it.zc;
        ov_beg = it.beg;
        ov_end = it.end;
    }

    for (int i = tid; i < zc; i += (CW + 1) * 32) {
        const uint32_t r = (uint32_t)(zs + i) * NC + y0;
        s_beg[i] = ov_end > ov_beg ? ov_beg : A.row_start[r];
        s_end[i] = ov_end > ov_beg ? ov_end : A.row_start[r + YB];
    }
    if (tid == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(full_p + s, 32);       // the 32 producer lanes' cp.async arrivals
            mbar_init(empty_p + s, CW);      // one arrival per consumer warp
        }
        for (int r = 0; r < R; ++r) {
            mbar_init(full_phi + r, 1);      // the issuing lane's arrive.expect_tx; the bulk copies complete the bytes
            mbar_init(empty_phi + r, CW);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp == CW) {
        // ------------------------------ producer warp ------------------------------
        // phi plane index pi <-> global plane zs - 1 + pi (periodic), ring slot pi % R; step k reads
        // pi = k .. k+3 and is the last reader of pi = k.
        int issued = 0;
        const int nplanes = zc + 3;
        auto issue_plane = [&](int pi) {
            const int slot = pi % R;
            if (lane == 0) {
                int zp = zs - 1 + pi;
                zp = zp < 0 ? zp + NC : (zp >= NC ? zp - NC : zp);
                float *dst = ring + slot * SLAB;
                const float *plane = A.phi + (size_t)zp * NC * NC;
                mbar_arrive_expect_tx(full_phi + slot, (unsigned)(SLAB * 4));
                // slab rows y0-1 .. y0+YB+1: contiguous in the mesh except where y wraps
                if (y0 == 0) {
                    bulk_g2s(dst, plane + (size_t)(NC - 1) * NC, NC * 4, full_phi + slot);
                    bulk_g2s(dst + NC, plane, (SR - 1) * NC * 4, full_phi + slot);
                } else if (y0 + YB + 1 >= NC) {
                    const int head = NC - (y0 - 1);                    // rows y0-1 .. NC-1
                    bulk_g2s(dst, plane + (size_t)(y0 - 1) * NC, head * NC * 4, full_phi + slot);
                    bulk_g2s(dst + head * NC, plane, (SR - head) * NC * 4, full_phi + slot);
                } else {
                    bulk_g2s(dst, plane + (size_t)(y0 - 1) * NC, SLAB * 4, full_phi + slot);
                }
            }
        };
        const unsigned u_stage = smem_u32(stage);
        // The producer never blocks on one resource while the other could advance: it polls.  Particle
        // sub-steps run up to S ahead of the consumers, phi planes up to R - 4 steps ahead.
        int s = 0, k = 0;            // next sub-step to stage; the z step it belongs to
        uint32_t off = 0;            // particles of step k already staged
        unsigned idle = 0;
        // permutation entries of the NEXT sub-step, loaded right after the previous one was issued: by
        // the time its staging buffer is free they have long arrived, so the perm -> particle chain of
        // dependent loads never stalls the producer
        constexpr int PB = (CAP + 31) / 32;
        uint32_t pj[PB];
        auto load_perm = [&]() {
            if (k >= zc) return;
            const uint32_t beg = s_beg[k], n = s_end[k] - beg;
            const uint32_t cnt = min((uint32_t)CAP, n - off);
            const uint32_t *perm = A.perm + beg + off;
#pragma unroll
            for (int b = 0; b < PB; ++b) {
                const uint32_t i = b * 32 + lane;
                pj[b] = i < cnt ? __ldg(perm + i) : 0u;
            }
        };
        load_perm();
        while (issued < nplanes || k < zc) {
            bool progressed = false;
            if (issued < nplanes) {
                bool can = issued < R;
                if (!can) {
                    const unsigned ready = lane == 0 ? (unsigned)mbar_try_wait(empty_phi + issued % R, (unsigned)((issued / R - 1) & 1)) : 0u;
                    can = __shfl_sync(0xffffffffu, ready, 0) != 0u;
                }
                if (can) {
                    issue_plane(issued);
                    ++issued;
                    progressed = true;
                }
            }
            if (k < zc) {
                const int st = s % S;
                bool can = s < S;
                if (!can) {
                    const unsigned ready = lane == 0 ? (unsigned)mbar_try_wait(empty_p + st, (unsigned)((s / S - 1) & 1)) : 0u;
                    can = __shfl_sync(0xffffffffu, ready, 0) != 0u;
                }
                if (can) {
                    const uint32_t n = s_end[k] - s_beg[k];
                    const uint32_t cnt = min((uint32_t)CAP, n - off);
                    const unsigned dst = u_stage + (unsigned)st * (7 * CAP * 4);
#pragma unroll
                    for (int b = 0; b < PB; ++b) {
                        const uint32_t i = b * 32 + lane;
                        if (i < cnt) {
                            const unsigned d = dst + i * 4;
                            const uint32_t j = pj[b];
                            cp4(d + 0 * CAP * 4, A.px + j);
                            cp4(d + 1 * CAP * 4, A.py + j);
                            cp4(d + 2 * CAP * 4, A.pz + j);
                            cp4(d + 3 * CAP * 4, A.vx + j);
                            cp4(d + 4 * CAP * 4, A.vy + j);
                            cp4(d + 5 * CAP * 4, A.vz + j);
                            cp4(d + 6 * CAP * 4, A.id_in + j);
                        }
                    }
                    mbar_arrive_on_cp_async(full_p + st);
                    ++s;
                    off += CAP;
                    if (off >= n) {
                        ++k;
                        off = 0;
                    }
                    load_perm();
                    progressed = true;
                }
            }
            if (progressed) {
                idle = 0;
            } else if (++idle > (1u << 24)) {
                atomicExch(err, 1u);
                return;
            }
        }
        return;
    }

    // ------------------------------ consumer warps ------------------------------
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
            slot = slot >= R ? slot - R : slot;
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
        if constexpr (SONLY) {
            reinterpret_cast<float4 *>(A.pos_out)[id] = make_float4(sx, sy, sz, 0.0f);   // one 16-byte store, not three 4-byte ones
            return;
        }
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

    // PP particles of one lane, phase by phase (same arithmetic and rounding points as `update`)
    auto update_n = [&](int s0, uint32_t pbase, const uint32_t (&ii)[PP], const bool (&ok)[PP], const float *sp) {
        float x[PP], y[PP], z[PP], vx[PP], vy[PP], vz[PP], sx[PP], sy[PP], sz[PP];
        uint32_t id[PP], kold[PP];
#pragma unroll
        for (int j = 0; j < PP; ++j) {
            const uint32_t i = ok[j] ? ii[j] : 0u;
            x[j] = ok[j] ? sp[i] : 0.f; y[j] = ok[j] ? sp[CAP + i] : 0.f; z[j] = ok[j] ? sp[2 * CAP + i] : 0.f;
            vx[j] = sp[3 * CAP + i]; vy[j] = sp[4 * CAP + i]; vz[j] = sp[5 * CAP + i];
            id[j] = __float_as_uint(sp[6 * CAP + i]);
        }
#pragma unroll
        for (int j = 0; j < PP; ++j) {
            const int xc = pm_cell(x[j], NC), yc = pm_cell(y[j], NC), zcell = pm_cell(z[j], NC);
            kold[j] = ((uint32_t)zcell * NC + yc) * NC + xc;
            const double d_x = (double)x[j] - (double)xc, d_y = (double)y[j] - (double)yc, d_z = (double)z[j] - (double)zcell;
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
            int xo[4];
            {
                const int a1 = xc + 1 == NC ? 0 : xc + 1;
                xo[0] = xc == 0 ? NC - 1 : xc - 1; xo[1] = xc; xo[2] = a1; xo[3] = a1 + 1 == NC ? 0 : a1 + 1;
            }
            int ry = yc - y0;
            ry = ry < 0 ? 0 : (ry > YB - 1 ? YB - 1 : ry);
            const float *row0 = ring + ry * NC;
            float v[4][4][4];
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                int slot = s0 + a;
                slot = slot >= R ? slot - R : slot;
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
            sx[j] = pm_gp<0>(v, t); sy[j] = pm_gp<1>(v, t); sz[j] = pm_gp<2>(v, t);
        }
#pragma unroll
        for (int j = 0; j < PP; ++j) {
            pm_push(x[j], vx[j], sx[j], A.k_kick, A.da, A.aa, A.raa, A.f_a1, NC, nullptr);
            pm_push(y[j], vy[j], sy[j], A.k_kick, A.da, A.aa, A.raa, A.f_a1, NC, nullptr);
            pm_push(z[j], vz[j], sz[j], A.k_kick, A.da, A.aa, A.raa, A.f_a1, NC, nullptr);
        }
#pragma unroll
        for (int j = 0; j < PP; ++j) {
            if (ok[j]) {
                const uint32_t p = pbase + ii[j];
                A.pos_out[p] = x[j]; A.pos_out[A.sout + p] = y[j]; A.pos_out[2 * A.sout + p] = z[j];
                A.vel_out[p] = vx[j]; A.vel_out[A.sout + p] = vy[j]; A.vel_out[2 * A.sout + p] = vz[j];
                A.id_out[p] = id[j];
                const uint32_t knew = pm_key(x[j], y[j], z[j], NC, 0, NC);
                A.keys_out[p] = knew;
                if (A.mover_cnt) {
                    const unsigned act = __activemask();
                    const uint32_t tile = p / PM_SORT_TILE;
                    const unsigned same = __match_any_sync(act, tile);
                    const unsigned mv = __ballot_sync(act, knew != kold[j]) & same;
                    if (mv && (int)(threadIdx.x & 31) == __ffs(same) - 1) atomicAdd(A.mover_cnt + tile, __popc(mv));
                }
            }
        }
    };

    int s = 0;
    for (int k = 0; k < zc; ++k) {
        // phi planes pi = k .. k+3 (k = 0: all four; afterwards only the newest)
        for (int pi = (k == 0 ? 0 : k + 3); pi <= k + 3; ++pi)
            if (!mbar_wait(full_phi + pi % R, (unsigned)((pi / R) & 1), err)) return;
        const int s0 = k % R;
        const uint32_t beg = s_beg[k], n = s_end[k] - beg;
        uint32_t off = 0;
        do {
            const uint32_t cnt = min((uint32_t)CAP, n - off);
            const int st = s % S;
            if (!mbar_wait(full_p + st, (unsigned)((s / S) & 1), err)) return;
            const float *sp = stage + st * 7 * CAP;
            const uint32_t nb = (cnt + 32 * PP - 1) / (32 * PP);
            // rotate the first batch with the sub-step so the odd batch does not always hit warp 0
            for (uint32_t b = (uint32_t)((warp + CW - s % CW) % CW); b < nb; b += CW) {
                if constexpr (PP == 1) {
                    const uint32_t i = b * 32 + lane;
                    if (i < cnt)
                        update(s0, beg + off + i, sp[i], sp[CAP + i], sp[2 * CAP + i], sp[3 * CAP + i], sp[4 * CAP + i],
                               sp[5 * CAP + i], __float_as_uint(sp[6 * CAP + i]));
                } else {
                    uint32_t ii[PP];
                    bool ok[PP];
#pragma unroll
                    for (int j = 0; j < PP; ++j) {
                        ii[j] = (b * PP + j) * 32 + lane;
                        ok[j] = ii[j] < cnt;
                    }
                    if (ok[0]) update_n(s0, beg + off, ii, ok, sp);       // (ok[j] implies ok[0])
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(empty_p + st);
            ++s;
            off += CAP;
        } while (off < n);
        __syncwarp();
        if (lane == 0) mbar_arrive(empty_phi + s0);      // step k was the last reader of plane pi = k
    }
}
