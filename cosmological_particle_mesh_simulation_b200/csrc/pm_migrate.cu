// pm_migrate.cu -- particle migration between slabs through peer memory (no collective, one host read).
//
// After the drift a rank holds particles whose z cell now belongs to another rank (the
// reference is single-process and has no counterpart).  pm_slab_gather left, per destination, the
// number of leavers (LEAVE_COUNTS) and their storage slots.  The NCCL route (slab.py) then needs an
// all-to-all of the counts, a device->host read and an all-to-all-v of the records.  Here, once the
// peers' buffers are mapped (pm_slab_peer_import + pm_slab_peer_aux_import, or the *_set variants):
//   pm_slab_migrate_counts_push   one kernel: my row of the count matrix -- leavers per destination, my
//                                 flag-wait timeout counter, my list-overflow flag, my free capacity --
//                                 stored into EVERY rank's matrix, then this step's flag word
//   pm_slab_migrate_counts_read   wait for all rows, ONE copy of the matrix to the host.  Every rank
//                                 now holds the same matrix: message sizes and offsets need no further
//                                 exchange, and an overflow or a timed-out wait anywhere is seen -- and
//                                 raised -- by all ranks in the same step
//   pm_slab_migrate_push          leave lists put in ascending slot order (bitonic sort in shared
//                                 memory: the lists are short), records packed STRAIGHT into the
//                                 destination's receive buffer at the offset the matrix implies, flag
//   pm_slab_migrate_wait          wait for every sender's flag; pm_slab_migrate_unpack as before
#include <string.h>

#include "pm_internal.cuh"

namespace {

struct Guard {
    int prev = -1;
    bool active = false;
    int enter(int dev)
    {
        if (cudaGetDevice(&prev) != cudaSuccess) return PM_ERR_NO_DEVICE;
        if (dev != prev) {
            cudaError_t e = cudaSetDevice(dev);
            if (e != cudaSuccess) return (int)e;
            active = true;
        }
        return PM_OK;
    }
    ~Guard()
    {
        if (active) cudaSetDevice(prev);
    }
};

struct MatPtrs {
    uint32_t *mat[PM_PEER_MAX];     // every rank's count matrix
    uint32_t *flag[PM_PEER_MAX];    // every rank's flag words
};

// thread s: my row into rank s's matrix, then my flag word of this step on rank s
__global__ void k_mig_counts_push(MatPtrs P, int nranks, int rank, const uint32_t *__restrict__ leave_cnt,
                                  const uint32_t *__restrict__ timeouts, uint32_t leave_cap, uint32_t room, int slot,
                                  uint32_t epoch)
{
    const int s = threadIdx.x;
    if (s >= nranks) return;
    uint32_t *row = P.mat[s] + (size_t)rank * PM_MIG_ROW;
    uint32_t over = 0;
    for (int d = 0; d < nranks; ++d) {
        const uint32_t c = leave_cnt[d];
        row[d] = c;
        over |= c > leave_cap ? 1u : 0u;
    }
    row[PM_PEER_MAX + 0] = *timeouts;
    row[PM_PEER_MAX + 1] = over;
    row[PM_PEER_MAX + 2] = room;
    __threadfence_system();
    volatile uint32_t *w = P.flag[s] + (size_t)slot * PM_PEER_MAX + rank;
    *w = epoch;
}

// One CTA per destination: its leave list (<= 4096 slots) in ascending order, in place.
__global__ void __launch_bounds__(1024) k_leave_sort(uint32_t *__restrict__ leave_slot, const uint32_t *__restrict__ leave_cnt,
                                                     int64_t leave_cap)
{
    __shared__ uint32_t s[4096];
    const int d = blockIdx.x;
    uint32_t n = leave_cnt[d];
    if ((int64_t)n > leave_cap) n = (uint32_t)leave_cap;
    if (n < 2u || n > 4096u) return;          // longer lists: the host sorts them with the radix sort
    uint32_t *list = leave_slot + (size_t)d * leave_cap;
    uint32_t m = 2;
    while (m < n) m <<= 1;
    for (uint32_t i = threadIdx.x; i < m; i += 1024) s[i] = i < n ? list[i] : 0xffffffffu;
    __syncthreads();
    for (uint32_t k = 2; k <= m; k <<= 1)
        for (uint32_t j = k >> 1; j > 0; j >>= 1) {
            for (uint32_t i = threadIdx.x; i < m; i += 1024) {
                const uint32_t l = i ^ j;
                if (l > i) {
                    const uint32_t a = s[i], b = s[l];
                    const bool up = (i & k) == 0;
                    if ((a > b) == up) {
                        s[i] = b;
                        s[l] = a;
                    }
                }
            }
            __syncthreads();
        }
    for (uint32_t i = threadIdx.x; i < n; i += 1024) list[i] = s[i];
}

__global__ void __launch_bounds__(256) k_migrate_pack_to(const float *__restrict__ pos, const float *__restrict__ vel,
                                                         const uint32_t *__restrict__ id, int64_t stride,
                                                         const uint32_t *__restrict__ slots, int64_t n, float *rec)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t s = slots[i];
    float *r = rec + i * 7;
    r[0] = pos[s]; r[1] = pos[stride + s]; r[2] = pos[2 * stride + s];
    r[3] = vel[s]; r[4] = vel[stride + s]; r[5] = vel[2 * stride + s];
    r[6] = __uint_as_float(id[s]);
}

}  // namespace

#define PM_MIG_ENTER(cond)                                   \
    if (!(p && p->slab && (cond))) return PM_ERR_INVALID;    \
    Guard guard_;                                            \
    {                                                        \
        int rc_ = guard_.enter(p->device);                   \
        if (rc_ != PM_OK) return rc_;                        \
    }                                                        \
    cudaStream_t st = pm_cu(stream)
#define PM_TRY(expr)                      \
    do {                                  \
        int rc_ = (expr);                 \
        if (rc_ != PM_OK) return rc_;     \
    } while (0)
#define PM_MIG_READY (p->peers_set == p->nranks && p->aux_set == p->nranks)

extern "C" {

// offsets[0..2]: phi buffer (ghost planes), migration receive buffer, count matrix -- inside the workspace
// that pm_slab_peer_export published
int pm_slab_peer_aux_export(pm_plan *p, uint64_t *offsets)
{
    if (!p || !p->slab || !offsets) return PM_ERR_INVALID;
    offsets[0] = (uint64_t)((char *)p->mesh2 - p->ws);
    offsets[1] = (uint64_t)((char *)p->mig_recv - p->ws);
    offsets[2] = (uint64_t)((char *)p->mig_matrix - p->ws);
    return PM_OK;
}

static int aux_store(pm_plan *p, int peer, void *mesh2, void *mig_recv, void *matrix)
{
    if (!p->peer_mig_matrix[peer]) ++p->aux_set;
    if (!p->peer_mesh2[peer]) ++p->ghosts_set;
    p->peer_mesh2[peer] = (float *)mesh2;
    p->peer_mig_recv[peer] = (float *)mig_recv;
    p->peer_mig_matrix[peer] = (uint32_t *)matrix;
    return PM_OK;
}

int pm_slab_peer_aux_import(pm_plan *p, int peer, const uint64_t *offsets)
{
    if (!p || !p->slab || !offsets || peer < 0 || peer >= p->nranks || peer >= PM_PEER_MAX) return PM_ERR_INVALID;
    if (peer == p->rank) return aux_store(p, peer, p->mesh2, p->mig_recv, p->mig_matrix);
    if (!p->peer_ipc[peer]) return PM_ERR_INVALID;
    char *base = (char *)p->peer_ipc[peer];
    return aux_store(p, peer, base + offsets[0], base + offsets[1], base + offsets[2]);
}

int pm_slab_peer_aux_set(pm_plan *p, int peer, void *mesh2, void *mig_recv, void *matrix)
{
    if (!p || !p->slab || peer < 0 || peer >= p->nranks || peer >= PM_PEER_MAX || !mesh2 || !mig_recv || !matrix)
        return PM_ERR_INVALID;
    return aux_store(p, peer, mesh2, mig_recv, matrix);
}

int pm_slab_aux_buffers(pm_plan *p, void **mesh2, void **mig_recv, void **matrix)
{
    if (!p || !p->slab || !mesh2 || !mig_recv || !matrix) return PM_ERR_INVALID;
    *mesh2 = p->mesh2;
    *mig_recv = p->mig_recv;
    *matrix = p->mig_matrix;
    return PM_OK;
}

int pm_slab_migrate_counts_push(pm_plan *p, pm_stream_t stream)
{
    PM_MIG_ENTER(PM_MIG_READY);
    MatPtrs P;
    for (int s = 0; s < PM_PEER_MAX; ++s) {
        P.mat[s] = s < p->nranks ? p->peer_mig_matrix[s] : nullptr;
        P.flag[s] = s < p->nranks ? p->peer_flag_of[s] : nullptr;
    }
    const int64_t room = p->np_cap - p->rtotal;
    PM_LAUNCH(k_mig_counts_push, 1, 32, 0, st, P, p->nranks, p->rank, (const uint32_t *)p->leave_cnt,
              (const uint32_t *)(p->peer_flags + (size_t)PM_PEER_SLOTS * PM_PEER_MAX), (uint32_t)p->leave_cap,
              (uint32_t)(room < 0 ? 0 : room), PM_SLOT_MIG_COUNTS, ++p->peer_epoch_sig[PM_SLOT_MIG_COUNTS]);
    PM_CHECK_LAUNCH();
    return PM_OK;
}

// matrix_h: [nranks][nranks + 3] -- row s = rank s's leavers per destination, then its timeout count, its
// overflow flag and its free particle capacity.  Synchronises the stream.
int pm_slab_migrate_counts_read(pm_plan *p, uint32_t *matrix_h, pm_stream_t stream)
{
    PM_MIG_ENTER(PM_MIG_READY && matrix_h);
    PM_TRY(pm_k_peer_wait(p, PM_SLOT_MIG_COUNTS, ++p->peer_epoch_wait[PM_SLOT_MIG_COUNTS], st));
    PM_CUDA(cudaMemcpyAsync(p->mig_matrix_h, p->mig_matrix, sizeof(uint32_t) * PM_PEER_MAX * PM_MIG_ROW,
                            cudaMemcpyDeviceToHost, st));
    PM_CUDA(cudaStreamSynchronize(st));
    const int P = p->nranks;
    for (int s = 0; s < P; ++s) {
        for (int d = 0; d < P; ++d) matrix_h[(size_t)s * (P + 3) + d] = p->mig_matrix_h[(size_t)s * PM_MIG_ROW + d];
        for (int e = 0; e < 3; ++e) matrix_h[(size_t)s * (P + 3) + P + e] = p->mig_matrix_h[(size_t)s * PM_MIG_ROW + PM_PEER_MAX + e];
    }
    return PM_OK;
}

// send_counts[d]: my leavers to rank d; dest_offsets[d]: record offset of my block inside rank d's receive
// buffer (the leavers of the ranks below me towards d) -- both from the matrix every rank holds.
int pm_slab_migrate_push(pm_plan *p, const int64_t *send_counts, const int64_t *dest_offsets, pm_stream_t stream)
{
    PM_MIG_ENTER(PM_MIG_READY && send_counts && dest_offsets);
    bool any = false, big = false;
    for (int d = 0; d < p->nranks; ++d) {
        if (send_counts[d] < 0 || send_counts[d] > p->leave_cap || dest_offsets[d] < 0 ||
            dest_offsets[d] + send_counts[d] > (int64_t)p->nranks * p->leave_cap)
            return PM_ERR_UNSUPPORTED;
        any |= send_counts[d] > 1;
        big |= send_counts[d] > 4096;
    }
    if (any) PM_LAUNCH(k_leave_sort, p->nranks, 1024, 0, st, p->leave_slot, (const uint32_t *)p->leave_cnt, p->leave_cap);
    const int c = p->rcur;
    for (int d = 0; d < p->nranks; ++d) {
        const int64_t n = send_counts[d];
        if (n == 0) continue;
        const uint32_t *slots = p->leave_slot + (size_t)d * p->leave_cap;
        if (big && n > 4096) {
            PM_TRY(pm_k_sort_u32(p, slots, p->leave_sorted, n, st));
            slots = p->leave_sorted;
        }
        PM_LAUNCH(k_migrate_pack_to, (unsigned)((n + 255) / 256), 256, 0, st, p->rpos[c], p->rvel[c], p->rid[c],
                  p->rstride, slots, n, p->peer_mig_recv[d] + dest_offsets[d] * 7);
    }
    PM_CHECK_LAUNCH();
    return pm_k_peer_signal(p, PM_SLOT_MIG_DATA, ++p->peer_epoch_sig[PM_SLOT_MIG_DATA], st);
}

int pm_slab_migrate_wait(pm_plan *p, pm_stream_t stream)
{
    PM_MIG_ENTER(PM_MIG_READY);
    return pm_k_peer_wait(p, PM_SLOT_MIG_DATA, ++p->peer_epoch_wait[PM_SLOT_MIG_DATA], st);
}

}  // extern "C"
