// pm_slab.cu -- extern "C" entry points of the slab-decomposed (multi-GPU) step.
//
// One plan per rank; rank r owns mesh planes [r*Nc/P, (r+1)*Nc/P) along array axis 0 (z, the
// reference's positions[2]) and the particles whose z cell lies in them.  Every
// exchange is done by the caller (torch.distributed over NCCL, or plain copies between the plans
// of a single-GPU rank loop in the tests) on buffers that live inside the plans: the entry points
// below only compute, and pm_slab_buffer() hands out the addresses to send from / receive into.
//
//   pm_slab_deposit      keys -> sort -> rows -> CIC deposit into nzl+1 planes
//        exchange  RHO_GHOST_SEND (plane nzl) -> rank+1's RHO_GHOST_RECV
//   pm_slab_ghost_add    plane 0 += RHO_GHOST_RECV
//   pm_slab_fft_rows_forward   rows R2C on the local planes
//   for each of C chunks of kx columns (the chunks pipeline against each other):
//     pm_slab_fft_y_forward(c)  y pass on the chunk's columns, pack per destination rank
//        exchange  all-to-all chunk c of FFT_SEND_MAIN -> FFT_RECV_MAIN (+ SIDE with chunk 0)
//     pm_slab_fft_z(c)          z forward + Green + z inverse on the transposed chunk, in place
//        exchange  all-to-all chunk c of FFT_RECV_MAIN -> FFT_SEND_MAIN   (the way back)
//     pm_slab_fft_y_inverse(c)  unpack + y inverse
//   pm_slab_fft_rows_inverse   rows C2R -> phi planes 1..nzl
//        exchange  PHI_HI_SEND (own last plane) -> rank+1's PHI_LO_RECV (its plane 0)
//                  PHI_LO_SEND (own first two planes) -> rank-1's PHI_HI_RECV (its planes nzl+1, nzl+2)
//   pm_slab_gather       gather + kick + drift into the other buffer set; leavers -> leave lists
//        host reads LEAVE_COUNTS, exchanges counts
//   pm_slab_migrate_pack records of the leavers, per destination, ascending slot order
//        exchange  all-to-all-v MIG_SEND -> MIG_RECV (7 floats per particle)
//   pm_slab_migrate_unpack  append the arrivals and their keys
#include "pm_internal.cuh"
#include <cstring>

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
}  // namespace

#define PM_SLAB_ENTER(cond)                         \
    if (!(p && p->slab && (cond))) return PM_ERR_INVALID; \
    Guard guard_;                                   \
    {                                               \
        int rc_ = guard_.enter(p->device);          \
        if (rc_ != PM_OK) return rc_;               \
    }                                               \
    cudaStream_t st = pm_cu(stream)
#define PM_TRY(expr)                      \
    do {                                  \
        int rc_ = (expr);                 \
        if (rc_ != PM_OK) return rc_;     \
    } while (0)

extern "C" {

int pm_slab_buffer(pm_plan *p, int which, void **ptr, size_t *bytes)
{
    if (!p || !p->slab || !ptr || !bytes) return PM_ERR_INVALID;
    const size_t plane = (size_t)p->nc * p->nc, pb = plane * sizeof(float);
    const size_t h = p->nc / 2;
    const size_t main_n = (size_t)p->nzl * p->nc * h, side_n = (size_t)p->nzl * p->nc;
    switch (which) {
        case PM_BUF_RHO: *ptr = p->mesh; *bytes = pb * p->nzl; break;
        case PM_BUF_RHO_GHOST_SEND: *ptr = p->mesh + plane * p->nzl; *bytes = pb; break;
        // the phi buffer is dead between the gather and the inverse transform; the z-pass array
        // tbuf[1] cannot be used here because peers may already be storing into it (pm_slab_fft_push)
        case PM_BUF_RHO_GHOST_RECV: *ptr = p->mesh2; *bytes = pb; break;
        case PM_BUF_FFT_SEND_MAIN: *ptr = p->tbuf[0]; *bytes = main_n * sizeof(float2); break;
        case PM_BUF_FFT_SEND_SIDE: *ptr = p->tbuf[0] + main_n; *bytes = side_n * sizeof(float2); break;
        case PM_BUF_FFT_RECV_MAIN: *ptr = p->tbuf[1]; *bytes = main_n * sizeof(float2); break;
        case PM_BUF_FFT_RECV_SIDE: *ptr = p->tbuf[1] + main_n; *bytes = side_n * sizeof(float2); break;
        case PM_BUF_PHI: *ptr = p->mesh2 + plane; *bytes = pb * p->nzl; break;
        case PM_BUF_PHI_LO_SEND: *ptr = p->mesh2 + plane; *bytes = 2 * pb; break;
        case PM_BUF_PHI_HI_SEND: *ptr = p->mesh2 + plane * p->nzl; *bytes = pb; break;
        case PM_BUF_PHI_LO_RECV: *ptr = p->mesh2; *bytes = pb; break;
        case PM_BUF_PHI_HI_RECV: *ptr = p->mesh2 + plane * (p->nzl + 1); *bytes = 2 * pb; break;
        case PM_BUF_MIG_SEND: *ptr = p->mig_send; *bytes = (size_t)p->nranks * p->leave_cap * 28; break;
        case PM_BUF_MIG_RECV: *ptr = p->mig_recv; *bytes = (size_t)p->nranks * p->leave_cap * 28; break;
        case PM_BUF_LEAVE_COUNTS: *ptr = p->leave_cnt; *bytes = (size_t)p->nranks * 4; break;
        case PM_BUF_PEER_FLAGS: *ptr = p->peer_flags; *bytes = (size_t)(PM_PEER_SLOTS + 1) * PM_PEER_MAX * 4; break;
        default: return PM_ERR_INVALID;
    }
    return PM_OK;
}

int pm_slab_load(pm_plan *p, const float *pos_d, const float *vel_d, const uint32_t *ids_d, int64_t np,
                 pm_stream_t stream)
{
    PM_SLAB_ENTER(np >= 0 && np <= p->np_cap && (np == 0 || (pos_d && vel_d && ids_d)));
    p->rcur = 0;
    p->rnp = p->rtotal = np;
    p->rkeys_valid = false;
    p->rsorted_n = 0;
    if (np == 0) return PM_OK;
    const size_t w = (size_t)np * sizeof(float), pitch = (size_t)p->rstride * sizeof(float);
    PM_CUDA(cudaMemcpy2DAsync(p->rpos[0], pitch, pos_d, w, w, 3, cudaMemcpyDeviceToDevice, st));
    PM_CUDA(cudaMemcpy2DAsync(p->rvel[0], pitch, vel_d, w, w, 3, cudaMemcpyDeviceToDevice, st));
    PM_CUDA(cudaMemcpyAsync(p->rid[0], ids_d, (size_t)np * 4, cudaMemcpyDeviceToDevice, st));
    return PM_OK;
}

int64_t pm_slab_count(const pm_plan *p) { return p ? p->rnp : 0; }
int64_t pm_slab_entries(const pm_plan *p) { return p ? p->rtotal : 0; }

int pm_slab_deposit(pm_plan *p, double mass, pm_stream_t stream)
{
    PM_SLAB_ENTER(true);
    const int64_t n = p->rtotal;
    if (!p->rkeys_valid) {
        PM_TRY(pm_k_cell_keys(p, p->rpos[p->rcur], n, p->rstride, p->keys, nullptr, st));
        p->rsorted_n = 0;
    }
    PM_TRY(pm_k_sort(p, n, p->rsorted_n, st));
    PM_TRY(pm_k_row_offsets(p, n, st));
    return pm_k_deposit_slab(p, p->rpos[p->rcur], mass, p->mesh, st);
}

int pm_slab_ghost_add(pm_plan *p, pm_stream_t stream)
{
    PM_SLAB_ENTER(true);
    return pm_k_ghost_add(p, p->mesh, p->mesh2, st);
}

// Chunk c of C: byte ranges inside FFT_SEND_MAIN / FFT_RECV_MAIN are [c, c+1) * bytes / C.
static int chunk_ok(const pm_plan *p, int c, int C)
{
    const int h = p->nc / 2;
    return C >= 1 && c >= 0 && c < C && h % C == 0 && (h / C) % pm_fft_cols_per_tile(p->nc) == 0;
}

// <rho> of the WHOLE mesh (all ranks): total particles * mass / Nc^3, which only the caller knows.
// The forward transform subtracts it (pm_internal.cuh, rho_mean_d); unset = 0 = plain transform.
int pm_slab_set_rho_mean(pm_plan *p, double mean)
{
    if (!p || !p->slab || !(mean == mean)) return PM_ERR_INVALID;
    p->rho_mean_hint = mean;
    return PM_OK;
}

int pm_slab_fft_rows_forward(pm_plan *p, pm_stream_t stream)
{
    PM_SLAB_ENTER(true);
    if (p->rho_mean_hint != p->rho_mean_hint) p->rho_mean_hint = 0.0;   // NaN: never told
    PM_TRY(pm_k_rho_mean(p, p->mesh, 0, 1.0, st));
    return pm_k_fft_slab_rows_fwd(p, p->mesh, st);
}

int pm_slab_fft_y_forward(pm_plan *p, int c, int C, pm_stream_t stream)
{
    PM_SLAB_ENTER(chunk_ok(p, c, C));
    const size_t main_n = (size_t)p->nzl * p->nc * (p->nc / 2);
    return pm_k_fft_slab_y_fwd_pack(p, c, C, p->tbuf[0] + main_n / C * c, p->tbuf[0] + main_n, st);
}

int pm_slab_fft_z(pm_plan *p, int c, int C, double a, double omega_m0, pm_stream_t stream)
{
    PM_SLAB_ENTER(a != 0.0 && chunk_ok(p, c, C));
    const size_t main_n = (size_t)p->nzl * p->nc * (p->nc / 2);
    return pm_k_fft_slab_z_chunk(p, c, C, p->tbuf[1] + main_n / C * c, p->tbuf[1] + main_n, a, omega_m0, st);
}

int pm_slab_fft_y_inverse(pm_plan *p, int c, int C, pm_stream_t stream)
{
    PM_SLAB_ENTER(chunk_ok(p, c, C));
    const size_t main_n = (size_t)p->nzl * p->nc * (p->nc / 2);
    return pm_k_fft_slab_unpack_y_inv(p, c, C, p->tbuf[0] + main_n / C * c, p->tbuf[0] + main_n, st);
}

int pm_slab_fft_rows_inverse(pm_plan *p, pm_stream_t stream)
{
    PM_SLAB_ENTER(true);
    return pm_k_fft_slab_rows_inv(p, p->mesh2 + (size_t)p->nc * p->nc, st);
}

// ---- peer-memory transposes ------------------------------------------------------------------
// Every rank publishes where its z-pass array (tbuf[1]) and its flag words live; after that
//   pm_slab_fft_y_forward_local(c)   y pass of chunk c, left in the local spectrum
//   pm_slab_fft_push(c)              local spectrum chunk -> block [rank] of EVERY peer's z-pass array
//   pm_slab_peer_signal(c)           "my blocks of chunk c have landed" -> flag word on every peer
//   pm_slab_peer_wait(c)             spin until all P flags of chunk c carry this step's epoch
//   pm_slab_fft_z(c)                 unchanged
//   pm_slab_peer_signal(SLOTS/2+c)   "my z pass of chunk c is done"
//   pm_slab_peer_wait(SLOTS/2+c); pm_slab_fft_pull(c)   every peer's block [rank] -> local spectrum
//   pm_slab_fft_y_inverse_local(c)
// No staging buffer, no collective: the all-to-all is the stores / loads of the two copy kernels.
int pm_slab_peer_export(pm_plan *p, void *handle64, uint64_t *recv_offset, uint64_t *flags_offset)
{
    if (!p || !p->slab || !handle64 || !recv_offset || !flags_offset) return PM_ERR_INVALID;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
    Guard g;
    int rc = g.enter(p->device);
    if (rc != PM_OK) return rc;
    cudaIpcMemHandle_t h;
    PM_CUDA(cudaIpcGetMemHandle(&h, p->ws));
    memcpy(handle64, &h, 64);
    *recv_offset = (uint64_t)((char *)p->tbuf[1] - p->ws);
    *flags_offset = (uint64_t)((char *)p->peer_flags - p->ws);
    return PM_OK;
}

static int peer_store(pm_plan *p, int peer, void *recv, void *flags)
{
    if (!p->peer_recv[peer]) ++p->peers_set;
    p->peer_recv[peer] = (float2 *)recv;
    p->peer_flag_of[peer] = (uint32_t *)flags;
    return PM_OK;
}

int pm_slab_peer_import(pm_plan *p, int peer, const void *handle64, uint64_t recv_offset,
                        uint64_t flags_offset)
{
    if (!p || !p->slab || peer < 0 || peer >= p->nranks || peer >= PM_PEER_MAX) return PM_ERR_INVALID;
    if (peer == p->rank) return peer_store(p, peer, p->tbuf[1], p->peer_flags);
    if (!handle64) return PM_ERR_INVALID;
    Guard g;
    int rc = g.enter(p->device);
    if (rc != PM_OK) return rc;
    if (p->peer_ipc[peer]) {
        cudaIpcCloseMemHandle(p->peer_ipc[peer]);
        p->peer_ipc[peer] = nullptr;
    }
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    void *base = nullptr;
    PM_CUDA(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
    p->peer_ipc[peer] = base;
    return peer_store(p, peer, (char *)base + recv_offset, (char *)base + flags_offset);
}

// Plans of several ranks inside ONE process (single-GPU rank loop of the tests, or one process
// driving several GPUs with peer access enabled): plain device pointers.
int pm_slab_peer_set(pm_plan *p, int peer, void *recv_main, void *flags)
{
    if (!p || !p->slab || peer < 0 || peer >= p->nranks || peer >= PM_PEER_MAX || !recv_main || !flags)
        return PM_ERR_INVALID;
    return peer_store(p, peer, recv_main, flags);
}

int pm_slab_peer_signal(pm_plan *p, int slot, pm_stream_t stream)
{
    PM_SLAB_ENTER(slot >= 0 && slot < PM_PEER_SLOTS && p->peers_set == p->nranks);
    return pm_k_peer_signal(p, slot, ++p->peer_epoch_sig[slot], st);
}

int pm_slab_peer_wait(pm_plan *p, int slot, pm_stream_t stream)
{
    PM_SLAB_ENTER(slot >= 0 && slot < PM_PEER_SLOTS && p->peers_set == p->nranks);
    return pm_k_peer_wait(p, slot, ++p->peer_epoch_wait[slot], st);
}

// Unmap the other ranks' workspaces (every rank does this, then a barrier, before any plan is
// destroyed: freeing memory that another process still has mapped is undefined).
int pm_slab_peer_release(pm_plan *p)
{
    if (!p || !p->slab) return PM_ERR_INVALID;
    Guard g;
    int rc = g.enter(p->device);
    if (rc != PM_OK) return rc;
    cudaStreamSynchronize(0);
    for (int s = 0; s < PM_PEER_MAX; ++s) {
        if (p->peer_ipc[s]) cudaIpcCloseMemHandle(p->peer_ipc[s]);
        p->peer_ipc[s] = nullptr;
        p->peer_recv[s] = nullptr;
        p->peer_flag_of[s] = nullptr;
        p->peer_mesh2[s] = nullptr;
        p->peer_mig_matrix[s] = nullptr;
        p->peer_mig_recv[s] = nullptr;
    }
    p->peers_set = 0;
    p->ghosts_set = 0;
    p->aux_set = 0;
    return PM_OK;
}

// Number of pm_slab_peer_wait calls that gave up (2 s without the peers' flags): 0 on a healthy run.
int pm_slab_peer_timeouts(pm_plan *p, uint32_t *timeouts)
{
    if (!p || !p->slab || !timeouts) return PM_ERR_INVALID;
    Guard g;
    int rc = g.enter(p->device);
    if (rc != PM_OK) return rc;
    PM_CUDA(cudaMemcpy(timeouts, p->peer_flags + (size_t)PM_PEER_SLOTS * PM_PEER_MAX, 4, cudaMemcpyDeviceToHost));
    return PM_OK;
}

// ---- ghost planes through the same peer mappings (EXPERIMENTAL: not yet validated on hardware) ----
// Instead of ncclSend/Recv: copy the plane into the neighbour's phi buffer, then one flag word to that
// neighbour; the neighbour waits for that one word.  The density ghost lands in the neighbour's phi
// plane 0 (dead at that time), exactly where the NCCL path receives it.
int pm_slab_peer_ghost_export(pm_plan *p, uint64_t *mesh2_offset)
{
    if (!p || !p->slab || !mesh2_offset) return PM_ERR_INVALID;
    *mesh2_offset = (uint64_t)((char *)p->mesh2 - p->ws);
    return PM_OK;
}

static int ghost_store(pm_plan *p, int peer, void *mesh2)
{
    if (!p->peer_mesh2[peer]) ++p->ghosts_set;
    p->peer_mesh2[peer] = (float *)mesh2;
    return PM_OK;
}

// after pm_slab_peer_import(peer): uses that mapping
int pm_slab_peer_ghost_import(pm_plan *p, int peer, uint64_t mesh2_offset)
{
    if (!p || !p->slab || peer < 0 || peer >= p->nranks || peer >= PM_PEER_MAX) return PM_ERR_INVALID;
    if (peer == p->rank) return ghost_store(p, peer, p->mesh2);
    if (!p->peer_ipc[peer]) return PM_ERR_INVALID;
    return ghost_store(p, peer, (char *)p->peer_ipc[peer] + mesh2_offset);
}

int pm_slab_peer_ghost_set(pm_plan *p, int peer, void *mesh2)
{
    if (!p || !p->slab || peer < 0 || peer >= p->nranks || peer >= PM_PEER_MAX || !mesh2) return PM_ERR_INVALID;
    return ghost_store(p, peer, mesh2);
}

#define PM_GHOST_READY (p->peers_set == p->nranks && p->ghosts_set == p->nranks)

int pm_slab_ghost_push_rho(pm_plan *p, pm_stream_t stream)
{
    PM_SLAB_ENTER(PM_GHOST_READY);
    const size_t plane = (size_t)p->nc * p->nc;
    const int up = (p->rank + 1) % p->nranks;
    PM_TRY(pm_k_peer_put(p, p->peer_mesh2[up], p->mesh + plane * p->nzl, plane, st));
    return pm_k_peer_signal_to(p, PM_SLOT_GHOST_RHO, up, ++p->peer_epoch_sig[PM_SLOT_GHOST_RHO], st);
}

int pm_slab_ghost_wait_rho(pm_plan *p, pm_stream_t stream)
{
    PM_SLAB_ENTER(PM_GHOST_READY);
    const int dn = (p->rank + p->nranks - 1) % p->nranks;
    return pm_k_peer_wait_from(p, PM_SLOT_GHOST_RHO, dn, ++p->peer_epoch_wait[PM_SLOT_GHOST_RHO], st);
}

int pm_slab_ghost_push_phi(pm_plan *p, pm_stream_t stream)
{
    PM_SLAB_ENTER(PM_GHOST_READY);
    const size_t plane = (size_t)p->nc * p->nc;
    const int up = (p->rank + 1) % p->nranks, dn = (p->rank + p->nranks - 1) % p->nranks;
    // my last owned plane is rank+1's plane z0-1; my first two owned planes close rank-1's stencil
    PM_TRY(pm_k_peer_put(p, p->peer_mesh2[up], p->mesh2 + plane * p->nzl, plane, st));
    PM_TRY(pm_k_peer_signal_to(p, PM_SLOT_GHOST_PHI_UP, up, ++p->peer_epoch_sig[PM_SLOT_GHOST_PHI_UP], st));
    PM_TRY(pm_k_peer_put(p, p->peer_mesh2[dn] + plane * (p->nzl + 1), p->mesh2 + plane, 2 * plane, st));
    return pm_k_peer_signal_to(p, PM_SLOT_GHOST_PHI_DN, dn, ++p->peer_epoch_sig[PM_SLOT_GHOST_PHI_DN], st);
}

int pm_slab_ghost_wait_phi(pm_plan *p, pm_stream_t stream)
{
    PM_SLAB_ENTER(PM_GHOST_READY);
    const int up = (p->rank + 1) % p->nranks, dn = (p->rank + p->nranks - 1) % p->nranks;
    PM_TRY(pm_k_peer_wait_from(p, PM_SLOT_GHOST_PHI_UP, dn, ++p->peer_epoch_wait[PM_SLOT_GHOST_PHI_UP], st));
    return pm_k_peer_wait_from(p, PM_SLOT_GHOST_PHI_DN, up, ++p->peer_epoch_wait[PM_SLOT_GHOST_PHI_DN], st);
}

int pm_slab_fft_y_forward_local(pm_plan *p, int c, int C, pm_stream_t stream)
{
    PM_SLAB_ENTER(chunk_ok(p, c, C));
    return pm_k_fft_slab_y_fwd(p, c, C, st);
}

int pm_slab_fft_push(pm_plan *p, int c, int C, pm_stream_t stream)
{
    PM_SLAB_ENTER(chunk_ok(p, c, C) && p->peers_set == p->nranks);
    return pm_k_fft_slab_push(p, c, C, st);
}

int pm_slab_fft_pull(pm_plan *p, int c, int C, pm_stream_t stream)
{
    PM_SLAB_ENTER(chunk_ok(p, c, C) && p->peers_set == p->nranks);
    return pm_k_fft_slab_pull(p, c, C, st);
}

// The same two legs fused into the y passes: the forward y pass stores its output into the peers'
// z-pass arrays, the inverse y pass loads its input from them (pm_fft.cu, PeerArgs).
int pm_slab_fft_y_forward_push(pm_plan *p, int c, int C, pm_stream_t stream)
{
    PM_SLAB_ENTER(chunk_ok(p, c, C) && p->peers_set == p->nranks);
    return pm_k_fft_slab_y_fwd_push(p, c, C, st);
}

int pm_slab_fft_y_inverse_pull(pm_plan *p, int c, int C, pm_stream_t stream)
{
    PM_SLAB_ENTER(chunk_ok(p, c, C) && p->peers_set == p->nranks);
    return pm_k_fft_slab_y_inv_pull(p, c, C, st);
}

int pm_slab_fft_y_inverse_local(pm_plan *p, int c, int C, pm_stream_t stream)
{
    PM_SLAB_ENTER(chunk_ok(p, c, C));
    return pm_k_fft_slab_y_inv(p, c, C, st);
}

int pm_slab_gather(pm_plan *p, double a, double f_a1, double da, pm_stream_t stream)
{
    PM_SLAB_ENTER(true);
    PM_TRY(pm_k_gather_kick_drift_slab(p, p->mesh2, a, f_a1, da, st));
    p->rcur ^= 1;
    p->rtotal = p->rnp;       // the kernel wrote the live particles, leavers now carry dead keys
    p->rkeys_valid = true;
    p->rsorted_n = p->rnp;    // ... in the order of this step's sort (arrivals are appended behind)
    return PM_OK;
}

int pm_slab_migrate_pack(pm_plan *p, const int64_t *counts_h, pm_stream_t stream)
{
    PM_SLAB_ENTER(counts_h != nullptr);
    int64_t off = 0;
    for (int d = 0; d < p->nranks; ++d) {
        if (counts_h[d] < 0 || counts_h[d] > p->leave_cap) return PM_ERR_UNSUPPORTED;  // list overflow
        PM_TRY(pm_k_migrate_pack(p, d, counts_h[d], off, st));
        off += counts_h[d];
    }
    return PM_OK;
}

int pm_slab_migrate_unpack(pm_plan *p, int64_t n_arrive, int64_t n_leave, pm_stream_t stream)
{
    PM_SLAB_ENTER(n_arrive >= 0 && n_leave >= 0 && n_leave <= p->rnp);
    if (p->rtotal + n_arrive > p->np_cap) return PM_ERR_UNSUPPORTED;   // particle capacity exceeded
    PM_TRY(pm_k_migrate_unpack(p, n_arrive, st));
    p->rtotal += n_arrive;
    p->rnp += n_arrive - n_leave;
    return PM_OK;
}

int pm_slab_export(pm_plan *p, float *pos_d, float *vel_d, uint32_t *ids_d, uint32_t *live_d,
                   pm_stream_t stream)
{
    PM_SLAB_ENTER(p->rtotal == 0 || (pos_d && vel_d && ids_d && live_d));
    return pm_k_export(p, pos_d, vel_d, ids_d, live_d, st);
}

}  // extern "C"
