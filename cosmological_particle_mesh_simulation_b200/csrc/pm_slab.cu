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
        case PM_BUF_RHO_GHOST_RECV: *ptr = p->tbuf[1]; *bytes = pb; break;
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
    return pm_k_ghost_add(p, p->mesh, reinterpret_cast<const float *>(p->tbuf[1]), st);
}

// Chunk c of C: byte ranges inside FFT_SEND_MAIN / FFT_RECV_MAIN are [c, c+1) * bytes / C.
static int chunk_ok(const pm_plan *p, int c, int C)
{
    const int h = p->nc / 2;
    return C >= 1 && c >= 0 && c < C && h % C == 0 && (h / C) % pm_fft_cols_per_tile(p->nc) == 0;
}

int pm_slab_fft_rows_forward(pm_plan *p, pm_stream_t stream)
{
    PM_SLAB_ENTER(true);
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
