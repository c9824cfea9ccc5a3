// pm_internal.cuh -- shared declarations of libpmstep.so (not part of the public ABI).
#pragma once

#include <cuda_runtime.h>
#include <cufft.h>
#include <stdint.h>

#include "pmstep.h"

// Every kernel launch of this library goes through PM_LAUNCH so that pm_launch_count() is an
// honest count (bench.py reports it as gpu_launches).
extern unsigned long long g_pm_launches;
#define PM_LAUNCH(kernel, grid, block, smem, stream, ...)              \
    do {                                                               \
        kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);    \
        ++g_pm_launches;                                               \
    } while (0)

#define PM_CUDA(expr)                                   \
    do {                                                \
        cudaError_t e_ = (expr);                        \
        if (e_ != cudaSuccess) return (int)e_;          \
    } while (0)

#define PM_CHECK_LAUNCH()                               \
    do {                                                \
        cudaError_t e_ = cudaPeekAtLastError();         \
        if (e_ != cudaSuccess) return (int)e_;          \
    } while (0)

extern int g_pm_last_cufft;
#define PM_CUFFT(expr)                                  \
    do {                                                \
        cufftResult r_ = (expr);                        \
        if (r_ != CUFFT_SUCCESS) {                      \
            g_pm_last_cufft = (int)r_;                  \
            return PM_ERR_CUFFT;                        \
        }                                               \
    } while (0)

// Function attributes (dynamic shared-memory opt-in, carveout) are per DEVICE: guard their one-time
// set-up by a bit per device ordinal, not by a process-wide flag, so that a plan on a second GPU of
// the same process opts in there as well.  Setting them twice from two threads is harmless.
#include <atomic>
#define PM_ONCE_PER_DEVICE_BEGIN(dev)                                                   \
    {                                                                                   \
        static std::atomic<unsigned long long> once_{0ull};                             \
        const unsigned long long bit_ = 1ull << ((unsigned)(dev) & 63u);                \
        if (!(once_.load(std::memory_order_acquire) & bit_)) {
#define PM_ONCE_PER_DEVICE_END()                                                        \
            once_.fetch_or(bit_, std::memory_order_release);                            \
        }                                                                               \
    }

#define PM_PEER_MAX 16
#define PM_PEER_SLOTS 24      // 0..7 "chunk pushed", 8..15 "z pass done", 16..18 ghost planes (pm_slab_ghost_*)
#define PM_SLOT_GHOST_RHO 16  // density ghost plane written into rank+1
#define PM_SLOT_GHOST_PHI_UP 17   // my last phi plane written into rank+1
#define PM_SLOT_GHOST_PHI_DN 18   // my first two phi planes written into rank-1
#define PM_SLOT_MIG_COUNTS 19     // my row of the migration count matrix written into every rank (pm_migrate.cu)
#define PM_SLOT_MIG_DATA 20       // my leavers' records written into their destinations
#define PM_MIG_ROW (PM_PEER_MAX + 4)   // words per matrix row: counts per destination, timeouts, overflow, room, pad

// Per-step scalars of a resident step as read from DEVICE memory by the kernels of a captured CUDA graph
// (k_fft_cols<FUSED>: green_scale; the tiled gather kernels: the rest).  One small kernel node writes
// them; its by-value argument is the only thing that changes between replays (pm_api.cu).
struct PmStepParams {
    double k_kick, da, aa, raa, f_a1;
    float green_scale;
    float pad;
};

#define PM_HOST_CHUNKS 4
#define PM_DEP_MAX_SLOTS 1024     // crowded deposit tiles per step that get a scratch slot (pm_deposit_tiles.cuh)
struct pm_plan {
    int nc;           // N_CELLS
    int64_t np_cap;   // particle capacity
    int device;
    int sm_count;
    int key_bits;     // bits of the largest cell key (+ the DEAD key in slab mode)

    // slab decomposition along array axis 0 (z): this rank owns planes [z0, z0 + nzl)
    int rank, nranks, nzl, z0;
    bool slab;

    char *ws;         // one workspace allocation; everything below points into it
    size_t ws_bytes;

    // deposit scratch
    uint32_t *keys, *iota, *keys_sorted, *order_sorted;  // iota[i] = i, written once
    uint64_t *sort_tmp;     // second ping-pong buffer of the full radix sort (the first is inc_a)
    void *sort_ctl;         // SortCtl (pm_sort.cu): mover count, mode, error flag, barrier counter -- device resident
    uint32_t *sort_hist;    // per-CTA digit histograms of k_radix_sort [PM_SORT_MAX_GRID][<=512], their column prefixes, column totals
    uint32_t *row_start;  // nzl*nc*dep_nseg + 1 offsets into the sorted particle list

    // incremental sort (pm_sort.cu): stayers / movers as (key << 32 | slot), tile tables
    uint64_t *inc_a, *inc_b, *inc_bs;
    int64_t inc_bcap;       // capacity of inc_b / inc_bs in entries
    uint32_t *inc_tile;     // movers per 2048-entry tile -> exclusive offsets; [ntiles] = total
    uint32_t *inc_split;    // merge-path split of every output tile boundary
    int sort_mode;          // PM_SORT_AUTO / PM_SORT_FULL
    int64_t rsorted_n;      // the first rsorted_n entries of set rcur are stored in the order of
                            // the previous sort and keys_sorted[] still holds the keys they had
    bool gather_ws;         // with gather_tiled: the warp-specialised mbarrier/bulk-copy kernel (pm_gather_ws.cuh); PM_GATHER_WS=0: k_gather_tiled
    bool gather_tiled;      // resident gather through shared-memory phi slabs (pm_gather_tiled.cuh) where supported
    bool inc_counted;       // inc_tile already holds this step's movers per tile (counted by the resident gather)
    bool rows_valid;        // row_start matches keys_sorted (set by the sort's merge or pm_k_row_offsets)
    int64_t sort_last_n;    // entries of the last pm_k_sort (mode and mover count live in sort_ctl)
    int dep_nseg;         // segments per mesh row in the deposit (1 unless the mesh is wide)

    // Poisson scratch
    float *mesh;          // nc^3: rho when the caller does not keep it
    float *mesh2;         // nc^3: phi
    float2 *spec;         // nc*nc*(nc/2+1) half spectrum
    // The forward transform runs on rho - <rho> (the DC mode is zeroed by the Green's factor, so the
    // potential is the same in exact arithmetic; in float32 the 1e8-sized DC lineage otherwise leaks
    // rounding noise into the lowest-k modes, which the 1/k^2 factor amplifies: measured 1.6e-5
    // relative L2 in phi at 512^3 with it, < 3e-6 without).  rho_mean_d: the mean as a device float
    // read by the row-pass kernels; rho_mean_hint: what the caller knows it to be (Np*mass/Nc^3 for a
    // deposit of this library), NaN = unknown -> pm_k_poisson reduces the mesh (deterministic sum).
    float *rho_mean_d;
    double *mean_ws;      // [1024] block partial sums of that reduction
    double rho_mean_hint;
    void *fft_work;
    size_t fft_work_bytes;
    float *sin2;          // sin^2(pi i / nc), i < nc
    float *sin2rev;       // the same in the digit-reversed order of the hand-written FFT
    float2 *tw;           // exp(-2 pi i m / nc)
    bool own_fft;         // power-of-two mesh: pm_fft.cu path; otherwise cuFFT
    bool fft_v2;          // two-stage register-resident transforms (pm_fft2.cuh), meshes 256..1024
    bool fft_zmix;        // with fft_v2: the fused z pass still runs the radix-8.8.8 kernel (faster at 80 registers)
    int fft_v3;           // with fft_v2: y passes by the cp.async-pipelined persistent kernel (pm_fft3.cuh); value = ring depth (0 off, 2, 3)
    bool fft_fuse;        // x and y passes of a direction in one persistent launch (k_fft_plane)
    int fft_lag;          // planes between the producer and the consumer pass of that launch
    unsigned *fft_sync;   // [0] error flag, then per direction: ticket + per-plane counters
    cufftHandle r2c, c2r;
    bool have_fft;
    // diagnostic backend 2 (pm_plan_set_fft_backend): the reference's own transform precision, float64
    // D2Z/Z2D through cuFFT on buffers allocated on first use (pm_poisson.cu, pm_k_poisson_f64)
    bool fft_f64, f64_ready;
    cufftHandle d2z, z2d;
    double *f64_mesh;     // [nc^3]
    double *f64_spec;     // [nc^2 * (nc/2+1)] complex
    // Poisson options (pm_plan_set_poisson_options; north_star (2)): both 0 = the reference's scheme
    int deconv;           // phi_k /= W(k)^deconv, W = CIC window
    int kgrad;            // 1: accelerations from -i k phi_k (three force meshes), interpolated by the gather
    float *dec_tab;       // [nc] per-axis 1/W_i^deconv
    float *k_tab;         // [nc] wavenumbers 2*pi*fftfreq (0 at Nyquist)
    float2 *spec2;        // second half spectrum (one force component at a time)
    float *fmesh;         // [3][nc^3]: 2 * (-dphi/dx_d), the "s" of pm_push

    // resident particle state (pm_particles_load / pm_step_resident / pm_particles_store), two
    // buffer sets; set rcur holds the particles in the cell order of the previous step's sort.
    float *rpos[2], *rvel[2];
    uint32_t *rid[2];     // original index of the particle in each slot
    int rcur;
    int64_t rnp;          // live particles in set rcur
    int64_t rstride;      // distance between the x, y, z rows of the resident sets (= np_cap)
    bool rkeys_valid;     // p->keys already holds the keys of set rcur (written by the last gather)
    bool rsort_done;      // keys_sorted / order_sorted / row_start describe set rcur as it is now (pm_api.cu)
    cudaStream_t s_main, s_up, s_down;
    cudaEvent_t ev_a, ev_b, ev_c;
    cudaEvent_t ev_chunk[PM_HOST_CHUNKS];   // pm_step_host: un-permuted particle ranges ready for download
    cudaEvent_t ev_upchunk[PM_HOST_CHUNKS]; // pm_step_host: velocity ranges uploaded
    cudaEvent_t ev_x;                       // pm_step_host (split route): x coordinates uploaded
    bool sort_rows_only;      // pm_step_host (split route): the next full sort orders by mesh ROW only (see pm_k_sort)

    // slab-mode scratch (nranks > 1, or a 1-rank slab plan used to test the slab kernels)
    float2 *tbuf[2];        // all-to-all staging: [nranks][nzl][nyl][nc/2] (+ side [nranks][nzl][nyl])
    uint32_t *leave_cnt;    // [nranks] particles leaving to each rank (written by the gather kernel)
    uint32_t *leave_slot;   // [nranks][leave_cap] their storage slots
    uint32_t *leave_sorted; // [leave_cap] one destination's slots in ascending order
    int64_t leave_cap;
    float *mig_send, *mig_recv;  // [nranks*leave_cap][7] packed (x,y,z,vx,vy,vz,id) records
    int64_t rtotal;         // entries of the current buffer set incl. dead (left) and arrived ones

    // peer-memory transposes of the distributed FFT (pm_slab_peer_*): the pack kernel stores its
    // blocks straight into the peers' z-pass arrays and the unpack kernel loads from them, over
    // NVLink (CUDA IPC mappings between the one-process-per-GPU ranks, plain pointers between the
    // plans of a single-process rank loop); flag words replace the collective as the barrier
    uint32_t *peer_flags;                   // [PM_PEER_SLOTS + 1][PM_PEER_MAX] in ws; last row [0] = timeouts
    float2 *peer_recv[PM_PEER_MAX];         // rank s's tbuf[1] as seen from this device
    uint32_t *peer_flag_of[PM_PEER_MAX];    // rank s's peer_flags
    void *peer_ipc[PM_PEER_MAX];            // cudaIpcOpenMemHandle mappings to close
    uint32_t peer_epoch_sig[PM_PEER_SLOTS], peer_epoch_wait[PM_PEER_SLOTS];
    int peers_set;                          // how many of the nranks entries are filled in
    float *peer_mesh2[PM_PEER_MAX];         // rank s's phi buffer (ghost planes pushed through peer memory)
    int ghosts_set;
    uint32_t *mig_matrix;                   // [PM_PEER_MAX][PM_MIG_ROW] in ws: every rank's leave counts etc. (pm_migrate.cu)
    uint32_t *mig_matrix_h;                 // pinned host copy
    uint32_t *peer_mig_matrix[PM_PEER_MAX]; // rank s's matrix / receive buffer as seen from this device
    float *peer_mig_recv[PM_PEER_MAX];
    int aux_set;

    // tile deposit (pm_deposit_tiles.cuh): scratch slots, queue and counters of the heavy tiles
    bool deposit_tiles;     // false: k_deposit_rows (PM_DEPOSIT=rows, meshes the tile kernel does not take)
    unsigned long long *dep_scratch;
    uint32_t *dep_ctl, *dep_slot_tile;   // ctl: 16 words, then slot_done[S] (zeroed with ctl every step), slot_tile[S], slot_items[S]
    void *dep_items;
    // work list of the tiled gather (pm_gather_ws.cuh, k_gather_items): heavy items from the front, light from the back
    void *gat_items;
    uint32_t *gat_ctl;      // [0] heavy items, [1] light items, [2] overflow flag
    int gat_cap;
    bool peer_dma;          // transport "peer": the transposes by cudaMemcpy3DAsync (copy engines) instead of copy kernels
    bool gather_items;      // PM_GATHER_ITEMS=0: fixed (row block, z chunk) grid as before
    void *diag;             // 64 bytes of device scratch for diagnostics (pm_plan_block_stats)

    // CUDA-graph replay of the resident step (pm_step_resident; pm_api.cu)
    PmStepParams *step_params_d;   // device copy of the per-step scalars (always allocated)
    PmStepParams *graph_params;    // == step_params_d while a step is being CAPTURED (kernels then read it), else nullptr
    bool use_graph;                // PM_GRAPH=0 turns replay off
    void *graph_exec[2];           // cudaGraphExec_t per buffer-set parity
    void *graph_node[2];           // the parameter kernel node of each (a handle into graph_tpl, which must outlive it)
    void *graph_tpl[2];            // cudaGraph_t the executable was instantiated from
    float *graph_rho[2];           // what each was captured for
    double graph_mass[2], graph_omega[2];
    int64_t graph_np[2];
    cudaStream_t graph_stream[2];
    int graph_replays;             // statistics

    // optional per-stage timing of pm_step (pm_plan_profile_begin/read)
    cudaEvent_t *prof_ev;   // prof_cap * (PM_NUM_STAGES + 1) events
    int prof_cap, prof_n;
};

// Record the boundary event `k` (0 = before the first stage) of the step being profiled.
static inline void pm_prof_mark(pm_plan *p, int k, cudaStream_t st)
{
    if (p->prof_ev && p->prof_n < p->prof_cap)
        cudaEventRecord(p->prof_ev[(size_t)p->prof_n * (PM_NUM_STAGES + 1) + k], st);
}

static inline cudaStream_t pm_cu(pm_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

// pm_sort.cu
#ifndef PM_SORT_TILE
#define PM_SORT_TILE 2048
#endif
#define PM_SORT_MAX_GRID 1024   /* sort_hist holds PM_SORT_MAX_GRID * 512 words */
int pm_k_sort_stats(pm_plan *p, int64_t *entries, int64_t *movers, int *mode);
int64_t pm_sort_tiles(int64_t np);
int64_t pm_sort_mover_capacity(int64_t np);
int pm_k_sort(pm_plan *p, int64_t np, int64_t n_old, cudaStream_t st);
int pm_k_sort_u32(pm_plan *p, const uint32_t *in, uint32_t *out, int64_t count, cudaStream_t st);

// pm_particles.cu
int pm_k_cell_keys(pm_plan *p, const float *pos, int64_t np, int64_t stride, uint32_t *keys,
                   uint32_t *order, cudaStream_t st, bool rows_only = false);
int pm_k_row_offsets(pm_plan *p, int64_t np, cudaStream_t st);
int pm_k_deposit(pm_plan *p, const float *pos, int64_t stride, double mass, float *rho,
                 cudaStream_t st);
int pm_k_gather_kick_drift(pm_plan *p, float *pos, float *vel, int64_t np, const float *phi,
                           double a_val, double f_a1, double da, float *acc, cudaStream_t st);
int pm_k_gather_kick_drift_resident(pm_plan *p, const float *phi, double a_val, double f_a1,
                                    double da, cudaStream_t st);
int pm_k_unpermute(pm_plan *p, float *pos_out, float *vel_out, cudaStream_t st);
int pm_k_poisson_tables(pm_plan *p);
int pm_gather_item_capacity(int nc, int64_t np);
int pm_k_reorder_rows(pm_plan *p, const float *in, const uint32_t *order, int64_t np, float *out, cudaStream_t st);
bool pm_unpermute_aos_ok(const pm_plan *p, const float *pos_out, const float *vel_out);
int pm_k_unpermute_scatter_aos(pm_plan *p, cudaStream_t st);
int pm_k_aos_rows_range(pm_plan *p, int64_t i0, int64_t i1, float *pos_out, float *vel_out, cudaStream_t st);
void pm_gather_step_scalars(double a_val, double f_a1, double da, PmStepParams *out);
bool pm_gather_sums_ok(const pm_plan *p);
int pm_k_gather_sums(pm_plan *p, const float *phi, cudaStream_t st);
int pm_k_push_rows(pm_plan *p, const float *pos_in, const float *vel_in, int64_t i0, int64_t i1, double a_val, double f_a1,
                   double da, float *pos_out, float *vel_out, cudaStream_t st);
bool pm_gather_graphable(const pm_plan *p);
int pm_k_block_stats(pm_plan *p, int rows_per_block, int cap, int64_t *out4, cudaStream_t st);
void pm_gather_tile_shape(int *rows_per_block, int *cap);
// slab mode (pm_slab.cu / pm_particles.cu)
int pm_k_deposit_slab(pm_plan *p, const float *pos, double mass, float *rho, cudaStream_t st);
int pm_deposit_segments(int nc);
int pm_k_gather_kick_drift_slab(pm_plan *p, const float *phi, double a_val, double f_a1, double da,
                                cudaStream_t st);
int pm_k_iota(uint32_t *out, int64_t n, cudaStream_t st);

// pm_fft.cu
bool pm_fft_supported(int nc);
int pm_peer_timeout_init();
int pm_k_fft_tables(pm_plan *p);
int pm_k_sin2rev_install(pm_plan *p, const float *sin2_h);
int pm_k_poisson_own(pm_plan *p, const float *rho, double a, double omega_m0, float *phi,
                     cudaStream_t st);
// slab pieces of the same transform, cut into C chunks of kx columns so that the all-to-all of
// one chunk overlaps the passes of its neighbours (pm_slab.cu sequences them)
int pm_k_fft_slab_rows_fwd(pm_plan *p, const float *rho, cudaStream_t st);
int pm_k_fft_slab_y_fwd_pack(pm_plan *p, int c, int C, float2 *send_main_c, float2 *send_side,
                             cudaStream_t st);
int pm_k_fft_slab_z_chunk(pm_plan *p, int c, int C, float2 *main_t_c, float2 *side_t, double a,
                          double omega_m0, cudaStream_t st);
int pm_k_fft_slab_unpack_y_inv(pm_plan *p, int c, int C, const float2 *back_main_c,
                               const float2 *back_side, cudaStream_t st);
int pm_k_fft_slab_rows_inv(pm_plan *p, float *phi, cudaStream_t st);
// peer-memory variant: y pass / push into the peers' z-pass arrays / pull from them / y inverse
int pm_k_fft_slab_y_fwd(pm_plan *p, int c, int C, cudaStream_t st);
int pm_k_fft_slab_push(pm_plan *p, int c, int C, cudaStream_t st);
int pm_k_fft_slab_pull(pm_plan *p, int c, int C, cudaStream_t st);
int pm_k_fft_slab_y_inv(pm_plan *p, int c, int C, cudaStream_t st);
// ... and with the transposes fused into the y passes themselves
int pm_k_fft_slab_y_fwd_push(pm_plan *p, int c, int C, cudaStream_t st);
int pm_k_fft_slab_y_inv_pull(pm_plan *p, int c, int C, cudaStream_t st);
int pm_k_peer_signal(pm_plan *p, int slot, uint32_t epoch, cudaStream_t st);
int pm_k_peer_wait(pm_plan *p, int slot, uint32_t epoch, cudaStream_t st);
// ghost planes through peer memory: plane copy into a neighbour + flag to that neighbour only
int pm_k_peer_put(pm_plan *p, float *dst, const float *src, size_t nfloat, cudaStream_t st);
int pm_k_peer_signal_to(pm_plan *p, int slot, int dst_rank, uint32_t epoch, cudaStream_t st);
int pm_k_peer_wait_from(pm_plan *p, int slot, int src_rank, uint32_t epoch, cudaStream_t st);
int pm_fft_cols_per_tile(int nc);
int pm_k_power_spectrum(pm_plan *p, const float *rho, int nbins, double *psum, double *pcnt,
                        cudaStream_t st);

// pm_poisson.cu
int pm_k_sin2_table(pm_plan *p);
int pm_k_rho_mean(pm_plan *p, const float *rho, size_t n, double total_cells, cudaStream_t st);
int pm_k_fourier_grid(pm_plan *p, float *fgrid, cudaStream_t st);
int pm_k_poisson(pm_plan *p, const float *rho, double a, double omega_m0, float *phi,
                 cudaStream_t st);

// slab helpers (pm_particles.cu)
int pm_k_ghost_add(pm_plan *p, float *plane, const float *ghost, cudaStream_t st);
int pm_k_migrate_pack(pm_plan *p, int dest, int64_t count, int64_t rec_offset, cudaStream_t st);
int pm_k_migrate_unpack(pm_plan *p, int64_t n_arrive, cudaStream_t st);
int pm_k_export(pm_plan *p, float *pos_out, float *vel_out, uint32_t *id_out, uint32_t *live_out,
                cudaStream_t st);
