// pm_api.cu -- the extern "C" surface declared in include/pmstep.h.
#include <math.h>
#include <stdio.h>
#include <new>
#include <stdlib.h>
#include <string.h>

#include "pm_internal.cuh"

unsigned long long g_pm_launches = 0;
int g_pm_last_cufft = 0;

namespace {

const size_t kAlign = 512;
inline size_t align_up(size_t v) { return (v + kAlign - 1) / kAlign * kAlign; }

int key_bits_for(int nc)
{
    const uint64_t m = (uint64_t)nc * nc * nc;
    int b = 1;
    while (((uint64_t)1 << b) < m) ++b;
    return b;
}

// Sizes of every workspace segment, in the order they are carved.
struct Layout {
    size_t keys, iota, keys_sorted, order_sorted, cub, row_start, mesh, mesh2, spec, fft, sin2, tw,
        rpos, rvel, rid, tbuf, leave_cnt, leave_slot, leave_sorted, mig, peer_flags, inc_a, inc_b, inc_tile, fft_sync,
        gat, total;
    int gat_cap;
    int64_t leave_cap, inc_bcap;
};

struct Geometry {
    int nc, rank, nranks;
    bool slab;   // slab-mode buffers (ghost planes, all-to-all staging, migration lists)
};

bool config_ok(int nc, int64_t np)
{
    if (nc < 4 || nc > 2048) return false;
    if (np < 0 || np > (int64_t)0xfffffff0u) return false;
    return true;
}

int make_fft_plans(int nc, cufftHandle *r2c, cufftHandle *c2r, size_t *work)
{
    size_t w1 = 0, w2 = 0;
    PM_CUFFT(cufftCreate(r2c));
    PM_CUFFT(cufftCreate(c2r));
    PM_CUFFT(cufftSetAutoAllocation(*r2c, 0));
    PM_CUFFT(cufftSetAutoAllocation(*c2r, 0));
    PM_CUFFT(cufftMakePlan3d(*r2c, nc, nc, nc, CUFFT_R2C, &w1));
    PM_CUFFT(cufftMakePlan3d(*c2r, nc, nc, nc, CUFFT_C2R, &w2));
    *work = w1 > w2 ? w1 : w2;
    return PM_OK;
}

int compute_layout(const Geometry &g, int64_t np, size_t fft_work, Layout *L)
{
    const int nc = g.nc, nzl = nc / g.nranks;
    const size_t plane = (size_t)nc * nc;
    const size_t npad = (size_t)((np + 3) / 4 * 4);
    L->keys = align_up(npad * 4);
    L->iota = align_up(npad * 4);
    L->keys_sorted = align_up(npad * 4);
    L->order_sorted = align_up(npad * 4);
    L->cub = align_up(npad * 8) + align_up(256) + align_up(((size_t)PM_SORT_MAX_GRID * 2 + 1) * 512 * 4);   // sort_tmp, sort_ctl, sort_hist (+ prefixes, totals)
    L->row_start = align_up(((size_t)nzl * nc * pm_deposit_segments(nc) + 1) * 4);
    L->mesh = align_up(plane * (nzl + (g.slab ? 1 : 0)) * 4);    // rho (+ ghost plane)
    L->mesh2 = align_up(plane * (nzl + (g.slab ? 3 : 0)) * 4);   // phi (+ 1 + 2 ghost planes)
    L->spec = align_up((size_t)nzl * nc * (nc / 2 + 1) * sizeof(float2));
    L->fft = align_up(fft_work);
    L->sin2 = align_up((size_t)nc * 4);
    L->tw = align_up((size_t)nc * 8);
    L->rpos = align_up(3 * npad * 4);
    L->rvel = align_up(3 * npad * 4);
    L->rid = align_up(npad * 4);
    L->tbuf = g.slab ? L->spec : 0;
    // migration: room for 1/32 of the capacity (at least 65536) to leave towards each rank per step
    L->leave_cap = g.slab ? (int64_t)((npad / 32 > 65536) ? npad / 32 : 65536) : 0;
    L->leave_cnt = g.slab ? align_up((size_t)g.nranks * 4) : 0;
    L->leave_slot = g.slab ? align_up((size_t)g.nranks * L->leave_cap * 4) : 0;
    L->leave_sorted = g.slab ? align_up((size_t)L->leave_cap * 4) : 0;
    L->mig = g.slab ? align_up((size_t)g.nranks * L->leave_cap * 7 * 4) : 0;
    L->peer_flags = g.slab ? align_up((size_t)(PM_PEER_SLOTS + 1) * PM_PEER_MAX * 4) + align_up((size_t)PM_PEER_MAX * PM_MIG_ROW * 4) : 0;
    // incremental sort: stayers (np), movers in and out (capacity each), tile tables
    L->inc_bcap = pm_sort_mover_capacity((int64_t)npad);
    L->inc_a = align_up(npad * 8);
    L->inc_b = align_up((size_t)L->inc_bcap * 8);
    L->inc_tile = align_up(((size_t)pm_sort_tiles((int64_t)npad) + 2) * 4);
    L->fft_sync = align_up((size_t)(2 * (nc + 1) + 1) * 4);
    // gather work list: 2 per base chunk + the pieces of split chunks (pm_particles.cu, pm_gather_item_threshold)
    L->gat_cap = pm_gather_item_capacity(nc, (int64_t)npad);
    L->gat = align_up((size_t)L->gat_cap * 16) + kAlign;
    L->total = L->gat + align_up((size_t)PM_DEP_MAX_SLOTS * 8192 * 8) /* dep_scratch */ + align_up(64 + 3 * PM_DEP_MAX_SLOTS * 4) /* dep ctl + slot tables */ +
               align_up((size_t)16384 * 16) /* dep_items */ + kAlign /* diag */ + align_up(1024 * sizeof(double) + 64) /* mean */ + L->fft_sync + L->keys + L->iota + L->keys_sorted + L->order_sorted + L->cub + L->row_start +
               L->mesh + L->mesh2 + L->spec + L->fft + 2 * L->sin2 + L->tw +
               2 * (L->rpos + L->rvel + L->rid) + 2 * L->tbuf + L->leave_cnt + L->leave_slot +
               L->leave_sorted + 2 * L->mig + L->peer_flags + L->inc_a + 2 * L->inc_b + 2 * L->inc_tile;
    return PM_OK;
}

struct DeviceGuard {
    int prev = -1;
    bool active = false;
    int enter(int dev)
    {
        if (cudaGetDevice(&prev) != cudaSuccess) return PM_ERR_NO_DEVICE;
        if (dev >= 0 && dev != prev) {
            cudaError_t e = cudaSetDevice(dev);
            if (e != cudaSuccess) return (int)e;
            active = true;
        }
        return PM_OK;
    }
    ~DeviceGuard()
    {
        if (active) cudaSetDevice(prev);
    }
};

}  // namespace

extern "C" {

const char *pm_version(void) { return "pmstep 0.2 (sm_100a; hand-written sort, deposit, 5-pass real FFT with fused Green's function, gather; cuFFT only for non-power-of-two meshes and the diagnostic backends)"; }

const char *pm_error_string(int code)
{
    switch (code) {
        case PM_OK: return "ok";
        case PM_ERR_INVALID: return "invalid argument";
        case PM_ERR_UNSUPPORTED: return "unsupported configuration";
        case PM_ERR_NOMEM: return "workspace allocation failed";
        case PM_ERR_CUFFT: return "cuFFT error (see pm_last_cufft_status)";
        case PM_ERR_NO_DEVICE: return "no usable CUDA device";
        default: return code > 0 ? cudaGetErrorString((cudaError_t)code) : "unknown error";
    }
}

int pm_last_cufft_status(void) { return g_pm_last_cufft; }
uint64_t pm_launch_count(void) { return g_pm_launches; }

size_t pm_plan_workspace_bytes(int n_cells, int64_t np_capacity)
{
    if (!config_ok(n_cells, np_capacity)) return 0;
    cufftHandle a, b;
    size_t work = 0;
    if (make_fft_plans(n_cells, &a, &b, &work) != PM_OK) return 0;
    cufftDestroy(a);
    cufftDestroy(b);
    Layout L;
    Geometry g{n_cells, 0, 1, false};
    compute_layout(g, np_capacity, work, &L);
    return L.total;
}

static int plan_create(pm_plan **out, const Geometry &g, int64_t np_capacity, int device)
{
    if (!out) return PM_ERR_INVALID;
    *out = nullptr;
    const int n_cells = g.nc;
    if (!config_ok(n_cells, np_capacity)) return PM_ERR_UNSUPPORTED;
    if (g.nranks < 1 || g.rank < 0 || g.rank >= g.nranks || n_cells % g.nranks != 0)
        return PM_ERR_INVALID;
    if (g.slab && (!pm_fft_supported(n_cells) || (n_cells / g.nranks) % 16 != 0))
        return PM_ERR_UNSUPPORTED;   // slab mode runs the hand-written FFT on 16-column tiles
    // 32-bit cell keys and mesh offsets: a slab plan needs (nzl+3)*Nc^2 < 2^32, a whole-mesh plan Nc^3
    if (g.slab ? ((uint64_t)(n_cells / g.nranks + 3) * n_cells * n_cells >= (1ull << 32))
               : ((uint64_t)n_cells * n_cells * n_cells >= (1ull << 32)))
        return PM_ERR_UNSUPPORTED;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return PM_ERR_NO_DEVICE;
    }
    DeviceGuard guard;
    int rc = guard.enter(device);
    if (rc != PM_OK) return rc;

    pm_plan *p = new (std::nothrow) pm_plan();
    if (!p) return PM_ERR_NOMEM;
    memset(p, 0, sizeof(*p));
    p->nc = n_cells;
    p->np_cap = np_capacity;
    p->rank = g.rank;
    p->nranks = g.nranks;
    p->nzl = n_cells / g.nranks;
    p->z0 = g.rank * p->nzl;
    p->slab = g.slab;
    p->dep_nseg = pm_deposit_segments(n_cells);
    p->key_bits = g.slab ? 32 : key_bits_for(n_cells);
    p->rstride = (np_capacity + 3) / 4 * 4;
    cudaGetDevice(&p->device);
    cudaDeviceGetAttribute(&p->sm_count, cudaDevAttrMultiProcessorCount, p->device);

    size_t work = 0;
    if (!g.slab) {
        rc = make_fft_plans(n_cells, &p->r2c, &p->c2r, &work);
        if (rc != PM_OK) {
            pm_plan_destroy(p);
            return rc;
        }
        p->have_fft = true;
    }
    Layout L;
    compute_layout(g, np_capacity, work, &L);
    p->ws_bytes = L.total;
    if (cudaMalloc((void **)&p->ws, L.total) != cudaSuccess) {
        cudaGetLastError();
        pm_plan_destroy(p);
        return PM_ERR_NOMEM;
    }
    char *c = p->ws;
    p->keys = (uint32_t *)c;          c += L.keys;
    p->iota = (uint32_t *)c;          c += L.iota;
    p->keys_sorted = (uint32_t *)c;   c += L.keys_sorted;
    p->order_sorted = (uint32_t *)c;  c += L.order_sorted;
    p->sort_tmp = (uint64_t *)c;
    p->sort_ctl = c + align_up((size_t)((np_capacity + 3) / 4 * 4) * 8);
    p->sort_hist = (uint32_t *)((char *)p->sort_ctl + align_up(256));
    c += L.cub;
    p->row_start = (uint32_t *)c;     c += L.row_start;
    p->mesh = (float *)c;             c += L.mesh;
    p->mesh2 = (float *)c;            c += L.mesh2;
    p->spec = (float2 *)c;            c += L.spec;
    p->fft_work = c;                  c += L.fft;
    p->fft_work_bytes = L.fft;
    p->sin2 = (float *)c;             c += L.sin2;
    p->sin2rev = (float *)c;          c += L.sin2;
    p->tw = (float2 *)c;              c += L.tw;
    for (int k = 0; k < 2; ++k) {
        p->rpos[k] = (float *)c;      c += L.rpos;
        p->rvel[k] = (float *)c;      c += L.rvel;
        p->rid[k] = (uint32_t *)c;    c += L.rid;
    }
    if (g.slab) {
        p->tbuf[0] = (float2 *)c;     c += L.tbuf;
        p->tbuf[1] = (float2 *)c;     c += L.tbuf;
        p->leave_cnt = (uint32_t *)c; c += L.leave_cnt;
        p->leave_slot = (uint32_t *)c; c += L.leave_slot;
        p->leave_sorted = (uint32_t *)c; c += L.leave_sorted;
        p->mig_send = (float *)c;     c += L.mig;
        p->mig_recv = (float *)c;     c += L.mig;
        p->leave_cap = L.leave_cap;
        p->peer_flags = (uint32_t *)c;
        p->mig_matrix = (uint32_t *)(c + align_up((size_t)(PM_PEER_SLOTS + 1) * PM_PEER_MAX * 4));
        c += L.peer_flags;
    }
    p->inc_a = (uint64_t *)c;         c += L.inc_a;
    p->inc_b = (uint64_t *)c;         c += L.inc_b;
    p->inc_bs = (uint64_t *)c;        c += L.inc_b;
    p->inc_tile = (uint32_t *)c;      c += L.inc_tile;
    p->inc_split = (uint32_t *)c;     c += L.inc_tile;
    p->inc_bcap = L.inc_bcap;
    p->fft_sync = (unsigned *)c;      c += L.fft_sync;
    p->diag = c;                      p->step_params_d = (PmStepParams *)(c + 256);  c += kAlign;
    p->dep_scratch = (unsigned long long *)c;  c += align_up((size_t)PM_DEP_MAX_SLOTS * 8192 * 8);
    p->dep_ctl = (uint32_t *)c;       p->dep_slot_tile = (uint32_t *)(c + 64 + PM_DEP_MAX_SLOTS * 4);  c += align_up(64 + 3 * PM_DEP_MAX_SLOTS * 4);
    p->dep_items = c;                 c += align_up((size_t)16384 * 16);
    p->mean_ws = (double *)c;         p->rho_mean_d = (float *)(c + 1024 * sizeof(double));
    c += align_up(1024 * sizeof(double) + 64);
    p->gat_items = c;                 p->gat_ctl = (uint32_t *)(c + align_up((size_t)L.gat_cap * 16));  c += L.gat;
    p->gat_cap = L.gat_cap;
    p->rho_mean_hint = NAN;
    {
        const char *fu = getenv("PM_FFT_FUSE");   // "0": separate row and y launches (A/B checks)
        const char *lg = getenv("PM_FFT_LAG");
        const char *v2 = getenv("PM_FFT_V2");     // "0": the three-stage radix-8 kernels
        p->fft_v2 = !(v2 && strcmp(v2, "0") == 0);
        // measured at 512^3 (profiles/r01_notes.md): separate row and y launches 0.46 ms per
        // direction, the persistent plane launch 0.48 ms -> off unless PM_FFT_FUSE=1
        p->fft_fuse = (fu && strcmp(fu, "1") == 0);
        const char *zm = getenv("PM_FFT_ZMIX");   // "0": two-stage kernel for the fused z pass too
        p->fft_zmix = !(zm && strcmp(zm, "0") == 0);
        const char *v3 = getenv("PM_FFT_V3");     // ring depth of the pipelined y passes: 0 (off), 2, 3
        p->fft_v3 = v3 ? atoi(v3) : 2;
        if (p->fft_v3 != 2 && p->fft_v3 != 3) p->fft_v3 = 0;
        p->fft_lag = lg ? atoi(lg) : 12;
        if (p->fft_lag < 1) p->fft_lag = 1;
        if (p->fft_lag > n_cells) p->fft_lag = n_cells;
    }
    {
        const char *gt = getenv("PM_GATHER_TILED");   // "0": one thread per particle, scattered phi loads
        p->gather_tiled = !(gt && strcmp(gt, "0") == 0);
        const char *gw = getenv("PM_GATHER_WS");      // "0": the barrier-per-step tiled kernel (A/B checks)
        p->gather_ws = !(gw && strcmp(gw, "0") == 0);
    }
    {
        const char *pd = getenv("PM_PEER_DMA");       // "1": copy engines for the "peer" transport's transposes
        p->peer_dma = pd && strcmp(pd, "0") != 0;
        const char *gi = getenv("PM_GATHER_ITEMS");   // "0": the gather's fixed grid, no work list
        p->gather_items = !(gi && strcmp(gi, "0") == 0);
        const char *gr = getenv("PM_GRAPH");     // "0": never replay the resident step as a CUDA graph
        p->use_graph = !(gr && strcmp(gr, "0") == 0);
    }
    {
        const char *dp = getenv("PM_DEPOSIT");   // "rows": one warp per output row (k_deposit_rows; A/B checks)
        p->deposit_tiles = !(dp && strcmp(dp, "rows") == 0);
    }
    {
        const char *sm = getenv("PM_SORT");   // "full" forces the radix sort of every entry (A/B checks)
        p->sort_mode = (sm && strcmp(sm, "full") == 0) ? PM_SORT_FULL : PM_SORT_AUTO;
    }

    if (p->have_fft && (cufftSetWorkArea(p->r2c, p->fft_work) != CUFFT_SUCCESS ||
                        cufftSetWorkArea(p->c2r, p->fft_work) != CUFFT_SUCCESS)) {
        pm_plan_destroy(p);
        return PM_ERR_CUFFT;
    }
    rc = pm_k_sin2_table(p);
    if (rc == PM_OK && g.slab) rc = pm_peer_timeout_init();     // PM_PEER_TIMEOUT_MS (per device: a __device__ variable)
    {
        const char *be = getenv("PM_FFT_BACKEND");  // "cufft" forces the library path (A/B checks)
        p->own_fft = pm_fft_supported(n_cells) && (g.slab || !(be && (strcmp(be, "cufft") == 0 || strcmp(be, "f64") == 0)));
        p->fft_f64 = !g.slab && be && strcmp(be, "f64") == 0 && n_cells <= 256;   // diagnostic float64 transforms (pm_poisson.cu)
    }
    if (rc == PM_OK && pm_fft_supported(n_cells)) rc = pm_k_fft_tables(p);
    if (rc == PM_OK) rc = (int)cudaMemset(p->fft_sync, 0, sizeof(unsigned));
    if (rc == PM_OK) rc = (int)cudaMemset(p->sort_ctl, 0, 256);
    if (rc == PM_OK) rc = (int)cudaMemset(p->dep_scratch, 0, (size_t)PM_DEP_MAX_SLOTS * 8192 * 8);   // all-zero between steps (k_deposit_items zeroes what it converts)
    if (rc == PM_OK && p->peer_flags)
        rc = (int)cudaMemset(p->peer_flags, 0, (size_t)(PM_PEER_SLOTS + 1) * PM_PEER_MAX * 4);
    if (rc == PM_OK && p->mig_matrix) rc = (int)cudaMemset(p->mig_matrix, 0, (size_t)PM_PEER_MAX * PM_MIG_ROW * 4);
    if (rc == PM_OK && g.slab && cudaHostAlloc((void **)&p->mig_matrix_h, (size_t)PM_PEER_MAX * PM_MIG_ROW * 4, cudaHostAllocDefault) != cudaSuccess) {
        cudaGetLastError();
        p->mig_matrix_h = nullptr;
        rc = PM_ERR_NOMEM;
    }
    if (rc == PM_OK) rc = pm_k_iota(p->iota, np_capacity, 0);
    if (rc == PM_OK) rc = (int)cudaStreamSynchronize(0);
    if (rc == PM_OK) rc = (int)cudaStreamCreateWithFlags(&p->s_main, cudaStreamNonBlocking);
    if (rc == PM_OK) rc = (int)cudaStreamCreateWithFlags(&p->s_up, cudaStreamNonBlocking);
    if (rc == PM_OK) rc = (int)cudaStreamCreateWithFlags(&p->s_down, cudaStreamNonBlocking);
    if (rc == PM_OK) rc = (int)cudaEventCreateWithFlags(&p->ev_a, cudaEventDisableTiming);
    if (rc == PM_OK) rc = (int)cudaEventCreateWithFlags(&p->ev_b, cudaEventDisableTiming);
    if (rc == PM_OK) rc = (int)cudaEventCreateWithFlags(&p->ev_c, cudaEventDisableTiming);
    for (int k = 0; k < PM_HOST_CHUNKS && rc == PM_OK; ++k) rc = (int)cudaEventCreateWithFlags(&p->ev_chunk[k], cudaEventDisableTiming);
    for (int k = 0; k < PM_HOST_CHUNKS && rc == PM_OK; ++k) rc = (int)cudaEventCreateWithFlags(&p->ev_upchunk[k], cudaEventDisableTiming);
    if (rc == PM_OK) rc = (int)cudaEventCreateWithFlags(&p->ev_x, cudaEventDisableTiming);
    if (rc != PM_OK) {
        pm_plan_destroy(p);
        return rc;
    }
    *out = p;
    return PM_OK;
}

int pm_plan_create(pm_plan **out, int n_cells, int64_t np_capacity, int device)
{
    Geometry g{n_cells, 0, 1, false};
    return plan_create(out, g, np_capacity, device);
}

int pm_plan_create_slab(pm_plan **out, int n_cells, int64_t np_capacity, int device, int rank,
                        int nranks)
{
    Geometry g{n_cells, rank, nranks, true};
    return plan_create(out, g, np_capacity, device);
}

int pm_plan_destroy(pm_plan *p)
{
    if (!p) return PM_OK;
    DeviceGuard guard;
    guard.enter(p->device);
    if (p->have_fft) {
        cufftDestroy(p->r2c);
        cufftDestroy(p->c2r);
    }
    if (p->f64_ready) {
        cufftDestroy(p->d2z);
        cufftDestroy(p->z2d);
    }
    if (p->dec_tab) cudaFree(p->dec_tab);
    if (p->spec2) cudaFree(p->spec2);
    if (p->fmesh) cudaFree(p->fmesh);
    if (p->f64_mesh) cudaFree(p->f64_mesh);
    if (p->f64_spec) cudaFree(p->f64_spec);
    if (p->ev_a) cudaEventDestroy(p->ev_a);
    if (p->ev_b) cudaEventDestroy(p->ev_b);
    if (p->ev_c) cudaEventDestroy(p->ev_c);
    for (int k = 0; k < PM_HOST_CHUNKS; ++k)
        if (p->ev_chunk[k]) cudaEventDestroy(p->ev_chunk[k]);
    for (int k = 0; k < PM_HOST_CHUNKS; ++k)
        if (p->ev_upchunk[k]) cudaEventDestroy(p->ev_upchunk[k]);
    if (p->ev_x) cudaEventDestroy(p->ev_x);
    if (p->s_main) cudaStreamDestroy(p->s_main);
    if (p->s_up) cudaStreamDestroy(p->s_up);
    if (p->s_down) cudaStreamDestroy(p->s_down);
    for (int s = 0; s < PM_PEER_MAX; ++s)
        if (p->peer_ipc[s]) cudaIpcCloseMemHandle(p->peer_ipc[s]);
    for (int k = 0; k < 2; ++k) {
        if (p->graph_exec[k]) cudaGraphExecDestroy((cudaGraphExec_t)p->graph_exec[k]);
        if (p->graph_tpl[k]) cudaGraphDestroy((cudaGraph_t)p->graph_tpl[k]);
    }
    if (p->mig_matrix_h) cudaFreeHost(p->mig_matrix_h);
    if (p->ws) cudaFree(p->ws);
    if (p->prof_ev) {
        for (int i = 0; i < p->prof_cap * (PM_NUM_STAGES + 1); ++i) cudaEventDestroy(p->prof_ev[i]);
        free(p->prof_ev);
    }
    delete p;
    return PM_OK;
}

int pm_plan_n_cells(const pm_plan *p) { return p ? p->nc : 0; }

int pm_plan_set_fft_backend(pm_plan *p, int backend)
{
    if (!p || backend < 0 || backend > 2) return PM_ERR_INVALID;
    if (backend == 0 && !pm_fft_supported(p->nc)) return PM_ERR_UNSUPPORTED;
    if (backend == 1 && !p->have_fft) return PM_ERR_UNSUPPORTED;   // slab plans carry no cuFFT plan
    if (backend == 2 && (p->slab || p->nc > 256)) return PM_ERR_UNSUPPORTED;   // diagnostic: small single-GPU meshes
    p->own_fft = (backend == 0);
    p->fft_f64 = (backend == 2);
    return PM_OK;
}

int pm_plan_set_sin2_table(pm_plan *p, const float *sin2_h)
{
    if (!p || !sin2_h) return PM_ERR_INVALID;
    DeviceGuard guard;
    const int rc = guard.enter(p->device);
    if (rc != PM_OK) return rc;
    PM_CUDA(cudaDeviceSynchronize());          // no solve may be reading the tables
    PM_CUDA(cudaMemcpy(p->sin2, sin2_h, sizeof(float) * p->nc, cudaMemcpyHostToDevice));
    return pm_k_sin2rev_install(p, sin2_h);
}

int pm_plan_set_poisson_options(pm_plan *p, int deconvolve, int kspace_gradient)
{
    if (!p || deconvolve < 0 || deconvolve > 2 || kspace_gradient < 0 || kspace_gradient > 1) return PM_ERR_INVALID;
    if ((deconvolve || kspace_gradient) && (p->slab || !p->have_fft)) return PM_ERR_UNSUPPORTED;
    DeviceGuard guard;
    int rc = guard.enter(p->device);
    if (rc != PM_OK) return rc;
    PM_CUDA(cudaDeviceSynchronize());
    const size_t cells = (size_t)p->nc * p->nc * p->nc;
    if ((deconvolve || kspace_gradient) && !p->dec_tab) {
        PM_CUDA(cudaMalloc(&p->dec_tab, sizeof(float) * 2 * p->nc));
        p->k_tab = p->dec_tab + p->nc;
    }
    if (kspace_gradient && !p->fmesh) {
        // allocated when the option is first asked for, never during a step
        if (cudaMalloc(&p->spec2, (size_t)p->nc * p->nc * (p->nc / 2 + 1) * sizeof(float2)) != cudaSuccess ||
            cudaMalloc(&p->fmesh, 3 * cells * sizeof(float)) != cudaSuccess) {
            cudaGetLastError();
            if (p->spec2) cudaFree(p->spec2);
            p->spec2 = nullptr; p->fmesh = nullptr;
            return PM_ERR_NOMEM;
        }
    }
    p->deconv = deconvolve;
    p->kgrad = kspace_gradient;
    if (deconvolve || kspace_gradient) rc = pm_k_poisson_tables(p);
    return rc;
}

int pm_plan_poisson_options(const pm_plan *p, int *deconvolve, int *kspace_gradient)
{
    if (!p || !deconvolve || !kspace_gradient) return PM_ERR_INVALID;
    *deconvolve = p->deconv;
    *kspace_gradient = p->kgrad;
    return PM_OK;
}

int pm_plan_gather_items(pm_plan *p, int64_t *heavy, int64_t *light, int *overflow)
{
    if (!p || !heavy || !light || !overflow) return PM_ERR_INVALID;
    DeviceGuard guard;
    const int rc = guard.enter(p->device);
    if (rc != PM_OK) return rc;
    uint32_t h[4] = {0, 0, 0, 0};
    PM_CUDA(cudaDeviceSynchronize());
    PM_CUDA(cudaMemcpy(h, p->gat_ctl, sizeof(h), cudaMemcpyDeviceToHost));
    *heavy = h[0]; *light = h[1]; *overflow = (int)h[2];
    return PM_OK;
}

int pm_plan_set_sort_mode(pm_plan *p, int mode)
{
    if (!p || (mode != PM_SORT_AUTO && mode != PM_SORT_FULL)) return PM_ERR_INVALID;
    p->sort_mode = mode;
    return PM_OK;
}

int pm_plan_sort_stats(const pm_plan *p, int64_t *entries, int64_t *movers, int *mode)
{
    if (!p) return PM_ERR_INVALID;
    DeviceGuard guard;
    int rc = guard.enter(p->device);
    if (rc != PM_OK) return rc;
    return pm_k_sort_stats(const_cast<pm_plan *>(p), entries, movers, mode);
}

int pm_plan_fft_backend(const pm_plan *p) { return p ? (p->own_fft ? 0 : (p->fft_f64 ? 2 : 1)) : PM_ERR_INVALID; }

int pm_plan_set_fft_variant(pm_plan *p, int two_stage)
{
    if (!p) return PM_ERR_INVALID;
    p->fft_v2 = (two_stage != 0);
    p->fft_zmix = (two_stage == 1 || two_stage >= 3);
    p->fft_v3 = two_stage == 3 ? 2 : two_stage == 4 ? 3 : 0;
    return PM_OK;
}

int pm_plan_set_gather_tiled(pm_plan *p, int tiled)
{
    if (!p) return PM_ERR_INVALID;
    p->gather_tiled = (tiled != 0);
    return PM_OK;
}

int pm_plan_gather_tile(const pm_plan *p, int *rows_per_block, int *cap)
{
    if (!p || !rows_per_block || !cap) return PM_ERR_INVALID;
    pm_gather_tile_shape(rows_per_block, cap);
    return PM_OK;
}

int pm_plan_block_stats(pm_plan *p, int rows_per_block, int cap, int64_t *out4, pm_stream_t stream)
{
    if (!p || !out4 || cap < 0) return PM_ERR_INVALID;
    DeviceGuard guard;
    int rc = guard.enter(p->device);
    if (rc != PM_OK) return rc;
    return pm_k_block_stats(p, rows_per_block, cap, out4, pm_cu(stream));
}

int pm_plan_set_fft_fuse(pm_plan *p, int fuse, int lag)
{
    if (!p || lag < 0 || lag > p->nc) return PM_ERR_INVALID;
    p->fft_fuse = (fuse != 0);
    if (lag > 0) p->fft_lag = lag;
    return PM_OK;
}

int pm_plan_fft_sync_errors(pm_plan *p)
{
    if (!p) return PM_ERR_INVALID;
    DeviceGuard guard;
    int rc = guard.enter(p->device);
    if (rc != PM_OK) return rc;
    unsigned v = 0;
    PM_CUDA(cudaMemcpy(&v, p->fft_sync, sizeof(unsigned), cudaMemcpyDeviceToHost));
    return v ? 1 : 0;
}
int64_t pm_plan_np_capacity(const pm_plan *p) { return p ? p->np_cap : 0; }

#define PM_ARGS(cond)                     \
    do {                                  \
        if (!(cond)) return PM_ERR_INVALID; \
    } while (0)
#define PM_TRY(expr)                      \
    do {                                  \
        int rc_ = (expr);                 \
        if (rc_ != PM_OK) return rc_;     \
    } while (0)

int pm_fourier_grid(pm_plan *p, float *fgrid_d, pm_stream_t stream)
{
    PM_ARGS(p && fgrid_d);
    DeviceGuard guard;
    PM_TRY(guard.enter(p->device));
    return pm_k_fourier_grid(p, fgrid_d, pm_cu(stream));
}

int pm_cell_keys(pm_plan *p, const float *pos_d, int64_t np, uint32_t *keys_d, pm_stream_t stream)
{
    PM_ARGS(p && (np == 0 || (pos_d && keys_d)) && np >= 0 && np <= p->np_cap);
    DeviceGuard guard;
    PM_TRY(guard.enter(p->device));
    return pm_k_cell_keys(p, pos_d, np, np, keys_d, nullptr, pm_cu(stream));
}

int pm_sort_by_cell(pm_plan *p, const float *pos_d, int64_t np, uint32_t *keys_sorted_d,
                    uint32_t *order_d, pm_stream_t stream)
{
    PM_ARGS(p && (np == 0 || pos_d) && np >= 0 && np <= p->np_cap);
    DeviceGuard guard;
    PM_TRY(guard.enter(p->device));
    cudaStream_t st = pm_cu(stream);
    p->rkeys_valid = false;
    p->rsorted_n = 0;
    p->rsort_done = false;
    PM_TRY(pm_k_cell_keys(p, pos_d, np, np, p->keys, nullptr, st));
    PM_TRY(pm_k_sort(p, np, 0, st));
    if (np == 0) return PM_OK;
    if (keys_sorted_d)
        PM_CUDA(cudaMemcpyAsync(keys_sorted_d, p->keys_sorted, (size_t)np * 4,
                                cudaMemcpyDeviceToDevice, st));
    if (order_d)
        PM_CUDA(cudaMemcpyAsync(order_d, p->order_sorted, (size_t)np * 4, cudaMemcpyDeviceToDevice, st));
    return PM_OK;
}

int pm_deposit_cic(pm_plan *p, const float *pos_d, int64_t np, double mass, float *rho_d,
                   pm_stream_t stream)
{
    PM_ARGS(p && (np == 0 || pos_d) && rho_d && np >= 0 && np <= p->np_cap);
    DeviceGuard guard;
    PM_TRY(guard.enter(p->device));
    cudaStream_t st = pm_cu(stream);
    p->rkeys_valid = false;
    p->rsorted_n = 0;
    p->rsort_done = false;
    PM_TRY(pm_k_cell_keys(p, pos_d, np, np, p->keys, nullptr, st));
    PM_TRY(pm_k_sort(p, np, 0, st));
    PM_TRY(pm_k_row_offsets(p, np, st));
    return pm_k_deposit(p, pos_d, np, mass, rho_d, st);
}

int pm_poisson(pm_plan *p, const float *rho_d, double a, double omega_m0, float *phi_d,
               pm_stream_t stream)
{
    PM_ARGS(p && rho_d && phi_d && a != 0.0);
    DeviceGuard guard;
    PM_TRY(guard.enter(p->device));
    p->rho_mean_hint = NAN;   // an arbitrary mesh: its mean is measured
    return pm_k_poisson(p, rho_d, a, omega_m0, phi_d, pm_cu(stream));
}

int pm_power_spectrum(pm_plan *p, const float *rho_d, int nbins, double *psum_d, double *pcnt_d,
                      pm_stream_t stream)
{
    PM_ARGS(p && rho_d && psum_d && pcnt_d && nbins >= 2 && nbins <= 4096);
    if (p->slab || !pm_fft_supported(p->nc)) return PM_ERR_UNSUPPORTED;
    DeviceGuard guard;
    PM_TRY(guard.enter(p->device));
    return pm_k_power_spectrum(p, rho_d, nbins, psum_d, pcnt_d, pm_cu(stream));
}

int pm_gather_kick_drift(pm_plan *p, float *pos_d, float *vel_d, int64_t np, const float *phi_d,
                         double a_val, double f_a1, double da, float *acc_d, pm_stream_t stream)
{
    PM_ARGS(p && (np == 0 || (pos_d && vel_d)) && phi_d && np >= 0);
    DeviceGuard guard;
    PM_TRY(guard.enter(p->device));
    return pm_k_gather_kick_drift(p, pos_d, vel_d, np, phi_d, a_val, f_a1, da, acc_d, pm_cu(stream));
}

int pm_step(pm_plan *p, float *pos_d, float *vel_d, int64_t np, double mass, double a, double da,
            double f_a1, double omega_m0, float *rho_d, pm_stream_t stream)
{
    PM_ARGS(p && (np == 0 || (pos_d && vel_d)) && np >= 0 && np <= p->np_cap && a != 0.0);
    DeviceGuard guard;
    PM_TRY(guard.enter(p->device));
    cudaStream_t st = pm_cu(stream);
    float *rho = rho_d ? rho_d : p->mesh;
    p->rkeys_valid = false;
    p->rsorted_n = 0;
    p->rsort_done = false;
    pm_prof_mark(p, 0, st);
    PM_TRY(pm_k_cell_keys(p, pos_d, np, np, p->keys, nullptr, st));
    pm_prof_mark(p, PM_STAGE_KEYS + 1, st);
    PM_TRY(pm_k_sort(p, np, 0, st));
    pm_prof_mark(p, PM_STAGE_SORT + 1, st);
    PM_TRY(pm_k_row_offsets(p, np, st));
    pm_prof_mark(p, PM_STAGE_ROWS + 1, st);
    PM_TRY(pm_k_deposit(p, pos_d, np, mass, rho, st));
    pm_prof_mark(p, PM_STAGE_DEPOSIT + 1, st);
    p->rho_mean_hint = (double)np * mass / ((double)p->nc * p->nc * p->nc);   // CIC weights sum to 1 (density.py:28-47)
    PM_TRY(pm_k_poisson(p, rho, a, omega_m0, p->mesh2, st));  // marks R2C, GREEN, C2R itself
    PM_TRY(pm_k_gather_kick_drift(p, pos_d, vel_d, np, p->mesh2, a, f_a1, da, nullptr, st));
    pm_prof_mark(p, PM_STAGE_GATHER + 1, st);
    if (p->prof_ev && p->prof_n < p->prof_cap) ++p->prof_n;
    return PM_OK;
}

int pm_plan_profile_begin(pm_plan *p, int max_steps)
{
    PM_ARGS(p && max_steps >= 0 && max_steps <= 4096);
    DeviceGuard guard;
    PM_TRY(guard.enter(p->device));
    if (p->prof_ev) {
        for (int i = 0; i < p->prof_cap * (PM_NUM_STAGES + 1); ++i) cudaEventDestroy(p->prof_ev[i]);
        free(p->prof_ev);
        p->prof_ev = nullptr;
    }
    p->prof_cap = p->prof_n = 0;
    if (max_steps == 0) return PM_OK;
    const int n = max_steps * (PM_NUM_STAGES + 1);
    p->prof_ev = (cudaEvent_t *)calloc(n, sizeof(cudaEvent_t));
    if (!p->prof_ev) return PM_ERR_NOMEM;
    for (int i = 0; i < n; ++i) PM_CUDA(cudaEventCreate(&p->prof_ev[i]));
    p->prof_cap = max_steps;
    return PM_OK;
}

int pm_plan_profile_read(pm_plan *p, float *ms, int *n_steps)
{
    PM_ARGS(p && ms && n_steps);
    DeviceGuard guard;
    PM_TRY(guard.enter(p->device));
    *n_steps = p->prof_n;
    for (int s = 0; s < p->prof_n; ++s) {
        cudaEvent_t *e = p->prof_ev + (size_t)s * (PM_NUM_STAGES + 1);
        PM_CUDA(cudaEventSynchronize(e[PM_NUM_STAGES]));
        for (int k = 0; k < PM_NUM_STAGES; ++k)
            PM_CUDA(cudaEventElapsedTime(&ms[s * PM_NUM_STAGES + k], e[k], e[k + 1]));
    }
    p->prof_n = 0;
    return PM_OK;
}

// ---- resident particle state -------------------------------------------------------------------
// The two halves of a resident step.  Set rcur is stored in the cell order of the previous sort; the
// first half orders it anew and deposits, the second solves for the potential and moves the particles
// into the other set.  p->rsort_done: (keys_sorted, order_sorted, row_start) describe set rcur as it
// is now -- a second deposit of the same state (the caller wants the density again) must not sort
// again, because the incremental sort compares against the keys the list was LAST sorted by.
static int resident_sort(pm_plan *p, cudaStream_t st)
{
    const int64_t np = p->rnp;
    if (p->rsort_done) return PM_OK;
    if (!p->rkeys_valid) {
        PM_TRY(pm_k_cell_keys(p, p->rpos[p->rcur], np, p->rstride, p->keys, nullptr, st));
        p->rsorted_n = 0;
    }
    pm_prof_mark(p, PM_STAGE_KEYS + 1, st);
    PM_TRY(pm_k_sort(p, np, p->rsorted_n, st));
    pm_prof_mark(p, PM_STAGE_SORT + 1, st);
    PM_TRY(pm_k_row_offsets(p, np, st));
    pm_prof_mark(p, PM_STAGE_ROWS + 1, st);
    p->rsort_done = true;
    return PM_OK;
}

static int resident_deposit(pm_plan *p, double mass, float *rho, cudaStream_t st)
{
    pm_prof_mark(p, 0, st);
    PM_TRY(resident_sort(p, st));
    PM_TRY(pm_k_deposit(p, p->rpos[p->rcur], p->rstride, mass, rho, st));
    pm_prof_mark(p, PM_STAGE_DEPOSIT + 1, st);
    return PM_OK;
}

static int resident_advance(pm_plan *p, const float *rho, double rho_mean, double a, double da, double f_a1,
                            double omega_m0, cudaStream_t st)
{
    const int64_t np = p->rnp;
    PM_TRY(resident_sort(p, st));           // no-op after resident_deposit
    p->rho_mean_hint = rho_mean;
    PM_TRY(pm_k_poisson(p, rho, a, omega_m0, p->mesh2, st));
    PM_TRY(pm_k_gather_kick_drift_resident(p, p->mesh2, a, f_a1, da, st));
    pm_prof_mark(p, PM_STAGE_GATHER + 1, st);
    if (p->prof_ev && p->prof_n < p->prof_cap) ++p->prof_n;
    p->rcur ^= 1;
    p->rkeys_valid = (np > 0);
    p->rsorted_n = np;   // set rcur is stored in the order of this step's sort
    p->rsort_done = false;
    return PM_OK;
}

static int resident_step(pm_plan *p, double mass, double a, double da, double f_a1,
                         double omega_m0, float *rho_d, cudaStream_t st)
{
    float *rho = rho_d ? rho_d : p->mesh;
    PM_TRY(resident_deposit(p, mass, rho, st));
    const double mean = (double)p->rnp * mass / ((double)p->nc * p->nc * p->nc);   // CIC weights sum to 1 (density.py:28-47)
    return resident_advance(p, rho, mean, a, da, f_a1, omega_m0, st);
}

int pm_particles_load(pm_plan *p, const float *pos_d, const float *vel_d, int64_t np,
                      pm_stream_t stream)
{
    PM_ARGS(p && (np == 0 || (pos_d && vel_d)) && np >= 0 && np <= p->np_cap);
    DeviceGuard guard;
    PM_TRY(guard.enter(p->device));
    cudaStream_t st = pm_cu(stream);
    const size_t b = (size_t)np * 3 * sizeof(float);
    p->rcur = 0;
    p->rnp = p->rtotal = np;
    p->rkeys_valid = false;
    p->rsorted_n = 0;
    p->rsort_done = false;
    if (np == 0) return PM_OK;
    (void)b;
    const size_t w = (size_t)np * sizeof(float), pitch = (size_t)p->rstride * sizeof(float);
    PM_CUDA(cudaMemcpy2DAsync(p->rpos[0], pitch, pos_d, w, w, 3, cudaMemcpyDeviceToDevice, st));
    PM_CUDA(cudaMemcpy2DAsync(p->rvel[0], pitch, vel_d, w, w, 3, cudaMemcpyDeviceToDevice, st));
    PM_CUDA(cudaMemcpyAsync(p->rid[0], p->iota, (size_t)np * 4, cudaMemcpyDeviceToDevice, st));
    return PM_OK;
}

int pm_particles_store(pm_plan *p, float *pos_d, float *vel_d, pm_stream_t stream)
{
    PM_ARGS(p && (p->rnp == 0 || (pos_d && vel_d)));
    DeviceGuard guard;
    PM_TRY(guard.enter(p->device));
    return pm_k_unpermute(p, pos_d, vel_d, pm_cu(stream));
}

int64_t pm_particles_count(const pm_plan *p) { return p ? p->rnp : 0; }

int pm_particles_order(pm_plan *p, uint32_t *ids_d, pm_stream_t stream)
{
    PM_ARGS(p && (p->rnp == 0 || ids_d));
    DeviceGuard guard;
    PM_TRY(guard.enter(p->device));
    if (p->rnp)
        PM_CUDA(cudaMemcpyAsync(ids_d, p->rid[p->rcur], (size_t)p->rnp * 4, cudaMemcpyDeviceToDevice,
                                pm_cu(stream)));
    return PM_OK;
}

// ---- the resident step as a CUDA graph -----------------------------------------------------------
// In steady state a resident step is a fixed sequence of ~20 launches whose only step-dependent inputs
// are six scalars (the gather's kick/drift factors and the Green's factor): the sort decides between its
// two ways on the device, the buffer sets alternate with period two.  So two graphs are captured (one
// per buffer-set parity) from the very code path that runs eagerly, with the kernels told to read the
// scalars from PmStepParams in device memory, and a replay only re-parameterises the one-thread kernel
// node that writes that struct.  Reference: the loop body src/pmesh.py:56-63.
__global__ void k_set_step_params(PmStepParams v, PmStepParams *dst) { *dst = v; }

static void graph_drop(pm_plan *p, int k)
{
    if (p->graph_exec[k]) cudaGraphExecDestroy((cudaGraphExec_t)p->graph_exec[k]);
    if (p->graph_tpl[k]) cudaGraphDestroy((cudaGraph_t)p->graph_tpl[k]);
    p->graph_exec[k] = nullptr;
    p->graph_node[k] = nullptr;
    p->graph_tpl[k] = nullptr;
}

static PmStepParams step_params_for(const pm_plan *p, double a, double da, double f_a1, double omega_m0)
{
    PmStepParams v;
    memset(&v, 0, sizeof(v));
    pm_gather_step_scalars(a, f_a1, da, &v);
    const double m = (double)p->nc * p->nc * p->nc;
    v.green_scale = (float)(-3 * omega_m0 / 8 / a / m);
    return v;
}

static bool graph_eligible(const pm_plan *p)
{
    // steady state only: keys and mover counts come from the previous gather, the state is not yet sorted,
    // nothing is being profiled, own FFT, tiled gather
    return p->use_graph && !p->slab && p->own_fft && !p->deconv && !p->kgrad && !p->prof_ev && p->rnp > 0 && p->rkeys_valid && p->inc_counted &&
           p->rsorted_n == p->rnp && !p->rsort_done && p->sort_mode == PM_SORT_AUTO && pm_gather_graphable(p);
}

static int resident_step_graphed(pm_plan *p, double mass, double a, double da, double f_a1, double omega_m0,
                                 float *rho_d, cudaStream_t st)
{
    if (!graph_eligible(p) || !st) return resident_step(p, mass, a, da, f_a1, omega_m0, rho_d, st);   // no capture on the legacy stream
    const int k = p->rcur;
    float *rho = rho_d ? rho_d : p->mesh;
    PmStepParams v = step_params_for(p, a, da, f_a1, omega_m0);
    if (p->graph_exec[k] && (p->graph_rho[k] != rho || p->graph_mass[k] != mass || p->graph_omega[k] != omega_m0 ||
                             p->graph_np[k] != p->rnp || p->graph_stream[k] != st))
        graph_drop(p, k);
    if (!p->graph_exec[k]) {
        // capture this step's own launch sequence
        cudaGraph_t graph = nullptr;
        if (cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
            cudaGetLastError();
            return resident_step(p, mass, a, da, f_a1, omega_m0, rho_d, st);
        }
        k_set_step_params<<<1, 1, 0, st>>>(v, p->step_params_d);
        ++g_pm_launches;
        p->graph_params = p->step_params_d;
        const bool s_keys = p->rkeys_valid, s_cnt = p->inc_counted, s_done = p->rsort_done;
        const int s_cur = p->rcur;
        const int64_t s_n = p->rsorted_n;
        const int rc = resident_step(p, mass, a, da, f_a1, omega_m0, rho_d, st);
        p->graph_params = nullptr;
        const cudaError_t ce = cudaStreamEndCapture(st, &graph);
        // the captured launches did not run: put the bookkeeping back to "before the step"
        p->rkeys_valid = s_keys; p->inc_counted = s_cnt; p->rsort_done = s_done; p->rcur = s_cur; p->rsorted_n = s_n;
        if (rc != PM_OK || ce != cudaSuccess || !graph) {
            cudaGetLastError();
            if (graph) cudaGraphDestroy(graph);
            p->use_graph = false;            // this configuration cannot be captured: stay eager
            return resident_step(p, mass, a, da, f_a1, omega_m0, rho_d, st);
        }
        cudaGraphExec_t exec = nullptr;
        cudaGraphNode_t node = nullptr;
        {
            size_t n = 0;
            cudaGraphGetNodes(graph, nullptr, &n);
            cudaGraphNode_t *nodes = (cudaGraphNode_t *)malloc(sizeof(cudaGraphNode_t) * (n ? n : 1));
            if (nodes && cudaGraphGetNodes(graph, nodes, &n) == cudaSuccess)
                for (size_t i = 0; i < n && !node; ++i) {
                    cudaGraphNodeType t;
                    cudaKernelNodeParams kp;
                    if (cudaGraphNodeGetType(nodes[i], &t) == cudaSuccess && t == cudaGraphNodeTypeKernel &&
                        cudaGraphKernelNodeGetParams(nodes[i], &kp) == cudaSuccess && kp.func == (void *)k_set_step_params)
                        node = nodes[i];
                }
            free(nodes);
        }
        if (!node || cudaGraphInstantiate(&exec, graph, 0) != cudaSuccess) {
            cudaGetLastError();
            cudaGraphDestroy(graph);
            p->use_graph = false;
            return resident_step(p, mass, a, da, f_a1, omega_m0, rho_d, st);
        }
        // the node handle that cudaGraphExecKernelNodeSetParams takes belongs to this graph: keep it alive
        p->graph_exec[k] = exec; p->graph_node[k] = node; p->graph_tpl[k] = graph;
        p->graph_rho[k] = rho; p->graph_mass[k] = mass; p->graph_omega[k] = omega_m0; p->graph_np[k] = p->rnp;
        p->graph_stream[k] = st;
    } else {
        PmStepParams *dst = p->step_params_d;
        void *args[2] = {&v, &dst};
        cudaKernelNodeParams kp;
        memset(&kp, 0, sizeof(kp));
        kp.func = (void *)k_set_step_params;
        kp.gridDim = dim3(1); kp.blockDim = dim3(1); kp.sharedMemBytes = 0; kp.kernelParams = args; kp.extra = nullptr;
        PM_CUDA(cudaGraphExecKernelNodeSetParams((cudaGraphExec_t)p->graph_exec[k], (cudaGraphNode_t)p->graph_node[k], &kp));
    }
    PM_CUDA(cudaGraphLaunch((cudaGraphExec_t)p->graph_exec[k], st));
    ++p->graph_replays;
    // what resident_step leaves behind on the host
    p->inc_counted = true;      // the gather counted the movers of the next sort
    p->rcur ^= 1;
    p->rkeys_valid = true;
    p->rsorted_n = p->rnp;
    p->rsort_done = false;
    p->rows_valid = true;
    p->sort_last_n = p->rnp;
    return PM_OK;
}

int pm_plan_graph_replays(const pm_plan *p) { return p ? p->graph_replays : PM_ERR_INVALID; }

int pm_plan_set_graph(pm_plan *p, int on)
{
    if (!p) return PM_ERR_INVALID;
    p->use_graph = (on != 0);
    return PM_OK;
}

int pm_step_resident(pm_plan *p, double mass, double a, double da, double f_a1, double omega_m0,
                     float *rho_d, pm_stream_t stream)
{
    PM_ARGS(p && a != 0.0);
    DeviceGuard guard;
    PM_TRY(guard.enter(p->device));
    return resident_step_graphed(p, mass, a, da, f_a1, omega_m0, rho_d, pm_cu(stream));
}

int pm_resident_deposit(pm_plan *p, double mass, float *rho_d, pm_stream_t stream)
{
    PM_ARGS(p && rho_d && !p->slab);
    DeviceGuard guard;
    PM_TRY(guard.enter(p->device));
    return resident_deposit(p, mass, rho_d, pm_cu(stream));
}

int pm_resident_advance(pm_plan *p, const float *rho_d, double rho_mean, double a, double da, double f_a1,
                        double omega_m0, pm_stream_t stream)
{
    PM_ARGS(p && rho_d && a != 0.0 && !p->slab);
    DeviceGuard guard;
    PM_TRY(guard.enter(p->device));
    return resident_advance(p, rho_d, rho_mean, a, da, f_a1, omega_m0, pm_cu(stream));
}

// first particle of range k of a host-buffer step's PM_HOST_CHUNKS ranges (multiples of 64 particles)
static inline int64_t host_chunk_begin(int64_t np, int k)
{
    return k >= PM_HOST_CHUNKS ? np : ((np * k / PM_HOST_CHUNKS) & ~(int64_t)63);
}

int pm_step_host_range(int64_t np, int k, int64_t *i0, int64_t *i1, int *n_ranges)
{
    if (np < 0 || k < 0 || k >= PM_HOST_CHUNKS || !i0 || !i1) return PM_ERR_INVALID;
    *i0 = host_chunk_begin(np, k);
    *i1 = host_chunk_begin(np, k + 1);
    if (n_ranges) *n_ranges = PM_HOST_CHUNKS;
    return PM_OK;
}

int pm_host_register(void *ptr, size_t bytes)
{
    if (!ptr || bytes == 0) return PM_ERR_INVALID;
    if (cudaHostRegister(ptr, bytes, cudaHostRegisterDefault) != cudaSuccess) {
        cudaGetLastError();      // not an error of this library's work: leave no state behind
        return PM_ERR_UNSUPPORTED;
    }
    return PM_OK;
}

int pm_host_unregister(void *ptr)
{
    if (!ptr) return PM_ERR_INVALID;
    if (cudaHostUnregister(ptr) != cudaSuccess) {
        cudaGetLastError();
        return PM_ERR_UNSUPPORTED;
    }
    return PM_OK;
}

int pm_step_host(pm_plan *p, float *pos_h, float *vel_h, int64_t np, double mass, double a,
                 double da, double f_a1, double omega_m0, float *rho_h)
{
    PM_ARGS(p && (np == 0 || (pos_h && vel_h)) && np >= 0 && np <= p->np_cap && a != 0.0);
    DeviceGuard guard;
    PM_TRY(guard.enter(p->device));
    const size_t pbytes = (size_t)np * 3 * sizeof(float);
    const size_t mbytes = (size_t)p->nc * p->nc * p->nc * sizeof(float);
    // This call works on the plan's own streams, and it returns only when its results are in host memory: wait
    // first for whatever the caller still has in flight on this device.  A pm_step / density() enqueued on the
    // caller's stream a moment ago uses the same workspace (keys, permutation, meshes), and nothing else orders
    // the plan's streams behind it -- with the sort now starting 2.4 ms into the call, such a step's gather was
    // still reading the permutation this call's sort overwrites (tests/test_gpu_parity.py, the 512^3 case).
    PM_CUDA(cudaDeviceSynchronize());
    // The host owns the state, so every call starts from the caller's particle order: upload into
    // set 0, sort, put the POSITIONS in cell order into set 1, deposit, Poisson solve.  The velocity upload
    // (s_up, after the positions, in ranges) overlaps all of that; the density download (s_down) overlaps
    // the Poisson solve.  Then the step is split where the velocities enter it: the gather proper -- the
    // stencil sums of every particle, positions and phi only -- runs on the cell-ordered set as soon as phi is
    // there and stores its three floats per particle at the particle's ORIGINAL index; kick and drift
    // (pm_push, the function the fused kernels call: same bits) then run in the caller's order, range by range
    // behind the velocity upload, each range downloaded while the next is pushed.  No velocity re-ordering, no
    // un-permute of the results, and the device-to-host link starts right after the last velocity range has
    // arrived (PM_HOST_SPLIT=0, or a plan without the warp-specialised gather: the fused resident step
    // followed by an un-permute, as before).
    p->rcur = 0;
    p->rnp = np;
    p->rkeys_valid = false;
    p->rsorted_n = 0;
    p->rsort_done = false;
    const size_t w = (size_t)np * sizeof(float), pitch = (size_t)p->rstride * sizeof(float);
    (void)pbytes;
    // PM_HOST_TIMING=1: a timeline of this call on stderr (events on the three streams; diagnostic only)
    static const bool timing = getenv("PM_HOST_TIMING") && atoi(getenv("PM_HOST_TIMING")) != 0;
    cudaEvent_t tl[8] = {};
    auto tmark = [&](int k, cudaStream_t s) {
        if (timing && cudaEventCreate(&tl[k]) == cudaSuccess) cudaEventRecord(tl[k], s);
    };
    tmark(0, p->s_main);
    cudaStream_t st = p->s_main;
    static const bool split_off = getenv("PM_HOST_SPLIT") && atoi(getenv("PM_HOST_SPLIT")) == 0;
    const bool want_split = np && !split_off && pm_gather_sums_ok(p);
    if (np) {
        // positions first and alone on the link: everything up to the gather needs only them.  (Issued
        // together, the two uploads share the link and the positions arrive last: measured 7.3 ms instead
        // of 3.6 ms before the first kernel could start.)  On the split route the sort is by mesh ROW, which
        // needs y and z only: those two rows go first and the keys and the sort run under the upload of x.
        if (want_split) {
            PM_CUDA(cudaMemcpy2DAsync(p->rpos[0] + p->rstride, pitch, pos_h + np, w, w, 2, cudaMemcpyHostToDevice, st));
            PM_CUDA(cudaEventRecord(p->ev_c, st));
            PM_CUDA(cudaStreamWaitEvent(p->s_up, p->ev_c, 0));
            PM_CUDA(cudaMemcpyAsync(p->rpos[0], pos_h, w, cudaMemcpyHostToDevice, p->s_up));
            PM_CUDA(cudaEventRecord(p->ev_x, p->s_up));
        } else {
            PM_CUDA(cudaMemcpy2DAsync(p->rpos[0], pitch, pos_h, w, w, 3, cudaMemcpyHostToDevice, st));
            PM_CUDA(cudaEventRecord(p->ev_c, st));
            PM_CUDA(cudaStreamWaitEvent(p->s_up, p->ev_c, 0));
        }
        for (int k = 0; k < PM_HOST_CHUNKS; ++k) {
            const int64_t i0 = host_chunk_begin(np, k), i1 = host_chunk_begin(np, k + 1);
            if (i1 > i0)
                PM_CUDA(cudaMemcpy2DAsync(p->rvel[0] + i0, pitch, vel_h + i0, w, (size_t)(i1 - i0) * sizeof(float), 3,
                                          cudaMemcpyHostToDevice, p->s_up));
            PM_CUDA(cudaEventRecord(p->ev_upchunk[k], p->s_up));
        }
    }
    PM_CUDA(cudaEventRecord(p->ev_a, p->s_up));
    tmark(1, st);             // positions uploaded (split route: y and z)
    tmark(2, p->s_up);        // velocities uploaded
    PM_TRY(pm_k_cell_keys(p, p->rpos[0], np, p->rstride, p->keys, nullptr, st, want_split));
    p->sort_rows_only = want_split;      // grouped by mesh row is all the split route needs (pm_k_sort)
    {
        const int rc = pm_k_sort(p, np, 0, st);
        p->sort_rows_only = false;
        if (rc != PM_OK) return rc;
    }
    PM_TRY(pm_k_row_offsets(p, np, st));
    if (want_split) PM_CUDA(cudaStreamWaitEvent(st, p->ev_x, 0));     // from here on x is needed
    if (np) {
        // The caller's order is arbitrary, and reading it through the sort permutation costs the deposit and
        // above all the gather a random 4-byte access per value (gather: 2.8 ms instead of 0.5 ms).  So the
        // particles are put in cell order ONCE -- positions now, velocities when they have arrived -- into
        // set 1, which then is a resident state like any other: identity permutation, ids = the caller's index.
        PM_TRY(pm_k_reorder_rows(p, p->rpos[0], p->order_sorted, np, p->rpos[1], st));
        PM_CUDA(cudaMemcpyAsync(p->rid[1], p->order_sorted, (size_t)np * 4, cudaMemcpyDeviceToDevice, st));
        PM_CUDA(cudaMemcpyAsync(p->order_sorted, p->iota, (size_t)np * 4, cudaMemcpyDeviceToDevice, st));
    }
    p->rcur = 1;
    PM_TRY(pm_k_deposit(p, p->rpos[1], p->rstride, mass, p->mesh, st));
    if (rho_h) {
        PM_CUDA(cudaEventRecord(p->ev_b, st));
        PM_CUDA(cudaStreamWaitEvent(p->s_down, p->ev_b, 0));
        PM_CUDA(cudaMemcpyAsync(rho_h, p->mesh, mbytes, cudaMemcpyDeviceToHost, p->s_down));
    }
    p->rho_mean_hint = (double)np * mass / ((double)p->nc * p->nc * p->nc);
    PM_TRY(pm_k_poisson(p, p->mesh, a, omega_m0, p->mesh2, st));
    tmark(3, st);             // potential ready
    bool split = want_split;
    if (split) {
        const int rc = pm_k_gather_sums(p, p->mesh2, st);   // set 1 (cell order) -> sums at the original index
        if (rc == PM_ERR_UNSUPPORTED) split = false;        // nothing was launched: the fused route below
        else if (rc != PM_OK) return rc;
    }
    if (split) {
        tmark(4, st);         // gather done
        p->inc_counted = false;
        for (int k = 0; k < PM_HOST_CHUNKS; ++k) {
            const int64_t i0 = host_chunk_begin(np, k), i1 = host_chunk_begin(np, k + 1);
            if (i1 <= i0) continue;
            PM_CUDA(cudaStreamWaitEvent(st, p->ev_upchunk[k], 0));
            // caller's order: uploaded positions / velocities of set 0 -> dense [3][np] rows in set 1
            PM_TRY(pm_k_push_rows(p, p->rpos[0], p->rvel[0], i0, i1, a, f_a1, da, p->rpos[1], p->rvel[1], st));
            PM_CUDA(cudaEventRecord(p->ev_chunk[k], st));
            PM_CUDA(cudaStreamWaitEvent(p->s_down, p->ev_chunk[k], 0));
            const size_t cw = (size_t)(i1 - i0) * sizeof(float);
            PM_CUDA(cudaMemcpy2DAsync(pos_h + i0, w, p->rpos[1] + i0, w, cw, 3, cudaMemcpyDeviceToHost, p->s_down));
            PM_CUDA(cudaMemcpy2DAsync(vel_h + i0, w, p->rvel[1] + i0, w, cw, 3, cudaMemcpyDeviceToHost, p->s_down));
        }
    } else {
    PM_CUDA(cudaStreamWaitEvent(st, p->ev_a, 0));
    if (np) PM_TRY(pm_k_reorder_rows(p, p->rvel[0], p->rid[1], np, p->rvel[1], st));
    PM_TRY(pm_k_gather_kick_drift_resident(p, p->mesh2, a, f_a1, da, st));     // set 1 -> set 0
    tmark(4, st);             // gather done
    p->rcur = 0;
    // un-permute into set 1 as dense [3][np] arrays (the unpermute kernels write stride np)
    if (np && pm_unpermute_aos_ok(p, p->rpos[1], p->rvel[1])) {
        // one sector-wide scatter of all particles into the idle half-spectrum buffer, then ranges of the
        // caller's order are streamed into the six rows and downloaded while the next range is formed:
        // the device-to-host link starts ~0.2 ms after the gather instead of after a full un-permute
        PM_TRY(pm_k_unpermute_scatter_aos(p, st));
        for (int k = 0; k < PM_HOST_CHUNKS; ++k) {
            const int64_t i0 = host_chunk_begin(np, k), i1 = host_chunk_begin(np, k + 1);
            if (i1 <= i0) continue;
            PM_TRY(pm_k_aos_rows_range(p, i0, i1, p->rpos[1], p->rvel[1], st));
            PM_CUDA(cudaEventRecord(p->ev_chunk[k], st));
            PM_CUDA(cudaStreamWaitEvent(p->s_down, p->ev_chunk[k], 0));
            const size_t cw = (size_t)(i1 - i0) * sizeof(float);
            PM_CUDA(cudaMemcpy2DAsync(pos_h + i0, w, p->rpos[1] + i0, w, cw, 3, cudaMemcpyDeviceToHost, p->s_down));
            PM_CUDA(cudaMemcpy2DAsync(vel_h + i0, w, p->rvel[1] + i0, w, cw, 3, cudaMemcpyDeviceToHost, p->s_down));
        }
    } else if (np) {
        PM_TRY(pm_k_unpermute(p, p->rpos[1], p->rvel[1], st));
        PM_CUDA(cudaEventRecord(p->ev_c, st));
        PM_CUDA(cudaMemcpyAsync(pos_h, p->rpos[1], 3 * w, cudaMemcpyDeviceToHost, st));
        PM_CUDA(cudaStreamWaitEvent(p->s_up, p->ev_c, 0));
        PM_CUDA(cudaMemcpyAsync(vel_h, p->rvel[1], 3 * w, cudaMemcpyDeviceToHost, p->s_up));
    }
    }
    tmark(5, st);             // un-permute kernels done
    tmark(6, p->s_down);      // downloads done (chunked path)
    PM_CUDA(cudaStreamSynchronize(p->s_main));
    PM_CUDA(cudaStreamSynchronize(p->s_up));
    PM_CUDA(cudaStreamSynchronize(p->s_down));
    if (timing && tl[0]) {
        float t[8] = {};
        for (int k = 1; k <= 6; ++k)
            if (tl[k]) cudaEventElapsedTime(&t[k], tl[0], tl[k]);
        fprintf(stderr, "pm_step_host timeline (ms): pos up %.2f | vel up %.2f | phi ready %.2f | gather done %.2f | un-permute done %.2f | download done %.2f\n",
                t[1], t[2], t[3], t[4], t[5], t[6]);
        for (int k = 0; k < 8; ++k)
            if (tl[k]) cudaEventDestroy(tl[k]);
    }
    p->rnp = 0;  // the resident buffers were scratch for this call
    p->rkeys_valid = false;
    return PM_OK;
}

}  // extern "C"
