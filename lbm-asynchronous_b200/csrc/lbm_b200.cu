// lbm_b200.cu -- host side of the C ABI declared in include/lbm_b200.h.
//
// Owns the device-resident state of one run (the reference's cells / tmp_cells / obstacles / av_vels,
// SerialCode/d2q9-bgk.c:531-543,610) and drives the sm_100a kernels of lbm_kernels.cuh:
//   * lattices: 2 x 9 SoA planes per row slab, row pitch a multiple of 32 floats;
//   * obstacles: 1 bit per cell;
//   * the `for tt` loop (SerialCode:166-169) as CUDA graphs covering GRAPH_STEPS timesteps (16 launches of
//     step2_kernel, which advances two steps per pass over HBM, or 32 single-step launches), the step
//     index living in device memory so that one graph serves the whole run;
//   * av_vels: exact integer sums of |u| per step kept on the device for the whole run, combined in
//     a fixed way by the host afterwards (replaces the MPI_Reduce of MPI/d2q9-bgk.c:298-309);
//   * several slabs: halo rings + flags in peer-mapped memory, written by the neighbour's step kernel
//     (replaces MPI_Isend/Irecv/Waitall|Testall, MPI_Waitall/d2q9-bgk.c:225-253).
// There is no CPU fallback anywhere in this file: without a CUDA device every compute entry point
// returns LBM_ENODEVICE.
#include "../../include/lbm_b200.h"
#include "lbm_kernels.cuh"
#include "lbm_tma_kernel.cuh"
#include "lbm_fused2_kernel.cuh"
#include "lbm_cluster_kernel.cuh"
#include "lbm_ll_kernel.cuh"
#include "lbm_band_kernel.cuh"

#include <cudaTypedefs.h>
#include <unistd.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <new>
#include <vector>

using namespace lbm;

namespace {

constexpr int GRAPH_STEPS = 32;      // kernel nodes per graph (even: lattice parity is restored)
constexpr uint32_t HANDLE_MAGIC = 0x4c424d48u; // "LBMH"

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return code;
}

#define CU(call)                                                                                                   \
    do {                                                                                                           \
        cudaError_t e__ = (call);                                                                                  \
        if (e__ != cudaSuccess)                                                                                    \
            return fail(e__ == cudaErrorMemoryAllocation ? LBM_ENOMEM : LBM_ECUDA, "%s failed: %s (%s:%d)", #call, \
                        cudaGetErrorString(e__), __FILE__, __LINE__);                                              \
    } while (0)

struct HaloHandle { // what lbm_halo_export writes (<= LBM_HALO_HANDLE_BYTES)
    uint32_t magic;
    int32_t pid;
    int32_t device;
    int32_t pitch;
    int32_t ring;
    int32_t config;     // everything both sides of a link must agree on (see halo_config_word)
    uint64_t local_ptr; // valid inside process `pid` only
    uint64_t bytes;
    cudaIpcMemHandle_t ipc;
};
static_assert(sizeof(HaloHandle) <= LBM_HALO_HANDLE_BYTES, "handle too large");

struct Slab {
    int device = 0;
    int row0 = 0, row1 = 0, rows = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    float* lat[2] = {nullptr, nullptr}; // 9 planes x rows x pitch each
    uint32_t* obst = nullptr;           // rows x opitch
    unsigned long long* fluid_dev = nullptr;
    long long fluid = 0;
    // halo block: [south ring][north ring][flags: 2 x 16 u64][obstacle bits of my row 0 and my last row]
    // (one allocation: one IPC handle)
    char* halo_block = nullptr;
    size_t halo_bytes = 0;
    float* ring_s = nullptr;
    float* ring_n = nullptr;
    unsigned long long* flag_s = nullptr;
    unsigned long long* flag_n = nullptr;
    uint32_t* edge_obst = nullptr;          // inside the halo block: [2][opitch], for the neighbours to copy
    uint32_t* obst_halo = nullptr;          // [2][opitch]: the south neighbour's last row, the north neighbour's row 0
    unsigned long long* arrive = nullptr;   // [2] last-arriver counters of halo_arrive()
    // the neighbours' sides facing me
    float* peer_ring_s = nullptr;           // south neighbour's north ring
    unsigned long long* peer_flag_s = nullptr;
    float* peer_ring_n = nullptr;           // north neighbour's south ring
    unsigned long long* peer_flag_n = nullptr;
    void* ipc_open[2] = {nullptr, nullptr}; // mappings to close
    int* ctrl = nullptr;
    int* error = nullptr;
    unsigned long long* sums = nullptr;
    unsigned long long** sums_ref = nullptr; // device word holding `sums` (what the graph kernels read)
    size_t sums_steps = 0;
    unsigned long long* state_sums = nullptr; // SUM_WORDS u64 + 1 double
    cudaGraphExec_t graph[2] = {nullptr, nullptr}; // by parity of the source lattice
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    std::vector<cudaEvent_t> dbg_events; // LBM_DEBUG: one per graph replay of the last run
    // launch geometry of step_vec4_kernel / step_scalar_kernel
    int tw_shift = 0, nbx = 0, ngroups = 0, nxv = 0, block = 0, nslots = 1;
    unsigned grid = 0;    // every row (mode 0)
    unsigned grid_b = 0;  // boundary rows only (mode 1)
    bool vec4 = false;
    int accel_row = -1;
    // whole runs in one cooperative launch (step_loop_kernel), L2-resident single-slab grids
    bool use_loop = false;
    unsigned loop_grid = 0;
    int loop_vec = 4, loop_block = 256, loop_tw_shift = 0, loop_nbx = 0, loop_nby = 0, loop_nxv = 0, loop_ntiles = 0;
    unsigned* loop_barrier = nullptr;
    // whole runs in one launch of step_cluster_kernel: the lattice lives in the shared memory of one cluster
    bool use_cluster = false;
    int cl_size = 0, cl_rpc = 0, cl_threads = 0;
    size_t cl_smem = 0;
    // whole runs in one cooperative launch of step_ll_kernel: a row per CTA, cells in registers, packets through L2
    bool use_ll = false;
    int ll_var = 0;              // 0: one cell per thread, 1: four, 2: two
    int ll_block = 0;
    size_t ll_smem = 0;
    uint4* ll_packets = nullptr; // [2 directions][2 parities][rows][pitch]
    uint4* ll_recv_s = nullptr;  // row slabs: packet areas [2][pitch] inside halo_block, written by the neighbours
    uint4* ll_recv_n = nullptr;
    uint4* peer_ll_s = nullptr;  // the south neighbour's ll_recv_n / the north neighbour's ll_recv_s (peer memory)
    uint4* peer_ll_n = nullptr;
    // whole runs in one cooperative launch of step_band_kernel: a band of rows per CTA, neighbour flags
    bool use_band = false;
    unsigned band_grid = 0;
    int band_per_sm = 0;            // > 0: band_grid = band_per_sm CTAs on each SM, rows dealt out per SM
    unsigned* band_flags = nullptr; // [band_grid][32]
    // interior rows through step_tma_kernel
    bool use_tma = false;
    CUtensorMap tmap[2];  // per lattice: boxes TMA_TX wide
    CUtensorMap tmapw[2]; // per lattice: boxes TMA_TXW wide (x-shifted planes)
    int tma_ntx = 0, tma_ntiles = 0, tma_dq = 0, tma_dr = 0;
    unsigned tma_grid = 0;
    // two timesteps per launch through step2_kernel
    bool use_f2 = false;
    CUtensorMap f2map[2];  // per lattice: boxes TMA_TX x SROWS
    CUtensorMap f2mapw[2]; // per lattice: boxes TMA_TXW x SROWS
    int f2_nsx = 0, f2_seg_h = 0, f2_nseg = 0, f2_nunits = 0;
    unsigned f2_grid = 0;
};

} // namespace

struct lbm_lattice {
    lbm_param_t p;
    lbm_options_t opt;
    int nslabs = 0;          // slabs held by this process
    int rank = 0, nranks = 1; // ring position of slab 0 / ring size (per-process mode), else 0 / nslabs
    bool per_process = false;
    bool connected = false;
    bool poisoned = false;    // an lbm_run failed after it had started to queue work
    bool interleaved = false; // some slabs share a device (and therefore a stream): step-major launches only
    Slab* slabs = nullptr;
    int pitch = 0, opitch = 0, ring = 2;
    int cur = 0;             // which lattice holds the current state
    long long steps_done = 0;
    long long epoch = 0;     // halo epochs done (every slab of the lattice counts the same sequence)
    long long launches = 0;
    long long run_first = 0; // absolute index of the first step of the last lbm_run
    int run_iters = 0;
    float w0 = 0, w1 = 0, w2 = 0;   // initial state, SerialCode:546-548
    float w1a = 0, w2a = 0;         // accelerate_flow weights, SerialCode:222-223
    unsigned long long timeout_ns = 30ull * 1000000000ull;
    void (*kernel)(StepArgs) = nullptr;                 // all rows, or the boundary rows next to step_tma_kernel
    void (*loop_kernel[2])(LoopArgs) = {nullptr, nullptr}; // step_loop_kernel: [0] 1 cell per thread, [1] 4 cells
    void (*tma_kernel)(CUtensorMap, CUtensorMap, TmaArgs) = nullptr; // interior rows (null: `kernel` does every row)
    int tma_ty = 0, tma_stages = 0, tma_minb = 0, tma_resident = 0, sm_count = 0;
    int loop_resident[2] = {0, 0};
    void (*band_kernel)(BandArgs) = nullptr;       // step_band_kernel (null: not available for this lattice)
    int band_block = 0, band_resident = 0;
    void (*ll_kernel[3])(LLArgs) = {nullptr, nullptr, nullptr}; // step_ll_kernel, 1 / 4 / 2 cells per thread (null: not available)
    int ll_resident[3] = {0, 0, 0};                   // resident CTAs per SM at the lattice's row length
    unsigned ll_flags = 0;                         // packet flags handed out so far (never reused)
    void (*cluster_kernel)(ClusterArgs) = nullptr; // step_cluster_kernel (null: not available for this lattice)
    int cl_cpt = 0, cl_vert = 0, cl_maxt = 0;      // its cells per thread, collision shape, thread limit
    int cl_sync = 0;                               // barrier fencing (tuning knob LBM_CL_SYNC)
    void (*f2_kernel)(CUtensorMap, CUtensorMap, Fused2Args) = nullptr; // pairs of steps (null: single steps only)
    int f2_r = 0, f2_srows = 0, f2_stages = 0, f2_minb = 0, f2_resident = 0;
    int f2_iter_rows = 0; // rows a CTA advances per iteration of its marching loop
    size_t f2_smem = 0;
    int prio_high = 0; // numerically lowest = most urgent stream / kernel-node priority of the device
    size_t tma_smem = 0;
};

namespace {

// ---- kernel table -------------------------------------------------------------------------------
typedef void (*step_fn)(StepArgs);

template <bool STRICT, int HINT>
step_fn vec4_by_block(int block, int minb)
{
    switch (block) {
    case 128: return minb >= 4 ? step_vec4_kernel<STRICT, HINT, 128, 4> : step_vec4_kernel<STRICT, HINT, 128, 1>;
    case 512: return step_vec4_kernel<STRICT, HINT, 512, 1>;
    default: return minb >= 2 ? step_vec4_kernel<STRICT, HINT, 256, 2> : step_vec4_kernel<STRICT, HINT, 256, 1>;
    }
}
template <bool STRICT>
step_fn vec4_by_hint(int hint, int block, int minb)
{
    switch (hint) {
    case 0: return vec4_by_block<STRICT, 0>(block, minb);
    case 1: return vec4_by_block<STRICT, 1>(block, minb);
    default: return vec4_by_block<STRICT, 2>(block, minb);
    }
}
step_fn scalar_kernel(bool strict, int block)
{
    if (block == 128) return strict ? step_scalar_kernel<true, 128> : step_scalar_kernel<false, 128>;
    return strict ? step_scalar_kernel<true, 256> : step_scalar_kernel<false, 256>;
}

typedef void (*tma_fn)(CUtensorMap, CUtensorMap, TmaArgs);
struct TmaChoice {
    int ty, stages, minb;
    tma_fn fn;
};
template <bool STRICT>
bool tma_by_shape(int ty, int stages, int minb, TmaChoice* c)
{
#define LBM_TMA_CASE(TY_, ST_, MB_)                               \
    if (ty == TY_ && stages == ST_ && minb == MB_) {              \
        *c = {TY_, ST_, MB_, step_tma_kernel<STRICT, TY_, ST_, MB_>}; \
        return true;                                              \
    }
    LBM_TMA_CASE(8, 2, 3)
    LBM_TMA_CASE(8, 2, 2)
    LBM_TMA_CASE(8, 3, 2)
    LBM_TMA_CASE(8, 3, 1)
    LBM_TMA_CASE(8, 4, 1)
    LBM_TMA_CASE(4, 3, 4)
    LBM_TMA_CASE(4, 4, 4)
    LBM_TMA_CASE(4, 4, 3)
    LBM_TMA_CASE(4, 6, 2)
    LBM_TMA_CASE(16, 2, 1)
    LBM_TMA_CASE(12, 3, 1)
    LBM_TMA_CASE(12, 2, 1)
    LBM_TMA_CASE(10, 4, 1)
    LBM_TMA_CASE(10, 3, 1)
    LBM_TMA_CASE(16, 3, 1)
#undef LBM_TMA_CASE
    return false;
}

typedef void (*f2_fn)(CUtensorMap, CUtensorMap, Fused2Args);
struct F2Choice {
    int r, srows, stages, minb;
    f2_fn fn;
};
template <bool STRICT>
bool f2_by_shape(int r, int srows, int stages, int minb, F2Choice* c)
{
#define LBM_F2_CASE(R_, S_, N_, M_)                                      \
    if (r == R_ && srows == S_ && stages == N_ && minb == M_) {          \
        *c = {R_, S_, N_, M_, step2_kernel<STRICT, R_, S_, N_, M_>};     \
        return true;                                                     \
    }
    LBM_F2_CASE(8, 4, 3, 2)
    LBM_F2_CASE(16, 8, 3, 1)
    LBM_F2_CASE(12, 4, 4, 1)
#undef LBM_F2_CASE
    return false;
}

struct ClusterChoice {
    int cpt, vert, maxt;
    void (*fn)(ClusterArgs);
};
template <bool STRICT>
bool cluster_by_shape(int cpt, int vert, int maxt, ClusterChoice* c)
{
#define LBM_CL_CASE(C_, V_, T_)                                                      \
    if (cpt == C_ && vert == V_ && maxt == T_) {                                     \
        *c = {C_, V_, T_, step_cluster_kernel<STRICT, C_, (V_ != 0), T_>};           \
        return true;                                                                 \
    }
    LBM_CL_CASE(4, 1, 256)
    LBM_CL_CASE(4, 0, 256)
    LBM_CL_CASE(4, 0, 512)
    LBM_CL_CASE(4, 0, 1024)
    LBM_CL_CASE(2, 0, 512)
    LBM_CL_CASE(2, 0, 1024)
    LBM_CL_CASE(1, 0, 1024)
#undef LBM_CL_CASE
    return false;
}

// opt.kernel:
//   0            library default.  Single steps: step_tma_kernel (strict: 10-row tiles, 4 stages, 1 CTA/SM; fast: 8-row
//                tiles) for the interior rows when nx % 4 == 0, nx >= 128 and the slab has >= 3 rows,
//                step_vec4_kernel / step_scalar_kernel for the rest.  Fast flavour with nx % 4 == 0, nx >= 128, every
//                slab >= 8 rows, halo_lag == 0: step2_kernel, TWO timesteps per launch (8 warps, stages of 4 rows, 3
//                stages, 2 CTAs/SM; an odd last step of a run takes the single-step kernels)
//   2RRSNM       step2_kernel with RR warps (rows per iteration), S rows per stage, N stages, M CTAs per SM asked of
//                the compiler (208432 216831 212441)
//   1TTSM        step_tma_kernel with TT rows per tile, S stages, M resident CTAs per SM asked of the compiler
//                (10823 10822 10832 10831 10841 10434 10444 10443 10462 11621 11631 11231 11221 11041 11031)
//   H M (10..39) step_vec4_kernel for every row: hint = H-1 (0 plain, 1 ld.nc.no_allocate, 2 + st.cs), min blocks M
//   99           step_scalar_kernel for every row
//   200          step_loop_kernel (all steps of a run in one cooperative launch); also the default for
//                single-slab grids whose two lattices fit in L2 (<= LOOP_MAX_CELLS cells).  201 / 204 force
//                its 1-cell / 4-cell per thread mapping (default: 1 cell up to LOOP_VEC4_CELLS cells per slab)
//   3000         step_cluster_kernel (all steps of a run in one launch, the lattice resident in the shared memory of one
//                cluster of up to 16 CTAs); also the default for single-slab grids of up to CLUSTER_MAX_CELLS cells.
//                3CVM forces C cells per thread (1, 2, 4), V = 1 both pairs of a thread side by side through the
//                collision, M x 256 threads per CTA at most (3411 3401 3402 3404 3202 3204 3104)
//   500          step_band_kernel (all steps of a run in one cooperative launch, a band of rows per CTA, neighbour flags
//                instead of a grid barrier); also the default for single-slab grids of up to LOOP_MAX_CELLS cells
//                that step_ll_kernel does not take and that have a row for every resident CTA.  5BM: B x 128 threads per CTA, M CTAs per SM asked of the
//                compiler (522 514 521; 521 runs both pairs of a thread side by side through the collision)
//   400          step_ll_kernel (all steps of a run in one cooperative launch, a row per CTA, cells in registers, rows
//                exchanging flagged 16-byte packets through L2, slabs on several GPUs through each other's memory); also
//                the default for slabs of up to LL_MAX_CELLS cells (one cell per thread, nx <= 1024) or LL4_MAX_CELLS
//                cells (two cells per thread, nx <= 1024) whose rows are all resident at once.  401 / 404 / 402 force one /
//                four / two cells per thread
constexpr long long LL_MAX_CELLS = 70000;   // up to 256 x 256: one cell per thread
constexpr int LL_CPT[3] = {1, 4, 2};           // cells per thread of step_ll_kernel's variants
constexpr int LL_MAXT[3] = {1024, 256, 512};   // their thread limits
constexpr long long LL4_MAX_CELLS = 300000; // up to 1024 x 256 (a quarter of the shipped 1024 x 1024 case): two
constexpr int LL_SLOTS = 8;               // slots of the per-step sums (one RED per CTA, step and word)
constexpr long long CLUSTER_MAX_CELLS = 32768; // 128 x 256: above, 16 SMs have more arithmetic than the whole GPU has latency
constexpr int CLUSTER_MAX_CTAS = 16;
constexpr int BAND_MIN_PER_SM = 3;                // step_band_kernel by default: rows per slab >= 3 x SMs
constexpr long long LOOP_HALO_MAX_CELLS = 1400000; // slabs on several GPUs: the same bound (2048 x 1024 on 2 GPUs: 15.7 us from
                                                   // the resident step loop against 24.7 from the step graphs)
constexpr long long LOOP_MAX_CELLS = 1400000; // 2 x 36 B x cells <= ~100 MB of the 126 MB L2
constexpr long long LOOP_VEC4_CELLS = 70000;  // up to 256 x 256: one cell per thread (<= 512 CTAs of 128 threads) beats four;
                                              // above, the one-counter grid barrier gets too slow for that many CTAs
struct KernelChoice {
    bool band;
    int band_block, band_minb;
    bool ll;
    bool cluster;
    int cl_cpt, cl_vert, cl_maxt; // 0: chosen from the grid's shape
    bool loop;
    bool vec4;
    int hint, block, minb;
    bool tma;
    int tma_ty, tma_stages, tma_minb;
    bool f2;
    int f2_r, f2_srows, f2_stages, f2_minb;
};
bool k_band_default(int kernel) { return kernel == 0; }
KernelChoice choose_kernel(const lbm_options_t& o, int nx)
{
    KernelChoice k;
    k.vec4 = (nx % 4 == 0) && o.kernel != 99;
    const bool band_code = (o.kernel >= 500 && o.kernel < 600);
    const bool ll_code = (o.kernel == 400 || o.kernel == 401 || o.kernel == 402 || o.kernel == 404);
    const bool cl_code = (o.kernel >= 3000 && o.kernel < 4000) || ll_code || band_code; // everything else as the default
    k.band = k_band_default(o.kernel) || band_code;
    k.band_block = 128, k.band_minb = 4;
    if (band_code && o.kernel != 500) k.band_block = 128 * ((o.kernel / 10) % 10), k.band_minb = o.kernel % 10;
    k.ll = (o.kernel == 0 || ll_code);
    k.cluster = (o.kernel >= 3000 && o.kernel < 4000); // not a default: step_ll_kernel is faster wherever both apply
    k.cl_cpt = k.cl_vert = k.cl_maxt = 0;
    if (k.cluster && o.kernel != 3000) k.cl_cpt = (o.kernel / 100) % 10, k.cl_vert = (o.kernel / 10) % 10, k.cl_maxt = 256 * (o.kernel % 10);
    k.loop = (o.kernel == 0 || cl_code || o.kernel == 200 || o.kernel == 201 || o.kernel == 204);
    k.hint = 0;
    k.minb = 1;
    k.block = (o.block == 128 || o.block == 256 || o.block == 512) ? o.block : 256;
    k.tma = k.vec4 && nx >= TMA_TX && (o.kernel == 0 || cl_code || (o.kernel >= 200 && o.kernel <= 204) || o.kernel >= 10000);
    // best of the sweeps (profiles/r02_kernel_sweep.md).  With the packed collision both flavours want few, fat warps:
    // the strict flavour 10 consumer warps with ~124 registers each (both pairs of a thread side by side through the
    // collision) and a 4-stage pipeline of 10-row tiles -- 92 GLUPS at 8192^2, HBM bound; the fast flavour the 8-row
    // tiles of round 1
    if (o.arith == LBM_ARITH_FAST)
        k.tma_ty = 8, k.tma_stages = 4, k.tma_minb = 1;
    else
        k.tma_ty = 10, k.tma_stages = 4, k.tma_minb = 1;
    // pairs of steps (step2_kernel): half the HBM traffic, but bound by the latency of the collision's dependency
    // chains at 16 warps of 128 registers per SM.  Default for the fast flavour (93-96 GLUPS against 90 from single
    // steps); the strict flavour is faster on single steps (92 against 83-85), where it reaches the HBM roofline.
    // An explicit 2RRSNM code selects it for either; an explicit single-step variant (1TTSM, H M, 99) or a
    // deterministic halo lag (defined per single step) switches it off
    const bool f2_default = (o.kernel == 0 || cl_code || (o.kernel >= 200 && o.kernel <= 204)) && o.arith == LBM_ARITH_FAST;
    k.f2 = k.tma && (f2_default || o.kernel >= 200000) && o.halo_lag == 0;
    k.f2_r = 8, k.f2_srows = 4, k.f2_stages = 3, k.f2_minb = 2;
    if (o.kernel >= 200000) {
        k.f2_r = (o.kernel - 200000) / 1000;
        k.f2_srows = (o.kernel / 100) % 10;
        k.f2_stages = (o.kernel / 10) % 10;
        k.f2_minb = o.kernel % 10;
    } else if (o.kernel >= 10000) {
        k.tma_ty = (o.kernel - 10000) / 100;
        k.tma_stages = (o.kernel / 10) % 10;
        k.tma_minb = o.kernel % 10;
    } else if (o.kernel > 0 && o.kernel < 99) {
        const int h = o.kernel / 10, m = o.kernel % 10;
        if (h >= 1 && h <= 3) k.hint = h - 1;
        k.minb = m;
    }
    if (!k.vec4 && k.block == 512) k.block = 256;
    return k;
}

int next_pow2_shift(int v) // smallest s with (1<<s) >= v
{
    int s = 0;
    while ((1 << s) < v) s++;
    return s;
}

int check_device_available()
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
        cudaGetLastError();
        return fail(LBM_ENODEVICE, "no CUDA device available (%s): this library has no CPU path",
                    e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    }
    return LBM_OK;
}

// the reference's constants must be what the device code hard-wires
bool constants_ok()
{
    volatile float c_sq = 1.f / 3.f, w0 = 4.f / 9.f, w1 = 1.f / 9.f, w2 = 1.f / 36.f;
    volatile float c2 = 2.f * c_sq, c4 = 2.f * c_sq * c_sq;
    auto bits = [](float f) {
        uint32_t u;
        memcpy(&u, &f, 4);
        return u;
    };
    return bits(c_sq) == 0x3eaaaaabu && bits(w0) == 0x3ee38e39u && bits(w1) == 0x3de38e39u && bits(w2) == 0x3ce38e39u &&
           bits(c2) == 0x3f2aaaabu && bits(c4) == 0x3e638e3au && bits(1.f / c_sq) == 0x40400000u &&
           bits(1.f / c2) == 0x3fc00000u && bits(1.f / c4) == 0x408fffffu;
}

int validate_params(const lbm_param_t* p)
{
    if (!p) return fail(LBM_EINVAL, "params is NULL");
    if (p->nx < 1 || p->ny < 2) return fail(LBM_EINVAL, "grid too small: nx=%d ny=%d (need nx >= 1, ny >= 2)", p->nx, p->ny);
    if (static_cast<long long>(p->nx) + 32 > 0x7fffffffLL) return fail(LBM_EINVAL, "nx too large");
    return LBM_OK;
}

size_t plane_floats(const lbm_lattice* L, const Slab& s) { return static_cast<size_t>(s.rows) * L->pitch; }
bool uses_halo_cfg(const lbm_lattice* L) { return L->per_process || L->nslabs > 1; }

void slab_geometry(lbm_lattice* L, Slab& s, const KernelChoice& k)
{
    s.vec4 = k.vec4;
    s.block = k.block;
    s.nxv = k.vec4 ? L->p.nx / 4 : L->p.nx;
    int sh = next_pow2_shift(s.nxv);
    const int bsh = next_pow2_shift(s.block);
    if (sh > bsh) sh = bsh;
    s.tw_shift = sh;
    const int tw = 1 << sh, th = s.block >> sh;
    s.nbx = (s.nxv + tw - 1) / tw;
    s.ngroups = (s.rows + th - 1) / th;
    s.grid = static_cast<unsigned>(s.nbx) * static_cast<unsigned>(s.ngroups);
    s.grid_b = static_cast<unsigned>(s.nbx) * (s.rows >= 2 ? 2u : 1u);
    // step loop: tiles strided over the CTAs that are resident at once.  With several slabs every slab must have
    // a GPU of its own (the resident kernels exchange flags), and every rank must take the same decisions, so they
    // are based on the nominal slab height ceil(ny / slabs), not on this slab's own row count
    s.use_loop = false;
    const int total_slabs = L->per_process ? L->nranks : L->nslabs;
    const long long nominal_cells = static_cast<long long>(L->p.nx) * ((L->p.ny + total_slabs - 1) / total_slabs);
    // default: single slabs, and synchronous slabs of up to LOOP_HALO_MAX_CELLS cells on GPUs of their own that
    // step_ll_kernel does not take (the shipped 1024 x 1024 on 2 GPUs: 11.1 us per step against 13.6 from the step
    // graphs, profiles/r02_small_grids.md); kernel codes 200/201/204 ask for it explicitly
    if (k.loop && L->opt.use_graph && !L->interleaved &&
        ((L->opt.kernel >= 200 && L->opt.kernel <= 204) || (total_slabs == 1 && nominal_cells <= LOOP_MAX_CELLS) ||
         (total_slabs > 1 && L->opt.halo_mode == LBM_HALO_SYNC && L->opt.halo_lag == 0 && nominal_cells <= LOOP_HALO_MAX_CELLS))) {
        bool v4 = k.vec4 && nominal_cells >= LOOP_VEC4_CELLS;
        if (L->opt.kernel == 201) v4 = false;
        if (L->opt.kernel == 204) v4 = k.vec4;
        s.loop_vec = v4 ? 4 : 1;
        s.loop_block = v4 ? 256 : 128;
        s.loop_nxv = v4 ? L->p.nx / 4 : L->p.nx;
        int lsh = next_pow2_shift(s.loop_nxv);
        const int lbsh = next_pow2_shift(s.loop_block);
        if (lsh > lbsh) lsh = lbsh;
        s.loop_tw_shift = lsh;
        const int ltw = 1 << lsh, lth = s.loop_block >> lsh;
        s.loop_nbx = (s.loop_nxv + ltw - 1) / ltw;
        const long long ntiles = static_cast<long long>(s.loop_nbx) * ((s.rows + lth - 1) / lth);
        long long g = static_cast<long long>(L->loop_resident[v4 ? 1 : 0]) * L->sm_count;
        if (g > ntiles) g = ntiles;
        if (g >= 1 && ntiles <= 0x7fffffffLL && L->loop_kernel[v4 ? 1 : 0]) {
            s.loop_ntiles = static_cast<int>(ntiles);
            s.loop_nby = (s.rows + lth - 1) / lth;
            s.loop_grid = static_cast<unsigned>(g);
            s.use_loop = true;
        }
    }
    // the lattice resident in one cluster's shared memory: single slab, the shape fixed in common_setup()
    s.use_cluster = false;
    if (k.cluster && L->cluster_kernel && L->opt.use_graph && !L->interleaved && total_slabs == 1 && L->nslabs == 1 && s.cl_size > 0)
        s.use_cluster = true;
    // a row per CTA, all rows resident (cooperative launch).  With several slabs every slab needs a GPU of its own
    // (the resident kernels wait for one another's packets) and every rank must take the same decision: it is based on
    // the nominal slab height
    s.use_ll = false;
    {
        const bool ll_asked = (L->opt.kernel == 400 || L->opt.kernel == 401 || L->opt.kernel == 402 || L->opt.kernel == 404);
        const int nominal_rows = (L->p.ny + total_slabs - 1) / total_slabs;
        const bool halo_ok = !uses_halo_cfg(L) || (L->opt.halo_mode == LBM_HALO_SYNC && L->opt.halo_lag == 0);
        if (k.ll && L->opt.use_graph && !L->interleaved && halo_ok) {
            // one cell per thread for the smallest grids, two beyond (1024 x 256: 4.6 us per step against 6.1 with four
            // and 5.3 from step_loop_kernel); four on request only
            const int order[3] = {0, 2, 1};
            for (int oi = 0; oi < 3 && !s.use_ll; oi++) {
                const int v = order[oi];
                if (!L->ll_kernel[v]) continue;
                if (L->opt.kernel == 401 && v != 0) continue;
                if (L->opt.kernel == 404 && v != 1) continue;
                if (L->opt.kernel == 402 && v != 2) continue;
                if (v == 1 && L->opt.kernel != 404) continue;
                if (!ll_asked && nominal_cells > (v ? LL4_MAX_CELLS : LL_MAX_CELLS)) continue;
                if (static_cast<long long>(L->ll_resident[v]) * L->sm_count < nominal_rows) continue;
                const int block = (L->p.nx / LL_CPT[v] + 31) / 32 * 32;
                s.use_ll = true;
                s.ll_var = v;
                s.ll_block = block;
                s.ll_smem = (2 * 6 * static_cast<size_t>(block) + 2 * (block / 32) * 4) * sizeof(float);
            }
        }
        if (ll_asked && !s.use_ll && getenv("LBM_DEBUG")) fprintf(stderr, "[lbm] step_ll_kernel asked for but not applicable\n");
    }
    // a band of rows per CTA (cooperative launch, neighbour flags).  Full SMs only: per_sm CTAs on every SM, the rows
    // dealt out per SM.  With several slabs every slab needs a GPU of its own and every rank must take the same
    // decision: it is based on the nominal slab height
    s.use_band = false;
    {
        const bool asked = (L->opt.kernel >= 500 && L->opt.kernel < 600);
        const int nominal_rows = (L->p.ny + total_slabs - 1) / total_slabs;
        const bool halo_ok = !uses_halo_cfg(L) || (L->opt.halo_mode == LBM_HALO_SYNC && L->opt.halo_lag == 0);
        const long long max_cells = (total_slabs == 1) ? LOOP_MAX_CELLS : LOOP_HALO_MAX_CELLS;
        if (k.band && L->band_kernel && !s.use_ll && k.vec4 && L->opt.use_graph && !L->interleaved && halo_ok && L->sm_count > 0 &&
            (asked || nominal_cells <= max_cells)) {
            const int per_nominal = std::min(L->band_resident, nominal_rows / L->sm_count);
            // default: from three rows per SM (444 rows on 148 SMs); below, step_loop_kernel's finer tiles spread a row
            // over several CTAs and win (profiles/r02_small_grids.md).  As long as every row can have a CTA of its own
            // (rows <= resident CTAs) it gets one and the hardware's placement balances the SMs to within one row; with
            // more rows every SM is full and the rows are dealt out per SM (two rows in one CTA run one after the other:
            // 512 x 512 with 444 CTAs 6.4 us, with 512 CTAs 5.0 us)
            // and rows of more than one pass of the CTA: with a single pass per row and step (nx <= 512) the chain
            // wait - row - fence - flag of a one-row band is all latency (512 x 512: 6.3 us against 5.2)
            if (asked || (per_nominal >= BAND_MIN_PER_SM && L->p.nx > 4 * L->band_block)) {
                int per_sm = (static_cast<long long>(s.rows) <= static_cast<long long>(L->band_resident) * L->sm_count) ? 0 : L->band_resident;
                long long g = per_sm >= 1 ? static_cast<long long>(per_sm) * L->sm_count : s.rows;
                if (const char* t = getenv("LBM_BAND_GRID")) {
                    const long long v = atoll(t);
                    if (v >= 1 && v <= std::min<long long>(s.rows, static_cast<long long>(L->band_resident) * L->sm_count)) g = v, per_sm = 0;
                }
                if (getenv("LBM_BAND_SM") && atoi(getenv("LBM_BAND_SM")) == 0) per_sm = 0;
                s.use_band = true;
                s.band_grid = static_cast<unsigned>(g);
                s.band_per_sm = per_sm >= 1 ? per_sm : 0;
            }
        }
    }
    s.use_tma = !s.use_loop && k.tma && L->tma_kernel && s.rows >= 3;
    if (s.use_tma) {
        const int interior = s.rows - 2;
        s.tma_ntx = (L->p.nx + TMA_TX - 1) / TMA_TX;
        const long long nty = (interior + L->tma_ty - 1) / L->tma_ty;
        const long long ntiles = nty * s.tma_ntx;
        if (ntiles > 0x7fffffffLL) {
            s.use_tma = false;
        } else {
            s.tma_ntiles = static_cast<int>(ntiles);
            // persistent CTAs: as many as fit on the device at once, never more than tiles
            long long g = static_cast<long long>(L->tma_resident) * L->sm_count;
            if (g > ntiles) g = ntiles;
            s.tma_grid = static_cast<unsigned>(g);
            s.tma_dq = static_cast<int>(s.tma_grid / s.tma_ntx);
            s.tma_dr = static_cast<int>(s.tma_grid % s.tma_ntx);
        }
    }
    // pairs of steps (step2_kernel): every slab of the lattice must take the same decision, so it depends on the
    // nominal slab height only
    s.use_f2 = s.use_tma && k.f2 && L->f2_kernel && (L->p.ny / total_slabs) >= 8 && s.rows >= 8;
    if (s.use_f2) {
        s.f2_nsx = (L->p.nx + F2_CORE - 1) / F2_CORE;
        const int interior = s.rows - 4;
        const long long g_max = static_cast<long long>(L->f2_resident) * L->sm_count;
        // segment height: CTAs take units round-robin, so the launch lasts (units per CTA) x (iterations per unit,
        // each R rows; the unit's two extra intermediate rows and ~half an iteration of pipeline fill included);
        // pick the segment count that minimises it
        int best_nseg = 1;
        double best_cost = 1e300;
        if (const char* e = getenv("LBM_F2_SEG")) {
            const int h = atoi(e);
            if (h > 0) best_nseg = (interior + h - 1) / h, best_cost = -1.0;
        }
        if (best_cost > 0.0) {
            const int lo = std::max(1, interior / 640), hi = std::max(1, interior / 24);
            for (int nseg = lo; nseg <= hi; nseg++) {
                const int h = (interior + nseg - 1) / nseg;
                const long long units = static_cast<long long>((interior + h - 1) / h) * s.f2_nsx;
                const long long waves = (units + g_max - 1) / g_max;
                const int iters = (h + 2 + L->f2_iter_rows - 1) / L->f2_iter_rows;
                const double cost = static_cast<double>(waves) * (iters + 0.5);
                if (cost < best_cost - 1e-9) best_cost = cost, best_nseg = nseg;
            }
        }
        s.f2_seg_h = (interior + best_nseg - 1) / best_nseg;
        s.f2_nseg = (interior + s.f2_seg_h - 1) / s.f2_seg_h;
        const long long units = 2LL * s.f2_nsx + static_cast<long long>(s.f2_nseg) * s.f2_nsx;
        if (units > 0x7fffffffLL) {
            s.use_f2 = false;
        } else {
            s.f2_nunits = static_cast<int>(units);
            s.f2_grid = static_cast<unsigned>(std::min<long long>(g_max, units));
        }
    }
    if (getenv("LBM_DEBUG") && s.use_f2)
        fprintf(stderr, "[lbm] slab rows %d..%d dev %d: step2_kernel R %d, %d rows/stage, %d stages, %d resident/SM, %zu B smem, grid %u, %d strips x %d segments of %d rows (+ %d boundary units)\n",
                s.row0, s.row1 - 1, s.device, L->f2_r, L->f2_srows, L->f2_stages, L->f2_resident, L->f2_smem, s.f2_grid, s.f2_nsx, s.f2_nseg,
                s.f2_seg_h, 2 * s.f2_nsx);
    if (getenv("LBM_DEBUG"))
        fprintf(stderr, "[lbm] slab rows %d..%d dev %d: %s, grid %u x %d thr, boundary grid %u; loop %d (%d cell/thread, grid %u, %d tiles); tma %d (TY %d, %d stages, %d CTA/SM wanted, %d resident, %zu B smem, grid %u, %d tiles)\n",
                s.row0, s.row1 - 1, s.device, s.vec4 ? "vec4" : "scalar", s.grid, s.block, s.grid_b, s.use_loop ? 1 : 0, s.loop_vec, s.loop_grid,
                s.loop_ntiles, s.use_tma ? 1 : 0, L->tma_ty,
                L->tma_stages, L->tma_minb, L->tma_resident, L->tma_smem, s.tma_grid, s.tma_ntiles);
    if (getenv("LBM_DEBUG"))
        fprintf(stderr, "[lbm] slab rows %d..%d dev %d: single-launch kernels: ll %d (variant %d, %d threads), band %d (grid %u, %d per SM), cluster %d\n",
                s.row0, s.row1 - 1, s.device, s.use_ll ? 1 : 0, s.ll_var, s.ll_block, s.use_band ? 1 : 0, s.band_grid, s.band_per_sm,
                s.use_cluster ? 1 : 0);
    // spread the per-step global atomics over several addresses when there are many CTAs
    const unsigned ctas = s.use_tma ? std::max(s.tma_grid + s.grid_b, s.f2_grid) : s.grid;
    int slots = 1;
    while (slots < 64 && static_cast<unsigned>(slots) * 1024u < ctas) slots <<= 1;
    s.nslots = slots;
    if (s.use_ll || s.use_band) s.nslots = LL_SLOTS; // one RED per CTA, step and word: spread them
}

// the lattice as a 3-D tensor (x, y, plane) for the TMA unit; boxes are (TMA_TX | TMA_TXW) x box_rows x 1
int encode_maps(lbm_lattice* L, Slab& s, int box_rows, CUtensorMap narrow[2], CUtensorMap wide[2])
{
    static PFN_cuTensorMapEncodeTiled encode = nullptr;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        CU(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
        if (q != cudaDriverEntryPointSuccess || !fn) return fail(LBM_ECUDA, "the driver does not export cuTensorMapEncodeTiled");
        encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled>(fn);
    }
    const size_t pf = plane_floats(L, s);
    for (int i = 0; i < 2; i++) {
        const cuuint64_t gdim[3] = {static_cast<cuuint64_t>(L->p.nx), static_cast<cuuint64_t>(s.rows), Q};
        const cuuint64_t gstride[2] = {static_cast<cuuint64_t>(L->pitch) * sizeof(float), static_cast<cuuint64_t>(pf) * sizeof(float)};
        const cuuint32_t estride[3] = {1, 1, 1};
        CUtensorMapL2promotion promo = CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
        if (const char* e = getenv("LBM_TMA_L2PROMO")) { // tuning: 0 none, 64, 128, 256 (default)
            const int v = atoi(e);
            promo = v == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE
                           : (v == 64 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B : (v == 128 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : promo));
        }
        for (int w = 0; w < 2; w++) {
            const cuuint32_t box[3] = {static_cast<cuuint32_t>(w ? TMA_TXW : TMA_TX), static_cast<cuuint32_t>(box_rows), 1};
            const CUresult r = encode(w ? &wide[i] : &narrow[i], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, s.lat[i], gdim, gstride, box,
                                      estride, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                      promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) return fail(LBM_ECUDA, "cuTensorMapEncodeTiled failed with CUresult %d", static_cast<int>(r));
        }
    }
    return LBM_OK;
}

int make_tensor_maps(lbm_lattice* L, Slab& s)
{
    int rc = encode_maps(L, s, L->tma_ty, s.tmap, s.tmapw);
    if (!rc && s.use_f2) rc = encode_maps(L, s, L->f2_srows, s.f2map, s.f2mapw);
    return rc;
}

// LBM_DEBUG=1: wall-clock phases of lattice creation on stderr
struct DebugTimer {
    bool on;
    double t0;
    static double now()
    {
        timespec ts;
        clock_gettime(CLOCK_MONOTONIC, &ts);
        return ts.tv_sec + ts.tv_nsec * 1e-9;
    }
    DebugTimer() : on(getenv("LBM_DEBUG") != nullptr), t0(now()) {}
    void lap(const char* what)
    {
        if (!on) return;
        const double t = now();
        fprintf(stderr, "[lbm] %-28s %8.3f ms\n", what, (t - t0) * 1e3);
        t0 = t;
    }
};

// obst_rows_host: the slab's rows of the reference's int map (packed == false), or of the packed bit map
// ([rows][opitch] words, bit x%32 of word x/32)
int alloc_slab(lbm_lattice* L, Slab& s, const void* obst_rows_host, bool packed)
{
    DebugTimer dbg;
    CU(cudaSetDevice(s.device));
    CU(cudaStreamCreateWithFlags(&s.own_stream, cudaStreamNonBlocking));
    s.stream = s.own_stream;
    CU(cudaEventCreate(&s.ev0));
    CU(cudaEventCreate(&s.ev1));
    const size_t pf = plane_floats(L, s);
    dbg.lap("stream + events");
    CU(cudaMalloc(&s.lat[0], 2 * pf * Q * sizeof(float))); // both lattices in one allocation
    s.lat[1] = s.lat[0] + pf * Q;
    dbg.lap("cudaMalloc lattices");
    CU(cudaMalloc(&s.obst, static_cast<size_t>(s.rows) * L->opitch * sizeof(uint32_t)));
    CU(cudaMemsetAsync(s.obst, 0, static_cast<size_t>(s.rows) * L->opitch * sizeof(uint32_t), s.stream));
    CU(cudaMalloc(&s.fluid_dev, sizeof(unsigned long long)));
    CU(cudaMemsetAsync(s.fluid_dev, 0, sizeof(unsigned long long), s.stream));
    CU(cudaMalloc(&s.ctrl, 4 * sizeof(int)));
    CU(cudaMemsetAsync(s.ctrl, 0, 4 * sizeof(int), s.stream));
    CU(cudaMalloc(&s.error, sizeof(int)));
    CU(cudaMemsetAsync(s.error, 0, sizeof(int), s.stream));
    CU(cudaMalloc(&s.state_sums, (SUM_WORDS + 1) * sizeof(unsigned long long)));
    CU(cudaMalloc(&s.loop_barrier, 128));

    // obstacle map first: the int rows (4 B per cell) are staged in the memory of the second lattice
    // (36 B per cell, not initialised yet) -- no temporary allocation, no cudaFree on the creation path --
    // packed to bits on the device, and the fluid cells counted
    {
        const size_t n = static_cast<size_t>(s.rows) * L->p.nx;
        dbg.lap("small allocs");
        if (packed) {
            // already one bit per cell in the device's own layout (opitch == ceil(nx / 32) words per row)
            const size_t nw = static_cast<size_t>(s.rows) * L->opitch;
            CU(cudaMemcpyAsync(s.obst, obst_rows_host, nw * sizeof(uint32_t), cudaMemcpyHostToDevice, s.stream));
            const size_t blocks = std::min<size_t>((nw + 255) / 256, 148 * 8);
            sanitize_obstacle_bits_kernel<<<static_cast<unsigned>(blocks), 256, 0, s.stream>>>(s.obst, L->p.nx, s.rows, L->opitch, s.fluid_dev);
        } else {
            int* staging = reinterpret_cast<int*>(s.lat[1]);
            CU(cudaMemcpyAsync(staging, obst_rows_host, n * sizeof(int), cudaMemcpyHostToDevice, s.stream));
            const size_t warps = static_cast<size_t>(s.rows) * ((L->p.nx + 31) / 32);
            const size_t blocks = std::min<size_t>((warps * 32 + 255) / 256, 148 * 16);
            pack_obstacles_kernel<<<static_cast<unsigned>(blocks), 256, 0, s.stream>>>(staging, s.obst, L->p.nx, s.rows,
                                                                                        L->opitch, s.fluid_dev);
        }
        L->launches++;
        CU(cudaGetLastError());
        // initial state: every cell, obstacles included (SerialCode:551-567); both lattices
        const float w[Q] = {L->w0, L->w1, L->w1, L->w1, L->w1, L->w2, L->w2, L->w2, L->w2};
        for (int i = 0; i < 2; i++)
            for (int k = 0; k < Q; k++) {
                fill_kernel<<<592, 256, 0, s.stream>>>(s.lat[i] + k * pf, pf, w[k]);
                L->launches++;
            }
        CU(cudaGetLastError());
        unsigned long long fl = 0;
        CU(cudaMemcpyAsync(&fl, s.fluid_dev, sizeof fl, cudaMemcpyDeviceToHost, s.stream));
        CU(cudaStreamSynchronize(s.stream));
        dbg.lap("H2D + pack + fill (sync)");
        s.fluid = static_cast<long long>(fl);
    }
    return LBM_OK;
}

// floats in one halo ring of this lattice
size_t ring_floats_of(const lbm_lattice* L) { return static_cast<size_t>(L->ring) * RING_ENTRIES * L->pitch; }

int alloc_halo(lbm_lattice* L, Slab& s)
{
    CU(cudaSetDevice(s.device));
    const size_t ring_floats = ring_floats_of(L);
    const size_t obst_bytes = 2 * static_cast<size_t>(L->opitch) * sizeof(uint32_t);
    const size_t ll_off = (2 * ring_floats * sizeof(float) + 2 * 128 + obst_bytes + 15) / 16 * 16;
    const size_t ll_bytes = 2 * static_cast<size_t>(L->pitch) * sizeof(uint4); // one boundary: [2 parities][pitch] packets
    s.halo_bytes = ll_off + 2 * ll_bytes;
    CU(cudaMalloc(&s.halo_block, s.halo_bytes));
    // step_ll_kernel's packet areas: zeroed here, before anybody can know the address (flag 0 is never waited for)
    s.ll_recv_s = reinterpret_cast<uint4*>(s.halo_block + ll_off);
    s.ll_recv_n = reinterpret_cast<uint4*>(s.halo_block + ll_off + ll_bytes);
    CU(cudaMemsetAsync(s.ll_recv_s, 0, 2 * ll_bytes, s.stream));
    s.ring_s = reinterpret_cast<float*>(s.halo_block);
    s.ring_n = s.ring_s + ring_floats;
    s.flag_s = reinterpret_cast<unsigned long long*>(s.halo_block + 2 * ring_floats * sizeof(float));
    s.flag_n = s.flag_s + 16;
    s.edge_obst = reinterpret_cast<uint32_t*>(s.halo_block + 2 * ring_floats * sizeof(float) + 256);
    CU(cudaMemsetAsync(s.flag_s, 0, 256, s.stream));
    // obstacle bits of my first and last row, for the neighbours (they recompute those rows' intermediate step)
    CU(cudaMemcpyAsync(s.edge_obst, s.obst, L->opitch * sizeof(uint32_t), cudaMemcpyDeviceToDevice, s.stream));
    CU(cudaMemcpyAsync(s.edge_obst + L->opitch, s.obst + static_cast<size_t>(s.rows - 1) * L->opitch, L->opitch * sizeof(uint32_t),
                       cudaMemcpyDeviceToDevice, s.stream));
    CU(cudaMalloc(&s.obst_halo, obst_bytes));
    CU(cudaMemsetAsync(s.obst_halo, 0, obst_bytes, s.stream));
    CU(cudaMalloc(&s.arrive, 2 * sizeof(unsigned long long)));
    CU(cudaMemsetAsync(s.arrive, 0, 2 * sizeof(unsigned long long), s.stream));
    // both rings start in the uniform initial state (MPI_Testall_OptimizedVersion/d2q9-bgk.c:784-824): entries
    // 0..2 hold planes 0,1,3; entries 3..5 and 6..8 the crossing planes (2,5,6 south ring / 4,7,8 north ring)
    const float w[RING_ENTRIES] = {L->w0, L->w1, L->w1, L->w1, L->w2, L->w2, L->w1, L->w2, L->w2};
    for (int slot = 0; slot < L->ring; slot++)
        for (int e = 0; e < RING_ENTRIES; e++) {
            const size_t off = (static_cast<size_t>(slot) * RING_ENTRIES + e) * L->pitch;
            fill_kernel<<<64, 256, 0, s.stream>>>(s.ring_s + off, L->pitch, w[e]);
            fill_kernel<<<64, 256, 0, s.stream>>>(s.ring_n + off, L->pitch, w[e]);
            L->launches += 2;
        }
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(s.stream));
    return LBM_OK;
}

void destroy_graphs(Slab& s)
{
    for (int i = 0; i < 2; i++)
        if (s.graph[i]) {
            cudaGraphExecDestroy(s.graph[i]);
            s.graph[i] = nullptr;
        }
}

HaloCfg make_halo_cfg(const lbm_lattice* L, const Slab& s)
{
    HaloCfg h;
    memset(&h, 0, sizeof h);
    h.on = (L->per_process || L->nslabs > 1) ? 1 : 0;
    if (h.on) {
        h.hs.recv_ring = s.ring_s, h.hs.send_ring = s.peer_ring_s, h.hs.wait = s.flag_s, h.hs.signal = s.peer_flag_s;
        h.hn.recv_ring = s.ring_n, h.hn.send_ring = s.peer_ring_n, h.hn.wait = s.flag_n, h.hn.signal = s.peer_flag_n;
    }
    h.wait = (L->opt.halo_mode == LBM_HALO_SYNC) ? 1 : 0;
    h.ring = L->ring;
    h.lag = (L->opt.halo_mode == LBM_HALO_SYNC) ? L->opt.halo_lag : 0;
    h.arrive = s.arrive;
    h.slot_stride = static_cast<unsigned long long>(RING_ENTRIES) * L->pitch;
    h.timeout_ns = L->timeout_ns;
    h.error = s.error;
    return h;
}

StepArgs make_args(const lbm_lattice* L, const Slab& s, int src, int step_offset, int epoch_offset)
{
    StepArgs a;
    memset(&a, 0, sizeof a);
    a.mode = s.use_tma ? 1 : 0;
    a.pf = plane_floats(L, s);
    a.in = s.lat[src];
    a.out = s.lat[src ^ 1];
    a.h = make_halo_cfg(L, s);
    a.obst = s.obst;
    a.ctrl = s.ctrl;
    a.sums_ref = s.sums_ref;
    a.nslots = s.nslots;
    a.step_offset = step_offset;
    a.epoch_offset = epoch_offset;
    a.nx = L->p.nx, a.nxv = s.nxv, a.rows = s.rows, a.pitch = L->pitch, a.opitch = L->opitch;
    a.tw_shift = s.tw_shift, a.nbx = s.nbx, a.ngroups = s.ngroups;
    a.accel_row = s.accel_row;
    a.omega = L->p.omega, a.w1a = L->w1a, a.w2a = L->w2a;
    return a;
}

TmaArgs make_tma_args(const lbm_lattice* L, const Slab& s, int src, int step_offset)
{
    TmaArgs a;
    memset(&a, 0, sizeof a);
    const size_t pf = plane_floats(L, s);
    const float* in = s.lat[src];
    for (int k = 0; k < Q; k++) a.out[k] = s.lat[src ^ 1] + k * pf;
    a.west[0] = in + 1 * pf, a.west[1] = in + 5 * pf, a.west[2] = in + 8 * pf;
    a.east[0] = in + 3 * pf, a.east[1] = in + 6 * pf, a.east[2] = in + 7 * pf;
    a.obst = s.obst;
    a.ctrl = s.ctrl;
    a.sums_ref = s.sums_ref;
    a.nslots = s.nslots;
    a.step_offset = step_offset;
    a.nx = L->p.nx, a.pitch = L->pitch, a.opitch = L->opitch;
    a.y_first = 1, a.y_end = s.rows - 1;
    a.ntx = s.tma_ntx, a.ntiles = s.tma_ntiles, a.dq = s.tma_dq, a.dr = s.tma_dr;
    a.accel_row = s.accel_row;
    a.omega = L->p.omega, a.w1a = L->w1a, a.w2a = L->w2a;
    return a;
}

Fused2Args make_f2_args(const lbm_lattice* L, const Slab& s, int src, int step_offset, int epoch_offset)
{
    Fused2Args a;
    memset(&a, 0, sizeof a);
    a.in = s.lat[src];
    a.out = s.lat[src ^ 1];
    a.pf = plane_floats(L, s);
    a.h = make_halo_cfg(L, s);
    a.obst = s.obst;
    a.obst_halo = s.obst_halo;
    a.ctrl = s.ctrl;
    a.sums_ref = s.sums_ref;
    a.nslots = s.nslots;
    a.step_offset = step_offset, a.epoch_offset = epoch_offset;
    a.nx = L->p.nx, a.rows = s.rows, a.pitch = L->pitch, a.opitch = L->opitch;
    a.nsx = s.f2_nsx, a.seg_h = s.f2_seg_h, a.nseg = s.f2_nseg, a.nunits = s.f2_nunits;
    a.accel_row = s.accel_row;
    a.omega = L->p.omega, a.w1a = L->w1a, a.w2a = L->w2a;
    return a;
}

// one timestep of one slab on its stream, outside a graph: the boundary rows (or every row), then
// the interior rows
int launch_step(lbm_lattice* L, Slab& s, int src, int step_offset, int epoch_offset)
{
    const StepArgs a = make_args(L, s, src, step_offset, epoch_offset);
    L->kernel<<<s.use_tma ? s.grid_b : s.grid, s.block, 0, s.stream>>>(a);
    L->launches++;
    if (s.use_tma) {
        const TmaArgs t = make_tma_args(L, s, src, step_offset);
        L->tma_kernel<<<s.tma_grid, 32 * L->tma_ty + 32, L->tma_smem, s.stream>>>(s.tmap[src], s.tmapw[src], t);
        L->launches++;
    }
    CU(cudaGetLastError());
    return LBM_OK;
}

// two timesteps of one slab in one launch of step2_kernel
int launch_pair(lbm_lattice* L, Slab& s, int src, int step_offset, int epoch_offset)
{
    const Fused2Args a = make_f2_args(L, s, src, step_offset, epoch_offset);
    L->f2_kernel<<<s.f2_grid, 32 * L->f2_r, L->f2_smem, s.stream>>>(s.f2map[src], s.f2mapw[src], a);
    L->launches++;
    CU(cudaGetLastError());
    return LBM_OK;
}

// timesteps and halo epochs one step graph covers
int graph_epochs(const Slab& s) { return s.use_f2 ? GRAPH_STEPS / 2 : GRAPH_STEPS; }

int build_graph(lbm_lattice* L, Slab& s, int parity)
{
    CU(cudaSetDevice(s.device));
    cudaGraph_t g;
    CU(cudaGraphCreate(&g, 0));
    std::vector<cudaGraphNode_t> prev;
    if (s.use_f2) {
        // GRAPH_STEPS / 2 launches of step2_kernel, two timesteps each, one after the other
        for (int j = 0; j < GRAPH_STEPS / 2; j++) {
            const int src = (parity + j) & 1;
            Fused2Args a = make_f2_args(L, s, src, 2 * j, j);
            void* kp[3] = {&s.f2map[src], &s.f2mapw[src], &a};
            cudaKernelNodeParams np = {};
            np.func = reinterpret_cast<void*>(L->f2_kernel);
            np.gridDim = dim3(s.f2_grid, 1, 1);
            np.blockDim = dim3(32 * L->f2_r, 1, 1);
            np.sharedMemBytes = static_cast<unsigned>(L->f2_smem);
            np.kernelParams = kp;
            cudaGraphNode_t node;
            CU(cudaGraphAddKernelNode(&node, g, prev.data(), prev.size(), &np));
            prev.assign(1, node);
        }
    }
    // single steps -- per step: the boundary-row kernel (or the all-row kernel) and, beside it, the interior-row
    // TMA kernel; both depend on both kernels of the previous step
    for (int j = 0; j < GRAPH_STEPS && !s.use_f2; j++) {
        std::vector<cudaGraphNode_t> cur;
        const int src = (parity + j) & 1;
        {
            StepArgs a = make_args(L, s, src, j, j);
            void* kp[1] = {&a};
            cudaKernelNodeParams np = {};
            np.func = reinterpret_cast<void*>(L->kernel);
            np.gridDim = dim3(s.use_tma ? s.grid_b : s.grid, 1, 1);
            np.blockDim = dim3(s.block, 1, 1);
            np.kernelParams = kp;
            cudaGraphNode_t node;
            CU(cudaGraphAddKernelNode(&node, g, prev.data(), prev.size(), &np));
            if (s.use_tma) {
                // the few boundary CTAs must not queue behind the persistent interior kernel (which fills
                // every SM): their halo rows and flags are what the neighbour GPU's next step waits for
                cudaKernelNodeAttrValue v;
                memset(&v, 0, sizeof v);
                v.priority = L->prio_high;
                CU(cudaGraphKernelNodeSetAttribute(node, cudaKernelNodeAttributePriority, &v));
            }
            cur.push_back(node);
        }
        if (s.use_tma) {
            TmaArgs t = make_tma_args(L, s, src, j);
            void* kp[3] = {&s.tmap[src], &s.tmapw[src], &t};
            cudaKernelNodeParams np = {};
            np.func = reinterpret_cast<void*>(L->tma_kernel);
            np.gridDim = dim3(s.tma_grid, 1, 1);
            np.blockDim = dim3(32 * L->tma_ty + 32, 1, 1);
            np.sharedMemBytes = static_cast<unsigned>(L->tma_smem);
            np.kernelParams = kp;
            cudaGraphNode_t node;
            CU(cudaGraphAddKernelNode(&node, g, prev.data(), prev.size(), &np));
            cur.push_back(node);
        }
        prev = cur;
    }
    {
        int* ctrl = s.ctrl;
        int by = GRAPH_STEPS, by_epochs = graph_epochs(s);
        void* kp[3] = {&ctrl, &by, &by_epochs};
        cudaKernelNodeParams np = {};
        np.func = reinterpret_cast<void*>(advance_ctrl_kernel);
        np.gridDim = dim3(1, 1, 1);
        np.blockDim = dim3(32, 1, 1);
        np.kernelParams = kp;
        cudaGraphNode_t node;
        CU(cudaGraphAddKernelNode(&node, g, prev.data(), prev.size(), &np));
    }
    cudaError_t e = cudaGraphInstantiate(&s.graph[parity], g, 0);
    cudaGraphDestroy(g);
    if (e != cudaSuccess) return fail(LBM_ECUDA, "cudaGraphInstantiate failed: %s", cudaGetErrorString(e));
    return LBM_OK;
}

int ensure_sums(lbm_lattice* L, Slab& s, size_t steps)
{
    (void)L;
    CU(cudaSetDevice(s.device));
    if (!s.sums_ref) CU(cudaMalloc(&s.sums_ref, sizeof(unsigned long long*)));
    if (steps > s.sums_steps) {
        if (s.sums) {
            CU(cudaStreamSynchronize(s.stream));
            CU(cudaFree(s.sums));
        }
        s.sums = nullptr;
        size_t cap = s.sums_steps ? s.sums_steps : 8192;
        while (cap < steps) cap *= 2;
        CU(cudaMalloc(&s.sums, cap * s.nslots * SUM_WORDS * sizeof(unsigned long long)));
        s.sums_steps = cap;
        // the step graphs read the base through sums_ref: nothing to rebuild
        CU(cudaMemcpyAsync(s.sums_ref, &s.sums, sizeof(unsigned long long*), cudaMemcpyHostToDevice, s.stream));
        CU(cudaStreamSynchronize(s.stream)); // &s.sums must stay valid until the copy has happened
    }
    CU(cudaMemsetAsync(s.sums, 0, steps * s.nslots * SUM_WORDS * sizeof(unsigned long long), s.stream));
    return LBM_OK;
}

bool uses_halo(const lbm_lattice* L) { return L->per_process || L->nslabs > 1; }

// Build, instantiate and upload the step graphs of both lattice parities when the lattice is created /
// connected, not inside the first lbm_run that needs them: cudaGraphInstantiate was measured to take
// anything from 1 ms to 100 ms on a busy host, and a run that starts with the device idle behind it
// loses that much (profiles/r01_variance.md).
int prepare_graphs(lbm_lattice* L)
{
    if (!L->opt.use_graph || L->interleaved) return LBM_OK;
    for (int i = 0; i < L->nslabs; i++) {
        Slab& s = L->slabs[i];
        if (s.use_loop) continue;
        int rc = ensure_sums(L, s, 8192);
        if (rc) return rc;
        for (int parity = 0; parity < 2; parity++) {
            if (s.graph[parity]) continue;
            rc = build_graph(L, s, parity);
            if (rc) return rc;
            CU(cudaGraphUpload(s.graph[parity], s.stream));
        }
        CU(cudaStreamSynchronize(s.stream));
    }
    return LBM_OK;
}

int common_setup(lbm_lattice* L, const lbm_param_t* params, const lbm_options_t* opt)
{
    L->p = *params;
    if (opt)
        L->opt = *opt;
    else
        lbm_default_options(&L->opt);
    if (L->opt.arith != LBM_ARITH_STRICT && L->opt.arith != LBM_ARITH_FAST) return fail(LBM_EINVAL, "bad arith option %d", L->opt.arith);
    if (L->opt.halo_mode != LBM_HALO_SYNC && L->opt.halo_mode != LBM_HALO_ASYNC) return fail(LBM_EINVAL, "bad halo_mode %d", L->opt.halo_mode);
    if (L->opt.halo_lag < 0 || (L->opt.halo_lag & 1) || L->opt.halo_lag > 64)
        return fail(LBM_EINVAL, "halo_lag must be even and in [0, 64], got %d", L->opt.halo_lag);
    if (L->opt.halo_mode == LBM_HALO_ASYNC && L->opt.halo_lag != 0)
        return fail(LBM_EINVAL, "halo_lag is only meaningful with LBM_HALO_SYNC");
    L->pitch = (params->nx + 31) / 32 * 32;
    L->opitch = L->pitch / 32;
    // Ring depth.  Slots are labelled by d = (producer step + 1): the consumer's step s reads d = s - lag,
    // the producer's step t writes d = t + 1.  In sync mode the consumer of step s waits until the
    // producer has finished step s-lag-1, and symmetrically the producer (a consumer in the other
    // direction) cannot start step t before its neighbour finished step t-lag-1.  So while the producer
    // writes d = t+1 the neighbour is reading some d' in [t-2*lag, t+1], and d' = t+1 is only read after
    // the write was signalled: no slot is overwritten while live iff ring > (t+1) - (t-2*lag), i.e.
    // ring >= 2*lag + 2.  (Async mode: lag = 0, ring = 2 -- the A/B parity buffers of the reference,
    // SURVEY.md App. C-3; races there are the mode's definition.)
    L->ring = 2 * L->opt.halo_lag + 2;
    // exactly the reference's expressions (float arithmetic, left to right)
    L->w0 = params->density * 4.f / 9.f;  // SerialCode:546
    L->w1 = params->density / 9.f;        // :547
    L->w2 = params->density / 36.f;       // :548
    L->w1a = params->density * params->accel / 9.f;  // :222
    L->w2a = params->density * params->accel / 36.f; // :223
    if (const char* t = getenv("LBM_HALO_TIMEOUT_MS")) {
        const long long ms = atoll(t);
        if (ms > 0) L->timeout_ns = static_cast<unsigned long long>(ms) * 1000000ull;
    }
    {
        int lo = 0, hi = 0;
        if (cudaDeviceGetStreamPriorityRange(&lo, &hi) == cudaSuccess) L->prio_high = hi;
        cudaGetLastError();
    }
    const KernelChoice k = choose_kernel(L->opt, params->nx);
    const bool strict = L->opt.arith == LBM_ARITH_STRICT;
    if (k.vec4)
        L->kernel = strict ? vec4_by_hint<true>(k.hint, k.block, k.minb) : vec4_by_hint<false>(k.hint, k.block, k.minb);
    else
        L->kernel = scalar_kernel(strict, k.block);
    if (k.loop && !L->interleaved) {
        if (uses_halo_cfg(L)) {
            L->loop_kernel[0] = strict ? step_loop_kernel<true, 128, 1, true> : step_loop_kernel<false, 128, 1, true>;
            L->loop_kernel[1] = strict ? step_loop_kernel<true, 256, 4, true> : step_loop_kernel<false, 256, 4, true>;
        } else {
            L->loop_kernel[0] = strict ? step_loop_kernel<true, 128, 1, false> : step_loop_kernel<false, 128, 1, false>;
            L->loop_kernel[1] = strict ? step_loop_kernel<true, 256, 4, false> : step_loop_kernel<false, 256, 4, false>;
        }
        CU(cudaSetDevice(L->slabs[0].device));
        int sms = 0, coop = 0;
        CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, L->slabs[0].device));
        CU(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, L->slabs[0].device));
        L->sm_count = sms;
        for (int v = 0; v < 2; v++) {
            int resident = 0;
            CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, reinterpret_cast<const void*>(L->loop_kernel[v]),
                                                             v ? 256 : 128, 0));
            L->loop_resident[v] = coop ? resident : 0;
            if (!coop) L->loop_kernel[v] = nullptr;
        }
    }
    if (k.band && k.vec4 && !L->interleaved) {
        const bool halo = uses_halo_cfg(L);
        void (*fn)(BandArgs) = nullptr;
        const int bb = k.band_block, bm = k.band_minb;
#define LBM_BAND_CASE(B_, M_, V_)                                                                                         \
    if (bb == B_ && bm == M_)                                                                                             \
        fn = halo ? (strict ? step_band_kernel<true, B_, M_, V_, true> : step_band_kernel<false, B_, M_, V_, true>)      \
                  : (strict ? step_band_kernel<true, B_, M_, V_, false> : step_band_kernel<false, B_, M_, V_, false>);
        LBM_BAND_CASE(256, 2, false)
        LBM_BAND_CASE(128, 4, false)
        LBM_BAND_CASE(256, 1, true)
        LBM_BAND_CASE(512, 1, false)
#undef LBM_BAND_CASE
        if (!fn) return fail(LBM_EINVAL, "no step_band_kernel variant with %d threads and %d CTAs per SM", bb, bm);
        bool ok = true;
        for (int i = 0; i < L->nslabs; i++) {
            CU(cudaSetDevice(L->slabs[i].device));
            int sms = 0, coop = 0, resident = 0;
            CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, L->slabs[i].device));
            CU(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, L->slabs[i].device));
            if (i == 0 || sms < L->sm_count) L->sm_count = sms;
            if (coop) CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, reinterpret_cast<const void*>(fn), bb, 0));
            if (i == 0 || resident < L->band_resident) L->band_resident = resident;
            ok = ok && coop && resident >= 1;
        }
        if (ok) L->band_kernel = fn, L->band_block = bb;
    }
    if (k.ll && !L->interleaved) {
        const bool halo = uses_halo_cfg(L);
        void (*fn[3])(LLArgs);
        if (halo) {
            fn[0] = strict ? step_ll_kernel<true, 1, false, 1024, 1, true> : step_ll_kernel<false, 1, false, 1024, 1, true>;
            fn[1] = strict ? step_ll_kernel<true, 4, false, 256, 2, true> : step_ll_kernel<false, 4, false, 256, 2, true>;
            fn[2] = strict ? step_ll_kernel<true, 2, false, 512, 2, true> : step_ll_kernel<false, 2, false, 512, 2, true>;
        } else {
            fn[0] = strict ? step_ll_kernel<true, 1, false, 1024, 1, false> : step_ll_kernel<false, 1, false, 1024, 1, false>;
            fn[1] = strict ? step_ll_kernel<true, 4, false, 256, 2, false> : step_ll_kernel<false, 4, false, 256, 2, false>;
            fn[2] = strict ? step_ll_kernel<true, 2, false, 512, 2, false> : step_ll_kernel<false, 2, false, 512, 2, false>;
        }
        for (int i = 0; i < L->nslabs; i++) {
            CU(cudaSetDevice(L->slabs[i].device));
            int sms = 0, coop = 0;
            CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, L->slabs[i].device));
            CU(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, L->slabs[i].device));
            if (i == 0 || sms < L->sm_count) L->sm_count = sms;
            for (int v = 0; v < 3; v++) {
                const int cpt = LL_CPT[v];
                int resident = 0;
                if (coop && params->nx % cpt == 0 && params->nx / cpt <= LL_MAXT[v]) {
                    const int block = (params->nx / cpt + 31) / 32 * 32;
                    const size_t smem = (2 * 6 * static_cast<size_t>(block) + 2 * (block / 32) * 4) * sizeof(float);
                    CU(cudaFuncSetAttribute(reinterpret_cast<const void*>(fn[v]), cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
                    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, reinterpret_cast<const void*>(fn[v]), block, smem));
                }
                if (i == 0 || resident < L->ll_resident[v]) L->ll_resident[v] = resident;
            }
        }
        for (int v = 0; v < 3; v++) L->ll_kernel[v] = L->ll_resident[v] > 0 ? fn[v] : nullptr;
    }
    if (k.cluster && !L->interleaved && !uses_halo_cfg(L) && L->nslabs == 1) {
        // one cluster of C CTAs (a power of two <= 16 and <= rows), rows dealt out in blocks of rpc = ceil(rows / C)
        Slab& s0 = L->slabs[0];
        const int rows = params->ny, nx = params->nx;
        const long long cells = static_cast<long long>(nx) * rows;
        const bool asked = (L->opt.kernel >= 3000 && L->opt.kernel < 4000);
        CU(cudaSetDevice(s0.device));
        int cl_ok = 0;
        CU(cudaDeviceGetAttribute(&cl_ok, cudaDevAttrClusterLaunch, s0.device));
        for (int C = CLUSTER_MAX_CTAS; cl_ok && C >= 1 && !L->cluster_kernel && (asked || cells <= CLUSTER_MAX_CELLS); C >>= 1) {
            if (C > rows) continue;
            const int rpc = (rows + C - 1) / C;
            int cpt = k.cl_cpt, vert = k.cl_vert, maxt = k.cl_maxt;
            if (cpt == 0) {
                // few fat threads: four cells per thread whenever the row length allows it
                cpt = (nx % 4 == 0) ? 4 : ((nx % 2 == 0) ? 2 : 1);
                const long long t = static_cast<long long>(rpc) * (nx / cpt);
                if (cpt == 4) maxt = (t <= 256) ? 256 : ((t <= 512) ? 512 : 1024), vert = (maxt == 256);
                else if (cpt == 2) maxt = (t <= 512) ? 512 : 1024, vert = 0;
                else maxt = 1024, vert = 0;
            }
            if (cpt < 1 || nx % cpt) {
                if (asked && k.cl_cpt) return fail(LBM_EINVAL, "step_cluster_kernel: nx = %d is not a multiple of %d cells per thread", nx, cpt);
                break;
            }
            const long long threads = static_cast<long long>(rpc) * (nx / cpt);
            const size_t smem = 2 * static_cast<size_t>(Q) * rpc * nx * sizeof(float); // two copies
            if (threads > maxt || smem > 220 * 1024) {
                if (C == CLUSTER_MAX_CTAS && asked)
                    return fail(LBM_EINVAL, "step_cluster_kernel: %d x %d does not fit (%lld threads per CTA of %d allowed, %zu B of shared memory)",
                                nx, rows, threads, maxt, smem);
                break; // fewer CTAs only make it worse
            }
            ClusterChoice c;
            const bool ok = strict ? cluster_by_shape<true>(cpt, vert, maxt, &c) : cluster_by_shape<false>(cpt, vert, maxt, &c);
            if (!ok) return fail(LBM_EINVAL, "no step_cluster_kernel variant with %d cells per thread, vert %d, %d threads", cpt, vert, maxt);
            const void* fn = reinterpret_cast<const void*>(c.fn);
            CU(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
            if (C > 8) CU(cudaFuncSetAttribute(fn, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
            cudaLaunchConfig_t cfg;
            memset(&cfg, 0, sizeof cfg);
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = static_cast<unsigned>(C), at[0].val.clusterDim.y = 1, at[0].val.clusterDim.z = 1;
            cfg.gridDim = dim3(C), cfg.blockDim = dim3(static_cast<unsigned>((threads + 31) / 32 * 32));
            cfg.dynamicSmemBytes = smem, cfg.attrs = at, cfg.numAttrs = 1;
            int nclusters = 0;
            if (cudaOccupancyMaxActiveClusters(&nclusters, fn, &cfg) != cudaSuccess || nclusters < 1) {
                cudaGetLastError();
                continue; // this device does not co-schedule that many CTAs of this size: try half
            }
            L->cluster_kernel = c.fn;
            L->cl_cpt = cpt, L->cl_vert = vert, L->cl_maxt = maxt;
            s0.cl_size = C, s0.cl_rpc = rpc, s0.cl_threads = static_cast<int>(cfg.blockDim.x), s0.cl_smem = smem;
        }
        if (asked && !L->cluster_kernel) return fail(LBM_ECUDA, "step_cluster_kernel cannot be launched on this device");
        if (const char* t = getenv("LBM_CL_SYNC")) L->cl_sync = atoi(t);
        if (getenv("LBM_DEBUG") && L->cluster_kernel)
            fprintf(stderr, "[lbm] step_cluster_kernel: %d CTAs x %d threads, %d rows per CTA, %d cells per thread (vert %d, <= %d threads), %zu B smem\n",
                    s0.cl_size, s0.cl_threads, s0.cl_rpc, L->cl_cpt, L->cl_vert, L->cl_maxt, s0.cl_smem);
    }
    if (k.tma) {
        TmaChoice c;
        const bool ok = strict ? tma_by_shape<true>(k.tma_ty, k.tma_stages, k.tma_minb, &c)
                               : tma_by_shape<false>(k.tma_ty, k.tma_stages, k.tma_minb, &c);
        if (!ok)
            return fail(LBM_EINVAL, "no step_tma_kernel variant with %d rows per tile, %d stages, %d CTAs per SM", k.tma_ty,
                        k.tma_stages, k.tma_minb);
        L->tma_kernel = c.fn;
        L->tma_ty = c.ty, L->tma_stages = c.stages, L->tma_minb = c.minb;
        L->tma_smem = static_cast<size_t>(c.stages) * stage_floats(c.ty) * sizeof(float);
        // the same kernel runs on every device of the lattice: attributes and occupancy per device
        for (int i = 0; i < L->nslabs; i++) {
            CU(cudaSetDevice(L->slabs[i].device));
            CU(cudaFuncSetAttribute(reinterpret_cast<const void*>(c.fn), cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    static_cast<int>(L->tma_smem)));
            int resident = 0, sms = 0;
            CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, reinterpret_cast<const void*>(c.fn), 32 * c.ty + 32,
                                                             L->tma_smem));
            CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, L->slabs[i].device));
            if (resident < 1) return fail(LBM_ECUDA, "step_tma_kernel does not fit on an SM (%zu bytes of shared memory)", L->tma_smem);
            if (i == 0 || resident < L->tma_resident) L->tma_resident = resident;
            if (i == 0 || sms < L->sm_count) L->sm_count = sms;
        }
    }
    if (k.f2 && L->tma_kernel) {
        F2Choice c;
        const bool ok = strict ? f2_by_shape<true>(k.f2_r, k.f2_srows, k.f2_stages, k.f2_minb, &c)
                               : f2_by_shape<false>(k.f2_r, k.f2_srows, k.f2_stages, k.f2_minb, &c);
        if (!ok)
            return fail(LBM_EINVAL, "no step2_kernel variant with %d consumer warps, %d rows per stage, %d stages, %d CTAs per SM", k.f2_r,
                        k.f2_srows, k.f2_stages, k.f2_minb);
        L->f2_kernel = c.fn;
        L->f2_r = c.r, L->f2_srows = c.srows, L->f2_stages = c.stages, L->f2_minb = c.minb;
        L->f2_iter_rows = c.r;
        L->f2_smem = (static_cast<size_t>(c.stages) * stage_floats(c.srows) + static_cast<size_t>(c.r + 2) * F2_B2ROW) * sizeof(float) + 64;
        for (int i = 0; i < L->nslabs; i++) {
            CU(cudaSetDevice(L->slabs[i].device));
            CU(cudaFuncSetAttribute(reinterpret_cast<const void*>(c.fn), cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    static_cast<int>(L->f2_smem)));
            int resident = 0;
            CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, reinterpret_cast<const void*>(c.fn), 32 * c.r, L->f2_smem));
            if (resident < 1) return fail(LBM_ECUDA, "step2_kernel does not fit on an SM (%zu bytes of shared memory)", L->f2_smem);
            if (i == 0 || resident < L->f2_resident) L->f2_resident = resident;
        }
    }
    for (int i = 0; i < L->nslabs; i++) slab_geometry(L, L->slabs[i], k);
    return LBM_OK;
}

void free_slab(Slab& s)
{
    cudaSetDevice(s.device);
    destroy_graphs(s);
    for (int i = 0; i < 2; i++)
        if (s.ipc_open[i]) cudaIpcCloseMemHandle(s.ipc_open[i]);
    cudaFree(s.lat[0]); // both lattices
    cudaFree(s.obst);
    cudaFree(s.fluid_dev);
    cudaFree(s.halo_block);
    cudaFree(s.obst_halo);
    cudaFree(s.arrive);
    cudaFree(s.ctrl);
    cudaFree(s.error);
    cudaFree(s.sums);
    cudaFree(s.sums_ref);
    cudaFree(s.state_sums);
    cudaFree(s.loop_barrier);
    cudaFree(s.ll_packets);
    cudaFree(s.band_flags);
    if (s.ev0) cudaEventDestroy(s.ev0);
    if (s.ev1) cudaEventDestroy(s.ev1);
    if (s.own_stream) cudaStreamDestroy(s.own_stream);
    cudaGetLastError();
}

int check_error_flags(lbm_lattice* L)
{
    for (int i = 0; i < L->nslabs; i++) {
        Slab& s = L->slabs[i];
        CU(cudaSetDevice(s.device));
        int e = 0;
        CU(cudaMemcpyAsync(&e, s.error, sizeof e, cudaMemcpyDeviceToHost, s.stream));
        CU(cudaStreamSynchronize(s.stream));
        if (e) {
            CU(cudaMemsetAsync(s.error, 0, sizeof(int), s.stream));
            return fail(LBM_ETIMEOUT, "halo wait timed out on slab %d (rows %d..%d): a neighbour never delivered its row", i,
                        s.row0, s.row1 - 1);
        }
    }
    return LBM_OK;
}

// push this slab's boundary rows of the current lattice into the neighbours' rings (every slot, all nine
// entries) and set their flags to "every epoch so far delivered": used after lbm_upload_cells
int push_boundary_rows(lbm_lattice* L, Slab& s)
{
    CU(cudaSetDevice(s.device));
    RingFillArgs a;
    a.lat = s.lat[L->cur];
    a.pf = plane_floats(L, s);
    a.ring_s = s.peer_ring_s, a.ring_n = s.peer_ring_n;
    a.nx = L->p.nx, a.rows = s.rows, a.pitch = L->pitch, a.nring = L->ring;
    a.slot_stride = static_cast<unsigned long long>(RING_ENTRIES) * L->pitch;
    ring_fill_kernel<<<(L->p.nx + 255) / 256, 256, 0, s.stream>>>(a);
    L->launches++;
    CU(cudaGetLastError());
    const unsigned long long v = static_cast<unsigned long long>(L->epoch);
    CU(cudaMemcpyAsync(s.peer_flag_s, &v, sizeof v, cudaMemcpyHostToDevice, s.stream));
    CU(cudaMemcpyAsync(s.peer_flag_n, &v, sizeof v, cudaMemcpyHostToDevice, s.stream));
    CU(cudaStreamSynchronize(s.stream));
    return LBM_OK;
}

// what both ends of a halo link must agree on, packed into the handle: arithmetic flavour, halo mode and lag,
// whether the lattice advances in pairs of steps (the epoch sequence differs), which resident kernel it runs, and the
// tuning options that change how many CTAs signal a halo row (opt.block, opt.kernel)
int32_t halo_config_word(const lbm_lattice* L)
{
    return static_cast<int32_t>((L->opt.arith & 0xf) | ((L->opt.halo_mode & 0xf) << 4) | ((L->opt.halo_lag & 0xff) << 8) |
                                ((L->slabs[0].use_f2 ? 1 : 0) << 16) | ((L->slabs[0].use_loop ? 1 : 0) << 17) |
                                ((L->slabs[0].use_ll ? 1 + L->slabs[0].ll_var : 0) << 18) |
                                // CTA shape and kernel code decide how many CTAs signal a halo row: both ends must agree
                                (((L->opt.block / 128) & 0x7) << 20) | ((L->slabs[0].use_band ? 1 : 0) << 23) |
                                ((static_cast<unsigned>(L->opt.kernel) * 2654435761u >> 25) << 24 & 0x7f000000));
}

// the halo protocol stores and adds into the neighbour device's memory from inside kernels: peer access is not
// enough, the link must also carry native atomics (NVLink does; PCIe peers may not)
int check_peer_link(int device, int peer)
{
    if (device == peer) return LBM_OK;
    int can = 0;
    cudaDeviceCanAccessPeer(&can, device, peer);
    if (!can) return fail(LBM_ENODEVICE, "device %d cannot access device %d as a peer", device, peer);
    int atomics = 0;
    if (cudaDeviceGetP2PAttribute(&atomics, cudaDevP2PAttrNativeAtomicSupported, device, peer) != cudaSuccess) {
        cudaGetLastError();
        atomics = 0;
    }
    if (!atomics)
        return fail(LBM_ENODEVICE, "device %d has no native peer atomics to device %d (the halo flags need them; NVLink provides them)",
                    device, peer);
    return LBM_OK;
}

// copy the obstacle bits of the neighbours' rows that touch this slab (step2_kernel recomputes those rows'
// intermediate step): `south_edge` / `north_edge` point at the neighbours' edge_obst blocks ([2][opitch])
int fetch_halo_obstacles(lbm_lattice* L, Slab& s, const uint32_t* south_edge, const uint32_t* north_edge)
{
    CU(cudaSetDevice(s.device));
    const size_t row = static_cast<size_t>(L->opitch) * sizeof(uint32_t);
    // the south neighbour's LAST row (second row of its block), the north neighbour's row 0 (first row of its block)
    CU(cudaMemcpyAsync(s.obst_halo, south_edge + L->opitch, row, cudaMemcpyDefault, s.stream));
    CU(cudaMemcpyAsync(s.obst_halo + L->opitch, north_edge, row, cudaMemcpyDefault, s.stream));
    CU(cudaStreamSynchronize(s.stream));
    return LBM_OK;
}

} // namespace

// =================================================================================================
// C ABI
// =================================================================================================
extern "C" {

void lbm_default_options(lbm_options_t* opt)
{
    if (!opt) return;
    opt->arith = LBM_ARITH_STRICT;
    opt->halo_mode = LBM_HALO_SYNC;
    opt->halo_lag = 0;
    opt->use_graph = 1;
    opt->kernel = 0;
    opt->block = 0;
}

const char* lbm_last_error(void) { return g_err; }

int lbm_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int lbm_partition(int ny, int nslabs, int* starts)
{
    if (!starts || nslabs < 1 || ny < 1) return fail(LBM_EINVAL, "lbm_partition: bad arguments");
    const int base = ny / nslabs, rem = ny % nslabs;
    int row = 0;
    for (int r = 0; r < nslabs; r++) {
        // the remainder rows go to the last slabs, so the last slab is never the smallest
        const int rows = base + (r >= nslabs - rem ? 1 : 0);
        starts[r] = row;
        row += rows;
        if (nslabs > 1) {
            const int need = (r == nslabs - 1) ? 3 : 2;
            if (rows < need)
                return fail(LBM_EINVAL, "ny=%d is too small for %d slabs (slab %d would own %d rows, needs %d)", ny, nslabs, r,
                            rows, need);
        }
    }
    starts[nslabs] = ny;
    return LBM_OK;
}

static int create_common(const lbm_param_t* params, const lbm_options_t* opt, int nslabs, const int* devices,
                         const int* starts, const void* obstacles /* rows of slab 0 onwards, contiguous */, bool packed,
                         bool per_process, int rank, int nranks, lbm_lattice_t** out)
{
    if (!out) return fail(LBM_EINVAL, "out is NULL");
    *out = nullptr;
    int rc = validate_params(params);
    if (rc) return rc;
    if (!obstacles) return fail(LBM_EINVAL, "obstacles is NULL");
    if (!constants_ok()) return fail(LBM_EINVAL, "host float arithmetic does not reproduce the reference's constants");
    rc = check_device_available();
    if (rc) return rc;
    lbm_lattice* L = new (std::nothrow) lbm_lattice();
    if (!L) return fail(LBM_ENOMEM, "out of host memory");
    L->nslabs = nslabs;
    L->per_process = per_process;
    L->rank = rank;
    L->nranks = nranks;
    L->slabs = new (std::nothrow) Slab[nslabs];
    if (!L->slabs) {
        delete L;
        return fail(LBM_ENOMEM, "out of host memory");
    }
    for (int i = 0; i < nslabs; i++) {
        Slab& s = L->slabs[i];
        s.device = devices[i];
        s.row0 = starts[i];
        s.row1 = starts[i + 1];
        s.rows = s.row1 - s.row0;
        const int g = params->ny - 2; // the driven row, SerialCode:226
        s.accel_row = (g >= s.row0 && g < s.row1) ? g - s.row0 : -1;
    }
    for (int i = 0; i < nslabs; i++)
        for (int j = 0; j < i; j++)
            if (devices[j] == devices[i]) L->interleaved = true; // slabs sharing a device: step-major launches on one stream
    rc = common_setup(L, params, opt);
    for (int i = 0; i < nslabs && !rc; i++) {
        Slab& s = L->slabs[i];
        const size_t row_bytes = packed ? static_cast<size_t>(L->opitch) * sizeof(uint32_t) : static_cast<size_t>(params->nx) * sizeof(int);
        rc = alloc_slab(L, s, static_cast<const char*>(obstacles) + static_cast<size_t>(s.row0 - starts[0]) * row_bytes, packed);
        if (!rc && s.use_tma) rc = make_tensor_maps(L, s);
        if (!rc && uses_halo(L)) rc = alloc_halo(L, s);
    }
    if (rc) {
        char keep[sizeof g_err];
        memcpy(keep, g_err, sizeof keep);
        lbm_destroy(L);
        memcpy(g_err, keep, sizeof keep);
        return rc;
    }
    *out = L;
    return LBM_OK;
}

static int create_on_impl(const lbm_param_t* params, const void* obstacles, bool packed, int nslabs, const int* devices,
                          const lbm_options_t* opt, lbm_lattice_t** out)
{
    if (nslabs < 1 || nslabs > 1024) return fail(LBM_EINVAL, "bad slab count %d", nslabs);
    if (!devices) return fail(LBM_EINVAL, "devices is NULL");
    int rc = validate_params(params);
    if (rc) return rc;
    rc = check_device_available();
    if (rc) return rc;
    const int ndev = lbm_device_count();
    for (int i = 0; i < nslabs; i++)
        if (devices[i] < 0 || devices[i] >= ndev)
            return fail(LBM_ENODEVICE, "slab %d asks for CUDA device %d but only %d device(s) are visible", i, devices[i], ndev);
    std::vector<int> starts(nslabs + 1);
    rc = lbm_partition(params->ny, nslabs, starts.data());
    if (rc) return rc;
    lbm_lattice_t* L = nullptr;
    rc = create_common(params, opt, nslabs, devices, starts.data(), obstacles, packed, false, 0, nslabs, &L);
    if (rc) return rc;
    if (nslabs > 1) {
        // slabs that share a device share a stream: their kernels then run one after the other in
        // launch order, so a halo wait is always already satisfied (kernels that spin on one another
        // must never depend on being co-resident on one GPU)
        for (int i = 0; i < nslabs; i++)
            for (int j = 0; j < i; j++)
                if (L->slabs[j].device == L->slabs[i].device) {
                    L->slabs[i].stream = L->slabs[j].stream;
                    L->interleaved = true;
                    break;
                }
        // peer access between neighbouring devices, then wire the ring (periodic: MPI/d2q9-bgk.c:253-254)
        for (int i = 0; i < nslabs && !rc; i++) {
            Slab& s = L->slabs[i];
            Slab& south = L->slabs[(i - 1 + nslabs) % nslabs];
            Slab& north = L->slabs[(i + 1) % nslabs];
            for (Slab* nb : {&south, &north}) {
                if (nb->device == s.device) continue;
                rc = check_peer_link(s.device, nb->device);
                if (rc) break;
                cudaSetDevice(s.device);
                cudaError_t e = cudaDeviceEnablePeerAccess(nb->device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
                    rc = fail(LBM_ECUDA, "cudaDeviceEnablePeerAccess(%d -> %d): %s", s.device, nb->device, cudaGetErrorString(e));
                    break;
                }
                cudaGetLastError();
            }
            if (rc) break;
            s.peer_ring_s = south.ring_n, s.peer_flag_s = south.flag_n;
            s.peer_ring_n = north.ring_s, s.peer_flag_n = north.flag_s;
            s.peer_ll_s = south.ll_recv_n, s.peer_ll_n = north.ll_recv_s;
            rc = fetch_halo_obstacles(L, s, south.edge_obst, north.edge_obst);
        }
        if (rc) {
            char keep[sizeof g_err];
            memcpy(keep, g_err, sizeof keep);
            lbm_destroy(L);
            memcpy(g_err, keep, sizeof keep);
            return rc;
        }
        L->connected = true;
    } else {
        L->connected = true;
    }
    rc = prepare_graphs(L);
    if (rc) {
        char keep[sizeof g_err];
        memcpy(keep, g_err, sizeof keep);
        lbm_destroy(L);
        memcpy(g_err, keep, sizeof keep);
        return rc;
    }
    *out = L;
    return LBM_OK;
}

int lbm_create_on(const lbm_param_t* params, const int* obstacles, int nslabs, const int* devices,
                  const lbm_options_t* opt, lbm_lattice_t** out)
{
    return create_on_impl(params, obstacles, false, nslabs, devices, opt, out);
}

static int create_impl(const lbm_param_t* params, const void* obstacles, bool packed, int ngpus, const lbm_options_t* opt,
                       lbm_lattice_t** out)
{
    if (ngpus < 1) return fail(LBM_EINVAL, "ngpus must be >= 1, got %d", ngpus);
    int rc = check_device_available();
    if (rc) return rc;
    const int ndev = lbm_device_count();
    if (ngpus > ndev) return fail(LBM_ENODEVICE, "%d GPUs requested but only %d CUDA device(s) are visible", ngpus, ndev);
    std::vector<int> devices(ngpus);
    for (int i = 0; i < ngpus; i++) devices[i] = i;
    return create_on_impl(params, obstacles, packed, ngpus, devices.data(), opt, out);
}

int lbm_create(const lbm_param_t* params, const int* obstacles, int ngpus, const lbm_options_t* opt, lbm_lattice_t** out)
{
    return create_impl(params, obstacles, false, ngpus, opt, out);
}

int lbm_create_packed(const lbm_param_t* params, const unsigned* obstacle_bits, int ngpus, const lbm_options_t* opt,
                      lbm_lattice_t** out)
{
    return create_impl(params, obstacle_bits, true, ngpus, opt, out);
}

size_t lbm_packed_words_per_row(int nx) { return nx > 0 ? (static_cast<size_t>(nx) + 31) / 32 : 0; }

static int create_slab_impl(const lbm_param_t* params, const void* obstacle_rows, bool packed, int row0, int row1, int rank, int nranks,
                            int device, const lbm_options_t* opt, lbm_lattice_t** out)
{
    int rc = validate_params(params);
    if (rc) return rc;
    if (nranks < 1 || rank < 0 || rank >= nranks) return fail(LBM_EINVAL, "bad rank %d of %d", rank, nranks);
    if (row0 < 0 || row1 > params->ny || row1 - row0 < 1) return fail(LBM_EINVAL, "bad row range [%d, %d)", row0, row1);
    const int g = params->ny - 2;
    if (nranks > 1 && g >= row0 && g < row1 && (g == row0 || g == row1 - 1))
        return fail(LBM_EINVAL, "the driven row ny-2 must be an interior row of its slab (use lbm_partition)");
    if (nranks == 1 && (row0 != 0 || row1 != params->ny)) return fail(LBM_EINVAL, "a single rank must own every row");
    rc = check_device_available();
    if (rc) return rc;
    if (device < 0 || device >= lbm_device_count()) return fail(LBM_ENODEVICE, "CUDA device %d is not visible", device);
    const int starts[2] = {row0, row1};
    return create_common(params, opt, 1, &device, starts, obstacle_rows, packed, true, rank, nranks, out);
}

int lbm_create_slab(const lbm_param_t* params, const int* obstacle_rows, int row0, int row1, int rank, int nranks, int device,
                    const lbm_options_t* opt, lbm_lattice_t** out)
{
    return create_slab_impl(params, obstacle_rows, false, row0, row1, rank, nranks, device, opt, out);
}

int lbm_create_slab_packed(const lbm_param_t* params, const unsigned* obstacle_bit_rows, int row0, int row1, int rank, int nranks,
                           int device, const lbm_options_t* opt, lbm_lattice_t** out)
{
    return create_slab_impl(params, obstacle_bit_rows, true, row0, row1, rank, nranks, device, opt, out);
}

int lbm_halo_export(lbm_lattice_t* L, void* handle)
{
    if (!L || !handle) return fail(LBM_EINVAL, "NULL argument");
    if (!L->per_process) return fail(LBM_EINVAL, "lbm_halo_export is for lbm_create_slab lattices");
    Slab& s = L->slabs[0];
    HaloHandle h;
    memset(&h, 0, sizeof h);
    h.magic = HANDLE_MAGIC;
    h.pid = static_cast<int32_t>(getpid());
    h.device = s.device;
    h.pitch = L->pitch;
    h.ring = L->ring;
    h.config = halo_config_word(L);
    h.local_ptr = reinterpret_cast<uint64_t>(s.halo_block);
    h.bytes = s.halo_bytes;
    CU(cudaSetDevice(s.device));
    CU(cudaIpcGetMemHandle(&h.ipc, s.halo_block));
    memset(handle, 0, LBM_HALO_HANDLE_BYTES);
    memcpy(handle, &h, sizeof h);
    return LBM_OK;
}

int lbm_halo_connect(lbm_lattice_t* L, const void* south_handle, const void* north_handle)
{
    if (!L || !south_handle || !north_handle) return fail(LBM_EINVAL, "NULL argument");
    if (!L->per_process) return fail(LBM_EINVAL, "lbm_halo_connect is for lbm_create_slab lattices");
    if (L->connected) return fail(LBM_EINVAL, "already connected");
    Slab& s = L->slabs[0];
    CU(cudaSetDevice(s.device));
    HaloHandle h[2];
    memcpy(&h[0], south_handle, sizeof(HaloHandle));
    memcpy(&h[1], north_handle, sizeof(HaloHandle));
    char* base[2] = {nullptr, nullptr};
    for (int i = 0; i < 2; i++) {
        if (h[i].magic != HANDLE_MAGIC) return fail(LBM_EINVAL, "neighbour handle %d is not a halo handle", i);
        if (h[i].pitch != L->pitch || h[i].ring != L->ring || h[i].bytes != s.halo_bytes)
            return fail(LBM_EINVAL, "neighbour %d was created with a different nx / halo_lag", i);
        if (h[i].config != halo_config_word(L))
            return fail(LBM_EINVAL, "neighbour %d was created with different options (arith / halo_mode / halo_lag / kernel): 0x%x vs 0x%x",
                        i, h[i].config, halo_config_word(L));
        if (h[i].device != s.device) {
            int rc = check_peer_link(s.device, h[i].device);
            if (rc) return rc;
        }
        if (h[i].pid == static_cast<int32_t>(getpid())) {
            base[i] = reinterpret_cast<char*>(h[i].local_ptr); // same process (e.g. a ring of one)
            if (h[i].device != s.device) {
                cudaError_t e = cudaDeviceEnablePeerAccess(h[i].device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
                    return fail(LBM_ECUDA, "cudaDeviceEnablePeerAccess: %s", cudaGetErrorString(e));
                cudaGetLastError();
            }
        } else if (i == 1 && h[0].pid == h[1].pid && h[0].local_ptr == h[1].local_ptr && base[0]) {
            base[1] = base[0]; // ring of two: both neighbours are the same rank
        } else {
            void* p = nullptr;
            cudaError_t e = cudaIpcOpenMemHandle(&p, h[i].ipc, cudaIpcMemLazyEnablePeerAccess);
            if (e != cudaSuccess)
                return fail(LBM_ENODEVICE, "cannot map the halo ring of the neighbour on device %d (pid %d): %s", h[i].device,
                            h[i].pid, cudaGetErrorString(e));
            s.ipc_open[i] = p;
            base[i] = static_cast<char*>(p);
        }
    }
    const size_t ring_floats = ring_floats_of(L);
    auto ring_s_of = [&](char* b) { return reinterpret_cast<float*>(b); };
    auto ring_n_of = [&](char* b) { return reinterpret_cast<float*>(b) + ring_floats; };
    auto flag_s_of = [&](char* b) { return reinterpret_cast<unsigned long long*>(b + 2 * ring_floats * sizeof(float)); };
    auto flag_n_of = [&](char* b) { return flag_s_of(b) + 16; };
    // my south neighbour receives my row 0 in ITS north ring; my north neighbour my last row in ITS south ring
    s.peer_ring_s = ring_n_of(base[0]), s.peer_flag_s = flag_n_of(base[0]);
    s.peer_ring_n = ring_s_of(base[1]), s.peer_flag_n = flag_s_of(base[1]);
    {
        const size_t ll_off = reinterpret_cast<char*>(s.ll_recv_s) - s.halo_block, ll_bytes = 2 * static_cast<size_t>(L->pitch) * sizeof(uint4);
        s.peer_ll_s = reinterpret_cast<uint4*>(base[0] + ll_off + ll_bytes); // the south neighbour's ll_recv_n
        s.peer_ll_n = reinterpret_cast<uint4*>(base[1] + ll_off);            // the north neighbour's ll_recv_s
    }
    auto edge_obst_of = [&](char* b) { return reinterpret_cast<const uint32_t*>(b + 2 * ring_floats * sizeof(float) + 256); };
    int rc = fetch_halo_obstacles(L, s, edge_obst_of(base[0]), edge_obst_of(base[1]));
    if (rc) return rc;
    L->connected = true;
    return prepare_graphs(L);
}

int lbm_set_stream(lbm_lattice_t* L, void* cuda_stream)
{
    if (!L) return fail(LBM_EINVAL, "NULL lattice");
    Slab& s0 = L->slabs[0];
    cudaStream_t old = s0.stream;
    cudaStream_t now = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : s0.own_stream;
    CU(cudaSetDevice(s0.device));
    CU(cudaStreamSynchronize(old));
    for (int i = 0; i < L->nslabs; i++)
        if (L->slabs[i].stream == old) L->slabs[i].stream = now;
    return LBM_OK;
}

int lbm_run(lbm_lattice_t* L, int iters)
{
    if (!L) return fail(LBM_EINVAL, "NULL lattice");
    if (iters < 0) return fail(LBM_EINVAL, "iters must be >= 0");
    if (!L->connected) return fail(LBM_EINVAL, "lbm_halo_connect has not been called");
    if (L->steps_done + iters > 0x7ffffff0LL) return fail(LBM_EINVAL, "step counter would overflow");
    if (L->poisoned) return fail(LBM_ECUDA, "an earlier lbm_run failed half-way: the slabs are no longer at the same step");
    if (iters == 0) {
        L->run_first = L->steps_done;
        L->run_iters = 0;
        return LBM_OK;
    }
    const int first = static_cast<int>(L->steps_done);
    const int parity = L->cur;
    const bool use_graphs = L->opt.use_graph && !L->interleaved;
    bool all_loop = true, f2 = true;
    bool ll = true;
    for (int i = 0; i < L->nslabs; i++) ll = ll && L->slabs[i].use_ll;
    bool band = !ll;
    for (int i = 0; i < L->nslabs; i++) band = band && L->slabs[i].use_band;
    const bool cluster = !ll && !band && (L->nslabs == 1 && L->slabs[0].use_cluster);
    for (int i = 0; i < L->nslabs; i++) all_loop = all_loop && L->slabs[i].use_loop, f2 = f2 && L->slabs[i].use_f2;
    if (cluster || ll || band) all_loop = true; // no graphs, no per-step launches
    if (band) {
        f2 = false;
        for (int i = 0; i < L->nslabs; i++) {
            Slab& s = L->slabs[i];
            if (s.band_flags) continue;
            CU(cudaSetDevice(s.device));
            CU(cudaMalloc(&s.band_flags, (static_cast<size_t>(s.band_grid) * 32 + 1 + 2 * 1024 + 2) * sizeof(unsigned)));
        }
        if (uses_halo(L) && static_cast<unsigned long long>(L->ll_flags) + static_cast<unsigned long long>(iters) + 2ull > 0x7fffffffull)
            return fail(LBM_EINVAL, "packet flags exhausted (2^31 steps on one multi-GPU lattice)");
    }
    if (ll) {
        f2 = false;
        // packet buffers (zeroed: flag 0 is never waited for); flags are never reused by a lattice
        const bool wrap = static_cast<unsigned long long>(L->ll_flags) + static_cast<unsigned long long>(iters) + 2ull > 0x7fffffffull;
        if (wrap && uses_halo(L)) return fail(LBM_EINVAL, "packet flags exhausted (2^31 steps on one multi-GPU lattice)");
        for (int i = 0; i < L->nslabs; i++) {
            Slab& s = L->slabs[i];
            CU(cudaSetDevice(s.device));
            const size_t bytes = 4 * static_cast<size_t>(s.rows) * L->pitch * sizeof(uint4);
            if (!s.ll_packets) {
                CU(cudaMalloc(&s.ll_packets, bytes));
                CU(cudaMemsetAsync(s.ll_packets, 0, bytes, s.stream));
            } else if (wrap) {
                CU(cudaMemsetAsync(s.ll_packets, 0, bytes, s.stream));
            }
        }
        if (wrap) L->ll_flags = 0;
    }
    // everything that can fail without having touched the lattice comes first: a failure here leaves the run
    // un-started and the call can be repeated
    for (int i = 0; i < L->nslabs; i++) {
        Slab& s = L->slabs[i];
        int rc = ensure_sums(L, s, static_cast<size_t>(iters));
        if (rc) return rc;
        if (use_graphs && !all_loop && iters >= GRAPH_STEPS && !s.graph[parity]) {
            rc = build_graph(L, s, parity);
            if (rc) return rc;
        }
    }
    // from here on a failure leaves slabs at different steps (or the driven row accelerated twice on a retry)
    struct Poison {
        lbm_lattice* L;
        bool armed = true;
        ~Poison()
        {
            if (armed) L->poisoned = true;
        }
    } poison{L};
    L->run_first = L->steps_done;
    L->run_iters = iters;

    // lattices that advance in pairs of steps re-deliver their boundary rows at the start of every run (halo
    // epoch e0): the accelerate_flow() pre-pass below changes row ny-2, which the first slab reads as its far
    // south halo row, after the previous run's last step delivered it
    const bool start_push = f2 && uses_halo(L);
    const long long e0 = L->epoch;
    const int epoch_base = static_cast<int>(e0 + (start_push ? 1 : 0)); // epoch of the first step kernel
    for (int i = 0; i < L->nslabs; i++) {
        Slab& s = L->slabs[i];
        CU(cudaSetDevice(s.device));
        set_ctrl_kernel<<<1, 32, 0, s.stream>>>(s.ctrl, first, first, first + iters - 1, epoch_base);
        L->launches++;
        if (i == 0) CU(cudaEventRecord(s.ev0, s.stream));
        // accelerate_flow() of the first step (SerialCode:209); every later step's is applied by the
        // previous step's store
        if (s.accel_row >= 0) {
            AccelArgs aa;
            const size_t pf = plane_floats(L, s);
            const size_t off = static_cast<size_t>(s.accel_row) * L->pitch;
            for (int k = 0; k < Q; k++) aa.f[k] = s.lat[L->cur] + k * pf + off;
            aa.obst_row = s.obst + static_cast<size_t>(s.accel_row) * L->opitch;
            aa.nx = L->p.nx;
            aa.w1a = L->w1a, aa.w2a = L->w2a;
            accelerate_row_kernel<<<(L->p.nx + 255) / 256, 256, 0, s.stream>>>(aa);
            L->launches++;
        }
        CU(cudaGetLastError());
    }
    if (start_push) {
        for (int i = 0; i < L->nslabs; i++) {
            Slab& s = L->slabs[i];
            CU(cudaSetDevice(s.device));
            HaloPushArgs pa;
            memset(&pa, 0, sizeof pa);
            pa.lat = s.lat[L->cur];
            pa.pf = plane_floats(L, s);
            pa.h = make_halo_cfg(L, s);
            pa.ctrl = s.ctrl;
            pa.epoch_offset = -1;
            pa.nx = L->p.nx, pa.rows = s.rows, pa.pitch = L->pitch;
            halo_push_kernel<<<(L->p.nx / 4 + 127) / 128, 128, 0, s.stream>>>(pa);
            L->launches++;
            CU(cudaGetLastError());
        }
    }
    int done = 0;        // timesteps queued
    long long passes = 0; // lattice swaps queued
    long long epochs = 0; // halo epochs queued after epoch_base
    if (band) {
        // every step of this run in ONE cooperative launch of step_band_kernel per slab, a band of rows per CTA; slabs
        // on different GPUs run at the same time and exchange step_ll_kernel's (unshifted) packets at their boundaries
        const unsigned flag_base = L->ll_flags + 1u;
        if (uses_halo(L))
            for (int i = 0; i < L->nslabs; i++) {
                Slab& s = L->slabs[i];
                CU(cudaSetDevice(s.device));
                LLSeedArgs sa;
                sa.lat = s.lat[parity];
                sa.pf = plane_floats(L, s);
                sa.halo_send_s = s.peer_ll_s, sa.halo_send_n = s.peer_ll_n;
                sa.seed_flag = flag_base;
                sa.nx = L->p.nx, sa.rows = s.rows, sa.pitch = L->pitch;
                sa.unshifted = 1;
                ll_seed_kernel<<<(L->p.nx + 127) / 128, 128, 0, s.stream>>>(sa);
                L->launches++;
                CU(cudaGetLastError());
            }
        for (int i = 0; i < L->nslabs; i++) {
            Slab& s = L->slabs[i];
            CU(cudaSetDevice(s.device));
            BandArgs a;
            memset(&a, 0, sizeof a);
            a.lat[0] = s.lat[0], a.lat[1] = s.lat[1];
            a.pf = plane_floats(L, s);
            a.obst = s.obst;
            a.sums = s.sums;
            a.nslots = s.nslots;
            a.flags = s.band_flags;
            a.error = s.error;
            a.timeout_ns = L->timeout_ns;
            a.first_step = first, a.nsteps = iters, a.last_step = first + iters - 1;
            a.src = parity;
            a.nx = L->p.nx, a.nxv = L->p.nx / 4, a.rows = s.rows, a.pitch = L->pitch, a.opitch = L->opitch;
            a.accel_row = s.accel_row;
            a.omega = L->p.omega, a.w1a = L->w1a, a.w2a = L->w2a;
            if (uses_halo(L)) {
                a.halo_recv_s = s.ll_recv_s, a.halo_recv_n = s.ll_recv_n;
                a.halo_send_s = s.peer_ll_s, a.halo_send_n = s.peer_ll_n;
            }
            a.flag_base = flag_base;
            // full SMs: the rows are dealt out per SM (the CTAs find their SM themselves)
            a.per_sm = s.band_per_sm;
            a.nsm = L->sm_count;
            a.sm_table = s.band_flags + static_cast<size_t>(s.band_grid) * 32;
            CU(cudaMemsetAsync(s.band_flags, 0, (static_cast<size_t>(s.band_grid) * 32 + 1 + 2 * 1024 + 2) * sizeof(unsigned), s.stream));
            void* kp[1] = {&a};
            CU(cudaLaunchCooperativeKernel(reinterpret_cast<const void*>(L->band_kernel), dim3(s.band_grid),
                                           dim3(static_cast<unsigned>(L->band_block)), kp, 0, s.stream));
            L->launches++;
        }
        if (uses_halo(L)) L->ll_flags += static_cast<unsigned>(iters) + 1u;
        done = iters;
        passes = iters;
        epochs = iters;
    } else if (ll) {
        // every step of this run in ONE cooperative launch of step_ll_kernel per slab, a CTA per row; slabs on
        // different GPUs run at the same time and exchange packets (the seed packets of the current state first)
        const unsigned seed_flag = L->ll_flags + 1u, flag_base = L->ll_flags + 1u;
        if (uses_halo(L))
            for (int i = 0; i < L->nslabs; i++) {
                Slab& s = L->slabs[i];
                CU(cudaSetDevice(s.device));
                LLSeedArgs sa;
                sa.lat = s.lat[parity];
                sa.pf = plane_floats(L, s);
                sa.halo_send_s = s.peer_ll_s, sa.halo_send_n = s.peer_ll_n;
                sa.seed_flag = seed_flag;
                sa.nx = L->p.nx, sa.rows = s.rows, sa.pitch = L->pitch;
                sa.unshifted = 0;
                ll_seed_kernel<<<(L->p.nx + 127) / 128, 128, 0, s.stream>>>(sa);
                L->launches++;
                CU(cudaGetLastError());
            }
        for (int i = 0; i < L->nslabs; i++) {
            Slab& s = L->slabs[i];
            CU(cudaSetDevice(s.device));
            LLArgs a;
            memset(&a, 0, sizeof a);
            a.lat[0] = s.lat[0], a.lat[1] = s.lat[1];
            a.pf = plane_floats(L, s);
            a.obst = s.obst;
            a.sums = s.sums;
            a.nslots = s.nslots;
            const size_t half = 2 * static_cast<size_t>(s.rows) * L->pitch;
            a.pk_north = s.ll_packets, a.pk_south = s.ll_packets + half;
            if (uses_halo(L)) {
                a.halo_recv_s = s.ll_recv_s, a.halo_recv_n = s.ll_recv_n;
                a.halo_send_s = s.peer_ll_s, a.halo_send_n = s.peer_ll_n;
            }
            a.seed_flag = seed_flag;
            a.flag_base = flag_base;
            a.error = s.error;
            a.timeout_ns = L->timeout_ns;
            a.first_step = first, a.nsteps = iters, a.last_step = first + iters - 1;
            a.src = parity;
            a.nx = L->p.nx, a.rows = s.rows, a.pitch = L->pitch, a.opitch = L->opitch;
            a.accel_row = s.accel_row;
            a.omega = L->p.omega, a.w1a = L->w1a, a.w2a = L->w2a;
            void* kp[1] = {&a};
            CU(cudaLaunchCooperativeKernel(reinterpret_cast<const void*>(L->ll_kernel[s.ll_var]), dim3(static_cast<unsigned>(s.rows)),
                                           dim3(static_cast<unsigned>(s.ll_block)), kp, s.ll_smem, s.stream));
            L->launches++;
        }
        L->ll_flags += static_cast<unsigned>(iters) + 1u;
        done = iters;
        passes = iters;
        epochs = iters;
    } else if (cluster) {
        // every step of this run in ONE launch of step_cluster_kernel: the lattice goes to the cluster's shared
        // memory, comes back after the last step
        Slab& s = L->slabs[0];
        CU(cudaSetDevice(s.device));
        ClusterArgs a;
        memset(&a, 0, sizeof a);
        a.lat[0] = s.lat[0], a.lat[1] = s.lat[1];
        a.pf = plane_floats(L, s);
        a.obst = s.obst;
        a.sums = s.sums;
        a.nslots = s.nslots;
        a.first_step = first, a.nsteps = iters, a.last_step = first + iters - 1;
        a.src = parity;
        a.nx = L->p.nx, a.rows = s.rows, a.pitch = L->pitch, a.opitch = L->opitch;
        a.rpc = s.cl_rpc;
        a.accel_row = s.accel_row;
        a.sync_mode = L->cl_sync;
        a.omega = L->p.omega, a.w1a = L->w1a, a.w2a = L->w2a;
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof cfg);
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = static_cast<unsigned>(s.cl_size), at[0].val.clusterDim.y = 1, at[0].val.clusterDim.z = 1;
        cfg.gridDim = dim3(static_cast<unsigned>(s.cl_size)), cfg.blockDim = dim3(static_cast<unsigned>(s.cl_threads));
        cfg.dynamicSmemBytes = s.cl_smem, cfg.stream = s.stream, cfg.attrs = at, cfg.numAttrs = 1;
        CU(cudaLaunchKernelEx(&cfg, L->cluster_kernel, a));
        L->launches++;
        done = iters;
        passes = iters;
        epochs = iters;
    } else if (all_loop) {
        // every step of this run in cooperative launches of step_loop_kernel, one per slab (slabs on different
        // GPUs run at the same time and exchange halo rows and flags); more than one launch per slab only if
        // the 32-bit barrier counter would overflow
        long long max_steps = iters;
        for (int i = 0; i < L->nslabs; i++) max_steps = std::min<long long>(max_steps, 0xffffffffLL / L->slabs[i].loop_grid - 1);
        while (done < iters) {
            const int n = static_cast<int>(std::min<long long>(iters - done, max_steps));
            for (int i = 0; i < L->nslabs; i++) {
                Slab& s = L->slabs[i];
                CU(cudaSetDevice(s.device));
                LoopArgs a;
                memset(&a, 0, sizeof a);
                a.lat[0] = s.lat[0], a.lat[1] = s.lat[1];
                a.pf = plane_floats(L, s);
                a.h = make_halo_cfg(L, s);
                a.obst = s.obst;
                a.sums = s.sums + static_cast<size_t>(done) * s.nslots * SUM_WORDS;
                a.barrier = s.loop_barrier;
                a.nslots = s.nslots;
                a.first_step = first + done, a.nsteps = n, a.last_step = first + iters - 1;
                a.first_epoch = epoch_base + done;
                a.src = (parity + done) & 1;
                a.nx = L->p.nx, a.nxv = s.loop_nxv, a.rows = s.rows, a.pitch = L->pitch, a.opitch = L->opitch;
                a.tw_shift = s.loop_tw_shift, a.nbx = s.loop_nbx, a.nby = s.loop_nby, a.ntiles = s.loop_ntiles;
                {
                    const int nb = (s.loop_nby >= 2 ? 2 : 1) * s.loop_nbx;
                    a.nboundary = (a.h.on && static_cast<long long>(s.loop_grid) > nb && s.loop_ntiles > nb) ? nb : 0;
                }
                a.accel_row = s.accel_row;
                a.omega = L->p.omega, a.w1a = L->w1a, a.w2a = L->w2a;
                CU(cudaMemsetAsync(s.loop_barrier, 0, sizeof(unsigned), s.stream));
                void* kp[1] = {&a};
                CU(cudaLaunchCooperativeKernel(reinterpret_cast<const void*>(L->loop_kernel[s.loop_vec == 4 ? 1 : 0]),
                                               dim3(s.loop_grid), dim3(s.loop_block), kp, 0, s.stream));
                L->launches++;
            }
            done += n;
            passes += n;
            epochs += n;
        }
    }
    if (use_graphs && !all_loop) {
        while (iters - done >= GRAPH_STEPS) {
            for (int i = 0; i < L->nslabs; i++) {
                Slab& s = L->slabs[i];
                CU(cudaSetDevice(s.device));
                CU(cudaGraphLaunch(s.graph[parity], s.stream));
                L->launches += (s.use_f2 ? GRAPH_STEPS / 2 : GRAPH_STEPS * (s.use_tma ? 2 : 1)) + 1;
                if (i == 0 && getenv("LBM_DEBUG") && s.dbg_events.size() < 256) {
                    cudaEvent_t e;
                    CU(cudaEventCreate(&e));
                    CU(cudaEventRecord(e, s.stream));
                    s.dbg_events.push_back(e);
                }
            }
            done += GRAPH_STEPS;
            passes += f2 ? GRAPH_STEPS / 2 : GRAPH_STEPS; // even: the graph's source lattice is `parity` again
            epochs += f2 ? GRAPH_STEPS / 2 : GRAPH_STEPS;
        }
    }
    // what is left (or everything, without graphs): plain launches, slabs interleaved pass by pass.  The ctrl words
    // already point at step `first + done` / epoch `epoch_base + epochs` (the graphs advance them).
    const int rest = iters - done;
    int rest_epochs = 0;
    if (rest > 0) {
        const int pairs = f2 ? rest / 2 : 0;
        for (int j = 0; j < pairs; j++)
            for (int i = 0; i < L->nslabs; i++) {
                Slab& s = L->slabs[i];
                CU(cudaSetDevice(s.device));
                int rc = launch_pair(L, s, static_cast<int>((parity + passes + j) & 1), 2 * j, j);
                if (rc) return rc;
            }
        const int singles = rest - 2 * pairs;
        for (int j = 0; j < singles; j++)
            for (int i = 0; i < L->nslabs; i++) {
                Slab& s = L->slabs[i];
                CU(cudaSetDevice(s.device));
                int rc = launch_step(L, s, static_cast<int>((parity + passes + pairs + j) & 1), 2 * pairs + j, pairs + j);
                if (rc) return rc;
            }
        rest_epochs = pairs + singles;
        passes += rest_epochs;
        epochs += rest_epochs;
        for (int i = 0; i < L->nslabs; i++) {
            Slab& s = L->slabs[i];
            CU(cudaSetDevice(s.device));
            advance_ctrl_kernel<<<1, 32, 0, s.stream>>>(s.ctrl, rest, rest_epochs);
            L->launches++;
            CU(cudaGetLastError());
        }
    }
    // slab 0's stop event is recorded after it has seen every other slab's stream finish
    for (int i = 1; i < L->nslabs; i++) {
        Slab& s = L->slabs[i];
        if (s.stream == L->slabs[0].stream) continue;
        CU(cudaSetDevice(s.device));
        CU(cudaEventRecord(s.ev1, s.stream));
        CU(cudaSetDevice(L->slabs[0].device));
        CU(cudaStreamWaitEvent(L->slabs[0].stream, s.ev1, 0));
    }
    CU(cudaSetDevice(L->slabs[0].device));
    CU(cudaEventRecord(L->slabs[0].ev1, L->slabs[0].stream));
    L->steps_done += iters;
    L->epoch = epoch_base + epochs;
    L->cur = static_cast<int>((parity + passes) & 1);
    poison.armed = false;
    return LBM_OK;
}

int lbm_sync(lbm_lattice_t* L)
{
    if (!L) return fail(LBM_EINVAL, "NULL lattice");
    for (int i = 0; i < L->nslabs; i++) {
        Slab& s = L->slabs[i];
        CU(cudaSetDevice(s.device));
        CU(cudaStreamSynchronize(s.stream));
    }
    return check_error_flags(L);
}

int lbm_tot_u_sums(lbm_lattice_t* L, long long* sums, long long* nonfinite, int iters)
{
    if (!L || (!sums && iters != 0)) return fail(LBM_EINVAL, "NULL argument");
    if (iters < 0 || iters > L->run_iters) return fail(LBM_EINVAL, "the last lbm_run call made %d steps, %d asked for", L->run_iters, iters);
    int rc = lbm_sync(L);
    if (rc) return rc;
    for (int t = 0; t < iters; t++) {
        sums[2 * t] = sums[2 * t + 1] = 0;
        if (nonfinite) nonfinite[t] = 0;
    }
    std::vector<unsigned long long> host;
    for (int i = 0; i < L->nslabs; i++) {
        Slab& s = L->slabs[i];
        CU(cudaSetDevice(s.device));
        const size_t n = static_cast<size_t>(iters) * s.nslots * SUM_WORDS;
        host.resize(n);
        if (n) CU(cudaMemcpy(host.data(), s.sums, n * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
        for (int t = 0; t < iters; t++)
            for (int j = 0; j < s.nslots; j++) {
                const unsigned long long* w = &host[(static_cast<size_t>(t) * s.nslots + j) * SUM_WORDS];
                sums[2 * t] += static_cast<long long>(w[0]);
                sums[2 * t + 1] += static_cast<long long>(w[1]);
                if (nonfinite) nonfinite[t] += static_cast<long long>(w[2]);
            }
    }
    return LBM_OK;
}

float lbm_av_from_sums(long long lo, long long hi, long long nonfinite, long long fluid_cells)
{
    if (nonfinite) return NAN;
    // total = lo + hi * 2^24 in units of 2^-40: exact in 128-bit integers, rounded once to double,
    // scaled (exactly) and rounded to the reference's float tot_u; then tot_u / (float)tot_cells as
    // SerialCode/d2q9-bgk.c:457
    const unsigned __int128 total = static_cast<unsigned __int128>(static_cast<unsigned long long>(lo)) +
                                    (static_cast<unsigned __int128>(static_cast<unsigned long long>(hi)) << FIX_SPLIT);
    const double hi64 = static_cast<double>(static_cast<unsigned long long>(total >> 64));
    const double lo64 = static_cast<double>(static_cast<unsigned long long>(total));
    const double tot = (hi64 * 18446744073709551616.0 + lo64) * (1.0 / 1099511627776.0);
    const float tot_u = static_cast<float>(tot);
    return tot_u / static_cast<float>(fluid_cells);
}

long long lbm_fluid_cells(const lbm_lattice_t* L)
{
    if (!L) return 0;
    long long n = 0;
    for (int i = 0; i < L->nslabs; i++) n += L->slabs[i].fluid;
    return n;
}

long long lbm_steps_done(const lbm_lattice_t* L) { return L ? L->steps_done : 0; }

int lbm_av_vels(lbm_lattice_t* L, float* av_vels, int iters)
{
    if (!L || (!av_vels && iters != 0)) return fail(LBM_EINVAL, "NULL argument");
    if (L->per_process && L->nranks > 1)
        return fail(LBM_EINVAL, "one slab per process: add lbm_tot_u_sums over the ranks and use lbm_av_from_sums");
    if (iters == 0) return LBM_OK; // maxIters = 0: the reference writes an empty av_vels.dat
    std::vector<long long> sums(2 * static_cast<size_t>(iters > 0 ? iters : 0)), bad(iters > 0 ? iters : 0);
    int rc = lbm_tot_u_sums(L, sums.data(), bad.data(), iters);
    if (rc) return rc;
    const long long fluid = lbm_fluid_cells(L);
    for (int t = 0; t < iters; t++) av_vels[t] = lbm_av_from_sums(sums[2 * t], sums[2 * t + 1], bad[t], fluid);
    return LBM_OK;
}

static int run_state_kernel(lbm_lattice_t* L, Slab& s, float* d_out[4], bool want_sums, bool want_density)
{
    StateArgs a;
    memset(&a, 0, sizeof a);
    const size_t pf = plane_floats(L, s);
    for (int k = 0; k < Q; k++) a.f[k] = s.lat[L->cur] + k * pf;
    a.obst = s.obst;
    a.nx = L->p.nx, a.rows = s.rows, a.pitch = L->pitch, a.opitch = L->opitch;
    a.density = L->p.density;
    a.u_x = d_out[0], a.u_y = d_out[1], a.u = d_out[2], a.pressure = d_out[3];
    CU(cudaMemsetAsync(s.state_sums, 0, (SUM_WORDS + 1) * sizeof(unsigned long long), s.stream));
    a.sums = want_sums ? s.state_sums : nullptr;
    a.density_sum = want_density ? reinterpret_cast<double*>(s.state_sums + SUM_WORDS) : nullptr;
    const size_t n = static_cast<size_t>(L->p.nx) * s.rows;
    state_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, s.stream>>>(a);
    L->launches++;
    CU(cudaGetLastError());
    return LBM_OK;
}

int lbm_av_velocity(lbm_lattice_t* L, float* av)
{
    if (!L || !av) return fail(LBM_EINVAL, "NULL argument");
    if (L->per_process && L->nranks > 1) return fail(LBM_EINVAL, "lbm_av_velocity needs a single-process lattice");
    long long lo = 0, hi = 0, bad = 0;
    for (int i = 0; i < L->nslabs; i++) {
        Slab& s = L->slabs[i];
        CU(cudaSetDevice(s.device));
        float* none[4] = {nullptr, nullptr, nullptr, nullptr};
        int rc = run_state_kernel(L, s, none, true, false);
        if (rc) return rc;
        unsigned long long w[SUM_WORDS];
        CU(cudaMemcpyAsync(w, s.state_sums, sizeof w, cudaMemcpyDeviceToHost, s.stream));
        CU(cudaStreamSynchronize(s.stream));
        lo += static_cast<long long>(w[0]), hi += static_cast<long long>(w[1]), bad += static_cast<long long>(w[2]);
    }
    *av = lbm_av_from_sums(lo, hi, bad, lbm_fluid_cells(L));
    return LBM_OK;
}

int lbm_total_density(lbm_lattice_t* L, double* total)
{
    if (!L || !total) return fail(LBM_EINVAL, "NULL argument");
    double t = 0.0;
    for (int i = 0; i < L->nslabs; i++) {
        Slab& s = L->slabs[i];
        CU(cudaSetDevice(s.device));
        float* none[4] = {nullptr, nullptr, nullptr, nullptr};
        int rc = run_state_kernel(L, s, none, false, true);
        if (rc) return rc;
        double d = 0.0;
        CU(cudaMemcpyAsync(&d, s.state_sums + SUM_WORDS, sizeof d, cudaMemcpyDeviceToHost, s.stream));
        CU(cudaStreamSynchronize(s.stream));
        t += d;
    }
    *total = t;
    return LBM_OK;
}

int lbm_final_state(lbm_lattice_t* L, float* u_x, float* u_y, float* u, float* pressure)
{
    if (!L) return fail(LBM_EINVAL, "NULL lattice");
    int rc = lbm_sync(L);
    if (rc) return rc;
    float* host[4] = {u_x, u_y, u, pressure};
    const int base_row = L->slabs[0].row0;
    for (int i = 0; i < L->nslabs; i++) {
        Slab& s = L->slabs[i];
        CU(cudaSetDevice(s.device));
        const size_t n = static_cast<size_t>(L->p.nx) * s.rows;
        // scratch: the lattice that does NOT hold the current state (9 planes of rows*pitch >= 4 planes of
        // rows*nx floats); the next lbm_run overwrites it anyway.  No allocation on this path.
        float* scratch = s.lat[L->cur ^ 1];
        float* dev[4] = {nullptr, nullptr, nullptr, nullptr};
        for (int q = 0; q < 4; q++)
            if (host[q]) dev[q] = scratch + static_cast<size_t>(q) * n;
        rc = run_state_kernel(L, s, dev, false, false);
        const size_t off = static_cast<size_t>(s.row0 - base_row) * L->p.nx;
        for (int q = 0; q < 4 && !rc; q++)
            if (host[q]) {
                cudaError_t e = cudaMemcpyAsync(host[q] + off, dev[q], n * sizeof(float), cudaMemcpyDeviceToHost, s.stream);
                if (e != cudaSuccess) rc = fail(LBM_ECUDA, "final-state download failed: %s", cudaGetErrorString(e));
            }
        cudaError_t es = cudaStreamSynchronize(s.stream);
        if (!rc && es != cudaSuccess) rc = fail(LBM_ECUDA, "final-state download failed: %s", cudaGetErrorString(es));
        if (rc) return rc;
    }
    return LBM_OK;
}

static int layout_copy(lbm_lattice_t* L, lbm_speed_t* cells, bool download)
{
    int rc = lbm_sync(L);
    if (rc) return rc;
    const int base_row = L->slabs[0].row0;
    for (int i = 0; i < L->nslabs; i++) {
        Slab& s = L->slabs[i];
        CU(cudaSetDevice(s.device));
        const size_t n = static_cast<size_t>(L->p.nx) * s.rows;
        // AoS staging in the lattice that does not hold the current state (9*rows*pitch >= 9*rows*nx floats)
        float* aos = s.lat[L->cur ^ 1];
        LayoutArgs a;
        const size_t pf = plane_floats(L, s);
        for (int k = 0; k < Q; k++) a.f[k] = s.lat[L->cur] + k * pf;
        a.aos = aos;
        a.nx = L->p.nx, a.rows = s.rows, a.pitch = L->pitch;
        float* host = reinterpret_cast<float*>(cells) + static_cast<size_t>(s.row0 - base_row) * L->p.nx * Q;
        cudaError_t e;
        if (download) {
            soa_to_aos_kernel<<<1184, 256, 0, s.stream>>>(a);
            e = cudaMemcpyAsync(host, aos, n * Q * sizeof(float), cudaMemcpyDeviceToHost, s.stream);
        } else {
            e = cudaMemcpyAsync(aos, host, n * Q * sizeof(float), cudaMemcpyHostToDevice, s.stream);
            aos_to_soa_kernel<<<1184, 256, 0, s.stream>>>(a);
        }
        L->launches++;
        cudaError_t e2 = cudaStreamSynchronize(s.stream);
        if (e != cudaSuccess || e2 != cudaSuccess)
            return fail(LBM_ECUDA, "cell %s failed: %s", download ? "download" : "upload",
                        cudaGetErrorString(e != cudaSuccess ? e : e2));
    }
    return LBM_OK;
}

int lbm_download_cells(lbm_lattice_t* L, lbm_speed_t* cells)
{
    if (!L || !cells) return fail(LBM_EINVAL, "NULL argument");
    return layout_copy(L, cells, true);
}

int lbm_upload_cells(lbm_lattice_t* L, const lbm_speed_t* cells)
{
    if (!L || !cells) return fail(LBM_EINVAL, "NULL argument");
    if (!L->connected) return fail(LBM_EINVAL, "lbm_halo_connect has not been called");
    int rc = layout_copy(L, const_cast<lbm_speed_t*>(cells), false);
    if (rc) return rc;
    if (uses_halo(L))
        for (int i = 0; i < L->nslabs; i++) {
            rc = push_boundary_rows(L, L->slabs[i]);
            if (rc) return rc;
        }
    return LBM_OK;
}

int lbm_last_run_ms(lbm_lattice_t* L, float* ms)
{
    if (!L || !ms) return fail(LBM_EINVAL, "NULL argument");
    int rc = lbm_sync(L);
    if (rc) return rc;
    if (L->run_iters == 0) {
        *ms = 0.f;
        return LBM_OK;
    }
    CU(cudaSetDevice(L->slabs[0].device));
    CU(cudaEventElapsedTime(ms, L->slabs[0].ev0, L->slabs[0].ev1));
    if (!L->slabs[0].dbg_events.empty()) {
        // LBM_DEBUG: device time of every graph replay (a gap = the host was late with the next launch)
        fprintf(stderr, "[lbm] graph replays (ms):");
        cudaEvent_t prev = L->slabs[0].ev0;
        for (cudaEvent_t e : L->slabs[0].dbg_events) {
            float t = 0.f;
            cudaEventElapsedTime(&t, prev, e);
            fprintf(stderr, " %.2f", t);
            prev = e;
        }
        fprintf(stderr, "\n");
        for (cudaEvent_t e : L->slabs[0].dbg_events) cudaEventDestroy(e);
        L->slabs[0].dbg_events.clear();
    }
    return LBM_OK;
}

long long lbm_kernel_launches(const lbm_lattice_t* L) { return L ? L->launches : 0; }

int lbm_selftest(int device, unsigned long long pairs, unsigned long long seed, unsigned long long* mismatches)
{
    if (!mismatches) return fail(LBM_EINVAL, "NULL argument");
    int rc = check_device_available();
    if (rc) return rc;
    CU(cudaSetDevice(device));
    unsigned long long* d = nullptr;
    CU(cudaMalloc(&d, 2 * sizeof(unsigned long long)));
    CU(cudaMemset(d, 0, 2 * sizeof(unsigned long long)));
    const unsigned blocks = 148 * 8, threads = 256;
    const unsigned long long per_thread = (pairs + static_cast<unsigned long long>(blocks) * threads - 1) / (static_cast<unsigned long long>(blocks) * threads);
    selftest_kernel<<<blocks, threads>>>(per_thread, seed, d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e == cudaSuccess) e = cudaMemcpy(mismatches, d, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (e != cudaSuccess) return fail(LBM_ECUDA, "self-test kernel failed: %s", cudaGetErrorString(e));
    return LBM_OK;
}

int lbm_selftest_collide(int device, int arith, unsigned long long sets, unsigned long long seed, unsigned long long* mismatches)
{
    if (!mismatches) return fail(LBM_EINVAL, "NULL argument");
    if (arith != LBM_ARITH_STRICT && arith != LBM_ARITH_FAST) return fail(LBM_EINVAL, "unknown arithmetic flavour %d", arith);
    int rc = check_device_available();
    if (rc) return rc;
    CU(cudaSetDevice(device));
    unsigned long long* d = nullptr;
    CU(cudaMalloc(&d, 2 * sizeof(unsigned long long)));
    CU(cudaMemset(d, 0, 2 * sizeof(unsigned long long)));
    const unsigned blocks = 148 * 4, threads = 128;
    const unsigned long long per_thread = (sets + static_cast<unsigned long long>(blocks) * threads - 1) / (static_cast<unsigned long long>(blocks) * threads);
    if (arith == LBM_ARITH_STRICT) selftest_collide_kernel<true><<<blocks, threads>>>(per_thread, seed, 1.85f, d);
    else selftest_collide_kernel<false><<<blocks, threads>>>(per_thread, seed, 1.85f, d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e == cudaSuccess) e = cudaMemcpy(mismatches, d, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (e != cudaSuccess) return fail(LBM_ECUDA, "collision self-test kernel failed: %s", cudaGetErrorString(e));
    return LBM_OK;
}

int lbm_num_slabs(const lbm_lattice_t* L) { return L ? L->nslabs : 0; }

int lbm_slab_info(const lbm_lattice_t* L, int i, int* row0, int* row1, int* device)
{
    if (!L || i < 0 || i >= L->nslabs) return fail(LBM_EINVAL, "bad slab index");
    if (row0) *row0 = L->slabs[i].row0;
    if (row1) *row1 = L->slabs[i].row1;
    if (device) *device = L->slabs[i].device;
    return LBM_OK;
}

void lbm_destroy(lbm_lattice_t* L)
{
    if (!L) return;
    if (L->slabs) {
        for (int i = 0; i < L->nslabs; i++) {
            cudaSetDevice(L->slabs[i].device);
            cudaDeviceSynchronize();
        }
        cudaGetLastError();
        for (int i = 0; i < L->nslabs; i++) free_slab(L->slabs[i]);
        delete[] L->slabs;
    }
    delete L;
}

} // extern "C"
