// lbm_fused2_kernel.cuh -- TWO timesteps per pass over HBM (temporal blocking of the step kernel).
//
// One launch advances the slab from step t to step t+2: every population is read from HBM once and written
// once per TWO lattice updates (36 B per update instead of the one-pass 72 B), the intermediate step lives in
// shared memory only.  Arithmetic, operation order and the |u| sums of both steps are exactly those of
// step_tma_kernel / step_vec4_kernel applied twice (update_cell / accelerate_cell / acc_speed), so the strict
// flavour stays bit-identical to SerialCode/d2q9-bgk.c:207-458 after every pair of steps.
//
// Decomposition.  The slab is cut into column strips of F2_CORE = 120 cells and row segments; a work unit is
// (strip, rows [ya, yb)).  For a unit the CTA marches in y ("sliding window"):
//   phase 0  step t -> t+1 on the 128 columns [x0-4, x0+124) of row r1 (the strip plus what step t+2 will pull
//            from its x neighbours); sources are the nine planes staged by TMA exactly as in step_tma_kernel (the
//            y part of the pull shift done by the copy engine, the x part on the way out of shared memory);
//            results go to a ring of R+2 rows in shared memory (`buf2`);
//   phase 1  step t+1 -> t+2 on the 120 core columns of row r2 = r1 - 1, pulling from the ring; results go to
//            HBM with 128-bit stores.
// R warps take one row each per iteration (4 cells per lane); a CTA barrier separates the phases.  Marching
// means the redundant work is the x apron only (128 of 120 columns in phase 0, 30 of 32 lanes in phase 1) plus
// two rows per segment.  NSTAGES stages of SROWS rows are in flight; thread 0 refills a stage as soon as the
// barrier after phase 0 says every warp has its row in registers (no producer warp: 16 warps per SM are four
// per scheduler and leave 128 registers per thread).
//
// Rows near the slab's first / last row (output rows 0,1 and rows-2,rows-1) are "boundary units": their step-t
// source rows include rows -2,-1 / rows,rows+1, i.e. the periodic wrap of a single slab or -- with several
// slabs -- the two rows the neighbouring GPU delivered into this slab's halo ring.  The CTA loads those
// stages with ordinary loads (wrapping x itself) after waiting for the neighbour's flag, and
// also stores the rows the neighbour will need into ITS ring over NVLink, and the last CTA to finish a side
// bumps the neighbour's flag.  Boundary units come first in the unit order, so the exchange for step t+2
// overlaps the interior work.  This replaces, for pairs of steps, MPI_Isend/Irecv/Waitall of
// MPI_Waitall/d2q9-bgk.c:225-253 with a two-row halo: the neighbour's boundary row at step t+1 is recomputed
// locally from (row -1: planes 0,1,3 and the three planes that cross towards me; row -2: the three crossing planes).
#pragma once

#include "lbm_tma_kernel.cuh"

namespace lbm {

constexpr int F2_CORE = 120;          // columns of a strip that a unit finally writes (30 lanes x 4 cells)
constexpr int F2_W1 = 128;            // columns of a strip computed for the intermediate step
constexpr int F2_B2ROW = Q * F2_W1;   // floats per ring row of intermediate results: [plane][128]

struct Fused2Args {
    const float* in;             // plane 0 of the source lattice (state after step t-1)
    float* out;                  // plane 0 of the destination lattice (state after step t+1)
    size_t pf;                   // floats per plane
    HaloCfg h;
    const uint32_t* obst;        // [rows][opitch]
    const uint32_t* obst_halo;   // several slabs: [2][opitch] obstacle bits of the row south of row 0 / north of row rows-1
    const int* ctrl;             // [0] step index of step_offset 0, [1] first step held by sums[], [2] last step of the run,
                                 // [3] halo epoch of epoch_offset 0
    unsigned long long* const* sums_ref;
    int nslots, step_offset, epoch_offset;
    int nx, rows, pitch, opitch;
    int nsx;                     // strips
    int seg_h, nseg;             // interior units: rows [2 + s*seg_h, min(2 + (s+1)*seg_h, rows-2))
    int nunits;                  // 2*nsx boundary units, then nseg*nsx interior units (x fastest)
    int accel_row;
    float omega, w1a, w2a;
};

// position of plane k inside its triple (0,1,3) / (2,5,6) / (4,7,8)
__host__ __device__ constexpr int triple_index(int k) { return (k == 0 || k == 2 || k == 4) ? 0 : ((k == 1 || k == 5 || k == 7) ? 1 : 2); }

__device__ __forceinline__ float4 lds4(uint32_t addr)
{
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ float lds1(uint32_t addr)
{
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts4(uint32_t addr, float a, float b, float c, float d)
{
    asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// update4() for this kernel: collide4, the |u| of the four cells straight into one 64-bit total (any split into the
// two words the host adds up is equivalent), accelerate_flow() at store time on the driven row.  Cells whose |u| is
// not finite (never in a healthy run) are counted straight into shared memory: no register for them.
template <bool STRICT>
__device__ __forceinline__ void update4_total(const float (&t)[Q][4], uint32_t obits, bool counted, bool accel, float omega, float w1a,
                                              float w2a, float (&o)[Q][4], unsigned long long& total, unsigned long long* bad_counter)
{
    float speed[4];
    collide4<STRICT, false>(t, obits, omega, o, speed);
    unsigned nbad = 0u;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const bool cnt = counted && !((obits >> j) & 1u);
        const bool bad = !(speed[j] < FIX_LIMIT);
        total += (bad || !cnt) ? 0ull : __float2ull_rn(speed[j] * FIX_SCALE);
        nbad += (bad && cnt) ? 1u : 0u;
    }
    if (nbad) atomicAdd(bad_counter, static_cast<unsigned long long>(nbad));
    if (accel) {
#pragma unroll
        for (int j = 0; j < 4; j++) {
            float oc[Q];
#pragma unroll
            for (int k = 0; k < Q; k++) oc[k] = o[k][j];
            accelerate_cell(oc, (obits >> j) & 1u, w1a, w2a);
#pragma unroll
            for (int k = 0; k < Q; k++) o[k][j] = oc[k];
        }
    }
}

// source row of plane K for lattice row g of a boundary unit: a lattice row, the periodic wrap of a single
// slab, or an entry of the halo ring the neighbouring GPU filled (near row: planes 0,1,3 -> entries 0..2, the
// three planes crossing towards this slab -> 3..5; far row: the crossing planes -> 6..8)
template <int K>
__device__ __forceinline__ const float* f2_source_row(const Fused2Args& a, int g, int epoch, bool& from_ring)
{
    from_ring = false;
    const size_t pitch = a.pitch;
    if (g < 0) {
        if (a.h.on) {
            from_ring = true;
            const int e = (g == -2) ? 6 + triple_index(K) : (dir_cy(K) == 0 ? triple_index(K) : 3 + triple_index(K));
            return a.h.hs.recv_ring + static_cast<size_t>(ring_slot(epoch, a.h.ring)) * a.h.slot_stride + e * pitch;
        }
        g += a.rows;
    } else if (g >= a.rows) {
        if (a.h.on) {
            from_ring = true;
            const int e = (g == a.rows + 1) ? 6 + triple_index(K) : (dir_cy(K) == 0 ? triple_index(K) : 3 + triple_index(K));
            return a.h.hn.recv_ring + static_cast<size_t>(ring_slot(epoch, a.h.ring)) * a.h.slot_stride + e * pitch;
        }
        g -= a.rows;
    }
    return a.in + K * a.pf + static_cast<size_t>(g) * pitch;
}

template <int K, int SROWS>
__device__ __forceinline__ void f2_load_plane_generic(const Fused2Args& a, float* stage, int x0, int ya, int epoch, int lane)
{
    constexpr int WK = plane_width(K);
    const int xs = (dir_cx(K) == 1) ? x0 - 8 : x0 - 4;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        bool ring;
        const float* src = f2_source_row<K>(a, ya - 1 + j - dir_cy(K), epoch, ring);
        float* dst = stage + plane_offset(K, SROWS) + j * WK;
#pragma unroll
        for (int i0 = 0; i0 < WK / 4; i0 += 32) {
            const int i = i0 + lane;
            if (i < WK / 4) {
                int x = xs + 4 * i;
                x = x < 0 ? x + a.nx : (x >= a.nx ? x - a.nx : x); // periodic in x (SerialCode:259-262); nx % 4 == 0
                const float4 v = ring ? __ldcg(reinterpret_cast<const float4*>(src + x)) : __ldg(reinterpret_cast<const float4*>(src + x));
                *reinterpret_cast<float4*>(dst + 4 * i) = v;
            }
        }
    }
}

struct F2Unit {
    int kind; // 0: rows 0,1 (south boundary), 1: rows rows-2,rows-1 (north boundary), 2: interior segment
    int x0, ya, yb;
};
__device__ __forceinline__ F2Unit f2_unit(const Fused2Args& a, int u)
{
    F2Unit r;
    if (u < a.nsx) {
        r.kind = 0, r.x0 = u * F2_CORE, r.ya = 0, r.yb = 2;
    } else if (u < 2 * a.nsx) {
        r.kind = 1, r.x0 = (u - a.nsx) * F2_CORE, r.ya = a.rows - 2, r.yb = a.rows;
    } else {
        const int v = u - 2 * a.nsx;
        const int seg = v / a.nsx;
        r.kind = 2, r.x0 = (v - seg * a.nsx) * F2_CORE;
        r.ya = 2 + seg * a.seg_h;
        r.yb = min(r.ya + a.seg_h, a.rows - 2);
    }
    return r;
}

// boundary units of a slab with neighbours: the four cells just stored (row y of the slab's first / last two rows) go
// straight into the ring of the neighbour that needs them for ITS next pair of steps (peer memory over NVLink)
__device__ __forceinline__ void f2_push_rows(const Fused2Args& a, bool south, int y, int xg, int epoch, const float (&o)[Q][4])
{
    const size_t pitch = a.pitch;
    const bool near = south ? (y == 0) : (y == a.rows - 1);
    float* ring = (south ? a.h.hs.send_ring : a.h.hn.send_ring) + static_cast<size_t>(ring_slot(epoch + 1, a.h.ring)) * a.h.slot_stride + xg;
    // planes crossing towards the neighbour: 4,7,8 southwards, 2,5,6 northwards
    const float4 ca = south ? make_float4(o[4][0], o[4][1], o[4][2], o[4][3]) : make_float4(o[2][0], o[2][1], o[2][2], o[2][3]);
    const float4 cb = south ? make_float4(o[7][0], o[7][1], o[7][2], o[7][3]) : make_float4(o[5][0], o[5][1], o[5][2], o[5][3]);
    const float4 cc = south ? make_float4(o[8][0], o[8][1], o[8][2], o[8][3]) : make_float4(o[6][0], o[6][1], o[6][2], o[6][3]);
    if (near) {
        *reinterpret_cast<float4*>(ring + 0 * pitch) = make_float4(o[0][0], o[0][1], o[0][2], o[0][3]);
        *reinterpret_cast<float4*>(ring + 1 * pitch) = make_float4(o[1][0], o[1][1], o[1][2], o[1][3]);
        *reinterpret_cast<float4*>(ring + 2 * pitch) = make_float4(o[3][0], o[3][1], o[3][2], o[3][3]);
        *reinterpret_cast<float4*>(ring + 3 * pitch) = ca;
        *reinterpret_cast<float4*>(ring + 4 * pitch) = cb;
        *reinterpret_cast<float4*>(ring + 5 * pitch) = cc;
    } else {
        *reinterpret_cast<float4*>(ring + 6 * pitch) = ca;
        *reinterpret_cast<float4*>(ring + 7 * pitch) = cb;
        *reinterpret_cast<float4*>(ring + 8 * pitch) = cc;
    }
}

// dispatch f2_load_plane_generic on a run-time plane index
template <int SROWS>
__device__ __forceinline__ void f2_load_plane_any(int k, const Fused2Args& a, float* stage, int x0, int ya, int epoch, int lane)
{
    switch (k) {
    case 0: f2_load_plane_generic<0, SROWS>(a, stage, x0, ya, epoch, lane); break;
    case 1: f2_load_plane_generic<1, SROWS>(a, stage, x0, ya, epoch, lane); break;
    case 2: f2_load_plane_generic<2, SROWS>(a, stage, x0, ya, epoch, lane); break;
    case 3: f2_load_plane_generic<3, SROWS>(a, stage, x0, ya, epoch, lane); break;
    case 4: f2_load_plane_generic<4, SROWS>(a, stage, x0, ya, epoch, lane); break;
    case 5: f2_load_plane_generic<5, SROWS>(a, stage, x0, ya, epoch, lane); break;
    case 6: f2_load_plane_generic<6, SROWS>(a, stage, x0, ya, epoch, lane); break;
    case 7: f2_load_plane_generic<7, SROWS>(a, stage, x0, ya, epoch, lane); break;
    default: f2_load_plane_generic<8, SROWS>(a, stage, x0, ya, epoch, lane); break;
    }
}

// tmap / tmapw: boxes of TMA_TX x SROWS x 1 and TMA_TXW x SROWS x 1 over the SOURCE lattice.
// R warps, no dedicated producer: thread 0 refills the stages an iteration has consumed right after the barrier
// that ends its phase 0 (every warp then holds its row in registers), so the copy of iteration c + NSTAGES*SROWS/R
// runs behind phase 1 of iteration c and everything after it.  16 warps per SM (2 CTAs of 8, or 1 of 16) is four per
// scheduler: 128 registers per thread.
template <bool STRICT, int R, int SROWS, int NSTAGES, int MINB>
__global__ void __launch_bounds__(32 * R, MINB)
    step2_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ CUtensorMap tmapw, const Fused2Args a)
{
    static_assert(R % SROWS == 0 && SROWS >= 4, "a stage holds a whole number of warps' rows and at least a boundary unit");
    constexpr int RB = R + 2;                          // ring rows of intermediate results
    constexpr int STAGE = stage_floats(SROWS);         // floats per stage
    constexpr uint32_t STAGE_BYTES = stage_tx_bytes(SROWS);
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* stages = reinterpret_cast<float*>(smem_raw);            // [NSTAGES][STAGE]
    const uint32_t stages_s = smem_u32(stages);
    const uint32_t buf2_s = stages_s + NSTAGES * STAGE * 4;        // [RB][Q][128] floats (+ a few floats of slack)
    __shared__ __align__(8) uint64_t full_bar[NSTAGES];
    __shared__ unsigned long long s_acc[2][3];
    __shared__ unsigned s_readers[NSTAGES];            // warps that have read the stage into registers

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < NSTAGES; s++) mbar_init(&full_bar[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < 6) s_acc[tid / 3][tid % 3] = 0ull;
    if (tid < NSTAGES) s_readers[tid] = 0u;
    __syncthreads();

    const int t1 = a.ctrl[0] + a.step_offset;          // the first of the two steps
    const bool live2 = (t1 + 1) != a.ctrl[2];          // the run's last step is not followed by accelerate_flow
    const int epoch = a.ctrl[3] + a.epoch_offset;
    const size_t pitch = a.pitch;

    unsigned long long acc_a = 0ull, acc_b = 0ull;     // per-thread |u| totals (units of 2^-40) of step t1 / t1+1
    int nbase = 0;                                     // TMA stages consumed by this CTA's earlier units
    bool primed = false;                               // the TMA pipeline has been started
    // request stage number `seq` of this CTA's stage sequence (the current unit starts at sequence number
    // nbase_cur and has nst_cur stages; the units after it are u_cur + gridDim.x, ...)
    auto refill = [&](int seq, int u_cur, int nbase_cur, int nst_cur) {
        int i = seq - nbase_cur, uu = u_cur, nn = nst_cur;
        F2Unit un = f2_unit(a, uu);
        while (i >= nn) {
            i -= nn;
            uu += gridDim.x;
            if (uu >= a.nunits) return;
            un = f2_unit(a, uu);
            nn = (un.yb - un.ya + 2 + SROWS - 1) / SROWS;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); // the buffer was last touched by ordinary loads / stores
        const int s = seq % NSTAGES;
        mbar_expect_tx(&full_bar[s], STAGE_BYTES);
        float* dst = stages + s * STAGE;
        const int ys = un.ya - 1 + i * SROWS;
#pragma unroll
        for (int k = 0; k < Q; k++)
            tma_load_3d(dst + plane_offset(k, SROWS), dir_cx(k) == 0 ? &tmap : &tmapw, &full_bar[s],
                        dir_cx(k) == 1 ? un.x0 - 8 : un.x0 - 4, ys - dir_cy(k), k);
    };
    for (int u = blockIdx.x; u < a.nunits; u += gridDim.x) {
        const F2Unit un = f2_unit(a, u);
        const int x0 = un.x0, ya = un.ya, yb = un.yb;
        const int rows1 = yb - ya + 2;                       // intermediate rows ya-1 .. yb
        const int nst = (rows1 + SROWS - 1) / SROWS;
        const int nc = (rows1 + R - 1) / R;
        const bool tma_unit = un.kind == 2;
        if (!tma_unit) {
            // boundary unit (always ahead of this CTA's interior units): four rows into stage buffer 0 with ordinary
            // loads, after the neighbour has delivered the rows of this epoch
            if (a.h.on && a.h.wait && tid == 0) halo_wait(a.h, epoch, un.kind == 0, un.kind == 1);
            __syncthreads();
            for (int k = warp; k < Q; k += R) f2_load_plane_any<SROWS>(k, a, stages, x0, ya, epoch, lane);
            __syncthreads();
        } else if (!primed) {
            primed = true;
            __syncthreads(); // nobody reads the boundary units' buffer any more
            if (tid == 0) {
                for (int j = 0; j < NSTAGES; j++) refill(j, u, 0, nst);
            }
        }
        // x wrap of TMA-staged strips: the copy engine zero-fills columns outside [0, nx); the intermediate
        // cells x = -1 and x = nx (periodic images of nx-1 and 0) are needed by step t+2, so the lanes that own
        // them (and their inner neighbours) fetch the wrapped populations themselves
        const bool west = tma_unit && (x0 == 0);
        const bool east = tma_unit && (x0 + F2_CORE >= a.nx);
        const int le = (a.nx - x0 + 4) >> 2;                 // lane whose first cell is x = nx
        const bool pw0 = west && lane == 0, pw1 = west && lane == 1;
        const bool pe0 = east && lane == le, pe1 = east && lane == le - 1;
        const bool push_halo = !tma_unit && a.h.on;

        for (int c = 0; c < nc; c++) {
            const int q = c * R + warp;                      // this warp's intermediate row, relative to ya-1
#pragma unroll 1
            for (int phase = 0; phase < 2; phase++) {
                const int rel = q - phase;                   // row of this phase, relative to ya-1
                const int y = ya - 1 + rel;                  // slab row (-1 / rows possible in phase 0 of boundary units)
                const int xg = x0 + 4 * lane - (phase ? 0 : 4);
                const bool active = phase ? (rel >= 1 && y < yb) : (q < rows1);
                float4 v[Q];
                float sh[Q];
                float pa[3] = {0.f, 0.f, 0.f}, pb[3] = {0.f, 0.f, 0.f};
                uint32_t oword = 0u;
                const int xo = xg < 0 ? xg + a.nx : (xg >= a.nx ? xg - a.nx : xg);
                if (active) {
                    const uint32_t* orow = (y < 0) ? (a.h.on ? a.obst_halo : a.obst + static_cast<size_t>(a.rows - 1) * a.opitch)
                                                   : (y >= a.rows ? (a.h.on ? a.obst_halo + a.opitch : a.obst)
                                                                  : a.obst + static_cast<size_t>(y) * a.opitch);
                    oword = __ldg(orow + (xo >> 5));
                }
                if (phase == 0) {
                    int s = 0;
                    if (tma_unit && q < nst * SROWS) {
                        const int n = nbase + q / SROWS;
                        s = n % NSTAGES;
                        if (active && (west || east)) {
                            // rows of the triples' members: cy = 0, +1, -1  ->  y, y-1, y+1
                            const size_t r0 = static_cast<size_t>(y) * pitch, rm = r0 - pitch, rp = r0 + pitch;
                            if (pw0 || pw1) {
                                const size_t col = pw0 ? a.nx - 2 : a.nx - 1;
                                pa[0] = __ldg(a.in + 1 * a.pf + r0 + col), pa[1] = __ldg(a.in + 5 * a.pf + rm + col),
                                pa[2] = __ldg(a.in + 8 * a.pf + rp + col);
                                if (pw0) {
                                    const size_t c2 = a.nx - 1;
                                    pb[0] = __ldg(a.in + 0 * a.pf + r0 + c2), pb[1] = __ldg(a.in + 2 * a.pf + rm + c2),
                                    pb[2] = __ldg(a.in + 4 * a.pf + rp + c2);
                                }
                            }
                            if (pe0 || pe1) {
                                const size_t col = pe0 ? 1 : 0;
                                pa[0] = __ldg(a.in + 3 * a.pf + r0 + col), pa[1] = __ldg(a.in + 6 * a.pf + rm + col),
                                pa[2] = __ldg(a.in + 7 * a.pf + rp + col);
                                if (pe0) {
                                    pb[0] = __ldg(a.in + 0 * a.pf + r0), pb[1] = __ldg(a.in + 2 * a.pf + rm),
                                    pb[2] = __ldg(a.in + 4 * a.pf + rp);
                                }
                            }
                        }
                        mbar_wait(&full_bar[s], (n / NSTAGES) & 1);
                    }
                    if (active) {
                        const uint32_t st = stages_s + (s * STAGE) * 4;
                        const int rs = q % SROWS;
#pragma unroll
                        for (int k = 0; k < Q; k++) {
                            const uint32_t row = st + (plane_offset(k, SROWS) + rs * plane_width(k) + 4 * lane) * 4;
                            if (dir_cx(k) == 0) {
                                v[k] = lds4(row);
                                sh[k] = 0.f;
                            } else if (dir_cx(k) == 1) { // staged four columns further west
                                v[k] = lds4(row + 16);
                                sh[k] = lds1(row + 12);
                            } else {
                                v[k] = lds4(row);
                                sh[k] = lds1(row + 16);
                            }
                        }
                    }
                    if (tma_unit && q < nst * SROWS) {
                        // this warp's row of the stage is on its way into registers (a warp's shared-memory
                        // instructions are performed in order, so the counter below is bumped after the loads above
                        // have read the buffer): the LAST of the stage's SROWS reader warps asks the copy engine for
                        // the stage that will reuse the buffer -- a full iteration before it is needed
                        __syncwarp();
                        if (lane == 0) {
                            const int n = nbase + q / SROWS, sb = n % NSTAGES;
                            if (atomicAdd(&s_readers[sb], 1u) == SROWS - 1) {
                                s_readers[sb] = 0u;
                                refill(n + NSTAGES, u, nbase, nst);
                            }
                        }
                    }
                }
                if (phase == 1 && active) {
                    // ring rows of the intermediate step: plane k comes from row rel - cy_k, column x - cx_k
                    const uint32_t r0 = buf2_s + ((rel % RB) * F2_B2ROW + 4 * lane) * 4;
                    const uint32_t rm = buf2_s + (((rel - 1) % RB) * F2_B2ROW + 4 * lane) * 4;
                    const uint32_t rp = buf2_s + (((rel + 1) % RB) * F2_B2ROW + 4 * lane) * 4;
#pragma unroll
                    for (int k = 0; k < Q; k++) {
                        const uint32_t row = (dir_cy(k) == 0 ? r0 : (dir_cy(k) == 1 ? rm : rp)) + k * F2_W1 * 4;
                        v[k] = lds4(row + 16);
                        sh[k] = dir_cx(k) == 0 ? 0.f : (dir_cx(k) == 1 ? lds1(row + 12) : lds1(row + 32));
                    }
                }

                if (active) {
                    float t[Q][4];
#pragma unroll
                    for (int k = 0; k < Q; k++) {
                        const float e[6] = {sh[k], v[k].x, v[k].y, v[k].z, v[k].w, sh[k]};
#pragma unroll
                        for (int j = 0; j < 4; j++) t[k][j] = e[j + 1 - dir_cx(k)]; // cell xg+j pulls column xg+j-cx
                    }
                    if (phase == 0 && (west || east)) {
                        if (pw0) t[1][3] = pa[0], t[5][3] = pa[1], t[8][3] = pa[2], t[0][3] = pb[0], t[2][3] = pb[1], t[4][3] = pb[2];
                        if (pw1) t[1][0] = pa[0], t[5][0] = pa[1], t[8][0] = pa[2];
                        if (pe0) t[3][0] = pa[0], t[6][0] = pa[1], t[7][0] = pa[2], t[0][0] = pb[0], t[2][0] = pb[1], t[4][0] = pb[2];
                        if (pe1) t[3][3] = pa[0], t[6][3] = pa[1], t[7][3] = pa[2];
                    }
                    if (phase == 0 && (west || east)) {
                        // columns the copy engine zero-filled and nobody will read (x < -1, x > nx): give them a fluid at
                        // rest, or their 0 / 0 would drag the whole warp through the IEEE slow paths on every row
#pragma unroll
                        for (int j = 0; j < 4; j++) {
                            const int x = xg + j;
                            if (x < -1 || x > a.nx) {
                                t[0][j] = 0.04f;
#pragma unroll
                                for (int k = 1; k < Q; k++) t[k][j] = k < 5 ? 0.01f : 0.0025f;
                            }
                        }
                    }
                    const uint32_t obits = (oword >> (xo & 31)) & 0xfu;
                    // cells whose |u| this unit owns: its core columns, and in phase 0 its own rows only
                    const bool counted = phase ? (lane < 30 && xg < a.nx)
                                               : (y >= ya && y < yb && lane >= 1 && lane <= 30 && xg < a.nx);
                    const bool accel = (a.accel_row >= 0) && (y == a.accel_row) && (phase == 0 || live2);
                    float o[Q][4];
                    unsigned long long tot = 0ull;
                    update4_total<STRICT>(t, obits, counted, accel, a.omega, a.w1a, a.w2a, o, tot, &s_acc[phase][2]);
                    if (phase) acc_b += tot;
                    else acc_a += tot;

                    if (phase == 0) {
                        const uint32_t dst = buf2_s + ((q % RB) * F2_B2ROW + 4 * lane) * 4;
#pragma unroll
                        for (int k = 0; k < Q; k++) sts4(dst + k * F2_W1 * 4, o[k][0], o[k][1], o[k][2], o[k][3]);
                    } else if (lane < 30 && xg < a.nx) {
                        float* dst = a.out + static_cast<size_t>(y) * pitch + xg;
#pragma unroll
                        for (int k = 0; k < Q; k++)
                            *reinterpret_cast<float4*>(dst + k * a.pf) = make_float4(o[k][0], o[k][1], o[k][2], o[k][3]);
                        if (push_halo) f2_push_rows(a, un.kind == 0, y, xg, epoch, o);
                    }
                }
                // phase 0 -> 1: the ring rows are complete and every warp holds its staged row in registers;
                // phase 1 -> next iteration: the ring rows may be overwritten
                __syncthreads();
            }
        }
        if (!tma_unit && a.h.on && tid == 0) halo_arrive(a.h, un.kind == 0, un.kind == 1, static_cast<unsigned>(a.nsx));
        if (tma_unit) nbase += nst;
    }

    // ---------------- |u| sums of both steps (s_acc[step][0], units of 2^-40) ----------------
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
        acc_a += __shfl_xor_sync(0xffffffffu, acc_a, s);
        acc_b += __shfl_xor_sync(0xffffffffu, acc_b, s);
    }
    if (lane == 0) {
        atomicAdd(&s_acc[0][0], acc_a);
        atomicAdd(&s_acc[1][0], acc_b);
    }
    __syncthreads();
    if (tid < 2) {
        // total = lo + hi * 2^24 (what the host forms); any split with the same total is equivalent
        unsigned long long* out =
            *a.sums_ref + (static_cast<size_t>(t1 + tid - a.ctrl[1]) * a.nslots + (blockIdx.x & (a.nslots - 1))) * SUM_WORDS;
        const unsigned long long tot = s_acc[tid][0];
        atomicAdd(&out[0], tot & ((1ull << FIX_SPLIT) - 1ull));
        atomicAdd(&out[1], tot >> FIX_SPLIT);
        if (s_acc[tid][2]) atomicAdd(&out[2], s_acc[tid][2]);
    }
}

// ---- start of every lbm_run on a lattice that advances in pairs of steps: (re)deliver this slab's two boundary
// rows to both neighbours.  Needed because accelerate_flow() of the run's first step has just changed row ny-2 in
// place (the last slab's second-to-last row, which the first slab reads as its far south halo row), and after
// lbm_upload_cells.  Counts as one halo epoch like a pair of steps: wait for the neighbours' previous epoch,
// store the nine ring entries per side, last CTA bumps the flags.
struct HaloPushArgs {
    const float* lat; // plane 0 of the current lattice
    size_t pf;
    HaloCfg h;
    const int* ctrl;
    int epoch_offset;
    int nx, rows, pitch;
};
__global__ void __launch_bounds__(128) halo_push_kernel(const HaloPushArgs a)
{
    const int epoch = a.ctrl[3] + a.epoch_offset;
    if (a.h.wait && threadIdx.x == 0) halo_wait(a.h, epoch, true, true);
    __syncthreads();
    const int x = 4 * (blockIdx.x * blockDim.x + threadIdx.x);
    if (x < a.nx) {
        const size_t pitch = a.pitch;
        const size_t slot = static_cast<size_t>(ring_slot(epoch + 1, a.h.ring)) * a.h.slot_stride;
        const size_t last = static_cast<size_t>(a.rows - 1) * pitch;
        auto ld = [&](int k, size_t off) { return __ldcg(reinterpret_cast<const float4*>(a.lat + k * a.pf + off + x)); };
        auto st = [&](float* ring, int e, float4 v) { *reinterpret_cast<float4*>(ring + slot + e * pitch + x) = v; };
        // row 0 / 1 -> the south neighbour's north ring: near 0,1,3 and 4,7,8; far 4,7,8
        st(a.h.hs.send_ring, 0, ld(0, 0)), st(a.h.hs.send_ring, 1, ld(1, 0)), st(a.h.hs.send_ring, 2, ld(3, 0));
        st(a.h.hs.send_ring, 3, ld(4, 0)), st(a.h.hs.send_ring, 4, ld(7, 0)), st(a.h.hs.send_ring, 5, ld(8, 0));
        st(a.h.hs.send_ring, 6, ld(4, pitch)), st(a.h.hs.send_ring, 7, ld(7, pitch)), st(a.h.hs.send_ring, 8, ld(8, pitch));
        // row rows-1 / rows-2 -> the north neighbour's south ring: near 0,1,3 and 2,5,6; far 2,5,6
        st(a.h.hn.send_ring, 0, ld(0, last)), st(a.h.hn.send_ring, 1, ld(1, last)), st(a.h.hn.send_ring, 2, ld(3, last));
        st(a.h.hn.send_ring, 3, ld(2, last)), st(a.h.hn.send_ring, 4, ld(5, last)), st(a.h.hn.send_ring, 5, ld(6, last));
        st(a.h.hn.send_ring, 6, ld(2, last - pitch)), st(a.h.hn.send_ring, 7, ld(5, last - pitch)), st(a.h.hn.send_ring, 8, ld(6, last - pitch));
    }
    __syncthreads();
    if (threadIdx.x == 0) halo_arrive(a.h, true, true, gridDim.x);
}

} // namespace lbm
