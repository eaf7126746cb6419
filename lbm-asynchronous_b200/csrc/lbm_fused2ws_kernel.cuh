// lbm_fused2ws_kernel.cuh -- step2_kernel (two timesteps per pass over HBM, lbm_fused2_kernel.cuh) with the two phases
// on DIFFERENT warps and no CTA barrier in the marching loop.
//
// step2_kernel runs phase 0 (step t -> t+1 into the shared-memory ring) and phase 1 (t+1 -> t+2 out of it) on the same
// warps, separated by __syncthreads(): every warp waits for the slowest of eight, twice per row (ncu: "barrier" 15 %
// of all stall samples, on a kernel that is bound by latency, not by issue slots or HBM).  Here a CTA of R warps is
// split into R/2 PRODUCER warps that only run phase 0 and R/2 CONSUMER warps that only run phase 1.  They meet in the
// ring through two mbarriers per ring row:
//   row_full[slot]   32 arrivals: the lanes of the producer warp that wrote the row;
//   row_empty[slot]  96 arrivals: the lanes of the three consumer tasks that read the row (as row r-1, r, r+1).
// A consumer task waits for its three rows, a producer waits for the slot it is about to overwrite; with R + 2 ring
// rows the producers may run up to R - 1 rows ahead, so neither side normally waits.  mbarrier.try_wait suspends the
// warp in hardware: no polling loop competes for issue slots (the rejected "flags" variant polled shared memory).
// Dependencies are acyclic (consumer task q needs producer rows <= q, producer row q needs consumer tasks <= q-R+1),
// and every wait targets the phase the waiter's own later arrival completes, so a barrier can never run two phases
// ahead of a waiter.
//
// The staging of step-t rows by TMA (one stage = the R/2 rows of one producer iteration, refilled by the last of its
// reader warps), the strip / segment decomposition, the x wrap, the arithmetic (collide4) and the boundary units
// (rows next to the slab edges: generic loads, halo wait, peer stores; they keep CTA barriers -- four rows per unit,
// ahead of the interior units) are step2_kernel's.
#pragma once

#include "lbm_fused2_kernel.cuh"

namespace lbm {

template <bool STRICT, int R, int NSTAGES, int MINB>
__global__ void __launch_bounds__(32 * R, MINB)
    step2ws_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ CUtensorMap tmapw, const Fused2Args a)
{
    constexpr int NP = R / 2;                          // producer warps = consumer warps = rows per stage
    constexpr int SROWS = NP;
    static_assert(R % 2 == 0 && SROWS >= 4, "a stage holds the rows of one producer iteration and at least a boundary unit");
    constexpr int RB = R + 2;                          // ring rows of intermediate results
    constexpr int STAGE = stage_floats(SROWS);         // floats per stage
    constexpr uint32_t STAGE_BYTES = stage_tx_bytes(SROWS);
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* stages = reinterpret_cast<float*>(smem_raw);            // [NSTAGES][STAGE]
    const uint32_t stages_s = smem_u32(stages);
    const uint32_t buf2_s = stages_s + NSTAGES * STAGE * 4;        // [RB][Q][128] floats (+ a few floats of slack)
    __shared__ __align__(8) uint64_t full_bar[NSTAGES];
    __shared__ __align__(8) uint64_t row_full[RB], row_empty[RB];
    __shared__ unsigned long long s_acc[2][3];
    __shared__ unsigned s_readers[NSTAGES];            // producer warps that have read the stage into registers

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < NSTAGES; s++) mbar_init(&full_bar[s], 1);
#pragma unroll
        for (int s = 0; s < RB; s++) mbar_init(&row_full[s], 32), mbar_init(&row_empty[s], 96);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < 6) s_acc[tid / 3][tid % 3] = 0ull;
    if (tid < NSTAGES) s_readers[tid] = 0u;
    __syncthreads();

    const int t1 = a.ctrl[0] + a.step_offset;          // the first of the two steps
    const bool live2 = (t1 + 1) != a.ctrl[2];          // the run's last step is not followed by accelerate_flow
    const int epoch = a.ctrl[3] + a.epoch_offset;
    const size_t pitch = a.pitch;
    const bool producer = warp < NP;
    const int pw = producer ? warp : warp - NP;        // this warp's row inside an iteration of its group

    unsigned long long acc = 0ull;                     // per-thread |u| total (units of 2^-40) of this warp's step
    int nbase = 0;                                     // TMA stages consumed by this CTA's earlier units
    int gbase = 0;                                     // ring rows produced by this CTA's earlier interior units
    bool primed = false;                               // the TMA pipeline has been started
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

    // the nine staged planes of one row -> registers (the x part of the pull shift happens here)
    auto load_staged = [&](uint32_t st, int rs, float4 (&v)[Q], float (&sh)[Q]) {
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
    };
    // ring rows rel-1, rel, rel+1 of the intermediate step -> registers: plane k comes from row rel - cy_k, column x - cx_k
    auto load_ring = [&](uint32_t r0, uint32_t rm, uint32_t rp, float4 (&v)[Q], float (&sh)[Q]) {
#pragma unroll
        for (int k = 0; k < Q; k++) {
            const uint32_t row = (dir_cy(k) == 0 ? r0 : (dir_cy(k) == 1 ? rm : rp)) + (k * F2_W1 + 4 * lane) * 4;
            v[k] = lds4(row + 16);
            sh[k] = dir_cx(k) == 0 ? 0.f : (dir_cx(k) == 1 ? lds1(row + 12) : lds1(row + 32));
        }
    };
    auto gather = [&](const float4 (&v)[Q], const float (&sh)[Q], float (&t)[Q][4]) {
#pragma unroll
        for (int k = 0; k < Q; k++) {
            const float e[6] = {sh[k], v[k].x, v[k].y, v[k].z, v[k].w, sh[k]};
#pragma unroll
            for (int j = 0; j < 4; j++) t[k][j] = e[j + 1 - dir_cx(k)]; // cell xg+j pulls column xg+j-cx
        }
    };
    auto obstacle_word = [&](int y, int xo) {
        const uint32_t* orow = (y < 0) ? (a.h.on ? a.obst_halo : a.obst + static_cast<size_t>(a.rows - 1) * a.opitch)
                                       : (y >= a.rows ? (a.h.on ? a.obst_halo + a.opitch : a.obst)
                                                      : a.obst + static_cast<size_t>(y) * a.opitch);
        return __ldg(orow + (xo >> 5));
    };

    for (int u = blockIdx.x; u < a.nunits; u += gridDim.x) {
        const F2Unit un = f2_unit(a, u);
        const int x0 = un.x0, ya = un.ya, yb = un.yb;
        const int rows1 = yb - ya + 2;                       // intermediate rows ya-1 .. yb
        const int nst = (rows1 + SROWS - 1) / SROWS;

        if (un.kind != 2) {
            // ---------------- boundary unit (always ahead of this CTA's interior units): CTA barriers ----------------
            // four intermediate rows from stage buffer 0, filled with ordinary loads after the neighbour has delivered
            // the rows of this epoch; warps 0..3 take a row each; ring slots 0..3
            if (a.h.on && a.h.wait && tid == 0) halo_wait(a.h, epoch, un.kind == 0, un.kind == 1);
            __syncthreads();
            for (int k = warp; k < Q; k += R) f2_load_plane_any<SROWS>(k, a, stages, x0, ya, epoch, lane);
            __syncthreads();
            const int q = warp;
#pragma unroll 1
            for (int phase = 0; phase < 2; phase++) {
                const int rel = q - phase, y = ya - 1 + rel;
                const int xg = x0 + 4 * lane - (phase ? 0 : 4);
                const bool active = phase ? (rel >= 1 && y < yb) : (q < rows1);
                if (active) {
                    const int xo = xg < 0 ? xg + a.nx : (xg >= a.nx ? xg - a.nx : xg);
                    const uint32_t oword = obstacle_word(y, xo);
                    float4 v[Q];
                    float sh[Q], t[Q][4], o[Q][4];
                    if (phase == 0) load_staged(stages_s, q % SROWS, v, sh);
                    else load_ring(buf2_s + (rel % RB) * F2_B2ROW * 4, buf2_s + ((rel - 1) % RB) * F2_B2ROW * 4,
                                   buf2_s + ((rel + 1) % RB) * F2_B2ROW * 4, v, sh);
                    gather(v, sh, t);
                    const uint32_t obits = (oword >> (xo & 31)) & 0xfu;
                    const bool counted = phase ? (lane < 30 && xg < a.nx) : (y >= ya && y < yb && lane >= 1 && lane <= 30 && xg < a.nx);
                    const bool accel = (a.accel_row >= 0) && (y == a.accel_row) && (phase == 0 || live2);
                    unsigned long long tot = 0ull;
                    update4_total<STRICT>(t, obits, counted, accel, a.omega, a.w1a, a.w2a, o, tot, &s_acc[phase][2]);
#pragma unroll
                    for (int s = 16; s > 0; s >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, s);
                    if (lane == 0) atomicAdd(&s_acc[phase][0], tot);
                    if (phase == 0) {
                        const uint32_t dst = buf2_s + ((q % RB) * F2_B2ROW + 4 * lane) * 4;
#pragma unroll
                        for (int k = 0; k < Q; k++) sts4(dst + k * F2_W1 * 4, o[k][0], o[k][1], o[k][2], o[k][3]);
                    } else if (lane < 30 && xg < a.nx) {
                        float* dst = a.out + static_cast<size_t>(y) * pitch + xg;
#pragma unroll
                        for (int k = 0; k < Q; k++)
                            *reinterpret_cast<float4*>(dst + k * a.pf) = make_float4(o[k][0], o[k][1], o[k][2], o[k][3]);
                        if (a.h.on) f2_push_rows(a, un.kind == 0, y, xg, epoch, o);
                    }
                }
                __syncthreads();
            }
            if (a.h.on && tid == 0) halo_arrive(a.h, un.kind == 0, un.kind == 1, static_cast<unsigned>(a.nsx));
            continue;
        }

        // ---------------- interior unit: producers and consumers, no CTA barrier ----------------
        if (!primed) {
            primed = true;
            __syncthreads(); // nobody reads the boundary units' buffer or ring rows any more
            if (tid == 0) {
                for (int j = 0; j < NSTAGES; j++) refill(j, u, 0, nst);
            }
        }
        if (producer) {
            // x wrap of TMA-staged strips: the copy engine zero-fills columns outside [0, nx); the intermediate
            // cells x = -1 and x = nx (periodic images of nx-1 and 0) are needed by step t+2, so the lanes that own
            // them (and their inner neighbours) fetch the wrapped populations themselves
            const bool west = (x0 == 0), east = (x0 + F2_CORE >= a.nx);
            const int le = (a.nx - x0 + 4) >> 2;             // lane whose first cell is x = nx
            const bool pw0 = west && lane == 0, pw1 = west && lane == 1;
            const bool pe0 = east && lane == le, pe1 = east && lane == le - 1;
            const int xg = x0 + 4 * lane - 4;
            const int xo = xg < 0 ? xg + a.nx : (xg >= a.nx ? xg - a.nx : xg);
#pragma unroll 1
            for (int q = pw; q < nst * SROWS; q += NP) {     // this warp's intermediate row, relative to ya-1
                const int y = ya - 1 + q;
                const bool active = q < rows1;
                const int n = nbase + q / SROWS, s = n % NSTAGES;
                uint32_t oword = 0u;
                float pa[3] = {0.f, 0.f, 0.f}, pb[3] = {0.f, 0.f, 0.f};
                if (active) {
                    oword = __ldg(a.obst + static_cast<size_t>(y) * a.opitch + (xo >> 5));
                    if (west || east) {
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
                }
                mbar_wait(&full_bar[s], (n / NSTAGES) & 1);
                float4 v[Q];
                float sh[Q];
                if (active) load_staged(stages_s + (s * STAGE) * 4, q % SROWS, v, sh);
                // this warp's row of the stage is on its way into registers (a warp's shared-memory instructions are
                // performed in order, so the counter below is bumped after the loads above have read the buffer): the
                // LAST of the stage's reader warps asks the copy engine for the stage that will reuse the buffer
                __syncwarp();
                if (lane == 0) {
                    if (atomicAdd(&s_readers[s], 1u) == SROWS - 1) {
                        s_readers[s] = 0u;
                        refill(n + NSTAGES, u, nbase, nst);
                    }
                }
                if (!active) continue;
                float t[Q][4], o[Q][4];
                gather(v, sh, t);
                if (west || east) {
                    if (pw0) t[1][3] = pa[0], t[5][3] = pa[1], t[8][3] = pa[2], t[0][3] = pb[0], t[2][3] = pb[1], t[4][3] = pb[2];
                    if (pw1) t[1][0] = pa[0], t[5][0] = pa[1], t[8][0] = pa[2];
                    if (pe0) t[3][0] = pa[0], t[6][0] = pa[1], t[7][0] = pa[2], t[0][0] = pb[0], t[2][0] = pb[1], t[4][0] = pb[2];
                    if (pe1) t[3][3] = pa[0], t[6][3] = pa[1], t[7][3] = pa[2];
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
                // cells whose |u| this unit owns: its core columns, its own rows
                const bool counted = (y >= ya && y < yb && lane >= 1 && lane <= 30 && xg < a.nx);
                const bool accel = (y == a.accel_row);
                update4_total<STRICT>(t, obits, counted, accel, a.omega, a.w1a, a.w2a, o, acc, &s_acc[0][2]);
                // the ring slot of this row: free once the three consumer tasks of its previous tenant have read it
                const int g = gbase + q, slot = g % RB;
                mbar_wait(&row_empty[slot], ((g / RB) & 1) ^ 1);
                const uint32_t dst = buf2_s + (slot * F2_B2ROW + 4 * lane) * 4;
#pragma unroll
                for (int k = 0; k < Q; k++) sts4(dst + k * F2_W1 * 4, o[k][0], o[k][1], o[k][2], o[k][3]);
                mbar_arrive(&row_full[slot]);
            }
        } else {
            const int xg = x0 + 4 * lane;
            const bool in_core = lane < 30 && xg < a.nx;
            // task q reads the ring rows q-2, q-1, q and writes slab row ya-1 + (q-1); tasks 0, 1, rows1, rows1+1 only
            // take part in the hand-back of the rows (every row is released by exactly three tasks)
#pragma unroll 1
            for (int q = pw; q < rows1 + 2; q += NP) {
                const int rel = q - 1, y = ya - 1 + rel;
                const bool valid = rel >= 1 && y < yb;
                uint32_t oword = 0u;
                if (valid) oword = __ldg(a.obst + static_cast<size_t>(y) * a.opitch + (xg >> 5));
                const int g2 = gbase + q - 2, g1 = g2 + 1, g0 = g2 + 2;   // ring rows rel-1, rel, rel+1
                const bool e2 = q - 2 >= 0 && q - 2 < rows1, e1 = q - 1 >= 0 && q - 1 < rows1, e0 = q < rows1;
                const int s2 = (g2 + RB) % RB, s1 = (g1 + RB) % RB, s0 = g0 % RB;
                if (e2) mbar_wait(&row_full[s2], (g2 / RB) & 1);
                if (e1) mbar_wait(&row_full[s1], (g1 / RB) & 1);
                if (e0) mbar_wait(&row_full[s0], (g0 / RB) & 1);
                float4 v[Q];
                float sh[Q];
                if (valid) load_ring(buf2_s + s1 * F2_B2ROW * 4, buf2_s + s2 * F2_B2ROW * 4, buf2_s + s0 * F2_B2ROW * 4, v, sh);
                if (e2) mbar_arrive(&row_empty[s2]);
                if (e1) mbar_arrive(&row_empty[s1]);
                if (e0) mbar_arrive(&row_empty[s0]);
                if (!valid) continue;
                float t[Q][4], o[Q][4];
                gather(v, sh, t);
                const uint32_t obits = (oword >> (xg & 31)) & 0xfu;
                const bool accel = (y == a.accel_row) && live2;
                update4_total<STRICT>(t, obits, in_core, accel, a.omega, a.w1a, a.w2a, o, acc, &s_acc[1][2]);
                if (in_core) {
                    float* dst = a.out + static_cast<size_t>(y) * pitch + xg;
#pragma unroll
                    for (int k = 0; k < Q; k++)
                        *reinterpret_cast<float4*>(dst + k * a.pf) = make_float4(o[k][0], o[k][1], o[k][2], o[k][3]);
                }
            }
        }
        nbase += nst;
        gbase += rows1;
    }

    // ---------------- |u| sums of both steps (s_acc[step][0], units of 2^-40) ----------------
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
    if (lane == 0) atomicAdd(&s_acc[producer ? 0 : 1][0], acc);
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

} // namespace lbm
