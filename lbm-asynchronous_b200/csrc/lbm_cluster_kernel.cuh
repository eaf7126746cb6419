// lbm_cluster_kernel.cuh -- step_cluster_kernel: every timestep of a run with the lattice RESIDENT IN SHARED MEMORY
// of one thread-block cluster (the reference's smallest shipped grids, where its README quotes its
// headline speed-ups, /root/reference/README.md:126-128).
//
// For a 128 x 128 grid a timestep is well under 1 us of arithmetic; step_loop_kernel (lbm_kernels.cuh) spends more than
// that on its grid-wide barrier (fence + global atomic + acquire spin: three L2 round trips) and on reading the lattice
// back from L2.  Here the lattice never leaves the SMs between the first and the last step of an lbm_run call:
//   * one cluster of C <= 16 CTAs (16 is the non-portable maximum), CTA c owns the rows [c*rpc, (c+1)*rpc) of the
//     slab in its own shared memory: TWO copies of the nine SoA planes, [2][9][rpc][nx] floats, ping-ponged;
//   * a thread owns CPT consecutive cells of one row for the whole run (obstacle bits, row addresses and the
//     addresses of the row above / below -- in the neighbouring CTA's shared memory for the CTA's first / last row,
//     `mapa` -- are computed once);
//   * a step = pull the nine populations of the owned cells out of (distributed) shared memory with
//     ld.shared::cluster, collide (the same arithmetic as every other step kernel: the strict flavour stays
//     bit-identical to SerialCode), store the new populations into the other copy (own shared memory only), and
//     tell the three CTAs that will read them -- itself and the owners of the rows above and below -- through
//     mbarriers in their shared memory (one remote arrive per warp and reader).  A CTA starts step s+1 when its own
//     mbarrier of step s is complete: no cluster-wide barrier, a CTA only ever waits for its two neighbours, and one
//     wait per step covers both hazards (a neighbour's arrival for step s comes after its reads of step s, so the copy
//     read at step s may be overwritten at step s+1);
//   * |u| of the new state is computed AFTER the arrive, i.e. while the neighbours' arrivals are in flight;
//     integer warp reduction, per-warp parts in shared memory, one RED per CTA, step and word to sums[step], two steps late;
//   * accelerate-at-store as in the other kernels; the last step of the launch is stored to the destination lattice
//     in global memory instead of shared memory.
// A first version (one copy of the lattice, two hardware cluster barriers per step) measured 2.1-2.5 us per step at
// 128 x 128 against 2.4 us from step_loop_kernel: barrier.cluster.arrive.release is MEMBAR.ALL.GPU + UCGABAR_ARV and
// every wait costs ~0.25 us even when it is already complete (profiles/r02_small_grids.md).
// Replaces the timestep loop of SerialCode/d2q9-bgk.c:187-194 (accelerate_flow / propagate / rebound / collision /
// av_velocity per step) for grids of up to CLUSTER_MAX_CELLS cells on one GPU.
#pragma once

#include "lbm_kernels.cuh"

namespace lbm {

struct ClusterArgs {
    float* lat[2];       // two lattices in global memory, 9 planes each: [src] holds the state before first_step
    size_t pf;           // floats per plane
    const uint32_t* obst;
    unsigned long long* sums; // [nsteps][nslots][SUM_WORDS] of this launch
    int nslots;
    int first_step, nsteps, last_step; // absolute indices; no accelerate-at-store at last_step
    int src;
    int nx, rows, pitch, opitch;
    int rpc;             // rows per CTA (the same for every CTA; the last CTAs may own fewer or none)
    int accel_row;
    int sync_mode;       // 0: mbarrier arrivals with release semantics at cluster scope; 1: CTA-scope fence + relaxed arrive
    float omega, w1a, w2a;
};

__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank)
{
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ float lds_cluster(uint32_t addr)
{
    float v;
    asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ float2 lds_cluster2(uint32_t addr)
{
    float2 v;
    asm volatile("ld.shared::cluster.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ float4 lds_cluster4(uint32_t addr)
{
    float4 v;
    asm volatile("ld.shared::cluster.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
    return v;
}
// arrive with release semantics at cluster scope (ptxas: MEMBAR.ALL.GPU + UCGABAR_ARV -- it also waits for the
// thread's outstanding global-memory operations)
__device__ __forceinline__ void cluster_arrive()
{
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
}
__device__ __forceinline__ void cluster_wait()
{
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
// one arrival on an mbarrier of any CTA of the cluster (shared::cluster address); release at cluster scope: the
// warp's earlier shared-memory stores (made visible to this lane by __syncwarp) are visible to whoever acquires it
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar)
{
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar) : "memory");
}
// the same after a CTA-scope fence only (tuning knob): the stores went to this CTA's own shared memory, the one place
// the other CTAs read them from
__device__ __forceinline__ void mbar_arrive_cluster_ctafence(uint32_t bar)
{
    asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity)
{
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!done);
}

// The collision in two halves, so that the new populations can be published before |u| of the new state is formed.
// relax_cells: populations of 2*NP cells (bounce-back applied); returns true when a cell was outside the early window
// of lbm_collide4.cuh and update_cell()'s guarded code ran instead (speed[] is then final).
template <bool STRICT, int NP>
__device__ __forceinline__ bool relax_cells(const float (&t)[Q][2 * NP], uint32_t obits, float omega, float (&o)[Q][2 * NP], float (&speed)[2 * NP])
{
    constexpr int mirror[Q] = {0, 3, 4, 1, 2, 7, 8, 5, 6};
    PV<NP> tp[Q], rho, mx, my;
#pragma unroll
    for (int k = 0; k < Q; k++)
#pragma unroll
        for (int h = 0; h < NP; h++) tp[k].v[h] = pk(t[k][2 * h], t[k][2 * h + 1]);
    vmoments<STRICT>(tp, rho, mx, my);
    if (!all_in_window(rho, mx, my, obits, 9.5367431640625e-07f /* 2^-20 */, 1048576.f /* 2^20 */, 2.f)) {
#pragma unroll
        for (int j = 0; j < 2 * NP; j++) {
            float tj[Q], oc[Q];
#pragma unroll
            for (int k = 0; k < Q; k++) tj[k] = t[k][j];
            speed[j] = update_cell<STRICT>(tj, (obits >> j) & 1u, omega, oc);
#pragma unroll
            for (int k = 0; k < Q; k++) o[k][j] = oc[k];
        }
        return true;
    }
    PV<NP> c[Q];
    vrelax<STRICT>(tp, rho, mx, my, omega, c);
#pragma unroll
    for (int h = 0; h < NP; h++) {
        const bool solid0 = (obits >> (2 * h)) & 1u, solid1 = (obits >> (2 * h + 1)) & 1u;
#pragma unroll
        for (int k = 0; k < Q; k++) {
            float c0, c1;
            upk(c[k].v[h], c0, c1);
            o[k][2 * h] = solid0 ? t[mirror[k]][2 * h] : c0;
            o[k][2 * h + 1] = solid1 ? t[mirror[k]][2 * h + 1] : c1;
        }
    }
    return false;
}
// speed_cells: |u| of the stored populations (SerialCode:425-452), the second half of collide_block()
template <bool STRICT, int NP>
__device__ __forceinline__ void speed_cells(const float (&o)[Q][2 * NP], uint32_t obits, float (&speed)[2 * NP])
{
    PV<NP> c[Q], r2, nx_, ny_;
#pragma unroll
    for (int k = 0; k < Q; k++)
#pragma unroll
        for (int h = 0; h < NP; h++) c[k].v[h] = pk(o[k][2 * h], o[k][2 * h + 1]);
    vmoments<STRICT>(c, r2, nx_, ny_);
    const bool late = all_in_window(r2, nx_, ny_, obits, 4.76837158203125e-07f /* 2^-21 */, 2097152.f /* 2^21 */, 4.f);
    if constexpr (STRICT) {
        PV<NP> vx, vy;
        vdiv2(nx_, ny_, r2, vx, vy);
        vspeed_from_sq(vadd(vmul_rounded(vx, vx), vmul_rounded(vy, vy)), speed);
    } else {
        vspeed_from_sq(vfma(nx_, nx_, vmul(ny_, ny_)), speed);
#pragma unroll
        for (int h = 0; h < NP; h++) {
            float r0, r1;
            upk(r2.v[h], r0, r1);
            speed[2 * h] = __fdividef(speed[2 * h], r0), speed[2 * h + 1] = __fdividef(speed[2 * h + 1], r1);
        }
    }
    if (!late) {
#pragma unroll
        for (int j = 0; j < 2 * NP; j++) {
            float cj[Q];
#pragma unroll
            for (int k = 0; k < Q; k++) cj[k] = o[k][j];
            if (!((obits >> j) & 1u)) {
                if constexpr (STRICT) speed[j] = speed_strict(cj);
                else speed[j] = speed_fast_guarded(cj);
            }
        }
    }
}

// CPT cells of a thread, first half: new populations (collision / bounce-back).  Returns true if speed[] is final.
template <bool STRICT, int CPT, bool VERT>
__device__ __forceinline__ bool cells_relax(const float (&t)[Q][CPT], uint32_t obits, float omega, float (&o)[Q][CPT], float (&speed)[CPT])
{
    if constexpr (CPT == 4 && VERT) {
        return relax_cells<STRICT, 2>(t, obits, omega, o, speed);
    } else if constexpr (CPT == 4) {
        bool done = true;
#pragma unroll
        for (int h = 0; h < 2; h++) {
            float t2[Q][2], o2[Q][2], s2[2];
#pragma unroll
            for (int k = 0; k < Q; k++) t2[k][0] = t[k][2 * h], t2[k][1] = t[k][2 * h + 1];
            const bool d = relax_cells<STRICT, 1>(t2, (obits >> (2 * h)) & 3u, omega, o2, s2);
#pragma unroll
            for (int k = 0; k < Q; k++) o[k][2 * h] = o2[k][0], o[k][2 * h + 1] = o2[k][1];
            speed[2 * h] = s2[0], speed[2 * h + 1] = s2[1];
            done = done && d; // a pair that took the vector code has no speed yet: speed_cells() redoes all four
        }
        return done;
    } else if constexpr (CPT == 2) {
        return relax_cells<STRICT, 1>(t, obits & 3u, omega, o, speed);
    } else {
        float tj[Q], oc[Q];
#pragma unroll
        for (int k = 0; k < Q; k++) tj[k] = t[k][0];
        speed[0] = update_cell<STRICT>(tj, obits & 1u, omega, oc);
#pragma unroll
        for (int k = 0; k < Q; k++) o[k][0] = oc[k];
        return true;
    }
}
// second half: |u| of the new (pre-accelerate) populations
template <bool STRICT, int CPT, bool VERT>
__device__ __forceinline__ void cells_speed(const float (&o)[Q][CPT], uint32_t obits, float (&speed)[CPT])
{
    if constexpr (CPT == 4 && VERT) {
        speed_cells<STRICT, 2>(o, obits, speed);
    } else if constexpr (CPT == 4) {
#pragma unroll
        for (int h = 0; h < 2; h++) {
            float o2[Q][2], s2[2];
#pragma unroll
            for (int k = 0; k < Q; k++) o2[k][0] = o[k][2 * h], o2[k][1] = o[k][2 * h + 1];
            speed_cells<STRICT, 1>(o2, (obits >> (2 * h)) & 3u, s2);
            speed[2 * h] = s2[0], speed[2 * h + 1] = s2[1];
        }
    } else if constexpr (CPT == 2) {
        speed_cells<STRICT, 1>(o, obits & 3u, speed);
    }
}

template <int CPT>
__device__ __forceinline__ void store_cells(float* dst, size_t plane_stride, const float (&o)[Q][CPT])
{
#pragma unroll
    for (int k = 0; k < Q; k++) {
        if constexpr (CPT == 4) {
            *reinterpret_cast<float4*>(dst + k * plane_stride) = make_float4(o[k][0], o[k][1], o[k][2], o[k][3]);
        } else if constexpr (CPT == 2) {
            *reinterpret_cast<float2*>(dst + k * plane_stride) = make_float2(o[k][0], o[k][1]);
        } else {
            dst[k * plane_stride] = o[k][0];
        }
    }
}

// dynamic shared memory: [2][9][rpc][nx] floats (nx % CPT == 0; 16-byte aligned rows when CPT == 4)
template <bool STRICT, int CPT, bool VERT, int MAXT>
__global__ void __launch_bounds__(MAXT, 1) step_cluster_kernel(const ClusterArgs a)
{
    extern __shared__ __align__(16) float cl_smem[];
    __shared__ __align__(8) unsigned long long s_bar[2];
    __shared__ unsigned s_part[3][32][4]; // per-warp |u| sums of a step (lo, hi, non-finite), three steps in flight
    const int tid = threadIdx.x;
    const int lane = tid & 31;

    const int rank = static_cast<int>(cluster_ctarank());
    const int nx = a.nx, rpc = a.rpc;
    const int nact = (a.rows + rpc - 1) / rpc; // CTAs that own rows: ranks 0 .. nact-1
    const bool active = rank < nact;
    const int ipr = nx / CPT; // threads per row
    const int lrow_raw = tid / ipr;
    const int x0 = (tid - lrow_raw * ipr) * CPT;
    const int row0 = rank * rpc;
    const int my_rows = max(0, min(rpc, a.rows - row0));
    const bool valid = lrow_raw < my_rows;
    const int lrow = valid ? lrow_raw : 0;
    const int gy = min(row0 + lrow, a.rows - 1); // clamped: idle threads compute on a legal address and store nothing
    const int ys = (gy == 0) ? a.rows - 1 : gy - 1; // SerialCode:257-258
    const int yn = (gy == a.rows - 1) ? 0 : gy + 1;
    const uint32_t plane_bytes = static_cast<uint32_t>(rpc) * nx * 4u;
    const uint32_t copy_bytes = Q * plane_bytes;
    const uint32_t base = static_cast<uint32_t>(__cvta_generic_to_shared(cl_smem));
    const uint32_t addr_c = mapa_u32(base + (static_cast<uint32_t>(gy % rpc) * nx + x0) * 4u, static_cast<uint32_t>(gy / rpc));
    const uint32_t addr_s = mapa_u32(base + (static_cast<uint32_t>(ys % rpc) * nx + x0) * 4u, static_cast<uint32_t>(ys / rpc));
    const uint32_t addr_n = mapa_u32(base + (static_cast<uint32_t>(yn % rpc) * nx + x0) * 4u, static_cast<uint32_t>(yn / rpc));
    // the cell west of x0 / east of x0 + CPT - 1, periodic (SerialCode:259-262), as byte offsets from x0
    const int dxw = ((x0 == 0) ? nx - 1 : x0 - 1) * 4 - x0 * 4;
    const int dxe = ((x0 + CPT == nx) ? 0 : x0 + CPT) * 4 - x0 * 4;
    // a warp that holds 32 consecutive threads of one row gets the west / east cell from the neighbouring lane
    const bool by_shuffle = (CPT > 1) && (ipr % 32 == 0);
    const bool wrap_in_warp = (ipr == 32);
    float* own = cl_smem + static_cast<size_t>(lrow) * nx + x0; // plane 0, copy 0 of the owned cells (own shared memory)
    const size_t plane_floats = static_cast<size_t>(rpc) * nx;
    const size_t copy_floats = Q * plane_floats;
    const size_t goff = static_cast<size_t>(gy) * a.pitch + x0;
    const uint32_t obits = (__ldg(a.obst + static_cast<size_t>(gy) * a.opitch + (x0 >> 5)) >> (x0 & 31)) & ((1u << CPT) - 1u);

    // mbarriers: s_bar[s & 1] collects, for step s, one arrival per warp of this CTA and of the two neighbouring
    // CTAs (cyclically among the CTAs that own rows; they may coincide)
    const uint32_t bar0 = static_cast<uint32_t>(__cvta_generic_to_shared(&s_bar[0]));
    const uint32_t nwarps = blockDim.x >> 5;
    if (tid == 0) {
        mbar_init(bar0, 3u * nwarps);
        mbar_init(bar0 + 8u, 3u * nwarps);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    const uint32_t bar_s = mapa_u32(bar0, static_cast<uint32_t>(active ? (rank + nact - 1) % nact : rank));
    const uint32_t bar_n = mapa_u32(bar0, static_cast<uint32_t>(active ? (rank + 1) % nact : rank));

    // the state before first_step: global -> copy 0 in own shared memory
    if (valid) {
        const float* in = a.lat[a.src & 1];
#pragma unroll
        for (int k = 0; k < Q; k++) {
            if constexpr (CPT == 4) {
                *reinterpret_cast<float4*>(own + k * plane_floats) = __ldcg(reinterpret_cast<const float4*>(in + k * a.pf + goff));
            } else if constexpr (CPT == 2) {
                *reinterpret_cast<float2*>(own + k * plane_floats) = __ldcg(reinterpret_cast<const float2*>(in + k * a.pf + goff));
            } else {
                own[k * plane_floats] = __ldcg(in + k * a.pf + goff);
            }
        }
    }
    cluster_arrive();
    cluster_wait();

    float* out = a.lat[(a.src + a.nsteps) & 1];
    const bool on_accel_row = valid && (row0 + lrow == a.accel_row);
    const int nsteps = active ? a.nsteps : 0; // CTAs without rows only take part in the two cluster barriers
    for (int s = 0; s < nsteps; s++) {
        if (s > 0) mbar_wait_cluster(bar0 + 8u * ((s - 1) & 1), static_cast<uint32_t>(((s - 1) >> 1) & 1));
        // the sums of step s-2 are complete: every warp of this CTA formed them before it arrived for step s-1
        if (s >= 2 && tid < 3) {
            unsigned long long v = 0ull;
            for (uint32_t w = 0; w < nwarps; w++) v += s_part[(s - 2) % 3][w][tid]; // rewritten at step s+1, behind this warp's arrival for step s
            if (v) atomicAdd(a.sums + (static_cast<size_t>(s - 2) * a.nslots + (blockIdx.x & (a.nslots - 1))) * SUM_WORDS + tid, v);
        }
        // ---- pull: planes 1,5,8 come from the west, 3,6,7 from the east; 2,5,6 from the row below (south), 4,7,8
        // from the row above (SerialCode:263-279)
        const uint32_t rd = (s & 1) ? copy_bytes : 0u;
        float t[Q][CPT];
#pragma unroll
        for (int k = 0; k < Q; k++) {
            const bool from_south = (k == 2 || k == 5 || k == 6), from_north = (k == 4 || k == 7 || k == 8);
            const bool from_west = (k == 1 || k == 5 || k == 8), from_east = (k == 3 || k == 6 || k == 7);
            const uint32_t ad = (from_south ? addr_s : (from_north ? addr_n : addr_c)) + rd + static_cast<uint32_t>(k) * plane_bytes;
            float v[CPT];
            if constexpr (CPT == 4) {
                const float4 q = lds_cluster4(ad);
                v[0] = q.x, v[1] = q.y, v[2] = q.z, v[3] = q.w;
            } else if constexpr (CPT == 2) {
                const float2 q = lds_cluster2(ad);
                v[0] = q.x, v[1] = q.y;
            } else {
                v[0] = (from_west || from_east) ? 0.f : lds_cluster(ad);
            }
            if (from_west) {
                float w;
                if (by_shuffle) {
                    w = __shfl_sync(0xffffffffu, v[CPT - 1], (lane + 31) & 31);
                    if (lane == 0 && !wrap_in_warp) w = lds_cluster(ad + static_cast<uint32_t>(dxw));
                } else {
                    w = lds_cluster(ad + static_cast<uint32_t>(dxw));
                }
                t[k][0] = w;
#pragma unroll
                for (int j = 1; j < CPT; j++) t[k][j] = v[j - 1];
            } else if (from_east) {
                float e;
                if (by_shuffle) {
                    e = __shfl_sync(0xffffffffu, v[0], (lane + 1) & 31);
                    if (lane == 31 && !wrap_in_warp) e = lds_cluster(ad + static_cast<uint32_t>(dxe));
                } else {
                    e = lds_cluster(ad + static_cast<uint32_t>(dxe));
                }
#pragma unroll
                for (int j = 0; j + 1 < CPT; j++) t[k][j] = v[j + 1];
                t[k][CPT - 1] = e;
            } else {
#pragma unroll
                for (int j = 0; j < CPT; j++) t[k][j] = v[j];
            }
        }

        const bool last = (s + 1 == nsteps);
        const bool accel = on_accel_row && (a.first_step + s != a.last_step);
        float o[Q][CPT], speed[CPT];
        const bool speed_done = cells_relax<STRICT, CPT, VERT>(t, obits, a.omega, o, speed);
        if (valid) {
            float* dst = last ? out + goff : own + ((s & 1) ? 0 : copy_floats);
            const size_t ps = last ? a.pf : plane_floats;
            if (accel) {
                // accelerate_flow() of the next step on the values being stored (|u| below is of the collided state)
                float oa[Q][CPT];
#pragma unroll
                for (int j = 0; j < CPT; j++) {
                    float oc[Q];
#pragma unroll
                    for (int k = 0; k < Q; k++) oc[k] = o[k][j];
                    accelerate_cell(oc, (obits >> j) & 1u, a.w1a, a.w2a);
#pragma unroll
                    for (int k = 0; k < Q; k++) oa[k][j] = oc[k];
                }
                store_cells<CPT>(dst, ps, oa);
            } else {
                store_cells<CPT>(dst, ps, o);
            }
        }
        if (!last) {
            // publish: this warp's part of the new state is in place
            __syncwarp();
            if (lane == 0) {
                const uint32_t off = 8u * (s & 1);
                if (a.sync_mode == 0) {
                    mbar_arrive_cluster(bar0 + off);
                    mbar_arrive_cluster(bar_s + off);
                    mbar_arrive_cluster(bar_n + off);
                } else {
                    asm volatile("fence.acq_rel.cta;" ::: "memory");
                    mbar_arrive_cluster_ctafence(bar0 + off);
                    mbar_arrive_cluster_ctafence(bar_s + off);
                    mbar_arrive_cluster_ctafence(bar_n + off);
                }
            }
        }
        // ---- |u| of the new state, while the neighbours' arrivals are on their way
        if (!speed_done) cells_speed<STRICT, CPT, VERT>(o, obits, speed);
        SpeedAcc acc = {0u, 0u, 0u};
#pragma unroll
        for (int j = 0; j < CPT; j++) acc_speed(acc, speed[j], valid && !((obits >> j) & 1u));
        const unsigned lo = __reduce_add_sync(0xffffffffu, acc.lo);
        const unsigned hi = __reduce_add_sync(0xffffffffu, acc.hi);
        const unsigned nbad = __reduce_add_sync(0xffffffffu, acc.bad);
        if (lane == 0) {
            unsigned* pp = s_part[s % 3][tid >> 5];
            pp[0] = lo, pp[1] = hi, pp[2] = nbad;
        }
    }
    // nobody leaves while a neighbour may still read its shared memory or arrive on its mbarriers
    cluster_arrive();
    cluster_wait();
    if (tid < 3) {
        for (int s = max(0, nsteps - 2); s < nsteps; s++) {
            unsigned long long v = 0ull;
            for (uint32_t w = 0; w < nwarps; w++) v += s_part[s % 3][w][tid];
            if (v) atomicAdd(a.sums + (static_cast<size_t>(s) * a.nslots + (blockIdx.x & (a.nslots - 1))) * SUM_WORDS + tid, v);
        }
    }
}

} // namespace lbm
