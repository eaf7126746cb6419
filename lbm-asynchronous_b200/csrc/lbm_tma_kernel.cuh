// lbm_tma_kernel.cuh -- the interior-row timestep kernel: TMA-staged, persistent, warp-specialised.
//
// Same arithmetic as step_vec4_kernel (update_cell / accelerate_cell / acc_speed of lbm_kernels.cuh);
// what differs is how the nine populations reach the registers:
//
//   * the lattice (9 planes x rows x pitch) is described to the TMA unit as a 3-D tensor (x, y, plane);
//   * a tile is TX=128 cells x TY rows.  For plane k the producer asks for the box whose rows start at
//     y0 - cy_k: the y part of the pull-streaming shift (SerialCode/d2q9-bgk.c:257-272) is done by the
//     copy engine.  TMA wants the innermost start coordinate 16-byte aligned (tools/tma_probe.cu: an
//     odd x start raises "illegal instruction"), so the x part is done on the way out of shared
//     memory: planes 1,5,8 (need x-1) are staged 132 wide from x0-4, planes 3,6,7 (need x+1) 132 wide
//     from x0, and a thread reads one aligned float4 plus one scalar per shifted plane.  No shuffles,
//     no warp-edge special cases, no per-plane global address arithmetic;
//   * columns outside [0, nx) are zero-filled by TMA; the periodic wrap in x (SerialCode:259-262) is
//     patched by the two threads per row that own column 0 / nx-1 with three scalar loads each;
//   * rows outside the slab never occur: this kernel owns rows [1, rows-1) only.  Rows 0 and rows-1
//     (periodic wrap in y, or the halo ring of a neighbouring GPU) belong to step_vec4_kernel /
//     step_scalar_kernel in boundary mode, which runs concurrently as a second graph branch -- the
//     interior / boundary split of MPI_Waitall/d2q9-bgk.c:234-253;
//   * CTAs are persistent (grid = resident CTAs per SM x SMs) and walk the tile list with stride
//     gridDim.x; one producer warp keeps STAGES tiles in flight per CTA through full/empty mbarriers,
//     TY consumer warps (one tile row each, 4 cells per thread) compute and store with 128-bit STG;
//   * |u| sums: per-thread integer accumulators over all tiles of the CTA, one reduction per launch.
#pragma once

#include <cuda.h>

#include "lbm_kernels.cuh"

namespace lbm {

constexpr int TMA_TX = 128;  // cells per tile row (one warp x 4 cells)
constexpr int TMA_TXW = 132; // staged width of the x-shifted planes (one float4 column of apron)

// pull offsets (cx, cy) of the nine directions, SerialCode/d2q9-bgk.c:9-15
__host__ __device__ constexpr int dir_cx(int k) { return k == 1 || k == 5 || k == 8 ? 1 : (k == 3 || k == 6 || k == 7 ? -1 : 0); }
__host__ __device__ constexpr int dir_cy(int k) { return k == 2 || k == 5 || k == 6 ? 1 : (k == 4 || k == 7 || k == 8 ? -1 : 0); }
__host__ __device__ constexpr int plane_width(int k) { return dir_cx(k) == 0 ? TMA_TX : TMA_TXW; }
// float offset of staged plane k inside a stage with TY rows (every plane starts 128-byte aligned, as
// the TMA destination must be)
__host__ __device__ constexpr int plane_offset(int k, int ty)
{
    int off = 0;
    for (int i = 0; i < k; i++) off += (plane_width(i) * ty + 31) / 32 * 32;
    return off;
}
__host__ __device__ constexpr int stage_floats(int ty) { return plane_offset(Q, ty); }
// bytes the nine TMA boxes of one stage deliver (no padding): what the full barrier expects
__host__ __device__ constexpr unsigned stage_tx_bytes(int ty)
{
    unsigned n = 0;
    for (int i = 0; i < Q; i++) n += static_cast<unsigned>(plane_width(i) * ty) * sizeof(float);
    return n;
}

struct TmaArgs {
    float* out[Q];           // destination planes
    const float* west[3];    // source planes 1,5,8 (periodic column nx-1 for x0 == 0)
    const float* east[3];    // source planes 3,6,7 (periodic column 0 for x0+4 == nx)
    const uint32_t* obst;
    const int* ctrl;
    unsigned long long* const* sums_ref; // see StepArgs
    int nslots, step_offset;
    int nx, pitch, opitch;
    int y_first, y_end;      // rows [y_first, y_end) of the slab are this kernel's
    int ntx, ntiles;         // tiles per row band, tiles in total
    int dq, dr;              // gridDim.x / ntx, gridDim.x % ntx  (tile walk without divisions)
    int accel_row;           // slab row that gets accelerate_flow at store time, or -1
    float omega, w1a, w2a;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2)
{
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
            smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}

// tmap: boxes of TMA_TX x TY x 1 (planes 0,2,4); tmapw: boxes of TMA_TXW x TY x 1 (the x-shifted planes)
template <bool STRICT, int TY, int STAGES, int MINB>
__global__ void __launch_bounds__(32 * TY + 32, MINB)
    step_tma_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ CUtensorMap tmapw, const TmaArgs a)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* tiles = reinterpret_cast<float*>(smem_raw); // [STAGES][stage_floats(TY)]
    __shared__ __align__(8) uint64_t full_bar[STAGES];
    __shared__ __align__(8) uint64_t empty_bar[STAGES];
    __shared__ unsigned long long s_acc[3];
    constexpr int STAGE = stage_floats(TY);      // floats per stage
    constexpr uint32_t STAGE_BYTES = stage_tx_bytes(TY);

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; s++) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], TY);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < 3) s_acc[tid] = 0ull;
    __syncthreads();

    unsigned long long acc_lo = 0ull, acc_hi = 0ull; // per-thread sums over all tiles of this CTA
    unsigned acc_bad = 0u;

    // tile walk: tile = blockIdx.x + it * gridDim.x  ->  (by, bx) without divisions after the first
    int tile = blockIdx.x;
    int by = tile / a.ntx, bx = tile - by * a.ntx;

    if (warp == TY) {
        // ---------------- producer warp: one elected lane feeds the stages ----------------
        if (lane == 0) {
            int it = 0;
            for (; tile < a.ntiles; tile += gridDim.x, it++) {
                const int s = it % STAGES;
                if (it >= STAGES) mbar_wait(&empty_bar[s], ((it / STAGES) - 1) & 1);
                const int x0 = bx * TMA_TX, y0 = a.y_first + by * TY;
                mbar_expect_tx(&full_bar[s], STAGE_BYTES);
                float* dst = tiles + s * STAGE;
#pragma unroll
                for (int k = 0; k < Q; k++)
                    tma_load_3d(dst + plane_offset(k, TY), dir_cx(k) == 0 ? &tmap : &tmapw, &full_bar[s],
                                dir_cx(k) == 1 ? x0 - 4 : x0, y0 - dir_cy(k), k);
                bx += a.dr, by += a.dq;
                if (bx >= a.ntx) bx -= a.ntx, by++;
            }
        }
    } else {
        // ---------------- consumer warps: warp w owns tile row w ----------------
        const size_t pitch = a.pitch;
        int step = 0;
        bool accel_live = false;
        if (a.accel_row >= 0) {
            step = a.ctrl[0] + a.step_offset;
            accel_live = (step != a.ctrl[2]); // the run's last step is not followed by accelerate_flow
        }
        int it = 0;
        for (; tile < a.ntiles; tile += gridDim.x, it++) {
            const int s = it % STAGES;
            const int x0 = bx * TMA_TX + lane * 4;
            const int y = a.y_first + by * TY + warp;
            const bool valid = (x0 < a.nx) && (y < a.y_end);
            const size_t roff = static_cast<size_t>(y) * pitch;

            // issued before the wait so that their latency hides behind it
            uint32_t oword = 0u;
            float w1 = 0.f, w5 = 0.f, w8 = 0.f, e3 = 0.f, e6 = 0.f, e7 = 0.f;
            const bool west = valid && (x0 == 0);
            const bool east = valid && (x0 + 4 == a.nx);
            if (valid) oword = __ldg(a.obst + static_cast<size_t>(y) * a.opitch + (x0 >> 5));
            if (west) {
                const size_t xw = a.nx - 1;
                w1 = __ldg(a.west[0] + roff + xw);
                w5 = __ldg(a.west[1] + roff - pitch + xw);
                w8 = __ldg(a.west[2] + roff + pitch + xw);
            }
            if (east) {
                e3 = __ldg(a.east[0] + roff);
                e6 = __ldg(a.east[1] + roff - pitch);
                e7 = __ldg(a.east[2] + roff + pitch);
            }

            mbar_wait(&full_bar[s], (it / STAGES) & 1);
            const float* st = tiles + s * STAGE;
            float4 v[Q];
            float sh[Q]; // the fifth value of an x-shifted plane: column x0-1 (planes 1,5,8) or x0+4 (planes 3,6,7)
#pragma unroll
            for (int k = 0; k < Q; k++) {
                const float* row = st + plane_offset(k, TY) + warp * plane_width(k);
                if (dir_cx(k) == 0) {
                    v[k] = *reinterpret_cast<const float4*>(row + 4 * lane);
                    sh[k] = 0.f;
                } else if (dir_cx(k) == 1) { // staged from x0-4
                    v[k] = *reinterpret_cast<const float4*>(row + 4 + 4 * lane);
                    sh[k] = row[3 + 4 * lane];
                } else {                      // staged from x0
                    v[k] = *reinterpret_cast<const float4*>(row + 4 * lane);
                    sh[k] = row[4 + 4 * lane];
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty_bar[s]); // this warp's row of stage s is in registers

            if (west) sh[1] = w1, sh[5] = w5, sh[8] = w8;
            if (east) sh[3] = e3, sh[6] = e6, sh[7] = e7;

            const uint32_t obits = (oword >> (x0 & 31)) & 0xfu;
            float t[Q][4];
#pragma unroll
            for (int k = 0; k < Q; k++) {
                const float e[6] = {sh[k], v[k].x, v[k].y, v[k].z, v[k].w, sh[k]};
#pragma unroll
                for (int j = 0; j < 4; j++) t[k][j] = e[j + 1 - dir_cx(k)]; // cell x0+j pulls column x0+j-cx
            }
            float o[Q][4];
            SpeedAcc acc = {0u, 0u, 0u}; // this tile's four cells: lo < 2^26, hi < 2^20
            update4<STRICT, (MINB == 1 && TY <= 12)>(t, obits, valid, accel_live && (y == a.accel_row), a.omega, a.w1a, a.w2a, o, acc);
            acc_lo += acc.lo, acc_hi += acc.hi, acc_bad += acc.bad;

            if (valid) {
#pragma unroll
                for (int k = 0; k < Q; k++)
                    *reinterpret_cast<float4*>(a.out[k] + roff + x0) = make_float4(o[k][0], o[k][1], o[k][2], o[k][3]);
            }
            bx += a.dr, by += a.dq;
            if (bx >= a.ntx) bx -= a.ntx, by++;
        }
    }

    // ---------------- one |u| reduction per launch ----------------
    // 64-bit parts: warp tree over shuffles; the bad-cell count fits REDUX
#pragma unroll
    for (int sh = 16; sh > 0; sh >>= 1) {
        acc_lo += __shfl_xor_sync(0xffffffffu, acc_lo, sh);
        acc_hi += __shfl_xor_sync(0xffffffffu, acc_hi, sh);
    }
    const unsigned nbad = __reduce_add_sync(0xffffffffu, acc_bad);
    if (lane == 0) {
        atomicAdd(&s_acc[0], acc_lo);
        atomicAdd(&s_acc[1], acc_hi);
        if (nbad) atomicAdd(&s_acc[2], static_cast<unsigned long long>(nbad));
    }
    __syncthreads();
    if (tid == 0) {
        const int s_abs = a.ctrl[0] + a.step_offset;
        unsigned long long* out =
            *a.sums_ref + (static_cast<size_t>(s_abs - a.ctrl[1]) * a.nslots + (blockIdx.x & (a.nslots - 1))) * SUM_WORDS;
        atomicAdd(&out[0], s_acc[0]);
        atomicAdd(&out[1], s_acc[1]);
        if (s_acc[2]) atomicAdd(&out[2], s_acc[2]);
    }
}

} // namespace lbm
