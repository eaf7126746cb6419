// lbm_ll_kernel.cuh -- step_ll_kernel: every timestep of a run in one cooperative launch, the lattice held in
// REGISTERS, rows exchanging their boundary populations through flagged 16-byte packets in L2 (the
// reference's small shipped grids, /root/reference/README.md:126-128).
//
// A timestep of a 128 x 128 grid is ~0.7 us of dependent arithmetic for one cell per thread; what step_loop_kernel
// adds to that is a grid-wide barrier (fence + atomic + acquire spin = three L2 round trips, ~1.2 us) and reading
// the lattice back from L2.  A cell only depends on the rows above and below it, so here
//   * CTA y owns row y for the whole run, thread x owns cell (x, y): its nine populations live in registers;
//   * after the collision the six populations that leave the cell sideways or diagonally are exchanged inside the
//     row through shared memory (one __syncthreads per step), and each thread then stores TWO 16-byte packets with
//     one 128-bit store each: {f2(x), f5(x-1), f6(x+1), flag} for the cell above it, {f4(x), f7(x+1), f8(x-1), flag}
//     for the cell below it -- exactly what that cell pulls across the row boundary (SerialCode/d2q9-bgk.c:263-279),
//     with flag = a step counter.  st / ld .relaxed.gpu .b128 are single-copy atomic: the flag travels WITH the
//     data, so there is no fence, no separate flag and no barrier -- the consumer polls its two packets until the flag
//     is the step it needs (one L2 round trip after the store lands), the "LL" protocol of collective libraries;
//   * packets are double buffered by step parity: row y overwrites its packet of step s when it has finished step
//     s+2, which needed row y+-1's packet of step s+1, which row y+-1 stored after ALL its threads had read step s
//     (the store comes after that step's __syncthreads);
//   * |u| of the new state is formed after the packets have been stored (it overlaps their flight); per-warp partial
//     sums in shared memory are added up one step later by three threads and go to sums[step][slot] with one RED each;
//   * accelerate-at-store as in the other kernels (applied before the exchange); the last step of the launch stores
//     the cell to the destination lattice in global memory.
// Cooperative launch: every CTA must be resident (rows <= resident CTAs), which also bounds the grids it takes.
// A poll that does not complete within the lattice's time-out sets the error word and gives up (lbm_sync reports it).
// Replaces the timestep loop of SerialCode/d2q9-bgk.c:187-194 for small single-GPU grids.
#pragma once

#include "lbm_kernels.cuh"

namespace lbm {

struct LLArgs {
    float* lat[2];       // two lattices in global memory, 9 planes each: [src] holds the state before first_step
    size_t pf;           // floats per plane
    const uint32_t* obst;
    unsigned long long* sums; // [nsteps][nslots][SUM_WORDS] of this launch
    int nslots;
    uint4* pk_north;     // [2][rows][pitch] packets of row r for row r+1
    uint4* pk_south;     // [2][rows][pitch] packets of row r for row r-1
    unsigned flag_base;  // flag of step s of this launch = flag_base + s + 1 (never reused by a lattice)
    int* error;
    unsigned long long timeout_ns;
    int first_step, nsteps, last_step; // absolute indices; no accelerate-at-store at last_step
    int src;
    int nx, rows, pitch, opitch;
    int accel_row;
    float omega, w1a, w2a;
};

__device__ __forceinline__ void st_packet(uint4* p, float a, float b, float c, unsigned flag)
{
    asm volatile(
        "{\n\t.reg .b128 q;\n\t.reg .b64 lo, hi;\n\t"
        "mov.b64 lo, {%1,%2};\n\tmov.b64 hi, {%3,%4};\n\tmov.b128 q, {lo, hi};\n\t"
        "st.relaxed.gpu.global.b128 [%0], q;\n\t}" ::"l"(p),
        "r"(__float_as_uint(a)), "r"(__float_as_uint(b)), "r"(__float_as_uint(c)), "r"(flag)
        : "memory");
}
__device__ __forceinline__ uint4 ld_packet(const uint4* p)
{
    uint4 v;
    asm volatile(
        "{\n\t.reg .b128 q;\n\t.reg .b64 lo, hi;\n\t"
        "ld.relaxed.gpu.global.b128 q, [%4];\n\t"
        "mov.b128 {lo, hi}, q;\n\tmov.b64 {%0,%1}, lo;\n\tmov.b64 {%2,%3}, hi;\n\t}"
        : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
        : "l"(p)
        : "memory");
    return v;
}

// |u| of a cell's new populations (SerialCode:425-452)
template <bool STRICT>
__device__ __forceinline__ float speed_cell(const float c[Q])
{
    if constexpr (STRICT) return speed_strict(c);
    else return speed_fast_guarded(c);
}

// dynamic shared memory: float xch[2][6][blockDim.x] (sideways populations, ping-pong by step parity)
//                        + unsigned part[2][blockDim.x / 32][4] (per-warp |u| sums, ping-pong)
template <bool STRICT>
__global__ void __launch_bounds__(1024) step_ll_kernel(const LLArgs a)
{
    extern __shared__ __align__(16) float ll_smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nthr = blockDim.x, nwarps = nthr >> 5;
    float* xch = ll_smem;
    unsigned* part = reinterpret_cast<unsigned*>(ll_smem + 2 * 6 * nthr);

    const int nx = a.nx, y = blockIdx.x;
    const bool valid = tid < nx;
    const int x = valid ? tid : nx - 1;
    const int xw = (x == 0) ? nx - 1 : x - 1; // SerialCode:259-262
    const int xe = (x == nx - 1) ? 0 : x + 1;
    const int ys = (y == 0) ? a.rows - 1 : y - 1; // :257-258
    const int yn = (y == a.rows - 1) ? 0 : y + 1;
    const size_t pitch = a.pitch;
    const bool solid = (__ldg(a.obst + static_cast<size_t>(y) * a.opitch + (x >> 5)) >> (x & 31)) & 1u;
    const bool on_accel_row = (y == a.accel_row);

    // the populations that stream into the cell at first_step: pulled from the source lattice
    float t[Q];
    {
        const float* in = a.lat[a.src & 1];
        const int col[Q] = {x, xw, x, xe, x, xw, xe, xe, xw};
        const int row[Q] = {y, y, ys, y, yn, ys, ys, yn, yn};
#pragma unroll
        for (int k = 0; k < Q; k++) t[k] = __ldcg(in + k * a.pf + static_cast<size_t>(row[k]) * pitch + col[k]);
    }
    float* out = a.lat[(a.src + a.nsteps) & 1];
    const size_t rows_pitch = static_cast<size_t>(a.rows) * pitch;
    uint4* const my_north = a.pk_north + static_cast<size_t>(y) * pitch + x;
    uint4* const my_south = a.pk_south + static_cast<size_t>(y) * pitch + x;
    const uint4* const from_south = a.pk_north + static_cast<size_t>(ys) * pitch + x; // what the row below sends up
    const uint4* const from_north = a.pk_south + static_cast<size_t>(yn) * pitch + x; // what the row above sends down
    bool lost = false;

    for (int s = 0; s < a.nsteps; s++) {
        const int par = s & 1;
        const bool last = (s + 1 == a.nsteps);
        // ---- collision / bounce-back (update_cell without |u|), accelerate_flow() of the next step
        float c[Q], o[Q];
        collide_cell<STRICT, false>(t, a.omega, c);
        o[0] = solid ? t[0] : c[0];
        o[1] = solid ? t[3] : c[1];
        o[2] = solid ? t[4] : c[2];
        o[3] = solid ? t[1] : c[3];
        o[4] = solid ? t[2] : c[4];
        o[5] = solid ? t[7] : c[5];
        o[6] = solid ? t[8] : c[6];
        o[7] = solid ? t[5] : c[7];
        o[8] = solid ? t[6] : c[8];
        float f[Q]; // what is stored / sent: accelerated on the driven row
#pragma unroll
        for (int k = 0; k < Q; k++) f[k] = o[k];
        if (on_accel_row && (a.first_step + s != a.last_step)) accelerate_cell(f, solid, a.w1a, a.w2a);

        float* xs = xch + static_cast<size_t>(par) * 6 * nthr;
        if (!last) {
            // ---- sideways exchange inside the row
            xs[0 * nthr + tid] = f[1], xs[1 * nthr + tid] = f[3], xs[2 * nthr + tid] = f[5];
            xs[3 * nthr + tid] = f[6], xs[4 * nthr + tid] = f[7], xs[5 * nthr + tid] = f[8];
        } else if (valid) {
            const size_t off = static_cast<size_t>(y) * pitch + x;
#pragma unroll
            for (int k = 0; k < Q; k++) out[k * a.pf + off] = f[k];
        }
        __syncthreads();
        const unsigned flag = a.flag_base + static_cast<unsigned>(s) + 1u;
        float t1 = 0.f, t3 = 0.f;
        if (!last) {
            t1 = xs[0 * nthr + xw], t3 = xs[1 * nthr + xe];
            const float p5 = xs[2 * nthr + xw], p6 = xs[3 * nthr + xe], p7 = xs[4 * nthr + xe], p8 = xs[5 * nthr + xw];
            if (valid) {
                st_packet(my_north + par * rows_pitch, f[2], p5, p6, flag);
                st_packet(my_south + par * rows_pitch, f[4], p7, p8, flag);
            }
        }
        // ---- the |u| sums of the previous step are complete in shared memory (this step's barrier): three threads
        // add up the warps' parts
        if (s > 0 && tid < 3) {
            const unsigned* pp = part + static_cast<size_t>(par ^ 1) * nwarps * 4;
            unsigned long long v = 0ull;
            for (int w = 0; w < nwarps; w++) v += pp[w * 4 + tid];
            if (v) atomicAdd(a.sums + (static_cast<size_t>(s - 1) * a.nslots + (blockIdx.x & (a.nslots - 1))) * SUM_WORDS + tid, v);
        }
        // ---- |u| of the collided state, while the packets travel
        {
            const float sp = speed_cell<STRICT>(c);
            SpeedAcc acc = {0u, 0u, 0u};
            acc_speed(acc, sp, valid && !solid);
            const unsigned lo = __reduce_add_sync(0xffffffffu, acc.lo);
            const unsigned hi = __reduce_add_sync(0xffffffffu, acc.hi);
            const unsigned nbad = __reduce_add_sync(0xffffffffu, acc.bad);
            if (lane == 0) {
                unsigned* pp = part + (static_cast<size_t>(par) * nwarps + warp) * 4;
                pp[0] = lo, pp[1] = hi, pp[2] = nbad;
            }
        }
        if (last) break;
        // ---- wait for the two packets of this step from the rows below and above
        {
            const uint4* ps = from_south + par * rows_pitch;
            const uint4* pn = from_north + par * rows_pitch;
            uint4 S = ld_packet(ps), N = ld_packet(pn);
            bool ok_s = (S.w == flag), ok_n = (N.w == flag);
            if (!(ok_s && ok_n) && !lost) {
                const unsigned long long t0 = globaltimer_ns();
                unsigned spins = 0;
                while (true) {
                    if (!ok_s) S = ld_packet(ps), ok_s = (S.w == flag);
                    if (!ok_n) N = ld_packet(pn), ok_n = (N.w == flag);
                    if (ok_s && ok_n) break;
                    if ((++spins & 255u) == 0u &&
                        (globaltimer_ns() - t0 > a.timeout_ns || *reinterpret_cast<volatile const int*>(a.error))) {
                        atomicExch(a.error, 1);
                        lost = true; // the run is lost: finish without waiting any more
                        break;
                    }
                }
            }
            t[0] = f[0], t[1] = t1, t[3] = t3;
            t[2] = __uint_as_float(S.x), t[5] = __uint_as_float(S.y), t[6] = __uint_as_float(S.z);
            t[4] = __uint_as_float(N.x), t[7] = __uint_as_float(N.y), t[8] = __uint_as_float(N.z);
        }
    }
    // the last step's sums
    __syncthreads();
    if (tid < 3 && a.nsteps > 0) {
        const unsigned* pp = part + static_cast<size_t>((a.nsteps - 1) & 1) * nwarps * 4;
        unsigned long long v = 0ull;
        for (int w = 0; w < nwarps; w++) v += pp[w * 4 + tid];
        if (v) atomicAdd(a.sums + (static_cast<size_t>(a.nsteps - 1) * a.nslots + (blockIdx.x & (a.nslots - 1))) * SUM_WORDS + tid, v);
    }
}

} // namespace lbm
