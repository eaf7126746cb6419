// lbm_band_kernel.cuh -- step_band_kernel: every timestep of a run in one cooperative launch for grids that live in
// L2 but not on the SMs (the reference's 1024 x 1024 case), WITHOUT a grid-wide barrier.
//
// step_loop_kernel (lbm_kernels.cuh) deals tiles out round-robin and crosses a one-counter barrier every step: all
// warps of the GPU load together, compute together and wait together (ncu, round 1: barrier stalls 3.8 per issue,
// issue slots 51 % busy).  A row only depends on the rows above and below it, so here
//   * CTA b owns a contiguous band of rows for the whole run (rows / CTAs, balanced to within one row);
//   * a step = first the band's first and last row (the only rows other CTAs read), then a CTA barrier, a fence and
//     a release store of the step number into the CTA's flag, then the interior rows;
//   * before a step the CTA waits until the flags of the bands below and above say that THEIR boundary rows of the
//     previous step are published.  That one condition covers both hazards: what this step reads from the
//     neighbours exists, and the neighbours have finished reading -- in their previous step's boundary rows -- the
//     rows of the lattice this step overwrites (two lattices, ping-pong).  The flags were published a whole band
//     interior earlier, so the wait is normally over before it starts and CTAs drift apart instead of marching in
//     lockstep: one CTA's loads overlap another's arithmetic on the same SM;
//   * tile shape, gather (pull4: LDG.128 through L2 + shuffles), arithmetic (update4) and |u| sums as in
//     step_vec4_kernel / step_loop_kernel: the strict flavour stays bit-identical to SerialCode.
// A wait that does not complete within the lattice's time-out sets the error word and gives up.
// Replaces the timestep loop of SerialCode/d2q9-bgk.c:187-194 for single-GPU grids of up to LOOP_MAX_CELLS cells.
#pragma once

#include "lbm_ll_kernel.cuh" // st_packet / ld_packet / ll_seed_kernel: the packets that cross a GPU boundary

namespace lbm {

struct BandArgs {
    float* lat[2];       // two lattices, 9 planes each
    size_t pf;           // floats per plane
    const uint32_t* obst;
    unsigned long long* sums; // [nsteps][nslots][SUM_WORDS] of this launch
    int nslots;
    unsigned* flags;     // [gridDim.x][32] (a 128-byte line per CTA), zeroed before the launch: word 0 = steps of this
                         // launch whose boundary rows the CTA has published
    int* error;
    unsigned long long timeout_ns;
    int first_step, nsteps, last_step; // absolute indices; no accelerate-at-store at last_step
    int src;             // lattice that holds the state before first_step
    int nx, nxv, rows, pitch, opitch;
    int accel_row;
    float omega, w1a, w2a;
    // row slabs on several GPUs (HALO): the slab's first / last row exchange step_ll_kernel's UNSHIFTED packets with the
    // neighbouring GPUs ({f2, f5, f6} / {f4, f7, f8} of one cell + flag, st / ld.relaxed.sys.b128 into / out of the packet
    // areas [2][pitch] of the halo blocks; the consumer reads the packets of the cells west and east of its own as well).
    // Step s consumes flag_base + s from slot (s + 1) & 1 (s = 0: what ll_seed_kernel sent) and produces flag_base + s + 1
    // in slot s & 1.
    const uint4* halo_recv_s;
    const uint4* halo_recv_n;
    uint4* halo_send_s;
    uint4* halo_send_n;
    unsigned flag_base;
    // SM-aware band sizes (per_sm > 0: the grid is exactly per_sm CTAs on each of nsm SMs): CTAs find out which SM they
    // run on and the rows are dealt out per SM first, so that every SM has the same number of rows to within one
    int per_sm, nsm;
    unsigned* sm_table;  // [1 + 2 * 1024 + 2] zeroed before the launch: dense SM counter, per-%smid CTA counter, per-%smid dense
                         // id + 1, CTAs that have a ticket, placement-irregular flag
};

__device__ __forceinline__ void st_release_gpu_u32(unsigned* p, unsigned v)
{
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

template <bool STRICT, int BLOCK, int MINB, bool VERT, bool HALO>
__global__ void __launch_bounds__(BLOCK, MINB) step_band_kernel(const BandArgs a)
{
    __shared__ unsigned long long s_part[2][BLOCK / 32][3];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int G = gridDim.x;
    int b = blockIdx.x, r0, nr;
    if (a.per_sm > 0) {
        // which SM am I on, and the how-manieth CTA there?  (cooperative launch at full occupancy: exactly per_sm each)
        __shared__ int s_b;
        if (tid == 0) {
            unsigned smid;
            asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
            smid &= 1023u;
            const unsigned rank = atomicAdd(a.sm_table + 1 + smid, 1u);
            unsigned* dense_p = a.sm_table + 1 + 1024 + smid;
            if (rank == 0) atomicExch(dense_p, atomicAdd(a.sm_table, 1u) + 1u);
            unsigned dense;
            while ((dense = ld_acquire_gpu_u32(dense_p)) == 0u) {
            }
            // one grid-wide rendezvous per launch: if the placement is not what a full-occupancy cooperative launch
            // gives (more CTAs on an SM than expected, fewer SMs in use), everybody falls back to bands by blockIdx
            unsigned* arrived = a.sm_table + 1 + 2 * 1024, *irregular = arrived + 1;
            if (rank >= static_cast<unsigned>(a.per_sm)) atomicExch(irregular, 1u);
            __threadfence();
            atomicAdd(arrived, 1u);
            while (ld_acquire_gpu_u32(arrived) < static_cast<unsigned>(G)) {
            }
            const bool bad = ld_acquire_gpu_u32(irregular) != 0u || ld_acquire_gpu_u32(a.sm_table) != static_cast<unsigned>(a.nsm);
            s_b = bad ? -1 : static_cast<int>(dense - 1u) * a.per_sm + static_cast<int>(rank);
        }
        __syncthreads();
        b = s_b;
    }
    if (a.per_sm > 0 && b >= 0) {
        // rows per SM first (balanced to within one), then per CTA of the SM
        const int sm = b / a.per_sm, k = b - sm * a.per_sm;
        const int qs = a.rows / a.nsm, rs = a.rows % a.nsm;
        const int sm_r0 = sm * qs + min(sm, rs), sm_nr = qs + (sm < rs ? 1 : 0);
        const int qc = sm_nr / a.per_sm, rc = sm_nr % a.per_sm;
        r0 = sm_r0 + k * qc + min(k, rc);
        nr = qc + (k < rc ? 1 : 0); // >= 1: the host asks for this only with rows >= CTAs
    } else {
        b = blockIdx.x;
        const int q = a.rows / G, rem = a.rows % G;
        r0 = b * q + min(b, rem);
        nr = q + (b < rem ? 1 : 0); // >= 1: the host launches at most `rows` CTAs
    }
    const unsigned* flag_s = a.flags + static_cast<size_t>((b + G - 1) % G) * 32;
    const unsigned* flag_n = a.flags + static_cast<size_t>((b + 1) % G) * 32;
    unsigned* flag_own = a.flags + static_cast<size_t>(b) * 32;
    const size_t pitch = a.pitch;
    const HaloCfg h = HaloCfg{}; // single slab: periodic in y inside the lattice
    bool lost = false; // a wait gave up: the run is lost, nobody waits any more

    for (int s = 0; s < a.nsteps; s++) {
        const float* in = a.lat[(a.src + s) & 1];
        float* out = a.lat[(a.src + s + 1) & 1];
        const bool accel_live = (a.first_step + s != a.last_step);
        if (s > 0) {
            if (tid == 0 && !lost) {
                const unsigned need = static_cast<unsigned>(s);
                bool ok_s = false, ok_n = false;
                const unsigned long long t0 = globaltimer_ns();
                unsigned spins = 0;
                while (true) {
                    if (!ok_s) ok_s = ld_acquire_gpu_u32(flag_s) >= need;
                    if (!ok_n) ok_n = ld_acquire_gpu_u32(flag_n) >= need;
                    if (ok_s && ok_n) break;
                    if ((++spins & 255u) == 0u &&
                        (globaltimer_ns() - t0 > a.timeout_ns || *reinterpret_cast<volatile const int*>(a.error))) {
                        atomicExch(a.error, 1);
                        lost = true;
                        break;
                    }
                }
            }
            __syncthreads();
        }
        unsigned long long acc_lo = 0ull, acc_hi = 0ull;
        unsigned acc_bad = 0u;
        // the band's rows, the two that other CTAs read first; every thread four cells at a time
        const int publish_after = min(1, nr - 1);
#pragma unroll 1
        for (int i = 0; i < nr; i++) {
            const int y = (i == 0) ? r0 : ((i == 1) ? r0 + nr - 1 : r0 + i - 1);
            const PullRows rows = pull_rows(in, a.pf, a.rows, pitch, h, y, 0);
            const size_t roff = static_cast<size_t>(y) * pitch;
            const bool accel = accel_live && (y == a.accel_row);
#pragma unroll 1
            for (int c0 = 0; c0 < a.nxv; c0 += BLOCK) {
                const int c_raw = c0 + tid;
                const bool valid = c_raw < a.nxv;
                const int c = min(c_raw, a.nxv - 1);
                const uint32_t oword = __ldg(a.obst + static_cast<size_t>(y) * a.opitch + (c >> 3));
                float t[Q][4];
                pull4<3>(rows, c, a.nxv, a.nx, lane, tid, BLOCK, t); // through L2: the source changes every step
                if (HALO && (y == 0 || y == a.rows - 1)) {
                    // what crosses the GPU boundary comes out of the neighbour's packets (the periodic wrap inside the
                    // slab that pull4 just read for those planes is not the neighbour row)
                    const int x0 = 4 * c;
                    const int off[6] = {(x0 == 0) ? a.nx - 1 : x0 - 1, x0, x0 + 1, x0 + 2, x0 + 3, (x0 + 4 == a.nx) ? 0 : x0 + 4};
                    const unsigned want = a.flag_base + static_cast<unsigned>(s);
                    const size_t slot = static_cast<size_t>((s + 1) & 1) * pitch;
#pragma unroll
                    for (int side = 0; side < 2; side++) {
                        if (side == 0 ? (y != 0) : (y != a.rows - 1)) continue;
                        const uint4* src = (side == 0 ? a.halo_recv_s : a.halo_recv_n) + slot;
                        uint4 P[6];
                        unsigned pending = 0x3fu, spins = 0;
                        unsigned long long t0 = 0ull;
                        while (pending != 0u && !lost) {
#pragma unroll
                            for (int i = 0; i < 6; i++)
                                if (pending & (1u << i)) P[i] = ld_packet<true>(src + off[i]);
#pragma unroll
                            for (int i = 0; i < 6; i++)
                                if ((pending & (1u << i)) && P[i].w == want) pending &= ~(1u << i);
                            if (pending != 0u && (++spins & 255u) == 0u) {
                                const unsigned long long now = globaltimer_ns();
                                if (t0 == 0ull) t0 = now;
                                if (now - t0 > a.timeout_ns || *reinterpret_cast<volatile const int*>(a.error)) {
                                    atomicExch(a.error, 1);
                                    lost = true;
                                }
                            }
                        }
#pragma unroll
                        for (int j = 0; j < 4; j++) {
                            if (side == 0) {
                                t[2][j] = __uint_as_float(P[1 + j].x), t[5][j] = __uint_as_float(P[j].y), t[6][j] = __uint_as_float(P[2 + j].z);
                            } else {
                                t[4][j] = __uint_as_float(P[1 + j].x), t[7][j] = __uint_as_float(P[2 + j].y), t[8][j] = __uint_as_float(P[j].z);
                            }
                        }
                    }
                }
                const uint32_t obits = (oword >> ((c & 7) * 4)) & 0xfu;
                float o[Q][4];
                SpeedAcc acc = {0u, 0u, 0u};
                update4<STRICT, VERT>(t, obits, valid, accel, a.omega, a.w1a, a.w2a, o, acc);
                if (valid) store4<3>(out, a.pf, roff + 4 * c, o);
                if (HALO && valid && s + 1 < a.nsteps && (y == 0 || y == a.rows - 1)) {
                    const unsigned flag = a.flag_base + static_cast<unsigned>(s) + 1u;
                    const size_t slot = static_cast<size_t>(s & 1) * pitch + 4 * c;
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        if (y == 0) st_packet<true>(a.halo_send_s + slot + j, o[4][j], o[7][j], o[8][j], flag);
                        if (y == a.rows - 1) st_packet<true>(a.halo_send_n + slot + j, o[2][j], o[5][j], o[6][j], flag);
                    }
                }
                acc_lo += acc.lo, acc_hi += acc.hi, acc_bad += acc.bad;
            }
            if (i == publish_after && s + 1 < a.nsteps) {
                __syncthreads(); // every thread's stores of the boundary rows have been issued
                if (tid == 0) {
                    __threadfence();
                    st_release_gpu_u32(flag_own, static_cast<unsigned>(s + 1));
                }
            }
        }

        // this CTA's |u| sums of the step
#pragma unroll
        for (int sh = 16; sh > 0; sh >>= 1) {
            acc_lo += __shfl_xor_sync(0xffffffffu, acc_lo, sh);
            acc_hi += __shfl_xor_sync(0xffffffffu, acc_hi, sh);
        }
        const unsigned nbad = __reduce_add_sync(0xffffffffu, acc_bad);
        if (lane == 0) {
            unsigned long long* pp = s_part[s & 1][warp];
            pp[0] = acc_lo, pp[1] = acc_hi, pp[2] = nbad;
        }
        __syncthreads(); // the step's stores are visible to this CTA's next step; the parts are complete
        if (tid < 3) {
            unsigned long long v = 0ull;
            for (int w = 0; w < BLOCK / 32; w++) v += s_part[s & 1][w][tid];
            if (v) atomicAdd(a.sums + (static_cast<size_t>(s) * a.nslots + (b & (a.nslots - 1))) * SUM_WORDS + tid, v);
        }
    }
}

} // namespace lbm
