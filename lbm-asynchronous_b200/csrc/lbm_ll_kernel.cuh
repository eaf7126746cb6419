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
//   * CPT = 2 / 4: a thread owns two / four consecutive cells (packed collision of lbm_collide4.cuh in two halves, CPT
//     packets per direction): a half / a quarter of the threads per row, so grids with more rows than CTAs of nx threads
//     can be resident.  Two cells per thread (512 threads for a 1024-cell row at 64 registers, two CTAs per SM) is the
//     default beyond the smallest grids: twice the warps of the four-cell variant for the same instructions per cell
//     (1024 x 256: 4.6 us per step against 6.1, profiles/r02_small_grids.md);
//   * row slabs on several GPUs (HALO): the slab's first / last row store their packets straight into the neighbour
//     GPU's packet area (peer memory over NVLink, st.relaxed.sys.b128) and poll their own, which the neighbour
//     writes -- the same protocol at system scope, replacing MPI_Isend/Irecv/Waitall of MPI_Waitall/d2q9-bgk.c:225-253
//     for small slabs.  A run starts with ll_seed_kernel sending the packets of the CURRENT state (no step produced
//     them), which the boundary rows' first gather waits for: that is also what lines the GPUs up.
// Cooperative launch: every CTA must be resident (rows <= resident CTAs), which also bounds the grids it takes.
// A poll that does not complete within the lattice's time-out sets the error word and gives up (lbm_sync reports it).
// Replaces the timestep loop of SerialCode/d2q9-bgk.c:187-194 for small single-GPU grids.
#pragma once

#include "lbm_cluster_kernel.cuh" // cells_relax / cells_speed: the collision in two halves

namespace lbm {

struct LLArgs {
    float* lat[2];       // two lattices in global memory, 9 planes each: [src] holds the state before first_step
    size_t pf;           // floats per plane
    const uint32_t* obst;
    unsigned long long* sums; // [nsteps][nslots][SUM_WORDS] of this launch
    int nslots;
    uint4* pk_north;     // [2][rows][pitch] packets of row r for row r+1
    uint4* pk_south;     // [2][rows][pitch] packets of row r for row r-1
    // row slabs on several GPUs (HALO): packet areas [2][pitch] of the slab's two boundaries
    const uint4* halo_recv_s; // what the south neighbour's last row sends up (it writes it: peer stores)
    const uint4* halo_recv_n; // what the north neighbour's first row sends down
    uint4* halo_send_s;       // the south neighbour's halo_recv_n (peer memory)
    uint4* halo_send_n;       // the north neighbour's halo_recv_s
    unsigned seed_flag;  // flag of the packets ll_seed_kernel sent for the state before first_step (parity slot 1)
    unsigned flag_base;  // flag of step s of this launch = flag_base + s + 1 (never reused by a lattice)
    int* error;
    unsigned long long timeout_ns;
    int first_step, nsteps, last_step; // absolute indices; no accelerate-at-store at last_step
    int src;
    int nx, rows, pitch, opitch;
    int accel_row;
    float omega, w1a, w2a;
};

template <bool SYS>
__device__ __forceinline__ void st_packet(uint4* p, float a, float b, float c, unsigned flag)
{
    if constexpr (SYS) {
        asm volatile(
            "{\n\t.reg .b128 q;\n\t.reg .b64 lo, hi;\n\t"
            "mov.b64 lo, {%1,%2};\n\tmov.b64 hi, {%3,%4};\n\tmov.b128 q, {lo, hi};\n\t"
            "st.relaxed.sys.global.b128 [%0], q;\n\t}" ::"l"(p),
            "r"(__float_as_uint(a)), "r"(__float_as_uint(b)), "r"(__float_as_uint(c)), "r"(flag)
            : "memory");
    } else {
        asm volatile(
            "{\n\t.reg .b128 q;\n\t.reg .b64 lo, hi;\n\t"
            "mov.b64 lo, {%1,%2};\n\tmov.b64 hi, {%3,%4};\n\tmov.b128 q, {lo, hi};\n\t"
            "st.relaxed.gpu.global.b128 [%0], q;\n\t}" ::"l"(p),
            "r"(__float_as_uint(a)), "r"(__float_as_uint(b)), "r"(__float_as_uint(c)), "r"(flag)
            : "memory");
    }
}
template <bool SYS>
__device__ __forceinline__ uint4 ld_packet(const uint4* p)
{
    uint4 v;
    if constexpr (SYS) {
        asm volatile(
            "{\n\t.reg .b128 q;\n\t.reg .b64 lo, hi;\n\t"
            "ld.relaxed.sys.global.b128 q, [%4];\n\t"
            "mov.b128 {lo, hi}, q;\n\tmov.b64 {%0,%1}, lo;\n\tmov.b64 {%2,%3}, hi;\n\t}"
            : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
            : "l"(p)
            : "memory");
    } else {
        asm volatile(
            "{\n\t.reg .b128 q;\n\t.reg .b64 lo, hi;\n\t"
            "ld.relaxed.gpu.global.b128 q, [%4];\n\t"
            "mov.b128 {lo, hi}, q;\n\tmov.b64 {%0,%1}, lo;\n\tmov.b64 {%2,%3}, hi;\n\t}"
            : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
            : "l"(p)
            : "memory");
    }
    return v;
}

// Wait for the N packets at ps[0..N) and pn[0..N) to carry `flag` (sys_s / sys_n: written by a peer GPU).  Every
// missing packet is re-read per round, all loads of a round in flight together.  Gives up (lost = true, error word
// set) when the lattice's time-out passes: the run is lost, later waits return at once.
template <int N, bool HALO>
__device__ __forceinline__ void wait_packets(const uint4* ps, const uint4* pn, bool want_s, bool want_n, bool sys_s, bool sys_n, unsigned flag,
                                             uint4 (&S)[N], uint4 (&NN)[N], bool& lost, const LLArgs& a)
{
    // first round: straight-line loads, all in flight together (the common case needs no second one with four cells
    // per thread, one or two with one cell)
    if (want_s) {
#pragma unroll
        for (int j = 0; j < N; j++) S[j] = (HALO && sys_s) ? ld_packet<true>(ps + j) : ld_packet<false>(ps + j);
    }
    if (want_n) {
#pragma unroll
        for (int j = 0; j < N; j++) NN[j] = (HALO && sys_n) ? ld_packet<true>(pn + j) : ld_packet<false>(pn + j);
    }
    unsigned pending = 0u;
#pragma unroll
    for (int j = 0; j < N; j++) {
        if (want_s && S[j].w != flag) pending |= 1u << j;
        if (want_n && NN[j].w != flag) pending |= 1u << (N + j);
    }
    unsigned long long t0 = 0ull;
    unsigned spins = 0;
    while (pending != 0u && !lost) {
#pragma unroll
        for (int j = 0; j < N; j++) {
            if (pending & (1u << j)) S[j] = (HALO && sys_s) ? ld_packet<true>(ps + j) : ld_packet<false>(ps + j);
            if (pending & (1u << (N + j))) NN[j] = (HALO && sys_n) ? ld_packet<true>(pn + j) : ld_packet<false>(pn + j);
        }
#pragma unroll
        for (int j = 0; j < N; j++) {
            if ((pending & (1u << j)) && S[j].w == flag) pending &= ~(1u << j);
            if ((pending & (1u << (N + j))) && NN[j].w == flag) pending &= ~(1u << (N + j));
        }
        if ((++spins & 255u) == 0u) { // not on the fast path: the clock and the error word cost an L2 round trip
            const unsigned long long now = globaltimer_ns();
            if (t0 == 0ull) t0 = now;
            if (now - t0 > a.timeout_ns || *reinterpret_cast<volatile const int*>(a.error)) {
                atomicExch(a.error, 1);
                lost = true;
            }
        }
    }
}

// dynamic shared memory: float xch[2][6][blockDim.x] (what leaves a thread's cells sideways, ping-pong by step parity)
//                        + unsigned part[2][blockDim.x / 32][4] (per-warp |u| sums, ping-pong)
template <bool STRICT, int CPT, bool VERT, int MAXT, int MINB, bool HALO>
__global__ void __launch_bounds__(MAXT, MINB) step_ll_kernel(const LLArgs a)
{
    extern __shared__ __align__(16) float ll_smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nthr = blockDim.x, nwarps = nthr >> 5;
    float* xch = ll_smem;
    unsigned* part = reinterpret_cast<unsigned*>(ll_smem + 2 * 6 * nthr);

    const int nx = a.nx, y = blockIdx.x;
    const int ipr = nx / CPT; // threads that own cells
    const bool valid = tid < ipr;
    const int xi = valid ? tid : ipr - 1;
    const int x0 = xi * CPT;
    const int tw = (xi == 0) ? ipr - 1 : xi - 1; // the threads that own the cells west / east of mine (periodic, SerialCode:259-262)
    const int te = (xi == ipr - 1) ? 0 : xi + 1;
    const int ys = (y == 0) ? a.rows - 1 : y - 1; // :257-258
    const int yn = (y == a.rows - 1) ? 0 : y + 1;
    const size_t pitch = a.pitch;
    const uint32_t obits = (__ldg(a.obst + static_cast<size_t>(y) * a.opitch + (x0 >> 5)) >> (x0 & 31)) & ((1u << CPT) - 1u);
    const bool on_accel_row = (y == a.accel_row);
    const bool peer_s = HALO && (y == 0), peer_n = HALO && (y == a.rows - 1); // rows whose neighbour row lives on another GPU

    const size_t rows_pitch = static_cast<size_t>(a.rows) * pitch;
    // where my packets go / come from, parity slot 0; + par * stride for the other
    uint4* const my_north = peer_n ? a.halo_send_n + x0 : a.pk_north + static_cast<size_t>(y) * pitch + x0;
    uint4* const my_south = peer_s ? a.halo_send_s + x0 : a.pk_south + static_cast<size_t>(y) * pitch + x0;
    const uint4* const from_south = peer_s ? a.halo_recv_s + x0 : a.pk_north + static_cast<size_t>(ys) * pitch + x0;
    const uint4* const from_north = peer_n ? a.halo_recv_n + x0 : a.pk_south + static_cast<size_t>(yn) * pitch + x0;
    const size_t stride_s = peer_s ? pitch : rows_pitch, stride_n = peer_n ? pitch : rows_pitch;
    bool lost = false;

    // the populations that stream into the cells at first_step: pulled from the source lattice; across a slab
    // boundary from the packets ll_seed_kernel of the neighbouring GPU sent (parity slot 1)
    float t[Q][CPT];
    {
        const float* in = a.lat[a.src & 1];
#pragma unroll
        for (int j = 0; j < CPT; j++) {
            const int x = x0 + j;
            const int xw = (x == 0) ? nx - 1 : x - 1, xe = (x == nx - 1) ? 0 : x + 1;
            const int col[Q] = {x, xw, x, xe, x, xw, xe, xe, xw};
            const int row[Q] = {y, y, ys, y, yn, ys, ys, yn, yn};
#pragma unroll
            for (int k = 0; k < Q; k++) t[k][j] = __ldcg(in + k * a.pf + static_cast<size_t>(row[k]) * pitch + col[k]);
        }
        if (HALO && (peer_s || peer_n)) {
            uint4 S[CPT], N[CPT];
            wait_packets<CPT, HALO>(from_south + stride_s, from_north + stride_n, peer_s, peer_n, true, true, a.seed_flag, S, N, lost, a);
#pragma unroll
            for (int j = 0; j < CPT; j++) {
                if (peer_s) t[2][j] = __uint_as_float(S[j].x), t[5][j] = __uint_as_float(S[j].y), t[6][j] = __uint_as_float(S[j].z);
                if (peer_n) t[4][j] = __uint_as_float(N[j].x), t[7][j] = __uint_as_float(N[j].y), t[8][j] = __uint_as_float(N[j].z);
            }
        }
    }
    float* out = a.lat[(a.src + a.nsteps) & 1];

    for (int s = 0; s < a.nsteps; s++) {
        const int par = s & 1;
        const bool last = (s + 1 == a.nsteps);
        // ---- collision / bounce-back without |u|; f = what is stored / sent: accelerate_flow() of the next step applied
        // on the driven row
        float o[Q][CPT], speed[CPT];
        bool speed_done;
        if constexpr (CPT == 1) {
            // one cell: the scalar collision without its |u| half (formed from o after the packets have left)
            float tc[Q], c[Q];
#pragma unroll
            for (int k = 0; k < Q; k++) tc[k] = t[k][0];
            collide_cell<STRICT, false>(tc, a.omega, c);
            constexpr int mirror[Q] = {0, 3, 4, 1, 2, 7, 8, 5, 6};
            const bool solid = obits & 1u;
#pragma unroll
            for (int k = 0; k < Q; k++) o[k][0] = solid ? tc[mirror[k]] : c[k];
            speed_done = false;
        } else {
            speed_done = cells_relax<STRICT, CPT, VERT>(t, obits, a.omega, o, speed);
        }
        float f[Q][CPT];
#pragma unroll
        for (int k = 0; k < Q; k++)
#pragma unroll
            for (int j = 0; j < CPT; j++) f[k][j] = o[k][j];
        if (on_accel_row && (a.first_step + s != a.last_step)) {
#pragma unroll
            for (int j = 0; j < CPT; j++) {
                float fc[Q];
#pragma unroll
                for (int k = 0; k < Q; k++) fc[k] = f[k][j];
                accelerate_cell(fc, (obits >> j) & 1u, a.w1a, a.w2a);
#pragma unroll
                for (int k = 0; k < Q; k++) f[k][j] = fc[k];
            }
        }

        float* xs = xch + static_cast<size_t>(par) * 6 * nthr;
        if (!last) {
            // ---- what leaves my cells sideways: f1, f5, f8 of the last cell go east, f3, f6, f7 of the first go west
            xs[0 * nthr + tid] = f[1][CPT - 1], xs[1 * nthr + tid] = f[3][0], xs[2 * nthr + tid] = f[5][CPT - 1];
            xs[3 * nthr + tid] = f[6][0], xs[4 * nthr + tid] = f[7][0], xs[5 * nthr + tid] = f[8][CPT - 1];
        } else if (valid) {
            float* dst = out + static_cast<size_t>(y) * pitch + x0;
            store_cells<CPT>(dst, a.pf, f);
        }
        __syncthreads();
        const unsigned flag = a.flag_base + static_cast<unsigned>(s) + 1u;
        float f1w = 0.f, f3e = 0.f;
        if (!last) {
            f1w = xs[0 * nthr + tw], f3e = xs[1 * nthr + te];
            const float f5w = xs[2 * nthr + tw], f6e = xs[3 * nthr + te], f7e = xs[4 * nthr + te], f8w = xs[5 * nthr + tw];
            if (valid) {
                uint4* pn = my_north + par * stride_n;
                uint4* ps = my_south + par * stride_s;
#pragma unroll
                for (int j = 0; j < CPT; j++) {
                    const float a5 = (j > 0) ? f[5][j > 0 ? j - 1 : 0] : f5w, a6 = (j + 1 < CPT) ? f[6][j + 1 < CPT ? j + 1 : 0] : f6e;
                    const float a7 = (j + 1 < CPT) ? f[7][j + 1 < CPT ? j + 1 : 0] : f7e, a8 = (j > 0) ? f[8][j > 0 ? j - 1 : 0] : f8w;
                    if (peer_n) st_packet<true>(pn + j, f[2][j], a5, a6, flag);
                    else st_packet<false>(pn + j, f[2][j], a5, a6, flag);
                    if (peer_s) st_packet<true>(ps + j, f[4][j], a7, a8, flag);
                    else st_packet<false>(ps + j, f[4][j], a7, a8, flag);
                }
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
            if constexpr (CPT == 1) {
                float c[Q];
#pragma unroll
                for (int k = 0; k < Q; k++) c[k] = o[k][0];
                speed[0] = STRICT ? speed_strict(c) : speed_fast_guarded(c); // an obstacle cell's value is not counted
            } else {
                if (!speed_done) cells_speed<STRICT, CPT, VERT>(o, obits, speed);
            }
            SpeedAcc acc = {0u, 0u, 0u};
#pragma unroll
            for (int j = 0; j < CPT; j++) acc_speed(acc, speed[j], valid && !((obits >> j) & 1u));
            const unsigned lo = __reduce_add_sync(0xffffffffu, acc.lo);
            const unsigned hi = __reduce_add_sync(0xffffffffu, acc.hi);
            const unsigned nbad = __reduce_add_sync(0xffffffffu, acc.bad);
            if (lane == 0) {
                unsigned* pp = part + (static_cast<size_t>(par) * nwarps + warp) * 4;
                pp[0] = lo, pp[1] = hi, pp[2] = nbad;
            }
        }
        if (last) break;
        // ---- wait for this step's packets from the rows below and above, assemble the next step's gather
        {
            uint4 S[CPT], N[CPT];
            wait_packets<CPT, HALO>(from_south + par * stride_s, from_north + par * stride_n, true, true, peer_s, peer_n, flag, S, N, lost, a);
#pragma unroll
            for (int j = 0; j < CPT; j++) {
                t[0][j] = f[0][j];
                t[1][j] = (j > 0) ? f[1][j > 0 ? j - 1 : 0] : f1w;
                t[3][j] = (j + 1 < CPT) ? f[3][j + 1 < CPT ? j + 1 : 0] : f3e;
                t[2][j] = __uint_as_float(S[j].x), t[5][j] = __uint_as_float(S[j].y), t[6][j] = __uint_as_float(S[j].z);
                t[4][j] = __uint_as_float(N[j].x), t[7][j] = __uint_as_float(N[j].y), t[8][j] = __uint_as_float(N[j].z);
            }
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

// Row slabs on several GPUs, start of a run: the packets of the CURRENT state of the slab's first and last row go to
// the neighbouring GPUs (parity slot 1, flag `seed_flag`), exactly what a step would have sent
struct LLSeedArgs {
    const float* lat;    // current lattice of the slab
    size_t pf;
    uint4* halo_send_s;  // the south neighbour's packet area for what comes from the north (peer memory)
    uint4* halo_send_n;
    unsigned seed_flag;
    int nx, rows, pitch;
    int unshifted;       // 1: a packet holds the populations of its own cell (step_band_kernel shifts at the consumer), 0: already
                         // shifted for the cell that pulls them (step_ll_kernel)
};
__global__ void ll_seed_kernel(const LLSeedArgs a)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= a.nx) return;
    const int xw = a.unshifted ? x : ((x == 0) ? a.nx - 1 : x - 1), xe = a.unshifted ? x : ((x == a.nx - 1) ? 0 : x + 1);
    const size_t pitch = a.pitch;
    {   // row 0 -> the last row of the south neighbour pulls f4(x), f7(x+1), f8(x-1) from it
        const float* r = a.lat;
        st_packet<true>(a.halo_send_s + pitch + x, __ldcg(r + 4 * a.pf + x), __ldcg(r + 7 * a.pf + xe), __ldcg(r + 8 * a.pf + xw), a.seed_flag);
    }
    {   // last row -> the first row of the north neighbour pulls f2(x), f5(x-1), f6(x+1)
        const float* r = a.lat + static_cast<size_t>(a.rows - 1) * pitch;
        st_packet<true>(a.halo_send_n + pitch + x, __ldcg(r + 2 * a.pf + x), __ldcg(r + 5 * a.pf + xw), __ldcg(r + 6 * a.pf + xe), a.seed_flag);
    }
}

} // namespace lbm
