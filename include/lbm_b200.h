/*
 * lbm_b200.h -- C ABI of liblbm_b200.so: the D2Q9-BGK lattice-Boltzmann timestep of
 * Xinran1205/LBM-Asynchronous as hand-written sm_100a CUDA kernels.
 *
 * The reference has no FFI/plugin boundary (each variant is one C translation unit); the seam this
 * library occupies is the body of main()'s `for tt` loop plus the device residency of the lattices
 * (SURVEY.md 8b).  Each entry point names the reference interface it replaces.  Plain pointers and
 * sizes only; host-visible layouts are the reference's (AoS t_speed cells, int obstacles), the SoA
 * planes / obstacle bitmask / halo rings are internal.
 *
 * Error convention: every function returns 0 on success and a non-zero LBM_E* code on failure;
 * lbm_last_error() then returns a human-readable message for the calling thread, which the C host
 * program hands to its die() (reference convention: SerialCode/d2q9-bgk.c:745-751).  There is no
 * CPU fallback: without a usable CUDA device every compute entry point fails with LBM_ENODEVICE.
 */
#ifndef LBM_B200_H
#define LBM_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LBM_NSPEEDS 9

/* == t_param, SerialCode/d2q9-bgk.c:66-75 (same members, same order, same types) */
typedef struct {
    int   nx;           /* no. of cells in x-direction */
    int   ny;           /* no. of cells in y-direction */
    int   maxIters;     /* no. of iterations */
    int   reynolds_dim; /* dimension for Reynolds number */
    float density;      /* density per link */
    float accel;        /* density redistribution */
    float omega;        /* relaxation parameter */
} lbm_param_t;

/* == t_speed, SerialCode/d2q9-bgk.c:78-81 */
typedef struct {
    float speeds[LBM_NSPEEDS];
} lbm_speed_t;

/* opaque: the device-resident lattices (cells + tmp_cells of the reference), obstacle bitmask,
 * halo rings and per-step velocity sums of one run, on one or several GPUs */
typedef struct lbm_lattice lbm_lattice_t;

enum {
    LBM_OK = 0,
    LBM_EINVAL = 1,    /* bad argument */
    LBM_ENODEVICE = 2, /* no CUDA device / not enough devices / no peer access */
    LBM_ECUDA = 3,     /* a CUDA runtime call or kernel failed */
    LBM_ENOMEM = 4,
    LBM_ETIMEOUT = 5   /* a halo wait gave up (neighbour never delivered) */
};

/* arithmetic flavour of the collision */
enum {
    LBM_ARITH_STRICT = 0, /* operation order and IEEE rounding of SerialCode/d2q9-bgk.c:306-407: the
                             lattice is bit-identical to the reference's after every step */
    LBM_ARITH_FAST = 1    /* same formula, fused multiply-adds and reciprocal constants allowed:
                             differs from the reference by fp32 rounding only */
};

/* halo protocol between row slabs (only meaningful with more than one slab) */
enum {
    LBM_HALO_SYNC = 0,  /* boundary rows wait for the neighbour's row of the previous step
                           == MPI_Waitall, MPI_Waitall/d2q9-bgk.c:225-253 */
    LBM_HALO_ASYNC = 1  /* boundary rows never wait and use whatever the halo holds
                           == un-waited MPI_Testall, MPI_Testall_OptimizedVersion/d2q9-bgk.c:263-290 */
};

typedef struct {
    int arith;       /* LBM_ARITH_*            default LBM_ARITH_STRICT */
    int halo_mode;   /* LBM_HALO_*             default LBM_HALO_SYNC */
    int halo_lag;    /* LBM_HALO_SYNC only: boundary rows at step t use the neighbour row of step
                        t-halo_lag (even, >= 0; 0 = exact).  The deterministic stale-halo mode. */
    int use_graph;   /* 1: keep the step loop on the device (default): one cooperative launch per run for small
                        grids and slabs (cells in registers, rows exchanging flagged packets -- also between GPUs) and
                        for grids that live in L2, CUDA graphs of 32 steps otherwise; 0: one plain launch per step */
    int kernel;      /* kernel variant, 0 = library default; for tuning/bench only.  The codes are listed above
                        choose_kernel() in csrc/lbm_b200.cu and in DESIGN.md 3 (400/401/404 step_ll_kernel, 500/5BM
                        step_band_kernel, 3000/3CVM step_cluster_kernel, 200/201/204 step_loop_kernel, 1TTSM
                        step_tma_kernel, 2RRSNM step2_kernel, HM step_vec4_kernel, 99 step_scalar_kernel) */
    int block;       /* threads per CTA, 0 = default */
} lbm_options_t;

void lbm_default_options(lbm_options_t* opt);

const char* lbm_last_error(void);
/* number of usable CUDA devices (0 when there is none; never fails) */
int lbm_device_count(void);

/* Balanced row partition used by the multi-GPU path: slab r owns rows [starts[r], starts[r+1]).
 * starts[] has nslabs+1 entries.  Replaces the decomposition of MPI/d2q9-bgk.c:661-688; the driven
 * row ny-2 is always an interior row of the last slab (that slab owns >= 3 rows, every other >= 2).
 * Fails with LBM_EINVAL when ny is too small for nslabs. */
int lbm_partition(int ny, int nslabs, int* starts);

/* ---- one process driving `ngpus` devices (0..ngpus-1) ---------------------------------------
 * Replaces initialise()'s lattice set-up (SerialCode/d2q9-bgk.c:531-567; slabs, halos and
 * obstacle scatter of MPI/d2q9-bgk.c:661-829 when ngpus > 1).  `obstacles` is the reference's
 * int[ny*nx] map (non-zero = blocked).  The lattice starts in the uniform equilibrium state. */
int lbm_create(const lbm_param_t* params, const int* obstacles, int ngpus, const lbm_options_t* opt,
               lbm_lattice_t** out);

/* Same from a PACKED obstacle map: one bit per cell, row jj = lbm_packed_words_per_row(nx) 32-bit words, cell
 * (ii, jj) is bit ii % 32 of word ii / 32 of row jj (bits beyond nx are ignored).  This is the device's own
 * layout, so creation uploads 1/32 of the bytes of the int map and needs no packing pass.  Replaces the obstacle
 * read + scatter of MPI/d2q9-bgk.c:730-829 (rank 0 reads the file into an int map and sends int slabs). */
int lbm_create_packed(const lbm_param_t* params, const unsigned* obstacle_bits, int ngpus, const lbm_options_t* opt,
                      lbm_lattice_t** out);
size_t lbm_packed_words_per_row(int nx);

/* Same, with an explicit device per slab: devices[i] is the CUDA device of slab i.  A device may be
 * named more than once (tests on a box with fewer GPUs than slabs): slabs that share a device run
 * one after the other on one stream, so the halo protocol is exercised without kernels that would
 * have to be co-resident. */
int lbm_create_on(const lbm_param_t* params, const int* obstacles, int nslabs, const int* devices,
                  const lbm_options_t* opt, lbm_lattice_t** out);

/* ---- one process per GPU (torchrun / MPI-style launch) -------------------------------------
 * This process owns slab `rank` of `nranks`: global rows [row0, row1) (take them from
 * lbm_partition) on CUDA device `device`.  `obstacle_rows` holds only those rows,
 * int[(row1-row0)*nx].  After creating, every rank exports its halo handle, the handles are
 * exchanged by whatever transport the host has (torch.distributed, MPI, files ...), and each rank
 * connects to its two ring neighbours (rank-1 and rank+1, periodic).  Replaces MPI_Init/rank
 * set-up and the Isend/Irecv pairing of MPI_Waitall/d2q9-bgk.c:225-230. */
#define LBM_HALO_HANDLE_BYTES 128
int lbm_create_slab(const lbm_param_t* params, const int* obstacle_rows, int row0, int row1, int rank,
                    int nranks, int device, const lbm_options_t* opt, lbm_lattice_t** out);
/* lbm_create_slab from this rank's rows of the packed map (see lbm_create_packed) */
int lbm_create_slab_packed(const lbm_param_t* params, const unsigned* obstacle_bit_rows, int row0, int row1, int rank,
                           int nranks, int device, const lbm_options_t* opt, lbm_lattice_t** out);
int lbm_halo_export(lbm_lattice_t* lat, void* handle /* LBM_HALO_HANDLE_BYTES */);
int lbm_halo_connect(lbm_lattice_t* lat, const void* south_handle /* rank-1 */, const void* north_handle /* rank+1 */);

/* run the library's kernels on a caller-owned stream (slab 0 only; e.g. torch's current stream so
 * that torch.cuda.Event brackets them).  NULL restores the library's own stream. */
int lbm_set_stream(lbm_lattice_t* lat, void* cuda_stream);

/* ---- the hot path ---------------------------------------------------------------------------
 * lbm_run == `iters` iterations of  { timestep(); av_vels[tt] = av_velocity(); }
 * (SerialCode/d2q9-bgk.c:166-169; fusion_more() of OpenMP/d2q9-bgk.c:216-230).  Asynchronous:
 * returns once the work is queued.  May be called repeatedly; step indices continue. */
int lbm_run(lbm_lattice_t* lat, int iters);
/* block until everything queued so far has finished; reports kernel faults / halo time-outs */
int lbm_sync(lbm_lattice_t* lat);
/* av_vels of the last lbm_run call: av_vels[tt] = tot_u / (float)tot_cells, tt in [0, iters).
 * Multi-GPU in one process: already combined over slabs (replaces MPI_Reduce + divide,
 * MPI/d2q9-bgk.c:298-309).  Synchronises. */
int lbm_av_vels(lbm_lattice_t* lat, float* av_vels, int iters);
/* one-process-per-GPU: this lattice's exact integer sums of |u| over its fluid cells for each step of
 * the last lbm_run call, for the host to add over ranks (replaces MPI_Reduce(MPI_SUM),
 * MPI/d2q9-bgk.c:298-309).  Every cell's fp32 |u| is counted in units of 2^-40; sums[2*tt] is the
 * sum of the low 24 bits of those counts, sums[2*tt+1] the sum of the rest (total = lo + hi * 2^24),
 * nonfinite[tt] the number of fluid cells whose |u| was NaN or >= 4.  Integer addition is
 * associative: the totals do not depend on the decomposition, the kernel variant or scheduling.
 * Synchronises. */
int lbm_tot_u_sums(lbm_lattice_t* lat, long long* sums /* 2*iters */, long long* nonfinite /* iters or NULL */, int iters);
/* av_vels[tt] from the (rank-added) sums: tot_u = (float)((lo + hi*2^24) * 2^-40), returned value
 * tot_u / (float)fluid_cells as SerialCode/d2q9-bgk.c:457; NaN when nonfinite != 0 */
float lbm_av_from_sums(long long lo, long long hi, long long nonfinite, long long fluid_cells);
long long lbm_fluid_cells(const lbm_lattice_t* lat); /* non-obstacle cells of this lattice/slab */
long long lbm_steps_done(const lbm_lattice_t* lat);

/* av_velocity() of the current state (SerialCode/d2q9-bgk.c:409-458), as calc_reynolds needs it
 * (:637-642).  Single-process lattices only.  Synchronises. */
int lbm_av_velocity(lbm_lattice_t* lat, float* av);
/* total_density(), SerialCode/d2q9-bgk.c:644-660 (summed in double).  Synchronises. */
int lbm_total_density(lbm_lattice_t* lat, double* total);

/* per-cell values write_values() prints (SerialCode/d2q9-bgk.c:679-724): float[rows*nx] each, row
 * major, this lattice's rows only (all ny rows for lbm_create lattices).  Obstacle cells:
 * u_x = u_y = u = 0, pressure = density * c_sq.  Any pointer may be NULL.  Synchronises. */
int lbm_final_state(lbm_lattice_t* lat, float* u_x, float* u_y, float* u, float* pressure);

/* the lattice in the reference's AoS layout, this lattice's rows only (tests, checkpoints).
 * Replaces the slab gather of MPI/d2q9-bgk.c:265-295.  Synchronises. */
int lbm_download_cells(lbm_lattice_t* lat, lbm_speed_t* cells);
int lbm_upload_cells(lbm_lattice_t* lat, const lbm_speed_t* cells);

/* milliseconds the device spent in the last lbm_run call (CUDA events on slab 0's stream) */
int lbm_last_run_ms(lbm_lattice_t* lat, float* ms);
/* kernels the library launched so far (graph nodes counted individually) */
long long lbm_kernel_launches(const lbm_lattice_t* lat);
/* Device self-test of the arithmetic building blocks of the strict flavour: the kernel's division and
 * square-root sequences are compared with div.rn.f32 / sqrt.rn.f32 on about `pairs` pseudo-random
 * operand sets drawn from the operand window they are specified for; mismatches[0] counts
 * differing quotients, mismatches[1] differing roots.  Both must be 0. */
int lbm_selftest(int device, unsigned long long pairs, unsigned long long seed, unsigned long long* mismatches /* [2] */);
/* Device self-test of the four-cell collision the step kernels run (packed fp32 instructions, one basic block,
 * lbm_collide4.cuh) against the scalar per-cell code with its guarded IEEE paths, on about `sets` pseudo-random
 * groups of four cells (lattice-like, rough and arbitrary populations, random obstacle bits): mismatches[0] counts
 * differing population words, mismatches[1] differing |u| words.  Both must be 0 for arith = LBM_ARITH_STRICT (the
 * packed code is the reference's operation sequence, SerialCode/d2q9-bgk.c:325-401,425-452, two cells per
 * instruction); the fast flavour may differ from its scalar form by contraction only. */
int lbm_selftest_collide(int device, int arith, unsigned long long sets, unsigned long long seed, unsigned long long* mismatches /* [2] */);
/* rows [row0,row1) and device of slab `i` (i < lbm_num_slabs) */
int lbm_num_slabs(const lbm_lattice_t* lat);
int lbm_slab_info(const lbm_lattice_t* lat, int i, int* row0, int* row1, int* device);

/* finalise(), SerialCode/d2q9-bgk.c:615-634 */
void lbm_destroy(lbm_lattice_t* lat);

#ifdef __cplusplus
}
#endif
#endif /* LBM_B200_H */
