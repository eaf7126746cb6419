/*
 * lbm_oracle.c -- CPU restatement of the reference's D2Q9-BGK timestep (see lbm_oracle.h).
 *
 * TEST INFRASTRUCTURE ONLY; parity pinned against the reference's SerialCode binary and its
 * shipped golden files (tests/test_oracle.py).  Every function cites the reference lines it
 * restates.  All arithmetic is fp32 with the reference's operation order; this file is compiled
 * with -ffp-contract=off (the reference is built with gcc -std=c99, which implies the same), so
 * no multiply-add is fused and the results are bit-identical to the reference binary's.
 *
 * Direction numbering (SerialCode/d2q9-bgk.c:9-15): 0 rest, 1 E, 2 N, 3 W, 4 S, 5 NE, 6 NW,
 * 7 SW, 8 SE.
 */
#include "lbm_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define Q ORACLE_NSPEEDS
#define AT(base, nx, ii, jj) ((base) + ((size_t)(ii) + (size_t)(jj) * (size_t)(nx)) * Q)

/* ------------------------------------------------------------------------------------------
 * per-cell pieces
 * ---------------------------------------------------------------------------------------- */

/* rho, u_x, u_y of one cell: SerialCode/d2q9-bgk.c:325-349 (and again :425-449, :693-717).
 * Density is a left-to-right sum starting from 0.f; each velocity bracket is left-to-right. */
static inline void moments(const float* f, float* rho, float* ux, float* uy)
{
    float d = 0.f;
    for (int k = 0; k < Q; k++) d += f[k];
    *rho = d;
    *ux = (f[1] + f[5] + f[8] - (f[3] + f[6] + f[7])) / d;
    *uy = (f[2] + f[5] + f[6] - (f[4] + f[7] + f[8])) / d;
}

/* BGK relaxation of one fluid cell: SerialCode/d2q9-bgk.c:306-401.  t = streamed-in values,
 * out = post-collision values. */
static inline void collide_cell(const float* t, float omega, float* out)
{
    const float c_sq = 1.f / 3.f; /* :308 */
    const float w0 = 4.f / 9.f;   /* :309 */
    const float w1 = 1.f / 9.f;   /* :310 */
    const float w2 = 1.f / 36.f;  /* :311 */

    float rho, ux, uy;
    moments(t, &rho, &ux, &uy);
    const float u_sq = ux * ux + uy * uy; /* :352 */

    float u[Q]; /* :355-363 */
    u[0] = 0.f;
    u[1] = ux;
    u[2] = uy;
    u[3] = -ux;
    u[4] = -uy;
    u[5] = ux + uy;
    u[6] = -ux + uy;
    u[7] = -ux - uy;
    u[8] = ux - uy;

    float d_equ[Q];
    d_equ[0] = w0 * rho * (1.f - u_sq / (2.f * c_sq)); /* :367-368 */
    for (int k = 1; k < Q; k++) {                       /* :370-393, same expression 8 times */
        const float w = (k <= 4) ? w1 : w2;
        d_equ[k] = w * rho * (1.f + u[k] / c_sq + (u[k] * u[k]) / (2.f * c_sq * c_sq) - u_sq / (2.f * c_sq));
    }
    for (int k = 0; k < Q; k++) /* :396-401 */
        out[k] = t[k] + omega * (d_equ[k] - t[k]);
}

/* bounce-back permutation of one obstacle cell: SerialCode/d2q9-bgk.c:287-299
 * (1<->3, 2<->4, 5<->7, 6<->8; speed 0 is not touched). */
static inline void rebound_cell(const float* t, float* out)
{
    out[1] = t[3];
    out[2] = t[4];
    out[3] = t[1];
    out[4] = t[2];
    out[5] = t[7];
    out[6] = t[8];
    out[7] = t[5];
    out[8] = t[6];
}

/* accelerate one cell of the driven row: SerialCode/d2q9-bgk.c:229-241 */
static inline void accelerate_cell(float* f, float w1, float w2)
{
    if ((f[3] - w1) > 0.f && (f[6] - w2) > 0.f && (f[7] - w2) > 0.f) {
        f[1] += w1;
        f[5] += w2;
        f[8] += w2;
        f[3] -= w1;
        f[6] -= w2;
        f[7] -= w2;
    }
}

/* pull the nine populations that stream into cell (ii,jj) of a grid with periodic wrap in both
 * directions: SerialCode/d2q9-bgk.c:257-272 */
static inline void pull_periodic(const float* cells, int nx, int ny, int ii, int jj, float* t)
{
    const int y_n = (jj + 1) % ny;
    const int x_e = (ii + 1) % nx;
    const int y_s = (jj == 0) ? (ny - 1) : (jj - 1);
    const int x_w = (ii == 0) ? (nx - 1) : (ii - 1);
    t[0] = AT(cells, nx, ii, jj)[0];
    t[1] = AT(cells, nx, x_w, jj)[1];
    t[2] = AT(cells, nx, ii, y_s)[2];
    t[3] = AT(cells, nx, x_e, jj)[3];
    t[4] = AT(cells, nx, ii, y_n)[4];
    t[5] = AT(cells, nx, x_w, y_s)[5];
    t[6] = AT(cells, nx, x_e, y_s)[6];
    t[7] = AT(cells, nx, x_e, y_n)[7];
    t[8] = AT(cells, nx, x_w, y_n)[8];
}

/* ------------------------------------------------------------------------------------------
 * the serial program's passes
 * ---------------------------------------------------------------------------------------- */

void oracle_init_cells(const oracle_param* p, float* cells)
{
    /* SerialCode/d2q9-bgk.c:546-567 */
    const float w0 = p->density * 4.f / 9.f;
    const float w1 = p->density / 9.f;
    const float w2 = p->density / 36.f;
    const size_t n = (size_t)p->nx * (size_t)p->ny;
    for (size_t c = 0; c < n; c++) {
        float* f = cells + c * Q;
        f[0] = w0;
        f[1] = f[2] = f[3] = f[4] = w1;
        f[5] = f[6] = f[7] = f[8] = w2;
    }
}

void oracle_accelerate_flow(const oracle_param* p, float* cells, const int* obstacles)
{
    /* SerialCode/d2q9-bgk.c:216-246: second row from the top only */
    const float w1 = p->density * p->accel / 9.f;
    const float w2 = p->density * p->accel / 36.f;
    const int jj = p->ny - 2;
    for (int ii = 0; ii < p->nx; ii++)
        if (!obstacles[ii + jj * p->nx]) accelerate_cell(AT(cells, p->nx, ii, jj), w1, w2);
}

void oracle_propagate(const oracle_param* p, const float* cells, float* tmp_cells)
{
    /* SerialCode/d2q9-bgk.c:248-277: every cell, obstacles included */
    for (int jj = 0; jj < p->ny; jj++)
        for (int ii = 0; ii < p->nx; ii++) pull_periodic(cells, p->nx, p->ny, ii, jj, AT(tmp_cells, p->nx, ii, jj));
}

void oracle_rebound(const oracle_param* p, float* cells, const float* tmp_cells, const int* obstacles)
{
    /* SerialCode/d2q9-bgk.c:279-304 */
    for (int jj = 0; jj < p->ny; jj++)
        for (int ii = 0; ii < p->nx; ii++)
            if (obstacles[ii + jj * p->nx]) rebound_cell(AT(tmp_cells, p->nx, ii, jj), AT(cells, p->nx, ii, jj));
}

void oracle_collision(const oracle_param* p, float* cells, const float* tmp_cells, const int* obstacles)
{
    /* SerialCode/d2q9-bgk.c:306-407 */
    for (int jj = 0; jj < p->ny; jj++)
        for (int ii = 0; ii < p->nx; ii++)
            if (!obstacles[ii + jj * p->nx])
                collide_cell(AT(tmp_cells, p->nx, ii, jj), p->omega, AT(cells, p->nx, ii, jj));
}

void oracle_timestep(const oracle_param* p, float* cells, float* tmp_cells, const int* obstacles)
{
    /* SerialCode/d2q9-bgk.c:207-214 */
    oracle_accelerate_flow(p, cells, obstacles);
    oracle_propagate(p, cells, tmp_cells);
    oracle_rebound(p, cells, tmp_cells, obstacles);
    oracle_collision(p, cells, tmp_cells, obstacles);
}

float oracle_av_velocity(const oracle_param* p, const float* cells, const int* obstacles)
{
    /* SerialCode/d2q9-bgk.c:409-458 */
    int tot_cells = 0;
    float tot_u = 0.f;
    for (int jj = 0; jj < p->ny; jj++)
        for (int ii = 0; ii < p->nx; ii++)
            if (!obstacles[ii + jj * p->nx]) {
                float rho, ux, uy;
                moments(AT(cells, p->nx, ii, jj), &rho, &ux, &uy);
                tot_u += sqrtf((ux * ux) + (uy * uy));
                ++tot_cells;
            }
    return tot_u / (float)tot_cells;
}

double oracle_tot_u_f64(const oracle_param* p, const float* cells, const int* obstacles, int* tot_cells)
{
    /* same per-cell fp32 value as av_velocity (:425-452), accumulated in double */
    int n = 0;
    double tot = 0.0;
    for (int jj = 0; jj < p->ny; jj++)
        for (int ii = 0; ii < p->nx; ii++)
            if (!obstacles[ii + jj * p->nx]) {
                float rho, ux, uy;
                moments(AT(cells, p->nx, ii, jj), &rho, &ux, &uy);
                tot += (double)sqrtf((ux * ux) + (uy * uy));
                ++n;
            }
    if (tot_cells) *tot_cells = n;
    return tot;
}

float oracle_total_density(const oracle_param* p, const float* cells)
{
    /* SerialCode/d2q9-bgk.c:644-660 */
    float total = 0.f;
    const size_t n = (size_t)p->nx * (size_t)p->ny * Q;
    for (size_t i = 0; i < n; i++) total += cells[i];
    return total;
}

float oracle_calc_reynolds(const oracle_param* p, const float* cells, const int* obstacles)
{
    /* SerialCode/d2q9-bgk.c:637-642 */
    const float viscosity = 1.f / 6.f * (2.f / p->omega - 1.f);
    return oracle_av_velocity(p, cells, obstacles) * p->reynolds_dim / viscosity;
}

void oracle_run(const oracle_param* p, float* cells, float* tmp_cells, const int* obstacles, int iters,
                float* av_vels)
{
    /* SerialCode/d2q9-bgk.c:166-169 */
    for (int tt = 0; tt < iters; tt++) {
        oracle_timestep(p, cells, tmp_cells, obstacles);
        av_vels[tt] = oracle_av_velocity(p, cells, obstacles);
    }
}

void oracle_final_state(const oracle_param* p, const float* cells, const int* obstacles, float* u_x, float* u_y,
                        float* u, float* pressure)
{
    /* SerialCode/d2q9-bgk.c:679-724 */
    const float c_sq = 1.f / 3.f;
    for (int jj = 0; jj < p->ny; jj++)
        for (int ii = 0; ii < p->nx; ii++) {
            const size_t c = (size_t)ii + (size_t)jj * p->nx;
            if (obstacles[c]) {
                u_x[c] = u_y[c] = u[c] = 0.f;
                pressure[c] = p->density * c_sq;
            } else {
                float rho, ux, uy;
                moments(cells + c * Q, &rho, &ux, &uy);
                u_x[c] = ux;
                u_y[c] = uy;
                u[c] = sqrtf((ux * ux) + (uy * uy));
                pressure[c] = rho * c_sq;
            }
        }
}

/* ------------------------------------------------------------------------------------------
 * fused single pass (the OpenMP program)
 * ---------------------------------------------------------------------------------------- */

float oracle_fused_step(const oracle_param* p, float* in, float* out, const int* obstacles)
{
    /* OpenMP/d2q9-bgk.c:304-321: accelerate pre-pass on the source lattice */
    oracle_accelerate_flow(p, in, obstacles);

    /* OpenMP/d2q9-bgk.c:334-495: one sweep; the reference reduces per-thread partial sums, so
     * does this (row-blocked static schedule). */
    const int nx = p->nx, ny = p->ny;
    float tot_u = 0.f;
    int tot_cells = 0;
#pragma omp parallel for reduction(+ : tot_u, tot_cells) schedule(static)
    for (int jj = 0; jj < ny; jj++) {
        for (int ii = 0; ii < nx; ii++) {
            float t[Q];
            pull_periodic(in, nx, ny, ii, jj, t);
            float* o = AT(out, nx, ii, jj);
            if (!obstacles[ii + jj * nx]) {
                collide_cell(t, p->omega, o);
                float rho, ux, uy;
                moments(o, &rho, &ux, &uy); /* :450-475: from the stored values */
                tot_u += sqrtf((ux * ux) + (uy * uy));
                ++tot_cells;
            } else {
                o[0] = t[0]; /* OpenMP/d2q9-bgk.c:484 */
                rebound_cell(t, o);
            }
        }
    }
    return tot_u / (float)tot_cells;
}

/* ------------------------------------------------------------------------------------------
 * row-decomposed runs (the MPI programs), emulated rank by rank in one process
 * ---------------------------------------------------------------------------------------- */

int oracle_reference_partition(int ny, int nranks, int* starts)
{
    /* MPI/d2q9-bgk.c:661-688 (identical in the other MPI variants) */
    if (nranks < 1) return -1;
    const int basic = (ny - 3) / nranks;
    const int rem = (ny - 3) % nranks;
    int row = 0;
    for (int r = 0; r < nranks; r++) {
        int rows = basic + (r < rem ? 1 : 0);
        if (r == nranks - 1) rows += 3;
        if (rows < 1) return -1;
        starts[r] = row;
        row += rows;
    }
    starts[nranks] = row; /* == ny */
    return 0;
}

typedef struct {
    int rows;     /* owned rows; the slab has rows+2 rows, row 0 and rows+1 are halos */
    int g0;       /* first owned global row */
    float* cur;   /* (rows+2)*nx*Q */
    float* nxt;
    float* hist_first; /* ring of the row-1 copies "sent" at steps t, t-1, ... : (lag+1)*nx*Q */
    float* hist_last;  /* same for row `rows` */
} slab_t;

/* One call of the MPI variants' fusion_more() over slab rows [r0,r1]:
 * MPI_Waitall/d2q9-bgk.c:352-555 == MPI_Testall_OptimizedVersion/d2q9-bgk.c:407-610.
 * No y wrap is needed (rows 1..rows never leave the slab: y_s=jj-1, y_n=jj+1, :460-463).
 * Returns the fp32 sum of |u| over the fluid cells of those rows. */
static float fused_rows(const oracle_param* p, const slab_t* s, const int* obstacles, int r0, int r1)
{
    const int nx = p->nx;
    float tot_u = 0.f;
    for (int jj = r0; jj <= r1; jj++) {
        const int gj = s->g0 + jj - 1; /* obstacle slab has no halo rows, :479 */
        for (int ii = 0; ii < nx; ii++) {
            const int x_e = (ii + 1) % nx;
            const int x_w = (ii == 0) ? (nx - 1) : (ii - 1);
            float t[Q];
            t[0] = AT(s->cur, nx, ii, jj)[0];
            t[1] = AT(s->cur, nx, x_w, jj)[1];
            t[2] = AT(s->cur, nx, ii, jj - 1)[2];
            t[3] = AT(s->cur, nx, x_e, jj)[3];
            t[4] = AT(s->cur, nx, ii, jj + 1)[4];
            t[5] = AT(s->cur, nx, x_w, jj - 1)[5];
            t[6] = AT(s->cur, nx, x_e, jj - 1)[6];
            t[7] = AT(s->cur, nx, x_e, jj + 1)[7];
            t[8] = AT(s->cur, nx, x_w, jj + 1)[8];
            float* o = AT(s->nxt, nx, ii, jj);
            if (!obstacles[ii + gj * nx]) {
                collide_cell(t, p->omega, o);
                float rho, ux, uy;
                moments(o, &rho, &ux, &uy);
                tot_u += sqrtf((ux * ux) + (uy * uy));
            } else {
                rebound_cell(t, o); /* speed 0 left alone, Optimized:596-606 */
            }
        }
    }
    return tot_u;
}

int oracle_run_decomposed(const oracle_param* p, const int* obstacles, int nranks, const int* starts, int halo_lag,
                          int iters, float* cells_out, float* av_vels)
{
    const int nx = p->nx, ny = p->ny;
    if (nranks < 1 || halo_lag < 0 || (halo_lag & 1)) return -1;
    if (starts[0] != 0 || starts[nranks] != ny) return -1;
    const size_t rowsz = (size_t)nx * Q;
    const int ring = halo_lag + 1;

    slab_t* sl = (slab_t*)calloc((size_t)nranks, sizeof(slab_t));
    if (!sl) return -2;
    int bad = 0;
    for (int r = 0; r < nranks; r++) {
        sl[r].g0 = starts[r];
        sl[r].rows = starts[r + 1] - starts[r];
        if (sl[r].rows < 1) bad = 1;
    }
    /* the driven row ny-2 must be an interior row of the last slab (rows >= 3 there), which is
     * what the reference's "+3" guarantees (Optimized:749-759) */
    if (sl[nranks - 1].rows < 3) bad = 1;
    if (bad) {
        free(sl);
        return -1;
    }

    oracle_param slab_p = *p;
    for (int r = 0; r < nranks; r++) {
        const size_t n = (size_t)(sl[r].rows + 2) * rowsz;
        sl[r].cur = (float*)malloc(n * sizeof(float));
        sl[r].nxt = (float*)malloc(n * sizeof(float));
        sl[r].hist_first = (float*)malloc((size_t)ring * rowsz * sizeof(float));
        sl[r].hist_last = (float*)malloc((size_t)ring * rowsz * sizeof(float));
        /* both lattices, halo rows included, start uniform: Optimized:784-824 */
        slab_p.ny = sl[r].rows + 2;
        oracle_init_cells(&slab_p, sl[r].cur);
        oracle_init_cells(&slab_p, sl[r].nxt);
    }

    int fluid = 0; /* numberOfNonObstacles, Optimized:870-880 */
    for (size_t c = 0; c < (size_t)nx * ny; c++) fluid += (obstacles[c] != 1);

    const float w1 = p->density * p->accel / 9.f;
    const float w2 = p->density * p->accel / 36.f;

    for (int tt = 0; tt < iters; tt++) {
        /* (A) every rank "sends" its first and last owned row of the current lattice
         * (Optimized:263-264); keep them in a ring so a lagged receiver can pick an older one */
        const int slot = tt % ring;
        for (int r = 0; r < nranks; r++) {
            memcpy(sl[r].hist_first + (size_t)slot * rowsz, sl[r].cur + rowsz, rowsz * sizeof(float));
            memcpy(sl[r].hist_last + (size_t)slot * rowsz, sl[r].cur + (size_t)sl[r].rows * rowsz,
                   rowsz * sizeof(float));
        }
        /* (A') what has "arrived" in the halo rows of the current lattice when the boundary rows are
         * computed: the message of step tt-halo_lag (Optimized:267-268,279-280; App. C).  With
         * halo_lag == 0 this is MPI_Waitall (MPI_Waitall:243). */
        if (tt - halo_lag >= 0) {
            const int from = (tt - halo_lag) % ring;
            for (int r = 0; r < nranks; r++) {
                const int up = (r - 1 + nranks) % nranks; /* Optimized:253-254 */
                const int down = (r + 1) % nranks;
                memcpy(sl[r].cur, sl[up].hist_last + (size_t)from * rowsz, rowsz * sizeof(float));
                memcpy(sl[r].cur + (size_t)(sl[r].rows + 1) * rowsz, sl[down].hist_first + (size_t)from * rowsz,
                       rowsz * sizeof(float));
            }
        }
        float step_sum = 0.f;
        for (int r = 0; r < nranks; r++) {
            slab_t* s = &sl[r];
            /* (B) accelerate (last rank only, local row ny-3 == global ny-2), Optimized:417-441 */
            if (r == nranks - 1) {
                const int lj = s->rows - 1; /* slab row of global ny-2 */
                for (int ii = 0; ii < nx; ii++)
                    if (!obstacles[ii + (ny - 2) * nx]) accelerate_cell(AT(s->cur, nx, ii, lj), w1, w2);
            }
            /* interior rows 2..rows-1, then the two boundary rows, Optimized:271-290.  A one-row
             * slab computes its row once (the reference would do it twice and double count,
             * SURVEY.md App. D -- not replicated). */
            float tot_in = 0.f, tot_bd = 0.f;
            if (s->rows >= 3) tot_in = fused_rows(p, s, obstacles, 2, s->rows - 1);
            tot_bd = fused_rows(p, s, obstacles, 1, 1);
            if (s->rows >= 2) tot_bd += fused_rows(p, s, obstacles, s->rows, s->rows);
            step_sum += tot_in + tot_bd; /* Optimized:293 then MPI_Reduce(SUM), :368 */
        }
        av_vels[tt] = step_sum / (float)fluid; /* Optimized:370-374 */
        for (int r = 0; r < nranks; r++) { /* swap, Optimized:304-306 */
            float* t = sl[r].cur;
            sl[r].cur = sl[r].nxt;
            sl[r].nxt = t;
        }
    }

    /* gather in rank order, Optimized:331-361 */
    for (int r = 0; r < nranks; r++)
        memcpy(cells_out + (size_t)sl[r].g0 * rowsz, sl[r].cur + rowsz, (size_t)sl[r].rows * rowsz * sizeof(float));

    for (int r = 0; r < nranks; r++) {
        free(sl[r].cur);
        free(sl[r].nxt);
        free(sl[r].hist_first);
        free(sl[r].hist_last);
    }
    free(sl);
    return 0;
}
