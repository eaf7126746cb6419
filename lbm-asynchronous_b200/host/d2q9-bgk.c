/*
 * d2q9-bgk -- host program of the B200 D2Q9-BGK lattice-Boltzmann solver.
 *
 * Same command line, input formats, output files and stdout lines as the reference programs
 * (every variant of Xinran1205/LBM-Asynchronous shares them; cited below against
 * SerialCode/d2q9-bgk.c), so the reference's check/check.py validates the results unchanged:
 *
 *     d2q9-bgk <paramfile> <obstaclefile>      ->  final_state.dat, av_vels.dat in the cwd
 *
 * All numerical work happens in liblbm_b200.so (include/lbm_b200.h) on the GPU(s); this file is
 * plain C99 host glue: parse, call, time, print.  There is no CPU path: without a CUDA device the
 * program stops with an error.
 *
 * Behaviour beyond the reference is selected by environment variables, so the command line stays
 * the reference's:
 *   LBM_GPUS=N               row slabs on N GPUs of this box (default 1)
 *   LBM_HALO_MODE=sync|async halo protocol between slabs (sync == MPI_Waitall variant, async ==
 *                            MPI_Testall variant: boundary rows never wait), default sync
 *   LBM_HALO_LAG=k           sync only: deterministic staleness of k (even) steps
 *   LBM_ARITH=strict|fast    collision arithmetic (default strict: bit-identical to SerialCode)
 *   LBM_SKIP_FINAL_STATE=1   do not write final_state.dat (87 bytes per cell of text)
 *   LBM_KERNEL, LBM_BLOCK    kernel variant / CTA size (tuning)
 *   LBM_PARSE_ONLY=1         read both input files, print the number of blocked cells, an FNV-1a hash of the packed
 *                            obstacle map and the time the parse took, and exit (no GPU needed; for tooling/tests)
 *   LBM_ANIMATION_EVERY=N    write animation_data/velocity_magnitude_%06d.dat after every N-th timestep
 *                            (tt = 0, N, 2N, ...), the frames Visualization/animation.py reads.  Off by
 *                            default: the reference ships with these calls commented out
 *                            (SerialCode/d2q9-bgk.c:171-173, write_animation_data :802-849)
 */
#include <fcntl.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <sys/time.h>
#include <unistd.h>

#include "lbm_b200.h"

#define FINAL_STATE_FILE "final_state.dat" /* SerialCode/d2q9-bgk.c:62 */
#define AV_VELS_FILE "av_vels.dat"         /* :63 */

/* error convention of the reference: message on stderr, exit status 1 (SerialCode:745-751) */
static void die(const char* message, const int line, const char* file)
{
    fprintf(stderr, "Error at line %d of file %s:\n", line, file);
    fprintf(stderr, "%s\n", message);
    fflush(stderr);
    exit(EXIT_FAILURE);
}
#define DIE(msg) die((msg), __LINE__, __FILE__)

static void usage(const char* exe) /* SerialCode:753-757 */
{
    fprintf(stderr, "Usage: %s <paramfile> <obstaclefile>\n", exe);
    exit(EXIT_FAILURE);
}

static void die_lbm(const char* what, int line)
{
    char msg[768];
    snprintf(msg, sizeof msg, "%s: %s", what, lbm_last_error());
    die(msg, line, __FILE__);
}
#define LBM_CALL(call)                          \
    do {                                        \
        if ((call) != LBM_OK) die_lbm(#call, __LINE__); \
    } while (0)

static double wall_seconds(void)
{
    struct timeval t;
    gettimeofday(&t, NULL);
    return t.tv_sec + t.tv_usec / 1000000.0;
}

/* the seven values of the parameter file, in file order (SerialCode:480-506) */
static void read_params(const char* path, lbm_param_t* p)
{
    FILE* fp = fopen(path, "r");
    if (!fp) {
        char msg[1024];
        snprintf(msg, sizeof msg, "could not open input parameter file: %s", path);
        DIE(msg);
    }
    struct {
        const char* fmt;
        void* dst;
        const char* err;
    } fields[7] = {
        {"%d\n", &p->nx, "could not read param file: nx"},
        {"%d\n", &p->ny, "could not read param file: ny"},
        {"%d\n", &p->maxIters, "could not read param file: maxIters"},
        {"%d\n", &p->reynolds_dim, "could not read param file: reynolds_dim"},
        {"%f\n", &p->density, "could not read param file: density"},
        {"%f\n", &p->accel, "could not read param file: accel"},
        {"%f\n", &p->omega, "could not read param file: omega"},
    };
    for (int i = 0; i < 7; i++)
        if (fscanf(fp, fields[i].fmt, fields[i].dst) != 1) DIE(fields[i].err);
    fclose(fp);
}

/* ---- obstacle file -> packed bit map --------------------------------------------------------------------
 * The reference reads the file with  while ((retval = fscanf(fp, "%d %d %d\n", &xx, &yy, &blocked)) != EOF)
 * (SerialCode/d2q9-bgk.c:588-601): a stream of whitespace-separated integers taken three at a time, four checks
 * per triple, the first failing triple (in file order) ends the program.  One fscanf per line on one core takes
 * ~1 s per 3 M lines; the 32768 x 32768 synthetic channel has 5.4 M.  Same semantics here, but the file is
 * mapped and tokenised by all host cores: pass 1 counts the integer tokens of each chunk (and finds the first
 * byte that is not part of an integer), a prefix sum tells each chunk where its first triple starts, pass 2
 * checks the triples and sets bits.  The map is 1 bit per cell (what the device keeps) instead of the reference's
 * int per cell. */
typedef struct {
    const char* base;
    size_t begin, end, size; /* this chunk: [begin, end) of [0, size) */
    int nx, ny;
    uint32_t* bits;
    size_t words_per_row;
    /* pass 1 */
    size_t ntokens;
    size_t garbage; /* offset of the first byte that starts no integer, or SIZE_MAX */
    /* pass 2 */
    size_t first_token; /* global index of this chunk's first token */
    size_t total_tokens; /* tokens before the first garbage byte of the whole file */
    size_t err_triple;  /* first failing triple, or SIZE_MAX */
    int err_kind;       /* 1 x range, 2 y range, 3 blocked value */
} obst_job;

static int is_space(char c) { return c == ' ' || c == '\n' || c == '\t' || c == '\r' || c == '\v' || c == '\f'; }

/* one "%d" conversion at *pos: skips white space, then [+-]digits.  1: value stored; 0: not an integer
 * (*pos = the offending byte); -1: only white space left before `limit` */
static int scan_int(const char* s, size_t* pos, size_t limit, long long* value)
{
    size_t p = *pos;
    while (p < limit && is_space(s[p])) p++;
    if (p >= limit) {
        *pos = p;
        return -1;
    }
    size_t q = p;
    int neg = 0;
    if (s[q] == '+' || s[q] == '-') neg = s[q] == '-', q++;
    if (q >= limit || s[q] < '0' || s[q] > '9') {
        *pos = p;
        return 0;
    }
    long long v = 0;
    while (q < limit && s[q] >= '0' && s[q] <= '9') {
        if (v < (1LL << 40)) v = v * 10 + (s[q] - '0');
        q++;
    }
    *value = neg ? -v : v;
    *pos = q;
    return 1;
}

static void* obst_pass1(void* arg)
{
    obst_job* j = (obst_job*)arg;
    size_t p = j->begin, n = 0;
    j->garbage = SIZE_MAX;
    for (;;) {
        /* a token belongs to the chunk its first byte lies in; tokens never contain white space and chunks are
         * cut at white space, so scanning up to the end of the file never reads a token twice */
        size_t q = p;
        while (q < j->end && is_space(j->base[q])) q++;
        if (q >= j->end) break;
        long long v;
        p = q;
        const int r = scan_int(j->base, &p, j->size, &v);
        if (r != 1) {
            j->garbage = q;
            break;
        }
        n++;
    }
    j->ntokens = n;
    return NULL;
}

static void* obst_pass2(void* arg)
{
    obst_job* j = (obst_job*)arg;
    j->err_triple = SIZE_MAX;
    j->err_kind = 0;
    size_t p = j->begin;
    size_t tok = j->first_token;
    /* skip the tail of a triple that started in an earlier chunk */
    long long v;
    while (tok % 3 != 0 && tok < j->first_token + j->ntokens) {
        scan_int(j->base, &p, j->size, &v);
        tok++;
    }
    while (tok < j->first_token + j->ntokens && tok + 3 <= j->total_tokens) {
        long long t[3];
        for (int i = 0; i < 3; i++) scan_int(j->base, &p, j->size, &t[i]); /* may run into the next chunk */
        const size_t triple = tok / 3;
        int kind = 0;
        if (t[0] < 0 || t[0] > j->nx - 1) kind = 1;
        else if (t[1] < 0 || t[1] > j->ny - 1) kind = 2;
        else if (t[2] != 1) kind = 3;
        if (kind) {
            j->err_triple = triple;
            j->err_kind = kind;
            return NULL;
        }
        __atomic_fetch_or(&j->bits[(size_t)t[1] * j->words_per_row + ((size_t)t[0] >> 5)], 1u << (t[0] & 31), __ATOMIC_RELAXED);
        tok += 3;
    }
    return NULL;
}

static uint32_t* read_obstacles(const char* path, const lbm_param_t* p)
{
    const size_t wpr = ((size_t)p->nx + 31) / 32;
    uint32_t* bits = calloc(wpr * (size_t)p->ny, sizeof(uint32_t));
    if (!bits) DIE("cannot allocate column memory for obstacles");
    int fd = open(path, O_RDONLY);
    if (fd < 0) {
        char msg[1024];
        snprintf(msg, sizeof msg, "could not open input obstacles file: %s", path);
        DIE(msg);
    }
    struct stat st;
    if (fstat(fd, &st) != 0) DIE("could not open input obstacles file");
    const size_t size = (size_t)st.st_size;
    if (size == 0) {
        close(fd);
        return bits;
    }
    const char* base = mmap(NULL, size, PROT_READ, MAP_PRIVATE, fd, 0);
    if (base == MAP_FAILED) DIE("could not map the obstacles file");
    long ncpu = sysconf(_SC_NPROCESSORS_ONLN);
    int nthreads = (int)(ncpu < 1 ? 1 : (ncpu > 64 ? 64 : ncpu));
    if (size < (size_t)nthreads * 65536) nthreads = (int)(size / 65536) + 1;
    obst_job* jobs = calloc((size_t)nthreads, sizeof *jobs);
    pthread_t* tids = calloc((size_t)nthreads, sizeof *tids);
    if (!jobs || !tids) DIE("cannot allocate memory for the input threads");
    size_t cut = 0;
    for (int i = 0; i < nthreads; i++) {
        size_t end = (i == nthreads - 1) ? size : size / (size_t)nthreads * (size_t)(i + 1);
        if (end < cut) end = cut;
        while (end < size && !is_space(base[end])) end++; /* never inside a token */
        obst_job j = {base, cut, end, size, p->nx, p->ny, bits, wpr, 0, SIZE_MAX, 0, 0, SIZE_MAX, 0};
        jobs[i] = j;
        cut = end;
    }
    for (int pass = 1; pass <= 2; pass++) {
        for (int i = 0; i < nthreads; i++)
            if (pthread_create(&tids[i], NULL, pass == 1 ? obst_pass1 : obst_pass2, &jobs[i]) != 0) {
                (pass == 1 ? obst_pass1 : obst_pass2)(&jobs[i]);
                tids[i] = 0;
            }
        for (int i = 0; i < nthreads; i++)
            if (tids[i]) pthread_join(tids[i], NULL);
        if (pass == 1) {
            /* tokens before the first byte that is not part of an integer: all the reference would ever convert */
            size_t total = 0;
            int stopped = 0;
            for (int i = 0; i < nthreads; i++) {
                jobs[i].first_token = total;
                if (stopped) jobs[i].ntokens = 0;
                total += jobs[i].ntokens;
                if (jobs[i].garbage != SIZE_MAX) stopped = 1;
            }
            for (int i = 0; i < nthreads; i++) jobs[i].total_tokens = total;
        }
    }
    size_t total = jobs[0].total_tokens;
    int stopped = 0;
    for (int i = 0; i < nthreads; i++) stopped |= jobs[i].garbage != SIZE_MAX;
    size_t err_triple = SIZE_MAX;
    int err_kind = 0;
    for (int i = 0; i < nthreads; i++)
        if (jobs[i].err_triple < err_triple) err_triple = jobs[i].err_triple, err_kind = jobs[i].err_kind;
    munmap((void*)base, size);
    close(fd);
    free(jobs);
    free(tids);
    /* the first failing triple in file order; an incomplete last triple (or a non-integer) comes after every
     * complete one */
    if (err_kind == 1) DIE("obstacle x-coord out of range");
    if (err_kind == 2) DIE("obstacle y-coord out of range");
    if (err_kind == 3) DIE("obstacle blocked value should be 1");
    if (stopped || total % 3 != 0) DIE("expected 3 values per line in obstacle file");
    return bits;
}

static int obstacle_at(const uint32_t* bits, size_t words_per_row, int ii, int jj)
{
    return (int)((bits[(size_t)jj * words_per_row + ((size_t)ii >> 5)] >> (ii & 31)) & 1u);
}

/* final_state.dat: `ii jj u_x u_y u pressure obstacle`, jj outer / ii inner (SerialCode:679-724);
 * av_vels.dat: `tt:\tvalue` (:735-738).
 * The reference formats 87 bytes per cell with one fprintf per cell on one core -- at 1024 x 1024 that
 * is 91 MB of text and takes far longer than the simulation.  Same bytes here, but the rows are
 * formatted by all host cores (snprintf into per-band buffers) and written in order. */
typedef struct {
    const lbm_param_t* p;
    const uint32_t* obstacles; /* packed, see read_obstacles */
    const float *ux, *uy, *u, *pressure;
    int row0, row1;
    char* buf;
    size_t len;
    int failed;
} format_job;

static void* format_rows(void* arg)
{
    format_job* j = (format_job*)arg;
    const int nx = j->p->nx;
    const size_t wpr = ((size_t)nx + 31) / 32;
    const size_t cap = (size_t)(j->row1 - j->row0) * (size_t)nx * 112 + 16; /* a line is at most 2*11+4*20+1+7 bytes */
    j->buf = malloc(cap);
    if (!j->buf) {
        j->failed = 1;
        return NULL;
    }
    size_t at = 0;
    for (int jj = j->row0; jj < j->row1; jj++)
        for (int ii = 0; ii < nx; ii++) {
            const size_t c = (size_t)ii + (size_t)jj * nx;
            at += (size_t)snprintf(j->buf + at, cap - at, "%d %d %.12E %.12E %.12E %.12E %d\n", ii, jj, j->ux[c], j->uy[c], j->u[c],
                                   j->pressure[c], obstacle_at(j->obstacles, wpr, ii, jj));
        }
    j->len = at;
    return NULL;
}

static void write_final_state(const lbm_param_t* p, const uint32_t* obstacles, const float* ux, const float* uy, const float* u,
                              const float* pressure)
{
    FILE* fp = fopen(FINAL_STATE_FILE, "w");
    if (!fp) DIE("could not open file output file");
    long ncpu = sysconf(_SC_NPROCESSORS_ONLN);
    int nthreads = (int)(ncpu < 1 ? 1 : (ncpu > 64 ? 64 : ncpu));
    /* one band of rows per thread and round; at most ~1 M cells per band keeps the text buffers bounded
     * (112 MB each) */
    const size_t cells_per_band = 1u << 20;
    int band_rows = (int)(cells_per_band / (size_t)p->nx);
    const int even_rows = (p->ny + nthreads - 1) / nthreads;
    if (band_rows > even_rows) band_rows = even_rows;
    if (band_rows < 1) band_rows = 1;
    format_job* jobs = calloc((size_t)nthreads, sizeof *jobs);
    pthread_t* tids = calloc((size_t)nthreads, sizeof *tids);
    if (!jobs || !tids) DIE("cannot allocate memory for the output threads");
    int row = 0;
    while (row < p->ny) {
        int n = 0;
        for (; n < nthreads && row < p->ny; n++) {
            const int r1 = row + band_rows < p->ny ? row + band_rows : p->ny;
            format_job j = {p, obstacles, ux, uy, u, pressure, row, r1, NULL, 0, 0};
            jobs[n] = j;
            row = r1;
            if (pthread_create(&tids[n], NULL, format_rows, &jobs[n]) != 0) {
                format_rows(&jobs[n]); /* no thread: format here */
                tids[n] = 0;
            }
        }
        for (int i = 0; i < n; i++) {
            if (tids[i]) pthread_join(tids[i], NULL);
            if (jobs[i].failed) DIE("cannot allocate memory for the output buffers");
            if (fwrite(jobs[i].buf, 1, jobs[i].len, fp) != jobs[i].len) DIE("could not write file output file");
            free(jobs[i].buf);
        }
    }
    free(jobs);
    free(tids);
    fclose(fp);
}

static void write_av_vels(const lbm_param_t* p, const float* av_vels)
{
    FILE* fp = fopen(AV_VELS_FILE, "w");
    if (!fp) DIE("could not open file output file");
    for (int tt = 0; tt < p->maxIters; tt++) fprintf(fp, "%d:\t%.12E\n", tt, av_vels[tt]);
    fclose(fp);
}

/* one animation frame: header, then |u| of every cell in row-major order, "%.6E" (SerialCode:802-849) */
static void write_animation_frame(const lbm_param_t* p, const float* u, int timestep)
{
    char filename[256];
    snprintf(filename, sizeof filename, "animation_data/velocity_magnitude_%06d.dat", timestep);
    FILE* fp = fopen(filename, "w");
    if (!fp) DIE("could not open animation data file");
    fprintf(fp, "# nx=%d ny=%d timestep=%d\n", p->nx, p->ny, timestep);
    const size_t n = (size_t)p->nx * (size_t)p->ny;
    for (size_t c = 0; c < n; c++) fprintf(fp, "%.6E\n", u[c]);
    fclose(fp);
    printf("Written animation data for timestep %d\n", timestep);
}

static int env_int(const char* name, int dflt)
{
    const char* v = getenv(name);
    return (v && *v) ? atoi(v) : dflt;
}

int main(int argc, char* argv[])
{
    if (argc != 3) usage(argv[0]);
    const char* paramfile = argv[1];
    const char* obstaclefile = argv[2];

    /* ---- init: parse, create the device lattice (SerialCode:156-159) ---- */
    const double tot_tic = wall_seconds();
    lbm_param_t params;
    read_params(paramfile, &params);
    if (params.nx < 1 || params.ny < 2 || params.maxIters < 0) DIE("grid size / iteration count out of range");
    const double parse_tic = wall_seconds();
    uint32_t* obstacles = read_obstacles(obstaclefile, &params); /* 1 bit per cell */
    const double parse_toc = wall_seconds();
    if (env_int("LBM_PARSE_ONLY", 0)) {
        const size_t nw = (((size_t)params.nx + 31) / 32) * (size_t)params.ny;
        unsigned long long blocked = 0, h = 1469598103934665603ULL;
        for (size_t i = 0; i < nw; i++) {
            blocked += (unsigned long long)__builtin_popcount(obstacles[i]);
            for (int b = 0; b < 4; b++) h = (h ^ ((obstacles[i] >> (8 * b)) & 0xffu)) * 1099511628211ULL;
        }
        printf("parsed: nx=%d ny=%d maxIters=%d blocked=%llu fnv1a=%016llx obstacle_parse_seconds=%.6f\n", params.nx, params.ny,
               params.maxIters, blocked, h, parse_toc - parse_tic);
        free(obstacles);
        return EXIT_SUCCESS;
    }

    lbm_options_t opt;
    lbm_default_options(&opt);
    const char* s;
    if ((s = getenv("LBM_ARITH")) && strcmp(s, "fast") == 0) opt.arith = LBM_ARITH_FAST;
    if ((s = getenv("LBM_HALO_MODE")) && strcmp(s, "async") == 0) opt.halo_mode = LBM_HALO_ASYNC;
    opt.halo_lag = env_int("LBM_HALO_LAG", 0);
    opt.kernel = env_int("LBM_KERNEL", 0);
    opt.block = env_int("LBM_BLOCK", 0);
    opt.use_graph = env_int("LBM_GRAPH", 1);
    const int ngpus = env_int("LBM_GPUS", 1);
    const int skip_final = env_int("LBM_SKIP_FINAL_STATE", 0);

    lbm_lattice_t* lat = NULL;
    LBM_CALL(lbm_create_packed(&params, obstacles, ngpus, &opt, &lat));
    float* av_vels = malloc(sizeof(float) * (size_t)(params.maxIters > 0 ? params.maxIters : 1));
    if (!av_vels) DIE("cannot allocate memory for av_vels");
    LBM_CALL(lbm_sync(lat));
    const double init_toc = wall_seconds();

    /* ---- compute: the whole `for tt` loop runs on the device (SerialCode:166-169) ---- */
    const int frame_every = env_int("LBM_ANIMATION_EVERY", 0);
    float device_ms = 0.f;
    if (frame_every <= 0) {
        LBM_CALL(lbm_run(lat, params.maxIters));
        LBM_CALL(lbm_sync(lat));
        LBM_CALL(lbm_last_run_ms(lat, &device_ms));
    } else {
        /* frames after timestep tt = 0, N, 2N, ...: run up to the next frame, collect that chunk's av_vels */
        mkdir("animation_data", 0777);
        float* frame = malloc(sizeof(float) * (size_t)params.nx * (size_t)params.ny);
        if (!frame) DIE("cannot allocate memory for the animation frame");
        int done = 0;
        while (done < params.maxIters) {
            const int next_frame_tt = ((done + frame_every - 1) / frame_every) * frame_every; /* first tt >= done on the grid */
            int chunk = next_frame_tt - done + 1;
            if (chunk > params.maxIters - done) chunk = params.maxIters - done;
            float ms = 0.f;
            LBM_CALL(lbm_run(lat, chunk));
            LBM_CALL(lbm_av_vels(lat, av_vels + done, chunk));
            LBM_CALL(lbm_last_run_ms(lat, &ms));
            device_ms += ms;
            done += chunk;
            if ((done - 1) % frame_every == 0) {
                LBM_CALL(lbm_final_state(lat, NULL, NULL, frame, NULL));
                write_animation_frame(&params, frame, done - 1);
            }
        }
        free(frame);
    }
    const double comp_toc = wall_seconds();

    /* ---- collate: per-step sums -> av_vels, moments of the final state -> host (the MPI variants'
     * gather + MPI_Reduce, MPI/d2q9-bgk.c:265-309) ---- */
    if (frame_every <= 0) LBM_CALL(lbm_av_vels(lat, av_vels, params.maxIters));
    const size_t n = (size_t)params.nx * (size_t)params.ny;
    float *ux = NULL, *uy = NULL, *u = NULL, *pressure = NULL;
    if (!skip_final) {
        ux = malloc(n * sizeof(float));
        uy = malloc(n * sizeof(float));
        u = malloc(n * sizeof(float));
        pressure = malloc(n * sizeof(float));
        if (!ux || !uy || !u || !pressure) DIE("cannot allocate memory for the final state");
        LBM_CALL(lbm_final_state(lat, ux, uy, u, pressure));
    }
    float av_final = 0.f;
    LBM_CALL(lbm_av_velocity(lat, &av_final));
    const double col_toc = wall_seconds();

    /* ---- report (SerialCode:194-201); calc_reynolds :637-642 ---- */
    const float viscosity = 1.f / 6.f * (2.f / params.omega - 1.f);
    const float reynolds = av_final * params.reynolds_dim / viscosity;
    printf("==done==\n");
    printf("Reynolds number:\t\t%.12E\n", reynolds);
    printf("Elapsed Init time:\t\t\t%.6lf (s)\n", init_toc - tot_tic);
    printf("Elapsed Compute time:\t\t\t%.6lf (s)\n", comp_toc - init_toc);
    printf("Elapsed Collate time:\t\t\t%.6lf (s)\n", col_toc - comp_toc);
    printf("Elapsed Total time:\t\t\t%.6lf (s)\n", col_toc - tot_tic);
    /* extra, machine-readable (after the reference's lines; 72 B per lattice update) */
    if (params.maxIters > 0 && device_ms > 0.f) {
        const double lups = (double)n * params.maxIters / (device_ms * 1e-3);
        printf("B200: gpus=%d slabs=%d arith=%s halo=%s MLUPS=%.1f GB/s=%.1f kernel_ms_per_step=%.6f launches=%lld\n", ngpus,
               lbm_num_slabs(lat), opt.arith == LBM_ARITH_STRICT ? "strict" : "fast",
               opt.halo_mode == LBM_HALO_SYNC ? "sync" : "async", lups * 1e-6, lups * 72e-9, device_ms / params.maxIters,
               lbm_kernel_launches(lat));
    }
    fflush(stdout);

    if (!skip_final) write_final_state(&params, obstacles, ux, uy, u, pressure);
    write_av_vels(&params, av_vels);

    /* finalise (SerialCode:615-634) */
    lbm_destroy(lat);
    free(ux);
    free(uy);
    free(u);
    free(pressure);
    free(av_vels);
    free(obstacles);
    return EXIT_SUCCESS;
}
