/*
 * minimpi.c -- see mpi.h.  TEST / BASELINE INFRASTRUCTURE ONLY.
 *
 * Shared region layout: header {barrier state} followed by one FIFO per ordered rank pair that the
 * reference programs can use: ring neighbours (halo rows, the d2q9-bgk.c of MPI, MPI_Waitall and MPI_Testall_OptimizedVersion Isend/Irecv/Sendrecv) and
 * any pair involving rank 0 (obstacle scatter, final gather, reduce).  A FIFO is a byte ring of
 * records {tag, bytes, consumed} + payload; head/tail are C11-style atomics in shared memory.
 */
#define _GNU_SOURCE
#include "mpi.h"

#include <errno.h>
#include <sched.h>
#include <signal.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/time.h>
#include <sys/wait.h>
#include <unistd.h>

#define RING_NEIGHBOUR (32u << 20) /* default bytes between ring neighbours; env MINIMPI_EAGER_BYTES overrides it:
                                      the bound on how far a sender of halo rows can run ahead of its receiver */
#define RING_ROOT (192u << 20)     /* bytes: one obstacle / lattice slab */
#define TAG_REDUCE (-1001)
#define ALIGN 64u

typedef struct {
    uint64_t head;  /* consumer side: offset of the oldest record (monotonic) */
    uint64_t tail;  /* producer side: offset past the newest record (monotonic) */
    uint64_t bytes; /* capacity of data[] */
    uint64_t pad[5];
    unsigned char data[];
} ring_t;

typedef struct {
    int32_t tag;
    uint32_t consumed;
    uint64_t bytes;  /* payload bytes */
    uint64_t total;  /* record size incl. header and padding; 0 = wrap marker */
    uint64_t pad;
} rec_t;

typedef struct {
    volatile int barrier_count;
    volatile int barrier_sense;
    volatile int abort_flag;
    int nranks;
} shared_hdr;

struct minimpi_request {
    int is_recv, done, peer, tag;
    void* buf;
    uint64_t bytes;
    struct minimpi_request* next; /* posted-receive list, in posting order */
};

static int g_rank = 0, g_size = 1, g_inited = 0;
static unsigned char* g_shm = NULL;
static shared_hdr* g_hdr = NULL;
static uint64_t* g_ring_off = NULL; /* [src*size+dst] -> offset in g_shm, 0 = no ring */
static struct minimpi_request *g_posted_head = NULL, *g_posted_tail = NULL;
static pid_t* g_children = NULL;
static int g_local_sense = 0;

static void die(const char* msg)
{
    fprintf(stderr, "minimpi[rank %d]: %s\n", g_rank, msg);
    if (g_hdr) g_hdr->abort_flag = 1;
    _exit(70);
}

static ring_t* ring_of(int src, int dst)
{
    const uint64_t off = g_ring_off[(size_t)src * g_size + dst];
    if (!off) die("no FIFO between these ranks (only ring neighbours and rank 0 may talk)");
    return (ring_t*)(g_shm + off);
}

static uint64_t round_up(uint64_t v, uint64_t a) { return (v + a - 1) / a * a; }

/* ---- FIFO ------------------------------------------------------------------------------------ */
static int ring_try_push(ring_t* r, int tag, const void* buf, uint64_t bytes)
{
    const uint64_t need = round_up(sizeof(rec_t) + bytes, ALIGN);
    if (need * 2 > r->bytes) die("message larger than half a FIFO");
    uint64_t tail = r->tail;
    const uint64_t head = __atomic_load_n(&r->head, __ATOMIC_ACQUIRE);
    uint64_t pos = tail % r->bytes;
    uint64_t skip = 0;
    if (pos + need > r->bytes) skip = r->bytes - pos; /* record must be contiguous: wrap */
    if (tail + skip + need - head > r->bytes) return 0; /* full */
    if (skip) {
        if (skip >= sizeof(rec_t)) {
            rec_t* w = (rec_t*)(r->data + pos);
            w->tag = 0, w->consumed = 1, w->bytes = 0, w->total = 0; /* wrap marker */
        }
        tail += skip;
        pos = 0;
    }
    rec_t* rec = (rec_t*)(r->data + pos);
    rec->tag = tag, rec->consumed = 0, rec->bytes = bytes, rec->total = need;
    memcpy(rec + 1, buf, bytes);
    __atomic_store_n(&r->tail, tail + need, __ATOMIC_RELEASE);
    return 1;
}

/* first unconsumed record with `tag`; copies it out and marks it; then retires consumed records at the head */
static int ring_try_pop(ring_t* r, int tag, void* buf, uint64_t bytes)
{
    const uint64_t tail = __atomic_load_n(&r->tail, __ATOMIC_ACQUIRE);
    uint64_t at = r->head;
    int found = 0;
    while (at < tail) {
        const uint64_t pos = at % r->bytes;
        if (r->bytes - pos < sizeof(rec_t)) { at += r->bytes - pos; continue; }
        rec_t* rec = (rec_t*)(r->data + pos);
        if (rec->total == 0) { at += r->bytes - pos; continue; } /* wrap marker */
        if (!rec->consumed && rec->tag == tag) {
            if (rec->bytes > bytes) die("message truncated: receive buffer too small");
            memcpy(buf, rec + 1, rec->bytes);
            rec->consumed = 1;
            found = 1;
            break;
        }
        at += rec->total;
    }
    /* retire */
    uint64_t head = r->head;
    while (head < tail) {
        const uint64_t pos = head % r->bytes;
        if (r->bytes - pos < sizeof(rec_t)) { head += r->bytes - pos; continue; }
        rec_t* rec = (rec_t*)(r->data + pos);
        if (rec->total == 0) { head += r->bytes - pos; continue; }
        if (!rec->consumed) break;
        head += rec->total;
    }
    if (head != r->head) __atomic_store_n(&r->head, head, __ATOMIC_RELEASE);
    return found;
}

/* ---- progress: complete posted receives in posting order ---------------------------------------- */
static void progress(void)
{
    if (g_hdr->abort_flag) _exit(71);
    struct minimpi_request* prev = NULL;
    struct minimpi_request* q = g_posted_head;
    while (q) {
        struct minimpi_request* next = q->next;
        /* MPI matching is ordered per (source, tag): an earlier posted receive with the same source and
         * tag must match first -- it does, because the list is walked in posting order */
        if (ring_try_pop(ring_of(q->peer, g_rank), q->tag, q->buf, q->bytes)) {
            q->done = 1;
            if (prev) prev->next = next; else g_posted_head = next;
            if (g_posted_tail == q) g_posted_tail = prev;
        } else {
            prev = q;
        }
        q = next;
    }
}

static void relax(void) { sched_yield(); }

/* ---- API ------------------------------------------------------------------------------------- */
int MPI_Init(int* argc, char*** argv)
{
    (void)argc, (void)argv;
    if (g_inited) return MPI_SUCCESS;
    const char* np = getenv("MINIMPI_NP");
    g_size = np ? atoi(np) : 1;
    if (g_size < 1 || g_size > 1024) g_size = 1;
    uint64_t neighbour_cap = RING_NEIGHBOUR;
    if (getenv("MINIMPI_EAGER_BYTES")) {
        neighbour_cap = strtoull(getenv("MINIMPI_EAGER_BYTES"), NULL, 10);
        if (neighbour_cap < 4096) neighbour_cap = 4096;
    }
    /* region size and ring offsets */
    g_ring_off = calloc((size_t)g_size * g_size, sizeof(uint64_t));
    uint64_t off = round_up(sizeof(shared_hdr), 4096);
    for (int s = 0; s < g_size; s++)
        for (int d = 0; d < g_size; d++) {
            if (s == d && g_size > 1) continue; /* a ring of one talks to itself */
            const int neighbour = (d == (s + 1) % g_size) || (d == (s - 1 + g_size) % g_size);
            const int root = (s == 0 || d == 0);
            if (!neighbour && !root) continue;
            const uint64_t cap = root ? RING_ROOT : neighbour_cap;
            g_ring_off[(size_t)s * g_size + d] = off;
            off += round_up(sizeof(ring_t) + cap, 4096);
        }
    g_shm = mmap(NULL, off, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
    if (g_shm == MAP_FAILED) { perror("minimpi: mmap"); _exit(72); }
    g_hdr = (shared_hdr*)g_shm;
    g_hdr->nranks = g_size;
    for (int s = 0; s < g_size; s++)
        for (int d = 0; d < g_size; d++) {
            const uint64_t o = g_ring_off[(size_t)s * g_size + d];
            if (!o) continue;
            ring_t* r = (ring_t*)(g_shm + o);
            r->head = r->tail = 0;
            r->bytes = ((s == 0 || d == 0) ? RING_ROOT : neighbour_cap);
        }
    fflush(stdout);
    fflush(stderr);
    g_children = calloc((size_t)g_size, sizeof(pid_t));
    for (int r = 1; r < g_size; r++) {
        pid_t pid = fork();
        if (pid < 0) { perror("minimpi: fork"); _exit(73); }
        if (pid == 0) { g_rank = r; free(g_children); g_children = NULL; break; }
        g_children[r] = pid;
    }
    g_inited = 1;
    return MPI_SUCCESS;
}

int MPI_Barrier(MPI_Comm comm)
{
    (void)comm;
    progress(); /* every MPI call makes progress, also for the rank that arrives last and never spins */
    g_local_sense = !g_local_sense;
    if (__atomic_add_fetch(&g_hdr->barrier_count, 1, __ATOMIC_ACQ_REL) == g_size) {
        g_hdr->barrier_count = 0;
        __atomic_store_n(&g_hdr->barrier_sense, g_local_sense, __ATOMIC_RELEASE);
    } else {
        while (__atomic_load_n(&g_hdr->barrier_sense, __ATOMIC_ACQUIRE) != g_local_sense) { progress(); relax(); }
    }
    progress();
    return MPI_SUCCESS;
}

int MPI_Finalize(void)
{
    if (!g_inited) return MPI_SUCCESS;
    /* receives that were never completed (the reference's un-waited MPI_Testall leaks four per step) are
     * dropped here: their buffers have usually been freed by the caller's finalise() already */
    g_posted_head = g_posted_tail = NULL;
    MPI_Barrier(MPI_COMM_WORLD);
    fflush(stdout);
    fflush(stderr);
    if (g_rank != 0) _exit(0); /* children end here: rank 0 alone returns to main() and writes the results */
    int bad = 0;
    for (int r = 1; r < g_size; r++) {
        int st = 0;
        if (waitpid(g_children[r], &st, 0) < 0 || !WIFEXITED(st) || WEXITSTATUS(st) != 0) bad = 1;
    }
    if (bad) { fprintf(stderr, "minimpi: a rank did not exit cleanly\n"); _exit(74); }
    g_inited = 0;
    return MPI_SUCCESS;
}

int MPI_Comm_size(MPI_Comm comm, int* size) { (void)comm; *size = g_size; return MPI_SUCCESS; }
int MPI_Comm_rank(MPI_Comm comm, int* rank) { (void)comm; *rank = g_rank; return MPI_SUCCESS; }

int MPI_Type_contiguous(int count, MPI_Datatype oldtype, MPI_Datatype* newtype)
{
    *newtype = (MPI_Datatype)((oldtype & 0xffffff) * count); /* kind 0: opaque bytes */
    return MPI_SUCCESS;
}
int MPI_Type_commit(MPI_Datatype* type) { (void)type; return MPI_SUCCESS; }

static uint64_t nbytes(int count, MPI_Datatype t) { return (uint64_t)count * (uint64_t)(t & 0xffffff); }

int MPI_Send(const void* buf, int count, MPI_Datatype type, int dest, int tag, MPI_Comm comm)
{
    (void)comm;
    ring_t* r = ring_of(g_rank, dest);
    while (!ring_try_push(r, tag, buf, nbytes(count, type))) { progress(); relax(); }
    return MPI_SUCCESS;
}

int MPI_Isend(const void* buf, int count, MPI_Datatype type, int dest, int tag, MPI_Comm comm, MPI_Request* request)
{
    MPI_Send(buf, count, type, dest, tag, comm); /* eager: the payload is in the FIFO when this returns */
    struct minimpi_request* q = calloc(1, sizeof *q);
    q->done = 1;
    *request = q;
    return MPI_SUCCESS;
}

int MPI_Irecv(void* buf, int count, MPI_Datatype type, int source, int tag, MPI_Comm comm, MPI_Request* request)
{
    (void)comm;
    struct minimpi_request* q = calloc(1, sizeof *q);
    q->is_recv = 1, q->peer = source, q->tag = tag, q->buf = buf, q->bytes = nbytes(count, type);
    if (g_posted_tail) g_posted_tail->next = q; else g_posted_head = q;
    g_posted_tail = q;
    *request = q;
    return MPI_SUCCESS;
}

int MPI_Wait(MPI_Request* request, MPI_Status* status)
{
    (void)status;
    struct minimpi_request* q = *request;
    if (!q) return MPI_SUCCESS;
    while (!q->done) { progress(); if (!q->done) relax(); }
    free(q);
    *request = MPI_REQUEST_NULL;
    return MPI_SUCCESS;
}

int MPI_Waitall(int count, MPI_Request requests[], MPI_Status statuses[])
{
    (void)statuses;
    for (int i = 0; i < count; i++) MPI_Wait(&requests[i], MPI_STATUS_IGNORE);
    return MPI_SUCCESS;
}

int MPI_Testall(int count, MPI_Request requests[], int* flag, MPI_Status statuses[])
{
    (void)statuses;
    progress();
    int all = 1;
    for (int i = 0; i < count; i++)
        if (requests[i] && !requests[i]->done) all = 0;
    if (all)
        for (int i = 0; i < count; i++)
            if (requests[i]) { free(requests[i]); requests[i] = MPI_REQUEST_NULL; }
    *flag = all;
    return MPI_SUCCESS; /* requests that are not complete stay posted (and are leaked by the caller) */
}

int MPI_Recv(void* buf, int count, MPI_Datatype type, int source, int tag, MPI_Comm comm, MPI_Status* status)
{
    MPI_Request q;
    MPI_Irecv(buf, count, type, source, tag, comm, &q);
    return MPI_Wait(&q, status);
}

int MPI_Sendrecv(const void* sendbuf, int sendcount, MPI_Datatype sendtype, int dest, int sendtag, void* recvbuf,
                 int recvcount, MPI_Datatype recvtype, int source, int recvtag, MPI_Comm comm, MPI_Status* status)
{
    MPI_Request q;
    MPI_Irecv(recvbuf, recvcount, recvtype, source, recvtag, comm, &q);
    MPI_Send(sendbuf, sendcount, sendtype, dest, sendtag, comm);
    return MPI_Wait(&q, status);
}

int MPI_Reduce(const void* sendbuf, void* recvbuf, int count, MPI_Datatype type, MPI_Op op, int root, MPI_Comm comm)
{
    if (op != MPI_SUM || root != 0) die("MPI_Reduce: only MPI_SUM to root 0");
    const int is_float = (type >> 24) == 1;
    if (g_rank != 0) return MPI_Send(sendbuf, count, type, 0, TAG_REDUCE, comm);
    memcpy(recvbuf, sendbuf, nbytes(count, type));
    void* tmp = malloc(nbytes(count, type));
    for (int r = 1; r < g_size; r++) { /* rank order: deterministic */
        MPI_Recv(tmp, count, type, r, TAG_REDUCE, comm, MPI_STATUS_IGNORE);
        if (is_float) for (int i = 0; i < count; i++) ((float*)recvbuf)[i] += ((float*)tmp)[i];
        else for (int i = 0; i < count; i++) ((int*)recvbuf)[i] += ((int*)tmp)[i];
    }
    free(tmp);
    return MPI_SUCCESS;
}

double MPI_Wtime(void)
{
    struct timeval t;
    gettimeofday(&t, NULL);
    return t.tv_sec + t.tv_usec * 1e-6;
}

int MPI_Abort(MPI_Comm comm, int errorcode)
{
    (void)comm;
    if (g_hdr) g_hdr->abort_flag = 1;
    _exit(errorcode ? errorcode : 1);
}
