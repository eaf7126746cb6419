/* Self-check of oracle/minimpi (the shared-memory stand-in for the MPI subset the reference uses):
 * ring exchange with both neighbours being the same rank (P = 2: tags must keep the two directions apart),
 * un-waited receives completed by a later call's progress, Sendrecv, large messages, Reduce in rank order. */
#include <mpi.h>
#include <stdio.h>
#include <stdlib.h>

int main(int argc, char** argv)
{
    MPI_Init(&argc, &argv);
    int rank, size, bad = 0;
    MPI_Comm_size(MPI_COMM_WORLD, &size);
    MPI_Comm_rank(MPI_COMM_WORLD, &rank);
    MPI_Datatype row;
    MPI_Type_contiguous(9, MPI_FLOAT, &row);
    MPI_Type_commit(&row);
    const int up = (rank - 1 + size) % size, down = (rank + 1) % size;
    const int n = 1000;
    float* a = malloc(sizeof(float) * 9 * n);
    float* b = malloc(sizeof(float) * 9 * n);
    float* ra = malloc(sizeof(float) * 9 * n);
    float* rb = malloc(sizeof(float) * 9 * n);
    for (int step = 0; step < 50; step++) {
        for (int i = 0; i < 9 * n; i++) a[i] = rank * 1000.f + step + 0.25f, b[i] = rank * 1000.f + step + 0.5f;
        MPI_Request req[4];
        MPI_Isend(a, n, row, up, 11, MPI_COMM_WORLD, &req[2]);   /* my first row -> up, tag 11 */
        MPI_Isend(b, n, row, down, 10, MPI_COMM_WORLD, &req[3]); /* my last row -> down, tag 10 */
        MPI_Irecv(ra, n, row, up, 10, MPI_COMM_WORLD, &req[0]);  /* up's last row */
        MPI_Irecv(rb, n, row, down, 11, MPI_COMM_WORLD, &req[1]); /* down's first row */
        MPI_Waitall(4, req, MPI_STATUSES_IGNORE);
        if (ra[0] != up * 1000.f + step + 0.5f || ra[9 * n - 1] != ra[0]) bad++;
        if (rb[0] != down * 1000.f + step + 0.25f || rb[9 * n - 1] != rb[0]) bad++;
    }
    /* un-waited receive: posted now, completed by the progress of a later call (Barrier) */
    {
        MPI_Request r;
        int flag = 0, v = rank + 7, got = -1;
        MPI_Irecv(&got, 1, MPI_INT, up, 77, MPI_COMM_WORLD, &r);
        MPI_Testall(1, &r, &flag, MPI_STATUSES_IGNORE); /* may or may not be there yet */
        MPI_Send(&v, 1, MPI_INT, down, 77, MPI_COMM_WORLD);
        MPI_Barrier(MPI_COMM_WORLD);
        MPI_Barrier(MPI_COMM_WORLD);
        if (got != up + 7) bad++;
    }
    /* Sendrecv both ways with the same tag (the plain MPI variant, MPI/d2q9-bgk.c:224-231) */
    {
        float s = (float)rank, r1 = -1.f, r2 = -1.f;
        MPI_Sendrecv(&s, 1, MPI_FLOAT, up, 0, &r1, 1, MPI_FLOAT, down, 0, MPI_COMM_WORLD, MPI_STATUS_IGNORE);
        MPI_Sendrecv(&s, 1, MPI_FLOAT, down, 0, &r2, 1, MPI_FLOAT, up, 0, MPI_COMM_WORLD, MPI_STATUS_IGNORE);
        if (r1 != (float)down || r2 != (float)up) bad++;
    }
    /* a large message to rank 0 (the final gather) and a float reduce in rank order */
    {
        const int big = 3 * 1000 * 1000;
        float* m = malloc(sizeof(float) * big);
        if (rank != 0) {
            for (int i = 0; i < big; i++) m[i] = (float)(rank + i % 7);
            MPI_Send(m, big, MPI_FLOAT, 0, 5, MPI_COMM_WORLD);
        } else {
            for (int r = 1; r < size; r++) {
                MPI_Recv(m, big, MPI_FLOAT, r, 5, MPI_COMM_WORLD, MPI_STATUS_IGNORE);
                if (m[0] != (float)r || m[big - 1] != (float)(r + (big - 1) % 7)) bad++;
            }
        }
        free(m);
        float mine[3] = {1.f + rank, 0.1f * rank, 1e-8f}, tot[3] = {0, 0, 0};
        MPI_Reduce(mine, tot, 3, MPI_FLOAT, MPI_SUM, 0, MPI_COMM_WORLD);
        if (rank == 0) {
            float want[3] = {0, 0, 0};
            for (int r = 0; r < size; r++) want[0] += 1.f + r, want[1] += 0.1f * r, want[2] += 1e-8f;
            if (tot[0] != want[0] || tot[1] != want[1] || tot[2] != want[2]) bad++;
        }
    }
    int allbad = 0;
    MPI_Reduce(&bad, &allbad, 1, MPI_INT, MPI_SUM, 0, MPI_COMM_WORLD);
    if (rank == 0) printf("minimpi selfcheck: %d ranks, %d failures\n", size, allbad);
    free(a), free(b), free(ra), free(rb);
    MPI_Finalize();
    return allbad ? 1 : 0;
}
