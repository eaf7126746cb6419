/*
 * gen_channel -- writes the synthetic channel-flow inputs of the benchmark configurations
 *  in the reference's own file formats, so that the same files drive
 * d2q9-bgk and the reference programs:
 *
 *     gen_channel <nx> <ny> <maxIters> <paramfile> <obstaclefile> [p=0.005] [seed=42]
 *
 * Geometry: rows 0 and ny-1 fully blocked (channel walls), x periodic; every cell of rows 1..ny-3
 * blocked independently with probability p (row ny-2, the driven row, stays clear).  Random
 * numbers: SplitMix64 seeded with `seed`, one draw per cell of rows 1..ny-3 in y-major order,
 * blocked iff (draw >> 11) * 2^-53 < p.  The numpy generator in synthetic.py produces the same map.
 * Physics values as dataSet/input_128x128.params: reynolds_dim 10, density 0.1, accel 0.005,
 * omega 1.85.
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

static uint64_t splitmix64(uint64_t* state)
{
    uint64_t z = (*state += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

int main(int argc, char** argv)
{
    if (argc < 6) {
        fprintf(stderr, "Usage: %s <nx> <ny> <maxIters> <paramfile> <obstaclefile> [p] [seed]\n", argv[0]);
        return EXIT_FAILURE;
    }
    const int nx = atoi(argv[1]), ny = atoi(argv[2]), iters = atoi(argv[3]);
    const double p = argc > 6 ? atof(argv[6]) : 0.005;
    uint64_t state = argc > 7 ? strtoull(argv[7], NULL, 10) : 42ull;
    if (nx < 1 || ny < 4) {
        fprintf(stderr, "need nx >= 1 and ny >= 4\n");
        return EXIT_FAILURE;
    }
    FILE* fp = fopen(argv[4], "w");
    if (!fp) {
        perror(argv[4]);
        return EXIT_FAILURE;
    }
    fprintf(fp, "%d\n%d\n%d\n%d\n%s\n%s\n%s\n", nx, ny, iters, 10, "0.1", "0.005", "1.85");
    fclose(fp);
    fp = fopen(argv[5], "w");
    if (!fp) {
        perror(argv[5]);
        return EXIT_FAILURE;
    }
    for (int x = 0; x < nx; x++) fprintf(fp, "%d %d 1\n", x, 0);
    for (int y = 1; y <= ny - 3; y++)
        for (int x = 0; x < nx; x++) {
            const uint64_t r = splitmix64(&state);
            if ((double)(r >> 11) * (1.0 / 9007199254740992.0) < p) fprintf(fp, "%d %d 1\n", x, y);
        }
    for (int x = 0; x < nx; x++) fprintf(fp, "%d %d 1\n", x, ny - 1);
    fclose(fp);
    return EXIT_SUCCESS;
}
