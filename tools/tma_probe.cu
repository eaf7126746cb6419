// tma_probe.cu -- stand-alone check of the TMA mechanics step_tma_kernel relies on (3-D tensor map over
// [plane][row][pitch] floats, 128 x TY x 1 boxes, negative / out-of-range start coordinates -> zero fill).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o /tmp/tma_probe tools/tma_probe.cu
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../lbm-asynchronous_b200/csrc/lbm_tma_kernel.cuh"

using namespace lbm;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

template <int TY, int W>
__global__ void probe(const __grid_constant__ CUtensorMap tmap, float* out, int c0, int c1, int c2, int nloads)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* tile = reinterpret_cast<float*>(smem_raw);
    __shared__ __align__(8) uint64_t bar;
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_expect_tx(&bar, nloads * TY * W * 4);
        for (int k = 0; k < nloads; k++) tma_load_3d(tile + k * ((TY * W + 31) / 32 * 32), &tmap, &bar, c0, c1, c2 + k);
    }
    mbar_wait(&bar, 0);
    for (int k = 0; k < nloads; k++)
        for (int i = threadIdx.x; i < TY * W; i += blockDim.x) out[k * TY * W + i] = tile[k * ((TY * W + 31) / 32 * 32) + i];
}

template <int TY, int W>
int run(int argc, char** argv)
{
    const int nx = 1000, rows = 40, pitch = 1024;
    const size_t pf = (size_t)rows * pitch;
    std::vector<float> h(9 * pf);
    for (int k = 0; k < 9; k++) for (int y = 0; y < rows; y++) for (int x = 0; x < pitch; x++) h[k * pf + (size_t)y * pitch + x] = k * 1000000.f + y * 1000.f + x;
    float *d, *o;
    CK(cudaMalloc(&d, h.size() * 4));
    CK(cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&o, 9 * TY * W * 4));
    void* fn = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    auto encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled>(fn);
    CUtensorMap map;
    const cuuint64_t gdim[3] = {(cuuint64_t)nx, (cuuint64_t)rows, 9};
    const cuuint64_t gstr[2] = {(cuuint64_t)pitch * 4, (cuuint64_t)pf * 4};
    const cuuint32_t box[3] = {W, TY, 1}, es[3] = {1, 1, 1};
    CUresult r = encode(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode -> %d\n", (int)r);
    const int smem = 9 * ((TY * W + 31) / 32 * 32) * 4;
    CK((cudaFuncSetAttribute(probe<TY, W>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)));
    printf("TY=%d W=%d\n", TY, W);
    struct Case { int c0, c1, c2, n; };
    std::vector<Case> cases;
    if (argc >= 5) cases.push_back({atoi(argv[1]), atoi(argv[2]), atoi(argv[3]), atoi(argv[4])});
    else cases = {{0, 0, 0, 1}, {128, 8, 3, 1}, {896, 35, 2, 1}, {256, -1, 0, 9}};
    for (auto c : cases) {
        probe<TY, W><<<1, 256, smem>>>(map, o, c.c0, c.c1, c.c2, c.n);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("case (%d,%d,%d,%d): %s\n", c.c0, c.c1, c.c2, c.n, cudaGetErrorString(e)); return 1; }
        std::vector<float> got(c.n * TY * W);
        CK(cudaMemcpy(got.data(), o, got.size() * 4, cudaMemcpyDeviceToHost));
        long bad = 0;
        for (int k = 0; k < c.n; k++) for (int j = 0; j < TY; j++) for (int i = 0; i < W; i++) {
            const int x = c.c0 + i, y = c.c1 + j, p = c.c2 + k;
            const float want = (x < 0 || x >= nx || y < 0 || y >= rows || p < 0 || p >= 9) ? 0.f : p * 1000000.f + y * 1000.f + x;
            if (got[(k * TY + j) * W + i] != want) bad++;
        }
        printf("case (%d,%d,%d) x%d: %ld mismatches\n", c.c0, c.c1, c.c2, c.n, bad);
    }
    return 0;
}

int main(int argc, char** argv)
{
    const int ty = argc > 5 ? atoi(argv[5]) : 8, w = argc > 6 ? atoi(argv[6]) : 128;
    if (ty == 8 && w == 128) return run<8, 128>(argc, argv);
    if (ty == 8 && w == 132) return run<8, 132>(argc, argv);
    if (ty == 4 && w == 128) return run<4, 128>(argc, argv);
    if (ty == 4 && w == 132) return run<4, 132>(argc, argv);
    if (ty == 16 && w == 132) return run<16, 132>(argc, argv);
    return 2;
}
