// Stand-alone check of block_radix_sort_desc_hi32 (ffx_kernels.cuh) against std::sort.
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I fast-forward-indexes_b200/csrc -o tools/radix_check tools/radix_check.cu
#include <algorithm>
#include <cstdio>
#include <random>
#include <vector>

#include "ffx_kernels.cuh"

template <int EMAX>
__global__ void sort_kernel(unsigned long long *g, int n, int lists) {
    extern __shared__ unsigned long long sk[];
    for (int l = blockIdx.x; l < lists; l += gridDim.x) {
        for (int i = threadIdx.x; i < n; i += blockDim.x) sk[i] = g[(size_t)l * n + i];
        __syncthreads();
        const size_t hist_off = ((size_t)n * 8 + 15) & ~(size_t)15;
        ffx::block_radix_sort_desc_hi32<EMAX>(sk, n, reinterpret_cast<unsigned short *>(reinterpret_cast<unsigned char *>(sk) + hist_off));
        for (int i = threadIdx.x; i < n; i += blockDim.x) g[(size_t)l * n + i] = sk[i];
        __syncthreads();
    }
}

template <int EMAX>
int run(int T, int n, int lists, int distinct) {
    std::mt19937_64 rng(n * 31 + T);
    std::vector<unsigned long long> h((size_t)lists * n), want;
    for (int l = 0; l < lists; l++)
        for (int i = 0; i < n; i++) {
            unsigned long long hi = distinct ? (rng() % distinct) * 2654435761ull : rng();
            h[(size_t)l * n + i] = ((hi & 0xffffffffull) << 32) | (unsigned)(~(unsigned)i);
        }
    want = h;
    for (int l = 0; l < lists; l++) std::sort(want.begin() + (size_t)l * n, want.begin() + (size_t)(l + 1) * n, std::greater<unsigned long long>());
    unsigned long long *d;
    cudaMalloc(&d, h.size() * 8);
    cudaMemcpy(d, h.data(), h.size() * 8, cudaMemcpyHostToDevice);
    size_t smem = (((size_t)n * 8 + 15) & ~(size_t)15) + (T / 32) * 512 + 128;
    cudaFuncSetAttribute(sort_kernel<EMAX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    sort_kernel<EMAX><<<148, T, smem>>>(d, n, lists);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(h.data(), d, h.size() * 8, cudaMemcpyDeviceToHost);
    cudaFree(d);
    size_t bad = 0;
    for (size_t i = 0; i < h.size(); i++) bad += h[i] != want[i];
    printf("EMAX %2d T %4d n %5d lists %4d distinct %6d : %s, %zu wrong (%s)\n", EMAX, T, n, lists, distinct,
           bad ? "FAIL" : "ok", bad, cudaGetErrorString(e));
    return bad != 0;
}

int main() {
    int bad = 0;
    for (int distinct : {0, 7, 300}) {
        bad += run<8>(1024, 5000, 600, distinct);
        bad += run<8>(512, 2200, 600, distinct);
        bad += run<8>(1024, 8192, 300, distinct);
        bad += run<8>(256, 2048, 600, distinct);
        bad += run<16>(448, 5000, 600, distinct);
        bad += run<16>(256, 4096, 600, distinct);
        bad += run<16>(512, 6000, 400, distinct);
        bad += run<8>(1024, 3000, 600, distinct);
    }
    return bad ? 1 : 0;
}
