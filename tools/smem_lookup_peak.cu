// smem_lookup_peak.cu — the ceiling the ADC kernels are measured against: conflict-free 32-bit
// shared-memory look-ups per second on this GPU, with everything else stripped away.
//
// One CTA of 1024 threads per SM holds a 96 KB table (the C4 shape: 96 x 256 floats).  Every
// thread performs dependent-free random look-ups `table[line * 32 + lane]` (line from a cheap
// per-thread LCG, so the 32 lanes of a warp always hit 32 different banks — what the XOR-swizzled
// ADC kernel achieves for real codes) and accumulates them.  Per look-up: one integer op for the
// address, one LDS, one FADD — the same 2-3 issue slots the ADC kernel cannot avoid — so the
// number is the practical peak of "look-ups", not of the LDS pipe alone.
//
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o smem_lookup_peak tools/smem_lookup_peak.cu
//   ./smem_lookup_peak            -> one JSON line
#include <cstdio>
#include <cuda_runtime.h>

constexpr int kThreads = 1024;
constexpr int kLines = 96 * 256 / 32;  // 128-byte lines of the table
constexpr int kUnroll = 32;

__global__ void __launch_bounds__(kThreads, 1) lookup_kernel(float *out, int iters) {
    extern __shared__ float table[];
    for (int i = threadIdx.x; i < kLines * 32; i += kThreads) table[i] = static_cast<float>(i & 1023) * 1e-3f;
    __syncthreads();
    const unsigned lane = threadIdx.x & 31;
    unsigned state = threadIdx.x * 2654435761u + blockIdx.x * 40503u + 1u;
    const unsigned base = static_cast<unsigned>(__cvta_generic_to_shared(table)) + lane * 4u;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < kUnroll; u += 4) {
            state = state * 1664525u + 1013904223u;
            // the four bytes of one 32-bit draw are four code values: address = table + 128 * code + 4 * lane
            // through dp4a (the ADC kernels' address arithmetic), tables 32 KB apart as an immediate
            float v0, v1, v2, v3;
            const unsigned t = (u / 4) % 3 * 32768u;
            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v0) : "r"(__dp4a(state, 0x00000080u, base + t)));
            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v1) : "r"(__dp4a(state, 0x00008000u, base + t)));
            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v2) : "r"(__dp4a(state, 0x00800000u, base + t)));
            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v3) : "r"(__dp4a(state, 0x80000000u, base + t)));
            a0 += v0;
            a1 += v1;
            a2 += v2;
            a3 += v3;
        }
    }
    out[blockIdx.x * kThreads + threadIdx.x] = (a0 + a1) + (a2 + a3);
}

int main() {
    int dev = 0, sms = 0, clock_khz = 0;
    cudaSetDevice(dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&clock_khz, cudaDevAttrClockRate, dev);
    const size_t smem = kLines * 32 * sizeof(float);
    cudaFuncSetAttribute(lookup_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    float *out = nullptr;
    cudaMalloc(&out, static_cast<size_t>(sms) * kThreads * sizeof(float));
    const int iters = 2000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    double best = 0.0, mean = 0.0;
    const int reps = 12;
    for (int rep = 0; rep < reps + 3; rep++) {
        cudaEventRecord(e0);
        lookup_kernel<<<sms, kThreads, smem>>>(out, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep < 3) continue;  // warm-up
        const double rate = static_cast<double>(sms) * kThreads * iters * kUnroll / (ms * 1e-3);
        best = rate > best ? rate : best;
        mean += rate / reps;
    }
    const cudaError_t err = cudaGetLastError();
    printf("{\"smem_lookups_per_s_best\": %.4g, \"smem_lookups_per_s_sustained\": %.4g, \"sms\": %d, \"threads_per_sm\": %d, "
           "\"table_bytes\": %zu, \"lds_pipe_nominal_per_s\": %.4g, \"sm_clock_khz_max\": %d, \"cuda_error\": \"%s\"}\n",
           best, mean, sms, kThreads, smem, static_cast<double>(sms) * 32.0 * clock_khz * 1e3, clock_khz,
           cudaGetErrorString(err));
    return err == cudaSuccess ? 0 : 1;
}
