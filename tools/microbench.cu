// microbench.cu — measures the secondary ceilings of the interpolation hot path on the B200
// (DESIGN.md §4): FP64 issue rate (fused and unfused), divergent global gathers from an
// L2-resident table, and random shared-memory gathers. Standalone binary; not part of the library.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench tools/microbench.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

template <int MODE>  // 0: DFMA chains, 1: DMUL+DADD (unfused pairs), 2: DADD only
__global__ void fp64_kernel(double* out, int iters, double seed) {
    double a[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] = seed + k + threadIdx.x * 1e-9;
    const double m = 1.0000001, c = 1e-7;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            if (MODE == 0) a[k] = __fma_rn(a[k], m, c);
            else if (MODE == 1) a[k] = __dadd_rn(__dmul_rn(a[k], m), c);
            else a[k] = __dadd_rn(a[k], c);
        }
    }
    double s = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += a[k];
    if (s == 123.456) out[0] = s;
}

__device__ __forceinline__ unsigned hash32(unsigned x) {
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x;
}

// Each thread performs `per` dependent-free random 8-byte loads from a table of `n` doubles.
template <int VEC>  // VEC doubles contiguous per load group (1 -> LDG.64 singles, 4 -> 4 consecutive LDG.64)
__global__ void gather_kernel(const double* __restrict__ tab, unsigned n, int per, double* out) {
    unsigned tid = blockIdx.x * blockDim.x + threadIdx.x;
    double s = 0;
    for (int i = 0; i < per; i += 8) {
        double v[8][VEC];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            unsigned idx = hash32(tid * 977u + (i + k) * 131071u) % (n - VEC);
#pragma unroll
            for (int j = 0; j < VEC; ++j) v[k][j] = __ldg(tab + idx + j);
        }
#pragma unroll
        for (int k = 0; k < 8; ++k)
#pragma unroll
            for (int j = 0; j < VEC; ++j) s += v[k][j];
    }
    if (s == 123.456) out[0] = s;
}

__global__ void smem_gather_kernel(const double* __restrict__ tab, int n, int per, double* out) {
    extern __shared__ double sm[];
    for (int i = threadIdx.x; i < n; i += blockDim.x) sm[i] = tab[i];
    __syncthreads();
    unsigned tid = blockIdx.x * blockDim.x + threadIdx.x;
    double s = 0;
    for (int i = 0; i < per; i += 8) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            unsigned idx = hash32(tid * 977u + (i + k) * 131071u) % (unsigned)n;
            s += sm[idx];
        }
    }
    if (s == 123.456) out[0] = s;
}

template <class F>
float time_ms(F f, int reps = 5) {
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    f();
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        CK(cudaEventRecord(a));
        f();
        CK(cudaEventRecord(b));
        CK(cudaEventSynchronize(b));
        float ms; CK(cudaEventElapsedTime(&ms, a, b));
        if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int sms = p.multiProcessorCount;
    printf("{\"device\": \"%s\", \"sms\": %d", p.name, sms);
    double* out; CK(cudaMalloc(&out, 64));
    // ---- FP64 issue rate
    {
        const int iters = 4096, blocks = sms * 8, threads = 256;
        double ops = double(blocks) * threads * iters * 8;
        float t0 = time_ms([&] { fp64_kernel<0><<<blocks, threads>>>(out, iters, 1.0); });
        float t1 = time_ms([&] { fp64_kernel<1><<<blocks, threads>>>(out, iters, 1.0); });
        float t2 = time_ms([&] { fp64_kernel<2><<<blocks, threads>>>(out, iters, 1.0); });
        printf(", \"dfma_Tinstr_s\": %.3f, \"dmul_dadd_Tinstr_s\": %.3f, \"dadd_Tinstr_s\": %.3f", ops / t0 * 1e-9, 2 * ops / t1 * 1e-9, ops / t2 * 1e-9);
    }
    // ---- divergent global gathers from an 8 MB (L2-resident) table
    {
        const unsigned n = 1000000;
        double* tab; CK(cudaMalloc(&tab, n * sizeof(double))); CK(cudaMemset(tab, 0, n * sizeof(double)));
        const int per = 256, blocks = sms * 16, threads = 256;
        double loads = double(blocks) * threads * per;
        float t1 = time_ms([&] { gather_kernel<1><<<blocks, threads>>>(tab, n, per, out); });
        float t4 = time_ms([&] { gather_kernel<4><<<blocks, threads>>>(tab, n, per, out); });
        printf(", \"gather8B_random_Gloads_s\": %.2f, \"gather4x8B_contig_Ggroups_s\": %.2f", loads / t1 * 1e-6, loads / t4 * 1e-6);
        const unsigned nbig = 200000000;  // 1.6 GB: HBM-resident
        double* big; CK(cudaMalloc(&big, size_t(nbig) * sizeof(double))); CK(cudaMemset(big, 0, size_t(nbig) * sizeof(double)));
        float tb = time_ms([&] { gather_kernel<1><<<blocks, threads>>>(big, nbig, per, out); });
        printf(", \"gather8B_random_hbm_Gloads_s\": %.2f", loads / tb * 1e-6);
        CK(cudaFree(big)); CK(cudaFree(tab));
    }
    // ---- random shared-memory gathers (32 KB tile)
    {
        const int n = 4096, per = 2048, blocks = sms * 4, threads = 256;
        double* tab; CK(cudaMalloc(&tab, n * sizeof(double))); CK(cudaMemset(tab, 0, n * sizeof(double)));
        double loads = double(blocks) * threads * per;
        float t = time_ms([&] { smem_gather_kernel<<<blocks, threads, n * sizeof(double)>>>(tab, n, per, out); });
        printf(", \"smem_gather8B_random_Gloads_s\": %.2f", loads / t * 1e-6);
        CK(cudaFree(tab));
    }
    printf("}\n");
    return 0;
}
