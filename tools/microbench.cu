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


// ---- second-generation probes: cheap index generation (LCG + mulhi) so the load path, not the
// integer pipe, is what saturates.
__device__ __forceinline__ unsigned lcg(unsigned& x) { x = x * 1664525u + 1013904223u; return x; }

struct __align__(32) D4 { double x, y, z, w; };
struct __align__(16) D2 { double x, y; };

// MODE 0: one 32-byte aligned LDG.256 per row (window layout); 1: 4 x LDG.64 at an arbitrary 8-byte
// offset; 2: 2 x LDG.128 at a 16-byte aligned offset; 3: one LDG.128 (pair layout); 4: one LDG.64.
template <int MODE>
__global__ void row_gather_kernel(const double* __restrict__ tab, unsigned nrows_or_n, int per, double* out) {
    unsigned x = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
    double s = 0;
    for (int i = 0; i < per; i += 8) {
        double v[8][4];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            unsigned r = __umulhi(lcg(x), nrows_or_n);
            if (MODE == 0) {
                D4 q = reinterpret_cast<const D4*>(tab)[r];
                v[k][0] = q.x; v[k][1] = q.y; v[k][2] = q.z; v[k][3] = q.w;
            } else if (MODE == 1) {
#pragma unroll
                for (int j = 0; j < 4; ++j) v[k][j] = __ldg(tab + r + j);
            } else if (MODE == 2) {
                D2 a = reinterpret_cast<const D2*>(tab)[r], b = reinterpret_cast<const D2*>(tab)[r + 1];
                v[k][0] = a.x; v[k][1] = a.y; v[k][2] = b.x; v[k][3] = b.y;
            } else if (MODE == 3) {
                D2 a = reinterpret_cast<const D2*>(tab)[r];
                v[k][0] = a.x; v[k][1] = a.y; v[k][2] = 0; v[k][3] = 0;
            } else {
                v[k][0] = __ldg(tab + r); v[k][1] = v[k][2] = v[k][3] = 0;
            }
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) s += (v[k][0] + v[k][1]) + (v[k][2] + v[k][3]);
    }
    if (s == 123.456) out[0] = s;
}

// Quad-contiguous gathers (cubic_quad4.cuh): the four lanes of a quad load four consecutive 32-byte sectors (LDG.256
// each) starting at a random sector (ALIGNED = 0: any sector, the 128 bytes straddle two lines 75 % of the time;
// 1: a multiple of four sectors = one line). Counts sectors: the L2 -> L1 rate the coefficient layout can reach.
template <int ALIGNED>
__global__ void quad_gather_kernel(const double* __restrict__ tab, unsigned nsectors, int per, double* out) {
    unsigned x = ((blockIdx.x * blockDim.x + threadIdx.x) >> 2) * 2654435761u + 12345u;  // quad-uniform stream
    const unsigned j = threadIdx.x & 3u;
    double s = 0;
    for (int i = 0; i < per; i += 8) {
        D4 v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            unsigned r = __umulhi(lcg(x), nsectors - 4);
            if (ALIGNED) r &= ~3u;
            v[k] = reinterpret_cast<const D4*>(tab)[r + j];
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) s += (v[k].x + v[k].y) + (v[k].z + v[k].w);
    }
    if (s == 123.456) out[0] = s;
}

// The same with W consecutive 128-byte lines per quad (aligned to W lines): W loads per lane, chunks of W*128 bytes.
template <int W>
__global__ void chunk_gather_kernel(const double* __restrict__ tab, unsigned nsectors, int per, double* out) {
    unsigned x = ((blockIdx.x * blockDim.x + threadIdx.x) >> 2) * 2654435761u + 12345u;  // quad-uniform stream
    const unsigned j = threadIdx.x & 3u;
    double s = 0;
    for (int i = 0; i < per; i += 8 / W) {
        D4 v[8];
#pragma unroll
        for (int k = 0; k < 8 / W; ++k) {
            unsigned r = __umulhi(lcg(x), nsectors - 4 * W) & ~(4u * W - 1u);
#pragma unroll
            for (int w = 0; w < W; ++w) v[k * W + w] = reinterpret_cast<const D4*>(tab)[r + 4 * w + j];
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) s += (v[k].x + v[k].y) + (v[k].z + v[k].w);
    }
    if (s == 123.456) out[0] = s;
}

// Random shared-memory gathers with cheap indices. W = 1: LDS.64, 2: LDS.128 (16-byte aligned pair).
template <int W>
__global__ void smem_gather2_kernel(const double* __restrict__ tab, int n, int per, double* out) {
    extern __shared__ __align__(16) double sm[];
    for (int i = threadIdx.x; i < n; i += blockDim.x) sm[i] = tab[i];
    __syncthreads();
    unsigned x = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 777u;
    double s0 = 0, s1 = 0;
    for (int i = 0; i < per; i += 8) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            if (W == 1) {
                unsigned idx = __umulhi(lcg(x), (unsigned)n);
                s0 += sm[idx];
            } else {
                unsigned idx = __umulhi(lcg(x), (unsigned)(n / 2));
                D2 q = reinterpret_cast<const D2*>(sm)[idx];
                s0 += q.x; s1 += q.y;
            }
        }
    }
    if (s0 + s1 == 123.456) out[0] = s0;
}

// Division probes. MODE 0: __ddiv_rn(a, b); 1: reciprocal-residual sequence (5 FP64 ops, b fixed).
template <int MODE>
__global__ void div_kernel(double* out, int iters, double b, double rb) {
    double a[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] = 1.0 + k + threadIdx.x * 1e-3;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            double q;
            if (MODE == 0) q = __ddiv_rn(a[k], b);
            else {
                double q0 = __dmul_rn(a[k], rb);
                double e0 = __fma_rn(-q0, b, a[k]);
                double q1 = __fma_rn(e0, rb, q0);
                double e1 = __fma_rn(-q1, b, a[k]);
                q = __fma_rn(e1, rb, q1);
            }
            a[k] = __dadd_rn(q, 1.0);
        }
    }
    double s = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += a[k];
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

    // ---- row gathers from L2-resident tables, cheap indices
    {
        const int per = 256, blocks = sms * 16, threads = 256;
        double rows = double(blocks) * threads * per;
        const unsigned n32 = 4000000;  // 32 MB window layout of a 100^3 grid (4 doubles per row)
        double* tab; CK(cudaMalloc(&tab, size_t(n32) * 8 + 64)); CK(cudaMemset(tab, 0, size_t(n32) * 8 + 64));
        float t0 = time_ms([&] { row_gather_kernel<0><<<blocks, threads>>>(tab, n32 / 4, per, out); });
        float t1 = time_ms([&] { row_gather_kernel<1><<<blocks, threads>>>(tab, 1000000 - 4, per, out); });
        float t2 = time_ms([&] { row_gather_kernel<2><<<blocks, threads>>>(tab, 500000 - 2, per, out); });
        float t3 = time_ms([&] { row_gather_kernel<3><<<blocks, threads>>>(tab, 1000000, per, out); });
        float t4 = time_ms([&] { row_gather_kernel<4><<<blocks, threads>>>(tab, 1000000, per, out); });
        float t5 = time_ms([&] { row_gather_kernel<4><<<blocks, threads>>>(tab, 4096, per, out); });
        float t6 = time_ms([&] { row_gather_kernel<0><<<blocks, threads>>>(tab, 1024, per, out); });
        printf(", \"row32B_ldg256_L2_Grows_s\": %.2f, \"row4x_ldg64_L2_Grows_s\": %.2f, \"row2x_ldg128_L2_Grows_s\": %.2f"
               ", \"pair_ldg128_L2_Gloads_s\": %.2f, \"ldg64_L2_Gloads_s\": %.2f, \"ldg64_L1hit_Gloads_s\": %.2f, \"ldg256_L1hit_Grows_s\": %.2f",
               rows / t0 * 1e-6, rows / t1 * 1e-6, rows / t2 * 1e-6, rows / t3 * 1e-6, rows / t4 * 1e-6, rows / t5 * 1e-6, rows / t6 * 1e-6);
        {   // the same from HBM: 128-byte lines at random from a 2 GB table (16-value hypercube blocks of a 4-D multilinear grid)
            const size_t big_sectors = size_t(1) << 26;  // 2 GiB
            double* big; CK(cudaMalloc(&big, big_sectors * 32)); CK(cudaMemset(big, 0, big_sectors * 32));
            float h1 = time_ms([&] { quad_gather_kernel<1><<<blocks, threads>>>(big, unsigned(big_sectors), per, out); });
            float h0 = time_ms([&] { quad_gather_kernel<0><<<blocks, threads>>>(big, unsigned(big_sectors), per, out); });
            printf(", \"quad128B_aligned_hbm_Glines_s\": %.2f, \"quad128B_unaligned_hbm_Gquads_s\": %.2f", rows / 4 / h1 * 1e-6, rows / 4 / h0 * 1e-6);
            float c1 = time_ms([&] { chunk_gather_kernel<1><<<blocks, threads>>>(big, unsigned(big_sectors), per, out); });
            float c2 = time_ms([&] { chunk_gather_kernel<2><<<blocks, threads>>>(big, unsigned(big_sectors), per, out); });
            float c4 = time_ms([&] { chunk_gather_kernel<4><<<blocks, threads>>>(big, unsigned(big_sectors), per, out); });
            printf(", \"chunk128B_hbm_TBs\": %.3f, \"chunk256B_hbm_TBs\": %.3f, \"chunk512B_hbm_TBs\": %.3f", rows * 32 / c1 * 1e-9, rows * 32 / c2 * 1e-9, rows * 32 / c4 * 1e-9);
            CK(cudaFree(big));
        }
        float q0 = time_ms([&] { quad_gather_kernel<0><<<blocks, threads>>>(tab, n32 / 4, per, out); });
        float q1 = time_ms([&] { quad_gather_kernel<1><<<blocks, threads>>>(tab, n32 / 4, per, out); });
        printf(", \"quad128B_ldg256_L2_Gsectors_s\": %.2f, \"quad128B_aligned_ldg256_L2_Gsectors_s\": %.2f", rows / q0 * 1e-6, rows / q1 * 1e-6);
        CK(cudaFree(tab));
    }
    // ---- shared-memory gathers, cheap indices (96 KB tile, 2 CTAs per SM)
    {
        const int n = 12288, per = 4096, blocks = sms * 2, threads = 512;
        double* tab; CK(cudaMalloc(&tab, n * sizeof(double))); CK(cudaMemset(tab, 0, n * sizeof(double)));
        CK(cudaFuncSetAttribute(smem_gather2_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, n * 8));
        CK(cudaFuncSetAttribute(smem_gather2_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, n * 8));
        double loads = double(blocks) * threads * per;
        float t1 = time_ms([&] { smem_gather2_kernel<1><<<blocks, threads, n * sizeof(double)>>>(tab, n, per, out); });
        float t2 = time_ms([&] { smem_gather2_kernel<2><<<blocks, threads, n * sizeof(double)>>>(tab, n, per, out); });
        printf(", \"lds64_random_Gloads_s\": %.2f, \"lds128_random_Gloads_s\": %.2f", loads / t1 * 1e-6, loads / t2 * 1e-6);
        CK(cudaFree(tab));
    }
    // ---- FP64 division
    {
        const int iters = 1024, blocks = sms * 8, threads = 256;
        double ops = double(blocks) * threads * iters * 8;
        const double b = 100.0 / 99.0;
        float t0 = time_ms([&] { div_kernel<0><<<blocks, threads>>>(out, iters, b, 1.0 / b); });
        float t1 = time_ms([&] { div_kernel<1><<<blocks, threads>>>(out, iters, b, 1.0 / b); });
        printf(", \"ddiv_rn_Gdiv_s\": %.2f, \"recip_residual_div_Gdiv_s\": %.2f", ops / t0 * 1e-6, ops / t1 * 1e-6);
    }
    printf("}\n");
    return 0;
}
