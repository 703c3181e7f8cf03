// Secondary rooflines of the line-integration path on this GPU: random 4-byte
// gathers from an L2-resident volume, fp32 red.global.add to random voxels of it, dependent-free
// fp64 FMA issue.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o peaks peaks.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t lcg(uint32_t& s) { s = s * 1664525u + 1013904223u; return s; }

__global__ void gather_kernel(const float* __restrict__ vol, uint32_t nvox, int iters, float* out)
{
    uint32_t s = blockIdx.x * blockDim.x + threadIdx.x + 1;
    float acc = 0.f;
    for (int k = 0; k < iters; k++) {
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; j++) v[j] = __ldg(vol + (lcg(s) % nvox));
#pragma unroll
        for (int j = 0; j < 8; j++) acc += v[j];
    }
    if (acc == 123.456f) out[0] = acc;
}

// like a ray: consecutive gathers of a thread hit neighbouring voxels (azimuth fastest)
__global__ void gather_walk_kernel(const float* __restrict__ vol, uint32_t nvox, int iters, float* out)
{
    uint32_t s = blockIdx.x * blockDim.x + threadIdx.x + 1;
    float acc = 0.f;
    uint32_t idx = lcg(s) % nvox;
    for (int k = 0; k < iters * 8; k++) {
        acc += __ldg(vol + idx);
        const uint32_t r = lcg(s) >> 30;           // a step in azimuth (x2), elevation or radius
        idx = (idx + (r < 2 ? 1u : (r == 2 ? 360u : 64800u))) % nvox;
    }
    if (acc == 123.456f) out[0] = acc;
}

__global__ void red_kernel(float* vol, uint32_t nvox, int iters)
{
    uint32_t s = blockIdx.x * blockDim.x + threadIdx.x + 1;
    for (int k = 0; k < iters * 8; k++) atomicAdd(vol + (lcg(s) % nvox), 1.0f);
}

__global__ void dfma_kernel(double* out, int iters)
{
    double a[8], b = 1.0000001, c = 1e-9;
#pragma unroll
    for (int j = 0; j < 8; j++) a[j] = threadIdx.x + j;
    for (int k = 0; k < iters; k++) {
#pragma unroll
        for (int j = 0; j < 8; j++) a[j] = fma(a[j], b, c);
    }
    double sum = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) sum += a[j];
    if (sum == 123.456) out[0] = sum;
}

template <class F> float time_ms(F launch, int reps = 5)
{
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    launch(); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < reps; r++) {
        cudaEventRecord(a); launch(); cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
    }
    return best;
}

int main()
{
    const uint32_t nvox = 200u * 180u * 360u;          // the config-5 volume: 51.8 MB
    float* vol; float* out; double* dout;
    cudaMalloc(&vol, nvox * sizeof(float)); cudaMemset(vol, 0, nvox * sizeof(float));
    cudaMalloc(&out, 64); cudaMalloc(&dout, 64);
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int blocks = sms * 8, threads = 256, iters = 2000;
    const double n = (double)blocks * threads * iters * 8;
    float ms = time_ms([&] { gather_kernel<<<blocks, threads>>>(vol, nvox, iters, out); });
    printf("{\"l2_random_gather_per_s\": %.4g, ", n / (ms * 1e-3));
    ms = time_ms([&] { gather_walk_kernel<<<blocks, threads>>>(vol, nvox, iters, out); });
    printf("\"l2_raylike_gather_per_s\": %.4g, ", n / (ms * 1e-3));
    ms = time_ms([&] { red_kernel<<<blocks, threads>>>(vol, nvox, iters / 4); });
    printf("\"l2_random_red_f32_per_s\": %.4g, ", n / 4 / (ms * 1e-3));
    ms = time_ms([&] { dfma_kernel<<<blocks, threads>>>(dout, iters * 4); });
    printf("\"fp64_fma_per_s\": %.4g, \"sms\": %d}\n", n * 4 / (ms * 1e-3), sms);
    return 0;
}
