// Stand-alone sweep over the all-pairs kernel table (allpairs.cuh): times every variant on
// synthetic bodies and prints interactions/s next to the pure-FMA issue rate of the device.
// Usage: tune_allpairs [N=262144] [reps=3]
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
#include <algorithm>

#include "../procedural-universe_b200/csrc/allpairs.cuh"

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

template <bool PACKED>
__global__ void __launch_bounds__(256, 2) k_probe(float* out, int iters, float a, float b)
{
    float2 acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = make_float2(threadIdx.x * 1e-3f + k, threadIdx.x * 2e-3f - k);
    const float2 a2 = make_float2(a, a), b2 = make_float2(b, b);
    for (int i = 0; i < iters; ++i)
    {
#pragma unroll
        for (int rep = 0; rep < 4; ++rep)
#pragma unroll
            for (int k = 0; k < 8; ++k)
            {
                if (PACKED) acc[k] = __ffma2_rn(acc[k], a2, b2);
                else { acc[k].x = fmaf(acc[k].x, a, b); acc[k].y = fmaf(acc[k].y, a, b); }
            }
    }
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += acc[k].x + acc[k].y;
    if (s == 123.456f) out[0] = s;
}

// MUFU.RSQ rate probe
__global__ void __launch_bounds__(256, 2) k_probe_rsq(float* out, int iters)
{
    float acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = 1.5f + threadIdx.x * 1e-3f + k;
    for (int i = 0; i < iters; ++i)
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] = nb::rsqrt_approx(acc[k]);
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += acc[k];
    if (s == 123.456f) out[0] = s;
}

int main(int argc, char** argv)
{
    const int n = argc > 1 ? atoi(argv[1]) : 262144;
    const int reps = argc > 2 ? atoi(argv[2]) : 3;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    printf("device %s sm_%d%d SMs %d clock %d MHz\n", prop.name, prop.major, prop.minor, prop.multiProcessorCount, prop.clockRate / 1000);

    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float* dscratch; CK(cudaMalloc(&dscratch, 64));
    {
        const int blocks = prop.multiProcessorCount * 8, iters = 8192;
        for (int v = 0; v < 2; ++v)
        {
            float best = 1e30f;
            for (int r = 0; r < 4; ++r)
            {
                CK(cudaEventRecord(e0));
                if (v == 0) k_probe<false><<<blocks, 256>>>(dscratch, iters, 0.999f, 1e-3f);
                else k_probe<true><<<blocks, 256>>>(dscratch, iters, 0.999f, 1e-3f);
                CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
                float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
                if (r > 0) best = std::min(best, ms);
            }
            const double fmas = (double)blocks * 256 * iters * 4.0 * 16.0;
            printf("probe %-6s: %.3f ms  %.2f TFLOP/s fp32\n", v ? "FFMA2" : "FFMA", best, 2.0 * fmas / (best * 1e-3) * 1e-12);
        }
        float best = 1e30f;
        for (int r = 0; r < 4; ++r)
        {
            CK(cudaEventRecord(e0));
            k_probe_rsq<<<blocks, 256>>>(dscratch, 2048);
            CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
            if (r > 0) best = std::min(best, ms);
        }
        printf("probe MUFU.RSQ: %.3f ms  %.2f Gop/s (%.1f per clk per SM at %d MHz)\n", best,
               (double)blocks * 256 * 2048 * 8 / (best * 1e-3) * 1e-9,
               (double)blocks * 256 * 2048 * 8 / (best * 1e-3) / prop.multiProcessorCount / (prop.clockRate * 1e3), prop.clockRate / 1000);
    }

    std::vector<float4> hp(n);
    srand(1234);
    for (int i = 0; i < n; ++i)
    {
        auto u = []() { return (float)rand() / (float)RAND_MAX; };
        hp[i] = make_float4(1400.f * (u() - 0.5f), 1400.f * (u() - 0.5f), 60.f * (u() - 0.5f), 6.674e-11f * (1e28f + u() * 1e30f));
    }
    float4* dp; CK(cudaMalloc(&dp, n * sizeof(float4)));
    CK(cudaMemcpy(dp, hp.data(), n * sizeof(float4), cudaMemcpyHostToDevice));
    float hwmax = 0.f;
    for (int i = 0; i < n; ++i) hwmax = std::max(hwmax, hp[i].w);
    float* dwmax; CK(cudaMalloc(&dwmax, sizeof(float)));
    CK(cudaMemcpy(dwmax, &hwmax, sizeof(float), cudaMemcpyHostToDevice));
    const int max_splits = 16;
    double* dout; CK(cudaMalloc(&dout, (size_t)max_splits * 3 * n * sizeof(double)));
    std::vector<double> ref(3 * (size_t)n), cur((size_t)max_splits * 3 * n);

    int count = 0;
    const nb::AllPairsKernel* table = nb::allpairs_table(&count);
    const float sc = 10.0f * nb::kPreScale;
    bool have_ref = false;
    for (int k = 0; k < count; ++k)
    {
        const nb::AllPairsKernel& K = table[k];
        int occ = 0;
        if (K.smem_bytes > 0) CK(cudaFuncSetAttribute((const void*)K.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, K.smem_bytes));
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, (const void*)K.fn, K.threads, K.smem_bytes));
        const long slots = (long)occ * prop.multiProcessorCount;
        const long tb = ((long)n + K.threads * K.targets - 1) / (K.threads * K.targets);
        const int tile = K.variant == 2 ? K.threads : 2 * K.threads;
        int splits = 1; double best_eff = -1;
        for (int s = 1; s <= max_splits; ++s)
        {
            if ((long)n / s < 4 * tile) break;
            const double waves = (double)(tb * s) / slots, eff = waves / std::ceil(waves);
            if (eff > best_eff + 0.005) { best_eff = eff; splits = s; }
        }
        long chunk = ((long)n + splits - 1) / splits; chunk = (chunk + tile - 1) / tile * tile;
        dim3 grid((unsigned)tb, (unsigned)splits);
        float best = 1e30f;
        for (int r = 0; r < reps + 1; ++r)
        {
            CK(cudaEventRecord(e0));
            K.fn<<<grid, K.threads, K.smem_bytes>>>(dp, n, 0, n, (int)chunk, dout, sc, 10.0f, dwmax);
            CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
            CK(cudaGetLastError());
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
            if (r > 0 || reps == 0) best = std::min(best, ms);
        }
        CK(cudaMemcpy(cur.data(), dout, (size_t)splits * 3 * n * sizeof(double), cudaMemcpyDeviceToHost));
        std::vector<double> sum(3 * (size_t)n, 0.0);
        for (int s = 0; s < splits; ++s)
            for (size_t i = 0; i < 3 * (size_t)n; ++i) sum[i] += cur[(size_t)s * 3 * n + i];
        double maxrel = 0;
        if (!have_ref) { ref = sum; have_ref = true; }
        else
            for (int i = 0; i < n; ++i)
            {
                double d2 = 0, r2 = 0;
                for (int c = 0; c < 3; ++c) { const double d = sum[(size_t)c * n + i] - ref[(size_t)c * n + i]; d2 += d * d; r2 += ref[(size_t)c * n + i] * ref[(size_t)c * n + i]; }
                maxrel = std::max(maxrel, std::sqrt(d2 / r2));
            }
        const double inter = (double)n * (double)n;
        printf("[%2d] %-36s occ %d splits %2d eff %.3f : %8.3f ms  %.3e int/s  %.2f TFLOP/s(20)  maxrel-vs-first %.2e\n", k, K.name, occ,
               splits, best_eff, best, inter / (best * 1e-3), inter * 20 / (best * 1e-3) * 1e-12, maxrel);
    }
    return 0;
}
