// Issue-port probe for sm_100a: can ALU / MUFU / LDS instructions issue in the shadow of packed f32x2
// FMA instructions?  Each variant runs a long unrolled loop of independent instruction streams on every
// SM (8 warps per SMSP) and reports SMSP cycles per loop iteration.  Used to size the Barnes-Hut walk
// (DESIGN.md K7): the walk is issue-bound, so what matters is whether its ~15 bookkeeping instructions per
// node visit cost issue cycles of their own once the FP work is packed.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/issue_probe tools/issue_probe.cu && tools/issue_probe
#include <cstdio>
#include <cuda_runtime.h>

#define FMA2(acc, a, b) asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(a), "l"(b))
#define FMA1(acc, a, b) asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(acc) : "f"(a), "f"(b))
#define ALU(x, y) asm volatile("lop3.b32 %0, %0, %1, %0, 0x96;" : "+r"(x) : "r"(y))
#define IAD(x, y) asm volatile("add.u32 %0, %0, %1;" : "+r"(x) : "r"(y))
#define MUFU(x) asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(x))
#define SETP(p, x, y) asm volatile("{ .reg .pred q; setp.gt.f32 q, %1, %2; selp.u32 %0, 1, %0, q; }" : "+r"(p) : "f"(x), "f"(y))

template <int NF2, int NF1, int NALU, int NMUFU>
__global__ void __launch_bounds__(1024) k_probe(int iters, float seed, unsigned long long* out, long long* cycles)
{
    unsigned long long acc2[8];
    float acc1[8], m[4];
    unsigned int ia[8];
    const unsigned long long a = ((unsigned long long)__float_as_uint(seed) << 32) | __float_as_uint(seed * 0.5f);
    const unsigned long long b = ((unsigned long long)__float_as_uint(1.0f - seed) << 32) | __float_as_uint(seed);
#pragma unroll
    for (int k = 0; k < 8; ++k) { acc2[k] = a + k; acc1[k] = seed * k; ia[k] = threadIdx.x + k; }
#pragma unroll
    for (int k = 0; k < 4; ++k) m[k] = seed + k + 1.0f;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it)
    {
#pragma unroll
        for (int r = 0; r < 4; ++r)
        {
            // interleave the streams the way a compiler would schedule them
#pragma unroll
            for (int k = 0; k < 8; ++k)
            {
                if (k < NF2) FMA2(acc2[k], a, b);
                if (k < NF1) FMA1(acc1[k], seed, seed);
                if (k < NALU) ALU(ia[k], ia[(k + 1) & 7]);
                if (k < NMUFU) MUFU(m[k & 3]);
            }
        }
    }
    const long long t1 = clock64();
    unsigned long long s = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += acc2[k] + __float_as_uint(acc1[k]) + ia[k];
#pragma unroll
    for (int k = 0; k < 4; ++k) s += __float_as_uint(m[k]);
    if (s == 0x1234567ull) out[0] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
}

template <int NF2, int NF1, int NALU, int NMUFU>
static void run(const char* name)
{
    unsigned long long* out;
    long long* cyc;
    cudaMalloc(&out, 8);
    cudaMalloc(&cyc, 8);
    const int iters = 4096;
    k_probe<NF2, NF1, NALU, NMUFU><<<148, 1024>>>(16, 0.3f, out, cyc);
    k_probe<NF2, NF1, NALU, NMUFU><<<148, 1024>>>(iters, 0.3f, out, cyc);
    long long c = 0;
    cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    // 1024 threads = 32 warps per SM = 8 per SMSP; per iteration each warp issues 4 x (NF2 + NF1 + NALU + NMUFU) instructions
    const double per_iter = (double)c / iters / 4.0 / 8.0;      // SMSP cycles per (NF2 + NF1 + NALU + NMUFU)-instruction group
    printf("%-44s f32x2 %d  f32 %d  alu %d  mufu %d : %6.2f SMSP cycles per group (%d instructions)\n", name, NF2, NF1, NALU, NMUFU, per_iter,
           NF2 + NF1 + NALU + NMUFU);
    cudaFree(out);
    cudaFree(cyc);
}

int main()
{
    run<8, 0, 0, 0>("packed FMA only");
    run<0, 8, 0, 0>("scalar FMA only");
    run<0, 0, 8, 0>("ALU only");
    run<0, 0, 0, 4>("MUFU only");
    run<8, 0, 4, 0>("packed FMA + 4 ALU");
    run<8, 0, 8, 0>("packed FMA + 8 ALU");
    run<0, 8, 4, 0>("scalar FMA + 4 ALU");
    run<0, 8, 8, 0>("scalar FMA + 8 ALU");
    run<8, 0, 0, 2>("packed FMA + 2 MUFU");
    run<8, 0, 0, 4>("packed FMA + 4 MUFU");
    run<8, 0, 8, 2>("packed FMA + 8 ALU + 2 MUFU");
    run<6, 4, 8, 2>("6 packed + 4 scalar FMA + 8 ALU + 2 MUFU");
    run<4, 8, 8, 2>("4 packed + 8 scalar FMA + 8 ALU + 2 MUFU");
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
