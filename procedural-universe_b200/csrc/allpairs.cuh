// K1 -- tiled all-pairs acceleration kernels for sm_100a.
//
// Replaces the hot loop of the reference's BruteForceCPU::Exec (src/Sim/BruteForceCPU.cpp:25-43)
// with Phys::Gravity (src/Sim/Physics.hpp:25-35) inlined:
//
//     F_i += -(G m_j m_i) / (d^2 + S) * normalize(p_i - p_j),   d^2 = |p_i - p_j|^2 in fp32
//
// The kernels accumulate the ACCELERATION a_i = F_i / m_i (G m_j m_i does not fit fp32):
//
//     a_i = sum_j  w_j * (p_j - p_i) / (|d| (d^2 + S)),     w_j = G m_j
//
// with ONE special-function op per interaction:  1 / (|d| (d^2+S)) = rsqrt(d^2 (d^2+S)^2).
// To keep d^2 (d^2+S)^2 inside fp32 for separations up to ~1.3e9 program units everything under
// the rsqrt is pre-scaled by 2^-54 (t' = (d^2+S) 2^-27 comes for free out of an FMA) and w_j is
// pre-multiplied by 2^-27 when the tile is staged.  A tiny epsilon added by the last FMA makes the
// self term and coincident bodies contribute exactly zero (dx = 0 times a finite number), which is
// what the reference gets from normalising a zero vector (SimpleMath Normalize, zero -> 0).
//
// 13 FMA-pipe operations + 1 MUFU per interaction.  Packed variants issue them as FFMA2 / FADD2 /
// FMUL2 (fma.rn.f32x2, new on sm_100) so the FMA pipe, not the issue slot, is the limiter.
//
// Accumulation: fp32 inside one source tile (<= 512 terms per accumulator), fp64 across tiles.
// Output: out[split][3][tgt_count] doubles, split = blockIdx.y (source-range split; partials are
// summed in fixed order by the kick-drift kernel, so results are run-to-run deterministic).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace nb
{

constexpr float kPreScale = 7.450580596923828e-09f;   // 2^-27
constexpr float kEps = 1.0e-37f;

__device__ __forceinline__ float rsqrt_approx(float x)
{
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// ------------------------------------------------------------------------------------------------
// Variant 0: scalar FFMA, sources as float4 {x,y,z,w'} in shared memory (LDS.128 broadcast).
// ------------------------------------------------------------------------------------------------
template <int THREADS, int T, int MINB>
__global__ void __launch_bounds__(THREADS, MINB)
k_allpairs_scalar(const float4* __restrict__ posw, int n, int tgt_first, int tgt_count, int src_chunk,
                  double* __restrict__ out, float sc, float soft, const float* __restrict__ wmax)
{
    constexpr int TILE = 2 * THREADS;
    __shared__ float4 sm[2][TILE];

    const int tid = threadIdx.x;
    const int j0 = blockIdx.y * src_chunk;
    const int j1 = min(n, j0 + src_chunk);
    const int base = blockIdx.x * (THREADS * T) + tid;

    float px[T], py[T], pz[T];
    double dax[T], day[T], daz[T];
#pragma unroll
    for (int t = 0; t < T; ++t)
    {
        const int li = base + t * THREADS;
        float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
        if (li < tgt_count) p = posw[tgt_first + li];
        px[t] = p.x; py[t] = p.y; pz[t] = p.z;
        dax[t] = day[t] = daz[t] = 0.0;
    }

    const int ntiles = (j1 - j0 + TILE - 1) / TILE;
    float4 pre[2];
    auto fetch = [&](int tile) {
#pragma unroll
        for (int k = 0; k < 2; ++k)
        {
            const int j = j0 + tile * TILE + 2 * tid + k;
            float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
            if (j < j1) s = posw[j];
            s.w *= kPreScale;
            pre[k] = s;
        }
    };
    auto stash = [&](int buf) {
        sm[buf][2 * tid] = pre[0];
        sm[buf][2 * tid + 1] = pre[1];
    };

    if (ntiles > 0) { fetch(0); stash(0); }
    __syncthreads();

    for (int tile = 0; tile < ntiles; ++tile)
    {
        const int buf = tile & 1;
        if (tile + 1 < ntiles) fetch(tile + 1);

        float ax[T], ay[T], az[T];
#pragma unroll
        for (int t = 0; t < T; ++t) ax[t] = ay[t] = az[t] = 0.f;

#pragma unroll 4
        for (int j = 0; j < TILE; ++j)
        {
            const float4 s = sm[buf][j];
#pragma unroll
            for (int t = 0; t < T; ++t)
            {
                const float dx = s.x - px[t];
                const float dy = s.y - py[t];
                const float dz = s.z - pz[t];
                float d2 = dx * dx;
                d2 = fmaf(dy, dy, d2);
                d2 = fmaf(dz, dz, d2);
                const float tt = fmaf(d2, kPreScale, sc);
                const float u = d2 * tt;
                const float x = fmaf(u, tt, kEps);
                const float r = rsqrt_approx(x);
                const float sw = s.w * r;
                ax[t] = fmaf(sw, dx, ax[t]);
                ay[t] = fmaf(sw, dy, ay[t]);
                az[t] = fmaf(sw, dz, az[t]);
            }
        }
#pragma unroll
        for (int t = 0; t < T; ++t)
        {
            dax[t] += (double)ax[t];
            day[t] += (double)ay[t];
            daz[t] += (double)az[t];
        }
        if (tile + 1 < ntiles) stash(buf ^ 1);
        __syncthreads();
    }

    double* o = out + (size_t)blockIdx.y * 3 * (size_t)tgt_count;
#pragma unroll
    for (int t = 0; t < T; ++t)
    {
        const int li = base + t * THREADS;
        if (li < tgt_count)
        {
            o[li] = dax[t];
            o[(size_t)tgt_count + li] = day[t];
            o[2 * (size_t)tgt_count + li] = daz[t];
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Variant 1: packed f32x2 over PAIRS OF SOURCES.  Shared memory holds, per source pair,
// A = {x0,x1,y0,y1} and B = {z0,z1,w0',w1'}; the target position is duplicated in a register
// pair once per kernel, so there is no per-interaction packing work.
// ------------------------------------------------------------------------------------------------
template <int THREADS, int T, int MINB>
__global__ void __launch_bounds__(THREADS, MINB)
k_allpairs_srcpair(const float4* __restrict__ posw, int n, int tgt_first, int tgt_count, int src_chunk,
                   double* __restrict__ out, float sc, float soft, const float* __restrict__ wmax)
{
    constexpr int TILE = 2 * THREADS;      // sources per tile
    constexpr int PAIRS = THREADS;         // source pairs per tile
    __shared__ float4 smA[2][PAIRS];
    __shared__ float4 smB[2][PAIRS];

    const int tid = threadIdx.x;
    const int j0 = blockIdx.y * src_chunk;
    const int j1 = min(n, j0 + src_chunk);
    const int base = blockIdx.x * (THREADS * T) + tid;

    float2 npx[T], npy[T], npz[T];         // {-p,-p}
    double dax[T], day[T], daz[T];
#pragma unroll
    for (int t = 0; t < T; ++t)
    {
        const int li = base + t * THREADS;
        float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
        if (li < tgt_count) p = posw[tgt_first + li];
        npx[t] = make_float2(-p.x, -p.x);
        npy[t] = make_float2(-p.y, -p.y);
        npz[t] = make_float2(-p.z, -p.z);
        dax[t] = day[t] = daz[t] = 0.0;
    }

    const float2 c2 = make_float2(kPreScale, kPreScale);
    const float2 sc2 = make_float2(sc, sc);
    const float2 eps2 = make_float2(kEps, kEps);

    const int ntiles = (j1 - j0 + TILE - 1) / TILE;
    float4 preA, preB;
    auto fetch = [&](int tile) {
        const int j = j0 + tile * TILE + 2 * tid;
        float4 s0 = make_float4(0.f, 0.f, 0.f, 0.f), s1 = s0;
        if (j < j1) s0 = posw[j];
        if (j + 1 < j1) s1 = posw[j + 1];
        preA = make_float4(s0.x, s1.x, s0.y, s1.y);
        preB = make_float4(s0.z, s1.z, s0.w * kPreScale, s1.w * kPreScale);
    };
    auto stash = [&](int buf) {
        smA[buf][tid] = preA;
        smB[buf][tid] = preB;
    };

    if (ntiles > 0) { fetch(0); stash(0); }
    __syncthreads();

    for (int tile = 0; tile < ntiles; ++tile)
    {
        const int buf = tile & 1;
        if (tile + 1 < ntiles) fetch(tile + 1);

        float2 ax[T], ay[T], az[T];
#pragma unroll
        for (int t = 0; t < T; ++t) ax[t] = ay[t] = az[t] = make_float2(0.f, 0.f);

#pragma unroll 4
        for (int jp = 0; jp < PAIRS; ++jp)
        {
            const float4 A = smA[buf][jp];
            const float4 B = smB[buf][jp];
            const float2 sx = make_float2(A.x, A.y), sy = make_float2(A.z, A.w);
            const float2 sz = make_float2(B.x, B.y), sw = make_float2(B.z, B.w);
#pragma unroll
            for (int t = 0; t < T; ++t)
            {
                const float2 dx = __fadd2_rn(sx, npx[t]);
                const float2 dy = __fadd2_rn(sy, npy[t]);
                const float2 dz = __fadd2_rn(sz, npz[t]);
                float2 d2 = __fmul2_rn(dx, dx);
                d2 = __ffma2_rn(dy, dy, d2);
                d2 = __ffma2_rn(dz, dz, d2);
                const float2 tt = __ffma2_rn(d2, c2, sc2);
                const float2 u = __fmul2_rn(d2, tt);
                const float2 x = __ffma2_rn(u, tt, eps2);
                const float2 r = make_float2(rsqrt_approx(x.x), rsqrt_approx(x.y));
                const float2 s = __fmul2_rn(sw, r);
                ax[t] = __ffma2_rn(s, dx, ax[t]);
                ay[t] = __ffma2_rn(s, dy, ay[t]);
                az[t] = __ffma2_rn(s, dz, az[t]);
            }
        }
#pragma unroll
        for (int t = 0; t < T; ++t)
        {
            dax[t] += (double)(ax[t].x + ax[t].y);
            day[t] += (double)(ay[t].x + ay[t].y);
            daz[t] += (double)(az[t].x + az[t].y);
        }
        if (tile + 1 < ntiles) stash(buf ^ 1);
        __syncthreads();
    }

    double* o = out + (size_t)blockIdx.y * 3 * (size_t)tgt_count;
#pragma unroll
    for (int t = 0; t < T; ++t)
    {
        const int li = base + t * THREADS;
        if (li < tgt_count)
        {
            o[li] = dax[t];
            o[(size_t)tgt_count + li] = day[t];
            o[2 * (size_t)tgt_count + li] = daz[t];
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Variant 2: packed f32x2 over PAIRS OF TARGETS.  Shared memory holds every source duplicated,
// A = {x,x,y,y}, B = {z,z,w',w'}; a thread owns T targets as T/2 register pairs.  Half the
// registers per target of variant 1 (more targets per thread), twice the LDS per interaction.
// ------------------------------------------------------------------------------------------------
template <int THREADS, int T, int MINB>
__global__ void __launch_bounds__(THREADS, MINB)
k_allpairs_tgtpair(const float4* __restrict__ posw, int n, int tgt_first, int tgt_count, int src_chunk,
                   double* __restrict__ out, float sc, float soft, const float* __restrict__ wmax)
{
    static_assert(T % 2 == 0, "targets are processed in pairs");
    constexpr int TP = T / 2;
    constexpr int TILE = THREADS;          // sources per tile (2 float4 each)
    __shared__ float4 smA[2][TILE];
    __shared__ float4 smB[2][TILE];

    const int tid = threadIdx.x;
    const int j0 = blockIdx.y * src_chunk;
    const int j1 = min(n, j0 + src_chunk);
    const int base = blockIdx.x * (THREADS * T) + tid;

    float2 npx[TP], npy[TP], npz[TP];
    double dax[T], day[T], daz[T];
#pragma unroll
    for (int p = 0; p < TP; ++p)
    {
        float4 q[2];
#pragma unroll
        for (int k = 0; k < 2; ++k)
        {
            const int li = base + (2 * p + k) * THREADS;
            q[k] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (li < tgt_count) q[k] = posw[tgt_first + li];
            dax[2 * p + k] = day[2 * p + k] = daz[2 * p + k] = 0.0;
        }
        npx[p] = make_float2(-q[0].x, -q[1].x);
        npy[p] = make_float2(-q[0].y, -q[1].y);
        npz[p] = make_float2(-q[0].z, -q[1].z);
    }

    const float2 c2 = make_float2(kPreScale, kPreScale);
    const float2 sc2 = make_float2(sc, sc);
    const float2 eps2 = make_float2(kEps, kEps);

    const int ntiles = (j1 - j0 + TILE - 1) / TILE;
    float4 pre;
    auto fetch = [&](int tile) {
        const int j = j0 + tile * TILE + tid;
        pre = make_float4(0.f, 0.f, 0.f, 0.f);
        if (j < j1) pre = posw[j];
        pre.w *= kPreScale;
    };
    auto stash = [&](int buf) {
        smA[buf][tid] = make_float4(pre.x, pre.x, pre.y, pre.y);
        smB[buf][tid] = make_float4(pre.z, pre.z, pre.w, pre.w);
    };

    if (ntiles > 0) { fetch(0); stash(0); }
    __syncthreads();

    for (int tile = 0; tile < ntiles; ++tile)
    {
        const int buf = tile & 1;
        if (tile + 1 < ntiles) fetch(tile + 1);

        float2 ax[TP], ay[TP], az[TP];
#pragma unroll
        for (int p = 0; p < TP; ++p) ax[p] = ay[p] = az[p] = make_float2(0.f, 0.f);

#pragma unroll 4
        for (int j = 0; j < TILE; ++j)
        {
            const float4 A = smA[buf][j];
            const float4 B = smB[buf][j];
            const float2 sx = make_float2(A.x, A.y), sy = make_float2(A.z, A.w);
            const float2 sz = make_float2(B.x, B.y), sw = make_float2(B.z, B.w);
#pragma unroll
            for (int p = 0; p < TP; ++p)
            {
                const float2 dx = __fadd2_rn(sx, npx[p]);
                const float2 dy = __fadd2_rn(sy, npy[p]);
                const float2 dz = __fadd2_rn(sz, npz[p]);
                float2 d2 = __fmul2_rn(dx, dx);
                d2 = __ffma2_rn(dy, dy, d2);
                d2 = __ffma2_rn(dz, dz, d2);
                const float2 tt = __ffma2_rn(d2, c2, sc2);
                const float2 u = __fmul2_rn(d2, tt);
                const float2 x = __ffma2_rn(u, tt, eps2);
                const float2 r = make_float2(rsqrt_approx(x.x), rsqrt_approx(x.y));
                const float2 s = __fmul2_rn(sw, r);
                ax[p] = __ffma2_rn(s, dx, ax[p]);
                ay[p] = __ffma2_rn(s, dy, ay[p]);
                az[p] = __ffma2_rn(s, dz, az[p]);
            }
        }
#pragma unroll
        for (int p = 0; p < TP; ++p)
        {
            dax[2 * p] += (double)ax[p].x; dax[2 * p + 1] += (double)ax[p].y;
            day[2 * p] += (double)ay[p].x; day[2 * p + 1] += (double)ay[p].y;
            daz[2 * p] += (double)az[p].x; daz[2 * p + 1] += (double)az[p].y;
        }
        if (tile + 1 < ntiles) stash(buf ^ 1);
        __syncthreads();
    }

    double* o = out + (size_t)blockIdx.y * 3 * (size_t)tgt_count;
#pragma unroll
    for (int t = 0; t < T; ++t)
    {
        const int li = base + t * THREADS;
        if (li < tgt_count)
        {
            o[li] = dax[t];
            o[(size_t)tgt_count + li] = day[t];
            o[2 * (size_t)tgt_count + li] = daz[t];
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Variant 3: variant 1 with the source weight folded under the rsqrt -- 12 FMA-pipe ops + 1 MUFU.
//
//     w / (|d| (d^2+S)) = rsqrt( d^2 * (a (d^2+S))^2 ),   a = 1 / w
//
// so the multiply by w disappears if the tile holds a and b = S a per source:
//     t = fma(d2, a, b);  u = d2 * t;  x = fma(u, t, eps);  s = rsqrt(x);  acc += s * d.
// To stay inside fp32 the weights are normalised by W = 2^ceil(log2(max_j w_j)) (exact scaling,
// undone exactly in the epilogue) and a carries the same 2^-27 pre-scale as the other variants:
// a = 2^-27 W / w >= 2^-27.  Light bodies have a large `a`; their term overflows to rsqrt(inf) = 0
// only where it is < 1e-20 of a unit-weight body's at the same distance.  Massless bodies
// (a = inf) are clamped to a = 1e18 so that their own self term stays finite (0 * finite).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float pow2_ceil(float m)
{
    // smallest power of two >= m for normal m > 0 (exact powers of two map to themselves * 2; harmless)
    return __int_as_float((__float_as_int(m) & 0x7f800000) + 0x00800000);
}

template <int THREADS, int T, int MINB>
__global__ void __launch_bounds__(THREADS, MINB)
k_allpairs_fold(const float4* __restrict__ posw, int n, int tgt_first, int tgt_count, int src_chunk,
                double* __restrict__ out, float sc, float soft, const float* __restrict__ wmax)
{
    constexpr int TILE = 2 * THREADS;
    constexpr int PAIRS = THREADS;
    __shared__ float4 smA[2][PAIRS];   // {x0,x1,y0,y1}
    __shared__ float4 smB[2][PAIRS];   // {z0,z1,a0,a1}
    __shared__ float2 smC[2][PAIRS];   // {b0,b1}

    const int tid = threadIdx.x;
    const int j0 = blockIdx.y * src_chunk;
    const int j1 = min(n, j0 + src_chunk);
    const int base = blockIdx.x * (THREADS * T) + tid;
    const float W = pow2_ceil(fmaxf(*wmax, 1.17549435e-38f));
    const float cW = kPreScale * W;

    float2 npx[T], npy[T], npz[T];
    double dax[T], day[T], daz[T];
#pragma unroll
    for (int t = 0; t < T; ++t)
    {
        const int li = base + t * THREADS;
        float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
        if (li < tgt_count) p = posw[tgt_first + li];
        npx[t] = make_float2(-p.x, -p.x);
        npy[t] = make_float2(-p.y, -p.y);
        npz[t] = make_float2(-p.z, -p.z);
        dax[t] = day[t] = daz[t] = 0.0;
    }
    const float2 eps2 = make_float2(kEps, kEps);

    const int ntiles = (j1 - j0 + TILE - 1) / TILE;
    float4 preA, preB;
    float2 preC;
    auto fetch = [&](int tile) {
        const int j = j0 + tile * TILE + 2 * tid;
        float4 s0 = make_float4(0.f, 0.f, 0.f, 0.f), s1 = s0;
        if (j < j1) s0 = posw[j];
        if (j + 1 < j1) s1 = posw[j + 1];
        const float a0 = fminf(__fdiv_rn(cW, s0.w), 1e18f);   // w = 0 (padding, massless) -> clamp
        const float a1 = fminf(__fdiv_rn(cW, s1.w), 1e18f);
        preA = make_float4(s0.x, s1.x, s0.y, s1.y);
        preB = make_float4(s0.z, s1.z, a0, a1);
        preC = make_float2(soft * a0, soft * a1);
    };
    auto stash = [&](int buf) {
        smA[buf][tid] = preA;
        smB[buf][tid] = preB;
        smC[buf][tid] = preC;
    };

    if (ntiles > 0) { fetch(0); stash(0); }
    __syncthreads();

    for (int tile = 0; tile < ntiles; ++tile)
    {
        const int buf = tile & 1;
        if (tile + 1 < ntiles) fetch(tile + 1);

        float2 ax[T], ay[T], az[T];
#pragma unroll
        for (int t = 0; t < T; ++t) ax[t] = ay[t] = az[t] = make_float2(0.f, 0.f);

#pragma unroll 4
        for (int jp = 0; jp < PAIRS; ++jp)
        {
            const float4 A = smA[buf][jp];
            const float4 B = smB[buf][jp];
            const float2 sb = smC[buf][jp];
            const float2 sx = make_float2(A.x, A.y), sy = make_float2(A.z, A.w);
            const float2 sz = make_float2(B.x, B.y), sa = make_float2(B.z, B.w);
#pragma unroll
            for (int t = 0; t < T; ++t)
            {
                const float2 dx = __fadd2_rn(sx, npx[t]);
                const float2 dy = __fadd2_rn(sy, npy[t]);
                const float2 dz = __fadd2_rn(sz, npz[t]);
                float2 d2 = __fmul2_rn(dx, dx);
                d2 = __ffma2_rn(dy, dy, d2);
                d2 = __ffma2_rn(dz, dz, d2);
                const float2 tt = __ffma2_rn(d2, sa, sb);
                const float2 u = __fmul2_rn(d2, tt);
                const float2 x = __ffma2_rn(u, tt, eps2);
                const float2 s = make_float2(rsqrt_approx(x.x), rsqrt_approx(x.y));
                ax[t] = __ffma2_rn(s, dx, ax[t]);
                ay[t] = __ffma2_rn(s, dy, ay[t]);
                az[t] = __ffma2_rn(s, dz, az[t]);
            }
        }
#pragma unroll
        for (int t = 0; t < T; ++t)
        {
            dax[t] += (double)(ax[t].x + ax[t].y);
            day[t] += (double)(ay[t].x + ay[t].y);
            daz[t] += (double)(az[t].x + az[t].y);
        }
        if (tile + 1 < ntiles) stash(buf ^ 1);
        __syncthreads();
    }

    double* o = out + (size_t)blockIdx.y * 3 * (size_t)tgt_count;
    const double Wd = (double)W * (double)kPreScale;   // rsqrt(x) = (w/W) 2^27 / (|d| (d^2+S))
#pragma unroll
    for (int t = 0; t < T; ++t)
    {
        const int li = base + t * THREADS;
        if (li < tgt_count)
        {
            o[li] = dax[t] * Wd;
            o[(size_t)tgt_count + li] = day[t] * Wd;
            o[2 * (size_t)tgt_count + li] = daz[t] * Wd;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Variant 4: variant 3 tuned for the issue port.  ncu on variants 1/3 shows the FMA pipe ~85 %
// busy with `math pipe throttle` the only stall: on sm_100 every packed f32x2 op and every MUFU
// holds the issue port for 2 cycles and every other instruction for 1, so the floor per
// interaction is 12 (FMA) + 2 (MUFU) + whatever else the loop issues.  This variant trims the
// "whatever else": the fp64 running sums live in shared memory (touched once per 512-source tile)
// instead of 6 T registers, which lets a thread own more targets (fewer LDS per interaction)
// inside the 128-register budget, and b = S a of two source pairs shares one LDS.128.
// ------------------------------------------------------------------------------------------------
// MIX selects, per group of operations, packed f32x2 (bit clear) or two scalar instructions (bit set):
// bit 0 the three subtractions, bit 1 the d^2 chain, bit 2 t / u / x, bit 3 the three accumulations.  Measured
// on B200 (tools/issue_probe.cu): an FFMA2 costs ~2.2 issue cycles, two scalar FFMA 2.0 -- packed math saves
// issue slots, not pipe time, so the best mix depends on what else competes for the issue port.
template <int THREADS, int T, int MINB, int UNROLL, int MIX = 0>
__global__ void __launch_bounds__(THREADS, MINB)
k_allpairs_fold2(const float4* __restrict__ posw, int n, int tgt_first, int tgt_count, int src_chunk,
                 double* __restrict__ out, float sc, float soft, const float* __restrict__ wmax)
{
    constexpr int TILE = 2 * THREADS;
    constexpr int PAIRS = THREADS;
    extern __shared__ __align__(16) unsigned char dyn_smem[];
    float4 (*smA)[PAIRS] = reinterpret_cast<float4 (*)[PAIRS]>(dyn_smem);                        // {x0,x1,y0,y1}
    float4 (*smB)[PAIRS] = reinterpret_cast<float4 (*)[PAIRS]>(dyn_smem + 2 * PAIRS * 16);       // {z0,z1,a0,a1}
    float4 (*smC)[PAIRS / 2] = reinterpret_cast<float4 (*)[PAIRS / 2]>(dyn_smem + 4 * PAIRS * 16);   // {b of pair 2k, b of pair 2k+1}
    double (*sacc)[THREADS] = reinterpret_cast<double (*)[THREADS]>(dyn_smem + 5 * PAIRS * 16);  // [3T][THREADS]

    const int tid = threadIdx.x;
    const int j0 = blockIdx.y * src_chunk;
    const int j1 = min(n, j0 + src_chunk);
    const int base = blockIdx.x * (THREADS * T) + tid;
    const float W = pow2_ceil(fmaxf(*wmax, 1.17549435e-38f));
    const float cW = kPreScale * W;

    float2 npx[T], npy[T], npz[T];
#pragma unroll
    for (int t = 0; t < T; ++t)
    {
        const int li = base + t * THREADS;
        float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
        if (li < tgt_count) p = posw[tgt_first + li];
        npx[t] = make_float2(-p.x, -p.x);
        npy[t] = make_float2(-p.y, -p.y);
        npz[t] = make_float2(-p.z, -p.z);
#pragma unroll
        for (int c = 0; c < 3; ++c) sacc[3 * t + c][tid] = 0.0;
    }
    const float2 eps2 = make_float2(kEps, kEps);

    const int ntiles = (j1 - j0 + TILE - 1) / TILE;
    float4 preA, preB;
    float2 preC;
    auto fetch = [&](int tile) {
        const int j = j0 + tile * TILE + 2 * tid;
        float4 s0 = make_float4(0.f, 0.f, 0.f, 0.f), s1 = s0;
        if (j < j1) s0 = posw[j];
        if (j + 1 < j1) s1 = posw[j + 1];
        const float a0 = fminf(__fdiv_rn(cW, s0.w), 1e18f);
        const float a1 = fminf(__fdiv_rn(cW, s1.w), 1e18f);
        preA = make_float4(s0.x, s1.x, s0.y, s1.y);
        preB = make_float4(s0.z, s1.z, a0, a1);
        preC = make_float2(soft * a0, soft * a1);
    };
    auto stash = [&](int buf) {
        smA[buf][tid] = preA;
        smB[buf][tid] = preB;
        reinterpret_cast<float2*>(&smC[buf][0])[tid] = preC;
    };

    if (ntiles > 0) { fetch(0); stash(0); }
    __syncthreads();

    for (int tile = 0; tile < ntiles; ++tile)
    {
        const int buf = tile & 1;
        if (tile + 1 < ntiles) fetch(tile + 1);

        float2 ax[T], ay[T], az[T];
#pragma unroll
        for (int t = 0; t < T; ++t) ax[t] = ay[t] = az[t] = make_float2(0.f, 0.f);

#pragma unroll UNROLL
        for (int jq = 0; jq < PAIRS / 2; ++jq)
        {
            const float4 C = smC[buf][jq];
#pragma unroll
            for (int h = 0; h < 2; ++h)
            {
                const float4 A = smA[buf][2 * jq + h];
                const float4 B = smB[buf][2 * jq + h];
                const float2 sx = make_float2(A.x, A.y), sy = make_float2(A.z, A.w);
                const float2 sz = make_float2(B.x, B.y), sa = make_float2(B.z, B.w);
                const float2 sb = h == 0 ? make_float2(C.x, C.y) : make_float2(C.z, C.w);
#pragma unroll
                for (int t = 0; t < T; ++t)
                {
                    float2 dx, dy, dz, d2, tt, u, x;
                    if (MIX & 1)
                    {
                        dx = make_float2(sx.x + npx[t].x, sx.y + npx[t].y);
                        dy = make_float2(sy.x + npy[t].x, sy.y + npy[t].y);
                        dz = make_float2(sz.x + npz[t].x, sz.y + npz[t].y);
                    }
                    else
                    {
                        dx = __fadd2_rn(sx, npx[t]);
                        dy = __fadd2_rn(sy, npy[t]);
                        dz = __fadd2_rn(sz, npz[t]);
                    }
                    if (MIX & 2)
                    {
                        d2 = make_float2(dx.x * dx.x, dx.y * dx.y);
                        d2 = make_float2(fmaf(dy.x, dy.x, d2.x), fmaf(dy.y, dy.y, d2.y));
                        d2 = make_float2(fmaf(dz.x, dz.x, d2.x), fmaf(dz.y, dz.y, d2.y));
                    }
                    else
                    {
                        d2 = __fmul2_rn(dx, dx);
                        d2 = __ffma2_rn(dy, dy, d2);
                        d2 = __ffma2_rn(dz, dz, d2);
                    }
                    if (MIX & 4)
                    {
                        tt = make_float2(fmaf(d2.x, sa.x, sb.x), fmaf(d2.y, sa.y, sb.y));
                        u = make_float2(d2.x * tt.x, d2.y * tt.y);
                        x = make_float2(fmaf(u.x, tt.x, kEps), fmaf(u.y, tt.y, kEps));
                    }
                    else
                    {
                        tt = __ffma2_rn(d2, sa, sb);
                        u = __fmul2_rn(d2, tt);
                        x = __ffma2_rn(u, tt, eps2);
                    }
                    const float2 s = make_float2(rsqrt_approx(x.x), rsqrt_approx(x.y));
                    if (MIX & 8)
                    {
                        ax[t] = make_float2(fmaf(s.x, dx.x, ax[t].x), fmaf(s.y, dx.y, ax[t].y));
                        ay[t] = make_float2(fmaf(s.x, dy.x, ay[t].x), fmaf(s.y, dy.y, ay[t].y));
                        az[t] = make_float2(fmaf(s.x, dz.x, az[t].x), fmaf(s.y, dz.y, az[t].y));
                    }
                    else
                    {
                        ax[t] = __ffma2_rn(s, dx, ax[t]);
                        ay[t] = __ffma2_rn(s, dy, ay[t]);
                        az[t] = __ffma2_rn(s, dz, az[t]);
                    }
                }
            }
        }
#pragma unroll
        for (int t = 0; t < T; ++t)
        {
            sacc[3 * t + 0][tid] += (double)(ax[t].x + ax[t].y);
            sacc[3 * t + 1][tid] += (double)(ay[t].x + ay[t].y);
            sacc[3 * t + 2][tid] += (double)(az[t].x + az[t].y);
        }
        if (tile + 1 < ntiles) stash(buf ^ 1);
        __syncthreads();
    }

    double* o = out + (size_t)blockIdx.y * 3 * (size_t)tgt_count;
    const double Wd = (double)W * (double)kPreScale;
#pragma unroll
    for (int t = 0; t < T; ++t)
    {
        const int li = base + t * THREADS;
        if (li < tgt_count)
        {
            o[li] = sacc[3 * t + 0][tid] * Wd;
            o[(size_t)tgt_count + li] = sacc[3 * t + 1][tid] * Wd;
            o[2 * (size_t)tgt_count + li] = sacc[3 * t + 2][tid] * Wd;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Launch table shared by the library and the tuner.
// ------------------------------------------------------------------------------------------------
struct AllPairsKernel
{
    const char* name;
    int variant;       // 0 scalar, 1 source pairs, 2 target pairs, 3 source pairs with folded weight
    int threads;
    int targets;       // targets per thread
    int min_blocks;    // __launch_bounds__ min blocks per SM
    int smem_bytes;    // dynamic shared memory per CTA (0: the kernel only uses static shared memory)
    void (*fn)(const float4*, int, int, int, int, double*, float, float, const float*);
};

#define NB_AP_ENTRY(KERNEL, VAR, TH, T, MB) \
    { #KERNEL "<" #TH "," #T "," #MB ">", VAR, TH, T, MB, 0, KERNEL<TH, T, MB> }
#define NB_AP_ENTRY4(TH, T, MB, UN) \
    { "k_allpairs_fold2<" #TH "," #T "," #MB "," #UN ">", 3, TH, T, MB, 5 * TH * 16 + 3 * T * TH * 8, k_allpairs_fold2<TH, T, MB, UN> }
#define NB_AP_ENTRY5(TH, T, MB, UN, MIX) \
    { "k_allpairs_fold2<" #TH "," #T "," #MB "," #UN ",mix" #MIX ">", 3, TH, T, MB, 5 * TH * 16 + 3 * T * TH * 8, k_allpairs_fold2<TH, T, MB, UN, MIX> }

inline const AllPairsKernel* allpairs_table(int* count)
{
    static const AllPairsKernel table[] = {
        // index 0 = library default (fastest at 1e-5 parity).  The round-1 tuning sweep covered 38 entries
        // (profiles/r1_tune_sweep_a.txt); what stays is one representative per kernel family and the neighbours
        // of the default in (threads, targets per thread, blocks per SM, unroll), all kept under test.
        NB_AP_ENTRY4(256, 4, 2, 2),                      // 0
        NB_AP_ENTRY(k_allpairs_scalar, 0, 256, 4, 2),    // 1: scalar FFMA, weight multiplied in: 56 % of peak
        NB_AP_ENTRY(k_allpairs_srcpair, 1, 256, 4, 2),   // 2: packed over source pairs, 13 ops: 65.7 %
        NB_AP_ENTRY(k_allpairs_tgtpair, 2, 256, 4, 2),   // 3: packed over target pairs
        NB_AP_ENTRY(k_allpairs_fold, 3, 256, 4, 2),      // 4: weight folded under the rsqrt: 68.6 %
        NB_AP_ENTRY4(256, 2, 2, 2),                      // 5
        NB_AP_ENTRY4(256, 4, 2, 1),                      // 6
        NB_AP_ENTRY4(128, 4, 4, 2),                      // 7
        NB_AP_ENTRY4(256, 6, 1, 1),                      // 8
        NB_AP_ENTRY4(256, 8, 1, 1),                      // 9
        NB_AP_ENTRY5(256, 4, 2, 2, 8),                   // 10: scalar accumulations (50.2 TFLOP/s against 51.0 packed at 256 K bodies)
        NB_AP_ENTRY5(256, 4, 2, 2, 15),                  // 11: all scalar (45.9): every mix of scalar and packed groups is slower than all packed
    };
    *count = (int)(sizeof(table) / sizeof(table[0]));
    return table;
}

}  // namespace nb
