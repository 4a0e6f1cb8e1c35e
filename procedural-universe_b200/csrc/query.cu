// Nearest-body query on device-resident positions.
//
// What is being replaced: Maths::ClosestParticle (reference src/Core/Maths.hpp:62-85), an O(N)
// scan over the caller's std::vector<Particle> that the sandbox runs on the UI thread; it is the
// one Sim-adjacent function the reference's own tests pin (test/MathsTests.cpp:4-33).
// Semantics kept exactly:
//   * d = Vector3::DistanceSquared(pos, p.Position) in fp32, ((dx*dx + dy*dy) + dz*dz), no FMA;
//   * scan in index order with a strict `d < best`, best starting at FLT_MAX: the FIRST index among
//     equal minima wins, distances that are FLT_MAX, +inf or NaN never win, and with no winner the
//     reported id stays 0.
// How: HBM-bound single pass over the float4 position array (16 B/body).  Every candidate becomes
// the 64-bit key (bits(d) << 32 | index); d >= 0, so the unsigned order of the keys is "smaller
// distance first, then smaller index" and min() over keys in ANY order equals the serial scan.
#include <algorithm>
#include <cfloat>
#include <cstring>

#include "nb_internal.h"

namespace nb
{

__device__ __forceinline__ unsigned long long closest_key(float4 p, float qx, float qy, float qz, unsigned int i)
{
    const float dx = __fsub_rn(qx, p.x), dy = __fsub_rn(qy, p.y), dz = __fsub_rn(qz, p.z);
    const float d = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
    // strict d < FLT_MAX also rejects +inf and NaN (comparison false)
    if (!(d < FLT_MAX)) return 0xFFFFFFFFFFFFFFFFull;
    return ((unsigned long long)__float_as_uint(d) << 32) | i;
}

__global__ void __launch_bounds__(256)
k_closest(const float4* __restrict__ posw, unsigned int n, float qx, float qy, float qz, unsigned long long* __restrict__ best)
{
    __shared__ unsigned long long red[8];
    unsigned long long k = 0xFFFFFFFFFFFFFFFFull;
    const size_t stride = (size_t)gridDim.x * 256u;
    size_t i = (size_t)blockIdx.x * 256u + threadIdx.x;
    // two independent loads in flight per thread
    for (; i + stride < n; i += 2 * stride)
    {
        const float4 a = posw[i], b = posw[i + stride];
        k = min(k, min(closest_key(a, qx, qy, qz, (unsigned int)i), closest_key(b, qx, qy, qz, (unsigned int)(i + stride))));
    }
    if (i < n) k = min(k, closest_key(posw[i], qx, qy, qz, (unsigned int)i));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) k = min(k, __shfl_xor_sync(0xffffffffu, k, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = k;
    __syncthreads();
    if (threadIdx.x == 0)
    {
#pragma unroll
        for (int w = 1; w < 8; ++w) k = min(k, red[w]);
        atomicMin(best, k);
    }
}

}  // namespace nb

using namespace nb;

extern "C" int nb_closest_particle(nb_handle h, const float pos[3], size_t* index, float* dist_sq)
{
    NB_REQUIRE(h != nullptr && pos != nullptr && index != nullptr, NB_ERR_ARG, "null argument");
    NB_REQUIRE(h->n > 0, NB_ERR_STATE, "not initialised (the reference dereferences an unset pointer on an empty vector)");
    NB_REQUIRE(h->exchanged, NB_ERR_STATE, "positions of remote ranks are stale");
    NB_CUDA(cudaSetDevice(h->cfg.device));
    if (h->p2p_attached) NB_CHECK(p2p_wait(h));
    unsigned long long* d_best = nullptr;
    NB_CUDA(cudaMalloc(&d_best, sizeof(unsigned long long)));
    // no winner -> id 0 like the reference (size_t id = 0; float distance = FLT_MAX)
    const float fmax = FLT_MAX;
    unsigned int fbits;
    memcpy(&fbits, &fmax, sizeof(fbits));
    const unsigned long long init = (unsigned long long)fbits << 32;
    cudaError_t e = cudaMemcpyAsync(d_best, &init, sizeof(init), cudaMemcpyHostToDevice, h->stream);
    if (e == cudaSuccess)
    {
        const size_t want = (h->n + 511) / 512;
        const int blocks = (int)std::min<size_t>(want, (size_t)h->sm_count * 8);
        k_closest<<<blocks > 0 ? blocks : 1, 256, 0, h->stream>>>(h->posw, (unsigned int)h->n, pos[0], pos[1], pos[2], d_best);
        e = cudaGetLastError();
    }
    unsigned long long best = init;
    if (e == cudaSuccess) e = cudaMemcpyAsync(&best, d_best, sizeof(best), cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    cudaFree(d_best);
    NB_CUDA(e);
    ++h->total_launches;
    *index = (size_t)(best & 0xFFFFFFFFull);
    if (dist_sq)
    {
        const unsigned int bits = (unsigned int)(best >> 32);
        memcpy(dist_sq, &bits, sizeof(bits));
    }
    return NB_OK;
}
