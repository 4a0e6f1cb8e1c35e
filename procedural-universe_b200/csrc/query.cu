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
#include <new>

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

// ------------------------------------------------------------------------------------------------
// Cross-check kernels (parity hooks, not on the step path).
//
// k_direct_accel: the reference's all-pairs law restated operation by operation for a handful of
// targets against ALL sources (BruteForceCPU::Exec, BruteForceCPU.cpp:25-43 + Phys::Gravity,
// Physics.hpp:25-35): diff = p_i - p_j and Normalize in fp32 without fusion, d^2 = ((dx^2 + dy^2) + dz^2)
// in fp32, f = -(G m_j) / ((double)d^2 + S) and the accumulation in fp64.  Usable at any N (64 targets x
// 64 M sources is 4e9 pair evaluations), so the fast all-pairs kernel and the tree walk can be checked on
// the device at sizes where no CPU reference finishes.  The only difference from the reference is the
// weight: posw.w carries G m_j rounded to fp32 (6e-8 relative).
// ------------------------------------------------------------------------------------------------
constexpr int DA_GROUP = 8;

__global__ void __launch_bounds__(256)
k_direct_accel(const float4* __restrict__ posw, int n, const unsigned int* __restrict__ bodies, int nt, int chunk, double S,
               double* __restrict__ out /* [nt][gridDim.x][3] */)
{
    __shared__ double red[DA_GROUP][3][8];
    const int g0 = blockIdx.y * DA_GROUP;
    float4 me[DA_GROUP];
    unsigned int idx[DA_GROUP];
#pragma unroll
    for (int k = 0; k < DA_GROUP; ++k)
    {
        idx[k] = bodies[min(g0 + k, nt - 1)];
        me[k] = posw[idx[k]];
    }
    double ax[DA_GROUP], ay[DA_GROUP], az[DA_GROUP];
#pragma unroll
    for (int k = 0; k < DA_GROUP; ++k) ax[k] = ay[k] = az[k] = 0.0;
    const int j0 = blockIdx.x * chunk, j1 = min(n, j0 + chunk);
    for (int j = j0 + threadIdx.x; j < j1; j += 256)
    {
        const float4 s = posw[j];
#pragma unroll
        for (int k = 0; k < DA_GROUP; ++k)
        {
            if ((unsigned int)j == idx[k]) continue;
            const float dx = __fsub_rn(me[k].x, s.x), dy = __fsub_rn(me[k].y, s.y), dz = __fsub_rn(me[k].z, s.z);
            const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
            const float len = __fsqrt_rn(d2);
            float ux = 0.f, uy = 0.f, uz = 0.f;            // XMVector3Normalize: zero vector stays zero
            if (len > 0.f) { ux = __fdiv_rn(dx, len); uy = __fdiv_rn(dy, len); uz = __fdiv_rn(dz, len); }
            const double f = -(double)s.w / ((double)d2 + S);
            ax[k] += f * (double)ux; ay[k] += f * (double)uy; az[k] += f * (double)uz;
        }
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int k = 0; k < DA_GROUP; ++k)
    {
        double x = ax[k], y = ay[k], z = az[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
        {
            x += __shfl_down_sync(0xffffffffu, x, o);
            y += __shfl_down_sync(0xffffffffu, y, o);
            z += __shfl_down_sync(0xffffffffu, z, o);
        }
        if (lane == 0) { red[k][0][warp] = x; red[k][1][warp] = y; red[k][2][warp] = z; }
    }
    __syncthreads();
    if (threadIdx.x < DA_GROUP * 3 && g0 + threadIdx.x / 3 < nt)
    {
        const int k = threadIdx.x / 3, c = threadIdx.x % 3;
        double v = 0.0;
        for (int w = 0; w < 8; ++w) v += red[k][c][w];
        out[((size_t)(g0 + k) * gridDim.x + blockIdx.x) * 3 + c] = v;
    }
}

// acc planes [3][count] of the owned range -> acc3[k][3] for the listed GLOBAL body indices (NaN if not owned)
__global__ void k_gather_accel(const double* __restrict__ acc, int first, int count, const unsigned int* __restrict__ bodies, int nt,
                               double* __restrict__ out)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nt) return;
    const long long li = (long long)bodies[k] - first;
    const bool mine = li >= 0 && li < count;
    const double nan = __longlong_as_double(0x7ff8000000000000ll);
    out[3 * k] = mine ? acc[li] : nan;
    out[3 * k + 1] = mine ? acc[(size_t)count + li] : nan;
    out[3 * k + 2] = mine ? acc[2 * (size_t)count + li] : nan;
}

// Position-sensitive, order-free 64-bit checksum: sum over elements of mix(index, bits) mod 2^64.  The sum
// commutes, so any reduction order gives the same value; two arrays agree iff (up to 2^-64) they are bitwise equal.
__device__ __forceinline__ unsigned long long mix64(unsigned long long x)
{
    x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
    return x;
}

__global__ void __launch_bounds__(256)
k_hash_words(const unsigned int* __restrict__ words, size_t nwords, size_t offset, unsigned long long salt,
             unsigned long long* __restrict__ out)
{
    // word i is keyed by its GLOBAL position offset + i, so the checksums of the shards of an array add up
    // (mod 2^64) to the checksum of the whole array
    unsigned long long acc = 0ull;
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < nwords; i += (size_t)gridDim.x * 256)
    {
        const size_t g = offset + i;
        acc += mix64(((unsigned long long)words[i] << 32 | (unsigned long long)(g & 0xffffffffull)) ^ mix64(salt + (g >> 32)));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(out, acc);
}

}  // namespace nb

using namespace nb;

extern "C" int nb_direct_accel(nb_handle h, const uint32_t* bodies, size_t k, double* acc3)
{
    NB_REQUIRE(h != nullptr && bodies != nullptr && acc3 != nullptr, NB_ERR_ARG, "null argument");
    NB_REQUIRE(h->n > 0, NB_ERR_STATE, "not initialised");
    NB_REQUIRE(k > 0 && k <= 65535u * DA_GROUP, NB_ERR_ARG, "between 1 and 524280 targets");
    NB_REQUIRE(h->exchanged, NB_ERR_STATE, "positions of remote ranks are stale");
    for (size_t i = 0; i < k; ++i) NB_REQUIRE(bodies[i] < h->n, NB_ERR_ARG, "body index out of range");
    NB_CUDA(cudaSetDevice(h->cfg.device));
    if (h->p2p_attached) NB_CHECK(p2p_wait(h));
    const int chunks = (int)std::min<size_t>(2 * (size_t)h->sm_count, (h->n + 4095) / 4096);
    const int chunk = (int)((h->n + chunks - 1) / chunks);
    const int groups = (int)((k + DA_GROUP - 1) / DA_GROUP);
    unsigned int* d_idx = nullptr;
    double* d_out = nullptr;
    const size_t words = k * (size_t)chunks * 3;
    cudaError_t e = cudaMalloc(&d_idx, k * sizeof(unsigned int));
    if (e == cudaSuccess) e = cudaMalloc(&d_out, words * sizeof(double));
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_idx, bodies, k * sizeof(unsigned int), cudaMemcpyHostToDevice, h->stream);
    if (e == cudaSuccess)
    {
        k_direct_accel<<<dim3(chunks, groups), 256, 0, h->stream>>>(h->posw, (int)h->n, d_idx, (int)k, chunk, h->cfg.softening, d_out);
        e = cudaGetLastError();
    }
    double* part = new (std::nothrow) double[words];
    if (e == cudaSuccess && part == nullptr) { cudaFree(d_idx); cudaFree(d_out); NB_REQUIRE(false, NB_ERR_NOMEM, "out of host memory"); }
    if (e == cudaSuccess) e = cudaMemcpyAsync(part, d_out, words * sizeof(double), cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    cudaFree(d_idx);
    cudaFree(d_out);
    if (e == cudaSuccess)
        for (size_t i = 0; i < k; ++i)
            for (int c = 0; c < 3; ++c)
            {
                double v = 0.0;                             // chunk order: deterministic
                for (int b = 0; b < chunks; ++b) v += part[(i * chunks + b) * 3 + c];
                acc3[3 * i + c] = v;
            }
    delete[] part;
    NB_CUDA(e);
    ++h->total_launches;
    return NB_OK;
}

// acc planes of the owned range -> acc3[k][3] for the listed global body indices
static int gather_accel(nb_handle h, const uint32_t* bodies, size_t k, double* acc3)
{
    unsigned int* d_idx = nullptr;
    double* d_out = nullptr;
    cudaError_t e = cudaMalloc(&d_idx, k * sizeof(unsigned int));
    if (e == cudaSuccess) e = cudaMalloc(&d_out, 3 * k * sizeof(double));
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_idx, bodies, k * sizeof(unsigned int), cudaMemcpyHostToDevice, h->stream);
    if (e == cudaSuccess)
    {
        k_gather_accel<<<(int)((k + 255) / 256), 256, 0, h->stream>>>(h->acc, (int)h->first, (int)h->count, d_idx, (int)k, d_out);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(acc3, d_out, 3 * k * sizeof(double), cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    cudaFree(d_idx);
    cudaFree(d_out);
    NB_CUDA(e);
    ++h->total_launches;
    return NB_OK;
}

extern "C" int nb_get_accel_of(nb_handle h, const uint32_t* bodies, size_t k, double* acc3)
{
    NB_REQUIRE(h != nullptr && bodies != nullptr && acc3 != nullptr && k > 0, NB_ERR_ARG, "null argument");
    NB_REQUIRE(h->n > 0, NB_ERR_STATE, "not initialised");
    NB_CUDA(cudaSetDevice(h->cfg.device));
    if (!h->acc_valid) NB_CHECK(nb_compute_accel(h));
    return gather_accel(h, bodies, k, acc3);
}

extern "C" int nb_get_step_accel_of(nb_handle h, const uint32_t* bodies, size_t k, double* acc3)
{
    NB_REQUIRE(h != nullptr && bodies != nullptr && acc3 != nullptr && k > 0, NB_ERR_ARG, "null argument");
    NB_REQUIRE(h->n > 0, NB_ERR_STATE, "not initialised");
    NB_REQUIRE(h->acc_is_last_step, NB_ERR_STATE, "no nb_step since the last nb_init_* / nb_compute_accel");
    NB_CUDA(cudaSetDevice(h->cfg.device));
    return gather_accel(h, bodies, k, acc3);
}

extern "C" int nb_state_hash(nb_handle h, uint64_t hash2[2])
{
    NB_REQUIRE(h != nullptr && hash2 != nullptr, NB_ERR_ARG, "null argument");
    NB_REQUIRE(h->n > 0, NB_ERR_STATE, "not initialised");
    NB_REQUIRE(h->exchanged, NB_ERR_STATE, "positions of remote ranks are stale");
    NB_CUDA(cudaSetDevice(h->cfg.device));
    if (h->p2p_attached) NB_CHECK(p2p_wait(h));
    unsigned long long* d = nullptr;
    NB_CUDA(cudaMalloc(&d, 2 * sizeof(unsigned long long)));
    cudaError_t e = cudaMemsetAsync(d, 0, 2 * sizeof(unsigned long long), h->stream);
    if (e == cudaSuccess)
    {
        const int blocks = h->sm_count * 8;
        // [0]: positions + weights of ALL bodies (identical on every rank after the exchange)
        k_hash_words<<<blocks, 256, 0, h->stream>>>(reinterpret_cast<const unsigned int*>(h->posw), h->n * 4, 0, 0ull, d);
        // [1]: velocities of the owned range, every word keyed by (component, GLOBAL body index): the values of
        //      all ranks add up (mod 2^64) to the single-handle value
        for (int c = 0; c < 3; ++c)
            k_hash_words<<<blocks, 256, 0, h->stream>>>(reinterpret_cast<const unsigned int*>(h->vel + (size_t)c * h->count), h->count * 2,
                                                        h->first * 2, 0x9e3779b97f4a7c15ull * (unsigned long long)(c + 1), d + 1);
        e = cudaGetLastError();
    }
    unsigned long long out[2] = {0ull, 0ull};
    if (e == cudaSuccess) e = cudaMemcpyAsync(out, d, sizeof(out), cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    cudaFree(d);
    NB_CUDA(e);
    h->total_launches += 4;
    hash2[0] = out[0];
    hash2[1] = out[1];
    return NB_OK;
}

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
