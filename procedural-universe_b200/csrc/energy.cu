// Energy diagnostic.  The reference has none; this is the quantity its
// integrator conserves.  The pair force  f(r) = G ma mb / (r^2 + S)  (Physics.hpp:33-34) along the
// unit separation integrates to the pair potential
//     U(r) = -(G ma mb / sqrt(S)) * atan(sqrt(S) / r),
// and because positions advance by v dt / position_scale (BruteForceCPU.cpp:66) the conserved
// quantity is   E = sum_i 1/2 m_i v_i^2  +  position_scale * sum_{i<j} U(r_ij).
// Evaluated exactly (all pairs, fp64) over owned targets x all sources; each rank returns its
// share (half of every pair it touches), the host adds the shares.
#include <algorithm>
#include <vector>

#include "nb_internal.h"

namespace nb
{

__global__ void __launch_bounds__(256)
k_energy(const float4* __restrict__ posw, int n, int first, int count, const double* __restrict__ vel,
         const double* __restrict__ mass, double inv_sqrt_s, double sqrt_s, double inv_G, double* __restrict__ out2)
{
    __shared__ float4 tile[256];
    __shared__ double red[2][256];
    const int li = blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = li < count;
    float4 me = make_float4(0.f, 0.f, 0.f, 0.f);
    double m = 0.0, ke = 0.0;
    if (live)
    {
        me = posw[first + li];
        m = mass[li];
        const double vx = vel[li], vy = vel[(size_t)count + li], vz = vel[2 * (size_t)count + li];
        ke = 0.5 * m * (vx * vx + vy * vy + vz * vz);
    }
    double pe = 0.0;
    for (int j0 = 0; j0 < n; j0 += 256)
    {
        const int j = j0 + threadIdx.x;
        tile[threadIdx.x] = (j < n) ? posw[j] : make_float4(0.f, 0.f, 0.f, 0.f);
        __syncthreads();
        const int lim = min(256, n - j0);
        if (live)
            for (int k = 0; k < lim; ++k)
            {
                if (j0 + k == first + li) continue;
                const float4 s = tile[k];
                const double dx = (double)s.x - (double)me.x, dy = (double)s.y - (double)me.y, dz = (double)s.z - (double)me.z;
                const double r = sqrt(dx * dx + dy * dy + dz * dz);
                // G m_i m_j = m_i * w_j
                const double a = (r > 0.0) ? atan(sqrt_s / r) : 1.5707963267948966;
                pe -= (double)s.w * a;
            }
        __syncthreads();
    }
    pe *= 0.5 * m * inv_sqrt_s;
    red[0][threadIdx.x] = ke;
    red[1][threadIdx.x] = pe;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1)
    {
        if (threadIdx.x < s)
        {
            red[0][threadIdx.x] += red[0][threadIdx.x + s];
            red[1][threadIdx.x] += red[1][threadIdx.x + s];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0)
    {
        out2[2 * blockIdx.x] = red[0][0];
        out2[2 * blockIdx.x + 1] = red[1][0];
    }
}

int energy(nb_sim* h, double* ke, double* pe)
{
    const int blocks = (int)((h->count + 255) / 256);
    double* d = nullptr;
    NB_CUDA(cudaMalloc(&d, (size_t)blocks * 2 * sizeof(double)));
    const double s = h->cfg.softening;
    k_energy<<<blocks, 256, 0, h->stream>>>(h->posw, (int)h->n, (int)h->first, (int)h->count, h->vel, h->mass,
                                            1.0 / sqrt(s), sqrt(s), 1.0 / h->cfg.G, d);
    cudaError_t e = cudaGetLastError();
    double* hbuf = new double[(size_t)blocks * 2];
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    if (e == cudaSuccess) e = cudaMemcpy(hbuf, d, (size_t)blocks * 2 * sizeof(double), cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (e != cudaSuccess)
    {
        delete[] hbuf;
        set_error("nb_energy: %s", cudaGetErrorString(e));
        return NB_ERR_CUDA;
    }
    // fixed-order host sum of the block partials (deterministic)
    double k = 0.0, p = 0.0;
    for (int b = 0; b < blocks; ++b) { k += hbuf[2 * b]; p += hbuf[2 * b + 1]; }
    delete[] hbuf;
    if (ke) *ke = k;
    if (pe) *pe = p * h->cfg.position_scale;
    return NB_OK;
}

// ------------------------------------------------------------------------------------------------
// Sampled estimator for sizes where the exact pair sum is out of reach (config 5: 64 M bodies).
// Kinetic energy exactly over the owned bodies; potential from the owned bodies whose GLOBAL index
// is a multiple of `stride`, each against ALL sources:  PE ~= stride * sum_{i in sample} 1/2 m_i phi_i.
// The sample is a fixed set of bodies, so E(t) - E(0) is taken over the same bodies at both times.
// Layout: a block takes ES_GROUP samples and one chunk of sources; a thread loads each source once
// and applies it to the group's samples held in registers.
// ------------------------------------------------------------------------------------------------
constexpr int ES_GROUP = 8;

__global__ void __launch_bounds__(256)
k_kinetic(int count, const double* __restrict__ vel, const double* __restrict__ mass, double* __restrict__ out)
{
    __shared__ double red[256];
    double ke = 0.0;
    for (size_t li = (size_t)blockIdx.x * 256 + threadIdx.x; li < (size_t)count; li += (size_t)gridDim.x * 256)
    {
        const double vx = vel[li], vy = vel[(size_t)count + li], vz = vel[2 * (size_t)count + li];
        ke += 0.5 * mass[li] * (vx * vx + vy * vy + vz * vz);
    }
    red[threadIdx.x] = ke;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1)
    {
        if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[blockIdx.x] = red[0];
}

__global__ void __launch_bounds__(256)
k_potential_sampled(const float4* __restrict__ posw, int n, int first_sample, int stride, int nsamples, int chunk,
                    double sqrt_s, double* __restrict__ out /* [nsamples][gridDim.x] */)
{
    __shared__ double red[ES_GROUP][8];
    const int g0 = blockIdx.y * ES_GROUP;
    float4 me[ES_GROUP];
    int idx[ES_GROUP];
#pragma unroll
    for (int k = 0; k < ES_GROUP; ++k)
    {
        const int s = min(g0 + k, nsamples - 1);
        idx[k] = first_sample + s * stride;
        me[k] = posw[idx[k]];
    }
    double phi[ES_GROUP];
#pragma unroll
    for (int k = 0; k < ES_GROUP; ++k) phi[k] = 0.0;
    const int j0 = blockIdx.x * chunk, j1 = min(n, j0 + chunk);
    for (int j = j0 + threadIdx.x; j < j1; j += 256)
    {
        const float4 s = posw[j];
#pragma unroll
        for (int k = 0; k < ES_GROUP; ++k)
        {
            const double dx = (double)s.x - (double)me[k].x, dy = (double)s.y - (double)me[k].y, dz = (double)s.z - (double)me[k].z;
            const double r = sqrt(dx * dx + dy * dy + dz * dz);
            const double a = (r > 0.0) ? atan(sqrt_s / r) : 1.5707963267948966;
            if (j != idx[k]) phi[k] -= (double)s.w * a;
        }
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int k = 0; k < ES_GROUP; ++k)
    {
        double v = phi[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if (lane == 0) red[k][warp] = v;
    }
    __syncthreads();
    if (threadIdx.x < ES_GROUP && g0 + threadIdx.x < nsamples)
    {
        double v = 0.0;
        for (int w = 0; w < 8; ++w) v += red[threadIdx.x][w];
        out[(size_t)(g0 + threadIdx.x) * gridDim.x + blockIdx.x] = v;
    }
}

int energy_sampled(nb_sim* h, size_t stride, double* ke, double* pe, size_t* nsamples_out)
{
    const size_t first = h->first, count = h->count, n = h->n;
    // owned bodies with global index % stride == 0
    const size_t s0 = (first + stride - 1) / stride * stride;
    const size_t ns = s0 < first + count ? (first + count - 1 - s0) / stride + 1 : 0;
    if (nsamples_out) *nsamples_out = ns;
    const int kblocks = (int)std::min<size_t>(1024, (count + 255) / 256);
    const int chunks = (int)std::min<size_t>(296, (n + 4095) / 4096);
    const int chunk = (int)((n + chunks - 1) / chunks);
    const int groups = (int)((ns + ES_GROUP - 1) / ES_GROUP);
    double* d = nullptr;
    const size_t words = (size_t)kblocks + ns * (size_t)chunks;
    NB_CUDA(cudaMalloc(&d, words * sizeof(double)));
    k_kinetic<<<kblocks, 256, 0, h->stream>>>((int)count, h->vel, h->mass, d);
    if (ns > 0)
        k_potential_sampled<<<dim3(chunks, groups), 256, 0, h->stream>>>(h->posw, (int)n, (int)s0, (int)stride, (int)ns, chunk,
                                                                         sqrt(h->cfg.softening), d + kblocks);
    cudaError_t e = cudaGetLastError();
    std::vector<double> hbuf(words), m(ns);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    if (e == cudaSuccess) e = cudaMemcpy(hbuf.data(), d, words * sizeof(double), cudaMemcpyDeviceToHost);
    for (size_t k = 0; k < ns && e == cudaSuccess; ++k)
        e = cudaMemcpy(&m[k], h->mass + (s0 + k * stride - first), sizeof(double), cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (e != cudaSuccess)
    {
        set_error("nb_energy_sampled: %s", cudaGetErrorString(e));
        return NB_ERR_CUDA;
    }
    double k = 0.0, p = 0.0;
    for (int b = 0; b < kblocks; ++b) k += hbuf[b];
    const double inv_sqrt_s = 1.0 / sqrt(h->cfg.softening);
    for (size_t i = 0; i < ns; ++i)
    {
        double phi = 0.0;
        for (int c = 0; c < chunks; ++c) phi += hbuf[kblocks + i * chunks + c];
        p += 0.5 * m[i] * inv_sqrt_s * phi;
    }
    if (ke) *ke = k;
    if (pe) *pe = p * (double)stride * h->cfg.position_scale;
    return NB_OK;
}


}  // namespace nb
