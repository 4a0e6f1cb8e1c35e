// Energy diagnostic.  The reference has none; this is the quantity its
// integrator conserves.  The pair force  f(r) = G ma mb / (r^2 + S)  (Physics.hpp:33-34) along the
// unit separation integrates to the pair potential
//     U(r) = -(G ma mb / sqrt(S)) * atan(sqrt(S) / r),
// and because positions advance by v dt / position_scale (BruteForceCPU.cpp:66) the conserved
// quantity is   E = sum_i 1/2 m_i v_i^2  +  position_scale * sum_{i<j} U(r_ij).
// Evaluated exactly (all pairs, fp64) over owned targets x all sources; each rank returns its
// share (half of every pair it touches), the host adds the shares.
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

}  // namespace nb
