// K2 (fused kick-drift) and K9 (AoS Particle <-> SoA) kernels.
//
// Integrator of the reference, BruteForceCPU.cpp:61-73 and BarnesHut.cpp:81-95 (symplectic Euler):
//     a = Forces / Mass;  Velocity += a * dt;                       (double)
//     vel = (Velocity * dt) / Phys::StarSystemScale;                (double, true division)
//     Position += Vector3((float)vel.x, (float)vel.y, (float)vel.z) (float add)
// reproduced operation for operation; the only difference is where `a` comes from.
#include "nb_internal.h"

namespace nb
{

// One thread per owned body.  Sums the all-pairs partials in split order (deterministic), kicks the
// fp64 velocity, drifts the fp32 position and writes the new float4 straight into the position
// array every kernel (and the NCCL all-gather) reads -- no separate "publish" pass.
__global__ void __launch_bounds__(256)
k_kick_drift(float4* __restrict__ posw, int first, int count, double* __restrict__ vel,
             const double* __restrict__ acc_part, int splits, double* __restrict__ acc, double dt,
             double pos_scale)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const size_t plane = (size_t)count;
    double ax, ay, az;
    if (acc_part != nullptr)
    {
        ax = ay = az = 0.0;
        for (int s = 0; s < splits; ++s)
        {
            const double* p = acc_part + (size_t)s * 3 * plane;
            ax += p[i];
            ay += p[plane + i];
            az += p[2 * plane + i];
        }
        acc[i] = ax; acc[plane + i] = ay; acc[2 * plane + i] = az;
    }
    else
    {
        ax = acc[i]; ay = acc[plane + i]; az = acc[2 * plane + i];
    }
    double vx = vel[i], vy = vel[plane + i], vz = vel[2 * plane + i];
    // Velocity += a * dt: a rounded product, then a rounded sum, as the reference's g++ build (-ffp-contract=off)
    // computes it -- not a DFMA, which differs by an ulp every few steps
    vx = __dadd_rn(vx, __dmul_rn(ax, dt)); vy = __dadd_rn(vy, __dmul_rn(ay, dt)); vz = __dadd_rn(vz, __dmul_rn(az, dt));
    vel[i] = vx; vel[plane + i] = vy; vel[2 * plane + i] = vz;
    float4 p = posw[first + i];
    p.x += (float)((vx * dt) / pos_scale);
    p.y += (float)((vy * dt) / pos_scale);
    p.z += (float)((vz * dt) / pos_scale);
    posw[first + i] = p;
}

// Sums the partials only (parity hook nb_compute_accel).
__global__ void __launch_bounds__(256)
k_reduce_partials(const double* __restrict__ acc_part, int splits, int count, double* __restrict__ acc)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 3 * count) return;
    double a = 0.0;
    for (int s = 0; s < splits; ++s) a += acc_part[(size_t)s * 3 * (size_t)count + i];
    acc[i] = a;
}

// AoS -> SoA.  Every body's position and G*m go to posw (all ranks need all sources); velocity and
// mass only for the owned range.
__global__ void __launch_bounds__(256)
k_unpack_aos(const unsigned char* __restrict__ aos, size_t stride, int begin, int end, int first, int count,
             float4* __restrict__ posw, double* __restrict__ vel, double* __restrict__ mass, double G,
             int* __restrict__ wmax_bits)
{
    const int i = begin + blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = i < end;
    float w = 0.f;
    const unsigned char* rec = aos + (size_t)(live ? i : begin) * stride;
    const double m = *reinterpret_cast<const double*>(rec + NB_OFF_MASS);
    if (live) w = fmaxf((float)(G * m), 0.f);
    // max_j w_j for the weight-folded all-pairs kernel: non-negative floats order like their bits
    const int wmax = __reduce_max_sync(0xffffffffu, __float_as_int(w));
    if ((threadIdx.x & 31) == 0) atomicMax(wmax_bits, wmax);
    if (!live) return;
    const float* pos = reinterpret_cast<const float*>(rec + NB_OFF_POSITION);
    posw[i] = make_float4(pos[0], pos[1], pos[2], (float)(G * m));
    const int li = i - first;
    if (li >= 0 && li < count)
    {
        const double* v = reinterpret_cast<const double*>(rec + NB_OFF_VELOCITY);
        vel[li] = v[0];
        vel[(size_t)count + li] = v[1];
        vel[2 * (size_t)count + li] = v[2];
        mass[li] = m;
    }
}

// SoA -> AoS for the owned range: Position, Velocity, Forces only (colours are never touched).
__global__ void __launch_bounds__(256)
k_pack_aos(unsigned char* __restrict__ aos, size_t stride, int first, int count,
           const float4* __restrict__ posw, const double* __restrict__ vel,
           const double* __restrict__ mass, const double* __restrict__ acc, int forces_zero)
{
    const int li = blockIdx.x * blockDim.x + threadIdx.x;
    if (li >= count) return;
    unsigned char* rec = aos + (size_t)(first + li) * stride;
    const float4 p = posw[first + li];
    float* pos = reinterpret_cast<float*>(rec + NB_OFF_POSITION);
    pos[0] = p.x; pos[1] = p.y; pos[2] = p.z;
    double* v = reinterpret_cast<double*>(rec + NB_OFF_VELOCITY);
    v[0] = vel[li];
    v[1] = vel[(size_t)count + li];
    v[2] = vel[2 * (size_t)count + li];
    double* f = reinterpret_cast<double*>(rec + NB_OFF_FORCES);
    if (forces_zero)
    {
        f[0] = f[1] = f[2] = 0.0;
    }
    else
    {
        const double m = mass[li];
        f[0] = m * acc[li];
        f[1] = m * acc[(size_t)count + li];
        f[2] = m * acc[2 * (size_t)count + li];
    }
}

static inline int blocks_for(size_t n, int threads) { return (int)((n + threads - 1) / threads); }

int launch_kick_drift(nb_sim* h, float dt)
{
    const bool partials = (h->cfg.mode == NB_MODE_ALLPAIRS);
    k_kick_drift<<<blocks_for(h->count, 256), 256, 0, h->stream>>>(
        h->posw, (int)h->first, (int)h->count, h->vel, partials ? h->acc_part : nullptr, h->ap_splits,
        h->acc, (double)dt, h->cfg.position_scale);
    NB_CUDA(cudaGetLastError());
    ++h->last_launches;
    return NB_OK;
}

int launch_reduce_partials(nb_sim* h)
{
    k_reduce_partials<<<blocks_for(3 * h->count, 256), 256, 0, h->stream>>>(h->acc_part, h->ap_splits,
                                                                           (int)h->count, h->acc);
    NB_CUDA(cudaGetLastError());
    ++h->last_launches;
    return NB_OK;
}

int launch_unpack_aos(nb_sim* h, size_t stride, size_t begin, size_t end)
{
    if (begin == 0 && end == h->n) NB_CUDA(cudaMemsetAsync(h->wmax, 0, sizeof(float), h->stream));
    k_unpack_aos<<<blocks_for(end - begin, 256), 256, 0, h->stream>>>(
        static_cast<const unsigned char*>(h->d_aos), stride, (int)begin, (int)end, (int)h->first, (int)h->count,
        h->posw, h->vel, h->mass, h->cfg.G, reinterpret_cast<int*>(h->wmax));
    NB_CUDA(cudaGetLastError());
    ++h->last_launches;
    return NB_OK;
}

int launch_pack_aos(nb_sim* h, size_t stride, bool forces_zero)
{
    k_pack_aos<<<blocks_for(h->count, 256), 256, 0, h->stream>>>(
        static_cast<unsigned char*>(h->d_aos), stride, (int)h->first, (int)h->count, h->posw, h->vel,
        h->mass, h->acc, forces_zero ? 1 : 0);
    NB_CUDA(cudaGetLastError());
    ++h->last_launches;
    return NB_OK;
}

// Forces the lazily loaded kernels of this file into the context (CUDA 12 loads a kernel at its first
// launch, and that load can wait for the device to drain -- fatal if it happens while another handle of
// the same process sits in a peer-flag wait; see p2p.cu).
int preload_integrate()
{
    cudaFuncAttributes a;
    NB_CUDA(cudaFuncGetAttributes(&a, reinterpret_cast<const void*>(k_kick_drift)));
    NB_CUDA(cudaFuncGetAttributes(&a, reinterpret_cast<const void*>(k_reduce_partials)));
    NB_CUDA(cudaFuncGetAttributes(&a, reinterpret_cast<const void*>(k_unpack_aos)));
    NB_CUDA(cudaFuncGetAttributes(&a, reinterpret_cast<const void*>(k_pack_aos)));
    return NB_OK;
}

}  // namespace nb
