// K8 -- device galaxy seeder.
//
// Same layout and distributions as GalaxySeeder<Particle>::Seed (reference
// src/Sim/GalaxySeeder.cpp:43-143): bodies [0, 2*61*floor(0.4 N / 60)) are the two spiral arms
// (61 segments each, mass 1e20, velocity along the segment normal), the rest is the disk
// (x, y ~ U(-2000, 2000), z ~ N(0, 16), kept if inside the radius-720 sphere; solid-body tangent
// velocity times U(0.8, 1.2) * 1e14; mass ~ U(1e28, 1e30)); positions are divided by `scale` and
// rotated by the random orientation, velocities are not rotated (GalaxySeeder.cpp:91-93).
//
// The reference draws everything from ONE serial minstd_rand0 stream with data-dependent draw
// counts, which cannot be reproduced bit for bit in parallel without replaying the stream; this
// kernel uses a counter-based generator keyed by (seed, body, draw) instead, so every body is
// independent and all ranks can generate all bodies without communication.  It is
// distribution-equivalent, not stream-equivalent; nb_seed_galaxy_host is the bit-exact one.
// The 122 segment frames and the orientation matrix are computed on the host with the same
// fp32 operation order as the host seeder and handed to the kernel.
#include <cmath>
#include <cstring>

#include "nb_internal.h"

namespace nb
{

struct SegmentFrame
{
    float sx[3], ex[3], sy[3], ey[3];
    float vel[3];
};

struct SeedParams
{
    float rot[3][3];
    float inv_scale_is_div;   // scale (positions are divided, like the reference)
    unsigned long long seed;
    int n, first, count;
    int arm_bodies;           // 2 * 61 * per_segment
    int per_segment;
    double G;
};

__constant__ SegmentFrame c_segments[122];

__device__ __forceinline__ unsigned long long mix64(unsigned long long z)
{
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

struct Counter
{
    unsigned long long key;
    unsigned int ctr;
    __device__ Counter(unsigned long long seed, unsigned int body) : key(mix64(seed ^ ((unsigned long long)body << 32 | 0x5bd1e995u))), ctr(0) {}
    __device__ unsigned long long next() { return mix64(key + 0x632BE59BD9B4E019ull * (unsigned long long)(++ctr)); }
    __device__ float unif() { return (float)(next() >> 40) * (1.0f / 16777216.0f); }                 // [0,1)
    __device__ double unifd() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }          // [0,1)
    __device__ float normal()
    {
        const float u1 = ((float)(next() >> 40) + 1.0f) * (1.0f / 16777216.0f);   // (0,1]
        const float u2 = unif();
        float s, c;
        sincospif(2.0f * u2, &s, &c);
        return sqrtf(-2.0f * logf(u1)) * c;
    }
};

__global__ void __launch_bounds__(256)
k_seed_galaxy(SeedParams P, float4* __restrict__ posw, double* __restrict__ vel, double* __restrict__ mass,
              int* __restrict__ wmax_bits)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.n) return;
    // heaviest possible body: G * 1e30 (every rank seeds all bodies, so a constant is exact enough)
    if (i == 0) atomicMax(wmax_bits, __float_as_int((float)(P.G * 1e30)));
    Counter rng(P.seed, (unsigned int)i);
    float px, py, pz;
    double vx, vy, vz, m;
    if (i < P.arm_bodies)
    {
        const SegmentFrame& f = c_segments[i / P.per_segment];
        const float tx = rng.normal() * 0.2f + 0.5f;
        const float ty = rng.unif() * (0.5f - 0.2f) + 0.2f;
        px = ((f.ex[0] - f.sx[0]) * tx + f.sx[0]) + ((f.ey[0] - f.sy[0]) * ty + f.sy[0]);
        py = ((f.ex[1] - f.sx[1]) * tx + f.sx[1]) + ((f.ey[1] - f.sy[1]) * ty + f.sy[1]);
        pz = rng.normal() * 16.0f;
        vx = (double)f.vel[0]; vy = (double)f.vel[1]; vz = (double)f.vel[2];
        m = 1e20;
    }
    else
    {
        for (;;)
        {
            px = rng.unif() * 4000.0f - 2000.0f;
            py = rng.unif() * 4000.0f - 2000.0f;
            pz = rng.normal() * 16.0f;
            if ((px * px + py * py) + pz * pz <= 720.0f * 720.0f) break;
        }
        // tangent = pos x (0,0,1) = (y, -x, 0)
        const double k = (rng.unifd() * (1.2 - 0.8) + 0.8) * 1e14;
        vx = (double)py * k; vy = (double)(-px) * k; vz = 0.0 * k;
        m = rng.unifd() * (1e30 - 1e28) + 1e28;
    }
    const float qx = px / P.inv_scale_is_div, qy = py / P.inv_scale_is_div, qz = pz / P.inv_scale_is_div;
    const float rx = (qz * P.rot[2][0]) + qy * P.rot[1][0] + qx * P.rot[0][0];
    const float ry = (qz * P.rot[2][1]) + qy * P.rot[1][1] + qx * P.rot[0][1];
    const float rz = (qz * P.rot[2][2]) + qy * P.rot[1][2] + qx * P.rot[0][2];
    posw[i] = make_float4(rx, ry, rz, (float)(P.G * m));
    const int li = i - P.first;
    if (li >= 0 && li < P.count)
    {
        vel[li] = vx;
        vel[(size_t)P.count + li] = vy;
        vel[2 * (size_t)P.count + li] = vz;
        mass[li] = m;
    }
}

static void host_rotation(unsigned long long seed, float rot[3][3])
{
    // three uniform angles from the same counter generator (host copy of mix64)
    auto mix = [](unsigned long long z) {
        z += 0x9E3779B97F4A7C15ull;
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    };
    float ang[3];
    for (int k = 0; k < 3; ++k)
        ang[k] = (float)(mix(seed * 3 + k + 0xA5A5A5A5ull) >> 40) * (1.0f / 16777216.0f) * (2.0f * 3.141592654f);
    const float hp = ang[1] * 0.5f, hy = ang[0] * 0.5f, hr = ang[2] * 0.5f;
    const float sp = sinf(hp), cp = cosf(hp), sy = sinf(hy), cy = cosf(hy), sr = sinf(hr), cr = cosf(hr);
    const float qx = (cr * sp) * cy + (sr * cp) * sy;
    const float qy = (cr * cp) * sy - (sr * sp) * cy;
    const float qz = (sr * cp) * cy - (cr * sp) * sy;
    const float qw = (cr * cp) * cy + (sr * sp) * sy;
    rot[0][0] = 1.f - 2.f * (qy * qy + qz * qz); rot[0][1] = 2.f * (qx * qy + qw * qz); rot[0][2] = 2.f * (qx * qz - qw * qy);
    rot[1][0] = 2.f * (qx * qy - qw * qz); rot[1][1] = 1.f - 2.f * (qx * qx + qz * qz); rot[1][2] = 2.f * (qy * qz + qw * qx);
    rot[2][0] = 2.f * (qx * qz + qw * qy); rot[2][1] = 2.f * (qy * qz - qw * qx); rot[2][2] = 1.f - 2.f * (qx * qx + qy * qy);
}

int seed_galaxy_device(nb_sim* h, size_t n, uint64_t seed, float scale)
{
    SegmentFrame frames[122];
    int s = 0;
    const float offsets[2] = {0.0f, 3.14f};
    for (int arm = 0; arm < 2; ++arm)
    {
        const float offset = offsets[arm];
        for (float angle = 0.0f, r = 2.0f; angle < 6.0f && s < 61 * (arm + 1); angle += 0.1f, r += 7.2f, ++s)
        {
            const float sp[3] = {cosf(angle + offset) * r, sinf(angle + offset) * r, 0.0f};
            const float sn[3] = {cosf(angle + offset + 0.1f) * (r + 10.0f), sinf(angle + offset + 0.1f) * (r + 10.0f), 0.0f};
            float nrm[3] = {sp[0] - sn[0], sp[1] - sn[1], sp[2] - sn[2]};
            const float mag = sqrtf((nrm[0] * nrm[0] + nrm[1] * nrm[1]) + nrm[2] * nrm[2]);
            for (int k = 0; k < 3; ++k) nrm[k] = nrm[k] / mag;
            float tan[3] = {nrm[1] * 1.0f - nrm[2] * 0.0f, nrm[2] * 0.0f - nrm[0] * 1.0f, 0.0f};
            const float tl = sqrtf((tan[0] * tan[0] + tan[1] * tan[1]) + tan[2] * tan[2]);
            for (int k = 0; k < 3; ++k) tan[k] = tan[k] / tl;
            SegmentFrame& f = frames[s];
            for (int k = 0; k < 3; ++k)
            {
                f.sx[k] = sp[k] - tan[k] * 140.0f;
                f.ex[k] = sp[k] + tan[k] * 140.0f;
                f.sy[k] = sp[k] - nrm[k] * 400.0f;
                f.ey[k] = sp[k] + nrm[k] * 400.0f;
                f.vel[k] = nrm[k] * 2e16f * (1000.0f / mag);
            }
        }
    }
    if (s != 122) { set_error("seed_galaxy_device: expected 122 arm segments, built %d", s); return NB_ERR_STATE; }
    NB_CUDA(cudaMemcpyToSymbolAsync(c_segments, frames, sizeof(frames), 0, cudaMemcpyHostToDevice, h->stream));

    SeedParams P;
    std::memset(&P, 0, sizeof(P));
    host_rotation(seed, P.rot);
    P.inv_scale_is_div = scale;
    P.seed = seed;
    P.n = (int)n; P.first = (int)h->first; P.count = (int)h->count;
    P.per_segment = (int)std::floor(((float)n * 0.4f) / 60);
    P.arm_bodies = 2 * 61 * P.per_segment;
    if (P.per_segment < 1) { P.per_segment = 1; P.arm_bodies = 0; }
    if ((size_t)P.arm_bodies > n) P.arm_bodies = (int)n;
    P.G = h->cfg.G;
    k_seed_galaxy<<<(int)((n + 255) / 256), 256, 0, h->stream>>>(P, h->posw, h->vel, h->mass, reinterpret_cast<int*>(h->wmax));
    NB_CUDA(cudaGetLastError());
    NB_CUDA(cudaStreamSynchronize(h->stream));
    return NB_OK;
}

}  // namespace nb
