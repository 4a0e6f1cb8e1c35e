// K8 -- the reference's particle seeders on the device, bit for bit.
//
// GalaxySeeder<T>::Seed (reference src/Sim/GalaxySeeder.cpp:43-80, CreateSpiralArm :109-143, AddParticle
// :83-106) draws everything from ONE serial minstd_rand0 stream, and how many draws a body takes depends
// on the values drawn: libstdc++'s normal_distribution is Marsaglia's polar method (a rejection loop over
// PAIRS of draws, the second variate cached for the next call), and the disk loop rejects ~90 % of its
// attempts (`continue`, :65-69).  seed_host.cpp replays that stream on one core; this file PARSES it in
// parallel and then generates every body independently:
//
//  * the stream is random access: draw t is x0 * 16807^t mod (2^31 - 1), a modular power (table of
//    16807^(2^i), <= 40 multiplications);
//  * the consumer is a small state machine.  The stream falls into three phases (arm 1, arm 2, disk) and each
//    phase into UNITS whose length is a function of the draws at the unit's start only:
//      arm unit      = two consecutive arm bodies: the first draws the polar pair(s) of distx, the second uses
//                      the cached variate; distz alternates the same way (which body of the pair generates
//                      depends on the parity of the arm's body count -- two unit layouts);
//      disk unit     = two consecutive ATTEMPTS: the first generates the distz pair, the second uses the cached
//                      variate; an accepted attempt takes 7 more draws (k, mass, colour).
//    A cached variate always comes from the pair that sits right before the previous body's three colour draws,
//    so a unit can be generated from its start position alone;
//  * parse: the phase's stream is cut into chunks of 2048 positions.  For every chunk and every possible
//    entry offset o < 128 (the first unit of the chunk starts at chunk + o) one thread walks the units of the
//    chunk: table[chunk][o] = (exit offset into the next chunk, units -- or accepted bodies -- counted).
//    Tables compose (128 chunks -> 1, twice), a single thread walks the top level, and the entries are pushed
//    back down: every chunk then knows where its first unit starts and how many units / bodies precede it;
//  * generation: one thread per unit jumps to the unit's first draw and replays the reference's arithmetic --
//    libstdc++'s generate_canonical / uniform_real / normal_distribution, glibc's logf (restated below: its
//    table-driven algorithm, checked exhaustively against the host libm for every float in (0, 1]),
//    correctly rounded sqrt / division, DirectXMath's scalar operation order without fused multiply-adds.
//
// RandomSeeder (RandomSeeder.cpp:13-40) and StarSystemSeeder (StarSystemSeeder.cpp:18-55) take a fixed 11
// draws per body: one thread per body, no parse.
//
// The arm geometry (61 segment frames per arm: libm sinf / cosf) and the orientation matrix are computed on the
// host with the operation order of seed_host.cpp and handed to the kernels.
#include <cmath>
#include <cstring>
#include <limits>
#include <vector>

#include "nb_internal.h"

namespace nb
{

// ------------------------------------------------------------------------------------------------
// minstd_rand0
// ------------------------------------------------------------------------------------------------
constexpr unsigned int kLcgM = 2147483647u;

__host__ __device__ __forceinline__ unsigned int mulmod(unsigned int a, unsigned int b)
{
    unsigned long long p = (unsigned long long)a * (unsigned long long)b;   // < 2^62
    p = (p & kLcgM) + (p >> 31);                                            // < 2^32
    p = (p & kLcgM) + (p >> 31);                                            // <= 2^31
    return p >= kLcgM ? (unsigned int)(p - kLcgM) : (unsigned int)p;
}

struct SegmentFrame
{
    float sx[3], ex[3], sy[3], ey[3];
    float vel[3];
};

// Everything a seeding run needs, in device memory (two galaxies of a collision use two of these).
struct SeedConst
{
    unsigned int pow2[48];      // 16807^(2^i) mod M
    unsigned int x0;            // engine state after seeding (seed % M, 0 -> 1)
    SegmentFrame seg[122];
    float rot[4][4];
    float scale;
    float red[2], green[2], blue[2];
    int layout;
    unsigned long long stride;
    double canon_div;           // (double)((long double)2147483646.0 * 2147483646.0L): generate_canonical<double>'s divisor
    unsigned long long n;       // bodies
    unsigned long long arm;     // bodies per arm (61 * floor((float)n * 0.4f / 60))
    unsigned long long per_segment;
};

// the draw with 1-based index t: x0 * 16807^t
__device__ __forceinline__ unsigned int lcg_at(const SeedConst* __restrict__ K, unsigned long long t)
{
    unsigned int x = K->x0;
    for (int i = 0; t != 0ull; ++i, t >>= 1)
        if (t & 1ull) x = mulmod(x, K->pow2[i]);
    return x;
}

// A sequential reader positioned at draw t (the next call returns draw t).
struct Lcg
{
    unsigned int x;
    int used = 0;            // draws taken so far
    __device__ Lcg(const SeedConst* __restrict__ K, unsigned long long t) : x(lcg_at(K, t - 1)) {}
    __device__ __forceinline__ unsigned int next() { x = mulmod(x, 16807u); ++used; return x; }
    __device__ __forceinline__ void skip(int k) { for (int i = 0; i < k; ++i) next(); }
};

// ------------------------------------------------------------------------------------------------
// libstdc++ 13 distributions (bits/random.h, bits/random.tcc), every operation rounded on its own
// ------------------------------------------------------------------------------------------------
// generate_canonical<float, 24>: one draw; the range 2^31 - 2 rounds to 2^31 in float
__device__ __forceinline__ float canonical_f(unsigned int draw)
{
    const float r = __fmul_rn(__uint2float_rn(draw - 1u), 4.656612873077392578125e-10f);   // / 2^31, exact
    return r >= 1.0f ? __uint_as_float(0x3f7fffffu) : r;
}

__device__ __forceinline__ float uniform_f(unsigned int draw, float a, float b)
{
    return __fadd_rn(__fmul_rn(canonical_f(draw), __fsub_rn(b, a)), a);
}

// generate_canonical<double, 53>: two draws
__device__ __forceinline__ double canonical_d(unsigned int d1, unsigned int d2, double div)
{
    const double sum = __dadd_rn((double)(d1 - 1u), __dmul_rn((double)(d2 - 1u), 2147483646.0));
    const double r = __ddiv_rn(sum, div);
    return r >= 1.0 ? __longlong_as_double(0x3fefffffffffffffll) : r;
}

__device__ __forceinline__ double uniform_d(unsigned int d1, unsigned int d2, double div, double a, double b)
{
    return __dadd_rn(__dmul_rn(canonical_d(d1, d2, div), __dsub_rn(b, a)), a);
}

// one coordinate of the polar method: result_type(2.0) * canonical - 1.0 (the subtraction is in double)
__device__ __forceinline__ float polar_coord(unsigned int draw)
{
    return __double2float_rn((double)__fmul_rn(2.0f, canonical_f(draw)) - 1.0);
}

__device__ __forceinline__ bool polar_rejects(float x, float y, float* r2_out)
{
    const float r2 = __fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y));
    *r2_out = r2;
    return r2 > 1.0f || r2 == 0.0f;
}

// glibc 2.27+ logf (sysdeps/ieee754/flt-32/e_logf.c, the ARM optimized-routines algorithm) for normal
// 0 < x <= 1: table of 16 (1/c, log c), degree-3 polynomial in double, one rounding to float at the end.
// Checked against the host libm for all 1 056 964 609 floats of that range (fused and unfused evaluation
// round to the same float everywhere in it).
__device__ __forceinline__ float glibc_logf(float x)
{
    const double T[16][2] = {
        {0x1.661ec79f8f3bep+0, -0x1.57bf7808caadep-2}, {0x1.571ed4aaf883dp+0, -0x1.2bef0a7c06ddbp-2},
        {0x1.49539f0f010bp+0, -0x1.01eae7f513a67p-2},  {0x1.3c995b0b80385p+0, -0x1.b31d8a68224e9p-3},
        {0x1.30d190c8864a5p+0, -0x1.6574f0ac07758p-3}, {0x1.25e227b0b8eap+0, -0x1.1aa2bc79c81p-3},
        {0x1.1bb4a4a1a343fp+0, -0x1.a4e76ce8c0e5ep-4}, {0x1.12358f08ae5bap+0, -0x1.1973c5a611cccp-4},
        {0x1.0953f419900a7p+0, -0x1.252f438e10c1ep-5}, {0x1p+0, 0x0p+0},
        {0x1.e608cfd9a47acp-1, 0x1.aa5aa5df25984p-5},  {0x1.ca4b31f026aap-1, 0x1.c5e53aa362eb4p-4},
        {0x1.b2036576afce6p-1, 0x1.526e57720db08p-3},  {0x1.9c2d163a1aa2dp-1, 0x1.bc2860d22477p-3},
        {0x1.886e6037841edp-1, 0x1.1058bc8a07ee1p-2},  {0x1.767dcf5534862p-1, 0x1.4043057b6ee09p-2}};
    const unsigned int ix = __float_as_uint(x);
    if (ix == 0x3f800000u) return 0.0f;
    const unsigned int tmp = ix - 0x3f330000u;
    const int i = (int)((tmp >> 19) & 15u);
    const int k = (int)tmp >> 23;
    const unsigned int iz = ix - (tmp & 0xff800000u);
    const double z = (double)__uint_as_float(iz);
    const double r = __dadd_rn(__dmul_rn(z, T[i][0]), -1.0);
    const double y0 = __dadd_rn(T[i][1], __dmul_rn((double)k, 0x1.62e42fefa39efp-1));
    const double r2 = __dmul_rn(r, r);
    double y = __dadd_rn(__dmul_rn(0x1.5575b0be00b6ap-2, r), -0x1.ffffef20a4123p-2);
    y = __dadd_rn(__dmul_rn(-0x1.00ea348b88334p-2, r2), y);
    y = __dadd_rn(__dmul_rn(y, r2), __dadd_rn(y0, r));
    return __double2float_rn(y);
}

// the two variates of an accepted pair: *ret is returned now, *saved by the next call
__device__ __forceinline__ void polar_variates(float x, float y, float r2, float* ret, float* saved)
{
    const float mult = __fsqrt_rn(__fdiv_rn(__fmul_rn(-2.0f, glibc_logf(r2)), r2));
    *saved = __fmul_rn(x, mult);
    *ret = __fmul_rn(y, mult);
}

__device__ __forceinline__ float normal_apply(float v, float mean, float stddev) { return __fadd_rn(__fmul_rn(v, stddev), mean); }

// a generating call of normal_distribution on a sequential reader
__device__ __forceinline__ void normal_generate(Lcg& g, float* ret, float* saved)
{
    float x, y, r2;
    do
    {
        x = polar_coord(g.next());
        y = polar_coord(g.next());
    } while (polar_rejects(x, y, &r2));
    polar_variates(x, y, r2, ret, saved);
}

// ------------------------------------------------------------------------------------------------
// records
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void put_record(const SeedConst* __restrict__ K, unsigned char* __restrict__ out, unsigned long long i, float px,
                                           float py, float pz, float r, float g, float b, double vx, double vy, double vz, double mass)
{
    unsigned char* rec = out + i * K->stride;
    float* f = reinterpret_cast<float*>(rec);
    f[0] = px; f[1] = py; f[2] = pz;
    f[3] = r; f[4] = g; f[5] = b; f[6] = 1.0f;
    if (K->layout == NB_LAYOUT_LWPARTICLE)
    {
        f[7] = 1.0f;                                   // AddParticleScale(p, 1.0f)
        return;
    }
    f[7] = r; f[8] = g; f[9] = b; f[10] = 1.0f;       // OriginalColour
    double* d = reinterpret_cast<double*>(rec + NB_OFF_VELOCITY);
    d[0] = vx; d[1] = vy; d[2] = vz;
    d[3] = 0.0; d[4] = 0.0; d[5] = 0.0;                // Forces
    d[6] = mass;
}

// GalaxySeeder::AddParticle (GalaxySeeder.cpp:83-106): / scale, Vector3::Transform by the orientation, colour
__device__ __forceinline__ void add_particle(const SeedConst* __restrict__ K, unsigned char* __restrict__ out, unsigned long long i, Lcg& g,
                                             float x, float y, float z, double vx, double vy, double vz, double mass)
{
    const float qx = __fdiv_rn(x, K->scale), qy = __fdiv_rn(y, K->scale), qz = __fdiv_rn(z, K->scale);
    float r[4];
#pragma unroll
    for (int c = 0; c < 4; ++c)
    {
        float s = __fadd_rn(__fmul_rn(qz, K->rot[2][c]), K->rot[3][c]);
        s = __fadd_rn(__fmul_rn(qy, K->rot[1][c]), s);
        s = __fadd_rn(__fmul_rn(qx, K->rot[0][c]), s);
        r[c] = s;
    }
    // Color(DistR(Gen), DistG(Gen), DistB(Gen)): g++ evaluates the arguments right to left
    const float cb = uniform_f(g.next(), K->blue[0], K->blue[1]);
    const float cg = uniform_f(g.next(), K->green[0], K->green[1]);
    const float cr = uniform_f(g.next(), K->red[0], K->red[1]);
    put_record(K, out, i, __fdiv_rn(r[0], r[3]), __fdiv_rn(r[1], r[3]), __fdiv_rn(r[2], r[3]), cr, cg, cb, vx, vy, vz, mass);
}

__device__ __forceinline__ float lerp1(float a, float b, float t) { return __fadd_rn(__fmul_rn(__fsub_rn(b, a), t), a); }

// body j of arm `arm` (GalaxySeeder.cpp:128-141) with its lerp parameters and height already drawn
__device__ __forceinline__ void arm_body(const SeedConst* __restrict__ K, unsigned char* __restrict__ out, int arm, unsigned long long j, Lcg& g,
                                         float tx, float ty, float z)
{
    const SegmentFrame& f = K->seg[arm * 61 + (int)(j / K->per_segment)];
    const float px = __fadd_rn(lerp1(f.sx[0], f.ex[0], tx), lerp1(f.sy[0], f.ey[0], ty));
    const float py = __fadd_rn(lerp1(f.sx[1], f.ex[1], tx), lerp1(f.sy[1], f.ey[1], ty));
    add_particle(K, out, (unsigned long long)arm * K->arm + j, g, px, py, z, (double)f.vel[0], (double)f.vel[1], (double)f.vel[2], 1e20);
}

// One arm unit starting at draw `t`: bodies j and j + 1 of the arm (only j if `single`).
//   hz1 == false:  [ty][distx pairs][distz pairs][colour x3]   [ty][colour x3]
//   hz1 == true :  [ty][distx pairs][colour x3]                [ty][distz pairs][colour x3]
//                  and the first body's height is the variate cached by the pair at draws t-5, t-4.
// Returns the number of draws the unit took.
__device__ int arm_unit(const SeedConst* __restrict__ K, unsigned char* __restrict__ out, int arm, unsigned long long j, bool hz1, bool single,
                        unsigned long long t)
{
    float zsaved = 0.f;
    if (hz1)
    {
        Lcg p(K, t - 5);
        const float x = polar_coord(p.next()), y = polar_coord(p.next());
        float r2, ret;
        polar_rejects(x, y, &r2);
        polar_variates(x, y, r2, &ret, &zsaved);
    }
    Lcg g(K, t);
    // Lerp(sx, ex, distx(Gen)) + Lerp(sy, ey, disty(Gen)): the right operand is evaluated first
    float ty = uniform_f(g.next(), 0.2f, 0.5f);
    float xr, xs;
    normal_generate(g, &xr, &xs);
    float z;
    if (hz1) z = zsaved;
    else
    {
        float zr;
        normal_generate(g, &zr, &zsaved);
        z = zr;
    }
    arm_body(K, out, arm, j, g, normal_apply(xr, 0.5f, 0.2f), ty, normal_apply(z, 0.0f, 16.0f));
    if (single) return g.used;
    ty = uniform_f(g.next(), 0.2f, 0.5f);
    if (hz1)
    {
        float dummy;
        normal_generate(g, &z, &dummy);
    }
    else z = zsaved;
    arm_body(K, out, arm, j + 1, g, normal_apply(xs, 0.5f, 0.2f), ty, normal_apply(z, 0.0f, 16.0f));
    return g.used;
}

// One disk attempt after x, y, z are known (GalaxySeeder.cpp:60-78); returns whether it was accepted.
__device__ __forceinline__ bool disk_inside(float x, float y, float z)
{
    const float dx = __fsub_rn(0.f, x), dy = __fsub_rn(0.f, y), dz = __fsub_rn(0.f, z);
    const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
    return !(d2 > __fmul_rn(720.0f, 720.0f));
}

__device__ __forceinline__ void disk_body(const SeedConst* __restrict__ K, unsigned char* __restrict__ out, unsigned long long i, Lcg& g, float x,
                                          float y, float z)
{
    const unsigned int k1 = g.next(), k2 = g.next(), m1 = g.next(), m2 = g.next();
    if (i >= K->n) { g.skip(3); return; }
    // tangent = Cross(pos - Centre, (0, 0, 1)), component by component without fusion
    const float nx = __fsub_rn(x, 0.f), ny = __fsub_rn(y, 0.f), nz = __fsub_rn(z, 0.f);
    const float tx = __fsub_rn(__fmul_rn(ny, 1.0f), __fmul_rn(nz, 0.0f));
    const float ty = __fsub_rn(__fmul_rn(nz, 0.0f), __fmul_rn(nx, 1.0f));
    const float tz = __fsub_rn(__fmul_rn(nx, 0.0f), __fmul_rn(ny, 0.0f));
    const double k = __dmul_rn(uniform_d(k1, k2, K->canon_div, 0.8, 1.2), 1e14);
    const double mass = uniform_d(m1, m2, K->canon_div, 1e28, 1e30);
    add_particle(K, out, i, g, x, y, z, __dmul_rn((double)tx, k), __dmul_rn((double)ty, k), __dmul_rn((double)tz, k), mass);
}

// One disk unit starting at draw `t`: two attempts, `first` bodies accepted before it.
__device__ void disk_unit(const SeedConst* __restrict__ K, unsigned char* __restrict__ out, unsigned long long first, unsigned long long t)
{
    Lcg g(K, t);
    float x = uniform_f(g.next(), -2000.0f, 2000.0f);
    float y = uniform_f(g.next(), -2000.0f, 2000.0f);
    float zr, zs;
    normal_generate(g, &zr, &zs);
    float z = normal_apply(zr, 0.0f, 16.0f);
    if (disk_inside(x, y, z))
    {
        disk_body(K, out, first, g, x, y, z);
        ++first;
    }
    if (first >= K->n) return;                 // `while (local < n)` ended with the first attempt
    x = uniform_f(g.next(), -2000.0f, 2000.0f);
    y = uniform_f(g.next(), -2000.0f, 2000.0f);
    z = normal_apply(zs, 0.0f, 16.0f);
    if (disk_inside(x, y, z)) disk_body(K, out, first, g, x, y, z);
}

// ------------------------------------------------------------------------------------------------
// parse
// ------------------------------------------------------------------------------------------------
constexpr int SP_CHUNK = 2048;        // stream positions per chunk
constexpr int SP_ENTRIES = 128;       // entry offsets per chunk == upper bound of a unit's length
constexpr int SP_WINDOW = SP_CHUNK + SP_ENTRIES + 8;
constexpr int SP_THREADS = 256;
constexpr int SP_GROUP = 128;         // tables merged per level
enum { UNIT_ARM_HZ0 = 0, UNIT_ARM_HZ1 = 1, UNIT_DISK = 2 };
enum { SE_OK = 0, SE_UNIT_TOO_LONG = 1, SE_RANGE_SHORT = 2 };

struct SeedState
{
    unsigned long long phase_start;    // first draw of the phase being parsed
    unsigned long long next_start;     // first draw after the phase (written by the emit kernel)
    unsigned long long found;          // units (arms) / accepted bodies (disk) inside the parsed range
    int error;
};

// number of draws a polar loop starting at window position p consumes; 0 if it runs out of the window
__device__ __forceinline__ int polar_len(const unsigned int* __restrict__ d, int p, float* x_out, float* y_out, float* r2_out)
{
    for (int q = p; q + 1 < SP_WINDOW; q += 2)
    {
        const float x = polar_coord(d[q]), y = polar_coord(d[q + 1]);
        if (!polar_rejects(x, y, r2_out))
        {
            *x_out = x; *y_out = y;
            return q + 2 - p;
        }
    }
    return 0;
}

// length of the unit starting at window position p and what it counts for (1 unit, or 0..2 accepted bodies);
// 0 if the unit does not fit the window
__device__ int unit_len(int kind, const unsigned int* __restrict__ d, int p, int* counted)
{
    float x, y, r2;
    if (kind != UNIT_DISK)
    {
        *counted = 1;
        const int a = polar_len(d, p + 1, &x, &y, &r2);                           // distx
        if (a == 0) return 0;
        const int zpos = kind == UNIT_ARM_HZ0 ? p + 1 + a : p + 1 + a + 3 + 1;      // distz: first or second body
        if (zpos + 1 >= SP_WINDOW) return 0;
        const int b = polar_len(d, zpos, &x, &y, &r2);
        if (b == 0) return 0;
        return 8 + a + b;
    }
    if (p + 4 >= SP_WINDOW) return 0;
    const float ax = uniform_f(d[p], -2000.0f, 2000.0f), ay = uniform_f(d[p + 1], -2000.0f, 2000.0f);
    const int b = polar_len(d, p + 2, &x, &y, &r2);
    if (b == 0) return 0;
    float zr, zs;
    polar_variates(x, y, r2, &zr, &zs);
    int q = p + 2 + b, c = 0;
    if (disk_inside(ax, ay, normal_apply(zr, 0.0f, 16.0f))) { q += 7; ++c; }
    if (q + 1 >= SP_WINDOW) return 0;
    const float bx = uniform_f(d[q], -2000.0f, 2000.0f), by = uniform_f(d[q + 1], -2000.0f, 2000.0f);
    q += 2;
    if (disk_inside(bx, by, normal_apply(zs, 0.0f, 16.0f))) { q += 7; ++c; }
    *counted = c;
    return q - p;
}

// fills the chunk's draw window and, for every position of the chunk, the length / count of the unit that
// would start there
__device__ void chunk_window(const SeedConst* __restrict__ K, SeedState* __restrict__ S, int kind, unsigned long long chunk, unsigned int* d,
                             unsigned char* len, unsigned char* cnt)
{
    const unsigned long long base = S->phase_start + chunk * SP_CHUNK;
    constexpr int per_thread = (SP_WINDOW + SP_THREADS - 1) / SP_THREADS;
    {
        const int w0 = threadIdx.x * per_thread;
        if (w0 < SP_WINDOW)
        {
            unsigned int x = lcg_at(K, base + w0);
            for (int k = 0; k < per_thread && w0 + k < SP_WINDOW; ++k)
            {
                d[w0 + k] = x;
                x = mulmod(x, 16807u);
            }
        }
    }
    __syncthreads();
    for (int p = threadIdx.x; p < SP_CHUNK; p += SP_THREADS)
    {
        int c = 0;
        const int l = unit_len(kind, d, p, &c);
        if (l == 0 || l > SP_ENTRIES) { atomicExch(&S->error, SE_UNIT_TOO_LONG); len[p] = 1; cnt[p] = 0; }
        else { len[p] = (unsigned char)l; cnt[p] = (unsigned char)c; }
    }
    __syncthreads();
}

// table[chunk][o] = exit offset (low 8 bits) | counted << 8
__global__ void __launch_bounds__(SP_THREADS)
k_seed_tables(const SeedConst* __restrict__ K, SeedState* __restrict__ S, int kind, unsigned int* __restrict__ table)
{
    __shared__ unsigned int d[SP_WINDOW];
    __shared__ unsigned char len[SP_CHUNK], cnt[SP_CHUNK];
    chunk_window(K, S, kind, blockIdx.x, d, len, cnt);
    if (threadIdx.x < SP_ENTRIES)
    {
        int p = threadIdx.x;
        unsigned int c = 0;
        while (p < SP_CHUNK)
        {
            c += cnt[p];
            p += len[p];
        }
        table[(size_t)blockIdx.x * SP_ENTRIES + threadIdx.x] = (unsigned int)(p - SP_CHUNK) | (c << 8);
    }
}

// composes SP_GROUP consecutive tables into one
__global__ void __launch_bounds__(SP_ENTRIES)
k_seed_merge(const unsigned int* __restrict__ in, size_t n_in, unsigned int* __restrict__ out)
{
    const size_t g = blockIdx.x;
    unsigned int o = threadIdx.x, c = 0;
    for (size_t i = g * SP_GROUP; i < (g + 1) * SP_GROUP && i < n_in; ++i)
    {
        const unsigned int e = in[i * SP_ENTRIES + o];
        o = e & 255u;
        c += e >> 8;
    }
    out[g * SP_ENTRIES + threadIdx.x] = o | (c << 8);
}

// top level: a single thread walks the level-2 tables; entry[g] = {offset, counted before}
__global__ void k_seed_top(const unsigned int* __restrict__ table, size_t n, uint2* __restrict__ entry_o, unsigned long long* __restrict__ entry_c,
                           SeedState* __restrict__ S, unsigned long long needed)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    unsigned int o = 0;
    unsigned long long c = 0;
    for (size_t g = 0; g < n; ++g)
    {
        entry_o[g] = make_uint2(o, 0u);
        entry_c[g] = c;
        const unsigned int e = table[g * SP_ENTRIES + o];
        o = e & 255u;
        c += e >> 8;
    }
    S->found = c;
    if (c < needed) atomicExch(&S->error, SE_RANGE_SHORT);
}

// pushes the entries of one level down to the SP_GROUP tables below each
__global__ void __launch_bounds__(128)
k_seed_expand(const unsigned int* __restrict__ table, size_t n_lower, const uint2* __restrict__ up_o, const unsigned long long* __restrict__ up_c,
              size_t n_upper, uint2* __restrict__ lo_o, unsigned long long* __restrict__ lo_c)
{
    const size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_upper) return;
    unsigned int o = up_o[g].x;
    unsigned long long c = up_c[g];
    for (size_t i = g * SP_GROUP; i < (g + 1) * SP_GROUP && i < n_lower; ++i)
    {
        lo_o[i] = make_uint2(o, 0u);
        lo_c[i] = c;
        const unsigned int e = table[i * SP_ENTRIES + o];
        o = e & 255u;
        c += e >> 8;
    }
}

// generation: every chunk walks its units from its entry and hands one unit to each thread
__global__ void __launch_bounds__(SP_THREADS)
k_seed_emit(const SeedConst* __restrict__ K, SeedState* __restrict__ S, int kind, int arm, const uint2* __restrict__ entry_o,
            const unsigned long long* __restrict__ entry_c, unsigned long long needed, unsigned long long first_body,
            unsigned char* __restrict__ out)
{
    // needed: arm phases -- full units of the arm (the odd tail body is generated by the thread of the last unit);
    //         disk -- bodies of the disk.  first_body: index of the disk's first body.
    const unsigned long long before = entry_c[blockIdx.x];
    if (before >= needed) return;
    __shared__ unsigned int d[SP_WINDOW];
    __shared__ unsigned char len[SP_CHUNK], cnt[SP_CHUNK];
    __shared__ unsigned short start[SP_CHUNK / 4 + 2];
    __shared__ unsigned short pref[SP_CHUNK / 4 + 2];
    __shared__ int units;
    chunk_window(K, S, kind, blockIdx.x, d, len, cnt);
    if (threadIdx.x == 0)
    {
        int p = (int)entry_o[blockIdx.x].x, k = 0, c = 0;
        while (p < SP_CHUNK)
        {
            start[k] = (unsigned short)p;
            pref[k] = (unsigned short)c;
            c += cnt[p];
            p += len[p];
            ++k;
        }
        start[k] = (unsigned short)p;          // where the next chunk's first unit starts (may exceed the chunk)
        units = k;
    }
    __syncthreads();
    const unsigned long long base = S->phase_start + (unsigned long long)blockIdx.x * SP_CHUNK;
    for (int k = threadIdx.x; k < units; k += SP_THREADS)
    {
        const unsigned long long idx = before + pref[k];          // units / bodies before this unit
        if (idx >= needed) continue;
        const unsigned long long t = base + start[k];
        if (kind == UNIT_DISK)
        {
            disk_unit(K, out, first_body + idx, t);
            continue;
        }
        const bool hz1 = kind == UNIT_ARM_HZ1;
        arm_unit(K, out, arm, 2 * idx, hz1, false, t);
        if (idx + 1 == needed)
        {
            // last full unit of the arm: the phase ends here, or after the odd tail body
            unsigned long long end = base + start[k + 1];
            if (K->arm & 1ull) end += (unsigned long long)arm_unit(K, out, arm, 2 * needed, hz1, true, end);
            S->next_start = end;
        }
    }
}

// Random (11 draws per body) and StarSystem (11 draws per body after the star): one thread per body
__global__ void __launch_bounds__(256)
k_seed_fixed(const SeedConst* __restrict__ K, int kind, unsigned char* __restrict__ out)
{
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= K->n) return;
    if (kind == NB_SEEDER_RANDOM)
    {
        Lcg g(K, 11ull * i + 1ull);
        unsigned int r[11];
#pragma unroll
        for (int k = 0; k < 11; ++k) r[k] = g.next();
        // uniform_real_distribution<double>(-500.0f, 500.0), narrowed to float, then / Scale
        const float px = __fdiv_rn(__double2float_rn(uniform_d(r[0], r[1], K->canon_div, -500.0, 500.0)), K->scale);
        const float py = __fdiv_rn(__double2float_rn(uniform_d(r[2], r[3], K->canon_div, -500.0, 500.0)), K->scale);
        const float pz = __fdiv_rn(__double2float_rn(uniform_d(r[4], r[5], K->canon_div, -500.0, 500.0)), K->scale);
        const float len = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(px, px), __fmul_rn(py, py)), __fmul_rn(pz, pz)));
        float nx = 0.f, ny = 0.f, nz = 0.f;
        if (isinf(len)) nx = ny = nz = __uint_as_float(0x7fc00000u);
        else if (len != 0.f) { nx = __fdiv_rn(px, len); ny = __fdiv_rn(py, len); nz = __fdiv_rn(pz, len); }
        const double k = (double)10000000000000000.0f;
        const double mass = uniform_d(r[6], r[7], K->canon_div, 1e20, 1e30);
        const float cb = uniform_f(r[8], 0.2f, 1.0f), cg = uniform_f(r[9], 0.2f, 1.0f), cr = uniform_f(r[10], 0.2f, 1.0f);
        put_record(K, out, i, px, py, pz, cr, cg, cb, __dmul_rn((double)nx, k), __dmul_rn((double)ny, k), __dmul_rn((double)nz, k), mass);
        return;
    }
    if (i == 0)
    {
        put_record(K, out, 0, 0.f, 0.f, 0.f, 0.6f, 1.0f, 1.0f, 0.0, 0.0, 0.0, 1e30);
        return;
    }
    const double AU = 1.15e12, M = 1000.0, StarSystemScale = 20 * AU;
    Lcg g(K, 11ull * (i - 1ull) + 1ull);
    unsigned int r[11];
#pragma unroll
    for (int k = 0; k < 11; ++k) r[k] = g.next();
    const float pz = __double2float_rn(__ddiv_rn(uniform_d(r[0], r[1], K->canon_div, 4.0 * AU * M, 7.0 * AU * M), StarSystemScale));
    const double vy = __ddiv_rn(uniform_d(r[2], r[3], K->canon_div, 1 * AU * M, 5 * AU * M), 20.0);
    const double vx = uniform_d(r[4], r[5], K->canon_div, 1 * AU * M, 5 * AU * M);
    const double mass = uniform_d(r[6], r[7], K->canon_div, 1e10, 1e26);
    const float cb = uniform_f(r[8], 0.2f, 1.0f), cg = uniform_f(r[9], 0.2f, 1.0f), cr = uniform_f(r[10], 0.0f, 0.4f);
    put_record(K, out, i, 0.f, 0.f, pz, cr, cg, cb, vx, vy, 0.0, mass);
}

// the two-galaxy scene: shift and approach (nb_seed_collision_host)
__global__ void __launch_bounds__(256)
k_seed_collide(unsigned char* __restrict__ aos, unsigned long long n, unsigned long long half, unsigned long long stride, float separation,
               double approach_speed)
{
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned char* rec = aos + i * stride;
    const float sign = i < half ? -1.0f : 1.0f;
    float* pos = reinterpret_cast<float*>(rec + NB_OFF_POSITION);
    double* vel = reinterpret_cast<double*>(rec + NB_OFF_VELOCITY);
    pos[0] = __fadd_rn(pos[0], __fmul_rn(__fmul_rn(sign, 0.5f), separation));
    vel[0] = __dsub_rn(vel[0], __dmul_rn((double)sign, approach_speed));
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
namespace
{

struct HostLcg
{
    unsigned int x;
    explicit HostLcg(uint64_t seed) { x = (unsigned int)((uint32_t)seed % kLcgM); if (x == 0) x = 1; }
    unsigned int next() { x = mulmod(x, 16807u); return x; }
};

float host_canonical_f(HostLcg& g)
{
    const float sum = (float)(uint64_t)(g.next() - 1u);
    float r = sum / 2147483648.0f;
    if (r >= 1.0f) r = std::nextafter(1.0f, 0.0f);
    return r;
}

// the orientation (GalaxySeeder.cpp:51-53) and the 2 x 61 segment frames (:116-126), operation order of seed_host.cpp
bool host_geometry(uint64_t seed, SeedConst* K)
{
    HostLcg g(seed);
    const float two_pi = 2.0f * 3.141592654f;
    const float roll = host_canonical_f(g) * (two_pi - 0.0f) + 0.0f;
    const float pitch = host_canonical_f(g) * (two_pi - 0.0f) + 0.0f;
    const float yaw = host_canonical_f(g) * (two_pi - 0.0f) + 0.0f;
    const float hp = pitch * 0.5f, hy = yaw * 0.5f, hr = roll * 0.5f;
    const float sp = sinf(hp), cp = cosf(hp), sy = sinf(hy), cy = cosf(hy), sr = sinf(hr), cr = cosf(hr);
    const float qx = (cr * sp) * cy + (sr * cp) * sy;
    const float qy = (cr * cp) * sy - (sr * sp) * cy;
    const float qz = (sr * cp) * cy - (cr * sp) * sy;
    const float qw = (cr * cp) * cy + (sr * sp) * sy;
    const float xx = qx * qx, yy = qy * qy, zz = qz * qz, xy = qx * qy, xz = qx * qz, yz = qy * qz, wx = qw * qx, wy = qw * qy, wz = qw * qz;
    for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c) K->rot[r][c] = (r == c) ? 1.f : 0.f;
    K->rot[0][0] = 1.f - 2.f * (yy + zz); K->rot[0][1] = 2.f * (xy + wz); K->rot[0][2] = 2.f * (xz - wy);
    K->rot[1][0] = 2.f * (xy - wz); K->rot[1][1] = 1.f - 2.f * (xx + zz); K->rot[1][2] = 2.f * (yz + wx);
    K->rot[2][0] = 2.f * (xz + wy); K->rot[2][1] = 2.f * (yz - wx); K->rot[2][2] = 1.f - 2.f * (xx + yy);

    const float offsets[2] = {0.0f, 3.14f};
    for (int arm = 0; arm < 2; ++arm)
    {
        const float offset = offsets[arm];
        int s = 0;
        for (float angle = 0.0f, r = 2.0f; angle < 6.0f; angle += 0.1f, r += 7.2f, ++s)
        {
            if (s >= 61) return false;
            const float spx = cosf(angle + offset) * r, spy = sinf(angle + offset) * r;
            const float snx = cosf(angle + offset + 0.1f) * (r + 10.0f), sny = sinf(angle + offset + 0.1f) * (r + 10.0f);
            float nrm[3] = {spx - snx, spy - sny, 0.0f - 0.0f};
            const float mag = sqrtf((nrm[0] * nrm[0] + nrm[1] * nrm[1]) + nrm[2] * nrm[2]);
            if (mag == 0.f) nrm[0] = nrm[1] = nrm[2] = 0.f;
            else for (int k = 0; k < 3; ++k) nrm[k] = nrm[k] / mag;
            float tan[3] = {nrm[1] * 1.0f - nrm[2] * 0.0f, nrm[2] * 0.0f - nrm[0] * 1.0f, nrm[0] * 0.0f - nrm[1] * 0.0f};
            const float tl = sqrtf((tan[0] * tan[0] + tan[1] * tan[1]) + tan[2] * tan[2]);
            if (tl == 0.f) tan[0] = tan[1] = tan[2] = 0.f;
            else for (int k = 0; k < 3; ++k) tan[k] = tan[k] / tl;
            const float sp3[3] = {spx, spy, 0.0f};
            SegmentFrame& f = K->seg[arm * 61 + s];
            for (int k = 0; k < 3; ++k)
            {
                f.sx[k] = sp3[k] - tan[k] * 140.0f;
                f.ex[k] = sp3[k] + tan[k] * 140.0f;
                f.sy[k] = sp3[k] - nrm[k] * 400.0f;
                f.ey[k] = sp3[k] + nrm[k] * 400.0f;
                f.vel[k] = (nrm[k] * 2e16f) * (1000.0f / mag);
            }
        }
        if (s != 61) return false;            // the float loop `angle < 6.0f; angle += 0.1f` runs 61 times
    }
    return true;
}

void host_constants(uint64_t seed, SeedConst* K)
{
    std::memset(K, 0, sizeof(*K));
    unsigned int p = 16807u;
    for (int i = 0; i < 48; ++i) { K->pow2[i] = p; p = mulmod(p, p); }
    K->x0 = HostLcg(seed).x;
    K->canon_div = (double)((long double)2147483646.0 * 2147483646.0L);
}

inline float clamp01(float v) { return v < 0.0f ? 0.0f : (v > 1.0f ? 1.0f : v); }

struct Scratch
{
    SeedConst* K = nullptr;
    SeedState* S = nullptr;
    unsigned int* table[3] = {nullptr, nullptr, nullptr};
    uint2* eo[3] = {nullptr, nullptr, nullptr};
    unsigned long long* ec[3] = {nullptr, nullptr, nullptr};
    size_t cap0 = 0;
    ~Scratch()
    {
        cudaFree(K); cudaFree(S);
        for (int l = 0; l < 3; ++l) { cudaFree(table[l]); cudaFree(eo[l]); cudaFree(ec[l]); }
    }
    int reserve(size_t chunks)
    {
        if (K == nullptr) { NB_CUDA(cudaMalloc(&K, sizeof(SeedConst))); NB_CUDA(cudaMalloc(&S, sizeof(SeedState))); }
        if (chunks <= cap0) return NB_OK;
        for (int l = 0; l < 3; ++l) { cudaFree(table[l]); cudaFree(eo[l]); cudaFree(ec[l]); table[l] = nullptr; eo[l] = nullptr; ec[l] = nullptr; }
        size_t m = chunks;
        for (int l = 0; l < 3; ++l)
        {
            NB_CUDA(cudaMalloc(&table[l], m * SP_ENTRIES * sizeof(unsigned int)));
            NB_CUDA(cudaMalloc(&eo[l], m * sizeof(uint2)));
            NB_CUDA(cudaMalloc(&ec[l], m * sizeof(unsigned long long)));
            m = (m + SP_GROUP - 1) / SP_GROUP;
        }
        cap0 = chunks;
        return NB_OK;
    }
};

// Parses one phase starting at draw `start` and generates its bodies.  needed: full units (arms) or bodies
// (disk).  Returns the first draw after the phase through *next (arms only).
int run_phase(Scratch& sc, cudaStream_t st, int kind, int arm, unsigned long long start, unsigned long long needed,
              unsigned long long first_body, double positions_per_needed, unsigned char* d_out, unsigned long long* next)
{
    if (needed == 0) { if (next) *next = start; return NB_OK; }
    double estimate = (double)needed * positions_per_needed * 1.03 + 16384.0;
    for (int attempt = 0; attempt < 6; ++attempt, estimate *= 2.0)
    {
        const size_t chunks = (size_t)(estimate / SP_CHUNK) + 1;
        NB_CHECK(sc.reserve(chunks));
        SeedState s0;
        s0.phase_start = start; s0.next_start = 0; s0.found = 0; s0.error = SE_OK;
        NB_CUDA(cudaMemcpyAsync(sc.S, &s0, sizeof(s0), cudaMemcpyHostToDevice, st));
        const size_t n1 = (chunks + SP_GROUP - 1) / SP_GROUP, n2 = (n1 + SP_GROUP - 1) / SP_GROUP;
        k_seed_tables<<<(unsigned int)chunks, SP_THREADS, 0, st>>>(sc.K, sc.S, kind, sc.table[0]);
        k_seed_merge<<<(unsigned int)n1, SP_ENTRIES, 0, st>>>(sc.table[0], chunks, sc.table[1]);
        k_seed_merge<<<(unsigned int)n2, SP_ENTRIES, 0, st>>>(sc.table[1], n1, sc.table[2]);
        k_seed_top<<<1, 32, 0, st>>>(sc.table[2], n2, sc.eo[2], sc.ec[2], sc.S, needed);
        k_seed_expand<<<(unsigned int)((n2 + 127) / 128), 128, 0, st>>>(sc.table[1], n1, sc.eo[2], sc.ec[2], n2, sc.eo[1], sc.ec[1]);
        k_seed_expand<<<(unsigned int)((n1 + 127) / 128), 128, 0, st>>>(sc.table[0], chunks, sc.eo[1], sc.ec[1], n1, sc.eo[0], sc.ec[0]);
        NB_CUDA(cudaGetLastError());
        SeedState s1;
        NB_CUDA(cudaMemcpyAsync(&s1, sc.S, sizeof(s1), cudaMemcpyDeviceToHost, st));
        NB_CUDA(cudaStreamSynchronize(st));
        if (s1.error == SE_UNIT_TOO_LONG)
        {
            set_error("device seeder: a unit of the random stream is longer than %d draws (a polar rejection loop of > 50 rounds)", SP_ENTRIES);
            return NB_ERR_STATE;
        }
        if (s1.error == SE_RANGE_SHORT) continue;       // the estimate of the phase's length was short: parse a longer range
        k_seed_emit<<<(unsigned int)chunks, SP_THREADS, 0, st>>>(sc.K, sc.S, kind, arm, sc.eo[0], sc.ec[0], needed, first_body, d_out);
        NB_CUDA(cudaGetLastError());
        NB_CUDA(cudaMemcpyAsync(&s1, sc.S, sizeof(s1), cudaMemcpyDeviceToHost, st));
        NB_CUDA(cudaStreamSynchronize(st));
        if (s1.error != SE_OK) { set_error("device seeder: generation failed (%d)", s1.error); return NB_ERR_STATE; }
        if (next) *next = s1.next_start;
        return NB_OK;
    }
    set_error("device seeder: the random stream did not yield enough bodies");
    return NB_ERR_STATE;
}

}  // namespace

// `d_out`: device buffer of n records (stride bytes each, zero-initialised by the caller).
int seed_records_device(cudaStream_t st, int kind, unsigned char* d_out, size_t n, size_t stride, uint64_t seed, const nb_seed_options& o)
{
    if (n == 0) return NB_OK;
    Scratch sc;                                   // freed on return: seeding is an Init-time operation
    NB_CHECK(sc.reserve(1));
    SeedConst K;
    // RandomSeeder and StarSystemSeeder construct a fresh default_random_engine and ignore `seed`
    host_constants(kind == NB_SEEDER_GALAXY ? seed : 1u, &K);
    K.scale = o.scale;
    K.red[0] = clamp01(o.red[0]); K.red[1] = clamp01(o.red[1]);
    K.green[0] = clamp01(o.green[0]); K.green[1] = clamp01(o.green[1]);
    K.blue[0] = clamp01(o.blue[0]); K.blue[1] = clamp01(o.blue[1]);
    K.layout = o.layout;
    K.stride = stride;
    K.n = n;
    if (kind != NB_SEEDER_GALAXY)
    {
        NB_CUDA(cudaMemcpyAsync(sc.K, &K, sizeof(K), cudaMemcpyHostToDevice, st));
        k_seed_fixed<<<(unsigned int)((n + 255) / 256), 256, 0, st>>>(sc.K, kind, d_out);
        NB_CUDA(cudaGetLastError());
        NB_CUDA(cudaStreamSynchronize(st));
        return NB_OK;
    }
    if (!host_geometry(seed, &K)) { set_error("device seeder: the spiral-arm loop did not run 61 times"); return NB_ERR_STATE; }
    const float per_loop = std::floor(((float)n * (0.8f / 2)) / 60);
    K.per_segment = per_loop >= 1.0f ? (unsigned long long)per_loop : 0ull;
    K.arm = 61ull * K.per_segment;
    if (K.per_segment == 0) K.per_segment = 1;            // never divided by when arm == 0
    NB_CUDA(cudaMemcpyAsync(sc.K, &K, sizeof(K), cudaMemcpyHostToDevice, st));
    const unsigned long long A = K.arm;
    const bool odd = (A & 1ull) != 0;
    unsigned long long t = 4;                                // draws 1..3 are the orientation
    // arm 1: distz starts without a cached variate; arm 2 starts with one iff arm 1 had an odd number of bodies
    NB_CHECK(run_phase(sc, st, UNIT_ARM_HZ0, 0, t, A / 2, 0, 13.1, d_out, &t));
    NB_CHECK(run_phase(sc, st, odd ? UNIT_ARM_HZ1 : UNIT_ARM_HZ0, 1, t, A / 2, 0, 13.1, d_out, &t));
    // the disk always starts without a cached variate (distz was called 2A times)
    NB_CHECK(run_phase(sc, st, UNIT_DISK, 0, t, (unsigned long long)n - 2 * A, 2 * A, 39.3, d_out, nullptr));
    return NB_OK;
}

int seed_galaxy_device(nb_sim* h, size_t n, uint64_t seed, float scale)
{
    nb_seed_options o;
    nb_seed_default_options(&o);
    o.scale = scale;
    const size_t stride = NB_PARTICLE_STRIDE;
    NB_CHECK(reserve_aos(h, n * stride));
    NB_CUDA(cudaMemsetAsync(h->d_aos, 0, n * stride, h->stream));
    NB_CHECK(seed_records_device(h->stream, NB_SEEDER_GALAXY, static_cast<unsigned char*>(h->d_aos), n, stride, seed, o));
    h->d_aos_stride = stride;
    h->last_launches = 0;
    NB_CHECK(launch_unpack_aos(h, stride, 0, n));
    NB_CUDA(cudaStreamSynchronize(h->stream));
    return NB_OK;
}

int seed_collision_device(nb_sim* h, size_t n, uint64_t seed, float scale, float separation, double approach_speed)
{
    nb_seed_options o;
    nb_seed_default_options(&o);
    o.scale = scale;
    const size_t stride = NB_PARTICLE_STRIDE, half = n / 2;
    unsigned char* aos = nullptr;
    NB_CHECK(reserve_aos(h, n * stride));
    aos = static_cast<unsigned char*>(h->d_aos);
    NB_CUDA(cudaMemsetAsync(aos, 0, n * stride, h->stream));
    NB_CHECK(seed_records_device(h->stream, NB_SEEDER_GALAXY, aos, half, stride, seed, o));
    NB_CHECK(seed_records_device(h->stream, NB_SEEDER_GALAXY, aos + half * stride, n - half, stride, seed + 1, o));
    k_seed_collide<<<(unsigned int)((n + 255) / 256), 256, 0, h->stream>>>(aos, n, half, stride, separation, approach_speed);
    NB_CUDA(cudaGetLastError());
    h->d_aos_stride = stride;
    h->last_launches = 0;
    NB_CHECK(launch_unpack_aos(h, stride, 0, n));
    NB_CUDA(cudaStreamSynchronize(h->stream));
    return NB_OK;
}

__global__ void __launch_bounds__(256) k_scale_masses(unsigned char* __restrict__ aos, unsigned long long n, unsigned long long stride, double factor)
{
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double* m = reinterpret_cast<double*>(aos + i * stride + NB_OFF_MASS);
    *m = __dmul_rn(*m, factor);
}

int launch_scale_masses(nb_sim* h, double factor)
{
    k_scale_masses<<<(unsigned int)((h->n + 255) / 256), 256, 0, h->stream>>>(static_cast<unsigned char*>(h->d_aos), h->n, h->d_aos_stride, factor);
    NB_CUDA(cudaGetLastError());
    return NB_OK;
}

}  // namespace nb

using namespace nb;

extern "C" int nb_seed_device(int kind, int device, void* particles, size_t n, size_t stride, uint64_t seed, const nb_seed_options* opt)
{
    nb_seed_options o;
    nb_seed_default_options(&o);
    if (opt != nullptr)
    {
        NB_REQUIRE(opt->struct_size == sizeof(nb_seed_options), NB_ERR_ARG, "nb_seed_device: nb_seed_options.struct_size mismatch");
        o = *opt;
    }
    const size_t min_stride = (o.layout == NB_LAYOUT_LWPARTICLE) ? NB_LW_PARTICLE_STRIDE : NB_PARTICLE_STRIDE;
    NB_REQUIRE(!(particles == nullptr && n > 0) && (o.layout == NB_LAYOUT_PARTICLE || o.layout == NB_LAYOUT_LWPARTICLE) && stride >= min_stride &&
                   stride % (o.layout == NB_LAYOUT_PARTICLE ? 8 : 4) == 0 && o.scale != 0.0f,
               NB_ERR_ARG, "nb_seed_device: bad argument (null buffer, unknown layout, stride below the record size or misaligned, or zero scale)");
    NB_REQUIRE(kind == NB_SEEDER_RANDOM || kind == NB_SEEDER_GALAXY || kind == NB_SEEDER_STARSYSTEM, NB_ERR_ARG, "nb_seed_device: unknown seeder kind");
    NB_REQUIRE(!(kind == NB_SEEDER_STARSYSTEM && n == 0), NB_ERR_ARG, "nb_seed_device: the star-system seeder needs at least one particle");
    if (n == 0) return NB_OK;
    NB_CUDA(cudaSetDevice(device));
    unsigned char* d = nullptr;
    NB_CUDA(cudaMalloc(&d, n * stride));
    cudaStream_t st = nullptr;
    cudaError_t e = cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaMemsetAsync(d, 0, n * stride, st);
    int rc = NB_OK;
    if (e == cudaSuccess) rc = seed_records_device(st, kind, d, n, stride, seed, o);
    if (e == cudaSuccess && rc == NB_OK) e = cudaMemcpyAsync(particles, d, n * stride, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess && rc == NB_OK) e = cudaStreamSynchronize(st);
    if (st) cudaStreamDestroy(st);
    cudaFree(d);
    NB_CHECK(rc);
    NB_CUDA(e);
    return NB_OK;
}
