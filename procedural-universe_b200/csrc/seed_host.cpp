// Host particle seeders: the procedural initial conditions that feed the simulation
// (IParticleSeeder, reference src/Sim/IParticleSeeder.hpp:12-50): GalaxySeeder, RandomSeeder and
// StarSystemSeeder, each for the reference's two record types (Particle, 104 bytes, and the
// renderer's LWParticle, 32 bytes -- src/Render/Misc/Particle.hpp:8-25).
//
// Follows GalaxySeeder<Particle>::Seed / CreateSpiralArm / AddParticle
// (reference src/Sim/GalaxySeeder.cpp:43-80, 109-143, 83-106) and produces, bit for bit, what the
// reference produces when it is built with g++ / libstdc++ (the oracle build in this repo):
//
//  * std::default_random_engine is implementation-defined; under libstdc++ it is minstd_rand0,
//    x <- 16807 x mod (2^31 - 1).  The distributions are restated below exactly as libstdc++ 13
//    implements them (bits/random.h, bits/random.tcc): generate_canonical with one engine draw for
//    float and two for double, uniform_real = canonical * (b - a) + a, normal = Marsaglia polar
//    with the second variate cached inside the distribution object.
//  * Several draws of the reference sit in expressions whose evaluation order the language leaves
//    open (GalaxySeeder.cpp:53, :95, :137).  The order written out here is the one g++ 13 emits
//    (function arguments and the operands of the overloaded + right to left), verified
//    against the oracle build by tests/test_seeder.py.
//  * Vector math follows DirectXMath's scalar semantics as restated in DESIGN.md ("SimpleMath
//    semantics"): no fused multiply-add anywhere (this file is compiled with -ffp-contract=off).
//
// Serial by construction (one LCG stream with data-dependent draw counts): ~25 engine draws per
// body.  The device seeder in seed_device.cu is the parallel, distribution-equivalent variant.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>

#include "nb_internal.h"

namespace
{

struct V3 { float x, y, z; };

inline V3 sub(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 add(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 mul(V3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline float dot3(V3 a) { return (a.x * a.x + a.y * a.y) + a.z * a.z; }
inline V3 cross(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline V3 lerp(V3 a, V3 b, float t) { return {(b.x - a.x) * t + a.x, (b.y - a.y) * t + a.y, (b.z - a.z) * t + a.z}; }
inline V3 normalize(V3 a)
{
    const float len = sqrtf(dot3(a));
    if (len == 0.f) return {0.f, 0.f, 0.f};
    if (std::isinf(len))
    {
        const float q = std::numeric_limits<float>::quiet_NaN();
        return {q, q, q};
    }
    return {a.x / len, a.y / len, a.z / len};
}

// minstd_rand0
struct Lcg
{
    uint32_t x;
    explicit Lcg(uint32_t seed)
    {
        x = seed % 2147483647u;
        if (x == 0) x = 1;
    }
    uint32_t next()
    {
        x = (uint32_t)(((uint64_t)x * 16807ull) % 2147483647ull);
        return x;
    }
};

// std::generate_canonical<float, 24>: one draw, range 2^31 - 2 rounds to 2^31 in float.
inline float canonical_f(Lcg& g)
{
    const float sum = (float)(uint64_t)(g.next() - 1u);
    float r = sum / 2147483648.0f;
    if (r >= 1.0f) r = std::nextafter(1.0f, 0.0f);
    return r;
}

// std::generate_canonical<double, 53>: two draws.
inline double canonical_d(Lcg& g)
{
    const long double range = 2147483646.0L;
    double sum = 0.0, tmp = 1.0;
    for (int k = 0; k < 2; ++k)
    {
        sum += (double)(uint64_t)(g.next() - 1u) * tmp;
        tmp = (double)((long double)tmp * range);
    }
    double r = sum / tmp;
    if (r >= 1.0) r = std::nextafter(1.0, 0.0);
    return r;
}

inline float uniform_f(Lcg& g, float a, float b) { return canonical_f(g) * (b - a) + a; }
inline double uniform_d(Lcg& g, double a, double b) { return canonical_d(g) * (b - a) + a; }

// std::normal_distribution<float>
struct NormalF
{
    float mean, stddev;
    float saved = 0.f;
    bool have = false;
    NormalF(float m, float s) : mean(m), stddev(s) {}
    float operator()(Lcg& g)
    {
        float ret;
        if (have)
        {
            have = false;
            ret = saved;
        }
        else
        {
            float x, y, r2;
            do
            {
                x = (float)((double)(2.0f * canonical_f(g)) - 1.0);
                y = (float)((double)(2.0f * canonical_f(g)) - 1.0);
                r2 = x * x + y * y;
            } while (r2 > 1.0 || r2 == 0.0);
            const float mult = std::sqrt(-2 * std::log(r2) / r2);
            saved = x * mult;
            have = true;
            ret = y * mult;
        }
        return ret * stddev + mean;
    }
};

struct Rot
{
    float m[4][4];
};

// Matrix::CreateFromYawPitchRoll -> XMMatrixRotationRollPitchYaw (quaternion route), libm sinf/cosf.
Rot yaw_pitch_roll(float yaw, float pitch, float roll)
{
    const float hp = pitch * 0.5f, hy = yaw * 0.5f, hr = roll * 0.5f;
    const float sp = sinf(hp), cp = cosf(hp);
    const float sy = sinf(hy), cy = cosf(hy);
    const float sr = sinf(hr), cr = cosf(hr);
    const float qx = (cr * sp) * cy + (sr * cp) * sy;
    const float qy = (cr * cp) * sy - (sr * sp) * cy;
    const float qz = (sr * cp) * cy - (cr * sp) * sy;
    const float qw = (cr * cp) * cy + (sr * sp) * sy;
    const float xx = qx * qx, yy = qy * qy, zz = qz * qz;
    const float xy = qx * qy, xz = qx * qz, yz = qy * qz;
    const float wx = qw * qx, wy = qw * qy, wz = qw * qz;
    Rot R;
    for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c) R.m[r][c] = (r == c) ? 1.f : 0.f;
    R.m[0][0] = 1.f - 2.f * (yy + zz);
    R.m[0][1] = 2.f * (xy + wz);
    R.m[0][2] = 2.f * (xz - wy);
    R.m[1][0] = 2.f * (xy - wz);
    R.m[1][1] = 1.f - 2.f * (xx + zz);
    R.m[1][2] = 2.f * (yz + wx);
    R.m[2][0] = 2.f * (xz + wy);
    R.m[2][1] = 2.f * (yz - wx);
    R.m[2][2] = 1.f - 2.f * (xx + yy);
    return R;
}

// Vector3::Transform -> XMVector3TransformCoord
V3 transform(V3 v, const Rot& M)
{
    float r[4];
    for (int c = 0; c < 4; ++c)
    {
        float s = v.z * M.m[2][c] + M.m[3][c];
        s = v.y * M.m[1][c] + s;
        s = v.x * M.m[0][c] + s;
        r[c] = s;
    }
    return {r[0] / r[3], r[1] / r[3], r[2] / r[3]};
}

// One output record in either of the reference's layouts.  For LWParticle the velocity, mass,
// forces and original colour setters are no-ops (the SFINAE AddParticle* overloads,
// Particle.hpp:48-82) but their ARGUMENTS are still evaluated, so the random draws are the same.
struct Writer
{
    unsigned char* out;
    size_t stride;
    int layout;

    void put(size_t i, V3 p, const float col[4], const double vel[3], double mass) const
    {
        unsigned char* rec = out + i * stride;
        const float pos3[3] = {p.x, p.y, p.z};
        std::memcpy(rec + 0, pos3, sizeof(pos3));
        std::memcpy(rec + 12, col, 16);
        if (layout == NB_LAYOUT_LWPARTICLE)
        {
            const float one = 1.0f;   // AddParticleScale(p, 1.0f)
            std::memcpy(rec + NB_LW_OFF_SCALE, &one, sizeof(one));
            return;
        }
        const double zero[3] = {0.0, 0.0, 0.0};
        std::memcpy(rec + 28, col, 16);
        std::memcpy(rec + NB_OFF_VELOCITY, vel, 24);
        std::memcpy(rec + NB_OFF_FORCES, zero, sizeof(zero));
        std::memcpy(rec + NB_OFF_MASS, &mass, sizeof(mass));
    }
};

inline float clamp01(float v) { return v < 0.0f ? 0.0f : (v > 1.0f ? 1.0f : v); }

struct Seeder
{
    Writer w;
    size_t n;
    float scale;
    size_t local = 0;
    Lcg gen;
    NormalF distz{0.0f, 16.0f};
    Rot orientation;
    // DistR / DistG / DistB, GalaxySeeder.cpp:14-16; Set{Red,Green,Blue}Dist clamp to [0,1] (:24-41).
    float red[2] = {0.0f, 1.0f}, green[2] = {0.0f, 1.0f}, blue[2] = {0.0f, 1.0f};

    Seeder(void* p, size_t n_, size_t stride_, int layout_, uint64_t seed, float scale_)
        : w{static_cast<unsigned char*>(p), stride_, layout_}, n(n_), scale(scale_), gen((uint32_t)seed)
    {
    }

    // GalaxySeeder.cpp:83-106
    bool add_particle(V3 pos, double vx, double vy, double vz, double mass)
    {
        if (local >= n) return false;
        const V3 scaled = {pos.x / scale, pos.y / scale, pos.z / scale};
        const V3 p = transform(scaled, orientation);
        // Color(DistR(Gen), DistG(Gen), DistB(Gen)): g++ evaluates the arguments right to left.
        const float b = uniform_f(gen, blue[0], blue[1]);
        const float g = uniform_f(gen, green[0], green[1]);
        const float r = uniform_f(gen, red[0], red[1]);
        const float col[4] = {r, g, b, 1.0f};
        const double vel[3] = {vx, vy, vz};
        w.put(local, p, col, vel, mass);
        ++local;
        return true;
    }

    // GalaxySeeder.cpp:109-143
    void spiral_arm(float offset, float dist)
    {
        NormalF distx(0.5f, 0.2f);
        const float maxAngle = 6.0f;
        const int loops = (int)std::floor(maxAngle / 0.1f);
        const float numPerLoop = std::floor(((float)n * dist) / loops);
        for (float angle = 0.0f, r = 2.0f; angle < 6.0f; angle += 0.1f, r += 7.2f)
        {
            const V3 spiral = {cosf(angle + offset) * r, sinf(angle + offset) * r, 0.0f};
            const V3 spiraln = {cosf(angle + offset + 0.1f) * (r + 10.0f), sinf(angle + offset + 0.1f) * (r + 10.0f), 0.0f};
            V3 normal = sub(spiral, spiraln);
            const float mag = sqrtf(dot3(normal));
            normal = normalize(normal);
            V3 tangent = cross(normal, V3{0.0f, 0.0f, 1.0f});
            tangent = normalize(tangent);
            const V3 sx = sub(spiral, mul(tangent, 140.0f));
            const V3 ex = add(spiral, mul(tangent, 140.0f));
            const V3 sy = sub(spiral, mul(normal, 400.0f));
            const V3 ey = add(spiral, mul(normal, 400.0f));
            for (int i = 0; i < numPerLoop; ++i)
            {
                // Lerp(sx, ex, distx(Gen)) + Lerp(sy, ey, disty(Gen)): g++ evaluates the right operand first.
                const float ty = uniform_f(gen, 0.2f, 0.5f);
                const float tx = distx(gen);
                V3 position = add(lerp(sx, ex, tx), lerp(sy, ey, ty));
                const V3 velocity = mul(mul(normal, 2e16f), 1000.0f / mag);
                position.z = distz(gen);
                add_particle(position, (double)velocity.x, (double)velocity.y, (double)velocity.z, 1e20);
            }
        }
    }

    // GalaxySeeder.cpp:43-80
    void seed()
    {
        const float two_pi = 2.0f * 3.141592654f;
        // CreateFromYawPitchRoll(dist_rot(Gen), dist_rot(Gen), dist_rot(Gen)): right to left.
        const float roll = uniform_f(gen, 0.0f, two_pi);
        const float pitch = uniform_f(gen, 0.0f, two_pi);
        const float yaw = uniform_f(gen, 0.0f, two_pi);
        orientation = yaw_pitch_roll(yaw, pitch, roll);

        const float arm_dist = 0.8f;
        spiral_arm(0.0f, arm_dist / 2);
        spiral_arm(3.14f, arm_dist / 2);

        while (local < n)
        {
            V3 pos;
            pos.x = uniform_f(gen, -2000.0f, 2000.0f);
            pos.y = uniform_f(gen, -2000.0f, 2000.0f);
            pos.z = distz(gen);
            // DistanceSquared(pos, Centre) with Centre = 0: d = Centre - pos
            const V3 d = {0.f - pos.x, 0.f - pos.y, 0.f - pos.z};
            if (dot3(d) > 720.0f * 720.0f) continue;
            const V3 norm = sub(pos, V3{0.f, 0.f, 0.f});
            const V3 tangent = cross(norm, V3{0.0f, 0.0f, 1.0f});
            const double k = uniform_d(gen, 0.8, 1.2) * 1e14;
            const double vx = (double)tangent.x * k, vy = (double)tangent.y * k, vz = (double)tangent.z * k;
            const double mass = uniform_d(gen, 1e28, 1e30);
            if (!add_particle(pos, vx, vy, vz, mass)) break;
        }
    }
};


// RandomSeeder<T>::Seed, reference src/Sim/RandomSeeder.cpp:13-40.  The reference constructs a
// fresh default_random_engine and never uses its `seed` argument; so does this.
void seed_random(const Writer& w, size_t n, float scale)
{
    Lcg gen(1u);   // std::default_random_engine{} == minstd_rand0 with default_seed 1
    for (size_t i = 0; i < n; ++i)
    {
        // uniform_real_distribution<double>(-500.0f, 500.0), narrowed to float, then / Scale
        V3 p;
        p.x = (float)uniform_d(gen, -500.0, 500.0) / scale;
        p.y = (float)uniform_d(gen, -500.0, 500.0) / scale;
        p.z = (float)uniform_d(gen, -500.0, 500.0) / scale;
        const V3 nrm = normalize(p);
        // Vec3d vel(normal); vel *= 10000000000000000.0f  (double * float-literal-as-double)
        const double k = (double)10000000000000000.0f;
        const double vel[3] = {(double)nrm.x * k, (double)nrm.y * k, (double)nrm.z * k};
        const double mass = uniform_d(gen, 1e20, 1e30);
        // Color(dist_col, dist_col, dist_col, 1.0f): arguments right to left under g++.
        const float b = uniform_f(gen, 0.2f, 1.0f);
        const float g = uniform_f(gen, 0.2f, 1.0f);
        const float r = uniform_f(gen, 0.2f, 1.0f);
        const float col[4] = {r, g, b, 1.0f};
        w.put(i, p, col, vel, mass);
    }
}

// StarSystemSeeder<T>::Seed, reference src/Sim/StarSystemSeeder.cpp:18-55: a 1e30 star at the
// origin and n-1 bodies strung along +z with velocities in the xy plane.  Ignores `seed` and
// `scale` like the reference.
void seed_starsystem(const Writer& w, size_t n)
{
    const double AU = 1.15e12, M = 1000.0, StarSystemScale = 20 * AU;   // Physics.hpp:11-16
    {
        const float col[4] = {0.6f, 1.0f, 1.0f, 1.0f};
        const double vel[3] = {0.0, 0.0, 0.0};
        w.put(0, V3{0.f, 0.f, 0.f}, col, vel, 1e30);
    }
    Lcg gen(1u);
    const double rad_lo = 4.0 * AU * M, rad_hi = 7.0 * AU * M;
    const double vel_lo = 1 * AU * M, vel_hi = 5 * AU * M;
    for (size_t i = 1; i < n; ++i)
    {
        V3 p = {0.f, 0.f, 0.f};
        p.z = (float)(uniform_d(gen, rad_lo, rad_hi) / StarSystemScale);
        // Vec3d vel(dist_vel(gen), dist_vel(gen) / 20.0, 0): constructor arguments right to left.
        const double vy = uniform_d(gen, vel_lo, vel_hi) / 20.0;
        const double vx = uniform_d(gen, vel_lo, vel_hi);
        const double vel[3] = {vx, vy, 0.0};
        const double mass = uniform_d(gen, 1e10, 1e26);
        // Color(dist_red, dist_col, dist_col, 1.0f), right to left.
        const float b = uniform_f(gen, 0.2f, 1.0f);
        const float g = uniform_f(gen, 0.2f, 1.0f);
        const float r = uniform_f(gen, 0.0f, 0.4f);
        const float col[4] = {r, g, b, 1.0f};
        w.put(i, p, col, vel, mass);
    }
}

}  // namespace

namespace nb
{

int seed_galaxy_host(void* particles, size_t n, size_t stride, uint64_t seed, float scale)
{
    Seeder s(particles, n, stride, NB_LAYOUT_PARTICLE, seed, scale);
    s.seed();
    return NB_OK;
}

}  // namespace nb

extern "C" int nb_seed_default_options(nb_seed_options* opt)
{
    if (opt == nullptr)
    {
        nb::set_error("nb_seed_default_options: null argument");
        return NB_ERR_ARG;
    }
    std::memset(opt, 0, sizeof(*opt));
    opt->struct_size = (uint32_t)sizeof(*opt);
    opt->layout = NB_LAYOUT_PARTICLE;
    opt->scale = 1.0f;
    opt->red[1] = opt->green[1] = opt->blue[1] = 1.0f;
    return NB_OK;
}

extern "C" int nb_seed_host(int kind, void* particles, size_t n, size_t stride, uint64_t seed,
                            const nb_seed_options* opt)
{
    nb_seed_options o;
    nb_seed_default_options(&o);
    if (opt != nullptr)
    {
        if (opt->struct_size != sizeof(nb_seed_options))
        {
            nb::set_error("nb_seed_host: nb_seed_options.struct_size mismatch");
            return NB_ERR_ARG;
        }
        o = *opt;
    }
    const size_t min_stride = (o.layout == NB_LAYOUT_LWPARTICLE) ? NB_LW_PARTICLE_STRIDE : NB_PARTICLE_STRIDE;
    if ((particles == nullptr && n > 0) || (o.layout != NB_LAYOUT_PARTICLE && o.layout != NB_LAYOUT_LWPARTICLE) ||
        stride < min_stride || stride % 4 != 0 || !(o.scale != 0.0f))
    {
        nb::set_error("nb_seed_host: bad argument (null buffer, unknown layout, stride below the record size, or zero scale)");
        return NB_ERR_ARG;
    }
    const Writer w{static_cast<unsigned char*>(particles), stride, o.layout};
    switch (kind)
    {
        case NB_SEEDER_RANDOM:
            seed_random(w, n, o.scale);
            return NB_OK;
        case NB_SEEDER_GALAXY:
        {
            Seeder s(particles, n, stride, o.layout, seed, o.scale);
            s.red[0] = clamp01(o.red[0]), s.red[1] = clamp01(o.red[1]);
            s.green[0] = clamp01(o.green[0]), s.green[1] = clamp01(o.green[1]);
            s.blue[0] = clamp01(o.blue[0]), s.blue[1] = clamp01(o.blue[1]);
            s.seed();
            return NB_OK;
        }
        case NB_SEEDER_STARSYSTEM:
            if (n == 0)
            {
                // the reference writes Particles[0] unconditionally (StarSystemSeeder.cpp:20)
                nb::set_error("nb_seed_host: the star-system seeder needs at least one particle");
                return NB_ERR_ARG;
            }
            seed_starsystem(w, n);
            return NB_OK;
        default:
            nb::set_error("nb_seed_host: unknown seeder kind");
            return NB_ERR_ARG;
    }
}

extern "C" int nb_seed_collision_host(void* particles, size_t n, size_t stride, uint64_t seed, float scale,
                                      float separation, double approach_speed)
{
    if (particles == nullptr || stride < NB_PARTICLE_STRIDE || stride % 8 != 0 || n < 2)
    {
        nb::set_error("nb_seed_collision_host: bad argument");
        return NB_ERR_ARG;
    }
    unsigned char* base = static_cast<unsigned char*>(particles);
    const size_t half = n / 2;
    nb::seed_galaxy_host(base, half, stride, seed, scale);
    nb::seed_galaxy_host(base + half * stride, n - half, stride, seed + 1, scale);
    for (size_t i = 0; i < n; ++i)
    {
        unsigned char* rec = base + i * stride;
        const float sign = (i < half) ? -1.0f : 1.0f;
        float pos[3];
        double vel[3];
        std::memcpy(pos, rec + NB_OFF_POSITION, sizeof(pos));
        std::memcpy(vel, rec + NB_OFF_VELOCITY, sizeof(vel));
        pos[0] += sign * 0.5f * separation;
        vel[0] -= (double)sign * approach_speed;
        std::memcpy(rec + NB_OFF_POSITION, pos, sizeof(pos));
        std::memcpy(rec + NB_OFF_VELOCITY, vel, sizeof(vel));
    }
    return NB_OK;
}
