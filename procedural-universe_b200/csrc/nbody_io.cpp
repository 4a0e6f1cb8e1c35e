// `.nbody` particle files -- the reference's checkpoint format, byte for byte.
//
// Writer: SimulationState::RunSimulation dumps its std::vector<Particle> record by record with
// file.write(&p, sizeof(p)) (reference src/States/Simulation/SimulationState.cpp:317-331): no
// header, no count, 104 bytes per body including colours, forces and the 4 padding bytes at 44.
// Reader: SimulationState::InitParticlesFromFile (:229-277) reads whole records until the stream
// fails (a trailing partial record is dropped), then recentres on the centre of mass:
//     long double TotalMass += Mass;   Vec3d CentreOfMass += Vec3d(pos) * Mass;    (:252-260)
//     CentreOfMass /= TotalMass;       -> Vec3<double>::operator/=(double), Vec3.hpp:80-84
//     Position -= Vector3((float)c.x, (float)c.y, (float)c.z)                      (:262-270)
// TotalMass is `long double` (80-bit under g++ on x86-64, the build the parity oracle uses; 64-bit
// under MSVC) and is converted to double by the /= call.  Pure host I/O: no device work here.
#include <cerrno>
#include <cstdio>
#include <cstring>

#include "nb_internal.h"

extern "C" int nb_nbody_recentre(void* particles, size_t n, size_t stride)
{
    if (particles == nullptr || stride < NB_PARTICLE_STRIDE || stride % 8 != 0)
    {
        nb::set_error("nb_nbody_recentre: bad argument");
        return NB_ERR_ARG;
    }
    unsigned char* base = static_cast<unsigned char*>(particles);
    long double total = 0.0L;
    double cx = 0.0, cy = 0.0, cz = 0.0;
    for (size_t i = 0; i < n; ++i)
    {
        float pos[3];
        double mass;
        std::memcpy(pos, base + i * stride + NB_OFF_POSITION, sizeof(pos));
        std::memcpy(&mass, base + i * stride + NB_OFF_MASS, sizeof(mass));
        total += mass;
        cx += (double)pos[0] * mass;
        cy += (double)pos[1] * mass;
        cz += (double)pos[2] * mass;
    }
    const double t = (double)total;
    cx /= t; cy /= t; cz /= t;          // n == 0 or zero total mass: NaN, as in the reference
    const float c[3] = {(float)cx, (float)cy, (float)cz};
    for (size_t i = 0; i < n; ++i)
    {
        float pos[3];
        std::memcpy(pos, base + i * stride + NB_OFF_POSITION, sizeof(pos));
        pos[0] -= c[0]; pos[1] -= c[1]; pos[2] -= c[2];
        std::memcpy(base + i * stride + NB_OFF_POSITION, pos, sizeof(pos));
    }
    return NB_OK;
}

extern "C" int nb_nbody_save(const char* path, const void* particles, size_t n, size_t stride)
{
    if (path == nullptr || (particles == nullptr && n > 0) || stride < NB_PARTICLE_STRIDE)
    {
        nb::set_error("nb_nbody_save: bad argument");
        return NB_ERR_ARG;
    }
    std::FILE* f = std::fopen(path, "wb");
    if (!f)
    {
        nb::set_error("nb_nbody_save: cannot open %s: %s", path, std::strerror(errno));
        return NB_ERR_ARG;
    }
    const unsigned char* base = static_cast<const unsigned char*>(particles);
    bool ok = true;
    if (stride == NB_PARTICLE_STRIDE)
        ok = std::fwrite(base, NB_PARTICLE_STRIDE, n, f) == n;
    else
        for (size_t i = 0; i < n && ok; ++i) ok = std::fwrite(base + i * stride, NB_PARTICLE_STRIDE, 1, f) == 1;
    if (std::fclose(f) != 0) ok = false;
    if (!ok)
    {
        nb::set_error("nb_nbody_save: short write to %s", path);
        return NB_ERR_ARG;
    }
    return NB_OK;
}

extern "C" int nb_nbody_count(const char* path, size_t* n)
{
    if (path == nullptr || n == nullptr)
    {
        nb::set_error("nb_nbody_count: null argument");
        return NB_ERR_ARG;
    }
    std::FILE* f = std::fopen(path, "rb");
    if (!f)
    {
        nb::set_error("Could not read particle file %s", path);
        return NB_ERR_ARG;
    }
    std::fseek(f, 0, SEEK_END);
    const long long bytes = std::ftell(f);
    std::fclose(f);
    *n = bytes < 0 ? 0 : (size_t)bytes / NB_PARTICLE_STRIDE;     // whole records only
    return NB_OK;
}

extern "C" int nb_nbody_load(const char* path, void* particles, size_t capacity, size_t stride, size_t* n_read, int recentre)
{
    if (path == nullptr || (particles == nullptr && capacity > 0) || stride < NB_PARTICLE_STRIDE || stride % 8 != 0)
    {
        nb::set_error("nb_nbody_load: bad argument");
        return NB_ERR_ARG;
    }
    std::FILE* f = std::fopen(path, "rb");
    if (!f)
    {
        nb::set_error("Could not read particle file %s", path);      // the reference's LOGE text (:235)
        return NB_ERR_ARG;
    }
    unsigned char* base = static_cast<unsigned char*>(particles);
    size_t got = 0;
    if (stride == NB_PARTICLE_STRIDE)
        got = std::fread(base, NB_PARTICLE_STRIDE, capacity, f);
    else
        while (got < capacity && std::fread(base + got * stride, NB_PARTICLE_STRIDE, 1, f) == 1) ++got;
    std::fclose(f);
    if (n_read) *n_read = got;
    if (recentre) return nb_nbody_recentre(particles, got, stride);
    return NB_OK;
}
