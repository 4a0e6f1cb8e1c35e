// Test driver for the C++ adapter (procedural-universe_b200/host/B200Sim).  TEST INFRASTRUCTURE.
//
// Compiled together with the reference's own src/Sim sources and the adapter into
// oracle/_ref/libb200_adapter_test.so (oracle/build_ref.sh).  It drives BOTH the reference sims and
// the B200 adapter through the reference's own INBodySim interface and factory names, exactly as
// SimulationState does (Init on a std::vector<Particle>, then Update(dt) per frame,
// SimulationState.cpp:52-53, 218-227), so the parity test reads like the reference's usage.
#include <cstdint>
#include <cstdlib>
#include <chrono>
#include <cstring>
#include <new>
#include <iostream>
#include <sstream>
#include <vector>

#include "Sim/INBodySim.hpp"
#include "Sim/BruteForceCPU.hpp"
#include "Sim/BarnesHut.hpp"
#include "Sim/Octree.hpp"
#include "Core/Event.hpp"

#include "B200Sim.hpp"

// INBodySim.cpp itself cannot be compiled headless (it pulls in BruteForceGPU / D3D); this is the
// same switch without the D3D sim.
// Zeroed storage: CThreadPool starts its workers before it clears HaveWork[] (ThreadPool.hpp:48-49),
// so on a recycled heap block a worker can run on garbage.  The sims are released, never deleted.
template <class Sim>
static Sim* NewInZeroedStorage(ID3D11DeviceContext* context)
{
    return new (std::calloc(1, sizeof(Sim))) Sim(context);
}

std::unique_ptr<INBodySim> CreateNBodySim(ID3D11DeviceContext* context, ENBodySim type)
{
    switch (type)
    {
        case ENBodySim::BruteForceCPU: return std::unique_ptr<INBodySim>(NewInZeroedStorage<BruteForceCPU>(context));
        case ENBodySim::BarnesHut:     return std::unique_ptr<INBodySim>(NewInZeroedStorage<BarnesHut>(context));
        default:                       return nullptr;
    }
}

namespace
{
    struct Quiet
    {
        std::streambuf* old;
        std::ostringstream sink;
        Quiet() : old(std::cout.rdbuf(sink.rdbuf())) {}
        ~Quiet() { std::cout.rdbuf(old); }
        std::string text() { return sink.str(); }
    };
}

extern "C"
{
    // impl: 0 = reference sim from CreateNBodySim, 1 = CreateB200NBodySim.
    // type: ENBodySim value (0 BruteForceCPU, 1 BruteForceGPU, 2 BarnesHut).
    // Returns 0 on success; the constructor's log line is copied to log_out (if given).
    int adapter_run(void* aos, size_t n, int impl, int type, float dt, int steps, float theta, int recolour,
                    char* log_out, size_t log_cap)
    {
        std::vector<Particle> particles(n);
        std::memcpy(static_cast<void*>(particles.data()), aos, n * sizeof(Particle));
        Octree::Theta = theta;
        std::unique_ptr<INBodySim> sim;
        std::string log;
        {
            Quiet q;
            sim = impl == 0 ? CreateNBodySim(nullptr, static_cast<ENBodySim>(type))
                            : CreateB200NBodySim(nullptr, static_cast<ENBodySim>(type));
            log = q.text();
        }
        if (log_out && log_cap) { std::strncpy(log_out, log.c_str(), log_cap - 1); log_out[log_cap - 1] = 0; }
        if (!sim) return 1;
        if (impl == 0)
        {
            // reference pools: keep W | n -- reach into the known concrete types
            if (type == 0) static_cast<BruteForceCPU*>(sim.get())->Pool.SetNumWorkers(n % 4 == 0 ? 4 : 1);
            if (type == 2) static_cast<BarnesHut*>(sim.get())->Pool.SetNumWorkers(n % 4 == 0 ? 4 : 1);
        }
        {
            // theta through the reference's own event plumbing, after construction
            FloatEventData ev(theta);
            EventStream::Report(EEvent::BHThetaChanged, ev);
        }
        sim->Init(particles);
        for (int s = 0; s < steps; ++s)
        {
            if (recolour) particles[s % n].Colour = DirectX::SimpleMath::Color(0.5f, 0.25f, 0.125f);   // UI picking
            sim->Update(dt);
        }
        std::memcpy(aos, particles.data(), n * sizeof(Particle));
        if (impl == 1) static_cast<B200Sim*>(sim.get())->Shutdown();
        sim.release();   // never destroy a reference sim on glibc (~CThreadPool deadlocks)
        return 0;
    }

    // SimulationState::RunBenchmark's protocol (SimulationState.cpp:334-362) for ONE sim: create it through the
    // factory, Init, then `frames` x Update(1.0f) timed with a steady clock; returns ms per frame (< 0 on failure).
    // workers: pool size of the reference sims (0 = leave the reference's default, hardware_concurrency() - 1).
    double adapter_benchmark(void* aos, size_t n, int impl, int type, int frames, int workers, float theta, float dt)
    {
        std::vector<Particle> particles(n);
        std::memcpy(static_cast<void*>(particles.data()), aos, n * sizeof(Particle));
        Octree::Theta = theta;
        std::unique_ptr<INBodySim> sim;
        {
            Quiet q;
            sim = impl == 0 ? CreateNBodySim(nullptr, static_cast<ENBodySim>(type))
                            : CreateB200NBodySim(nullptr, static_cast<ENBodySim>(type));
        }
        if (!sim) return -1.0;
        if (impl == 0 && workers > 0)
        {
            if (type == 0) static_cast<BruteForceCPU*>(sim.get())->Pool.SetNumWorkers(workers);
            if (type == 2) static_cast<BarnesHut*>(sim.get())->Pool.SetNumWorkers(workers);
        }
        {
            FloatEventData ev(theta);
            EventStream::Report(EEvent::BHThetaChanged, ev);
        }
        sim->Init(particles);
        const auto t0 = std::chrono::steady_clock::now();
        for (int f = 0; f < frames; ++f) sim->Update(dt);   // RunBenchmark passes 1.0f
        const auto t1 = std::chrono::steady_clock::now();
        if (impl == 1) static_cast<B200Sim*>(sim.get())->Shutdown();
        sim.release();   // never destroy a reference sim on glibc (~CThreadPool deadlocks)
        return std::chrono::duration<double, std::milli>(t1 - t0).count() / frames;
    }

    // Sim->RenderDebug(view, proj) through the interface (SimulationState.cpp:81) after `steps` Updates,
    // on a sim that was given a (never dereferenced) D3D context so that it owns a debug Cube.  Returns
    // the number of cubes drawn; the first `cap` {x, y, z, size} records are copied to out4.
    long adapter_debug_cubes(void* aos, size_t n, int impl, float dt, int steps, float theta, float* out4, size_t cap)
    {
        std::vector<Particle> particles(n);
        std::memcpy(static_cast<void*>(particles.data()), aos, n * sizeof(Particle));
        Octree::Theta = theta;
        ID3D11DeviceContext* context = reinterpret_cast<ID3D11DeviceContext*>(static_cast<uintptr_t>(16));
        std::unique_ptr<INBodySim> sim;
        {
            Quiet q;
            sim = impl == 0 ? CreateNBodySim(context, ENBodySim::BarnesHut) : CreateB200NBodySim(context, ENBodySim::BarnesHut);
        }
        if (!sim) return -1;
        if (impl == 0) static_cast<BarnesHut*>(sim.get())->Pool.SetNumWorkers(n % 4 == 0 ? 4 : 1);
        sim->Init(particles);
        for (int s = 0; s < steps; ++s) sim->Update(dt);
        Cube::Drawn().clear();
        sim->RenderDebug(DirectX::SimpleMath::Matrix(), DirectX::SimpleMath::Matrix());
        const std::vector<float>& d = Cube::Drawn();
        const size_t cubes = d.size() / 4;
        if (out4 && cap) std::memcpy(out4, d.data(), sizeof(float) * 4 * (cubes < cap ? cubes : cap));
        if (impl == 1) static_cast<B200Sim*>(sim.get())->Shutdown();
        sim.release();
        return static_cast<long>(cubes);
    }
}
