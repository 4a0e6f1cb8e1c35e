// C-ABI driver around the reference's own CPU simulation path.  TEST INFRASTRUCTURE ONLY.
//
// This translation unit is compiled together with the reference's unmodified
// src/Sim/{BruteForceCPU,BarnesHut,Octree}.cpp, src/Services/Log.cpp and src/Core/Event.cpp
// (taken where they lie under /root/reference, never copied into this repository) against
// the headless stand-in headers in oracle/ref_shim/.  The result, oracle/_ref/libpu_ref.so,
// is the parity oracle "kind: reference": every number it returns was computed by the
// reference's code.  Nothing here is shipped or measured as the product; only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs load it.
//
// Hazards of the reference that this driver works around:
//   * sims are heap-allocated once and never destroyed (~CThreadPool blocks forever in
//     pthread_cond_destroy on glibc, src/Core/ThreadPool.hpp:54-60);
//   * the worker count is always set so that it divides N, which keeps the remainder
//     mis-indexing of BruteForceCPU.cpp:56-57 / BarnesHut.cpp:70-77 dormant;
//   * private members are reached with -fno-access-control (Exec, Pool, Tree, Children...).
#include <cstdint>
#include <cstdlib>
#include <new>
#include <cstring>
#include <vector>
#include <chrono>
#include <thread>
#include <iostream>
#include <sstream>

#include "Sim/BruteForceCPU.hpp"
#include "Sim/BarnesHut.hpp"
#include "Sim/Octree.hpp"
#include "Sim/Physics.hpp"
#include "Sim/IParticleSeeder.hpp"
#include "Core/Event.hpp"
#include "Core/Maths.hpp"

const DirectX::SimpleMath::Vector3 DirectX::SimpleMath::Vector3::Zero;

namespace
{
    // The sim constructors log "[Info] <name>" to stdout (BruteForceCPU.cpp:17,
    // BarnesHut.cpp:12); keep that out of the callers' JSON output.
    struct CoutSilencer
    {
        std::streambuf* old;
        std::ostringstream sink;
        CoutSilencer() : old(std::cout.rdbuf(sink.rdbuf())) {}
        ~CoutSilencer() { std::cout.rdbuf(old); }
    };

    uint32_t g_bruteSpawned = 0, g_treeSpawned = 0;   // worker threads the pool actually started

    // CThreadPool's constructor starts worker i BEFORE it writes HaveWork[i] = false
    // (ThreadPool.hpp:48-49): on a recycled heap block the worker can read a stale `true` and run
    // the user function on an uninitialised work item (seen as a segfault in BruteForceCPU::Exec
    // once torch had churned the heap).  Constructing the sim in zeroed storage makes the flag
    // false before any worker exists; the sims are never freed (see ~CThreadPool above).
    template <class Sim>
    Sim* NewInZeroedStorage()
    {
        void* mem = std::calloc(1, sizeof(Sim));
        return new (mem) Sim(nullptr);
    }

    BruteForceCPU* BruteSim()
    {
        static BruteForceCPU* sim = nullptr;
        if (!sim)
        {
            CoutSilencer q;
            sim = NewInZeroedStorage<BruteForceCPU>();
            g_bruteSpawned = sim->Pool.GetNumWorkers();
        }
        return sim;
    }

    BarnesHut* TreeSim()
    {
        static BarnesHut* sim = nullptr;
        if (!sim)
        {
            CoutSilencer q;
            sim = NewInZeroedStorage<BarnesHut>();
            g_treeSpawned = sim->Pool.GetNumWorkers();
        }
        return sim;
    }

    // Largest power of two <= min(request, pool capacity, n).
    uint32_t PickWorkers(int request, uint32_t spawned, size_t n)
    {
        uint32_t cap = spawned;
        if (request > 0 && static_cast<uint32_t>(request) < cap) cap = request;
        if (cap < 1) cap = 1;
        uint32_t w = 1;
        while (w * 2 <= cap && (n % (w * 2)) == 0) w *= 2;
        return w;
    }

    std::vector<Particle> ToVector(const void* aos, size_t n)
    {
        std::vector<Particle> v(n);
        if (n) std::memcpy(static_cast<void*>(v.data()), aos, n * sizeof(Particle));
        return v;
    }

    void FromVector(const std::vector<Particle>& v, void* aos)
    {
        if (!v.empty()) std::memcpy(aos, v.data(), v.size() * sizeof(Particle));
    }

    template <class T>
    void SeedWithColours(std::vector<T>& v, int kind, uint64_t seed, float scale, const float* rgb6)
    {
        auto seeder = CreateParticleSeeder(v, static_cast<EParticleSeeder>(kind), scale);
        if (rgb6)
        {
            // IParticleSeeder::Set{Red,Green,Blue}Dist (IParticleSeeder.hpp:24-26; only GalaxySeeder
            // overrides them, GalaxySeeder.cpp:24-41)
            seeder->SetRedDist(rgb6[0], rgb6[1]);
            seeder->SetGreenDist(rgb6[2], rgb6[3]);
            seeder->SetBlueDist(rgb6[4], rgb6[5]);
        }
        seeder->Seed(seed);
    }

    std::unique_ptr<Octree> BuildTree(std::vector<Particle>& v, double theta)
    {
        // Same sequence as BarnesHut::Update, src/Sim/BarnesHut.cpp:46-56, with the bounds of
        // BarnesHut::BarnesHut, :14-19.
        Octree::Theta = theta;
        const float size = 4000.0f;
        BoundingCube bounds = { { -size, -size, -size }, { +size, +size, +size } };
        std::unique_ptr<Octree> tree = std::make_unique<Octree>(bounds);
        Particle* p = v.data();
        for (size_t i = 0; i < v.size(); ++i, ++p) tree->Add(p);
        tree->CalculateMass();
        return tree;
    }

    void CountNodes(Octree* t, int64_t* nodes, int64_t* internal, int* maxDepth)
    {
        ++*nodes;
        if (t->Depth > *maxDepth) *maxDepth = t->Depth;
        if (!t->IsLeaf)
        {
            ++*internal;
            for (auto& c : t->Children) CountNodes(c.get(), nodes, internal, maxDepth);
        }
    }

    // Instrumented restatement of the control flow of Octree::CalculateForce
    // (src/Sim/Octree.cpp:107-145) that only counts evaluations.
    void CountEvals(Octree* t, Particle* p, int64_t* cellEvals, int64_t* leafEvals, int64_t* visits)
    {
        ++*visits;
        if (t->NumParticles == 1)
        {
            if (p != t->P && !t->Bounds.Contains(p)) ++*leafEvals;
        }
        else
        {
            float r = (p->Position - t->CentreOfMass).Length();
            float d = t->Bounds.BottomRight.x - t->Bounds.TopLeft.x;
            if (d / r < Octree::Theta) ++*cellEvals;
            else if (!t->IsLeaf)
                for (auto& c : t->Children) CountEvals(c.get(), p, cellEvals, leafEvals, visits);
        }
    }
}

extern "C"
{
    int ref_sizeof_particle() { return static_cast<int>(sizeof(Particle)); }

    void ref_particle_offsets(int* out5)
    {
        out5[0] = offsetof(Particle, Position);
        out5[1] = offsetof(Particle, Colour);
        out5[2] = offsetof(Particle, Velocity);
        out5[3] = offsetof(Particle, Forces);
        out5[4] = offsetof(Particle, Mass);
    }

    int ref_hardware_workers() { BruteSim(); return static_cast<int>(g_bruteSpawned); }

    // kind: 0 Random, 1 Galaxy, 2 StarSystem (EParticleSeeder, src/Sim/IParticleSeeder.hpp:12-17)
    void ref_seed(void* aos, size_t n, int kind, uint64_t seed, float scale)
    {
        std::vector<Particle> v = ToVector(aos, n);
        auto seeder = CreateParticleSeeder(v, static_cast<EParticleSeeder>(kind), scale);
        seeder->Seed(seed);
        FromVector(v, aos);
    }

    // The same for the renderer's 32-byte LWParticle (src/Render/Misc/Particle.hpp:20-25), which is
    // what Galaxy.cpp:61, GalaxyTarget.cpp:109, StarTarget.cpp:209 and UniverseTarget.cpp:98 seed.
    int ref_sizeof_lwparticle() { return static_cast<int>(sizeof(LWParticle)); }

    void ref_seed_ex(void* aos, size_t n, int kind, uint64_t seed, float scale, const float* rgb6)
    {
        std::vector<Particle> v = ToVector(aos, n);
        SeedWithColours(v, kind, seed, scale, rgb6);
        FromVector(v, aos);
    }

    void ref_seed_lw(void* lw, size_t n, int kind, uint64_t seed, float scale, const float* rgb6)
    {
        std::vector<LWParticle> v(n);
        if (n) std::memcpy(static_cast<void*>(v.data()), lw, n * sizeof(LWParticle));
        SeedWithColours(v, kind, seed, scale, rgb6);
        if (n) std::memcpy(lw, v.data(), n * sizeof(LWParticle));
    }

    // Maths::ClosestParticle (src/Core/Maths.hpp:62-85): index of the particle nearest to `pos`
    // (first minimum of the fp32 squared distance).
    uint64_t ref_closest_particle(const void* aos, size_t n, const float* pos3)
    {
        std::vector<Particle> v = ToVector(aos, n);
        size_t id = 0;
        Maths::ClosestParticle(DirectX::SimpleMath::Vector3(pos3[0], pos3[1], pos3[2]), v, &id);
        return static_cast<uint64_t>(id);
    }

    // Forces[target] accumulated by the reference's own BruteForceCPU::Exec for each listed target
    // (src/Sim/BruteForceCPU.cpp:25-43); out3 = nt x 3 doubles (force, not acceleration).
    void ref_bruteforce_forces(const void* aos, size_t n, const int64_t* targets, size_t nt, double* out3)
    {
        std::vector<Particle> v = ToVector(aos, n);
        BruteForceCPU* sim = BruteSim();
        sim->Init(v);
        for (size_t k = 0; k < nt; ++k)
        {
            const size_t i = static_cast<size_t>(targets[k]);
            v[i].Forces = Vec3d();
            sim->Exec({ i, 1 });
            out3[3 * k + 0] = v[i].Forces.x;
            out3[3 * k + 1] = v[i].Forces.y;
            out3[3 * k + 2] = v[i].Forces.z;
            v[i].Forces = Vec3d();
        }
    }

    // steps x BruteForceCPU::Update(dt) in place; returns seconds spent in the Update loop
    // (steady_clock, like SimulationState::RunBenchmark, SimulationState.cpp:346-351).
    double ref_bruteforce_run(void* aos, size_t n, float dt, int steps, int workers, int* workersUsed)
    {
        std::vector<Particle> v = ToVector(aos, n);
        BruteForceCPU* sim = BruteSim();
        const uint32_t spawned = g_bruteSpawned;
        const uint32_t w = PickWorkers(workers, spawned, n);
        sim->Pool.SetNumWorkers(w);
        if (workersUsed) *workersUsed = static_cast<int>(w);
        sim->Init(v);
        const auto t0 = std::chrono::steady_clock::now();
        for (int s = 0; s < steps; ++s) sim->Update(dt);
        const auto t1 = std::chrono::steady_clock::now();
        FromVector(v, aos);
        return std::chrono::duration<double>(t1 - t0).count();
    }

    // Timed block of `count` consecutive targets starting at `first`, split over the pool exactly as
    // BruteForceCPU::Update does (one Dispatch per worker + Join), without the integrator.  Used as
    // the bounded CPU-baseline sample at large N.  Forces of the block are written to out3 if given.
    double ref_bruteforce_block(const void* aos, size_t n, size_t first, size_t count, int workers,
                                int* workersUsed, double* out3)
    {
        std::vector<Particle> v = ToVector(aos, n);
        BruteForceCPU* sim = BruteSim();
        const uint32_t spawned = g_bruteSpawned;
        // any worker count that divides the block (the pool itself has no power-of-two constraint):
        // the largest w <= min(request, spawned) with count % w == 0 and the block aligned to count / w
        uint32_t w = spawned;
        if (workers > 0 && static_cast<uint32_t>(workers) < w) w = workers;
        if (w < 1) w = 1;
        while (w > 1 && ((count % w) != 0 || (first % (count / w)) != 0)) --w;
        if (workersUsed) *workersUsed = static_cast<int>(w);
        sim->Init(v);
        for (size_t k = 0; k < count; ++k) v[first + k].Forces = Vec3d();
        const size_t loops = count / w;
        const size_t base = first / loops;
        const auto t0 = std::chrono::steady_clock::now();
        sim->Pool.SetNumWorkers(w);
        for (uint32_t t = 0; t < w; ++t) sim->Pool.Dispatch(t, { base + t, loops });
        sim->Pool.Join();
        const auto t1 = std::chrono::steady_clock::now();
        if (out3)
            for (size_t k = 0; k < count; ++k)
            {
                out3[3 * k + 0] = v[first + k].Forces.x;
                out3[3 * k + 1] = v[first + k].Forces.y;
                out3[3 * k + 2] = v[first + k].Forces.z;
            }
        return std::chrono::duration<double>(t1 - t0).count();
    }

    // steps x BarnesHut::Update(dt) in place at the given theta; returns seconds in the loop.
    double ref_barneshut_run(void* aos, size_t n, float dt, int steps, double theta, int workers, int* workersUsed)
    {
        std::vector<Particle> v = ToVector(aos, n);
        BarnesHut* sim = TreeSim();
        const uint32_t spawned = g_treeSpawned;
        const uint32_t w = PickWorkers(workers, spawned, n);
        sim->Pool.SetNumWorkers(w);
        if (workersUsed) *workersUsed = static_cast<int>(w);
        Octree::Theta = theta;
        sim->Init(v);
        const auto t0 = std::chrono::steady_clock::now();
        for (int s = 0; s < steps; ++s) sim->Update(dt);
        const auto t1 = std::chrono::steady_clock::now();
        FromVector(v, aos);
        return std::chrono::duration<double>(t1 - t0).count();
    }

    // Tree built as BarnesHut::Update builds it, then Octree::CalculateForce for each listed target.
    // Returns seconds spent building (Add + CalculateMass); evalSeconds gets the traversal time.
    double ref_barneshut_forces(const void* aos, size_t n, double theta, const int64_t* targets, size_t nt,
                                double* out3, double* evalSeconds)
    {
        std::vector<Particle> v = ToVector(aos, n);
        const auto t0 = std::chrono::steady_clock::now();
        std::unique_ptr<Octree> tree = BuildTree(v, theta);
        const auto t1 = std::chrono::steady_clock::now();
        for (size_t k = 0; k < nt; ++k)
        {
            Vec3d f = tree->CalculateForce(&v[static_cast<size_t>(targets[k])]);
            out3[3 * k + 0] = f.x; out3[3 * k + 1] = f.y; out3[3 * k + 2] = f.z;
        }
        const auto t2 = std::chrono::steady_clock::now();
        if (evalSeconds) *evalSeconds = std::chrono::duration<double>(t2 - t1).count();
        return std::chrono::duration<double>(t1 - t0).count();
    }

    // Topology probe.  For every body: depth of the leaf that holds it (-1 if the body is in no
    // leaf, i.e. outside the root cube) and its child-index path packed 3 bits per level, first
    // level in the most significant used position (digit = z*4 + y*2 + x, Octree.cpp:25-47).
    // stats4 = {nodes, internal nodes, max depth, root NumParticles}.
    void ref_octree_paths(const void* aos, size_t n, int32_t* leafDepth, uint64_t* path, int64_t* stats4)
    {
        std::vector<Particle> v = ToVector(aos, n);
        std::unique_ptr<Octree> tree = BuildTree(v, 0.5);
        for (size_t i = 0; i < n; ++i)
        {
            Particle* p = &v[i];
            Octree* t = tree.get();
            uint64_t code = 0;
            int depth = -1;
            if (t->Bounds.Contains(p))
            {
                while (true)
                {
                    if (t->NumParticles == 1 && t->P == p) { depth = t->Depth; break; }
                    if (t->IsLeaf) { depth = -2; break; }  // inconsistent: not found
                    int next = -1;
                    for (int c = 0; c < 8; ++c)
                        if (t->Children[c]->Bounds.Contains(p)) { next = c; break; }
                    if (next < 0) { depth = -3; break; }   // fell into a rounding gap
                    code = (code << 3) | static_cast<uint64_t>(next);
                    t = t->Children[next].get();
                }
            }
            leafDepth[i] = depth;
            path[i] = code;
        }
        if (stats4)
        {
            int64_t nodes = 0, internal = 0; int maxDepth = 0;
            CountNodes(tree.get(), &nodes, &internal, &maxDepth);
            stats4[0] = nodes; stats4[1] = internal; stats4[2] = maxDepth; stats4[3] = tree->NumParticles;
        }
    }

    // What Octree::RenderDebug draws (Octree.cpp:147-175): one cube per OCCUPIED LEAF, at
    // pos = (TopLeft + BottomRight) / 2 with size = BottomRight.x - TopLeft.x.  Collected by the same
    // recursion; out4[k] = {pos.x, pos.y, pos.z, size}, body[k] = index of the leaf's particle.
    static void CollectLeafCubes(Octree* t, const Particle* base, float* out4, int64_t* body, size_t* count)
    {
        if (t->IsLeaf && t->NumParticles > 0)
        {
            auto avg = (t->Bounds.TopLeft + t->Bounds.BottomRight) / 2;
            const size_t k = (*count)++;
            out4[4 * k + 0] = avg.x; out4[4 * k + 1] = avg.y; out4[4 * k + 2] = avg.z;
            out4[4 * k + 3] = t->Bounds.BottomRight.x - t->Bounds.TopLeft.x;
            body[k] = static_cast<int64_t>(t->P - base);
        }
        if (!t->IsLeaf)
            for (auto& child : t->Children) CollectLeafCubes(child.get(), base, out4, body, count);
    }

    size_t ref_octree_leaf_cubes(const void* aos, size_t n, float* out4, int64_t* body)
    {
        std::vector<Particle> v = ToVector(aos, n);
        std::unique_ptr<Octree> tree = BuildTree(v, 0.5);
        size_t count = 0;
        CollectLeafCubes(tree.get(), v.data(), out4, body, &count);
        return count;
    }

    // Mass / centre of mass / population of the cell reached from the root by `depth` digits of
    // `path` (same packing as above).  Returns 0 on success, 1 if the path leaves the tree.
    int ref_octree_cell(const void* aos, size_t n, int depth, uint64_t path, double* mass, float* com3,
                        int32_t* numParticles, float* width)
    {
        static std::vector<Particle> cached;
        static std::unique_ptr<Octree> tree;
        static const void* cachedKey = nullptr;
        static size_t cachedN = 0;
        if (cachedKey != aos || cachedN != n || !tree)
        {
            cached = ToVector(aos, n);
            tree = BuildTree(cached, 0.5);
            cachedKey = aos; cachedN = n;
        }
        Octree* t = tree.get();
        for (int l = 0; l < depth; ++l)
        {
            if (t->IsLeaf) return 1;
            const int digit = static_cast<int>((path >> (3 * (depth - 1 - l))) & 7u);
            t = t->Children[digit].get();
        }
        *mass = t->TotalMass;
        com3[0] = t->CentreOfMass.x; com3[1] = t->CentreOfMass.y; com3[2] = t->CentreOfMass.z;
        *numParticles = t->NumParticles;
        *width = t->Bounds.BottomRight.x - t->Bounds.TopLeft.x;
        return 0;
    }

    void ref_octree_cell_reset() { /* next ref_octree_cell call with a new buffer rebuilds */ }

    // Work counters of the reference traversal for the listed targets:
    // out3 = {accepted-cell evaluations, leaf (pair) evaluations, node visits}, summed.
    void ref_barneshut_work(const void* aos, size_t n, double theta, const int64_t* targets, size_t nt, int64_t* out3)
    {
        std::vector<Particle> v = ToVector(aos, n);
        std::unique_ptr<Octree> tree = BuildTree(v, theta);
        out3[0] = out3[1] = out3[2] = 0;
        for (size_t k = 0; k < nt; ++k)
            CountEvals(tree.get(), &v[static_cast<size_t>(targets[k])], &out3[0], &out3[1], &out3[2]);
    }

    // Theta through the reference's own event plumbing (BarnesHut.cpp:29-31).
    void ref_report_theta(float theta)
    {
        TreeSim();
        FloatEventData data(theta);
        EventStream::Report(EEvent::BHThetaChanged, data);
    }

    double ref_get_theta() { return Octree::Theta; }
}
