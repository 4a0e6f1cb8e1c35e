// nbody_headless -- the reference's headless entry points on the B200 engine, over the C ABI only.
//
//   nbody_headless -c [-t seconds] [-s timestep] [-p particles] [-f file.nbody]
//       = `nbody.exe --compute` (reference src/App/NBody.cpp:52-89): App::RunSimulation(timestep/60,
//         simtime, particles, file) -> SimulationState::RunSimulation (SimulationState.cpp:279-332):
//         StarSystem seeder, Barnes-Hut sim, Update(dt) until `simtime` seconds of wall clock have
//         passed, "[Info] Running... (k iterations)" once a second, then the particle vector is
//         dumped to data/<ticks>.nbody.  Same flags, same defaults (10 s, 0.02, 4000 bodies).
//   nbody_headless --benchmark [-p particles]
//       = SimulationState::RunBenchmark (:334-362): for every sim, Init, then 10 x Update(1.0f);
//         reports ms per frame (the reference shows the number in its UI).
//
// Differences, all deliberate:
//   * a file given with -f is NOT overwritten by the seeder (the reference re-seeds after loading,
//     SimulationState.cpp:283-289, which discards what it loaded);
//   * state stays on the device between steps; the host array is read back once at the end
//     (the reference's sims rewrite the caller's vector every Update);
//   * extra flags for scripted runs: --steps K (fixed step count instead of wall clock), --out PATH,
//     --sim bh|allpairs, --seeder starsystem|galaxy|random, --seed S, --theta T.
// Log lines keep the reference's format "[Info] text" / "[Error] text" (Services/Log.cpp, pinned by
// test/LogTests.cpp).
#include <sys/stat.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "nbody_b200.h"

namespace
{
void LogInfo(const std::string& s) { std::printf("[Info] %s\n", s.c_str()); std::fflush(stdout); }
void LogError(const std::string& s) { std::printf("[Error] %s\n", s.c_str()); std::fflush(stdout); }

struct Options
{
    bool compute = false, benchmark = false;
    int simtime = 10, particles = 4000;          // NBody.cpp:55
    float timestep = 0.02f;                      // NBody.cpp:56
    std::string file, out, sim = "bh", seeder = "starsystem";
    long steps = -1;
    unsigned long long seed = 0;
    float theta = 2.0f;                          // Octree::Theta, Octree.cpp:5
};

bool Parse(int argc, char** argv, Options& o)
{
    for (int i = 1; i < argc; ++i)
    {
        const std::string a = argv[i];
        auto value = [&](const char* what) -> const char* {
            if (i + 1 >= argc) { LogError(std::string("missing value for ") + what); return nullptr; }
            return argv[++i];
        };
        const char* v = nullptr;
        if (a == "-c" || a == "--compute") o.compute = true;
        else if (a == "--benchmark") o.benchmark = true;
        else if (a == "-t" || a == "--simtime") { if (!(v = value("simtime"))) return false; o.simtime = std::atoi(v); }
        else if (a == "-s" || a == "--timestep") { if (!(v = value("timestep"))) return false; o.timestep = (float)std::atof(v); }
        else if (a == "-p" || a == "--particles") { if (!(v = value("particles"))) return false; o.particles = std::atoi(v); }
        else if (a == "-f" || a == "--file") { if (!(v = value("file"))) return false; o.file = v; }
        else if (a == "--steps") { if (!(v = value("steps"))) return false; o.steps = std::atol(v); }
        else if (a == "--out") { if (!(v = value("out"))) return false; o.out = v; }
        else if (a == "--sim") { if (!(v = value("sim"))) return false; o.sim = v; }
        else if (a == "--seeder") { if (!(v = value("seeder"))) return false; o.seeder = v; }
        else if (a == "--seed") { if (!(v = value("seed"))) return false; o.seed = std::strtoull(v, nullptr, 10); }
        else if (a == "--theta") { if (!(v = value("theta"))) return false; o.theta = (float)std::atof(v); }
        else { LogError("unknown option " + a); return false; }
    }
    return true;
}

bool Check(int rc, const char* what)
{
    if (rc == NB_OK) return true;
    LogError(std::string(what) + ": " + nb_last_error());
    return false;
}

int SeederKind(const std::string& s)
{
    if (s == "galaxy") return NB_SEEDER_GALAXY;
    if (s == "random") return NB_SEEDER_RANDOM;
    return NB_SEEDER_STARSYSTEM;                 // RunSimulation's choice, SimulationState.cpp:287
}

nb_handle Create(int mode, float theta)
{
    nb_config cfg;
    nb_default_config(&cfg);
    cfg.mode = mode;
    cfg.theta = theta;
    nb_handle h = nullptr;
    if (!Check(nb_create(&cfg, &h), "nb_create")) return nullptr;
    // the sims' constructor log line (BruteForceCPU.cpp:17, BarnesHut.cpp:12), with the engine named
    LogInfo(mode == NB_MODE_BARNESHUT ? "Barnes-Hut (B200)" : "Brute Force (B200)");
    return h;
}

int RunSimulation(const Options& o)
{
    std::vector<unsigned char> particles((size_t)o.particles * NB_PARTICLE_STRIDE);
    size_t n = (size_t)o.particles;
    bool loaded = false;
    if (!o.file.empty())
    {
        // InitParticlesFromFile: "data/" + fname, all records, recentred (SimulationState.cpp:229-277)
        const std::string path = "data/" + o.file;
        size_t count = 0;
        if (nb_nbody_count(path.c_str(), &count) != NB_OK) LogError("Could not read particle file " + o.file);
        else
        {
            particles.assign(count * NB_PARTICLE_STRIDE, 0);
            if (Check(nb_nbody_load(path.c_str(), particles.data(), count, NB_PARTICLE_STRIDE, &n, 1), "nb_nbody_load"))
            {
                LogInfo("Read " + std::to_string(n) + " particles from file");
                loaded = true;
            }
        }
    }
    if (!loaded)
    {
        n = (size_t)o.particles;
        particles.assign(n * NB_PARTICLE_STRIDE, 0);
        if (!Check(nb_seed_host(SeederKind(o.seeder), particles.data(), n, NB_PARTICLE_STRIDE, o.seed, nullptr), "nb_seed_host")) return 1;
    }
    if (n == 0) { LogError("no particles"); return 1; }

    nb_handle h = Create(o.sim == "allpairs" ? NB_MODE_ALLPAIRS : NB_MODE_BARNESHUT, o.theta);
    if (!h) return 1;
    if (!Check(nb_init_aos(h, particles.data(), n, NB_PARTICLE_STRIDE), "nb_init_aos")) return 1;

    const float dt = o.timestep * (1.0f / 60.0f);                 // NBody.cpp:87
    using clock = std::chrono::steady_clock;
    const auto start = clock::now();
    auto tick = start;
    long iterations = 0;
    for (;;)
    {
        ++iterations;                                             // counted before the exit test, as in the reference (:303)
        const auto now = clock::now();
        if (o.steps >= 0 ? iterations > o.steps
                         : std::chrono::duration_cast<std::chrono::seconds>(now - start).count() >= o.simtime) break;
        if (std::chrono::duration_cast<std::chrono::seconds>(now - tick).count() >= 1)
        {
            tick = now;
            LogInfo("Running... (" + std::to_string(iterations) + " iterations)");
        }
        if (!Check(nb_step(h, dt, 1), "nb_step")) return 1;
        if (o.steps < 0 && !Check(nb_sync(h), "nb_sync")) return 1;   // wall-clock mode: do not queue ahead of the clock
    }
    if (!Check(nb_read_aos(h, particles.data(), n, NB_PARTICLE_STRIDE), "nb_read_aos")) return 1;
    const double secs = std::chrono::duration<double>(clock::now() - start).count();
    LogInfo("Ran " + std::to_string(iterations - 1) + " iterations in " + std::to_string(secs) + " s");

    std::string out = o.out;
    if (out.empty())
    {
        if (mkdir("data", 0777) != 0 && errno != EEXIST) LogError("Failed to create data directory");
        const long long ticks = std::chrono::duration_cast<std::chrono::nanoseconds>(clock::now().time_since_epoch()).count() / 100;
        out = "data/" + std::to_string(ticks) + ".nbody";         // QueryPerformanceCounter ticks in the reference (:320)
    }
    if (!Check(nb_nbody_save(out.c_str(), particles.data(), n, NB_PARTICLE_STRIDE), "nb_nbody_save")) return 1;
    LogInfo("Wrote " + out);
    nb_destroy(h);
    return 0;
}

int RunBenchmark(const Options& o)
{
    LogInfo("Running benchmark");
    const size_t n = (size_t)o.particles;
    std::vector<unsigned char> seeded(n * NB_PARTICLE_STRIDE, 0);
    // the sandbox's default particle set: galaxy seeder (SimulationState.cpp:107-110)
    if (!Check(nb_seed_host(o.seeder == "starsystem" ? NB_SEEDER_GALAXY : SeederKind(o.seeder), seeded.data(), n, NB_PARTICLE_STRIDE,
                            o.seed ? o.seed : 1, nullptr), "nb_seed_host")) return 1;
    const int modes[2] = {NB_MODE_ALLPAIRS, NB_MODE_BARNESHUT};
    const char* names[2] = {"Brute Force (B200)", "Barnes-Hut (B200)"};
    for (int m = 0; m < 2; ++m)
    {
        std::vector<unsigned char> particles = seeded;            // every sim starts from the same Particles
        nb_handle h = Create(modes[m], o.theta);
        if (!h) return 1;
        if (!Check(nb_init_aos(h, particles.data(), n, NB_PARTICLE_STRIDE), "nb_init_aos")) return 1;
        const int numFrames = 10;                                 // :339
        const auto t0 = std::chrono::steady_clock::now();
        for (int frame = 0; frame < numFrames; ++frame)
            if (!Check(nb_update_aos(h, particles.data(), n, NB_PARTICLE_STRIDE, 1.0f), "nb_update_aos")) return 1;   // Update(1.0f), :349
        const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count() / numFrames;
        char line[160];
        std::snprintf(line, sizeof(line), "Benchmark %s: %.3f ms/frame (%zu particles, %d frames)", names[m], ms, n, numFrames);
        LogInfo(line);
        nb_destroy(h);
    }
    LogInfo("Benchmark finished");
    return 0;
}
}  // namespace

int main(int argc, char** argv)
{
    Options o;
    if (!Parse(argc, argv, o)) return 2;
    if (o.benchmark) return RunBenchmark(o);
    if (o.compute) return RunSimulation(o);
    LogError("nothing to do: pass -c (precompute a simulation) or --benchmark; the interactive renderer is out of scope");
    return 2;
}
