/*
 * nbody_b200.h -- C ABI of the B200-native N-body engine (libnbody_b200.so).
 *
 * This is the drop-in boundary for the particle-simulation hot path of
 * matty9090/Procedural-Universe (src/Sim).  The reference has no FFI of its own: its sims sit
 * behind the C++ interface
 *
 *     class INBodySim { virtual void Init(std::vector<Particle>&); virtual void Update(float dt); }
 *                                                     (reference src/Sim/INBodySim.hpp:19-25)
 *     class IParticleSeeder { virtual void Seed(uint64_t seed); }
 *                                                     (reference src/Sim/IParticleSeeder.hpp:19-27)
 *
 * Every entry point below names the reference member it stands in for.  The C++ adapter
 * `B200Sim : INBodySim` (procedural-universe_b200/host/B200Sim.hpp) is a ~100-line wrapper over
 * exactly these calls; INTEGRATION.md shows the factory patch a maintainer would add.
 *
 * Conventions
 *   - plain C, POD arguments only; every function returns NB_OK (0) or a negative nb_status and
 *     leaves a message for nb_last_error() (the reference's sims return void and log failures,
 *     BruteForceGPU.cpp:25-33; the adapter turns a non-zero status into LOGE and keeps the last
 *     good state);
 *   - a handle is driven by one host thread at a time (the reference calls Update from the UI
 *     thread only, SimulationState.cpp:52-53);
 *   - there is NO CPU fallback: with no CUDA device or with a kernel image that does not match the
 *     device (sm_100a only) nb_create fails with NB_ERR_CUDA.
 *
 * Particle record (reference src/Render/Misc/Particle.hpp:8-18, g++/MSVC x64 layout, 104 bytes):
 *   Position float3 @0 | Colour float4 @12 | OriginalColour float4 @28 | pad @44 |
 *   Velocity double3 @48 | Forces double3 @72 | Mass double @96
 * The engine reads Position, Velocity and Mass and writes Position, Velocity and Forces; the
 * colour fields are never touched (the UI recolours picked particles in place,
 * SimulationState.cpp:364-396).
 */
#ifndef NBODY_B200_H
#define NBODY_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NB_ABI_VERSION 1

#if defined(__GNUC__)
#define NB_API __attribute__((visibility("default")))
#else
#define NB_API
#endif

typedef struct nb_sim* nb_handle;

typedef enum nb_status
{
    NB_OK = 0,
    NB_ERR_ARG = -1,      /* null / out-of-range argument, wrong stride, handle not initialised */
    NB_ERR_CUDA = -2,     /* CUDA runtime error, no device, or no sm_100a kernel image */
    NB_ERR_NCCL = -3,     /* NCCL missing or failed */
    NB_ERR_STATE = -4,    /* call not valid in the handle's mode / before nb_init_* */
    NB_ERR_NOMEM = -5
} nb_status;

/* Which reference sim the handle stands in for (ENBodySim, INBodySim.hpp:11-17). */
typedef enum nb_mode
{
    NB_MODE_ALLPAIRS = 0,   /* BruteForceCPU / BruteForceGPU: exact O(N^2) sum          */
    NB_MODE_BARNESHUT = 2   /* BarnesHut: per-step octree rebuild + theta-walk           */
} nb_mode;

#define NB_PARTICLE_STRIDE 104
#define NB_OFF_POSITION 0
#define NB_OFF_VELOCITY 48
#define NB_OFF_FORCES 72
#define NB_OFF_MASS 96

typedef struct nb_config
{
    uint32_t struct_size;    /* = sizeof(nb_config); set by nb_default_config                      */
    int32_t  device;         /* CUDA device ordinal                                                */
    int32_t  mode;           /* nb_mode                                                            */
    float    theta;          /* Octree::Theta (Octree.cpp:5 default 2.0; benchmarks use 0.5)       */
    double   G;              /* Phys::G  = 6.674e-11   (Physics.hpp:9)                             */
    double   softening;      /* Phys::S  = 10, ADDED to d^2 (Physics.hpp:10,34)                    */
    double   position_scale; /* Phys::StarSystemScale = 2.3e13 (Physics.hpp:13,16)                 */
    float    bounds;         /* half-width of the fixed octree root cube = 4000 (BarnesHut.cpp:14) */
    int32_t  rank;           /* this handle integrates bodies [rank*N/world, (rank+1)*N/world)     */
    int32_t  world;          /* number of cooperating handles (one per GPU); 1 = single GPU        */
    void*    stream;         /* cudaStream_t to launch on; NULL = a private non-blocking stream    */
    int32_t  source_splits;  /* all-pairs: source-range splits per target block, 0 = auto          */
    int32_t  kernel_variant; /* all-pairs inner-loop variant, 0 = default          */
} nb_config;

/* ---- lifetime ------------------------------------------------------------------------------ */
NB_API int nb_abi_version(void);
NB_API const char* nb_last_error(void);
NB_API int nb_default_config(nb_config* cfg);
/* CreateNBodySim(ctx, type) -- INBodySim.cpp:7-27 */
NB_API int nb_create(const nb_config* cfg, nb_handle* out);
NB_API int nb_destroy(nb_handle h);
/* BHThetaChanged event -> Octree::Theta -- BarnesHut.cpp:29-31 */
NB_API int nb_set_theta(nb_handle h, float theta);

/* ---- INBodySim::Init ----------------------------------------------------------------------- */
/* Init(std::vector<Particle>&) -- BruteForceCPU.cpp:20-23, BarnesHut.cpp:39-42.  `particles` is a
 * HOST array of n records of `stride` bytes laid out as above (stride >= 104).  The data is
 * copied to the device; the caller keeps ownership. */
NB_API int nb_init_aos(nb_handle h, const void* particles, size_t n, size_t stride);
/* Same from structure-of-arrays HOST buffers: pos[3n] float, vel[3n] double, mass[n] double. */
NB_API int nb_init_soa(nb_handle h, const float* pos3, const double* vel3, const double* mass, size_t n);

/* ---- IParticleSeeder::Seed ----------------------------------------------------------------- */
/* GalaxySeeder<Particle>(particles, scale).Seed(seed) -- GalaxySeeder.cpp:43-80 -- into a HOST
 * AoS buffer, bit-identical to the reference built with g++/libstdc++ (the random engine and
 * distributions are implementation-defined).  Does not need a handle. */
NB_API int nb_seed_galaxy_host(void* particles, size_t n, size_t stride, uint64_t seed, float scale);
/* All three reference seeders (EParticleSeeder, IParticleSeeder.hpp:12-17; factory
 * CreateParticleSeeder<T>(particles, type, scale), :29-50) for both record types the reference
 * instantiates them with: T = Particle (the simulation, SimulationState.cpp:107-110) and
 * T = LWParticle (the renderer's 32-byte record, Render/Misc/Particle.hpp:20-25; Galaxy.cpp:61,
 * GalaxyTarget.cpp:109, StarTarget.cpp:209, UniverseTarget.cpp:98).  Output is bit-identical to the
 * reference built with g++/libstdc++:
 *   NB_SEEDER_RANDOM      RandomSeeder.cpp:13-40     cube +-500/scale, radial velocity 1e16, mass
 *                                                    U(1e20,1e30); IGNORES `seed` like the reference
 *   NB_SEEDER_GALAXY      GalaxySeeder.cpp:43-143    (nb_seed_galaxy_host is this with defaults)
 *   NB_SEEDER_STARSYSTEM  StarSystemSeeder.cpp:18-55 1e30 star + bodies on +z; IGNORES `seed`, `scale`
 * `opt` may be NULL (Particle layout, scale 1, colour ranges [0,1]). */
typedef enum nb_seeder_kind { NB_SEEDER_RANDOM = 0, NB_SEEDER_GALAXY = 1, NB_SEEDER_STARSYSTEM = 2 } nb_seeder_kind;
typedef enum nb_record_layout { NB_LAYOUT_PARTICLE = 0, NB_LAYOUT_LWPARTICLE = 1 } nb_record_layout;
#define NB_LW_PARTICLE_STRIDE 32   /* Position float3 @0 | Colour float4 @12 | Scale float @28 */
#define NB_LW_OFF_SCALE 28
typedef struct nb_seed_options
{
    uint32_t struct_size;   /* = sizeof(nb_seed_options); set by nb_seed_default_options            */
    int32_t  layout;        /* nb_record_layout                                                     */
    float    scale;         /* the seeders' `scale` constructor argument (positions are divided)    */
    float    red[2];        /* GalaxySeeder::SetRedDist(low, hi), clamped to [0,1] (GalaxySeeder.cpp:24-41) */
    float    green[2];      /* SetGreenDist                                                         */
    float    blue[2];       /* SetBlueDist                                                          */
} nb_seed_options;
NB_API int nb_seed_default_options(nb_seed_options* opt);
NB_API int nb_seed_host(int kind, void* particles, size_t n, size_t stride, uint64_t seed, const nb_seed_options* opt);
/* Two-galaxy collision scene used by config 5 (defined by this repo): galaxies seeded
 * with `seed` and `seed+1`, n/2 bodies each, offset by -/+ `separation`/2 along x and approaching
 * each other with `approach_speed` (velocity units). */
NB_API int nb_seed_collision_host(void* particles, size_t n, size_t stride, uint64_t seed, float scale,
                           float separation, double approach_speed);
/* The same three seeders ON THE DEVICE, producing bit for bit the records nb_seed_host (and the reference built
 * with g++ / libstdc++) produces.  GalaxySeeder's single minstd_rand0 stream with data-dependent draw counts
 * (polar-method rejection, cached normal variates, the disk's rejection loop, GalaxySeeder.cpp:43-143) is parsed in
 * parallel: LCG jump-ahead, per-chunk transition tables over all entry offsets, a composed scan, then one thread per
 * pair of bodies (csrc/seed_device.cu).  nb_seed_device copies the records to a HOST array (bytes the seeders do
 * not write are zero); `device` is the CUDA device to use. */
NB_API int nb_seed_device(int kind, int device, void* particles, size_t n, size_t stride, uint64_t seed, const nb_seed_options* opt);
/* GalaxySeeder<Particle>(particles, scale).Seed(seed) straight into the handle -- no host round trip; every rank
 * of a multi-GPU run seeds all bodies (it needs all positions) and keeps its own shard's velocities. */
NB_API int nb_seed_galaxy_device(nb_handle h, size_t n, uint64_t seed, float scale);
/* nb_seed_collision_host's scene straight into the handle. */
NB_API int nb_seed_collision_device(nb_handle h, size_t n, uint64_t seed, float scale, float separation, double approach_speed);
/* One-GPU Barnes-Hut steps of up to 2^21 bodies replay their ~40 launches as two CUDA graphs (tree build, walk),
 * captured at the first step after Init / a theta change: same kernels, same arguments, same results.  On by
 * default (environment NB_GRAPHS=0 disables it process-wide); this switches it per handle. */
NB_API int nb_enable_graphs(nb_handle h, int on);
/* Multiplies every body's Mass by `factor` in the device image and re-derives the device state from it (a scene
 * variant: the same bodies with the total mass of a smaller scene).  Needs the image (nb_init_aos / nb_seed_*_device). */
NB_API int nb_scale_masses(nb_handle h, double factor);
/* Records of the handle's device image of the Particle array (as seeded / last uploaded / last written back) for a
 * list of body indices: records[k][104].  Lets a caller check a few bodies of a 7 GB scene without reading it back. */
NB_API int nb_get_aos_records(nb_handle h, const uint32_t* bodies, size_t k, void* records);

/* ---- INBodySim::Update --------------------------------------------------------------------- */
/* nsteps x Update(dt) on device-resident state -- BruteForceCPU.cpp:45-74 / BarnesHut.cpp:44-96:
 * accelerations, then kick-drift  v += a*dt;  pos += (float)(v*dt/position_scale).
 * Asynchronous on the handle's stream; with world > 1 it performs the per-step position exchange. */
NB_API int nb_step(nb_handle h, float dt, int nsteps);
/* Update(dt) with the reference's host-memory contract: upload Position/Velocity/Mass from the
 * caller's AoS array, one step, write Position/Velocity/Forces back, synchronous on return
 * (SimulationState.cpp:52-60 copies the array to a vertex buffer right after Update). */
NB_API int nb_update_aos(nb_handle h, void* particles, size_t n, size_t stride, float dt);
/* Page-locks / unlocks a caller-owned host array (cudaHostRegister) so that nb_update_aos copies at
 * full PCIe rate; optional.  The adapter pins the std::vector it was Init'ed with. */
NB_API int nb_host_register(void* ptr, size_t bytes);
NB_API int nb_host_unregister(void* ptr);
/* Wait for everything queued on the handle's stream. */
NB_API int nb_sync(nb_handle h);

/* ---- read back ------------------------------------------------------------------------------ */
/* Writes Position, Velocity, Forces of all bodies this handle owns into the HOST AoS array
 * (records [first, first+count) of nb_owned_range).  Forces follow the reference: zero in
 * all-pairs mode (BruteForceCPU.cpp:72), mass * acceleration of the last step in Barnes-Hut
 * mode (BarnesHut.cpp:75,108). */
NB_API int nb_read_aos(nb_handle h, void* particles, size_t n, size_t stride);
NB_API int nb_read_soa(nb_handle h, float* pos3, double* vel3);
NB_API int nb_owned_range(nb_handle h, size_t* first, size_t* count);
NB_API int nb_num_bodies(nb_handle h, size_t* n);
/* The static partition of n bodies over `world` handles: rank r owns [r*n/world, (r+1)*n/world)
 * -- the same block partition the reference uses over its worker threads
 * (BruteForceCPU.cpp:47-54, without its remainder bug).  Host-only, needs no device. */
NB_API int nb_shard_range(size_t n, int rank, int world, size_t* first, size_t* count);

/* ---- parity hooks --------------------------------------------------------------------------- */
/* Accelerations (Forces / Mass of the reference) of the owned bodies for the CURRENT positions and
 * theta, without integrating: acc3[3*count] doubles.  nb_get_accel evaluates them if no evaluation of
 * the current positions exists (every nb_step / nb_init_* / nb_set_theta invalidates the last one). */
NB_API int nb_compute_accel(nb_handle h);
NB_API int nb_get_accel(nb_handle h, double* acc3);
/* The same accelerations for a list of GLOBAL body indices (NaN for bodies this handle does not own):
 * acc3[3*k].  At 16 M bodies nb_get_accel is a 400 MB read-back; parity checks want a few hundred. */
NB_API int nb_get_accel_of(nb_handle h, const uint32_t* bodies, size_t k, double* acc3);
/* The accelerations the LAST nb_step applied in its kick (Forces / Mass of that Update: evaluated at that
 * step's PRE-drift positions), for a list of global body indices -- what the Forces write-back of
 * nb_update_aos carries, without the 104-byte-per-body read-back.  After the first step of a run these are
 * the accelerations of the initial conditions at no extra force pass (one all-pairs pass over 16 M bodies
 * is 110 s on two GPUs).  NB_ERR_STATE before the first step. */
NB_API int nb_get_step_accel_of(nb_handle h, const uint32_t* bodies, size_t k, double* acc3);
/* Cross-check at any N: BruteForceCPU::Exec (BruteForceCPU.cpp:25-43) with Phys::Gravity (Physics.hpp:25-35)
 * restated operation by operation on the device -- fp32 difference / DistanceSquared / Normalize without
 * fusion, fp64 force and accumulation -- for the listed bodies (any bodies, owned or not) against ALL
 * sources of the current positions.  O(k * N); independent of the handle's mode and of the fast kernels. */
NB_API int nb_direct_accel(nb_handle h, const uint32_t* bodies, size_t k, double* acc3);
/* 64-bit checksums of the device state: hash2[0] over {x, y, z, G m} of ALL n bodies (equal on every rank
 * after an exchange, equal between runs iff the positions are bitwise equal), hash2[1] over the fp64
 * velocities of the OWNED bodies, keyed by global body index so that the values of all ranks add up
 * (mod 2^64) to the single-handle value. */
NB_API int nb_state_hash(nb_handle h, uint64_t hash2[2]);
/* Barnes-Hut topology of the last build: number of in-bounds bodies, their 63-bit Morton codes in
 * sorted order and the body index of each sorted slot (bodies outside the root cube are dropped
 * as sources, Octree.cpp:58-62). */
NB_API int nb_get_morton(nb_handle h, uint64_t* codes, uint32_t* order, size_t* n_inbounds);
/* Radix-tree topology: for each of the n_inbounds-1 internal nodes the left/right child
 * (>= 0: internal node index, < 0: ~sorted leaf slot), the common-prefix length in bits, the
 * mass and centre of mass.  Any pointer may be NULL. */
NB_API int nb_get_tree(nb_handle h, int32_t* left, int32_t* right, int32_t* prefix_bits, double* mass,
                float* com3, size_t* n_internal);
/* BarnesHut::RenderDebug -> Octree::RenderDebug (BarnesHut.cpp:98-101, Octree.cpp:147-175): the reference
 * draws one cube per OCCUPIED LEAF of its octree.  For the current positions: cells4[4*k] = {centre.x,
 * centre.y, centre.z, size} of the leaf holding in-bounds body body[k], k in Morton order (the order of
 * the reference's recursion over children 0..7).  Either output may be NULL. */
NB_API int nb_get_leaf_cells(nb_handle h, float* cells4, uint32_t* body, size_t* n_inbounds);
/* Counters of the last traversal summed over owned targets: {accepted cells, pair (leaf)
 * evaluations, node visits}. */
NB_API int nb_get_walk_stats(nb_handle h, uint64_t stats3[3]);
/* Profiling hook: lane occupancy of the instrumented traversal nb_get_walk_stats just ran --
 * hist33[k] = warp iterations (node visits by a warp) during which k of the 32 lanes were at work
 * (the others were parked inside a subtree they had accepted as a whole). */
NB_API int nb_get_walk_occupancy(nb_handle h, uint64_t hist33[33]);
/* Of the same instrumented walk: how the visits with few lanes awake are spread over a warp's lanes.  out7 =
 * {sum, max-per-warp summed over warps} of the per-lane counts of visits with <= 4, <= 8, <= 16 lanes awake, then the
 * number of warps.  max / (sum / 32) is the imbalance a per-lane treatment of those visits would meet. */
NB_API int nb_get_walk_sparse_load(nb_handle h, uint64_t out7[7]);
/* Kinetic and potential energy of the conserved quantity of this force law,
 * E = sum 1/2 m v^2 + position_scale * sum_{i<j} U(r), U = -(G ma mb / sqrt(S)) atan(sqrt(S)/r),
 * potential by exact pair sum over owned targets x all sources (O(N^2/world)). */
NB_API int nb_energy(nb_handle h, double* kinetic, double* potential);
/* The same quantity where the exact pair sum is out of reach (config 5, 64 M bodies): kinetic energy
 * exactly over the owned bodies, potential estimated from the owned bodies whose GLOBAL index is a
 * multiple of `stride`, each against all sources, scaled by `stride`.  stride = 1 is the exact sum.
 * The sample is a fixed set of bodies, so drift |E(t)-E(0)|/|E(0)| compares the same bodies. */
NB_API int nb_energy_sampled(nb_handle h, size_t stride, double* kinetic, double* potential, size_t* nsamples);

/* ---- queries ---------------------------------------------------------------------------------- */
/* Maths::ClosestParticle(pos, particles, &id) -- Core/Maths.hpp:62-85, pinned by the reference's
 * test/MathsTests.cpp:4-33 -- on the handle's CURRENT device-resident positions: index of the body
 * with the smallest fp32 DistanceSquared to `pos`; the first index wins ties (strict <, scan in
 * index order); 0 when no body has a finite distance below FLT_MAX.  dist_sq may be NULL. */
NB_API int nb_closest_particle(nb_handle h, const float pos[3], size_t* index, float* dist_sq);

/* ---- .nbody particle files (host I/O, no handle) ------------------------------------------------ */
/* The reference's checkpoint format: the raw std::vector<Particle>, 104 bytes per body, no header
 * (writer SimulationState.cpp:317-331, reader :229-277). */
NB_API int nb_nbody_save(const char* path, const void* particles, size_t n, size_t stride);
/* Number of whole records in the file. */
NB_API int nb_nbody_count(const char* path, size_t* n);
/* Reads up to `capacity` records; a trailing partial record is dropped like the reference's read
 * loop does.  recentre != 0 applies InitParticlesFromFile's mass-weighted recentring (:252-270). */
NB_API int nb_nbody_load(const char* path, void* particles, size_t capacity, size_t stride, size_t* n_read, int recentre);
/* The recentring alone: Position -= (float3)(sum pos*Mass / sum Mass), sums in index order. */
NB_API int nb_nbody_recentre(void* particles, size_t n, size_t stride);

/* ---- multi-GPU plumbing ---------------------------------------------------------------------- */
/* One handle per process/GPU.  The host (torch.distributed, MPI, ...) moves 128 opaque bytes from
 * rank 0 to everyone; the library then owns an NCCL communicator and all-gathers the float4
 * {x,y,z,G*m} of the owned bodies after every kick-drift. */
NB_API int nb_comm_unique_id(uint8_t id[128]);
NB_API int nb_comm_init(nb_handle h, const uint8_t id[128]);
/* Raw device pointers for hosts that prefer to run the exchange themselves (e.g. through
 * torch.distributed.all_gather_into_tensor): the float4 array of all n bodies. */
NB_API int nb_device_posw(nb_handle h, void** dev_ptr, size_t* bytes);
/* Fused exchange over NVLink peer memory (one process per GPU, one box): after nb_init_*, every rank
 * exports NB_P2P_HANDLE_BYTES opaque bytes (CUDA IPC handles of its two position buffers, its
 * flag array, its acceleration array and the two arrays of the sorted Morton order), the launcher all-gathers them in rank order, every rank
 * attaches.  From then on the kick-drift kernel stores each new position directly into every rank's
 * position array and no collective is launched (csrc/p2p.cu); takes precedence over nb_comm_init.
 * In Barnes-Hut mode nb_step additionally balances the traversal: the Morton-ordered target list is
 * dealt out to the ranks block by block and every rank stores the accelerations it computed straight
 * into their owner's array, and shards the per-step sort: every rank sorts the bodies of one Morton-key
 * range and stores its segment into every rank's sorted arrays. */
#define NB_MAX_PEERS 16
#define NB_P2P_HANDLE_BYTES 384
NB_API int nb_p2p_export(nb_handle h, uint8_t handles[NB_P2P_HANDLE_BYTES]);
NB_API int nb_p2p_attach(nb_handle h, const uint8_t* all_handles /* world x NB_P2P_HANDLE_BYTES */);
/* Same for handles that live in ONE process (peers[r] = the handle of rank r). */
NB_API int nb_p2p_attach_local(nb_handle h, const nb_handle* peers);
/* Tells the handle that the host exchanged positions itself after the last step. */
NB_API int nb_mark_exchanged(nb_handle h);

/* ---- measurement ----------------------------------------------------------------------------- */
/* Device time in milliseconds (CUDA events on the handle's stream) of the last nb_step call and of
 * its dominant kernel, and how many kernels that call launched. */
NB_API int nb_last_step_timing(nb_handle h, float* total_ms, float* force_kernel_ms, int* launches);
/* The same two figures averaged over the last `max_steps` force passes (at most 64 are kept; 0 = all kept):
 * the library records three CUDA events per step on the handle's stream and reads them only here, so a
 * timed loop needs no host synchronisation between steps. */
NB_API int nb_step_timing_mean(nb_handle h, int max_steps, float* force_kernel_ms, float* build_ms, int* steps_averaged);
/* Mean period of the last steps on the device: from the begin of one force pass to the begin of the next (force
 * pass + kick-drift + exchange + whatever the caller put between the steps), over the same ring of events. */
NB_API int nb_step_period_mean(nb_handle h, int max_steps, float* period_ms, int* steps_averaged);
/* Barnes-Hut: device time of the tree build (Morton + sort + Karras + reduction) of that step;
 * force_kernel_ms above is then the traversal alone.  0 in all-pairs mode. */
NB_API int nb_last_build_timing(nb_handle h, float* build_ms);
/* Pure-FFMA issue-rate probe: sustained FP32 FLOP/s of this device as measured now. */
NB_API int nb_probe_fp32_peak(nb_handle h, double* flops_per_s);

#ifdef __cplusplus
}
#endif
#endif /* NBODY_B200_H */
