// Internal state of one engine handle.  Not part of the ABI.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include <string>

#include "../../include/nbody_b200.h"

namespace nb
{

void set_error(const char* fmt, ...);

#define NB_CUDA(call)                                                                              \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            nb::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_));   \
            return NB_ERR_CUDA;                                                                    \
        }                                                                                          \
    } while (0)

#define NB_CHECK(call)                                                                             \
    do {                                                                                           \
        int s_ = (call);                                                                           \
        if (s_ != NB_OK) return s_;                                                                \
    } while (0)

#define NB_REQUIRE(cond, status, msg)                                                              \
    do {                                                                                           \
        if (!(cond)) {                                                                             \
            nb::set_error("%s: %s", __func__, msg);                                                \
            return status;                                                                         \
        }                                                                                          \
    } while (0)

// Barnes-Hut working set of one handle (tree.cu).
struct TreeBuffers
{
    size_t capacity = 0;            // bodies the buffers were sized for
    unsigned long long* keys[2] = {nullptr, nullptr};   // Morton codes, ping-pong for the radix sort
    unsigned int* vals[2] = {nullptr, nullptr};         // body index per slot, ping-pong
    uint32_t* hist = nullptr;       // radix-sort block histograms
    size_t hist_words = 0;
    uint32_t* desc = nullptr;       // onesweep: [8][tiles][256] tile descriptors + [8][256] digit counts + 8 tickets
    size_t desc_words = 0;
    uint32_t* counters = nullptr;   // [0] in-bounds bodies, [1..] scratch
    int2* child = nullptr;          // [n-1] {left, right} child of internal node (< n internal, >= n leaf slot + n)
    int32_t* parent = nullptr;      // [2n-1]   parent of internal nodes then of leaves
    int32_t* prefix = nullptr;      // [n-1]    common-prefix length in bits (0..63; 64+ = duplicate codes)
    int32_t* range = nullptr;       // [n-1] first sorted slot covered
    int32_t* range_hi = nullptr;    // [n-1] last sorted slot covered
    uint32_t* flags = nullptr;      // [n-1]    arrival counters of the bottom-up pass
    double* nsum = nullptr;         // [n-1][4] {G M, G M x, G M y, G M z} of internal nodes, fp64, one 32-byte record each
    float4* walk_a = nullptr;       // [2n-1] {com.x, com.y, com.z, G*M}
    int4* walk_b = nullptr;         // [2n-1] {open threshold (float bits), first slot, next-if-open, next-if-skip}
    unsigned long long* stats = nullptr;   // [3] accepted cells, pair evals, node visits
    unsigned int* cnt = nullptr;    // [n+1] owning nodes per first slot
    unsigned int* pref = nullptr;   // [n+1] exclusive scan of cnt
    int* rank = nullptr;            // [2n]  pre-order rank of every traversal record
    unsigned short* meta = nullptr; // [n-1] octree level of an internal node | 0x100 if it owns octree cells
    unsigned int* tlist = nullptr;  // [n]   owned targets in Morton order (world > 1)
    unsigned long long* keys_final = nullptr;   // [n] sorted keys assembled from every rank's segment (p2p.cu)
    unsigned int* vals_final = nullptr;         // [n]
    unsigned long long* splitters = nullptr;    // [NB_MAX_PEERS + 1] first key of every rank's segment of the NEXT sort
    const unsigned long long* skeys = nullptr;  // the sorted keys / body indices of the last build
    const unsigned int* svals = nullptr;
    bool dist_ready = false;        // splitters are valid: the next collective build may shard the sort
    int sort_bits_done = 0;
    int cur = 0;                    // which ping-pong half holds the sorted result
    size_t n_inbounds_host = 0;
    bool built = false;
    bool split_ids = false;         // internal node ids are split positions (k_build_up), not Karras's range ends
};

}  // namespace nb

namespace nb
{
// Where a balanced walk stores accelerations: rank r's acc[3][count_r] planes and its first body.
struct AccTable
{
    double* acc[NB_MAX_PEERS];        // rank r's two acc buffers, [2][3][count_r]
    int first[NB_MAX_PEERS + 1];      // first[r] = r * n / world; first[world] = n
    int world, rank;
    int parity;                       // which of the two buffers this step writes
};
}  // namespace nb

constexpr int NB_TIMING_RING = 64;

struct nb_sim
{
    nb_config cfg;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;

    size_t n = 0;          // all bodies (sources)
    size_t first = 0;      // owned (target) range
    size_t count = 0;

    float4* posw = nullptr;      // [n]        {x, y, z, (float)(G*m)}  == posw_buf[posw_cur]
    float4* posw_buf[2] = {nullptr, nullptr};   // second buffer only with the fused P2P exchange
    int posw_cur = 0;
    double* vel = nullptr;       // [3][count] velocity planes of owned bodies
    double* mass = nullptr;      // [count]
    double* acc = nullptr;       // [3][count] accelerations of the last force evaluation (== acc_base + acc_cur * 3 * count)
    double* acc_base = nullptr;  // the allocation: one buffer, or two once peer memory is prepared (p2p.cu)
    int acc_cur = 0;
    bool acc_two = false;
    double* acc_part = nullptr;  // [splits][3][count] all-pairs partial sums
    size_t acc_part_splits = 0;

    float* wmax = nullptr;       // [1] max_j G*m_j over all sources (bit pattern, kept by atomicMax)
    void* d_aos = nullptr;       // device image of the caller's AoS array
    size_t d_aos_bytes = 0;
    size_t d_aos_stride = 0;     // record stride of the image (0: no image yet)

    int ap_kernel = 0;           // index into allpairs_table()
    int ap_splits = 1;
    bool acc_valid = false;
    bool forces_from_last_step = false;
    bool acc_is_last_step = false;   // h->acc still holds what the last nb_step kicked with (no nb_compute_accel since)
    bool exchanged = true;       // posw of remote ranks is current

    cudaEvent_t ev[2] = {nullptr, nullptr};   // begin / end of the last nb_step (or nb_compute_accel) call
    // {force pass begin, dominant kernel begin, dominant kernel end} of the last NB_TIMING_RING force passes:
    // recorded every step without a host sync, read back after the timed region (nb_step_timing_mean)
    cudaEvent_t ring[NB_TIMING_RING][3] = {};
    unsigned long long ring_pos = 0;          // force passes recorded so far
    // Small scenes (one GPU, Barnes-Hut): the ~40 launches of a step are replayed as two CUDA graphs -- tree build and
    // walk -- captured once per (body count, theta); the reference's own range is <= 50 000 bodies (UI.cpp:75), where
    // a step is launch-bound
    cudaGraphExec_t graph_build = nullptr, graph_walk = nullptr;
    size_t graph_n = 0;
    float graph_theta = 0.f;
    int graph_launches = 0;
    bool graph_failed = false;
    bool timing_valid = false;
    int last_launches = 0;
    unsigned long long total_launches = 0;

    nb::TreeBuffers tree;

    void* nccl_comm = nullptr;   // ncclComm_t

    // fused kick-drift + exchange over peer memory (p2p.cu)
    bool p2p_attached = false;
    bool p2p_ipc = false;
    unsigned int p2p_step = 0;
    unsigned int p2p_acc_step = 0;
    unsigned int p2p_sort_step = 0;
    // [4][NB_MAX_PEERS] step counters raised by the peers (positions, accelerations, sort counts, sort
    // data), then [NB_MAX_PEERS] segment sizes of the sharded sort
    unsigned int* p2p_flags = nullptr;
    void* p2p_report = nullptr;                   // pinned host record written by a peer wait that timed out
    void* peer_skeys[NB_MAX_PEERS] = {};
    void* peer_svals[NB_MAX_PEERS] = {};
    void* peer_posw[2][NB_MAX_PEERS] = {};
    void* peer_flags[NB_MAX_PEERS] = {};
    void* peer_acc[NB_MAX_PEERS] = {};            // every rank's acc[3][count] (balanced Barnes-Hut walk)
};

namespace nb
{
// allpairs / integrator (nb_api.cu, integrate.cu)
int launch_allpairs(nb_sim* h);
int launch_kick_drift(nb_sim* h, float dt);
int launch_unpack_aos(nb_sim* h, size_t stride, size_t begin, size_t end);
int launch_reduce_partials(nb_sim* h);
int launch_pack_aos(nb_sim* h, size_t stride, bool forces_zero);
int choose_allpairs_config(nb_sim* h);
int preload_integrate();
int preload_tree();
int preload_allpairs(nb_sim* h);
int preload_step_kernels(nb_sim* h);

// tree.cu
int tree_reserve(nb_sim* h);
void tree_release(nb_sim* h);
int tree_build(nb_sim* h, bool collective = false);
int tree_walk(nb_sim* h, bool balanced = false, bool instrumented = false);

// nccl_dl.cpp
int comm_unique_id(uint8_t id[128]);
int comm_init(nb_sim* h, const uint8_t id[128]);
int comm_allgather_posw(nb_sim* h);
void comm_destroy(nb_sim* h);

// p2p.cu
int p2p_prepare(nb_sim* h);
int p2p_wait(nb_sim* h);
int p2p_kick_drift_push(nb_sim* h, float dt);
struct AccTable;
int p2p_acc_table(const nb_sim* h, AccTable* out);
int p2p_acc_exchange(nb_sim* h);
bool p2p_describe_timeout(const nb_sim* h, char* out, size_t cap);
int p2p_sort_exchange(nb_sim* h, const unsigned long long* keys_local, const unsigned int* vals_local, const unsigned int* count_dev);
void p2p_release(nb_sim* h);

// seed_host.cpp / seed_device.cu
int seed_galaxy_host(void* particles, size_t n, size_t stride, uint64_t seed, float scale);
int seed_galaxy_device(nb_sim* h, size_t n, uint64_t seed, float scale);
int seed_collision_device(nb_sim* h, size_t n, uint64_t seed, float scale, float separation, double approach_speed);
int reserve_aos(nb_sim* h, size_t bytes);
int launch_scale_masses(nb_sim* h, double factor);

// energy.cu
int energy(nb_sim* h, double* ke, double* pe);
int energy_sampled(nb_sim* h, size_t stride, double* ke, double* pe, size_t* nsamples);

// probe.cu
int probe_fp32_peak(nb_sim* h, double* flops);
}  // namespace nb
