// C ABI entry points (include/nbody_b200.h): handle lifetime, Init / Update / read-back.
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <new>

#include <cstdlib>

#include "nb_internal.h"
#include "allpairs.cuh"

namespace nb
{

static thread_local char g_error[512] = "";

void set_error(const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

static void graphs_release(nb_sim* h)
{
    if (h->graph_build) cudaGraphExecDestroy(h->graph_build);
    if (h->graph_walk) cudaGraphExecDestroy(h->graph_walk);
    h->graph_build = h->graph_walk = nullptr;
    h->graph_n = 0;
}

static int free_state(nb_sim* h)
{
    graphs_release(h);
    p2p_release(h);
    cudaFree(h->posw_buf[0]); cudaFree(h->posw_buf[1]);
    h->posw_buf[0] = h->posw_buf[1] = nullptr; h->posw = nullptr; h->posw_cur = 0;
    cudaFree(h->p2p_flags); h->p2p_flags = nullptr;
    if (h->p2p_report) { cudaFreeHost(h->p2p_report); h->p2p_report = nullptr; }
    cudaFree(h->vel); h->vel = nullptr;
    cudaFree(h->mass); h->mass = nullptr;
    cudaFree(h->acc_base); h->acc_base = nullptr; h->acc = nullptr; h->acc_cur = 0; h->acc_two = false;
    cudaFree(h->acc_part); h->acc_part = nullptr;
    h->acc_part_splits = 0;
    tree_release(h);
    h->n = h->first = h->count = 0;
    h->acc_valid = false;
    return NB_OK;
}

int preload_allpairs(nb_sim* h)
{
    int count = 0;
    const AllPairsKernel* table = allpairs_table(&count);
    cudaFuncAttributes a;
    NB_CUDA(cudaFuncGetAttributes(&a, reinterpret_cast<const void*>(table[h->ap_kernel].fn)));
    return NB_OK;
}

int reserve_aos(nb_sim* h, size_t bytes)
{
    if (h->d_aos_bytes >= bytes) return NB_OK;
    cudaFree(h->d_aos);
    h->d_aos = nullptr;
    h->d_aos_bytes = 0;
    NB_CUDA(cudaMalloc(&h->d_aos, bytes));
    h->d_aos_bytes = bytes;
    return NB_OK;
}

// Picks the all-pairs kernel and the number of source-range splits so that the grid fills whole
// waves of resident CTAs (148 SMs x CTAs/SM): a 1 M-body step is only ~3.5 waves of 1024-target
// CTAs, and a ragged last wave would idle 14 % of the machine.
int choose_allpairs_config(nb_sim* h)
{
    int count = 0;
    const AllPairsKernel* table = allpairs_table(&count);
    int idx = h->cfg.kernel_variant;
    if (idx < 0 || idx >= count) idx = 0;
    h->ap_kernel = idx;
    const AllPairsKernel& k = table[idx];

    int occ = 0;
    if (k.smem_bytes > 0) NB_CUDA(cudaFuncSetAttribute((const void*)k.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, k.smem_bytes));
    NB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, (const void*)k.fn, k.threads, k.smem_bytes));
    if (occ < 1) occ = 1;
    const long slots = (long)occ * h->sm_count;
    const long tgt_blocks = ((long)h->count + (long)k.threads * k.targets - 1) / ((long)k.threads * k.targets);
    const int tile = (k.variant == 2) ? k.threads : 2 * k.threads;

    int splits = h->cfg.source_splits;
    if (splits <= 0)
    {
        splits = 1;
        if (tgt_blocks < 8 * slots)
        {
            const long max_by_chunk = (long)(h->n / (size_t)(4 * tile));   // >= 4 tiles per chunk
            long max_splits = max_by_chunk < 1 ? 1 : (max_by_chunk > 64 ? 64 : max_by_chunk);
            // partial buffers: splits * 24 B per owned body, keep below 1 GiB
            const long max_by_mem = (long)((1ull << 30) / (24ull * (h->count ? h->count : 1)));
            if (max_by_mem < max_splits) max_splits = max_by_mem < 1 ? 1 : max_by_mem;
            double best = -1.0;
            for (long s = 1; s <= max_splits; ++s)
            {
                const double waves = (double)(tgt_blocks * s) / (double)slots;
                const double full = (double)((tgt_blocks * s + slots - 1) / slots);
                const double eff = waves / full;
                if (eff > best + 0.005) { best = eff; splits = (int)s; }
            }
        }
    }
    h->ap_splits = splits;
    return NB_OK;
}

int launch_allpairs(nb_sim* h)
{
    int count = 0;
    const AllPairsKernel* table = allpairs_table(&count);
    const AllPairsKernel& k = table[h->ap_kernel];
    const int splits = h->ap_splits;
    if (h->acc_part_splits < (size_t)splits)
    {
        cudaFree(h->acc_part);
        h->acc_part = nullptr;
        NB_CUDA(cudaMalloc(&h->acc_part, (size_t)splits * 3 * h->count * sizeof(double)));
        h->acc_part_splits = splits;
    }
    const int tile = (k.variant == 2) ? k.threads : 2 * k.threads;
    // chunk = ceil(n / splits) rounded up to whole tiles
    long chunk = ((long)h->n + splits - 1) / splits;
    chunk = (chunk + tile - 1) / tile * tile;
    const long tgt_blocks = ((long)h->count + (long)k.threads * k.targets - 1) / ((long)k.threads * k.targets);
    dim3 grid((unsigned)tgt_blocks, (unsigned)splits, 1);
    const float sc = (float)(h->cfg.softening * (double)kPreScale);
    k.fn<<<grid, k.threads, k.smem_bytes, h->stream>>>(h->posw, (int)h->n, (int)h->first, (int)h->count, (int)chunk,
                                           h->acc_part, sc, (float)h->cfg.softening, h->wmax);
    NB_CUDA(cudaGetLastError());
    ++h->last_launches;
    return NB_OK;
}

// `timed`: bracket the force pass in the next slot of the event ring -- {pass begin, dominant kernel
// (all-pairs kernel / tree walk) begin, dominant kernel end}; the tree build lies between the first two.
// `balanced`: Barnes-Hut inside nb_step with peer memory attached -- every rank walks an interleaved
// share of ALL targets and stores into the owners' arrays; the exchange that follows makes sure this
// rank's own accelerations are complete before the kick-drift reads them.
// Captures `body` (launches on h->stream only) into an executable graph.
template <typename F>
static int capture_graph(nb_sim* h, cudaGraphExec_t* out, F body)
{
    cudaGraph_t g = nullptr;
    NB_CUDA(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
    const int rc = body();
    const cudaError_t e = cudaStreamEndCapture(h->stream, &g);
    if (rc != NB_OK) { if (g) cudaGraphDestroy(g); return rc; }
    NB_CUDA(e);
    const cudaError_t ei = cudaGraphInstantiate(out, g, 0);
    cudaGraphDestroy(g);
    NB_CUDA(ei);
    return NB_OK;
}

static bool graphs_wanted(const nb_sim* h)
{
    static const bool enabled = [] { const char* v = std::getenv("NB_GRAPHS"); return v == nullptr || v[0] != '0'; }();
    return enabled && !h->graph_failed && h->cfg.world == 1 && h->cfg.mode == NB_MODE_BARNESHUT && h->n <= ((size_t)1 << 21);
}

// The force pass of a small one-GPU Barnes-Hut step as two graph launches (same kernels, same order, same arguments).
static int compute_forces_graphed(nb_sim* h)
{
    if (h->graph_build == nullptr || h->graph_n != h->n || h->graph_theta != h->cfg.theta)
    {
        graphs_release(h);
        NB_CHECK(tree_reserve(h));                     // allocations and attribute calls happen outside the capture
        const int before = h->last_launches;
        int rc = capture_graph(h, &h->graph_build, [&] { return tree_build(h, false); });
        if (rc == NB_OK) rc = capture_graph(h, &h->graph_walk, [&] { return tree_walk(h, false); });
        h->graph_launches = h->last_launches - before;
        h->last_launches = before;
        if (rc != NB_OK)
        {
            graphs_release(h);
            h->graph_failed = true;                    // this handle launches directly from now on
            cudaGetLastError();
            return rc;
        }
        h->graph_n = h->n;
        h->graph_theta = h->cfg.theta;
    }
    cudaEvent_t* slot = h->ring[h->ring_pos % NB_TIMING_RING];
    NB_CUDA(cudaEventRecord(slot[0], h->stream));
    NB_CUDA(cudaGraphLaunch(h->graph_build, h->stream));
    NB_CUDA(cudaEventRecord(slot[1], h->stream));
    NB_CUDA(cudaGraphLaunch(h->graph_walk, h->stream));
    NB_CUDA(cudaEventRecord(slot[2], h->stream));
    ++h->ring_pos;
    h->last_launches += h->graph_launches;
    h->tree.built = true;
    return NB_OK;
}

static int compute_forces(nb_sim* h, bool timed, bool balanced = false)
{
    if (timed && !balanced && graphs_wanted(h))
    {
        if (compute_forces_graphed(h) == NB_OK) return NB_OK;
        // capture failed: fall through to direct launches
    }
    cudaEvent_t* slot = h->ring[h->ring_pos % NB_TIMING_RING];
    if (timed) NB_CUDA(cudaEventRecord(slot[0], h->stream));
    if (h->cfg.mode == NB_MODE_ALLPAIRS)
    {
        if (timed) NB_CUDA(cudaEventRecord(slot[1], h->stream));
        NB_CHECK(launch_allpairs(h));
    }
    else
    {
        NB_CHECK(tree_build(h, balanced));
        if (timed) NB_CUDA(cudaEventRecord(slot[1], h->stream));
        NB_CHECK(tree_walk(h, balanced));
    }
    if (timed)
    {
        NB_CUDA(cudaEventRecord(slot[2], h->stream));
        ++h->ring_pos;
    }
    if (balanced && h->cfg.mode == NB_MODE_BARNESHUT) NB_CHECK(p2p_acc_exchange(h));
    return NB_OK;
}

static int set_bodies(nb_sim* h, size_t n)
{
    NB_REQUIRE(n > 0 && n < (size_t)0x7fffffff, NB_ERR_ARG, "body count must be in [1, 2^31)");
    if (n != h->n)
    {
        free_state(h);
        const size_t world = (size_t)h->cfg.world, rank = (size_t)h->cfg.rank;
        h->n = n;
        h->first = rank * n / world;
        h->count = (rank + 1) * n / world - h->first;
        NB_REQUIRE(h->count > 0, NB_ERR_ARG, "fewer bodies than ranks");
        NB_CUDA(cudaMalloc(&h->posw_buf[0], n * sizeof(float4)));
        h->posw = h->posw_buf[0];
        h->posw_cur = 0;
        NB_CUDA(cudaMalloc(&h->vel, 3 * h->count * sizeof(double)));
        NB_CUDA(cudaMalloc(&h->mass, h->count * sizeof(double)));
        NB_CUDA(cudaMalloc(&h->acc_base, 3 * h->count * sizeof(double)));
        h->acc = h->acc_base;
        h->acc_cur = 0;
        NB_CUDA(cudaMemsetAsync(h->acc, 0, 3 * h->count * sizeof(double), h->stream));
        if (h->cfg.mode == NB_MODE_BARNESHUT) NB_CHECK(tree_reserve(h));
    }
    NB_CHECK(choose_allpairs_config(h));
    h->acc_valid = false;
    h->forces_from_last_step = false;
    h->acc_is_last_step = false;
    h->exchanged = true;
    return NB_OK;
}

}  // namespace nb

using namespace nb;

extern "C" {

int nb_abi_version(void) { return NB_ABI_VERSION; }

const char* nb_last_error(void) { return nb::g_error; }

int nb_default_config(nb_config* cfg)
{
    NB_REQUIRE(cfg != nullptr, NB_ERR_ARG, "null config");
    std::memset(cfg, 0, sizeof(*cfg));
    cfg->struct_size = (uint32_t)sizeof(nb_config);
    cfg->device = 0;
    cfg->mode = NB_MODE_ALLPAIRS;
    cfg->theta = 2.0f;                 // Octree.cpp:5
    cfg->G = 6.674e-11;                // Physics.hpp:9
    cfg->softening = 1e1;              // Physics.hpp:10
    cfg->position_scale = 20 * 1.15e12;   // Physics.hpp:13,16
    cfg->bounds = 4000.0f;             // BarnesHut.cpp:14
    cfg->rank = 0;
    cfg->world = 1;
    cfg->stream = nullptr;
    cfg->source_splits = 0;
    cfg->kernel_variant = 0;
    return NB_OK;
}

int nb_create(const nb_config* cfg, nb_handle* out)
{
    NB_REQUIRE(cfg != nullptr && out != nullptr, NB_ERR_ARG, "null argument");
    NB_REQUIRE(cfg->struct_size == sizeof(nb_config), NB_ERR_ARG, "nb_config size mismatch (ABI)");
    NB_REQUIRE(cfg->mode == NB_MODE_ALLPAIRS || cfg->mode == NB_MODE_BARNESHUT, NB_ERR_ARG, "unknown mode");
    NB_REQUIRE(cfg->world >= 1 && cfg->rank >= 0 && cfg->rank < cfg->world, NB_ERR_ARG, "bad rank/world");
    NB_REQUIRE(cfg->theta > 0.f && cfg->bounds > 0.f && cfg->position_scale > 0.0, NB_ERR_ARG, "bad constants");
    *out = nullptr;

    int ndev = 0;
    NB_CUDA(cudaGetDeviceCount(&ndev));
    NB_REQUIRE(cfg->device >= 0 && cfg->device < ndev, NB_ERR_CUDA, "no such CUDA device");
    NB_CUDA(cudaSetDevice(cfg->device));
    cudaDeviceProp prop;
    NB_CUDA(cudaGetDeviceProperties(&prop, cfg->device));
    {
        // The library carries sm_100a SASS only; refuse anything else up front instead of failing
        // at the first launch.
        int count = 0;
        const AllPairsKernel* table = allpairs_table(&count);
        cudaFuncAttributes attr;
        cudaError_t e = cudaFuncGetAttributes(&attr, (const void*)table[0].fn);
        if (e != cudaSuccess)
        {
            nb::set_error("nb_create: no kernel image for device %d (%s, sm_%d%d): %s -- this library is built for sm_100a only",
                          cfg->device, prop.name, prop.major, prop.minor, cudaGetErrorString(e));
            cudaGetLastError();
            return NB_ERR_CUDA;
        }
    }

    nb_sim* h = new (std::nothrow) nb_sim();
    NB_REQUIRE(h != nullptr, NB_ERR_NOMEM, "out of host memory");
    h->cfg = *cfg;
    h->sm_count = prop.multiProcessorCount;
    if (cfg->stream != nullptr)
    {
        h->stream = static_cast<cudaStream_t>(cfg->stream);
        h->own_stream = false;
    }
    else
    {
        cudaError_t e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
        if (e != cudaSuccess) { delete h; nb::set_error("cudaStreamCreate: %s", cudaGetErrorString(e)); return NB_ERR_CUDA; }
        h->own_stream = true;
    }
    {
        cudaError_t e = cudaMalloc(&h->wmax, sizeof(float));
        if (e == cudaSuccess) e = cudaMemset(h->wmax, 0, sizeof(float));
        if (e != cudaSuccess) { nb::set_error("cudaMalloc: %s", cudaGetErrorString(e)); nb_destroy(h); return NB_ERR_CUDA; }
    }
    for (int i = 0; i < 2 + 3 * NB_TIMING_RING; ++i)
    {
        cudaEvent_t* slot = i < 2 ? &h->ev[i] : &h->ring[(i - 2) / 3][(i - 2) % 3];
        cudaError_t e = cudaEventCreate(slot);
        if (e != cudaSuccess) { nb::set_error("cudaEventCreate: %s", cudaGetErrorString(e)); nb_destroy(h); return NB_ERR_CUDA; }
    }
    *out = h;
    return NB_OK;
}

int nb_destroy(nb_handle h)
{
    if (h == nullptr) return NB_OK;
    cudaSetDevice(h->cfg.device);
    cudaStreamSynchronize(h->stream);
    comm_destroy(h);
    free_state(h);
    cudaFree(h->d_aos);
    cudaFree(h->wmax);
    for (int i = 0; i < 2; ++i)
        if (h->ev[i]) cudaEventDestroy(h->ev[i]);
    for (int i = 0; i < NB_TIMING_RING; ++i)
        for (int k = 0; k < 3; ++k)
            if (h->ring[i][k]) cudaEventDestroy(h->ring[i][k]);
    if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
    delete h;
    return NB_OK;
}

int nb_set_theta(nb_handle h, float theta)
{
    NB_REQUIRE(h != nullptr, NB_ERR_ARG, "null handle");
    NB_REQUIRE(theta > 0.f, NB_ERR_ARG, "theta must be positive");
    h->cfg.theta = theta;
    h->acc_valid = false;
    return NB_OK;
}

int nb_init_aos(nb_handle h, const void* particles, size_t n, size_t stride)
{
    NB_REQUIRE(h != nullptr && particles != nullptr, NB_ERR_ARG, "null argument");
    NB_REQUIRE(stride >= NB_PARTICLE_STRIDE && stride % 8 == 0, NB_ERR_ARG, "stride must be >= 104 and a multiple of 8");
    NB_CUDA(cudaSetDevice(h->cfg.device));
    NB_CHECK(set_bodies(h, n));
    NB_CHECK(reserve_aos(h, n * stride));
    NB_CUDA(cudaMemcpyAsync(h->d_aos, particles, n * stride, cudaMemcpyHostToDevice, h->stream));
    h->d_aos_stride = stride;
    h->last_launches = 0;
    NB_CHECK(launch_unpack_aos(h, stride, 0, n));
    NB_CUDA(cudaStreamSynchronize(h->stream));
    return NB_OK;
}

int nb_init_soa(nb_handle h, const float* pos3, const double* vel3, const double* mass, size_t n)
{
    NB_REQUIRE(h != nullptr && pos3 != nullptr && vel3 != nullptr && mass != nullptr, NB_ERR_ARG, "null argument");
    NB_CUDA(cudaSetDevice(h->cfg.device));
    NB_CHECK(set_bodies(h, n));
    // Build a packed 104-byte image on the host side of the copy and reuse the AoS path; Init is
    // not on the hot path (the reference re-Inits only on particle-count changes).
    const size_t stride = NB_PARTICLE_STRIDE;
    unsigned char* tmp = nullptr;
    NB_CUDA(cudaMallocHost(&tmp, n * stride));
    std::memset(tmp, 0, n * stride);
    for (size_t i = 0; i < n; ++i)
    {
        unsigned char* rec = tmp + i * stride;
        std::memcpy(rec + NB_OFF_POSITION, pos3 + 3 * i, 3 * sizeof(float));
        std::memcpy(rec + NB_OFF_VELOCITY, vel3 + 3 * i, 3 * sizeof(double));
        std::memcpy(rec + NB_OFF_MASS, mass + i, sizeof(double));
    }
    int rc = reserve_aos(h, n * stride);
    if (rc == NB_OK)
    {
        cudaError_t e = cudaMemcpyAsync(h->d_aos, tmp, n * stride, cudaMemcpyHostToDevice, h->stream);
        if (e != cudaSuccess) { nb::set_error("cudaMemcpyAsync: %s", cudaGetErrorString(e)); rc = NB_ERR_CUDA; }
    }
    if (rc == NB_OK) { h->d_aos_stride = stride; h->last_launches = 0; rc = launch_unpack_aos(h, stride, 0, n); }
    cudaStreamSynchronize(h->stream);
    cudaFreeHost(tmp);
    return rc;
}

int nb_step(nb_handle h, float dt, int nsteps)
{
    NB_REQUIRE(h != nullptr, NB_ERR_ARG, "null handle");
    NB_REQUIRE(h->n > 0, NB_ERR_STATE, "nb_init_* has not been called");
    NB_REQUIRE(nsteps >= 0, NB_ERR_ARG, "negative step count");
    NB_CUDA(cudaSetDevice(h->cfg.device));
    // checked BEFORE anything is launched: a rank that integrated but could not exchange would diverge
    NB_REQUIRE(h->p2p_attached || h->nccl_comm == nullptr || h->n % (size_t)h->cfg.world == 0, NB_ERR_ARG,
               "the NCCL exchange needs the body count to be divisible by the number of ranks (use the peer-memory exchange for ragged shards)");
    h->last_launches = 0;
    h->timing_valid = false;
    NB_CUDA(cudaEventRecord(h->ev[0], h->stream));
    for (int s = 0; s < nsteps; ++s)
    {
        NB_REQUIRE(h->exchanged, NB_ERR_STATE,
                   "world > 1 without nb_comm_init: call nb_mark_exchanged after exchanging positions");
        if (h->p2p_attached) NB_CHECK(p2p_wait(h));          // every peer's positions of the last step are in
        NB_CHECK(compute_forces(h, true, h->p2p_attached && h->cfg.world > 1 && h->cfg.mode == NB_MODE_BARNESHUT));
        h->acc_valid = false;
        h->forces_from_last_step = true;
        h->acc_is_last_step = true;
        if (h->p2p_attached)
        {
            NB_CHECK(p2p_kick_drift_push(h, dt));            // ONE kernel: integrate + store into every rank
        }
        else
        {
            NB_CHECK(launch_kick_drift(h, dt));
            if (h->cfg.world > 1)
            {
                if (h->nccl_comm != nullptr) NB_CHECK(comm_allgather_posw(h));
                else h->exchanged = false;
            }
        }
    }
    NB_CUDA(cudaEventRecord(h->ev[1], h->stream));
    h->timing_valid = nsteps > 0;
    h->total_launches += (unsigned long long)h->last_launches;
    return NB_OK;
}

int nb_update_aos(nb_handle h, void* particles, size_t n, size_t stride, float dt)
{
    NB_REQUIRE(h != nullptr && particles != nullptr, NB_ERR_ARG, "null argument");
    NB_REQUIRE(stride >= NB_PARTICLE_STRIDE && stride % 8 == 0, NB_ERR_ARG, "stride must be >= 104 and a multiple of 8");
    NB_CUDA(cudaSetDevice(h->cfg.device));
    // A single handle re-reads the whole array (the caller may have edited any record).  A shard
    // handle (world > 1) owns records [first, first+count) of the caller's array: it re-reads those
    // and keeps the other ranks' positions from the last exchange.
    const bool fresh = (n != h->n);
    if (fresh) NB_CHECK(set_bodies(h, n));
    NB_CHECK(reserve_aos(h, n * stride));
    h->d_aos_stride = stride;
    const bool whole = fresh || h->cfg.world == 1;
    const size_t begin = whole ? 0 : h->first, end = whole ? n : h->first + h->count;
    NB_REQUIRE(whole || h->exchanged, NB_ERR_STATE, "positions of remote ranks are stale");
    unsigned char* host_in = static_cast<unsigned char*>(particles);
    NB_CUDA(cudaMemcpyAsync(static_cast<unsigned char*>(h->d_aos) + begin * stride, host_in + begin * stride,
                            (end - begin) * stride, cudaMemcpyHostToDevice, h->stream));
    h->last_launches = 0;
    NB_CHECK(launch_unpack_aos(h, stride, begin, end));
    if (whole) h->exchanged = true;
    const int unpack_launches = h->last_launches;
    NB_CHECK(nb_step(h, dt, 1));
    NB_CHECK(launch_pack_aos(h, stride, h->cfg.mode == NB_MODE_ALLPAIRS));
    h->last_launches += unpack_launches;
    h->total_launches += (unsigned long long)unpack_launches + 1ull;
    unsigned char* host = static_cast<unsigned char*>(particles);
    NB_CUDA(cudaMemcpyAsync(host + h->first * stride, static_cast<unsigned char*>(h->d_aos) + h->first * stride,
                            h->count * stride, cudaMemcpyDeviceToHost, h->stream));
    NB_CUDA(cudaStreamSynchronize(h->stream));
    return NB_OK;
}

int nb_host_register(void* ptr, size_t bytes)
{
    NB_REQUIRE(ptr != nullptr && bytes > 0, NB_ERR_ARG, "null argument");
    NB_CUDA(cudaHostRegister(ptr, bytes, cudaHostRegisterDefault));
    return NB_OK;
}

int nb_host_unregister(void* ptr)
{
    NB_REQUIRE(ptr != nullptr, NB_ERR_ARG, "null argument");
    NB_CUDA(cudaHostUnregister(ptr));
    return NB_OK;
}

int nb_sync(nb_handle h)
{
    NB_REQUIRE(h != nullptr, NB_ERR_ARG, "null handle");
    const cudaError_t e = cudaStreamSynchronize(h->stream);
    if (e != cudaSuccess)
    {
        char why[256];
        if (p2p_describe_timeout(h, why, sizeof(why))) nb::set_error("nb_sync: %s (%s)", why, cudaGetErrorString(e));
        else nb::set_error("nb_sync: %s", cudaGetErrorString(e));
        return NB_ERR_CUDA;
    }
    return NB_OK;
}

int nb_read_aos(nb_handle h, void* particles, size_t n, size_t stride)
{
    NB_REQUIRE(h != nullptr && particles != nullptr, NB_ERR_ARG, "null argument");
    NB_REQUIRE(h->n > 0 && n == h->n, NB_ERR_ARG, "array length differs from the initialised body count");
    NB_REQUIRE(stride >= NB_PARTICLE_STRIDE && stride % 8 == 0, NB_ERR_ARG, "stride must be >= 104 and a multiple of 8");
    NB_CUDA(cudaSetDevice(h->cfg.device));
    NB_CHECK(reserve_aos(h, n * stride));
    // Seed the device image with the caller's records so untouched fields survive the round trip.
    unsigned char* host = static_cast<unsigned char*>(particles);
    unsigned char* dev = static_cast<unsigned char*>(h->d_aos);
    NB_CUDA(cudaMemcpyAsync(dev + h->first * stride, host + h->first * stride, h->count * stride,
                            cudaMemcpyHostToDevice, h->stream));
    const bool zero = (h->cfg.mode == NB_MODE_ALLPAIRS) || !h->forces_from_last_step;
    NB_CHECK(launch_pack_aos(h, stride, zero));
    NB_CUDA(cudaMemcpyAsync(host + h->first * stride, dev + h->first * stride, h->count * stride,
                            cudaMemcpyDeviceToHost, h->stream));
    NB_CUDA(cudaStreamSynchronize(h->stream));
    return NB_OK;
}

int nb_read_soa(nb_handle h, float* pos3, double* vel3)
{
    NB_REQUIRE(h != nullptr, NB_ERR_ARG, "null handle");
    NB_REQUIRE(h->n > 0, NB_ERR_STATE, "not initialised");
    NB_CUDA(cudaSetDevice(h->cfg.device));
    NB_CUDA(cudaStreamSynchronize(h->stream));
    if (pos3 != nullptr)
    {
        float4* tmp = nullptr;
        NB_CUDA(cudaMallocHost(&tmp, h->count * sizeof(float4)));
        cudaError_t e = cudaMemcpy(tmp, h->posw + h->first, h->count * sizeof(float4), cudaMemcpyDeviceToHost);
        if (e == cudaSuccess)
            for (size_t i = 0; i < h->count; ++i)
            {
                pos3[3 * i] = tmp[i].x; pos3[3 * i + 1] = tmp[i].y; pos3[3 * i + 2] = tmp[i].z;
            }
        cudaFreeHost(tmp);
        NB_CUDA(e);
    }
    if (vel3 != nullptr)
    {
        double* tmp = nullptr;
        NB_CUDA(cudaMallocHost(&tmp, 3 * h->count * sizeof(double)));
        cudaError_t e = cudaMemcpy(tmp, h->vel, 3 * h->count * sizeof(double), cudaMemcpyDeviceToHost);
        if (e == cudaSuccess)
            for (size_t i = 0; i < h->count; ++i)
            {
                vel3[3 * i] = tmp[i];
                vel3[3 * i + 1] = tmp[h->count + i];
                vel3[3 * i + 2] = tmp[2 * h->count + i];
            }
        cudaFreeHost(tmp);
        NB_CUDA(e);
    }
    return NB_OK;
}

int nb_owned_range(nb_handle h, size_t* first, size_t* count)
{
    NB_REQUIRE(h != nullptr, NB_ERR_ARG, "null handle");
    if (first) *first = h->first;
    if (count) *count = h->count;
    return NB_OK;
}

int nb_shard_range(size_t n, int rank, int world, size_t* first, size_t* count)
{
    NB_REQUIRE(world >= 1 && rank >= 0 && rank < world, NB_ERR_ARG, "bad rank/world");
    const size_t f = (size_t)rank * n / (size_t)world;
    if (first) *first = f;
    if (count) *count = ((size_t)rank + 1) * n / (size_t)world - f;
    return NB_OK;
}

int nb_num_bodies(nb_handle h, size_t* n)
{
    NB_REQUIRE(h != nullptr && n != nullptr, NB_ERR_ARG, "null argument");
    *n = h->n;
    return NB_OK;
}

int nb_compute_accel(nb_handle h)
{
    NB_REQUIRE(h != nullptr, NB_ERR_ARG, "null handle");
    NB_REQUIRE(h->n > 0, NB_ERR_STATE, "not initialised");
    NB_REQUIRE(h->exchanged, NB_ERR_STATE, "positions of remote ranks are stale");
    NB_CUDA(cudaSetDevice(h->cfg.device));
    h->last_launches = 0;
    h->timing_valid = false;
    if (h->p2p_attached) NB_CHECK(p2p_wait(h));
    NB_CUDA(cudaEventRecord(h->ev[0], h->stream));
    NB_CHECK(compute_forces(h, true));
    if (h->cfg.mode == NB_MODE_ALLPAIRS) NB_CHECK(launch_reduce_partials(h));
    NB_CUDA(cudaEventRecord(h->ev[1], h->stream));
    h->timing_valid = true;
    h->acc_valid = true;
    h->acc_is_last_step = false;
    h->total_launches += (unsigned long long)h->last_launches;
    return NB_OK;
}

int nb_get_accel(nb_handle h, double* acc3)
{
    NB_REQUIRE(h != nullptr && acc3 != nullptr, NB_ERR_ARG, "null argument");
    NB_REQUIRE(h->n > 0, NB_ERR_STATE, "not initialised");
    NB_CUDA(cudaSetDevice(h->cfg.device));
    // acc_valid is cleared by every step, Init and theta change: what is returned always belongs to the
    // CURRENT positions and theta (forces_from_last_step only decides what the Forces write-back holds)
    if (!h->acc_valid) NB_CHECK(nb_compute_accel(h));
    NB_CUDA(cudaStreamSynchronize(h->stream));
    double* tmp = nullptr;
    NB_CUDA(cudaMallocHost(&tmp, 3 * h->count * sizeof(double)));
    cudaError_t e = cudaMemcpy(tmp, h->acc, 3 * h->count * sizeof(double), cudaMemcpyDeviceToHost);
    if (e == cudaSuccess)
        for (size_t i = 0; i < h->count; ++i)
        {
            acc3[3 * i] = tmp[i];
            acc3[3 * i + 1] = tmp[h->count + i];
            acc3[3 * i + 2] = tmp[2 * h->count + i];
        }
    cudaFreeHost(tmp);
    NB_CUDA(e);
    return NB_OK;
}

int nb_energy(nb_handle h, double* kinetic, double* potential)
{
    NB_REQUIRE(h != nullptr, NB_ERR_ARG, "null handle");
    NB_REQUIRE(h->n > 0, NB_ERR_STATE, "not initialised");
    NB_CUDA(cudaSetDevice(h->cfg.device));
    if (h->p2p_attached) NB_CHECK(p2p_wait(h));
    return nb::energy(h, kinetic, potential);
}

int nb_energy_sampled(nb_handle h, size_t stride, double* kinetic, double* potential, size_t* nsamples)
{
    NB_REQUIRE(h != nullptr, NB_ERR_ARG, "null handle");
    NB_REQUIRE(stride >= 1, NB_ERR_ARG, "stride must be >= 1");
    NB_REQUIRE(h->n > 0, NB_ERR_STATE, "not initialised");
    NB_REQUIRE(h->exchanged, NB_ERR_STATE, "positions of remote ranks are stale");
    NB_CUDA(cudaSetDevice(h->cfg.device));
    if (h->p2p_attached) NB_CHECK(p2p_wait(h));
    return nb::energy_sampled(h, stride, kinetic, potential, nsamples);
}

int nb_comm_unique_id(uint8_t id[128]) { return nb::comm_unique_id(id); }

int nb_comm_init(nb_handle h, const uint8_t id[128])
{
    NB_REQUIRE(h != nullptr && id != nullptr, NB_ERR_ARG, "null argument");
    NB_CUDA(cudaSetDevice(h->cfg.device));
    return nb::comm_init(h, id);
}

int nb_device_posw(nb_handle h, void** dev_ptr, size_t* bytes)
{
    NB_REQUIRE(h != nullptr && dev_ptr != nullptr, NB_ERR_ARG, "null argument");
    NB_REQUIRE(h->n > 0, NB_ERR_STATE, "not initialised");
    *dev_ptr = h->posw;
    if (bytes) *bytes = h->n * sizeof(float4);
    return NB_OK;
}

int nb_mark_exchanged(nb_handle h)
{
    NB_REQUIRE(h != nullptr, NB_ERR_ARG, "null handle");
    h->exchanged = true;
    return NB_OK;
}

int nb_last_step_timing(nb_handle h, float* total_ms, float* force_kernel_ms, int* launches)
{
    NB_REQUIRE(h != nullptr, NB_ERR_ARG, "null handle");
    NB_REQUIRE(h->timing_valid, NB_ERR_STATE, "no timed call yet");
    NB_CUDA(cudaSetDevice(h->cfg.device));
    NB_CUDA(cudaEventSynchronize(h->ev[1]));
    if (total_ms) NB_CUDA(cudaEventElapsedTime(total_ms, h->ev[0], h->ev[1]));
    if (force_kernel_ms)
    {
        NB_REQUIRE(h->ring_pos > 0, NB_ERR_STATE, "no timed force pass yet");
        cudaEvent_t* slot = h->ring[(h->ring_pos - 1) % NB_TIMING_RING];
        NB_CUDA(cudaEventElapsedTime(force_kernel_ms, slot[1], slot[2]));
    }
    if (launches) *launches = h->last_launches;
    return NB_OK;
}

int nb_step_timing_mean(nb_handle h, int max_steps, float* force_kernel_ms, float* build_ms, int* steps_averaged)
{
    NB_REQUIRE(h != nullptr, NB_ERR_ARG, "null handle");
    NB_REQUIRE(h->timing_valid && h->ring_pos > 0, NB_ERR_STATE, "no timed call yet");
    NB_CUDA(cudaSetDevice(h->cfg.device));
    NB_CUDA(cudaEventSynchronize(h->ev[1]));
    unsigned long long k = h->ring_pos < (unsigned long long)NB_TIMING_RING ? h->ring_pos : (unsigned long long)NB_TIMING_RING;
    if (max_steps > 0 && (unsigned long long)max_steps < k) k = (unsigned long long)max_steps;
    double kernel = 0.0, build = 0.0;
    for (unsigned long long i = 0; i < k; ++i)
    {
        cudaEvent_t* slot = h->ring[(h->ring_pos - 1 - i) % NB_TIMING_RING];
        float a = 0.f, b = 0.f;
        NB_CUDA(cudaEventElapsedTime(&a, slot[1], slot[2]));
        NB_CUDA(cudaEventElapsedTime(&b, slot[0], slot[1]));
        kernel += a;
        build += b;
    }
    if (force_kernel_ms) *force_kernel_ms = (float)(kernel / (double)k);
    if (build_ms) *build_ms = (float)(build / (double)k);
    if (steps_averaged) *steps_averaged = (int)k;
    return NB_OK;
}

int nb_step_period_mean(nb_handle h, int max_steps, float* period_ms, int* steps_averaged)
{
    NB_REQUIRE(h != nullptr && period_ms != nullptr, NB_ERR_ARG, "null argument");
    NB_REQUIRE(h->timing_valid && h->ring_pos > 1, NB_ERR_STATE, "fewer than two timed force passes");
    NB_CUDA(cudaSetDevice(h->cfg.device));
    NB_CUDA(cudaEventSynchronize(h->ev[1]));
    unsigned long long k = h->ring_pos < (unsigned long long)NB_TIMING_RING ? h->ring_pos : (unsigned long long)NB_TIMING_RING;
    if (max_steps > 0 && (unsigned long long)max_steps < k) k = (unsigned long long)max_steps;
    NB_REQUIRE(k > 1, NB_ERR_STATE, "fewer than two timed force passes");
    float ms = 0.f;
    NB_CUDA(cudaEventElapsedTime(&ms, h->ring[(h->ring_pos - k) % NB_TIMING_RING][0], h->ring[(h->ring_pos - 1) % NB_TIMING_RING][0]));
    *period_ms = ms / (float)(k - 1);
    if (steps_averaged) *steps_averaged = (int)(k - 1);
    return NB_OK;
}

int nb_last_build_timing(nb_handle h, float* build_ms)
{
    NB_REQUIRE(h != nullptr && build_ms != nullptr, NB_ERR_ARG, "null argument");
    NB_REQUIRE(h->timing_valid, NB_ERR_STATE, "no timed call yet");
    NB_CUDA(cudaSetDevice(h->cfg.device));
    NB_CUDA(cudaEventSynchronize(h->ev[1]));
    NB_REQUIRE(h->ring_pos > 0, NB_ERR_STATE, "no timed force pass yet");
    cudaEvent_t* slot = h->ring[(h->ring_pos - 1) % NB_TIMING_RING];
    NB_CUDA(cudaEventElapsedTime(build_ms, slot[0], slot[1]));
    return NB_OK;
}

int nb_probe_fp32_peak(nb_handle h, double* flops_per_s)
{
    NB_REQUIRE(h != nullptr && flops_per_s != nullptr, NB_ERR_ARG, "null argument");
    NB_CUDA(cudaSetDevice(h->cfg.device));
    return nb::probe_fp32_peak(h, flops_per_s);
}

int nb_seed_galaxy_host(void* particles, size_t n, size_t stride, uint64_t seed, float scale)
{
    NB_REQUIRE(particles != nullptr, NB_ERR_ARG, "null argument");
    NB_REQUIRE(stride >= NB_PARTICLE_STRIDE && stride % 8 == 0, NB_ERR_ARG, "stride must be >= 104 and a multiple of 8");
    return nb::seed_galaxy_host(particles, n, stride, seed, scale);
}

int nb_seed_galaxy_device(nb_handle h, size_t n, uint64_t seed, float scale)
{
    NB_REQUIRE(h != nullptr, NB_ERR_ARG, "null handle");
    NB_CUDA(cudaSetDevice(h->cfg.device));
    NB_CHECK(set_bodies(h, n));
    return nb::seed_galaxy_device(h, n, seed, scale);
}

int nb_seed_collision_device(nb_handle h, size_t n, uint64_t seed, float scale, float separation, double approach_speed)
{
    NB_REQUIRE(h != nullptr, NB_ERR_ARG, "null handle");
    NB_REQUIRE(n >= 2, NB_ERR_ARG, "the collision scene needs at least two bodies");
    NB_CUDA(cudaSetDevice(h->cfg.device));
    NB_CHECK(set_bodies(h, n));
    return nb::seed_collision_device(h, n, seed, scale, separation, approach_speed);
}

int nb_enable_graphs(nb_handle h, int on)
{
    NB_REQUIRE(h != nullptr, NB_ERR_ARG, "null handle");
    h->graph_failed = (on == 0);
    if (!on) graphs_release(h);
    return NB_OK;
}

int nb_scale_masses(nb_handle h, double factor)
{
    NB_REQUIRE(h != nullptr, NB_ERR_ARG, "null handle");
    NB_REQUIRE(h->n > 0 && h->d_aos != nullptr && h->d_aos_stride >= NB_PARTICLE_STRIDE, NB_ERR_STATE,
               "no device image of the Particle array (nb_init_aos / nb_seed_*_device create it)");
    NB_CUDA(cudaSetDevice(h->cfg.device));
    NB_CHECK(nb::launch_scale_masses(h, factor));
    h->last_launches = 0;
    NB_CHECK(launch_unpack_aos(h, h->d_aos_stride, 0, h->n));
    h->acc_valid = false;
    h->tree.built = false;
    NB_CUDA(cudaStreamSynchronize(h->stream));
    return NB_OK;
}

int nb_get_aos_records(nb_handle h, const uint32_t* bodies, size_t k, void* records)
{
    NB_REQUIRE(h != nullptr && bodies != nullptr && records != nullptr, NB_ERR_ARG, "null argument");
    NB_REQUIRE(h->n > 0 && h->d_aos != nullptr && h->d_aos_stride >= NB_PARTICLE_STRIDE, NB_ERR_STATE,
               "no device image of the Particle array (nb_init_aos / nb_seed_*_device / nb_update_aos create it)");
    NB_CUDA(cudaSetDevice(h->cfg.device));
    NB_CUDA(cudaStreamSynchronize(h->stream));
    for (size_t i = 0; i < k; ++i)
    {
        NB_REQUIRE(bodies[i] < h->n, NB_ERR_ARG, "body index out of range");
        NB_CUDA(cudaMemcpy(static_cast<unsigned char*>(records) + i * NB_PARTICLE_STRIDE,
                           static_cast<const unsigned char*>(h->d_aos) + (size_t)bodies[i] * h->d_aos_stride, NB_PARTICLE_STRIDE,
                           cudaMemcpyDeviceToHost));
    }
    return NB_OK;
}

}  // extern "C"
