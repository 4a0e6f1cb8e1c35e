// Fused kick-drift + position exchange over NVLink peer memory.
//
// The baseline exchange is "kick-drift kernel, then ncclAllGather" (nccl_dl.cpp).  This variant is
// ONE kernel for both: every thread integrates its body and stores the new float4 {x, y, z, G m}
// straight into the position array of EVERY rank (its own and P-1 peer-mapped ones), so the bytes
// cross NVLink / NVSwitch while the rest of the grid is still integrating, and no separate
// collective is launched.  What remains of the collective is its synchronisation: a per-step flag
// each rank raises on all peers once its stores are out (release, system scope) and a one-block
// wait at the head of the next force pass (acquire).
//
// Positions are double-buffered: step s reads buffer s&1 and writes buffer (s+1)&1 everywhere.  A
// rank can only reach the kick-drift of step s+1 (which overwrites buffer s&1 on its peers) after
// every peer has raised flag s+1, i.e. after every peer has finished its own step-s force pass over
// buffer s&1 -- so no rank ever writes into a buffer a peer is still reading.
//
// Peers are attached either through CUDA IPC handles (one process per GPU; the 192 bytes per rank
// travel over the launcher's torch.distributed / MPI) or, for handles living in one process, by
// handle (nb_p2p_attach_local; used by the single-GPU test).
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "nb_internal.h"

namespace nb
{

struct PeerTable
{
    float4* posw[2][NB_MAX_PEERS];
    unsigned int* flags[NB_MAX_PEERS];
    unsigned long long* skeys[NB_MAX_PEERS];
    unsigned int* svals[NB_MAX_PEERS];
    int world, rank;
};

// rows of the flag array; the segment sizes of the sharded sort follow the rows
enum { ROW_POS = 0, ROW_ACC = 1, ROW_SORT_COUNT = 2, ROW_SORT_DATA = 3, FLAG_ROWS = 4 };
constexpr int kFlagWords = (FLAG_ROWS + 2) * NB_MAX_PEERS;
constexpr int kSegCountAt = FLAG_ROWS * NB_MAX_PEERS;          // bodies every rank sorted this step ...
constexpr int kSegOobAt = (FLAG_ROWS + 1) * NB_MAX_PEERS;      // ... and how many of them lie outside the cube

__global__ void __launch_bounds__(256)
k_kick_drift_push(PeerTable pt, int cur, int first, int count, double* __restrict__ vel,
                  const double* __restrict__ acc_part, int splits, double* __restrict__ acc, double dt,
                  double pos_scale)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const size_t plane = (size_t)count;
    double ax, ay, az;
    if (acc_part != nullptr)
    {
        ax = ay = az = 0.0;
        for (int s = 0; s < splits; ++s)
        {
            const double* p = acc_part + (size_t)s * 3 * plane;
            ax += p[i]; ay += p[plane + i]; az += p[2 * plane + i];
        }
        acc[i] = ax; acc[plane + i] = ay; acc[2 * plane + i] = az;
    }
    else
    {
        ax = acc[i]; ay = acc[plane + i]; az = acc[2 * plane + i];
    }
    double vx = vel[i], vy = vel[plane + i], vz = vel[2 * plane + i];
    // Velocity += a * dt: a rounded product, then a rounded sum, as the reference's g++ build (-ffp-contract=off)
    // computes it -- not a DFMA, which differs by an ulp every few steps
    vx = __dadd_rn(vx, __dmul_rn(ax, dt)); vy = __dadd_rn(vy, __dmul_rn(ay, dt)); vz = __dadd_rn(vz, __dmul_rn(az, dt));
    vel[i] = vx; vel[plane + i] = vy; vel[2 * plane + i] = vz;
    float4 p = pt.posw[cur][pt.rank][first + i];
    p.x += (float)((vx * dt) / pos_scale);
    p.y += (float)((vy * dt) / pos_scale);
    p.z += (float)((vz * dt) / pos_scale);
    const int nxt = cur ^ 1;
#pragma unroll 1
    for (int r = 0; r < pt.world; ++r)
    {
        // start with the own copy, then walk the peers round-robin so the ranks do not all hammer
        // the same destination GPU at the same time
        const int dst = (pt.rank + r) % pt.world;
        pt.posw[nxt][dst][first + i] = p;
    }
}

// Raised after the producing kernel has completed (stream order): my stores of `step` are out.
// `slot` selects the counter row: 0 = positions (after the push), NB_MAX_PEERS = accelerations (after a
// balanced walk).
__global__ void k_signal(PeerTable pt, unsigned int step, int slot)
{
    const int r = threadIdx.x;
    if (r < pt.world)
    {
        __threadfence_system();
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(pt.flags[r] + slot + pt.rank), "r"(step) : "memory");
    }
}

// Every peer has raised its counter to `step` (positions: head of the next force pass; accelerations:
// before the kick-drift; sort rows: see p2p_sort_exchange).  A peer that never arrives must not hang the
// GPU: after `timeout_ns` the waiter records what it was waiting for in pinned host memory and traps, so
// the host gets an error (and a message) instead of a wedged device.
struct WaitReport { int rank, row, peer; unsigned int want, have; int tripped; };

__global__ void k_wait(const unsigned int* my_flags, int world, unsigned int step, int rank, int row,
                       unsigned long long timeout_ns, WaitReport* report)
{
    const int r = threadIdx.x;
    if (r < world)
    {
        unsigned long long t0 = 0ull;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        unsigned int v, spins = 0;
        do
        {
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(my_flags + r) : "memory");
            if (v < step && (++spins & 1023u) == 0u)
            {
                unsigned long long t1;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
                if (t1 - t0 > timeout_ns)
                {
                    if (report != nullptr && atomicExch(&report->tripped, 1) == 0)
                    {
                        report->rank = rank; report->row = row; report->peer = r; report->want = step; report->have = v;
                        __threadfence_system();
                    }
                    __trap();
                }
            }
        } while (v < step);
    }
}

static unsigned long long wait_timeout_ns()
{
    // default 10 minutes: far above any legitimate wait (one all-pairs step of 16 M bodies on 2 GPUs is ~1 min)
    static unsigned long long ns = 0ull;
    if (ns == 0ull)
    {
        const char* e = std::getenv("NB_P2P_TIMEOUT_MS");
        const double ms = e ? std::atof(e) : 600000.0;
        ns = (unsigned long long)((ms > 1.0 ? ms : 1.0) * 1.0e6);
    }
    return ns;
}

static int launch_wait(nb_sim* h, int row, unsigned int step)
{
    k_wait<<<1, 32, 0, h->stream>>>(h->p2p_flags + row * NB_MAX_PEERS, h->cfg.world, step, h->cfg.rank, row, wait_timeout_ns(),
                                    static_cast<WaitReport*>(h->p2p_report));
    NB_CUDA(cudaGetLastError());
    return NB_OK;
}

// Text for nb_last_error when a wait tripped (the CUDA error itself only says "unspecified launch failure").
bool p2p_describe_timeout(const nb_sim* h, char* out, size_t cap)
{
    const WaitReport* r = static_cast<const WaitReport*>(h->p2p_report);
    if (r == nullptr || !r->tripped) return false;
    static const char* rows[] = {"positions", "accelerations", "sort counts", "sort data"};
    std::snprintf(out, cap, "peer exchange timed out: rank %d waited for the %s of rank %d, step %u (has %u)", r->rank,
                  r->row >= 0 && r->row < 4 ? rows[r->row] : "?", r->peer, r->want, r->have);
    return true;
}

int p2p_prepare(nb_sim* h)
{
    if (h->posw_buf[1] == nullptr)
    {
        NB_CUDA(cudaMalloc(&h->posw_buf[1], h->n * sizeof(float4)));
        NB_CUDA(cudaMemcpyAsync(h->posw_buf[1], h->posw_buf[0], h->n * sizeof(float4), cudaMemcpyDeviceToDevice, h->stream));
    }
    if (!h->acc_two)
    {
        // Two acceleration buffers: a balanced walk of step s stores into buffer s & 1 of every owner while
        // the owner may still be reading step s-1's (Forces write-back after its kick-drift).
        double* two = nullptr;
        const size_t one = 3 * h->count * sizeof(double);
        NB_CUDA(cudaMalloc(&two, 2 * one));
        NB_CUDA(cudaMemcpyAsync(two, h->acc, one, cudaMemcpyDeviceToDevice, h->stream));
        NB_CUDA(cudaMemsetAsync(reinterpret_cast<unsigned char*>(two) + one, 0, one, h->stream));
        NB_CUDA(cudaStreamSynchronize(h->stream));
        cudaFree(h->acc_base);
        h->acc_base = two;
        h->acc = two;
        h->acc_cur = 0;
        h->acc_two = true;
    }
    if (h->p2p_report == nullptr)
    {
        NB_CUDA(cudaHostAlloc(&h->p2p_report, sizeof(WaitReport), cudaHostAllocMapped | cudaHostAllocPortable));
        std::memset(h->p2p_report, 0, sizeof(WaitReport));
    }
    if (h->p2p_flags == nullptr)
    {
        NB_CUDA(cudaMalloc(&h->p2p_flags, kFlagWords * sizeof(unsigned int)));
    }
    NB_CUDA(cudaMemsetAsync(h->p2p_flags, 0, kFlagWords * sizeof(unsigned int), h->stream));
    if (h->cfg.mode == NB_MODE_BARNESHUT && h->tree.keys_final == nullptr)
    {
        // where every rank's sorted segment lands (sharded sort of the Barnes-Hut build)
        NB_CHECK(tree_reserve(h));
        NB_CUDA(cudaMalloc(&h->tree.keys_final, h->n * sizeof(unsigned long long)));
        NB_CUDA(cudaMalloc(&h->tree.vals_final, h->n * sizeof(unsigned int)));
        NB_CUDA(cudaMalloc(&h->tree.splitters, (NB_MAX_PEERS + 1) * sizeof(unsigned long long)));
    }
    NB_CUDA(cudaStreamSynchronize(h->stream));
    return NB_OK;
}

static PeerTable make_table(const nb_sim* h)
{
    PeerTable pt;
    std::memset(&pt, 0, sizeof(pt));
    pt.world = h->cfg.world;
    pt.rank = h->cfg.rank;
    for (int r = 0; r < h->cfg.world; ++r)
    {
        pt.posw[0][r] = static_cast<float4*>(h->peer_posw[0][r]);
        pt.posw[1][r] = static_cast<float4*>(h->peer_posw[1][r]);
        pt.flags[r] = static_cast<unsigned int*>(h->peer_flags[r]);
        pt.skeys[r] = static_cast<unsigned long long*>(h->peer_skeys[r]);
        pt.svals[r] = static_cast<unsigned int*>(h->peer_svals[r]);
    }
    return pt;
}

int p2p_wait(nb_sim* h)
{
    if (h->p2p_step == 0) return NB_OK;
    NB_CHECK(launch_wait(h, ROW_POS, h->p2p_step));
    ++h->last_launches;
    return NB_OK;
}

int p2p_kick_drift_push(nb_sim* h, float dt)
{
    const PeerTable pt = make_table(h);
    const bool partials = (h->cfg.mode == NB_MODE_ALLPAIRS);
    const int blocks = (int)((h->count + 255) / 256);
    k_kick_drift_push<<<blocks, 256, 0, h->stream>>>(pt, h->posw_cur, (int)h->first, (int)h->count, h->vel,
                                                     partials ? h->acc_part : nullptr, h->ap_splits, h->acc, (double)dt,
                                                     h->cfg.position_scale);
    NB_CUDA(cudaGetLastError());
    ++h->p2p_step;
    k_signal<<<1, 32, 0, h->stream>>>(pt, h->p2p_step, ROW_POS * NB_MAX_PEERS);
    NB_CUDA(cudaGetLastError());
    h->last_launches += 2;
    h->posw_cur ^= 1;
    h->posw = h->posw_buf[h->posw_cur];
    return NB_OK;
}

int p2p_acc_table(const nb_sim* h, AccTable* out)
{
    std::memset(out, 0, sizeof(*out));
    out->world = h->cfg.world;
    out->rank = h->cfg.rank;
    for (int r = 0; r < h->cfg.world; ++r)
    {
        NB_REQUIRE(h->peer_acc[r] != nullptr, NB_ERR_STATE, "peer acceleration arrays are not attached");
        out->acc[r] = static_cast<double*>(h->peer_acc[r]);
        out->first[r] = (int)((size_t)r * h->n / (size_t)h->cfg.world);
    }
    out->first[h->cfg.world] = (int)h->n;
    out->parity = (int)((h->p2p_acc_step + 1u) & 1u);     // the step p2p_acc_exchange is about to count
    return NB_OK;
}

// After a balanced walk: tell every rank that my share of ITS accelerations is stored, then wait until
// every rank has said the same about mine.
int p2p_acc_exchange(nb_sim* h)
{
    const PeerTable pt = make_table(h);
    ++h->p2p_acc_step;
    h->acc_cur = (int)(h->p2p_acc_step & 1u);             // what the walk just filled (p2p_acc_table's parity)
    h->acc = h->acc_base + (size_t)h->acc_cur * 3 * h->count;
    k_signal<<<1, 32, 0, h->stream>>>(pt, h->p2p_acc_step, ROW_ACC * NB_MAX_PEERS);
    NB_CHECK(launch_wait(h, ROW_ACC, h->p2p_acc_step));
    NB_CUDA(cudaGetLastError());
    h->last_launches += 2;
    return NB_OK;
}

// ---- sharded sort of the Barnes-Hut build -------------------------------------------------------
// Every rank has sorted the in-bounds bodies whose Morton key falls into ITS key range plus the
// out-of-bounds bodies (key ~0) of ITS index range (tree.cu); locally the latter sort to the end, in body
// order.  The key ranges tile the key space in rank order, so the global sorted array is
//     [in-bounds part of rank 0] ... [in-bounds part of rank P-1] [out-of-bounds part of rank 0] ... [of rank P-1]:
// (1) every rank tells every rank its two sizes, (2) every rank stores its two parts at their offsets into
// every rank's final arrays.  No collective is launched; the bytes cross NVLink as plain stores.
__global__ void k_publish_count(PeerTable pt, const unsigned int* __restrict__ count_dev, unsigned int step)
{
    const int r = threadIdx.x;
    if (r < pt.world)
    {
        pt.flags[r][kSegCountAt + pt.rank] = count_dev[0];
        pt.flags[r][kSegOobAt + pt.rank] = count_dev[1];
        __threadfence_system();
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(pt.flags[r] + ROW_SORT_COUNT * NB_MAX_PEERS + pt.rank), "r"(step) : "memory");
    }
}

__global__ void __launch_bounds__(256)
k_push_segment(PeerTable pt, const unsigned long long* __restrict__ keys, const unsigned int* __restrict__ vals,
               const unsigned int* seg_count /* my copy of every rank's segment size */)
{
    const unsigned int* seg_oob = seg_count + (kSegOobAt - kSegCountAt);
    unsigned int inb_before = 0, inb_all = 0, oob_before = 0;
    for (int q = 0; q < pt.world; ++q)
    {
        const unsigned int inb = seg_count[q] - seg_oob[q];
        if (q < pt.rank) { inb_before += inb; oob_before += seg_oob[q]; }
        inb_all += inb;
    }
    const unsigned int cnt = seg_count[pt.rank], inb_mine = cnt - seg_oob[pt.rank];
    for (unsigned int j = blockIdx.x * 256u + threadIdx.x; j < cnt; j += gridDim.x * 256u)
    {
        const unsigned long long k = keys[j];
        const unsigned int v = vals[j];
        const unsigned int at = j < inb_mine ? inb_before + j : inb_all + oob_before + (j - inb_mine);
#pragma unroll 1
        for (int r = 0; r < pt.world; ++r)
        {
            const int dst = (pt.rank + r) % pt.world;
            pt.skeys[dst][at] = k;
            pt.svals[dst][at] = v;
        }
    }
}

int p2p_sort_exchange(nb_sim* h, const unsigned long long* keys_local, const unsigned int* vals_local, const unsigned int* count_dev)
{
    const PeerTable pt = make_table(h);
    for (int r = 0; r < h->cfg.world; ++r)
        NB_REQUIRE(pt.skeys[r] != nullptr && pt.svals[r] != nullptr, NB_ERR_STATE, "peer sort buffers are not attached");
    ++h->p2p_sort_step;
    cudaStream_t st = h->stream;
    k_publish_count<<<1, 32, 0, st>>>(pt, count_dev, h->p2p_sort_step);
    NB_CHECK(launch_wait(h, ROW_SORT_COUNT, h->p2p_sort_step));
    const int blocks = (int)std::min<size_t>((h->n + 255) / 256, (size_t)h->sm_count * 8);
    k_push_segment<<<blocks, 256, 0, st>>>(pt, keys_local, vals_local, h->p2p_flags + kSegCountAt);
    k_signal<<<1, 32, 0, st>>>(pt, h->p2p_sort_step, ROW_SORT_DATA * NB_MAX_PEERS);
    NB_CHECK(launch_wait(h, ROW_SORT_DATA, h->p2p_sort_step));
    NB_CUDA(cudaGetLastError());
    h->last_launches += 5;
    return NB_OK;
}

// Everything a step can launch is loaded BEFORE the first peer wait can be in flight.  CUDA 12 loads a
// kernel lazily at its first launch and that load may wait for the whole context to drain; with several
// ranks driven by one host thread (nb_p2p_attach_local) rank 0 would then spin on a flag that rank 1 can
// never raise, because rank 1's first launch is stuck behind rank 0's spinning wait.
int preload_step_kernels(nb_sim* h)
{
    cudaFuncAttributes a;
    NB_CUDA(cudaFuncGetAttributes(&a, reinterpret_cast<const void*>(k_kick_drift_push)));
    NB_CUDA(cudaFuncGetAttributes(&a, reinterpret_cast<const void*>(k_signal)));
    NB_CUDA(cudaFuncGetAttributes(&a, reinterpret_cast<const void*>(k_wait)));
    NB_CUDA(cudaFuncGetAttributes(&a, reinterpret_cast<const void*>(k_publish_count)));
    NB_CUDA(cudaFuncGetAttributes(&a, reinterpret_cast<const void*>(k_push_segment)));
    NB_CHECK(preload_integrate());
    NB_CHECK(preload_tree());
    NB_CHECK(preload_allpairs(h));
    return NB_OK;
}

void p2p_release(nb_sim* h)
{
    if (h->p2p_ipc)
        for (int r = 0; r < h->cfg.world; ++r)
        {
            if (r == h->cfg.rank) continue;
            for (int b = 0; b < 2; ++b)
                if (h->peer_posw[b][r]) cudaIpcCloseMemHandle(h->peer_posw[b][r]);
            if (h->peer_flags[r]) cudaIpcCloseMemHandle(h->peer_flags[r]);
            if (h->peer_acc[r]) cudaIpcCloseMemHandle(h->peer_acc[r]);
            if (h->peer_skeys[r]) cudaIpcCloseMemHandle(h->peer_skeys[r]);
            if (h->peer_svals[r]) cudaIpcCloseMemHandle(h->peer_svals[r]);
        }
    std::memset(h->peer_posw, 0, sizeof(h->peer_posw));
    std::memset(h->peer_flags, 0, sizeof(h->peer_flags));
    std::memset(h->peer_acc, 0, sizeof(h->peer_acc));
    std::memset(h->peer_skeys, 0, sizeof(h->peer_skeys));
    std::memset(h->peer_svals, 0, sizeof(h->peer_svals));
    h->tree.dist_ready = false;
    h->p2p_sort_step = 0;
    h->p2p_attached = false;
    h->p2p_ipc = false;
    h->p2p_step = 0;
    h->p2p_acc_step = 0;
}

}  // namespace nb

using namespace nb;

extern "C" {

int nb_p2p_export(nb_handle h, uint8_t handles[NB_P2P_HANDLE_BYTES])
{
    NB_REQUIRE(h != nullptr && handles != nullptr, NB_ERR_ARG, "null argument");
    NB_REQUIRE(h->n > 0, NB_ERR_STATE, "nb_init_* has not been called");
    NB_REQUIRE(h->cfg.world <= NB_MAX_PEERS, NB_ERR_ARG, "too many ranks for the peer table");
    NB_CUDA(cudaSetDevice(h->cfg.device));
    NB_CHECK(p2p_prepare(h));
    cudaIpcMemHandle_t m[6];
    std::memset(m, 0, sizeof(m));
    NB_CUDA(cudaIpcGetMemHandle(&m[0], h->posw_buf[0]));
    NB_CUDA(cudaIpcGetMemHandle(&m[1], h->posw_buf[1]));
    NB_CUDA(cudaIpcGetMemHandle(&m[2], h->p2p_flags));
    NB_CUDA(cudaIpcGetMemHandle(&m[3], h->acc_base));
    if (h->cfg.mode == NB_MODE_BARNESHUT)
    {
        NB_CUDA(cudaIpcGetMemHandle(&m[4], h->tree.keys_final));
        NB_CUDA(cudaIpcGetMemHandle(&m[5], h->tree.vals_final));
    }
    static_assert(sizeof(m) == NB_P2P_HANDLE_BYTES, "IPC handle size");
    std::memcpy(handles, m, sizeof(m));
    return NB_OK;
}

int nb_p2p_attach(nb_handle h, const uint8_t* all_handles)
{
    NB_REQUIRE(h != nullptr && all_handles != nullptr, NB_ERR_ARG, "null argument");
    NB_REQUIRE(h->n > 0 && h->posw_buf[1] != nullptr, NB_ERR_STATE, "call nb_p2p_export first");
    NB_CUDA(cudaSetDevice(h->cfg.device));
    p2p_release(h);
    for (int r = 0; r < h->cfg.world; ++r)
    {
        if (r == h->cfg.rank)
        {
            h->peer_posw[0][r] = h->posw_buf[0];
            h->peer_posw[1][r] = h->posw_buf[1];
            h->peer_flags[r] = h->p2p_flags;
            h->peer_acc[r] = h->acc_base;
            h->peer_skeys[r] = h->tree.keys_final;
            h->peer_svals[r] = h->tree.vals_final;
            continue;
        }
        cudaIpcMemHandle_t m[6];
        std::memcpy(m, all_handles + (size_t)r * NB_P2P_HANDLE_BYTES, sizeof(m));
        NB_CUDA(cudaIpcOpenMemHandle(&h->peer_posw[0][r], m[0], cudaIpcMemLazyEnablePeerAccess));
        NB_CUDA(cudaIpcOpenMemHandle(&h->peer_posw[1][r], m[1], cudaIpcMemLazyEnablePeerAccess));
        NB_CUDA(cudaIpcOpenMemHandle(&h->peer_flags[r], m[2], cudaIpcMemLazyEnablePeerAccess));
        NB_CUDA(cudaIpcOpenMemHandle(&h->peer_acc[r], m[3], cudaIpcMemLazyEnablePeerAccess));
        if (h->cfg.mode == NB_MODE_BARNESHUT)
        {
            NB_CUDA(cudaIpcOpenMemHandle(&h->peer_skeys[r], m[4], cudaIpcMemLazyEnablePeerAccess));
            NB_CUDA(cudaIpcOpenMemHandle(&h->peer_svals[r], m[5], cudaIpcMemLazyEnablePeerAccess));
        }
    }
    NB_CHECK(preload_step_kernels(h));
    h->p2p_ipc = true;
    h->p2p_attached = true;
    return NB_OK;
}

int nb_p2p_attach_local(nb_handle h, const nb_handle* peers)
{
    NB_REQUIRE(h != nullptr && peers != nullptr, NB_ERR_ARG, "null argument");
    NB_REQUIRE(h->n > 0, NB_ERR_STATE, "nb_init_* has not been called");
    NB_REQUIRE(h->cfg.world <= NB_MAX_PEERS, NB_ERR_ARG, "too many ranks for the peer table");
    NB_CUDA(cudaSetDevice(h->cfg.device));
    p2p_release(h);
    for (int r = 0; r < h->cfg.world; ++r)
    {
        nb_sim* p = peers[r];
        NB_REQUIRE(p != nullptr && p->n == h->n && p->cfg.rank == r && p->cfg.world == h->cfg.world, NB_ERR_ARG,
                   "peer handles must be the ranks 0..world-1 of the same body set");
        NB_CUDA(cudaSetDevice(p->cfg.device));
        NB_CHECK(p->posw_buf[1] == nullptr || p->p2p_flags == nullptr || !p->acc_two ? p2p_prepare(p) : NB_OK);
        if (p->cfg.device != h->cfg.device)
        {
            NB_CUDA(cudaSetDevice(h->cfg.device));
            cudaError_t e = cudaDeviceEnablePeerAccess(p->cfg.device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) NB_CUDA(e);
            cudaGetLastError();
        }
        h->peer_posw[0][r] = p->posw_buf[0];
        h->peer_posw[1][r] = p->posw_buf[1];
        h->peer_flags[r] = p->p2p_flags;
        h->peer_acc[r] = p->acc_base;
        h->peer_skeys[r] = p->tree.keys_final;
        h->peer_svals[r] = p->tree.vals_final;
    }
    NB_CUDA(cudaSetDevice(h->cfg.device));
    NB_CHECK(preload_step_kernels(h));
    h->p2p_ipc = false;
    h->p2p_attached = true;
    return NB_OK;
}

}  // extern "C"
