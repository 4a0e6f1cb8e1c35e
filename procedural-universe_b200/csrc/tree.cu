// Barnes-Hut on the GPU: K3 Morton encode, K4 radix sort, K5 Karras radix tree, K6 bottom-up
// mass / centre-of-mass reduction, K7 warp-cooperative stackless traversal.
//
// What is being replaced: BarnesHut::Update (reference src/Sim/BarnesHut.cpp:44-96) rebuilds a
// pointer octree every step (Octree::Add, Octree.cpp:53-84, eager 8-way Split :16-51), runs
// Octree::CalculateMass (:86-105) and then Octree::CalculateForce (:107-145) per particle.
// The reference's octree semantics that this file reproduces:
//   * fixed root cube [-bounds, bounds)^3, half-open cells, child index z*4 + y*2 + x;
//   * a cell with >= 2 bodies is internal, a single body sits in a leaf, empty children contribute
//     nothing, bodies outside the root are dropped as sources but still receive forces;
//   * acceptance  width / r < theta  tested top-down at EVERY internal cell (also cells that
//     contain the target), leaves are always evaluated directly, the target itself is skipped.
// How: bodies are sorted by the 63-bit Morton code of their level-21 cell; the Karras binary radix
// tree over the sorted codes contains every internal octree cell as the node whose common prefix
// first reaches that cell's level.  A radix-tree node with prefix delta lies in the octree cell of
// level L = floor((delta-1)/3) (the 64-bit key has one leading zero); it OWNS octree cells iff
// L > L(parent).  Owned chains of single-child cells share mass and centre of mass and their widths
// shrink with depth, so "some cell of the chain is accepted" == "the deepest one is":
// accept iff (width_L / theta)^2 < r^2.  Nodes that own no cell are never visited: the traversal
// pointers (next-if-opened / next-if-skipped) jump over them.
//
// Determinism: the sort is a stable LSD radix sort, the tree is a pure function of the sorted
// keys, and the bottom-up pass adds (left + right) in fp64 -- no order-dependent atomics -- so
// every rank that holds the same positions builds bit-identical trees.
#include <atomic>
#include <cfloat>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "nb_internal.h"
#include "allpairs.cuh"   // kPreScale, kEps, rsqrt_approx

namespace nb
{

constexpr int kLevels = 21;
constexpr unsigned long long kOutside = 0xFFFFFFFFFFFFFFFFull;
constexpr int kEnd = -1;
constexpr int kWalkStatWords = 3 + 33 + 8;   // {cells, pairs, visits} + lane-occupancy histogram [0..32] + sparse-visit load per warp: {sum, max} x K = 4, 8, 16 + warps + spare

// counters[] slots
enum { C_INBOUNDS = 0, C_TOTAL = 1, C_SEG = 2, C_SEG_OOB = 3, C_ROOT = 4, C_WORDS = 8 };   // C_SEG: bodies this rank sorts (sharded sort), C_SEG_OOB: of which outside the cube

// ------------------------------------------------------------------------------------------------
// K3: Morton codes.  Same cells as the comparison descent of the CPU restatement (oracle/nbody_port.c,
// port_morton_one): cell corners -B + k * 2B / 2^level are exact in fp64, so the digits are the
// ones Octree::Add's Contains() tests select.  Bodies outside the root get the key ~0.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long spread3(unsigned int v)
{
    unsigned long long x = v & 0x1fffffull;
    x = (x | x << 32) & 0x1f00000000ffffull;
    x = (x | x << 16) & 0x1f0000ff0000ffull;
    x = (x | x << 8) & 0x100f00f00f00f00full;
    x = (x | x << 4) & 0x10c30c30c30c30c3ull;
    x = (x | x << 2) & 0x1249249249249249ull;
    return x;
}

// Cell index of p along one axis at level 21: the k with lo(k) <= p < lo(k+1), lo(k) = -B + k * cell,
// cell = 2B / 2^21.  B is a float, so k * cell (<= 45 significant bits) and lo(k) are exact in fp64 and
// this is the same k the 21-step comparison descent of the restatement (port_morton_one) arrives at:
// estimate by one multiplication, then settle with the exact corner comparisons.
__device__ __forceinline__ unsigned int axis_cell(double p, double B, double cell, double inv_cell)
{
    int k = (int)((p + B) * inv_cell);
    k = max(0, min(k, (1 << kLevels) - 1));
    while (k > 0 && p < __dadd_rn(__dmul_rn((double)k, cell), -B)) --k;
    while (k < (1 << kLevels) - 1 && p >= __dadd_rn(__dmul_rn((double)(k + 1), cell), -B)) ++k;
    return (unsigned int)k;
}

constexpr int MORTON_ITEMS = 4;   // bodies per thread: all loads are issued before the first use

__global__ void __launch_bounds__(256)
k_morton(const float4* __restrict__ posw, int n, double B, double cell, double inv_cell, unsigned long long* __restrict__ keys,
         unsigned int* __restrict__ vals, unsigned int* __restrict__ counters,
         const unsigned long long* __restrict__ splitters, int seg, int nseg, unsigned int* __restrict__ in_segment,
         int oob_first, int oob_count)
{
    // sharded sort: in_segment[i] = 1 iff this rank sorts body i -- an in-bounds body whose key lies in the rank's
    // key range [splitters[seg], splitters[seg+1]) (the last range is open-ended), or an out-of-bounds body
    // (key ~0, they all sort to the end in body order) of the rank's INDEX range [oob_first, oob_first + oob_count):
    // after a dispersal most bodies are outside, and one rank must not end up sorting and shipping all of them.
    unsigned long long seg_lo = 0ull, seg_hi = 0ull;
    if (in_segment != nullptr)
    {
        seg_lo = splitters[seg];
        seg_hi = seg + 1 < nseg ? splitters[seg + 1] : 0ull;
    }
    const int base = blockIdx.x * (256 * MORTON_ITEMS) + threadIdx.x;
    float4 p[MORTON_ITEMS];
#pragma unroll
    for (int k = 0; k < MORTON_ITEMS; ++k)
    {
        const int i = base + k * 256;
        p[k] = i < n ? posw[i] : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    const float b = (float)B;
    int inside_count = 0, oob_mine = 0;
#pragma unroll
    for (int k = 0; k < MORTON_ITEMS; ++k)
    {
        const int i = base + k * 256;
        if (i < n)
        {
            const bool inside = p[k].x >= -b && p[k].y >= -b && p[k].z >= -b && p[k].x < b && p[k].y < b && p[k].z < b;
            unsigned long long key = kOutside;
            if (inside)
            {
                const unsigned int qx = axis_cell((double)p[k].x, B, cell, inv_cell), qy = axis_cell((double)p[k].y, B, cell, inv_cell),
                                   qz = axis_cell((double)p[k].z, B, cell, inv_cell);
                key = spread3(qx) | (spread3(qy) << 1) | (spread3(qz) << 2);
                ++inside_count;
            }
            keys[i] = key;
            vals[i] = (unsigned int)i;
            if (in_segment != nullptr)
            {
                const bool mine = inside ? (key >= seg_lo && (seg + 1 >= nseg || key < seg_hi))
                                         : (i >= oob_first && i < oob_first + oob_count);
                in_segment[i] = mine ? 1u : 0u;
                oob_mine += (mine && !inside) ? 1 : 0;
            }
        }
    }
    // one atomic per block, not per warp: 16 M bodies would otherwise queue 512 K updates on one address
    __shared__ int block_count, block_oob;
    if (threadIdx.x == 0) { block_count = 0; block_oob = 0; }
    __syncthreads();
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
    {
        inside_count += __shfl_down_sync(0xffffffffu, inside_count, o);
        oob_mine += __shfl_down_sync(0xffffffffu, oob_mine, o);
    }
    if ((threadIdx.x & 31) == 0 && inside_count) atomicAdd(&block_count, inside_count);
    if ((threadIdx.x & 31) == 0 && oob_mine) atomicAdd(&block_oob, oob_mine);
    __syncthreads();
    if (threadIdx.x == 0 && block_count) atomicAdd(&counters[C_INBOUNDS], (unsigned int)block_count);
    if (threadIdx.x == 0 && block_oob) atomicAdd(&counters[C_SEG_OOB], (unsigned int)block_oob);
}

// ------------------------------------------------------------------------------------------------
// K4: stable LSD radix sort, 8 bits per pass, (key 64, value 32).  Hand-rolled, no CUB.
// Per pass: block histograms -> exclusive scan in (digit, block) order -> stable scatter.
// A block owns a contiguous tile; warp w owns a contiguous 512-key chunk of it and walks it in
// rounds of 32 consecutive keys, so the order (block, warp, round, lane) is the input order.
// ------------------------------------------------------------------------------------------------
constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_ITEMS = 16;
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;

__global__ void __launch_bounds__(RS_THREADS)
k_rs_hist(const unsigned long long* __restrict__ keys, int n, int shift, unsigned int* __restrict__ hist, int tiles,
          const unsigned int* __restrict__ n_dev)
{
    // n_dev: the key count lives on the device (sharded sort: this rank's segment); tiles past it are idle
    if (n_dev != nullptr) n = (int)*n_dev;
    const int base = blockIdx.x * RS_TILE;
    if (base >= n) return;
    __shared__ unsigned int h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
#pragma unroll
    for (int r = 0; r < RS_ITEMS; ++r)
    {
        const int i = base + r * RS_THREADS + threadIdx.x;
        if (i < n) atomicAdd(&h[(unsigned int)(keys[i] >> shift) & 255u], 1u);
    }
    __syncthreads();
    hist[(size_t)threadIdx.x * tiles + blockIdx.x] = h[threadIdx.x];
}

// one block per digit: in-place exclusive scan of the digit's row, total to totals[digit]
__global__ void __launch_bounds__(256)
k_rs_scan_rows(unsigned int* __restrict__ hist, int tiles, unsigned int* __restrict__ totals,
               const unsigned int* __restrict__ n_dev = nullptr)
{
    __shared__ unsigned int warp_sums[8];
    __shared__ unsigned int carry;
    unsigned int* row = hist + (size_t)blockIdx.x * tiles;
    if (n_dev != nullptr) tiles = min(tiles, (int)((*n_dev + RS_TILE - 1) / RS_TILE));   // the row stride stays the full tile count
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < tiles; base += 256)
    {
        const int i = base + threadIdx.x;
        const unsigned int v = i < tiles ? row[i] : 0u;
        unsigned int x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1)
        {
            const unsigned int y = __shfl_up_sync(0xffffffffu, x, o);
            if ((threadIdx.x & 31) >= o) x += y;
        }
        if ((threadIdx.x & 31) == 31) warp_sums[threadIdx.x >> 5] = x;
        __syncthreads();
        unsigned int wprefix = 0;
        for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) wprefix += warp_sums[w];
        const unsigned int c = carry;
        if (i < tiles) row[i] = c + wprefix + x - v;
        __syncthreads();
        if (threadIdx.x == 255) carry = c + wprefix + x;
        __syncthreads();
    }
    if (threadIdx.x == 0) totals[blockIdx.x] = carry;
}

__global__ void __launch_bounds__(256) k_rs_scan_totals(unsigned int* __restrict__ totals)
{
    __shared__ unsigned int s[256];
    s[threadIdx.x] = totals[threadIdx.x];
    __syncthreads();
    if (threadIdx.x == 0)
    {
        unsigned int run = 0;
        for (int d = 0; d < 256; ++d) { const unsigned int v = s[d]; s[d] = run; run += v; }
    }
    __syncthreads();
    totals[threadIdx.x] = s[threadIdx.x];
}

// Onesweep-style sort (Adinets & Merrill 2022), NB_SORT=onesweep: ONE histogram pass counts all eight digits of every key,
// and the scatter of pass p finds "keys of my digit in the tiles before mine" by decoupled look-back over per-tile
// descriptors (2 flag bits + 30-bit count) instead of a histogram kernel and two scan kernels per pass.  Tiles are
// handed out by an atomic ticket, so a tile only ever waits for tiles whose blocks are already running.
// Measured at 16 M bodies: it removes 8 x 61 us of histogram + scan launches and adds 80 us of one-off histogram plus
// ~50 us of look-back per pass (one thread per digit walks back through the ~440 resident tiles, eight entries per
// round trip) -- 1.75 ms either way, so the three-kernel passes stay the default; a warp-parallel look-back is what
// it would take to turn the saved launches into time.
constexpr int RS_PASSES = 8;
constexpr unsigned int RS_FLAG_LOCAL = 1u << 30, RS_FLAG_PREFIX = 2u << 30, RS_VALUE_MASK = (1u << 30) - 1u;

__global__ void __launch_bounds__(256)
k_rs_hist_all(const unsigned long long* __restrict__ keys, int n, const unsigned int* __restrict__ n_dev, unsigned int* __restrict__ ghist)
{
    if (n_dev != nullptr) n = (int)*n_dev;
    __shared__ unsigned int h[RS_PASSES][256];
    for (int k = threadIdx.x; k < RS_PASSES * 256; k += 256) (&h[0][0])[k] = 0;
    __syncthreads();
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < (size_t)n; i += (size_t)gridDim.x * 256)
    {
        const unsigned long long key = keys[i];
#pragma unroll
        for (int p = 0; p < RS_PASSES; ++p) atomicAdd(&h[p][(unsigned int)(key >> (8 * p)) & 255u], 1u);
    }
    __syncthreads();
    for (int k = threadIdx.x; k < RS_PASSES * 256; k += 256)
    {
        const unsigned int v = (&h[0][0])[k];
        if (v) atomicAdd(&ghist[k], v);
    }
}

// one block: ghist[p][d] -> exclusive prefix over d, for every pass p
__global__ void __launch_bounds__(256) k_rs_scan_all(unsigned int* __restrict__ ghist)
{
    __shared__ unsigned int s[256];
    for (int p = 0; p < RS_PASSES; ++p)
    {
        s[threadIdx.x] = ghist[p * 256 + threadIdx.x];
        __syncthreads();
        if (threadIdx.x == 0)
        {
            unsigned int run = 0;
            for (int d = 0; d < 256; ++d) { const unsigned int v = s[d]; s[d] = run; run += v; }
        }
        __syncthreads();
        ghist[p * 256 + threadIdx.x] = s[threadIdx.x];
        __syncthreads();
    }
}

// Scatter of one tile.  (1) every key gets its rank among the keys of the same digit that precede
// it in the tile (warp-level match.any per round + running per-warp counters; the order is warp,
// round, lane = input order, so the sort is stable); (2) the tile is reordered by digit in shared
// memory; (3) thread t writes tile slot t + 256 k, so keys of one digit leave as contiguous runs
// (mean run 16 keys = 128 B for uniform digits, far longer in the high passes) instead of 8-byte
// scattered stores.
constexpr size_t RS_SCATTER_SMEM = RS_TILE * (sizeof(unsigned long long) + sizeof(unsigned int) + sizeof(unsigned short)) +
                                   (RS_WARPS + 2) * 256 * sizeof(unsigned int);

template <bool ONESWEEP>
__global__ void __launch_bounds__(RS_THREADS, 3)
k_rs_scatter(const unsigned long long* __restrict__ keys_in, const unsigned int* __restrict__ vals_in,
             unsigned long long* __restrict__ keys_out, unsigned int* __restrict__ vals_out, int n, int shift,
             const unsigned int* __restrict__ hist, const unsigned int* __restrict__ totals, int tiles,
             const unsigned int* __restrict__ n_dev, unsigned int* desc = nullptr, unsigned int* ticket = nullptr)
{
    // desc != nullptr: onesweep -- `totals` holds this pass's exclusive digit offsets, the tile index comes from the
    // ticket, and the per-tile prefix from look-back over desc[tile][digit]
    if (n_dev != nullptr) n = (int)*n_dev;
    __shared__ int tile_id;
    if (ONESWEEP)
    {
        if (threadIdx.x == 0) tile_id = (int)atomicAdd(ticket, 1u);
        __syncthreads();
    }
    const int tile = ONESWEEP ? tile_id : (int)blockIdx.x;
    if ((long long)tile * RS_TILE >= n) return;
    extern __shared__ __align__(16) unsigned char rs_smem[];
    unsigned long long* skeys = reinterpret_cast<unsigned long long*>(rs_smem);            // tile, input order
    unsigned int* svals = reinterpret_cast<unsigned int*>(skeys + RS_TILE);
    unsigned int (*wh)[256] = reinterpret_cast<unsigned int (*)[256]>(svals + RS_TILE);   // [RS_WARPS][256]
    unsigned int* lbase = &wh[RS_WARPS][0];      // first digit-ordered slot of each digit
    unsigned int* gbase = lbase + 256;           // global position of digit-ordered slot s of digit d = gbase[d] + s
    unsigned short* perm = reinterpret_cast<unsigned short*>(gbase + 256);                 // digit-ordered slot -> input slot
    __shared__ unsigned int warp_tot[RS_WARPS];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int k = threadIdx.x; k < RS_WARPS * 256; k += RS_THREADS) (&wh[0][0])[k] = 0;

    const int tile_base = tile * RS_TILE;
    const int wbase = warp * (RS_ITEMS * 32);    // warp w owns input slots [512 w, 512 w + 512)
#pragma unroll
    for (int r = 0; r < RS_ITEMS; ++r)
    {
        const int ip = wbase + r * 32 + lane;
        const bool ok = tile_base + ip < n;
        skeys[ip] = ok ? keys_in[tile_base + ip] : 0ull;
        svals[ip] = ok ? vals_in[tile_base + ip] : 0u;
    }
    __syncthreads();

    const unsigned int lt = (1u << lane) - 1u;
    unsigned int packed[RS_ITEMS];               // digit << 16 | rank among the warp's earlier keys of that digit
#pragma unroll
    for (int r = 0; r < RS_ITEMS; ++r)
    {
        const int ip = wbase + r * 32 + lane;
        const bool ok = tile_base + ip < n;
        // invalid lanes get a private pseudo-digit so they never match anyone
        const unsigned int d = ok ? ((unsigned int)(skeys[ip] >> shift) & 255u) : (256u + lane);
        const unsigned int peers = __match_any_sync(0xffffffffu, d);
        const unsigned int before = ok ? wh[warp][d] : 0u;
        __syncwarp();
        packed[r] = (d << 16) | (before + __popc(peers & lt));
        if (ok && lane == __ffs(peers) - 1) wh[warp][d] = before + __popc(peers);
        __syncwarp();
    }
    __syncthreads();
    {
        // thread d: exclusive scan over the warps, then over the digits (block scan of the totals)
        const int d = threadIdx.x;   // RS_THREADS == 256 digits
        unsigned int run = 0;
#pragma unroll
        for (int w = 0; w < RS_WARPS; ++w) { const unsigned int c = wh[w][d]; wh[w][d] = run; run += c; }
        unsigned int x = run;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1)
        {
            const unsigned int y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) warp_tot[warp] = x;
        __syncthreads();
        unsigned int wprefix = 0;
        for (int w = 0; w < warp; ++w) wprefix += warp_tot[w];
        const unsigned int first = wprefix + x - run;
        lbase[d] = first;
        unsigned int before;                      // keys of digit d in the tiles before this one
        if (ONESWEEP)
        {
            volatile unsigned int* dd = desc + d;
            dd[(size_t)tile * 256] = RS_FLAG_LOCAL | run;
            before = 0;
            // look back eight tiles per round trip: the loads of a batch are independent; the batch is then consumed
            // nearest tile first, re-reading an entry that was not published yet
            for (int t = tile - 1; t >= 0;)
            {
                constexpr int LB = 8;
                unsigned int v[LB];
#pragma unroll
                for (int k = 0; k < LB; ++k) v[k] = t - k >= 0 ? dd[(size_t)(t - k) * 256] : (2u << 30);
                bool done = false;
#pragma unroll
                for (int k = 0; k < LB; ++k)
                {
                    if (done) continue;
                    while ((v[k] >> 30) == 0u) v[k] = dd[(size_t)(t - k) * 256];
                    before += v[k] & RS_VALUE_MASK;
                    done = (v[k] >> 30) == 2u;
                }
                if (done) break;
                t -= LB;
            }
            dd[(size_t)tile * 256] = RS_FLAG_PREFIX | (before + run);
        }
        else before = hist[(size_t)d * tiles + tile];
        gbase[d] = totals[d] + before - first;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < RS_ITEMS; ++r)
    {
        const int ip = wbase + r * 32 + lane;
        if (tile_base + ip < n)
        {
            const unsigned int d = packed[r] >> 16;
            perm[lbase[d] + wh[warp][d] + (packed[r] & 0xffffu)] = (unsigned short)ip;
        }
    }
    __syncthreads();
    const int live = min(RS_TILE, n - tile_base);
#pragma unroll
    for (int k = 0; k < RS_ITEMS; ++k)
    {
        const int slot = k * RS_THREADS + threadIdx.x;
        if (slot < live)
        {
            const int ip = perm[slot];
            const unsigned long long kk = skeys[ip];
            const unsigned int pos = gbase[(unsigned int)(kk >> shift) & 255u] + slot;
            keys_out[pos] = kk;
            vals_out[pos] = svals[ip];
        }
    }
}

// Small scenes (n <= RS_TILE, the reference's interactive range -- its default is 1 000 bodies): all eight passes in
// ONE block.  The tile lives in shared memory, ping-ponging between two buffers; a pass ranks exactly as
// k_rs_scatter does, and for a single tile the digit-ordered slot IS the output position.  Replaces 32 launches.
constexpr size_t RS_SMALL_SMEM = 2 * RS_TILE * (sizeof(unsigned long long) + sizeof(unsigned int)) + RS_TILE * sizeof(unsigned short) +
                                 (RS_WARPS + 1) * 256 * sizeof(unsigned int);

__global__ void __launch_bounds__(RS_THREADS)
k_rs_small(const unsigned long long* __restrict__ keys_in, const unsigned int* __restrict__ vals_in,
           unsigned long long* __restrict__ keys_out, unsigned int* __restrict__ vals_out, int n)
{
    extern __shared__ __align__(16) unsigned char rs_smem[];
    unsigned long long* kbuf0 = reinterpret_cast<unsigned long long*>(rs_smem);
    unsigned long long* kbuf1 = kbuf0 + RS_TILE;
    unsigned int* vbuf0 = reinterpret_cast<unsigned int*>(kbuf1 + RS_TILE);
    unsigned int* vbuf1 = vbuf0 + RS_TILE;
    unsigned int (*wh)[256] = reinterpret_cast<unsigned int (*)[256]>(vbuf1 + RS_TILE);   // [RS_WARPS][256]
    unsigned int* lbase = &wh[RS_WARPS][0];
    unsigned short* perm = reinterpret_cast<unsigned short*>(lbase + 256);
    __shared__ unsigned int warp_tot[RS_WARPS];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wbase = warp * (RS_ITEMS * 32);
    const unsigned int lt = (1u << lane) - 1u;
#pragma unroll
    for (int r = 0; r < RS_ITEMS; ++r)
    {
        const int ip = wbase + r * 32 + lane;
        kbuf0[ip] = ip < n ? keys_in[ip] : 0ull;
        vbuf0[ip] = ip < n ? vals_in[ip] : 0u;
    }
    for (int pass = 0; pass < 8; ++pass)
    {
        const unsigned long long* skeys = (pass & 1) ? kbuf1 : kbuf0;
        const unsigned int* svals = (pass & 1) ? vbuf1 : vbuf0;
        unsigned long long* okeys = (pass & 1) ? kbuf0 : kbuf1;
        unsigned int* ovals = (pass & 1) ? vbuf0 : vbuf1;
        const int shift = 8 * pass;
        for (int k = threadIdx.x; k < RS_WARPS * 256; k += RS_THREADS) (&wh[0][0])[k] = 0;
        __syncthreads();
        unsigned int packed[RS_ITEMS];               // digit << 16 | rank among the warp's earlier keys of that digit
#pragma unroll
        for (int r = 0; r < RS_ITEMS; ++r)
        {
            const int ip = wbase + r * 32 + lane;
            const bool ok = ip < n;
            const unsigned int d = ok ? ((unsigned int)(skeys[ip] >> shift) & 255u) : (256u + lane);
            const unsigned int peers = __match_any_sync(0xffffffffu, d);
            const unsigned int before = ok ? wh[warp][d] : 0u;
            __syncwarp();
            packed[r] = (d << 16) | (before + __popc(peers & lt));
            if (ok && lane == __ffs(peers) - 1) wh[warp][d] = before + __popc(peers);
            __syncwarp();
        }
        __syncthreads();
        {
            const int d = threadIdx.x;
            unsigned int run = 0;
#pragma unroll
            for (int w = 0; w < RS_WARPS; ++w) { const unsigned int c = wh[w][d]; wh[w][d] = run; run += c; }
            unsigned int x = run;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1)
            {
                const unsigned int y = __shfl_up_sync(0xffffffffu, x, o);
                if (lane >= o) x += y;
            }
            if (lane == 31) warp_tot[warp] = x;
            __syncthreads();
            unsigned int wprefix = 0;
            for (int w = 0; w < warp; ++w) wprefix += warp_tot[w];
            lbase[d] = wprefix + x - run;
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < RS_ITEMS; ++r)
        {
            const int ip = wbase + r * 32 + lane;
            if (ip < n)
            {
                const unsigned int d = packed[r] >> 16;
                perm[lbase[d] + wh[warp][d] + (packed[r] & 0xffffu)] = (unsigned short)ip;
            }
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < RS_ITEMS; ++k)
        {
            const int slot = k * RS_THREADS + threadIdx.x;
            if (slot < n)
            {
                const int ip = perm[slot];
                okeys[slot] = skeys[ip];
                ovals[slot] = svals[ip];
            }
        }
        __syncthreads();
    }
    for (int i = threadIdx.x; i < n; i += RS_THREADS)
    {
        keys_out[i] = kbuf0[i];                       // eight passes: the result is back in the first buffer
        vals_out[i] = vbuf0[i];
    }
}

// ------------------------------------------------------------------------------------------------
// K5: Karras radix tree (HPG 2012, section 3) over the m sorted in-bounds keys.
// Node ids: internal node i -> i, leaf slot j -> leaf_base + j.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int delta_fn(const unsigned long long* __restrict__ k, int m, int i, int j)
{
    if (j < 0 || j >= m) return -1;
    const unsigned long long x = k[i] ^ k[j];
    if (x != 0ull) return __clzll((long long)x);
    return 64 + __clz(i ^ j);
}

__device__ __forceinline__ int level_of(int prefix_bits)
{
    return prefix_bits >= 64 ? kLevels : min(kLevels, (prefix_bits - 1) / 3);
}

// meta[id] of an internal node = octree level of its cell | 0x100 if it OWNS octree cells (its level is
// deeper than its parent's).  Written by the parent's thread, which knows both children's ranges and
// therefore their prefixes; the same thread counts owning nodes per first slot (pre-order ranks).
constexpr unsigned short kOwns = 0x100;
constexpr unsigned short kIsLeft = 0x200;     // the node is its parent's LEFT child (same first slot as the parent)

__global__ void __launch_bounds__(256)
k_karras(const unsigned long long* __restrict__ keys, const unsigned int* __restrict__ counters, int leaf_base,
         int2* __restrict__ child, int* __restrict__ prefix, int* __restrict__ parent,
         unsigned int* __restrict__ flags, int* __restrict__ first_slot, unsigned short* __restrict__ meta,
         unsigned int* __restrict__ cnt, int* __restrict__ last_slot)
{
    const int m = (int)counters[C_INBOUNDS];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m - 1) return;
    const int d = (delta_fn(keys, m, i, i + 1) - delta_fn(keys, m, i, i - 1)) >= 0 ? 1 : -1;
    const int dmin = delta_fn(keys, m, i, i - d);
    int lmax = 2;
    while (delta_fn(keys, m, i, i + lmax * d) > dmin) lmax *= 2;
    int l = 0;
    for (int t = lmax / 2; t >= 1; t /= 2)
        if (delta_fn(keys, m, i, i + (l + t) * d) > dmin) l += t;
    const int j = i + l * d;
    const int dnode = delta_fn(keys, m, i, j);
    int s = 0, t = l;
    do
    {
        t = (t + 1) / 2;
        if (delta_fn(keys, m, i, i + (s + t) * d) > dnode) s += t;
    } while (t > 1);
    const int gamma = i + s * d + (d < 0 ? d : 0);
    const int lo = min(i, j), hi = max(i, j);
    const int cl = (lo == gamma) ? leaf_base + gamma : gamma;
    const int cr = (hi == gamma + 1) ? leaf_base + gamma + 1 : gamma + 1;
    child[i] = make_int2(cl, cr);
    prefix[i] = dnode;
    first_slot[i] = lo;
    last_slot[i] = hi;
    parent[cl] = i;
    parent[cr] = i;
    flags[i] = 0;
    const int level = level_of(dnode);
    if (i == 0)
    {
        parent[0] = kEnd;
        meta[0] = (unsigned short)level | kOwns;
        atomicAdd(&cnt[0], 1u);
    }
    if (cl < leaf_base)
    {
        const int lc = level_of(delta_fn(keys, m, lo, gamma));
        meta[cl] = (unsigned short)lc | (lc > level ? kOwns : (unsigned short)0) | kIsLeft;
        if (lc > level) atomicAdd(&cnt[lo], 1u);
    }
    if (cr < leaf_base)
    {
        const int lc = level_of(delta_fn(keys, m, gamma + 1, hi));
        meta[cr] = (unsigned short)lc | (lc > level ? kOwns : (unsigned short)0);
        if (lc > level) atomicAdd(&cnt[gamma + 1], 1u);
    }
}

// ------------------------------------------------------------------------------------------------
// K6: bottom-up weight / weighted-position sums in fp64 (Octree::CalculateMass, Octree.cpp:86-105,
// which accumulates the centre of mass in fp32 and overflows for large total mass -- fp64 here).
// One thread per leaf climbs; the second arrival at a node combines (left + right).
// ------------------------------------------------------------------------------------------------
// Node sums are 32-byte records {w, w x, w y, w z} (one sector per node).  A climbing thread carries
// the sums of the subtree it has just completed in registers, so at every level it reads only the
// SIBLING's record; the child / parent links are fetched alongside the arrival atomic.
__global__ void __launch_bounds__(256)
k_bottom_up(const float4* __restrict__ posw, const unsigned int* __restrict__ order,
            const unsigned int* __restrict__ counters, int leaf_base, const int2* __restrict__ child,
            const int* __restrict__ parent, unsigned int* __restrict__ flags, double* nsum)
{
    const int m = (int)counters[C_INBOUNDS];
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m || m < 2) return;
    int id = leaf_base + j;
    int node = parent[id];
    double w, sx, sy, sz;
    {
        const float4 p = posw[order[j]];
        w = (double)p.w;
        sx = w * (double)p.x; sy = w * (double)p.y; sz = w * (double)p.z;
    }
    for (;;)
    {
        const int2 c = child[node];
        const int up = parent[node];
        // Arrival counter.  Release (level >= 2) publishes the record this thread stored one level below;
        // a leaf has stored nothing, so level 1 is relaxed.  No acquire half (it costs an L1 invalidate per
        // level): the sibling's record is read with ld.global.cg -- L2, where the release made it visible
        // before the counter moved -- and that load cannot issue before the branch on `arrived` resolves.
        unsigned int arrived;
        if (id >= leaf_base)
            asm volatile("atom.add.relaxed.gpu.global.u32 %0, [%1], %2;" : "=r"(arrived) : "l"(flags + node), "r"(1u) : "memory");
        else
            asm volatile("atom.add.release.gpu.global.u32 %0, [%1], %2;" : "=r"(arrived) : "l"(flags + node), "r"(1u) : "memory");
        if (arrived == 0u) return;                       // first arrival: the sibling subtree is not done
        const int sib = (c.x == id) ? c.y : c.x;
        double ow, ox, oy, oz;
        if (sib >= leaf_base)
        {
            const float4 p = posw[order[sib - leaf_base]];
            ow = (double)p.w;
            ox = ow * (double)p.x; oy = ow * (double)p.y; oz = ow * (double)p.z;
        }
        else
        {
            // written by another SM earlier in this kernel: read through L2, never a stale L1 line
            const double2* rec = reinterpret_cast<const double2*>(nsum + 4 * (size_t)sib);
            const double2 a = __ldcg(rec), b = __ldcg(rec + 1);
            ow = a.x; ox = a.y; oy = b.x; oz = b.y;
        }
        w += ow; sx += ox; sy += oy; sz += oz;           // fp64 addition commutes: left + right either way
        double2* rec = reinterpret_cast<double2*>(nsum + 4 * (size_t)node);
        rec[0] = make_double2(w, sx);
        rec[1] = make_double2(sy, sz);
        if (up == kEnd) return;
        id = node;
        node = up;
    }
}

// K6, production form: the same reduction with BLOCK-LOCAL arrival counters.  A block owns 128 consecutive leaves;
// a node whose whole range lies inside them (all but the few that straddle a block boundary: a node is numbered by
// an end of its range, so its id lies in the block too) is reduced through shared memory -- arrival counter and
// the sums a sibling has to read -- instead of a global atomic and a 32-byte record fetched back from L2.  Nodes
// that straddle blocks use the global protocol of k_bottom_up.  Every node's sums still go to the global record
// (k_finalize and the parents outside the block read them); the additions are the same fp64 (left + right), so the
// result is bitwise the one k_bottom_up computes.  Measured at 16 M bodies (whole build): 4.19 ms with k_bottom_up,
// 4.53 / 4.10 / 4.09 ms with blocks of 1024 / 256 / 128 leaves -- big blocks stay resident until their last climber
// is done, and what bounds the climb is the dependent child / parent loads of every level, which both forms share.
constexpr int BU_LEAVES = 128;

__global__ void __launch_bounds__(BU_LEAVES)
k_bottom_up_local(const float4* __restrict__ posw, const unsigned int* __restrict__ order,
                  const unsigned int* __restrict__ counters, int leaf_base, const int2* __restrict__ child,
                  const int* __restrict__ parent, const int* __restrict__ first_slot, const int* __restrict__ last_slot,
                  unsigned int* __restrict__ flags, double* nsum)
{
    __shared__ unsigned int sflags[BU_LEAVES];
    __shared__ double ssum[4][BU_LEAVES];
    const int m = (int)counters[C_INBOUNDS];
    const int b0 = blockIdx.x * BU_LEAVES, b1 = b0 + BU_LEAVES;
    if (b0 >= m || m < 2) return;
    sflags[threadIdx.x] = 0u;
    __syncthreads();
    const int j = b0 + threadIdx.x;
    if (j >= m) return;
    int id = leaf_base + j;
    int node = parent[id];
    double w, sx, sy, sz;
    {
        const float4 p = posw[order[j]];
        w = (double)p.w;
        sx = w * (double)p.x; sy = w * (double)p.y; sz = w * (double)p.z;
    }
    for (;;)
    {
        const int2 c = child[node];
        const int up = parent[node];
        const bool local = first_slot[node] >= b0 && last_slot[node] < b1;
        const int sib = (c.x == id) ? c.y : c.x;
        double ow, ox, oy, oz;
        if (local)
        {
            if (id < leaf_base)
            {
                // the sums of the range just completed, for the sibling's thread if it arrives second
                const int k = id - b0;
                ssum[0][k] = w; ssum[1][k] = sx; ssum[2][k] = sy; ssum[3][k] = sz;
                __threadfence_block();
            }
            if (atomicAdd(&sflags[node - b0], 1u) == 0u) return;
            __threadfence_block();
            if (sib >= leaf_base)
            {
                const float4 p = posw[order[sib - leaf_base]];
                ow = (double)p.w;
                ox = ow * (double)p.x; oy = ow * (double)p.y; oz = ow * (double)p.z;
            }
            else
            {
                const volatile double* vs = &ssum[0][0];
                const int k = sib - b0;
                ow = vs[k]; ox = vs[BU_LEAVES + k]; oy = vs[2 * BU_LEAVES + k]; oz = vs[3 * BU_LEAVES + k];
            }
        }
        else
        {
            unsigned int arrived;
            if (id >= leaf_base)
                asm volatile("atom.add.relaxed.gpu.global.u32 %0, [%1], %2;" : "=r"(arrived) : "l"(flags + node), "r"(1u) : "memory");
            else
                asm volatile("atom.add.release.gpu.global.u32 %0, [%1], %2;" : "=r"(arrived) : "l"(flags + node), "r"(1u) : "memory");
            if (arrived == 0u) return;
            if (sib >= leaf_base)
            {
                const float4 p = posw[order[sib - leaf_base]];
                ow = (double)p.w;
                ox = ow * (double)p.x; oy = ow * (double)p.y; oz = ow * (double)p.z;
            }
            else
            {
                const double2* rec = reinterpret_cast<const double2*>(nsum + 4 * (size_t)sib);
                const double2 a = __ldcg(rec), b = __ldcg(rec + 1);
                ow = a.x; ox = a.y; oy = b.x; oz = b.y;
            }
        }
        w += ow; sx += ox; sy += oy; sz += oz;           // fp64 addition commutes: left + right either way
        double2* rec = reinterpret_cast<double2*>(nsum + 4 * (size_t)node);
        rec[0] = make_double2(w, sx);
        rec[1] = make_double2(sy, sz);
        if (up == kEnd) return;
        id = node;
        node = up;
    }
}

// ------------------------------------------------------------------------------------------------
// K5 + K6 fused (NB_BUILD=agglomerative; measured slower than the two passes, kept with its tests): agglomerative construction (Apetrei, "Fast and Simple Agglomerative LBVH Construction",
// 2014) -- the radix tree is found WHILE its sums are reduced, bottom-up, instead of by two binary searches per
// node beforehand.  A thread starts at a leaf with the range [j, j]; a finished range [l, r] joins the neighbour
// it shares the longer prefix with: the node that splits at r (range becomes its LEFT child) or at l - 1 (its
// RIGHT child).  The two children of a node meet through one 64-bit exchange word {outer bound, node id}: the
// first to arrive leaves its half and retires, the second takes it, adds the sibling's sums to the ones it
// carries in registers, writes the node (children, parent links, prefix, first slot, the children's octree level
// and "owns a cell" bit, the owning-node counts -- everything k_karras wrote) and climbs on.  Internal node ids are
// SPLIT POSITIONS here (k_karras numbers a node by an end of its range); it is the same tree, and nb_get_tree
// renumbers for the topology parity hook.  The root's id is left in counters[C_ROOT].
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_build_up(const float4* __restrict__ posw, const unsigned int* __restrict__ order, const unsigned long long* __restrict__ keys,
           unsigned int* __restrict__ counters, int leaf_base, int2* __restrict__ child, int* __restrict__ prefix,
           int* __restrict__ parent, int* __restrict__ first_slot, unsigned short* __restrict__ meta, unsigned int* __restrict__ cnt,
           unsigned long long* meet, double* nsum)
{
    const int m = (int)counters[C_INBOUNDS];
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m || m < 2) return;
    int l = j, r = j, id = leaf_base + j;
    double w, sx, sy, sz;
    {
        const float4 p = posw[order[j]];
        w = (double)p.w;
        sx = w * (double)p.x; sy = w * (double)p.y; sz = w * (double)p.z;
    }
    for (;;)
    {
        if (l == 0 && r == m - 1)
        {
            parent[id] = kEnd;
            meta[id] = (unsigned short)level_of(delta_fn(keys, m, l, r)) | kOwns;
            atomicAdd(&cnt[0], 1u);
            counters[C_ROOT] = (unsigned int)id;
            return;
        }
        // longer common prefix = closer; at the ends of the key list there is only one neighbour
        const int dl = l > 0 ? delta_fn(keys, m, l - 1, l) : -1;
        const int dr = r < m - 1 ? delta_fn(keys, m, r, r + 1) : -1;
        const bool to_right = dr > dl;                   // the node splitting at r: this range is its left child
        const int node = to_right ? r : l - 1;
        const unsigned long long mine = ((unsigned long long)(unsigned int)(to_right ? l : r) << 32) | (unsigned int)id;
        // Release (internal ranges) publishes the sums record this thread stored when it formed `id`; a leaf has stored
        // nothing.  No acquire half: the sibling's record is read with ld.global.cg after the branch on `other`
        // (same reasoning as k_bottom_up).
        unsigned long long other;
        if (id >= leaf_base)
            asm volatile("atom.relaxed.gpu.global.exch.b64 %0, [%1], %2;" : "=l"(other) : "l"(meet + node), "l"(mine) : "memory");
        else
            asm volatile("atom.release.gpu.global.exch.b64 %0, [%1], %2;" : "=l"(other) : "l"(meet + node), "l"(mine) : "memory");
        if (other == ~0ull) return;                       // first arrival: the sibling range is not finished
        const int sib = (int)(unsigned int)(other & 0xffffffffull);
        const int bound = (int)(unsigned int)(other >> 32);
        double ow, ox, oy, oz;
        if (sib >= leaf_base)
        {
            const float4 p = posw[order[sib - leaf_base]];
            ow = (double)p.w;
            ox = ow * (double)p.x; oy = ow * (double)p.y; oz = ow * (double)p.z;
        }
        else
        {
            const double2* rec = reinterpret_cast<const double2*>(nsum + 4 * (size_t)sib);
            const double2 a = __ldcg(rec), b = __ldcg(rec + 1);
            ow = a.x; ox = a.y; oy = b.x; oz = b.y;
        }
        w += ow; sx += ox; sy += oy; sz += oz;           // fp64 addition commutes: left + right either way
        double2* rec = reinterpret_cast<double2*>(nsum + 4 * (size_t)node);
        rec[0] = make_double2(w, sx);
        rec[1] = make_double2(sy, sz);
        const int cl = to_right ? id : sib, cr = to_right ? sib : id;
        if (to_right) r = bound; else l = bound;
        const int dnode = delta_fn(keys, m, l, r);
        const int level = level_of(dnode);
        child[node] = make_int2(cl, cr);
        prefix[node] = dnode;
        first_slot[node] = l;
        parent[cl] = node;
        parent[cr] = node;
        if (cl < leaf_base)
        {
            const int lc = level_of(delta_fn(keys, m, l, node));
            meta[cl] = (unsigned short)lc | (lc > level ? kOwns : (unsigned short)0) | kIsLeft;
            if (lc > level) atomicAdd(&cnt[l], 1u);
        }
        if (cr < leaf_base)
        {
            const int lc = level_of(delta_fn(keys, m, node + 1, r));
            meta[cr] = (unsigned short)lc | (lc > level ? kOwns : (unsigned short)0);
            if (lc > level) atomicAdd(&cnt[node + 1], 1u);
        }
        id = node;
    }
}

// ------------------------------------------------------------------------------------------------
// K6b: traversal records.  a = {com.xyz, G M 2^-27}, b = {open threshold, next-if-opened,
// next-if-skipped, body}.  Threshold: leaves -1 (always evaluated), owning nodes (width/theta)^2,
// nodes that own no octree cell are skipped by the pointers.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool owns_cell(int id, const unsigned short* __restrict__ meta) { return (meta[id] & kOwns) != 0; }

// K6b: depth-first (pre-order) layout of the nodes the traversal can visit -- leaves and the
// radix-tree nodes that own an octree cell.  Opening a node then means "next record" and skipping
// a subtree is a forward jump, so a walk reads memory front to back (L1-friendly).
// rank(v) = #records before v in pre-order:
//   every leaf slot < first(v) precedes v; an owning node u precedes v iff first(u) < first(v), or
//   first(u) == first(v) and u is a proper ancestor of v (same first <=> chain of left children).
// With cnt[f] = #owning nodes whose range starts at slot f and P = exclusive scan of cnt:
//   rank(node v) = first(v) + P[first(v)] + #owning proper ancestors with the same first
//   rank(leaf j) = j + P[j + 1]
__global__ void __launch_bounds__(256)
k_scan_apply(const unsigned int* __restrict__ in, int n, const unsigned int* __restrict__ block_prefix,
             unsigned int* __restrict__ out)
{
    __shared__ unsigned int warp_sums[8];
    __shared__ unsigned int carry;
    if (threadIdx.x == 0) carry = block_prefix[blockIdx.x];
    __syncthreads();
    const int base = blockIdx.x * 4096;
    for (int k = 0; k < 16; ++k)
    {
        const int i = base + k * 256 + threadIdx.x;
        const unsigned int v = i < n ? in[i] : 0u;
        unsigned int x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1)
        {
            const unsigned int y = __shfl_up_sync(0xffffffffu, x, o);
            if ((threadIdx.x & 31) >= o) x += y;
        }
        if ((threadIdx.x & 31) == 31) warp_sums[threadIdx.x >> 5] = x;
        __syncthreads();
        unsigned int wprefix = 0;
        for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) wprefix += warp_sums[w];
        const unsigned int c = carry;
        if (i < n) out[i] = c + wprefix + x - v;
        __syncthreads();
        if (threadIdx.x == 255) carry = c + wprefix + x;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256)
k_rank(unsigned int* __restrict__ counters, int n, int leaf_base, const int2* __restrict__ child,
       const unsigned short* __restrict__ meta, const int* __restrict__ parent, const int* __restrict__ first_slot,
       const unsigned int* __restrict__ pref, int* __restrict__ rank)
{
    const int m = (int)counters[C_INBOUNDS];
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t == 0) counters[C_TOTAL] = (unsigned int)m + pref[n];
    if (t < m) rank[leaf_base + t] = t + (int)pref[t + 1];
    if (t < m - 1 && owns_cell(t, meta))
    {
        // climb while the node is a left child: one dependent pair of loads per level (parent link, then the parent's
        // meta, which carries both "owns a cell" and whether the climb goes on)
        int above = 0, v = t;
        unsigned short mv = meta[t];
        while (mv & kIsLeft)
        {
            v = parent[v];
            mv = meta[v];
            above += (mv & kOwns) ? 1 : 0;
        }
        const int f = first_slot[t];
        rank[t] = f + (int)pref[f] + above;
    }
}

// The record after the subtree of `id` in pre-order (kEnd if none).
__device__ __forceinline__ int after_subtree(int id, int leaf_base, const int2* __restrict__ child,
                                             const unsigned short* __restrict__ meta, const int* __restrict__ parent)
{
    for (;;)
    {
        const int p = parent[id];
        if (p == kEnd) return kEnd;
        const int2 c = child[p];
        if (c.x == id)
        {
            int r = c.y;
            while (r < leaf_base && !owns_cell(r, meta)) r = child[r].x;
            return r;
        }
        id = p;
    }
}

// Traversal records, 32 bytes each, at their pre-order rank: {com.xyz, G M 2^-27} and
// {open threshold, rank of the record after this subtree, body, 0}.  Threshold: leaves -1 (always
// evaluated), owning nodes (width / theta)^2.
__global__ void __launch_bounds__(256)
k_finalize(const float4* __restrict__ posw, const unsigned int* __restrict__ order, const unsigned int* __restrict__ counters,
           int leaf_base, const int2* __restrict__ child,
           const unsigned short* __restrict__ meta, const int* __restrict__ parent, const double* __restrict__ nsum,
           float root_width, float inv_theta,
           const int* __restrict__ rank, float4* __restrict__ nodes)
{
    const int m = (int)counters[C_INBOUNDS];
    const int total = (int)counters[C_TOTAL];
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t == 0)
    {
        // sentinel after the last record: weightless, accepted by everybody, "skips" to itself -- a walk that checks
        // for the end only every other visit (k_walk2) may land on it once
        nodes[2 * (size_t)total] = make_float4(0.f, 0.f, 0.f, 0.f);
        nodes[2 * (size_t)total + 1] = make_float4(-1.0f, __int_as_float(total), __int_as_float(-1), 0.f);
    }
    if (t < m)
    {
        const int r = rank[leaf_base + t];
        const unsigned int body = order[t];
        const float4 p = posw[body];
        nodes[2 * (size_t)r] = make_float4(p.x, p.y, p.z, p.w * kPreScale);
        nodes[2 * (size_t)r + 1] = make_float4(-1.0f, __int_as_float(r + 1), __int_as_float((int)body), 0.f);
    }
    if (t < m - 1 && owns_cell(t, meta))
    {
        const int r = rank[t];
        const int nxt = after_subtree(t, leaf_base, child, meta, parent);
        const int skip = nxt == kEnd ? total : rank[nxt];
        const double2* rec = reinterpret_cast<const double2*>(nsum + 4 * (size_t)t);
        const double2 s0 = rec[0], s1 = rec[1];
        const double w = s0.x;
        const double inv = w != 0.0 ? 1.0 / w : 0.0;
        const float width = ldexpf(root_width, -(int)(meta[t] & 0xff));
        const float lim = width * inv_theta;
        nodes[2 * (size_t)r] = make_float4((float)(s0.y * inv), (float)(s1.x * inv), (float)(s1.y * inv), (float)w * kPreScale);
        nodes[2 * (size_t)r + 1] = make_float4(lim * lim, __int_as_float(skip), __int_as_float(-1), 0.f);
    }
}

// ------------------------------------------------------------------------------------------------
// K7: warp-cooperative stackless traversal.  A warp owns 32 consecutive targets of the Morton-
// sorted list and walks the union of what its lanes need; node records are warp-uniform loads.
// Every lane applies its OWN acceptance test (the reference's per-particle decisions, not a
// group criterion): a lane that accepts a cell the warp still has to open for another lane parks
// until the walk has left that subtree (cur >= the record after it).
// Interaction: same law as all-pairs, a += G M (c - p) / (|d| (d^2 + S)); the self term and
// coincident bodies vanish through the epsilon (see allpairs.cuh).
// ------------------------------------------------------------------------------------------------
// BALANCED (world > 1 with peer memory attached): the Morton-ordered list of ALL bodies is dealt out
// block by block -- this rank takes blocks rank, rank + world, ... -- so every rank walks the same mix
// of dense and sparse regions, and each lane stores its acceleration straight into the acc array of the
// body's owner (local or over NVLink); the owners' kick-drift waits for the accelerations (p2p.cu).
template <bool STATS, int GROUP, bool BALANCED, bool SKIP_UNUSED = false>
__global__ void __launch_bounds__(256)
k_walk(const float4* __restrict__ posw, const unsigned int* __restrict__ order, const unsigned int* __restrict__ tlist,
       int ntargets, const unsigned int* __restrict__ counters, const float4* __restrict__ nodes,
       int first, int count, float sc, double* __restrict__ acc, unsigned long long* __restrict__ stats, AccTable owners)
{
    // GROUP consecutive lanes share one traversal pointer (32 = the whole warp, the default).  Smaller
    // groups walk a smaller union of subtrees, but the warp runs until its slowest group is done and the
    // vote costs more: measured at 16 M bodies 37.6 ms (32) / 42.0 ms (16) / 40.9 ms (8).
    const int t = BALANCED ? (blockIdx.x * owners.world + owners.rank) * 256 + threadIdx.x
                           : blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = t < ntargets;
    unsigned int body = 0;
    float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
    if (valid)
    {
        body = order[(!BALANCED && tlist) ? tlist[t] : (unsigned int)t];
        p = posw[body];
    }
    const int total = (int)counters[C_TOTAL];
    const unsigned int gmask = GROUP == 32 ? 0xffffffffu : ((GROUP == 16 ? 0xffffu : 0xffu) << (threadIdx.x & (32 - GROUP) & 31));
    int cur = 0;
    int parked = valid ? 0 : 0x7fffffff;   // the lane is idle while cur < parked; lanes without a target never wake
    float ax = 0.f, ay = 0.f, az = 0.f;
    unsigned int n_cells = 0, n_leaves = 0, n_visits = 0;
    unsigned int n_sparse4 = 0, n_sparse8 = 0, n_sparse16 = 0;   // STATS: this lane's visits with <= 4 / 8 / 16 lanes of the warp awake
    for (;;)
    {
        const bool live = cur < total;
        if (GROUP == 32) { if (!live) break; }
        else if (!__any_sync(0xffffffffu, live)) break;
        const size_t rec = 2 * (size_t)(live ? cur : 0);
        const float4 a = nodes[rec];
        const float2 b = *reinterpret_cast<const float2*>(nodes + rec + 1);
        const bool active = live && cur >= parked;
        const float dx = a.x - p.x, dy = a.y - p.y, dz = a.z - p.z;
        float d2 = dx * dx;
        d2 = fmaf(dy, dy, d2);
        d2 = fmaf(dz, dz, d2);
        const float thr = b.x;
        const int skip = __float_as_int(b.y);
        const bool accept = d2 > thr;
        const bool use = active && accept;
        // SKIP_UNUSED (kernel_variant 3, kept for the record): a record that no lane of the warp evaluates costs the
        // traversal instructions only -- but the extra vote and branch on EVERY visit cost more than the skipped
        // interactions save: 44.4 ms against 39.1 ms at 16 M bodies
        if (!SKIP_UNUSED || __any_sync(0xffffffffu, use))
        {
            const float tt = fmaf(d2, kPreScale, sc);
            const float u = d2 * tt;
            const float x = fmaf(u, tt, kEps);
            const float s = use ? a.w * rsqrt_approx(x) : 0.f;
            ax = fmaf(s, dx, ax);
            ay = fmaf(s, dy, ay);
            az = fmaf(s, dz, az);
        }
        if (use) parked = skip;
        if (STATS)
        {
            // lane-occupancy histogram of the walk: stats[3 + k] counts warp iterations with k lanes at work
            const unsigned int busy = __ballot_sync(0xffffffffu, active);
            if ((threadIdx.x & 31) == 0) atomicAdd(&stats[3 + __popc(busy)], 1ull);
            if (active)
            {
                const int awake = __popc(busy);
                n_sparse4 += awake <= 4; n_sparse8 += awake <= 8; n_sparse16 += awake <= 16;
            }
        }
        if (STATS && active)
        {
            ++n_visits;
            if (accept)
            {
                if (thr < 0.f) n_leaves += ((unsigned int)__float_as_int(nodes[rec + 1].z) != body);
                else ++n_cells;
            }
        }
        bool open;
        if (GROUP == 32) open = __any_sync(0xffffffffu, active && !accept);
        else open = (__ballot_sync(0xffffffffu, active && !accept) & gmask) != 0u;
        if (live) cur = open ? cur + 1 : skip;
    }
    if (valid && BALANCED)
    {
        // owner of `body` under the block partition first[r] = r * n / world
        int o = (int)(((unsigned long long)body * (unsigned long long)owners.world) / (unsigned long long)owners.first[owners.world]);
        while (o > 0 && (int)body < owners.first[o]) --o;
        while (o + 1 < owners.world && (int)body >= owners.first[o + 1]) ++o;
        const size_t cnt = (size_t)(owners.first[o + 1] - owners.first[o]);
        const size_t li = (size_t)((int)body - owners.first[o]);
        double* dst = owners.acc[o] + (size_t)owners.parity * 3 * cnt;
        dst[li] = (double)ax;
        dst[cnt + li] = (double)ay;
        dst[2 * cnt + li] = (double)az;
    }
    else if (valid)
    {
        const int li = (int)body - first;
        if (li >= 0 && li < count)
        {
            acc[li] = (double)ax;
            acc[(size_t)count + li] = (double)ay;
            acc[2 * (size_t)count + li] = (double)az;
        }
    }
    if (STATS)
    {
        unsigned long long c = n_cells, l = n_leaves, v = n_visits;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
        {
            c += __shfl_down_sync(0xffffffffu, c, o);
            l += __shfl_down_sync(0xffffffffu, l, o);
            v += __shfl_down_sync(0xffffffffu, v, o);
        }
        if ((threadIdx.x & 31) == 0) { atomicAdd(&stats[0], c); atomicAdd(&stats[1], l); atomicAdd(&stats[2], v); }
        // how evenly the sparse visits are spread over the lanes of a warp: sum and max per warp (a per-lane treatment
        // of those visits would run as long as the busiest lane)
        const unsigned int sp[3] = {n_sparse4, n_sparse8, n_sparse16};
#pragma unroll
        for (int k = 0; k < 3; ++k)
        {
            const unsigned int sum = __reduce_add_sync(0xffffffffu, sp[k]), mx = __reduce_max_sync(0xffffffffu, sp[k]);
            if ((threadIdx.x & 31) == 0) { atomicAdd(&stats[36 + 2 * k], (unsigned long long)sum); atomicAdd(&stats[37 + 2 * k], (unsigned long long)mx); }
        }
        if ((threadIdx.x & 31) == 0) atomicAdd(&stats[42], 1ull);
    }
}

// K7, production form.  Same traversal, same per-lane decisions and the same per-target summation order as k_walk
// (the results are bitwise equal), trimmed from 27 to 25 SASS instructions per visit:
//   * the end of the record list is tested every OTHER visit: a sentinel record sits at index `total`
//     (weightless, accepted by every lane, skipping to itself), so the one extra visit is harmless;
//   * two visits per loop trip also remove the register move of the loop-carried traversal pointer;
//   * the weight is masked (FSEL) instead of the whole interaction being predicated, which stops ptxas from
//     re-reading the softening constant under a predicate.
// LOAD256 fetches the 32-byte record with ONE 256-bit load (LDG.E.256, new on sm_100): 24 instructions per visit, but
// measured slower (39.4 ms against 38.5 ms at 16 M bodies).  Measured at 16 M bodies on one B200: 39.2 ms (k_walk),
// 38.5 ms (this kernel); also tried and rejected: skipping the interaction when no lane uses the record (44.4 ms), and
// requesting record cur + 1 ahead of the decision (57.9 ms) -- the loop is bound by instruction issue, not by latency.
// TPW (targets per warp) < 32 serves small scenes: with a few thousand bodies the GPU is nearly empty and a warp's
// walk is a chain of dependent record loads, so fewer targets per warp = a smaller union = a shorter chain, on more
// warps.  A lane's result does not depend on which targets share its warp, so every TPW gives bitwise the same forces.
template <bool BALANCED, bool LOAD256 = false, int TPW = 32>
__global__ void __launch_bounds__(256)
k_walk2(const float4* __restrict__ posw, const unsigned int* __restrict__ order, const unsigned int* __restrict__ tlist,
        int ntargets, const unsigned int* __restrict__ counters, const float4* __restrict__ nodes,
        int first, int count, float sc, double* __restrict__ acc, AccTable owners)
{
    const int gt = blockIdx.x * blockDim.x + threadIdx.x;
    const int t = BALANCED ? (blockIdx.x * owners.world + owners.rank) * 256 + threadIdx.x
                           : (TPW == 32 ? gt : (gt >> 5) * TPW + (gt & 31));
    const bool valid = t < ntargets && (TPW == 32 || (gt & 31) < TPW);
    unsigned int body = 0;
    float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
    if (valid)
    {
        body = order[(!BALANCED && tlist) ? tlist[t] : (unsigned int)t];
        p = posw[body];
    }
    const int total = (int)counters[C_TOTAL];
    int cur = 0;
    int parked = valid ? 0 : 0x7fffffff;
    float ax = 0.f, ay = 0.f, az = 0.f;
#define NB_VISIT(CUR, NEXT)                                                                                              \
    {                                                                                                                    \
        float ax_, ay_, az_, aw_, thr_, skipf_, b2_, b3_;                                                                \
        if (LOAD256)                                                                                                     \
            asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"                                          \
                         : "=f"(ax_), "=f"(ay_), "=f"(az_), "=f"(aw_), "=f"(thr_), "=f"(skipf_), "=f"(b2_), "=f"(b3_)    \
                         : "l"(nodes + 2 * (size_t)(CUR)));                                                              \
        else                                                                                                             \
        {                                                                                                                \
            const float4 ra_ = nodes[2 * (size_t)(CUR)];                                                                 \
            const float2 rb_ = *reinterpret_cast<const float2*>(nodes + 2 * (size_t)(CUR) + 1);                          \
            ax_ = ra_.x; ay_ = ra_.y; az_ = ra_.z; aw_ = ra_.w; thr_ = rb_.x; skipf_ = rb_.y;                            \
        }                                                                                                                \
        const bool active = (CUR) >= parked;                                                                             \
        const float dx = ax_ - p.x, dy = ay_ - p.y, dz = az_ - p.z;                                                      \
        float d2 = dx * dx;                                                                                              \
        d2 = fmaf(dy, dy, d2);                                                                                           \
        d2 = fmaf(dz, dz, d2);                                                                                           \
        const int skip = __float_as_int(skipf_);                                                                         \
        const bool accept = d2 > thr_;                                                                                   \
        const bool use = active && accept;                                                                               \
        const float tt = fmaf(d2, kPreScale, sc);                                                                        \
        const float u = d2 * tt;                                                                                         \
        const float x = fmaf(u, tt, kEps);                                                                               \
        const float s = (use ? aw_ : 0.f) * rsqrt_approx(x);                                                             \
        ax = fmaf(s, dx, ax);                                                                                            \
        ay = fmaf(s, dy, ay);                                                                                            \
        az = fmaf(s, dz, az);                                                                                            \
        if (use) parked = skip;                                                                                          \
        const bool open = __any_sync(0xffffffffu, active && !accept);                                                    \
        NEXT = open ? (CUR) + 1 : skip;                                                                                  \
    }
    if (total > 0)
    {
        do
        {
            int mid;
            NB_VISIT(cur, mid)
            NB_VISIT(mid, cur)
        } while (cur < total);
    }
#undef NB_VISIT
    if (valid && BALANCED)
    {
        int o = (int)(((unsigned long long)body * (unsigned long long)owners.world) / (unsigned long long)owners.first[owners.world]);
        while (o > 0 && (int)body < owners.first[o]) --o;
        while (o + 1 < owners.world && (int)body >= owners.first[o + 1]) ++o;
        const size_t cnt = (size_t)(owners.first[o + 1] - owners.first[o]);
        const size_t li = (size_t)((int)body - owners.first[o]);
        double* dst = owners.acc[o] + (size_t)owners.parity * 3 * cnt;
        dst[li] = (double)ax;
        dst[cnt + li] = (double)ay;
        dst[2 * cnt + li] = (double)az;
    }
    else if (valid)
    {
        const int li = (int)body - first;
        if (li >= 0 && li < count)
        {
            acc[li] = (double)ax;
            acc[(size_t)count + li] = (double)ay;
            acc[2 * (size_t)count + li] = (double)az;
        }
    }
}

// Debug export: what BarnesHut::RenderDebug -> Octree::RenderDebug draws (Octree.cpp:147-175) -- one
// cube per occupied leaf.  A body's leaf sits one level below the deepest cell it shares with a
// neighbour in the sorted order (level 0, the root cube, if it is alone; level 21 for duplicates).
__device__ __forceinline__ unsigned int compact3(unsigned long long x)
{
    x &= 0x1249249249249249ull;
    x = (x | x >> 2) & 0x10c30c30c30c30c3ull;
    x = (x | x >> 4) & 0x100f00f00f00f00full;
    x = (x | x >> 8) & 0x1f0000ff0000ffull;
    x = (x | x >> 16) & 0x1f00000000ffffull;
    x = (x | x >> 32) & 0x1fffffull;
    return (unsigned int)x;
}

__global__ void __launch_bounds__(256)
k_leaf_cells(const unsigned long long* __restrict__ keys, const unsigned int* __restrict__ counters, double B,
             float4* __restrict__ cells)
{
    const int m = (int)counters[C_INBOUNDS];
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    const int shared = max(delta_fn(keys, m, j, j - 1), delta_fn(keys, m, j, j + 1));
    const int depth = shared < 0 ? 0 : min(kLevels, level_of(shared) + 1);
    const unsigned long long k = keys[j];
    const int drop = kLevels - depth;                               // axis bits below the leaf's level
    const double width = ldexp(2.0 * B, -depth);
    const double x = -B + (double)(compact3(k) >> drop) * width;
    const double y = -B + (double)(compact3(k >> 1) >> drop) * width;
    const double z = -B + (double)(compact3(k >> 2) >> drop) * width;
    const double half = 0.5 * width;
    cells[j] = make_float4((float)(x + half), (float)(y + half), (float)(z + half), (float)width);
}

// Owned targets in Morton order (world > 1): slots whose body index lies in [first, first+count).
__global__ void __launch_bounds__(256)
k_select_flags(const unsigned int* __restrict__ order, int n, int first, int count, unsigned int* __restrict__ flag)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < n)
    {
        const int b = (int)order[s] - first;
        flag[s] = (b >= 0 && b < count) ? 1u : 0u;
    }
}

// Exclusive scan of flag[] -> positions, compaction.  One block per 4096 slots + row scan reuse.
__global__ void __launch_bounds__(256)
k_block_sums(const unsigned int* __restrict__ flag, int n, unsigned int* __restrict__ sums)
{
    __shared__ unsigned int s[256];
    unsigned int c = 0;
    const int base = blockIdx.x * 4096;
    for (int k = 0; k < 16; ++k)
    {
        const int i = base + k * 256 + threadIdx.x;
        if (i < n) c += flag[i];
    }
    s[threadIdx.x] = c;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1)
    {
        if (threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) sums[blockIdx.x] = s[0];
}

__global__ void __launch_bounds__(256)
k_compact(const unsigned int* __restrict__ flag, int n, const unsigned int* __restrict__ block_prefix,
          unsigned int* __restrict__ tlist)
{
    __shared__ unsigned int warp_sums[8];
    __shared__ unsigned int carry;
    if (threadIdx.x == 0) carry = block_prefix[blockIdx.x];
    __syncthreads();
    const int base = blockIdx.x * 4096;
    for (int k = 0; k < 16; ++k)
    {
        const int i = base + k * 256 + threadIdx.x;
        const unsigned int v = i < n ? flag[i] : 0u;
        unsigned int x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1)
        {
            const unsigned int y = __shfl_up_sync(0xffffffffu, x, o);
            if ((threadIdx.x & 31) >= o) x += y;
        }
        if ((threadIdx.x & 31) == 31) warp_sums[threadIdx.x >> 5] = x;
        __syncthreads();
        unsigned int wprefix = 0;
        for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) wprefix += warp_sums[w];
        const unsigned int c = carry;
        if (v) tlist[c + wprefix + x - 1] = (unsigned int)i;
        __syncthreads();
        if (threadIdx.x == 255) carry = c + wprefix + x;
        __syncthreads();
    }
}

// Sharded sort, step 1: this rank's bodies (in_segment == 1) compacted in body order -- stable, so equal
// keys keep their body order through the local sort exactly as in the full sort.  The last block leaves
// the segment size in counters[C_SEG].
__global__ void __launch_bounds__(256)
k_compact_pairs(const unsigned int* __restrict__ flag, int n, const unsigned int* __restrict__ block_prefix,
                const unsigned long long* __restrict__ keys_in, const unsigned int* __restrict__ vals_in,
                unsigned long long* __restrict__ keys_out, unsigned int* __restrict__ vals_out, unsigned int* __restrict__ counters)
{
    __shared__ unsigned int warp_sums[8];
    __shared__ unsigned int carry;
    if (threadIdx.x == 0) carry = block_prefix[blockIdx.x];
    __syncthreads();
    const int base = blockIdx.x * 4096;
    for (int k = 0; k < 16; ++k)
    {
        const int i = base + k * 256 + threadIdx.x;
        const unsigned int v = i < n ? flag[i] : 0u;
        unsigned int x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1)
        {
            const unsigned int y = __shfl_up_sync(0xffffffffu, x, o);
            if ((threadIdx.x & 31) >= o) x += y;
        }
        if ((threadIdx.x & 31) == 31) warp_sums[threadIdx.x >> 5] = x;
        __syncthreads();
        unsigned int wprefix = 0;
        for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) wprefix += warp_sums[w];
        const unsigned int c = carry;
        if (v)
        {
            const unsigned int at = c + wprefix + x - 1;
            keys_out[at] = keys_in[i];
            vals_out[at] = vals_in[i];
        }
        __syncthreads();
        if (threadIdx.x == 255) carry = c + wprefix + x;
        __syncthreads();
    }
    if (threadIdx.x == 0 && blockIdx.x == gridDim.x - 1) counters[C_SEG] = carry;
}

// Key ranges of the NEXT sharded sort: equal shares of this step's in-bounds bodies.  Every rank holds the
// same sorted keys, so every rank derives the same splitters.
__global__ void k_splitters(const unsigned long long* __restrict__ skeys, const unsigned int* __restrict__ counters, int world,
                            unsigned long long* __restrict__ splitters)
{
    const int r = threadIdx.x;
    if (r > world) return;
    const unsigned int m = counters[C_INBOUNDS];
    splitters[r] = (r == 0 || m == 0) ? 0ull : (r == world ? kOutside : skeys[(size_t)r * m / world]);
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static inline int blocks_for(size_t n, int threads) { return (int)((n + threads - 1) / threads); }

void tree_release(nb_sim* h)
{
    TreeBuffers& t = h->tree;
    for (int k = 0; k < 2; ++k) { cudaFree(t.keys[k]); cudaFree(t.vals[k]); t.keys[k] = nullptr; t.vals[k] = nullptr; }
    cudaFree(t.hist); cudaFree(t.desc); cudaFree(t.counters); cudaFree(t.child); cudaFree(t.parent); cudaFree(t.prefix);
    cudaFree(t.range); cudaFree(t.range_hi); cudaFree(t.flags); cudaFree(t.nsum); cudaFree(t.walk_a);
    cudaFree(t.walk_b); cudaFree(t.stats); cudaFree(t.cnt); cudaFree(t.pref); cudaFree(t.rank); cudaFree(t.tlist); cudaFree(t.meta);
    cudaFree(t.keys_final); cudaFree(t.vals_final); cudaFree(t.splitters);
    t = TreeBuffers();
}

int tree_reserve(nb_sim* h)
{
    TreeBuffers& t = h->tree;
    const size_t n = h->n;
    // Opt in to > 48 KB of dynamic shared memory once per device, here (Init time) and not in the step:
    // the call can wait for the device to drain, which must never happen while another handle of this
    // process sits in a peer-flag wait (several ranks driven by one host thread).
    static std::atomic<bool> opted_in[64];
    const int dev = h->cfg.device;
    if (dev < 0 || dev >= 64 || !opted_in[dev].load(std::memory_order_acquire))
    {
        NB_CUDA(cudaFuncSetAttribute(k_rs_scatter<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RS_SCATTER_SMEM));
        NB_CUDA(cudaFuncSetAttribute(k_rs_scatter<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RS_SCATTER_SMEM));
        NB_CUDA(cudaFuncSetAttribute(k_rs_small, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RS_SMALL_SMEM));
        if (dev >= 0 && dev < 64) opted_in[dev].store(true, std::memory_order_release);    // idempotent: a second caller repeats it at worst
    }
    if (t.capacity >= n) return NB_OK;
    tree_release(h);
    const size_t tiles = (n + RS_TILE - 1) / RS_TILE;
    for (int k = 0; k < 2; ++k)
    {
        NB_CUDA(cudaMalloc(&t.keys[k], n * sizeof(unsigned long long)));
        NB_CUDA(cudaMalloc(&t.vals[k], n * sizeof(unsigned int)));
    }
    t.hist_words = 256 * tiles + 256 + tiles + 16;
    // onesweep: per-pass tile descriptors [8][tiles][256], then the 8 x 256 digit counts, then 8 tickets
    t.desc_words = (size_t)RS_PASSES * tiles * 256 + RS_PASSES * 256 + 16;
    NB_CUDA(cudaMalloc(&t.desc, t.desc_words * sizeof(unsigned int)));
    NB_CUDA(cudaMalloc(&t.hist, t.hist_words * sizeof(unsigned int)));
    NB_CUDA(cudaMalloc(&t.counters, C_WORDS * sizeof(unsigned int)));
    NB_CUDA(cudaMalloc(&t.child, n * sizeof(int2)));
    NB_CUDA(cudaMalloc(&t.parent, 2 * n * sizeof(int)));
    NB_CUDA(cudaMalloc(&t.prefix, n * sizeof(int)));
    NB_CUDA(cudaMalloc(&t.range, n * sizeof(int)));                   // first slot of every node's range
    NB_CUDA(cudaMalloc(&t.range_hi, n * sizeof(int)));                // last slot
    NB_CUDA(cudaMalloc(&t.cnt, (n + 1) * sizeof(unsigned int)));
    NB_CUDA(cudaMalloc(&t.pref, (n + 1) * sizeof(unsigned int)));
    NB_CUDA(cudaMalloc(&t.rank, 2 * n * sizeof(int)));
    NB_CUDA(cudaMalloc(&t.meta, n * sizeof(unsigned short)));
    NB_CUDA(cudaMalloc(&t.tlist, n * sizeof(unsigned int)));
    NB_CUDA(cudaMalloc(&t.flags, n * sizeof(unsigned int)));
    NB_CUDA(cudaMalloc(&t.nsum, 4 * n * sizeof(double)));
    NB_CUDA(cudaMalloc(&t.walk_a, 4 * n * sizeof(float4)));           // 2n records x 32 B
    NB_CUDA(cudaMalloc(&t.stats, kWalkStatWords * sizeof(unsigned long long)));
    NB_CUDA(cudaMemsetAsync(t.stats, 0, kWalkStatWords * sizeof(unsigned long long), h->stream));
    t.capacity = n;
    return NB_OK;
}

int tree_build(nb_sim* h, bool collective)
{
    NB_CHECK(tree_reserve(h));
    TreeBuffers& t = h->tree;
    const int n = (int)h->n;
    const int tiles = (n + RS_TILE - 1) / RS_TILE;
    cudaStream_t st = h->stream;
    // collective: called from nb_step with peer memory attached, i.e. every rank is building right now.
    // From the second such build on the sort is sharded by Morton-key range (splitters from the last one).
    const bool peers = collective && h->p2p_attached && h->cfg.world > 1 && t.keys_final != nullptr;
    const bool sharded = peers && t.dist_ready;

    NB_CUDA(cudaMemsetAsync(t.counters, 0, C_WORDS * sizeof(unsigned int), st));
    const double cell = std::ldexp((double)h->cfg.bounds, 1 - kLevels);      // 2B / 2^21, exact
    k_morton<<<blocks_for(n, 256 * MORTON_ITEMS), 256, 0, st>>>(h->posw, n, (double)h->cfg.bounds, cell, 1.0 / cell, t.keys[0], t.vals[0],
                                                               t.counters, t.splitters, h->cfg.rank, h->cfg.world,
                                                               sharded ? t.flags : nullptr, (int)h->first, (int)h->count);
    ++h->last_launches;

    unsigned int* totals = t.hist + (size_t)256 * tiles;
    int src = 0;
    const unsigned int* n_dev = nullptr;
    if (sharded)
    {
        // this rank's key range, compacted in body order into the second buffer: flags -> block sums -> scan -> pairs
        const int blocks = (n + 4095) / 4096;
        unsigned int* sums = t.hist;                  // scratch until the first histogram pass
        k_block_sums<<<blocks, 256, 0, st>>>(t.flags, n, sums);
        k_rs_scan_rows<<<1, 256, 0, st>>>(sums, blocks, sums + blocks);
        k_compact_pairs<<<blocks, 256, 0, st>>>(t.flags, n, sums, t.keys[0], t.vals[0], t.keys[1], t.vals[1], t.counters);
        h->last_launches += 3;
        src = 1;
        n_dev = t.counters + C_SEG;
    }
    if (!sharded && n <= RS_TILE)
    {
        k_rs_small<<<1, RS_THREADS, RS_SMALL_SMEM, st>>>(t.keys[0], t.vals[0], t.keys[1], t.vals[1], n);
        ++h->last_launches;
        src = 1;
    }
    else
    {
    static const bool three_kernel_sort = [] { const char* v = std::getenv("NB_SORT"); return v == nullptr || std::strcmp(v, "onesweep") != 0; }();
    if (three_kernel_sort || (size_t)n >= ((size_t)1 << 30))
    {
        // per pass a histogram kernel, two scan kernels and the scatter
        for (int pass = 0; pass < 8; ++pass)
        {
            const int shift = 8 * pass;
            k_rs_hist<<<tiles, RS_THREADS, 0, st>>>(t.keys[src], n, shift, t.hist, tiles, n_dev);
            k_rs_scan_rows<<<256, 256, 0, st>>>(t.hist, tiles, totals, n_dev);
            k_rs_scan_totals<<<1, 256, 0, st>>>(totals);
            k_rs_scatter<false><<<tiles, RS_THREADS, RS_SCATTER_SMEM, st>>>(t.keys[src], t.vals[src], t.keys[src ^ 1], t.vals[src ^ 1], n, shift,
                                                                           t.hist, totals, tiles, n_dev);
            h->last_launches += 4;
            src ^= 1;
        }
    }
    else
    {
        unsigned int* ghist = t.desc + (size_t)RS_PASSES * tiles * 256;
        unsigned int* tickets = ghist + RS_PASSES * 256;
        NB_CUDA(cudaMemsetAsync(t.desc, 0, t.desc_words * sizeof(unsigned int), st));
        k_rs_hist_all<<<h->sm_count * 4, 256, 0, st>>>(t.keys[src], n, n_dev, ghist);
        k_rs_scan_all<<<1, 256, 0, st>>>(ghist);
        h->last_launches += 2;
        for (int pass = 0; pass < RS_PASSES; ++pass)
        {
            k_rs_scatter<true><<<tiles, RS_THREADS, RS_SCATTER_SMEM, st>>>(t.keys[src], t.vals[src], t.keys[src ^ 1], t.vals[src ^ 1], n, 8 * pass,
                                                                          nullptr, ghist + pass * 256, tiles, n_dev,
                                                                          t.desc + (size_t)pass * tiles * 256, tickets + pass);
            ++h->last_launches;
            src ^= 1;
        }
    }
    }
    t.cur = src;
    NB_CUDA(cudaGetLastError());
    if (sharded)
    {
        // every rank's sorted segment -> every rank's final arrays, at its offset (p2p.cu)
        NB_CHECK(p2p_sort_exchange(h, t.keys[src], t.vals[src], t.counters + C_SEG));   // {total, out-of-bounds} = C_SEG, C_SEG_OOB
        t.skeys = t.keys_final;
        t.svals = t.vals_final;
    }
    else
    {
        t.skeys = t.keys[src];
        t.svals = t.vals[src];
    }
    if (peers)
    {
        k_splitters<<<1, 32, 0, st>>>(t.skeys, t.counters, h->cfg.world, t.splitters);
        ++h->last_launches;
        t.dist_ready = true;
    }

    const int leaf_base = n;
    const int words = n + 1;
    NB_CUDA(cudaMemsetAsync(t.cnt, 0, (size_t)words * sizeof(unsigned int), st));
    // NB_BUILD=agglomerative selects the fused construction + reduction (k_build_up).  Measured at 16 M bodies it is
    // SLOWER than the two passes (2.07 ms against 0.53 + 0.98 ms): every level of its climb waits for the neighbour
    // keys of a range it has only just learnt, on top of the exchange and the sibling record, where k_karras's binary
    // searches run at full occupancy and k_bottom_up's climb carries only three dependent accesses per level.
    static const bool two_pass = [] { const char* v = std::getenv("NB_BUILD"); return v == nullptr || std::strcmp(v, "agglomerative") != 0; }();
    t.split_ids = !two_pass;
    if (two_pass)
    {
        // Karras's two binary searches per node, then the reduction
        k_karras<<<blocks_for(n, 256), 256, 0, st>>>(t.skeys, t.counters, leaf_base, t.child, t.prefix, t.parent, t.flags, t.range,
                                                    t.meta, t.cnt, t.range_hi);
        static const bool global_reduce = [] { const char* v = std::getenv("NB_REDUCE"); return v != nullptr && std::strcmp(v, "global") == 0; }();
        if (global_reduce)
            k_bottom_up<<<blocks_for(n, 256), 256, 0, st>>>(h->posw, t.svals, t.counters, leaf_base, t.child, t.parent, t.flags, t.nsum);
        else
            k_bottom_up_local<<<blocks_for(n, BU_LEAVES), BU_LEAVES, 0, st>>>(h->posw, t.svals, t.counters, leaf_base, t.child, t.parent, t.range,
                                                                             t.range_hi, t.flags, t.nsum);
    }
    else
    {
        // the sort's spare ping-pong half is free now: the children's meeting words live there
        unsigned long long* meet = t.keys[t.cur ^ 1];
        NB_CUDA(cudaMemsetAsync(meet, 0xFF, (size_t)n * sizeof(unsigned long long), st));
        k_build_up<<<blocks_for(n, 256), 256, 0, st>>>(h->posw, t.svals, t.skeys, t.counters, leaf_base, t.child, t.prefix, t.parent, t.range,
                                                      t.meta, t.cnt, meet, t.nsum);
    }
    // pre-order ranks: owning nodes per first slot were counted by k_karras; exclusive scan, rank, then the records
    {
        const int sblocks = (words + 4095) / 4096;
        unsigned int* sums = t.hist;                 // radix-sort scratch is free again
        unsigned int* total = t.hist + sblocks;
        k_block_sums<<<sblocks, 256, 0, st>>>(t.cnt, words, sums);
        k_rs_scan_rows<<<1, 256, 0, st>>>(sums, sblocks, total);
        k_scan_apply<<<sblocks, 256, 0, st>>>(t.cnt, words, sums, t.pref);
        k_rank<<<blocks_for(n, 256), 256, 0, st>>>(t.counters, n, leaf_base, t.child, t.meta, t.parent, t.range, t.pref, t.rank);
        h->last_launches += 4;
    }
    k_finalize<<<blocks_for(n, 256), 256, 0, st>>>(h->posw, t.svals, t.counters, leaf_base, t.child, t.meta,
                                                  t.parent, t.nsum, 2.0f * h->cfg.bounds, 1.0f / h->cfg.theta, t.rank, t.walk_a);
    h->last_launches += 3;
    NB_CUDA(cudaGetLastError());
    t.built = true;
    return NB_OK;
}

int tree_walk(nb_sim* h, bool balanced, bool g_walk_stats)
{
    TreeBuffers& t = h->tree;
    const int n = (int)h->n;
    cudaStream_t st = h->stream;
    const unsigned int* order = t.svals;
    const float sc = (float)(h->cfg.softening * (double)kPreScale);
    AccTable owners;
    std::memset(&owners, 0, sizeof(owners));
    if (balanced && !g_walk_stats)
    {
        NB_CHECK(p2p_acc_table(h, &owners));
        const int all_blocks = blocks_for(n, 256);
        const int blocks = (all_blocks - h->cfg.rank + h->cfg.world - 1) / h->cfg.world;
        if (blocks > 0 && h->cfg.kernel_variant == 4)
            k_walk<false, 32, true><<<blocks, 256, 0, st>>>(h->posw, order, nullptr, n, t.counters, t.walk_a, (int)h->first,
                                                           (int)h->count, sc, h->acc, t.stats, owners);
        else if (blocks > 0)
            k_walk2<true><<<blocks, 256, 0, st>>>(h->posw, order, nullptr, n, t.counters, t.walk_a, (int)h->first,
                                                 (int)h->count, sc, h->acc, owners);
        ++h->last_launches;
        NB_CUDA(cudaGetLastError());
        return NB_OK;
    }
    const unsigned int* tlist = nullptr;
    int ntargets = n;
    if (h->cfg.world > 1)
    {
        // owned bodies in Morton order: flags -> block sums -> scan -> compaction
        const int blocks = (n + 4095) / 4096;
        unsigned int* flag = t.flags;                 // free again after the bottom-up pass
        unsigned int* sums = t.hist;                  // radix-sort scratch is free here
        unsigned int* total = t.hist + blocks;
        k_select_flags<<<blocks_for(n, 256), 256, 0, st>>>(order, n, (int)h->first, (int)h->count, flag);
        k_block_sums<<<blocks, 256, 0, st>>>(flag, n, sums);
        k_rs_scan_rows<<<1, 256, 0, st>>>(sums, blocks, total);
        k_compact<<<blocks, 256, 0, st>>>(flag, n, sums, t.tlist);
        h->last_launches += 4;
        tlist = t.tlist;
        ntargets = (int)h->count;
    }
    const int blocks = blocks_for(ntargets, 256);
#define NB_WALK(STATS, GROUP)                                                                                            \
    k_walk<STATS, GROUP, false><<<blocks, 256, 0, st>>>(h->posw, order, tlist, ntargets, t.counters, t.walk_a, (int)h->first, \
                                                        (int)h->count, sc, h->acc, t.stats, owners)
    // kernel_variant 1 / 2: 16- / 8-lane groups (measured slower, kept for the record: DESIGN.md K7)
    const int group = h->cfg.kernel_variant == 1 ? 16 : (h->cfg.kernel_variant == 2 ? 8 : 32);
    if (g_walk_stats)
    {
        NB_CUDA(cudaMemsetAsync(t.stats, 0, kWalkStatWords * sizeof(unsigned long long), st));
        NB_WALK(true, 32);
    }
    else if (h->cfg.kernel_variant == 3)
        k_walk<false, 32, false, true><<<blocks, 256, 0, st>>>(h->posw, order, tlist, ntargets, t.counters, t.walk_a, (int)h->first,
                                                              (int)h->count, sc, h->acc, t.stats, owners);
    else if (h->cfg.kernel_variant == 4) NB_WALK(false, 32);         // the round-1 loop (one visit per trip, two loads)
    else if (h->cfg.kernel_variant == 5)      // one 256-bit load per record (measured: 39.4 ms against 38.5 ms at 16 M bodies)
        k_walk2<false, true><<<blocks, 256, 0, st>>>(h->posw, order, tlist, ntargets, t.counters, t.walk_a, (int)h->first, (int)h->count, sc,
                                                    h->acc, owners);
    else if (group == 32 && ntargets < 20000)       // 2 targets per warp
        k_walk2<false, false, 2><<<blocks_for((size_t)((ntargets + 1) / 2) * 32, 256), 256, 0, st>>>(
            h->posw, order, tlist, ntargets, t.counters, t.walk_a, (int)h->first, (int)h->count, sc, h->acc, owners);
    else if (group == 32 && ntargets < 150000)      // 8 targets per warp
        k_walk2<false, false, 8><<<blocks_for((size_t)((ntargets + 7) / 8) * 32, 256), 256, 0, st>>>(
            h->posw, order, tlist, ntargets, t.counters, t.walk_a, (int)h->first, (int)h->count, sc, h->acc, owners);
    else if (group == 32)
        k_walk2<false><<<blocks, 256, 0, st>>>(h->posw, order, tlist, ntargets, t.counters, t.walk_a, (int)h->first, (int)h->count, sc,
                                              h->acc, owners);
    else if (group == 16) NB_WALK(false, 16);
    else NB_WALK(false, 8);
#undef NB_WALK
    ++h->last_launches;
    NB_CUDA(cudaGetLastError());
    return NB_OK;
}

// Forces the lazily loaded kernels of this file into the context (CUDA 12 loads a kernel at its first
// launch, and that load can wait for the device to drain -- fatal if it happens while another handle of
// the same process sits in a peer-flag wait; see p2p.cu).
int preload_tree()
{
    cudaFuncAttributes a;
    NB_CUDA(cudaFuncGetAttributes(&a, reinterpret_cast<const void*>(k_morton)));
    NB_CUDA(cudaFuncGetAttributes(&a, reinterpret_cast<const void*>(k_rs_hist)));
    NB_CUDA(cudaFuncGetAttributes(&a, reinterpret_cast<const void*>(k_rs_scan_rows)));
    NB_CUDA(cudaFuncGetAttributes(&a, reinterpret_cast<const void*>(k_rs_scan_totals)));
    NB_CUDA(cudaFuncGetAttributes(&a, reinterpret_cast<const void*>(k_rs_scatter<false>)));
    NB_CUDA(cudaFuncGetAttributes(&a, reinterpret_cast<const void*>(k_rs_scatter<true>)));
    NB_CUDA(cudaFuncGetAttributes(&a, reinterpret_cast<const void*>(k_rs_small)));
    NB_CUDA(cudaFuncGetAttributes(&a, reinterpret_cast<const void*>(k_rs_hist_all)));
    NB_CUDA(cudaFuncGetAttributes(&a, reinterpret_cast<const void*>(k_rs_scan_all)));
    NB_CUDA(cudaFuncGetAttributes(&a, reinterpret_cast<const void*>(k_karras)));
    NB_CUDA(cudaFuncGetAttributes(&a, reinterpret_cast<const void*>(k_bottom_up)));
    NB_CUDA(cudaFuncGetAttributes(&a, reinterpret_cast<const void*>(k_build_up)));
    NB_CUDA(cudaFuncGetAttributes(&a, reinterpret_cast<const void*>(k_bottom_up_local)));
    NB_CUDA(cudaFuncGetAttributes(&a, reinterpret_cast<const void*>(k_scan_apply)));
    NB_CUDA(cudaFuncGetAttributes(&a, reinterpret_cast<const void*>(k_rank)));
    NB_CUDA(cudaFuncGetAttributes(&a, reinterpret_cast<const void*>(k_finalize)));
    NB_CUDA(cudaFuncGetAttributes(&a, reinterpret_cast<const void*>((k_walk<false, 32, false>))));
    NB_CUDA(cudaFuncGetAttributes(&a, reinterpret_cast<const void*>((k_walk<false, 32, true>))));
    NB_CUDA(cudaFuncGetAttributes(&a, reinterpret_cast<const void*>((k_walk<false, 16, false>))));
    NB_CUDA(cudaFuncGetAttributes(&a, reinterpret_cast<const void*>((k_walk<false, 8, false>))));
    NB_CUDA(cudaFuncGetAttributes(&a, reinterpret_cast<const void*>((k_walk<true, 32, false>))));
    NB_CUDA(cudaFuncGetAttributes(&a, reinterpret_cast<const void*>((k_walk2<false>))));
    NB_CUDA(cudaFuncGetAttributes(&a, reinterpret_cast<const void*>((k_walk2<true>))));
    NB_CUDA(cudaFuncGetAttributes(&a, reinterpret_cast<const void*>((k_walk2<false, true>))));
    NB_CUDA(cudaFuncGetAttributes(&a, reinterpret_cast<const void*>((k_walk2<false, false, 2>))));
    NB_CUDA(cudaFuncGetAttributes(&a, reinterpret_cast<const void*>((k_walk2<false, false, 8>))));
    NB_CUDA(cudaFuncGetAttributes(&a, reinterpret_cast<const void*>((k_walk<false, 32, false, true>))));
    NB_CUDA(cudaFuncGetAttributes(&a, reinterpret_cast<const void*>(k_leaf_cells)));
    NB_CUDA(cudaFuncGetAttributes(&a, reinterpret_cast<const void*>(k_select_flags)));
    NB_CUDA(cudaFuncGetAttributes(&a, reinterpret_cast<const void*>(k_block_sums)));
    NB_CUDA(cudaFuncGetAttributes(&a, reinterpret_cast<const void*>(k_compact)));
    NB_CUDA(cudaFuncGetAttributes(&a, reinterpret_cast<const void*>(k_compact_pairs)));
    NB_CUDA(cudaFuncGetAttributes(&a, reinterpret_cast<const void*>(k_splitters)));
    return NB_OK;
}

}  // namespace nb

using namespace nb;

extern "C" {

int nb_get_morton(nb_handle h, uint64_t* codes, uint32_t* order, size_t* n_inbounds)
{
    NB_REQUIRE(h != nullptr, NB_ERR_ARG, "null handle");
    NB_REQUIRE(h->cfg.mode == NB_MODE_BARNESHUT, NB_ERR_STATE, "handle is not in Barnes-Hut mode");
    NB_REQUIRE(h->n > 0, NB_ERR_STATE, "not initialised");
    NB_CUDA(cudaSetDevice(h->cfg.device));
    if (!h->tree.built) NB_CHECK(tree_build(h));
    NB_CUDA(cudaStreamSynchronize(h->stream));
    unsigned int m = 0;
    NB_CUDA(cudaMemcpy(&m, h->tree.counters + C_INBOUNDS, sizeof(m), cudaMemcpyDeviceToHost));
    if (n_inbounds) *n_inbounds = m;
    if (codes) NB_CUDA(cudaMemcpy(codes, h->tree.skeys, (size_t)m * sizeof(uint64_t), cudaMemcpyDeviceToHost));
    if (order) NB_CUDA(cudaMemcpy(order, h->tree.svals, (size_t)m * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    return NB_OK;
}

int nb_get_tree(nb_handle h, int32_t* left, int32_t* right, int32_t* prefix_bits, double* mass, float* com3, size_t* n_internal)
{
    NB_REQUIRE(h != nullptr, NB_ERR_ARG, "null handle");
    NB_REQUIRE(h->cfg.mode == NB_MODE_BARNESHUT, NB_ERR_STATE, "handle is not in Barnes-Hut mode");
    NB_REQUIRE(h->n > 0, NB_ERR_STATE, "not initialised");
    NB_CUDA(cudaSetDevice(h->cfg.device));
    if (!h->tree.built) NB_CHECK(tree_build(h));
    NB_CUDA(cudaStreamSynchronize(h->stream));
    unsigned int m = 0;
    NB_CUDA(cudaMemcpy(&m, h->tree.counters + C_INBOUNDS, sizeof(m), cudaMemcpyDeviceToHost));
    const size_t k = m >= 2 ? m - 1 : 0;
    if (n_internal) *n_internal = k;
    if (k == 0) return NB_OK;
    const size_t n = h->n;
    const int leaf_base = (int)n;
    std::vector<int2> ch(k);
    NB_CUDA(cudaMemcpy(ch.data(), h->tree.child, k * sizeof(int2), cudaMemcpyDeviceToHost));
    // Node ids of the parity hook are Karras's (a node is numbered by an end of its range: a left child by its last
    // slot, a right child by its first, the root 0).  k_build_up numbers a node by its split position: the left
    // child of the node splitting at g is then Karras's node g, the right child g + 1.
    std::vector<int> kid(k);
    if (h->tree.split_ids)
    {
        std::vector<int> par(k);
        NB_CUDA(cudaMemcpy(par.data(), h->tree.parent, k * sizeof(int), cudaMemcpyDeviceToHost));
        for (size_t v = 0; v < k; ++v) kid[v] = par[v] == kEnd ? 0 : (ch[(size_t)par[v]].x == (int)v ? par[v] : par[v] + 1);
    }
    else
        for (size_t v = 0; v < k; ++v) kid[v] = (int)v;
    if (left || right)
        for (size_t v = 0; v < k; ++v)
        {
            const int cl = ch[v].x, cr = ch[v].y;
            const int il = h->tree.split_ids ? (int)v : cl, ir = h->tree.split_ids ? (int)v + 1 : cr;     // Karras ids of internal children
            if (left) left[kid[v]] = cl >= leaf_base ? ~(cl - leaf_base) : il;
            if (right) right[kid[v]] = cr >= leaf_base ? ~(cr - leaf_base) : ir;
        }
    if (prefix_bits)
    {
        std::vector<int> pf(k);
        NB_CUDA(cudaMemcpy(pf.data(), h->tree.prefix, k * sizeof(int), cudaMemcpyDeviceToHost));
        for (size_t v = 0; v < k; ++v) prefix_bits[kid[v]] = pf[v];
    }
    if (mass || com3)
    {
        std::vector<double> s4(4 * k);
        NB_CUDA(cudaMemcpy(s4.data(), h->tree.nsum, 4 * k * sizeof(double), cudaMemcpyDeviceToHost));
        for (size_t v = 0; v < k; ++v)
        {
            const double w = s4[4 * v];
            const size_t i = (size_t)kid[v];
            if (mass) mass[i] = w / h->cfg.G;
            if (com3)
                for (int c = 0; c < 3; ++c) com3[3 * i + c] = (float)(w != 0.0 ? s4[4 * v + 1 + c] / w : 0.0);
        }
    }
    return NB_OK;
}

int nb_get_leaf_cells(nb_handle h, float* cells4, uint32_t* body, size_t* n_inbounds)
{
    NB_REQUIRE(h != nullptr, NB_ERR_ARG, "null handle");
    NB_REQUIRE(h->cfg.mode == NB_MODE_BARNESHUT, NB_ERR_STATE, "handle is not in Barnes-Hut mode");
    NB_REQUIRE(h->n > 0, NB_ERR_STATE, "not initialised");
    NB_CUDA(cudaSetDevice(h->cfg.device));
    if (!h->tree.built) NB_CHECK(tree_build(h));
    TreeBuffers& t = h->tree;
    NB_CUDA(cudaStreamSynchronize(h->stream));
    unsigned int m = 0;
    NB_CUDA(cudaMemcpy(&m, t.counters + C_INBOUNDS, sizeof(m), cudaMemcpyDeviceToHost));
    if (n_inbounds) *n_inbounds = m;
    if (m == 0) return NB_OK;
    if (cells4)
    {
        float4* d = t.walk_a;      // scratch: the traversal records are rewritten by every tree_build before a walk reads them
        k_leaf_cells<<<blocks_for(m, 256), 256, 0, h->stream>>>(t.skeys, t.counters, (double)h->cfg.bounds, d);
        NB_CUDA(cudaGetLastError());
        NB_CUDA(cudaMemcpyAsync(cells4, d, (size_t)m * sizeof(float4), cudaMemcpyDeviceToHost, h->stream));
        NB_CUDA(cudaStreamSynchronize(h->stream));
        ++h->total_launches;
    }
    if (body) NB_CUDA(cudaMemcpy(body, t.svals, (size_t)m * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    return NB_OK;
}

int nb_get_walk_stats(nb_handle h, uint64_t stats3[3])
{
    NB_REQUIRE(h != nullptr && stats3 != nullptr, NB_ERR_ARG, "null argument");
    NB_REQUIRE(h->cfg.mode == NB_MODE_BARNESHUT, NB_ERR_STATE, "handle is not in Barnes-Hut mode");
    NB_REQUIRE(h->n > 0, NB_ERR_STATE, "not initialised");
    NB_CUDA(cudaSetDevice(h->cfg.device));
    // one instrumented traversal of the current tree (the production walk carries no counters)
    NB_REQUIRE(h->exchanged, NB_ERR_STATE, "positions of remote ranks are stale");
    if (h->p2p_attached) NB_CHECK(p2p_wait(h));      // the peers' positions of the last step are in
    NB_CHECK(tree_build(h));
    NB_CHECK(tree_walk(h, false, true));
    NB_CUDA(cudaStreamSynchronize(h->stream));
    unsigned long long s[3];
    NB_CUDA(cudaMemcpy(s, h->tree.stats, sizeof(s), cudaMemcpyDeviceToHost));
    stats3[0] = s[0]; stats3[1] = s[1]; stats3[2] = s[2];
    return NB_OK;
}

int nb_get_walk_occupancy(nb_handle h, uint64_t hist33[33])
{
    NB_REQUIRE(h != nullptr && hist33 != nullptr, NB_ERR_ARG, "null argument");
    NB_REQUIRE(h->cfg.mode == NB_MODE_BARNESHUT, NB_ERR_STATE, "handle is not in Barnes-Hut mode");
    NB_REQUIRE(h->n > 0, NB_ERR_STATE, "not initialised");
    NB_CUDA(cudaSetDevice(h->cfg.device));
    NB_CUDA(cudaStreamSynchronize(h->stream));
    unsigned long long s[kWalkStatWords];
    NB_CUDA(cudaMemcpy(s, h->tree.stats, sizeof(s), cudaMemcpyDeviceToHost));
    for (int k = 0; k < 33; ++k) hist33[k] = s[3 + k];
    return NB_OK;
}

int nb_get_walk_sparse_load(nb_handle h, uint64_t out7[7])
{
    NB_REQUIRE(h != nullptr && out7 != nullptr, NB_ERR_ARG, "null argument");
    NB_REQUIRE(h->cfg.mode == NB_MODE_BARNESHUT, NB_ERR_STATE, "handle is not in Barnes-Hut mode");
    NB_REQUIRE(h->n > 0, NB_ERR_STATE, "not initialised");
    NB_CUDA(cudaSetDevice(h->cfg.device));
    NB_CUDA(cudaStreamSynchronize(h->stream));
    unsigned long long s[kWalkStatWords];
    NB_CUDA(cudaMemcpy(s, h->tree.stats, sizeof(s), cudaMemcpyDeviceToHost));
    for (int k = 0; k < 7; ++k) out7[k] = s[36 + k];
    return NB_OK;
}

}  // extern "C"
