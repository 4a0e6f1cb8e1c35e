/*
 * nbody_port.c -- CPU restatement of the reference's simulation path.  TEST INFRASTRUCTURE ONLY.
 *
 * Plain C restatement of matty9090/Procedural-Universe src/Sim (all-pairs gravity, the pointer
 * octree, the Barnes-Hut walk, the kick-drift integrator) plus the pieces the reference does not
 * have and this repository defines (Morton codes, Karras radix tree, energy).  Every function
 * cites the reference file:line it follows.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this; the product never does.
 *
 * PARITY PINNING: the reference's own tests do not touch src/Sim (test/ holds 3 log tests and 2
 * closest-particle tests), so there are no golden vectors to pin against.  Instead this file is
 * pinned against the reference's own code compiled headless (oracle/_ref/libpu_ref.so, built by
 * oracle/build_ref.sh): tests/test_oracle.py requires bit-identical forces, trajectories and
 * octree paths between the two, and tests/golden/ holds vectors generated from oracle/_ref.
 *
 * Build: gcc -O2 -ffp-contract=off -fPIC -shared (no -ffast-math): the float operation order
 * below is the definition of the arithmetic.
 *
 * SimpleMath semantics used (DirectXMath scalar path, see oracle/ref_shim/SimpleMath.h):
 *   length^2 = (x*x + y*y) + z*z;  Length = sqrtf;  Normalize = v / len (zero -> 0, inf -> NaN);
 *   V /= s  multiplies by 1.f/s;   DistanceSquared(a, b) uses d = b - a.
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define STRIDE 104
#define OFF_POS 0
#define OFF_VEL 48
#define OFF_FRC 72
#define OFF_MASS 96

/* Physics.hpp:9-16 */
static const double PHYS_G = 6.674e-11;
static const double PHYS_S = 1e1;
static const double PHYS_SCALE = 20 * 1.15e12;

typedef struct { float x, y, z; } v3f;

static const float* pos_of(const unsigned char* aos, size_t i) { return (const float*)(aos + i * STRIDE + OFF_POS); }
static double mass_of(const unsigned char* aos, size_t i) { return *(const double*)(aos + i * STRIDE + OFF_MASS); }

static v3f v3(const float* p) { v3f r = { p[0], p[1], p[2] }; return r; }
static v3f v3sub(v3f a, v3f b) { v3f r = { a.x - b.x, a.y - b.y, a.z - b.z }; return r; }
static float v3len2(v3f a) { return (a.x * a.x + a.y * a.y) + a.z * a.z; }
static v3f v3normalize(v3f a)
{
    const float len = sqrtf(v3len2(a));
    v3f r = { 0.f, 0.f, 0.f };
    if (len == 0.f) return r;
    if (isinf(len)) { r.x = r.y = r.z = NAN; return r; }
    r.x = a.x / len; r.y = a.y / len; r.z = a.z / len;
    return r;
}

/* Phys::Gravity(a, b) -- Physics.hpp:25-35: d = DistanceSquared (float), f = -(G ma mb)/(d + S). */
static double gravity(v3f pa, double ma, v3f pb, double mb)
{
    const double d = (double)v3len2(v3sub(pb, pa));
    return -(PHYS_G * ma * mb) / (d + PHYS_S);
}

/* ---------------------------------------------------------------------------------------------
 * All pairs -- BruteForceCPU::Exec, BruteForceCPU.cpp:25-43.  Forces (not accelerations) of
 * targets [first, first+count) accumulated in j order; out3 = count x 3 doubles.
 * ------------------------------------------------------------------------------------------- */
void port_allpairs_forces(const void* aos_, size_t n, size_t first, size_t count, double* out3)
{
    const unsigned char* aos = (const unsigned char*)aos_;
    for (size_t k = 0; k < count; ++k)
    {
        const size_t i = first + k;
        const v3f pi = v3(pos_of(aos, i));
        const double mi = mass_of(aos, i);
        double fx = 0.0, fy = 0.0, fz = 0.0;
        for (size_t j = 0; j < n; ++j)
        {
            if (i == j) continue;
            const v3f pj = v3(pos_of(aos, j));
            const v3f diff = v3normalize(v3sub(pi, pj));
            const double f = gravity(pj, mass_of(aos, j), pi, mi);
            fx += f * diff.x; fy += f * diff.y; fz += f * diff.z;
        }
        out3[3 * k] = fx; out3[3 * k + 1] = fy; out3[3 * k + 2] = fz;
    }
}

/* Integrator -- BruteForceCPU.cpp:61-73 / BarnesHut.cpp:81-95.  forces3 = n x 3 doubles. */
static void integrate(unsigned char* aos, size_t n, const double* forces3, float dt, int zero_forces)
{
    for (size_t i = 0; i < n; ++i)
    {
        unsigned char* rec = aos + i * STRIDE;
        float* pos = (float*)(rec + OFF_POS);
        double* vel = (double*)(rec + OFF_VEL);
        double* frc = (double*)(rec + OFF_FRC);
        const double m = *(double*)(rec + OFF_MASS);
        for (int c = 0; c < 3; ++c)
        {
            const double a = forces3[3 * i + c] / m;
            vel[c] += a * dt;
            const double step = (vel[c] * dt) / PHYS_SCALE;
            pos[c] += (float)step;
            frc[c] = zero_forces ? 0.0 : forces3[3 * i + c];
        }
    }
}

/* steps x BruteForceCPU::Update(dt) -- BruteForceCPU.cpp:45-74 (threading elided: the partition
 * over workers does not change any result when the worker count divides n). */
void port_allpairs_run(void* aos_, size_t n, float dt, int steps)
{
    unsigned char* aos = (unsigned char*)aos_;
    double* f = (double*)malloc(n * 3 * sizeof(double));
    for (int s = 0; s < steps; ++s)
    {
        port_allpairs_forces(aos, n, 0, n, f);
        integrate(aos, n, f, dt, 1);
    }
    free(f);
}

/* ---------------------------------------------------------------------------------------------
 * Pointer octree -- Octree.hpp:13-49, Octree.cpp:8-145.
 * ------------------------------------------------------------------------------------------- */
typedef struct node
{
    int depth, num;
    v3f lo, hi;              /* BoundingCube TopLeft / BottomRight */
    double total_mass;
    v3f com;
    int is_leaf;
    float size;
    long body;               /* P, -1 = none */
    struct node* child[8];
} node;

static const unsigned char* g_aos;   /* particles of the tree under construction */

static node* node_new(v3f lo, v3f hi, int depth)
{
    node* t = (node*)calloc(1, sizeof(node));
    t->lo = lo; t->hi = hi; t->depth = depth;
    t->size = hi.x - lo.x;        /* Octree.cpp:10 */
    t->is_leaf = 1; t->body = -1;
    return t;
}

static void node_free(node* t)
{
    if (!t) return;
    for (int i = 0; i < 8; ++i) node_free(t->child[i]);
    free(t);
}

/* BoundingCube::Contains -- Octree.hpp:18-22 (half open) */
static int contains(const node* t, v3f p)
{
    return p.x >= t->lo.x && p.y >= t->lo.y && p.z >= t->lo.z && p.x < t->hi.x && p.y < t->hi.y && p.z < t->hi.z;
}

/* Octree::Split -- Octree.cpp:16-51: eager 8 children, x fastest, corners by float accumulation */
static void split(node* t)
{
    const float size = t->size / 2;
    v3f cur = t->lo;
    int i = 0;
    for (int z = 0; z < 2; ++z)
    {
        cur.y = t->lo.y;
        for (int y = 0; y < 2; ++y)
        {
            cur.x = t->lo.x;
            for (int x = 0; x < 2; ++x, ++i)
            {
                v3f hi = { cur.x + size, cur.y + size, cur.z + size };
                t->child[i] = node_new(cur, hi, t->depth + 1);
                cur.x += size;
            }
            cur.y += size;
        }
        cur.z += size;
    }
    t->is_leaf = 0;
}

/* Octree::Add -- Octree.cpp:53-84 */
static void add(node* t, long b)
{
    const v3f p = v3(pos_of(g_aos, (size_t)b));
    if (t->num > 1)
    {
        for (int i = 0; i < 8; ++i)
            if (contains(t->child[i], p)) add(t->child[i], b);
    }
    else if (t->num == 1)
    {
        if (t->is_leaf) split(t);
        const v3f q = v3(pos_of(g_aos, (size_t)t->body));
        for (int i = 0; i < 8; ++i)
        {
            if (contains(t->child[i], p)) add(t->child[i], b);
            if (contains(t->child[i], q)) add(t->child[i], t->body);
        }
        t->body = -1;
    }
    else
    {
        t->body = b;
    }
    ++t->num;
}

/* Octree::CalculateMass -- Octree.cpp:86-105: float centre-of-mass accumulation, 1/(float)M */
static void calc_mass(node* t)
{
    if (t->num == 1)
    {
        t->com = v3(pos_of(g_aos, (size_t)t->body));
        t->total_mass = mass_of(g_aos, (size_t)t->body);
    }
    else if (!t->is_leaf)
    {
        for (int i = 0; i < 8; ++i)
        {
            node* c = t->child[i];
            calc_mass(c);
            t->total_mass += c->total_mass;
            const float w = (float)c->total_mass;
            t->com.x += c->com.x * w; t->com.y += c->com.y * w; t->com.z += c->com.z * w;
        }
        if (t->total_mass > 0.0)
        {
            const float r = 1.f / (float)t->total_mass;
            t->com.x *= r; t->com.y *= r; t->com.z *= r;
        }
    }
}

static double g_theta = 2.0;     /* Octree::Theta, Octree.cpp:5 */

/* Octree::CalculateForce -- Octree.cpp:107-145 */
static void calc_force(const node* t, long b, double* f3, int64_t* counters)
{
    const v3f p = v3(pos_of(g_aos, (size_t)b));
    const double mp = mass_of(g_aos, (size_t)b);
    if (counters) ++counters[2];
    if (t->num == 1)
    {
        if (b != t->body && !contains(t, p))
        {
            const v3f q = v3(pos_of(g_aos, (size_t)t->body));
            const double f = gravity(p, mp, q, mass_of(g_aos, (size_t)t->body));
            const v3f diff = v3normalize(v3sub(p, q));
            f3[0] += f * diff.x; f3[1] += f * diff.y; f3[2] += f * diff.z;
            if (counters) ++counters[1];
        }
    }
    else
    {
        const float r = sqrtf(v3len2(v3sub(p, t->com)));
        const float d = t->hi.x - t->lo.x;
        if (d / r < g_theta)
        {
            const double f = gravity(p, mp, t->com, t->total_mass);
            const v3f diff = v3normalize(v3sub(p, t->com));
            f3[0] += f * diff.x; f3[1] += f * diff.y; f3[2] += f * diff.z;
            if (counters) ++counters[0];
        }
        else if (!t->is_leaf)
        {
            /* `force += child->CalculateForce(p)`: each child's sum is formed separately and
             * then added (Octree.cpp:139-142) -- tree-shaped fp64 summation. */
            for (int i = 0; i < 8; ++i)
            {
                double c3[3] = { 0.0, 0.0, 0.0 };
                calc_force(t->child[i], b, c3, counters);
                f3[0] += c3[0]; f3[1] += c3[1]; f3[2] += c3[2];
            }
        }
    }
}

/* BarnesHut::BarnesHut / Update -- BarnesHut.cpp:14-19, 46-56 */
static node* build_tree(const unsigned char* aos, size_t n)
{
    const float size = 4000.0f;
    v3f lo = { -size, -size, -size }, hi = { +size, +size, +size };
    node* root = node_new(lo, hi, 0);
    g_aos = aos;
    for (size_t i = 0; i < n; ++i) add(root, (long)i);
    calc_mass(root);
    return root;
}

/* Forces on the listed targets from a tree built once (Octree::CalculateForce per target). */
void port_barneshut_forces(const void* aos_, size_t n, double theta, const int64_t* targets, size_t nt,
                           double* out3, int64_t* counters3)
{
    const unsigned char* aos = (const unsigned char*)aos_;
    node* root = build_tree(aos, n);
    g_theta = theta;
    if (counters3) counters3[0] = counters3[1] = counters3[2] = 0;
    for (size_t k = 0; k < nt; ++k)
    {
        double f3[3] = { 0.0, 0.0, 0.0 };
        calc_force(root, (long)targets[k], f3, counters3);
        out3[3 * k] = f3[0]; out3[3 * k + 1] = f3[1]; out3[3 * k + 2] = f3[2];
    }
    node_free(root);
}

/* steps x BarnesHut::Update(dt) -- BarnesHut.cpp:44-96 (Forces are left in place) */
void port_barneshut_run(void* aos_, size_t n, float dt, int steps, double theta)
{
    unsigned char* aos = (unsigned char*)aos_;
    double* f = (double*)malloc(n * 3 * sizeof(double));
    g_theta = theta;
    for (int s = 0; s < steps; ++s)
    {
        node* root = build_tree(aos, n);
        for (size_t i = 0; i < n; ++i)
        {
            double f3[3] = { 0.0, 0.0, 0.0 };
            calc_force(root, (long)i, f3, NULL);
            f[3 * i] = f3[0]; f[3 * i + 1] = f3[1]; f[3 * i + 2] = f3[2];
        }
        node_free(root);
        integrate(aos, n, f, dt, 0);
    }
    free(f);
}

/* Leaf depth and (z,y,x) digit path of every body in the pointer octree (-1 = outside the root). */
void port_octree_paths(const void* aos_, size_t n, int32_t* leaf_depth, uint64_t* path, int64_t* stats4)
{
    const unsigned char* aos = (const unsigned char*)aos_;
    node* root = build_tree(aos, n);
    int64_t maxdepth = 0;
    for (size_t i = 0; i < n; ++i)
    {
        const v3f p = v3(pos_of(aos, i));
        node* t = root;
        uint64_t code = 0;
        int depth = -1;
        if (contains(t, p))
        {
            for (;;)
            {
                if (t->num == 1 && t->body == (long)i) { depth = t->depth; break; }
                if (t->is_leaf) { depth = -2; break; }
                int next = -1;
                for (int c = 0; c < 8; ++c)
                    if (contains(t->child[c], p)) { next = c; break; }
                if (next < 0) { depth = -3; break; }
                code = (code << 3) | (uint64_t)next;
                t = t->child[next];
            }
        }
        leaf_depth[i] = depth;
        path[i] = code;
        if (depth > maxdepth) maxdepth = depth;
    }
    if (stats4) { stats4[0] = 0; stats4[1] = 0; stats4[2] = maxdepth; stats4[3] = root->num; }
    node_free(root);
}

/* ---------------------------------------------------------------------------------------------
 * Morton codes.  No reference counterpart: the reference descends by comparisons
 * (Octree.cpp:53-84 with the child order of :25-47, digit = z*4 + y*2 + x).  The code of a body is
 * the sequence of 21 such digits inside the fixed root cube [-4000, 4000)^3 (BarnesHut.cpp:14-19),
 * most significant digit first, obtained here by the SAME comparison descent on exact cell
 * corners (corner = -4000 + k * 8000 / 2^level evaluated in double, which is exact).  Bodies
 * outside the root get the code ~0 (they are dropped as sources, Octree.cpp:58-62).
 * ------------------------------------------------------------------------------------------- */
#define MORTON_LEVELS 21
#define MORTON_OUTSIDE 0xFFFFFFFFFFFFFFFFull

uint64_t port_morton_one(float x, float y, float z)
{
    const double B = 4000.0;
    if (!(x >= -B && y >= -B && z >= -B && x < B && y < B && z < B)) return MORTON_OUTSIDE;
    double lo[3] = { -B, -B, -B };
    double size = 2 * B;
    const double p[3] = { (double)x, (double)y, (double)z };
    uint64_t code = 0;
    for (int l = 0; l < MORTON_LEVELS; ++l)
    {
        size *= 0.5;
        unsigned digit = 0;
        for (int c = 0; c < 3; ++c)
        {
            const double mid = lo[c] + size;
            if (p[c] >= mid) { digit |= (1u << c); lo[c] = mid; }   /* bit0 = x, bit1 = y, bit2 = z */
        }
        code = (code << 3) | digit;
    }
    return code;
}

void port_morton(const void* aos_, size_t n, uint64_t* codes)
{
    const unsigned char* aos = (const unsigned char*)aos_;
    for (size_t i = 0; i < n; ++i)
    {
        const float* p = pos_of(aos, i);
        codes[i] = port_morton_one(p[0], p[1], p[2]);
    }
}

typedef struct { uint64_t code; uint32_t idx; } keyval;
static int cmp_keyval(const void* a, const void* b)
{
    const keyval* p = (const keyval*)a; const keyval* q = (const keyval*)b;
    if (p->code != q->code) return p->code < q->code ? -1 : 1;
    return p->idx < q->idx ? -1 : (p->idx > q->idx ? 1 : 0);
}

/* Stable sort by code (ties by body index); returns the number of in-bounds bodies. */
size_t port_morton_sorted(const void* aos, size_t n, uint64_t* sorted_codes, uint32_t* order)
{
    keyval* kv = (keyval*)malloc(n * sizeof(keyval));
    uint64_t* codes = (uint64_t*)malloc(n * sizeof(uint64_t));
    port_morton(aos, n, codes);
    size_t m = 0;
    for (size_t i = 0; i < n; ++i)
        if (codes[i] != MORTON_OUTSIDE) { kv[m].code = codes[i]; kv[m].idx = (uint32_t)i; ++m; }
    qsort(kv, m, sizeof(keyval), cmp_keyval);
    for (size_t i = 0; i < m; ++i) { sorted_codes[i] = kv[i].code; order[i] = kv[i].idx; }
    free(kv); free(codes);
    return m;
}

/* ---------------------------------------------------------------------------------------------
 * Karras radix tree over the sorted codes (T. Karras, "Maximizing Parallelism in the Construction
 * of BVHs, Octrees, and k-d Trees", HPG 2012, section 3).  delta(i, j) = length of the common
 * prefix of the 64-bit keys; equal keys are disambiguated by the slot index (64 + clz(i ^ j)).
 * Internal node i in [0, m-2]; children >= 0 are internal nodes, < 0 are ~leaf_slot.
 * prefix[i] = delta over the node's range (bits, counted on the 64-bit key whose top bit is
 * always 0 for a 63-bit code: octree level of the node = (prefix - 1) / 3).
 * ------------------------------------------------------------------------------------------- */
static int delta(const uint64_t* k, long m, long i, long j)
{
    if (j < 0 || j >= m) return -1;
    const uint64_t x = k[i] ^ k[j];
    if (x != 0) return __builtin_clzll(x);
    return 64 + __builtin_clz((uint32_t)i ^ (uint32_t)j);
}

void port_karras(const uint64_t* k, size_t m_, int32_t* left, int32_t* right, int32_t* prefix)
{
    const long m = (long)m_;
    for (long i = 0; i + 1 < m; ++i)
    {
        const int d = (delta(k, m, i, i + 1) - delta(k, m, i, i - 1)) >= 0 ? 1 : -1;
        const int dmin = delta(k, m, i, i - d);
        long lmax = 2;
        while (delta(k, m, i, i + lmax * d) > dmin) lmax *= 2;
        long l = 0;
        for (long t = lmax / 2; t >= 1; t /= 2)
            if (delta(k, m, i, i + (l + t) * d) > dmin) l += t;
        const long j = i + l * d;
        const int dnode = delta(k, m, i, j);
        long s = 0;
        long t = l;
        do
        {
            t = (t + 1) / 2;
            if (delta(k, m, i, i + (s + t) * d) > dnode) s += t;
        } while (t > 1);
        const long gamma = i + s * d + (d < 0 ? d : 0);
        const long lo = i < j ? i : j, hi = i < j ? j : i;
        left[i] = (lo == gamma) ? (int32_t)~gamma : (int32_t)gamma;
        right[i] = (hi == gamma + 1) ? (int32_t)~(gamma + 1) : (int32_t)(gamma + 1);
        prefix[i] = dnode;
    }
}

/* ---------------------------------------------------------------------------------------------
 * Energy (no reference counterpart): the pair force integrates to
 * U(r) = -(G ma mb / sqrt(S)) atan(sqrt(S) / r); positions advance by v dt / Scale, so the
 * conserved quantity is sum 1/2 m v^2 + Scale * sum_{i<j} U(r_ij).
 * ------------------------------------------------------------------------------------------- */
void port_energy(const void* aos_, size_t n, double* ke, double* pe)
{
    const unsigned char* aos = (const unsigned char*)aos_;
    double k = 0.0, u = 0.0;
    const double rs = sqrt(PHYS_S);
    for (size_t i = 0; i < n; ++i)
    {
        const double* v = (const double*)(aos + i * STRIDE + OFF_VEL);
        const double mi = mass_of(aos, i);
        k += 0.5 * mi * (v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
        const float* pi = pos_of(aos, i);
        double ui = 0.0;
        for (size_t j = i + 1; j < n; ++j)
        {
            const float* pj = pos_of(aos, j);
            const double dx = (double)pj[0] - pi[0], dy = (double)pj[1] - pi[1], dz = (double)pj[2] - pi[2];
            const double r = sqrt(dx * dx + dy * dy + dz * dz);
            const double a = r > 0.0 ? atan(rs / r) : 1.5707963267948966;
            ui -= mass_of(aos, j) * a;
        }
        u += ui * mi;
    }
    *ke = k;
    *pe = u * PHYS_G / rs * PHYS_SCALE;
}
