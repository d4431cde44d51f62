// dsb_api.cu -- host side of the C ABI declared in include/disimpy_b200.h.
// Owns device memory, the stream and the launch sequence; no PyTorch, no Python.
#include "../../include/disimpy_b200.h"
#include "dsb_fill.cuh"
#include "dsb_kernels.cuh"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <thread>
#include <new>
#include <string>
#include <vector>

#include <dlfcn.h>
#include <nvtx3/nvToolsExt.h>

namespace {

thread_local std::string g_err;

// NVTX range around every phase of a call (upload, RNG init, walk, reduction, read-back): shows up
// in nsys / ncu timelines, costs nothing when no tool is attached.
struct Range {
    explicit Range(const char *name) { nvtxRangePushA(name); }
    ~Range() { nvtxRangePop(); }
    Range(const Range &) = delete;
    Range &operator=(const Range &) = delete;
};

int fail(int code, const std::string &msg)
{
    g_err = msg;
    return code;
}

// Nothing may throw across the C boundary: entry points that allocate host memory run their body
// through this.
template <typename Fn>
int guarded(Fn fn)
{
    try {
        return fn();
    } catch (const std::bad_alloc &) {
        return fail(DSB_ENOMEM, "host allocation failed");
    } catch (const std::exception &e) {
        return fail(DSB_EINVAL, std::string("unexpected exception: ") + e.what());
    } catch (...) {
        return fail(DSB_EINVAL, "unexpected exception");
    }
}

#define DSB_CUDA(expr)                                                                            \
    do {                                                                                          \
        cudaError_t e_ = (expr);                                                                  \
        if (e_ != cudaSuccess)                                                                    \
            return fail(e_ == cudaErrorMemoryAllocation ? DSB_ENOMEM : DSB_ECUDA,                 \
                        std::string(#expr) + ": " + cudaGetErrorString(e_));                      \
    } while (0)

// ------------------------------------------------------------- xoroshiro128+ jump matrices

struct HState {
    uint64_t s0, s1;
};

inline uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }

inline void h_next(HState &s)
{
    uint64_t s0 = s.s0, s1 = s.s1;
    s1 ^= s0;
    s.s0 = rotl(s0, 55) ^ s1 ^ (s1 << 14);
    s.s1 = rotl(s1, 36);
}

// numba/cuda/random.py:102-126 (the published xoroshiro128+ 2^64 jump polynomial)
inline HState h_jump(HState s)
{
    static const uint64_t poly[2] = {0xbeac0467eba5facbULL, 0xd86b048b86aa9922ULL};
    HState acc = {0, 0};
    for (int i = 0; i < 2; ++i)
        for (int b = 0; b < 64; ++b) {
            if (poly[i] & (1ULL << b)) {
                acc.s0 ^= s.s0;
                acc.s1 ^= s.s1;
            }
            h_next(s);
        }
    return acc;
}

inline HState h_apply(const HState *cols, HState v)
{
    HState acc = {0, 0};
    for (int b = 0; b < 64; ++b)
        if ((v.s0 >> b) & 1) {
            acc.s0 ^= cols[b].s0;
            acc.s1 ^= cols[b].s1;
        }
    for (int b = 0; b < 64; ++b)
        if ((v.s1 >> b) & 1) {
            acc.s0 ^= cols[64 + b].s0;
            acc.s1 ^= cols[64 + b].s1;
        }
    return acc;
}

// pows[k*128 + b] = J^(2^k) e_b.  The jump is linear over GF(2), so J's columns are the jumps of
// the unit vectors and each further power is the previous matrix applied to its own columns.
const std::vector<HState> &jump_powers()
{
    static std::vector<HState> pows;
    static std::once_flag once;
    std::call_once(once, [] {
        pows.resize(64 * 128);
        for (int b = 0; b < 128; ++b) {
            HState e = {b < 64 ? (1ULL << b) : 0, b >= 64 ? (1ULL << (b - 64)) : 0};
            pows[b] = h_jump(e);
        }
        for (int k = 1; k < 64; ++k)
            for (int b = 0; b < 128; ++b) pows[k * 128 + b] = h_apply(&pows[(k - 1) * 128], pows[(k - 1) * 128 + b]);
    });
    return pows;
}

// numba/cuda/random.py:46-69
inline uint64_t splitmix64(uint64_t seed)
{
    uint64_t z = seed + 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

std::mutex g_pow_mu;
ulonglong2 *g_dev_pows[64] = {nullptr};  // per device ordinal

// The upload is ordered on `st` (the stream the caller launches rng_init_kernel on) and waited for
// before the table is published: later callers on other non-blocking streams may use it at once.
int device_jump_powers(int device, cudaStream_t st, ulonglong2 **out)
{
    if (device < 0 || device >= 64) return fail(DSB_EINVAL, "device ordinal out of range");
    std::lock_guard<std::mutex> lk(g_pow_mu);
    if (!g_dev_pows[device]) {
        const std::vector<HState> &p = jump_powers();
        ulonglong2 *d = nullptr;
        DSB_CUDA(cudaMalloc(&d, p.size() * sizeof(HState)));
        DSB_CUDA(cudaMemcpyAsync(d, p.data(), p.size() * sizeof(HState), cudaMemcpyHostToDevice, st));
        DSB_CUDA(cudaStreamSynchronize(st));
        g_dev_pows[device] = d;
    }
    *out = g_dev_pows[device];
    return DSB_OK;
}

int launch_rng_init(int device, uint64_t seed, uint64_t start, int64_t n, ulonglong2 *d_out, cudaStream_t st)
{
    if (n <= 0) return DSB_OK;
    ulonglong2 *pows = nullptr;
    int rc = device_jump_powers(device, st, &pows);
    if (rc) return rc;
    int64_t blocks = (n + 256 * dsb::kRngInitRun - 1) / (256 * dsb::kRngInitRun);
    dsb::rng_init_kernel<<<(unsigned)blocks, 256, 0, st>>>(splitmix64(seed), start, (long long)n, pows, d_out);
    DSB_CUDA(cudaGetLastError());
    return DSB_OK;
}

// ------------------------------------------------------------- device buffer cache

// cudaMalloc / cudaFree synchronise the device and cost from a few to tens of milliseconds per
// simulation() call (more when another allocator holds most of the memory), so buffers of
// destroyed handles are kept per device and size and handed to the next handle.
size_t cache_cap_from_env(size_t dflt, size_t divisor)
{
    const char *e = getenv("DISIMPY_B200_CACHE_MB");
    if (!e || !*e) return dflt;
    char *end = nullptr;
    const unsigned long long mb = strtoull(e, &end, 10);
    if (end == e) return dflt;
    return (size_t)mb * (size_t(1) << 20) / divisor;
}

struct BufferCache {
    std::mutex mu;
    std::multimap<std::pair<int, size_t>, void *> idle;
    std::map<void *, std::pair<int, size_t>> live;
    size_t idle_bytes = 0;
    // idle device memory kept for the next handle; DISIMPY_B200_CACHE_MB overrides (0: keep nothing)
    size_t max_idle_bytes = cache_cap_from_env(size_t(16) << 30, 1);
} g_cache;

cudaError_t cache_malloc(void **out, size_t bytes)
{
    int dev = 0;
    cudaGetDevice(&dev);
    bytes = std::max<size_t>(bytes, 1);
    {
        std::lock_guard<std::mutex> lk(g_cache.mu);
        auto it = g_cache.idle.find({dev, bytes});
        if (it != g_cache.idle.end()) {
            *out = it->second;
            g_cache.idle.erase(it);
            g_cache.idle_bytes -= bytes;
            g_cache.live[*out] = {dev, bytes};
            return cudaSuccess;
        }
    }
    cudaError_t e = cudaMalloc(out, bytes);
    if (e == cudaErrorMemoryAllocation) {  // give the idle buffers back and retry
        cudaGetLastError();
        dsb_release_cache();
        e = cudaMalloc(out, bytes);
    }
    if (e == cudaSuccess) {
        std::lock_guard<std::mutex> lk(g_cache.mu);
        g_cache.live[*out] = {dev, bytes};
    }
    return e;
}

void cache_free(void *p)
{
    if (!p) return;
    std::lock_guard<std::mutex> lk(g_cache.mu);
    auto it = g_cache.live.find(p);
    if (it == g_cache.live.end()) {
        cudaFree(p);
        return;
    }
    const std::pair<int, size_t> key = it->second;
    g_cache.live.erase(it);
    if (g_cache.idle_bytes + key.second > g_cache.max_idle_bytes) {
        cudaFree(p);
        return;
    }
    g_cache.idle.insert({key, p});
    g_cache.idle_bytes += key.second;
}

template <typename T>
cudaError_t cache_malloc(T **out, size_t bytes)
{
    return cache_malloc(reinterpret_cast<void **>(out), bytes);
}

// ------------------------------------------------------------- pinned host scratch

// The mesh re-layout below fills a few hundred MB of host arrays per handle.  Fresh pages for
// them cost more than the arithmetic (first-touch faults), and pageable memory uploads at a
// fraction of the PCIe rate, so these arrays are page-locked blocks kept by size like the device
// buffers above and handed from one handle to the next.
struct HostCache {
    std::mutex mu;
    std::multimap<size_t, void *> idle;
    std::map<void *, size_t> live;
    size_t idle_bytes = 0;
    size_t max_idle_bytes = cache_cap_from_env(size_t(4) << 30, 4);  // a quarter of the device cap
} g_host_cache;

void *host_cache_malloc(size_t bytes)
{
    bytes = std::max<size_t>(bytes, 1);
    {
        std::lock_guard<std::mutex> lk(g_host_cache.mu);
        auto it = g_host_cache.idle.find(bytes);
        if (it != g_host_cache.idle.end()) {
            void *p = it->second;
            g_host_cache.idle.erase(it);
            g_host_cache.idle_bytes -= bytes;
            g_host_cache.live[p] = bytes;
            return p;
        }
    }
    void *p = nullptr;
    if (cudaMallocHost(&p, bytes) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    std::lock_guard<std::mutex> lk(g_host_cache.mu);
    g_host_cache.live[p] = bytes;
    return p;
}

void host_cache_free(void *p)
{
    if (!p) return;
    std::lock_guard<std::mutex> lk(g_host_cache.mu);
    auto it = g_host_cache.live.find(p);
    if (it == g_host_cache.live.end()) {
        cudaFreeHost(p);
        return;
    }
    const size_t bytes = it->second;
    g_host_cache.live.erase(it);
    if (g_host_cache.idle_bytes + bytes > g_host_cache.max_idle_bytes) {
        cudaFreeHost(p);
        return;
    }
    g_host_cache.idle.insert({bytes, p});
    g_host_cache.idle_bytes += bytes;
}

// move-only array in that memory (uninitialised)
template <typename T>
struct HostBuf {
    T *p = nullptr;
    size_t n = 0;
    HostBuf() = default;
    explicit HostBuf(size_t count) : p(static_cast<T *>(host_cache_malloc(std::max<size_t>(count, 1) * sizeof(T)))), n(count) {}
    HostBuf(HostBuf &&o) noexcept : p(o.p), n(o.n) { o.p = nullptr; o.n = 0; }
    HostBuf &operator=(HostBuf &&o) noexcept
    {
        if (this != &o) {
            host_cache_free(p);
            p = o.p;
            n = o.n;
            o.p = nullptr;
            o.n = 0;
        }
        return *this;
    }
    HostBuf(const HostBuf &) = delete;
    HostBuf &operator=(const HostBuf &) = delete;
    ~HostBuf() { host_cache_free(p); }
    T &operator[](size_t i) { return p[i]; }
    const T &operator[](size_t i) const { return p[i]; }
    T *data() { return p; }
    const T *data() const { return p; }
    size_t size() const { return n; }
    bool ok() const { return p != nullptr; }
};

// ------------------------------------------------------------- mesh upload

// fn(begin, end) over [0, n) in contiguous pieces on a few host threads (the re-layout loops below
// are independent per element; a 1e6-triangle mesh has ~1e7 of them)
template <typename Fn>
void parallel_ranges(int64_t n, Fn fn, int64_t grain = 65536)
{
    unsigned hw = std::thread::hardware_concurrency();
    const int n_threads = (int)std::max<int64_t>(1, std::min<int64_t>({8, (int64_t)(hw ? hw : 1), n / grain}));
    if (n_threads <= 1) {
        fn((int64_t)0, n);
        return;
    }
    std::vector<std::thread> pool;
    for (int t = 0; t < n_threads; ++t) pool.emplace_back(fn, n * t / n_threads, n * (t + 1) / n_threads);
    for (auto &th : pool) th.join();
}


struct MeshBuffers {
    double *tri = nullptr;
    double *normal = nullptr;
    int *tri_idx = nullptr;
    uint4 *entry = nullptr;
    int2 *cell_rng = nullptr;
    double *xs = nullptr, *ys = nullptr, *zs = nullptr;
    dsb::MeshDev dev{};
    // host copies the initial-position sampler's column lists are built from on first use
    HostBuf<int2> h_cells;
    HostBuf<int> h_tri_idx;
    HostBuf<uint4> h_box;
    int64_t n_sv[3] = {0, 0, 0};
    int *col_start = nullptr, *col_cnt = nullptr;
    uint4 *col_entry = nullptr;
    dsb::FillColumns columns{};
    // the refined search grid (null when the walk searches the reference grid itself)
    uint4 *f_entry = nullptr;
    int2 *f_cell_rng = nullptr;
    double *f_xs = nullptr, *f_ys = nullptr, *f_zs = nullptr;
    int refine[3] = {1, 1, 1};
    int64_t data_bytes = 0, grid_bytes = 0;   // triangles + normals; list entries + cell ranges of the grid the walk searches
    std::mutex mu;   // the sampler's column lists are built on first use; handles may share one upload
    MeshBuffers() = default;
    MeshBuffers(const MeshBuffers &) = delete;
    MeshBuffers &operator=(const MeshBuffers &) = delete;
    ~MeshBuffers() { release(); }
    void release()
    {
        cache_free(f_entry);
        cache_free(f_cell_rng);
        cache_free(f_xs);
        cache_free(f_ys);
        cache_free(f_zs);
        cache_free(col_start);
        cache_free(col_cnt);
        cache_free(col_entry);
        cache_free(tri);
        cache_free(normal);
        cache_free(tri_idx);
        cache_free(entry);
        cache_free(cell_rng);
        cache_free(xs);
        cache_free(ys);
        cache_free(zs);
        f_entry = nullptr, f_cell_rng = nullptr, f_xs = f_ys = f_zs = nullptr;
        col_start = col_cnt = nullptr, col_entry = nullptr;
        tri = normal = nullptr, tri_idx = nullptr, entry = nullptr, cell_rng = nullptr, xs = ys = zs = nullptr;
        columns = dsb::FillColumns{};
        dev = dsb::MeshDev{};
        h_cells = HostBuf<int2>();
        h_tri_idx = HostBuf<int>();
        h_box = HostBuf<uint4>();
    }
};

// Re-lays the reference's mesh arrays out for the kernels: int64 indices become int32, and the
// three vertex gathers per test (simulations.py:100-118) become one 16-byte aligned 80-byte
// record per triangle holding A, B-A, C-A (72 bytes) and a pad, read with five 128-bit loads.
// How finely the search grid cuts the reference's cells, per axis.  The collision search reads the
// lists of the cells the remaining step segment overlaps; with cells much larger than the step most
// of a list is far from the segment.  Cutting a cell into sub-cells not smaller than 1.05 steps keeps
// a segment within two sub-cells per axis (the kernel's short-step variant) and shortens the lists
// it reads (measured on the 1e6-triangle mesh of BASELINE config 5, 1.92 um cells, 0.69 um steps).
// DISIMPY_B200_REFINE = "kx,ky,kz" overrides (tests force odd factors), "0" or "1" turns it off.
void choose_refinement(const dsb_mesh &m, double step_l, int64_t n_entries, int k[3])
{
    k[0] = k[1] = k[2] = 1;
    const char *env = getenv("DISIMPY_B200_REFINE");
    if (env && *env) {
        int a = 1, b = 1, c = 1;
        const int got = sscanf(env, "%d,%d,%d", &a, &b, &c);
        if (got == 1) b = c = a;
        if (got >= 1) {
            k[0] = std::min(std::max(a, 1), 8);
            k[1] = std::min(std::max(b, 1), 8);
            k[2] = std::min(std::max(c, 1), 8);
        }
    } else {
        if (!(step_l > 0)) return;
        const double *grid[3] = {m.xs, m.ys, m.zs};
        for (int a = 0; a < 3; ++a) {
            const double h = (grid[a][m.n_sv[a]] - grid[a][0]) / (double)m.n_sv[a];
            const double f = std::floor(h / (1.05 * step_l));
            k[a] = f >= 4.0 ? 4 : (f >= 1.0 ? (int)f : 1);
        }
    }
    // bounded tables: at most 2^24 sub-cells, and (a triangle is listed once per sub-cell its box
    // reaches into) not more than about 2^26 entries
    const int64_t n_cells = m.n_sv[0] * m.n_sv[1] * m.n_sv[2];
    for (;;) {
        const int64_t kk = (int64_t)k[0] * k[1] * k[2];
        if (kk == 1 || (n_cells * kk <= (int64_t(1) << 24) && n_entries * kk <= (int64_t(1) << 26))) break;
        int big = 0;
        for (int a = 1; a < 3; ++a)
            if (k[a] > k[big]) big = a;
        --k[big];
    }
}

// The refined search grid (dsb::SearchGrid): boundaries, per-sub-cell ranges and entry records.  A
// sub-cell lists, in the parent's order, the triangles of its parent cell whose bounding box, padded
// by 1e-9 of the voxel, reaches into it (closed intervals): a superset of the triangles that touch
// the sub-cell, a subset of the parent's list.
int build_search_grid(const dsb_mesh &m, const HostBuf<int2> &cells, const HostBuf<int> &tri_idx, const HostBuf<uint4> &box,
                      const int k[3], MeshBuffers &mb)
{
    Range nvtx("dsb: refined search grid");
    const int64_t n[3] = {m.n_sv[0], m.n_sv[1], m.n_sv[2]};
    const int64_t nf[3] = {n[0] * k[0], n[1] * k[1], n[2] * k[2]};
    const int64_t n_fine = nf[0] * nf[1] * nf[2];
    const double *grid[3] = {m.xs, m.ys, m.zs};
    std::vector<double> fb[3];
    for (int a = 0; a < 3; ++a) {
        fb[a].resize((size_t)nf[a] + 1);
        for (int64_t i = 0; i < n[a]; ++i)
            for (int j = 0; j < k[a]; ++j)
                fb[a][(size_t)(i * k[a] + j)] = j == 0 ? grid[a][i] : grid[a][i] + (grid[a][i + 1] - grid[a][i]) * ((double)j / k[a]);
        fb[a][(size_t)nf[a]] = grid[a][n[a]];
        for (int64_t i = 0; i < nf[a]; ++i)
            if (!(fb[a][(size_t)i] < fb[a][(size_t)i + 1])) return fail(DSB_EINVAL, "mesh: subvoxel boundaries are not increasing");
    }
    // padded bounding boxes of the triangles
    HostBuf<double> tb((size_t)m.n_faces * 6);
    if (!tb.ok()) return fail(DSB_ENOMEM, "mesh: no page-locked host memory for the search grid");
    double pad[3];
    for (int a = 0; a < 3; ++a) pad[a] = 1e-9 * std::fabs(grid[a][n[a]] - grid[a][0]);
    parallel_ranges(m.n_faces, [&](int64_t f0, int64_t f1) {
        for (int64_t f = f0; f < f1; ++f) {
            const int64_t *idx = m.faces + 3 * f;
            for (int a = 0; a < 3; ++a) {
                const double p = m.vertices[3 * idx[0] + a], q = m.vertices[3 * idx[1] + a], r = m.vertices[3 * idx[2] + a];
                double lo = std::min(p, std::min(q, r)) - pad[a], hi = std::max(p, std::max(q, r)) + pad[a];
                if (!(lo == lo) || !(hi == hi)) lo = -INFINITY, hi = INFINITY;   // NaN vertex: listed everywhere
                tb[(size_t)f * 6 + a] = lo;
                tb[(size_t)f * 6 + 3 + a] = hi;
            }
        }
    });
    HostBuf<int2> fcells((size_t)n_fine);
    if (!fcells.ok()) return fail(DSB_ENOMEM, "mesh: no page-locked host memory for the search grid");
    const int64_t n_coarse = n[0] * n[1] * n[2];
    auto for_sub_cells = [&](int64_t c, auto visit) {   // visit(fine index, triangle id) for every listed pair of parent c
        const int64_t cz = c % n[2], cy = (c / n[2]) % n[1], cx = c / (n[1] * n[2]);
        const int2 r = cells[(size_t)c];
        for (int sx = 0; sx < k[0]; ++sx)
            for (int sy = 0; sy < k[1]; ++sy)
                for (int sz = 0; sz < k[2]; ++sz) {
                    const int64_t fx = cx * k[0] + sx, fy = cy * k[1] + sy, fz = cz * k[2] + sz;
                    const double lo[3] = {fb[0][(size_t)fx], fb[1][(size_t)fy], fb[2][(size_t)fz]};
                    const double hi[3] = {fb[0][(size_t)fx + 1], fb[1][(size_t)fy + 1], fb[2][(size_t)fz + 1]};
                    const int64_t fi = (fx * nf[1] + fy) * nf[2] + fz;
                    for (int i = r.x; i < r.y; ++i) {
                        const int t = tri_idx[(size_t)i];
                        const double *b = &tb[(size_t)t * 6];
                        if (b[0] <= hi[0] && b[3] >= lo[0] && b[1] <= hi[1] && b[4] >= lo[1] && b[2] <= hi[2] && b[5] >= lo[2])
                            visit(fi, t);
                    }
                }
    };
    parallel_ranges(n_fine, [&](int64_t a, int64_t b) {
        for (int64_t i = a; i < b; ++i) fcells[(size_t)i] = make_int2(0, 0);
    });
    const int64_t grain = 8192;   // (k^3 x list length box tests per parent cell: worth a thread much earlier than a copy loop)
    parallel_ranges(n_coarse, [&](int64_t c0, int64_t c1) {
        for (int64_t c = c0; c < c1; ++c) for_sub_cells(c, [&](int64_t fi, int) { ++fcells[(size_t)fi].y; });
    }, grain);
    int64_t total = 0;
    for (int64_t i = 0; i < n_fine; ++i) {
        const int cnt = fcells[(size_t)i].y;
        if (total + cnt > 0x7fffffffLL) return fail(DSB_EINVAL, "mesh: refined search grid too large");
        fcells[(size_t)i] = make_int2((int)total, (int)total);   // .y is the fill cursor of the next pass
        total += cnt;
    }
    HostBuf<uint4> fentry((size_t)total + 1);
    if (!fentry.ok()) return fail(DSB_ENOMEM, "mesh: no page-locked host memory for the search grid");
    parallel_ranges(n_coarse, [&](int64_t c0, int64_t c1) {   // (a sub-cell belongs to one parent: no two threads share a cursor)
        for (int64_t c = c0; c < c1; ++c)
            for_sub_cells(c, [&](int64_t fi, int t) { fentry[(size_t)fcells[(size_t)fi].y++] = box[(size_t)t]; });
    }, grain);
    fentry[(size_t)total] = make_uint4(0u, 0u, 0u, 0u);
    DSB_CUDA(cache_malloc(&mb.f_entry, fentry.size() * sizeof(uint4)));
    DSB_CUDA(cache_malloc(&mb.f_cell_rng, fcells.size() * sizeof(int2)));
    DSB_CUDA(cache_malloc(&mb.f_xs, fb[0].size() * sizeof(double)));
    DSB_CUDA(cache_malloc(&mb.f_ys, fb[1].size() * sizeof(double)));
    DSB_CUDA(cache_malloc(&mb.f_zs, fb[2].size() * sizeof(double)));
    DSB_CUDA(cudaMemcpy(mb.f_entry, fentry.data(), fentry.size() * sizeof(uint4), cudaMemcpyHostToDevice));
    DSB_CUDA(cudaMemcpy(mb.f_cell_rng, fcells.data(), fcells.size() * sizeof(int2), cudaMemcpyHostToDevice));
    DSB_CUDA(cudaMemcpy(mb.f_xs, fb[0].data(), fb[0].size() * sizeof(double), cudaMemcpyHostToDevice));
    DSB_CUDA(cudaMemcpy(mb.f_ys, fb[1].data(), fb[1].size() * sizeof(double), cudaMemcpyHostToDevice));
    DSB_CUDA(cudaMemcpy(mb.f_zs, fb[2].data(), fb[2].size() * sizeof(double), cudaMemcpyHostToDevice));
    DSB_CUDA(cudaDeviceSynchronize());  // (pageable sources, see upload_mesh)
    mb.grid_bytes = (int64_t)(fentry.size() * sizeof(uint4) + fcells.size() * sizeof(int2));
    dsb::SearchGrid &g = mb.dev.fine;
    g.entry = mb.f_entry;
    g.cell_rng = mb.f_cell_rng;
    g.xs = mb.f_xs;
    g.ys = mb.f_ys;
    g.zs = mb.f_zs;
    g.len_xs = (int)nf[0] + 1;
    g.len_ys = (int)nf[1] + 1;
    g.len_zs = (int)nf[2] + 1;
    g.nsv1 = (int)nf[1];
    g.nsv2 = (int)nf[2];
    g.inv_hx = nf[0] / (grid[0][n[0]] - grid[0][0]);
    g.inv_hy = nf[1] / (grid[1][n[1]] - grid[1][0]);
    g.inv_hz = nf[2] / (grid[2][n[2]] - grid[2][0]);
    for (int a = 0; a < 3; ++a) mb.refine[a] = k[a];
    return DSB_OK;
}

int upload_mesh(const dsb_mesh &m, MeshBuffers &mb, double step_l = 0.0)
{
    Range nvtx("dsb: mesh re-layout + upload");
    if (!m.vertices || !m.faces || !m.xs || !m.ys || !m.zs || !m.subvoxel_indices ||
        (m.n_triangle_indices > 0 && !m.triangle_indices))
        return fail(DSB_EINVAL, "mesh: null array");
    if (m.n_faces <= 0 || m.n_vertices <= 0 || m.n_faces > 0x7fffffffLL || m.n_triangle_indices > 0x7fffffffLL)
        return fail(DSB_EINVAL, "mesh: sizes out of range");
    for (int k = 0; k < 3; ++k)
        if (m.n_sv[k] <= 0 || m.n_sv[k] > 1 << 20) return fail(DSB_EINVAL, "mesh: n_sv out of range");
    const int64_t n_cells = m.n_sv[0] * m.n_sv[1] * m.n_sv[2];
    if (n_cells > 0x7fffffffLL) return fail(DSB_EINVAL, "mesh: too many subvoxels");
    HostBuf<double> tri((size_t)m.n_faces * dsb::kTriStride);
    HostBuf<uint4> box((size_t)m.n_faces);
    HostBuf<int> tri_idx((size_t)m.n_triangle_indices);
    HostBuf<uint4> entry((size_t)m.n_triangle_indices + 1);
    HostBuf<int2> cells((size_t)n_cells);
    if (!tri.ok() || !box.ok() || !tri_idx.ok() || !entry.ok() || !cells.ok())
        return fail(DSB_ENOMEM, "mesh: no page-locked host memory for the re-layout");
    std::atomic<bool> bad_index{false};
    parallel_ranges(m.n_faces, [&](int64_t f0, int64_t f1) {
        for (int64_t f = f0; f < f1; ++f) {
            const int64_t *idx = m.faces + 3 * f;
            bool ok = true;
            for (int c = 0; c < 3; ++c) ok = ok && idx[c] >= 0 && idx[c] < m.n_vertices;
            if (!ok) {
                bad_index = true;
                continue;
            }
            const double *A = m.vertices + 3 * idx[0], *B = m.vertices + 3 * idx[1], *C = m.vertices + 3 * idx[2];
            double *o = &tri[(size_t)f * dsb::kTriStride];
            for (int c = 0; c < 3; ++c) {
                o[c] = A[c];
                o[3 + c] = B[c] - A[c];
                o[6 + c] = C[c] - A[c];
            }
            for (int c = 9; c < dsb::kTriStride; ++c) o[c] = 0.0;
        }
    });
    if (bad_index) return fail(DSB_EINVAL, "mesh: face index out of range");
    // Box of every triangle on a 15-bit grid over [0, xs[-1]] x [0, ys[-1]] x [0, zs[-1]], rounded
    // outwards by a grid unit, for the kernels' pre-test; stored as (lo, 32767 - hi) halfwords so
    // that "boxes meet" is one direction of comparison for all six numbers.
    const double tops[3] = {m.xs[m.n_sv[0]], m.ys[m.n_sv[1]], m.zs[m.n_sv[2]]};
    parallel_ranges(m.n_faces, [&](int64_t f0, int64_t f1) {
        for (int64_t f = f0; f < f1; ++f) {
            const int64_t *idx = m.faces + 3 * f;
            unsigned lo[3], hi[3];
            for (int k = 0; k < 3; ++k) {
                const double a = m.vertices[3 * idx[0] + k], b = m.vertices[3 * idx[1] + k], c = m.vertices[3 * idx[2] + k];
                const double mn = std::min(a, std::min(b, c)), mx = std::max(a, std::max(b, c));
                const double scale = 32767.0 / tops[k];
                double ql = std::floor(mn * scale) - 1.0, qh = std::ceil(mx * scale) + 1.0;
                if (!(ql > 0.0)) ql = 0.0;        // also NaN: never filtered out
                if (!(qh < 32767.0)) qh = 32767.0;
                if (!(mn == mn) || !(mx == mx)) ql = 0.0, qh = 32767.0;
                lo[k] = (unsigned)std::min(ql, 32767.0);
                hi[k] = 32767u - (unsigned)std::max(qh, 0.0);
            }
            // Seen along +x (the sampler's ray, dsb_fill.cuh) a triangle whose projection on the yz plane
            // is a sliver has a determinant that is mostly rounding error, and the reference's test may
            // then accept points far outside its box: such triangles are marked and always tested exactly.
            const double *o = &tri[(size_t)f * dsb::kTriStride];
            const double det = std::fma(o[7], o[5], -(o[8] * o[4]));  // as ray_triangle forms it for ray = (1,0,0)
            const double scale2_yz = (o[4] * o[4] + o[5] * o[5]) * (o[7] * o[7] + o[8] * o[8]);
            const bool edge_on = det != 0.0 && !(det * det > 1e-18 * scale2_yz);
            box[(size_t)f] = make_uint4((unsigned)f | (edge_on ? ~dsb::kEntryTriMask : 0u), lo[0] | (lo[1] << 16),
                                        lo[2] | (hi[0] << 16), hi[1] | (hi[2] << 16));
        }
    });
    entry[(size_t)m.n_triangle_indices] = make_uint4(0u, 0u, 0u, 0u);
    parallel_ranges(m.n_triangle_indices, [&](int64_t i0, int64_t i1) {
        for (int64_t i = i0; i < i1; ++i) {
            if (m.triangle_indices[i] < 0 || m.triangle_indices[i] >= m.n_faces) {
                bad_index = true;
                continue;
            }
            tri_idx[(size_t)i] = (int)m.triangle_indices[i];
            entry[(size_t)i] = box[(size_t)m.triangle_indices[i]];
        }
    });
    if (bad_index) return fail(DSB_EINVAL, "mesh: triangle index out of range");
    parallel_ranges(n_cells, [&](int64_t c0, int64_t c1) {
        for (int64_t c = c0; c < c1; ++c) {
            int64_t a = m.subvoxel_indices[2 * c], b = m.subvoxel_indices[2 * c + 1];
            if (a < 0 || b < a || b > m.n_triangle_indices) {
                bad_index = true;
                continue;
            }
            cells[(size_t)c] = make_int2((int)a, (int)b);
        }
    });
    if (bad_index) return fail(DSB_EINVAL, "mesh: subvoxel range out of bounds");
    DSB_CUDA(cache_malloc(&mb.tri, tri.size() * sizeof(double)));
    DSB_CUDA(cache_malloc(&mb.tri_idx, (tri_idx.size() + 1) * sizeof(int)));
    DSB_CUDA(cache_malloc(&mb.entry, entry.size() * sizeof(uint4)));
    DSB_CUDA(cudaMemcpy(mb.entry, entry.data(), entry.size() * sizeof(uint4), cudaMemcpyHostToDevice));
    DSB_CUDA(cache_malloc(&mb.cell_rng, cells.size() * sizeof(int2)));
    DSB_CUDA(cache_malloc(&mb.xs, (m.n_sv[0] + 1) * sizeof(double)));
    DSB_CUDA(cache_malloc(&mb.ys, (m.n_sv[1] + 1) * sizeof(double)));
    DSB_CUDA(cache_malloc(&mb.zs, (m.n_sv[2] + 1) * sizeof(double)));
    DSB_CUDA(cudaMemcpy(mb.tri, tri.data(), tri.size() * sizeof(double), cudaMemcpyHostToDevice));
    DSB_CUDA(cache_malloc(&mb.normal, sizeof(double) * 3 * (size_t)m.n_faces));
    dsb::tri_normal_kernel<<<(unsigned)((m.n_faces + 255) / 256), 256>>>(mb.tri, (long long)m.n_faces, mb.normal);
    DSB_CUDA(cudaGetLastError());
    DSB_CUDA(cudaDeviceSynchronize());
    if (tri_idx.size() > 0)
        DSB_CUDA(cudaMemcpy(mb.tri_idx, tri_idx.data(), tri_idx.size() * sizeof(int), cudaMemcpyHostToDevice));
    DSB_CUDA(cudaMemcpy(mb.cell_rng, cells.data(), cells.size() * sizeof(int2), cudaMemcpyHostToDevice));
    DSB_CUDA(cudaMemcpy(mb.xs, m.xs, (m.n_sv[0] + 1) * sizeof(double), cudaMemcpyHostToDevice));
    DSB_CUDA(cudaMemcpy(mb.ys, m.ys, (m.n_sv[1] + 1) * sizeof(double), cudaMemcpyHostToDevice));
    DSB_CUDA(cudaMemcpy(mb.zs, m.zs, (m.n_sv[2] + 1) * sizeof(double), cudaMemcpyHostToDevice));
    // copies from pageable memory may return before the DMA lands, and the handle's streams do not
    // synchronise with the legacy stream: wait here
    DSB_CUDA(cudaDeviceSynchronize());
    mb.data_bytes = (int64_t)(tri.size() * sizeof(double) + sizeof(double) * 3 * (size_t)m.n_faces);
    mb.grid_bytes = (int64_t)(entry.size() * sizeof(uint4) + cells.size() * sizeof(int2));
    dsb::MeshDev &d = mb.dev;
    d.tri = mb.tri;
    d.normal = mb.normal;
    d.tri_idx = mb.tri_idx;
    d.entry = mb.entry;
    d.cell_rng = mb.cell_rng;
    d.xs = mb.xs;
    d.ys = mb.ys;
    d.zs = mb.zs;
    d.len_xs = (int)m.n_sv[0] + 1;
    d.len_ys = (int)m.n_sv[1] + 1;
    d.len_zs = (int)m.n_sv[2] + 1;
    d.nsv1 = (int)m.n_sv[1];
    d.nsv2 = (int)m.n_sv[2];
    d.inv_hx = m.n_sv[0] / (m.xs[m.n_sv[0]] - m.xs[0]);
    d.inv_hy = m.n_sv[1] / (m.ys[m.n_sv[1]] - m.ys[0]);
    d.inv_hz = m.n_sv[2] / (m.zs[m.n_sv[2]] - m.zs[0]);
    const double *grid[3] = {m.xs, m.ys, m.zs};
    for (int k = 0; k < 3; ++k) {
        d.vox[k] = std::fabs(grid[k][m.n_sv[k]] - grid[k][0]);
        d.inv_vox[k] = 1.0 / d.vox[k];
        d.top[k] = grid[k][m.n_sv[k]];
        d.qscale[k] = 32767.0 / d.top[k];
    }
    d.perm_prob = m.perm_prob;
    // the grid the cooperative search walks: the reference grid itself ...
    d.fine.entry = mb.entry;
    d.fine.cell_rng = mb.cell_rng;
    d.fine.xs = mb.xs;
    d.fine.ys = mb.ys;
    d.fine.zs = mb.zs;
    d.fine.len_xs = d.len_xs;
    d.fine.len_ys = d.len_ys;
    d.fine.len_zs = d.len_zs;
    d.fine.nsv1 = d.nsv1;
    d.fine.nsv2 = d.nsv2;
    d.fine.inv_hx = d.inv_hx;
    d.fine.inv_hy = d.inv_hy;
    d.fine.inv_hz = d.inv_hz;
    for (int k = 0; k < 3; ++k) d.fine.margin[k] = 1e-11 * d.vox[k];
    // ... or a refinement of it
    int refine[3];
    choose_refinement(m, step_l, m.n_triangle_indices, refine);
    if (refine[0] * refine[1] * refine[2] > 1) {
        int rc = build_search_grid(m, cells, tri_idx, box, refine, mb);
        if (rc) return rc;
    }
    mb.h_cells = std::move(cells);
    mb.h_tri_idx = std::move(tri_idx);
    mb.h_box = std::move(box);
    for (int k = 0; k < 3; ++k) mb.n_sv[k] = m.n_sv[k];
    return DSB_OK;
}

// Column lists for the sampler's +x rays (dsb::FillColumns): per (y, z) column the distinct
// triangles of its cells, those of the last x cell first.  One pass over the cell lists.
int build_fill_columns(MeshBuffers &mb)
{
    std::lock_guard<std::mutex> lk(mb.mu);
    if (mb.columns.entry) return DSB_OK;
    Range nvtx("dsb: sampler column lists");
    const int64_t n0 = mb.n_sv[0], n1 = mb.n_sv[1], n2 = mb.n_sv[2];
    const int64_t n_cols = n1 * n2;
    std::vector<int> start((size_t)n_cols + 1), cnt((size_t)(n_cols * n0));
    // pieces of consecutive columns, each on its own thread with its own "seen in this column" marks
    const int n_pieces = (int)std::max<int64_t>(1, std::min<int64_t>({8, (int64_t)std::max(1u, std::thread::hardware_concurrency()),
                                                                    (int64_t)mb.h_tri_idx.size() / 262144}));
    std::vector<std::vector<uint4>> piece_entry((size_t)n_pieces);
    std::vector<std::vector<int>> piece_start((size_t)n_pieces);
    auto work = [&](int piece) {
        const int64_t col0 = n_cols * piece / n_pieces, col1 = n_cols * (piece + 1) / n_pieces;
        std::vector<uint4> &out = piece_entry[(size_t)piece];
        std::vector<int> &st = piece_start[(size_t)piece];
        std::vector<int64_t> seen_in(mb.h_box.size(), -1);
        for (int64_t col = col0; col < col1; ++col) {
            const int64_t y = col / n2, z = col % n2;
            const size_t begin = out.size();
            st.push_back((int)begin);
            for (int64_t x = n0 - 1; x >= 0; --x) {
                const int2 c = mb.h_cells[(size_t)((x * n1 + y) * n2 + z)];
                for (int i = c.x; i < c.y; ++i) {
                    const int t = mb.h_tri_idx[(size_t)i];
                    if (seen_in[(size_t)t] != col) {
                        seen_in[(size_t)t] = col;
                        out.push_back(mb.h_box[(size_t)t]);
                    }
                }
                cnt[(size_t)(col * n0 + x)] = (int)(out.size() - begin);
            }
        }
    };
    {
        std::vector<std::thread> pool;
        for (int piece = 1; piece < n_pieces; ++piece) pool.emplace_back(work, piece);
        work(0);
        for (auto &th : pool) th.join();
    }
    size_t total = 0;
    for (auto &v : piece_entry) total += v.size();
    HostBuf<uint4> entry(total + 1);
    if (!entry.ok()) return fail(DSB_ENOMEM, "mesh: no page-locked host memory for the column lists");
    {
        size_t at = 0;
        for (int piece = 0; piece < n_pieces; ++piece) {
            const int64_t col0 = n_cols * piece / n_pieces;
            for (size_t k = 0; k < piece_start[(size_t)piece].size(); ++k)
                start[(size_t)col0 + k] = (int)at + piece_start[(size_t)piece][k];
            std::copy(piece_entry[(size_t)piece].begin(), piece_entry[(size_t)piece].end(), entry.data() + at);
            at += piece_entry[(size_t)piece].size();
        }
    }
    start[(size_t)n_cols] = (int)total;
    entry[total] = make_uint4(0u, 0xffffffffu, 0xffffffffu, 0xffffffffu);
    DSB_CUDA(cache_malloc(&mb.col_start, start.size() * sizeof(int)));
    DSB_CUDA(cache_malloc(&mb.col_cnt, cnt.size() * sizeof(int)));
    DSB_CUDA(cache_malloc(&mb.col_entry, entry.size() * sizeof(uint4)));
    DSB_CUDA(cudaMemcpy(mb.col_start, start.data(), start.size() * sizeof(int), cudaMemcpyHostToDevice));
    DSB_CUDA(cudaMemcpy(mb.col_cnt, cnt.data(), cnt.size() * sizeof(int), cudaMemcpyHostToDevice));
    DSB_CUDA(cudaMemcpy(mb.col_entry, entry.data(), entry.size() * sizeof(uint4), cudaMemcpyHostToDevice));
    DSB_CUDA(cudaDeviceSynchronize());  // (as in upload_mesh)
    mb.columns.start = mb.col_start;
    mb.columns.cnt = mb.col_cnt;
    mb.columns.entry = mb.col_entry;
    mb.columns.n0 = (int)n0;
    return DSB_OK;
}

// ------------------------------------------------------------- uploaded meshes, kept between handles

// simulation() is usually called many times on one substrate (other protocols, other seeds), and each
// call hands the library the same mesh arrays again.  Re-laying them out, uploading them and refining
// the search grid costs 12 ms for the 1e5-triangle mesh of BASELINE config 4 next to a 97 ms walk, 70
// ms for the 1e6-triangle mesh.  The last uploads (four per device) are therefore kept, found again by a
// 64-bit fingerprint of every input array (plus sizes and the refinement the step length asks for),
// and shared by the handles that use them.  DISIMPY_B200_MESH_CACHE=0 turns this off.
uint64_t fingerprint(const void *data, size_t bytes, uint64_t seed)
{
    const size_t n_words = bytes / 8;
    const uint64_t *w = static_cast<const uint64_t *>(data);
    const int64_t n_chunks = (int64_t)std::max<size_t>(1, std::min<size_t>(64, n_words / 32768));
    std::vector<uint64_t> part((size_t)n_chunks, 0);
    parallel_ranges(n_chunks, [&](int64_t c0, int64_t c1) {
        for (int64_t c = c0; c < c1; ++c) {
            uint64_t h0 = seed ^ (uint64_t)c, h1 = ~seed, h2 = seed * 3, h3 = seed + 0x9E3779B97F4A7C15ULL;
            size_t i = n_words * (size_t)c / (size_t)n_chunks;
            const size_t end = n_words * (size_t)(c + 1) / (size_t)n_chunks;
            for (; i + 4 <= end; i += 4) {   // four independent lanes: multiply latency hidden
                h0 = (h0 ^ w[i]) * 0x9E3779B97F4A7C15ULL;
                h1 = (h1 ^ w[i + 1]) * 0xC2B2AE3D27D4EB4FULL;
                h2 = (h2 ^ w[i + 2]) * 0x165667B19E3779F9ULL;
                h3 = (h3 ^ w[i + 3]) * 0xD6E8FEB86659FD93ULL;
                h0 ^= h0 >> 29, h1 ^= h1 >> 31, h2 ^= h2 >> 27, h3 ^= h3 >> 30;
            }
            for (; i < end; ++i) h0 = (h0 ^ w[i]) * 0x9E3779B97F4A7C15ULL, h0 ^= h0 >> 29;
            part[(size_t)c] = (h0 * 31 + h1) ^ ((h2 * 17 + h3) << 1);
        }
    }, 1);
    uint64_t h = seed ^ (uint64_t)bytes;
    for (uint64_t v : part) h = (h ^ v) * 0x9E3779B97F4A7C15ULL, h ^= h >> 32;
    const unsigned char *tail = static_cast<const unsigned char *>(data) + n_words * 8;
    for (size_t i = 0; i < bytes % 8; ++i) h = (h ^ tail[i]) * 0x100000001B3ULL;
    return h;
}

struct MeshKey {
    int device = -1;
    uint64_t hash = 0;
    int64_t n_faces = 0, n_vertices = 0, n_tri = 0, n_sv[3] = {0, 0, 0};
    int refine[3] = {1, 1, 1};
    bool operator==(const MeshKey &o) const
    {
        return device == o.device && hash == o.hash && n_faces == o.n_faces && n_vertices == o.n_vertices && n_tri == o.n_tri &&
               n_sv[0] == o.n_sv[0] && n_sv[1] == o.n_sv[1] && n_sv[2] == o.n_sv[2] && refine[0] == o.refine[0] &&
               refine[1] == o.refine[1] && refine[2] == o.refine[2];
    }
};

std::mutex g_mesh_mu;
std::vector<std::pair<MeshKey, std::shared_ptr<MeshBuffers>>> g_mesh_cache;   // most recently used last
constexpr size_t kMeshCacheEntries = 4;

int shared_mesh(const dsb_mesh &m, int device, double step_l, std::shared_ptr<MeshBuffers> &out)
{
    const char *env = getenv("DISIMPY_B200_MESH_CACHE");
    const bool use_cache = !(env && env[0] == '0') && m.vertices && m.faces && m.xs && m.ys && m.zs && m.subvoxel_indices &&
                           (m.triangle_indices || m.n_triangle_indices == 0) && m.n_faces > 0 && m.n_vertices > 0 &&
                           m.n_sv[0] > 0 && m.n_sv[1] > 0 && m.n_sv[2] > 0 && m.n_sv[0] <= (1 << 20) && m.n_sv[1] <= (1 << 20) &&
                           m.n_sv[2] <= (1 << 20) && m.n_triangle_indices >= 0;
    MeshKey key;
    if (use_cache) {
        Range nvtx("dsb: mesh fingerprint");
        key.device = device;
        key.n_faces = m.n_faces, key.n_vertices = m.n_vertices, key.n_tri = m.n_triangle_indices;
        for (int a = 0; a < 3; ++a) key.n_sv[a] = m.n_sv[a];
        choose_refinement(m, step_l, m.n_triangle_indices, key.refine);
        uint64_t h = fingerprint(m.vertices, sizeof(double) * 3 * (size_t)m.n_vertices, 1);
        h = fingerprint(m.faces, sizeof(int64_t) * 3 * (size_t)m.n_faces, h);
        h = fingerprint(m.xs, sizeof(double) * (size_t)(m.n_sv[0] + 1), h);
        h = fingerprint(m.ys, sizeof(double) * (size_t)(m.n_sv[1] + 1), h);
        h = fingerprint(m.zs, sizeof(double) * (size_t)(m.n_sv[2] + 1), h);
        h = fingerprint(m.subvoxel_indices, sizeof(int64_t) * 2 * (size_t)(m.n_sv[0] * m.n_sv[1] * m.n_sv[2]), h);
        if (m.n_triangle_indices > 0) h = fingerprint(m.triangle_indices, sizeof(int64_t) * (size_t)m.n_triangle_indices, h);
        key.hash = h;
        std::lock_guard<std::mutex> lk(g_mesh_mu);
        for (size_t i = 0; i < g_mesh_cache.size(); ++i)
            if (g_mesh_cache[i].first == key) {
                out = g_mesh_cache[i].second;
                std::rotate(g_mesh_cache.begin() + (long)i, g_mesh_cache.begin() + (long)i + 1, g_mesh_cache.end());
                return DSB_OK;
            }
    }
    std::shared_ptr<MeshBuffers> mb = std::make_shared<MeshBuffers>();
    int rc = upload_mesh(m, *mb, step_l);
    if (rc) return rc;
    if (use_cache) {
        std::lock_guard<std::mutex> lk(g_mesh_mu);
        g_mesh_cache.emplace_back(key, mb);
        // at most kMeshCacheEntries per DEVICE (a device list uploads the same mesh once per device): drop that
        // device's least recently used one
        size_t on_device = 0;
        for (const auto &e : g_mesh_cache) on_device += e.first.device == key.device;
        if (on_device > kMeshCacheEntries)
            for (auto it = g_mesh_cache.begin(); it != g_mesh_cache.end(); ++it)
                if (it->first.device == key.device) {
                    g_mesh_cache.erase(it);
                    break;
                }
    }
    out = mb;
    return DSB_OK;
}

}  // namespace

// ------------------------------------------------------------- the handle

struct dsb_sim {
    dsb_params prm{};
    cudaStream_t stream = nullptr;
    cudaStream_t stream2 = nullptr;  // every other part of a part-by-part run (dsb_run_part)
    cudaEvent_t ev_rewind = nullptr, ev_parts = nullptr;
    int part_parity = 0;
    // sampler threads [fill_t0, fill_t1) of a multi-rank run (dsb_fill_shard_*)
    ulonglong2 *fill_rng = nullptr;
    double *fill_pts = nullptr;
    int *fill_totals = nullptr;
    int64_t fill_t0 = 0, fill_t1 = 0;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    double *d_grad = nullptr, *d_grad_chunked = nullptr, *d_pos = nullptr, *d_phases = nullptr, *d_partials = nullptr, *d_signal = nullptr;
    unsigned long long *d_rng = nullptr, *d_rng0 = nullptr;
    unsigned char *d_exc = nullptr;
    std::shared_ptr<MeshBuffers> mesh = std::make_shared<MeshBuffers>();   // possibly shared with other handles (shared_mesh)
    double perm_prob = 0.0;
    dsb::EllipsoidConsts ell{};   // DSB_ELLIPSOID: semi-axes and what the distance check derives from them
    // mesh walk with the walkers in cell order (resort_interval, sort_walkers); allocated at the first sort
    int *d_order = nullptr, *d_key = nullptr, *d_bins = nullptr;
    dsb::CellBins bins{};
    int64_t sorted_at = -1;   // time step at which d_order was made from the positions (-1: no order since the last rewind)
    bool sort_off = false;    // no room for the sort's buffers: index order from now on
    bool signal_from_phases = false;   // the launch that reached the last step left the signal partials to phases_signal_kernel
    int rank = 0;            // > 0: low-rank protocol, the walk carries `rank` virtual measurements
    double *d_u = nullptr;   // (n_meas, rank) coefficients of the real measurements
    int64_t t_cur = -1;  // -1: positions not set
    int64_t parts_done = 0;  // walkers advanced by dsb_run_part since the last rewind
    int grid = 0;
    double kernel_ms = 0.0;
    int64_t n_launches = 0;
    bool finalized = false;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> pending;
    cudaEvent_t timer0 = nullptr, timer1 = nullptr;
};

namespace {

// dynamic shared memory of walk_kernel<SUB, 0> for the analytic substrates: two gradient tiles and
// one position tile per warp
template <int SUB>
constexpr size_t many_meas_smem()
{
    return SUB == 4 ? 0
                    : sizeof(double) * (2 * dsb::kGradRows * dsb::grad_row_len(dsb::ChunkSteps<SUB>::value) +
                                        (dsb::kBlock / 32) * 3 * dsb::ChunkSteps<SUB>::value * 32);
}

template <int SUB, int MAXC>
void launch_walk_cells(const dsb::KParams &kp, int grid, cudaStream_t st)
{
    switch (kp.n_meas <= dsb::kMaxRegMeas ? kp.n_meas : 0) {
    case 1: dsb::walk_kernel<SUB, 1, MAXC><<<grid, dsb::kBlock, 0, st>>>(kp); break;
    case 2: dsb::walk_kernel<SUB, 2, MAXC><<<grid, dsb::kBlock, 0, st>>>(kp); break;
    case 3: dsb::walk_kernel<SUB, 3, MAXC><<<grid, dsb::kBlock, 0, st>>>(kp); break;
    case 4: dsb::walk_kernel<SUB, 4, MAXC><<<grid, dsb::kBlock, 0, st>>>(kp); break;
    default: {
        constexpr size_t smem = many_meas_smem<SUB>();
        if (smem > 48 * 1024) {  // opt in once per device (the attribute is per device and function)
            static std::mutex mu;
            static bool done[64] = {false};
            int dev = 0;
            cudaGetDevice(&dev);
            std::lock_guard<std::mutex> lk(mu);
            if (dev >= 0 && dev < 64 && !done[dev]) {
                cudaFuncSetAttribute(dsb::walk_kernel<SUB, 0, MAXC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                done[dev] = true;
            }
        }
        dsb::walk_kernel<SUB, 0, MAXC><<<grid, dsb::kBlock, smem, st>>>(kp);
        break;
    }
    }
}

template <int SUB>
void launch_walk(const dsb::KParams &kp, int grid, cudaStream_t st)
{
    if constexpr (SUB == 4) {
        // a step shorter than the smallest grid spacing overlaps at most 2 cells per axis: the
        // kernel variant with 8 cell slots per walker does the same work with shorter loops
        const dsb::SearchGrid &g = kp.mesh.fine;
        const double h_min = 1.0 / std::max(g.inv_hx, std::max(g.inv_hy, g.inv_hz));
        if (kp.step_l < h_min) {
            launch_walk_cells<4, dsb::kMaxCellsShortStep>(kp, grid, st);
            return;
        }
    }
    launch_walk_cells<SUB, dsb::kMaxCells>(kp, grid, st);
}

// Rank-revealing factorisation of the (M x K) gradient matrix, K = 3 n_t: greedy row-pivoted
// Gram-Schmidt, at most max_rank steps.  Succeeds when every row lies in the span of the chosen
// basis up to 1e-13 of the largest row norm; then A = U V with V (rank x K, orthogonal rows
// scaled to that norm, so that they look like gradients) and U (M x rank).
bool factor_low_rank(const double *A, int64_t M, int64_t K, int max_rank, std::vector<double> &U, std::vector<double> &V,
                     int &rank)
{
    std::vector<double> R(A, A + M * K), norm2((size_t)M);
    double top = 0.0;
    for (int64_t m = 0; m < M; ++m) {
        double acc = 0.0;
        for (int64_t k = 0; k < K; ++k) acc += R[m * K + k] * R[m * K + k];
        if (!std::isfinite(acc)) return false;  // NaN / inf samples: the general path keeps the reference's semantics
        norm2[(size_t)m] = acc;
        top = std::max(top, acc);
    }
    if (!(top > 0.0)) return false;
    const double scale = std::sqrt(top), tol2 = 1e-26 * top;
    std::vector<double> Q, C((size_t)(M * max_rank), 0.0);
    rank = 0;
    for (;;) {
        int64_t piv = 0;
        for (int64_t m = 1; m < M; ++m)
            if (norm2[(size_t)m] > norm2[(size_t)piv]) piv = m;
        if (norm2[(size_t)piv] <= tol2) break;
        if (rank == max_rank) return false;
        std::vector<double> q(R.begin() + piv * K, R.begin() + (piv + 1) * K);
        const double inv = 1.0 / std::sqrt(norm2[(size_t)piv]);
        for (auto &x : q) x *= inv;
        for (int pass = 0; pass < 2; ++pass)  // re-orthogonalise against the earlier rows once
            for (int j = 0; j < rank; ++j) {
                double d = 0.0;
                for (int64_t k = 0; k < K; ++k) d += q[(size_t)k] * Q[(size_t)(j * K + k)];
                for (int64_t k = 0; k < K; ++k) q[(size_t)k] -= d * Q[(size_t)(j * K + k)];
            }
        for (int64_t m = 0; m < M; ++m) {
            double d = 0.0, acc = 0.0;
            for (int64_t k = 0; k < K; ++k) d += R[m * K + k] * q[(size_t)k];
            for (int64_t k = 0; k < K; ++k) {
                R[m * K + k] -= d * q[(size_t)k];
                acc += R[m * K + k] * R[m * K + k];
            }
            C[(size_t)(m * max_rank + rank)] = d;
            norm2[(size_t)m] = acc;
        }
        Q.insert(Q.end(), q.begin(), q.end());
        ++rank;
    }
    if (rank == 0) return false;
    U.assign((size_t)(M * rank), 0.0);
    V.assign((size_t)(rank * K), 0.0);
    for (int j = 0; j < rank; ++j)
        for (int64_t k = 0; k < K; ++k) V[(size_t)(j * K + k)] = Q[(size_t)(j * K + k)] * scale;
    // coefficients by projecting the ORIGINAL rows (not the running sums of the elimination)
    for (int64_t m = 0; m < M; ++m)
        for (int j = 0; j < rank; ++j) {
            double d = 0.0;
            for (int64_t k = 0; k < K; ++k) d += A[m * K + k] * Q[(size_t)(j * K + k)];
            U[(size_t)(m * rank + j)] = d / scale;
        }
    // final check on what will actually be used: |A - U V| row by row
    for (int64_t m = 0; m < M; ++m) {
        double acc = 0.0;
        for (int64_t k = 0; k < K; ++k) {
            double v = A[m * K + k];
            for (int j = 0; j < rank; ++j) v -= U[(size_t)(m * rank + j)] * V[(size_t)(j * K + k)];
            acc += v * v;
        }
        if (!(acc <= tol2)) return false;
    }
    return true;
}

int check_params(const dsb_params *p, const double *gradient)
{
    if (!p || !gradient) return fail(DSB_EINVAL, "null params or gradient");
    if (p->substrate < DSB_FREE || p->substrate > DSB_MESH) return fail(DSB_EINVAL, "unknown substrate");
    if (p->n_walkers <= 0) return fail(DSB_EINVAL, "n_walkers must be positive");
    if (p->n_walkers > (int64_t)0x7fffffff * dsb::kBlock) return fail(DSB_EINVAL, "n_walkers too large for one shard");
    if (p->n_meas <= 0 || p->n_meas > 0x3fffffff) return fail(DSB_EINVAL, "n_meas out of range");
    if (p->n_t <= 0 || p->n_t > 0x7fffffff) return fail(DSB_EINVAL, "n_t out of range");
    if (p->walker_offset < 0) return fail(DSB_EINVAL, "walker_offset must be non-negative");
    if (p->max_iter < 1) return fail(DSB_EINVAL, "max_iter must be >= 1");
    if (!(p->step_l > 0) || !(p->dt > 0)) return fail(DSB_EINVAL, "step_l and dt must be positive");
    if ((p->substrate == DSB_SPHERE || p->substrate == DSB_CYLINDER) && !(p->radius > 0))
        return fail(DSB_EINVAL, "radius must be positive");
    return DSB_OK;
}

int sync_stats(dsb_sim *s)
{
    for (auto &pr : s->pending) {
        float ms = 0.f;
        DSB_CUDA(cudaEventSynchronize(pr.second));
        DSB_CUDA(cudaEventElapsedTime(&ms, pr.first, pr.second));
        s->kernel_ms += ms;
        cudaEventDestroy(pr.first);
        cudaEventDestroy(pr.second);
    }
    s->pending.clear();
    return DSB_OK;
}

}  // namespace

extern "C" {

const char *dsb_last_error(void) { return g_err.c_str(); }

int dsb_release_cache(void)
{
    {
        std::lock_guard<std::mutex> mk(g_mesh_mu);
        g_mesh_cache.clear();   // (meshes still used by live handles go when those are destroyed)
    }
    std::lock_guard<std::mutex> lk(g_cache.mu);
    int dev = 0;
    cudaGetDevice(&dev);
    for (auto &kv : g_cache.idle) {
        cudaSetDevice(kv.first.first);
        cudaFree(kv.second);
    }
    g_cache.idle.clear();
    g_cache.idle_bytes = 0;
    cudaSetDevice(dev);
    std::lock_guard<std::mutex> hk(g_host_cache.mu);
    for (auto &kv : g_host_cache.idle) cudaFreeHost(kv.second);
    g_host_cache.idle.clear();
    g_host_cache.idle_bytes = 0;
    return DSB_OK;
}
const char *dsb_version(void) { return "disimpy_b200 0.1 (sm_100a)"; }

int dsb_device_count(int32_t *count)
{
    if (!count) return fail(DSB_EINVAL, "null argument");
    *count = 0;
    int n = 0;
    DSB_CUDA(cudaGetDeviceCount(&n));
    *count = n;
    return DSB_OK;
}

int dsb_destroy(dsb_sim *s)
{
    if (!s) return DSB_OK;
    cudaSetDevice(s->prm.device);
    if (s->stream) cudaStreamSynchronize(s->stream);
    if (s->stream2) cudaStreamSynchronize(s->stream2);
    for (auto &pr : s->pending) {
        cudaEventDestroy(pr.first);
        cudaEventDestroy(pr.second);
    }
    if (s->timer0) cudaEventDestroy(s->timer0);
    if (s->timer1) cudaEventDestroy(s->timer1);
    cache_free(s->d_grad);
    cache_free(s->d_grad_chunked);
    cache_free(s->d_pos);
    cache_free(s->d_phases);
    cache_free(s->d_partials);
    cache_free(s->d_signal);
    cache_free(s->d_u);
    cache_free(s->d_rng);
    cache_free(s->d_rng0);
    cache_free(s->d_exc);
    cache_free(s->d_order);
    cache_free(s->d_key);
    cache_free(s->d_bins);
    cache_free(s->fill_rng);
    cache_free(s->fill_pts);
    cache_free(s->fill_totals);
    s->mesh.reset();
    if (s->stream) cudaStreamDestroy(s->stream);
    if (s->stream2) cudaStreamDestroy(s->stream2);
    if (s->ev_rewind) cudaEventDestroy(s->ev_rewind);
    if (s->ev_parts) cudaEventDestroy(s->ev_parts);
    delete s;
    return DSB_OK;
}

// *live follows the handle while it is being built, so that dsb_create can free it if anything throws
static int create_impl(const dsb_params *params, const double *gradient, dsb_sim **out, dsb_sim **live)
{
    Range nvtx("dsb_create: buffers + gradient upload + RNG init");
    if (!out) return fail(DSB_EINVAL, "null out pointer");
    *out = nullptr;
    int rc = check_params(params, gradient);
    if (rc) return rc;
    int n_dev = 0;
    DSB_CUDA(cudaGetDeviceCount(&n_dev));
    if (params->device < 0 || params->device >= n_dev) return fail(DSB_EINVAL, "no such CUDA device");
    DSB_CUDA(cudaSetDevice(params->device));
    dsb_sim *s = new (std::nothrow) dsb_sim();
    if (!s) return fail(DSB_ENOMEM, "host allocation failed");
    *live = s;
    s->prm = *params;
    const int64_t N = params->n_walkers, M = params->n_meas, T = params->n_t;
    s->grid = (int)((N + dsb::kBlock - 1) / dsb::kBlock);
#define DSB_TRY(expr)                 \
    do {                              \
        cudaError_t e_ = (expr);      \
        if (e_ != cudaSuccess) {      \
            std::string msg = std::string(#expr) + ": " + cudaGetErrorString(e_); \
            int code = e_ == cudaErrorMemoryAllocation ? DSB_ENOMEM : DSB_ECUDA;  \
            cudaGetLastError();       \
            dsb_destroy(s), *live = nullptr;           \
            return fail(code, msg);   \
        }                             \
    } while (0)
    DSB_TRY(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
    DSB_TRY(cudaStreamCreateWithFlags(&s->stream2, cudaStreamNonBlocking));
    DSB_TRY(cudaEventCreateWithFlags(&s->ev_rewind, cudaEventDisableTiming));
    DSB_TRY(cudaEventCreateWithFlags(&s->ev_parts, cudaEventDisableTiming));
    // Protocols whose gradient matrix has rank <= 4 (all PGSE-type ones) walk with that many virtual
    // measurements and are expanded at the end; DISIMPY_B200_LOWRANK=0 keeps the general path.
    std::vector<double> lr_u, lr_v;
    const char *lr_env = getenv("DISIMPY_B200_LOWRANK");
    // (the factorisation works on a copy of the whole gradient: not attempted above 1 GiB of it)
    if (M > dsb::kMaxRegMeas && !(lr_env && lr_env[0] == '0') && M * 3 * T <= (int64_t(1) << 27) &&
        factor_low_rank(gradient, M, 3 * T, (int)std::min<int64_t>(dsb::kMaxRank, M / 2), lr_u, lr_v, s->rank)) {
        gradient = lr_v.data();  // (a rank above M / 2 would not pay for the expansion)
    } else {
        s->rank = 0;
    }
    const int64_t Mw = s->rank > 0 ? s->rank : M;  // measurements the walk kernels see
    DSB_TRY(cache_malloc(&s->d_grad, sizeof(double) * 3 * Mw * T));
    DSB_TRY(cache_malloc(&s->d_pos, sizeof(double) * 3 * N));
    DSB_TRY(cache_malloc(&s->d_phases, sizeof(double) * Mw * N));
    DSB_TRY(cache_malloc(&s->d_partials, sizeof(double) * (M + 1) * s->grid));
    DSB_TRY(cache_malloc(&s->d_signal, sizeof(double) * 2 * (M + 1)));   // this shard's result, then room for the all-reduced one
    DSB_TRY(cache_malloc(&s->d_rng, sizeof(unsigned long long) * 2 * N));
    DSB_TRY(cache_malloc(&s->d_rng0, sizeof(unsigned long long) * 2 * N));
    DSB_TRY(cache_malloc(&s->d_exc, N));
    // uploads are ordered on the handle's own (non-blocking) stream: the kernels that read them run there
    DSB_TRY(cudaMemcpyAsync(s->d_grad, gradient, sizeof(double) * 3 * Mw * T, cudaMemcpyHostToDevice, s->stream));
    if (s->rank > 0) {
        DSB_TRY(cache_malloc(&s->d_u, sizeof(double) * M * s->rank));
        DSB_TRY(cudaMemcpyAsync(s->d_u, lr_u.data(), sizeof(double) * M * s->rank, cudaMemcpyHostToDevice, s->stream));
    }
    if (Mw > dsb::kMaxRegMeas) {
        // chunk-major copy for the many-measurement kernels: (chunk, measurement, step in chunk, xyz),
        // rows padded to kGradRowLen doubles, scaled by gamma * dt (the A operand of the phase GEMM)
        const int64_t C = dsb::chunk_steps(params->substrate), L = dsb::grad_row_len((int)C), n_chunks = (T + C - 1) / C;
        const double gamma_dt = params->dt * 267.513e6;
        std::vector<double> gc((size_t)(n_chunks * Mw * L), 0.0);
        for (int64_t m = 0; m < Mw; ++m)
            for (int64_t t = 0; t < T; ++t)
                for (int c = 0; c < 3; ++c)
                    gc[(size_t)(((t / C) * Mw + m) * L + (t % C) * 3 + c)] = gamma_dt * gradient[(m * T + t) * 3 + c];
        DSB_TRY(cache_malloc(&s->d_grad_chunked, gc.size() * sizeof(double)));
        DSB_TRY(cudaMemcpyAsync(s->d_grad_chunked, gc.data(), gc.size() * sizeof(double), cudaMemcpyHostToDevice, s->stream));
        DSB_TRY(cudaStreamSynchronize(s->stream));  // gc is a local
    }
    DSB_TRY(cudaStreamSynchronize(s->stream));  // the sources (caller's gradient, lr_u, lr_v) are free again
#undef DSB_TRY
    if (params->substrate == DSB_MESH) {
        s->perm_prob = params->mesh.perm_prob;
        rc = shared_mesh(params->mesh, params->device, params->step_l, s->mesh);
        if (rc) {
            std::string keep = g_err;
            dsb_destroy(s), *live = nullptr;
            return fail(rc, keep);
        }
    }
    if (params->substrate == DSB_ELLIPSOID) {
        // reciprocals of the semi-axes etc., computed on the device with the step kernel's own functions
        dsb::EllipsoidConsts *d_ell = nullptr;
        cudaError_t e = cudaMalloc(&d_ell, sizeof(dsb::EllipsoidConsts));
        if (e == cudaSuccess) {
            dsb::ellipsoid_consts_kernel<<<1, 1, 0, s->stream>>>(params->semiaxes[0], params->semiaxes[1], params->semiaxes[2], d_ell);
            e = cudaMemcpyAsync(&s->ell, d_ell, sizeof(dsb::EllipsoidConsts), cudaMemcpyDeviceToHost, s->stream);
        }
        if (e == cudaSuccess) e = cudaStreamSynchronize(s->stream);
        cudaFree(d_ell);
        if (e != cudaSuccess) {
            std::string msg = std::string("ellipsoid constants: ") + cudaGetErrorString(e);
            dsb_destroy(s), *live = nullptr;
            return fail(DSB_ECUDA, msg);
        }
    }
    s->prm.mesh = dsb_mesh{};  // host pointers are not kept
    rc = launch_rng_init(params->device, params->seed, (uint64_t)params->walker_offset, N,
                         reinterpret_cast<ulonglong2 *>(s->d_rng0), s->stream);
    if (rc) {
        std::string keep = g_err;
        dsb_destroy(s), *live = nullptr;
        return fail(rc, keep);
    }
    *out = s;
    *live = nullptr;
    return DSB_OK;
}

int dsb_create(const dsb_params *params, const double *gradient, dsb_sim **out)
{
    dsb_sim *live = nullptr;
    const int rc = guarded([&] { return create_impl(params, gradient, out, &live); });
    if (live) {  // an exception left a half-built handle behind
        std::string keep = g_err;
        dsb_destroy(live);
        g_err = keep;
    }
    return rc;
}

static int rewind_sim(dsb_sim *s)
{
    const int64_t N = s->prm.n_walkers;
    DSB_CUDA(cudaMemcpyAsync(s->d_rng, s->d_rng0, sizeof(unsigned long long) * 2 * N, cudaMemcpyDeviceToDevice, s->stream));
    DSB_CUDA(cudaMemsetAsync(s->d_exc, 0, N, s->stream));
    for (auto &pr : s->pending) {
        cudaEventDestroy(pr.first);
        cudaEventDestroy(pr.second);
    }
    s->pending.clear();
    DSB_CUDA(cudaEventRecord(s->ev_rewind, s->stream));
    s->t_cur = 0;
    s->sorted_at = -1;
    s->parts_done = 0;
    s->part_parity = 0;
    s->kernel_ms = 0.0;
    s->n_launches = 0;
    s->finalized = false;
    return DSB_OK;
}

int dsb_set_positions(dsb_sim *s, const double *positions)
{
    if (!s || !positions) return fail(DSB_EINVAL, "null argument");
    Range nvtx("dsb_set_positions: H2D");
    DSB_CUDA(cudaSetDevice(s->prm.device));
    DSB_CUDA(cudaMemcpyAsync(s->d_pos, positions, sizeof(double) * 3 * s->prm.n_walkers, cudaMemcpyHostToDevice, s->stream));
    return rewind_sim(s);
}

int dsb_set_positions_dev(dsb_sim *s, const double *positions_dev)
{
    if (!s || !positions_dev) return fail(DSB_EINVAL, "null argument");
    DSB_CUDA(cudaSetDevice(s->prm.device));
    DSB_CUDA(cudaMemcpyAsync(s->d_pos, positions_dev, sizeof(double) * 3 * s->prm.n_walkers, cudaMemcpyDeviceToDevice, s->stream));
    return rewind_sim(s);
}

#ifndef DSB_RESORT_MIN_WALKERS
#define DSB_RESORT_MIN_WALKERS 65536
#endif
#ifndef DSB_RESORT_MIN_MESH_BYTES
#define DSB_RESORT_MIN_MESH_BYTES (int64_t(64) << 20)   // half the L2: a smaller mesh stays resident whatever the order
#endif
#ifndef DSB_RESORT_DRIFT
#define DSB_RESORT_DRIFT 0.1   // of the voxel edge: rms drift per axis after which the walkers are sorted again
#endif
// A mesh walk advances its walkers in cell order (sort_walkers): 0 = no (index order), otherwise
// the number of steps after which a run sorts them again.  Measured on a B200
// (profiles/r02_ab_kbench_cell_order.txt): what pays is the L2, not the L1 -- with the walkers of
// the ~600 resident blocks in one slab of the mesh instead of all over it, the 1e6-triangle mesh of
// BASELINE config 5 (178 MB of lists and triangles against 126 MB of L2) walks 17-18 % faster, the
// 1e5-triangle mesh of config 4 (L2-resident anyway) +0.7 % from uniform initial positions and -1.9 %
// from extra-axonal ones (bench.py), hence the size test below; sorting again every 16-128 steps to keep
// the lanes of a WARP in the same cell is slower than not sorting at all (-20 % ... -4 %): every
// launch ends with its lanes draining one by one, ~4 steps' worth of time.  So: once at the start of
// a run, and again only when diffusion has spread a block's walkers over a good part of the voxel.
// Register phases only (the many-measurement kernel needs consecutive walkers for its tiles).
// DISIMPY_B200_RESORT=<steps> overrides the interval, 0 turns the sort off.
static int64_t resort_interval(const dsb_sim *s)
{
    const dsb_params &P = s->prm;
    if (P.substrate != DSB_MESH || s->sort_off) return 0;
    const int meas = s->rank > 0 ? s->rank : (int)P.n_meas;
    if (meas > dsb::kMaxRegMeas || P.n_walkers >= (int64_t(1) << 31)) return 0;
    if (const char *env = getenv("DISIMPY_B200_RESORT")) return std::max<int64_t>(0, atoll(env));
    const dsb::MeshDev &m = s->mesh->dev;
    const int64_t cells = int64_t(m.len_xs - 1) * (m.len_ys - 1) * (m.len_zs - 1);
    if (P.n_walkers < DSB_RESORT_MIN_WALKERS || cells < 4096) return 0;
    if (s->mesh->data_bytes + s->mesh->grid_bytes < DSB_RESORT_MIN_MESH_BYTES) return 0;
    const double edge = std::max(m.vox[0], std::max(m.vox[1], m.vox[2]));   // the slabs are cut across this one
    const double steps = 3.0 * (DSB_RESORT_DRIFT * edge / P.step_l) * (DSB_RESORT_DRIFT * edge / P.step_l);
    return (int64_t)std::min(1e9, std::max(256.0, steps));   // (a NaN lands on 256)
}

// walkers of the handle into cell order (d_order), on stream st
static int sort_walkers(dsb_sim *s, cudaStream_t st)
{
    const int64_t N = s->prm.n_walkers;
    const dsb::MeshDev &m = s->mesh->dev;
    if (!s->d_order || !s->d_key || !s->d_bins) {
        const int len[3] = {m.len_xs - 1, m.len_ys - 1, m.len_zs - 1};
        for (int a = 0; a < 3; ++a) {
            s->bins.n[a] = std::max(1, std::min(len[a], 128));
            s->bins.inv_vox[a] = m.inv_vox[a];
            s->bins.axis[a] = a;
        }
        // consecutive blocks share a slab across the longest edge of the voxel
        std::stable_sort(s->bins.axis, s->bins.axis + 3, [&](int a, int b) { return m.vox[a] > m.vox[b]; });
        if (!s->d_order) DSB_CUDA(cache_malloc(&s->d_order, sizeof(int) * N));
        if (!s->d_key) DSB_CUDA(cache_malloc(&s->d_key, sizeof(int) * N));
        if (!s->d_bins) DSB_CUDA(cache_malloc(&s->d_bins, sizeof(int) * s->bins.n[0] * s->bins.n[1] * s->bins.n[2]));
    }
    const int n_bins = s->bins.n[0] * s->bins.n[1] * s->bins.n[2];
    const unsigned blocks = (unsigned)((N + 255) / 256);
    cudaEvent_t e0, e1;   // (counted into the run's device time like the walk launches)
    DSB_CUDA(cudaEventCreate(&e0));
    DSB_CUDA(cudaEventCreate(&e1));
    DSB_CUDA(cudaEventRecord(e0, st));
    DSB_CUDA(cudaMemsetAsync(s->d_bins, 0, sizeof(int) * n_bins, st));
    dsb::cell_order_count<<<blocks, 256, 0, st>>>(s->d_pos, N, s->bins, s->d_key, s->d_bins);
    dsb::cell_order_scan<<<1, 1024, 0, st>>>(s->d_bins, n_bins);
    dsb::cell_order_scatter<<<blocks, 256, 0, st>>>(s->d_key, N, s->d_bins, s->d_order);
    DSB_CUDA(cudaGetLastError());
    DSB_CUDA(cudaEventRecord(e1, st));
    s->pending.emplace_back(e0, e1);
    s->n_launches += 3;
    return DSB_OK;
}

// one launch of the walk kernel: walkers [w0, w1) over time steps [t0, t1); `sorted`: all walkers
// of the handle, assigned to threads in the order of d_order
static int launch_walk_range(dsb_sim *s, cudaStream_t st, int64_t w0, int64_t w1, int64_t t0, int64_t t1, bool sorted = false)
{
    const dsb_params &P = s->prm;
    cudaEvent_t e0, e1;
    DSB_CUDA(cudaEventCreate(&e0));
    DSB_CUDA(cudaEventCreate(&e1));
    DSB_CUDA(cudaEventRecord(e0, st));
    dsb::KParams kp{};
    kp.n_walkers = P.n_walkers;
    kp.w_begin = w0;
    kp.w_end = w1;
    kp.n_blocks_total = s->grid;
    kp.n_meas = s->rank > 0 ? s->rank : (int)P.n_meas;
    kp.n_t = (int)P.n_t;
    kp.t0 = (int)t0;
    kp.t1 = (int)t1;
    kp.finalize = t1 == P.n_t && s->rank == 0 && !sorted;  // low-rank protocols are reduced by lowrank_signal_kernel
    if (t1 == P.n_t) s->signal_from_phases = sorted && s->rank == 0;
    kp.max_iter = (int)std::min<int64_t>(P.max_iter, 0x7fffffff);
    kp.step_l = P.step_l;
    kp.gamma_dt = P.dt * 267.513e6;  // dt * GAMMA (gradients.py:13), one rounding like the reference
    kp.eps = P.epsilon;
    kp.radius = P.radius;
    memcpy(kp.R, P.R, sizeof kp.R);
    memcpy(kp.Rinv, P.R_inv, sizeof kp.Rinv);
    kp.ell = s->ell;
    kp.grad = s->d_grad;
    kp.grad_chunked = s->d_grad_chunked;
    kp.pos = s->d_pos;
    kp.rng = s->d_rng;
    kp.phases = s->d_phases;
    kp.iter_exc = s->d_exc;
    kp.partials = s->d_partials;
    kp.mesh = s->mesh->dev;
    kp.mesh.perm_prob = s->perm_prob;   // (the upload may be shared with handles of another permeability)
    const int grid = (int)((w1 - w0 + dsb::kBlock - 1) / dsb::kBlock);
    if (sorted) kp.order = s->d_order;
    switch (P.substrate) {
    case DSB_FREE: launch_walk<0>(kp, grid, st); break;
    case DSB_SPHERE: launch_walk<1>(kp, grid, st); break;
    case DSB_CYLINDER: launch_walk<2>(kp, grid, st); break;
    case DSB_ELLIPSOID: launch_walk<3>(kp, grid, st); break;
    default: launch_walk<4>(kp, grid, st); break;
    }
    DSB_CUDA(cudaGetLastError());
    s->n_launches += 1;
    DSB_CUDA(cudaEventRecord(e1, st));
    s->pending.emplace_back(e0, e1);
    return DSB_OK;
}

static int launch_signal_reduction(dsb_sim *s)
{
    cudaEvent_t e0, e1;
    DSB_CUDA(cudaEventCreate(&e0));
    DSB_CUDA(cudaEventCreate(&e1));
    DSB_CUDA(cudaEventRecord(e0, s->stream));
    if (s->rank > 0) {
        dsb::KParams kp{};
        kp.n_walkers = s->prm.n_walkers;
        kp.n_blocks_total = s->grid;
        kp.phases = s->d_phases;
        kp.iter_exc = s->d_exc;
        kp.partials = s->d_partials;
        dsb::lowrank_signal_kernel<<<s->grid, dsb::kBlock, 0, s->stream>>>(kp, s->d_u, s->rank, (int)s->prm.n_meas);
        DSB_CUDA(cudaGetLastError());
        s->n_launches += 1;
    } else if (s->signal_from_phases) {
        dsb::KParams kp{};
        kp.n_walkers = s->prm.n_walkers;
        kp.n_blocks_total = s->grid;
        kp.n_meas = (int)s->prm.n_meas;
        kp.phases = s->d_phases;
        kp.iter_exc = s->d_exc;
        kp.partials = s->d_partials;
        switch (kp.n_meas) {
        case 1: dsb::phases_signal_kernel<1><<<s->grid, dsb::kBlock, 0, s->stream>>>(kp); break;
        case 2: dsb::phases_signal_kernel<2><<<s->grid, dsb::kBlock, 0, s->stream>>>(kp); break;
        case 3: dsb::phases_signal_kernel<3><<<s->grid, dsb::kBlock, 0, s->stream>>>(kp); break;
        default: dsb::phases_signal_kernel<4><<<s->grid, dsb::kBlock, 0, s->stream>>>(kp); break;
        }
        DSB_CUDA(cudaGetLastError());
        s->n_launches += 1;
    }
    dsb::reduce_partials_kernel<<<(unsigned)(s->prm.n_meas + 1), 256, 0, s->stream>>>(s->d_partials, s->grid, s->d_signal);
    DSB_CUDA(cudaGetLastError());
    DSB_CUDA(cudaEventRecord(e1, s->stream));
    s->pending.emplace_back(e0, e1);
    s->n_launches += 1;
    s->finalized = true;
    return DSB_OK;
}

int dsb_run(dsb_sim *s, int64_t t0, int64_t t1)
{
    if (!s) return fail(DSB_EINVAL, "null handle");
    if (s->t_cur < 0) return fail(DSB_ESTATE, "dsb_run before dsb_set_positions");
    if (s->parts_done > 0) return fail(DSB_ESTATE, "dsb_run after dsb_run_part: finish the run part by part");
    if (t0 != s->t_cur || t1 <= t0 || t1 > s->prm.n_t) return fail(DSB_EINVAL, "bad step range");
    Range nvtx("dsb_run: walk (+ signal reduction)");
    DSB_CUDA(cudaSetDevice(s->prm.device));
    int rc = DSB_OK;
    int64_t a = t0;
    const int64_t resort = resort_interval(s);
    // The order made at one call serves the following ones (it is a permutation of the walkers whatever
    // happened to them since); calls of a few steps that find none -- a trajectory written step by
    // step -- walk in index order.
    if (resort > 0 && (s->sorted_at >= 0 || t1 - t0 >= std::min<int64_t>(resort, 8))) {
        while (a < t1 && !rc) {
            if (s->sorted_at < 0 || a - s->sorted_at >= resort) {
                rc = sort_walkers(s, s->stream);
                if (rc == DSB_ENOMEM) {   // the cell order is an optimisation: without room for it, index order
                    cudaGetLastError();
                    s->sort_off = true;
                    s->sorted_at = -1;
                    rc = DSB_OK;
                    break;
                }
                if (rc) break;
                s->sorted_at = a;
            }
            const int64_t b = std::min<int64_t>(t1, s->sorted_at + resort);
            rc = launch_walk_range(s, s->stream, 0, s->prm.n_walkers, a, b, true);
            a = b;
        }
    }
    if (!rc && a < t1) rc = launch_walk_range(s, s->stream, 0, s->prm.n_walkers, a, t1);
    if (rc) return rc;
    if (t1 == s->prm.n_t) {
        rc = launch_signal_reduction(s);
        if (rc) return rc;
    }
    s->t_cur = t1;
    return DSB_OK;
}

int dsb_rewind(dsb_sim *s)
{
    if (!s) return fail(DSB_EINVAL, "null handle");
    DSB_CUDA(cudaSetDevice(s->prm.device));
    return rewind_sim(s);
}

int dsb_set_rng_part(dsb_sim *s, int64_t w0, int64_t w1, int64_t global_offset)
{
    if (!s) return fail(DSB_EINVAL, "null handle");
    if (w0 < 0 || w1 <= w0 || w1 > s->prm.n_walkers || global_offset < 0) return fail(DSB_EINVAL, "bad walker range");
    DSB_CUDA(cudaSetDevice(s->prm.device));
    return launch_rng_init(s->prm.device, s->prm.seed, (uint64_t)global_offset, w1 - w0,
                           reinterpret_cast<ulonglong2 *>(s->d_rng0) + w0, s->stream);
}

int dsb_set_positions_part(dsb_sim *s, int64_t w0, int64_t w1, const double *positions)
{
    if (!s || !positions) return fail(DSB_EINVAL, "null argument");
    if (w0 < 0 || w1 <= w0 || w1 > s->prm.n_walkers) return fail(DSB_EINVAL, "bad walker range");
    DSB_CUDA(cudaSetDevice(s->prm.device));
    // consecutive parts alternate between two streams, so that the next part's blocks fill the
    // SMs while the previous part drains
    cudaStream_t st = s->part_parity ? s->stream2 : s->stream;
    if (s->part_parity) DSB_CUDA(cudaStreamWaitEvent(st, s->ev_rewind, 0));
    DSB_CUDA(cudaMemcpyAsync(s->d_pos + 3 * w0, positions, sizeof(double) * 3 * (w1 - w0), cudaMemcpyHostToDevice, st));
    return DSB_OK;
}

int dsb_run_part(dsb_sim *s, int64_t w0, int64_t w1)
{
    if (!s) return fail(DSB_EINVAL, "null handle");
    if (s->t_cur != 0) return fail(DSB_ESTATE, "dsb_run_part needs a rewound handle (dsb_rewind)");
    if (w0 < 0 || w1 <= w0 || w1 > s->prm.n_walkers || w0 % dsb::kBlock != 0)
        return fail(DSB_EINVAL, "bad walker range (w0 must be a multiple of 128)");
    Range nvtx("dsb_run_part: walk");
    DSB_CUDA(cudaSetDevice(s->prm.device));
    cudaStream_t st = s->part_parity ? s->stream2 : s->stream;
    if (s->part_parity) DSB_CUDA(cudaStreamWaitEvent(st, s->ev_rewind, 0));
    int rc = launch_walk_range(s, st, w0, w1, 0, s->prm.n_t);
    if (rc) return rc;
    s->part_parity ^= 1;
    s->parts_done += w1 - w0;
    return DSB_OK;
}

int dsb_finish(dsb_sim *s)
{
    if (!s) return fail(DSB_EINVAL, "null handle");
    if (s->t_cur != 0 || s->parts_done != s->prm.n_walkers)
        return fail(DSB_ESTATE, "dsb_finish: the parts run so far do not cover every walker exactly once");
    DSB_CUDA(cudaSetDevice(s->prm.device));
    DSB_CUDA(cudaEventRecord(s->ev_parts, s->stream2));
    DSB_CUDA(cudaStreamWaitEvent(s->stream, s->ev_parts, 0));
    int rc = launch_signal_reduction(s);
    if (rc) return rc;
    s->t_cur = s->prm.n_t;
    s->parts_done = 0;
    return DSB_OK;
}

int dsb_sync(dsb_sim *s)
{
    if (!s) return fail(DSB_EINVAL, "null handle");
    DSB_CUDA(cudaSetDevice(s->prm.device));
    DSB_CUDA(cudaStreamSynchronize(s->stream));
    DSB_CUDA(cudaStreamSynchronize(s->stream2));
    return DSB_OK;
}

int dsb_get_signal(dsb_sim *s, double *signal, int64_t *n_valid)
{
    return guarded([&]() -> int {
        if (!s || !signal) return fail(DSB_EINVAL, "null argument");
        if (!s->finalized) return fail(DSB_ESTATE, "signal is available after the last time step only");
        Range nvtx("dsb_get_signal: wait + D2H");
        DSB_CUDA(cudaSetDevice(s->prm.device));
        std::vector<double> h((size_t)s->prm.n_meas + 1);
        DSB_CUDA(cudaMemcpyAsync(h.data(), s->d_signal, h.size() * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
        DSB_CUDA(cudaStreamSynchronize(s->stream));
        memcpy(signal, h.data(), sizeof(double) * (size_t)s->prm.n_meas);
        if (n_valid) *n_valid = (int64_t)llround(h.back());
        return DSB_OK;
    });
}

#define DSB_GETTER(name, type, member, count)                                                             \
    int name(dsb_sim *s, type *dst)                                                                       \
    {                                                                                                     \
        if (!s || !dst) return fail(DSB_EINVAL, "null argument");                                         \
        if (s->t_cur < 0) return fail(DSB_ESTATE, "no walker state yet");                                 \
        DSB_CUDA(cudaSetDevice(s->prm.device));                                                           \
        DSB_CUDA(cudaMemcpyAsync(dst, s->member, sizeof(type) * (size_t)(count), cudaMemcpyDeviceToHost, s->stream)); \
        DSB_CUDA(cudaStreamSynchronize(s->stream));                                                       \
        return DSB_OK;                                                                                    \
    }

DSB_GETTER(dsb_get_positions, double, d_pos, 3 * s->prm.n_walkers)
int dsb_get_phases(dsb_sim *s, double *dst)
{
    if (!s || !dst) return fail(DSB_EINVAL, "null argument");
    if (s->t_cur < 0) return fail(DSB_ESTATE, "no walker state yet");
    DSB_CUDA(cudaSetDevice(s->prm.device));
    const int64_t N = s->prm.n_walkers, M = s->prm.n_meas;
    const double *src = s->d_phases;
    double *d_full = nullptr;
    if (s->rank > 0) {  // expand the virtual measurements' phases to the real ones
        DSB_CUDA(cache_malloc(&d_full, sizeof(double) * (size_t)(M * N)));
        dsb::lowrank_expand_kernel<<<(unsigned)((N + 255) / 256), 256, 0, s->stream>>>(s->d_phases, s->d_u, s->rank, (int)M,
                                                                                         (long long)N, d_full);
        src = d_full;
    }
    cudaError_t e = cudaMemcpyAsync(dst, src, sizeof(double) * (size_t)(M * N), cudaMemcpyDeviceToHost, s->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s->stream);
    cache_free(d_full);
    if (e != cudaSuccess) return fail(DSB_ECUDA, cudaGetErrorString(e));
    return DSB_OK;
}
DSB_GETTER(dsb_get_iter_exc, uint8_t, d_exc, s->prm.n_walkers)
DSB_GETTER(dsb_get_rng_states, uint64_t, d_rng, 2 * s->prm.n_walkers)

int dsb_set_rng_states(dsb_sim *s, const uint64_t *states)
{
    if (!s || !states) return fail(DSB_EINVAL, "null argument");
    if (s->t_cur != 0 || s->parts_done != 0) return fail(DSB_ESTATE, "generator states can be replaced at t = 0 only");
    DSB_CUDA(cudaSetDevice(s->prm.device));
    DSB_CUDA(cudaMemcpyAsync(s->d_rng, states, sizeof(uint64_t) * 2 * s->prm.n_walkers, cudaMemcpyHostToDevice, s->stream));
    DSB_CUDA(cudaStreamSynchronize(s->stream));
    return DSB_OK;
}

int dsb_get_run_stats(dsb_sim *s, double *kernel_ms, int64_t *n_launches)
{
    if (!s) return fail(DSB_EINVAL, "null handle");
    DSB_CUDA(cudaSetDevice(s->prm.device));
    int rc = sync_stats(s);
    if (rc) return rc;
    if (kernel_ms) *kernel_ms = s->kernel_ms;
    if (n_launches) *n_launches = s->n_launches;
    return DSB_OK;
}

int dsb_timer_start(dsb_sim *s)
{
    if (!s) return fail(DSB_EINVAL, "null handle");
    DSB_CUDA(cudaSetDevice(s->prm.device));
    if (!s->timer0) {
        DSB_CUDA(cudaEventCreate(&s->timer0));
        DSB_CUDA(cudaEventCreate(&s->timer1));
    }
    DSB_CUDA(cudaEventRecord(s->timer0, s->stream));
    return DSB_OK;
}

int dsb_timer_stop(dsb_sim *s, double *elapsed_ms)
{
    if (!s || !elapsed_ms) return fail(DSB_EINVAL, "null argument");
    if (!s->timer0) return fail(DSB_ESTATE, "dsb_timer_stop before dsb_timer_start");
    DSB_CUDA(cudaSetDevice(s->prm.device));
    DSB_CUDA(cudaEventRecord(s->timer1, s->stream));
    DSB_CUDA(cudaEventSynchronize(s->timer1));
    float ms = 0.f;
    DSB_CUDA(cudaEventElapsedTime(&ms, s->timer0, s->timer1));
    *elapsed_ms = ms;
    return DSB_OK;
}

int dsb_measure_fp64_peak(int32_t device, double *dfma_per_second)
{
    if (!dfma_per_second) return fail(DSB_EINVAL, "null argument");
    DSB_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    DSB_CUDA(cudaGetDeviceProperties(&prop, device));
    const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 4096;
    double *d_out = nullptr;
    DSB_CUDA(cudaMalloc(&d_out, sizeof(double) * (size_t)blocks * threads));
    cudaEvent_t e0, e1;
    DSB_CUDA(cudaEventCreate(&e0));
    DSB_CUDA(cudaEventCreate(&e1));
    double best = 0.0;
    for (int rep = 0; rep < 5; ++rep) {
        DSB_CUDA(cudaEventRecord(e0));
        dsb::fp64_peak_kernel<<<blocks, threads>>>(d_out, iters, 1.0 + 1e-9 * rep);
        DSB_CUDA(cudaEventRecord(e1));
        DSB_CUDA(cudaEventSynchronize(e1));
        float ms = 0.f;
        DSB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        double rate = (double)blocks * threads * iters * dsb::kPeakChains / (ms * 1e-3);
        if (rep > 0 && rate > best) best = rate;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d_out);
    DSB_CUDA(cudaGetLastError());
    *dfma_per_second = best;
    return DSB_OK;
}

int dsb_measure_l2_peak(int32_t device, double *bytes_per_second)
{
    if (!bytes_per_second) return fail(DSB_EINVAL, "null argument");
    DSB_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    DSB_CUDA(cudaGetDeviceProperties(&prop, device));
    // a third of the L2: resident whatever the hashing over the two dies does
    const size_t bytes = std::max<size_t>((size_t)prop.l2CacheSize / 3, size_t(8) << 20) / 16 * 16;
    uint4 *d_buf = nullptr;
    unsigned *d_out = nullptr;
    DSB_CUDA(cudaMalloc(&d_buf, bytes));
    DSB_CUDA(cudaMalloc(&d_out, sizeof(unsigned)));
    DSB_CUDA(cudaMemset(d_buf, 1, bytes));
    cudaEvent_t e0, e1;
    DSB_CUDA(cudaEventCreate(&e0));
    DSB_CUDA(cudaEventCreate(&e1));
    const int blocks = prop.multiProcessorCount * 8, passes = 40;
    double best = 0.0;
    for (int rep = 0; rep < 4; ++rep) {
        DSB_CUDA(cudaEventRecord(e0));
        dsb::l2_peak_kernel<<<blocks, 256>>>(d_buf, (long long)(bytes / 16), passes, d_out);
        DSB_CUDA(cudaEventRecord(e1));
        DSB_CUDA(cudaEventSynchronize(e1));
        float ms = 0.f;
        DSB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        const double rate = (double)bytes * passes / (ms * 1e-3);
        if (rep > 0 && rate > best) best = rate;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d_buf);
    cudaFree(d_out);
    DSB_CUDA(cudaGetLastError());
    *bytes_per_second = best;
    return DSB_OK;
}

int dsb_selftest_sqrt(int32_t device, uint64_t seed, int32_t exp_lo, int32_t exp_hi, int64_t n, int64_t *n_mismatch,
                      double *first_mismatch)
{
    if (!n_mismatch || exp_lo < -969 || exp_hi > 1022 || exp_hi < exp_lo || n <= 0) return fail(DSB_EINVAL, "bad arguments");
    DSB_CUDA(cudaSetDevice(device));
    unsigned long long *d_bad = nullptr;
    double *d_first = nullptr;
    DSB_CUDA(cudaMalloc(&d_bad, sizeof(unsigned long long)));
    DSB_CUDA(cudaMalloc(&d_first, sizeof(double)));
    DSB_CUDA(cudaMemset(d_bad, 0, sizeof(unsigned long long)));
    DSB_CUDA(cudaMemset(d_first, 0, sizeof(double)));
    const int per_thread = 4096, threads = 256;
    const int64_t blocks = std::max<int64_t>(1, (n + (int64_t)per_thread * threads - 1) / ((int64_t)per_thread * threads));
    dsb::sqrt_selftest_kernel<<<(unsigned)std::min<int64_t>(blocks, 0x7fffffff), threads>>>(seed, exp_lo, exp_hi, per_thread, d_bad,
                                                                                           d_first);
    unsigned long long bad = 0;
    double first = 0.0;
    cudaError_t e = cudaMemcpy(&bad, d_bad, sizeof bad, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(&first, d_first, sizeof first, cudaMemcpyDeviceToHost);
    cudaFree(d_bad);
    cudaFree(d_first);
    if (e != cudaSuccess) return fail(DSB_ECUDA, cudaGetErrorString(e));
    *n_mismatch = (int64_t)bad;
    if (first_mismatch) *first_mismatch = first;
    return DSB_OK;
}

int dsb_selftest_device_function(int32_t device, int32_t op, int64_t n, const double *in, double *out)
{
    if (op < 0 || op >= dsb::kUnitOps || n <= 0 || !in || !out) return fail(DSB_EINVAL, "bad arguments");
    DSB_CUDA(cudaSetDevice(device));
    const size_t bytes_in = sizeof(double) * (size_t)n * dsb::unit_n_in(op), bytes_out = sizeof(double) * (size_t)n * dsb::unit_n_out(op);
    double *d_in = nullptr, *d_out = nullptr;
    DSB_CUDA(cudaMalloc(&d_in, bytes_in));
    cudaError_t e = cudaMalloc(&d_out, bytes_out);
    if (e == cudaSuccess) e = cudaMemcpy(d_in, in, bytes_in, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        dsb::device_function_kernel<<<(unsigned)((n + 127) / 128), 128>>>(op, n, d_in, d_out);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpy(out, d_out, bytes_out, cudaMemcpyDeviceToHost);
    cudaFree(d_in);
    cudaFree(d_out);
    if (e != cudaSuccess) return fail(e == cudaErrorMemoryAllocation ? DSB_ENOMEM : DSB_ECUDA, cudaGetErrorString(e));
    return DSB_OK;
}

int dsb_protocol_rank(dsb_sim *s) { return s ? s->rank : 0; }

int dsb_protocol_factor(const double *gradient, int64_t n_meas, int64_t n_t, int32_t max_rank, int32_t *rank, double *u,
                        double *v)
{
    return guarded([&]() -> int {
        if (!gradient || !rank || !u || !v || n_meas <= 0 || n_t <= 0 || max_rank <= 0) return fail(DSB_EINVAL, "bad arguments");
        std::vector<double> U, V;
        int r = 0;
        if (!factor_low_rank(gradient, n_meas, 3 * n_t, max_rank, U, V, r)) {
            *rank = 0;
            return DSB_OK;
        }
        *rank = r;
        std::copy(U.begin(), U.end(), u);
        std::copy(V.begin(), V.end(), v);
        return DSB_OK;
    });
}
void *dsb_stream(dsb_sim *s) { return s ? (void *)s->stream : nullptr; }
double *dsb_signal_dev(dsb_sim *s) { return s ? s->d_signal : nullptr; }

int dsb_simulate(const dsb_params *params, const double *gradient, const double *positions_in,
                 double *signal_out, int64_t *n_valid_out, double *positions_out, double *phases_out,
                 uint8_t *iter_exc_out)
{
    if (!positions_in || !signal_out) return fail(DSB_EINVAL, "null argument");
    dsb_sim *s = nullptr;
    int rc = dsb_create(params, gradient, &s);
    if (rc) return rc;
    rc = dsb_set_positions(s, positions_in);
    if (!rc) rc = dsb_run(s, 0, params->n_t);
    if (!rc) rc = dsb_get_signal(s, signal_out, n_valid_out);
    if (!rc && positions_out) rc = dsb_get_positions(s, positions_out);
    if (!rc && phases_out) rc = dsb_get_phases(s, phases_out);
    if (!rc && iter_exc_out) rc = dsb_get_iter_exc(s, iter_exc_out);
    std::string keep = g_err;
    dsb_destroy(s);
    if (rc) g_err = keep;
    return rc;
}

int dsb_rng_states(int32_t device, uint64_t seed, uint64_t subsequence_start, int64_t n, uint64_t *states)
{
    if (n < 0 || (n > 0 && !states)) return fail(DSB_EINVAL, "bad arguments");
    if (n == 0) return DSB_OK;
    DSB_CUDA(cudaSetDevice(device));
    ulonglong2 *d = nullptr;
    DSB_CUDA(cudaMalloc(&d, sizeof(ulonglong2) * (size_t)n));
    int rc = launch_rng_init(device, seed, subsequence_start, n, d, 0);
    if (!rc) {
        cudaError_t e = cudaMemcpy(states, d, sizeof(ulonglong2) * (size_t)n, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) rc = fail(DSB_ECUDA, cudaGetErrorString(e));
    }
    cudaFree(d);
    return rc;
}

int dsb_fill_mesh_sim(dsb_sim *s, const double *voxel_size, int intra, uint64_t seed, int64_t n_points, int64_t first,
                      int64_t cuda_bs)
{
    if (!s || !voxel_size || n_points <= 0 || cuda_bs <= 0 || first < 0) return fail(DSB_EINVAL, "bad arguments");
    if (s->prm.substrate != DSB_MESH) return fail(DSB_ESTATE, "dsb_fill_mesh_sim needs a mesh handle");
    if (first + s->prm.n_walkers > n_points || n_points > 0x7fffffffLL) return fail(DSB_EINVAL, "bad point range");
    Range nvtx("dsb_fill_mesh_sim: initial positions on the GPU");
    DSB_CUDA(cudaSetDevice(s->prm.device));
    const int64_t n_states = (n_points + cuda_bs - 1) / cuda_bs * cuda_bs;
    const int n_blocks = (int)((n_points + dsb::kCompactBlock - 1) / dsb::kCompactBlock);
    ulonglong2 *d_rng = nullptr;
    double *d_pts = nullptr;
    int *d_totals = nullptr;
    cudaError_t e = cache_malloc(&d_rng, sizeof(ulonglong2) * (size_t)n_states);
    if (e == cudaSuccess) e = cache_malloc(&d_pts, sizeof(double) * 3 * (size_t)n_points);
    if (e == cudaSuccess) e = cache_malloc(&d_totals, sizeof(int) * (size_t)(n_blocks + 1));
    int rc = e == cudaSuccess ? DSB_OK : fail(DSB_ENOMEM, cudaGetErrorString(e));
    if (!rc) rc = build_fill_columns(*s->mesh);
    if (!rc) rc = launch_rng_init(s->prm.device, seed, 0, n_states, d_rng, s->stream);
    int64_t have = 0, proposed = 0;
    // One round per iteration like the reference's host loop (simulations.py:554-579): every thread
    // proposes a point, the accepted ones are appended in thread order until there are n_points.
    // Only the threads up to the one that supplies the last point matter, so from the second round
    // on a round is run as a prefix sized from the acceptance rate so far, and continued in thread
    // order if that was not enough.
    for (int round = 0; !rc && have < n_points; ++round) {
        if (round > 100000) {
            rc = fail(DSB_ESTATE, "fill_mesh: no acceptable points (is the surface closed?)");
            break;
        }
        for (int64_t c0 = 0; !rc && c0 < n_points && have < n_points;) {
            int64_t want = n_points - c0;
            if (have > 0) {
                const double rate = (double)have / (double)proposed;
                const double need = (double)(n_points - have) / rate * 1.02 + 8192.0;
                if (need < (double)want) want = ((int64_t)need + dsb::kCompactBlock - 1) / dsb::kCompactBlock * dsb::kCompactBlock;
                want = std::min(want, n_points - c0);
            }
            const int chunk_blocks = (int)((want + dsb::kCompactBlock - 1) / dsb::kCompactBlock);
            double *pts = d_pts + 3 * c0;
            dsb::fill_mesh_kernel<<<(unsigned)((want + 127) / 128), 128, 0, s->stream>>>(
                s->mesh->dev, s->mesh->columns, voxel_size[0], voxel_size[1], voxel_size[2], intra, (long long)want,
                d_rng + c0, pts);
            dsb::fill_count_kernel<<<chunk_blocks, dsb::kCompactBlock, 0, s->stream>>>(pts, (long long)want, d_totals);
            dsb::fill_scan_kernel<<<1, 1024, 0, s->stream>>>(d_totals, chunk_blocks);
            dsb::fill_scatter_kernel<<<chunk_blocks, dsb::kCompactBlock, 0, s->stream>>>(
                pts, (long long)want, d_totals, (long long)have, (long long)first, (long long)s->prm.n_walkers, s->d_pos);
            int accepted = 0;
            e = cudaMemcpyAsync(&accepted, d_totals + chunk_blocks, sizeof(int), cudaMemcpyDeviceToHost, s->stream);
            if (e == cudaSuccess) e = cudaStreamSynchronize(s->stream);
            if (e == cudaSuccess) e = cudaGetLastError();
            if (e != cudaSuccess) {
                rc = fail(DSB_ECUDA, cudaGetErrorString(e));
                break;
            }
            have += accepted;
            proposed += want;
            c0 += want;
        }
    }
    cache_free(d_rng);
    cache_free(d_pts);
    cache_free(d_totals);
    if (rc) return rc;
    return rewind_sim(s);
}

int dsb_fill_shard_end(dsb_sim *s)
{
    if (!s) return fail(DSB_EINVAL, "null argument");
    cudaSetDevice(s->prm.device);
    cache_free(s->fill_rng);
    cache_free(s->fill_pts);
    cache_free(s->fill_totals);
    s->fill_rng = nullptr;
    s->fill_pts = nullptr;
    s->fill_totals = nullptr;
    s->fill_t0 = s->fill_t1 = 0;
    return DSB_OK;
}

int dsb_fill_shard_begin(dsb_sim *s, uint64_t seed, int64_t thread_begin, int64_t thread_end)
{
    if (!s || thread_begin < 0 || thread_end <= thread_begin || thread_end > 0x7fffffffLL)
        return fail(DSB_EINVAL, "bad arguments");
    if (s->prm.substrate != DSB_MESH) return fail(DSB_ESTATE, "dsb_fill_shard_begin needs a mesh handle");
    DSB_CUDA(cudaSetDevice(s->prm.device));
    dsb_fill_shard_end(s);
    const int64_t n = thread_end - thread_begin;
    const int n_blocks = (int)((n + dsb::kCompactBlock - 1) / dsb::kCompactBlock);
    cudaError_t e = cache_malloc(&s->fill_rng, sizeof(ulonglong2) * (size_t)n);
    if (e == cudaSuccess) e = cache_malloc(&s->fill_pts, sizeof(double) * 3 * (size_t)n);
    if (e == cudaSuccess) e = cache_malloc(&s->fill_totals, sizeof(int) * (size_t)(n_blocks + 1));
    int rc = e == cudaSuccess ? DSB_OK : fail(DSB_ENOMEM, cudaGetErrorString(e));
    if (!rc) rc = build_fill_columns(*s->mesh);
    if (!rc) rc = launch_rng_init(s->prm.device, seed, (uint64_t)thread_begin, n, s->fill_rng, s->stream);
    if (rc) {
        std::string keep = g_err;
        dsb_fill_shard_end(s);
        return fail(rc, keep);
    }
    s->fill_t0 = thread_begin;
    s->fill_t1 = thread_end;
    return DSB_OK;
}

int dsb_fill_shard_round(dsb_sim *s, const double *voxel_size, int intra, double *accepted_dev, int64_t *n_accepted)
{
    if (!s || !voxel_size || !accepted_dev || !n_accepted) return fail(DSB_EINVAL, "null argument");
    if (!s->fill_rng) return fail(DSB_ESTATE, "dsb_fill_shard_round before dsb_fill_shard_begin");
    DSB_CUDA(cudaSetDevice(s->prm.device));
    const int64_t n = s->fill_t1 - s->fill_t0;
    const int n_blocks = (int)((n + dsb::kCompactBlock - 1) / dsb::kCompactBlock);
    dsb::fill_mesh_kernel<<<(unsigned)((n + 127) / 128), 128, 0, s->stream>>>(
        s->mesh->dev, s->mesh->columns, voxel_size[0], voxel_size[1], voxel_size[2], intra, (long long)n, s->fill_rng,
        s->fill_pts);
    dsb::fill_count_kernel<<<n_blocks, dsb::kCompactBlock, 0, s->stream>>>(s->fill_pts, (long long)n, s->fill_totals);
    dsb::fill_scan_kernel<<<1, 1024, 0, s->stream>>>(s->fill_totals, n_blocks);
    dsb::fill_scatter_kernel<<<n_blocks, dsb::kCompactBlock, 0, s->stream>>>(s->fill_pts, (long long)n, s->fill_totals, 0LL,
                                                                            0LL, (long long)n, accepted_dev);
    int accepted = 0;
    DSB_CUDA(cudaMemcpyAsync(&accepted, s->fill_totals + n_blocks, sizeof(int), cudaMemcpyDeviceToHost, s->stream));
    DSB_CUDA(cudaStreamSynchronize(s->stream));
    DSB_CUDA(cudaGetLastError());
    *n_accepted = accepted;
    return DSB_OK;
}

int dsb_fill_mesh(int32_t device, const dsb_mesh *mesh, const double *voxel_size, int intra, uint64_t seed,
                  int64_t n_points, int64_t cuda_bs, double *points)
{
    return guarded([&]() -> int {
        if (!mesh || !voxel_size || !points || n_points <= 0 || cuda_bs <= 0) return fail(DSB_EINVAL, "bad arguments");
        DSB_CUDA(cudaSetDevice(device));
        MeshBuffers mb;
        int rc = upload_mesh(*mesh, mb);
        if (rc) {
            mb.release();
            return rc;
        }
        const int64_t n_states = (n_points + cuda_bs - 1) / cuda_bs * cuda_bs;
        ulonglong2 *d_rng = nullptr;
        double *d_pts = nullptr;
        int *d_count = nullptr;
        cudaError_t e = cudaMalloc(&d_rng, sizeof(ulonglong2) * (size_t)n_states);
        if (e == cudaSuccess) e = cudaMalloc(&d_pts, sizeof(double) * 3 * (size_t)n_points);
        if (e == cudaSuccess) e = cudaMalloc(&d_count, sizeof(int));
        if (e != cudaSuccess) rc = fail(DSB_ENOMEM, cudaGetErrorString(e));
        if (!rc) rc = build_fill_columns(mb);
        if (!rc) rc = launch_rng_init(device, seed, 0, n_states, d_rng, 0);
        std::vector<double> round_pts((size_t)n_points * 3);
        int64_t have = 0;
        const double inf = INFINITY;
        // The reference launches one round per host-loop iteration, keeps the accepted points of
        // every round in thread order and stops once it has enough (simulations.py:554-579).
        for (int round = 0; !rc && have < n_points; ++round) {
            if (round > 100000) {
                rc = fail(DSB_ESTATE, "fill_mesh: no acceptable points (is the surface closed?)");
                break;
            }
            dsb::fill_mesh_kernel<<<(unsigned)((n_points + 127) / 128), 128>>>(mb.dev, mb.columns, voxel_size[0], voxel_size[1],
                                                                                voxel_size[2], intra, (long long)n_points,
                                                                                d_rng, d_pts);
            e = cudaGetLastError();
            if (e == cudaSuccess)
                e = cudaMemcpy(round_pts.data(), d_pts, sizeof(double) * 3 * (size_t)n_points, cudaMemcpyDeviceToHost);
            if (e != cudaSuccess) {
                rc = fail(DSB_ECUDA, cudaGetErrorString(e));
                break;
            }
            for (int64_t i = 0; i < n_points && have < n_points; ++i)
                if (round_pts[3 * i] != inf) {
                    memcpy(points + 3 * have, &round_pts[3 * i], 3 * sizeof(double));
                    ++have;
                }
        }
        cudaFree(d_rng);
        cudaFree(d_pts);
        cudaFree(d_count);
        mb.release();
        return rc;
    });
}

int dsb_copy_signal_dev(dsb_sim *s, double *dst_dev)
{
    if (!s || !dst_dev) return fail(DSB_EINVAL, "null argument");
    if (!s->finalized) return fail(DSB_ESTATE, "signal is available after the last time step only");
    Range nvtx("dsb_copy_signal_dev: wait + D2D");
    DSB_CUDA(cudaSetDevice(s->prm.device));
    DSB_CUDA(cudaMemcpyAsync(dst_dev, s->d_signal, sizeof(double) * (size_t)(s->prm.n_meas + 1), cudaMemcpyDeviceToDevice,
                             s->stream));
    DSB_CUDA(cudaStreamSynchronize(s->stream));
    return DSB_OK;
}

int dsb_fill_mesh_multi(dsb_sim **sims, int32_t n_sims, const double *voxel_size, int intra, uint64_t seed, int64_t n_points)
{
    return guarded([&]() -> int {
        if (!sims || n_sims <= 0 || !voxel_size || n_points <= 0 || n_points > 0x7fffffffLL) return fail(DSB_EINVAL, "bad arguments");
        Range nvtx("dsb_fill_mesh_multi: sampler rounds over the device list");
        // handle k holds the global walkers [off_k, off_k + n_k) and evaluates the sampler threads with the
        // same numbers in every round: together the handles must tile [0, n_points)
        std::vector<int64_t> w0((size_t)n_sims), w1((size_t)n_sims);
        int64_t expect = 0;
        for (int k = 0; k < n_sims; ++k) {
            if (!sims[k] || sims[k]->prm.substrate != DSB_MESH) return fail(DSB_ESTATE, "dsb_fill_mesh_multi needs mesh handles");
            w0[(size_t)k] = sims[k]->prm.walker_offset;
            w1[(size_t)k] = w0[(size_t)k] + sims[k]->prm.n_walkers;
            if (w0[(size_t)k] != expect) return fail(DSB_EINVAL, "the handles' walker ranges must tile [0, n_points) in order");
            expect = w1[(size_t)k];
        }
        if (expect != n_points) return fail(DSB_EINVAL, "the handles' walker ranges must tile [0, n_points) in order");
        std::vector<double *> accepted((size_t)n_sims, nullptr);
        std::vector<int64_t> count((size_t)n_sims, 0);
        std::vector<int> rcs((size_t)n_sims, DSB_OK);
        std::vector<std::string> errs((size_t)n_sims);
        auto on_all = [&](auto fn) {   // fn(k) on one host thread per handle; errors collected per handle
            std::vector<std::thread> pool;
            for (int k = 0; k < n_sims; ++k)
                pool.emplace_back([&, k] {
                    rcs[(size_t)k] = fn(k);
                    if (rcs[(size_t)k]) errs[(size_t)k] = g_err;
                });
            for (auto &th : pool) th.join();
            for (int k = 0; k < n_sims; ++k)
                if (rcs[(size_t)k]) return fail(rcs[(size_t)k], errs[(size_t)k]);
            return (int)DSB_OK;
        };
        // accepted points travel device to device: direct peer access where the hardware offers it (NVLink /
        // NVSwitch on a B200 node); without it cudaMemcpyPeerAsync stages through the host
        for (int d = 0; d < n_sims; ++d)
            for (int r = 0; r < n_sims; ++r) {
                const int dd = sims[d]->prm.device, dr = sims[r]->prm.device;
                int can = 0;
                if (dd != dr && cudaDeviceCanAccessPeer(&can, dd, dr) == cudaSuccess && can) {
                    cudaSetDevice(dd);
                    cudaDeviceEnablePeerAccess(dr, 0);   // (cudaErrorPeerAccessAlreadyEnabled from an earlier call is fine)
                    cudaGetLastError();
                }
            }
        int rc = on_all([&](int k) -> int {
            int r = dsb_fill_shard_begin(sims[k], seed, w0[(size_t)k], w1[(size_t)k]);
            if (r) return r;
            cudaError_t e = cache_malloc(&accepted[(size_t)k], sizeof(double) * 3 * (size_t)(w1[(size_t)k] - w0[(size_t)k]));
            return e == cudaSuccess ? DSB_OK : fail(DSB_ENOMEM, cudaGetErrorString(e));
        });
        int64_t have = 0;
        for (int round = 0; !rc && have < n_points; ++round) {
            if (round > 100000) {
                rc = fail(DSB_ESTATE, "fill_mesh: no acceptable points (is the surface closed?)");
                break;
            }
            rc = on_all([&](int k) { return dsb_fill_shard_round(sims[k], voxel_size, intra, accepted[(size_t)k], &count[(size_t)k]); });
            if (rc) break;
            // the round's accepted points, concatenated in handle = thread order, are the global points
            // have, have + 1, ...: every handle copies its walkers' rows from wherever they were produced
            rc = on_all([&](int d) -> int {
                dsb_sim *dst = sims[d];
                DSB_CUDA(cudaSetDevice(dst->prm.device));
                int64_t start = have;
                for (int r = 0; r < n_sims; ++r) {
                    const int64_t a = std::max(start, w0[(size_t)d]), b = std::min(start + count[(size_t)r], w1[(size_t)d]);
                    if (a < b)
                        DSB_CUDA(cudaMemcpyPeerAsync(dst->d_pos + 3 * (a - w0[(size_t)d]), dst->prm.device,
                                                     accepted[(size_t)r] + 3 * (a - start), sims[r]->prm.device,
                                                     sizeof(double) * 3 * (size_t)(b - a), dst->stream));
                    start += count[(size_t)r];
                }
                DSB_CUDA(cudaStreamSynchronize(dst->stream));  // the sources are overwritten by the next round
                return DSB_OK;
            });
            for (int k = 0; k < n_sims; ++k) have += count[(size_t)k];
        }
        std::string keep = g_err;
        for (int k = 0; k < n_sims; ++k) {
            cudaSetDevice(sims[k]->prm.device);
            cache_free(accepted[(size_t)k]);
            dsb_fill_shard_end(sims[k]);
            if (!rc) {
                int r = rewind_sim(sims[k]);
                if (r) rc = r, keep = g_err;
            }
        }
        if (rc) g_err = keep;
        return rc;
    });
}

int dsb_simulate_multi(const dsb_params *params, const int32_t *devices, int32_t n_devices, const double *gradient,
                       const double *positions_in, double *signal_out, int64_t *n_valid_out, double *positions_out,
                       double *phases_out, uint8_t *iter_exc_out)
{
    return guarded([&]() -> int {
        if (!params || !devices || n_devices <= 0 || !positions_in || !signal_out) return fail(DSB_EINVAL, "null argument");
        Range nvtx("dsb_simulate_multi");
        const int64_t N = params->n_walkers, M = params->n_meas;
        if (N <= 0 || M <= 0) return fail(DSB_EINVAL, "n_walkers and n_meas must be positive");
        const int n_used = (int)std::min<int64_t>(n_devices, N);   // never an empty shard
        std::vector<int> rcs((size_t)n_used, DSB_OK);
        std::vector<std::string> errs((size_t)n_used);
        std::vector<std::vector<double>> sig((size_t)n_used, std::vector<double>((size_t)M, 0.0));
        std::vector<int64_t> valid((size_t)n_used, 0);
        std::vector<std::thread> pool;
        for (int k = 0; k < n_used; ++k)
            pool.emplace_back([&, k] {
                try {
                const int64_t lo = N * k / n_used, hi = N * (k + 1) / n_used;
                dsb_params p = *params;
                p.device = devices[k];
                p.n_walkers = hi - lo;
                p.walker_offset = params->walker_offset + lo;
                dsb_sim *s = nullptr;
                int rc = dsb_create(&p, gradient, &s);
                if (!rc) rc = dsb_set_positions(s, positions_in + 3 * lo);
                if (!rc) rc = dsb_run(s, 0, p.n_t);
                if (!rc) rc = dsb_get_signal(s, sig[(size_t)k].data(), &valid[(size_t)k]);
                if (!rc && positions_out) rc = dsb_get_positions(s, positions_out + 3 * lo);
                if (!rc && iter_exc_out) rc = dsb_get_iter_exc(s, iter_exc_out + lo);
                if (!rc && phases_out) {
                    std::vector<double> ph((size_t)(M * (hi - lo)));
                    rc = dsb_get_phases(s, ph.data());
                    if (!rc)
                        for (int64_t m = 0; m < M; ++m)
                            memcpy(phases_out + m * N + lo, ph.data() + m * (hi - lo), sizeof(double) * (size_t)(hi - lo));
                }
                if (rc) errs[(size_t)k] = g_err;
                rcs[(size_t)k] = rc;
                dsb_destroy(s);
                } catch (...) {   // (an exception must not leave a host thread: std::terminate)
                    rcs[(size_t)k] = DSB_ENOMEM;
                }
            });
        for (auto &th : pool) th.join();
        for (int k = 0; k < n_used; ++k)
            if (rcs[(size_t)k]) return fail(rcs[(size_t)k], errs[(size_t)k].empty() ? "host allocation failed" : errs[(size_t)k]);
        int64_t total_valid = 0;
        for (int64_t m = 0; m < M; ++m) signal_out[m] = 0.0;
        for (int k = 0; k < n_used; ++k) {   // fixed order: the sum does not depend on which device finished first
            for (int64_t m = 0; m < M; ++m) signal_out[m] += sig[(size_t)k][(size_t)m];
            total_valid += valid[(size_t)k];
        }
        if (n_valid_out) *n_valid_out = total_valid;
        return DSB_OK;
    });
}

}  // extern "C"

// ------------------------------------------------------------- the one collective, inside the library

// NCCL is bound at run time (dlopen), so that the library neither links against it nor needs it for
// single-GPU or device-list runs.  Only the five entry points the path needs.
namespace {

struct NcclUniqueId {
    char internal[128];
};

struct NcclApi {
    void *handle = nullptr;
    int (*GetUniqueId)(NcclUniqueId *) = nullptr;
    int (*CommInitRank)(void **, int, NcclUniqueId, int) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, void *, cudaStream_t) = nullptr;
    int (*CommDestroy)(void *) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    std::string path;
};

std::mutex g_nccl_mu;
NcclApi g_nccl;

int nccl_api(const char *hint, NcclApi **out)
{
    std::lock_guard<std::mutex> lk(g_nccl_mu);
    if (!g_nccl.handle) {
        std::vector<std::string> names;
        if (const char *env = getenv("DISIMPY_B200_NCCL_LIB")) names.push_back(env);
        if (hint && *hint) names.push_back(hint);
        names.push_back("libnccl.so.2");
        names.push_back("libnccl.so");
        std::string tried;
        for (const std::string &n : names) {
            void *h = dlopen(n.c_str(), RTLD_NOW | RTLD_LOCAL);
            if (h) {
                g_nccl.handle = h;
                g_nccl.path = n;
                break;
            }
            tried += n + "; ";
        }
        if (!g_nccl.handle) return fail(DSB_ESTATE, "NCCL library not found (tried " + tried + "set DISIMPY_B200_NCCL_LIB)");
        g_nccl.GetUniqueId = reinterpret_cast<decltype(g_nccl.GetUniqueId)>(dlsym(g_nccl.handle, "ncclGetUniqueId"));
        g_nccl.CommInitRank = reinterpret_cast<decltype(g_nccl.CommInitRank)>(dlsym(g_nccl.handle, "ncclCommInitRank"));
        g_nccl.AllReduce = reinterpret_cast<decltype(g_nccl.AllReduce)>(dlsym(g_nccl.handle, "ncclAllReduce"));
        g_nccl.CommDestroy = reinterpret_cast<decltype(g_nccl.CommDestroy)>(dlsym(g_nccl.handle, "ncclCommDestroy"));
        g_nccl.GetErrorString = reinterpret_cast<decltype(g_nccl.GetErrorString)>(dlsym(g_nccl.handle, "ncclGetErrorString"));
        if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllReduce || !g_nccl.CommDestroy || !g_nccl.GetErrorString) {
            dlclose(g_nccl.handle);
            g_nccl = NcclApi();
            return fail(DSB_ESTATE, "NCCL library lacks an expected symbol");
        }
    }
    *out = &g_nccl;
    return DSB_OK;
}

#define DSB_NCCL(api, expr)                                                                            \
    do {                                                                                               \
        int r_ = (expr);                                                                               \
        if (r_ != 0) return fail(DSB_ECUDA, std::string(#expr) + ": " + (api)->GetErrorString(r_));   \
    } while (0)

}  // namespace

extern "C" {

int dsb_nccl_unique_id(const char *nccl_library, uint8_t id_out[128])
{
    if (!id_out) return fail(DSB_EINVAL, "null argument");
    NcclApi *api = nullptr;
    int rc = nccl_api(nccl_library, &api);
    if (rc) return rc;
    NcclUniqueId id;
    DSB_NCCL(api, api->GetUniqueId(&id));
    memcpy(id_out, id.internal, 128);
    return DSB_OK;
}

int dsb_nccl_init(const char *nccl_library, int32_t device, int32_t rank, int32_t world_size, const uint8_t id[128],
                  dsb_comm **out)
{
    if (!id || !out || world_size < 1 || rank < 0 || rank >= world_size) return fail(DSB_EINVAL, "bad arguments");
    *out = nullptr;
    NcclApi *api = nullptr;
    int rc = nccl_api(nccl_library, &api);
    if (rc) return rc;
    DSB_CUDA(cudaSetDevice(device));
    NcclUniqueId uid;
    memcpy(uid.internal, id, 128);
    void *comm = nullptr;
    Range nvtx("dsb_nccl_init");
    DSB_NCCL(api, api->CommInitRank(&comm, world_size, uid, rank));
    *out = reinterpret_cast<dsb_comm *>(comm);
    return DSB_OK;
}

int dsb_nccl_destroy(dsb_comm *comm)
{
    if (!comm) return DSB_OK;
    NcclApi *api = nullptr;
    int rc = nccl_api(nullptr, &api);
    if (rc) return rc;
    DSB_NCCL(api, api->CommDestroy(comm));
    return DSB_OK;
}

int dsb_allreduce_signal(dsb_sim *s, dsb_comm *comm, double *signal, int64_t *n_valid)
{
    return guarded([&]() -> int {
        if (!s || !comm || !signal) return fail(DSB_EINVAL, "null argument");
        if (!s->finalized) return fail(DSB_ESTATE, "signal is available after the last time step only");
        NcclApi *api = nullptr;
        int rc = nccl_api(nullptr, &api);
        if (rc) return rc;
        Range nvtx("dsb_allreduce_signal: NCCL all-reduce + D2H");
        DSB_CUDA(cudaSetDevice(s->prm.device));
        const size_t count = (size_t)s->prm.n_meas + 1;
        // from the handle's own result buffer, on its stream: ordered after the reduction kernel, no host detour
        // (out of place, next to it: the shard's own result stays what dsb_get_signal returns, and the call can be repeated)
        double *global = s->d_signal + count;
        DSB_NCCL(api, api->AllReduce(s->d_signal, global, count, /*ncclFloat64*/ 8, /*ncclSum*/ 0, comm, s->stream));
        std::vector<double> h(count);
        DSB_CUDA(cudaMemcpyAsync(h.data(), global, count * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
        DSB_CUDA(cudaStreamSynchronize(s->stream));
        memcpy(signal, h.data(), sizeof(double) * (size_t)s->prm.n_meas);
        if (n_valid) *n_valid = (int64_t)llround(h.back());
        return DSB_OK;
    });
}

int dsb_allreduce_zeros(int32_t device, dsb_comm *comm, int64_t n_meas, double *signal, int64_t *n_valid)
{
    return guarded([&]() -> int {
        if (!comm || !signal || n_meas <= 0) return fail(DSB_EINVAL, "bad arguments");
        NcclApi *api = nullptr;
        int rc = nccl_api(nullptr, &api);
        if (rc) return rc;
        DSB_CUDA(cudaSetDevice(device));
        const size_t count = (size_t)n_meas + 1;
        double *d = nullptr;
        DSB_CUDA(cache_malloc(&d, count * sizeof(double)));
        std::vector<double> h(count);
        cudaError_t e = cudaMemset(d, 0, count * sizeof(double));
        int nrc = 0;
        if (e == cudaSuccess) nrc = api->AllReduce(d, d, count, 8, 0, comm, (cudaStream_t)0);
        if (e == cudaSuccess && nrc == 0) e = cudaMemcpy(h.data(), d, count * sizeof(double), cudaMemcpyDeviceToHost);
        cache_free(d);
        if (nrc != 0) return fail(DSB_ECUDA, std::string("ncclAllReduce: ") + api->GetErrorString(nrc));
        if (e != cudaSuccess) return fail(DSB_ECUDA, cudaGetErrorString(e));
        memcpy(signal, h.data(), sizeof(double) * (size_t)n_meas);
        if (n_valid) *n_valid = (int64_t)llround(h.back());
        return DSB_OK;
    });
}

}  // extern "C"
