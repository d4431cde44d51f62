// dsb_fill.cuh -- rejection sampler for initial positions inside / outside a closed triangular
// surface.  Replaces _cuda_fill_mesh (disimpy/simulations.py:421-502): same three
// uniform_float64 draws per thread per round, same +x ray parity test through the subvoxel
// grid (non-periodic lookups), same 1000-hit abandon rule.
#pragma once
#include "dsb_kernels.cuh"

namespace dsb {

constexpr int kMaxRayHits = 1000;  // simulations.py:462-467

// The reference's loop as it is: every listed triangle of every cell from the point's cell to the
// end of the grid along +x is ray-tested, distinct hits are counted, the thread gives up at 1000.
// Out of line: only points with 1000 or more crossings, or with cell ranges the column walk does
// not cover, come here.
__device__ __noinline__ bool fill_ray_parity_reference(const MeshDev &g, const Vec3 &pt, int lx, int ux, int ly, int uy,
                                                       int lz, int uz, bool &abandoned)
{
    const Vec3 ray = {1.0, 0.0, 0.0};
    int hits[kMaxRayHits];
    int n_hits = 0;
    abandoned = false;
    for (int x = lx; x < ux && !abandoned; ++x)
        for (int y = ly; y < uy && !abandoned; ++y)
            for (int z = lz; z < uz && !abandoned; ++z) {
                int2 c = __ldg(g.cell_rng + ((long long)x * g.nsv1 + y) * g.nsv2 + z);
                for (int i = c.x; i < c.y; ++i) {
                    if (n_hits >= kMaxRayHits) {
                        abandoned = true;
                        break;
                    }
                    int tri = __ldg(g.tri_idx + i);
                    double d = ray_triangle(load_tri(g.tri, tri), pt, ray);
                    if (d > 0) {
                        bool seen = false;
                        for (int j = 0; j < n_hits; ++j)
                            if (hits[j] == tri) {
                                seen = true;
                                break;
                            }
                        if (!seen) hits[n_hits++] = tri;
                    }
                }
            }
    return (n_hits & 1) == 1;
}

// What the sampler's +x rays need of the cell lists, built once per mesh on first use
// (build_fill_columns in dsb_api.cu): for every (y, z) column of cells, the distinct triangles listed
// anywhere in the column, ordered by the last x cell that lists them (descending).  The triangles
// the reference visits from cell lx to the end of the grid are then the first cnt[column][lx]
// entries: one contiguous, duplicate-free stretch that a warp reads coalesced.
struct FillColumns {
    const int *start;     // (n1 * n2 + 1,) first entry of every column
    const int *cnt;       // (n1 * n2, n0): entries of the column listed in some cell x' >= x
    const uint4 *entry;   // the walk's 16-byte list entries (triangle id | edge-on mark, 15-bit box)
    int n0;
};

constexpr int kFillWarps = 4;

// 32 proposed points per warp, one per lane; their columns are then scanned one point after the
// other by the whole warp, 64 entries per iteration.  A listed triangle is ray-tested only if its box
// (15-bit grid, rounded outwards) meets the ray's: a triangle the reference counts is hit at
// u, v in [0, 1], so the point's y and z lie within rounding of the triangle's, and its x below the
// triangle's top; edge-on triangles, for which rounding says nothing, are marked at upload and
// always tested.  Survivors of any point queue up in shared memory and are tested 32 at a time,
// all lanes busy.  The number of distinct hits is the reference's as long as it stays below its
// 1000-hit rule; beyond, and for the point configurations the column walk does not cover, the
// thread redoes the point with the reference's loop.
__global__ void __launch_bounds__(32 * kFillWarps) fill_mesh_kernel(const MeshDev g, const FillColumns fc, double vx,
                                                                  double vy, double vz, int intra, long long n_points,
                                                                  ulonglong2 *rng_states, double *points)
{
    __shared__ double s_pt[kFillWarps][3][32];
    __shared__ int s_hits[kFillWarps][32];
    __shared__ unsigned long long s_queue[kFillWarps][64];
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const long long id = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = id < n_points;
    const double inf = __longlong_as_double(0x7FF0000000000000LL);
    const Vec3 ray = {1.0, 0.0, 0.0};
    Vec3 pt = {0.0, 0.0, 0.0};
    int lx = 0, ly = 0, lz = 0, ux = 0, uy = 0, uz = 0;
    int first = 0, n = 0;
    bool alone = false;   // this point takes the reference's loop
    unsigned a0 = 0, a1 = 0, a2 = 0;
    if (live) {
        Rng rng = {rng_states[id].x, rng_states[id].y};
        pt.x = mul_(u01_f64(rng_next(rng)), vx);
        pt.y = mul_(u01_f64(rng_next(rng)), vy);
        pt.z = mul_(u01_f64(rng_next(rng)), vz);
        rng_states[id] = make_ulonglong2(rng.s0, rng.s1);
        lx = ll_overlap(g.xs, g.len_xs, fmin(pt.x, add_(pt.x, 1.0)), g.inv_hx);
        ly = ll_overlap(g.ys, g.len_ys, pt.y, g.inv_hy);
        lz = ll_overlap(g.zs, g.len_zs, pt.z, g.inv_hz);
        ux = ul_overlap(g.xs, g.len_xs, fmax(pt.x, add_(pt.x, 1.0)), g.inv_hx);
        uy = ul_overlap(g.ys, g.len_ys, pt.y, g.inv_hy);
        uz = ul_overlap(g.zs, g.len_zs, pt.z, g.inv_hz);
        if (ux > lx && uy > ly && uz > lz) {   // (otherwise no cell is visited: no hits)
            if (fc.entry != nullptr && ux == fc.n0 && uy == ly + 1 && uz == lz + 1) {
                const int col = ly * g.nsv2 + lz;
                first = __ldg(fc.start + col);
                n = __ldg(fc.cnt + (long long)col * fc.n0 + lx);
            } else {
                alone = true;
            }
        }
        // the ray's box on the 15-bit grid: [x - 2 units, end] x [y -+ 2 units] x [z -+ 2 units]
        const int qx = quantize(pt.x, g.qscale[0]), qy = quantize(pt.y, g.qscale[1]), qz = quantize(pt.z, g.qscale[2]);
        a0 = (32767u | (clamp15(qy + 3) << 16)) + kSwarH;
        a1 = (clamp15(qz + 3) | ((32767u - clamp15(qx - 2)) << 16)) + kSwarH;
        a2 = ((32767u - clamp15(qy - 2)) | ((32767u - clamp15(qz - 2)) << 16)) + kSwarH;
    }
    s_pt[w][0][lane] = pt.x;
    s_pt[w][1][lane] = pt.y;
    s_pt[w][2][lane] = pt.z;
    s_hits[w][lane] = 0;
    __syncwarp();

    int head = 0, n_queue = 0;
    auto test_queued = [&](int count) {   // exact tests of the `count` oldest queued survivors, one per lane
        if (lane < count) {
            const unsigned long long rec = s_queue[w][(head + lane) & 63];
            const int owner = (int)(rec >> 32);
            const Vec3 o = {s_pt[w][0][owner], s_pt[w][1][owner], s_pt[w][2][owner]};
            if (ray_triangle(load_tri(g.tri, (int)(unsigned)rec), o, ray) > 0) atomicAdd(&s_hits[w][owner], 1);
        }
        head = (head + count) & 63;
        n_queue -= count;
        __syncwarp();
    };
    for (int p = 0; p < 32; ++p) {
        const int n_p = __shfl_sync(full, n, p);
        if (n_p == 0) continue;
        const uint4 *list = fc.entry + __shfl_sync(full, first, p);
        const unsigned b0 = __shfl_sync(full, a0, p), b1 = __shfl_sync(full, a1, p), b2 = __shfl_sync(full, a2, p);
        for (int i0 = 0; i0 < n_p; i0 += 64) {
            uint4 e[2];
            bool meets[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int i = i0 + 32 * h + lane;
                e[h] = make_uint4(0u, 0xffffffffu, 0xffffffffu, 0xffffffffu);   // never meets
                if (i < n_p) e[h] = __ldg(list + i);
            }
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                meets[h] = ((b0 - e[h].y) & (b1 - e[h].z) & (b2 - e[h].w) & kSwarH) == kSwarH || (e[h].x & ~kEntryTriMask);
                const unsigned m = __ballot_sync(full, meets[h]);
                if (m == 0) continue;
                if (meets[h])
                    s_queue[w][(head + n_queue + __popc(m & ((1u << lane) - 1u))) & 63] =
                        (unsigned long long)(e[h].x & kEntryTriMask) | ((unsigned long long)p << 32);
                n_queue += __popc(m);
                __syncwarp();
                if (n_queue >= 32) test_queued(32);
            }
        }
    }
    if (n_queue > 0) test_queued(n_queue);

    if (!live) return;
    const int n_hits = s_hits[w][lane];
    bool abandoned = false;
    bool inside = (n_hits & 1) == 1;
    if (alone || n_hits >= kMaxRayHits) inside = fill_ray_parity_reference(g, pt, lx, ux, ly, uy, lz, uz, abandoned);
    const bool keep = !abandoned && (intra ? inside : !inside);
    points[3 * id] = keep ? pt.x : inf;
    points[3 * id + 1] = keep ? pt.y : inf;
    points[3 * id + 2] = keep ? pt.z : inf;
}

// ---- keeping the accepted points of a round on the device, in thread order -----------------

constexpr int kCompactBlock = 1024;

// accepted (x != inf) candidates of this block's 1024 candidates: count, and exclusive rank of
// the calling thread's candidate among them
__device__ __forceinline__ int block_rank(bool accepted, int &block_total)
{
    __shared__ int s_warp[kCompactBlock / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned m = __ballot_sync(0xffffffffu, accepted);
    if (lane == 0) s_warp[warp] = __popc(m);
    __syncthreads();
    int before = 0, total = 0;
    for (int k = 0; k < kCompactBlock / 32; ++k) {
        const int c = s_warp[k];
        before += k < warp ? c : 0;
        total += c;
    }
    __syncthreads();
    block_total = total;
    return before + __popc(m & ((1u << lane) - 1u));
}

__global__ void __launch_bounds__(kCompactBlock) fill_count_kernel(const double *points, long long n, int *block_totals)
{
    const long long i = (long long)blockIdx.x * kCompactBlock + threadIdx.x;
    const bool ok = i < n && points[3 * i] != __longlong_as_double(0x7FF0000000000000LL);
    int total;
    block_rank(ok, total);
    if (threadIdx.x == 0) block_totals[blockIdx.x] = total;
}

// exclusive prefix sums of the block totals, in place (one block; every thread takes a
// contiguous stretch); totals[n_blocks] receives the grand total
__global__ void __launch_bounds__(1024) fill_scan_kernel(int *totals, int n_blocks)
{
    __shared__ long long s_sum[1024];
    const int per = (n_blocks + 1023) / 1024;
    const int lo = min(threadIdx.x * per, n_blocks), hi = min(lo + per, n_blocks);
    long long acc = 0;
    for (int k = lo; k < hi; ++k) acc += totals[k];
    s_sum[threadIdx.x] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        long long run = 0;
        for (int k = 0; k < 1024; ++k) {
            const long long v = s_sum[k];
            s_sum[k] = run;
            run += v;
        }
        totals[n_blocks] = (int)run;
    }
    __syncthreads();
    int run = (int)s_sum[threadIdx.x];
    for (int k = lo; k < hi; ++k) {
        const int v = totals[k];
        totals[k] = run;
        run += v;
    }
}

// accepted candidate number r of this round (thread order) becomes global point `have + r`;
// points [first, first + count) are kept in out (count x 3)
__global__ void __launch_bounds__(kCompactBlock) fill_scatter_kernel(const double *points, long long n,
                                                                     const int *block_offsets, long long have,
                                                                     long long first, long long count, double *out)
{
    const long long i = (long long)blockIdx.x * kCompactBlock + threadIdx.x;
    const bool ok = i < n && points[3 * i] != __longlong_as_double(0x7FF0000000000000LL);
    int total;
    const int r = block_rank(ok, total);
    if (!ok) return;
    const long long g = have + block_offsets[blockIdx.x] + r - first;
    if (g >= 0 && g < count) {
        out[3 * g] = points[3 * i];
        out[3 * g + 1] = points[3 * i + 1];
        out[3 * g + 2] = points[3 * i + 2];
    }
}

}  // namespace dsb
