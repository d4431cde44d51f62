// dsb_fill.cuh -- rejection sampler for initial positions inside / outside a closed triangular
// surface.  Replaces _cuda_fill_mesh (disimpy/simulations.py:421-502): same three
// uniform_float64 draws per thread per round, same +x ray parity test through the subvoxel
// grid (non-periodic lookups), same 1000-hit abandon rule.
#pragma once
#include "dsb_kernels.cuh"

namespace dsb {

constexpr int kMaxRayHits = 1000;  // simulations.py:462-467

__global__ void __launch_bounds__(128) fill_mesh_kernel(const MeshDev g, double vx, double vy, double vz,
                                                        int intra, long long n_points, ulonglong2 *rng_states,
                                                        double *points)
{
    const long long id = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= n_points) return;
    const double inf = __longlong_as_double(0x7FF0000000000000LL);
    Rng rng = {rng_states[id].x, rng_states[id].y};
    Vec3 pt;
    pt.x = mul_(u01_f64(rng_next(rng)), vx);
    pt.y = mul_(u01_f64(rng_next(rng)), vy);
    pt.z = mul_(u01_f64(rng_next(rng)), vz);
    rng_states[id] = make_ulonglong2(rng.s0, rng.s1);
    const Vec3 ray = {1.0, 0.0, 0.0};
    const int lx = ll_overlap(g.xs, g.len_xs, fmin(pt.x, add_(pt.x, 1.0)), g.inv_hx);
    const int ly = ll_overlap(g.ys, g.len_ys, pt.y, g.inv_hy);
    const int lz = ll_overlap(g.zs, g.len_zs, pt.z, g.inv_hz);
    const int ux = ul_overlap(g.xs, g.len_xs, fmax(pt.x, add_(pt.x, 1.0)), g.inv_hx);
    const int uy = ul_overlap(g.ys, g.len_ys, pt.y, g.inv_hy);
    const int uz = ul_overlap(g.zs, g.len_zs, pt.z, g.inv_hz);
    int hits[kMaxRayHits];
    int n_hits = 0;
    bool abandoned = false;
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
    const bool inside = (n_hits & 1) == 1;
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
