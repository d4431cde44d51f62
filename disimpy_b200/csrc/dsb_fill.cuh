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

}  // namespace dsb
