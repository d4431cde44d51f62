// dsb_kernels.cuh -- the walk kernels.  One thread owns one walker for the whole launch: RNG
// state, FP64 position and (for up to DSB_MAX_REG_MEAS measurements) the phase accumulators
// stay in registers across all time steps of the launch; HBM is touched once at entry and once
// at exit.  Replaces the reference's one-launch-per-time-step kernels
// (disimpy/simulations.py:682-1013) and the host-side signal reduction (:1413-1421).
#pragma once
#include "dsb_geom.cuh"

namespace dsb {

#ifndef DSB_BLOCK
#define DSB_BLOCK 128
#endif
#ifndef DSB_MIN_BLOCKS
#define DSB_MIN_BLOCKS 1
#endif
#ifndef DSB_MESH_MIN_BLOCKS
#define DSB_MESH_MIN_BLOCKS 4
#endif
constexpr int kBlock = DSB_BLOCK;    // threads (= walkers) per CTA
constexpr int kMaxRegMeas = 4;       // measurements whose phase lives in registers
constexpr int kTimeChunk = 8;        // steps buffered per phase pass when n_meas is larger

struct MeshDev {
    const double *tri9;     // (n_faces, 9): A, B-A, C-A
    const int *tri_idx;     // (K,) triangle ids, cell after cell (reference order)
    const int2 *cell_rng;   // (n_cells,) [begin, end) into tri_idx
    const double *xs, *ys, *zs;
    int len_xs, len_ys, len_zs;
    int nsv1, nsv2;
    double inv_hx, inv_hy, inv_hz;  // guesses only, never part of a result
    double vox[3];      // |xs[-1] - xs[0]| per axis, the period of the lookup (simulations.py:660)
    double inv_vox[3];  // 1 / vox, guess only
    double top[3];      // xs[-1], ys[-1], zs[-1]: the image shift unit (simulations.py:943)
    double perm_prob;
};

struct KParams {
    long long n_walkers;
    int n_meas, n_t;
    int t0, t1;
    int finalize;  // t1 == n_t: also emit per-block sum(cos phase) partials
    long long max_iter;
    double step_l, gamma_dt, eps, radius;
    double R[9], Rinv[9], ax[3];
    const double *grad;         // (n_meas, n_t, 3)
    double *pos;                // (n_walkers, 3)
    unsigned long long *rng;    // (n_walkers, 2)
    double *phases;             // (n_meas, n_walkers)
    unsigned char *iter_exc;    // (n_walkers,)
    double *partials;           // (n_meas + 1, gridDim.x)
    MeshDev mesh;
};

// ---------------------------------------------------------------- one time step, per substrate

template <int SUB>
__device__ __forceinline__ bool walker_step(Vec3 &pos, Rng &rng, const KParams &p, const double *tab,
                                            const bool live);

// simulations.py:682-702
template <>
__device__ __forceinline__ bool walker_step<0>(Vec3 &pos, Rng &rng, const KParams &p, const double *tab,
                                               const bool live)
{
    if (!live) return false;
    Vec3 s = random_step(rng, tab);
    pos.x = fma_(s.x, p.step_l, pos.x);
    pos.y = fma_(s.y, p.step_l, pos.y);
    pos.z = fma_(s.z, p.step_l, pos.z);
    return false;
}

// simulations.py:705-756
template <>
__device__ __forceinline__ bool walker_step<1>(Vec3 &pos, Rng &rng, const KParams &p, const double *tab,
                                               const bool live)
{
    if (!live) return false;
    Vec3 s = random_step(rng, tab);
    double step_l = p.step_l;
    long long iter = 0;
    bool check = true;
    while (check && step_l > 0 && iter < p.max_iter) {
        ++iter;
        double d = line_sphere(pos, s, p.radius);
        if (d > 0 && d < step_l) {
            Vec3 n;
            n.x = -fma_(d, s.x, pos.x);
            n.y = -fma_(d, s.y, pos.y);
            n.z = -fma_(d, s.z, pos.z);
            n = normalize3(n);
            reflect(pos, s, d, n, p.eps);
            step_l = sub_(step_l, add_(d, p.eps));
        } else {
            check = false;
        }
    }
    pos.x = fma_(step_l, s.x, pos.x);
    pos.y = fma_(step_l, s.y, pos.y);
    pos.z = fma_(step_l, s.z, pos.z);
    return iter >= p.max_iter;
}

// simulations.py:759-816.  The position is rotated into the cylinder frame and back every
// step, like the reference (the rounding of the round trip is part of the trajectory).
template <>
__device__ __forceinline__ bool walker_step<2>(Vec3 &pos, Rng &rng, const KParams &p, const double *tab,
                                               const bool live)
{
    if (!live) return false;
    Vec3 s = random_step(rng, tab);
    Vec3 r0 = matvec3(p.R, pos);
    double step_l = p.step_l;
    long long iter = 0;
    bool check = true;
    while (check && step_l > 0 && iter < p.max_iter) {
        ++iter;
        double d = line_circle(r0, s, p.radius);
        if (d > 0 && d < step_l) {
            double X1 = fma_(d, s.y, r0.y), X2 = fma_(d, s.z, r0.z);
            double len = sqrt_(fma_(X2, X2, fma_(X1, X1, 0.0)));
            // normal = (0, -X1, -X2) / len; 0 / len is +0 for every finite positive len
            Vec3 n;
            double rc = rcp_refined(len);
            bool ok = len > 0 && div_fast(-X1, len, rc, n.y);
            ok = ok && div_fast(-X2, len, rc, n.z);
            n.x = 0.0;
            if (!ok) {
                n.x = div_(0.0, len);
                n.y = div_(-X1, len);
                n.z = div_(-X2, len);
            }
            reflect(r0, s, d, n, p.eps);
            step_l = sub_(step_l, add_(d, p.eps));
        } else {
            check = false;
        }
    }
    s = matvec3(p.Rinv, s);
    r0 = matvec3(p.Rinv, r0);
    pos.x = fma_(step_l, s.x, r0.x);
    pos.y = fma_(step_l, s.y, r0.y);
    pos.z = fma_(step_l, s.z, r0.z);
    return iter >= p.max_iter;
}

// simulations.py:819-875
template <>
__device__ __forceinline__ bool walker_step<3>(Vec3 &pos, Rng &rng, const KParams &p, const double *tab,
                                               const bool live)
{
    if (!live) return false;
    Vec3 s = random_step(rng, tab);
    Vec3 r0 = matvec3(p.R, pos);
    double step_l = p.step_l;
    long long iter = 0;
    bool check = true;
    while (check && step_l > 0 && iter < p.max_iter) {
        ++iter;
        double d = line_ellipsoid(r0, s, p.ax);
        if (d > 0 && d < step_l) {
            Vec3 n;
            n.x = div_(-fma_(d, s.x, r0.x), mul_(p.ax[0], p.ax[0]));
            n.y = div_(-fma_(d, s.y, r0.y), mul_(p.ax[1], p.ax[1]));
            n.z = div_(-fma_(d, s.z, r0.z), mul_(p.ax[2], p.ax[2]));
            n = normalize3(n);
            reflect(r0, s, d, n, p.eps);
            step_l = sub_(step_l, add_(d, p.eps));
        } else {
            check = false;
        }
    }
    s = matvec3(p.Rinv, s);
    r0 = matvec3(p.Rinv, r0);
    pos.x = fma_(step_l, s.x, r0.x);
    pos.y = fma_(step_l, s.y, r0.y);
    pos.z = fma_(step_l, s.z, r0.z);
    return iter >= p.max_iter;
}

__device__ __forceinline__ Tri load_tri(const double *tri9, int id)
{
    const double *q = tri9 + 9ll * id;
    Tri t;
    t.A.x = __ldg(q + 0); t.A.y = __ldg(q + 1); t.A.z = __ldg(q + 2);
    t.E1.x = __ldg(q + 3); t.E1.y = __ldg(q + 4); t.E1.z = __ldg(q + 5);
    t.E2.x = __ldg(q + 6); t.E2.y = __ldg(q + 7); t.E2.z = __ldg(q + 8);
    return t;
}

// ---- mesh: which grid cells does the remaining step segment overlap (simulations.py:929-934)

// floor(x / V) with the rounding of the reference's div.rn + floor.  A reciprocal multiply is
// within 4e-16 relative of the IEEE quotient, so whenever its fractional part is not within 1e-9
// of an integer the two floors agree; otherwise the real division decides.
__device__ __forceinline__ double floor_quotient(double x, double V, double invV)
{
    double qa = x * invV;
    double fl = floor(qa);
    double fr = qa - fl;
    if (fr > 1e-9 && fr < 1.0 - 1e-9 && fabs(qa) < 1e5) return fl;
    return floor(div_(x, V));
}

// One end of the periodic cell range of an axis: the period index nq = floor(x / V) and the
// cell index inside the base voxel (simulations.py:654-679).  upper == false: lower limit
// (_ll_subvoxel_overlap), true: exclusive upper limit (_ul_subvoxel_overlap).
__device__ __forceinline__ void axis_limit(const double *xs, int len, double V, double invV, double inv_h,
                                           double x, bool upper, double &nq, int &base)
{
    nq = floor_quotient(x, V, invV);
    double sh = fma_(-V, nq, x);
    base = upper ? ul_overlap(xs, len, sh, inv_h) : ll_overlap(xs, len, sh, inv_h);
}

struct AxisCells {
    double nq;       // period index of the first cell
    int base;        // its index inside the base voxel (may equal n: first cell of the next image)
    long long count; // number of cells, <= 0 when the range is empty
};

__device__ __forceinline__ AxisCells axis_cells(const double *xs, int len, double V, double invV, double inv_h,
                                                double a, double b)
{
    AxisCells r;
    double nq_hi;
    int base_hi;
    axis_limit(xs, len, V, invV, inv_h, fmin(a, b), false, r.nq, r.base);
    axis_limit(xs, len, V, invV, inv_h, fmax(a, b), true, nq_hi, base_hi);
    // global indices formed in FP64 like the reference (exact integers)
    long long lo = __double2ll_rz(fma_(r.nq, (double)(len - 1), (double)r.base));
    long long hi = __double2ll_rz(fma_(nq_hi, (double)(len - 1), (double)base_hi));
    r.count = hi - lo;
    return r;
}

// k-th cell of an axis range -> index inside the base voxel and periodic image number
// (what simulations.py:940-945 derives with floor(x / (len - 1)))
__device__ __forceinline__ void axis_cell(const AxisCells &r, int n, long long k, int &cell, double &image)
{
    long long j = (long long)r.base + k;
    long long q = j / n;  // base >= 0, k >= 0
    cell = (int)(j - q * n);
    image = r.nq + (double)q;
}

constexpr int kCoopCells = 8;  // cells per walker and collision iteration tested cooperatively

// Per-warp scratch: every lane publishes its ray and its short list of grid cells, then the 32
// lanes share ALL listed (cell, triangle) tests of the warp evenly.
struct MeshScratch {
    double ray[6][32];            // position, unit step
    int begin[kCoopCells][32];    // first entry of the cell in tri_idx
    int cum[kCoopCells][32];      // inclusive running number of entries over the lane's cells
    int image[3][kCoopCells][32]; // periodic image number of the cell per axis
};

// Closest triangle hit (d > 0) over every triangle listed in the cells the segment
// [pos, pos + step_l * s] overlaps, visited in the reference's order (cells x -> y -> z, entries
// ascending, strict "<" so the first minimum wins): simulations.py:936-983.
__device__ __forceinline__ void mesh_closest_hit(const MeshDev &g, MeshScratch &sc, const int lane, const bool need,
                                                 const Vec3 &pos, const Vec3 &s, const double step_l,
                                                 double &min_d, int &closest)
{
    const unsigned full = 0xffffffffu;
    const double inf = __longlong_as_double(0x7FF0000000000000LL);
    min_d = inf;
    int n_items = 0;
    bool solo = false;
    AxisCells ax, ay, az;
    if (need) {
        // end point of the remaining segment: x uses a separately rounded product, y and z are
        // fused (that is how the reference's kernel was compiled)
        double ex = add_(pos.x, mul_(step_l, s.x));
        double ey = fma_(step_l, s.y, pos.y);
        double ez = fma_(step_l, s.z, pos.z);
        ax = axis_cells(g.xs, g.len_xs, g.vox[0], g.inv_vox[0], g.inv_hx, pos.x, ex);
        ay = axis_cells(g.ys, g.len_ys, g.vox[1], g.inv_vox[1], g.inv_hy, pos.y, ey);
        az = axis_cells(g.zs, g.len_zs, g.vox[2], g.inv_vox[2], g.inv_hz, pos.z, ez);
        if (ax.count > 0 && ay.count > 0 && az.count > 0) {
            const bool small = ax.count <= kCoopCells && ay.count <= kCoopCells && az.count <= kCoopCells &&
                               ax.count * ay.count * az.count <= kCoopCells && fabs(ax.nq) < 1e9 &&
                               fabs(ay.nq) < 1e9 && fabs(az.nq) < 1e9;
            if (small) {
                int c = 0;
                for (int ix = 0; ix < (int)ax.count; ++ix) {
                    int cx;
                    double mx;
                    axis_cell(ax, g.len_xs - 1, ix, cx, mx);
                    for (int iy = 0; iy < (int)ay.count; ++iy) {
                        int cy;
                        double my;
                        axis_cell(ay, g.len_ys - 1, iy, cy, my);
                        for (int iz = 0; iz < (int)az.count; ++iz) {
                            int cz;
                            double mz;
                            axis_cell(az, g.len_zs - 1, iz, cz, mz);
                            int2 r = __ldg(g.cell_rng + ((long long)cx * g.nsv1 + cy) * g.nsv2 + cz);
                            n_items += r.y - r.x;
                            sc.begin[c][lane] = r.x;
                            sc.cum[c][lane] = n_items;
                            sc.image[0][c][lane] = (int)mx;
                            sc.image[1][c][lane] = (int)my;
                            sc.image[2][c][lane] = (int)mz;
                            ++c;
                        }
                    }
                }
                sc.ray[0][lane] = pos.x; sc.ray[1][lane] = pos.y; sc.ray[2][lane] = pos.z;
                sc.ray[3][lane] = s.x; sc.ray[4][lane] = s.y; sc.ray[5][lane] = s.z;
            } else {
                solo = true;
            }
        }
    }
    __syncwarp();

    // exclusive prefix sum of the lanes' item counts
    int incl = n_items;
#pragma unroll
    for (int st = 1; st < 32; st <<= 1) {
        int v = __shfl_up_sync(full, incl, st);
        if (lane >= st) incl += v;
    }
    const int excl = incl - n_items;
    const int total = __shfl_sync(full, incl, 31);

    for (int base = 0; base < total; base += 32) {
        const int j = base + lane;
        // owner = last lane whose first item index is <= j
        int owner = 0;
#pragma unroll
        for (int st = 16; st > 0; st >>= 1) {
            int cand = owner + st;
            int first = __shfl_sync(full, excl, cand & 31);
            if (cand < 32 && first <= j) owner = cand;
        }
        const int owner_first = __shfl_sync(full, excl, owner);
        double d = inf;
        int id = -1;
        if (j < total) {
            const int i = j - owner_first;
            int c = 0;
            while (i >= sc.cum[c][owner]) ++c;
            const int entry = sc.begin[c][owner] + i - (c ? sc.cum[c - 1][owner] : 0);
            id = __ldg(g.tri_idx + entry);
            const Tri tr = load_tri(g.tri9, id);
            const int mx = sc.image[0][c][owner], my = sc.image[1][c][owner], mz = sc.image[2][c][owner];
            Vec3 o, dir;
            // walker moved into the base voxel: r0 - shift_n * xs[-1] (simulations.py:943, 970-971)
            o.x = sub_(sc.ray[0][owner], mx == 0 ? 0.0 : mul_((double)mx, g.top[0]));
            o.y = sub_(sc.ray[1][owner], my == 0 ? 0.0 : mul_((double)my, g.top[1]));
            o.z = sub_(sc.ray[2][owner], mz == 0 ? 0.0 : mul_((double)mz, g.top[2]));
            dir.x = sc.ray[3][owner]; dir.y = sc.ray[4][owner]; dir.z = sc.ray[5][owner];
            const double t = ray_triangle(tr, o, dir);
            if (t > 0) d = t;
        }
        // segmented running minimum over lanes that serve the same owner; on ties the earlier
        // item (lower lane) wins, like the reference's strict "<" in visiting order
#pragma unroll
        for (int st = 1; st < 32; st <<= 1) {
            double dp = __shfl_up_sync(full, d, st);
            int ip = __shfl_up_sync(full, id, st);
            int op = __shfl_up_sync(full, owner, st);
            if (lane >= st && op == owner && dp <= d) {
                d = dp;
                id = ip;
            }
        }
        // every owner with items in this round takes the result of its segment's last lane
        const int lo = max(excl, base), hi = min(excl + n_items, base + 32);
        const bool mine = lo < hi;
        const int src = mine ? hi - 1 - base : lane;
        const double ds = __shfl_sync(full, d, src);
        const int is = __shfl_sync(full, id, src);
        if (mine && ds < min_d) {
            min_d = ds;
            closest = is;
        }
    }
    __syncwarp();

    if (solo) {  // more cells than the shared table holds: this lane walks its own cells
        for (long long ix = 0; ix < ax.count; ++ix) {
            int cx;
            double mx;
            axis_cell(ax, g.len_xs - 1, ix, cx, mx);
            const double tx = sub_(pos.x, mx == 0.0 ? 0.0 : mul_(mx, g.top[0]));
            for (long long iy = 0; iy < ay.count; ++iy) {
                int cy;
                double my;
                axis_cell(ay, g.len_ys - 1, iy, cy, my);
                const double ty = sub_(pos.y, my == 0.0 ? 0.0 : mul_(my, g.top[1]));
                for (long long iz = 0; iz < az.count; ++iz) {
                    int cz;
                    double mz;
                    axis_cell(az, g.len_zs - 1, iz, cz, mz);
                    const double tz = sub_(pos.z, mz == 0.0 ? 0.0 : mul_(mz, g.top[2]));
                    const Vec3 tr0 = {tx, ty, tz};
                    const int2 r = __ldg(g.cell_rng + ((long long)cx * g.nsv1 + cy) * g.nsv2 + cz);
                    for (int i = r.x; i < r.y; ++i) {
                        const int id = __ldg(g.tri_idx + i);
                        const double d = ray_triangle(load_tri(g.tri9, id), tr0, s);
                        if (d > 0 && d < min_d) {
                            closest = id;
                            min_d = d;
                        }
                    }
                }
            }
        }
    }
}

// simulations.py:878-1013.  Warp-synchronous: every lane of the warp must call it (live == false
// for lanes without a walker); the collision search of each iteration is shared by the warp.
template <>
__device__ __forceinline__ bool walker_step<4>(Vec3 &pos, Rng &rng, const KParams &p, const double *tab,
                                               const bool live)
{
    __shared__ MeshScratch s_scratch[kBlock / 32];
    MeshScratch &sc = s_scratch[threadIdx.x >> 5];
    const int lane = threadIdx.x & 31;
    const MeshDev &g = p.mesh;
    Vec3 s = random_step(rng, tab);
    double step_l = p.step_l;
    long long iter = 0;
    bool check = live;
    int closest = 0;
    for (;;) {
        const bool need = check && step_l > 0 && iter < p.max_iter;
        if (!__any_sync(0xffffffffu, need)) break;
        if (need) ++iter;
        double min_d;
        mesh_closest_hit(g, sc, lane, need, pos, s, step_l, min_d, closest);
        if (need) {
            if (min_d > step_l) {
                check = false;
            } else {
                double u = u01_f64(rng_next(rng));
                Vec3 n = triangle_normal(load_tri(g.tri9, closest));
                if (g.perm_prob < u)
                    reflect(pos, s, min_d, n, p.eps);
                else
                    cross_membrane(pos, s, min_d, n, p.eps);
                step_l = sub_(step_l, min_d);
            }
        }
    }
    pos.x = fma_(step_l, s.x, pos.x);
    pos.y = fma_(step_l, s.y, pos.y);
    pos.z = fma_(step_l, s.z, pos.z);
    return iter >= p.max_iter;
}

// ---------------------------------------------------------------- block reduction of the signal

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}

// phase(m) -> per-block sum of cos(phase) over walkers with a clear iter_exc flag, written to
// partials[m * gridDim.x + blockIdx.x]; row n_meas receives the number of such walkers.
// Fixed summation tree: results do not depend on scheduling.
template <typename PhaseFn>
__device__ __forceinline__ void block_signal(const KParams &p, bool valid, PhaseFn phase_of)
{
    __shared__ double s_part[kBlock / 32][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int rows = p.n_meas + 1;
    for (int m0 = 0; m0 < rows; m0 += 32) {
        int mt = min(32, rows - m0);
        for (int j = 0; j < mt; ++j) {
            int m = m0 + j;
            double v = 0.0;
            if (valid) v = (m < p.n_meas) ? cos(phase_of(m)) : 1.0;
            v = warp_sum(v);
            if (lane == 0) s_part[warp][j] = v;
        }
        __syncthreads();
        if (warp == 0 && lane < mt) {
            double acc = 0.0;
#pragma unroll
            for (int w = 0; w < kBlock / 32; ++w) acc += s_part[w][lane];
            p.partials[(long long)(m0 + lane) * gridDim.x + blockIdx.x] = acc;
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------- the walk kernel

// MR > 0: n_meas == MR phases in registers.  MR == 0: any n_meas; positions of kTimeChunk steps
// are buffered in registers, then each measurement's phase makes one round trip through its
// (coalesced, L2-resident) row of `phases` per chunk instead of one per step.
template <int SUB, int MR>
__global__ void __launch_bounds__(kBlock, SUB == 4 ? DSB_MESH_MIN_BLOCKS : DSB_MIN_BLOCKS) walk_kernel(const __grid_constant__ KParams p)
{
    __shared__ double s_tab[16];
    if (threadIdx.x < 16) s_tab[threadIdx.x] = __longlong_as_double((long long)c_sincos_tab[threadIdx.x]);
    __syncthreads();

    const long long w = (long long)blockIdx.x * kBlock + threadIdx.x;
    const bool active = w < p.n_walkers;
    const long long N = p.n_walkers;
    Vec3 pos = {0.0, 0.0, 0.0};
    Rng rng = {1ull, 1ull};
    bool exc = false;
    if (active) {
        pos.x = p.pos[3 * w];
        pos.y = p.pos[3 * w + 1];
        pos.z = p.pos[3 * w + 2];
        ulonglong2 st = reinterpret_cast<const ulonglong2 *>(p.rng)[w];
        rng.s0 = st.x;
        rng.s1 = st.y;
    }

    if constexpr (MR > 0) {
        double ph[MR];
#pragma unroll
        for (int m = 0; m < MR; ++m) ph[m] = (active && p.t0 > 0) ? p.phases[(long long)m * N + w] : 0.0;
        // the time loop is uniform over the block (the mesh step shares work inside each warp)
        for (int t = p.t0; t < p.t1; ++t) {
            exc |= walker_step<SUB>(pos, rng, p, s_tab, active);
#pragma unroll
            for (int m = 0; m < MR; ++m) {
                const double *g = p.grad + ((long long)m * p.n_t + t) * 3;
                double gx = __ldg(g), gy = __ldg(g + 1), gz = __ldg(g + 2);
                ph[m] = fma_(p.gamma_dt, fma_(gz, pos.z, fma_(gx, pos.x, mul_(gy, pos.y))), ph[m]);
            }
        }
        if (active) {
#pragma unroll
            for (int m = 0; m < MR; ++m) p.phases[(long long)m * N + w] = ph[m];
        }
        if (active) {
            if (exc) p.iter_exc[w] = 1;
            else exc = p.iter_exc[w] != 0;
        }
        if (p.finalize)
            block_signal(p, active && !exc, [&](int m) {
                double v = 0.0;
#pragma unroll
                for (int k = 0; k < MR; ++k)
                    if (k == m) v = ph[k];
                return v;
            });
    } else {
        if (active && p.t0 == 0)
            for (int m = 0; m < p.n_meas; ++m) p.phases[(long long)m * N + w] = 0.0;
        Vec3 buf[kTimeChunk];
        for (int t = p.t0; t < p.t1; t += kTimeChunk) {
            const int cnt = min(kTimeChunk, p.t1 - t);
#pragma unroll
            for (int k = 0; k < kTimeChunk; ++k)
                if (k < cnt) {
                    exc |= walker_step<SUB>(pos, rng, p, s_tab, active);
                    buf[k] = pos;
                }
            if (active)
                for (int m = 0; m < p.n_meas; ++m) {
                    double *row = p.phases + (long long)m * N + w;
                    double a = *row;
                    const double *g = p.grad + ((long long)m * p.n_t + t) * 3;
#pragma unroll
                    for (int k = 0; k < kTimeChunk; ++k)
                        if (k < cnt) {
                            double gx = __ldg(g + 3 * k), gy = __ldg(g + 3 * k + 1), gz = __ldg(g + 3 * k + 2);
                            a = fma_(p.gamma_dt, fma_(gz, buf[k].z, fma_(gx, buf[k].x, mul_(gy, buf[k].y))), a);
                        }
                    *row = a;
                }
        }
        if (active) {
            if (exc) p.iter_exc[w] = 1;
            else exc = p.iter_exc[w] != 0;
        }
        if (p.finalize)
            block_signal(p, active && !exc, [&](int m) { return p.phases[(long long)m * N + w]; });
    }

    if (active) {
        p.pos[3 * w] = pos.x;
        p.pos[3 * w + 1] = pos.y;
        p.pos[3 * w + 2] = pos.z;
        reinterpret_cast<ulonglong2 *>(p.rng)[w] = make_ulonglong2(rng.s0, rng.s1);
    }
}

// Sums the per-block partials of one measurement (or of the valid-walker count) in a fixed
// order: thread j takes blocks j, j + 256, ...; then a shared-memory tree.
__global__ void __launch_bounds__(256) reduce_partials_kernel(const double *partials, int n_blocks, double *out)
{
    __shared__ double s[256];
    const double *row = partials + (long long)blockIdx.x * n_blocks;
    double acc = 0.0;
    for (int b = threadIdx.x; b < n_blocks; b += 256) acc += row[b];
    s[threadIdx.x] = acc;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[blockIdx.x] = s[0];
}

// ---------------------------------------------------------------- FP64 issue-rate probe

constexpr int kPeakChains = 8;

// kPeakChains independent DFMA chains per thread; nothing else in the loop.  Its rate is the
// FP64 pipe's ceiling that bench.py's roofline is quoted against.
__global__ void __launch_bounds__(256) fp64_peak_kernel(double *out, int iters, double seed)
{
    double a[kPeakChains];
#pragma unroll
    for (int k = 0; k < kPeakChains; ++k) a[k] = seed + k + threadIdx.x * 1e-6;
    const double m = 1.0000001, c = 1e-7;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < kPeakChains; ++k) a[k] = __fma_rn(a[k], m, c);
    }
    double sum = 0.0;
#pragma unroll
    for (int k = 0; k < kPeakChains; ++k) sum += a[k];
    out[(long long)blockIdx.x * blockDim.x + threadIdx.x] = sum;
}

// ---------------------------------------------------------------- RNG state derivation

// state[i] = J^(start + i) * splitmix64(seed), J = the 2^64-step jump of xoroshiro128+ as a
// 128x128 matrix over GF(2).  pows[k] holds J^(2^k) column by column (column b = image of unit
// vector b); applying the matrices for the set bits of the index reproduces numba's sequential
// jump chain (numba/cuda/random.py:102-126, 225-241) without the O(N) host loop.
__global__ void __launch_bounds__(256) rng_init_kernel(unsigned long long z, unsigned long long start,
                                                       long long n, const ulonglong2 *pows,
                                                       ulonglong2 *out)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned long long idx = start + (unsigned long long)i;
    unsigned long long s0 = z, s1 = z;
    for (int k = 0; k < 64 && (idx >> k) != 0; ++k) {
        if (((idx >> k) & 1ull) == 0) continue;
        const ulonglong2 *col = pows + 128 * k;
        unsigned long long a0 = 0, a1 = 0;
#pragma unroll 8
        for (int b = 0; b < 64; ++b) {
            ulonglong2 c = __ldg(col + b);
            unsigned long long mask = 0ull - ((s0 >> b) & 1ull);
            a0 ^= c.x & mask;
            a1 ^= c.y & mask;
        }
#pragma unroll 8
        for (int b = 0; b < 64; ++b) {
            ulonglong2 c = __ldg(col + 64 + b);
            unsigned long long mask = 0ull - ((s1 >> b) & 1ull);
            a0 ^= c.x & mask;
            a1 ^= c.y & mask;
        }
        s0 = a0;
        s1 = a1;
    }
    out[i] = make_ulonglong2(s0, s1);
}

}  // namespace dsb
