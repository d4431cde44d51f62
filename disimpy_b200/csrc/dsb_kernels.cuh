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
// Blocks per SM the register-phase kernels of the free and analytic substrates are compiled for.  Left
// alone, the step generator's straight-line code (round 2: no branches around square roots and
// divisions any more) lets ptxas spend 94-120 registers on it; capped at 6 blocks (80 registers; the
// ellipsoid: 7, 72 registers) it is 2-12 % faster (profiles/r02_j_kbench_sqrt_fast.txt).
#ifndef DSB_MIN_BLOCKS
#define DSB_MIN_BLOCKS 6
#endif
#ifndef DSB_MIN_BLOCKS_ELLIPSOID
#define DSB_MIN_BLOCKS_ELLIPSOID 7
#endif
#ifndef DSB_MR0_MIN_BLOCKS
#define DSB_MR0_MIN_BLOCKS 4
#endif
#ifndef DSB_MESH_MIN_BLOCKS
#define DSB_MESH_MIN_BLOCKS 4
#endif
constexpr int kBlock = DSB_BLOCK;    // threads (= walkers) per CTA
constexpr int kMaxRegMeas = 4;       // measurements whose phase lives in registers
constexpr int kMaxRank = 16;         // largest rank of a protocol walked through virtual measurements
// Steps per chunk of the many-measurement kernels (n_meas > kMaxRegMeas): every phase makes one
// round trip through HBM per chunk, so longer chunks mean less traffic (16 n_meas / chunk bytes
// per walker-step); the mesh kernel's shared memory leaves room for 8 steps only.
#ifndef DSB_CHUNK
#define DSB_CHUNK 16   // analytic substrates (a multiple of 4)
#endif
template <int SUB>
struct ChunkSteps {
    static constexpr int value = SUB == 4 ? 8 : DSB_CHUNK;
};
__host__ __device__ constexpr int chunk_steps(int substrate) { return substrate == 4 ? 8 : DSB_CHUNK; }
__host__ __device__ constexpr int grad_row_len(int chunk) { return 3 * chunk + 4; }  // + padding (bank spread)
#ifndef DSB_GRADROWS
#define DSB_GRADROWS 8
#endif
constexpr int kGradRows = DSB_GRADROWS;  // measurements per gradient tile staged in shared memory (a multiple of 8)

// The grid the cooperative collision search walks: the reference's subvoxel grid, or a refinement of
// it (every cell cut into kx x ky x kz sub-cells whose lists hold the parent's triangles that reach
// into the sub-cell).  Same layout as the reference grid's arrays in MeshDev.
struct SearchGrid {
    const uint4 *entry;     // per list entry: triangle id + box, as MeshDev::entry
    const int2 *cell_rng;   // (cells,) [begin, end) into entry
    const double *xs, *ys, *zs;
    int len_xs, len_ys, len_zs;
    int nsv1, nsv2;
    double inv_hx, inv_hy, inv_hz;  // guesses only
    double margin[3];   // a segment end closer than this to a cell boundary sends its walker to the per-lane search
};

struct MeshDev {
    const double *tri;      // (n_faces, kTriStride): A, B-A, C-A, pad
    const int *tri_idx;     // (K,) triangle ids, cell after cell (reference order)
    const double *normal;   // (n_faces, 3) unit normals, computed once with the walk's own arithmetic (tri_normal_kernel)
    const uint4 *entry;     // (K,) per list entry: triangle id, box (lo, 32767 - hi) on a 15-bit grid, 3 x 2 halfwords
    const int2 *cell_rng;   // (n_cells,) [begin, end) into tri_idx
    const double *xs, *ys, *zs;
    int len_xs, len_ys, len_zs;
    int nsv1, nsv2;
    double inv_hx, inv_hy, inv_hz;  // guesses only, never part of a result
    double vox[3];      // |xs[-1] - xs[0]| per axis, the period of the lookup (simulations.py:660)
    double inv_vox[3];  // 1 / vox, guess only
    double top[3];      // xs[-1], ys[-1], zs[-1]: the image shift unit (simulations.py:943)
    double qscale[3];   // 32767 / top: the 15-bit grid of the box filter
    double perm_prob;
    SearchGrid fine;    // what mesh_closest_hit walks (the arrays above when the grid is not refined)
};

struct KParams {
    long long n_walkers;        // walkers of the handle (row length of `phases`)
    long long w_begin, w_end;   // walkers this launch advances; w_begin is a multiple of kBlock
    int n_blocks_total;         // blocks covering all n_walkers (row length of `partials`)
    int n_meas, n_t;
    int t0, t1;
    int finalize;  // t1 == n_t: also emit per-block sum(cos phase) partials
    int max_iter;  // clamped to INT_MAX by the host
    double step_l, gamma_dt, eps, radius;
    double R[9], Rinv[9];
    EllipsoidConsts ell;
    const double *grad;         // (n_meas, n_t, 3)
    const double *grad_chunked; // (ceil(n_t / chunk), n_meas, grad_row_len(chunk)): gamma dt g, zero padded
    double *pos;                // (n_walkers, 3)
    unsigned long long *rng;    // (n_walkers, 2)
    double *phases;             // (n_meas, n_walkers)
    unsigned char *iter_exc;    // (n_walkers,)
    double *partials;           // (n_meas + 1, n_blocks_total)
    const int *order;           // mesh walk only, may be null: thread j of the launch advances walker order[j] (cell order)
    MeshDev mesh;
};

// ---------------------------------------------------------------- one time step, per substrate

// simulations.py:682-702
__device__ __forceinline__ bool free_step(Vec3 &pos, const Vec3 &s, const KParams &p, const bool live)
{
    if (!live) return false;
    pos.x = fma_(s.x, p.step_l, pos.x);
    pos.y = fma_(s.y, p.step_l, pos.y);
    pos.z = fma_(s.z, p.step_l, pos.z);
    return false;
}

// ---- analytic substrates (sphere, cylinder, ellipsoid): a time step taken apart into the pieces
// the collision loop of the reference is made of, so that the time loop can either run them
// back to back (walker_step) or park colliding walkers and bounce them in batches (walk_kernel).

struct Flight {      // one time step in progress
    Vec3 s, r0;      // unit step and position, in the substrate's frame
    double step_l;   // length still to travel
    double d;        // distance to the wall found by the last probe
    int iter;        // intersection checks made in this step
};

// new random direction; position into the substrate frame (simulations.py:722-728, 778-787,
// 838-847).  The step is not rotated in, like the reference.
template <int SUB>
__device__ __forceinline__ void begin_step(const Vec3 &pos, const Vec3 &unit, const KParams &p, Flight &f)
{
    f.s = unit;
    if constexpr (SUB == 2 || SUB == 3) f.r0 = matvec3(p.R, pos);
    else f.r0 = pos;
    f.step_l = p.step_l;
    f.iter = 0;
}

// Loop guard + intersection check of the reference's while loop (simulations.py:730-733,
// 789-792, 849-852): true when the walker hits the wall within the remaining length (f.d = the
// distance), false when the step can be completed.
template <int SUB>
__device__ __forceinline__ bool probe(Flight &f, const KParams &p)
{
    if (!(f.step_l > 0) || f.iter >= p.max_iter) return false;
    ++f.iter;
    if constexpr (SUB == 1) f.d = line_sphere(f.r0, f.s, p.radius);
    else if constexpr (SUB == 2) f.d = line_circle(f.r0, f.s, p.radius);
    else f.d = line_ellipsoid(f.r0, f.s, p.ell);
    return f.d > 0 && f.d < f.step_l;
}

// specular reflection at the wall f.d ahead (simulations.py:734-741, 793-801, 853-860)
template <int SUB>
__device__ __forceinline__ void bounce(Flight &f, const KParams &p)
{
    Vec3 n;
    if constexpr (SUB == 1) {
        n.x = -fma_(f.d, f.s.x, f.r0.x);
        n.y = -fma_(f.d, f.s.y, f.r0.y);
        n.z = -fma_(f.d, f.s.z, f.r0.z);
        n = normalize3(n);
    } else if constexpr (SUB == 2) {
        double X1 = fma_(f.d, f.s.y, f.r0.y), X2 = fma_(f.d, f.s.z, f.r0.z);
        double len = sqrt_(fma_(X2, X2, fma_(X1, X1, 0.0)));
        // normal = (0, -X1, -X2) / len; 0 / len is +0 for every finite positive len
        double rc = rcp_refined(len);
        bool ok = len > 0 && div_fast(-X1, len, rc, n.y);
        ok = ok && div_fast(-X2, len, rc, n.z);
        n.x = 0.0;
        if (!ok) {
            n.x = div_(0.0, len);
            n.y = div_(-X1, len);
            n.z = div_(-X2, len);
        }
    } else {
        const double nx = -fma_(f.d, f.s.x, f.r0.x), ny = -fma_(f.d, f.s.y, f.r0.y), nz = -fma_(f.d, f.s.z, f.r0.z);
        bool ok = div_fast(nx, p.ell.axsq[0], p.ell.axsq_rc[0], n.x);
        ok &= div_fast(ny, p.ell.axsq[1], p.ell.axsq_rc[1], n.y);
        ok &= div_fast(nz, p.ell.axsq[2], p.ell.axsq_rc[2], n.z);
        if (!ok) {
            n.x = div_(nx, p.ell.axsq[0]);
            n.y = div_(ny, p.ell.axsq[1]);
            n.z = div_(nz, p.ell.axsq[2]);
        }
        n = normalize3(n);
    }
    reflect(f.r0, f.s, f.d, n, p.eps);
    f.step_l = sub_(f.step_l, add_(f.d, p.eps));
}

// back to the lab frame and the final move (simulations.py:742-745, 802-805, 861-864); returns
// the iter_exc flag of the step.  The round trip through the substrate frame is kept: its
// rounding is part of the reference's trajectory.
template <int SUB>
__device__ __forceinline__ bool end_step(Vec3 &pos, Flight &f, const KParams &p)
{
    if constexpr (SUB == 2 || SUB == 3) {
        f.s = matvec3(p.Rinv, f.s);
        f.r0 = matvec3(p.Rinv, f.r0);
    }
    pos.x = fma_(f.step_l, f.s.x, f.r0.x);
    pos.y = fma_(f.step_l, f.s.y, f.r0.y);
    pos.z = fma_(f.step_l, f.s.z, f.r0.z);
    return f.iter >= p.max_iter;
}

// simulations.py:705-756 (sphere), :759-816 (cylinder), :819-875 (ellipsoid)
template <int SUB>
__device__ __forceinline__ bool walker_step(Vec3 &pos, const Vec3 &unit, const KParams &p, const bool live)
{
    static_assert(SUB >= 1 && SUB <= 3, "analytic substrates only");
    if (!live) return false;
    Flight f;
    begin_step<SUB>(pos, unit, p, f);
    while (probe<SUB>(f, p)) bounce<SUB>(f, p);
    return end_step<SUB>(pos, f, p);
}

constexpr int kTriStride = 10;  // doubles per triangle record: A, B-A, C-A, pad (five 16-byte loads)

__device__ __forceinline__ Tri load_tri(const double *tri, int id)
{
    const double2 *q = reinterpret_cast<const double2 *>(tri + (long long)kTriStride * id);
    const double2 a = __ldg(q), b = __ldg(q + 1), c = __ldg(q + 2), d = __ldg(q + 3), e = __ldg(q + 4);
    Tri t;
    t.A.x = a.x; t.A.y = a.y; t.A.z = b.x;
    t.E1.x = b.y; t.E1.y = c.x; t.E1.z = c.y;
    t.E2.x = d.x; t.E2.y = d.y; t.E2.z = e.x;
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

// The common case without divisions and scans: x lies strictly inside a grid cell and its
// period index is unambiguous.  Then both _ll_ and _ul_subvoxel_overlap_periodic follow from the
// cell (ll = its global index, ul = that + 1).  Returns false when x is on (or within rounding
// of) a cell or period boundary; the caller then uses the generic lookups above.
__device__ __forceinline__ bool cell_of(const double *xs, int n, double V, double invV, double inv_h, double margin,
                                        double x, int &cell, int &image)
{
    const double q = x * invV;
    const double fl = floor(q);
    const double fr = q - fl;
    const double sh = fma_(-V, fl, x);  // the reference's xmin_shifted, given floor(x / V) == fl
    const int c = min(max(__double2int_rz(sh * inv_h), 0), n - 1);
    const double lo = __ldg(xs + c), hi = __ldg(xs + c + 1);
    cell = c;
    image = __double2int_rz(fl);
    return fr > 1e-9 && fr < 1.0 - 1e-9 && fabs(q) < 1e5 && sh - lo > margin && hi - sh > margin;
}

struct AxisSpan {
    int cell;   // first overlapped cell, index inside the base voxel
    int image;  // its periodic image number
    int count;  // cells overlapped (>= 1)
};

#ifndef DSB_SPAN
#define DSB_SPAN 3
#endif
constexpr int kMaxSpan = DSB_SPAN;  // cells per axis the cooperative search takes

// cells overlapped by [a, b] (a = walker, b = end of the remaining step) along one axis
__device__ __forceinline__ bool axis_span(const double *xs, int n, double V, double invV, double inv_h, double margin,
                                          double a, double b, AxisSpan &sp)
{
    int ca, ia, cb, ib;
    const bool oka = cell_of(xs, n, V, invV, inv_h, margin, a, ca, ia);
    const bool okb = cell_of(xs, n, V, invV, inv_h, margin, b, cb, ib);
    const bool a_first = !(b < a);
    sp.cell = a_first ? ca : cb;
    sp.image = a_first ? ia : ib;
    const long long d = (long long)(ib - ia) * n + (cb - ca);
    const long long cnt = (a_first ? d : -d) + 1;
    sp.count = (int)cnt;
    return oka && okb && cnt >= 1 && cnt <= kMaxSpan && cnt <= n;
}

constexpr int kSurvivorCap = 256;  // triangles per warp and search that pass the box filter
constexpr int kRangeCap = 128;     // non-empty cell lists per warp and search
constexpr int kEntryCap = 4096;    // list entries per warp and search
constexpr unsigned kSwarH = 0x80008000u;
// bit 31 of a list entry's triangle word: the triangle is (nearly) edge-on to the +x axis, the
// initial-position sampler must not trust its box (dsb_fill.cuh)
constexpr unsigned kEntryTriMask = 0x7fffffffu;

// Per-warp scratch of the collision search.
struct MeshScratch {
    double dir[3][32];     // unit step of each lane's walker
    double org[3][2][32];  // its position moved into the base voxel: image of the first cell, next image
    unsigned long long best_d[32];  // closest hit distance per walker (bit pattern of a positive double)
    int best_tri[32];               // smallest ...
    int best_tri_hi[32];            // ... and largest id among the triangles at that distance (different: a tie)
    union {                         // (the hit distances are written after the last use of the boxes)
        uint4 range_box[kRangeCap];     // segment box for the filter (3 words); owner lane | image flags << 8
        double hit[kSurvivorCap];       // distance found for the survivor (inf: none)
    };
    int2 range_pos[kRangeCap];      // number of the range's first entry in the warp's flat numbering; its place in the list
    unsigned start_bits[kEntryCap / 32 + 1];    // bit j: a range starts at flat entry j (+ a word of padding)
    unsigned long long survivor[kSurvivorCap];  // triangle (32) | owner lane (8) | image flags (3) << 8
};
static_assert(kSurvivorCap * sizeof(double) <= kRangeCap * sizeof(uint4), "hit[] must fit over range_box[]");

// 15-bit grid coordinate of x on an axis of the base voxel (scale = 32767 / xs[-1]); unclamped
__device__ __forceinline__ int quantize(double x, double scale) { return __double2int_rd(x * scale); }
__device__ __forceinline__ unsigned clamp15(int q) { return (unsigned)min(max(q, 0), 32767); }

// Cells per walker and search the cooperative path takes (template parameter of the mesh
// kernels): 12 in general; 8 when the step is shorter than the smallest grid spacing, so that a
// segment never overlaps more than 2 cells per axis (the host picks, launch_walk).
constexpr int kMaxCells = 12, kMaxCellsShortStep = 8;

// c-th cell (visiting order x -> y -> z) of the spans: place of its list range, image flags
struct CellWalk {
    int ix = 0, iy = 0, iz = 0;
    __device__ __forceinline__ void next(const AxisSpan &sy, const AxisSpan &sz)
    {
        const bool wz = ++iz == sz.count;
        iz = wz ? 0 : iz;
        iy += wz;
        const bool wy = iy == sy.count;
        iy = wy ? 0 : iy;
        ix += wy;
    }
    __device__ __forceinline__ int index(const SearchGrid &g, const AxisSpan &sx, const AxisSpan &sy, const AxisSpan &sz,
                                         int &flags) const
    {
        int cx = sx.cell + ix, cy = sy.cell + iy, cz = sz.cell + iz;
        const int fx = cx >= g.len_xs - 1, fy = cy >= g.len_ys - 1, fz = cz >= g.len_zs - 1;
        cx -= fx ? g.len_xs - 1 : 0;
        cy -= fy ? g.len_ys - 1 : 0;
        cz -= fz ? g.len_zs - 1 : 0;
        flags = fx | (fy << 1) | (fz << 2);
        return (cx * g.nsv1 + cy) * g.nsv2 + cz;
    }
};

// The reference's loops as they are, for one walker: any number of cells, any list length.
// Taken by walkers on (or within rounding of) a cell or period boundary, with steps longer than
// the cooperative search handles, or when its tables are full.  Kept out of line: it is rare,
// and the hot loop stays small.
__device__ __noinline__ void mesh_closest_hit_alone(const MeshDev &g, const Vec3 &pos, const Vec3 &s, double ex,
                                                    double ey, double ez, double &min_d, int &closest)
{
    const AxisCells ax = axis_cells(g.xs, g.len_xs, g.vox[0], g.inv_vox[0], g.inv_hx, pos.x, ex);
    const AxisCells ay = axis_cells(g.ys, g.len_ys, g.vox[1], g.inv_vox[1], g.inv_hy, pos.y, ey);
    const AxisCells az = axis_cells(g.zs, g.len_zs, g.vox[2], g.inv_vox[2], g.inv_hz, pos.z, ez);
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
                    const double d = ray_triangle(load_tri(g.tri, id), tr0, s);
                    if (d > 0 && d < min_d) {
                        closest = id;
                        min_d = d;
                    }
                }
            }
        }
    }
}

// Closest triangle hit (d > 0) over every triangle listed in the cells the segment
// [pos, pos + step_l * s] overlaps, visited in the reference's order (cells x -> y -> z, entries
// ascending, strict "<" so the first minimum wins): simulations.py:936-983.
//
// Only hits within the remaining length can change the walk (a larger minimum just ends the
// collision loop, simulations.py:986), so a triangle whose bounding box does not meet the box of
// the remaining segment is skipped.  That filter reads a 16-byte record per list entry (triangle
// id + box on a 15-bit grid, rounded outwards; the segment's box is rounded outwards too, so the
// filter only ever errs on the side of keeping a triangle).  The lists of all 32 walkers are
// numbered through and filtered 32 entries at a time (coalesced), the survivors are compacted
// (ballot) and tested exactly, again 32 at a time.
//
// The lists that are filtered are those of g.fine: the reference's grid, or a refinement of it whose
// sub-cell lists hold the parent cell's triangles that reach into the sub-cell (padded boxes, built
// at upload).  The sub-cells the segment overlaps lie inside the reference cells it overlaps, and
// their lists are subsets of the parents', so only triangles the reference tests are tested; and a
// triangle the reference hits within the remaining length is hit at a point of the segment, which
// lies in an overlapped sub-cell the triangle reaches into (the segment's ends keep a margin from
// the cell boundaries, else the walker searches alone), so it is tested.  Shorter lists, same hit.
// What the sub-cells do not preserve is the reference's visiting order, which decides between
// DIFFERENT triangles hit at bitwise the same distance (a ray through a shared edge): such a walker
// repeats its search alone with the reference's own loops.
template <int MAXC>
__device__ __forceinline__ void mesh_closest_hit(const MeshDev &g, MeshScratch &sc, const int lane, const bool need,
                                                 const Vec3 &pos, const Vec3 &s, const double step_l,
                                                 double &min_d, int &closest)
{
    const unsigned full = 0xffffffffu;
    const double inf = __longlong_as_double(0x7FF0000000000000LL);
    const SearchGrid &f = g.fine;
    min_d = inf;
    bool fast = false;
    AxisSpan sx = {0, 0, 0}, sy = {0, 0, 0}, sz = {0, 0, 0};
    double ex = 0.0, ey = 0.0, ez = 0.0;
    int n_ranges = 0, n_entries = 0, n_cells = 0;
    sc.best_d[lane] = 0x7FF0000000000000ULL;
    sc.best_tri[lane] = 0x7fffffff;
    sc.best_tri_hi[lane] = -1;
#pragma unroll
    for (int k = 0; k < kEntryCap / 1024; ++k) sc.start_bits[lane + 32 * k] = 0u;
    if (need) {
        // end point of the remaining segment: x uses a separately rounded product, y and z are
        // fused (that is how the reference's kernel was compiled)
        ex = add_(pos.x, mul_(step_l, s.x));
        ey = fma_(step_l, s.y, pos.y);
        ez = fma_(step_l, s.z, pos.z);
        fast = axis_span(f.xs, f.len_xs - 1, g.vox[0], g.inv_vox[0], f.inv_hx, f.margin[0], pos.x, ex, sx);
        fast &= axis_span(f.ys, f.len_ys - 1, g.vox[1], g.inv_vox[1], f.inv_hy, f.margin[1], pos.y, ey, sy);
        fast &= axis_span(f.zs, f.len_zs - 1, g.vox[2], g.inv_vox[2], f.inv_hz, f.margin[2], pos.z, ez, sz);
        n_cells = sx.count * sy.count * sz.count;
        fast = fast && n_cells <= MAXC;
        if constexpr (MAXC == kMaxCellsShortStep) fast = fast && sx.count <= 2 && sy.count <= 2 && sz.count <= 2;
    }
    // list ranges of this lane's cells: independent loads, all in flight together
    int2 rng[MAXC];
#pragma unroll
    for (int c = 0; c < MAXC; ++c) rng[c] = make_int2(0, 0);
    // Short steps (at most 2 cells per axis): the cells are the slots of a static 2 x 2 x 2 nest,
    // slot 4 ix + 2 iy + iz, visited in that order like the reference's x -> y -> z loops; a slot
    // beyond an axis' count is empty.  Everything about a slot but the cell coordinates is static.
    int off_x[2] = {0, 0}, off_y[2] = {0, 0}, off_z[2] = {0, 0};  // row offsets of the (wrapped) cells per axis
    int wrap_x = 0, wrap_y = 0, wrap_z = 0;                        // the second cell lies in the next image
    int n_cells_warp = 0;
    if constexpr (MAXC == kMaxCellsShortStep) {
        if (fast) {
            int c1 = sx.cell + 1;
            wrap_x = c1 >= f.len_xs - 1;
            off_x[0] = sx.cell * f.nsv1 * f.nsv2;
            off_x[1] = (wrap_x ? 0 : c1) * f.nsv1 * f.nsv2;
            c1 = sy.cell + 1;
            wrap_y = c1 >= f.len_ys - 1;
            off_y[0] = sy.cell * f.nsv2;
            off_y[1] = (wrap_y ? 0 : c1) * f.nsv2;
            c1 = sz.cell + 1;
            wrap_z = c1 >= f.len_zs - 1;
            off_z[0] = sz.cell;
            off_z[1] = wrap_z ? 0 : c1;
        }
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const int ix = c >> 2, iy = (c >> 1) & 1, iz = c & 1;
            if (fast && ix < sx.count && iy < sy.count && iz < sz.count)
                rng[c] = __ldg(f.cell_rng + off_x[ix] + off_y[iy] + off_z[iz]);
        }
    } else {
        // general spans: the loops stop at the largest cell count of the warp
        n_cells_warp = __reduce_max_sync(full, fast ? n_cells : 0);
        CellWalk cw;
#pragma unroll
        for (int c = 0; c < MAXC; ++c) {
            if (c >= n_cells_warp) break;
            if (fast && c < n_cells) {
                int flags;
                rng[c] = __ldg(f.cell_rng + cw.index(f, sx, sy, sz, flags));
            }
            cw.next(sy, sz);
        }
    }
#pragma unroll
    for (int c = 0; c < MAXC; ++c) {
        n_ranges += rng[c].y > rng[c].x;
        n_entries += rng[c].y - rng[c].x;
    }
    // this lane's place in the warp's numbering of ranges and entries
    int incl_r = n_ranges, incl_e = n_entries;
#pragma unroll
    for (int st = 1; st < 32; st <<= 1) {
        const int vr = __shfl_up_sync(full, incl_r, st), ve = __shfl_up_sync(full, incl_e, st);
        if (lane >= st) {
            incl_r += vr;
            incl_e += ve;
        }
    }
    // walkers past the tables' capacity search alone (they form a suffix of the lanes)
    const bool coop = fast && incl_r <= kRangeCap && incl_e <= kEntryCap;
    const int total = __reduce_max_sync(full, coop ? incl_e : 0);
    __syncwarp();
    if (coop && n_entries > 0) {
        sc.dir[0][lane] = s.x; sc.dir[1][lane] = s.y; sc.dir[2][lane] = s.z;
        // walker moved into the base voxel: r0 - shift_n * xs[-1] (simulations.py:943, 970-971)
        const double ox = sub_(pos.x, mul_((double)sx.image, g.top[0]));
        const double oy = sub_(pos.y, mul_((double)sy.image, g.top[1]));
        const double oz = sub_(pos.z, mul_((double)sz.image, g.top[2]));
        sc.org[0][0][lane] = ox;
        sc.org[1][0][lane] = oy;
        sc.org[2][0][lane] = oz;
        sc.org[0][1][lane] = sub_(pos.x, mul_((double)(sx.image + 1), g.top[0]));
        sc.org[1][1][lane] = sub_(pos.y, mul_((double)(sy.image + 1), g.top[1]));
        sc.org[2][1][lane] = sub_(pos.z, mul_((double)(sz.image + 1), g.top[2]));
        // box of the remaining segment on the 15-bit grid, padded by a thousandth of its length
        // and two grid units; in the next image every coordinate is one period (32767) lower
        const double pad = 1e-3 * step_l;
        const double dx = step_l * s.x, dy = step_l * s.y, dz = step_l * s.z;
        const int lox = quantize(ox + fmin(dx, 0.0) - pad, g.qscale[0]) - 2, hix = quantize(ox + fmax(dx, 0.0) + pad, g.qscale[0]) + 3;
        const int loy = quantize(oy + fmin(dy, 0.0) - pad, g.qscale[1]) - 2, hiy = quantize(oy + fmax(dy, 0.0) + pad, g.qscale[1]) + 3;
        const int loz = quantize(oz + fmin(dz, 0.0) - pad, g.qscale[2]) - 2, hiz = quantize(oz + fmax(dz, 0.0) + pad, g.qscale[2]) + 3;
        const unsigned hx[2] = {clamp15(hix), clamp15(hix - 32767)}, lx[2] = {32767u - clamp15(lox), 32767u - clamp15(lox - 32767)};
        const unsigned hy[2] = {clamp15(hiy), clamp15(hiy - 32767)}, ly[2] = {32767u - clamp15(loy), 32767u - clamp15(loy - 32767)};
        const unsigned hz[2] = {clamp15(hiz), clamp15(hiz - 32767)}, lz[2] = {32767u - clamp15(loz), 32767u - clamp15(loz - 32767)};
        int k = incl_r - n_ranges, first = incl_e - n_entries;
        auto add_range = [&](int c, int flags) {
            const int n = rng[c].y - rng[c].x;
            if (n > 0) {
                const int fx = flags & 1, fy = (flags >> 1) & 1, fz = flags >> 2;
                // triangle box (lo, 32767 - hi) <= these, halfword by halfword  <=>  the boxes meet
                sc.range_box[k] = make_uint4(((fx ? hx[1] : hx[0]) | ((fy ? hy[1] : hy[0]) << 16)) + kSwarH,
                                             ((fz ? hz[1] : hz[0]) | ((fx ? lx[1] : lx[0]) << 16)) + kSwarH,
                                             ((fy ? ly[1] : ly[0]) | ((fz ? lz[1] : lz[0]) << 16)) + kSwarH,
                                             (unsigned)lane | ((unsigned)flags << 8));
                sc.range_pos[k] = make_int2(first, rng[c].x);
                atomicOr(&sc.start_bits[first >> 5], 1u << (first & 31));
                ++k;
                first += n;
            }
        };
        if constexpr (MAXC == kMaxCellsShortStep) {
            // (All eight slots, unrolled and predicated, although a lane has entries in one to three of them --
            // 257 instructions per warp-step at 5 active lanes.  A loop over the lane's non-empty slots only, their
            // ranges parked in shared memory, was measured: -21 %, profiles/r02_n_kbench_range_loop.txt; eight votes and a
            // warp-uniform skip of the slots in which no lane has a list: -1.7 %, profiles/r02_al_kbench_slot_vote.txt.)
#pragma unroll
            for (int c = 0; c < 8; ++c)
                add_range(c, ((c >> 2) & wrap_x) | ((((c >> 1) & 1) & wrap_y) << 1) | (((c & 1) & wrap_z) << 2));
        } else {
            CellWalk cw;
#pragma unroll
            for (int c = 0; c < MAXC; ++c) {
                if (c >= n_cells_warp) break;
                int flags;
                cw.index(f, sx, sy, sz, flags);
                add_range(c, flags);
                cw.next(sy, sz);
            }
        }
    }
    __syncwarp();

    // box filter over the warp's entries, 2 x 32 per iteration (two independent loads in flight)
    int n_surv = 0, started = 0;  // warp-uniform: survivors so far, ranges that start before this round
    for (int j0 = 0; j0 < total; j0 += 64) {
        const unsigned starts0 = sc.start_bits[j0 >> 5], starts1 = sc.start_bits[(j0 >> 5) + 1];
        const int started1 = started + __popc(starts0);
        bool pass[2] = {false, false};
        unsigned long long rec[2] = {0ull, 0ull};
        uint4 a[2], b[2];
        int off[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int j = j0 + 32 * h + lane;
            a[h] = make_uint4(0u, 0u, 0u, 0u);
            b[h] = make_uint4(0u, 0xffffffffu, 0xffffffffu, 0xffffffffu);
            off[h] = 0;
            if (j < total) {
                const int r = (h ? started1 : started) + __popc((h ? starts1 : starts0) & (0xffffffffu >> (31 - lane))) - 1;
                a[h] = sc.range_box[r];
                const int2 rp = sc.range_pos[r];
                off[h] = j - rp.x;
                b[h] = __ldg(f.entry + rp.y + off[h]);
            }
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            pass[h] = ((a[h].x - b[h].y) & (a[h].y - b[h].z) & (a[h].z - b[h].w) & kSwarH) == kSwarH;
            rec[h] = (unsigned long long)b[h].x | ((unsigned long long)a[h].w << 32);
            const unsigned m = __ballot_sync(full, pass[h]);
            if (pass[h]) {
                const int k = n_surv + __popc(m & ((1u << lane) - 1u));
                if (k < kSurvivorCap) sc.survivor[k] = rec[h];
            }
            n_surv += __popc(m);
        }
        started = started1 + __popc(starts1);
    }
    const bool overflow = n_surv > kSurvivorCap;  // every walker of the warp then searches alone
    n_surv = overflow ? 0 : n_surv;
    __syncwarp();

    // exact tests of the survivors, 32 at a time
    for (int j = lane; j < n_surv; j += 32) {
        const unsigned long long w = sc.survivor[j];
        const unsigned hi = (unsigned)(w >> 32);
        const int owner = hi & 0xff, flags = (hi >> 8) & 7;
        Vec3 o, dir;
        o.x = sc.org[0][flags & 1][owner];
        o.y = sc.org[1][(flags >> 1) & 1][owner];
        o.z = sc.org[2][(flags >> 2) & 1][owner];
        dir.x = sc.dir[0][owner]; dir.y = sc.dir[1][owner]; dir.z = sc.dir[2][owner];
        const double t = ray_triangle(load_tri(g.tri, (int)((unsigned)w & kEntryTriMask)), o, dir);
        const bool is_hit = t > 0;
        sc.hit[j] = is_hit ? t : inf;
        if (is_hit) atomicMin(&sc.best_d[owner], (unsigned long long)__double_as_longlong(t));
    }
    __syncwarp();
    // which triangle: the one at the closest distance; two different triangles at bitwise the same
    // distance (a ray through a shared edge) are a tie the reference breaks by its visiting order
    for (int j = lane; j < n_surv; j += 32) {
        const unsigned long long w = sc.survivor[j];
        const int owner = (unsigned)(w >> 32) & 0xff;
        const double t = sc.hit[j];
        if (t < inf && (unsigned long long)__double_as_longlong(t) == sc.best_d[owner]) {
            const int id = (int)((unsigned)w & kEntryTriMask);
            atomicMin(&sc.best_tri[owner], id);
            atomicMax(&sc.best_tri_hi[owner], id);
        }
    }
    __syncwarp();

    const bool tie = coop && !overflow && sc.best_tri_hi[lane] > sc.best_tri[lane];
    if (coop && !overflow && !tie) {
        const double d = __longlong_as_double((long long)sc.best_d[lane]);
        if (d < inf) {
            min_d = d;
            closest = sc.best_tri[lane];
        }
    } else if (need) {  // boundary cases, long segments, table overflow, ties: this lane walks its own cells
        mesh_closest_hit_alone(g, pos, s, ex, ey, ez, min_d, closest);
    }
}

// unit normal of every triangle, once per mesh upload (same device functions the reference's
// per-collision computation is restated with: _cuda_triangle_normal, simulations.py:77-97)
__global__ void __launch_bounds__(256) tri_normal_kernel(const double *tri, long long n_faces, double *normal)
{
    const long long f = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= n_faces) return;
    const Vec3 n = triangle_normal(load_tri(tri, (int)f));
    normal[3 * f] = n.x;
    normal[3 * f + 1] = n.y;
    normal[3 * f + 2] = n.z;
}

// What happens at the closest triangle (simulations.py:986-997): one uniform draw (also when
// perm_prob == 0), then reflection or passage through the membrane.
__device__ __forceinline__ void mesh_collision(const MeshDev &g, Vec3 &pos, Vec3 &s, Rng &rng, double min_d,
                                               int closest, double eps)
{
    const double u = u01_f64(rng_next(rng));
    // the reference recomputes the normal at every collision (simulations.py:77-97); it only depends
    // on the triangle, so it is read from the table tri_normal_kernel filled with the same arithmetic
    const double *nq = g.normal + 3ll * closest;
    const Vec3 n = {__ldg(nq), __ldg(nq + 1), __ldg(nq + 2)};
    if (g.perm_prob < u)
        reflect(pos, s, min_d, n, eps);
    else
        cross_membrane(pos, s, min_d, n, eps);
}

// Time steps [t_begin, t_end) of a mesh walk for the 32 walkers of a warp; done(t) is called by
// a lane when its walker has completed step t (pos = the new position).
//
// The collision search is shared by the warp (mesh_closest_hit), so a second search for the one
// or two walkers that bounced costs about as much as the first one for all 32.  The lanes are
// therefore not kept in lock step: a walker that bounced stays in flight and takes part in the
// warp's next search together with the other lanes' next time steps.  Every walker still
// executes exactly its own sequence of operations (simulations.py:878-1013).
template <int MAXC, typename Done>
__device__ __forceinline__ void mesh_walk(const KParams &p, MeshScratch &sc, const double *tab, const bool active,
                                          const int t_begin, const int t_end, Vec3 &pos, Rng &rng, bool &exc, Done done)
{
    const int lane = threadIdx.x & 31;
    const MeshDev &g = p.mesh;
    int t = t_begin, iter = 0, closest = 0;
    bool in_flight = false;
    Vec3 s = {0.0, 0.0, 0.0};
    double step_l = 0.0;
    for (;;) {
        const bool fresh = active && !in_flight && t < t_end;
        if (!__any_sync(0xffffffffu, fresh || in_flight)) break;
        if (fresh) {
            s = random_step(rng, tab);
            step_l = p.step_l;
            iter = 0;
            in_flight = true;
        }
        const bool need = in_flight && step_l > 0 && iter < p.max_iter;
        if (need) ++iter;
        double min_d;
        mesh_closest_hit<MAXC>(g, sc, lane, need, pos, s, step_l, min_d, closest);
        if (in_flight) {
            if (need && !(min_d > step_l)) {
                mesh_collision(g, pos, s, rng, min_d, closest, p.eps);
                step_l = sub_(step_l, min_d);
            } else {
                pos.x = fma_(step_l, s.x, pos.x);
                pos.y = fma_(step_l, s.y, pos.y);
                pos.z = fma_(step_l, s.z, pos.z);
                exc |= iter >= p.max_iter;
                done(t);
                ++t;
                in_flight = false;
            }
        }
    }
}

// ---------------------------------------------------------------- TMA bulk copy + mbarrier

__device__ __forceinline__ unsigned smem_addr(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}

// one thread: announce `bytes` and start the bulk copy global -> shared that will deliver them
__device__ __forceinline__ void tma_load_1d(void *dst, const void *src, unsigned bytes, unsigned long long *bar)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_addr(dst)),
                 "l"(src), "r"(bytes), "r"(smem_addr(bar))
                 : "memory");
}

__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "DSB_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DSB_DONE_%=;\n"
        "bra DSB_WAIT_%=;\n"
        "DSB_DONE_%=:\n"
        "}\n" ::"r"(smem_addr(bar)),
        "r"(parity)
        : "memory");
}

// ---------------------------------------------------------------- block reduction of the signal

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}

// phase(m) -> per-block sum of cos(phase) over walkers with a clear iter_exc flag, written to
// partials[m * n_blocks_total + block]; row n_meas receives the number of such walkers.
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
            p.partials[(long long)(m0 + lane) * p.n_blocks_total + (int)(p.w_begin / kBlock) + blockIdx.x] = acc;
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------- the walk kernel

template <int SUB>
__device__ __forceinline__ bool time_step(Vec3 &pos, const Vec3 &unit, const KParams &p, const bool live)
{
    static_assert(SUB != 4, "the mesh walk has its own time loop (mesh_walk)");
    if constexpr (SUB == 0) return free_step(pos, unit, p, live);
    else return walker_step<SUB>(pos, unit, p, live);
}

// Round 2 measured three designs that give a warp 64 walkers for its 32 lanes so that colliding
// walkers can be bounced in full batches while the lanes go on with other walkers (walker state in
// shared memory every step; state in registers and swapped on collision; two walkers per lane in
// registers): all bit-exact, all slower than this kernel for the sphere and the cylinder (4.4e10-4.9e10
// against 5.2e10 walker-steps/s; the ellipsoid gained 7 % with one of them, with one measurement only):
// once the lanes of a warp are at different time steps, the gradient sample, the loop counter and the
// branches stop being warp-uniform, and that costs more than the ~140 instructions per warp-step the
// batched reflection saves (profiles/r02_a_kbench.txt, r02_g/h/i_kbench_pool_*.txt).
//
// Parked walkers per warp that trigger a bounce pass (1: collisions are handled inline, step by
// step).  Measured on a B200 (tools/kbench.py, 1e6 walkers): batching pays when the bounce is
// expensive (ellipsoid: +9 % at 6), not for the sphere and the cylinder (-3 % at 4), where the
// lanes idling next to parked walkers cost more than the bounce code saves.
#ifndef DSB_PARK
#define DSB_PARK 1
#endif
#ifndef DSB_MESH_SMEM_PHASES_FROM
#define DSB_MESH_SMEM_PHASES_FROM 2   // measurements from which the mesh walk keeps its phases in shared memory
#endif
#ifndef DSB_PARK_ELLIPSOID
#define DSB_PARK_ELLIPSOID 6
#endif
#ifndef DSB_PARK_CYLINDER
#define DSB_PARK_CYLINDER 3   // round 2, with the branch-free step generator: +1.4 % (2: +1.1 %, 4: -1.4 %, 5: -6 %)
#endif
template <int SUB>
struct ParkFlush {
    static constexpr int value = SUB == 3 ? DSB_PARK_ELLIPSOID : (SUB == 2 ? DSB_PARK_CYLINDER : DSB_PARK);
};

// D = A * B + C on the FP64 tensor cores: A 8x4 (row major), B 4x8 (column major), C/D 8x8.
// Lane l holds A[l / 4][l % 4], B[l % 4][l / 4] and C[l / 4][2 * (l % 4) + {0, 1}].
__device__ __forceinline__ void dmma_m8n8k4(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// MR > 0: n_meas == MR phases in registers.  MR == 0: any n_meas; positions of a chunk of steps
// are buffered in registers, then each measurement's phase makes one round trip through its
// (coalesced, L2-resident) row of `phases` per chunk instead of one per step.
template <int SUB, int MR, int MAXC = kMaxCells>
__global__ void __launch_bounds__(kBlock, SUB == 4 ? DSB_MESH_MIN_BLOCKS : (MR == 0 ? DSB_MR0_MIN_BLOCKS : (SUB == 3 ? DSB_MIN_BLOCKS_ELLIPSOID : DSB_MIN_BLOCKS))) walk_kernel(const __grid_constant__ KParams p)
{
    __shared__ __align__(16) double s_tab[16];
    if (threadIdx.x < 16) s_tab[threadIdx.x] = __longlong_as_double((long long)c_sincos_tab[threadIdx.x]);
    __syncthreads();

    long long w = p.w_begin + (long long)blockIdx.x * kBlock + threadIdx.x;
    const bool active = w < p.w_end;
    if constexpr (SUB == 4 && MR > 0) {
        // walkers in cell order (cell_order_* kernels below): the lanes of a warp search the same few lists
        if (p.order != nullptr && active) w = __ldg(p.order + w);
    }
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
        // The mesh search runs at the register cap: with more than one measurement the phases
        // live in shared memory during the walk (the registers they would hold cost the search
        // loop its load pipelining: -15 % at 3 measurements).
        constexpr bool kSmemPhases = SUB == 4 && MR >= DSB_MESH_SMEM_PHASES_FROM;
        __shared__ double s_ph[kSmemPhases ? MR : 1][kSmemPhases ? kBlock : 1];
        if constexpr (kSmemPhases) {
#pragma unroll
            for (int m = 0; m < MR; ++m) s_ph[m][threadIdx.x] = ph[m];
        }
        // phase accumulation with the post-step position (simulations.py:746-755 and alike)
        auto accumulate = [&](int t) {
#pragma unroll
            for (int m = 0; m < MR; ++m) {
                const double *g = p.grad + ((long long)m * p.n_t + t) * 3;
                double gx = __ldg(g), gy = __ldg(g + 1), gz = __ldg(g + 2);
                if constexpr (kSmemPhases)
                    s_ph[m][threadIdx.x] = fma_(p.gamma_dt, fma_(gz, pos.z, fma_(gx, pos.x, mul_(gy, pos.y))), s_ph[m][threadIdx.x]);
                else
                    ph[m] = fma_(p.gamma_dt, fma_(gz, pos.z, fma_(gx, pos.x, mul_(gy, pos.y))), ph[m]);
            }
        };
        if constexpr (SUB >= 1 && SUB <= 3 && ParkFlush<SUB>::value > 1) {
            // Collisions are rare per walker and step but not per warp: taken inline, the
            // reflection code would run for one or two lanes in most steps of every warp.  So
            // the lanes of a warp are not kept in lock step: a walker that hits the wall is
            // parked with its step in flight while the other lanes go on with their next steps,
            // and once ParkFlush<SUB>::value walkers of the warp are parked (ballot) they are bounced
            // together.  Every walker still executes exactly its own sequence of operations,
            // so results do not depend on the grouping.
            const unsigned full = 0xffffffffu;
            int t = p.t0;
            bool parked = false;
            Flight f;
            f.d = 0.0;
            for (;;) {
                const bool fresh = active && !parked && t < p.t1;
                const unsigned m_parked = __ballot_sync(full, parked);
                const unsigned m_fresh = __ballot_sync(full, fresh);
                if ((m_parked | m_fresh) == 0) break;
                bool moved;
                if (__popc(m_parked) >= ParkFlush<SUB>::value || m_fresh == 0) {
                    moved = parked;
                    if (parked) bounce<SUB>(f, p);
                } else {
                    moved = fresh;
                    if (fresh) begin_step<SUB>(pos, random_step(rng, s_tab), p, f);
                }
                if (moved) {
                    parked = probe<SUB>(f, p);
                    if (!parked) {
                        exc |= end_step<SUB>(pos, f, p);
                        accumulate(t);
                        ++t;
                    }
                }
            }
        } else if constexpr (SUB == 4) {
            __shared__ MeshScratch s_scratch[kBlock / 32];
            mesh_walk<MAXC>(p, s_scratch[threadIdx.x >> 5], s_tab, active, p.t0, p.t1, pos, rng, exc, accumulate);
            if constexpr (kSmemPhases) {
#pragma unroll
                for (int m = 0; m < MR; ++m) ph[m] = s_ph[m][threadIdx.x];
            }
        } else {
            // the time loop is uniform over the block
            for (int t = p.t0; t < p.t1; ++t) {
                exc |= time_step<SUB>(pos, random_step(rng, s_tab), p, active);
                accumulate(t);
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
        // Any number of measurements.  The phase update of a chunk of C steps is a matrix product:
        // Phi[m, i] += sum_k G'[m, k] X[k, i] with G' = gamma dt g (n_meas x 3C, the chunk-major
        // gradient copy) and X the positions of the chunk (3C x walkers).  It runs on the FP64
        // tensor cores (mma.sync m8n8k4: the same FLOP/s as DFMA on B200 from an eighth of the
        // instructions, which is what limited the DFMA version), one warp for its own 32 walkers:
        //  * X: every lane writes its walker's positions into the warp's tile in shared memory
        //    (swizzled columns: the B fragments are read without bank conflicts);
        //  * G': tiles of kGradRows rows are streamed through shared memory by TMA bulk copies
        //    (double buffered: the next tile, or the next chunk's first tile, travels while this one
        //    is used);
        //  * Phi: 8 x 8 accumulator tiles straight from / to the (n_meas, n_walkers) array with
        //    evict-first accesses, requested one group of 8 measurements ahead; one round trip
        //    through HBM per chunk.  (Measured in round 2, profiles/r02_c_kbench_variants.txt: with
        //    L2-resident hints -- ld/st.cg, 3 or 4 blocks per SM so that the resident tiles, 82-109 MB
        //    at 180 measurements, fit the 126 MB L2 -- the sphere runs at 1.02e10 walker-steps/s
        //    against 1.34e10 with evict-first: the tiles of retired blocks crowd out the live ones.
        //    Also measured without gain (profiles/r02_f_kbench_many_meas_variants.txt): the chunk's B
        //    fragments held in registers for 16 walkers at a time (0.54 instead of 1.25 shared-memory
        //    loads per mma, two passes: 1.29e10), A fragments by __ldg instead of TMA tiles (1.28e10).)
        // Summation order and roundings differ from the reference's fma chain at the 1e-16 level
        // (phases for n_meas > 4 agree to ~1e-13, not bit for bit; positions are not affected).
        // The ragged end of a run (fewer than C steps) uses the reference's formula.
        constexpr int C = ChunkSteps<SUB>::value, kRows = 3 * C, kRowLen = grad_row_len(C);
        // a run that starts with a whole chunk starts its accumulators at zero instead of reading
        // zeros it would have had to write first
        const bool first_chunk_whole = p.t0 == 0 && p.t1 >= C;
        if (active && p.t0 == 0 && !first_chunk_whole)
            for (int m = 0; m < p.n_meas; ++m) p.phases[(long long)m * N + w] = 0.0;
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        // the warp's position tile: rows k = 3 * step + axis, 32 columns = lanes, column c of row k
        // kept at c ^ ((k & 3) << 3).  The mesh walk keeps its positions in local memory while it
        // runs (its shared scratch is busy) and the tile reuses that scratch afterwards; the
        // analytic kernels get tile and gradient buffers from dynamic shared memory.
        double *xs, *s_grad;
        Vec3 buf[SUB == 4 ? C : 1];
        MeshScratch *scratch = nullptr;
        __shared__ unsigned long long s_bar[2];
        if constexpr (SUB == 4) {
            __shared__ MeshScratch s_scratch[kBlock / 32];
            static_assert(sizeof(MeshScratch) >= sizeof(double) * kRows * 32, "position tile must fit the scratch");
            scratch = &s_scratch[warp];
            xs = reinterpret_cast<double *>(scratch);
            s_grad = nullptr;  // the mesh kernel reads the gradient from global memory (see below)
        } else {
            extern __shared__ __align__(128) double s_dyn[];
            s_grad = s_dyn;
            xs = s_dyn + 2 * kGradRows * kRowLen + warp * kRows * 32;
        }
        auto x_at = [&](int row, int col) -> double & { return xs[row * 32 + (col ^ ((row & 3) << 3))]; };
        const double *tile = s_grad;
        int n_tiles = 0;  // tiles consumed so far by this block (buffer = n_tiles & 1, parity = bit 1)
        if (threadIdx.x == 0) {
            mbar_init(&s_bar[0], 1);
            mbar_init(&s_bar[1], 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        if (SUB != 4 && threadIdx.x == 0 && p.t0 % C == 0 && p.t1 - p.t0 >= C)  // first tile of the first chunk
            tma_load_1d(s_grad, p.grad_chunked + (long long)(p.t0 / C) * p.n_meas * kRowLen,
                        min(kGradRows, p.n_meas) * kRowLen * 8, &s_bar[0]);
        for (int t = p.t0; t < p.t1; t += C) {
            const int cnt = min(C, p.t1 - t);
            if constexpr (SUB == 4) {
                mesh_walk<MAXC>(p, *scratch, s_tab, active, t, t + cnt, pos, rng, exc, [&](int tt) { buf[tt - t] = pos; });
                __syncwarp();
                for (int k = 0; k < cnt; ++k) {
                    x_at(3 * k, lane) = buf[k].x;
                    x_at(3 * k + 1, lane) = buf[k].y;
                    x_at(3 * k + 2, lane) = buf[k].z;
                }
            } else {
#pragma unroll 1
                for (int k = 0; k < cnt; ++k) {
                    exc |= time_step<SUB>(pos, random_step(rng, s_tab), p, active);
                    x_at(3 * k, lane) = pos.x;
                    x_at(3 * k + 1, lane) = pos.y;
                    x_at(3 * k + 2, lane) = pos.z;
                }
            }
            __syncwarp();
            if (cnt == C && t % C == 0) {
                const int g8 = lane >> 2, t4 = lane & 3;  // row group and thread-in-group of the mma fragments
                const double *gc = p.grad_chunked + (long long)(t / C) * p.n_meas * kRowLen;
                const bool next_chunk = t + 2 * C <= p.t1;
                const long long w_warp = w - lane;  // first walker of the warp
                auto load_c = [&](int m0, double (&c)[4][2]) {
                    const int m = m0 + g8;
                    const bool row_ok = m < p.n_meas && !(first_chunk_whole && t == 0);
                    const double *row = p.phases + (long long)(row_ok ? m : 0) * N + w_warp + 2 * t4;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const long long wj = w_warp + 8 * j + 2 * t4;
                        c[j][0] = (row_ok && wj < p.w_end) ? __ldcs(row + 8 * j) : 0.0;
                        c[j][1] = (row_ok && wj + 1 < p.w_end) ? __ldcs(row + 8 * j + 1) : 0.0;
                    }
                };
                // B fragment of k-step q and walker tile j: row 4q + t4 (so row & 3 == t4 for every q),
                // column 8j + g8 -> one pointer per j, the k-step is an immediate offset
                const double *bcol[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) bcol[j] = xs + t4 * 32 + ((8 * j + g8) ^ (t4 << 3));
                double c[4][2], c_next[4][2];
                load_c(0, c_next);
                for (int m0 = 0; m0 < p.n_meas; m0 += 8) {
                    if (SUB != 4 && m0 % kGradRows == 0) {  // next tile
                        const int bufi = n_tiles & 1;
                        __syncthreads();  // everybody is done with the other buffer
                        if (threadIdx.x == 0) {
                            const int m_next = m0 + kGradRows;
                            if (m_next < p.n_meas)
                                tma_load_1d(s_grad + (bufi ^ 1) * kGradRows * kRowLen, gc + (long long)m_next * kRowLen,
                                            min(kGradRows, p.n_meas - m_next) * kRowLen * 8, &s_bar[bufi ^ 1]);
                            else if (next_chunk)
                                tma_load_1d(s_grad + (bufi ^ 1) * kGradRows * kRowLen, gc + (long long)p.n_meas * kRowLen,
                                            min(kGradRows, p.n_meas) * kRowLen * 8, &s_bar[bufi ^ 1]);
                        }
                        mbar_wait(&s_bar[bufi], (n_tiles >> 1) & 1);
                        tile = s_grad + bufi * kGradRows * kRowLen;
                        ++n_tiles;
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        c[j][0] = c_next[j][0];
                        c[j][1] = c_next[j][1];
                    }
                    if (m0 + 8 < p.n_meas) load_c(m0 + 8, c_next);
                    const int m = m0 + g8;  // the measurement of this lane's A and C fragments
                    const bool row_ok = m < p.n_meas;
                    double *row = p.phases + (long long)(row_ok ? m : 0) * N + w_warp + 2 * t4;
                    // The warps of a mesh block reach this pass at different times (their walks differ),
                    // so they do not share gradient tiles (that needs block-wide barriers): each reads
                    // its A fragments from the L1/L2-resident chunk-major copy directly.
                    const double *arow = SUB == 4 ? gc + (long long)(row_ok ? m : 0) * kRowLen + t4
                                                  : tile + ((m0 % kGradRows) + g8) * kRowLen + t4;
                    // operands of k-step q + 1 are read from shared memory while the four products of
                    // k-step q run
                    double a = row_ok ? (SUB == 4 ? __ldg(arow) : arow[0]) : 0.0, b[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) b[j] = bcol[j][0];
#pragma unroll
                    for (int q = 0; q < kRows / 4; ++q) {
                        double a_next = 0.0, b_next[4] = {0.0, 0.0, 0.0, 0.0};
                        if (q + 1 < kRows / 4) {
                            a_next = row_ok ? (SUB == 4 ? __ldg(arow + 4 * (q + 1)) : arow[4 * (q + 1)]) : 0.0;
#pragma unroll
                            for (int j = 0; j < 4; ++j) b_next[j] = bcol[j][128 * (q + 1)];
                        }
#pragma unroll
                        for (int j = 0; j < 4; ++j) dmma_m8n8k4(c[j][0], c[j][1], a, b[j]);
                        a = a_next;
#pragma unroll
                        for (int j = 0; j < 4; ++j) b[j] = b_next[j];
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const long long wj = w_warp + 8 * j + 2 * t4;
                        if (row_ok && wj < p.w_end) __stcs(row + 8 * j, c[j][0]);
                        if (row_ok && wj + 1 < p.w_end) __stcs(row + 8 * j + 1, c[j][1]);
                    }
                }
            } else if (active) {  // ragged end of the run, or a launch that does not start on a chunk boundary
                for (int m = 0; m < p.n_meas; ++m) {
                    double *row = p.phases + (long long)m * N + w;
                    double a = *row;
                    const double *g = p.grad + ((long long)m * p.n_t + t) * 3;
                    for (int k = 0; k < cnt; ++k) {
                        double gx = __ldg(g + 3 * k), gy = __ldg(g + 3 * k + 1), gz = __ldg(g + 3 * k + 2);
                        a = fma_(p.gamma_dt, fma_(gz, x_at(3 * k + 2, lane), fma_(gx, x_at(3 * k, lane), mul_(gy, x_at(3 * k + 1, lane)))), a);
                    }
                    *row = a;
                }
            }
            __syncwarp();  // the tile is rewritten by the next chunk
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

// ---------------------------------------------------------------- low-rank protocols

// When the (n_meas x 3 n_t) gradient matrix factors as U V with r <= kMaxRank rows in V (every
// PGSE-type protocol: one time profile, scaled and rotated per measurement, has rank <= 3; k
// different timings give rank <= 3k), the walk carries the r phases psi of the rows of V (in
// registers for r <= kMaxRegMeas, through the many-measurement kernels above that) and the
// n_meas real phases are phi[m, i] = sum_k U[m, k] psi[k, i], formed once at the end.

// per-block partial sums of cos(phase) from the phase array, for runs whose last launch did not
// advance the walkers in index order (same partials, same summation tree as walk_kernel's own)
template <int MR>
__global__ void __launch_bounds__(kBlock) phases_signal_kernel(const KParams p)
{
    const long long w = (long long)blockIdx.x * kBlock + threadIdx.x;
    const bool active = w < p.n_walkers;
    double ph[MR];
#pragma unroll
    for (int m = 0; m < MR; ++m) ph[m] = active ? p.phases[(long long)m * p.n_walkers + w] : 0.0;
    block_signal(p, active && p.iter_exc[w] == 0, [&](int m) {
        double v = 0.0;
#pragma unroll
        for (int k = 0; k < MR; ++k)
            if (k == m) v = ph[k];
        return v;
    });
}

// ---------------------------------------------------------------- walkers in cell order
//
// A mesh walk with the walkers in index order has the ~600 resident blocks spread over the whole
// mesh, and a mesh larger than the L2 is then read from HBM over and over.  A counting sort of the
// walkers by the grid cell they start in (cell_order_count -> cell_order_scan -> cell_order_scatter)
// puts the walkers of consecutive blocks into one slab of the mesh (dsb_api.cu, resort_interval,
// has the measurements and says when the sort is repeated).  Only the assignment of walkers to
// threads changes: every walker runs its own stream from its own state and writes its own slots,
// so results do not depend on the order (the order inside a cell is whatever the atomics give).

struct CellBins {
    int n[3];          // bins per axis (the reference grid's cells, at most 128 per axis)
    double inv_vox[3];
    int axis[3];       // sort order: axis[0] (the longest edge of the voxel) varies slowest
};

__device__ __forceinline__ int cell_bin(const double *pos, long long i, const CellBins &b)
{
    int idx[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        double f = pos[3 * i + a] * b.inv_vox[a];
        f -= floor(f);                                   // periodic image; positions inside the voxel are unchanged
        idx[a] = min(max((int)(f * b.n[a]), 0), b.n[a] - 1);   // (a NaN position converts to 0)
    }
    return (idx[b.axis[0]] * b.n[b.axis[1]] + idx[b.axis[1]]) * b.n[b.axis[2]] + idx[b.axis[2]];
}

__global__ void __launch_bounds__(256) cell_order_count(const double *pos, long long n, CellBins b, int *key, int *count)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int k = cell_bin(pos, i, b);
    key[i] = k;
    atomicAdd(count + k, 1);
}

// exclusive prefix sums of the bin counts, in place; one block (the array is a few MB at most)
__global__ void __launch_bounds__(1024) cell_order_scan(int *count, int n_bins)
{
    __shared__ int s_sum[1024];
    const int per = (n_bins + 1023) / 1024;
    const int b = min(threadIdx.x * per, n_bins), e = min(b + per, n_bins);
    int acc = 0;
    for (int i = b; i < e; ++i) acc += count[i];
    s_sum[threadIdx.x] = acc;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
        const int v = (int)threadIdx.x >= o ? s_sum[threadIdx.x - o] : 0;
        __syncthreads();
        s_sum[threadIdx.x] += v;
        __syncthreads();
    }
    int run = s_sum[threadIdx.x] - acc;
    for (int i = b; i < e; ++i) {
        const int c = count[i];
        count[i] = run;
        run += c;
    }
}

__global__ void __launch_bounds__(256) cell_order_scatter(const int *key, long long n, int *cursor, int *order)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    order[atomicAdd(cursor + key[i], 1)] = (int)i;
}

// per-block partial sums of cos(phi) over unflagged walkers (same layout as block_signal writes)
__global__ void __launch_bounds__(kBlock) lowrank_signal_kernel(const KParams p, const double *u, int rank, int n_real)
{
    const long long w = (long long)blockIdx.x * kBlock + threadIdx.x;
    const bool active = w < p.n_walkers;
    double psi[kMaxRank];
#pragma unroll
    for (int k = 0; k < kMaxRank; ++k) psi[k] = (active && k < rank) ? p.phases[(long long)k * p.n_walkers + w] : 0.0;
    const bool valid = active && p.iter_exc[w] == 0;
    KParams q = p;
    q.n_meas = n_real;
    q.w_begin = 0;
    block_signal(q, valid, [&](int m) {
        double ph = 0.0;
#pragma unroll
        for (int k = 0; k < kMaxRank; ++k)
            if (k < rank) ph = __fma_rn(__ldg(u + (long long)m * rank + k), psi[k], ph);
        return ph;
    });
}

// the (n_real, n_walkers) phase matrix itself, for callers that ask for per-walker output
__global__ void __launch_bounds__(256) lowrank_expand_kernel(const double *psi, const double *u, int rank, int n_real,
                                                             long long n_walkers, double *phases)
{
    const long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n_walkers) return;
    double v[kMaxRank];
#pragma unroll
    for (int k = 0; k < kMaxRank; ++k) v[k] = k < rank ? psi[(long long)k * n_walkers + w] : 0.0;
    for (int m = 0; m < n_real; ++m) {
        double ph = 0.0;
#pragma unroll
        for (int k = 0; k < kMaxRank; ++k)
            if (k < rank) ph = __fma_rn(__ldg(u + (long long)m * rank + k), v[k], ph);
        phases[(long long)m * n_walkers + w] = ph;
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

// ---------------------------------------------------------------- L2 bandwidth probe

// Every thread streams 16-byte loads over a buffer that fits the L2 (the host sizes it), again and again:
// after the first pass the bytes come from L2.  Its rate is the L2 -> SM ceiling the mesh roofline of
// bench.py is quoted against (measured here instead of taken from a document).
__global__ void __launch_bounds__(256) l2_peak_kernel(const uint4 *buf, long long n_vec, int passes, unsigned *out)
{
    const long long stride = (long long)gridDim.x * blockDim.x;
    unsigned acc = 0u;
    for (int pass = 0; pass < passes; ++pass)
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec; i += stride) {
            const uint4 v = __ldcg(buf + i);
            acc ^= v.x ^ v.y ^ v.z ^ v.w;
        }
    if (acc == 0x12345678u) out[0] = acc;   // (keeps the loads alive)
}

// ---------------------------------------------------------------- ellipsoid constants

__global__ void ellipsoid_consts_kernel(double a0, double a1, double a2, EllipsoidConsts *out)
{
    const double ax[3] = {a0, a1, a2};
    for (int k = 0; k < 3; ++k) {
        out->ax[k] = ax[k];
        out->ax_rc[k] = rcp_refined(ax[k]);
        out->axsq[k] = mul_(ax[k], ax[k]);
        out->ax_isq[k] = rcp_(out->axsq[k]);
        out->axsq_rc[k] = rcp_refined(out->axsq[k]);
    }
}

// ---------------------------------------------------------------- device functions one at a time

// The walk's device functions on rows of arguments, the way the reference's unit tests call its
// `_cuda_*` functions through small test kernels (disimpy/tests/test_simulations.py:23-360);
// layouts in include/disimpy_b200.h (dsb_selftest_device_function).
constexpr int kUnitOps = 15;
__host__ __device__ constexpr int unit_n_in(int op)
{
    constexpr int n[kUnitOps] = {6, 6, 3, 9, 12, 5, 7, 9, 15, 11, 11, 19, 19, 19, 19};
    return n[op];
}
__host__ __device__ constexpr int unit_n_out(int op)
{
    constexpr int n[kUnitOps] = {1, 3, 3, 3, 3, 1, 1, 1, 1, 6, 3, 1, 1, 1, 1};
    return n[op];
}

__global__ void __launch_bounds__(128) device_function_kernel(int op, long long n, const double *in, double *out)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double *a = in + i * unit_n_in(op);
    double *o = out + i * unit_n_out(op);
    auto vec = [&](int k) { return Vec3{a[k], a[k + 1], a[k + 2]}; };
    auto put = [&](int k, const Vec3 &v) { o[k] = v.x; o[k + 1] = v.y; o[k + 2] = v.z; };
    auto tri = [&](int k) {   // corner and edges, as the upload stores a triangle
        Tri t;
        t.A = vec(k);
        t.E1 = Vec3{sub_(a[k + 3], a[k]), sub_(a[k + 4], a[k + 1]), sub_(a[k + 5], a[k + 2])};
        t.E2 = Vec3{sub_(a[k + 6], a[k]), sub_(a[k + 7], a[k + 1]), sub_(a[k + 8], a[k + 2])};
        return t;
    };
    switch (op) {
    case 0: o[0] = dot3(vec(0), vec(3)); break;
    case 1: put(0, cross3(vec(0), vec(3))); break;
    case 2: put(0, normalize3(vec(0))); break;
    case 3: put(0, triangle_normal(tri(0))); break;
    case 4: put(0, matvec3(a, vec(9))); break;
    case 5: o[0] = line_circle(Vec3{0.0, a[0], a[1]}, Vec3{0.0, a[2], a[3]}, a[4]); break;
    case 6: o[0] = line_sphere(vec(0), vec(3), a[6]); break;
    case 7: {
        EllipsoidConsts e;   // as ellipsoid_consts_kernel fills them
        for (int k = 0; k < 3; ++k) {
            e.ax[k] = a[6 + k];
            e.ax_rc[k] = rcp_refined(e.ax[k]);
            e.axsq[k] = mul_(e.ax[k], e.ax[k]);
            e.ax_isq[k] = rcp_(e.axsq[k]);
            e.axsq_rc[k] = rcp_refined(e.axsq[k]);
        }
        o[0] = line_ellipsoid(vec(0), vec(3), e);
        break;
    }
    case 8: o[0] = ray_triangle(tri(0), vec(9), vec(12)); break;
    case 9: {
        Vec3 r0 = vec(0), s = vec(3);
        reflect(r0, s, a[6], vec(7), a[10]);
        put(0, r0);
        put(3, s);
        break;
    }
    case 10: {
        Vec3 r0 = vec(0);
        cross_membrane(r0, vec(3), a[6], vec(7), a[10]);
        put(0, r0);
        break;
    }
    default: {   // 11-14: the subvoxel range lookups, from the guess-and-fix-up searches the mesh walk uses
        const double *xs = a + 3;
        const int len = (int)a[2];
        const double V = fabs(xs[len - 1] - xs[0]), inv_h = (double)(len - 1) / (xs[len - 1] - xs[0]);
        const bool upper = op == 12 || op == 14;
        const double x = upper ? fmax(a[0], a[1]) : fmin(a[0], a[1]);
        if (op <= 12) {
            o[0] = (double)(upper ? ul_overlap(xs, len, x, inv_h) : ll_overlap(xs, len, x, inv_h));
        } else {
            double nq;
            int base;
            axis_limit(xs, len, V, 1.0 / V, inv_h, x, upper, nq, base);
            o[0] = (double)__double2ll_rz(fma_(nq, (double)(len - 1), (double)base));
        }
        break;
    }
    }
}

// ---------------------------------------------------------------- self-test of sqrt_fast

// sqrt_fast(x) against __dsqrt_rn(x), bit for bit, on pseudo-random arguments: thread i tests `per_thread`
// values whose exponent is uniform in [e_lo, e_hi] and whose mantissa bits come from a 64-bit mixer.
__global__ void __launch_bounds__(256) sqrt_selftest_kernel(unsigned long long seed, int e_lo, int e_hi, int per_thread,
                                                            unsigned long long *n_bad, double *first_bad)
{
    unsigned long long z = seed + 0x9E3779B97F4A7C15ULL * ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x + 1);
    int bad = 0;
    for (int k = 0; k < per_thread; ++k) {
        z += 0x9E3779B97F4A7C15ULL;
        unsigned long long v = z;
        v = (v ^ (v >> 30)) * 0xBF58476D1CE4E5B9ULL;
        v = (v ^ (v >> 27)) * 0x94D049BB133111EBULL;
        v ^= v >> 31;
        const unsigned long long mant = v & 0x000FFFFFFFFFFFFFULL;
        const int e = e_lo + (int)((v >> 52) % (unsigned long long)(e_hi - e_lo + 1));
        const double x = __longlong_as_double((long long)(((unsigned long long)(e + 1023) << 52) | mant));
        const double a = sqrt_fast(x), b = __dsqrt_rn(x);
        if (__double_as_longlong(a) != __double_as_longlong(b)) {
            if (bad == 0 && atomicAdd(n_bad, 0ull) == 0ull) *first_bad = x;
            ++bad;
        }
    }
    if (bad) atomicAdd(n_bad, (unsigned long long)bad);
}

// ---------------------------------------------------------------- RNG state derivation

// state[i] = J^(start + i) * splitmix64(seed), J = the 2^64-step jump of xoroshiro128+ as a
// 128x128 matrix over GF(2).  pows[k] holds J^(2^k) column by column (column b = image of unit
// vector b); applying the matrices for the set bits of the index reproduces numba's sequential
// jump chain (numba/cuda/random.py:102-126, 225-241) without the O(N) host loop.  A thread
// derives kRngInitRun consecutive states: the first from the bits of its index, the others by one
// more application of J each.
constexpr int kRngInitRun = 8;

__device__ __forceinline__ void gf2_apply(const ulonglong2 *col, unsigned long long &s0, unsigned long long &s1)
{
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

__global__ void __launch_bounds__(256) rng_init_kernel(unsigned long long z, unsigned long long start,
                                                       long long n, const ulonglong2 *pows,
                                                       ulonglong2 *out)
{
    const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * kRngInitRun;
    if (i >= n) return;
    const unsigned long long idx = start + (unsigned long long)i;
    unsigned long long s0 = z, s1 = z;
    for (int k = 0; k < 64 && (idx >> k) != 0; ++k)
        if ((idx >> k) & 1ull) gf2_apply(pows + 128 * k, s0, s1);
    out[i] = make_ulonglong2(s0, s1);
    for (int j = 1; j < kRngInitRun && i + j < n; ++j) {
        gf2_apply(pows, s0, s1);
        out[i + j] = make_ulonglong2(s0, s1);
    }
}

}  // namespace dsb
