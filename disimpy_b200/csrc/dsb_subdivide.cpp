// dsb_subdivide.cpp -- host-side uniform-grid binning of mesh triangles.
//
// Native replacement of _mesh_space_subdivision (disimpy/substrates.py:467-536) and its helpers
// _triangle_box_overlap (:290-368, Akenine-Moller separating-axis test), _interval_sv_overlap
// (:371-419), _triangle_aabb (:422-441), _box_subvoxel_overlap (:444-464).  The reference runs a
// Python loop over faces (minutes for 1e6 triangles); the output ORDER is part of the hot path's
// contract (ascending face index inside each cell decides ties between equally distant
// triangles), so this produces the same arrays element for element.
//
// Plain IEEE double arithmetic, one rounding per operation like the reference's CPU code (build
// with -ffp-contract=off; no FMA instructions are enabled for host code).
#include "../../include/disimpy_b200.h"

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

namespace {

struct V3 {
    double c[3];
};

inline V3 cross(const V3 &a, const V3 &b)
{
    V3 r;
    r.c[0] = a.c[1] * b.c[2] - a.c[2] * b.c[1];
    r.c[1] = a.c[2] * b.c[0] - a.c[0] * b.c[2];
    r.c[2] = a.c[0] * b.c[1] - a.c[1] * b.c[0];
    return r;
}

inline double dot(const V3 &a, const V3 &b) { return a.c[0] * b.c[0] + a.c[1] * b.c[1] + a.c[2] * b.c[2]; }

// substrates.py:290-368.  lo / hi are the box corners closest to / furthest from the origin.
bool triangle_box_overlap(const V3 tri[3], const V3 &lo, const V3 &hi)
{
    V3 c, h, v[3];
    for (int i = 0; i < 3; ++i) {
        c.c[i] = (0.0 + lo.c[i] + hi.c[i]) / 2;
        h.c[i] = std::fabs(hi.c[i] - lo.c[i]) / 2;
    }
    for (int k = 0; k < 3; ++k)
        for (int i = 0; i < 3; ++i) v[k].c[i] = tri[k].c[i] - c.c[i];

    // triangle AABB against the box -- note the reference requires ALL three axes to separate
    bool all_above = true, all_below = true;
    for (int i = 0; i < 3; ++i) {
        double mn = std::min(v[0].c[i], std::min(v[1].c[i], v[2].c[i]));
        double mx = std::max(v[0].c[i], std::max(v[1].c[i], v[2].c[i]));
        if (!(mn > h.c[i])) all_above = false;
        if (!(mx < -h.c[i])) all_below = false;
    }
    if (all_above || all_below) return false;

    // plane of the triangle against the box corners
    V3 f[3];
    for (int i = 0; i < 3; ++i) {
        f[0].c[i] = v[1].c[i] - v[0].c[i];
        f[1].c[i] = v[2].c[i] - v[1].c[i];
        f[2].c[i] = v[0].c[i] - v[2].c[i];
    }
    const V3 normal = cross(f[0], f[1]);
    static const int sgn[8][3] = {{1, 1, 1}, {-1, -1, -1}, {-1, 1, 1}, {1, -1, -1},
                                  {1, -1, 1}, {-1, 1, -1}, {1, 1, -1}, {-1, -1, 1}};
    bool in_plane = false, all_behind = true, none_behind = true;
    for (int k = 0; k < 8; ++k) {
        V3 d;
        for (int i = 0; i < 3; ++i) d.c[i] = v[0].c[i] - (sgn[k][i] > 0 ? h.c[i] : -h.c[i]);
        double dp = dot(normal, d);
        bool behind = false;
        if (dp == 0)
            in_plane = true;
        else
            behind = dp > 0;
        if (behind) none_behind = false;
        else all_behind = false;
    }
    if (!in_plane && (all_behind || none_behind)) return false;

    // nine cross-product axes
    for (int i = 0; i < 3; ++i) {
        V3 e = {{0.0, 0.0, 0.0}};
        e.c[i] = 1.0;
        for (int j = 0; j < 3; ++j) {
            V3 a = cross(e, f[j]);
            V3 aa = {{std::fabs(a.c[0]), std::fabs(a.c[1]), std::fabs(a.c[2])}};
            double r = dot(h, aa);
            double p0 = dot(a, v[0]), p1 = dot(a, v[1]), p2 = dot(a, v[2]);
            double mn = std::min(p0, std::min(p1, p2)), mx = std::max(p0, std::max(p1, p2));
            if (mn > r || mx < -r) return false;
        }
    }
    return true;
}

// substrates.py:371-419: [ll, ul) of grid cells overlapping the interval, never empty
void interval_overlap(const double *xs, int64_t len, double x1, double x2, int64_t &ll, int64_t &ul)
{
    const double xmin = std::min(x1, x2), xmax = std::max(x1, x2);
    if (xmin <= xs[0])
        ll = 0;
    else if (xmin >= xs[len - 1])
        ll = len - 1;
    else {
        ll = 0;
        for (int64_t i = 0; i < len; ++i)
            if (xs[i] > xmin) {
                ll = i - 1;
                break;
            }
    }
    if (xmax >= xs[len - 1])
        ul = len - 1;
    else if (xmax <= xs[0])
        ul = 0;
    else {
        ul = len - 1;
        for (int64_t i = 0; i < len; ++i)
            if (!(xs[i] < xmax)) {
                ul = i;
                break;
            }
    }
    if (ll == ul) {
        if (ll != len - 1)
            ul = ul + 1;
        else
            ll = ll - 1;
    }
}

struct Subdivision {
    std::vector<int64_t> triangle_indices;
};

thread_local std::string g_sub_err;

}  // namespace

extern "C" {

// Bins n_faces triangles into the n_sv grid.  subvoxel_indices_out: caller-allocated
// (prod(n_sv), 2) int64.  The triangle list is kept inside *handle_out until
// dsb_mesh_subdivide_fetch copies it out (n_triangle_indices_out entries) and frees it.
int dsb_mesh_subdivide(const double *vertices, int64_t n_vertices, const int64_t *faces, int64_t n_faces,
                       const double *xs, const double *ys, const double *zs, const int64_t *n_sv,
                       int64_t *subvoxel_indices_out, int64_t *n_triangle_indices_out, void **handle_out)
{
    if (!vertices || !faces || !xs || !ys || !zs || !n_sv || !subvoxel_indices_out || !n_triangle_indices_out ||
        !handle_out)
        return DSB_EINVAL;
    for (int k = 0; k < 3; ++k)
        if (n_sv[k] <= 0) return DSB_EINVAL;
    const int64_t n_cells = n_sv[0] * n_sv[1] * n_sv[2];
    for (int64_t i = 0; i < 3 * n_faces; ++i)
        if (faces[i] < 0 || faces[i] >= n_vertices) return DSB_EINVAL;

    unsigned n_thr = std::max(1u, std::min(std::thread::hardware_concurrency(), 64u));
    if (n_faces < 2048) n_thr = 1;
    std::vector<std::vector<std::pair<int64_t, int64_t>>> found(n_thr);  // (cell, face), face ascending
    auto work = [&](unsigned tid) {
        const int64_t f0 = n_faces * tid / n_thr, f1 = n_faces * (tid + 1) / n_thr;
        auto &out = found[tid];
        for (int64_t fi = f0; fi < f1; ++fi) {
            V3 tri[3];
            for (int k = 0; k < 3; ++k)
                for (int i = 0; i < 3; ++i) tri[k].c[i] = vertices[3 * faces[3 * fi + k] + i];
            int64_t lo[3], hi[3];
            const double *grid[3] = {xs, ys, zs};
            for (int i = 0; i < 3; ++i) {
                double mn = std::min(tri[0].c[i], std::min(tri[1].c[i], tri[2].c[i]));
                double mx = std::max(tri[0].c[i], std::max(tri[1].c[i], tri[2].c[i]));
                interval_overlap(grid[i], n_sv[i] + 1, mn, mx, lo[i], hi[i]);
            }
            for (int64_t x = lo[0]; x < hi[0]; ++x)
                for (int64_t y = lo[1]; y < hi[1]; ++y)
                    for (int64_t z = lo[2]; z < hi[2]; ++z) {
                        V3 blo = {{xs[x], ys[y], zs[z]}}, bhi = {{xs[x + 1], ys[y + 1], zs[z + 1]}};
                        if (triangle_box_overlap(tri, blo, bhi))
                            out.emplace_back(x * n_sv[1] * n_sv[2] + y * n_sv[2] + z, fi);
                    }
        }
    };
    std::vector<std::thread> pool;
    for (unsigned t = 1; t < n_thr; ++t) pool.emplace_back(work, t);
    work(0);
    for (auto &t : pool) t.join();

    // counting sort by cell, stable in face order (threads hold ascending face ranges)
    std::vector<int64_t> count((size_t)n_cells + 1, 0);
    int64_t total = 0;
    for (auto &v : found) {
        for (auto &pr : v) ++count[(size_t)pr.first + 1];
        total += (int64_t)v.size();
    }
    for (int64_t c = 0; c < n_cells; ++c) count[(size_t)c + 1] += count[(size_t)c];
    for (int64_t c = 0; c < n_cells; ++c) {
        subvoxel_indices_out[2 * c] = count[(size_t)c];
        subvoxel_indices_out[2 * c + 1] = count[(size_t)c + 1];
    }
    Subdivision *sub = new Subdivision();
    sub->triangle_indices.resize((size_t)total);
    std::vector<int64_t> cursor(count.begin(), count.end() - 1);
    for (auto &v : found)
        for (auto &pr : v) sub->triangle_indices[(size_t)cursor[(size_t)pr.first]++] = pr.second;
    *n_triangle_indices_out = total;
    *handle_out = sub;
    return DSB_OK;
}

int dsb_mesh_subdivide_fetch(void *handle, int64_t *triangle_indices_out)
{
    Subdivision *sub = static_cast<Subdivision *>(handle);
    if (!sub) return DSB_EINVAL;
    if (triangle_indices_out && !sub->triangle_indices.empty())
        memcpy(triangle_indices_out, sub->triangle_indices.data(), sub->triangle_indices.size() * sizeof(int64_t));
    delete sub;
    return DSB_OK;
}

// Single overlap test, exported for the unit tests that mirror tests/test_substrates.py:293-314.
int dsb_triangle_box_overlap(const double *triangle9, const double *box6)
{
    V3 tri[3], lo, hi;
    for (int k = 0; k < 3; ++k)
        for (int i = 0; i < 3; ++i) tri[k].c[i] = triangle9[3 * k + i];
    for (int i = 0; i < 3; ++i) {
        lo.c[i] = box6[i];
        hi.c[i] = box6[3 + i];
    }
    return triangle_box_overlap(tri, lo, hi) ? 1 : 0;
}

// Exported for the unit tests that mirror tests/test_substrates.py:337-344.
int dsb_interval_sv_overlap(const double *xs, int64_t len, double x1, double x2, int64_t *ll, int64_t *ul)
{
    if (!xs || len < 2 || !ll || !ul) return DSB_EINVAL;
    interval_overlap(xs, len, x1, x2, *ll, *ul);
    return DSB_OK;
}

}  // extern "C"
