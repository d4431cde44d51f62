// dsb_hostfill.cpp -- host-side initial positions for the analytic substrates.
//
// Native replacement of _fill_circle / _fill_sphere / _fill_ellipsoid
// (disimpy/simulations.py:353-399), which the reference JIT-compiles with Numba: sequential
// rejection sampling from the MT19937 stream that `_set_seed(seed)` / `np.random.seed(seed)`
// start (init_genrand seeding, 53-bit doubles from two 32-bit outputs), accepted points kept
// in stream order.  One rounding per operation, same expression order as the reference.
#include "../../include/disimpy_b200.h"

#include <cmath>
#include <cstdint>

namespace {

struct MT19937 {
    uint32_t mt[624];
    int idx;
    explicit MT19937(uint32_t seed)
    {
        mt[0] = seed;
        for (int i = 1; i < 624; ++i) mt[i] = 1812433253u * (mt[i - 1] ^ (mt[i - 1] >> 30)) + (uint32_t)i;
        idx = 624;
    }
    static uint32_t twist(uint32_t u, uint32_t v)
    {
        uint32_t y = (u & 0x80000000u) | (v & 0x7fffffffu);
        return (y >> 1) ^ ((v & 1u) ? 0x9908b0dfu : 0u);
    }
    void refill()
    {
        int k = 0;
        for (; k < 624 - 397; ++k) mt[k] = mt[k + 397] ^ twist(mt[k], mt[k + 1]);
        for (; k < 623; ++k) mt[k] = mt[k + 397 - 624] ^ twist(mt[k], mt[k + 1]);
        mt[623] = mt[396] ^ twist(mt[623], mt[0]);
        idx = 0;
    }
    uint32_t next32()
    {
        if (idx >= 624) refill();
        uint32_t y = mt[idx++];
        y ^= y >> 11;
        y ^= (y << 7) & 0x9d2c5680u;
        y ^= (y << 15) & 0xefc60000u;
        y ^= y >> 18;
        return y;
    }
    double next_double()  // genrand_res53, what random_sample() returns
    {
        uint32_t a = next32() >> 5, b = next32() >> 6;
        return (a * 67108864.0 + b) / 9007199254740992.0;
    }
};

}  // namespace

extern "C" {

// shape: 0 = disc (out is (n,2), scale[0] = radius), 1 = ball (out (n,3), scale[0] = radius),
// 2 = axis-aligned ellipsoid (out (n,3), scale = semi-axes).
int dsb_host_fill(int32_t shape, int64_t n, uint64_t seed, const double *scale, double *out)
{
    if (n < 0 || !scale || (n > 0 && !out) || shape < 0 || shape > 2 || seed > 0xffffffffULL) return DSB_EINVAL;
    MT19937 rng((uint32_t)seed);
    int64_t have = 0;
    if (shape == 0) {
        const double r = scale[0];
        while (have < n) {
            double x = (rng.next_double() - 0.5) * 2 * r;
            double y = (rng.next_double() - 0.5) * 2 * r;
            if (std::sqrt(x * x + y * y) < r) {
                out[2 * have] = x;
                out[2 * have + 1] = y;
                ++have;
            }
        }
    } else if (shape == 1) {
        const double r = scale[0];
        while (have < n) {
            double x = (rng.next_double() - 0.5) * 2 * r;
            double y = (rng.next_double() - 0.5) * 2 * r;
            double z = (rng.next_double() - 0.5) * 2 * r;
            if (std::sqrt(x * x + y * y + z * z) < r) {
                out[3 * have] = x;
                out[3 * have + 1] = y;
                out[3 * have + 2] = z;
                ++have;
            }
        }
    } else {
        const double a = scale[0], b = scale[1], c = scale[2];
        while (have < n) {
            double x = (rng.next_double() - 0.5) * 2 * a;
            double y = (rng.next_double() - 0.5) * 2 * b;
            double z = (rng.next_double() - 0.5) * 2 * c;
            double qx = x / a, qy = y / b, qz = z / c;
            if (qx * qx + qy * qy + qz * qz < 1) {
                out[3 * have] = x;
                out[3 * have + 1] = y;
                out[3 * have + 2] = z;
                ++have;
            }
        }
    }
    return DSB_OK;
}

}  // extern "C"
