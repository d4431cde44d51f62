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
#include <new>

namespace {

// MT19937 that works a whole state at a time: the twist, the tempering and the conversion to
// 53-bit doubles (genrand_res53, what random_sample() returns: two 32-bit outputs per double) are
// plain loops over arrays, which the compiler vectorises; the stream is the scalar generator's.
struct MT19937 {
    uint32_t mt[624];
    double buf[312];
    int pos;
    explicit MT19937(uint32_t seed)
    {
        mt[0] = seed;
        for (int i = 1; i < 624; ++i) mt[i] = 1812433253u * (mt[i - 1] ^ (mt[i - 1] >> 30)) + (uint32_t)i;
        pos = 312;
    }
    __attribute__((target_clones("avx2", "default"), optimize("O3"))) void refill()
    {
        uint32_t *m = mt;
        for (int k = 0; k < 227; ++k) {
            uint32_t y = (m[k] & 0x80000000u) | (m[k + 1] & 0x7fffffffu);
            m[k] = m[k + 397] ^ (y >> 1) ^ ((0u - (m[k + 1] & 1u)) & 0x9908b0dfu);
        }
        for (int k = 227; k < 623; ++k) {
            uint32_t y = (m[k] & 0x80000000u) | (m[k + 1] & 0x7fffffffu);
            m[k] = m[k - 227] ^ (y >> 1) ^ ((0u - (m[k + 1] & 1u)) & 0x9908b0dfu);
        }
        {
            uint32_t y = (m[623] & 0x80000000u) | (m[0] & 0x7fffffffu);
            m[623] = m[396] ^ (y >> 1) ^ ((0u - (m[0] & 1u)) & 0x9908b0dfu);
        }
        uint32_t t[624];
        for (int k = 0; k < 624; ++k) {
            uint32_t y = m[k];
            y ^= y >> 11;
            y ^= (y << 7) & 0x9d2c5680u;
            y ^= (y << 15) & 0xefc60000u;
            y ^= y >> 18;
            t[k] = y;
        }
        for (int k = 0; k < 312; ++k) buf[k] = ((t[2 * k] >> 5) * 67108864.0 + (t[2 * k + 1] >> 6)) / 9007199254740992.0;
        pos = 0;
    }
    double next_double()
    {
        if (pos >= 312) refill();
        return buf[pos++];
    }
};

}  // namespace

struct dsb_host_sampler {
    MT19937 rng;
    int shape;
    double scale[3];
    dsb_host_sampler(uint32_t seed, int shape_, const double *sc) : rng(seed), shape(shape_)
    {
        scale[0] = sc[0];
        scale[1] = shape_ == 2 ? sc[1] : 0.0;
        scale[2] = shape_ == 2 ? sc[2] : 0.0;
    }
    void next(int64_t n, double *out)
    {
        int64_t have = 0;
        if (shape == 0) {
            const double r = scale[0];
            while (have < n) {
                double x = (rng.next_double() - 0.5) * 2 * r;
                double y = (rng.next_double() - 0.5) * 2 * r;
                out[2 * have] = x;  // written unconditionally, kept by advancing: the accept
                out[2 * have + 1] = y;  // branch is a coin flip the predictor cannot learn
                have += std::sqrt(x * x + y * y) < r;
            }
        } else if (shape == 1) {
            const double r = scale[0];
            while (have < n) {
                double x = (rng.next_double() - 0.5) * 2 * r;
                double y = (rng.next_double() - 0.5) * 2 * r;
                double z = (rng.next_double() - 0.5) * 2 * r;
                out[3 * have] = x;
                out[3 * have + 1] = y;
                out[3 * have + 2] = z;
                have += std::sqrt(x * x + y * y + z * z) < r;
            }
        } else {
            const double a = scale[0], b = scale[1], c = scale[2];
            while (have < n) {
                double x = (rng.next_double() - 0.5) * 2 * a;
                double y = (rng.next_double() - 0.5) * 2 * b;
                double z = (rng.next_double() - 0.5) * 2 * c;
                double qx = x / a, qy = y / b, qz = z / c;
                out[3 * have] = x;
                out[3 * have + 1] = y;
                out[3 * have + 2] = z;
                have += qx * qx + qy * qy + qz * qz < 1;
            }
        }
    }
};

extern "C" {

// shape: 0 = disc (out is (n,2), scale[0] = radius), 1 = ball (out (n,3), scale[0] = radius),
// 2 = axis-aligned ellipsoid (out (n,3), scale = semi-axes).
int dsb_host_fill(int32_t shape, int64_t n, uint64_t seed, const double *scale, double *out)
{
    if (n < 0 || !scale || (n > 0 && !out) || shape < 0 || shape > 2 || seed > 0xffffffffULL) return DSB_EINVAL;
    dsb_host_sampler s((uint32_t)seed, shape, scale);
    s.next(n, out);
    return DSB_OK;
}

int dsb_host_sampler_create(int32_t shape, uint64_t seed, const double *scale, dsb_host_sampler **out)
{
    if (!out || !scale || shape < 0 || shape > 2 || seed > 0xffffffffULL) return DSB_EINVAL;
    *out = new (std::nothrow) dsb_host_sampler((uint32_t)seed, shape, scale);
    return *out ? DSB_OK : DSB_ENOMEM;
}

int dsb_host_sampler_next(dsb_host_sampler *sampler, int64_t n, double *out)
{
    if (!sampler || n < 0 || (n > 0 && !out)) return DSB_EINVAL;
    sampler->next(n, out);
    return DSB_OK;
}

int dsb_host_sampler_destroy(dsb_host_sampler *sampler)
{
    delete sampler;
    return DSB_OK;
}

}  // extern "C"
