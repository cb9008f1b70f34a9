// x0gen.cpp -- host helper: the reference mains' start vectors, element for element.
//
// seq/main.cpp:34-43 and par/L-BFGS-Wolfe.cu:458-465 draw x0 with libstdc++'s
//   std::mt19937 gen(seed); std::uniform_real_distribution<> dis(lo, hi); x0[i] = dis(gen);
// This restates that pipeline in plain integer/double arithmetic (MT19937 by Matsumoto &
// Nishimura; std::generate_canonical<double,53> = two 32-bit draws, low word first) so
// shards can be generated independently of the C++ standard library in use.
#include <math.h>
#include <stdint.h>

#include "../../include/lbfgsb200.h"

namespace {
struct Mt19937 {
    uint32_t mt[624];
    int idx;
    explicit Mt19937(uint32_t seed)
    {
        mt[0] = seed;
        for (int i = 1; i < 624; ++i) mt[i] = 1812433253u * (mt[i - 1] ^ (mt[i - 1] >> 30)) + (uint32_t)i;
        idx = 624;
    }
    uint32_t next()
    {
        if (idx >= 624) {
            for (int i = 0; i < 624; ++i) {
                uint32_t y = (mt[i] & 0x80000000u) | (mt[(i + 1) % 624] & 0x7fffffffu);
                mt[i] = mt[(i + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
            }
            idx = 0;
        }
        uint32_t y = mt[idx++];
        y ^= y >> 11;
        y ^= (y << 7) & 0x9d2c5680u;
        y ^= (y << 15) & 0xefc60000u;
        y ^= y >> 18;
        return y;
    }
};
} // namespace

extern "C" void lbfgsb200_x0_uniform(unsigned seed, double lo, double hi, size_t offset, size_t count,
                                     double *out_host)
{
    Mt19937 gen(seed);
    for (size_t i = 0; i < offset; ++i) { gen.next(); gen.next(); } // two draws per element
    for (size_t i = 0; i < count; ++i) {
        const double a = (double)gen.next();
        const double b = (double)gen.next();
        double c = (a + b * 4294967296.0) / 18446744073709551616.0;
        if (c >= 1.0) c = nextafter(1.0, 0.0);
        out_host[i] = c * (hi - lo) + lo;
    }
}
