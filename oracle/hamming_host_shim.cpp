// TEST INFRASTRUCTURE ONLY.  Host build of the matcher's distance arithmetic (csrc/hamming_core.cuh:
// prefix form + carry-save adders + weighted popcount accumulation) for the CPU test suite, which checks
// it against a plain popcount of the XOR on exhaustive bit patterns.  The product never loads this library.
#include <string.h>

#include "../67604-slam---video-navigation_b200/csrc/hamming_core.cuh"

using namespace slamfe;

static void load_row(const uint8_t *src, int desc_bytes, uint32_t (&w)[W])
{
    uint8_t buf[64];
    memset(buf, 0, sizeof(buf));
    memcpy(buf, src, desc_bytes);
    for (int k = 0; k < W; ++k)
        w[k] = (uint32_t)buf[4 * k] | ((uint32_t)buf[4 * k + 1] << 8) | ((uint32_t)buf[4 * k + 2] << 16) |
               ((uint32_t)buf[4 * k + 3] << 24);
}

template <int CS>
static uint32_t dist_cs(const uint32_t (&q)[W], const uint32_t (&t)[W])
{
    const row4 t0 = {t[0], t[1], t[2], t[3]}, t1 = {t[4], t[5], t[6], t[7]}, t2 = {t[8], t[9], t[10], t[11]},
               t3 = {t[12], t[13], t[14], t[15]};
    return hamming16_key<CS>(q, t0, t1, t2, t3);
}

// n pairs: rows i of q and t (stride bytes apart) -> key[i] = distance << 22 as the kernel computes it.
extern "C" void hamming_host_keys(const uint8_t *q, const uint8_t *t, long n, int stride, int desc_bytes, int cs,
                                  uint32_t *keys)
{
    for (long i = 0; i < n; ++i) {
        uint32_t a[W], b[W];
        load_row(q + i * stride, desc_bytes, a);
        load_row(t + i * stride, desc_bytes, b);
        to_prefix_form(a);
        to_prefix_form(b);
        keys[i] = cs == 7 ? dist_cs<7>(a, b) : cs == 8 ? dist_cs<8>(a, b) : cs == 10 ? dist_cs<10>(a, b) : dist_cs<9>(a, b);
    }
}
