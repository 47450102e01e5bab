// Arithmetic core of the Hamming matcher (used by hamming.cu; plain C++ under SLAMFE_HD so that
// oracle/hamming_host_shim.cpp can compile it for the host and the CPU suite can check the prefix-XOR
// carry-save distance against a plain popcount on exhaustive bit patterns).
#pragma once
#include <stdint.h>

#include "hd.cuh"
#include "slamfe.h"

namespace slamfe {

constexpr int W = 16;  // u32 words per aligned descriptor row (64 B)

#ifdef __CUDACC__
using row4 = uint4;
#else
struct row4 {
    uint32_t x, y, z, w;
};
#endif

// ---- prefix-XOR carry-save Hamming distance --------------------------------------------------
// Both operands are held in "prefix form": word k stays w[k] for k odd, k = 0 and k = 15, and
// becomes w[0]^w[1]^...^w[k] for k = 2, 4, ..., 14.  XOR-ing the prefix forms of a query and a
// train row therefore gives X[k] = x[k] (the plain XOR word) at the unchanged positions and the
// running parity word x[0]^...^x[k] at the even ones — which is exactly the "ones" output of a
// chain of full adders that consumes two new words per step:
//     ones_0 = x0;  (ones_k, twos_k) = full_add(ones_{k-1}, x_{2k-1}, x_{2k})   =>  ones_k = X[2k]
// The carry of step k only needs ones_{k-1}, x_{2k-1} and ones_k (x_{2k} = ones_{k-1}^x_{2k-1}^ones_k):
//     twos_k = maj(a, b, a^b^s) with a = X[2k-2], b = X[2k-1], s = X[2k]            (LOP3 0xD4)
// so a descriptor pair costs 16 XOR + 7 carries (23 LOP3) and 9 POPC:
//     d = popc(X14) + popc(X15) + 2 * sum_k popc(twos_k)
// and every further full adder on equal-weight words trades one more POPC for 2 LOP3 (CS = 8..10).
// POPC issues on the XU pipe (16 lanes/clk/SM), LOP3 on the ALU pipe (64 lanes/clk/SM).
// On the device the three boolean functions are inline-PTX lop3 (written as C expressions, ptxas
// folds the operand XORs into them and re-derives the carry with 3 LOP3 instead of 1); the host
// build (oracle/hamming_host_shim.cpp, CPU tests) uses the equivalent C expressions.
SLAMFE_HD uint32_t lop3_carry_prefix(uint32_t a, uint32_t b, uint32_t s)
{
#ifdef __CUDA_ARCH__
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, 0xD4;" : "=r"(d) : "r"(a), "r"(b), "r"(s));
    return d;
#else
    const uint32_t c = a ^ b ^ s;
    return (a & b) | (c & (a ^ b));
#endif
}
SLAMFE_HD uint32_t fa_sum(uint32_t a, uint32_t b, uint32_t c)
{
#ifdef __CUDA_ARCH__
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
#else
    return a ^ b ^ c;
#endif
}
SLAMFE_HD uint32_t fa_carry(uint32_t a, uint32_t b, uint32_t c)
{
#ifdef __CUDA_ARCH__
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, 0xE8;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
#else
    return (a & b) | (c & (a ^ b));
#endif
}

// In-register conversion of 16 aligned words to prefix form.
SLAMFE_HD void to_prefix_form(uint32_t (&w)[W])
{
    uint32_t run = w[0] ^ w[1];
#pragma unroll
    for (int k = 2; k <= 14; k += 2) {
        run ^= w[k];
        const uint32_t odd = w[k + 1];
        w[k] = run;
        run ^= odd;
    }
}

// acc + popc(x) * (WEIGHT << 22) as one IMAD: the multiply-add runs on the FMA pipe, which is
// otherwise idle here, instead of IADD3/LEA on the saturated ALU pipe.
template <uint32_t WEIGHT>
SLAMFE_HD uint32_t popc_mad(uint32_t x, uint32_t acc)
{
#ifdef __CUDA_ARCH__
    uint32_t d;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(__popc(x)), "n"(WEIGHT << SLAMFE_KEY_IDX_BITS), "r"(acc));
    return d;
#else
    return static_cast<uint32_t>(__builtin_popcount(x)) * (WEIGHT << SLAMFE_KEY_IDX_BITS) + acc;
#endif
}

// Returns distance << 22 (the key without its index bits).
template <int CS>
SLAMFE_HD uint32_t hamming16_key(const uint32_t (&q)[W], const row4 &t0, const row4 &t1, const row4 &t2,
                                 const row4 &t3)
{
    static_assert(CS >= 7 && CS <= 10, "CS = number of full adders per descriptor pair");
    const uint32_t x0 = q[0] ^ t0.x, x1 = q[1] ^ t0.y, x2 = q[2] ^ t0.z, x3 = q[3] ^ t0.w;
    const uint32_t x4 = q[4] ^ t1.x, x5 = q[5] ^ t1.y, x6 = q[6] ^ t1.z, x7 = q[7] ^ t1.w;
    const uint32_t x8 = q[8] ^ t2.x, x9 = q[9] ^ t2.y, x10 = q[10] ^ t2.z, x11 = q[11] ^ t2.w;
    const uint32_t x12 = q[12] ^ t3.x, x13 = q[13] ^ t3.y, x14 = q[14] ^ t3.z, x15 = q[15] ^ t3.w;
    const uint32_t c0 = lop3_carry_prefix(x0, x1, x2);
    const uint32_t c1 = lop3_carry_prefix(x2, x3, x4);
    const uint32_t c2 = lop3_carry_prefix(x4, x5, x6);
    const uint32_t c3 = lop3_carry_prefix(x6, x7, x8);
    const uint32_t c4 = lop3_carry_prefix(x8, x9, x10);
    const uint32_t c5 = lop3_carry_prefix(x10, x11, x12);
    const uint32_t c6 = lop3_carry_prefix(x12, x13, x14);
    uint32_t acc = popc_mad<1>(x15, popc_mad<1>(x14, 0u));
    if (CS == 7) {
        acc = popc_mad<2>(c0, acc); acc = popc_mad<2>(c1, acc); acc = popc_mad<2>(c2, acc);
        acc = popc_mad<2>(c3, acc); acc = popc_mad<2>(c4, acc); acc = popc_mad<2>(c5, acc);
        return popc_mad<2>(c6, acc);
    }
    const uint32_t p0 = fa_sum(c0, c1, c2), f0 = fa_carry(c0, c1, c2);
    if (CS == 8) {
        acc = popc_mad<2>(p0, acc); acc = popc_mad<4>(f0, acc); acc = popc_mad<2>(c3, acc);
        acc = popc_mad<2>(c4, acc); acc = popc_mad<2>(c5, acc);
        return popc_mad<2>(c6, acc);
    }
    const uint32_t p1 = fa_sum(c3, c4, c5), f1 = fa_carry(c3, c4, c5);
    if (CS == 9) {
        acc = popc_mad<2>(p0, acc); acc = popc_mad<4>(f0, acc); acc = popc_mad<2>(p1, acc);
        acc = popc_mad<4>(f1, acc);
        return popc_mad<2>(c6, acc);
    }
    const uint32_t p2 = fa_sum(p0, p1, c6), f2 = fa_carry(p0, p1, c6);
    acc = popc_mad<4>(f0, acc); acc = popc_mad<4>(f1, acc); acc = popc_mad<2>(p2, acc);
    return popc_mad<4>(f2, acc);
}

}  // namespace slamfe
