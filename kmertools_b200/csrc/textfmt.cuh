// textfmt.cuh — fixed-width decimal text of normalised rows, produced on the GPU ("next" row N2).
//
// The reference writes every value with format!("{:.6}") (composition/src/oligo.rs:130-143,210-214):
// the EXACT binary value of the f64 quotient, rounded half-to-even at the 6th decimal — always 8
// characters for values in [0,1] (NUMBER_SIZE, oligo.rs:12), followed by the delimiter or '\n'
// (9 bytes per value, oligo.rs:170).  format6() reproduces that rounding with integer arithmetic on
// the double's mantissa (128-bit product m * 10^6, shifted by the exponent), so the bytes are
// identical to Rust's / glibc's correctly rounded output.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ktb {

__host__ __device__ __forceinline__ uint64_t mulhi64(uint64_t a, uint64_t b) {
#ifdef __CUDA_ARCH__
    return __umul64hi(a, b);
#else
    return (uint64_t)(((unsigned __int128)a * b) >> 64);
#endif
}

// q in [0, 1] -> round_half_even(q * 10^6) computed exactly from the bits of q
__host__ __device__ __forceinline__ uint32_t micro_units(double q) {
    union { double d; uint64_t u; } cv;
    cv.d = q;
    const uint64_t bits = cv.u;
    const int E = (int)((bits >> 52) & 0x7FF);
    const uint64_t M = bits & ((1ULL << 52) - 1);
    if (E == 0) return 0;                       // zero / denormal: far below half a unit
    if (E >= 1023) return 1000000u;             // q == 1.0 (values above 1 do not occur)
    const uint64_t m = M | (1ULL << 52);
    const int s = 1075 - E;                     // q = m * 2^-s, s in [53, 1074]
    if (s >= 128) return 0;                     // q < 2^-75
    const uint64_t lo = m * 1000000ULL;
    const uint64_t hi = mulhi64(m, 1000000ULL); // P = hi:lo < 2^73
    uint64_t ip, rem_hi, rem_lo, half_hi, half_lo;
    if (s >= 64) {
        const int t = s - 64;                   // 0..63
        ip = (t >= 64) ? 0 : (hi >> t);
        rem_hi = (t == 0) ? 0 : (hi & ((1ULL << t) - 1));
        rem_lo = lo;
        half_hi = (t == 0) ? 0 : (1ULL << (t - 1));
        half_lo = (t == 0) ? (1ULL << 63) : 0;
    } else {                                    // 53 <= s <= 63
        ip = (hi << (64 - s)) | (lo >> s);
        rem_hi = 0;
        rem_lo = lo & ((1ULL << s) - 1);
        half_hi = 0;
        half_lo = 1ULL << (s - 1);
    }
    const bool gt = (rem_hi > half_hi) || (rem_hi == half_hi && rem_lo > half_lo);
    const bool eq = (rem_hi == half_hi) && (rem_lo == half_lo);
    if (gt || (eq && (ip & 1))) ++ip;
    return (uint32_t)ip;
}

// 8 characters "d.dddddd" for a value in [0,1]
__host__ __device__ __forceinline__ void format6(double q, char *out) {
    uint32_t u = micro_units(q);
    uint32_t whole = u / 1000000u;
    u -= whole * 1000000u;
    out[0] = (char)('0' + whole);
    out[1] = '.';
#pragma unroll
    for (int i = 7; i >= 2; --i) {
        const uint32_t d = u / 10u;
        out[i] = (char)('0' + (u - d * 10u));
        u = d;
    }
}

// counts (n x dim u32) + totals -> text rows of dim*9 bytes: "0.dddddd" + delim, last delimiter '\n'
__global__ void __launch_bounds__(256) format_norm_kernel(const uint32_t *counts, const uint64_t *totals,
                                                          uint8_t *text, uint64_t n, uint32_t dim, char delim,
                                                          int norm_mode, int canonical) {
    const uint64_t nel = n * (uint64_t)dim;
    for (uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; e < nel;
         e += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t row = e / dim;
        const uint32_t col = (uint32_t)(e - row * dim);
        uint64_t t = totals[row];
        if (norm_mode == 2 && !canonical) t *= 2;   // pybindings raw-mode quirk
        if (t < 1) t = 1;
        const double q = (double)counts[e] / (double)t;   // same IEEE division as oligo.rs:256
        char buf[8];
        format6(q, buf);
        uint8_t *o = text + e * 9;
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = (uint8_t)buf[i];
        o[8] = (uint8_t)((col + 1 == dim) ? '\n' : delim);
    }
}

}  // namespace ktb
