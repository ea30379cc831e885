// long_kernel.cuh — CTA-per-sequence kernel for medium / long sequences, second generation (round 2).
//
// Same arithmetic as seq_kernel (kernels.cuh; closed form of kmer/src/kmer.rs:80-106 +
// composition/src/oligo.rs:231-259).  Three modes; the first two remove most of the per-k-mer and per-column work:
//
//   MODE_K7  (k = 7, canonical).  The two strands of an odd k-mer differ in the top bit of their MIDDLE base
//            (m vs 3-m), so a 16-bit key whose most significant digit is the middle base orders the strands:
//                key(s) = [b3 b4 b5 b6 | b0 b1 b2 b3]      (two bytes of the 2-bit packed stream, b3 twice)
//            min(key(f), key(r)) is the strand whose middle base is A or C, its bit 15 is 0, and after masking
//            the duplicated b3 it IS the byte offset of a dense 8192-bin u32 histogram (32 KB).  Both bytes of
//            a key are byte-aligned in one of four phase-shifted views of the packed stream, so TWO keys are
//            built with one PRMT per strand, canonicalised with one VIMNMX.U16x2 and masked with one LOP3:
//            ~4 integer instructions per k-mer instead of ~9 (two funnel shifts, two masks, test, select,
//            shift, multiply-add in seq_kernel mode 4).
//            Write-out: the bins of consecutive ranks are NOT in consecutive banks in this layout, so the
//            (bin -> rank) permutation is done through a host-built SCHEDULE: the 8192 (bin, rank) pairs are
//            split into 256 groups of 32 in which all source banks and all destination banks are distinct
//            (edge colouring of a 256-regular bipartite multigraph, api.cu), so a warp moves one group with one
//            conflict-free LDS, one conflict-free re-initialising STS and one conflict-free STS into a 32 KB
//            row image, which ONE bulk asynchronous copy (cp.async.bulk.global.shared::cta, SASS UBLKCP) then
//            drains to HBM while the CTA already counts its next sequence.  seq_kernel's gather measured 2.5
//            wavefronts per LDS/STS and 8192 STGs per row on the LSU.
//            Bins hold the FLOAT 2^23 + count (initial value 0x4B000000, incremented with integer atomics), so
//            the count -> float conversion of the normalisation is free.
//
//   MODE_FWD (3 <= k <= 5, canonical, long sequences).  Counts the FORWARD code only (4^k bins) and folds the two
//            strands at write-out, row[rank(c)] = hist[c] + hist[rc(c)] (once per column instead of once per
//            k-mer): no reverse-complement packing, no second extraction and no min in the inner loop.  Long contigs
//            count into lane-private replicas of the bins (RS) and are handed out longest length class first.
//
//   MODE_K8  (even k whose packed rank-space histogram fits shared memory: k = 8).  seq_kernel mode 5's arithmetic — rank from
//            two small shared-memory tables, 16-bit counters packed two to a word, linear unpacking sweep with plain
//            stores (the row, 128.5 KB, does not fit beside the histogram) — inside this kernel's loop: look-back lane
//            instead of a priming chunk, look-ahead across sequences.
//
// MODE_K7 and MODE_FWD write rows as one bulk copy from shared memory.  MODE_K7 and MODE_K8 look two sequences ahead
// (work ticket, offsets by cp.async, first bases before the write-out).  f64 output keeps seq_kernel.
#pragma once
#include "kernels.cuh"

namespace ktb {

constexpr int MODE_K7 = 0, MODE_FWD = 1, MODE_K8 = 2;
constexpr int LONG_WARPS = 8;
constexpr uint32_t K7_BINS = 8192;
constexpr uint32_t FLOAT_2P23 = 0x4B000000u;

struct LongParams {
    const uint8_t *bases;        // 16-byte aligned
    const uint64_t *offsets;
    uint64_t n;
    uint64_t total_bases;
    void *out;                   // n x dim of 4-byte elements (u32 / f32)
    uint64_t *totals;            // optional
    const uint32_t *sched;       // MODE_K7: [dim] (bin byte offset | rank byte offset << 16), group-major (see api.cu)
                                 // MODE_FWD: [dim] (byte offset of c | byte offset of rc(c) << 16) in rank order
    const uint32_t *even_tab;    // MODE_K8: [even_words] bitmap words then u16 prefixes (seq_kernel mode 5 / 7 tables, api.cu)
    uint32_t even_words;
    uint32_t *out_list;          // MODE_K8: sequences with more than 65535 windows are appended here (second launch)
    unsigned long long *out_count;
    unsigned long long *counter; // dynamic work counter (zeroed before launch)
    const uint32_t *list;        // groups to process (short_kernel's rejects); nullptr = every group
    const unsigned long long *list_count;
    uint32_t group_shift;        // log2 of the sequences per group (SHORT_G = 16)
    uint32_t grab;               // consecutive work items per trip to the counter
    uint32_t k;
    uint32_t dim;
    uint32_t hist_words;         // MODE_K7: 8192; MODE_FWD: 4^k + 4 (the word at 4^k stays zero: rc slot of palindromes)
    int norm_mode;
    int canonical;
};

__device__ __forceinline__ uint32_t smem_u32addr(const void *p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
// shared -> global bulk asynchronous copy (TMA engine), tracked by the issuing thread's bulk async-group
__device__ __forceinline__ void bulk_store(void *gdst, const void *ssrc, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 :: "l"(gdst), "r"(smem_u32addr(ssrc)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
// all earlier bulk copies of this thread have finished READING shared memory
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// generic-proxy writes to shared memory become visible to the async proxy (the bulk copy engine)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// 32-bit window of the 64-bit value (hi:lo) starting at bit S (compile-time, 0..32)
template <int S>
__device__ __forceinline__ uint32_t win32(uint32_t lo, uint32_t hi) {
    if constexpr (S == 0) return lo;
    else if constexpr (S == 32) return hi;
    else return __funnelshift_r(lo, hi, S);
}

// The 16 k = 7 keys of one lane (windows ending at bases 0..15 of the current chunk), as histogram byte
// offsets.  off[e] belongs to the window that ENDS at base e (validity bit 15 - e of the window mask).
template <int PHI>
__device__ __forceinline__ void k7_keys_phase(uint32_t cf, uint32_t cf_prev, uint32_t rc, uint32_t rc_prev,
                                              uint32_t *off) {
    // forward strand: F64 = cf_prev:cf, base i (-16..15) at bit 2*(15-i).  View B holds the bytes [b3 b4 b5 b6] of the
    // windows ending at e = PHI+12, PHI+8, PHI+4, PHI in bytes 0..3; view A the bytes [b0 b1 b2 b3] of the same windows.
    const uint32_t B = win32<6 - 2 * PHI>(cf, cf_prev);
    const uint32_t A = win32<12 - 2 * PHI>(cf, cf_prev);
    // reverse strand: R64 = rc(cf):rc(cf_prev), base i at bit 2*(i+16), complemented.  Byte j of RB / RA belongs to e = PHI+4j.
    const uint32_t RB = win32<2 * PHI + 20>(rc_prev, rc);
    const uint32_t RA = win32<2 * PHI + 26>(rc_prev, rc);
    // pair 0: low half e = PHI+12, high half e = PHI+8;  pair 1: low half e = PHI+4, high half e = PHI
    const uint32_t f0 = __byte_perm(A, B, 0x5140), r0 = __byte_perm(RA, RB, 0x6273);
    const uint32_t f1 = __byte_perm(A, B, 0x7362), r1 = __byte_perm(RA, RB, 0x4051);
    // the strands differ in the top bit of the key (middle base m vs 3-m): the minimum has bit 15 clear
    const uint32_t k0 = __vminu2(f0, r0) & 0x7FFC7FFCu;
    const uint32_t k1 = __vminu2(f1, r1) & 0x7FFC7FFCu;
    off[PHI + 12] = k0 & 0xFFFFu;
    off[PHI + 8] = k0 >> 16;
    off[PHI + 4] = k1 & 0xFFFFu;
    off[PHI] = k1 >> 16;
}

// ---- write-out of one row into the row image (normalisation fused), histogram re-initialised on the way.
// SMALL: every count and the divisor are below 2^23 (exact in f32, magic-constant conversion valid).
// the first two schedule iterations of a warp (MODE_K7), requested before the barrier that precedes the write-out
struct K7Sched {
    uint4 A, B;
    template <int NW>
    __device__ __forceinline__ void load_nw(const uint32_t *sched, uint32_t warp, uint32_t lane) {
        const uint4 *sched4 = reinterpret_cast<const uint4 *>(sched) + lane;
        A = __ldg(sched4 + warp * 32);
        B = (warp + NW < (K7_BINS >> 7)) ? __ldg(sched4 + (warp + NW) * 32) : make_uint4(0, 0, 0, 0);
    }
};

template <int OUT, bool NORM, int MODE, int NW, bool SMALL, int RS>
__device__ __forceinline__ void long_write_row(const LongParams &p, uint8_t *hbytes, uint32_t *stage, uint64_t dv,
                                               const K7Sched &ks, uint64_t row_seq) {
    using T = typename OutT<OUT>::type;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float dF = (float)dv;
    const float rinv = __frcp_rn(dF);
    const double dD = (double)dv;
    (void)dD;
    if constexpr (MODE == MODE_K7) {
        uint8_t *sbytes = reinterpret_cast<uint8_t *>(stage);
        const uint4 *sched4 = reinterpret_cast<const uint4 *>(p.sched) + lane;
        constexpr uint32_t niter = K7_BINS >> 7;   // 128 (bin, rank) pairs per warp iteration
        const float nK = -8388608.0f * rinv;
        (void)nK;
        // one warp iteration = 128 (bin, rank) pairs = four conflict-free groups
        auto move4 = [&](const uint4 e4) {
            const uint32_t ee[4] = {e4.x, e4.y, e4.z, e4.w};
            uint32_t bits[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                uint32_t *bin = reinterpret_cast<uint32_t *>(hbytes + (ee[q] & 0xFFFFu));
                bits[q] = *bin;
                *bin = FLOAT_2P23;
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                T val;
                if constexpr (OUT == OUT_U32) {
                    val = bits[q] - FLOAT_2P23;
                } else if constexpr (SMALL) {
                    const float F = __uint_as_float(bits[q]);   // 2^23 + count
                    const float cnt = F - 8388608.0f;
                    if constexpr (NORM) {
                        const float q0 = fmaf(F, rinv, nK);     // RN(count * rinv), as quot_f32
                        const float rem = fmaf(-q0, dF, cnt);
                        val = fmaf(rem, rinv, q0);
                    } else {
                        val = cnt;
                    }
                } else {
                    const uint32_t cnt = bits[q] - FLOAT_2P23;
                    val = NORM ? (float)((double)cnt / dD) : (float)cnt;
                }
                *reinterpret_cast<T *>(sbytes + (ee[q] >> 16)) = val;
            }
        };
        // the schedule comes from L2: two iterations in flight, registers A / B alternate (no copies between them)
        static_assert(niter % (2 * NW) == 0 || NW == 10, "iterations per warp");
        // (each register set is reloaded right after the iteration that consumed it: no copies, and the load has the other
        // set's iteration — ~500 clocks at 24 warps per SM — to come back from L2)
        uint4 A = ks.A, B = ks.B;
        for (uint32_t w = warp; w < niter; w += 2 * NW) {
            move4(A);
            if (w + 2 * NW < niter) A = __ldg(sched4 + (w + 2 * NW) * 32);
            if (w + NW < niter) {
                move4(B);
                if (w + 3 * NW < niter) B = __ldg(sched4 + (w + 3 * NW) * 32);
            }
        }
    } else if constexpr (MODE == MODE_K8) {
        // rank space, 16-bit counters packed two to a word (rank r -> word r/2, half r&1): linear sweep, one 128-bit shared
        // load = 8 counts = two 128-bit stores; zeroed on the way.  dim % 8 == 0 (checked on the host).
        (void)ks; (void)stage;
        uint32_t *hist = reinterpret_cast<uint32_t *>(hbytes);
        T *row = reinterpret_cast<T *>(p.out) + row_seq * (uint64_t)p.dim;
        for (uint32_t w = tid * 4u; w < p.hist_words; w += NW * 32 * 4u) {
            const uint4 hv = *reinterpret_cast<const uint4 *>(hist + w);
            *reinterpret_cast<uint4 *>(hist + w) = make_uint4(0, 0, 0, 0);
            const uint32_t hw[4] = {hv.x, hv.y, hv.z, hv.w};
            T e[8];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                e[2 * q] = cvt_count<OUT, NORM, SMALL>(hw[q] & 0xFFFFu, dF, rinv, dD);
                e[2 * q + 1] = cvt_count<OUT, NORM, SMALL>(hw[q] >> 16, dF, rinv, dD);
            }
            T *dst = row + 2u * w;
            if constexpr (OUT == OUT_F32) {
                reinterpret_cast<float4 *>(dst)[0] = make_float4(e[0], e[1], e[2], e[3]);
                reinterpret_cast<float4 *>(dst)[1] = make_float4(e[4], e[5], e[6], e[7]);
            } else {
                reinterpret_cast<uint4 *>(dst)[0] = make_uint4(e[0], e[1], e[2], e[3]);
                reinterpret_cast<uint4 *>(dst)[1] = make_uint4(e[4], e[5], e[6], e[7]);
            }
        }
    } else if constexpr (RS == 0) {
        const uint32_t zero_off = (p.hist_words - 4u) << 2;   // byte offset of the word at 4^k
        for (uint32_t j = tid; j < p.dim; j += NW * 32) {
            const uint32_t e = __ldg(p.sched + j);
            uint32_t *a = reinterpret_cast<uint32_t *>(hbytes + (e & 0xFFFFu));
            uint32_t *b = reinterpret_cast<uint32_t *>(hbytes + (e >> 16));
            const uint32_t cnt = *a + *b;   // palindromes: b is the always-zero word
            *a = 0;
            if ((e >> 16) != zero_off) *b = 0;   // (the zero word is shared by every palindrome: never stored to)
            reinterpret_cast<T *>(stage)[j] = cvt_count<OUT, NORM, SMALL>(cnt, dF, rinv, dD);
        }
    } else {
        // 2^RS lane-private replicas per bin (bin c, replica r at byte (c << (2 + RS)) + 4 r): a column is the sum of the
        // replicas of c and of rc(c); lane l starts at replica l so the 32 lanes of a load hit 32 banks
        constexpr uint32_t R = 1u << RS;
        for (uint32_t j = tid; j < p.dim; j += NW * 32) {
            const uint32_t e = __ldg(p.sched + j);
            const uint32_t ca = (e & 0xFFFFu) << RS, cb = (e >> 16) << RS;
            uint32_t cnt = 0;
#pragma unroll 8
            for (uint32_t r = 0; r < R; ++r) {
                const uint32_t rr = ((r + lane) & (R - 1u)) << 2;
                cnt += *reinterpret_cast<const uint32_t *>(hbytes + ca + rr) + *reinterpret_cast<const uint32_t *>(hbytes + cb + rr);
            }
            reinterpret_cast<T *>(stage)[j] = cvt_count<OUT, NORM, SMALL>(cnt, dF, rinv, dD);
        }
        __syncthreads();   // every column has read its bins: clear them with a linear sweep
        for (uint32_t i = tid * 4u; i < p.hist_words; i += NW * 32 * 4u) *reinterpret_cast<uint4 *>(hbytes + 4u * i) = make_uint4(0, 0, 0, 0);
    }
}

// ---- longest first.  Contigs of an assembly span orders of magnitude in length, and the work queue hands them out in
// input order: with a 500 kbp contig drawn near the end, its CTA runs long after the others have drained (20 % of the SM
// time of a 5,000-contig batch, 5 % at 20,000: profiles/r2_ncu_contigs_summary.txt, sm__cycles_active avg against
// elapsed).  Two small kernels build the order the queue should use: sequences binned by the log2 of their length, the
// long classes first (within a class lengths differ by less than 2x, which is all longest-processing-time-first needs).
struct OrderParams {
    const uint64_t *offsets;
    uint64_t n;
    const uint32_t *list;                  // groups to take (short_kernel's rejects) or nullptr = every sequence
    const unsigned long long *list_count;
    uint32_t group_shift;
    uint32_t *order;                       // out: sequence indices, long classes first
    unsigned long long *cls;               // [0,64) count per class, [64,128) fill cursor per class, [128] total (zeroed)
};

__device__ __forceinline__ uint64_t order_items(const OrderParams &p) {
    return p.list ? ((uint64_t)*p.list_count << p.group_shift) : p.n;
}
__device__ __forceinline__ uint64_t order_seq_of(const OrderParams &p, uint64_t item) {
    if (!p.list) return item;
    const uint64_t gi = item >> p.group_shift;
    return ((uint64_t)p.list[gi] << p.group_shift) + (item - (gi << p.group_shift));
}
__device__ __forceinline__ uint32_t order_class(const OrderParams &p, uint64_t seq) {
    return 63u - (uint32_t)__clzll((long long)((p.offsets[seq + 1] - p.offsets[seq]) | 1ull));
}

__global__ void __launch_bounds__(256) order_count_kernel(const OrderParams p) {
    __shared__ uint32_t s_cnt[64];
    if (threadIdx.x < 64) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    const uint64_t nitems = order_items(p);
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nitems; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t seq = order_seq_of(p, i);
        if (seq < p.n) atomicAdd(&s_cnt[order_class(p, seq)], 1u);
    }
    __syncthreads();
    if (threadIdx.x < 64 && s_cnt[threadIdx.x]) atomicAdd(p.cls + threadIdx.x, (unsigned long long)s_cnt[threadIdx.x]);
}

__global__ void __launch_bounds__(256) order_scatter_kernel(const OrderParams p) {
    __shared__ unsigned long long s_base[64];
    if (threadIdx.x == 0) {
        unsigned long long run = 0;
        for (int c = 63; c >= 0; --c) { s_base[c] = run; run += p.cls[c]; }
        if (blockIdx.x == 0) p.cls[128] = run;
    }
    __syncthreads();
    const uint64_t nitems = order_items(p);
    const uint32_t lane = threadIdx.x & 31;
    // whole warps iterate together: the lanes of a class take their slots with ONE atomic (uniform lengths would otherwise
    // queue every sequence of the batch on one address)
    for (uint64_t i0 = (uint64_t)blockIdx.x * blockDim.x + (threadIdx.x & ~31u); i0 < nitems; i0 += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t i = i0 + lane;
        const uint64_t seq = i < nitems ? order_seq_of(p, i) : p.n;
        const bool valid = seq < p.n;
        const uint32_t c = valid ? order_class(p, seq) : 64u;
        const uint32_t peers = __match_any_sync(0xffffffffu, c);
        const uint32_t leader = (uint32_t)__ffs((int)peers) - 1u;
        unsigned long long first = 0;
        if (valid && lane == leader) first = atomicAdd(p.cls + 64 + c, (unsigned long long)__popc(peers));
        first = __shfl_sync(0xffffffffu, first, (int)leader);
        if (valid) p.order[s_base[c] + first + (unsigned long long)__popc(peers & ((1u << lane) - 1u))] = (uint32_t)seq;
    }
}

// KT: compile-time k of MODE_FWD (0 = p.k); RS: log2 of the lane-private replicas per bin (MODE_FWD, long contigs:
// k <= 4 has so few bins that the 32 lanes of an atomic keep hitting the same banks — 3.9 wavefronts per ATOMS on
// config 4 — so bin c of lane l lives at word (c << RS) + (l mod 2^RS): one wavefront per ATOMS, folded at write-out)
template <int OUT, bool NORM, int MODE, int NW, int KT = 0, int RS = 0>
__global__ void __launch_bounds__(NW * 32, (MODE == MODE_K7 || MODE == MODE_K8) ? 3 : (NW >= 8 ? 4 : 8)) long_kernel(const LongParams p) {
    static_assert(OUT == OUT_U32 || OUT == OUT_F32, "f64 rows keep seq_kernel");
    static_assert(NW == 4 || NW == 8 || NW == 10, "warps per CTA");
    extern __shared__ __align__(128) uint32_t lsm[];
    __shared__ unsigned long long s_la[5];      // thread 0's look-ahead stream of work items
    __shared__ __align__(8) unsigned long long s_nx[3][3];   // look-ahead work items (three slots): sequence or a marker, its two offsets
    __shared__ uint32_t s_total[2];

    uint32_t *hist = lsm;
    uint32_t *stage = lsm + ((p.hist_words + 31u) & ~31u);   // row image: dim x 4 bytes, 128-byte aligned
    uint8_t *hbytes = reinterpret_cast<uint8_t *>(hist);

    const int tid = threadIdx.x, lane = tid & 31;
    const uint32_t warp = tid >> 5;
    constexpr uint32_t FULL = 0xffffffffu;
    // MODE_K7 bins hold the float 2^23 + count (see header)
    constexpr uint32_t INIT = (MODE == MODE_K7) ? FLOAT_2P23 : 0u;

    const uint32_t gshift = p.group_shift;
    const uint64_t ngroups = p.list ? (uint64_t)*p.list_count : (p.n + (1ull << gshift) - 1) >> gshift;
    const uint64_t nitems = ngroups << gshift;
    if ((uint64_t)blockIdx.x >= nitems) return;
    for (uint32_t i = tid; i < p.hist_words; i += NW * 32) hist[i] = INIT;
    // MODE_K8: rank tables behind the histogram (bitmap words, then every second u16 prefix: seq_kernel mode 5's layout)
    uint32_t *s_bitmap = hist + ((p.hist_words + 3u) & ~3u);
    uint16_t *s_prefix = reinterpret_cast<uint16_t *>(s_bitmap + p.even_words);
    if constexpr (MODE == MODE_K8) {
        for (uint32_t i = tid; i < p.even_words; i += NW * 32) s_bitmap[i] = __ldg(p.even_tab + i);
        const uint16_t *gp = reinterpret_cast<const uint16_t *>(p.even_tab + p.even_words);
        for (uint32_t i = tid; i < p.even_words / 2; i += NW * 32) s_prefix[i] = gp[2 * i];
    }
    (void)s_bitmap; (void)s_prefix;
    if (tid == 0) { s_total[0] = 0; s_total[1] = 0; }
    __syncthreads();

    const uint32_t k = (MODE == MODE_K7) ? 7u : (MODE == MODE_K8) ? 8u : (KT ? (uint32_t)KT : p.k);
    const uint32_t kmaskR = ((1u << (2 * k)) - 1u) << (2 + RS);            // code pre-scaled to the byte offset of its bin
    const uint32_t lane_off = ((uint32_t)lane & ((1u << RS) - 1u)) << 2;   // this lane's replica
    (void)kmaskR; (void)lane_off;
    using T = typename OutT<OUT>::type;
    T *out = reinterpret_cast<T *>(p.out);
    const uint4 filler = make_uint4(0x41414141u, 0x41414141u, 0x41414141u, 0x41414141u);
    const unsigned long long grab = p.grab ? p.grab : 1u;
    uint32_t it = 0;

    // ---- work items, two sequences ahead.  A sequence costs three dependent trips to memory before its first k-mer
    // (work counter -> offsets -> bases); at one sequence per ~6 us and CTA that chain was a third of the time.  Thread 0
    // runs a look-ahead stream of items: while sequence i is counted it draws the ticket of sequence i + 2, fetches its
    // offsets during the write-out of i and publishes them in s_nx; every warp requests the first bases of sequence i + 1
    // (known since the previous iteration) BEFORE the write-out of i.
    constexpr unsigned long long ITEM_DONE = ~0ull, ITEM_SKIP = ~0ull - 1ull;
    // (the stream's state lives in shared memory, touched by thread 0 only, and the offsets travel global -> shared with
    // cp.async: nothing of the look-ahead occupies registers except the ticket that is in flight during the count loop)
    unsigned long long tkt = 0;
    bool want = false;
    if (tid == 0) { s_la[0] = 0; s_la[1] = 0; s_la[2] = 0; s_la[3] = ~0ull; s_la[4] = 0; }
    auto la_ticket = [&]() {
        want = false;
        if (tid == 0) {
            const unsigned long long pos = s_la[0], end = s_la[1], dry = s_la[2];
            want = !dry && pos == end;
            if (want) tkt = atomicAdd(p.counter, grab);
        }
    };
    auto la_resolve = [&](const uint32_t slot) {   // thread 0: next item of the stream -> s_nx[slot] (offsets asynchronously)
        if (tid != 0) return;
        // (the whole state is read at once: one shared-memory round trip on thread 0's path instead of a chain of them)
        unsigned long long pos = s_la[0];
        bool dry = s_la[2] != 0;
        const unsigned long long last_gi = s_la[3], last_g = s_la[4];
        if (want) {
            pos = tkt;
            dry = tkt >= nitems;
            s_la[1] = min(tkt + grab, (unsigned long long)nitems);
            if (dry) s_la[2] = 1;
        }
        unsigned long long nseq = ITEM_DONE;
        if (!dry) {
            s_la[0] = pos + 1;
            const uint64_t gi = pos >> gshift;
            uint64_t g = gi;
            if (p.list) {   // one look-up per group of 2^gshift items
                g = last_g;
                if (last_gi != gi) { g = p.list[gi]; s_la[3] = gi; s_la[4] = g; }
            }
            const uint64_t seq = (g << gshift) + (pos - (gi << gshift));
            nseq = seq < p.n ? seq : ITEM_SKIP;   // (padding of the last group)
        }
        s_nx[slot][0] = nseq;
        if (nseq < ITEM_SKIP) {
            const uint32_t dst = smem_u32addr(&s_nx[slot][1]);
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"(dst), "l"(p.offsets + nseq) : "memory");
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"(dst + 8u), "l"(p.offsets + nseq + 1) : "memory");
        } else {
            s_nx[slot][1] = 0; s_nx[slot][2] = 0;
        }
    };
    auto la_arrived = [&]() {   // before the barrier that publishes the slot
        if (tid == 0) asm volatile("cp.async.wait_all;" ::: "memory");
    };

    // what a warp does with a sequence: contiguous runs of whole 32-chunk steps per warp.  A warp that starts inside the
    // sequence loads, in lane 0 of its first step, the LAST chunk of the previous warp as look-back (that lane emits
    // nothing), so no warp has to load and decode a priming chunk on its own: such a warp covers 32 * steps - 1 new
    // chunks, and the step count carries NW - 1 chunks of slack.  Sequences of fewer steps than warps use the first warps.
    struct Plan { uint64_t cbase; uint32_t nch, w0, w1, flags; };   // flags: 1 = look-back lane, 2 = guarded loads
    auto make_plan = [&](const uint64_t s0, const uint64_t s1) -> Plan {
        Plan pl{0, 0, 0, 0, 0};
        if (s1 - s0 >= k) {
            pl.cbase = s0 >> 4;
            pl.nch = (uint32_t)(((s1 - 1) >> 4) - pl.cbase) + 1u;
            const uint32_t nsteps = (pl.nch + (NW - 1) + 31) >> 5;
            const uint32_t q0 = nsteps >= NW ? (warp * nsteps) / NW : min(warp, nsteps);
            const uint32_t q1 = nsteps >= NW ? ((warp + 1) * nsteps) / NW : min(warp + 1, nsteps);
            const uint32_t w0 = q0 > 0 ? (q0 << 5) - warp : 0u;   // lane 0 of the first step re-reads chunk w0
            const uint32_t w1 = min(pl.nch, (q1 << 5) - warp);
            if (q0 < q1 && w0 < w1) {
                pl.w0 = w0; pl.w1 = w1;
                pl.flags = (q0 > 0 ? 1u : 0u) | ((((pl.cbase + pl.nch) << 4) > p.total_bases) ? 2u : 0u);
            }
        }
        return pl;
    };
    auto fetch = [&](const Plan &pl, const uint32_t c) -> uint4 {
        if (c >= pl.w1) return filler;
        if (pl.flags & 2u) return load16_guarded(p.bases, (pl.cbase + c) << 4, p.total_bases);
        return __ldg(reinterpret_cast<const uint4 *>(p.bases) + pl.cbase + c);
    };

    // Only k = 7 looks ahead.  The small histograms of k <= 5 run 4 - 8 CTAs per SM, which hide those latencies already
    // (measured with look-ahead: -3 % on 10 kbp reads at k = 5), and a CTA that reserves two 500 kbp contigs in advance
    // lengthens the tail of the launch (-5 % on config 4).
    constexpr bool LA = (MODE == MODE_K7 || MODE == MODE_K8);
    // prologue: the first two items, with their latencies exposed once per CTA
    la_ticket(); la_resolve(0);
    if constexpr (LA) { la_ticket(); la_resolve(1); }
    la_arrived();
    __syncthreads();
    unsigned long long seq = s_nx[0][0];
    uint32_t head_mask, tail_mask;
    Plan pl;
    {
        const unsigned long long s0 = s_nx[0][1], s1 = s_nx[0][2];
        head_mask = 0xFFFFu >> (uint32_t)(s0 & 15);
        tail_mask = ~(0xFFFFu >> ((uint32_t)((s1 - 1) & 15) + 1u)) & 0xFFFFu;
        pl = make_plan(s0, s1);
    }
    __syncthreads();
    // three slots in rotation: the NEXT item's record (read in this iteration), the one thread 0 is filling, and the one the
    // current item came from (still being read by slower warps when thread 0 starts to fill)
    uint32_t slot = 1;
    // two steps of bases in flight per warp (two buffers used alternately, each reloaded right after it has been decoded)
    uint4 buf_a = fetch(pl, pl.w0 + lane);
    uint4 buf_b = fetch(pl, pl.w0 + 32 + lane);

    while (seq != ITEM_DONE) {
        if constexpr (LA) la_ticket();
        // MODE_K8: 16-bit counters — a sequence that could hold more than 65535 windows goes to the second launch (out_list)
        const bool over = MODE == MODE_K8 && pl.nch >= 4096u;
        const bool real = seq != ITEM_SKIP && !over;   // uniform for the CTA
        if (over && tid == 0) p.out_list[atomicAdd(p.out_count, 1ULL)] = (uint32_t)seq;
        uint32_t mine = 0;
        if (pl.w0 < pl.w1 && !over) {
            const uint32_t nch = pl.nch, w0 = pl.w0, w1 = pl.w1;
            const bool lookback = pl.flags & 1u;
            uint32_t carry_cf = 0, carry_vm = 0;
            auto do_step = [&](uint4 &buf, const uint32_t c0) {
                const uint32_t c = c0 + lane;
                uint32_t cf, vm;
                decode16(buf, cf, vm);
                if (c0 + 64 < w1) buf = fetch(pl, c + 64);
                if (c >= w1) { vm = 0; cf = (uint32_t)lane * 0x9E3779B1u; }   // idle lanes add 0 at scattered bins
                if (c == 0) vm &= head_mask;
                if (c == nch - 1) vm &= tail_mask;
                uint32_t cf_prev = __shfl_up_sync(FULL, cf, 1);
                uint32_t vm_prev = __shfl_up_sync(FULL, vm, 1);
                if (lane == 0) { cf_prev = carry_cf; vm_prev = carry_vm; }
                carry_cf = __shfl_sync(FULL, cf, 31);
                carry_vm = __shfl_sync(FULL, vm, 31);
                uint32_t vw = window_mask((vm_prev << 16) | vm, k) & 0xFFFFu;
                const bool silent = lookback && c0 == w0 && lane == 0;   // the look-back lane of this warp's first step
                if (silent) vw = 0;
                mine += __popc(vw);
                uint32_t off[16];   // histogram byte offsets; off[e] = window ending at base e
                if constexpr (MODE == MODE_K7) {
                    const uint32_t rc = revcomp_pack(cf), rc_prev = revcomp_pack(cf_prev);
                    k7_keys_phase<0>(cf, cf_prev, rc, rc_prev, off);
                    k7_keys_phase<1>(cf, cf_prev, rc, rc_prev, off);
                    k7_keys_phase<2>(cf, cf_prev, rc, rc_prev, off);
                    k7_keys_phase<3>(cf, cf_prev, rc, rc_prev, off);
                } else if constexpr (MODE == MODE_K8) {
                    // canonical code, then its rank from the two shared-memory tables (seq_kernel mode 5): off[e] = rank
                    const uint64_t F64 = ((uint64_t)cf_prev << 32) | cf;
                    const uint64_t R64 = ((((uint64_t)revcomp_pack(cf)) << 32) | revcomp_pack(cf_prev)) >> (2 * (17 - 8));
                    constexpr uint32_t kmask8 = 0xFFFFu;
#pragma unroll
                    for (int e = 0; e < 16; ++e) {
                        const uint32_t f = (e < 15 ? (uint32_t)(F64 >> (2 * (15 - e))) : cf) & kmask8;
                        const uint32_t r = (uint32_t)(R64 >> (2 * e)) & kmask8;
                        const uint32_t c = min(f, r);
                        const uint32_t w = c >> 5;
                        const uint2 bw = reinterpret_cast<const uint2 *>(s_bitmap)[w >> 1];   // words w & ~1, w | 1
                        const bool odd = (w & 1u) != 0u;
                        const uint32_t below = (odd ? bw.y : bw.x) & ((1u << (c & 31u)) - 1u);
                        off[e] = (uint32_t)s_prefix[w >> 1] + (uint32_t)__popc(below) + (odd ? (uint32_t)__popc(bw.x) : 0u);
                    }
                } else {
                    const uint64_t F64 = ((uint64_t)cf_prev << 32) | cf;
#pragma unroll
                    for (int e = 0; e < 16; ++e) {
                        constexpr int up = 2 + RS;                       // window e sits at bit 2 (15 - e) of F64
                        const int sh = 2 * (15 - e) - up;
                        const uint32_t x = sh >= 0 ? (uint32_t)(F64 >> (sh >= 0 ? sh : 0)) : (cf << (sh < 0 ? -sh : 0));
                        off[e] = (x & kmaskR) | lane_off;
                    }
                }
                // (MODE_K8: off[e] is a rank; its 16-bit counter is half (rank & 1) of word rank / 2)
                auto bump = [&](const uint32_t o, const uint32_t inc) {
                    if constexpr (MODE == MODE_K8) atomicAdd(hist + (o >> 1), inc << ((o & 1u) << 4));
                    else atomicAdd(reinterpret_cast<uint32_t *>(hbytes + o), inc);
                };
                if (__all_sync(FULL, vw == 0xFFFFu)) {
#pragma unroll
                    for (int e = 0; e < 16; ++e) bump(off[e], 1u);
                } else if (__all_sync(FULL, silent || vw == 0xFFFFu)) {   // full step behind a look-back lane
                    const uint32_t inc = silent ? 0u : 1u;
#pragma unroll
                    for (int e = 0; e < 16; ++e) bump(off[e], inc);
                } else {
#pragma unroll
                    for (int e = 0; e < 16; ++e) bump(off[e], (vw >> (15 - e)) & 1u);   // branch-free: an invalid window adds 0
                }
            };
            for (uint32_t c0 = w0; c0 < w1; c0 += 64) {
                do_step(buf_a, c0);
                if (c0 + 32 < w1) do_step(buf_b, c0 + 32);
            }
        }
        // the next sequence's first bases travel during the write-out of this one (its plan is worked out again afterwards:
        // cheaper than eight more live registers across the write-out)
        const uint32_t slot_w = slot == 2u ? 0u : slot + 1u;
        if constexpr (LA) {
            const Plan pl_n = make_plan(s_nx[slot][1], s_nx[slot][2]);
            buf_a = fetch(pl_n, pl_n.w0 + lane);
            buf_b = fetch(pl_n, pl_n.w0 + 32 + lane);
            la_resolve(slot_w);   // the item after that: its offsets travel during the write-out, too
        }
        if (real) {
            // ---- total = block sum of `mine`
            mine = __reduce_add_sync(FULL, mine);
            if (lane == 0 && mine) atomicAdd(&s_total[it & 1], mine);
            K7Sched ks;
            if constexpr (MODE == MODE_K7) ks.template load_nw<NW>(p.sched, warp, lane);
            // the row image is about to be overwritten: the bulk copy of the previous row must have read it
            if constexpr (MODE != MODE_K8) { if (tid == 0) bulk_wait_read(); }
            __syncthreads();
            const uint32_t total = s_total[it & 1];
            if (tid == 0) s_total[(it + 1) & 1] = 0;
            ++it;
            const uint64_t dv = norm_divisor(total, p.norm_mode, p.canonical);
            if (tid == 0 && p.totals) p.totals[seq] = total;
            if (dv < (1ULL << 23)) long_write_row<OUT, NORM, MODE, NW, true, RS>(p, hbytes, stage, dv, ks, seq);
            else long_write_row<OUT, NORM, MODE, NW, false, RS>(p, hbytes, stage, dv, ks, seq);
            if constexpr (MODE != MODE_K8) fence_async_smem();
        }
        if constexpr (!LA) { la_ticket(); la_resolve(slot); }   // (slot 1 <-> 2: not the one the current item was read from)
        la_arrived();
        __syncthreads();
        if constexpr (MODE != MODE_K8) { if (real && tid == 0) bulk_store(out + seq * (uint64_t)p.dim, stage, p.dim * 4u); }
        seq = s_nx[slot][0];
        {
            const unsigned long long s0 = s_nx[slot][1], s1 = s_nx[slot][2];
            head_mask = 0xFFFFu >> (uint32_t)(s0 & 15);
            tail_mask = ~(0xFFFFu >> ((uint32_t)((s1 - 1) & 15) + 1u)) & 0xFFFFu;
            pl = make_plan(s0, s1);
        }
        slot = slot_w;
        if constexpr (!LA) {
            buf_a = fetch(pl, pl.w0 + lane);
            buf_b = fetch(pl, pl.w0 + 32 + lane);
        }
    }
    if constexpr (MODE != MODE_K8) { if (tid == 0) bulk_wait_all(); }
}

}  // namespace ktb
