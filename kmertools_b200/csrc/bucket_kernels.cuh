// bucket_kernels.cuh — rows larger than shared memory (canonical k = 9, 10; raw k = 8..10): bucket, then count.
//
// north_star asks for "a sort-and-reduce fallback when 4^k/2 exceeds shared memory".  Round 1 counted these rows with
// global RED atomics in L2-sized waves (wave_kernel, kernels.cuh): 27 % of the HBM roofline on config 5ii, bounded by
// the RED issue rate of an SM (0.66 lanes/clock), the zeroing of the rows and ~200 grid barriers.  This file replaces
// it with a one-level radix partition on the high bits of the column index followed by shared-memory counting:
//
//   tile_prefix_kernel   tiles (15,872 bases) per sequence -> exclusive prefix, so a work item maps to (sequence, tile)
//   bucket_kernel        one CTA per tile: decode, k-mer windows, column index (canonical rank from shared-memory
//                        bitmap tables, or the forward code in raw mode), then a counting sort of the tile's indices
//                        by SEGMENT (index >> log2 S) in shared memory; the sorted tile (16-bit in-segment indices,
//                        runs padded to 16 bytes) leaves as ONE bulk asynchronous copy (UBLKCP) into a pool, and a
//                        (pool offset, count) descriptor per (tile, segment) is recorded.
//   count_kernel         one CTA per (sequence, segment): the runs of that segment from every tile of the sequence
//                        are counted with shared-memory atomics into an S-bin u32 histogram, normalisation is applied
//                        in place, and the finished S x 4 bytes of the row leave as ONE bulk copy (double-buffered:
//                        the copy engine drains segment i while the CTA zeroes and counts segment i+1).
//
// Every output byte is written exactly once, by the copy engine, from shared memory; there are no global atomics on
// rows, no zeroing of rows and no grid barriers.  The pool costs 2 bytes per k-mer, written and read once (0.8 GB on
// config 5ii next to the 4.2 GB of rows).  Same arithmetic as everything else (kmer/src/kmer.rs:80-106,
// composition/src/oligo.rs:231-259).
#pragma once
#include "long_kernel.cuh"

namespace ktb {

constexpr int BK_WARPS = 16;                 // bucket_kernel: 512 threads
constexpr int BK_STEPS_PER_WARP = 2;
constexpr int BK_CHUNKS_PER_STEP = 31;       // lane 0 of a step only provides the look-back chunk
constexpr int BK_TILE_CHUNKS = BK_WARPS * BK_STEPS_PER_WARP * BK_CHUNKS_PER_STEP;   // 992 chunks = 15,872 bases
constexpr int BK_STAGE_ENTRIES = 16384;      // 32 KB of u16: 15,872 indices + 16-byte padding of <= 64 runs
constexpr int BK_MAX_SEG = 64;
constexpr int CK_THREADS = 512;              // count_kernel

struct BucketParams {
    const uint8_t *bases;
    const uint64_t *offsets;
    uint64_t n;
    uint64_t total_bases;
    const uint32_t *tile_prefix;     // [n+1] exclusive prefix of tiles per sequence
    unsigned long long *counter;     // work counter (zeroed)
    unsigned long long *pool_top;    // next free pool entry (zeroed); always a multiple of 8
    uint16_t *pool;                  // sorted in-segment indices
    uint2 *runs;                     // [tile * nseg + seg] = (pool offset in entries / 8, count)
    unsigned long long *totals;      // [n] valid windows per sequence (zeroed)
    const uint32_t *rank_full;       // RANK 1
    const uint32_t *rank_tab;        // RANK 2: canonical-code bitmap + u32 prefix per pair of words (as wave_kernel)
    uint32_t tab_words;
    uint32_t k;
    uint32_t nseg;
    uint32_t log2_seg;               // S = 1 << log2_seg columns per segment
};

// number of 31-chunk steps / tiles of a sequence [a, b)
__device__ __forceinline__ uint32_t bk_chunks(uint64_t a, uint64_t b, uint32_t k) {
    return (b - a >= k) ? (uint32_t)(((b - 1) >> 4) - (a >> 4)) + 1u : 0u;
}

// exclusive prefix of tiles per sequence; one CTA of 1024 threads, contiguous chunk of sequences per thread
__global__ void __launch_bounds__(1024) tile_prefix_kernel(const uint64_t *offsets, uint64_t n, uint32_t k, uint32_t *tile_prefix) {
    __shared__ uint32_t s_part[1024];
    const uint32_t tid = threadIdx.x;
    const uint64_t per = (n + 1023) / 1024;
    const uint64_t lo = min(n, tid * per), hi = min(n, lo + per);
    uint32_t sum = 0;
    for (uint64_t i = lo; i < hi; ++i) {
        const uint32_t nch = bk_chunks(offsets[i], offsets[i + 1], k);
        sum += (nch + BK_TILE_CHUNKS - 1) / BK_TILE_CHUNKS;
    }
    s_part[tid] = sum;
    __syncthreads();
    for (uint32_t d = 1; d < 1024; d <<= 1) {   // Hillis-Steele inclusive scan
        const uint32_t v = (tid >= d) ? s_part[tid - d] : 0u;
        __syncthreads();
        s_part[tid] += v;
        __syncthreads();
    }
    uint32_t run = s_part[tid] - sum;
    for (uint64_t i = lo; i < hi; ++i) {
        tile_prefix[i] = run;
        const uint32_t nch = bk_chunks(offsets[i], offsets[i + 1], k);
        run += (nch + BK_TILE_CHUNKS - 1) / BK_TILE_CHUNKS;
    }
    if (tid == 1023) tile_prefix[n] = s_part[1023];
}

template <int RANK>
__global__ void __launch_bounds__(BK_WARPS * 32, 1) bucket_kernel(const BucketParams p) {
    extern __shared__ __align__(128) uint32_t bsm[];
    // layout: staging (32 KB) | rank tables (RANK 2)
    uint16_t *stage = reinterpret_cast<uint16_t *>(bsm);
    const uint32_t *s_tab = bsm + BK_STAGE_ENTRIES / 2;
    const uint32_t *s_prefix = s_tab + p.tab_words;
    __shared__ uint32_t s_cnt[BK_MAX_SEG];      // indices of this tile per segment
    __shared__ uint32_t s_base[BK_MAX_SEG];     // first staging entry of the segment's run (multiple of 8)
    __shared__ unsigned long long s_work[2];    // tile id, pool offset of the tile
    __shared__ uint32_t s_seq[2];               // sequence, tile within the sequence
    __shared__ uint32_t s_tot;
    __shared__ uint32_t s_copy;                 // entries of the sorted tile (runs padded to multiples of 8)

    const int tid = threadIdx.x, lane = tid & 31;
    const uint32_t warp = tid >> 5;
    constexpr uint32_t FULL = 0xffffffffu;
    if constexpr (RANK == 2) {
        uint32_t *tab = bsm + BK_STAGE_ENTRIES / 2;
        const uint32_t nw = p.tab_words + p.tab_words / 2;
        for (uint32_t i = tid; i < nw; i += blockDim.x) tab[i] = __ldg(p.rank_tab + i);
    }
    (void)s_prefix;
    const uint32_t k = p.k;
    const uint32_t kmask = (1u << (2 * k)) - 1u;
    const uint32_t seg_mask = (1u << p.log2_seg) - 1u;
    const uint64_t ntiles = p.tile_prefix[p.n];
    const uint4 filler = make_uint4(0x41414141u, 0x41414141u, 0x41414141u, 0x41414141u);

    for (;;) {
        __syncthreads();   // previous tile done with s_* (and the tables are loaded)
        if (tid == 0) {
            const unsigned long long t = atomicAdd(p.counter, 1ULL);
            s_work[0] = t;
            if (t < ntiles) {   // sequence of tile t: largest s with tile_prefix[s] <= t
                uint64_t lo = 0, hi = p.n;
                while (hi - lo > 1) {
                    const uint64_t mid = (lo + hi) >> 1;
                    if (p.tile_prefix[mid] <= t) lo = mid; else hi = mid;
                }
                s_seq[0] = (uint32_t)lo;
                s_seq[1] = (uint32_t)(t - p.tile_prefix[lo]);
            }
            s_tot = 0;
        }
        if (tid < BK_MAX_SEG) s_cnt[tid] = 0;
        __syncthreads();
        const unsigned long long tile_id = s_work[0];
        if (tile_id >= ntiles) break;
        const uint64_t seq = s_seq[0];
        const uint32_t tile = s_seq[1];
        const uint64_t q0 = p.offsets[seq], q1 = p.offsets[seq + 1];
        const uint64_t cbase = q0 >> 4;
        const uint32_t nch = (uint32_t)(((q1 - 1) >> 4) - cbase) + 1u;
        const uint32_t head_mask = 0xFFFFu >> (uint32_t)(q0 & 15);
        const uint32_t tail_mask = ~(0xFFFFu >> ((uint32_t)((q1 - 1) & 15) + 1u)) & 0xFFFFu;

        // ---- phase 1: column index of every window of the tile, position inside the segment's run
        uint32_t idx[BK_STEPS_PER_WARP][16];
        uint32_t pos[BK_STEPS_PER_WARP][8];    // two 16-bit positions per word
        uint32_t vws[BK_STEPS_PER_WARP];
        uint32_t mine = 0;
#pragma unroll
        for (int s = 0; s < BK_STEPS_PER_WARP; ++s) {
            // lane l > 0 owns chunk c; lane 0 holds the chunk before lane 1's (look-back only)
            const int64_t c = (int64_t)tile * BK_TILE_CHUNKS + (int64_t)(warp * BK_STEPS_PER_WARP + s) * BK_CHUNKS_PER_STEP + lane - 1;
            const bool inside = c >= 0 && c < (int64_t)nch;
            const uint4 v = inside ? load16_guarded(p.bases, (cbase + (uint64_t)c) << 4, p.total_bases) : filler;
            uint32_t cf, vm;
            decode16(v, cf, vm);
            if (!inside) vm = 0;
            if (c == 0) vm &= head_mask;
            if (c == (int64_t)nch - 1) vm &= tail_mask;
            const uint32_t cf_prev = __shfl_up_sync(FULL, cf, 1);
            const uint32_t vm_prev = __shfl_up_sync(FULL, vm, 1);
            uint32_t vw = window_mask((vm_prev << 16) | vm, k) & 0xFFFFu;
            if (lane == 0) vw = 0;
            vws[s] = vw;
            mine += __popc(vw);
            const uint64_t F64 = ((uint64_t)cf_prev << 32) | cf;
            uint64_t R64 = 0;
            if constexpr (RANK == 2) R64 = ((uint64_t)revcomp_pack(cf) << 32) | revcomp_pack(cf_prev);
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const uint32_t f = (uint32_t)(F64 >> (2 * (15 - j))) & kmask;
                if constexpr (RANK == 0) {
                    idx[s][j] = f;
                } else if constexpr (RANK == 1) {
                    idx[s][j] = __ldg(p.rank_full + f);
                } else {
                    const uint32_t r = (uint32_t)(R64 >> (2 * (17 + j - (int)k))) & kmask;
                    const uint32_t cc = min(f, r);
                    const uint32_t wd = cc >> 5;
                    const uint2 bw = reinterpret_cast<const uint2 *>(s_tab)[wd >> 1];
                    const bool odd = (wd & 1u) != 0u;
                    const uint32_t below = (odd ? bw.y : bw.x) & ((1u << (cc & 31u)) - 1u);
                    idx[s][j] = s_prefix[wd >> 1] + (uint32_t)__popc(below) + (odd ? (uint32_t)__popc(bw.x) : 0u);
                }
            }
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                uint32_t ps = 0;
                if (vw & (1u << (15 - j))) ps = atomicAdd(&s_cnt[idx[s][j] >> p.log2_seg], 1u);
                if (j & 1) pos[s][j >> 1] |= ps << 16; else pos[s][j >> 1] = ps;
            }
        }
#pragma unroll
        for (int sft = 16; sft > 0; sft >>= 1) mine += __shfl_xor_sync(FULL, mine, sft);
        if (lane == 0 && mine) atomicAdd(&s_tot, mine);
        // the staging buffer is about to be overwritten: the previous tile's bulk copy must have read it
        if (tid == 0) bulk_wait_read();
        __syncthreads();

        // ---- phase 2: run bases (16-byte aligned), pool reservation, descriptors
        if (warp == 0) {
            const uint32_t c0 = (2 * lane < (int)p.nseg) ? s_cnt[2 * lane] : 0u;
            const uint32_t c1 = (2 * lane + 1 < (int)p.nseg) ? s_cnt[2 * lane + 1] : 0u;
            const uint32_t a0 = (c0 + 7u) & ~7u, a1 = (c1 + 7u) & ~7u;
            uint32_t incl = a0 + a1;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t t = __shfl_up_sync(FULL, incl, d);
                if (lane >= d) incl += t;
            }
            const uint32_t total_aligned = __shfl_sync(FULL, incl, 31);
            unsigned long long off = 0;
            if (lane == 0) {
                off = total_aligned ? atomicAdd(p.pool_top, (unsigned long long)total_aligned) : 0ULL;
                s_work[1] = off;
                if (s_tot) atomicAdd(p.totals + seq, (unsigned long long)s_tot);
            }
            off = __shfl_sync(FULL, off, 0);
            const uint32_t b0 = incl - a0 - a1, b1 = b0 + a0;
            uint2 *rd = p.runs + tile_id * p.nseg;
            if (2 * lane < (int)p.nseg) { s_base[2 * lane] = b0; rd[2 * lane] = make_uint2((uint32_t)((off + b0) >> 3), c0); }
            if (2 * lane + 1 < (int)p.nseg) { s_base[2 * lane + 1] = b1; rd[2 * lane + 1] = make_uint2((uint32_t)((off + b1) >> 3), c1); }
            if (lane == 31) s_copy = total_aligned;
        }
        __syncthreads();

        // ---- phase 3: scatter the in-segment indices to their runs, one bulk copy of the sorted tile to the pool
#pragma unroll
        for (int s = 0; s < BK_STEPS_PER_WARP; ++s) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                if (vws[s] & (1u << (15 - j))) {
                    const uint32_t sg = idx[s][j] >> p.log2_seg;
                    const uint32_t ps = (j & 1) ? (pos[s][j >> 1] >> 16) : (pos[s][j >> 1] & 0xFFFFu);
                    stage[s_base[sg] + ps] = (uint16_t)(idx[s][j] & seg_mask);
                }
            }
        }
        fence_async_smem();
        __syncthreads();
        if (tid == 0 && s_copy) bulk_store(p.pool + s_work[1], stage, s_copy * 2u);
    }
    if (tid == 0) bulk_wait_all();
}

struct CountParams {
    const uint32_t *tile_prefix;   // [n+1]
    const uint16_t *pool;
    const uint2 *runs;
    const unsigned long long *totals_in;   // [n] from bucket_kernel
    uint64_t *totals_out;          // optional
    void *out;
    unsigned long long *counter;   // work counter (zeroed)
    uint64_t n;
    uint64_t dim;
    uint32_t nseg;
    uint32_t log2_seg;
    int norm_mode;
    int canonical;
};

// One CTA per (sequence, segment).  u32 / f32 rows leave shared memory as one bulk copy per segment; f64 rows are
// stored directly (two values per 128-bit store).
template <int OUT, bool NORM>
__global__ void __launch_bounds__(CK_THREADS, 1) count_kernel(const CountParams p) {
    extern __shared__ __align__(128) uint32_t csm[];
    __shared__ unsigned long long s_item;
    using T = typename OutT<OUT>::type;
    const uint32_t S = 1u << p.log2_seg;
    const int tid = threadIdx.x;
    const uint64_t nitems = p.n * p.nseg;
    uint32_t it = 0;
    for (;;) {
        if (tid == 0) s_item = atomicAdd(p.counter, 1ULL);
        __syncthreads();
        const unsigned long long item = s_item;
        __syncthreads();
        if (item >= nitems) break;
        const uint64_t seq = item / p.nseg;
        const uint32_t seg = (uint32_t)(item - seq * p.nseg);
        uint32_t *hist = csm + (size_t)(OUT == OUT_F64 ? 0 : (it & 1)) * S;
        ++it;
        const uint32_t cols = (uint32_t)min((uint64_t)S, p.dim - (uint64_t)seg * S);   // multiple of 4 on this path
        // the bulk copy that used this buffer two items ago must have read it (at most one newer copy in flight)
        if constexpr (OUT != OUT_F64) {
            if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            __syncthreads();
        }
        for (uint32_t i = tid * 4u; i < cols; i += CK_THREADS * 4u) *reinterpret_cast<uint4 *>(hist + i) = make_uint4(0, 0, 0, 0);
        __syncthreads();
        const uint32_t t0 = p.tile_prefix[seq], t1 = p.tile_prefix[seq + 1];
        for (uint32_t t = t0; t < t1; ++t) {
            const uint2 rd = __ldg(p.runs + (uint64_t)t * p.nseg + seg);
            const uint4 *src = reinterpret_cast<const uint4 *>(p.pool + ((uint64_t)rd.x << 3));
            const uint32_t cnt = rd.y;
            for (uint32_t i = tid; i * 8u < cnt; i += CK_THREADS) {
                const uint4 v = __ldg(src + i);
                const uint32_t w[4] = {v.x, v.y, v.z, v.w};
                const uint32_t left = cnt - i * 8u;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    if (2u * q < left) atomicAdd(hist + (w[q] & 0xFFFFu), 1u);
                    if (2u * q + 1u < left) atomicAdd(hist + (w[q] >> 16), 1u);
                }
            }
        }
        __syncthreads();
        const unsigned long long total = p.totals_in[seq];
        if (tid == 0 && seg == 0 && p.totals_out) p.totals_out[seq] = total;
        const uint64_t dv = norm_divisor(total, p.norm_mode, p.canonical);
        const float dF = (float)dv, rinv = __frcp_rn(dF);
        const double dD = (double)dv;
        const bool small = dv < (1ULL << 23);
        T *row = reinterpret_cast<T *>(p.out) + seq * p.dim + (uint64_t)seg * S;
        if constexpr (OUT == OUT_F64) {
            for (uint32_t i = tid * 2u; i < cols; i += CK_THREADS * 2u) {
                const uint2 c = *reinterpret_cast<const uint2 *>(hist + i);
                *reinterpret_cast<double2 *>(row + i) = make_double2(cvt_count<OUT_F64, NORM, false>(c.x, dF, rinv, dD),
                                                                     cvt_count<OUT_F64, NORM, false>(c.y, dF, rinv, dD));
            }
        } else {
            if constexpr (OUT == OUT_F32) {   // counts -> floats in place
                for (uint32_t i = tid * 4u; i < cols; i += CK_THREADS * 4u) {
                    const uint4 c = *reinterpret_cast<const uint4 *>(hist + i);
                    float4 o;
                    if (small) {
                        o.x = cvt_count<OUT_F32, NORM, true>(c.x, dF, rinv, dD); o.y = cvt_count<OUT_F32, NORM, true>(c.y, dF, rinv, dD);
                        o.z = cvt_count<OUT_F32, NORM, true>(c.z, dF, rinv, dD); o.w = cvt_count<OUT_F32, NORM, true>(c.w, dF, rinv, dD);
                    } else {
                        o.x = cvt_count<OUT_F32, NORM, false>(c.x, dF, rinv, dD); o.y = cvt_count<OUT_F32, NORM, false>(c.y, dF, rinv, dD);
                        o.z = cvt_count<OUT_F32, NORM, false>(c.z, dF, rinv, dD); o.w = cvt_count<OUT_F32, NORM, false>(c.w, dF, rinv, dD);
                    }
                    *reinterpret_cast<float4 *>(hist + i) = o;
                }
            }
            fence_async_smem();
            __syncthreads();
            if (tid == 0) bulk_store(row, hist, cols * 4u);
        }
    }
    if (tid == 0) bulk_wait_all();
}

}  // namespace ktb
