// bucket_kernels.cuh — rows larger than shared memory (canonical k = 9, 10; raw k = 8..10): bucket, then count.
//
// north_star asks for "a sort-and-reduce fallback when 4^k/2 exceeds shared memory".  Round 1 counted these rows with
// global RED atomics in L2-sized waves (wave_kernel, kernels.cuh): 27 % of the HBM roofline on config 5ii, bounded by
// the RED issue rate of an SM (0.66 lanes/clock), the zeroing of the rows and ~200 grid barriers.  A histogram
// distributed over the shared memory of a cluster was measured and ruled out: remote shared-memory atomics reach
// 0.18 per clock per SM (tools/microbench_dsmem.cu, profiles/r2_microbench_dsmem.txt).  This file is a one-level
// radix partition on the high bits of the k-mer CODE followed by shared-memory counting:
//
//   tile_prefix_kernel   tiles (7,936 bases) per sequence -> exclusive prefix, so a work item maps to (sequence, tile)
//   bucket_kernel        one CTA per tile, table-free: decode, windows, canonical code min(f, r) (or f in raw mode),
//                        then a counting sort of the tile's codes by SEGMENT (code >> 14) in shared memory.  The sorted
//                        tile (16-bit in-segment codes, runs padded to 16 bytes) leaves as ONE bulk asynchronous copy
//                        (UBLKCP) into the tile's slot of a pool; a (start, count) descriptor is kept per (tile,
//                        segment).  No rank tables here, so three CTAs share an SM and hide each other's barriers.
//   count_kernel         one CTA per (sequence, segment): loads the segment's 3 KB slice of the canonical-code bitmap
//                        (+ running ranks), turns every in-segment code of the sequence's runs into a column with one
//                        shared-memory look-up + popcount, counts with shared-memory atomics, applies the
//                        normalisation in place and writes ITS part of the row — the columns of a code segment are
//                        contiguous because the rank is monotone in the code — as one bulk copy.  The kernel is bound
//                        by those row writes, so the per-k-mer rank arithmetic hides under them.
//
// Every output byte is written exactly once, from shared memory; there are no global atomics on rows, no zeroing of
// rows and no grid barriers.  The pool costs 2 bytes per k-mer, written and read once (0.8 GB on config 5ii next to
// the 4.2 GB of rows).  Same arithmetic as everything else (kmer/src/kmer.rs:80-106, composition/src/oligo.rs:231-259).
#pragma once
#include "long_kernel.cuh"

namespace ktb {

constexpr int BK_WARPS = 8;                  // bucket_kernel: 256 threads, one 31-chunk step per warp
constexpr int BK_CHUNKS_PER_STEP = 31;       // lane 0 of a step only provides the look-back chunk
constexpr int BK_TILE_CHUNKS = BK_WARPS * BK_CHUNKS_PER_STEP;       // 248 chunks = 3,968 bases (positions fit 12 bits)
constexpr int BK_MAX_SEG = 128;
constexpr int BK_SEG_PER_LANE = BK_MAX_SEG / 32;   // phase 2: one warp scans the segment counters
constexpr int BK_TILE_CAP = BK_TILE_CHUNKS * 16 + BK_MAX_SEG * 8;   // pool entries per tile (runs padded to 8 entries)
constexpr int CK_THREADS = 256;              // count_kernel
constexpr int CK_WARPS = CK_THREADS / 32;

// everything bucket_kernel needs to know about a tile, written once by tile_prefix_kernel (32 bytes = two 128-bit
// loads that do not depend on each other, prefetched one tile ahead)
struct __align__(16) TileInfo {
    uint64_t q0, q1;        // sequence = bases[q0, q1)
    uint32_t seq, tile;     // sequence index, tile index inside the sequence
    uint32_t pad[2];
};

struct BucketParams {
    const uint8_t *bases;
    uint64_t n;
    uint64_t total_bases;
    const uint32_t *tile_prefix;     // [n+1] exclusive prefix of tiles per sequence
    const TileInfo *tiles;           // [tile_prefix[n]]
    uint16_t *pool;                  // tile t owns entries [t * BK_TILE_CAP, (t+1) * BK_TILE_CAP)
    uint32_t *runs;                  // [seg * ntiles + tile] = (first entry of the run / 8) << 16 | count
    unsigned long long *totals;      // [n] valid windows per sequence (zeroed)
    uint32_t k;
    uint32_t nseg;
    uint32_t log2_seg;               // codes per segment = 1 << log2_seg
    uint64_t seq_lo, seq_hi;         // this launch buckets the tiles of sequences [seq_lo, seq_hi) (one wave)
};

// chunks (16 aligned bytes) touched by sequence [a, b); 0 when it is shorter than k
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

// one thread per tile: its sequence is the largest s with tile_prefix[s] <= tile
__global__ void __launch_bounds__(256) tile_info_kernel(const uint64_t *offsets, uint64_t n, const uint32_t *tile_prefix, TileInfo *tiles) {
    const uint32_t ntiles = tile_prefix[n];
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < ntiles; t += gridDim.x * blockDim.x) {
        uint64_t lo = 0, hi = n;
        while (hi - lo > 1) {
            const uint64_t mid = (lo + hi) >> 1;
            if (tile_prefix[mid] <= t) lo = mid; else hi = mid;
        }
        TileInfo ti;
        ti.q0 = offsets[lo]; ti.q1 = offsets[lo + 1]; ti.seq = (uint32_t)lo; ti.tile = t - tile_prefix[lo];
        ti.pad[0] = ti.pad[1] = 0;
        tiles[t] = ti;
    }
}

// KT, LS: compile-time k and log2 of the segment size (0 = from the parameters)
template <bool CANON, int KT = 0, int LS = 0>
__global__ void __launch_bounds__(BK_WARPS * 32, 4) bucket_kernel(const BucketParams p) {
    __shared__ __align__(128) uint16_t stage[BK_TILE_CAP + BK_WARPS * 32];   // sorted tile; behind it one entry per thread that swallows its invalid windows
    __shared__ uint32_t s_cnt[BK_MAX_SEG + 1];  // codes of this tile per segment; [nseg] collects the invalid windows
    __shared__ uint32_t s_base[BK_MAX_SEG + 1]; // first staging entry of the segment's run (multiple of 8); [nseg] = trash
    __shared__ uint32_t s_tot;
    __shared__ uint32_t s_copy;                 // entries of the sorted tile (runs padded to multiples of 8)

    const int tid = threadIdx.x, lane = tid & 31;
    const uint32_t warp = tid >> 5;
    constexpr uint32_t FULL = 0xffffffffu;
    const uint32_t k = KT ? (uint32_t)KT : p.k;
    const uint32_t log2_seg = LS ? (uint32_t)LS : p.log2_seg;
    const uint32_t kmask = (1u << (2 * k)) - 1u;
    const uint32_t seg_mask = (1u << log2_seg) - 1u;
    const uint32_t trash = p.nseg;
    const uint64_t ntiles = p.tile_prefix[p.n];
    const uint4 filler = make_uint4(0x41414141u, 0x41414141u, 0x41414141u, 0x41414141u);

    // chunk of this lane in a tile: lane l > 0 owns chunk c, lane 0 holds the chunk before lane 1's (look-back only)
    auto chunk_of = [&](const TileInfo &ti) -> int64_t {
        return (int64_t)ti.tile * BK_TILE_CHUNKS + (int64_t)warp * BK_CHUNKS_PER_STEP + lane - 1;
    };
    auto load_chunk = [&](const TileInfo &ti) -> uint4 {
        const int64_t c = chunk_of(ti);
        const uint64_t cbase = ti.q0 >> 4;
        const int64_t nch = (int64_t)(((ti.q1 - 1) >> 4) - cbase) + 1;
        return (c >= 0 && c < nch) ? load16_guarded(p.bases, (cbase + (uint64_t)c) << 4, p.total_bases) : filler;
    };

    // tiles cost the same, so they are dealt out statically (tile b, b + grid, ...): the next tile's record and bases
    // are loaded while the current tile is processed; no work counter, no load on the critical path
    const uint64_t tile_end = p.tile_prefix[p.seq_hi];
    uint64_t tile_id = (uint64_t)p.tile_prefix[p.seq_lo] + blockIdx.x;
    if (tile_id >= tile_end) return;
    TileInfo ti = p.tiles[tile_id];
    uint4 v = load_chunk(ti);
    for (; tile_id < tile_end; tile_id += gridDim.x) {
        const bool more = tile_id + gridDim.x < tile_end;
        TileInfo nti = ti;
        if (more) nti = p.tiles[tile_id + gridDim.x];
        if (tid == 0) s_tot = 0;
        if (tid <= BK_MAX_SEG) s_cnt[tid] = 0;
        __syncthreads();
        const uint64_t q0 = ti.q0, q1 = ti.q1;
        const int64_t nch = (int64_t)(((q1 - 1) >> 4) - (q0 >> 4)) + 1;
        const int64_t c = chunk_of(ti);

        // ---- phase 1: code of every window of the tile, position inside its segment's run (atomics WITH return)
        uint32_t cf, vm;
        decode16(v, cf, vm);
        if (c < 0 || c >= nch) vm = 0;
        if (c == 0) vm &= 0xFFFFu >> (uint32_t)(q0 & 15);
        if (c == nch - 1) vm &= ~(0xFFFFu >> ((uint32_t)((q1 - 1) & 15) + 1u)) & 0xFFFFu;
        const uint32_t cf_prev = __shfl_up_sync(FULL, cf, 1);
        const uint32_t vm_prev = __shfl_up_sync(FULL, vm, 1);
        uint32_t vw = window_mask((vm_prev << 16) | vm, k) & 0xFFFFu;
        if (lane == 0) vw = 0;
        uint32_t mine = __popc(vw);
        // reverse strand, shifted once so that window j sits at bit 2j: base i of R64 is at bit 2(i + 16)
        uint32_t rlo = 0, rhi = 0;
        if constexpr (CANON) {
            const uint32_t rcf = revcomp_pack(cf);
            const uint32_t rcf_prev = __shfl_up_sync(FULL, rcf, 1);   // (lane 0 emits nothing)
            const uint64_t R64 = (((uint64_t)rcf << 32) | rcf_prev) >> (2 * (17 - (int)k));
            rlo = (uint32_t)R64; rhi = (uint32_t)(R64 >> 32);
        }
        uint32_t cp[16];   // per window: code (20 bits, k <= 10) | position in the run << 20 (12 bits)
        // steady state: every lane but lane 0 has 16 valid windows, so validity is ONE predicate per lane instead of a
        // bit test per window (warp-uniform choice; invalid windows queue up in the trash segment, no branch per atomic)
        const bool steady = __all_sync(FULL, lane == 0 || vw == 0xFFFFu);
        const bool lane_ok = lane != 0;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            uint32_t code = ((j < 15) ? __funnelshift_r(cf, cf_prev, 2 * (15 - j)) : cf) & kmask;
            if constexpr (CANON) code = min(code, ((j > 0) ? __funnelshift_r(rlo, rhi, 2 * j) : rlo) & kmask);
            cp[j] = code;
        }
        if (steady) {
            if (lane_ok) {   // (lane 0 sits the step out: no select per window)
#pragma unroll
                for (int j = 0; j < 16; ++j)
                    cp[j] += atomicAdd(&s_cnt[cp[j] >> log2_seg], 1u) << 20;   // (+: one shift-add; the fields do not overlap)
            }
        } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const uint32_t sg = (vw & (1u << (15 - j))) ? (cp[j] >> log2_seg) : trash;
                cp[j] += atomicAdd(&s_cnt[sg], 1u) << 20;   // (+: one shift-add; the fields do not overlap)
            }
        }
        if (more) v = load_chunk(nti);   // the next tile's bases are on their way while this tile is sorted
        mine = __reduce_add_sync(FULL, mine);
        if (lane == 0 && mine) atomicAdd(&s_tot, mine);
        // the staging buffer is about to be overwritten: the previous tile's bulk copy must have read it
        if (tid == 0) bulk_wait_read();
        __syncthreads();

        // ---- phase 2: run bases (16-byte aligned) and descriptors
        if (warp == 0) {
            uint32_t cn[BK_SEG_PER_LANE], sum = 0;
#pragma unroll
            for (int u = 0; u < BK_SEG_PER_LANE; ++u) {
                const uint32_t sg = BK_SEG_PER_LANE * lane + u;
                cn[u] = sg < p.nseg ? s_cnt[sg] : 0u;
                sum += (cn[u] + 7u) & ~7u;
            }
            uint32_t incl = sum;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t t = __shfl_up_sync(FULL, incl, d);
                if (lane >= d) incl += t;
            }
            uint32_t b = incl - sum;
            uint32_t *rd = p.runs + tile_id;   // descriptors of one segment are contiguous over the tiles (count_kernel's order)
#pragma unroll
            for (int u = 0; u < BK_SEG_PER_LANE; ++u) {
                const uint32_t sg = BK_SEG_PER_LANE * lane + u;
                if (sg < p.nseg) { s_base[sg] = b; rd[(uint64_t)sg * ntiles] = ((b >> 3) << 16) | cn[u]; }
                b += (cn[u] + 7u) & ~7u;
            }
            if (lane == 31) s_copy = incl;
            if (lane == 0 && s_tot) atomicAdd(p.totals + ti.seq, (unsigned long long)s_tot);
        }
        __syncthreads();

        // ---- phase 3: scatter the in-segment codes to their runs (branch-free: invalid windows land in the trash
        // entries), then one bulk copy of the sorted tile into its pool slot
        // (an invalid window reads the base of whatever segment its stale code names — harmless — and is redirected)
        const uint32_t trash_slot = BK_TILE_CAP + tid;   // (per thread: no two threads ever store to the same entry)
        if (steady) {
            if (lane_ok) {
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const uint32_t code = cp[j] & 0xFFFFFu;
                    stage[s_base[code >> log2_seg] + (cp[j] >> 20)] = (uint16_t)(code & seg_mask);
                }
            }
        } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const uint32_t code = cp[j] & 0xFFFFFu;
                const uint32_t slot = s_base[code >> log2_seg] + (cp[j] >> 20);
                stage[(vw & (1u << (15 - j))) ? slot : trash_slot] = (uint16_t)(code & seg_mask);
            }
        }
        fence_async_smem();
        __syncthreads();
        if (tid == 0 && s_copy) bulk_store(p.pool + tile_id * BK_TILE_CAP, stage, s_copy * 2u);
        ti = nti;
    }
    if (tid == 0) bulk_wait_all();
}

struct CountParams {
    const uint32_t *tile_prefix;   // [n+1]
    const uint16_t *pool;
    const uint32_t *runs;          // [seg * ntiles + tile]
    const unsigned long long *totals_in;   // [n] from bucket_kernel
    uint64_t *totals_out;          // optional
    void *out;
    unsigned long long *counter;   // work counter (zeroed)
    const uint32_t *rank_tab;      // canonical: bitmap of the canonical codes (tab_words) + u32 running rank per PAIR of words
    uint32_t tab_words;            // 4^k / 32
    uint64_t n;
    uint64_t dim;
    uint32_t nseg;
    uint32_t log2_seg;
    uint32_t hist_words;           // words of histogram memory: S, or 2 S (every segment alternates between two buffers)
    int norm_mode;
    int canonical;
#ifdef KTB_COUNT_PROBE
    uint32_t probe;                // 1 no atomics, 2 no zeroing, 4 no bulk copy, 8 no code loads
#endif
    uint64_t chunk_lo, chunk_hi;   // this launch serves the chunks (of CK_SEQ_CHUNK sequences) [chunk_lo, chunk_hi) (one wave)
};

// Diagnostic builds only (tools/gpu_probe_count.sh compiles a second library with -DKTB_COUNT_PROBE): phases of
// count_kernel can be switched off to attribute its time; the rows such a build writes are wrong by construction.
#ifdef KTB_COUNT_PROBE
#define PROBE(bit) ((p.probe & (bit)) != 0)
#else
#define PROBE(bit) false
#endif

#ifndef KTB_CK_SEQ_CHUNK
#define KTB_CK_SEQ_CHUNK 8
#endif
constexpr int CK_SEQ_CHUNK = KTB_CK_SEQ_CHUNK;   // consecutive sequences per work unit (multiple of 8)

// One CTA per (segment of the code space, chunk of CK_SEQ_CHUNK sequences).  The columns of the segment are
// [R, R + bins): R = rank of the segment's first code.  For every sequence of the chunk the runs of the segment (one
// per tile of the sequence, contiguous descriptors) are dealt to QUARTER-WARPS — a run of ~60 codes is eight 16-byte
// loads — counted with shared-memory atomics and written out as one bulk copy (u32 / f32; f64 parts are stored
// directly).  The first run of every quarter-warp is loaded BEFORE the CTA waits for the bulk copy of the previous part
// to leave the histogram, so that latency and the copy overlap.
template <int OUT, bool NORM, bool CANON>
__global__ void __launch_bounds__(CK_THREADS, 3) count_kernel(const CountParams p) {
    extern __shared__ __align__(128) uint32_t csm[];
    __shared__ unsigned long long s_unit;
    __shared__ uint32_t s_tp[CK_SEQ_CHUNK + 1];   // tile prefix of the chunk's sequences
    __shared__ unsigned long long s_tot[CK_SEQ_CHUNK];   // valid windows of the chunk's sequences
    __shared__ uint32_t s_trash[32];
    __shared__ uint32_t s_rd[CK_SEQ_CHUNK * (CK_THREADS / 8)];   // descriptor of every quarter-warp's first run, per sequence of the chunk
    static_assert(CK_SEQ_CHUNK % 8 == 0 && CK_THREADS == 256, "descriptor staging: 8 sequences x 32 quarter-warps per pass");
    using T = typename OutT<OUT>::type;
    const uint32_t S = 1u << p.log2_seg;                 // codes per segment
    const uint32_t wps = S / 32;                         // bitmap words per segment
    uint32_t *s_bits = csm + p.hist_words;               // CANON: bitmap of the segment's canonical codes
    uint32_t *s_pref = s_bits + wps;                     // CANON: columns before each word, relative to the segment's first
    const int tid = threadIdx.x;
    const uint32_t qw = tid >> 3, ql = tid & 7;          // quarter-warp (one per run), lane in it (16 bytes = 8 codes per load)
    constexpr uint32_t NQW = CK_THREADS / 8;
    const uint32_t ntiles = p.tile_prefix[p.n];
    const uint64_t nunits = (p.chunk_hi - p.chunk_lo) * p.nseg;
    // Kept in registers on purpose: re-reading a kernel parameter from the constant bank at the top of the sequence loop
    // (LDC) took the scoreboard of the code loads that are in flight across iterations and waited for them — 18 % of the
    // kernel's stall samples (profiles/r2_ncu_count_summary.txt) — which undid the prefetch.
    // (threadIdx.y is 0, which the assembler cannot know: the sums are values it has to keep, not parameters it can re-read)
    const uint16_t *pool = p.pool + threadIdx.y;
    const uint32_t hist_half = p.hist_words / 2 + threadIdx.y;
    // units differ in cost (segments hold between none and twice the average number of k-mers), so they are dealt out
    // dynamically; the counter is read one unit ahead to keep its round trip off the critical path
    unsigned long long next_unit = 0;
    if (tid == 0) next_unit = atomicAdd(p.counter, 1ULL);
    uint32_t cur_seg = 0xFFFFFFFFu;   // the segment whose tables are in shared memory
    uint64_t col0 = 0, col1 = 0;      // its columns
    for (;;) {
        __syncthreads();      // everyone is done with s_unit and the tables of the previous unit
        if (tid == 0) {
            s_unit = next_unit;
            next_unit = atomicAdd(p.counter, 1ULL);
        }
        __syncthreads();
        const unsigned long long unit = s_unit;
        if (unit >= nunits) break;
        // units are numbered chunk-major (consecutive units = the segments of one chunk of sequences), so CTAs that run side
        // by side work on segments of every size.  Measured against segment-major order (large segments first, tables kept
        // while a CTA stays in a segment; profiles/r2_sweeps.txt, batch r2v): 1.32 against 1.36 ms for u32 rows on config
        // 5ii, 1.50 against 1.49 ms for f32 rows.
        const uint64_t chunk = p.chunk_lo + unit / p.nseg;
        const uint32_t seg = (uint32_t)(unit % p.nseg);
        if constexpr (CANON) {
            if (seg != cur_seg) {
                const uint32_t *gpref = p.rank_tab + p.tab_words;
                col0 = gpref[(size_t)seg * (wps / 2)];
                col1 = (seg + 1 < p.nseg) ? (uint64_t)gpref[(size_t)(seg + 1) * (wps / 2)] : p.dim;
                for (uint32_t w = tid; w < wps; w += CK_THREADS) {
                    const uint32_t bits = __ldg(p.rank_tab + (size_t)seg * wps + w);
                    const uint32_t prev = (w & 1u) ? __ldg(p.rank_tab + (size_t)seg * wps + w - 1) : 0u;
                    s_bits[w] = bits;
                    s_pref[w] = __ldg(gpref + ((size_t)seg * wps + w) / 2) - (uint32_t)col0 + (uint32_t)__popc(prev);
                }
            }
        } else {
            col0 = (uint64_t)seg * S;
            col1 = min(p.dim, col0 + S);
        }
        cur_seg = seg;
        const uint32_t bins = (uint32_t)(col1 - col0);   // multiple of 4 for every k this path serves (checked on the host)
        if (bins == 0) continue;                         // uniform: a segment without canonical codes
        const uint32_t *runs = p.runs + (uint64_t)seg * ntiles;
        const uint64_t seq0 = chunk * CK_SEQ_CHUNK, seq_end = min(p.n, (chunk + 1) * CK_SEQ_CHUNK);
        // software pipeline over the sequences of the chunk: the descriptors of every quarter-warp's first run are
        // fetched for all sequences of the chunk at once, and the codes of sequence i+1 are requested at the top of
        // iteration i — a whole iteration (count, write-out, wait for the previous bulk copy, zeroing) ahead of their use
        // segments with at most S / 2 columns (the upper half of the code space: fewer and fewer codes are canonical)
        // alternate between the two halves of the histogram memory, so the CTA never waits for its own last bulk copy
        const bool two = bins * 2u <= p.hist_words;
        if (tid == 0) bulk_wait_read();   // buffers change roles between units
#pragma unroll
        for (uint32_t j0 = 0; j0 < (uint32_t)CK_SEQ_CHUNK; j0 += 8) {
            const uint32_t j = j0 + (tid >> 5), q = tid & 31;   // 8 sequences x NQW quarter-warps per pass
            const uint32_t a0 = p.tile_prefix[min(p.n, seq0 + j)], a1 = p.tile_prefix[min(p.n, seq0 + j + 1)];
            s_rd[j * NQW + q] = (a0 + q < a1) ? __ldg(runs + a0 + q) : 0u;
            if (q == 0) s_tp[j] = a0;
            if (q == 2 && j == (uint32_t)CK_SEQ_CHUNK - 1) s_tp[CK_SEQ_CHUNK] = a1;
            if (q == 1) s_tot[j] = p.totals_in[min(p.n - 1, seq0 + j)];   // (a load per sequence here would stall every thread)
        }
        __syncthreads();
        // this lane's first 2 x 16 bytes of the first run of the chunk's i-th sequence (a run of a 3,968-base tile has ~62
        // codes at 64 segments: lanes 0..7 of the quarter-warp cover 128 with two loads each, both requested a sequence ahead)
        const uint4 zero4 = make_uint4(0, 0, 0, 0);
        auto first_codes = [&](const uint32_t i, uint4 &a, uint4 &b) {
            const uint32_t rdi = s_rd[i * NQW + qw];
            const uint32_t c = rdi & 0xFFFFu;
            const uint4 *src = reinterpret_cast<const uint4 *>(pool + (uint64_t)(s_tp[i] + qw) * BK_TILE_CAP + ((uint64_t)(rdi >> 16) << 3));
            a = (8u * ql < c && !PROBE(8)) ? __ldg(src + ql) : zero4;
            b = (64u + 8u * ql < c && !PROBE(8)) ? __ldg(src + 8 + ql) : zero4;
        };
        // add `delta` to the bins of the first `left` of the eight codes in v
        // (branch-free inside: the codes behind the end of a run add 1 to a trash word of their own lane)
        auto tally = [&](const uint4 v, const uint32_t left, uint32_t *hist) {
            if (left == 0 || PROBE(1)) return;
            const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int e8 = 0; e8 < 8; ++e8) {
                const uint32_t e = (e8 & 1) ? (w[e8 >> 1] >> 16) : (w[e8 >> 1] & 0xFFFFu);
                uint32_t col = e;
                if constexpr (CANON) {
                    const uint32_t wd = (e >> 5) & (wps - 1u);   // (padding behind a run is arbitrary: stay inside the tables)
                    col = s_pref[wd] + (uint32_t)__popc(s_bits[wd] & ~(0xFFFFFFFFu << (e & 31u)));
                }
                atomicAdd((uint32_t)e8 < left ? hist + col : s_trash + (tid & 31), 1u);
            }
        };
        auto left_of = [&](const uint32_t cnt, const uint32_t first) -> uint32_t { return cnt > first ? cnt - first : 0u; };
        // (measured and not kept: u32 rows leave the histogram intact, so the next sequence could UN-COUNT the previous codes
        // with the same atomics instead of zeroing 64 KB — 1.61 - 1.70 ms against 1.57 ms on config 5ii: the second rank
        // look-up per code costs more than sixteen conflict-free 128-bit stores per thread)
        uint4 v0, v1;
        first_codes(0, v0, v1);
        for (uint64_t seq = seq0; seq < seq_end; ++seq) {
            const uint32_t si = (uint32_t)(seq - seq0);
            const uint32_t t0 = s_tp[si], t1 = s_tp[si + 1];
            const uint32_t rd = s_rd[si * NQW + qw];
            uint32_t r = t0 + qw;
            uint32_t cnt = rd & 0xFFFFu;
            // the bulk copy that last used this buffer must have read it: the previous part (one buffer) or the one
            // before it (two buffers, see `two`)
            uint32_t *hist = csm + (two ? (si & 1u) * hist_half : 0u);
            if (tid == 0) {
                if (two) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                else bulk_wait_read();
            }
            __syncthreads();
            if (!PROBE(2)) for (uint32_t i = tid * 4u; i < bins; i += CK_THREADS * 4u) *reinterpret_cast<uint4 *>(hist + i) = zero4;
            __syncthreads();
            // ---- count: one quarter-warp per run (sequences with more than NQW tiles take further rounds)
            const bool extra = cnt > 128u || t1 - t0 > NQW;
            if (r < t1) {
                tally(v0, left_of(cnt, 8u * ql), hist);
                tally(v1, left_of(cnt, 64u + 8u * ql), hist);
            }
            if (extra) {
                const uint4 *src = reinterpret_cast<const uint4 *>(pool + (uint64_t)r * BK_TILE_CAP + ((uint64_t)(rd >> 16) << 3));
                while (r < t1) {
                    for (uint32_t q = ql + 16; 8u * q < cnt; q += 8) tally(__ldg(src + q), cnt - 8u * q, hist);
                    r += NQW;
                    if (r < t1) {
                        const uint32_t rd2 = __ldg(runs + r);
                        cnt = rd2 & 0xFFFFu;
                        src = reinterpret_cast<const uint4 *>(pool + (uint64_t)r * BK_TILE_CAP + ((uint64_t)(rd2 >> 16) << 3));
                        for (uint32_t q = ql; 8u * q < cnt; q += 8) tally(__ldg(src + q), cnt - 8u * q, hist);
                        cnt = 0;   // (this run is done)
                    }
                }
            }
            if (seq + 1 < seq_end) first_codes(si + 1, v0, v1);
            // ---- normalise and write this part of the row
            const unsigned long long total = s_tot[si];
            if (tid == 0 && seg == 0 && p.totals_out) p.totals_out[seq] = total;
            T *row = reinterpret_cast<T *>(p.out) + seq * p.dim + col0;
            const bool bulk = OUT != OUT_F64 && (reinterpret_cast<uintptr_t>(row) & 15) == 0 && (bins & 3) == 0;
            if constexpr (OUT == OUT_U32) {   // counts leave as they are: one barrier between the last atomic and the copy
                fence_async_smem();
                __syncthreads();
                if (bulk) {
                    if (tid == 0 && !PROBE(4)) bulk_store(row, hist, bins * 4u);
                } else {
                    for (uint32_t i = tid; i < bins; i += CK_THREADS) row[i] = hist[i];
                }
                continue;
            }
            __syncthreads();
            const uint64_t dv = norm_divisor(total, p.norm_mode, p.canonical);
            const float dF = (float)dv, rinv = __frcp_rn(dF);
            const double dD = (double)dv;
            const bool small = dv < (1ULL << 23);
            if constexpr (OUT == OUT_F64) {
                for (uint32_t i = tid; i < bins; i += CK_THREADS) row[i] = cvt_count<OUT_F64, NORM, false>(hist[i], dF, rinv, dD);
            } else {
                if (bulk) {
                    if constexpr (OUT == OUT_F32) {   // counts -> floats in place
                        for (uint32_t i = tid * 4u; i < bins; i += CK_THREADS * 4u) {
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
                    if (tid == 0 && !PROBE(4)) bulk_store(row, hist, bins * 4u);
                } else {   // a part that does not start on a 16-byte boundary: plain coalesced stores
                    for (uint32_t i = tid; i < bins; i += CK_THREADS)
                        row[i] = small ? cvt_count<OUT, NORM, true>(hist[i], dF, rinv, dD) : cvt_count<OUT, NORM, false>(hist[i], dF, rinv, dD);
                }
            }
        }
    }
    if (tid == 0) bulk_wait_all();
}

}  // namespace ktb
