// kv.cu — K3/K4/K5: radix partition, in-partition sort, segmented reduce over 16-byte records.
//
// Replaces the reference's shuffle/sort/reduce for (key, value) records:
//   Splitter.partition            dampr/base.py:6-8          hash(key) % n_partitions
//   CSDatasetWriter.flush         dampr/dataset.py:236-253   route every record to its partition
//   SortedWriter._write_to_gzip   dampr/dataset.py:162-164   list.sort(key=itemgetter(0)) (stable)
//   MergeDataset.read             dampr/dataset.py:571-579   k-way merge of sorted runs
//   Dataset.grouped_read          dampr/dataset.py:429-433   group adjacent equal keys
//   ARReduce._reduce / Reduce     dampr/dampr.py:678-683, dampr/base.py:204-207  fold a group
//
// Design (B200): MSD radix partition on the top bits of the (transformed) key, at most 10 bits
// per level, two levels for up to ~3e9 records; every level is histogram -> scan -> scatter with
// per-CTA contiguous record ranges ("pieces") so ranks are stable and no atomics touch HBM. The
// scatter ranks records inside a warp with one ballot per digit bit, stages a tile in shared memory
// bucket-major and writes each bucket's run with one TMA bulk store (cp.async.bulk.global.shared::cta)
// or coalesced 16-byte stores. Leaves (whole segments, <= 4096 records) are sorted in shared memory
// by a counting sort on the next key bits, after which every record ranks itself inside its bin by
// (key, input position), which makes the whole sort stable and deterministic; the same kernel
// optionally folds each key group (segmented reduce, or a shared-memory hash aggregate for commutative
// integer folds) before anything is written back.
#include <algorithm>

#include "common.cuh"

namespace {

constexpr int P_THREADS = 512;
constexpr int P_WARPS = P_THREADS / 32;
constexpr int P_TILE = 4096;                 // records per tile
constexpr int P_PER_WARP = P_TILE / P_WARPS;  // 256
constexpr int P_ROUNDS = P_PER_WARP / 32;     // 8
constexpr int P_MAX_BITS = 10;
constexpr int P_MAX_NB = 1 << P_MAX_BITS;

constexpr int L_THREADS = 512;
constexpr int L_CAP = 4096;    // records per leaf chunk
constexpr int L_BINS = 8192;   // counting-sort bins

struct Piece {
    u64 start, end;
    u32 seg;
    u32 pad;
};

struct DigitSpec {
    int xf;     // key transform
    u64 base;   // subtracted after the transform
    int shift;  // digit = ((xf(key) - base) >> shift) & mask
    u32 mask;
    int owner_mod;  // if > 0: digit = mix64(key) % owner_mod (exchange partition)
};

__device__ __forceinline__ u32 digit_of(u64 key, const DigitSpec &d) {
    if (d.owner_mod > 0) return (u32)(mix64(key) % (u64)d.owner_mod);
    return (u32)(((key_xform(key, d.xf) - d.base) >> d.shift) & d.mask);
}

// ---- level histogram --------------------------------------------------------------------------
__global__ void __launch_bounds__(P_THREADS)
part_hist_kernel(const ulonglong2 *__restrict__ in, const Piece *__restrict__ pieces,
                 const u32 *__restrict__ cta_piece_begin, DigitSpec ds, u32 nb,
                 u32 *__restrict__ piece_hist) {
    __shared__ u32 sh[P_MAX_NB];
    const u32 pb = cta_piece_begin[blockIdx.x], pe = cta_piece_begin[blockIdx.x + 1];
    for (u32 p = pb; p < pe; ++p) {
        for (u32 b = threadIdx.x; b < nb; b += blockDim.x) sh[b] = 0;
        __syncthreads();
        const Piece pc = pieces[p];
        for (u64 i = pc.start + threadIdx.x; i < pc.end; i += blockDim.x) {
            u64 key = in[i].x;
            atomicAdd(&sh[digit_of(key, ds)], 1u);
        }
        __syncthreads();
        for (u32 b = threadIdx.x; b < nb; b += blockDim.x) piece_hist[(u64)p * nb + b] = sh[b];
        __syncthreads();
    }
}

// ---- per-segment scan: bucket bases, per-piece offsets, next-level segment offsets ------------
__global__ void part_scan_kernel(const u32 *__restrict__ piece_hist, const u32 *__restrict__ seg_piece_begin,
                                 const u64 *__restrict__ seg_off, u32 nb, u64 *__restrict__ piece_off,
                                 u64 *__restrict__ next_seg_off) {
    __shared__ u64 wsum[32];
    const u32 s = blockIdx.x;
    const u32 b = threadIdx.x;  // blockDim.x == nb rounded up to 32
    const u32 pb = seg_piece_begin[s], pe = seg_piece_begin[s + 1];
    u64 tot = 0;
    if (b < nb)
        for (u32 p = pb; p < pe; ++p) tot += piece_hist[(u64)p * nb + b];
    // block exclusive scan of tot
    u64 v = tot;
    for (int d = 1; d < 32; d <<= 1) {
        u64 o = __shfl_up_sync(0xFFFFFFFFu, v, d);
        if ((int)(threadIdx.x & 31) >= d) v += o;
    }
    if ((threadIdx.x & 31) == 31) wsum[threadIdx.x >> 5] = v;
    __syncthreads();
    u64 woff = 0;
    for (u32 w = 0; w < (threadIdx.x >> 5); ++w) woff += wsum[w];
    u64 excl = woff + v - tot;
    if (b < nb) {
        u64 run = seg_off[s] + excl;
        next_seg_off[(u64)s * nb + b] = run;
        for (u32 p = pb; p < pe; ++p) {
            u32 h = piece_hist[(u64)p * nb + b];
            piece_off[(u64)p * nb + b] = run;
            run += h;
        }
    }
}

// ---- level scatter ---------------------------------------------------------------------------
struct ScatterSmem {
    alignas(16) ulonglong2 stage[P_TILE];
    u64 run_off[P_MAX_NB];
    u16 warp_hist[P_WARPS][P_MAX_NB];
    u32 local_base[P_MAX_NB + 1];
    u32 tile_cnt[P_MAX_NB];
    u32 wsum[32];
};

template <bool USE_TMA>
__global__ void __launch_bounds__(P_THREADS, 2)
part_scatter_kernel(const ulonglong2 *__restrict__ in, ulonglong2 *__restrict__ out,
                    const Piece *__restrict__ pieces, const u32 *__restrict__ cta_piece_begin,
                    DigitSpec ds, u32 nb, const u64 *__restrict__ piece_off) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    ScatterSmem &s = *reinterpret_cast<ScatterSmem *>(smem_raw);
    const u32 tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const u32 pb = cta_piece_begin[blockIdx.x], pe = cta_piece_begin[blockIdx.x + 1];
    u32 dbits = 0;
    while ((1u << dbits) < nb) ++dbits;

    for (u32 p = pb; p < pe; ++p) {
        const Piece pc = pieces[p];
        for (u32 b = tid; b < nb; b += P_THREADS) s.run_off[b] = piece_off[(u64)p * nb + b];
        __syncthreads();
        for (u64 t0 = pc.start; t0 < pc.end; t0 += P_TILE) {
            const u32 tn = (u32)min((u64)P_TILE, pc.end - t0);
            for (u32 i = tid; i < P_WARPS * nb; i += P_THREADS) (&s.warp_hist[0][0])[(i / nb) * P_MAX_NB + (i % nb)] = 0;
            __syncthreads();
            // ---- load + stable rank inside the warp's contiguous slice -------------------------
            ulonglong2 rec[P_ROUNDS];
            u32 dig[P_ROUNDS];
            u32 rnk[P_ROUNDS];
            // all loads of the tile first (8 independent 16-byte loads in flight per thread) ...
#pragma unroll
            for (int r = 0; r < P_ROUNDS; ++r) {
                u32 li = warp * P_PER_WARP + r * 32 + lane;
                if (li < tn) rec[r] = in[t0 + li];
            }
            // the tile after this one: pull it into L2 while this one is ranked and staged
            if (t0 + P_TILE < pc.end) {
                const char *nxt = reinterpret_cast<const char *>(in + t0 + P_TILE);
                const u32 nbytes = (u32)min((u64)P_TILE, pc.end - t0 - P_TILE) * 16u;
                if (tid * 128u < nbytes) asm volatile("prefetch.global.L2 [%0];" ::"l"(nxt + tid * 128u));
            }
            // ... then the stable ranking. Pass 1 (registers only): digit and same-digit lane set per round,
            // from one ballot per digit bit (MATCH.ANY costs one pass per distinct value, i.e. ~32 passes
            // with 10-bit digits).
            // Per round one packed word: digit (10 bits) | leader lane << 10 | rank among the set's lanes << 15
            // | (set size - 1) << 20 | valid << 25.
#pragma unroll
            for (int r = 0; r < P_ROUNDS; ++r) {
                const u32 li = warp * P_PER_WARP + r * 32 + lane;
                const bool valid = li < tn;
                const u32 vmask = __ballot_sync(0xFFFFFFFFu, valid);
                const u32 d = valid ? digit_of(rec[r].x, ds) : 0u;
                u32 peers = vmask;
                for (u32 bit = 0; bit < dbits; ++bit) {
                    const u32 mine = (d >> bit) & 1u;
                    const u32 bal = __ballot_sync(0xFFFFFFFFu, mine);
                    peers &= mine ? bal : ~bal;
                }
                dig[r] = valid ? (d | (((u32)__ffs(peers) - 1u) << 10) | ((u32)__popc(peers & ((1u << lane) - 1u)) << 15) |
                                  (((u32)__popc(peers) - 1u) << 20) | (1u << 25))
                               : 0u;
            }
            // Pass 2: the leader of every lane set bumps the warp's digit counter. The eight shared-memory
            // atomics of a thread do not depend on each other, so they pipeline (same-address ones complete
            // in issue order); counters are u16 pairs updated through their 32-bit word.
            u32 *wh32 = reinterpret_cast<u32 *>(s.warp_hist[warp]);
#pragma unroll
            for (int r = 0; r < P_ROUNDS; ++r) {
                const u32 pk = dig[r];
                rnk[r] = 0;
                if ((pk >> 25) && ((pk >> 10) & 31u) == lane) {
                    const u32 d = pk & 1023u, c = ((pk >> 20) & 31u) + 1u;
                    const u32 w = atomicAdd(&wh32[d >> 1], (d & 1u) ? (c << 16) : c);
                    rnk[r] = (d & 1u) ? (w >> 16) : (w & 0xFFFFu);
                }
            }
#pragma unroll
            for (int r = 0; r < P_ROUNDS; ++r) {
                const u32 pk = dig[r];
                const u32 old = __shfl_sync(0xFFFFFFFFu, rnk[r], (pk >> 25) ? ((pk >> 10) & 31u) : lane);
                rnk[r] = old + ((pk >> 15) & 31u);
                dig[r] = (pk >> 25) ? (pk & 1023u) : 0xFFFFFFFFu;
            }
            __syncthreads();
            // ---- per-bucket exclusive scan over warps, then over buckets ------------------------
            for (u32 b = tid; b < nb; b += P_THREADS) {
                u32 run = 0;
#pragma unroll
                for (int w = 0; w < P_WARPS; ++w) {
                    u32 c = s.warp_hist[w][b];
                    s.warp_hist[w][b] = (u16)run;
                    run += c;
                }
                s.tile_cnt[b] = run;
            }
            __syncthreads();
            {
                // exclusive scan of tile_cnt[0..nb) with 512 threads, 2 buckets per thread
                u32 c0 = (2 * tid < nb) ? s.tile_cnt[2 * tid] : 0;
                u32 c1 = (2 * tid + 1 < nb) ? s.tile_cnt[2 * tid + 1] : 0;
                u32 v = c0 + c1;
                u32 tsum = v;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    u32 o = __shfl_up_sync(0xFFFFFFFFu, v, d);
                    if ((int)lane >= d) v += o;
                }
                if (lane == 31) s.wsum[warp] = v;
                __syncthreads();
                u32 woff = 0;
                for (u32 w = 0; w < warp; ++w) woff += s.wsum[w];
                u32 excl = woff + v - tsum;
                if (2 * tid < nb) s.local_base[2 * tid] = excl;
                if (2 * tid + 1 < nb) s.local_base[2 * tid + 1] = excl + c0;
            }
            __syncthreads();
            // ---- stage bucket-major in shared memory ------------------------------------------
#pragma unroll
            for (int r = 0; r < P_ROUNDS; ++r) {
                if (dig[r] != 0xFFFFFFFFu) {
                    u32 pos = s.local_base[dig[r]] + s.warp_hist[warp][dig[r]] + rnk[r];
                    s.stage[pos] = rec[r];
                }
            }
            if (USE_TMA) fence_proxy_async();
            __syncthreads();
            // ---- write every bucket's run ------------------------------------------------------
            if (USE_TMA) {
                for (u32 b = tid; b < nb; b += P_THREADS) {
                    u32 c = s.tile_cnt[b];
                    if (c) tma_store_1d(out + s.run_off[b], &s.stage[s.local_base[b]], c * 16u);
                }
                tma_store_commit();
                tma_store_wait_read();
            } else {
                for (u32 j = tid; j < tn; j += P_THREADS) {
                    ulonglong2 rc = s.stage[j];
                    u32 d = digit_of(rc.x, ds);
                    out[s.run_off[d] + (j - s.local_base[d])] = rc;
                }
            }
            __syncthreads();
            for (u32 b = tid; b < nb; b += P_THREADS) s.run_off[b] += s.tile_cnt[b];
            __syncthreads();
        }
    }
    if (USE_TMA) tma_store_wait_all();
}

// ---- leaf: counting sort + fix-up in shared memory, optional segmented reduce -----------------
struct LeafChunk {
    u64 start;      // first record
    u32 n;          // records (<= L_CAP)
    int bin_shift;  // bin = ((xf(key) - base) >> bin_shift) - bin_base
    u64 bin_base;
};

// 104 KB: two CTAs per SM.  Only the transformed key is kept (every transform is a bijection, the stored
// key is rebuilt with key_unxform on the way out), which is what makes the second CTA fit.
struct LeafSmem {
    alignas(16) u64 sk[L_CAP];   // key_xform(key) - base
    u64 val[L_CAP];
    u16 cnt[L_BINS];   // counts, then bin starts
    u16 cur[L_BINS];   // cursors
    u16 ord[L_CAP];    // sorted order -> record index
    u32 wsum[32];
    u32 total_groups;
};

__device__ __forceinline__ u64 apply_op(int op, u64 acc, u64 v) {
    switch (op) {
        case DAMPR_OP_SUM_I64: return acc + v;
        case DAMPR_OP_COUNT: return acc + v;
        case DAMPR_OP_SUM_F64: return (u64)__double_as_longlong(__longlong_as_double((long long)acc) + __longlong_as_double((long long)v));
        case DAMPR_OP_MIN_I64: return ((long long)v < (long long)acc) ? v : acc;
        case DAMPR_OP_MAX_I64: return ((long long)v > (long long)acc) ? v : acc;
        case DAMPR_OP_MIN_F64: return (__longlong_as_double((long long)v) < __longlong_as_double((long long)acc)) ? v : acc;
        case DAMPR_OP_MAX_F64: return (__longlong_as_double((long long)v) > __longlong_as_double((long long)acc)) ? v : acc;
        case DAMPR_OP_FIRST: return acc;
        case DAMPR_OP_LAST: return v;
    }
    return acc;
}

// REDUCE_OP < 0: sort only (records rewritten in place). Otherwise one record per key group is
// written compacted at the start of the chunk in `out` and the group count to chunk_groups[c].
__global__ void __launch_bounds__(L_THREADS, 2)
leaf_sort_kernel(ulonglong2 *__restrict__ data, ulonglong2 *__restrict__ out, const LeafChunk *__restrict__ chunks,
                 u32 nchunks, int xf, u64 base, int reduce_op, u32 *__restrict__ chunk_groups) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    LeafSmem &s = *reinterpret_cast<LeafSmem *>(smem_raw);
    const u32 tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (u32 c = blockIdx.x; c < nchunks; c += gridDim.x) {
        const LeafChunk ch = chunks[c];
        const u32 n = ch.n;
        if (c + gridDim.x < nchunks) {
            // pull the chunk this CTA handles next into L2 (one 128-byte line per thread per pass)
            const LeafChunk nx = chunks[c + gridDim.x];
            const char *base = reinterpret_cast<const char *>(data + nx.start);
            for (u32 off = tid * 128u; off < nx.n * 16u; off += L_THREADS * 128u)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(base + off));
        }
        {
            uint4 *z = reinterpret_cast<uint4 *>(s.cnt);  // 8 counters per 16-byte store
            for (u32 i = tid; i < L_BINS / 8; i += L_THREADS) z[i] = make_uint4(0, 0, 0, 0);
        }
        __syncthreads();
        // load + histogram
        {
            // L_CAP / L_THREADS independent loads in flight per thread
            ulonglong2 rr[L_CAP / L_THREADS];
#pragma unroll
            for (int k = 0; k < L_CAP / L_THREADS; ++k) {
                u32 i = tid + k * L_THREADS;
                if (i < n) rr[k] = data[ch.start + i];
            }
#pragma unroll
            for (int k = 0; k < L_CAP / L_THREADS; ++k) {
                u32 i = tid + k * L_THREADS;
                if (i < n) {
                    s.val[i] = rr[k].y;
                    s.sk[i] = key_xform(rr[k].x, xf) - base;
                }
            }
        }
        __syncthreads();
        // ---- commutative integer folds under the MIX order (a_group_by(...).sum()/count()/min()/max()):
        // no sort at all. Every record looks its key up in a shared-memory index (slot -> first record
        // holding the key) and folds its value into that record with a shared-memory atomic: O(1) per
        // record however many duplicates a key has (the sort path ranks inside bins: O(duplicates)).
        if (xf == DAMPR_KEY_MIX && (reduce_op == DAMPR_OP_SUM_I64 || reduce_op == DAMPR_OP_COUNT ||
                                    reduce_op == DAMPR_OP_MIN_I64 || reduce_op == DAMPR_OP_MAX_I64)) {
            for (u32 i = tid; i < L_BINS; i += L_THREADS) s.cur[i] = 0;  // cnt[] is the index (zeroed above)
            if (reduce_op == DAMPR_OP_COUNT)
                for (u32 i = tid; i < n; i += L_THREADS) s.val[i] = 1ULL;
            __syncthreads();
            for (u32 i = tid; i < n; i += L_THREADS) {
                const u64 ki = s.sk[i];
                u32 slot = (u32)(ki >> 13) & (L_BINS - 1);  // the partition consumed the top bits
                slot = (slot ^ (u32)(ki >> 37)) & (L_BINS - 1);
                for (;;) {
                    u32 cur = s.cnt[slot];
                    if (cur == 0) {
                        cur = atomicCAS(&s.cnt[slot], (unsigned short)0, (unsigned short)(i + 1));
                        if (cur == 0) {
                            s.cur[i] = 1;  // this record represents its key
                            break;
                        }
                    }
                    if (s.sk[cur - 1] == ki) {
                        unsigned long long *acc = &s.val[cur - 1];
                        const u64 v = s.val[i];
                        if (reduce_op == DAMPR_OP_MIN_I64) atomicMin((long long *)acc, (long long)v);
                        else if (reduce_op == DAMPR_OP_MAX_I64) atomicMax((long long *)acc, (long long)v);
                        else atomicAdd(acc, v);
                        break;
                    }
                    slot = (slot + 1) & (L_BINS - 1);
                }
            }
            __syncthreads();
            // compact the representatives (in record order)
            constexpr int IPT = L_CAP / L_THREADS;
            u32 flags = 0;
#pragma unroll
            for (int k = 0; k < IPT; ++k) {
                u32 p = tid * IPT + k;
                if (p < n && s.cur[p]) flags |= 1u << k;
            }
            u32 cntl = __popc(flags), v = cntl;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                u32 o = __shfl_up_sync(0xFFFFFFFFu, v, d);
                if ((int)lane >= d) v += o;
            }
            if (lane == 31) s.wsum[warp] = v;
            __syncthreads();
            u32 woff = 0;
            for (u32 w = 0; w < warp; ++w) woff += s.wsum[w];
            u32 gidx = woff + v - cntl;
            if (tid == L_THREADS - 1) s.total_groups = woff + v;
#pragma unroll
            for (int k = 0; k < IPT; ++k)
                if (flags & (1u << k)) {
                    const u32 p = tid * IPT + k;
                    out[ch.start + gidx++] = make_ulonglong2(key_unxform(s.sk[p] + base, xf), s.val[p]);
                }
            __syncthreads();
            if (tid == 0) chunk_groups[c] = s.total_groups;
            __syncthreads();
            continue;
        }
        // counting with u16 counters packed two per u32 word: use 32-bit atomics on the pair
        u32 *cnt32 = reinterpret_cast<u32 *>(s.cnt);
        for (u32 i = tid; i < n; i += L_THREADS) {
            u64 bin64 = (s.sk[i] >> ch.bin_shift) - ch.bin_base;
            u32 bin = (u32)min(bin64, (u64)(L_BINS - 1));
            atomicAdd(&cnt32[bin >> 1], (bin & 1) ? 0x10000u : 1u);
        }
        __syncthreads();
        // exclusive scan of cnt[0..L_BINS): 16 bins per thread
        {
            u32 loc[16];
            u32 sum = 0;
            {
                const uint4 *c4 = reinterpret_cast<const uint4 *>(s.cnt) + tid * 2;
                const uint4 a = c4[0], b = c4[1];
                const u32 w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    loc[2 * k] = w[k] & 0xFFFFu;
                    loc[2 * k + 1] = w[k] >> 16;
                    sum += loc[2 * k] + loc[2 * k + 1];
                }
            }
            u32 v = sum;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                u32 o = __shfl_up_sync(0xFFFFFFFFu, v, d);
                if ((int)lane >= d) v += o;
            }
            if (lane == 31) s.wsum[warp] = v;
            __syncthreads();
            u32 woff = 0;
            for (u32 w = 0; w < warp; ++w) woff += s.wsum[w];
            u32 run = woff + v - sum;
            {
                u32 w[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const u32 lo = run;
                    run += loc[2 * k];
                    w[k] = lo | (run << 16);
                    run += loc[2 * k + 1];
                }
                const uint4 a = make_uint4(w[0], w[1], w[2], w[3]), b = make_uint4(w[4], w[5], w[6], w[7]);
                uint4 *c4 = reinterpret_cast<uint4 *>(s.cnt) + tid * 2;  // bin starts
                uint4 *u4 = reinterpret_cast<uint4 *>(s.cur) + tid * 2;  // cursors
                c4[0] = a;
                c4[1] = b;
                u4[0] = a;
                u4[1] = b;
            }
        }
        __syncthreads();
        // scatter indices (order inside a bin is fixed below)
        u32 *cur32 = reinterpret_cast<u32 *>(s.cur);
        for (u32 i = tid; i < n; i += L_THREADS) {
            u64 bin64 = (s.sk[i] >> ch.bin_shift) - ch.bin_base;
            u32 bin = (u32)min(bin64, (u64)(L_BINS - 1));
            u32 old = atomicAdd(&cur32[bin >> 1], (bin & 1) ? 0x10000u : 1u);
            u32 pos = (bin & 1) ? (old >> 16) : (old & 0xFFFFu);
            s.ord[pos] = (u16)i;
        }
        __syncthreads();
        // order inside every bin by (key, input position): each record counts the members of its own
        // bin that precede it (thread per record: all lanes busy, loop length = bin population) and
        // takes that rank. cur[] is free after the scatter and receives the final order.
        for (u32 i = tid; i < n; i += L_THREADS) {
            const u64 ki = s.sk[i];
            u64 bin64 = (ki >> ch.bin_shift) - ch.bin_base;
            u32 bin = (u32)min(bin64, (u64)(L_BINS - 1));
            u32 st = s.cnt[bin];
            u32 en = (bin + 1 < L_BINS) ? s.cnt[bin + 1] : n;
            u32 r = 0;
            for (u32 j = st; j < en; ++j) {
                u32 o = s.ord[j];
                u64 ko = s.sk[o];
                r += (ko < ki || (ko == ki && o < i)) ? 1u : 0u;
            }
            s.cur[st + r] = (u16)i;
        }
        __syncthreads();
        const u16 *fin = s.cur;   // final order: position -> record index
        if (reduce_op < 0) {
            for (u32 i = tid; i < n; i += L_THREADS) {
                const u32 o = fin[i];
                out[ch.start + i] = make_ulonglong2(key_unxform(s.sk[o] + base, xf), s.val[o]);
            }
            __syncthreads();
            continue;
        }
        // ---- segmented reduce: one thread per group head walks its group ----------------------
        {
            // heads per thread-strided position; ranks via block scan over 8 consecutive positions
            constexpr int IPT = L_CAP / L_THREADS;  // 8
            u32 headbits = 0;
#pragma unroll
            for (int k = 0; k < IPT; ++k) {
                u32 p = tid * IPT + k;
                if (p < n) {
                    bool head = (p == 0) || (s.sk[fin[p]] != s.sk[fin[p - 1]]);
                    headbits |= head ? (1u << k) : 0u;
                }
            }
            u32 cntl = __popc(headbits);
            u32 v = cntl;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                u32 o = __shfl_up_sync(0xFFFFFFFFu, v, d);
                if ((int)lane >= d) v += o;
            }
            if (lane == 31) s.wsum[warp] = v;
            __syncthreads();
            u32 woff = 0;
            for (u32 w = 0; w < warp; ++w) woff += s.wsum[w];
            u32 gidx = woff + v - cntl;
            if (tid == L_THREADS - 1) s.total_groups = woff + v;
#pragma unroll
            for (int k = 0; k < IPT; ++k) {
                if (headbits & (1u << k)) {
                    u32 p = tid * IPT + k;
                    u64 ksk = s.sk[fin[p]];
                    u64 acc = (reduce_op == DAMPR_OP_COUNT) ? 1ULL : s.val[fin[p]];
                    for (u32 q = p + 1; q < n && s.sk[fin[q]] == ksk; ++q) {
                        u64 val = (reduce_op == DAMPR_OP_COUNT) ? 1ULL : s.val[fin[q]];
                        acc = apply_op(reduce_op, acc, val);
                    }
                    out[ch.start + gidx] = make_ulonglong2(key_unxform(ksk + base, xf), acc);
                    ++gidx;
                }
            }
            __syncthreads();
            if (tid == 0) chunk_groups[c] = s.total_groups;
            __syncthreads();
        }
    }
}

// reduce a whole range that holds a single key (all keys equal): one CTA
__global__ void single_group_reduce_kernel(const ulonglong2 *__restrict__ data, u64 start, u64 n, int op,
                                           ulonglong2 *__restrict__ out_rec) {
    __shared__ u64 part[32];
    u64 acc = 0;
    bool have = false;
    // FIRST/LAST depend on order: handled by thread 0 directly
    if (op == DAMPR_OP_FIRST || op == DAMPR_OP_LAST) {
        if (threadIdx.x == 0) {
            ulonglong2 r = data[start + (op == DAMPR_OP_FIRST ? 0 : n - 1)];
            *out_rec = r;
        }
        return;
    }
    if (op == DAMPR_OP_SUM_F64) {
        // deterministic: fixed strided partition + tree
        double a = 0.0;
        for (u64 i = threadIdx.x; i < n; i += blockDim.x) a += __longlong_as_double((long long)data[start + i].y);
        for (int d = 16; d > 0; d >>= 1) a += __shfl_down_sync(0xFFFFFFFFu, a, d);
        if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = (u64)__double_as_longlong(a);
        __syncthreads();
        if (threadIdx.x == 0) {
            double t = 0.0;
            for (u32 w = 0; w < blockDim.x / 32; ++w) t += __longlong_as_double((long long)part[w]);
            *out_rec = make_ulonglong2(data[start].x, (u64)__double_as_longlong(t));
        }
        return;
    }
    for (u64 i = threadIdx.x; i < n; i += blockDim.x) {
        u64 v = (op == DAMPR_OP_COUNT) ? 1ULL : data[start + i].y;
        acc = have ? apply_op(op, acc, v) : v;
        have = true;
    }
    // combine across lanes / warps (ops here are commutative)
    for (int d = 16; d > 0; d >>= 1) {
        u64 o = __shfl_down_sync(0xFFFFFFFFu, acc, d);
        u32 oh = __shfl_down_sync(0xFFFFFFFFu, (u32)have, d);
        if (oh) {
            acc = have ? apply_op(op, acc, o) : o;
            have = true;
        }
    }
    __shared__ u32 parth[32];
    if ((threadIdx.x & 31) == 0) {
        part[threadIdx.x >> 5] = acc;
        parth[threadIdx.x >> 5] = have;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        u64 t = 0;
        bool th = false;
        for (u32 w = 0; w < blockDim.x / 32; ++w)
            if (parth[w]) {
                t = th ? apply_op(op, t, part[w]) : part[w];
                th = true;
            }
        *out_rec = make_ulonglong2(data[start].x, t);
    }
}

// ---- misc kernels -------------------------------------------------------------------------------
__global__ void minmax_kernel(const ulonglong2 *__restrict__ in, u64 n, int xf, u64 *out /*[2]: min,max*/) {
    u64 mn = ~0ULL, mx = 0;
    for (u64 i = blockIdx.x * (u64)blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) {
        u64 k = key_xform(in[i].x, xf);
        mn = min(mn, k);
        mx = max(mx, k);
    }
    for (int d = 16; d > 0; d >>= 1) {
        mn = min(mn, __shfl_down_sync(0xFFFFFFFFu, mn, d));
        mx = max(mx, __shfl_down_sync(0xFFFFFFFFu, mx, d));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(&out[0], mn);
        atomicMax(&out[1], mx);
    }
}

__global__ void copy_records_kernel(const ulonglong2 *__restrict__ in, ulonglong2 *__restrict__ out, u64 n) {
    for (u64 i = blockIdx.x * (u64)blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) out[i] = in[i];
}

// exclusive scan of chunk group counts (single CTA, sequential over blocks of 1024)
__global__ void scan_u32_to_u64_kernel(const u32 *__restrict__ in, u64 *__restrict__ out, u32 n) {
    __shared__ u64 wsum[32];
    __shared__ u64 carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (u32 base = 0; base < n + 1; base += blockDim.x) {
        u32 i = base + threadIdx.x;
        u64 x = (i < n) ? in[i] : 0;
        u64 v = x;
        for (int d = 1; d < 32; d <<= 1) {
            u64 o = __shfl_up_sync(0xFFFFFFFFu, v, d);
            if ((int)(threadIdx.x & 31) >= d) v += o;
        }
        if ((threadIdx.x & 31) == 31) wsum[threadIdx.x >> 5] = v;
        __syncthreads();
        u64 woff = 0;
        for (u32 w = 0; w < (threadIdx.x >> 5); ++w) woff += wsum[w];
        u64 excl = carry + woff + v - x;
        if (i <= n) out[i] = excl;
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) carry = excl + x;
        __syncthreads();
    }
}

__global__ void gather_groups_kernel(const ulonglong2 *__restrict__ tmp, const LeafChunk *__restrict__ chunks,
                                     const u32 *__restrict__ chunk_groups, const u64 *__restrict__ chunk_out_off,
                                     u32 nchunks, ulonglong2 *__restrict__ out) {
    for (u32 c = blockIdx.x; c < nchunks; c += gridDim.x) {
        u32 g = chunk_groups[c];
        u64 src = chunks[c].start, dst = chunk_out_off[c];
        for (u32 i = threadIdx.x; i < g; i += blockDim.x) out[dst + i] = tmp[src + i];
    }
}

__global__ void group_heads_kernel(const ulonglong2 *__restrict__ in, u64 n, u64 *__restrict__ offsets, u64 cap,
                                   u64 *__restrict__ cursor_unused, u64 *__restrict__ count) {
    // two uses: offsets == nullptr -> count heads; otherwise heads are written at their rank,
    // which requires a scan: this kernel is launched with one CTA per 4096-record tile after
    // tile head counts were scanned into cursor_unused (tile base).
    __shared__ u32 wsum[32];
    const u64 t0 = (u64)blockIdx.x * 4096ULL;
    u32 flags = 0, cnt = 0;
    for (int k = 0; k < 8; ++k) {
        u64 i = t0 + (u64)threadIdx.x * 8 + k;
        if (i < n) {
            bool head = (i == 0) || (in[i].x != in[i - 1].x);
            flags |= head ? (1u << k) : 0u;
        }
    }
    cnt = __popc(flags);
    u32 v = cnt;
    for (int d = 1; d < 32; d <<= 1) {
        u32 o = __shfl_up_sync(0xFFFFFFFFu, v, d);
        if ((int)(threadIdx.x & 31) >= d) v += o;
    }
    if ((threadIdx.x & 31) == 31) wsum[threadIdx.x >> 5] = v;
    __syncthreads();
    u32 woff = 0;
    for (u32 w = 0; w < (threadIdx.x >> 5); ++w) woff += wsum[w];
    if (!offsets) {
        if (threadIdx.x == blockDim.x - 1) ((u32 *)count)[blockIdx.x] = woff + v;
        return;
    }
    u64 g = cursor_unused[blockIdx.x] + woff + v - cnt;
    for (int k = 0; k < 8; ++k)
        if (flags & (1u << k)) {
            if (g < cap) offsets[g] = t0 + (u64)threadIdx.x * 8 + k;
            ++g;
        }
}

// ---- host-side orchestration ------------------------------------------------------------------
static int bits_for(u64 n, u64 leaf_avg) {
    int b = 0;
    while (((n + leaf_avg - 1) / leaf_avg) > (1ULL << b)) ++b;
    return b;
}

struct SortOut {
    // sort only: result lives in `cur`. reduce: groups appended to `groups` (device) in key order
    ulonglong2 *groups = nullptr;
    u64 n_groups = 0;
};

static int g_use_tma = 1;

// One partition level over records [start, start+n) of `src` into `dst`: every segment of
// seg_off is split by the digit `ds` into nb sub-segments (stable). seg_off is replaced by the
// (S*nb + 1) offsets of the next level.
static int partition_level(dampr_ctx *ctx, const ulonglong2 *src, ulonglong2 *dst, u64 start, u64 n,
                           std::vector<u64> &seg_off, DigitSpec ds, u32 nb) {
    const int G = ctx->num_sms * 2;
    const u64 S = seg_off.size() - 1;
    // pieces: split [start, start+n) at CTA-range boundaries and at segment boundaries
    u64 R = (n + G - 1) / G;
    R = ((R + P_TILE - 1) / P_TILE) * P_TILE;
    std::vector<Piece> pieces;
    std::vector<u32> cta_pb(G + 1, 0), seg_pb(S + 1, 0);
    {
        u64 s_idx = 0;
        for (int c = 0; c < G; ++c) {
            u64 lo = start + std::min(n, (u64)c * R), hi = start + std::min(n, (u64)(c + 1) * R);
            cta_pb[c] = (u32)pieces.size();
            u64 pos = lo;
            while (pos < hi) {
                while (seg_off[s_idx + 1] <= pos) ++s_idx;
                u64 e = std::min(hi, seg_off[s_idx + 1]);
                pieces.push_back(Piece{pos, e, (u32)s_idx, 0});
                pos = e;
            }
        }
        cta_pb[G] = (u32)pieces.size();
        // seg_piece_begin: pieces are ordered by start, hence by segment
        u32 pi = 0;
        for (u64 s2 = 0; s2 < S; ++s2) {
            while (pi < pieces.size() && pieces[pi].seg < s2) ++pi;
            seg_pb[s2] = pi;
        }
        seg_pb[S] = (u32)pieces.size();
        // empty segments: seg_pb must be monotone; fix holes
        for (u64 s2 = S; s2-- > 0;)
            if (seg_pb[s2] > seg_pb[s2 + 1]) seg_pb[s2] = seg_pb[s2 + 1];
    }
    const u64 NP = pieces.size();
    DevBuf d_pieces, d_cta_pb, d_seg_pb, d_seg_off, d_hist, d_poff, d_next;
    CUDA_TRY(ctx, d_pieces.alloc(NP * sizeof(Piece)));
    CUDA_TRY(ctx, d_cta_pb.alloc((G + 1) * 4));
    CUDA_TRY(ctx, d_seg_pb.alloc((S + 1) * 4));
    CUDA_TRY(ctx, d_seg_off.alloc((S + 1) * 8));
    CUDA_TRY(ctx, d_hist.alloc(NP * nb * 4));
    CUDA_TRY(ctx, d_poff.alloc(NP * nb * 8));
    CUDA_TRY(ctx, d_next.alloc((S * nb + 1) * 8));
    CUDA_TRY(ctx, cudaMemcpyAsync(d_pieces.p, pieces.data(), NP * sizeof(Piece), cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(d_cta_pb.p, cta_pb.data(), (G + 1) * 4, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(d_seg_pb.p, seg_pb.data(), (S + 1) * 4, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(d_seg_off.p, seg_off.data(), (S + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
    {
        ScopedTimer tm(ctx, DAMPR_K_PART_HIST);
        part_hist_kernel<<<G, P_THREADS, 0, ctx->stream>>>(src, (const Piece *)d_pieces.p, (const u32 *)d_cta_pb.p, ds,
                                                          nb, (u32 *)d_hist.p);
    }
    {
        ScopedTimer tm(ctx, DAMPR_K_MISC);
        u32 thr = ((nb + 31) / 32) * 32;
        part_scan_kernel<<<(unsigned)S, thr, 0, ctx->stream>>>((const u32 *)d_hist.p, (const u32 *)d_seg_pb.p,
                                                              (const u64 *)d_seg_off.p, nb, (u64 *)d_poff.p,
                                                              (u64 *)d_next.p);
    }
    {
        size_t smem = sizeof(ScatterSmem);
        ScopedTimer tm(ctx, DAMPR_K_PART_SCATTER);
        if (g_use_tma) {
            cudaFuncSetAttribute(part_scatter_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            part_scatter_kernel<true><<<G, P_THREADS, smem, ctx->stream>>>(
                src, dst, (const Piece *)d_pieces.p, (const u32 *)d_cta_pb.p, ds, nb, (const u64 *)d_poff.p);
        } else {
            cudaFuncSetAttribute(part_scatter_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            part_scatter_kernel<false><<<G, P_THREADS, smem, ctx->stream>>>(
                src, dst, (const Piece *)d_pieces.p, (const u32 *)d_cta_pb.p, ds, nb, (const u64 *)d_poff.p);
        }
    }
    CUDA_TRY(ctx, cudaGetLastError());
    // next level's segment offsets
    std::vector<u64> next(S * nb + 1);
    CUDA_TRY(ctx, cudaMemcpyAsync(next.data(), d_next.p, S * nb * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    next[S * nb] = start + n;
    seg_off.swap(next);
    return DAMPR_OK;
}

// Sort records [start, start+n) that currently live in `cur` (scratch = `alt`, same indexing).
// Digits are taken from bit `top` downwards of (xf(key) - base). On return the sorted range is in
// `cur` (copied back if an odd number of levels ran). If reduce_op >= 0 the key groups of the range
// are appended, in order, to gout[*gcount...] and nothing is guaranteed about `cur`.
static int sort_range(dampr_ctx *ctx, ulonglong2 *cur, ulonglong2 *alt, u64 start, u64 n, int xf, u64 base,
                      int top, int reduce_op, ulonglong2 *gout, u64 *gcount, int depth) {
    if (n == 0) return DAMPR_OK;
    if (top <= 0 || depth > 12) {
        // all keys equal: already "sorted" (stable); a single group when reducing
        if (reduce_op >= 0) {
            ScopedTimer tm(ctx, DAMPR_K_SEG_REDUCE);
            single_group_reduce_kernel<<<1, 1024, 0, ctx->stream>>>(cur, start, n, reduce_op, gout + *gcount);
            *gcount += 1;
        }
        CUDA_TRY(ctx, cudaGetLastError());
        return DAMPR_OK;
    }
    // ---- plan the levels ---------------------------------------------------------------------
    int total_bits = std::min(top, bits_for(n, 2600));
    int nlev = (total_bits + P_MAX_BITS - 1) / P_MAX_BITS;
    std::vector<int> lev_bits;
    {
        int left = total_bits;
        for (int l = 0; l < nlev; ++l) {
            int b = (left + (nlev - l) - 1) / (nlev - l);
            lev_bits.push_back(b);
            left -= b;
        }
    }
    // segment offsets of the current level (host copy), relative to absolute record index
    std::vector<u64> seg_off{start, start + n};
    ulonglong2 *src = cur, *dst = alt;
    int consumed = 0;
    for (int l = 0; l < nlev; ++l) {
        const int bits = lev_bits[l];
        const u32 nb = 1u << bits;
        DigitSpec ds{xf, base, top - consumed - bits, nb - 1, 0};
        int rc = partition_level(ctx, src, dst, start, n, seg_off, ds, nb);
        if (rc) return rc;
        std::swap(src, dst);
        consumed += bits;
    }
    // data now in `src`
    // ---- leaves --------------------------------------------------------------------------------
    const int rem_top = top - consumed;  // bits left below the partition digits
    const u64 S = seg_off.size() - 1;
    std::vector<LeafChunk> chunks;
    struct Big {
        u64 start, n;
        size_t after_chunk;  // chunks.size() when encountered (ordering of reduce output)
        u64 seg;             // flattened digit value of the segment
    };
    std::vector<Big> bigs;
    {
        u64 s0 = 0;
        while (s0 < S) {
            u64 sz = seg_off[s0 + 1] - seg_off[s0];
            if (sz == 0) {
                ++s0;
                continue;
            }
            if (sz > L_CAP) {
                bigs.push_back(Big{seg_off[s0], sz, chunks.size(), s0});
                ++s0;
                continue;
            }
            u64 s1 = s0 + 1;
            u64 tot = sz;
            while (s1 < S && tot + (seg_off[s1 + 1] - seg_off[s1]) <= L_CAP && (s1 - s0) < (u64)L_BINS) {
                tot += seg_off[s1 + 1] - seg_off[s1];
                ++s1;
            }
            // bins: (segment index relative to s0) << k | next k key bits
            u64 nseg = s1 - s0;
            int k = 0;
            while (((nseg << (k + 1)) <= (u64)L_BINS) && (k + 1) <= rem_top) ++k;
            LeafChunk lc;
            lc.start = seg_off[s0];
            lc.n = (u32)tot;
            lc.bin_shift = rem_top - k;
            // value of ((sk - base) >> rem_top) for segment s0 is its index in the flattened
            // digit space; the whole range shares the bits above `top` (zero after subtracting base
            // for the outermost call, or equal for recursive calls), so take the low `consumed` bits
            lc.bin_base = 0;  // filled below from s0 (needs the digit prefix)
            lc.bin_base = ((u64)s0) << k;
            chunks.push_back(lc);
            s0 = s1;
        }
    }
    // the segment index of a record is ((sk-base) >> rem_top) & (2^consumed - 1) only if the bits
    // above `top` are zero; recursive calls pass a base that makes them zero (see below).
    const size_t nchunks = chunks.size();
    DevBuf d_chunks, d_cgroups, d_coff;
    ulonglong2 *leaf_out = (reduce_op >= 0) ? dst : src;  // reduce writes compact groups into dst
    if (nchunks) {
        CUDA_TRY(ctx, d_chunks.alloc(nchunks * sizeof(LeafChunk)));
        CUDA_TRY(ctx, cudaMemcpyAsync(d_chunks.p, chunks.data(), nchunks * sizeof(LeafChunk), cudaMemcpyHostToDevice,
                                      ctx->stream));
        if (reduce_op >= 0) {
            CUDA_TRY(ctx, d_cgroups.alloc(nchunks * 4));
            CUDA_TRY(ctx, d_coff.alloc((nchunks + 1) * 8));
        }
        size_t smem = sizeof(LeafSmem);
        CUDA_TRY(ctx, cudaFuncSetAttribute(leaf_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        u32 grid = (u32)std::min<size_t>(nchunks, (size_t)ctx->num_sms * 2);
        {
            ScopedTimer tm(ctx, DAMPR_K_LEAF_SORT);
            leaf_sort_kernel<<<grid, L_THREADS, smem, ctx->stream>>>(src, leaf_out, (const LeafChunk *)d_chunks.p,
                                                                    (u32)nchunks, xf, base, reduce_op,
                                                                    (u32 *)d_cgroups.p);
        }
        CUDA_TRY(ctx, cudaGetLastError());
    }
    if (reduce_op < 0) {
        // big segments: recurse in place (their data is in `src`; scratch is `dst`)
        for (auto &b : bigs) {
            // keys in the segment share all bits >= rem_top: new base clears them
            int rc = sort_range(ctx, src, dst, b.start, b.n, xf, base + (b.seg << rem_top), rem_top, -1, nullptr, nullptr,
                                depth + 1);
            if (rc) return rc;
        }
        if (src != cur) {
            ScopedTimer tm(ctx, DAMPR_K_MISC);
            copy_records_kernel<<<ctx->num_sms * 4, 256, 0, ctx->stream>>>(src + start, cur + start, n);
            CUDA_TRY(ctx, cudaGetLastError());
        }
        return DAMPR_OK;
    }
    // ---- reduce: gather chunk groups in order, interleaving the big segments --------------------
    std::vector<u32> cg(nchunks);
    if (nchunks) {
        CUDA_TRY(ctx, cudaMemcpyAsync(cg.data(), d_cgroups.p, nchunks * 4, cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    }
    // output offsets per chunk (host scan; big segments produce a variable number -> processed in order)
    size_t ci = 0;
    size_t bi = 0;
    while (ci < nchunks || bi < bigs.size()) {
        size_t run_end = (bi < bigs.size()) ? bigs[bi].after_chunk : nchunks;
        if (ci < run_end) {
            // gather chunks [ci, run_end)
            std::vector<u64> off(run_end - ci);
            u64 run = *gcount;
            for (size_t c = ci; c < run_end; ++c) {
                off[c - ci] = run;
                run += cg[c];
            }
            DevBuf d_off;
            CUDA_TRY(ctx, d_off.alloc(off.size() * 8));
            CUDA_TRY(ctx, cudaMemcpyAsync(d_off.p, off.data(), off.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
            {
                ScopedTimer tm(ctx, DAMPR_K_SEG_REDUCE);
                gather_groups_kernel<<<(unsigned)std::min<size_t>(run_end - ci, (size_t)ctx->num_sms * 8), 256, 0, ctx->stream>>>(
                    leaf_out, (const LeafChunk *)d_chunks.p + ci, (const u32 *)d_cgroups.p + ci, (const u64 *)d_off.p,
                    (u32)(run_end - ci), gout);
            }
            CUDA_TRY(ctx, cudaGetLastError());
            CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));  // d_off lifetime
            *gcount = run;
            ci = run_end;
        }
        if (bi < bigs.size() && bigs[bi].after_chunk == ci) {
            int rc = sort_range(ctx, src, dst, bigs[bi].start, bigs[bi].n, xf, base + (bigs[bi].seg << rem_top), rem_top,
                                reduce_op, gout, gcount,
                                depth + 1);
            if (rc) return rc;
            ++bi;
        }
    }
    return DAMPR_OK;
}

static int ensure_alt(dampr_ctx *ctx, dampr_kv *kv) {
    if (!kv->alt) {
        kv->alt = (ulonglong2 *)pool_alloc(ctx, (kv->capacity ? kv->capacity : 1) * sizeof(ulonglong2));
        if (!kv->alt) {
            ctx->err = "device allocation (kv scratch) failed";
            return DAMPR_ERR_NOMEM;
        }
    }
    return DAMPR_OK;
}

// key range -> (base, top)
static int key_range(dampr_ctx *ctx, dampr_kv *kv, int xf, u64 *base, int *top) {
    if (xf == DAMPR_KEY_MIX) {
        *base = 0;
        *top = 64;
        return DAMPR_OK;
    }
    u64 init[2] = {~0ULL, 0ULL};
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_scratch, init, 16, cudaMemcpyHostToDevice, ctx->stream));
    {
        ScopedTimer tm(ctx, DAMPR_K_MISC);
        minmax_kernel<<<ctx->num_sms * 4, 256, 0, ctx->stream>>>(kv->rec, kv->n, xf, ctx->d_scratch);
    }
    CUDA_TRY(ctx, cudaGetLastError());
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->h_scratch, ctx->d_scratch, 16, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    u64 mn = ctx->h_scratch[0], mx = ctx->h_scratch[1];
    *base = mn;
    u64 span = mx - mn;
    int t = 0;
    while (t < 64 && (span >> t) != 0) ++t;
    *top = t;
    return DAMPR_OK;
}

}  // namespace

extern "C" {

int32_t dampr_set_option(const char *name, int64_t value) {
    if (!name) return DAMPR_ERR_ARG;
    if (!strcmp(name, "scatter_tma")) {
        g_use_tma = value != 0;
        return DAMPR_OK;
    }
    if (!strcmp(name, "text_ctas")) {
        if (value < 2 || value > 4) return DAMPR_ERR_ARG;
        g_text_ctas = (int)value;
        return DAMPR_OK;
    }
    if (!strcmp(name, "text_kernel")) {
        if (value != 1 && value != 2) return DAMPR_ERR_ARG;
        g_text_kernel = (int)value;
        return DAMPR_OK;
    }
    return DAMPR_ERR_ARG;
}

int32_t dampr_kv_sort(dampr_ctx *ctx, dampr_kv *kv, int32_t key_xf) {
    ARG_CHECK(ctx, ctx && kv, "null");
    CtxScope scope_(ctx);
    ARG_CHECK(ctx, key_xf >= 0 && key_xf <= DAMPR_KEY_F64, "unknown key transform");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    wait_uploads(ctx);
    if (kv->n < 2) return DAMPR_OK;
    int rc = ensure_alt(ctx, kv);
    if (rc) return rc;
    u64 base;
    int top;
    rc = key_range(ctx, kv, key_xf, &base, &top);
    if (rc) return rc;
    return sort_range(ctx, kv->rec, kv->alt, 0, kv->n, key_xf, base, top, -1, nullptr, nullptr, 0);
}

int32_t dampr_kv_sort_reduce(dampr_ctx *ctx, dampr_kv *kv, int32_t key_xf, int32_t op, dampr_kv **out) {
    ARG_CHECK(ctx, ctx && kv && out, "null");
    CtxScope scope_(ctx);
    ARG_CHECK(ctx, key_xf >= 0 && key_xf <= DAMPR_KEY_F64, "unknown key transform");
    ARG_CHECK(ctx, op >= 0 && op <= DAMPR_OP_LAST, "unknown reduce op");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    wait_uploads(ctx);
    int rc = dampr_kv_create(ctx, kv->n, out);  // at most n groups
    if (rc) return rc;
    if (kv->n == 0) return DAMPR_OK;
    rc = ensure_alt(ctx, kv);
    if (rc) return rc;
    u64 base;
    int top;
    rc = key_range(ctx, kv, key_xf, &base, &top);
    if (rc) return rc;
    u64 g = 0;
    rc = sort_range(ctx, kv->rec, kv->alt, 0, kv->n, key_xf, base, top, op, (*out)->rec, &g, 0);
    if (rc) return rc;
    (*out)->n = g;
    // the partition levels ping-pong between the two buffers of `kv` and the leaves stage their group
    // records in whichever is free: the input is consumed
    kv->n = 0;
    return DAMPR_OK;
}

int32_t dampr_kv_reduce_by_key(dampr_ctx *ctx, dampr_kv *sorted, int32_t op, dampr_kv **out) {
    ARG_CHECK(ctx, ctx && sorted && out, "null");
    // a key-sorted run is reduced by the same machinery in RAW key order (the levels only re-discover
    // the order the input already has), on a copy: the caller keeps its sorted run
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    wait_uploads(ctx);
    dampr_kv *tmp = nullptr;
    int rc = dampr_kv_create(ctx, sorted->n, &tmp);
    if (rc) return rc;
    if (sorted->n) {
        cudaError_t e = cudaMemcpyAsync(tmp->rec, sorted->rec, sorted->n * sizeof(ulonglong2), cudaMemcpyDeviceToDevice,
                                        ctx->stream);
        if (e != cudaSuccess) {
            dampr_kv_destroy(ctx, tmp);
            ctx->err = std::string("reduce_by_key copy failed: ") + cudaGetErrorString(e);
            return DAMPR_ERR_CUDA;
        }
    }
    tmp->n = sorted->n;
    rc = dampr_kv_sort_reduce(ctx, tmp, DAMPR_KEY_RAW, op, out);
    dampr_kv_destroy(ctx, tmp);
    return rc;
}

int32_t dampr_kv_group_offsets(dampr_ctx *ctx, dampr_kv *sorted, uint64_t *offsets, uint64_t cap,
                               uint64_t *n_groups) {
    ARG_CHECK(ctx, ctx && sorted && n_groups, "null");
    CtxScope scope_(ctx);
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    wait_uploads(ctx);
    const u64 n = sorted->n;
    if (n == 0) {
        *n_groups = 0;
        if (offsets && cap >= 1) offsets[0] = 0;
        return DAMPR_OK;
    }
    const u64 ntiles = (n + 4095) / 4096;
    DevBuf d_cnt, d_base, d_offs;
    CUDA_TRY(ctx, d_cnt.alloc(ntiles * 4));
    CUDA_TRY(ctx, d_base.alloc((ntiles + 1) * 8));
    {
        ScopedTimer tm(ctx, DAMPR_K_SEG_REDUCE);
        group_heads_kernel<<<(unsigned)ntiles, 512, 0, ctx->stream>>>(sorted->rec, n, nullptr, 0, nullptr, (u64 *)d_cnt.p);
    }
    {
        ScopedTimer tm(ctx, DAMPR_K_MISC);
        scan_u32_to_u64_kernel<<<1, 1024, 0, ctx->stream>>>((const u32 *)d_cnt.p, (u64 *)d_base.p, (u32)ntiles);
    }
    CUDA_TRY(ctx, cudaGetLastError());
    u64 total = 0;
    CUDA_TRY(ctx, cudaMemcpyAsync(&total, (u64 *)d_base.p + ntiles, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    *n_groups = total;
    if (!offsets) return DAMPR_OK;
    ARG_CHECK(ctx, cap >= total + 1, "offsets array too small (need n_groups + 1)");
    CUDA_TRY(ctx, d_offs.alloc((total + 1) * 8));
    {
        ScopedTimer tm(ctx, DAMPR_K_SEG_REDUCE);
        group_heads_kernel<<<(unsigned)ntiles, 512, 0, ctx->stream>>>(sorted->rec, n, (u64 *)d_offs.p, total,
                                                                     (u64 *)d_base.p, nullptr);
    }
    CUDA_TRY(ctx, cudaGetLastError());
    CUDA_TRY(ctx, cudaMemcpyAsync(offsets, d_offs.p, total * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    offsets[total] = n;
    return DAMPR_OK;
}

}  // extern "C"

extern "C" {

// destination-contiguous split for the exchange: owner = mix64(key) % n_dest
int32_t dampr_kv_partition_by_owner(dampr_ctx *ctx, dampr_kv *kv, int32_t n_dest, dampr_kv **out,
                                    uint64_t *counts_host) {
    ARG_CHECK(ctx, ctx && kv && out && counts_host, "null");
    CtxScope scope_(ctx);
    ARG_CHECK(ctx, n_dest >= 1 && n_dest <= P_MAX_NB, "n_dest out of range");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    wait_uploads(ctx);
    int rc = dampr_kv_create(ctx, kv->n, out);
    if (rc) return rc;
    (*out)->n = kv->n;
    if (kv->n == 0) {
        for (int i = 0; i < n_dest; ++i) counts_host[i] = 0;
        return DAMPR_OK;
    }
    std::vector<u64> seg_off{0, kv->n};
    DigitSpec ds{DAMPR_KEY_RAW, 0, 0, 0, n_dest};
    rc = partition_level(ctx, kv->rec, (*out)->rec, 0, kv->n, seg_off, ds, (u32)n_dest);
    if (rc) return rc;
    for (int i = 0; i < n_dest; ++i) counts_host[i] = seg_off[i + 1] - seg_off[i];
    return DAMPR_OK;
}

// k-way merge of key-sorted runs (+ optional segmented reduce). The runs are concatenated in run
// order and pushed through the stable partition+leaf pipeline, which yields exactly the stable
// merge (ties by run order, then by position) that heapq.merge produces (dataset.py:571-579).
int32_t dampr_kv_merge(dampr_ctx *ctx, dampr_kv **runs, int32_t n_runs, int32_t key_xf, int32_t op,
                       dampr_kv **out) {
    ARG_CHECK(ctx, ctx && runs && out && n_runs >= 0, "null");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    wait_uploads(ctx);
    u64 total = 0;
    for (int i = 0; i < n_runs; ++i) {
        ARG_CHECK(ctx, runs[i] != nullptr, "null run");
        total += runs[i]->n;
    }
    dampr_kv *cat = nullptr;
    int rc = dampr_kv_create(ctx, total, &cat);
    if (rc) return rc;
    u64 off = 0;
    for (int i = 0; i < n_runs; ++i) {
        if (runs[i]->n)
            CUDA_TRY(ctx, cudaMemcpyAsync(cat->rec + off, runs[i]->rec, runs[i]->n * sizeof(ulonglong2),
                                          cudaMemcpyDeviceToDevice, ctx->stream));
        off += runs[i]->n;
    }
    cat->n = total;
    if (op < 0) {
        rc = dampr_kv_sort(ctx, cat, key_xf);
        if (rc) {
            dampr_kv_destroy(ctx, cat);
            return rc;
        }
        *out = cat;
        return DAMPR_OK;
    }
    rc = dampr_kv_sort_reduce(ctx, cat, key_xf, op, out);
    dampr_kv_destroy(ctx, cat);
    return rc;
}

}  // extern "C"
