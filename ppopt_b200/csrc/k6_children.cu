// K6 - next-level generation with superset pruning, plus the ordered compaction / scan plumbing around it.
//
// Replaces generate_children_sets + CombinationTester.check
// (/root/reference/src/ppopt/mp_solvers/solver_utils.py:29-55,154-166, called from mpqp_combinatorial.py:60-61) and the
// mpLP cardinality filter (mpqp_combinatorial.py:40-42).  The reference tests every child against EVERY stored
// infeasible tuple (linear scan, the measured bottleneck at MPC N=10).  Here a child C = P + {i} of a feasible parent P
// survives iff every k-subset C \ {j}, j in P, is itself in this level's FEASIBLE list - equivalent to "no subset of C is
// in the murder list" by induction over levels; the feasible masks of the level go into an
// open-addressing hash set (load <= 0.5), so each test is ~1.5 probes instead of a 22-step binary search.
// Children are written parent by parent in ascending i, i.e. in the reference's order.
#include "common.cuh"
#include "launch.h"

namespace ppgpu {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_CHUNK = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ long long block_exclusive_scan(long long v, long long* total) {
    __shared__ long long wsum[SCAN_THREADS / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    long long x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const long long y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
    }
    if (lane == 31) wsum[warp] = x;
    __syncthreads();
    long long wpre = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < SCAN_THREADS / 32; ++w) {
        if (w < warp) wpre += wsum[w];
        tot += wsum[w];
    }
    __syncthreads();
    *total = tot;
    return wpre + x - v;
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_block_sums_kernel(const long long* __restrict__ in, long long n,
                                                                        long long* __restrict__ bsum) {
    const long long base = (long long)blockIdx.x * SCAN_CHUNK;
    long long s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        const long long i = base + (long long)k * SCAN_THREADS + threadIdx.x;
        if (i < n) s += in[i];
    }
    long long tot;
    block_exclusive_scan(s, &tot);
    if (threadIdx.x == 0) bsum[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_bsums_kernel(long long* __restrict__ bsum, long long nb) {
    __shared__ long long carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (long long base = 0; base < nb; base += SCAN_THREADS) {
        const long long i = base + threadIdx.x;
        const long long v = i < nb ? bsum[i] : 0;
        long long tot;
        const long long ex = block_exclusive_scan(v, &tot);
        const long long carry = carry_s;
        if (i < nb) bsum[i] = carry + ex;
        __syncthreads();
        if (threadIdx.x == 0) carry_s = carry + tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) bsum[nb] = carry_s;
}

// out[i] = exclusive prefix of in (in and out may alias); out[n] = total
__global__ void __launch_bounds__(SCAN_THREADS) scan_apply_kernel(const long long* in, long long n,
                                                                   const long long* __restrict__ bsum, long long nb,
                                                                   long long* out) {
    const long long base = (long long)blockIdx.x * SCAN_CHUNK + (long long)threadIdx.x * SCAN_ITEMS;
    long long v[SCAN_ITEMS];
    long long s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        const long long i = base + k;
        v[k] = i < n ? in[i] : 0;
        s += v[k];
    }
    long long tot;
    long long ex = block_exclusive_scan(s, &tot) + bsum[blockIdx.x];
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        const long long i = base + k;
        if (i < n) out[i] = ex;
        ex += v[k];
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) out[n] = bsum[nb];
}

// NOTE: scan_block_sums_kernel sums items strided by thread while scan_apply_kernel assigns items contiguously per
// thread; both cover exactly the chunk [blockIdx*SCAN_CHUNK, +SCAN_CHUNK), so the block totals agree.

size_t scan_workspace_bytes(long long n) {
    const long long nb = (n + SCAN_CHUNK - 1) / SCAN_CHUNK + 1;
    // offsets, block sums (+ total), one result slot, and (K6) an open-addressing table of up to 4n int32 slots
    const long long table = 2 * n + 2 > 1024 ? 2 * n + 2 : 1024;  // >= 2048 int32 slots even for tiny levels
    return (size_t)(n + 1 + nb + 2 + table) * sizeof(long long);
}

// exclusive scan of vals[0..n) in place -> vals[0..n], vals[n] = total; bsum = scratch of nb+1
static cudaError_t scan_inplace(long long* vals, long long n, long long* bsum, cudaStream_t st) {
    const long long nb = (n + SCAN_CHUNK - 1) / SCAN_CHUNK;
    if (n == 0) return cudaMemsetAsync(vals, 0, sizeof(long long), st);
    scan_block_sums_kernel<<<(unsigned)nb, SCAN_THREADS, 0, st>>>(vals, n, bsum);
    scan_bsums_kernel<<<1, SCAN_THREADS, 0, st>>>(bsum, nb);
    scan_apply_kernel<<<(unsigned)nb, SCAN_THREADS, 0, st>>>(vals, n, bsum, nb, vals);
    return cudaGetLastError();
}

__global__ void select_flag_kernel(const uint8_t* __restrict__ status, long long n, uint8_t bits, uint8_t value,
                                   long long* __restrict__ flag) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) flag[i] = ((status[i] & bits) == value) ? 1 : 0;
}

__global__ void select_scatter_kernel(const uint8_t* __restrict__ status, long long n, uint8_t bits, uint8_t value,
                                      const long long* __restrict__ offs, long long* __restrict__ idx_out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && ((status[i] & bits) == value)) idx_out[offs[i]] = i;
}

// ordered compaction: indices i with (status[i] & bits) == value, ascending; *d_count = how many
cudaError_t select_indices(const uint8_t* status, long long n, uint8_t bits, uint8_t value, long long* idx_out,
                           long long* d_count, void* ws, size_t ws_bytes, cudaStream_t st) {
    if (ws_bytes + sizeof(long long) < (size_t)(n + 1 + (n + SCAN_CHUNK - 1) / SCAN_CHUNK + 3) * sizeof(long long)) return cudaErrorInvalidValue;
    long long* offs = (long long*)ws;
    long long* bsum = offs + n + 1;
    if (n == 0) return cudaMemsetAsync(d_count, 0, sizeof(long long), st);
    const unsigned blocks = (unsigned)((n + 255) / 256);
    select_flag_kernel<<<blocks, 256, 0, st>>>(status, n, bits, value, offs);
    cudaError_t e = scan_inplace(offs, n, bsum, st);
    if (e != cudaSuccess) return e;
    select_scatter_kernel<<<blocks, 256, 0, st>>>(status, n, bits, value, offs, idx_out);
    return cudaMemcpyAsync(d_count, offs + n, sizeof(long long), cudaMemcpyDeviceToDevice, st);
}

__global__ void root_level_kernel(int W, long long count, uint64_t* __restrict__ masks) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) {
        for (int w = 0; w < W; ++w) masks[i * W + w] = 0ull;
        masks[i * W + (i >> 6)] = 1ull << (i & 63);
    }
}

// level-1 candidates [eq..., i]; mpLP keeps only i that can still reach |A| = n (mpqp_combinatorial.py:40-42)
cudaError_t root_level(const DevProgram& P, uint64_t* masks, long long* h_count, cudaStream_t st) {
    long long count = P.mi;
    if (!P.is_qp) {
        long long lim = (long long)1 + P.m - P.n;  // kept iff ne + i < (ne + 1) + m - n
        if (lim < 0) lim = 0;
        if (count > lim) count = lim;
    }
    *h_count = count;
    if (count == 0) return cudaSuccess;
    root_level_kernel<<<(unsigned)((count + 255) / 256), 256, 0, st>>>(P.W, count, masks);
    return cudaGetLastError();
}

constexpr int MAXW = 4;

__global__ void gather_masks_kernel(const uint64_t* __restrict__ masks, const long long* __restrict__ idx, long long nf, int W,
                                    uint64_t* __restrict__ out) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e < nf * W) {
        const long long p = e / W;
        const int w = (int)(e - p * W);
        out[e] = masks[idx[p] * W + w];
    }
}

struct Mask4 { uint64_t w[MAXW]; };

// lexicographic compare of a stored key with a register-resident mask (fully unrolled: no dynamic indexing)
__device__ __forceinline__ int lex_cmp4(const uint64_t* __restrict__ a, const Mask4& b, int W) {
    int res = 0;
#pragma unroll
    for (int x = 0; x < MAXW; ++x) {
        if (x < W && res == 0) {
            const uint64_t av = a[x];
            const uint64_t d = av ^ b.w[x];
            if (d) {
                const uint64_t low = d & (~d + 1ull);
                res = (av & low) ? -1 : 1;
            }
        }
    }
    return res;
}

__device__ __forceinline__ unsigned hash_mask(const Mask4& k, int W) {
    uint64_t h = 0x9e3779b97f4a7c15ull;
#pragma unroll
    for (int x = 0; x < MAXW; ++x)
        if (x < W) { h ^= k.w[x]; h *= 0xff51afd7ed558ccdull; h ^= h >> 32; }
    return (unsigned)h;
}

// open-addressing set of this level's feasible masks: table slots hold an index into `feas` (-1 = empty)
__global__ void hash_insert_kernel(const uint64_t* __restrict__ feas, long long nf, int W, int* __restrict__ table, unsigned cap_mask) {
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= nf) return;
    Mask4 k;
#pragma unroll
    for (int x = 0; x < MAXW; ++x) k.w[x] = x < W ? feas[p * W + x] : 0ull;
    unsigned slot = hash_mask(k, W) & cap_mask;
    while (atomicCAS(&table[slot], -1, (int)p) != -1) slot = (slot + 1) & cap_mask;
}

__device__ __forceinline__ bool hash_contains(const uint64_t* __restrict__ feas, int W, const int* __restrict__ table,
                                              unsigned cap_mask, const Mask4& key) {
    unsigned slot = hash_mask(key, W) & cap_mask;
    for (;;) {
        const int e = __ldg(table + slot);
        if (e < 0) return false;
        bool same = true;
#pragma unroll
        for (int x = 0; x < MAXW; ++x)
            if (x < W && feas[(long long)e * W + x] != key.w[x]) same = false;
        if (same) return true;
        slot = (slot + 1) & cap_mask;
    }
}

__global__ void __launch_bounds__(128)
children_count_kernel(DevProgram P, const uint64_t* __restrict__ feas, long long p_lo, long long p_hi, int k_act,
                      uint64_t* __restrict__ survive, long long* __restrict__ counts,
                      unsigned long long* __restrict__ counters, const int* __restrict__ table, unsigned cap_mask) {
    const int lane = threadIdx.x & 31;
    const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const int W = P.W, mi = P.mi;
    // mpLP: a list of length L whose last element is g is dropped iff g >= L + m - n
    const bool lp = !P.is_qp;
    const int child_limit = lp ? (P.ne + k_act + 1) + P.m - P.n : 0x7fffffff;   // global index bound for the child
    const int sub_limit = lp ? (P.ne + k_act) + P.m - P.n : 0x7fffffff;         // ... for its k-subsets
    unsigned long long lookups = 0;
    for (long long p = p_lo + warp0; p < p_hi; p += nwarps) {
        uint64_t pm[MAXW];
#pragma unroll
        for (int w = 0; w < MAXW; ++w) pm[w] = w < W ? feas[p * W + w] : 0ull;
        int last = -1;
#pragma unroll
        for (int w = MAXW - 1; w >= 0; --w)
            if (last < 0 && pm[w]) last = w * 64 + 63 - __clzll((long long)pm[w]);
        long long cnt = 0;
        uint64_t sv[MAXW];
#pragma unroll
        for (int w = 0; w < MAXW; ++w) sv[w] = 0ull;
        for (int base = ((last + 1) >> 5) << 5; base < mi; base += 32) {
            const int i = base + lane;
            bool ok = i > last && i < mi && (P.ne + i) < child_limit;
            if (ok) {
                const bool sub_filtered = (P.ne + i) >= sub_limit;
                if (!sub_filtered) {
                    Mask4 child;
#pragma unroll
                    for (int w = 0; w < MAXW; ++w) child.w[w] = pm[w] | ((w == (i >> 6)) ? (1ull << (i & 63)) : 0ull);
                    // every k-subset obtained by dropping one parent element must be feasible at this level
#pragma unroll
                    for (int w = 0; w < MAXW; ++w) {
                        if (w < W) {
                            uint64_t bits = pm[w];
                            while (bits && ok) {
                                const int b = __ffsll((long long)bits) - 1;
                                bits &= bits - 1;
                                Mask4 sub;
#pragma unroll
                                for (int x = 0; x < MAXW; ++x) sub.w[x] = child.w[x] & ~((x == w) ? (1ull << b) : 0ull);
                                ++lookups;
                                if (!hash_contains(feas, W, table, cap_mask, sub)) ok = false;
                            }
                        }
                    }
                }
            }
            const unsigned bal = __ballot_sync(0xffffffffu, ok);
            cnt += __popc(bal);
#pragma unroll
            for (int w = 0; w < MAXW; ++w)
                if (w == (base >> 6)) sv[w] |= (uint64_t)bal << (base & 63);
        }
        if (lane == 0) {
            counts[p] = cnt;
            for (int w = 0; w < W; ++w) survive[p * W + w] = sv[w];
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) lookups += __shfl_xor_sync(0xffffffffu, lookups, o);
    if (lane == 0 && lookups) atomicAdd(&counters[CNT_K6_LOOKUPS], lookups);
}

__global__ void __launch_bounds__(128)
children_write_kernel(int W, const uint64_t* __restrict__ feas, const uint64_t* __restrict__ survive,
                      const long long* __restrict__ offsets, long long nf, uint64_t* __restrict__ children) {
    const int lane = threadIdx.x & 31;
    const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long p = warp0; p < nf; p += nwarps) {
        long long off = offsets[p];
        for (int w = 0; w < W; ++w) {
            const uint64_t x = survive[p * W + w];
            const int tot = __popcll(x);
            for (int r = lane; r < tot; r += 32) {
                const int b = mask_nth(&x, 1, r);
                uint64_t* dst = children + (off + r) * W;
                for (int y = 0; y < W; ++y) dst[y] = feas[p * W + y] | (y == w ? (1ull << b) : 0ull);
            }
            off += tot;
        }
    }
}

// feas_masks (nf x W, sorted), survive (nf x W), offsets (nf + 1, exclusive scan of the per-parent child counts)
static unsigned hash_capacity(long long nf) {
    unsigned cap = 1024;
    while ((long long)cap < 2 * nf) cap <<= 1;          // load factor in (0.25, 0.5]
    return cap;
}

// gathers the feasible parents and builds the hash set of their masks in the workspace (behind the scan scratch)
cudaError_t children_prepare(const DevProgram& P, const uint64_t* masks, const long long* feas_idx, long long nf,
                             uint64_t* feas_masks, void* ws, size_t ws_bytes, cudaStream_t st) {
    if (P.W > MAXW) return cudaErrorInvalidValue;
    if (nf <= 0) return cudaSuccess;
    if (nf >= (1ll << 30)) return cudaErrorInvalidValue;
    const long long nb = (nf + SCAN_CHUNK - 1) / SCAN_CHUNK;
    const unsigned cap = hash_capacity(nf);
    if (ws_bytes < (size_t)(nb + 2) * sizeof(long long) + (size_t)cap * sizeof(int)) return cudaErrorInvalidValue;
    int* table = reinterpret_cast<int*>((long long*)ws + nb + 2);
    const long long ne = nf * P.W;
    gather_masks_kernel<<<(unsigned)((ne + 255) / 256), 256, 0, st>>>(masks, feas_idx, nf, P.W, feas_masks);
    cudaError_t e = cudaMemsetAsync(table, 0xff, (size_t)cap * sizeof(int), st);
    if (e != cudaSuccess) return e;
    hash_insert_kernel<<<(unsigned)((nf + 255) / 256), 256, 0, st>>>(feas_masks, nf, P.W, table, cap - 1);
    return cudaGetLastError();
}

// surviving children (bit sets) and their number for the parents [p_lo, p_hi); other entries are left untouched
cudaError_t children_count_range(const DevProgram& P, const uint64_t* feas_masks, long long nf, int k_act, uint64_t* survive,
                                 long long* counts, long long p_lo, long long p_hi, void* ws, unsigned long long* counters,
                                 cudaStream_t st) {
    if (p_hi <= p_lo) return cudaSuccess;
    const long long nb = (nf + SCAN_CHUNK - 1) / SCAN_CHUNK;
    const unsigned cap = hash_capacity(nf);
    const int* table = reinterpret_cast<const int*>((long long*)ws + nb + 2);
    long long blocks = ((p_hi - p_lo) * 32 + 127) / 128;
    if (blocks > 148 * 64) blocks = 148 * 64;
    children_count_kernel<<<(unsigned)blocks, 128, 0, st>>>(P, feas_masks, p_lo, p_hi, k_act, survive, counts, counters, table,
                                                            cap - 1);
    return cudaGetLastError();
}

cudaError_t children_scan(long long* counts_to_offsets, long long nf, void* ws, cudaStream_t st) {
    if (nf == 0) return cudaMemsetAsync(counts_to_offsets, 0, sizeof(long long), st);
    return scan_inplace(counts_to_offsets, nf, (long long*)ws, st);
}

cudaError_t children_count(const DevProgram& P, const uint64_t* masks, const long long* feas_idx, long long nf, int k_act,
                           uint64_t* feas_masks, uint64_t* survive, long long* offsets, void* ws, size_t ws_bytes,
                           unsigned long long* counters, cudaStream_t st) {
    if (nf == 0) return cudaMemsetAsync(offsets, 0, sizeof(long long), st);
    cudaError_t e = children_prepare(P, masks, feas_idx, nf, feas_masks, ws, ws_bytes, st);
    if (e != cudaSuccess) return e;
    if ((e = children_count_range(P, feas_masks, nf, k_act, survive, offsets, 0, nf, ws, counters, st)) != cudaSuccess) return e;
    return children_scan(offsets, nf, ws, st);
}

cudaError_t children_write(const DevProgram& P, const uint64_t* feas_masks, const uint64_t* survive, const long long* offsets,
                           long long nf, uint64_t* children, cudaStream_t st) {
    if (nf == 0) return cudaSuccess;
    long long blocks = (nf * 32 + 127) / 128;
    if (blocks > 148 * 64) blocks = 148 * 64;
    children_write_kernel<<<(unsigned)blocks, 128, 0, st>>>(P.W, feas_masks, survive, offsets, nf, children);
    return cudaGetLastError();
}

// ---- certificates inherited from the previous level ------------------------------------------------------------------------
// The vertex walk (k2w_walk.cu) leaves, for every candidate it certifies, the WITNESS: the mask of all rows active at the
// certifying vertex (~n' rows, the candidate's own among them), and next to it the mask of a later vertex of the walk that
// holds the candidate as well (PPG_WITNESS_SLOTS masks per candidate).  A candidate C of the next level is feasible as soon
// as a witness of ONE of its parents C \ {b} also has row b active: the same vertex, the same exact statement ("the rows of
// C are nonbasic at a vertex whose basic slacks are >= -1e-8").  On the 100x30x6 program 85 % of the candidates of levels
// 4-5 are certified this way (CPU model scripts/witness_coverage_model.py: 84 % / 86 %), which removes them - the ones
// that are easy to reach - from the walk.  Parents are found in the open-addressing hash set K6 built when it generated
// this level (children_prepare: table slots -> index into the parent level's feasible list); parent_wit is the witness
// array gathered in that order.  One thread per candidate; rows are dropped from the highest down (CPU model: the middle
// positions inherit most often, the first one least).  An inherited witness is passed on (witness_out), so that
// certificates propagate down the levels without any further pivot.
__device__ __forceinline__ int hash_find(const uint64_t* __restrict__ feas, int W, const int* __restrict__ table,
                                         unsigned cap_mask, const Mask4& key) {
    unsigned slot = hash_mask(key, W) & cap_mask;
    for (;;) {
        const int e = __ldg(table + slot);
        if (e < 0) return -1;
        bool same = true;
#pragma unroll
        for (int x = 0; x < MAXW; ++x)
            if (x < W && feas[(long long)e * W + x] != key.w[x]) same = false;
        if (same) return e;
        slot = (slot + 1) & cap_mask;
    }
}

__global__ void __launch_bounds__(256)
inherit_kernel(int W, const uint64_t* __restrict__ masks, long long n, uint8_t* __restrict__ status,
               uint64_t* __restrict__ witness_out, const uint64_t* __restrict__ pfeas, const int* __restrict__ table,
               unsigned cap_mask, const uint64_t* __restrict__ pwit, unsigned long long* __restrict__ counters) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long lookups = 0;
    int got = 0;
    if (i < n) {
        const uint8_t sb = status[i];
        if ((sb & PPG_ST_RANK) && !(sb & PPG_ST_FEAS)) {
            Mask4 m;
#pragma unroll
            for (int x = 0; x < MAXW; ++x) m.w[x] = x < W ? masks[i * W + x] : 0ull;
#pragma unroll
            for (int w = MAXW - 1; w >= 0; --w) {
                if (w < W && !got) {
                    uint64_t bits = m.w[w];
                    while (bits && !got) {
                        const int b = 63 - __clzll((long long)bits);
                        bits &= ~(1ull << b);
                        Mask4 sub = m;
#pragma unroll
                        for (int x = 0; x < MAXW; ++x)
                            if (x == w) sub.w[x] &= ~(1ull << b);
                        ++lookups;
                        const int e = hash_find(pfeas, W, table, cap_mask, sub);
                        if (e < 0) continue;   // (cannot happen for a child K6 let through: every parent is feasible)
                        // a parent's witnesses hold the parent (or are empty: certified by the relaxation / the simplex)
#pragma unroll
                        for (int sl = 0; sl < PPG_WITNESS_SLOTS; ++sl) {
                            if (got) continue;
                            bool inside = true;
                            uint64_t wv[MAXW];
#pragma unroll
                            for (int x = 0; x < MAXW; ++x) {
                                wv[x] = x < W ? __ldg(pwit + ((long long)e * PPG_WITNESS_SLOTS + sl) * W + x) : 0ull;
                                if (m.w[x] & ~wv[x]) inside = false;
                            }
                            if (inside) {
                                got = 1;
                                status[i] = sb | PPG_ST_FEAS;
                                if (witness_out)
#pragma unroll
                                    for (int x = 0; x < MAXW; ++x)
                                        if (x < W) witness_out[i * (PPG_WITNESS_SLOTS * W) + x] = wv[x];
                            }
                        }
                    }
                }
            }
        }
    }
    got = __reduce_add_sync(0xffffffffu, got);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) lookups += __shfl_xor_sync(0xffffffffu, lookups, o);
    if ((threadIdx.x & 31) == 0 && lookups) {
        atomicAdd(&counters[CNT_INHERITED], (unsigned long long)got);
        atomicAdd(&counters[CNT_INHERIT_LOOKUPS], lookups);
    }
}

cudaError_t launch_inherit(const DevProgram& P, const uint64_t* masks, long long n, uint8_t* status, uint64_t* witness_out,
                           const uint64_t* parent_feas, long long parent_nf, const void* parent_ws, const uint64_t* parent_wit,
                           unsigned long long* counters, cudaStream_t st) {
    if (n <= 0 || parent_nf <= 0) return cudaSuccess;
    if (P.W > MAXW) return cudaErrorInvalidValue;
    // the table sits in the K6 workspace behind the scan scratch (children_prepare)
    const long long nb = (parent_nf + SCAN_CHUNK - 1) / SCAN_CHUNK;
    const unsigned cap = hash_capacity(parent_nf);
    const int* table = reinterpret_cast<const int*>((const long long*)parent_ws + nb + 2);
    inherit_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(P.W, masks, n, status, witness_out, parent_feas, table, cap - 1,
                                                                parent_wit, counters);
    return cudaGetLastError();
}

// ---- FP64 FMA peak of the device (roofline denominator for the LP kernels; not in MEASURED_PEAKS.json)
__global__ void __launch_bounds__(256) dfma_peak_kernel(int iters, double* sink) {
    double a0 = threadIdx.x * 1e-3, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double m = 0.999999, c = 1e-9;
    for (int i = 0; i < iters; ++i) {
        a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
        a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
    }
    const double s = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    if (s == 123.456) sink[0] = s;
}

cudaError_t measure_fp64_peak(int iters, double* tflops, cudaStream_t st) {
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    double* sink = nullptr;
    cudaError_t e = cudaMalloc(&sink, sizeof(double));
    if (e != cudaSuccess) return e;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    const int blocks = sms * 8;
    dfma_peak_kernel<<<blocks, 256, 0, st>>>(iters / 10 + 1, sink);
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(a, st);
        dfma_peak_kernel<<<blocks, 256, 0, st>>>(iters, sink);
        cudaEventRecord(b, st);
        cudaEventSynchronize(b);
        float ms = 0;
        cudaEventElapsedTime(&ms, a, b);
        if (ms < best) best = ms;
    }
    cudaEventDestroy(a); cudaEventDestroy(b);
    cudaFree(sink);
    e = cudaGetLastError();
    const double flops = 2.0 * 8.0 * (double)iters * 256.0 * blocks;
    *tflops = flops / (best * 1e-3) / 1e12;
    return e;
}

}  // namespace ppgpu
