// PMI co-occurrence counts without a dense [V,V] table (ref: utils/pmi.py:37-66 — the counting loop and the
// min_cooccurence filter; utils/pmi.py:89-97 — the row-major enumeration of the kept cells).
//
// The reference accumulates pair_count[centre, target] over a window around every non-PAD token into a dense int64
// [V,V] matrix (3.25 GB at V=20k, 20 GB at V=50k) and then walks all V^2 cells three times.  Here the counts are
// built row by row from the emitted (centre -> target) pairs, whose number is what the corpus fixes
// (tokens x (2w-1)), not V^2:
//
//   1. pmi_row_emissions   one thread per token: n = number of in-window targets; row_emit[centre] += n,
//                          word_count[centre] += 1                                    (64-bit integer atomics)
//   2. exclusive scan      row_start[V+1] (int64)
//   3. pmi_scatter_targets one thread per token reserves its n slots with ONE atomic on the row cursor and writes
//                          its targets into the row's segment of a flat int32 buffer
//   4. pmi_row_reduce      one CTA per centre row: the row's targets are counted with shared-memory integer atomics
//                          into a column-indexed counter array (whole vocabulary in shared memory up to 56k words,
//                          column chunks beyond), then swept in column order — cells with count >= min_count are
//                          written in row-major order without any sort
//   5. pmi_compact         kept cells -> final CSR (rowptr from a scan of the per-row kept counts)
//
// Counts are exact integers; the order in which atomics land never shows in the result.  A cell cannot exceed its
// row's emission count, which the host checks against 2^31 before step 4 (cells are int32, like the CSR output).
#include "common.cuh"

namespace mgnns {

constexpr int PMI_REDUCE_THREADS = 1024;

__device__ __forceinline__ bool pmi_centre_ok(int c, int V, int pad_id, int row_lo, int row_hi) {
    return c >= 0 && c < V && c != pad_id && c >= row_lo && c < row_hi;
}

__global__ void __launch_bounds__(256) pmi_row_emissions_kernel(const int32_t* __restrict__ tok, int64_t D, int L, int V,
                                                                int window, int pad_id, int row_lo, int row_hi,
                                                                unsigned long long* __restrict__ row_emit,
                                                                unsigned long long* __restrict__ wc) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= D * L) return;
    const int64_t doc = idx / L;
    const int i = (int)(idx - doc * L);
    const int32_t* row = tok + doc * L;
    const int c = row[i];
    if (!pmi_centre_ok(c, V, pad_id, row_lo, row_hi)) return;
    const int j0 = max(0, i - window), j1 = min(L, i + window);
    int n = 0;
    for (int j = j0; j < j1; ++j) {
        const int t = row[j];
        n += (j != i && t >= 0 && t < V) ? 1 : 0;
    }
    atomicAdd(wc + c, 1ull);
    if (n) atomicAdd(row_emit + c, (unsigned long long)n);
}

__global__ void __launch_bounds__(256) pmi_scatter_targets_kernel(const int32_t* __restrict__ tok, int64_t D, int L, int V,
                                                                  int window, int pad_id, int row_lo, int row_hi,
                                                                  const int64_t* __restrict__ row_start,
                                                                  unsigned long long* __restrict__ cursor,
                                                                  int32_t* __restrict__ targets) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= D * L) return;
    const int64_t doc = idx / L;
    const int i = (int)(idx - doc * L);
    const int32_t* row = tok + doc * L;
    const int c = row[i];
    if (!pmi_centre_ok(c, V, pad_id, row_lo, row_hi)) return;
    const int j0 = max(0, i - window), j1 = min(L, i + window);
    int n = 0;
    for (int j = j0; j < j1; ++j) {
        const int t = row[j];
        n += (j != i && t >= 0 && t < V) ? 1 : 0;
    }
    if (!n) return;
    int64_t pos = row_start[c] + (int64_t)atomicAdd(cursor + c, (unsigned long long)n);
    for (int j = j0; j < j1; ++j) {
        const int t = row[j];
        if (j != i && t >= 0 && t < V) targets[pos++] = t;
    }
}

// block-wide exclusive scan of one int per thread (1024 threads); returns the exclusive prefix, *total = block sum
__device__ __forceinline__ int block_excl_scan_1024(int v, int* warp_tot, int* total) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int s = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, s, o);
        if (lane >= o) s += t;
    }
    if (lane == 31) warp_tot[w] = s;
    __syncthreads();
    if (w == 0) {
        const int t = warp_tot[lane];
        int ts = t;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int q = __shfl_up_sync(0xffffffffu, ts, o);
            if (lane >= o) ts += q;
        }
        warp_tot[lane] = ts - t;
        if (lane == 31) *total = ts;
    }
    __syncthreads();
    return warp_tot[w] + (s - v);
}

__global__ void __launch_bounds__(PMI_REDUCE_THREADS) pmi_row_reduce_kernel(
    const int32_t* __restrict__ targets, const int64_t* __restrict__ row_start, int V, int chunk, int min_count,
    int32_t* __restrict__ tmp_col, int32_t* __restrict__ tmp_cnt, int32_t* __restrict__ row_nnz) {
    extern __shared__ int s_cnt[];                   // [chunk] column counters
    __shared__ int warp_tot[32];
    __shared__ int s_total;
    const int tid = threadIdx.x;
    for (int r = blockIdx.x; r < V; r += gridDim.x) {
        const int64_t e0 = row_start[r], e1 = row_start[r + 1];
        if (e0 == e1) {
            if (tid == 0) row_nnz[r] = 0;
            continue;
        }
        int kept = 0;                                // identical in every thread
        for (int c0 = 0; c0 < V; c0 += chunk) {
            const int cw = min(chunk, V - c0);
            for (int i = tid; i < cw; i += PMI_REDUCE_THREADS) s_cnt[i] = 0;
            __syncthreads();
            // four independent loads in flight per thread: the hub rows (millions of pairs) are latency bound otherwise
            for (int64_t e = e0 + tid; e < e1; e += 4 * PMI_REDUCE_THREADS) {
                int tv[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int64_t ee = e + (int64_t)q * PMI_REDUCE_THREADS;
                    tv[q] = (ee < e1) ? __ldg(targets + ee) - c0 : -1;
                }
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    if (tv[q] >= 0 && tv[q] < cw) atomicAdd(&s_cnt[tv[q]], 1);
            }
            __syncthreads();
            // ordered sweep: every thread owns a contiguous span (odd length: conflict-free strided reads)
            const int per = ((cw + PMI_REDUCE_THREADS - 1) / PMI_REDUCE_THREADS) | 1;
            const int lo = min(cw, tid * per), hi = min(cw, lo + per);
            int k = 0;
            for (int i = lo; i < hi; ++i) k += (s_cnt[i] >= min_count) ? 1 : 0;
            const int off = block_excl_scan_1024(k, warp_tot, &s_total);
            int64_t pos = e0 + kept + off;
            for (int i = lo; i < hi; ++i) {
                const int v = s_cnt[i];
                if (v >= min_count) {
                    tmp_col[pos] = c0 + i;
                    tmp_cnt[pos] = v;
                    ++pos;
                }
            }
            kept += s_total;
            __syncthreads();                         // s_cnt / s_total are reused by the next chunk or row
        }
        if (tid == 0) row_nnz[r] = kept;
    }
}

__global__ void __launch_bounds__(256) pmi_compact_kernel(const int32_t* __restrict__ tmp_col, const int32_t* __restrict__ tmp_cnt,
                                                          const int64_t* __restrict__ row_start, const int32_t* __restrict__ rowptr,
                                                          int V, int32_t* __restrict__ col, int32_t* __restrict__ cnt) {
    // one warp per row
    const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= V) return;
    const int lane = threadIdx.x & 31;
    const int64_t src = row_start[r];
    const int dst = rowptr[r], n = rowptr[r + 1] - dst;
    for (int i = lane; i < n; i += 32) {
        col[dst + i] = tmp_col[src + i];
        cnt[dst + i] = tmp_cnt[src + i];
    }
}

// single-CTA exclusive scan, int64 (n up to a few hundred thousand rows)
__global__ void __launch_bounds__(1024) exclusive_scan_i64_kernel(const int64_t* __restrict__ in, int64_t* __restrict__ out, int n) {
    __shared__ long long warp_tot[32];
    __shared__ long long carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int base = 0; base < n; base += 1024) {
        const int i = base + threadIdx.x;
        const long long v = (i < n) ? in[i] : 0;
        long long s = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const long long t = __shfl_up_sync(0xffffffffu, s, o);
            if (lane >= o) s += t;
        }
        if (lane == 31) warp_tot[w] = s;
        __syncthreads();
        if (w == 0) {
            const long long t = warp_tot[lane];
            long long ts = t;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const long long q = __shfl_up_sync(0xffffffffu, ts, o);
                if (lane >= o) ts += q;
            }
            warp_tot[lane] = ts - t;
        }
        __syncthreads();
        const long long excl = carry + warp_tot[w] + (s - v);
        if (i < n) out[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) out[n] = carry;
}

}  // namespace mgnns

using namespace mgnns;

extern "C" int mgnns_pmi_row_emissions(const int32_t* tokens, int64_t D, int L, int V, int window, int pad_id,
                                       int row_lo, int row_hi, int64_t* row_emit, int64_t* word_count, void* stream) {
    MG_REQUIRE(D >= 0 && L >= 1 && V >= 1 && window >= 0, "pmi_row_emissions: bad dimensions");
    MG_REQUIRE(row_emit && word_count, "pmi_row_emissions: null pointer");
    cudaStream_t st = as_stream(stream);
    MG_REQUIRE(cudaMemsetAsync(row_emit, 0, sizeof(int64_t) * V, st) == cudaSuccess &&
               cudaMemsetAsync(word_count, 0, sizeof(int64_t) * V, st) == cudaSuccess, "pmi_row_emissions: memset failed");
    if (D == 0) return 0;
    MG_REQUIRE(tokens, "pmi_row_emissions: null pointer");
    const int64_t blocks = (D * L + 255) / 256;
    MG_REQUIRE(blocks < (1LL << 31), "pmi_row_emissions: corpus too large for one launch");
    pmi_row_emissions_kernel<<<(unsigned)blocks, 256, 0, st>>>(tokens, D, L, V, window, pad_id, row_lo, row_hi,
                                                             reinterpret_cast<unsigned long long*>(row_emit),
                                                             reinterpret_cast<unsigned long long*>(word_count));
    MG_LAUNCH_CHECK("pmi_row_emissions");
    return 0;
}

extern "C" int mgnns_exclusive_scan_i64(const int64_t* in, int64_t* out, int n, void* stream) {
    MG_REQUIRE(in && out && n >= 0, "scan_i64: bad argument");
    exclusive_scan_i64_kernel<<<1, 1024, 0, as_stream(stream)>>>(in, out, n);
    MG_LAUNCH_CHECK("exclusive_scan_i64");
    return 0;
}

extern "C" int mgnns_pmi_scatter_targets(const int32_t* tokens, int64_t D, int L, int V, int window, int pad_id,
                                         int row_lo, int row_hi, const int64_t* row_start, int64_t* cursor,
                                         int32_t* targets, void* stream) {
    MG_REQUIRE(D >= 0 && L >= 1 && V >= 1 && window >= 0, "pmi_scatter_targets: bad dimensions");
    MG_REQUIRE(row_start && cursor, "pmi_scatter_targets: null pointer");
    cudaStream_t st = as_stream(stream);
    MG_REQUIRE(cudaMemsetAsync(cursor, 0, sizeof(int64_t) * V, st) == cudaSuccess, "pmi_scatter_targets: memset failed");
    if (D == 0) return 0;
    MG_REQUIRE(tokens && targets, "pmi_scatter_targets: null pointer");
    const int64_t blocks = (D * L + 255) / 256;
    MG_REQUIRE(blocks < (1LL << 31), "pmi_scatter_targets: corpus too large for one launch");
    pmi_scatter_targets_kernel<<<(unsigned)blocks, 256, 0, st>>>(tokens, D, L, V, window, pad_id, row_lo, row_hi, row_start,
                                                               reinterpret_cast<unsigned long long*>(cursor), targets);
    MG_LAUNCH_CHECK("pmi_scatter_targets");
    return 0;
}

extern "C" int mgnns_pmi_row_reduce(const int32_t* targets, const int64_t* row_start, int V, int min_count,
                                    int32_t* tmp_col, int32_t* tmp_cnt, int32_t* row_nnz, void* stream) {
    MG_REQUIRE(V >= 1, "pmi_row_reduce: bad dimensions");
    MG_REQUIRE(targets && row_start && tmp_col && tmp_cnt && row_nnz, "pmi_row_reduce: null pointer");
    // the whole vocabulary in shared memory when it fits (56k words), else column chunks; small vocabularies leave
    // room for several CTAs per SM
    constexpr int MAX_CHUNK = 56 * 1024;
    const int chunk = V < MAX_CHUNK ? V : MAX_CHUNK;
    const size_t smem = sizeof(int) * (size_t)chunk;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(pmi_row_reduce_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)(sizeof(int) * MAX_CHUNK));
        MG_REQUIRE(e == cudaSuccess, "pmi_row_reduce: cannot reserve shared memory: %s", cudaGetErrorString(e));
        configured = true;
    }
    int per_sm = (int)((227 * 1024) / (smem + 1024));
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 2) per_sm = 2;                      // 1024-thread CTAs: at most two per SM
    int grid = 148 * per_sm;
    if (grid > V) grid = V;
    pmi_row_reduce_kernel<<<grid, PMI_REDUCE_THREADS, smem, as_stream(stream)>>>(targets, row_start, V, chunk, min_count,
                                                                                 tmp_col, tmp_cnt, row_nnz);
    MG_LAUNCH_CHECK("pmi_row_reduce");
    return 0;
}

extern "C" int mgnns_pmi_compact(const int32_t* tmp_col, const int32_t* tmp_cnt, const int64_t* row_start,
                                 const int32_t* rowptr, int V, int32_t* col, int32_t* cnt, void* stream) {
    MG_REQUIRE(V >= 1 && row_start && rowptr, "pmi_compact: bad argument");
    pmi_compact_kernel<<<(V + 7) / 8, 256, 0, as_stream(stream)>>>(tmp_col, tmp_cnt, row_start, rowptr, V, col, cnt);
    MG_LAUNCH_CHECK("pmi_compact");
    return 0;
}
