// CSR SpMM (sum semiring) for the label-graph / word-graph convolution
// Y[b,i,:] = sum_e val[e] * X[b, col[e], :], plus dense->CSR conversion helpers.
//
// One CTA per (row tile, batch element); each thread owns one float4 of the
// feature dimension, so a warp reads 512 contiguous bytes of every neighbour
// row (128-bit loads, fully coalesced).  Neighbour indices/values of the row are
// staged once in shared memory and the edge loop is unrolled x4 so that four
// independent row reads are in flight per thread.  X_b stays L2-resident
// (12 MB at N=10k, F=300), so HBM traffic is X once + Y once.
#include "common.cuh"

namespace mgnns {

constexpr int SPMM_EDGE_CHUNK = 256;

template <bool VEC>
__global__ void __launch_bounds__(128) spmm_csr_kernel(
    int n_rows, const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col,
    const float* __restrict__ val, const float* __restrict__ X, int64_t ldx, int64_t strideX,
    float* __restrict__ Y, int64_t ldy, int64_t strideY, int F, int rows_per_cta) {
    __shared__ int32_t s_col[SPMM_EDGE_CHUNK];
    __shared__ float s_val[SPMM_EDGE_CHUNK];
    const int b = blockIdx.y;
    const float* Xb = X + (int64_t)b * strideX;
    float* Yb = Y + (int64_t)b * strideY;
    const int nvec = VEC ? (F >> 2) : F;
    const uint64_t keep = l2_policy_evict_last(), stream = l2_policy_evict_first();
    const int row0 = blockIdx.x * rows_per_cta;
    for (int r = row0; r < row0 + rows_per_cta && r < n_rows; ++r) {
        const int e0 = rowptr[r], e1 = rowptr[r + 1];
        for (int f0 = 0; f0 < nvec; f0 += blockDim.x) {
            const int f = f0 + threadIdx.x;
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int c0 = e0; c0 < e1; c0 += SPMM_EDGE_CHUNK) {
                const int cn = min(SPMM_EDGE_CHUNK, e1 - c0);
                __syncthreads();
                for (int i = threadIdx.x; i < cn; i += blockDim.x) {
                    s_col[i] = col[c0 + i];
                    s_val[i] = val[c0 + i];
                }
                __syncthreads();
                if (f < nvec) {
                    int e = 0;
                    if (VEC) {
                        for (; e + 4 <= cn; e += 4) {
                            float4 x0 = ldg4_l2(Xb + (int64_t)s_col[e + 0] * ldx + 4 * f, keep);
                            float4 x1 = ldg4_l2(Xb + (int64_t)s_col[e + 1] * ldx + 4 * f, keep);
                            float4 x2 = ldg4_l2(Xb + (int64_t)s_col[e + 2] * ldx + 4 * f, keep);
                            float4 x3 = ldg4_l2(Xb + (int64_t)s_col[e + 3] * ldx + 4 * f, keep);
                            float v0 = s_val[e], v1 = s_val[e + 1], v2 = s_val[e + 2], v3 = s_val[e + 3];
                            acc.x = fmaf(v0, x0.x, acc.x); acc.y = fmaf(v0, x0.y, acc.y);
                            acc.z = fmaf(v0, x0.z, acc.z); acc.w = fmaf(v0, x0.w, acc.w);
                            acc.x = fmaf(v1, x1.x, acc.x); acc.y = fmaf(v1, x1.y, acc.y);
                            acc.z = fmaf(v1, x1.z, acc.z); acc.w = fmaf(v1, x1.w, acc.w);
                            acc.x = fmaf(v2, x2.x, acc.x); acc.y = fmaf(v2, x2.y, acc.y);
                            acc.z = fmaf(v2, x2.z, acc.z); acc.w = fmaf(v2, x2.w, acc.w);
                            acc.x = fmaf(v3, x3.x, acc.x); acc.y = fmaf(v3, x3.y, acc.y);
                            acc.z = fmaf(v3, x3.z, acc.z); acc.w = fmaf(v3, x3.w, acc.w);
                        }
                        for (; e < cn; ++e) {
                            float4 x0 = ldg4_l2(Xb + (int64_t)s_col[e] * ldx + 4 * f, keep);
                            float v0 = s_val[e];
                            acc.x = fmaf(v0, x0.x, acc.x); acc.y = fmaf(v0, x0.y, acc.y);
                            acc.z = fmaf(v0, x0.z, acc.z); acc.w = fmaf(v0, x0.w, acc.w);
                        }
                    } else {
                        for (; e < cn; ++e)
                            acc.x = fmaf(s_val[e], __ldg(Xb + (int64_t)s_col[e] * ldx + f), acc.x);
                    }
                }
            }
            if (f < nvec) {
                if (VEC) stg4_l2(Yb + (int64_t)r * ldy + 4 * f, acc, stream);
                else Yb[(int64_t)r * ldy + f] = acc.x;
            }
        }
    }
}

// ---- dense -> CSR -----------------------------------------------------------
template <typename T, typename Pred>
__device__ __forceinline__ void row_count(const T* __restrict__ row, int n_cols, Pred keep, int32_t* out) {
    // one warp per row
    int lane = threadIdx.x & 31;
    int cnt = 0;
    for (int c = lane; c < n_cols; c += 32) cnt += keep(row[c]) ? 1 : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if (lane == 0) *out = cnt;
}

__global__ void dense_row_nnz_f32_kernel(const float* __restrict__ A, int n_rows, int n_cols, int64_t ld,
                                         int32_t* __restrict__ row_nnz) {
    int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= n_rows) return;
    row_count(A + (int64_t)r * ld, n_cols, [](float v) { return v != 0.f; }, row_nnz + r);
}

__global__ void count_row_nnz_i32_kernel(const int32_t* __restrict__ M, int n_rows, int n_cols, int min_count,
                                         int32_t* __restrict__ row_nnz) {
    int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= n_rows) return;
    row_count(M + (int64_t)r * n_cols, n_cols, [min_count](int32_t v) { return v >= min_count; }, row_nnz + r);
}

// ordered fill: warp per row, ballot-prefix keeps column order
template <typename T, typename Pred>
__device__ __forceinline__ void row_fill(const T* __restrict__ row, int n_cols, Pred keep, int base,
                                         int32_t* __restrict__ col, T* __restrict__ val) {
    int lane = threadIdx.x & 31;
    int off = base;
    for (int c0 = 0; c0 < n_cols; c0 += 32) {
        int c = c0 + lane;
        T v = (c < n_cols) ? row[c] : T(0);
        bool k = (c < n_cols) && keep(v);
        unsigned m = __ballot_sync(0xffffffffu, k);
        if (k) {
            int pos = off + __popc(m & ((1u << lane) - 1u));
            col[pos] = c;
            val[pos] = v;
        }
        off += __popc(m);
    }
}

__global__ void dense_fill_csr_f32_kernel(const float* __restrict__ A, int n_rows, int n_cols, int64_t ld,
                                          const int32_t* __restrict__ rowptr, int32_t* __restrict__ col,
                                          float* __restrict__ val) {
    int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= n_rows) return;
    row_fill(A + (int64_t)r * ld, n_cols, [](float v) { return v != 0.f; }, rowptr[r], col, val);
}

__global__ void count_fill_csr_i32_kernel(const int32_t* __restrict__ M, int n_rows, int n_cols, int min_count,
                                          const int32_t* __restrict__ rowptr, int32_t* __restrict__ col,
                                          int32_t* __restrict__ cnt) {
    int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= n_rows) return;
    row_fill(M + (int64_t)r * n_cols, n_cols, [min_count](int32_t v) { return v >= min_count; }, rowptr[r], col, cnt);
}

// single-CTA exclusive scan (n up to a few hundred thousand rows)
__global__ void __launch_bounds__(1024) exclusive_scan_i32_kernel(const int32_t* __restrict__ in,
                                                                  int32_t* __restrict__ out, int n) {
    __shared__ int32_t warp_tot[32];
    __shared__ int32_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int base = 0; base < n; base += 1024) {
        int i = base + threadIdx.x;
        int v = (i < n) ? in[i] : 0;
        int s = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, s, o);
            if (lane >= o) s += t;
        }
        if (lane == 31) warp_tot[w] = s;
        __syncthreads();
        if (w == 0) {
            int t = warp_tot[lane];
            int ts = t;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int q = __shfl_up_sync(0xffffffffu, ts, o);
                if (lane >= o) ts += q;
            }
            warp_tot[lane] = ts - t;  // exclusive prefix of warp totals
        }
        __syncthreads();
        int excl = carry + warp_tot[w] + (s - v);
        if (i < n) out[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) out[n] = carry;
}

}  // namespace mgnns

using namespace mgnns;

extern "C" int mgnns_spmm_csr_f32(int n_rows, const int32_t* rowptr, const int32_t* col, const float* val,
                                  const float* X, int64_t ldx, int64_t strideX,
                                  float* Y, int64_t ldy, int64_t strideY,
                                  int F, int batch, void* stream) {
    MG_REQUIRE(n_rows >= 0 && F >= 0 && batch >= 0, "spmm: negative dimension");
    if (n_rows == 0 || F == 0 || batch == 0) return 0;
    MG_REQUIRE(rowptr && X && Y, "spmm: null pointer");
    MG_REQUIRE(batch <= 65535, "spmm: batch too large (%d)", batch);
    const bool vec = (F % 4 == 0) && aligned16(X) && aligned16(Y) && (ldx % 4 == 0) && (ldy % 4 == 0) &&
                     (strideX % 4 == 0) && (strideY % 4 == 0);
    // enough CTAs to fill 148 SMs several times over, but few enough to amortise launch
    int rows_per_cta = 1;
    int64_t ctas = (int64_t)n_rows * batch;
    while (ctas / rows_per_cta > 148LL * 64 && rows_per_cta < 8) rows_per_cta *= 2;
    dim3 grid((n_rows + rows_per_cta - 1) / rows_per_cta, batch);
    cudaStream_t st = as_stream(stream);
    // one thread per float4 of the feature row: 75 float4 at F=300 -> 96 threads (78% of lanes busy instead of 59%;
    // measured 7.9 vs 8.5 ms on cfg 2).  Measured and rejected on the same shape: L1::no_allocate / L1::evict_last
    // load hints for cold / hot columns (11.5 / 9.4 ms) and a barrier-free warp-per-row variant (18 ms).
    int threads = 128;
    if (vec && F / 4 <= 96) threads = (F / 4 + 31) / 32 * 32;
    if (vec)
        spmm_csr_kernel<true><<<grid, threads, 0, st>>>(n_rows, rowptr, col, val, X, ldx, strideX, Y, ldy, strideY, F, rows_per_cta);
    else
        spmm_csr_kernel<false><<<grid, 128, 0, st>>>(n_rows, rowptr, col, val, X, ldx, strideX, Y, ldy, strideY, F, rows_per_cta);
    MG_LAUNCH_CHECK("spmm_csr");
    return 0;
}

extern "C" int mgnns_dense_row_nnz_f32(const float* A, int n_rows, int n_cols, int64_t ld, int32_t* row_nnz, void* stream) {
    if (n_rows == 0) return 0;
    MG_REQUIRE(A && row_nnz, "dense_row_nnz: null pointer");
    dense_row_nnz_f32_kernel<<<(n_rows + 7) / 8, 256, 0, as_stream(stream)>>>(A, n_rows, n_cols, ld, row_nnz);
    MG_LAUNCH_CHECK("dense_row_nnz");
    return 0;
}

extern "C" int mgnns_exclusive_scan_i32(const int32_t* in, int32_t* out, int n, void* stream) {
    MG_REQUIRE(in && out && n >= 0, "scan: bad argument");
    exclusive_scan_i32_kernel<<<1, 1024, 0, as_stream(stream)>>>(in, out, n);
    MG_LAUNCH_CHECK("exclusive_scan");
    return 0;
}

extern "C" int mgnns_dense_fill_csr_f32(const float* A, int n_rows, int n_cols, int64_t ld,
                                        const int32_t* rowptr, int32_t* col, float* val, void* stream) {
    if (n_rows == 0) return 0;
    MG_REQUIRE(A && rowptr, "dense_fill_csr: null pointer");
    dense_fill_csr_f32_kernel<<<(n_rows + 7) / 8, 256, 0, as_stream(stream)>>>(A, n_rows, n_cols, ld, rowptr, col, val);
    MG_LAUNCH_CHECK("dense_fill_csr");
    return 0;
}

extern "C" int mgnns_count_row_nnz_i32(const int32_t* M, int n_rows, int n_cols, int min_count,
                                       int32_t* row_nnz, void* stream) {
    if (n_rows == 0) return 0;
    MG_REQUIRE(M && row_nnz, "count_row_nnz: null pointer");
    count_row_nnz_i32_kernel<<<(n_rows + 7) / 8, 256, 0, as_stream(stream)>>>(M, n_rows, n_cols, min_count, row_nnz);
    MG_LAUNCH_CHECK("count_row_nnz");
    return 0;
}

extern "C" int mgnns_count_fill_csr_i32(const int32_t* M, int n_rows, int n_cols, int min_count,
                                        const int32_t* rowptr, int32_t* col, int32_t* cnt, void* stream) {
    if (n_rows == 0) return 0;
    MG_REQUIRE(M && rowptr, "count_fill_csr: null pointer");
    count_fill_csr_i32_kernel<<<(n_rows + 7) / 8, 256, 0, as_stream(stream)>>>(M, n_rows, n_cols, min_count, rowptr, col, cnt);
    MG_LAUNCH_CHECK("count_fill_csr");
    return 0;
}
