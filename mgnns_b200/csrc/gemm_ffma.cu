// fp32 CUDA-core GEMM with fused bias/activation epilogue, strided batch and
// batch-reduce (split-K over consecutive (A,B) pairs).  This is the exact-fp32
// contraction used for every small dense op on the path (projections, FFN,
// classifier tail, label-GCN X*W) and the parity reference for the tcgen05 path.
//
// Tile: BM x BN x 16, 256 threads, (4*RM) x (4*RN) outputs per thread, operands
// staged through shared memory (double buffered, register prefetch), 128-bit
// global loads whenever the contiguous dimension allows it.
#include "common.cuh"

namespace mgnns {

struct GemmParams {
    int M, N, K;
    const float* A; int64_t lda, strideA;
    const float* B; int64_t ldb, strideB;
    float* C; int64_t ldc, strideC;
    int reduce, accumulate;
    const float* bias; int act; float slope;
    int vecA, vecB, vecC;
    int ksplit;        // >1: grid.z splits K of a single (A,B) pair, partial tiles are atomically added into C
    int klen;          // K elements per split
    float* ws;         // ksplit>1 && ws: partial tiles go to ws[z][M][N] (deterministic two-pass split-K)
};

constexpr int BK = 16;
constexpr int PAD = 4;

// Load one BMN x BK operand tile into registers (as float4 pieces).
//   TRANS == false: operand stored [rows=mn][cols=k], k contiguous
//   TRANS == true : operand stored [rows=k][cols=mn], mn contiguous
template <int BMN, bool TRANS, int NLD>
__device__ __forceinline__ void load_tile(float4 (&r)[NLD], const float* __restrict__ P, int64_t ld,
                                          int mn0, int k0, int MN, int K, int vec, int tid) {
#pragma unroll
    for (int it = 0; it < NLD; ++it) {
        int i = tid + it * 256;
        int mn, k;
        if (TRANS) { k = i / (BMN / 4); mn = (i % (BMN / 4)) * 4; }
        else       { mn = i / (BK / 4); k = (i % (BK / 4)) * 4; }
        int gmn = mn0 + mn, gk = k0 + k;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (TRANS) {
            if (gk < K) {
                const float* src = P + (int64_t)gk * ld + gmn;
                if (vec && gmn + 3 < MN) v = ldg4(src);
                else {
                    if (gmn + 0 < MN) v.x = __ldg(src + 0);
                    if (gmn + 1 < MN) v.y = __ldg(src + 1);
                    if (gmn + 2 < MN) v.z = __ldg(src + 2);
                    if (gmn + 3 < MN) v.w = __ldg(src + 3);
                }
            }
        } else {
            if (gmn < MN) {
                const float* src = P + (int64_t)gmn * ld + gk;
                if (vec && gk + 3 < K) v = ldg4(src);
                else {
                    if (gk + 0 < K) v.x = __ldg(src + 0);
                    if (gk + 1 < K) v.y = __ldg(src + 1);
                    if (gk + 2 < K) v.z = __ldg(src + 2);
                    if (gk + 3 < K) v.w = __ldg(src + 3);
                }
            }
        }
        r[it] = v;
    }
}

template <int BMN, bool TRANS, int NLD>
__device__ __forceinline__ void store_tile(const float4 (&r)[NLD], float (*S)[BMN + PAD], int tid) {
#pragma unroll
    for (int it = 0; it < NLD; ++it) {
        int i = tid + it * 256;
        if (TRANS) {
            int k = i / (BMN / 4), mn = (i % (BMN / 4)) * 4;
            *reinterpret_cast<float4*>(&S[k][mn]) = r[it];
        } else {
            int mn = i / (BK / 4), k = (i % (BK / 4)) * 4;
            S[k + 0][mn] = r[it].x;
            S[k + 1][mn] = r[it].y;
            S[k + 2][mn] = r[it].z;
            S[k + 3][mn] = r[it].w;
        }
    }
}

template <int RM, int RN, bool TA, bool TB>
__global__ void __launch_bounds__(256) gemm_ffma_kernel(GemmParams p) {
    constexpr int BM = 64 * RM, BN = 64 * RN;
    constexpr int NLA = BM * BK / 4 / 256, NLB = BN * BK / 4 / 256;
    __shared__ __align__(16) float As[2][BK][BM + PAD];
    __shared__ __align__(16) float Bs[2][BK][BN + PAD];

    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int z = blockIdx.z;

    const int zsplit = p.ksplit > 1 ? z % p.ksplit : 0;       // K slice
    const int zpair = p.ksplit > 1 ? z / p.ksplit : z;          // (A,B) pair / output index
    const int kbase = p.ksplit > 1 ? zsplit * p.klen : 0;
    const int kend = p.ksplit > 1 ? min(p.K, kbase + p.klen) : p.K;
    const int ktiles = (kend - kbase + BK - 1) / BK;
    const int total = ktiles * p.reduce;

    // accumulators in pairs along n: one FFMA2 (two exact fp32 FMAs, scalar a x pair of b) per pair and k — the
    // product loop is issue bound, and this halves its FMA instructions (same bits as scalar FFMA)
    unsigned long long acc2[4 * RM][2 * RN];
#pragma unroll
    for (int i = 0; i < 4 * RM; ++i)
#pragma unroll
        for (int j = 0; j < 2 * RN; ++j) acc2[i][j] = 0ull;

    float4 ra[NLA], rb[NLB];
    auto fetch = [&](int t) {
        int r = t / ktiles, kt = t - r * ktiles;
        int64_t pair = p.ksplit > 1 ? zpair : (int64_t)z * p.reduce + r;
        // A is indexed by (m, k): TA means stored [k][m]
        load_tile<BM, TA, NLA>(ra, p.A + pair * p.strideA, p.lda, m0, kbase + kt * BK, p.M, kend, p.vecA, tid);
        // B is indexed by (k, n): stored [k][n] (n contiguous) unless TB
        load_tile<BN, !TB, NLB>(rb, p.B + pair * p.strideB, p.ldb, n0, kbase + kt * BK, p.N, kend, p.vecB, tid);
    };

    fetch(0);
    store_tile<BM, TA, NLA>(ra, As[0], tid);
    store_tile<BN, !TB, NLB>(rb, Bs[0], tid);
    __syncthreads();

    for (int t = 0; t < total; ++t) {
        const int cur = t & 1;
        if (t + 1 < total) fetch(t + 1);
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            float a[4 * RM];
            unsigned long long b2[2 * RN];
#pragma unroll
            for (int g = 0; g < RM; ++g) {
                float4 v = *reinterpret_cast<const float4*>(&As[cur][k][g * 64 + ty * 4]);
                a[g * 4 + 0] = v.x; a[g * 4 + 1] = v.y; a[g * 4 + 2] = v.z; a[g * 4 + 3] = v.w;
            }
#pragma unroll
            for (int g = 0; g < RN; ++g) {
                const ulonglong2 v = *reinterpret_cast<const ulonglong2*>(&Bs[cur][k][g * 64 + tx * 4]);
                b2[g * 2 + 0] = v.x; b2[g * 2 + 1] = v.y;
            }
#pragma unroll
            for (int i = 0; i < 4 * RM; ++i) {
                const unsigned long long aa = pack_f32x2(a[i], a[i]);
#pragma unroll
                for (int j = 0; j < 2 * RN; ++j) acc2[i][j] = fma_f32x2(aa, b2[j], acc2[i][j]);
            }
        }
        if (t + 1 < total) {
            store_tile<BM, TA, NLA>(ra, As[cur ^ 1], tid);
            store_tile<BN, !TB, NLB>(rb, Bs[cur ^ 1], tid);
        }
        __syncthreads();
    }

    float acc[4 * RM][4 * RN];
#pragma unroll
    for (int i = 0; i < 4 * RM; ++i)
#pragma unroll
        for (int j = 0; j < 2 * RN; ++j) unpack_f32x2(acc2[i][j], acc[i][2 * j], acc[i][2 * j + 1]);

    const bool to_ws = p.ksplit > 1 && p.ws != nullptr;
    float* Cz = to_ws ? p.ws + (int64_t)z * p.M * p.N : p.C + (int64_t)zpair * p.strideC;
    const int64_t ldc = to_ws ? p.N : p.ldc;
    const bool add_bias = !to_ws && p.bias != nullptr && (p.ksplit <= 1 || zsplit == 0);
    const int act = to_ws ? MGNNS_ACT_NONE : p.act;
    const int accumulate = to_ws ? 0 : p.accumulate;
    const int vecC = to_ws ? ((p.N & 3) == 0) : p.vecC;
#pragma unroll
    for (int gi = 0; gi < RM; ++gi)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            int row = m0 + gi * 64 + ty * 4 + i;
            if (row >= p.M) continue;
#pragma unroll
            for (int gj = 0; gj < RN; ++gj) {
                int col = n0 + gj * 64 + tx * 4;
                if (col >= p.N) continue;
                float v[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float x = acc[gi * 4 + i][gj * 4 + j];
                    if (add_bias && col + j < p.N) x += __ldg(p.bias + col + j);
                    v[j] = apply_act(x, act, p.slope);
                }
                float* dst = Cz + (int64_t)row * ldc + col;
                if (accumulate) {
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (col + j < p.N) atomicAdd(dst + j, v[j]);
                } else if (vecC && col + 3 < p.N) {
                    *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (col + j < p.N) dst[j] = v[j];
                }
            }
        }
}

template <int RM, int RN>
static int launch_gemm(const GemmParams& p, int transA, int transB, int nz, cudaStream_t st) {
    dim3 grid((p.N + 64 * RN - 1) / (64 * RN), (p.M + 64 * RM - 1) / (64 * RM), nz);
    if (!transA && !transB) gemm_ffma_kernel<RM, RN, false, false><<<grid, 256, 0, st>>>(p);
    else if (!transA && transB) gemm_ffma_kernel<RM, RN, false, true><<<grid, 256, 0, st>>>(p);
    else if (transA && !transB) gemm_ffma_kernel<RM, RN, true, false><<<grid, 256, 0, st>>>(p);
    else gemm_ffma_kernel<RM, RN, true, true><<<grid, 256, 0, st>>>(p);
    MG_LAUNCH_CHECK("gemm_ffma");
    return 0;
}

// second pass of the deterministic split-K: C = act(sum_z ws[z] + bias), fixed summation order
__global__ void splitk_reduce_kernel(const float* __restrict__ ws, int splits, int M, int N, float* __restrict__ C,
                                     int64_t ldc, const float* __restrict__ bias, int act, float slope) {
    const int64_t total = (int64_t)M * N;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int m = (int)(i / N), n = (int)(i - (int64_t)m * N);
        float v = 0.f;
        for (int z = 0; z < splits; ++z) v += ws[(int64_t)z * total + i];
        if (bias != nullptr) v += __ldg(bias + n);
        C[(int64_t)m * ldc + n] = apply_act(v, act, slope);
    }
}

__global__ void act_bwd_kernel(const float* __restrict__ y, const float* __restrict__ g,
                               float* __restrict__ out, int64_t n, int act, float slope) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        float yy = y[i], gg = g[i];
        float d = 1.f;
        if (act == MGNNS_ACT_RELU) d = yy > 0.f ? 1.f : 0.f;
        else if (act == MGNNS_ACT_LEAKY) d = yy > 0.f ? 1.f : slope;
        out[i] = gg * d;
    }
}

// out[n] += sum_m x[m,n]; block = 32 x 8, each block reduces a strip of rows.
__global__ void colsum_kernel(const float* __restrict__ x, int64_t M, int N, int64_t ld,
                              float* __restrict__ out, int rows_per_block) {
    __shared__ float part[8][33];
    int n = blockIdx.x * 32 + threadIdx.x;
    int64_t r0 = (int64_t)blockIdx.y * rows_per_block;
    int64_t r1 = r0 + rows_per_block;
    if (r1 > M) r1 = M;
    float s = 0.f;
    if (n < N)
        for (int64_t r = r0 + threadIdx.y; r < r1; r += 8) s += x[r * ld + n];
    part[threadIdx.y][threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.y == 0 && n < N) {
        float t = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) t += part[i][threadIdx.x];
        atomicAdd(out + n, t);
    }
}

}  // namespace mgnns

using namespace mgnns;

static int gemm_impl(int transA, int transB, int M, int N, int K,
                     const float* A, int64_t lda, int64_t strideA,
                     const float* B, int64_t ldb, int64_t strideB,
                     float* C, int64_t ldc, int64_t strideC,
                     int batch, int reduce, int accumulate,
                     const float* bias, int act, float slope, float* workspace, int64_t workspace_floats,
                     void* stream);

extern "C" int mgnns_gemm_f32(int transA, int transB, int M, int N, int K,
                              const float* A, int64_t lda, int64_t strideA,
                              const float* B, int64_t ldb, int64_t strideB,
                              float* C, int64_t ldc, int64_t strideC,
                              int batch, int reduce, int accumulate,
                              const float* bias, int act, float slope, void* stream) {
    return gemm_impl(transA, transB, M, N, K, A, lda, strideA, B, ldb, strideB, C, ldc, strideC, batch, reduce,
                     accumulate, bias, act, slope, nullptr, 0, stream);
}

// Same contraction for a single (A,B) pair with a caller-provided workspace: small products are split
// along K into `workspace` partial tiles and summed in a fixed order (bitwise run-to-run deterministic).
// mgnns_gemm_splitk_workspace() returns the number of floats needed (0 = no split will be used).
extern "C" int64_t mgnns_gemm_splitk_workspace(int M, int N, int K) {
    int64_t ctas = (int64_t)((M + 63) / 64) * ((N + 63) / 64);
    if (ctas >= 148 || K < 128) return 0;
    int split = (int)((148 * 2 + ctas - 1) / ctas);
    if (split > K / 64) split = K / 64;
    if (split > 8) split = 8;
    return split > 1 ? (int64_t)split * M * N : 0;
}

extern "C" int mgnns_gemm_f32_ws(int transA, int transB, int M, int N, int K,
                                 const float* A, int64_t lda, const float* B, int64_t ldb, float* C, int64_t ldc,
                                 const float* bias, int act, float slope, float* workspace, int64_t workspace_floats,
                                 void* stream) {
    return gemm_impl(transA, transB, M, N, K, A, lda, 0, B, ldb, 0, C, ldc, 0, 1, 1, 0, bias, act, slope, workspace,
                     workspace_floats, stream);
}

static int gemm_impl(int transA, int transB, int M, int N, int K,
                     const float* A, int64_t lda, int64_t strideA,
                     const float* B, int64_t ldb, int64_t strideB,
                     float* C, int64_t ldc, int64_t strideC,
                     int batch, int reduce, int accumulate,
                     const float* bias, int act, float slope, float* workspace, int64_t workspace_floats,
                     void* stream) {
    MG_REQUIRE(M >= 0 && N >= 0 && K >= 0 && batch >= 0, "gemm: negative dimension");
    MG_REQUIRE(reduce >= 1 && (batch % reduce) == 0, "gemm: batch (%d) must be a multiple of reduce (%d)", batch, reduce);
    MG_REQUIRE(!(accumulate && (bias != nullptr || act != MGNNS_ACT_NONE)), "gemm: accumulate excludes bias/activation");
    if (M == 0 || N == 0 || batch == 0) return 0;
    MG_REQUIRE(A && B && C, "gemm: null operand");
    MG_REQUIRE(batch / reduce <= 65535, "gemm: too many output batches (%d)", batch / reduce);
    GemmParams p;
    p.M = M; p.N = N; p.K = K;
    p.A = A; p.lda = lda; p.strideA = strideA;
    p.B = B; p.ldb = ldb; p.strideB = strideB;
    p.C = C; p.ldc = ldc; p.strideC = strideC;
    p.reduce = reduce; p.accumulate = accumulate;
    p.bias = bias; p.act = act; p.slope = slope;
    p.vecA = aligned16(A) && (lda % 4 == 0) && (strideA % 4 == 0);
    p.vecB = aligned16(B) && (ldb % 4 == 0) && (strideB % 4 == 0);
    p.vecC = aligned16(C) && (ldc % 4 == 0) && (strideC % 4 == 0);
    int nz = batch / reduce;
    cudaStream_t st = as_stream(stream);
    p.ksplit = 1;
    p.klen = K;
    p.ws = nullptr;
    // deterministic two-pass split-K into the caller's workspace
    if (workspace != nullptr && batch == 1 && !accumulate) {
        const int64_t need = mgnns_gemm_splitk_workspace(M, N, K);
        if (need > 0 && need <= workspace_floats) {
            int split = (int)(need / ((int64_t)M * N));
            int klen = ((K + split - 1) / split + BK - 1) / BK * BK;
            split = (K + klen - 1) / klen;
            p.ksplit = split;
            p.klen = klen;
            p.ws = workspace;
            if (int rc = launch_gemm<1, 1>(p, transA, transB, split, st)) return rc;
            int64_t total = (int64_t)M * N;
            int blocks = (int)((total + 255) / 256);
            if (blocks > 148 * 8) blocks = 148 * 8;
            splitk_reduce_kernel<<<blocks, 256, 0, st>>>(workspace, split, M, N, C, ldc, bias, act, slope);
            MG_LAUNCH_CHECK("splitk_reduce");
            return 0;
        }
    }
    // big tiles only when they still fill the machine
    int64_t big_ctas = (int64_t)((M + 127) / 128) * ((N + 127) / 128) * nz;
    if (big_ctas >= 148 * 2) return launch_gemm<2, 2>(p, transA, transB, nz, st);
    // a single small product (M = batch rows or a weight gradient with K = batch): split K across
    // grid.z so that more than a handful of SMs work on it; partial tiles are summed with atomics
    int64_t ctas = (int64_t)((M + 63) / 64) * ((N + 63) / 64) * nz;
    // (only for the weight-gradient shape, transA: forward products stay bitwise run-to-run deterministic)
    if (transA && reduce == 1 && !accumulate && act == MGNNS_ACT_NONE && ctas < 148 && K >= 128) {
        int split = (int)((148 * 2 + ctas - 1) / ctas);
        int max_split = K / 64;
        if (split > max_split) split = max_split;
        if (split > 16) split = 16;
        if (split > 1) {
            int klen = ((K + split - 1) / split + BK - 1) / BK * BK;
            split = (K + klen - 1) / klen;
            for (int b = 0; b < batch; ++b) {
                cudaError_t e = cudaMemset2DAsync(C + (int64_t)b * strideC, (size_t)ldc * sizeof(float), 0,
                                                  (size_t)N * sizeof(float), (size_t)M, st);
                MG_REQUIRE(e == cudaSuccess, "gemm: memset failed: %s", cudaGetErrorString(e));
            }
            p.ksplit = split;
            p.klen = klen;
            p.accumulate = 1;
            nz = split * batch;
        }
    }
    return launch_gemm<1, 1>(p, transA, transB, nz, st);
}

extern "C" int mgnns_act_bwd_f32(const float* y, const float* g, float* out, int64_t n,
                                 int act, float slope, void* stream) {
    if (n == 0) return 0;
    MG_REQUIRE(y && g && out, "act_bwd: null pointer");
    int64_t blocks = (n + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    act_bwd_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(y, g, out, n, act, slope);
    MG_LAUNCH_CHECK("act_bwd");
    return 0;
}

extern "C" int mgnns_colsum_f32(const float* x, int64_t M, int N, int64_t ld, float* out, void* stream) {
    if (M == 0 || N == 0) return 0;
    MG_REQUIRE(x && out, "colsum: null pointer");
    int rows_per_block = 256;
    int64_t gy = (M + rows_per_block - 1) / rows_per_block;
    while (gy > 65535) { rows_per_block *= 2; gy = (M + rows_per_block - 1) / rows_per_block; }
    dim3 grid((N + 31) / 32, (unsigned)gy);
    colsum_kernel<<<grid, dim3(32, 8), 0, as_stream(stream)>>>(x, M, N, ld, out, rows_per_block);
    MG_LAUNCH_CHECK("colsum");
    return 0;
}
