// Label-query element-wise attention, residual + custom LayerNorm, global
// spatial max, PMI co-occurrence counting, and the library-level bookkeeping.
#include "common.cuh"
#include <stdarg.h>
#include <atomic>

namespace mgnns {

static thread_local char g_error[512] = "";
static std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// ============================================================================
// Label attention (ref: models/Multi_GCN_Multihead_att.py:97-131).
// energy[b,c,h,d] = Q[c,h,d]*K[b,h,d]/sqrt(dh); softmax over d; (*) V[b,h,d].
// ============================================================================
constexpr int LA_MAXE = 4;  // elements per lane -> dh <= 128

__global__ void __launch_bounds__(256) label_attn_fwd_kernel(
    const float* __restrict__ Q, const float* __restrict__ K, const float* __restrict__ V, int64_t ldkv,
    int B, int C, int heads, int dh, float inv_scale, float p_drop, uint64_t seed, const uint64_t* seed_offset,
    float* __restrict__ out) {
    if (seed_offset != nullptr) seed += *seed_offset;
    const int lane = threadIdx.x & 31;
    const int64_t wid = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t total = (int64_t)B * C * heads;
    if (wid >= total) return;
    const int h = (int)(wid % heads);
    const int c = (int)((wid / heads) % C);
    const int b = (int)(wid / ((int64_t)heads * C));
    const int HD = heads * dh;
    const float inv_keep = 1.f / (1.f - p_drop);
    float e[LA_MAXE];
    float mx = -INFINITY;
#pragma unroll
    for (int i = 0; i < LA_MAXE; ++i) {
        int d = lane + 32 * i;
        e[i] = -INFINITY;
        if (d < dh) {
            e[i] = Q[(int64_t)c * HD + h * dh + d] * K[(int64_t)b * ldkv + h * dh + d] * inv_scale;
            mx = fmaxf(mx, e[i]);
        }
    }
    mx = warp_max(mx);
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < LA_MAXE; ++i) {
        int d = lane + 32 * i;
        e[i] = (d < dh) ? __expf(e[i] - mx) : 0.f;
        sum += e[i];
    }
    sum = warp_sum(sum);
    const float inv = 1.f / sum;
#pragma unroll
    for (int i = 0; i < LA_MAXE; ++i) {
        int d = lane + 32 * i;
        if (d < dh) {
            const int64_t oidx = ((int64_t)b * C + c) * HD + h * dh + d;
            float p = e[i] * inv;
            if (p_drop > 0.f) p = (uniform01(seed, (uint64_t)oidx) >= p_drop) ? p * inv_keep : 0.f;
            out[oidx] = p * V[(int64_t)b * ldkv + h * dh + d];
        }
    }
}

// One CTA per group of LA_BS samples; dQ is reduced in shared memory first.
constexpr int LA_BS = 4;
__global__ void __launch_bounds__(256) label_attn_bwd_kernel(
    const float* __restrict__ Q, const float* __restrict__ K, const float* __restrict__ V, int64_t ldkv,
    int B, int C, int heads, int dh, float inv_scale, float p_drop, uint64_t seed, const uint64_t* seed_offset,
    const float* __restrict__ gout, float* __restrict__ gQ, float* __restrict__ gK, float* __restrict__ gV,
    int64_t ldg) {
    extern __shared__ float sm[];
    if (seed_offset != nullptr) seed += *seed_offset;
    const int HD = heads * dh;
    float* sQ = sm;               // [C*HD]
    float* sK = sQ + C * HD;      // [HD]
    float* sV = sK + HD;          // [HD]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const float inv_keep = 1.f / (1.f - p_drop);
    for (int i = threadIdx.x; i < C * HD; i += blockDim.x) sQ[i] = 0.f;
    for (int bb = 0; bb < LA_BS; ++bb) {
        const int b = blockIdx.x * LA_BS + bb;
        if (b >= B) break;
        __syncthreads();
        for (int i = threadIdx.x; i < 2 * HD; i += blockDim.x) sK[i] = 0.f;  // sK and sV are contiguous
        __syncthreads();
        for (int pair = warp; pair < C * heads; pair += nw) {
            const int c = pair / heads, h = pair - c * heads;
            float e[LA_MAXE], kv[LA_MAXE], qv[LA_MAXE];
            float mx = -INFINITY;
#pragma unroll
            for (int i = 0; i < LA_MAXE; ++i) {
                int d = lane + 32 * i;
                e[i] = -INFINITY; kv[i] = 0.f; qv[i] = 0.f;
                if (d < dh) {
                    qv[i] = Q[(int64_t)c * HD + h * dh + d];
                    kv[i] = K[(int64_t)b * ldkv + h * dh + d];
                    e[i] = qv[i] * kv[i] * inv_scale;
                    mx = fmaxf(mx, e[i]);
                }
            }
            mx = warp_max(mx);
            float sum = 0.f;
#pragma unroll
            for (int i = 0; i < LA_MAXE; ++i) {
                int d = lane + 32 * i;
                e[i] = (d < dh) ? __expf(e[i] - mx) : 0.f;
                sum += e[i];
            }
            sum = warp_sum(sum);
            const float inv = 1.f / sum;
            float dp[LA_MAXE];
            float delta = 0.f;
#pragma unroll
            for (int i = 0; i < LA_MAXE; ++i) {
                int d = lane + 32 * i;
                dp[i] = 0.f;
                if (d < dh) {
                    const int64_t oidx = ((int64_t)b * C + c) * HD + h * dh + d;
                    const float p = e[i] * inv;
                    e[i] = p;
                    const float g = gout[oidx];
                    bool keep = true;
                    if (p_drop > 0.f) keep = uniform01(seed, (uint64_t)oidx) >= p_drop;
                    const float pt = keep ? p * inv_keep : 0.f;
                    atomicAdd(sV + h * dh + d, pt * g);
                    dp[i] = keep ? g * V[(int64_t)b * ldkv + h * dh + d] * inv_keep : 0.f;
                    delta += p * dp[i];
                }
            }
            delta = warp_sum(delta);
#pragma unroll
            for (int i = 0; i < LA_MAXE; ++i) {
                int d = lane + 32 * i;
                if (d < dh) {
                    const float de = e[i] * (dp[i] - delta) * inv_scale;
                    atomicAdd(sK + h * dh + d, de * qv[i]);
                    atomicAdd(sQ + c * HD + h * dh + d, de * kv[i]);
                }
            }
        }
        __syncthreads();
        for (int i = threadIdx.x; i < HD; i += blockDim.x) {
            gK[(int64_t)b * ldg + i] = sK[i];
            gV[(int64_t)b * ldg + i] = sV[i];
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < C * HD; i += blockDim.x) {
        float v = sQ[i];
        if (v != 0.f) atomicAdd(gQ + i, v);
    }
}

// ============================================================================
// Residual + custom LayerNorm (ref: models/submodules.py:142-156):
//   y = gamma * (z - mean) / (std_unbiased + eps) + beta,  z = x + res
// ============================================================================
__global__ void __launch_bounds__(256) add_layernorm_fwd_kernel(
    const float* __restrict__ x, const float* __restrict__ res, const float* __restrict__ gamma,
    const float* __restrict__ beta, int64_t rows, int D, float eps, float* __restrict__ y) {
    const int lane = threadIdx.x & 31;
    const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= rows) return;
    const float* xr = x + r * D;
    const float* rr = res ? res + r * D : nullptr;
    float s = 0.f;
    for (int d = lane; d < D; d += 32) s += xr[d] + (rr ? rr[d] : 0.f);
    const float mean = warp_sum(s) / (float)D;
    float q = 0.f;
    for (int d = lane; d < D; d += 32) {
        float z = xr[d] + (rr ? rr[d] : 0.f) - mean;
        q += z * z;
    }
    const float sd = sqrtf(warp_sum(q) / (float)(D - 1));
    const float rinv = 1.f / (sd + eps);
    for (int d = lane; d < D; d += 32) {
        float z = xr[d] + (rr ? rr[d] : 0.f);
        y[r * D + d] = gamma[d] * (z - mean) * rinv + beta[d];
    }
}

template <int NPL>
__global__ void __launch_bounds__(256) add_layernorm_bwd_kernel(
    const float* __restrict__ x, const float* __restrict__ res, const float* __restrict__ gamma,
    const float* __restrict__ gy, int64_t rows, int D, float eps,
    float* __restrict__ gz, float* __restrict__ ggamma, float* __restrict__ gbeta) {
    extern __shared__ float sm[];
    float* sg = sm;       // [D]
    float* sb = sm + D;   // [D]
    const int lane = threadIdx.x & 31;
    const int wpb = blockDim.x >> 5;
    for (int i = threadIdx.x; i < 2 * D; i += blockDim.x) sm[i] = 0.f;
    __syncthreads();
    float dg[NPL], db[NPL];
#pragma unroll
    for (int i = 0; i < NPL; ++i) { dg[i] = 0.f; db[i] = 0.f; }
    for (int64_t r = (int64_t)blockIdx.x * wpb + (threadIdx.x >> 5); r < rows; r += (int64_t)gridDim.x * wpb) {
        const float* xr = x + r * D;
        const float* rr = res ? res + r * D : nullptr;
        const float* gr = gy + r * D;
        float zc[NPL], a[NPL];
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < NPL; ++i) {
            int d = lane + 32 * i;
            zc[i] = (d < D) ? xr[d] + (rr ? rr[d] : 0.f) : 0.f;
            s += zc[i];
        }
        const float mean = warp_sum(s) / (float)D;
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < NPL; ++i) {
            int d = lane + 32 * i;
            zc[i] = (d < D) ? zc[i] - mean : 0.f;
            q += zc[i] * zc[i];
        }
        const float sd = sqrtf(warp_sum(q) / (float)(D - 1));
        const float rinv = 1.f / (sd + eps);
        float sa = 0.f, saz = 0.f;
#pragma unroll
        for (int i = 0; i < NPL; ++i) {
            int d = lane + 32 * i;
            float g = (d < D) ? gr[d] : 0.f;
            a[i] = (d < D) ? g * gamma[d] : 0.f;
            sa += a[i];
            saz += a[i] * zc[i];
            dg[i] += g * zc[i] * rinv;
            db[i] += g;
        }
        sa = warp_sum(sa) / (float)D;
        saz = warp_sum(saz);
        const float coef = rinv * rinv * saz / ((float)(D - 1) * sd);
#pragma unroll
        for (int i = 0; i < NPL; ++i) {
            int d = lane + 32 * i;
            if (d < D) gz[r * D + d] = rinv * (a[i] - sa) - coef * zc[i];
        }
    }
#pragma unroll
    for (int i = 0; i < NPL; ++i) {
        int d = lane + 32 * i;
        if (d < D) { atomicAdd(sg + d, dg[i]); atomicAdd(sb + d, db[i]); }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < D; i += blockDim.x) {
        atomicAdd(ggamma + i, sg[i]);
        atomicAdd(gbeta + i, sb[i]);
    }
}

// ============================================================================
// Global spatial max with first-index arg-max (ref: nn.MaxPool2d(14,14), model:302)
// ============================================================================
__global__ void __launch_bounds__(256) rowmax_kernel(const float* __restrict__ F, int64_t rows, int P,
                                                     float* __restrict__ pooled, int32_t* __restrict__ argmax) {
    const int lane = threadIdx.x & 31;
    const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= rows) return;
    const float* row = F + r * P;
    float best = -INFINITY;
    int bi = 0x7fffffff;
    for (int p0 = 0; p0 < P; p0 += 32 * 8) {
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            int p = p0 + lane + 32 * j;
            v[j] = (p < P) ? __ldg(row + p) : -INFINITY;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            int p = p0 + lane + 32 * j;
            // NaN propagates like torch's max pooling: a NaN beats everything
            if (p < P && (v[j] > best || (v[j] != v[j] && best == best) || bi == 0x7fffffff)) { best = v[j]; bi = p; }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        float ob = __shfl_xor_sync(0xffffffffu, best, o);
        int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        bool take;
        if (oi == 0x7fffffff) take = false;
        else if (bi == 0x7fffffff) take = true;
        else if (ob != ob && best == best) take = true;          // NaN wins
        else if (best != best && ob == ob) take = false;
        else take = (ob > best) || (ob == best && oi < bi) || (ob != ob && best != best && oi < bi);
        if (take) { best = ob; bi = oi; }
    }
    if (lane == 0) {
        pooled[r] = best;
        if (argmax) argmax[r] = bi;
    }
}

// Vectorised variant for P % 4 == 0 rows (14x14 = 196): one warp per row, NV4 128-bit loads per lane, the row
// maximum through one integer redux (order-preserving float->uint map), then the first position that attains it
// through a second redux.  Same tie-breaking (lowest index) and NaN rule (a NaN wins, first NaN's index) as above.
__device__ __forceinline__ uint32_t float_to_ordered(float f) {
    const uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ordered_to_float(uint32_t o) {
    return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o);
}

template <int NV4>
__global__ void __launch_bounds__(256) rowmax_vec_kernel(const float* __restrict__ F, int64_t rows, int P,
                                                         float* __restrict__ pooled, int32_t* __restrict__ argmax) {
    const int lane = threadIdx.x & 31;
    const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= rows) return;
    const float4* row = reinterpret_cast<const float4*>(F + r * P);
    const int n4 = P >> 2;
    float v[4 * NV4];
#pragma unroll
    for (int i = 0; i < NV4; ++i) {
        const int q = lane + 32 * i;
        float4 x = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
        if (q < n4) x = __ldcs(row + q);                 // streamed: the image-bank GEMM re-reads the map from HBM anyway
        v[4 * i + 0] = x.x; v[4 * i + 1] = x.y; v[4 * i + 2] = x.z; v[4 * i + 3] = x.w;
    }
    float m = -INFINITY;
    bool has_nan = false;
#pragma unroll
    for (int j = 0; j < 4 * NV4; ++j) {
        m = fmaxf(m, v[j]);                               // ignores NaN
        has_nan |= (v[j] != v[j]);
    }
    const float best = ordered_to_float(__reduce_max_sync(0xffffffffu, float_to_ordered(m)));
    const bool any_nan = __ballot_sync(0xffffffffu, has_nan) != 0u;
    uint32_t pos = 0xffffffffu;
#pragma unroll
    for (int i = NV4 - 1; i >= 0; --i)
#pragma unroll
        for (int e = 3; e >= 0; --e) {
            const float x = v[4 * i + e];
            const uint32_t p = 4u * (uint32_t)(lane + 32 * i) + (uint32_t)e;
            const bool hit = any_nan ? (x != x) : (x == best);
            if (hit && (lane + 32 * i) < n4) pos = p;    // descending scan: the lowest position is written last
        }
    pos = __reduce_min_sync(0xffffffffu, pos);
    if (lane == 0) {
        pooled[r] = any_nan ? __int_as_float(0x7fc00000) : best;
        if (argmax) argmax[r] = (int32_t)pos;
    }
}

__global__ void rowmax_bwd_kernel(const float* __restrict__ g, const int32_t* __restrict__ argmax,
                                  int64_t rows, int P, float* __restrict__ gF) {
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r < rows) gF[r * P + argmax[r]] += g[r];
}

// ============================================================================
// PMI co-occurrence counts (ref: utils/pmi.py:40-58).  One thread per token.
// ============================================================================
__global__ void __launch_bounds__(256) pmi_count_kernel(const int32_t* __restrict__ tok, int64_t D, int L, int V,
                                                        int window, int pad_id, int32_t* __restrict__ pair,
                                                        unsigned long long* __restrict__ wc) {
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= D * L) return;
    const int64_t doc = idx / L;
    const int i = (int)(idx - doc * L);
    const int32_t* row = tok + doc * L;
    const int c = row[i];
    if (c < 0 || c >= V || c == pad_id) return;
    atomicAdd(wc + c, 1ull);
    const int j0 = max(0, i - window), j1 = min(L, i + window);
    int32_t* prow = pair + (int64_t)c * V;
    for (int j = j0; j < j1; ++j) {
        if (j == i) continue;
        const int t = row[j];
        if (t >= 0 && t < V) atomicAdd(prow + t, 1);
    }
}

}  // namespace mgnns

using namespace mgnns;

extern "C" int mgnns_abi_version(void) { return MGNNS_ABI_VERSION; }
extern "C" const char* mgnns_last_error(void) { return g_error; }
extern "C" int64_t mgnns_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

extern "C" int mgnns_label_attn_fwd(const float* Q, const float* K, const float* V, int64_t ldkv,
                                    int B, int C, int heads, int dh, float inv_scale,
                                    float p_drop, uint64_t seed, const uint64_t* seed_offset, float* out, void* stream) {
    MG_REQUIRE(B >= 0 && C >= 1 && heads >= 1 && dh >= 1, "label_attn_fwd: bad dimensions");
    MG_REQUIRE(dh <= 32 * LA_MAXE, "label_attn_fwd: head dim %d > %d unsupported", dh, 32 * LA_MAXE);
    MG_REQUIRE(p_drop >= 0.f && p_drop < 1.f, "label_attn_fwd: p_drop must be in [0,1)");
    if (B == 0) return 0;
    MG_REQUIRE(Q && K && V && out, "label_attn_fwd: null pointer");
    int64_t warps = (int64_t)B * C * heads;
    int64_t blocks = (warps + 7) / 8;
    MG_REQUIRE(blocks < (1LL << 31), "label_attn_fwd: problem too large");
    label_attn_fwd_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(Q, K, V, ldkv, B, C, heads, dh, inv_scale,
                                                                           p_drop, seed, seed_offset, out);
    MG_LAUNCH_CHECK("label_attn_fwd");
    return 0;
}

extern "C" int mgnns_label_attn_bwd(const float* Q, const float* K, const float* V, int64_t ldkv,
                                    int B, int C, int heads, int dh, float inv_scale,
                                    float p_drop, uint64_t seed, const uint64_t* seed_offset, const float* grad_out,
                                    float* grad_Q, float* grad_K, float* grad_V, int64_t ldg, void* stream) {
    MG_REQUIRE(B >= 0 && C >= 1 && heads >= 1 && dh >= 1, "label_attn_bwd: bad dimensions");
    MG_REQUIRE(dh <= 32 * LA_MAXE, "label_attn_bwd: head dim %d > %d unsupported", dh, 32 * LA_MAXE);
    MG_REQUIRE(p_drop >= 0.f && p_drop < 1.f, "label_attn_bwd: p_drop must be in [0,1)");
    if (B == 0) return 0;
    MG_REQUIRE(Q && K && V && grad_out && grad_Q && grad_K && grad_V, "label_attn_bwd: null pointer");
    const int HD = heads * dh;
    size_t smem = sizeof(float) * ((size_t)C * HD + 2 * HD);
    MG_REQUIRE(smem <= 200 * 1024, "label_attn_bwd: C*heads*dh too large for shared memory");
    if (smem > 48 * 1024)
        cudaFuncSetAttribute(label_attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    label_attn_bwd_kernel<<<(B + LA_BS - 1) / LA_BS, 256, smem, as_stream(stream)>>>(
        Q, K, V, ldkv, B, C, heads, dh, inv_scale, p_drop, seed, seed_offset, grad_out, grad_Q, grad_K, grad_V, ldg);
    MG_LAUNCH_CHECK("label_attn_bwd");
    return 0;
}

extern "C" int mgnns_add_layernorm_fwd(const float* x, const float* res, const float* gamma, const float* beta,
                                       int64_t rows, int D, float eps, float* y, void* stream) {
    MG_REQUIRE(rows >= 0 && D >= 2, "add_layernorm_fwd: need D >= 2 (unbiased std)");
    if (rows == 0) return 0;
    MG_REQUIRE(x && gamma && beta && y, "add_layernorm_fwd: null pointer");
    int64_t blocks = (rows + 7) / 8;
    MG_REQUIRE(blocks < (1LL << 31), "add_layernorm_fwd: too many rows");
    add_layernorm_fwd_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(x, res, gamma, beta, rows, D, eps, y);
    MG_LAUNCH_CHECK("add_layernorm_fwd");
    return 0;
}

extern "C" int mgnns_add_layernorm_bwd(const float* x, const float* res, const float* gamma,
                                       const float* grad_y, int64_t rows, int D, float eps,
                                       float* grad_z, float* grad_gamma, float* grad_beta, void* stream) {
    MG_REQUIRE(rows >= 0 && D >= 2, "add_layernorm_bwd: need D >= 2 (unbiased std)");
    MG_REQUIRE(D <= 1024, "add_layernorm_bwd: D=%d > 1024 unsupported", D);
    if (rows == 0) return 0;
    MG_REQUIRE(x && gamma && grad_y && grad_z && grad_gamma && grad_beta, "add_layernorm_bwd: null pointer");
    int64_t blocks = (rows + 7) / 8;
    if (blocks > 148 * 4) blocks = 148 * 4;
    size_t smem = sizeof(float) * 2 * D;
    cudaStream_t st = as_stream(stream);
    const int npl = (D + 31) / 32;
    if (npl <= 4)
        add_layernorm_bwd_kernel<4><<<(unsigned)blocks, 256, smem, st>>>(x, res, gamma, grad_y, rows, D, eps, grad_z, grad_gamma, grad_beta);
    else if (npl <= 10)
        add_layernorm_bwd_kernel<10><<<(unsigned)blocks, 256, smem, st>>>(x, res, gamma, grad_y, rows, D, eps, grad_z, grad_gamma, grad_beta);
    else if (npl <= 16)
        add_layernorm_bwd_kernel<16><<<(unsigned)blocks, 256, smem, st>>>(x, res, gamma, grad_y, rows, D, eps, grad_z, grad_gamma, grad_beta);
    else
        add_layernorm_bwd_kernel<32><<<(unsigned)blocks, 256, smem, st>>>(x, res, gamma, grad_y, rows, D, eps, grad_z, grad_gamma, grad_beta);
    MG_LAUNCH_CHECK("add_layernorm_bwd");
    return 0;
}

extern "C" int mgnns_rowmax_f32(const float* F, int64_t rows, int P, float* pooled, int32_t* argmax, void* stream) {
    MG_REQUIRE(rows >= 0 && P >= 1, "rowmax: bad dimensions");
    if (rows == 0) return 0;
    MG_REQUIRE(F && pooled, "rowmax: null pointer");
    int64_t blocks = (rows + 7) / 8;
    MG_REQUIRE(blocks < (1LL << 31), "rowmax: too many rows");
    const int nv4 = (P / 4 + 31) / 32;
    if ((P & 3) == 0 && aligned16(F) && nv4 <= 4) {
        cudaStream_t st = as_stream(stream);
        if (nv4 == 1) rowmax_vec_kernel<1><<<(unsigned)blocks, 256, 0, st>>>(F, rows, P, pooled, argmax);
        else if (nv4 == 2) rowmax_vec_kernel<2><<<(unsigned)blocks, 256, 0, st>>>(F, rows, P, pooled, argmax);
        else if (nv4 == 3) rowmax_vec_kernel<3><<<(unsigned)blocks, 256, 0, st>>>(F, rows, P, pooled, argmax);
        else rowmax_vec_kernel<4><<<(unsigned)blocks, 256, 0, st>>>(F, rows, P, pooled, argmax);
    } else {
        rowmax_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(F, rows, P, pooled, argmax);
    }
    MG_LAUNCH_CHECK("rowmax");
    return 0;
}

extern "C" int mgnns_rowmax_bwd_f32(const float* grad_pooled, const int32_t* argmax, int64_t rows, int P,
                                    float* grad_F, void* stream) {
    if (rows == 0) return 0;
    MG_REQUIRE(grad_pooled && argmax && grad_F, "rowmax_bwd: null pointer");
    int64_t blocks = (rows + 255) / 256;
    rowmax_bwd_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(grad_pooled, argmax, rows, P, grad_F);
    MG_LAUNCH_CHECK("rowmax_bwd");
    return 0;
}

extern "C" int mgnns_pmi_count(const int32_t* tokens, int64_t D, int L, int V, int window, int pad_id,
                               int32_t* pair_count, int64_t* word_count, void* stream) {
    MG_REQUIRE(D >= 0 && L >= 1 && V >= 1 && window >= 0, "pmi_count: bad dimensions");
    if (D == 0) return 0;
    MG_REQUIRE(tokens && pair_count && word_count, "pmi_count: null pointer");
    int64_t blocks = (D * L + 255) / 256;
    MG_REQUIRE(blocks < (1LL << 31), "pmi_count: corpus too large for one launch");
    pmi_count_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(
        tokens, D, L, V, window, pad_id, pair_count, reinterpret_cast<unsigned long long*>(word_count));
    MG_LAUNCH_CHECK("pmi_count");
    return 0;
}
