// Single-query multi-head attention core over a memory bank
// (ref: models/submodules.py:106-119 with the projections of :68-70 folded by
// the caller: score_h(l) = <W_k,h^T q'_h, bank_l> / sqrt(d_k), ctx_h = sum_l p_l bank_l).
//
// Forward: one CTA per sample, one warp per bank row; each row is read ONCE with
// 128-bit loads and feeds both the score dot products and the online-softmax
// weighted sum (flash-style running max / sum per head, merged across warps at
// the end), so HBM traffic is the bank once + O(H*D) per sample.  Rows masked
// out (padding) are never loaded.
// Backward: two sweeps over the rows of the same sample (the second one hits
// L2): sweep A computes s = <u,k> and t = <dctx,k>; the softmax backward is done
// on the [H,L] table in shared memory; sweep B emits dbank rows and accumulates du.
#include "common.cuh"

namespace mgnns {

constexpr int AT_THREADS = 256;
constexpr int AT_WARPS = AT_THREADS / 32;
constexpr int HG = 4;  // heads processed together (register budget)

__device__ __forceinline__ float dot4(const float4& a, const float4& b) {
    return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w;
}
__device__ __forceinline__ void fma4(float4& acc, float s, const float4& v) {
    acc.x = fmaf(s, v.x, acc.x); acc.y = fmaf(s, v.y, acc.y);
    acc.z = fmaf(s, v.z, acc.z); acc.w = fmaf(s, v.w, acc.w);
}
__device__ __forceinline__ void scale4(float4& acc, float s) { acc.x *= s; acc.y *= s; acc.z *= s; acc.w *= s; }

template <int NV>
__global__ void __launch_bounds__(AT_THREADS) attn_q1_fwd_kernel(
    const float* __restrict__ u, const float* __restrict__ bank, const float* __restrict__ mask,
    int B, int H, int L, int D, float scale, float p_drop, uint64_t seed, const uint64_t* seed_offset,
    float* __restrict__ ctx, float* __restrict__ attn, float* __restrict__ psum, float* __restrict__ lse) {
    extern __shared__ __align__(16) float sm[];
    float* us = sm;                          // [HG][D]
    float* sc = us + HG * D;                 // [HG][L]
    float* wm = sc + HG * L;                 // [AT_WARPS][HG]
    float* ws = wm + AT_WARPS * HG;          // [AT_WARPS][HG]
    float* fin = ws + AT_WARPS * HG;         // [HG] final max, [HG] final 1/sum, [AT_WARPS][HG] factors
    float* wacc = fin + 2 * HG + AT_WARPS * HG;  // [AT_WARPS][HG][D]  (16B aligned: all counts are multiples of 4)

    const int b = blockIdx.x;
    const int h0 = blockIdx.y * HG;
    const int nh = min(HG, H - h0);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int D4 = D >> 2;
    const float inv_keep = 1.f / (1.f - p_drop);
    if (seed_offset != nullptr) seed += *seed_offset;      // device-resident epoch (CUDA-graph replays)

    for (int i = threadIdx.x; i < HG * D; i += AT_THREADS)
        us[i] = (i < nh * D) ? u[((int64_t)b * H + h0) * D + i] : 0.f;
    __syncthreads();

    float m[HG], s[HG];
    float4 acc[HG][NV];
#pragma unroll
    for (int h = 0; h < HG; ++h) {
        m[h] = -INFINITY; s[h] = 0.f;
#pragma unroll
        for (int v = 0; v < NV; ++v) acc[h][v] = make_float4(0.f, 0.f, 0.f, 0.f);
    }

    const float* bk = bank + (int64_t)b * L * D;
    const float* mk = mask ? mask + (int64_t)b * L : nullptr;

    auto load_row = [&](int l, float4 (&k)[NV]) {
        const float* row = bk + (int64_t)l * D;
#pragma unroll
        for (int v = 0; v < NV; ++v) {
            int c = lane + 32 * v;
            k[v] = (c < D4) ? ldg4(row + 4 * c) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    };

    float4 kcur[NV], knext[NV];
    int l = warp;
    bool cur_live = (l < L) && !(mk && mk[l] == 0.f);
    if (cur_live) load_row(l, kcur);
    while (l < L) {
        const int ln = l + AT_WARPS;
        const bool next_live = (ln < L) && !(mk && mk[ln] == 0.f);
        if (next_live) load_row(ln, knext);
        if (cur_live) {
            float dot[HG];
#pragma unroll
            for (int h = 0; h < HG; ++h) {
                float d = 0.f;
#pragma unroll
                for (int v = 0; v < NV; ++v) {
                    int c = lane + 32 * v;
                    if (c < D4) d += dot4(kcur[v], *reinterpret_cast<const float4*>(us + h * D + 4 * c));
                }
                dot[h] = warp_sum(d);
            }
#pragma unroll
            for (int h = 0; h < HG; ++h) {
                if (h < nh) {
                    const float sv = dot[h] * scale;
                    if (lane == 0) sc[h * L + l] = sv;
                    bool keep = true;
                    if (p_drop > 0.f) keep = uniform01(seed, ((uint64_t)b * H + h0 + h) * L + l) >= p_drop;
                    const float mnew = fmaxf(m[h], sv);
                    const float corr = __expf(m[h] - mnew);
                    const float e = __expf(sv - mnew);
                    s[h] = s[h] * corr + e;
                    const float w = keep ? e : 0.f;
#pragma unroll
                    for (int v = 0; v < NV; ++v) {
                        scale4(acc[h][v], corr);
                        fma4(acc[h][v], w, kcur[v]);
                    }
                    m[h] = mnew;
                }
            }
        } else if (lane < nh) {
            sc[lane * L + l] = -INFINITY;
        }
#pragma unroll
        for (int v = 0; v < NV; ++v) kcur[v] = knext[v];
        cur_live = next_live;
        l = ln;
    }

    // ---- merge the per-warp partial softmax states --------------------------------
    if (lane == 0) {
#pragma unroll
        for (int h = 0; h < HG; ++h) { wm[warp * HG + h] = m[h]; ws[warp * HG + h] = s[h]; }
    }
#pragma unroll
    for (int h = 0; h < HG; ++h)
#pragma unroll
        for (int v = 0; v < NV; ++v) {
            int c = lane + 32 * v;
            if (c < D4) *reinterpret_cast<float4*>(wacc + ((size_t)warp * HG + h) * D + 4 * c) = acc[h][v];
        }
    __syncthreads();
    if (threadIdx.x < HG) {
        const int h = threadIdx.x;
        float M = -INFINITY;
        for (int w = 0; w < AT_WARPS; ++w) M = fmaxf(M, wm[w * HG + h]);
        float S = 0.f;
        for (int w = 0; w < AT_WARPS; ++w) {
            // a warp that saw no live row has m = -inf, s = 0: its factor is 0 unless
            // every row is masked (M = -inf), which yields NaN like the reference softmax
            float f = __expf(wm[w * HG + h] - M);
            fin[2 * HG + w * HG + h] = f;
            S += ws[w * HG + h] * f;
        }
        fin[h] = M;
        fin[HG + h] = 1.f / S;
        if (h < nh) lse[(int64_t)b * H + h0 + h] = M + __logf(S);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < nh * D; i += AT_THREADS) {
        const int h = i / D, d = i - h * D;
        float v = 0.f;
#pragma unroll
        for (int w = 0; w < AT_WARPS; ++w) v = fmaf(wacc[((size_t)w * HG + h) * D + d], fin[2 * HG + w * HG + h], v);
        ctx[((int64_t)b * H + h0 + h) * D + d] = v * fin[HG + h] * inv_keep;
    }
    if (warp < nh) {
        const int h = warp;
        const float M = fin[h], invS = fin[HG + h];
        float tot = 0.f;
        for (int ll = lane; ll < L; ll += 32) {
            float p = __expf(sc[h * L + ll] - M) * invS;
            bool keep = true;
            if (p_drop > 0.f) keep = uniform01(seed, ((uint64_t)b * H + h0 + h) * L + ll) >= p_drop;
            float pt = keep ? p * inv_keep : 0.f;
            attn[((int64_t)(h0 + h) * B + b) * L + ll] = pt;
            tot += pt;
        }
        tot = warp_sum(tot);
        if (lane == 0) psum[(int64_t)b * H + h0 + h] = tot;
    }
}

template <int NV>
__global__ void __launch_bounds__(AT_THREADS) attn_q1_bwd_kernel(
    const float* __restrict__ u, const float* __restrict__ bank, const float* __restrict__ mask,
    const float* __restrict__ lse, const float* __restrict__ gctx, const float* __restrict__ gpsum,
    int B, int H, int L, int D, float scale, float p_drop, uint64_t seed, const uint64_t* seed_offset,
    float* __restrict__ gu, float* __restrict__ gbank) {
    extern __shared__ __align__(16) float sm[];
    float* us = sm;                      // [HG][D]
    float* gs = us + HG * D;             // [HG][D]
    float* sc = gs + HG * D;             // [HG][L]  scores, then scale*ds
    float* tt = sc + HG * L;             // [HG][L]  <dctx,k>, then dropped-out probabilities
    float* wacc = tt + HG * L;           // [AT_WARPS][HG][D]; requires (2*HG*L) % 4 == 0 -> always

    const int b = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int D4 = D >> 2;
    const float inv_keep = 1.f / (1.f - p_drop);
    const float* bk = bank + (int64_t)b * L * D;
    const float* mk = mask ? mask + (int64_t)b * L : nullptr;
    float* gb = gbank + (int64_t)b * L * D;
    if (seed_offset != nullptr) seed += *seed_offset;

    for (int h0 = 0; h0 < H; h0 += HG) {
        const int nh = min(HG, H - h0);
        __syncthreads();
        for (int i = threadIdx.x; i < HG * D; i += AT_THREADS) {
            us[i] = (i < nh * D) ? u[((int64_t)b * H + h0) * D + i] : 0.f;
            gs[i] = (i < nh * D) ? gctx[((int64_t)b * H + h0) * D + i] : 0.f;
        }
        __syncthreads();

        // ---- sweep A: s = <u,k>, t = <dctx,k> ---------------------------------------
        for (int l = warp; l < L; l += AT_WARPS) {
            const bool live = !(mk && mk[l] == 0.f);
            if (!live) {
                if (lane < HG) { sc[lane * L + l] = -INFINITY; tt[lane * L + l] = 0.f; }
                continue;
            }
            float4 k[NV];
#pragma unroll
            for (int v = 0; v < NV; ++v) {
                int c = lane + 32 * v;
                k[v] = (c < D4) ? ldg4(bk + (int64_t)l * D + 4 * c) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int h = 0; h < HG; ++h) {
                float ds_ = 0.f, dt_ = 0.f;
#pragma unroll
                for (int v = 0; v < NV; ++v) {
                    int c = lane + 32 * v;
                    if (c < D4) {
                        ds_ += dot4(k[v], *reinterpret_cast<const float4*>(us + h * D + 4 * c));
                        dt_ += dot4(k[v], *reinterpret_cast<const float4*>(gs + h * D + 4 * c));
                    }
                }
                ds_ = warp_sum(ds_);
                dt_ = warp_sum(dt_);
                if (lane == 0) { sc[h * L + l] = ds_ * scale; tt[h * L + l] = dt_; }
            }
        }
        __syncthreads();

        // ---- softmax backward on the [nh, L] table ----------------------------------
        if (warp < nh) {
            const int h = warp;
            const float lse_h = lse[(int64_t)b * H + h0 + h];
            const float gp = gpsum ? gpsum[(int64_t)b * H + h0 + h] : 0.f;
            float delta = 0.f;
            for (int ll = lane; ll < L; ll += 32) {
                const float sv = sc[h * L + ll];
                const float p = (sv == -INFINITY) ? 0.f : __expf(sv - lse_h);
                bool keep = true;
                if (p_drop > 0.f) keep = uniform01(seed, ((uint64_t)b * H + h0 + h) * L + ll) >= p_drop;
                const float dp = keep ? (tt[h * L + ll] + gp) * inv_keep : 0.f;
                delta += p * dp;
            }
            delta = warp_sum(delta);
            for (int ll = lane; ll < L; ll += 32) {
                const float sv = sc[h * L + ll];
                const float p = (sv == -INFINITY) ? 0.f : __expf(sv - lse_h);
                bool keep = true;
                if (p_drop > 0.f) keep = uniform01(seed, ((uint64_t)b * H + h0 + h) * L + ll) >= p_drop;
                const float dp = keep ? (tt[h * L + ll] + gp) * inv_keep : 0.f;
                sc[h * L + ll] = p * (dp - delta) * scale;
                tt[h * L + ll] = keep ? p * inv_keep : 0.f;
            }
        } else if (warp < HG) {
            for (int ll = lane; ll < L; ll += 32) { sc[warp * L + ll] = 0.f; tt[warp * L + ll] = 0.f; }
        }
        __syncthreads();

        // ---- sweep B: dbank rows and du ---------------------------------------------
        float4 du[HG][NV];
#pragma unroll
        for (int h = 0; h < HG; ++h)
#pragma unroll
            for (int v = 0; v < NV; ++v) du[h][v] = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int l = warp; l < L; l += AT_WARPS) {
            const bool live = !(mk && mk[l] == 0.f);
            float4 k[NV], dk[NV];
#pragma unroll
            for (int v = 0; v < NV; ++v) {
                int c = lane + 32 * v;
                dk[v] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (c < D4) {
                    if (h0 > 0) dk[v] = *reinterpret_cast<const float4*>(gb + (int64_t)l * D + 4 * c);
                    if (live) k[v] = ldg4(bk + (int64_t)l * D + 4 * c);
                }
            }
            if (live) {
#pragma unroll
                for (int h = 0; h < HG; ++h) {
                    const float dsv = sc[h * L + l], pt = tt[h * L + l];
#pragma unroll
                    for (int v = 0; v < NV; ++v) {
                        int c = lane + 32 * v;
                        if (c < D4) {
                            fma4(du[h][v], dsv, k[v]);
                            fma4(dk[v], pt, *reinterpret_cast<const float4*>(gs + h * D + 4 * c));
                            fma4(dk[v], dsv, *reinterpret_cast<const float4*>(us + h * D + 4 * c));
                        }
                    }
                }
            }
#pragma unroll
            for (int v = 0; v < NV; ++v) {
                int c = lane + 32 * v;
                if (c < D4) *reinterpret_cast<float4*>(gb + (int64_t)l * D + 4 * c) = dk[v];
            }
        }
#pragma unroll
        for (int h = 0; h < HG; ++h)
#pragma unroll
            for (int v = 0; v < NV; ++v) {
                int c = lane + 32 * v;
                if (c < D4) *reinterpret_cast<float4*>(wacc + ((size_t)warp * HG + h) * D + 4 * c) = du[h][v];
            }
        __syncthreads();
        for (int i = threadIdx.x; i < nh * D; i += AT_THREADS) {
            const int h = i / D, d = i - h * D;
            float v = 0.f;
#pragma unroll
            for (int w = 0; w < AT_WARPS; ++w) v += wacc[((size_t)w * HG + h) * D + d];
            gu[((int64_t)b * H + h0 + h) * D + d] = v;
        }
    }
}

static size_t fwd_smem(int L, int D) {
    return sizeof(float) * ((size_t)HG * D + (size_t)HG * L + 2 * AT_WARPS * HG + 2 * HG + AT_WARPS * HG +
                            (size_t)AT_WARPS * HG * D);
}
static size_t bwd_smem(int L, int D) {
    return sizeof(float) * ((size_t)2 * HG * D + (size_t)2 * HG * L + (size_t)AT_WARPS * HG * D);
}

}  // namespace mgnns

using namespace mgnns;

#define AT_DISPATCH(NVVAL, KERNEL, SMEM, GRID, ...)                                                        \
    do {                                                                                                   \
        cudaFuncSetAttribute(KERNEL<NVVAL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(SMEM));     \
        KERNEL<NVVAL><<<GRID, AT_THREADS, SMEM, st>>>(__VA_ARGS__);                                        \
    } while (0)

extern "C" int mgnns_attn_q1_fwd(const float* u, const float* bank, const float* mask,
                                 int B, int H, int L, int D, float scale, float p_drop, uint64_t seed, const uint64_t* seed_offset,
                                 float* ctx, float* attn, float* psum, float* lse, void* stream) {
    MG_REQUIRE(B >= 0 && H >= 1 && L >= 1 && D >= 4, "attn_q1_fwd: bad dimensions");
    MG_REQUIRE(D % 4 == 0 && D <= 512, "attn_q1_fwd: D=%d must be a multiple of 4 and <= 512", D);
    MG_REQUIRE(p_drop >= 0.f && p_drop < 1.f, "attn_q1_fwd: p_drop must be in [0,1)");
    if (B == 0) return 0;
    MG_REQUIRE(u && bank && ctx && attn && psum && lse, "attn_q1_fwd: null pointer");
    MG_REQUIRE(aligned16(u) && aligned16(bank), "attn_q1_fwd: u/bank must be 16-byte aligned");
    // keep the float4 regions of shared memory 16B aligned: pad L to a multiple of 4 is not needed
    // because HG == 4 makes HG*L a multiple of 4.
    const size_t smem = fwd_smem(L, D);
    MG_REQUIRE(smem <= 220 * 1024, "attn_q1_fwd: L=%d too long for the shared-memory score table", L);
    cudaStream_t st = as_stream(stream);
    dim3 grid(B, (H + HG - 1) / HG);
    const int nv = (D / 4 + 31) / 32;
    switch (nv) {
        case 1: AT_DISPATCH(1, attn_q1_fwd_kernel, smem, grid, u, bank, mask, B, H, L, D, scale, p_drop, seed, seed_offset, ctx, attn, psum, lse); break;
        case 2: AT_DISPATCH(2, attn_q1_fwd_kernel, smem, grid, u, bank, mask, B, H, L, D, scale, p_drop, seed, seed_offset, ctx, attn, psum, lse); break;
        case 3: AT_DISPATCH(3, attn_q1_fwd_kernel, smem, grid, u, bank, mask, B, H, L, D, scale, p_drop, seed, seed_offset, ctx, attn, psum, lse); break;
        default: AT_DISPATCH(4, attn_q1_fwd_kernel, smem, grid, u, bank, mask, B, H, L, D, scale, p_drop, seed, seed_offset, ctx, attn, psum, lse); break;
    }
    MG_LAUNCH_CHECK("attn_q1_fwd");
    return 0;
}

extern "C" int mgnns_attn_q1_bwd(const float* u, const float* bank, const float* mask, const float* lse,
                                 const float* grad_ctx, const float* grad_psum,
                                 int B, int H, int L, int D, float scale, float p_drop, uint64_t seed, const uint64_t* seed_offset,
                                 float* grad_u, float* grad_bank, void* stream) {
    MG_REQUIRE(B >= 0 && H >= 1 && L >= 1 && D >= 4, "attn_q1_bwd: bad dimensions");
    MG_REQUIRE(D % 4 == 0 && D <= 512, "attn_q1_bwd: D=%d must be a multiple of 4 and <= 512", D);
    MG_REQUIRE(p_drop >= 0.f && p_drop < 1.f, "attn_q1_bwd: p_drop must be in [0,1)");
    if (B == 0) return 0;
    MG_REQUIRE(u && bank && lse && grad_ctx && grad_u && grad_bank, "attn_q1_bwd: null pointer");
    MG_REQUIRE(aligned16(u) && aligned16(bank) && aligned16(grad_ctx) && aligned16(grad_bank),
               "attn_q1_bwd: operands must be 16-byte aligned");
    const size_t smem = bwd_smem(L, D);
    MG_REQUIRE(smem <= 220 * 1024, "attn_q1_bwd: L=%d too long for the shared-memory score table", L);
    cudaStream_t st = as_stream(stream);
    dim3 grid(B);
    const int nv = (D / 4 + 31) / 32;
    switch (nv) {
        case 1: AT_DISPATCH(1, attn_q1_bwd_kernel, smem, grid, u, bank, mask, lse, grad_ctx, grad_psum, B, H, L, D, scale, p_drop, seed, seed_offset, grad_u, grad_bank); break;
        case 2: AT_DISPATCH(2, attn_q1_bwd_kernel, smem, grid, u, bank, mask, lse, grad_ctx, grad_psum, B, H, L, D, scale, p_drop, seed, seed_offset, grad_u, grad_bank); break;
        case 3: AT_DISPATCH(3, attn_q1_bwd_kernel, smem, grid, u, bank, mask, lse, grad_ctx, grad_psum, B, H, L, D, scale, p_drop, seed, seed_offset, grad_u, grad_bank); break;
        default: AT_DISPATCH(4, attn_q1_bwd_kernel, smem, grid, u, bank, mask, lse, grad_ctx, grad_psum, B, H, L, D, scale, p_drop, seed, seed_offset, grad_u, grad_bank); break;
    }
    MG_LAUNCH_CHECK("attn_q1_bwd");
    return 0;
}
