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
constexpr int RING = 6; // ring slots per warp: two rows being consumed + four in flight (cp.async into a per-warp shared-memory ring)

// Each warp streams its rows (warp, warp+8, ...) through a private ring of RING row slots filled with
// 16-byte cp.async copies, so several rows (1.2 KB each at D=300) are in flight per warp without holding them in
// registers.
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src)
                 : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Sum eight per-lane partials over the warp and leave all eight totals in every lane.  Three halving steps
// (each lane sends the half of its values the partner keeps) bring it to one value per lane summed over eight
// lanes, two more butterfly steps finish the sum, eight indexed shuffles broadcast: 17 shuffles, 9 adds.
__device__ __forceinline__ void warp_reduce8(float (&v)[8], int lane) {
    const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4;
    float a[4], c[2], d;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float send = b4 ? v[i] : v[i + 4];
        const float keep = b4 ? v[i + 4] : v[i];
        a[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const float send = b3 ? a[i] : a[i + 2];
        const float keep = b3 ? a[i + 2] : a[i];
        c[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
    {
        const float send = b2 ? c[0] : c[1];
        const float keep = b2 ? c[1] : c[0];
        d = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
    d += __shfl_xor_sync(0xffffffffu, d, 2);
    d += __shfl_xor_sync(0xffffffffu, d, 1);
    // value i = 4*b4 + 2*b3 + b2 lives in the lanes whose bits (4,3,2) spell i
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __shfl_sync(0xffffffffu, d, ((i >> 2) & 1) * 16 + ((i >> 1) & 1) * 8 + (i & 1) * 4);
}

template <int NV>
__device__ __forceinline__ void ring_issue(float* slot, const float* row, int D4, int lane, bool live) {
    if (live) {
#pragma unroll
        for (int v = 0; v < NV; ++v) {
            const int c = lane + 32 * v;
            if (c < D4) cp_async16(slot + 4 * c, row + 4 * c);
        }
    }
    cp_async_commit();          // an empty group for a masked row keeps the group count uniform
}
template <int NV>
__device__ __forceinline__ void ring_read(const float* slot, int D4, int lane, float4 (&k)[NV]) {
#pragma unroll
    for (int v = 0; v < NV; ++v) {
        const int c = lane + 32 * v;
        k[v] = (c < D4) ? *reinterpret_cast<const float4*>(slot + 4 * c) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
}

__device__ __forceinline__ float dot4(const float4& a, const float4& b) {
    return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w;
}
__device__ __forceinline__ float dot4acc(float d, const float4& a, const float4& b) {
    d = fmaf(a.x, b.x, d); d = fmaf(a.y, b.y, d); d = fmaf(a.z, b.z, d);
    return fmaf(a.w, b.w, d);
}
__device__ __forceinline__ void fma4(float4& acc, float s, const float4& v) {
    acc.x = fmaf(s, v.x, acc.x); acc.y = fmaf(s, v.y, acc.y);
    acc.z = fmaf(s, v.z, acc.z); acc.w = fmaf(s, v.w, acc.w);
}
__device__ __forceinline__ void scale4(float4& acc, float s) { acc.x *= s; acc.y *= s; acc.z *= s; acc.w *= s; }

// MASKED / DROPOUT are compile-time: the image-bank launches (no mask) and eval-mode launches (no dropout) drop the
// live-row table, the keep table and every branch on them from the row loop, which is instruction-issue bound.
template <int NV, bool MASKED, bool DROPOUT>
__global__ void __launch_bounds__(AT_THREADS, 2) attn_q1_fwd_kernel(
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
    float* ring = wacc + (size_t)AT_WARPS * HG * D;   // [AT_WARPS][RING][D]
    float* kp = ring + (size_t)AT_WARPS * RING * D;    // [HG][L] dropout keep flags (1/0)

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
    const float* mk = MASKED ? mask + (int64_t)b * L : nullptr;

    float* myring = ring + (size_t)warp * RING * D;
    // live-row table in shared memory + the live bound Lb = 1 + last live row: padded text banks (mean 16 of
    // 100 rows) stop there instead of walking the mask to L
    unsigned char* lv = reinterpret_cast<unsigned char*>(kp + (size_t)HG * L);
    __shared__ int s_lb;
    if (threadIdx.x == 0) s_lb = 0;
    for (int i = threadIdx.x; i < HG * L; i += AT_THREADS) sc[i] = -INFINITY;
    __syncthreads();
    if (MASKED) {
        int last = 0;
        for (int l = threadIdx.x; l < L; l += AT_THREADS) {
            const bool on = mk[l] != 0.f;
            lv[l] = on ? 1 : 0;
            if (on) last = l + 1;
        }
        last = __reduce_max_sync(0xffffffffu, last);
        if (lane == 0 && last > 0) atomicMax(&s_lb, last);
        __syncthreads();
    }
    const int Lb = MASKED ? s_lb : L;
    auto row_live = [&](int l) { return (l < Lb) && (!MASKED || lv[l] != 0); };

    // dropout keep flags for this sample's [nh, L] probabilities, computed once (the counter-based generator is
    // ~25 integer instructions; per row and lane it used to be a quarter of the kernel's instruction stream)
    if (DROPOUT) {
        for (int i = threadIdx.x; i < nh * L; i += AT_THREADS) {
            const int h = i / L, l = i - h * L;
            kp[i] = uniform01(seed, ((uint64_t)b * H + h0 + h) * L + l) >= p_drop ? 1.f : 0.f;
        }
        __syncthreads();
    }

    // Two rows per iteration: the query vectors are read from shared memory once per pair, and the eight dot
    // products (2 rows x 4 heads) share one transposed butterfly reduction (17 shuffles instead of 40).
    // prologue: the first two pairs in flight
#pragma unroll
    for (int j = 0; j < RING - 2; ++j) {
        const int l = warp + j * AT_WARPS;
        ring_issue<NV>(myring + (j % RING) * D, bk + (int64_t)l * D, D4, lane, row_live(l));
    }
    int s0 = 0;                                          // ring slot of the pair's first row: 0, 2, 4, 0, ...
    const int full_v = D4 >> 5;                          // float4 columns that every lane owns
    for (int l = warp; l < Lb; l += 2 * AT_WARPS) {
        const int sp = (s0 >= 2) ? s0 - 2 : s0 + RING - 2;   // slot of the pair being prefetched (two pairs ahead)
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int lp = l + (RING - 2 + r) * AT_WARPS;
            ring_issue<NV>(myring + (sp + r) * D, bk + (int64_t)lp * D, D4, lane, row_live(lp));
        }
        cp_async_wait<RING - 2>();
        __syncwarp();
        const int lrow[2] = {l, l + AT_WARPS};
        const bool live[2] = {row_live(lrow[0]), row_live(lrow[1])};
        float4 k[2][NV];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            if (live[r]) ring_read<NV>(myring + (s0 + r) * D, D4, lane, k[r]);
            else {
#pragma unroll
                for (int v = 0; v < NV; ++v) k[r][v] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
        s0 = (s0 == RING - 2) ? 0 : s0 + 2;
        float dot[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) dot[i] = 0.f;
#pragma unroll
        for (int h = 0; h < HG; ++h)
#pragma unroll
            for (int v = 0; v < NV; ++v) {
                const int c = lane + 32 * v;
                if (v < full_v || c < D4) {
                    const float4 q = *reinterpret_cast<const float4*>(us + h * D + 4 * c);
                    dot[h] = dot4acc(dot[h], k[0][v], q);
                    dot[HG + h] = dot4acc(dot[HG + h], k[1][v], q);
                }
            }
        warp_reduce8(dot, lane);
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            if (live[r]) {
#pragma unroll
                for (int h = 0; h < HG; ++h) {
                    if (h < nh) {
                        const float sv = dot[r * HG + h] * scale;
                        if (lane == 0) sc[h * L + lrow[r]] = sv;
                        if (sv > m[h]) {                 // warp-uniform: rescale only when the running max moves
                            const float corr = __expf(m[h] - sv);
                            s[h] *= corr;
#pragma unroll
                            for (int v = 0; v < NV; ++v) scale4(acc[h][v], corr);
                            m[h] = sv;
                        }
                        const float e = __expf(sv - m[h]);
                        s[h] += e;
                        const float w = DROPOUT ? e * kp[h * L + lrow[r]] : e;
#pragma unroll
                        for (int v = 0; v < NV; ++v) fma4(acc[h][v], w, k[r][v]);
                    }
                }
            }                                        // masked rows keep the -inf the score table was initialised with
        }
        __syncwarp();                                // the slots are refilled by the next iteration's issue
    }
    cp_async_wait<0>();


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
            const bool keep = DROPOUT ? (kp[h * L + ll] != 0.f) : true;
            float pt = keep ? p * inv_keep : 0.f;
            attn[((int64_t)(h0 + h) * B + b) * L + ll] = pt;
            tot += pt;
        }
        tot = warp_sum(tot);
        if (lane == 0) psum[(int64_t)b * H + h0 + h] = tot;
    }
}

template <int NV>
__global__ void __launch_bounds__(AT_THREADS, 2) attn_q1_bwd_kernel(
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
    float* ring = wacc + (size_t)AT_WARPS * HG * D;   // [AT_WARPS][RING][D]

    const int b = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int D4 = D >> 2;
    const float inv_keep = 1.f / (1.f - p_drop);
    const float* bk = bank + (int64_t)b * L * D;
    const float* mk = mask ? mask + (int64_t)b * L : nullptr;
    float* gb = gbank + (int64_t)b * L * D;
    if (seed_offset != nullptr) seed += *seed_offset;
    float* myring = ring + (size_t)warp * RING * D;
    unsigned char* lv = reinterpret_cast<unsigned char*>(ring + (size_t)AT_WARPS * RING * D);
    __shared__ int s_lb;
    if (threadIdx.x == 0) s_lb = 0;
    __syncthreads();
    {
        int last = 0;
        for (int l = threadIdx.x; l < L; l += AT_THREADS) {
            const bool on = !(mk && mk[l] == 0.f);
            lv[l] = on ? 1 : 0;
            if (on) last = l + 1;
        }
        last = __reduce_max_sync(0xffffffffu, last);
        if (lane == 0 && last > 0) atomicMax(&s_lb, last);
    }
    __syncthreads();
    const int Lb = s_lb;                 // 1 + last live row
    auto row_live = [&](int l) { return (l < Lb) && lv[l] != 0; };

    for (int h0 = 0; h0 < H; h0 += HG) {
        const int nh = min(HG, H - h0);
        __syncthreads();
        for (int i = threadIdx.x; i < HG * D; i += AT_THREADS) {
            us[i] = (i < nh * D) ? u[((int64_t)b * H + h0) * D + i] : 0.f;
            gs[i] = (i < nh * D) ? gctx[((int64_t)b * H + h0) * D + i] : 0.f;
        }
        for (int i = threadIdx.x; i < HG * L; i += AT_THREADS) { sc[i] = -INFINITY; tt[i] = 0.f; }   // masked rows
        __syncthreads();

        // ---- sweep A: s = <u,k>, t = <dctx,k>; two rows per iteration --------------------
#pragma unroll
        for (int j = 0; j < RING - 2; ++j) {
            const int l = warp + j * AT_WARPS;
            ring_issue<NV>(myring + (j % RING) * D, bk + (int64_t)l * D, D4, lane, row_live(l));
        }
        int s0 = 0;                                      // ring slot of the pair's first row: 0, 2, 4, 0, ...
        const int full_v = D4 >> 5;
        for (int l = warp; l < Lb; l += 2 * AT_WARPS) {
            const int sp = (s0 >= 2) ? s0 - 2 : s0 + RING - 2;
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const int lp = l + (RING - 2 + r) * AT_WARPS;
                ring_issue<NV>(myring + (sp + r) * D, bk + (int64_t)lp * D, D4, lane, row_live(lp));
            }
            cp_async_wait<RING - 2>();
            __syncwarp();
            const int lrow[2] = {l, l + AT_WARPS};
            const bool live[2] = {row_live(lrow[0]), row_live(lrow[1])};
            float4 k[2][NV];
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                if (live[r]) ring_read<NV>(myring + (s0 + r) * D, D4, lane, k[r]);
                else {
#pragma unroll
                    for (int v = 0; v < NV; ++v) k[r][v] = make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
            s0 = (s0 == RING - 2) ? 0 : s0 + 2;
            float ds_[8], dt_[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) { ds_[i] = 0.f; dt_[i] = 0.f; }
#pragma unroll
            for (int h = 0; h < HG; ++h)
#pragma unroll
                for (int v = 0; v < NV; ++v) {
                    const int c = lane + 32 * v;
                    if (v < full_v || c < D4) {
                        const float4 q = *reinterpret_cast<const float4*>(us + h * D + 4 * c);
                        const float4 g = *reinterpret_cast<const float4*>(gs + h * D + 4 * c);
                        ds_[h] = dot4acc(ds_[h], k[0][v], q);
                        ds_[HG + h] = dot4acc(ds_[HG + h], k[1][v], q);
                        dt_[h] = dot4acc(dt_[h], k[0][v], g);
                        dt_[HG + h] = dot4acc(dt_[HG + h], k[1][v], g);
                    }
                }
            warp_reduce8(ds_, lane);
            warp_reduce8(dt_, lane);
            if (lane < HG) {
#pragma unroll
                for (int r = 0; r < 2; ++r) {
                    if (live[r]) {
                        // select this lane's head without dynamic register indexing
                        float sv = ds_[r * HG], tv = dt_[r * HG];
#pragma unroll
                        for (int h = 1; h < HG; ++h)
                            if (lane == h) { sv = ds_[r * HG + h]; tv = dt_[r * HG + h]; }
                        sc[lane * L + lrow[r]] = sv * scale;
                        tt[lane * L + lrow[r]] = tv;
                    }
                }
            }
            __syncwarp();
        }
        cp_async_wait<0>();
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

        // ---- sweep B: dbank rows and du (second read of the bank: mostly L2 hits).  One row per iteration: this
        // sweep is FMA-bound (36 FMAs per head and float4), pairing rows only costs registers.
        float4 du[HG][NV];
#pragma unroll
        for (int h = 0; h < HG; ++h)
#pragma unroll
            for (int v = 0; v < NV; ++v) du[h][v] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int j = 0; j < RING - 1; ++j) {
            const int l = warp + j * AT_WARPS;
            ring_issue<NV>(myring + (j % RING) * D, bk + (int64_t)l * D, D4, lane, row_live(l));
        }
        s0 = 0;                                          // ring slot of the current row
        for (int l = warp; l < L; l += AT_WARPS) {
            const int lp = l + (RING - 1) * AT_WARPS;
            ring_issue<NV>(myring + (s0 == 0 ? RING - 1 : s0 - 1) * D, bk + (int64_t)lp * D, D4, lane, row_live(lp));
            cp_async_wait<RING - 1>();
            __syncwarp();
            const bool live = row_live(l);
            float4 k[NV], dk[NV];
            if (live) ring_read<NV>(myring + s0 * D, D4, lane, k);
            s0 = (s0 == RING - 1) ? 0 : s0 + 1;
#pragma unroll
            for (int v = 0; v < NV; ++v) {
                const int c = lane + 32 * v;
                dk[v] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (c < D4 && h0 > 0) dk[v] = *reinterpret_cast<const float4*>(gb + (int64_t)l * D + 4 * c);
            }
            if (live) {
#pragma unroll
                for (int h = 0; h < HG; ++h) {
                    const float dsv = sc[h * L + l], pt = tt[h * L + l];
#pragma unroll
                    for (int v = 0; v < NV; ++v) {
                        const int c = lane + 32 * v;
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
                const int c = lane + 32 * v;
                if (c < D4) *reinterpret_cast<float4*>(gb + (int64_t)l * D + 4 * c) = dk[v];
            }
            __syncwarp();
        }
        cp_async_wait<0>();
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
                            (size_t)AT_WARPS * HG * D + (size_t)AT_WARPS * RING * D + (size_t)HG * L) + ((size_t)L + 15) / 16 * 16;
}
static size_t bwd_smem(int L, int D) {
    return sizeof(float) * ((size_t)2 * HG * D + (size_t)2 * HG * L + (size_t)AT_WARPS * HG * D + (size_t)AT_WARPS * RING * D) +
           ((size_t)L + 15) / 16 * 16;
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
#define AT_FWD(NVVAL, M, DR)                                                                                              \
    do {                                                                                                                   \
        cudaFuncSetAttribute(attn_q1_fwd_kernel<NVVAL, M, DR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);    \
        attn_q1_fwd_kernel<NVVAL, M, DR><<<grid, AT_THREADS, smem, st>>>(u, bank, mask, B, H, L, D, scale, p_drop, seed,   \
                                                                         seed_offset, ctx, attn, psum, lse);              \
    } while (0)
#define AT_FWD_FLAGS(NVVAL)                                    \
    do {                                                       \
        if (mask != nullptr && p_drop > 0.f) AT_FWD(NVVAL, true, true);       \
        else if (mask != nullptr) AT_FWD(NVVAL, true, false);  \
        else if (p_drop > 0.f) AT_FWD(NVVAL, false, true);     \
        else AT_FWD(NVVAL, false, false);                      \
    } while (0)
    switch (nv) {
        case 1: AT_FWD_FLAGS(1); break;
        case 2: AT_FWD_FLAGS(2); break;
        case 3: AT_FWD_FLAGS(3); break;
        default: AT_FWD_FLAGS(4); break;
    }
#undef AT_FWD_FLAGS
#undef AT_FWD
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
