// Single-query multi-head attention core on tensor-core fragments (the default path; attention.cu keeps the
// CUDA-core kernels for shapes outside this one's limits and as an independent cross-check).
// (ref: models/submodules.py:106-119 — attn = softmax(q.k^T / sqrt(d_k)), dropout, attn.v — with the projections
//  of :68-74 folded by the caller: score_h(l) = <u_h, bank_l> * scale, ctx_h = sum_l p_l bank_l.)
//
// The CUDA-core kernels were instruction-issue bound (ncu: 56 % issue-active, 23 % of DRAM): 96 FMAs per lane and
// bank row.  Here both products of a sample run as mma.sync.m16n8k8 TF32 fragments with the 3xTF32 split
// (hi = x & ~0x1fff, lo = x - hi; acc += hi*hi + lo*hi + hi*lo: fp32-class accuracy), ~2.6x fewer instructions:
//
//   scores  S[l,h]   = sum_d bank[l,d] u[h,d]       M = 16 bank rows, N = 8 heads, K = D in steps of 8
//   context C^T[d,h] = sum_l bank[l,d] p[l,h]       M = 16 features,  N = 8 heads, K = bank rows in steps of 8
//
// One CTA per sample (two resident per SM).  The sample's bank rows stream through shared memory in chunks of 32 rows,
// each fetched by ONE bulk asynchronous copy (cp.async.bulk + mbarrier: a chunk is 32*D*4 contiguous bytes), double
// buffered; the D = 300 row stride makes every fragment load bank-conflict free (300 = 12 mod 32).  Per chunk the
// eight warps (a) compute score partials — warp = (row tile, quarter of K) —, (b) run the online softmax, one warp
// per head, one lane per row, and (c) accumulate the context, each warp owning up to three 16-feature tiles whose
// accumulators are rescaled when the running maximum moves.  Padding rows of a masked (text) bank are never loaded:
// the loop stops at the last live row.
//
// Backward: sweep A computes S and T = <dctx, bank_l> with the same fragments (u and dctx stacked as one B operand),
// the softmax backward runs on the [H, L] tables in shared memory, and sweep B (second read of the bank, L2 hits)
// emits dbank = dS.u + P.dctx (M = bank rows, N = features, K = heads) and du^T = bank^T.dS.
#include "common.cuh"

namespace mgnns {

constexpr int TCA_R = 32;             // bank rows per chunk
constexpr int TCA_THREADS = 256;
constexpr int TCA_WARPS = 8;
constexpr int TCA_KSPLIT = 4;         // K quarters of the score product
constexpr int TCA_PAD = 16;           // floats after a chunk buffer (fragment loads may run past the last row)

__device__ __forceinline__ uint32_t tca_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tca_mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(tca_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void tca_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(tca_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tca_mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok) {
        asm volatile(
            "{\n\t"
            ".reg .pred P1;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, P1;\n\t"
            "}" : "=r"(ok) : "r"(tca_smem_u32(bar)), "r"(parity) : "memory");
    }
}
__device__ __forceinline__ void tca_bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(tca_smem_u32(dst)), "l"(src), "r"(bytes), "r"(tca_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tca_fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tca_split(float x, uint32_t& hi, uint32_t& lo) {
    hi = __float_as_uint(x) & 0xFFFFE000u;
    lo = __float_as_uint(x - __uint_as_float(hi));
}
__device__ __forceinline__ void tca_mma(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// c += a.b with both operands split (3xTF32)
__device__ __forceinline__ void tca_mma3(float (&c)[4], const float (&a)[4], float b0, float b1) {
    uint32_t ah[4], al[4], b0h, b0l, b1h, b1l;
#pragma unroll
    for (int i = 0; i < 4; ++i) tca_split(a[i], ah[i], al[i]);
    tca_split(b0, b0h, b0l);
    tca_split(b1, b1h, b1l);
    tca_mma(c, ah, b0h, b1h);
    tca_mma(c, al, b0h, b1h);
    tca_mma(c, ah, b0l, b1l);
}
// same with the A operand already split (it is reused across head tiles)
__device__ __forceinline__ void tca_mma3_pre(float (&c)[4], const uint32_t (&ah)[4], const uint32_t (&al)[4], float b0, float b1) {
    uint32_t b0h, b0l, b1h, b1l;
    tca_split(b0, b0h, b0l);
    tca_split(b1, b1h, b1l);
    tca_mma(c, ah, b0h, b1h);
    tca_mma(c, al, b0h, b1h);
    tca_mma(c, ah, b0l, b1l);
}

struct TcaDims {
    int B, H, L, D;
    int DU;        // row stride of the query tables (>= 8*KS, = 4 mod 8: conflict-free B fragments)
    int KS;        // K steps of the score product: ceil(D/8)
    int MT;        // 16-feature tiles: ceil(D/16)
    int LT;        // row stride of the [head, row] tables (>= L rounded up to a chunk, = 4 mod 8)
    int HR;        // head rows kept in shared memory: 4 when H <= 4 (fragment rows 4-7 read as zero), else 8 per head tile
};

// live-row table + live bound (1 + last live row) of a sample
__device__ __forceinline__ int tca_live_rows(const float* mk, int L, unsigned char* lv, int* s_lb) {
    if (threadIdx.x == 0) *s_lb = 0;
    __syncthreads();
    int last = 0;
    for (int l = threadIdx.x; l < L; l += TCA_THREADS) {
        const bool on = (mk == nullptr) || (mk[l] != 0.f);
        lv[l] = on ? 1 : 0;
        if (on) last = l + 1;
    }
    last = __reduce_max_sync(0xffffffffu, last);
    if ((threadIdx.x & 31) == 0 && last > 0) atomicMax(s_lb, last);
    __syncthreads();
    return *s_lb;
}

// score partials of one chunk: warp = (row tile mt of 16 rows, K quarter ks); NTT head tiles of 8 columns whose query
// rows come from q0 (tiles < NT) or q1 (tiles >= NT, backward only)
// PACK (H <= 4, one head tile per query table): the table's eight rows are [hi(q_0..3) | lo(q_0..3)], already exact
// TF32 splits, so one pair of MMAs (A_hi, A_lo) yields A.q_hi in columns 0-3 and A.q_lo in columns 4-7 — two MMAs
// and no B-side split instead of three MMAs and four split operations; the caller adds column h and h+4.
template <int NT, int NTT, bool PACK>
__device__ __forceinline__ void tca_score_partials(const float* __restrict__ cb, const float* __restrict__ q0,
                                                   const float* __restrict__ q1, float* __restrict__ part,
                                                   const TcaDims& dm, int warp, int lane) {
    const int g = lane >> 2, t = lane & 3;
    const int mt = warp & 1, ks = warp >> 1;
    const int ks4 = (dm.KS + TCA_KSPLIT - 1) / TCA_KSPLIT;
    const int kk0 = ks * ks4, kk1 = min(dm.KS, kk0 + ks4);
    // two accumulator sets for alternate K steps: halves the dependent-MMA chain (the kernel is latency bound)
    float c2[2][NTT][4];
#pragma unroll
    for (int q = 0; q < 2; ++q)
#pragma unroll
        for (int j = 0; j < NTT; ++j) { c2[q][j][0] = c2[q][j][1] = c2[q][j][2] = c2[q][j][3] = 0.f; }
    const float* ar0 = cb + (mt * 16 + g) * dm.D;
    const float* ar1 = ar0 + 8 * dm.D;
    auto step = [&](int kk, float (&c)[NTT][4]) {
        const int k0 = kk * 8 + t;
        float a[4];
        a[0] = ar0[k0];
        a[1] = ar1[k0];
        const bool tail = (k0 + 4 >= dm.D);            // only in the last K step when D % 8 == 4
        a[2] = tail ? 0.f : ar0[k0 + 4];
        a[3] = tail ? 0.f : ar1[k0 + 4];
        uint32_t ah[4], al[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) tca_split(a[i], ah[i], al[i]);
#pragma unroll
        for (int j = 0; j < NTT; ++j) {
            if (PACK) {
                const float* qp = (j < NT ? q0 : q1) + g * dm.DU + k0;
                const uint32_t b0 = __float_as_uint(qp[0]), b1 = __float_as_uint(qp[4]);
                tca_mma(c[j], ah, b0, b1);
                tca_mma(c[j], al, b0, b1);
            } else {
                const int hr = (j < NT ? j : j - NT) * 8 + g;
                const float* qp = (j < NT ? q0 : q1) + hr * dm.DU + k0;
                const bool on = hr < dm.HR;
                tca_mma3_pre(c[j], ah, al, on ? qp[0] : 0.f, on ? qp[4] : 0.f);
            }
        }
    };
    int kk = kk0;
    for (; kk + 2 <= kk1; kk += 2) {
        step(kk, c2[0]);
        step(kk + 1, c2[1]);
    }
    if (kk < kk1) step(kk, c2[0]);
    float c[NTT][4];
#pragma unroll
    for (int j = 0; j < NTT; ++j)
#pragma unroll
        for (int e = 0; e < 4; ++e) c[j][e] = c2[0][j][e] + c2[1][j][e];
    constexpr int NC = 8 * NTT;
    float* pp = part + (ks * TCA_R + mt * 16 + g) * NC + 2 * t;
#pragma unroll
    for (int j = 0; j < NTT; ++j) {
        *reinterpret_cast<float2*>(pp + j * 8) = make_float2(c[j][0], c[j][1]);
        *reinterpret_cast<float2*>(pp + 8 * NC + j * 8) = make_float2(c[j][2], c[j][3]);
    }
}

// acc^T[feature tile][head tile] += chunk^T . W for the warp's feature tiles; W[r][h] row-major with stride ws
// (forward: softmax weights of the chunk) or column-major tables T[h][row] with stride ws (backward: dS), COLMAJOR
template <int NT, bool COLMAJOR, bool PACK>
__device__ __forceinline__ void tca_accumulate_t(const float* __restrict__ cb, const float* __restrict__ w, int ws,
                                                 float (&acc)[3][NT][4], const TcaDims& dm, int warp, int lane) {
    const int g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const int mt = warp + TCA_WARPS * i;
        if (mt >= dm.MT) break;
#pragma unroll
        for (int kk = 0; kk < TCA_R / 8; ++kk) {
            const int r0 = kk * 8 + t;
            const float* p0 = cb + r0 * dm.D + mt * 16 + g;
            float a[4];
            a[0] = p0[0];
            a[1] = p0[8];
            a[2] = p0[4 * dm.D];
            a[3] = p0[4 * dm.D + 8];
            uint32_t ah[4], al[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) tca_split(a[k], ah[k], al[k]);
#pragma unroll
            for (int j = 0; j < NT; ++j) {
                float b0, b1;
                if (COLMAJOR) {
                    const float* wp = w + (j * 8 + g) * ws + r0;
                    const bool on = (j * 8 + g) < dm.HR;
                    b0 = on ? wp[0] : 0.f; b1 = on ? wp[4] : 0.f;
                } else {
                    const float* wp = w + r0 * ws + j * 8 + g;
                    b0 = wp[0]; b1 = wp[4 * ws];
                }
                if (PACK) {                              // columns are [hi | lo] of the weights: no split, two MMAs
                    tca_mma(acc[i][j], ah, __float_as_uint(b0), __float_as_uint(b1));
                    tca_mma(acc[i][j], al, __float_as_uint(b0), __float_as_uint(b1));
                } else {
                    tca_mma3_pre(acc[i][j], ah, al, b0, b1);
                }
            }
        }
    }
}

// =====================================================================================================================
// PACK: H <= 4 — query table and chunk weights carry [hi | lo] in the eight fragment columns (see tca_score_partials)
template <int NT, bool MASKED, bool DROPOUT, bool PACK>
__global__ void __launch_bounds__(TCA_THREADS, 2) attn_q1_tc_fwd_kernel(
    const float* __restrict__ u, const float* __restrict__ bank, const float* __restrict__ mask, TcaDims dm,
    float scale, float p_drop, uint64_t seed, const uint64_t* seed_offset,
    float* __restrict__ ctx, float* __restrict__ attn, float* __restrict__ psum, float* __restrict__ lse) {
    constexpr int NC = 8 * NT;
    extern __shared__ __align__(16) float sm[];
    const int CH = TCA_R * dm.D + TCA_PAD;
    float* chunk = sm;                                   // [2][CH]
    const int UR = PACK ? 8 : dm.HR;                     // rows of the query table
    float* us = chunk + 2 * CH;                          // [UR][DU]
    float* part = us + UR * dm.DU;                       // [4][R][NC]
    float* pw = part + TCA_KSPLIT * TCA_R * NC;          // [R][NC] softmax weights of the chunk
    float* sc = pw + TCA_R * NC;                         // [HR][LT] scaled scores (for the returned attention weights)
    float* stat = sc + dm.HR * dm.LT;                    // corr[NC], M[NC], S[NC]
    uint64_t* mbar = reinterpret_cast<uint64_t*>(stat + 3 * NC);   // 8-byte aligned: all counts above are even
    unsigned char* lv = reinterpret_cast<unsigned char*>(mbar + 2);
    __shared__ int s_lb;

    const int b = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int H = dm.H, L = dm.L, D = dm.D;
    const float inv_keep = 1.f / (1.f - p_drop);
    if (seed_offset != nullptr) seed += *seed_offset;
    const float* bk = bank + (int64_t)b * L * D;

    if (threadIdx.x == 0) {
        tca_mbar_init(&mbar[0], 1);
        tca_mbar_init(&mbar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = threadIdx.x; i < 2 * CH; i += TCA_THREADS) chunk[i] = 0.f;       // rows never loaded must be finite
    for (int i = threadIdx.x; i < UR * dm.DU; i += TCA_THREADS) {
        const int r = i / dm.DU, d = i - r * dm.DU;
        const int h = PACK ? (r & 3) : r;
        const float x = (h < H && d < D) ? u[((int64_t)b * H + h) * D + d] : 0.f;
        if (PACK) {
            uint32_t hi, lo;
            tca_split(x, hi, lo);
            us[i] = __uint_as_float(r < 4 ? hi : lo);
        } else {
            us[i] = x;
        }
    }
    for (int i = threadIdx.x; i < dm.HR * dm.LT; i += TCA_THREADS) sc[i] = -INFINITY;
    const int Lb = tca_live_rows(MASKED ? mask + (int64_t)b * L : nullptr, L, lv, &s_lb);   // has the barriers
    const int nchunks = (Lb + TCA_R - 1) / TCA_R;

    auto issue = [&](int c) {
        const int rows = min(TCA_R, Lb - c * TCA_R);
        const uint32_t bytes = (uint32_t)rows * D * 4u;
        tca_mbar_expect_tx(&mbar[c & 1], bytes);
        tca_bulk_load(chunk + (c & 1) * CH, bk + (int64_t)c * TCA_R * D, bytes, &mbar[c & 1]);
    };
    if (threadIdx.x == 0 && nchunks > 0) {
        tca_fence_proxy_async();                         // the zero fill above (generic proxy) before the bulk copy
        issue(0);
    }

    float m_run[NT], s_run[NT];
    float acc[3][NT][4];
#pragma unroll
    for (int j = 0; j < NT; ++j) { m_run[j] = -INFINITY; s_run[j] = 0.f; }
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < NT; ++j) acc[i][j][0] = acc[i][j][1] = acc[i][j][2] = acc[i][j][3] = 0.f;

    for (int c = 0; c < nchunks; ++c) {
        const float* cb = chunk + (c & 1) * CH;
        tca_mbar_wait(&mbar[c & 1], (c >> 1) & 1);
        tca_score_partials<NT, NT, PACK>(cb, us, us, part, dm, warp, lane);
        __syncthreads();                                 // partials complete; every warp is past the previous chunk
        if (threadIdx.x == 0 && c + 1 < nchunks) issue(c + 1);
        // ---- online softmax: warp = head (w, w+8), lane = row of the chunk
#pragma unroll
        for (int j = 0; j < NT; ++j) {
            const int hh = warp + 8 * j;
            if (PACK && hh >= 4) break;                  // columns 4-7 belong to heads 0-3 (their lo halves)
            const int l = c * TCA_R + lane;
            float s = 0.f;
#pragma unroll
            for (int q = 0; q < TCA_KSPLIT; ++q) {
                s += part[(q * TCA_R + lane) * NC + hh];
                if (PACK) s += part[(q * TCA_R + lane) * NC + hh + 4];
            }
            s *= scale;
            const bool live = (hh < H) && (l < Lb) && (!MASKED || lv[l] != 0);
            if (!live) s = -INFINITY;
            if (hh < H && l < L) sc[hh * dm.LT + l] = s;
            const float m_new = fmaxf(m_run[j], warp_max(s));
            const float e = live ? __expf(s - m_new) : 0.f;
            const float corr = (m_new == -INFINITY) ? 1.f : __expf(m_run[j] - m_new);
            s_run[j] = s_run[j] * corr + warp_sum(e);
            m_run[j] = m_new;
            float wgt = e;
            if (DROPOUT && live) wgt = (uniform01(seed, ((uint64_t)b * H + hh) * L + l) >= p_drop) ? e : 0.f;
            if (PACK) {
                uint32_t hi, lo;
                tca_split(wgt, hi, lo);
                pw[lane * NC + hh] = __uint_as_float(hi);
                pw[lane * NC + hh + 4] = __uint_as_float(lo);
                if (lane == 0) { stat[hh] = corr; stat[hh + 4] = corr; }
            } else {
                pw[lane * NC + hh] = wgt;
                if (lane == 0) stat[hh] = corr;
            }
        }
        __syncthreads();                                 // weights and correction factors visible
        // ---- context: rescale, then accumulate this chunk
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < NT; ++j) {
                const float c0 = stat[j * 8 + 2 * t], c1 = stat[j * 8 + 2 * t + 1];
                acc[i][j][0] *= c0; acc[i][j][1] *= c1; acc[i][j][2] *= c0; acc[i][j][3] *= c1;
            }
        tca_accumulate_t<NT, false, PACK>(cb, pw, NC, acc, dm, warp, lane);
    }
    __syncthreads();
    if (lane == 0) {
#pragma unroll
        for (int j = 0; j < NT; ++j) {
            const int hh = warp + 8 * j;
            if (PACK && hh >= 4) break;
            stat[NC + hh] = m_run[j];
            stat[2 * NC + hh] = s_run[j];
            if (hh < H) lse[(int64_t)b * H + hh] = m_run[j] + __logf(s_run[j]);
        }
    }
    __syncthreads();
    // ---- context out: C^T fragments -> ctx[b, h, d]
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const int mt = warp + TCA_WARPS * i;
        if (mt >= dm.MT) break;
#pragma unroll
        for (int j = 0; j < NT; ++j) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int d = mt * 16 + g + ((e & 2) ? 8 : 0);
                const int hh = j * 8 + 2 * t + (e & 1);
                float v = acc[i][j][e];
                if (PACK) v += __shfl_xor_sync(0xffffffffu, v, 2);      // column h+4 (the lo half) lives in lane t+2
                if (d < D && hh < (PACK ? min(H, 4) : H)) ctx[((int64_t)b * H + hh) * D + d] = v * inv_keep / stat[2 * NC + hh];
            }
        }
    }
    // ---- attention weights (after dropout, as the reference returns them) and their row sums
#pragma unroll
    for (int j = 0; j < NT; ++j) {
        const int hh = warp + 8 * j;
        if (hh >= H) continue;
        const float M = stat[NC + hh], invS = 1.f / stat[2 * NC + hh];
        float tot = 0.f;
        for (int l = lane; l < L; l += 32) {
            const float p = __expf(sc[hh * dm.LT + l] - M) * invS;
            bool keep = true;
            if (DROPOUT) keep = uniform01(seed, ((uint64_t)b * H + hh) * L + l) >= p_drop;
            const float pt = keep ? p * inv_keep : 0.f;
            attn[((int64_t)hh * dm.B + b) * L + l] = pt;
            tot += pt;
        }
        tot = warp_sum(tot);
        if (lane == 0) psum[(int64_t)b * H + hh] = tot;
    }
}

// =====================================================================================================================
// PACK4: H <= 4 — the dbank product takes ONE K step whose eight K slots are [dS of heads 0-3 | P of heads 0-3]
template <int NT, bool PACK4>
__global__ void __launch_bounds__(TCA_THREADS, 2) attn_q1_tc_bwd_kernel(
    const float* __restrict__ u, const float* __restrict__ bank, const float* __restrict__ mask,
    const float* __restrict__ lse, const float* __restrict__ gctx, const float* __restrict__ gpsum, TcaDims dm,
    float scale, float p_drop, uint64_t seed, const uint64_t* seed_offset,
    float* __restrict__ gu, float* __restrict__ gbank) {
    constexpr int NC = 8 * NT;
    extern __shared__ __align__(16) float sm[];
    const int CH = TCA_R * dm.D + TCA_PAD;
    float* chunk = sm;                                   // [2][CH]
    const int UR = PACK4 ? 8 : dm.HR;                    // PACK4: query tables hold [hi | lo] rows (see tca_score_partials)
    float* us = chunk + 2 * CH;                          // [UR][DU]
    float* gs = us + UR * dm.DU;                         // [UR][DU]
    float* part = gs + UR * dm.DU;                       // [4][R][2*NC]
    float* sc = part + TCA_KSPLIT * TCA_R * 2 * NC;      // [HR][LT] scores, then scale*dS
    float* tt = sc + dm.HR * dm.LT;                      // [HR][LT] <dctx,k>, then dropped-out probabilities
    uint64_t* mbar = reinterpret_cast<uint64_t*>(tt + dm.HR * dm.LT);
    unsigned char* lv = reinterpret_cast<unsigned char*>(mbar + 2);
    __shared__ int s_lb;

    const int b = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int H = dm.H, L = dm.L, D = dm.D;
    const float inv_keep = 1.f / (1.f - p_drop);
    if (seed_offset != nullptr) seed += *seed_offset;
    const float* bk = bank + (int64_t)b * L * D;
    float* gb = gbank + (int64_t)b * L * D;

    if (threadIdx.x == 0) {
        tca_mbar_init(&mbar[0], 1);
        tca_mbar_init(&mbar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = threadIdx.x; i < 2 * CH; i += TCA_THREADS) chunk[i] = 0.f;
    for (int i = threadIdx.x; i < UR * dm.DU; i += TCA_THREADS) {
        const int r = i / dm.DU, d = i - r * dm.DU;
        const int h = PACK4 ? (r & 3) : r;
        const bool in = (h < H && d < D);
        const float xu = in ? u[((int64_t)b * H + h) * D + d] : 0.f;
        const float xg = in ? gctx[((int64_t)b * H + h) * D + d] : 0.f;
        if (PACK4) {
            uint32_t hi, lo;
            tca_split(xu, hi, lo);
            us[i] = __uint_as_float(r < 4 ? hi : lo);
            tca_split(xg, hi, lo);
            gs[i] = __uint_as_float(r < 4 ? hi : lo);
        } else {
            us[i] = xu;
            gs[i] = xg;
        }
    }
    for (int i = threadIdx.x; i < dm.HR * dm.LT; i += TCA_THREADS) { sc[i] = -INFINITY; tt[i] = 0.f; }
    const int Lb = tca_live_rows(mask ? mask + (int64_t)b * L : nullptr, L, lv, &s_lb);
    const int nchunks = (Lb + TCA_R - 1) / TCA_R;

    int n_issued = 0;                                    // loads are numbered across both sweeps (thread 0 only)
    auto issue = [&](int c) {
        const int rows = min(TCA_R, Lb - c * TCA_R);
        const uint32_t bytes = (uint32_t)rows * D * 4u;
        const int slot = n_issued & 1;
        tca_mbar_expect_tx(&mbar[slot], bytes);
        tca_bulk_load(chunk + slot * CH, bk + (int64_t)c * TCA_R * D, bytes, &mbar[slot]);
        ++n_issued;
    };
    if (threadIdx.x == 0 && nchunks > 0) {
        tca_fence_proxy_async();
        issue(0);
    }

    // ---- sweep A: S = bank.u^T, T = bank.dctx^T
    int n_wait = 0;
    for (int c = 0; c < nchunks; ++c, ++n_wait) {
        const float* cb = chunk + (n_wait & 1) * CH;
        if (threadIdx.x == 0 && c + 1 < nchunks) issue(c + 1);      // the other buffer: its readers passed the barrier below
        tca_mbar_wait(&mbar[n_wait & 1], (n_wait >> 1) & 1);
        tca_score_partials<NT, 2 * NT, PACK4>(cb, us, gs, part, dm, warp, lane);
        __syncthreads();
        for (int i = threadIdx.x; i < TCA_R * 2 * NC; i += TCA_THREADS) {
            const int r = i / (2 * NC), col = i - r * (2 * NC);
            if (PACK4 && (col & 4)) continue;            // columns 4-7 / 12-15 are the lo halves, added below
            float v = 0.f;
#pragma unroll
            for (int q = 0; q < TCA_KSPLIT; ++q) {
                v += part[(q * TCA_R + r) * 2 * NC + col];
                if (PACK4) v += part[(q * TCA_R + r) * 2 * NC + col + 4];
            }
            const int l = c * TCA_R + r;
            // head tiles are interleaved in the partial table as [u tiles | dctx tiles]
            if (col < NC) {
                const bool live = (col < H) && (l < Lb) && lv[min(l, L - 1)] != 0;
                if (col < dm.HR) sc[col * dm.LT + l] = live ? v * scale : -INFINITY;
            } else if (col - NC < dm.HR) {
                tt[(col - NC) * dm.LT + l] = v;
            }
        }
        __syncthreads();
    }

    // ---- softmax backward on the [H, LT] tables; padding heads and rows become zeros
    for (int hh = warp; hh < dm.HR; hh += TCA_WARPS) {
        float* scr = sc + hh * dm.LT;
        float* ttr = tt + hh * dm.LT;
        if (hh >= H) {
            for (int l = lane; l < dm.LT; l += 32) { scr[l] = 0.f; ttr[l] = 0.f; }
            continue;
        }
        const float lse_h = lse[(int64_t)b * H + hh];
        const float gp = gpsum ? gpsum[(int64_t)b * H + hh] : 0.f;
        float delta = 0.f;
        for (int l = lane; l < dm.LT; l += 32) {
            const float sv = scr[l];
            const float p = (sv == -INFINITY) ? 0.f : __expf(sv - lse_h);
            bool keep = true;
            if (p_drop > 0.f && l < L) keep = uniform01(seed, ((uint64_t)b * H + hh) * L + l) >= p_drop;
            const float dp = keep ? (ttr[l] + gp) * inv_keep : 0.f;
            delta += p * dp;
        }
        delta = warp_sum(delta);
        for (int l = lane; l < dm.LT; l += 32) {
            const float sv = scr[l];
            const float p = (sv == -INFINITY) ? 0.f : __expf(sv - lse_h);
            bool keep = true;
            if (p_drop > 0.f && l < L) keep = uniform01(seed, ((uint64_t)b * H + hh) * L + l) >= p_drop;
            const float dp = keep ? (ttr[l] + gp) * inv_keep : 0.f;
            scr[l] = p * (dp - delta) * scale;
            ttr[l] = keep ? p * inv_keep : 0.f;
        }
    }
    __syncthreads();

    // ---- sweep B: du^T += chunk^T . dS ;  dbank chunk = dS.u + P.dctx
    float du[3][NT][4];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < NT; ++j) du[i][j][0] = du[i][j][1] = du[i][j][2] = du[i][j][3] = 0.f;
    if (threadIdx.x == 0 && nchunks > 0) issue(0);
    for (int c = 0; c < nchunks; ++c, ++n_wait) {
        const float* cb = chunk + (n_wait & 1) * CH;
        if (threadIdx.x == 0 && c + 1 < nchunks) issue(c + 1);
        tca_mbar_wait(&mbar[n_wait & 1], (n_wait >> 1) & 1);
        tca_accumulate_t<NT, true, false>(cb, sc + c * TCA_R, dm.LT, du, dm, warp, lane);
        // dbank: C tile = (16-row tile mt2, 8-feature tile nn); tiles dealt round-robin to the warps
        const int ntiles = 2 * dm.KS;
        for (int q = warp; q < ntiles; q += TCA_WARPS) {
            const int nn = q >> 1, mt2 = q & 1;
            const int l0 = c * TCA_R + mt2 * 16;
            float cc[4] = {0.f, 0.f, 0.f, 0.f};
            if (PACK4) {
                float a[4];
                a[0] = sc[t * dm.LT + l0 + g];
                a[1] = sc[t * dm.LT + l0 + g + 8];
                a[2] = tt[t * dm.LT + l0 + g];
                a[3] = tt[t * dm.LT + l0 + g + 8];
                uint32_t ah[4], al[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) tca_split(a[k], ah[k], al[k]);
                const int n = nn * 8 + g;
                const uint32_t b0h = __float_as_uint(us[t * dm.DU + n]), b0l = __float_as_uint(us[(t + 4) * dm.DU + n]);
                const uint32_t b1h = __float_as_uint(gs[t * dm.DU + n]), b1l = __float_as_uint(gs[(t + 4) * dm.DU + n]);
                tca_mma(cc, ah, b0h, b1h);
                tca_mma(cc, al, b0h, b1h);
                tca_mma(cc, ah, b0l, b1l);
            } else {
#pragma unroll
                for (int j = 0; j < 2 * NT; ++j) {
                    const float* T = (j < NT ? sc + j * 8 * dm.LT : tt + (j - NT) * 8 * dm.LT);
                    const float* U = (j < NT ? us + j * 8 * dm.DU : gs + (j - NT) * 8 * dm.DU);
                    float a[4];
                    a[0] = T[t * dm.LT + l0 + g];
                    a[1] = T[t * dm.LT + l0 + g + 8];
                    a[2] = T[(t + 4) * dm.LT + l0 + g];
                    a[3] = T[(t + 4) * dm.LT + l0 + g + 8];
                    tca_mma3(cc, a, U[t * dm.DU + nn * 8 + g], U[(t + 4) * dm.DU + nn * 8 + g]);
                }
            }
            const int col = nn * 8 + 2 * t;
            if (col < D) {                               // D % 4 == 0 and col even: col + 1 < D as well
                const int la = l0 + g, lb2 = l0 + g + 8;
                if (la < L) *reinterpret_cast<float2*>(gb + (int64_t)la * D + col) = make_float2(cc[0], cc[1]);
                if (lb2 < L) *reinterpret_cast<float2*>(gb + (int64_t)lb2 * D + col) = make_float2(cc[2], cc[3]);
            }
        }
        __syncthreads();                                 // every warp is done with this buffer before it is refilled
    }
    // rows past the last processed chunk carry no probability mass: zero gradient
    {
        const int64_t z0 = (int64_t)min(nchunks * TCA_R, L) * D, z1 = (int64_t)L * D;
        for (int64_t i = z0 + 4 * threadIdx.x; i < z1; i += 4 * TCA_THREADS)
            *reinterpret_cast<float4*>(gb + i) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const int mt = warp + TCA_WARPS * i;
        if (mt >= dm.MT) break;
#pragma unroll
        for (int j = 0; j < NT; ++j) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int d = mt * 16 + g + ((e & 2) ? 8 : 0);
                const int hh = j * 8 + 2 * t + (e & 1);
                if (d < D && hh < H) gu[((int64_t)b * H + hh) * D + d] = du[i][j][e];
            }
        }
    }
}

static TcaDims tca_dims(int B, int H, int L, int D) {
    TcaDims dm;
    dm.B = B; dm.H = H; dm.L = L; dm.D = D;
    dm.KS = (D + 7) / 8;
    dm.MT = (D + 15) / 16;
    dm.DU = dm.KS * 8;
    while (dm.DU % 8 != 4) dm.DU += 1;                   // = 4 mod 8 (and a multiple of 4): conflict-free fragment loads
    dm.LT = (L + TCA_R - 1) / TCA_R * TCA_R;
    while (dm.LT % 8 != 4) dm.LT += 1;
    dm.HR = H <= 4 ? 4 : 8 * ((H + 7) / 8);
    return dm;
}
static size_t tca_fwd_smem(const TcaDims& dm, int NT) {
    const size_t NC = 8 * NT;
    const size_t UR = dm.H <= 4 ? 8 : dm.HR;
    const size_t floats = 2 * ((size_t)TCA_R * dm.D + TCA_PAD) + UR * dm.DU + (size_t)TCA_KSPLIT * TCA_R * NC + TCA_R * NC +
                          (size_t)dm.HR * dm.LT + 3 * NC;
    return floats * 4 + 16 + ((size_t)dm.L + 15) / 16 * 16;
}
static size_t tca_bwd_smem(const TcaDims& dm, int NT) {
    const size_t NC = 8 * NT;
    const size_t UR = dm.H <= 4 ? 8 : dm.HR;
    const size_t floats = 2 * ((size_t)TCA_R * dm.D + TCA_PAD) + 2 * UR * dm.DU + (size_t)TCA_KSPLIT * TCA_R * 2 * NC +
                          2 * (size_t)dm.HR * dm.LT;
    return floats * 4 + 16 + ((size_t)dm.L + 15) / 16 * 16;
}

}  // namespace mgnns

using namespace mgnns;

// 1 if the tensor-core attention kernels cover this shape (the caller falls back to mgnns_attn_q1_* otherwise)
extern "C" int mgnns_attn_q1_tc_supported(int H, int L, int D) {
    if (H < 1 || H > 16 || L < 1 || D < 8 || D % 4 != 0 || D > 512) return 0;
    const TcaDims dm = tca_dims(1, H, L, D);
    const int NT = (H + 7) / 8;
    if (dm.MT > 3 * TCA_WARPS) return 0;
    return (tca_fwd_smem(dm, NT) <= 226 * 1024 && tca_bwd_smem(dm, NT) <= 226 * 1024) ? 1 : 0;
}

extern "C" int mgnns_attn_q1_tc_fwd(const float* u, const float* bank, const float* mask,
                                    int B, int H, int L, int D, float scale, float p_drop, uint64_t seed,
                                    const uint64_t* seed_offset, float* ctx, float* attn, float* psum, float* lse, void* stream) {
    MG_REQUIRE(B >= 0 && mgnns_attn_q1_tc_supported(H, L, D), "attn_q1_tc_fwd: unsupported shape H=%d L=%d D=%d", H, L, D);
    MG_REQUIRE(p_drop >= 0.f && p_drop < 1.f, "attn_q1_tc_fwd: p_drop must be in [0,1)");
    if (B == 0) return 0;
    MG_REQUIRE(u && bank && ctx && attn && psum && lse, "attn_q1_tc_fwd: null pointer");
    MG_REQUIRE(aligned16(bank), "attn_q1_tc_fwd: bank must be 16-byte aligned");
    const TcaDims dm = tca_dims(B, H, L, D);
    const int NT = (H + 7) / 8;
    const size_t smem = tca_fwd_smem(dm, NT);
    cudaStream_t st = as_stream(stream);
#define TCA_FWD(NTV, M, DR, PK)                                                                                           \
    do {                                                                                                                   \
        cudaFuncSetAttribute(attn_q1_tc_fwd_kernel<NTV, M, DR, PK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        attn_q1_tc_fwd_kernel<NTV, M, DR, PK><<<B, TCA_THREADS, smem, st>>>(u, bank, mask, dm, scale, p_drop, seed,          \
                                                                            seed_offset, ctx, attn, psum, lse);            \
    } while (0)
#define TCA_FWD_FLAGS(NTV, PK)                                       \
    do {                                                             \
        if (mask != nullptr && p_drop > 0.f) TCA_FWD(NTV, true, true, PK);   \
        else if (mask != nullptr) TCA_FWD(NTV, true, false, PK);     \
        else if (p_drop > 0.f) TCA_FWD(NTV, false, true, PK);        \
        else TCA_FWD(NTV, false, false, PK);                         \
    } while (0)
    if (H <= 4) TCA_FWD_FLAGS(1, true); else if (NT == 1) TCA_FWD_FLAGS(1, false); else TCA_FWD_FLAGS(2, false);
#undef TCA_FWD_FLAGS
#undef TCA_FWD
    MG_LAUNCH_CHECK("attn_q1_tc_fwd");
    return 0;
}

extern "C" int mgnns_attn_q1_tc_bwd(const float* u, const float* bank, const float* mask, const float* lse,
                                    const float* grad_ctx, const float* grad_psum,
                                    int B, int H, int L, int D, float scale, float p_drop, uint64_t seed,
                                    const uint64_t* seed_offset, float* grad_u, float* grad_bank, void* stream) {
    MG_REQUIRE(B >= 0 && mgnns_attn_q1_tc_supported(H, L, D), "attn_q1_tc_bwd: unsupported shape H=%d L=%d D=%d", H, L, D);
    MG_REQUIRE(p_drop >= 0.f && p_drop < 1.f, "attn_q1_tc_bwd: p_drop must be in [0,1)");
    if (B == 0) return 0;
    MG_REQUIRE(u && bank && lse && grad_ctx && grad_u && grad_bank, "attn_q1_tc_bwd: null pointer");
    MG_REQUIRE(aligned16(bank) && aligned16(grad_bank), "attn_q1_tc_bwd: bank / grad_bank must be 16-byte aligned");
    const TcaDims dm = tca_dims(B, H, L, D);
    const int NT = (H + 7) / 8;
    const size_t smem = tca_bwd_smem(dm, NT);
    cudaStream_t st = as_stream(stream);
#define TCA_BWD(NTV, PK)                                                                                                  \
    do {                                                                                                                   \
        cudaFuncSetAttribute(attn_q1_tc_bwd_kernel<NTV, PK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);      \
        attn_q1_tc_bwd_kernel<NTV, PK><<<B, TCA_THREADS, smem, st>>>(u, bank, mask, lse, grad_ctx, grad_psum, dm, scale,   \
                                                                     p_drop, seed, seed_offset, grad_u, grad_bank);       \
    } while (0)
    if (H <= 4) TCA_BWD(1, true);
    else if (NT == 1) TCA_BWD(1, false);
    else TCA_BWD(2, false);
#undef TCA_BWD
    MG_LAUNCH_CHECK("attn_q1_tc_bwd");
    return 0;
}
