// tcgen05 / TMEM / TMA dense layer:  C[M,N] = act(A[M,K] . W + bias)
// (ref: support = torch.matmul(input, weight) in GraphConvolution.forward,
//  models/Multi_GCN_Multihead_att.py:52-58, plus the activation the caller applies right after it,
//  :470-472; nn.Linear call sites with a large row count.)
//
//   A  row-major fp32 (K contiguous)              -> K-major UMMA operand, 128-byte swizzle via TMA
//   W  [K,N] (GraphConvolution layout, N contig.) -> MN-major operand (SWIZZLE_128B_BASE32B atoms)
//      [N,K] (nn.Linear layout, K contiguous)     -> K-major operand
//
// One persistent CTA per SM walks a contiguous range of (128-row tile, BN-column tile) work items.
// Warp 0 = TMA producer, warp 1 = MMA issuer (one elected lane), warp 2 = TMEM allocator,
// warps 4-7 = operand splitter, warps 8-11 = epilogue.  The fp32 accumulator lives in TMEM and is
// double buffered (2 x 256 columns), so the epilogue of one tile (tcgen05.ld -> bias -> activation
// -> 128-bit global stores) overlaps the MMAs of the next.
//
// Precision: SPLIT=false is plain TF32.  SPLIT=true is 3xTF32 (fp32-class): x = hi + lo with
// hi = x & ~0x1fff; acc += hi*hi + lo*hi + hi*lo.  W is split ONCE per call into a caller-provided
// workspace by a small prep kernel (it is re-read by every row tile), A is split in shared memory
// by the splitter warps as each tile lands — shared-memory bandwidth, not the tensor pipe, is what
// bounds the split scheme, so the weight half of that traffic is taken out of the main loop.
#include "tc_common.cuh"

namespace mgnns {
namespace tc {

constexpr int LK = 32;                     // fp32 elements per 128-byte swizzle row (K chunk)
constexpr int L_UMMA_K = 8;
constexpr int L_THREADS = 384;
constexpr int L_BM = 128;
constexpr int L_BN_MAX = 256;
constexpr int L_A_BYTES = L_BM * LK * 4;           // 16 KB
constexpr int L_B_BYTES = L_BN_MAX * LK * 4;       // 32 KB

struct LinParams {
    int M, N, K;
    const float* bias;
    int act;
    float slope;
    float* C;
    int64_t ldc;
    int n_tiles, BN, k_chunks, n_items;
    int* counter;                          // dynamic tile scheduler: next work item (zeroed before the launch)
};

template <bool SPLIT>
struct LinCfg {
    static constexpr int STAGE_BYTES = (L_A_BYTES + L_B_BYTES) * (SPLIT ? 2 : 1);
    static constexpr int STAGES = SPLIT ? 2 : 4;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 512 /*barriers + scheduler ring*/;
};

template <bool SPLIT, bool B_MN>
__global__ void __launch_bounds__(L_THREADS, 1)
tc_linear_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmBhi,
                 const __grid_constant__ CUtensorMap tmBlo, LinParams p) {
    using CF = LinCfg<SPLIT>;
    constexpr int STAGES = CF::STAGES;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * CF::STAGE_BYTES);
    uint64_t* full = bars;                       // TMA bytes landed
    uint64_t* ready = bars + STAGES;             // (SPLIT) A hi/lo written
    uint64_t* empty = bars + 2 * STAGES;         // MMAs that read the stage have completed
    uint64_t* tmem_full = bars + 3 * STAGES;     // [2] accumulator complete
    uint64_t* tmem_empty = tmem_full + 2;        // [2] accumulator drained by the epilogue
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
    SchedSmem* sched = reinterpret_cast<SchedSmem*>(smem + STAGES * CF::STAGE_BYTES + 256);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // stage layout: A_hi | B_hi | A_lo | B_lo
    auto stageA = [&](int s) { return smem + s * CF::STAGE_BYTES; };
    auto stageB = [&](int s) { return smem + s * CF::STAGE_BYTES + L_A_BYTES; };
    auto stageAlo = [&](int s) { return smem + s * CF::STAGE_BYTES + L_A_BYTES + L_B_BYTES; };
    auto stageBlo = [&](int s) { return smem + s * CF::STAGE_BYTES + 2 * L_A_BYTES + L_B_BYTES; };

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmA);
        prefetch_tmap(&tmBhi);
        if (SPLIT) prefetch_tmap(&tmBlo);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&ready[s], 4);
            mbar_init(&empty[s], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tmem_full[i], 1);
            mbar_init(&tmem_empty[i], 4);
        }
        sched_init(sched);
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // work items (n tile fastest: the A row tile is re-read from L2 right away) come from the dynamic scheduler
    const int BN = p.BN;

    if (warp == 0) {
        // ===================================================== TMA producer
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            const uint32_t tx = L_A_BYTES + (uint32_t)BN * 128u * (SPLIT ? 2u : 1u);
            SchedState ss;
            for (;;) {
                const int item = sched_produce(sched, ss, p.counter);
                if (item >= p.n_items) break;
                const int mt = item / p.n_tiles, nt = item - mt * p.n_tiles;
                for (int kc = 0; kc < p.k_chunks; ++kc) {
                    mbar_wait(&empty[stage], phase ^ 1);
                    mbar_expect_tx(&full[stage], tx);
                    tma_load_2d(stageA(stage), &tmA, &full[stage], kc * LK, mt * L_BM);
                    if (B_MN) {
                        // BN/32 boxes of 32(n) x 32(k) -> MN-major atoms 4 KB apart
                        for (int j = 0; j < BN / 32; ++j) {
                            tma_load_2d(stageB(stage) + j * 4096, &tmBhi, &full[stage], nt * BN + j * 32, kc * LK);
                            if (SPLIT)
                                tma_load_2d(stageBlo(stage) + j * 4096, &tmBlo, &full[stage], nt * BN + j * 32, kc * LK);
                        }
                    } else {
                        // one box of 32(k) x BN(n) rows
                        tma_load_2d(stageB(stage), &tmBhi, &full[stage], kc * LK, nt * BN);
                        if (SPLIT) tma_load_2d(stageBlo(stage), &tmBlo, &full[stage], kc * LK, nt * BN);
                    }
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================================================== MMA issuer
        if (lane == 0) {
            const uint32_t idesc = instr_desc_tf32(L_BM, BN, 0, B_MN ? 1 : 0);
            constexpr uint32_t A_LBO = 16, A_SBO = 1024, A_KSTEP = 32, A_LT = 2;
            constexpr uint32_t B_LBO = B_MN ? 4096 : 16, B_SBO = B_MN ? 512 : 1024, B_KSTEP = B_MN ? 1024 : 32;
            constexpr uint32_t B_LT = B_MN ? 1 : 2;
            int stage = 0;
            uint32_t phase = 0, tphase[2] = {0, 0};
            SchedState ss;
            for (int it = 0;; ++it) {
                const int item = sched_consume_thread(sched, ss);
                if (item >= p.n_items) break;
                const int ab = it & 1;
                const uint32_t tacc = tmem_base + (uint32_t)ab * L_BN_MAX;
                mbar_wait(&tmem_empty[ab], tphase[ab] ^ 1);
                tc_fence_after();
                uint32_t accumulate = 0;
                for (int kc = 0; kc < p.k_chunks; ++kc) {
                    mbar_wait(SPLIT ? &ready[stage] : &full[stage], phase);
                    tc_fence_after();
                    const uint32_t a_hi = smem_u32(stageA(stage)), b_hi = smem_u32(stageB(stage));
                    const uint32_t a_lo = smem_u32(stageAlo(stage)), b_lo = smem_u32(stageBlo(stage));
#pragma unroll
                    for (int ks = 0; ks < LK / L_UMMA_K; ++ks) {
                        const uint64_t dah = smem_desc(a_hi + ks * A_KSTEP, A_LBO, A_SBO, A_LT);
                        const uint64_t dbh = smem_desc(b_hi + ks * B_KSTEP, B_LBO, B_SBO, B_LT);
                        umma_tf32(tacc, dah, dbh, idesc, accumulate);
                        accumulate = 1;
                        if (SPLIT) {
                            const uint64_t dal = smem_desc(a_lo + ks * A_KSTEP, A_LBO, A_SBO, A_LT);
                            const uint64_t dbl = smem_desc(b_lo + ks * B_KSTEP, B_LBO, B_SBO, B_LT);
                            umma_tf32(tacc, dal, dbh, idesc, 1);
                            umma_tf32(tacc, dah, dbl, idesc, 1);
                        }
                    }
                    umma_commit(&empty[stage]);          // frees the stage once these MMAs have read it
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                umma_commit(&tmem_full[ab]);
                tphase[ab] ^= 1;
            }
        }
    } else if (warp >= 4 && warp < 8) {
        // ===================================================== operand splitter (3xTF32: A only)
        {
            const int t = threadIdx.x - 128;
            int stage = 0;
            uint32_t phase = 0;
            SchedState ss;
            for (;;) {
                const int item = sched_consume_warp(sched, ss, lane);
                if (item >= p.n_items) break;
                if (!SPLIT) continue;                    // plain TF32: only keep the scheduler ring moving
                for (int kc = 0; kc < p.k_chunks; ++kc) {
                    mbar_wait(&full[stage], phase);
                    float4* hi = reinterpret_cast<float4*>(stageA(stage));
                    float4* lo = reinterpret_cast<float4*>(stageAlo(stage));
                    constexpr int NV = L_A_BYTES / 16;
#pragma unroll
                    for (int i = 0; i < NV / 128; ++i) {
                        const int idx = t + i * 128;
                        float4 x = hi[idx];
                        float4 h, l;
                        h.x = __uint_as_float(__float_as_uint(x.x) & 0xFFFFE000u); l.x = x.x - h.x;
                        h.y = __uint_as_float(__float_as_uint(x.y) & 0xFFFFE000u); l.y = x.y - h.y;
                        h.z = __uint_as_float(__float_as_uint(x.z) & 0xFFFFE000u); l.z = x.z - h.z;
                        h.w = __uint_as_float(__float_as_uint(x.w) & 0xFFFFE000u); l.w = x.w - h.w;
                        hi[idx] = h;
                        lo[idx] = l;
                    }
                    fence_proxy_async();                 // generic-proxy writes -> visible to the tensor core's async proxy
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&ready[stage]);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp >= 8) {
        // ===================================================== epilogue (TMEM -> registers -> global)
        const int q = warp & 3;                          // TMEM lane quarter this warp may read
        const int row = q * 32 + lane;
        uint32_t tphase[2] = {0, 0};
        const bool vec_ok = ((p.ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.C) & 15u) == 0);
        SchedState ss;
        for (int it = 0;; ++it) {
            const int item = sched_consume_warp(sched, ss, lane);
            if (item >= p.n_items) break;
            const int ab = it & 1;
            const int mt = item / p.n_tiles, nt = item - mt * p.n_tiles;
            mbar_wait(&tmem_full[ab], tphase[ab]);
            tc_fence_after();
            const uint32_t taddr = tmem_base + (uint32_t)ab * L_BN_MAX + ((uint32_t)(q * 32) << 16);
            const int m = mt * L_BM + row;
            const int n_base = nt * BN;
            float* dst = p.C + (int64_t)m * p.ldc;
            for (int c0 = 0; c0 < BN; c0 += 32) {
                uint32_t r[32];
                tmem_ld16(taddr + c0, *reinterpret_cast<uint32_t(*)[16]>(&r[0]));
                tmem_ld16(taddr + c0 + 16, *reinterpret_cast<uint32_t(*)[16]>(&r[16]));
                tmem_ld_wait();
                if (m < p.M) {
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        const int n = n_base + c0 + j;
                        if (n >= p.N) break;
                        float v[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            float x = __uint_as_float(r[j + e]);
                            if (p.bias != nullptr && n + e < p.N) x += __ldg(p.bias + n + e);
                            v[e] = apply_act(x, p.act, p.slope);
                        }
                        if (vec_ok && n + 3 < p.N) {
                            *reinterpret_cast<float4*>(dst + n) = make_float4(v[0], v[1], v[2], v[3]);
                        } else {
#pragma unroll
                            for (int e = 0; e < 4; ++e)
                                if (n + e < p.N) dst[n + e] = v[e];
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[ab]);
            tphase[ab] ^= 1;
        }
    }

    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 2) tmem_dealloc(tmem_base, 512);
}

// =====================================================================================================
// Weight-gradient product  C[M,N] += A[K,M]^T . B[K,N]  with a long reduction (K = tokens / rows of a batch):
// both operands are row-major with the reduction index as the row, so both are MN-major UMMA operands
// (32x32 TMA boxes, SWIZZLE_128B_BASE32B).  Work item = (128-column tile of A, BN-column tile of B, slice of K);
// partial tiles are added to the zero-initialised C with fp32 atomics.  Same warp roles, dynamic scheduler and
// double-buffered TMEM accumulators as tc_linear_kernel; both operands are activations, so the splitter warps
// split A and B in shared memory (3xTF32).
// (ref: the dW of nn.LSTM's input / recurrent projections, model:179-184, computed by autograd in the reference)
// =====================================================================================================
struct WgradParams {
    int M, N, K;
    float* C;
    int64_t ldc;
    int m_tiles, n_tiles, BN, k_splits, chunks_per_split, k_chunks, n_items;
    int* counter;
};

template <bool SPLIT>
__global__ void __launch_bounds__(L_THREADS, 1)
tc_wgrad_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, WgradParams p) {
    using CF = LinCfg<SPLIT>;
    constexpr int STAGES = CF::STAGES;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * CF::STAGE_BYTES);
    uint64_t* full = bars;
    uint64_t* ready = bars + STAGES;
    uint64_t* empty = bars + 2 * STAGES;
    uint64_t* tmem_full = bars + 3 * STAGES;
    uint64_t* tmem_empty = tmem_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
    SchedSmem* sched = reinterpret_cast<SchedSmem*>(smem + STAGES * CF::STAGE_BYTES + 256);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // stage layout: A_hi | B_hi | A_lo | B_lo   (A and B contiguous, so are their low halves)
    auto stageA = [&](int s) { return smem + s * CF::STAGE_BYTES; };
    auto stageB = [&](int s) { return smem + s * CF::STAGE_BYTES + L_A_BYTES; };
    auto stageAlo = [&](int s) { return smem + s * CF::STAGE_BYTES + L_A_BYTES + L_B_BYTES; };
    auto stageBlo = [&](int s) { return smem + s * CF::STAGE_BYTES + 2 * L_A_BYTES + L_B_BYTES; };

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmA);
        prefetch_tmap(&tmB);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&ready[s], 4);
            mbar_init(&empty[s], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tmem_full[i], 1);
            mbar_init(&tmem_empty[i], 4);
        }
        sched_init(sched);
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int BN = p.BN;
    const int tiles = p.m_tiles * p.n_tiles;
    // item -> (k slice, m tile, n tile): consecutive items share the k slice (their operand rows are L2-hot)
    auto chunks_of = [&](int ks) { return max(0, min(p.chunks_per_split, p.k_chunks - ks * p.chunks_per_split)); };

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            const uint32_t tx = L_A_BYTES + (uint32_t)BN * 128u;
            SchedState ss;
            for (;;) {
                const int item = sched_produce(sched, ss, p.counter);
                if (item >= p.n_items) break;
                const int ks = item / tiles, t = item - ks * tiles;
                const int mt = t / p.n_tiles, nt = t - mt * p.n_tiles;
                const int nch = chunks_of(ks), kc0 = ks * p.chunks_per_split;
                for (int kc = 0; kc < nch; ++kc) {
                    mbar_wait(&empty[stage], phase ^ 1);
                    mbar_expect_tx(&full[stage], tx);
                    const int k0 = (kc0 + kc) * LK;
                    for (int j = 0; j < 4; ++j)
                        tma_load_2d(stageA(stage) + j * 4096, &tmA, &full[stage], mt * L_BM + j * 32, k0);
                    for (int j = 0; j < BN / 32; ++j)
                        tma_load_2d(stageB(stage) + j * 4096, &tmB, &full[stage], nt * BN + j * 32, k0);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = instr_desc_tf32(L_BM, BN, 1, 1);
            constexpr uint32_t LBO = 4096, SBO = 512, KSTEP = 1024, LT = 1;      // MN-major, 32-byte swizzle atoms
            int stage = 0;
            uint32_t phase = 0, tphase[2] = {0, 0};
            SchedState ss;
            for (int it = 0;; ++it) {
                const int item = sched_consume_thread(sched, ss);
                if (item >= p.n_items) break;
                const int nch = chunks_of(item / tiles);
                const int ab = it & 1;
                const uint32_t tacc = tmem_base + (uint32_t)ab * L_BN_MAX;
                mbar_wait(&tmem_empty[ab], tphase[ab] ^ 1);
                tc_fence_after();
                uint32_t accumulate = 0;
                for (int kc = 0; kc < nch; ++kc) {
                    mbar_wait(SPLIT ? &ready[stage] : &full[stage], phase);
                    tc_fence_after();
                    const uint32_t a_hi = smem_u32(stageA(stage)), b_hi = smem_u32(stageB(stage));
                    const uint32_t a_lo = smem_u32(stageAlo(stage)), b_lo = smem_u32(stageBlo(stage));
#pragma unroll
                    for (int k8 = 0; k8 < LK / L_UMMA_K; ++k8) {
                        const uint64_t dah = smem_desc(a_hi + k8 * KSTEP, LBO, SBO, LT);
                        const uint64_t dbh = smem_desc(b_hi + k8 * KSTEP, LBO, SBO, LT);
                        umma_tf32(tacc, dah, dbh, idesc, accumulate);
                        accumulate = 1;
                        if (SPLIT) {
                            const uint64_t dal = smem_desc(a_lo + k8 * KSTEP, LBO, SBO, LT);
                            const uint64_t dbl = smem_desc(b_lo + k8 * KSTEP, LBO, SBO, LT);
                            umma_tf32(tacc, dal, dbh, idesc, 1);
                            umma_tf32(tacc, dah, dbl, idesc, 1);
                        }
                    }
                    umma_commit(&empty[stage]);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                if (nch > 0) umma_commit(&tmem_full[ab]);
                else mbar_arrive(&tmem_full[ab]);
                tphase[ab] ^= 1;
            }
        }
    } else if (warp >= 4 && warp < 8) {
        const int t = threadIdx.x - 128;
        int stage = 0;
        uint32_t phase = 0;
        SchedState ss;
        for (;;) {
            const int item = sched_consume_warp(sched, ss, lane);
            if (item >= p.n_items) break;
            if (!SPLIT) continue;
            const int nch = chunks_of(item / tiles);
            const int nv = (L_A_BYTES + BN * 128) / 16;          // A then the used part of B, contiguous
            for (int kc = 0; kc < nch; ++kc) {
                mbar_wait(&full[stage], phase);
                float4* hi = reinterpret_cast<float4*>(stageA(stage));
                float4* lo = reinterpret_cast<float4*>(stageAlo(stage));
#pragma unroll 4
                for (int i = t; i < nv; i += 128) {
                    float4 x = hi[i];
                    float4 h, l;
                    h.x = __uint_as_float(__float_as_uint(x.x) & 0xFFFFE000u); l.x = x.x - h.x;
                    h.y = __uint_as_float(__float_as_uint(x.y) & 0xFFFFE000u); l.y = x.y - h.y;
                    h.z = __uint_as_float(__float_as_uint(x.z) & 0xFFFFE000u); l.z = x.z - h.z;
                    h.w = __uint_as_float(__float_as_uint(x.w) & 0xFFFFE000u); l.w = x.w - h.w;
                    hi[i] = h;
                    lo[i] = l;
                }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(&ready[stage]);
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp >= 8) {
        const int q = warp & 3;
        const int row = q * 32 + lane;
        uint32_t tphase[2] = {0, 0};
        SchedState ss;
        for (int it = 0;; ++it) {
            const int item = sched_consume_warp(sched, ss, lane);
            if (item >= p.n_items) break;
            const int ks = item / tiles, t = item - ks * tiles;
            const int mt = t / p.n_tiles, nt = t - mt * p.n_tiles;
            const int ab = it & 1;
            mbar_wait(&tmem_full[ab], tphase[ab]);
            tc_fence_after();
            if (chunks_of(ks) > 0) {
                const uint32_t taddr = tmem_base + (uint32_t)ab * L_BN_MAX + ((uint32_t)(q * 32) << 16);
                const int m = mt * L_BM + row;
                float* dst = p.C + (int64_t)m * p.ldc + nt * BN;
                for (int c0 = 0; c0 < BN; c0 += 16) {
                    uint32_t r[16];
                    tmem_ld16(taddr + c0, r);
                    tmem_ld_wait();
                    if (m < p.M) {
#pragma unroll
                        for (int j = 0; j < 16; ++j)
                            if (nt * BN + c0 + j < p.N) atomicAdd(dst + c0 + j, __uint_as_float(r[j]));
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[ab]);
            tphase[ab] ^= 1;
        }
    }

    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 2) tmem_dealloc(tmem_base, 512);
}

template <bool SPLIT>
static int launch_wgrad(const CUtensorMap& a, const CUtensorMap& b, const WgradParams& p, cudaStream_t st) {
    using CF = LinCfg<SPLIT>;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(tc_wgrad_kernel<SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, CF::SMEM_BYTES);
        MG_REQUIRE(e == cudaSuccess, "wgrad_tc: cannot reserve %d bytes of shared memory: %s", CF::SMEM_BYTES, cudaGetErrorString(e));
        configured = true;
    }
    int grid = tc_grid_limit();
    if (grid > p.n_items) grid = p.n_items;
    WgradParams q = p;
    q.counter = next_tile_counter(st);
    MG_REQUIRE(q.counter != nullptr, "wgrad_tc: cannot set up the tile counter");
    tc_wgrad_kernel<SPLIT><<<grid, L_THREADS, CF::SMEM_BYTES, st>>>(a, b, q);
    MG_LAUNCH_CHECK("wgrad_tc");
    return 0;
}

// W -> (hi, lo) with hi exactly representable in TF32
__global__ void split_tf32_kernel(const float* __restrict__ w, float* __restrict__ hi, float* __restrict__ lo, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        const float x = w[i];
        const float h = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
        hi[i] = h;
        lo[i] = x - h;
    }
}

template <bool SPLIT, bool B_MN>
static int launch_linear(const CUtensorMap& a, const CUtensorMap& bhi, const CUtensorMap& blo, const LinParams& p,
                         cudaStream_t st) {
    using CF = LinCfg<SPLIT>;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(tc_linear_kernel<SPLIT, B_MN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             CF::SMEM_BYTES);
        MG_REQUIRE(e == cudaSuccess, "linear_tc: cannot reserve %d bytes of shared memory: %s", CF::SMEM_BYTES,
                   cudaGetErrorString(e));
        configured = true;
    }
    int grid = tc_grid_limit();
    if (grid > p.n_items) grid = p.n_items;
    LinParams q = p;
    q.counter = next_tile_counter(st);
    MG_REQUIRE(q.counter != nullptr, "linear_tc: cannot set up the tile counter");
    tc_linear_kernel<SPLIT, B_MN><<<grid, L_THREADS, CF::SMEM_BYTES, st>>>(a, bhi, blo, q);
    MG_LAUNCH_CHECK("linear_tc");
    return 0;
}

}  // namespace tc
}  // namespace mgnns

using namespace mgnns;
using namespace mgnns::tc;

// floats of workspace mgnns_linear_tc needs for this weight (0 for plain TF32)
extern "C" int64_t mgnns_linear_tc_workspace(int N, int K, int64_t ldw, int w_is_kn, int precision) {
    if (!precision) return 0;
    const int64_t rows = w_is_kn ? K : N;
    return 2 * rows * ldw;
}

static int split_weight(const float* W, int N, int K, int64_t ldw, int w_is_kn, float* workspace, cudaStream_t st) {
    const int64_t n = (w_is_kn ? (int64_t)K : (int64_t)N) * ldw;
    int blocks = (int)((n + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    split_tf32_kernel<<<blocks, 256, 0, st>>>(W, workspace, workspace + n, n);
    MG_LAUNCH_CHECK("split_tf32");
    return 0;
}

// w_hi / w_lo: the weight (precision 0: both = W) or its TF32 split (precision 1)
static int linear_tc_presplit(const float* A, int64_t lda, const float* w_hi, const float* w_lo, int64_t ldw, int w_is_kn,
                              const float* bias, int act, float slope, int M, int N, int K, int precision,
                              float* C, int64_t ldc, cudaStream_t st) {
    LinParams p{};
    p.M = M; p.N = N; p.K = K;
    p.bias = bias; p.act = act; p.slope = slope;
    p.C = C; p.ldc = ldc;
    p.n_tiles = (N + L_BN_MAX - 1) / L_BN_MAX;
    p.BN = (((N + p.n_tiles - 1) / p.n_tiles) + 31) / 32 * 32;
    p.k_chunks = (K + LK - 1) / LK;
    const int64_t m_tiles = ((int64_t)M + L_BM - 1) / L_BM;
    MG_REQUIRE(m_tiles * p.n_tiles < (1LL << 31), "linear_tc: too many tiles");
    p.n_items = (int)(m_tiles * p.n_tiles);

    CUtensorMap ma, mbh, mbl;
    {
        uint64_t dims[2] = {(uint64_t)K, (uint64_t)M};
        uint64_t str[1] = {(uint64_t)lda * 4};
        uint32_t box[2] = {LK, L_BM};
        if (int rc = make_map(&ma, A, 2, dims, str, box)) return rc;
    }
    if (w_is_kn) {
        uint64_t dims[2] = {(uint64_t)N, (uint64_t)K};
        uint64_t str[1] = {(uint64_t)ldw * 4};
        uint32_t box[2] = {32, LK};
        if (int rc = make_map(&mbh, w_hi, 2, dims, str, box, true)) return rc;
        if (int rc = make_map(&mbl, w_lo, 2, dims, str, box, true)) return rc;
    } else {
        uint64_t dims[2] = {(uint64_t)K, (uint64_t)N};
        uint64_t str[1] = {(uint64_t)ldw * 4};
        uint32_t box[2] = {LK, (uint32_t)p.BN};
        if (int rc = make_map(&mbh, w_hi, 2, dims, str, box)) return rc;
        if (int rc = make_map(&mbl, w_lo, 2, dims, str, box)) return rc;
    }
    if (precision) return w_is_kn ? launch_linear<true, true>(ma, mbh, mbl, p, st) : launch_linear<true, false>(ma, mbh, mbl, p, st);
    return w_is_kn ? launch_linear<false, true>(ma, mbh, mbl, p, st) : launch_linear<false, false>(ma, mbh, mbl, p, st);
}

// C[M,N] = act(A[M,K] . W + bias).  w_is_kn: 1 = W is [K,N] (ldw >= N), 0 = W is [N,K] (ldw >= K).
// precision: 0 = TF32, 1 = 3xTF32 (needs mgnns_linear_tc_workspace() floats of 16-byte aligned workspace).
extern "C" int mgnns_linear_tc(const float* A, int64_t lda, const float* W, int64_t ldw, int w_is_kn,
                               const float* bias, int act, float slope, int M, int N, int K, int precision,
                               float* workspace, int64_t workspace_floats, float* C, int64_t ldc, void* stream) {
    MG_REQUIRE(M >= 0 && N >= 1 && K >= 1, "linear_tc: bad dimensions");
    if (M == 0) return 0;
    MG_REQUIRE(A && W && C, "linear_tc: null pointer");
    MG_REQUIRE(lda >= K && ldc >= N && ldw >= (w_is_kn ? N : K), "linear_tc: leading dimension too small");
    MG_REQUIRE((lda % 4) == 0 && (ldw % 4) == 0, "linear_tc: lda=%lld and ldw=%lld must be multiples of 4 (16-byte TMA strides)",
               (long long)lda, (long long)ldw);
    MG_REQUIRE(aligned16(A) && aligned16(W), "linear_tc: A and W must be 16-byte aligned");
    MG_REQUIRE(act == MGNNS_ACT_NONE || act == MGNNS_ACT_RELU || act == MGNNS_ACT_LEAKY, "linear_tc: bad activation %d", act);
    cudaStream_t st = as_stream(stream);
    const float* w_hi = W;
    const float* w_lo = W;
    if (precision) {
        const int64_t need = mgnns_linear_tc_workspace(N, K, ldw, w_is_kn, precision);
        MG_REQUIRE(workspace && workspace_floats >= need && aligned16(workspace),
                   "linear_tc: 3xTF32 needs %lld floats of aligned workspace", (long long)need);
        if (int rc = split_weight(W, N, K, ldw, w_is_kn, workspace, st)) return rc;
        w_hi = workspace;
        w_lo = workspace + need / 2;
    }
    return linear_tc_presplit(A, lda, w_hi, w_lo, ldw, w_is_kn, bias, act, slope, M, N, K, precision, C, ldc, st);
}

// C[M,N] = A[K,M]^T . B[K,N]  (C is overwritten; partial K slices are accumulated with fp32 atomics, so the
// summation order — not the value beyond fp32 rounding — varies run to run).  precision: 0 = TF32, 1 = 3xTF32.
extern "C" int mgnns_wgrad_tc(const float* A, int64_t lda, const float* B, int64_t ldb, int M, int N, int K, int precision,
                              float* C, int64_t ldc, void* stream) {
    MG_REQUIRE(M >= 1 && N >= 1 && K >= 0, "wgrad_tc: bad dimensions");
    MG_REQUIRE(A && B && C, "wgrad_tc: null pointer");
    MG_REQUIRE(lda >= M && ldb >= N && ldc >= N, "wgrad_tc: leading dimension too small");
    MG_REQUIRE((lda % 4) == 0 && (ldb % 4) == 0, "wgrad_tc: lda=%lld and ldb=%lld must be multiples of 4 (16-byte TMA strides)",
               (long long)lda, (long long)ldb);
    MG_REQUIRE(aligned16(A) && aligned16(B), "wgrad_tc: A and B must be 16-byte aligned");
    cudaStream_t st = as_stream(stream);
    cudaError_t e = cudaMemset2DAsync(C, (size_t)ldc * sizeof(float), 0, (size_t)N * sizeof(float), (size_t)M, st);
    MG_REQUIRE(e == cudaSuccess, "wgrad_tc: memset failed: %s", cudaGetErrorString(e));
    if (K == 0) return 0;
    WgradParams p{};
    p.M = M; p.N = N; p.K = K;
    p.C = C; p.ldc = ldc;
    p.m_tiles = (M + L_BM - 1) / L_BM;
    p.n_tiles = (N + L_BN_MAX - 1) / L_BN_MAX;
    p.BN = (((N + p.n_tiles - 1) / p.n_tiles) + 31) / 32 * 32;
    p.k_chunks = (K + LK - 1) / LK;
    const int tiles = p.m_tiles * p.n_tiles;
    int splits = (2 * sm_count() + tiles - 1) / tiles;               // about two work items per SM
    if (splits > p.k_chunks / 4) splits = p.k_chunks / 4;             // at least four K chunks per item
    if (splits < 1) splits = 1;
    p.chunks_per_split = (p.k_chunks + splits - 1) / splits;
    p.k_splits = (p.k_chunks + p.chunks_per_split - 1) / p.chunks_per_split;
    p.n_items = tiles * p.k_splits;
    CUtensorMap ma, mb;
    {
        uint64_t dims[2] = {(uint64_t)M, (uint64_t)K};
        uint64_t str[1] = {(uint64_t)lda * 4};
        uint32_t box[2] = {32, LK};
        if (int rc = make_map(&ma, A, 2, dims, str, box, true)) return rc;
    }
    {
        uint64_t dims[2] = {(uint64_t)N, (uint64_t)K};
        uint64_t str[1] = {(uint64_t)ldb * 4};
        uint32_t box[2] = {32, LK};
        if (int rc = make_map(&mb, B, 2, dims, str, box, true)) return rc;
    }
    return precision ? launch_wgrad<true>(ma, mb, p, st) : launch_wgrad<false>(ma, mb, p, st);
}
