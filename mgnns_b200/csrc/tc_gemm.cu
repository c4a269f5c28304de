// tcgen05 / TMEM / TMA contractions for the two streaming GEMMs of the image channel
// (ref: nn.Linear(2048,300) over the 196 positions of the trunk output,
//  models/Multi_GCN_Multihead_att.py:400-428, and its weight gradient).
//
//   FWD : bank_b[p, o]  = sum_c F_b[c, p] * W[o, c] + bias[o]          (M = p, N = o, K = c)
//         A = F_b  read straight from the NCHW map (M-contiguous -> MN-major UMMA operand),
//         B = W    (K-major).  One work item = (sample, 128-row tile of the 196 positions).
//   DW  : gW[o, c]     += sum_{b,p} F_b[c, p] * gbank_b[p, o]           (M = c, N = o, K = (b,p))
//         A = F_b (K-major), B = gbank_b (N-contiguous -> MN-major).  One work item =
//         (128-channel tile, group of samples); fp32 atomicAdd of the partial tile.
//
// Pipeline (one CTA per SM, persistent): warp 0 = TMA producer, warp 1 = MMA issuer (one elected
// lane), warp 2 = TMEM allocator, warps 4-7 = operand splitter, warps 8-11 = epilogue
// (tcgen05.ld -> bias -> global).  Operands are fp32 in shared memory (128-byte swizzle written by
// TMA, out-of-bounds rows/columns zero-filled), consumed as kind::tf32.
//
// Precision modes: SPLIT=false is plain TF32 (10-bit mantissa operands).  SPLIT=true is the
// "3xTF32" scheme: the splitter warps rewrite each landed tile in place as hi = x & ~0x1fff (exactly
// representable in tf32) and write lo = x - hi to a second buffer; the issuer accumulates
// hi*hi + lo*hi + hi*lo in the fp32 TMEM accumulator, which restores ~2^-21 relative accuracy
// (fp32-class) at three times the tensor work.
#include <atomic>
#include "tc_common.cuh"

namespace mgnns {
namespace tc {

// order-preserving float -> uint key (larger float <=> larger key); any NaN maps to the top key
__device__ __forceinline__ uint32_t ordered_key(float f) {
    const uint32_t u = __float_as_uint(f);
    const uint32_t k = u ^ ((uint32_t)((int32_t)u >> 31) | 0x80000000u);
    return ((u << 1) > 0xFF000000u) ? 0xFFC00000u : k;
}

constexpr int KCHUNK = 32;               // fp32 elements per 128-byte swizzle row
constexpr int UMMA_K = 8;                // tf32
constexpr int NTHREADS = 384;
constexpr int FWD = 0, DW = 1;

struct Params {
    int B, C, P, O;                      // samples, channels (2048), positions (196), outputs (300)
    const float* bias;                   // FWD
    float* out;                          // FWD: bank [B,P,O];  DW: gW [O,C]
    int n_items;                         // FWD: ceil(B * bps / 4);   DW: c_tiles * groups
    int bps;                             // FWD: 32-position boxes per sample, ceil(P/32)
    uint32_t* pooled_ord;                // FWD + SPLIT: [B,C] running spatial max as order-preserving uints (or NULL)
    int* counter;                        // dynamic tile scheduler: next work item (zeroed before the launch)
    int groups, samples_per_group;       // DW
};

template <int PROBLEM> struct Geo;
template <> struct Geo<FWD> {
    static constexpr int A_BYTES = 128 * KCHUNK * 4;          // 4 boxes of 32(p) x 32(c): MN-major
    static constexpr int BN = 304;                            // 300 outputs padded to a multiple of 16
    static constexpr int B_BYTES = BN * KCHUNK * 4;           // 2 boxes of 32(c) x 152(o): K-major
    static constexpr int N0 = 160, N1 = 144;
    static constexpr int A_MN = 1, B_MN = 0;
};
template <> struct Geo<DW> {
    static constexpr int A_BYTES = 128 * KCHUNK * 4;          // 1 box of 32(p) x 128(c): K-major
    static constexpr int BN = 320;
    static constexpr int B_BYTES = BN * KCHUNK * 4;           // 10 boxes of 32(o) x 32(p): MN-major
    static constexpr int N0 = 160, N1 = 160;
    static constexpr int A_MN = 0, B_MN = 1;
};

template <int PROBLEM, bool SPLIT>
struct Cfg {
    using G = Geo<PROBLEM>;
    static constexpr int STAGE_BYTES = (G::A_BYTES + G::B_BYTES) * (SPLIT ? 2 : 1);
    static constexpr int STAGES = (226 * 1024 - 1024) / STAGE_BYTES >= 4 ? 4 : (226 * 1024 - 1024) / STAGE_BYTES;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 512 /*barriers + scheduler ring*/;
};

template <int PROBLEM, bool SPLIT>
__global__ void __launch_bounds__(NTHREADS, 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmBlo, Params p) {
    // FWD with SPLIT: W is split once per call in global memory (hi via tmB, lo via tmBlo); only A is split here.
    constexpr bool B_PRESPLIT = SPLIT && PROBLEM == FWD;
    using G = Geo<PROBLEM>;
    using CF = Cfg<PROBLEM, SPLIT>;
    constexpr int STAGES = CF::STAGES;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * CF::STAGE_BYTES);
    uint64_t* full = bars;                       // TMA bytes landed
    uint64_t* ready = bars + STAGES;             // (SPLIT) hi/lo tiles written
    uint64_t* empty = bars + 2 * STAGES;         // MMAs that read the stage have completed
    uint64_t* tmem_full = bars + 3 * STAGES;     // accumulator complete
    uint64_t* tmem_empty = tmem_full + 1;        // accumulator drained by the epilogue
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 1);
    SchedSmem* sched = reinterpret_cast<SchedSmem*>(smem + STAGES * CF::STAGE_BYTES + 256);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    auto stageA = [&](int s) { return smem + s * CF::STAGE_BYTES; };
    auto stageB = [&](int s) { return smem + s * CF::STAGE_BYTES + G::A_BYTES; };
    auto stageAlo = [&](int s) { return smem + s * CF::STAGE_BYTES + G::A_BYTES + G::B_BYTES; };
    auto stageBlo = [&](int s) { return smem + s * CF::STAGE_BYTES + 2 * G::A_BYTES + G::B_BYTES; };

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
        mbar_init(tmem_full, 1);
        mbar_init(tmem_empty, 4);
        sched_init(sched);
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // k-iterations of one work item
    const int p_chunks = (p.P + KCHUNK - 1) / KCHUNK;
    auto item_kiters = [&](int item) -> int {
        if (PROBLEM == FWD) return p.C / KCHUNK;
        const int g = item % p.groups;
        const int b0 = g * p.samples_per_group;
        const int nb = max(0, min(p.samples_per_group, p.B - b0));
        return nb * p_chunks;
    };

    if (warp == 0) {
        // ===================================================== TMA producer
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            SchedState ss;
            for (;;) {
                const int item = sched_produce(sched, ss, p.counter);
                if (item >= p.n_items) break;
                const int kiters = item_kiters(item);
                for (int kk = 0; kk < kiters; ++kk) {
                    mbar_wait(&empty[stage], phase ^ 1);
                    mbar_expect_tx(&full[stage], G::A_BYTES + G::B_BYTES * (B_PRESPLIT ? 2 : 1));
                    uint8_t* a = stageA(stage);
                    uint8_t* b = stageB(stage);
                    if (PROBLEM == FWD) {
                        const int c0 = kk * KCHUNK;
                        // A: four 32(p) x 32(c) boxes -> MN-major atoms 4 KB apart.  The M tile is four consecutive
                        // boxes of the (sample, 32-position box) sequence, so it may straddle two samples; boxes past
                        // the last sample and positions >= P are zero-filled by TMA.
                        for (int j = 0; j < 4; ++j) {
                            const int g = item * 4 + j;
                            const int smp = g / p.bps, pb = g - smp * p.bps;
                            tma_load_3d(a + j * 4096, &tmA, &full[stage], pb * 32, c0, smp);
                        }
                        // B: two 152-row boxes of W[o, c0:c0+32]
                        tma_load_2d(b, &tmB, &full[stage], c0, 0);
                        tma_load_2d(b + 152 * 128, &tmB, &full[stage], c0, 152);
                        if (B_PRESPLIT) {
                            uint8_t* bl = stageBlo(stage);
                            tma_load_2d(bl, &tmBlo, &full[stage], c0, 0);
                            tma_load_2d(bl + 152 * 128, &tmBlo, &full[stage], c0, 152);
                        }
                    } else {
                        const int ct = item / p.groups, g = item % p.groups;
                        const int smp = g * p.samples_per_group + kk / p_chunks;
                        const int p0 = (kk % p_chunks) * KCHUNK;
                        // A: F_b[c, p0:p0+32] for 128 channels (K-major)
                        tma_load_3d(a, &tmA, &full[stage], p0, ct * 128, smp);
                        // B: ten 32(o) x 32(p) boxes of gbank_b -> MN-major atoms 4 KB apart
                        for (int j = 0; j < 10; ++j)
                            tma_load_3d(b + j * 4096, &tmB, &full[stage], j * 32, p0, smp);
                    }
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================================================== MMA issuer
        if (lane == 0) {
            constexpr uint32_t idesc0 = instr_desc_tf32(128, G::N0, G::A_MN, G::B_MN);
            constexpr uint32_t idesc1 = instr_desc_tf32(128, G::N1, G::A_MN, G::B_MN);
            // descriptor geometry
            //   K-major  : 8-row groups 1024 B apart (SBO), k-step = +32 B inside the 128 B swizzle row
            //   MN-major : 32-element atoms 4096 B apart (LBO), 4-k-row groups 512 B apart (SBO), k-step (8 rows) = +1024 B,
            //              layout SWIZZLE_128B_BASE32B
            constexpr uint32_t A_LBO = G::A_MN ? 4096 : 16, A_SBO = G::A_MN ? 512 : 1024, A_KSTEP = G::A_MN ? 1024 : 32;
            constexpr uint32_t B_LBO = G::B_MN ? 4096 : 16, B_SBO = G::B_MN ? 512 : 1024, B_KSTEP = G::B_MN ? 1024 : 32;
            constexpr uint32_t A_LT = G::A_MN ? 1 : 2, B_LT = G::B_MN ? 1 : 2;
            // second N part starts N0 rows (K-major: N0*128 B) or N0/32 atoms (MN-major: N0/32*4096 B) further
            constexpr uint32_t B_PART1 = G::B_MN ? (G::N0 / 32) * 4096 : G::N0 * 128;
            int stage = 0;
            uint32_t phase = 0, tphase = 0;
            SchedState ss;
            for (;;) {
                const int item = sched_consume_thread(sched, ss);
                if (item >= p.n_items) break;
                const int kiters = item_kiters(item);
                mbar_wait(tmem_empty, tphase ^ 1);
                tc_fence_after();
                uint32_t accumulate = 0;
                for (int kk = 0; kk < kiters; ++kk) {
                    mbar_wait(SPLIT ? &ready[stage] : &full[stage], phase);
                    tc_fence_after();
                    const uint32_t a_hi = smem_u32(stageA(stage)), b_hi = smem_u32(stageB(stage));
                    const uint32_t a_lo = smem_u32(stageAlo(stage)), b_lo = smem_u32(stageBlo(stage));
#pragma unroll
                    for (int ks = 0; ks < KCHUNK / UMMA_K; ++ks) {
                        const uint64_t dah = smem_desc(a_hi + ks * A_KSTEP, A_LBO, A_SBO, A_LT);
                        const uint64_t dbh0 = smem_desc(b_hi + ks * B_KSTEP, B_LBO, B_SBO, B_LT);
                        const uint64_t dbh1 = smem_desc(b_hi + B_PART1 + ks * B_KSTEP, B_LBO, B_SBO, B_LT);
                        umma_tf32(tmem_base, dah, dbh0, idesc0, accumulate);
                        umma_tf32(tmem_base + G::N0, dah, dbh1, idesc1, accumulate);
                        accumulate = 1;
                        if (SPLIT) {
                            const uint64_t dal = smem_desc(a_lo + ks * A_KSTEP, A_LBO, A_SBO, A_LT);
                            const uint64_t dbl0 = smem_desc(b_lo + ks * B_KSTEP, B_LBO, B_SBO, B_LT);
                            const uint64_t dbl1 = smem_desc(b_lo + B_PART1 + ks * B_KSTEP, B_LBO, B_SBO, B_LT);
                            umma_tf32(tmem_base, dal, dbh0, idesc0, 1);
                            umma_tf32(tmem_base + G::N0, dal, dbh1, idesc1, 1);
                            umma_tf32(tmem_base, dah, dbl0, idesc0, 1);
                            umma_tf32(tmem_base + G::N0, dah, dbl1, idesc1, 1);
                        }
                    }
                    umma_commit(&empty[stage]);          // frees the stage once these MMAs have read it
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                if (kiters > 0) umma_commit(tmem_full);
                else mbar_arrive(tmem_full);
                tphase ^= 1;
            }
        }
    } else if (warp >= 4 && warp < 8) {
        // ===================================================== operand splitter (3xTF32 only)
        {
            const int t = threadIdx.x - 128;
            int stage = 0;
            uint32_t phase = 0;
            uint32_t pool_o[8];
            int pool_smp[4], pool_lim[4], pool_c0 = 0;
            bool pool_pending = false;
            SchedState ss;
            for (;;) {
                const int item = sched_consume_warp(sched, ss, lane);
                if (item >= p.n_items) break;
                if (!SPLIT) continue;                    // plain TF32: nothing to split, only keep the scheduler ring moving
                const int kiters = item_kiters(item);
                if (PROBLEM == FWD && p.pooled_ord != nullptr) {
#pragma unroll
                    for (int box = 0; box < 4; ++box) {
                        const int g = item * 4 + box;
                        const int smp = g / p.bps, pb = g - smp * p.bps;
                        pool_smp[box] = smp;
                        pool_lim[box] = (smp < p.B) ? min(32, p.P - pb * 32) : 0;     // valid positions of the box
                    }
                }
                for (int kk = 0; kk < kiters; ++kk) {
                    mbar_wait(&full[stage], phase);
                    float4* hi = reinterpret_cast<float4*>(stageA(stage));
                    float4* lo = reinterpret_cast<float4*>(stageAlo(stage));
                    if (PROBLEM == FWD && p.pooled_ord != nullptr) {
                        // Fused 14x14 global max pool (ref: nn.MaxPool2d(14,14), model:302,:454,:486): this pass touches
                        // every element of the feature map anyway.  Float4 `idx` of the A region is box idx/256 (32
                        // positions), channel row (idx%256)/8, physical 16-byte chunk idx%8; SWIZZLE_128B with 32-byte
                        // atoms stores logical atom a of row r at physical atom a ^ (r & 3).  Each thread keeps one
                        // order-preserving uint per float4 (NaN maps above +inf, like torch's max pooling); the
                        // shuffles and atomics that finish the reduction run AFTER the stage has been handed to the
                        // MMA issuer, off the critical path.
                        const int qphys = t & 7;
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const int idx = t + i * 128;
                            const float4 x = hi[idx];
                            float4 h, l;
                            h.x = __uint_as_float(__float_as_uint(x.x) & 0xFFFFE000u); l.x = x.x - h.x;
                            h.y = __uint_as_float(__float_as_uint(x.y) & 0xFFFFE000u); l.y = x.y - h.y;
                            h.z = __uint_as_float(__float_as_uint(x.z) & 0xFFFFE000u); l.z = x.z - h.z;
                            h.w = __uint_as_float(__float_as_uint(x.w) & 0xFFFFE000u); l.w = x.w - h.w;
                            hi[idx] = h;
                            lo[idx] = l;
                            const int r = (t >> 3) + 16 * (i & 1);
                            const int qlog = (((qphys >> 1) ^ (r & 3)) << 1) | (qphys & 1);
                            pool_o[i] = (4 * qlog < pool_lim[i >> 1])
                                            ? max(max(ordered_key(x.x), ordered_key(x.y)), max(ordered_key(x.z), ordered_key(x.w)))
                                            : 0u;
                        }
                        pool_pending = true;
                        pool_c0 = kk * KCHUNK;
                    } else {
                        // A and B are contiguous, so are the lo buffers; a pre-split B needs no work here
                        constexpr int NV = (G::A_BYTES + (B_PRESPLIT ? 0 : G::B_BYTES)) / 16;
#pragma unroll 4
                        for (int i = t; i < NV; i += 128) {
                            float4 x = hi[i];
                            float4 h, l;
                            h.x = __uint_as_float(__float_as_uint(x.x) & 0xFFFFE000u); l.x = x.x - h.x;
                            h.y = __uint_as_float(__float_as_uint(x.y) & 0xFFFFE000u); l.y = x.y - h.y;
                            h.z = __uint_as_float(__float_as_uint(x.z) & 0xFFFFE000u); l.z = x.z - h.z;
                            h.w = __uint_as_float(__float_as_uint(x.w) & 0xFFFFE000u); l.w = x.w - h.w;
                            hi[i] = h;
                            lo[i] = l;
                        }
                    }
                    fence_proxy_async();                 // generic-proxy writes -> visible to the tensor core's async proxy
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&ready[stage]);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                    if (PROBLEM == FWD && pool_pending) {
                        // boxes of the same sample are merged in registers first (a tile spans at most a few samples);
                        // then the eight lanes of a channel row reduce and one lane issues a fire-and-forget atomicMax
#pragma unroll
                        for (int rp = 0; rp < 2; ++rp) {
                            const int r = (t >> 3) + 16 * rp;
                            uint32_t acc = 0;
                            int acc_smp = pool_smp[0];
#pragma unroll
                            for (int box = 0; box < 4; ++box) {
                                if (pool_smp[box] != acc_smp) {                     // warp-uniform
                                    acc = max(acc, __shfl_xor_sync(0xffffffffu, acc, 1));
                                    acc = max(acc, __shfl_xor_sync(0xffffffffu, acc, 2));
                                    acc = max(acc, __shfl_xor_sync(0xffffffffu, acc, 4));
                                    if ((t & 7) == 0 && acc != 0) atomicMax(p.pooled_ord + (int64_t)acc_smp * p.C + pool_c0 + r, acc);
                                    acc = 0;
                                    acc_smp = pool_smp[box];
                                }
                                acc = max(acc, pool_o[box * 2 + rp]);
                            }
                            acc = max(acc, __shfl_xor_sync(0xffffffffu, acc, 1));
                            acc = max(acc, __shfl_xor_sync(0xffffffffu, acc, 2));
                            acc = max(acc, __shfl_xor_sync(0xffffffffu, acc, 4));
                            if ((t & 7) == 0 && acc != 0) atomicMax(p.pooled_ord + (int64_t)acc_smp * p.C + pool_c0 + r, acc);
                        }
                        pool_pending = false;
                    }
                }
            }
        }
    } else if (warp >= 8) {
        // ===================================================== epilogue (TMEM -> registers -> global)
        const int q = warp & 3;                          // TMEM lane quarter this warp may read
        const int row = q * 32 + lane;
        uint32_t tphase = 0;
        SchedState ss;
        for (;;) {
            const int item = sched_consume_warp(sched, ss, lane);
            if (item >= p.n_items) break;
            mbar_wait(tmem_full, tphase);
            tc_fence_after();
            const int kiters = item_kiters(item);
            if (kiters > 0) {
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
                if (PROBLEM == FWD) {
                    // this warp's 32 TMEM lanes are exactly one 32-position box
                    const int g = item * 4 + q;
                    const int smp = g / p.bps, pb = g - smp * p.bps;
                    const int pp = (smp < p.B) ? pb * 32 + lane : p.P;
                    float* dst = p.out + ((int64_t)smp * p.P + pp) * p.O;
                    for (int c0 = 0; c0 < G::BN; c0 += 16) {
                        uint32_t r[16];
                        tmem_ld16(taddr + c0, r);
                        tmem_ld_wait();
                        if (pp < p.P) {
#pragma unroll
                            for (int j = 0; j < 16; j += 4) {
                                const int o = c0 + j;
                                if (o + 3 < p.O) {
                                    float4 v;
                                    v.x = __uint_as_float(r[j + 0]) + __ldg(p.bias + o + 0);
                                    v.y = __uint_as_float(r[j + 1]) + __ldg(p.bias + o + 1);
                                    v.z = __uint_as_float(r[j + 2]) + __ldg(p.bias + o + 2);
                                    v.w = __uint_as_float(r[j + 3]) + __ldg(p.bias + o + 3);
                                    *reinterpret_cast<float4*>(dst + o) = v;
                                } else {
                                    for (int e = 0; e < 4; ++e)
                                        if (o + e < p.O) dst[o + e] = __uint_as_float(r[j + e]) + __ldg(p.bias + o + e);
                                }
                            }
                        }
                    }
                } else {
                    const int ct = item / p.groups;
                    const int c = ct * 128 + row;
                    for (int c0 = 0; c0 < G::BN; c0 += 16) {
                        uint32_t r[16];
                        tmem_ld16(taddr + c0, r);
                        tmem_ld_wait();
                        if (c < p.C) {
#pragma unroll
                            for (int j = 0; j < 16; ++j) {
                                const int o = c0 + j;
                                if (o < p.O) atomicAdd(p.out + (int64_t)o * p.C + c, __uint_as_float(r[j]));
                            }
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tmem_empty);
            tphase ^= 1;
        }
    }

    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 2) tmem_dealloc(tmem_base, 512);
}


// Work counters for the dynamic tile scheduler (tc_common.cuh).  Eager launches cycle through a 256-slot ring;
// launches that are being captured into a CUDA graph get a slot of their own from a second pool that eager
// launches never touch (the slot address is baked into the graph and re-zeroed by the captured memset node on
// every replay), so a replaying graph and an eager kernel on another stream cannot share a counter.  Symbol
// addresses are per device, so the base pointers are looked up per device.
constexpr unsigned EAGER_SLOTS = 256, GRAPH_SLOTS = 4096;
__device__ int g_tile_counters[EAGER_SLOTS];
__device__ int g_graph_counters[GRAPH_SLOTS];

int* next_tile_counter(cudaStream_t st) {
    constexpr int MAX_DEV = 64;
    static std::atomic<int*> eager_base[MAX_DEV];
    static std::atomic<int*> graph_base[MAX_DEV];
    static std::atomic<unsigned> next_eager{0}, next_graph{0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= MAX_DEV) return nullptr;
    int* eb = eager_base[dev].load(std::memory_order_acquire);
    if (!eb) {
        void *pe = nullptr, *pg = nullptr;
        if (cudaGetSymbolAddress(&pe, g_tile_counters) != cudaSuccess) return nullptr;
        if (cudaGetSymbolAddress(&pg, g_graph_counters) != cudaSuccess) return nullptr;
        graph_base[dev].store(static_cast<int*>(pg), std::memory_order_release);
        eager_base[dev].store(static_cast<int*>(pe), std::memory_order_release);
        eb = static_cast<int*>(pe);
    }
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(st, &cap) != cudaSuccess) return nullptr;
    int* c;
    if (cap == cudaStreamCaptureStatusActive)
        c = graph_base[dev].load(std::memory_order_acquire) + (next_graph.fetch_add(1, std::memory_order_relaxed) % GRAPH_SLOTS);
    else
        c = eb + (next_eager.fetch_add(1, std::memory_order_relaxed) % EAGER_SLOTS);
    if (cudaMemsetAsync(c, 0, sizeof(int), st) != cudaSuccess) return nullptr;
    return c;
}

// order-preserving uint -> float, in place (the fused max pool accumulates with integer atomicMax)
__global__ void ordered_to_float_kernel(uint32_t* __restrict__ v, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const uint32_t o = v[i];
        v[i] = (o & 0x80000000u) ? (o & 0x7fffffffu) : ~o;
    }
}

// W -> (hi, lo) with hi exactly representable in TF32 (done once per call for the forward's weight operand)
__global__ void split_weight_tf32_kernel(const float* __restrict__ w, float* __restrict__ hi, float* __restrict__ lo, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        const float x = w[i];
        const float h = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
        hi[i] = h;
        lo[i] = x - h;
    }
}

template <int PROBLEM, bool SPLIT>
static int launch(const CUtensorMap& a, const CUtensorMap& b, const CUtensorMap& blo, const Params& p, cudaStream_t st,
                  int max_ctas = 0) {
    using CF = Cfg<PROBLEM, SPLIT>;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(tc_gemm_kernel<PROBLEM, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             CF::SMEM_BYTES);
        MG_REQUIRE(e == cudaSuccess, "tc_gemm: cannot reserve %d bytes of shared memory: %s", CF::SMEM_BYTES,
                   cudaGetErrorString(e));
        configured = true;
    }
    int grid = tc_grid_limit();
    if (max_ctas > 0 && grid > max_ctas) grid = max_ctas;
    if (grid > p.n_items) grid = p.n_items;
    Params q = p;
    q.counter = next_tile_counter(st);
    MG_REQUIRE(q.counter != nullptr, "tc_gemm: cannot set up the tile counter");
    tc_gemm_kernel<PROBLEM, SPLIT><<<grid, NTHREADS, CF::SMEM_BYTES, st>>>(a, b, blo, q);
    MG_LAUNCH_CHECK("tc_gemm");
    return 0;
}

}  // namespace tc
}  // namespace mgnns

using namespace mgnns;
using namespace mgnns::tc;

// bank[B,P,O] = fmap[B,C,P]^T . weight[O,C]^T + bias ; precision: 0 = tf32, 1 = 3xTF32 (fp32-class; needs
// 2*O*C floats of 16-byte aligned workspace for the split weight).  pooled (optional, [B,C], 3xTF32 only):
// global spatial max of the feature map, fused into the operand pass.
extern "C" int mgnns_imgbank_fwd_tc(const float* fmap, const float* weight, const float* bias, int B, int C, int P, int O,
                                    int precision, float* workspace, float* pooled, float* bank, void* stream) {
    return mgnns_imgbank_fwd_tc_capped(fmap, weight, bias, B, C, P, O, precision, workspace, pooled, bank, 0, stream);
}

// max_ctas > 0 caps the persistent grid: the kernel's CTAs take a whole SM each (225 KB of shared memory), so a grid
// on every SM blocks every small kernel of a concurrent stream for its whole duration; a few SMs left free keep the
// latency-bound chains of the training step moving
extern "C" int mgnns_imgbank_fwd_tc_capped(const float* fmap, const float* weight, const float* bias, int B, int C, int P,
                                           int O, int precision, float* workspace, float* pooled, float* bank, int max_ctas,
                                           void* stream) {
    MG_REQUIRE(B >= 0 && C >= 32 && P >= 1 && O >= 1, "imgbank_fwd_tc: bad dimensions");
    MG_REQUIRE(C % KCHUNK == 0, "imgbank_fwd_tc: C=%d must be a multiple of 32", C);
    MG_REQUIRE(O <= Geo<FWD>::BN && O % 4 == 0, "imgbank_fwd_tc: O=%d must be <= 304 and a multiple of 4", O);
    MG_REQUIRE(P % 4 == 0, "imgbank_fwd_tc: P=%d must be a multiple of 4 (16-byte TMA strides)", P);
    if (B == 0) return 0;
    MG_REQUIRE(fmap && weight && bias && bank, "imgbank_fwd_tc: null pointer");
    MG_REQUIRE(aligned16(fmap) && aligned16(weight) && aligned16(bank), "imgbank_fwd_tc: operands must be 16-byte aligned");
    MG_REQUIRE(pooled == nullptr || precision == 1, "imgbank_fwd_tc: the fused max pool needs the 3xTF32 mode (operand splitter pass)");
    cudaStream_t st = as_stream(stream);
    if (pooled) {
        cudaError_t e = cudaMemsetAsync(pooled, 0, (size_t)B * C * sizeof(float), st);
        MG_REQUIRE(e == cudaSuccess, "imgbank_fwd_tc: memset failed: %s", cudaGetErrorString(e));
    }
    const float* w_hi = weight;
    const float* w_lo = weight;
    if (precision) {
        MG_REQUIRE(workspace && aligned16(workspace), "imgbank_fwd_tc: 3xTF32 needs 2*O*C floats of aligned workspace");
        const int64_t n = (int64_t)O * C;
        int blocks = (int)((n + 255) / 256);
        if (blocks > 148 * 8) blocks = 148 * 8;
        split_weight_tf32_kernel<<<blocks, 256, 0, st>>>(weight, workspace, workspace + n, n);
        MG_LAUNCH_CHECK("split_weight_tf32");
        w_hi = workspace;
        w_lo = workspace + n;
    }
    CUtensorMap ma, mb, mbl;
    {
        uint64_t dims[3] = {(uint64_t)P, (uint64_t)C, (uint64_t)B};
        uint64_t str[2] = {(uint64_t)P * 4, (uint64_t)C * P * 4};
        uint32_t box[3] = {32, 32, 1};
        if (int rc = make_map(&ma, fmap, 3, dims, str, box, true)) return rc;
    }
    {
        uint64_t dims[2] = {(uint64_t)C, (uint64_t)O};
        uint64_t str[1] = {(uint64_t)C * 4};
        uint32_t box[2] = {32, 152};
        if (int rc = make_map(&mb, w_hi, 2, dims, str, box)) return rc;
        if (int rc = make_map(&mbl, w_lo, 2, dims, str, box)) return rc;
    }
    Params p{};
    p.B = B; p.C = C; p.P = P; p.O = O;
    p.bias = bias; p.out = bank;
    p.bps = (P + 31) / 32;
    const int64_t items = ((int64_t)B * p.bps + 3) / 4;
    MG_REQUIRE(items < (1LL << 29), "imgbank_fwd_tc: batch too large");
    p.n_items = (int)items;
    p.groups = 1; p.samples_per_group = 1;
    p.pooled_ord = reinterpret_cast<uint32_t*>(pooled);
    if (int rc = precision ? launch<FWD, true>(ma, mb, mbl, p, st, max_ctas) : launch<FWD, false>(ma, mb, mbl, p, st, max_ctas)) return rc;
    if (pooled) {
        const int64_t n = (int64_t)B * C;
        ordered_to_float_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(p.pooled_ord, n);
        MG_LAUNCH_CHECK("ordered_to_float");
    }
    return 0;
}

// gW[O,C] += sum_b gbank_b^T . fmap_b^T   (gW must be initialised by the caller)
extern "C" int mgnns_imgbank_dw_tc(const float* fmap, const float* gbank, int B, int C, int P, int O, int precision,
                                   float* gW, void* stream) {
    return mgnns_imgbank_dw_tc_capped(fmap, gbank, B, C, P, O, precision, gW, 0, stream);
}

// max_ctas > 0 caps the persistent grid (the dynamic tile scheduler makes the grid size a free parameter): a weight
// gradient that is off the critical path can run on a share of the SMs and leave the rest to the latency-bound chain
extern "C" int mgnns_imgbank_dw_tc_capped(const float* fmap, const float* gbank, int B, int C, int P, int O, int precision,
                                          float* gW, int max_ctas, void* stream) {
    MG_REQUIRE(B >= 0 && C >= 1 && P >= 1 && O >= 1, "imgbank_dw_tc: bad dimensions");
    MG_REQUIRE(O <= Geo<DW>::BN && O % 4 == 0, "imgbank_dw_tc: O=%d must be <= 320 and a multiple of 4", O);
    MG_REQUIRE(P % 4 == 0, "imgbank_dw_tc: P=%d must be a multiple of 4 (16-byte TMA strides)", P);
    if (B == 0) return 0;
    MG_REQUIRE(fmap && gbank && gW, "imgbank_dw_tc: null pointer");
    MG_REQUIRE(aligned16(fmap) && aligned16(gbank), "imgbank_dw_tc: operands must be 16-byte aligned");
    CUtensorMap ma, mb;
    {
        uint64_t dims[3] = {(uint64_t)P, (uint64_t)C, (uint64_t)B};
        uint64_t str[2] = {(uint64_t)P * 4, (uint64_t)C * P * 4};
        uint32_t box[3] = {32, 128, 1};
        if (int rc = make_map(&ma, fmap, 3, dims, str, box)) return rc;
    }
    {
        uint64_t dims[3] = {(uint64_t)O, (uint64_t)P, (uint64_t)B};
        uint64_t str[2] = {(uint64_t)O * 4, (uint64_t)P * O * 4};
        uint32_t box[3] = {32, 32, 1};
        if (int rc = make_map(&mb, gbank, 3, dims, str, box, true)) return rc;
    }
    Params p{};
    p.B = B; p.C = C; p.P = P; p.O = O;
    p.bias = nullptr; p.out = gW;
    const int c_tiles = (C + 127) / 128;
    // about four work items per CTA: the dynamic scheduler then absorbs CTAs that get their SM late (a concurrent
    // kernel of another stream) at the price of 4x the epilogue atomics (still < 5 % of an item's MMA time)
    int groups = 4 * sm_count() / c_tiles;
    if (groups < 1) groups = 1;
    if (groups > B) groups = B;
    p.groups = groups;
    p.samples_per_group = (B + groups - 1) / groups;
    p.bps = 1;
    p.pooled_ord = nullptr;
    p.n_items = c_tiles * groups;
    cudaStream_t st = as_stream(stream);
    return precision ? launch<DW, true>(ma, mb, mb, p, st, max_ctas) : launch<DW, false>(ma, mb, mb, p, st, max_ctas);
}
