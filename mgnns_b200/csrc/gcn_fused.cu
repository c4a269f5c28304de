// Fused graph-convolution layer  Y[b] = act((Â · X[b]) · W + bias)  for batched node features
// (ref: GraphConvolution.forward, models/Multi_GCN_Multihead_att.py:52-58: support = X·W; output = Â·support — the
//  same product re-associated, Â applied on the narrower side; activation as the caller applies it, :470-472.
//  BASELINE.json cfg 2: N = 10,000 word graph, 300 -> 512, batch 256).
//
// One persistent CTA per SM; a work item is (sample b, tile of 128 output rows).  The aggregated rows
// Z = Â·X[b] never exist in HBM: sixteen gather warps accumulate them 32 features (one 128-byte swizzle row) at a
// time straight into the shared-memory A operand of the tensor core — already split into the 3xTF32 hi / lo
// halves — while one elected thread issues tcgen05.mma against the weight, which TMA streams from L2 in 8-row
// K slices (MN-major, SWIZZLE_128B_BASE32B atoms), and four epilogue warps drain the 128 x N fp32 accumulator
// from TMEM (bias + activation, 128-bit stores).
//
//   warp 0      work scheduler (atomic counter -> smem ring) + TMA producer for W (hi and lo, one K=8 slice per stage)
//   warp 1      MMA issuer (one lane): per K=8 slice and per 256-column half: hi*hi + lo*hi + hi*lo
//   warp 2      TMEM allocator (512 columns = the whole 128 x 512 fp32 accumulator)
//   warps 4-7   epilogue
//   warps 8-23  gather: an item's rows are cut into segments of at most 128 edges (host-built plan), segments are
//               dealt round-robin to the warps, 8 lanes cover the 128 bytes of one neighbour row piece and the four
//               8-lane groups of a warp take every fourth edge (eight 128-bit loads in flight per lane); rows that
//               span several segments are combined through a shared-memory fp32 scratch tile with red.shared.
//
// Tiles are built from rows dealt by degree rank (tile t holds the rows ranked t, t+T, t+2T, ...), so every tile
// carries the same share of the edges and the hub rows of a power-law graph are spread over all tiles.
#include "tc_common.cuh"

namespace mgnns {
namespace tc {

constexpr int F_BM = 128;
constexpr int F_THREADS = 768;
constexpr int F_GATHER_WARPS = 16;
constexpr int F_FIRST_GATHER_WARP = 8;
constexpr int F_A_STAGES = 2;
constexpr int F_B_STAGES = 3;
constexpr int F_A_HALF = F_BM * 32 * 4;                 // 16 KB: 128 rows x 32 floats
constexpr int F_A_STAGE = 2 * F_A_HALF;                 // hi | lo
constexpr int F_N_MAX = 512;
constexpr int F_B_HALF = F_N_MAX * 8 * 4;               // 16 KB: 8 K rows x 512 columns
constexpr int F_B_STAGE = 2 * F_B_HALF;                 // hi | lo
constexpr int F_SCRATCH = F_BM * 32 * 4;                // 16 KB
constexpr int F_OFF_B = F_A_STAGES * F_A_STAGE;
constexpr int F_OFF_SCRATCH = F_OFF_B + F_B_STAGES * F_B_STAGE;
constexpr int F_OFF_BARS = F_OFF_SCRATCH + F_SCRATCH;
constexpr int F_SMEM_BYTES = F_OFF_BARS + 1024 + 1024;   // barriers + scheduler ring, alignment slack
constexpr int F_SEG_EDGES = 128;                        // the host plan never makes a longer segment
constexpr int F_SCHED_CONSUMERS = 1 + 4 + F_GATHER_WARPS;

struct FusedParams {
    const float* X;
    int64_t strideX;                 // elements between samples
    float* Y;
    int64_t ldy, strideY;
    const float* bias;
    int act;
    float slope;
    int K, N, N0, N1;                // N0 + N1 = N, N0 <= 256
    int k_chunks, n_k8;
    int n_tiles, n_items;
    const int* tile_seg_ptr;         // [n_tiles + 1]
    const int4* segs;                // {first edge, edges, row in tile, 1 = the row's only segment}
    const int2* edges;               // {byte offset of the neighbour row inside X[b], bits of the edge value}
    const int* tile_rows;            // [n_tiles * 128] output row or -1
    const int* tile_multi_ptr;       // [n_tiles + 1]
    const int* multi_rows;           // rows in tile that span several segments
    int* counter;
};

__device__ __forceinline__ void gather_bar() { asm volatile("bar.sync 1, %0;" ::"n"(F_GATHER_WARPS * 32) : "memory"); }

__device__ __forceinline__ void split_store(uint8_t* hi, uint8_t* lo, int r, int c, float4 v, bool split) {
    const uint32_t off = (uint32_t)r * 128u + (uint32_t)((c ^ (r & 7)) << 4);      // 128-byte swizzle, K-major
    if (split) {
        float4 h, l;
        h.x = __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u); l.x = v.x - h.x;
        h.y = __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u); l.y = v.y - h.y;
        h.z = __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u); l.z = v.z - h.z;
        h.w = __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u); l.w = v.w - h.w;
        *reinterpret_cast<float4*>(hi + off) = h;
        *reinterpret_cast<float4*>(lo + off) = l;
    } else {
        *reinterpret_cast<float4*>(hi + off) = v;
    }
}

template <bool SPLIT>
__global__ void __launch_bounds__(F_THREADS, 1)
gcn_fused_kernel(const __grid_constant__ CUtensorMap tmBhi, const __grid_constant__ CUtensorMap tmBlo, FusedParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + F_OFF_BARS);
    uint64_t* a_ready = bars;                              // [2] gather warps have written the chunk
    uint64_t* a_empty = bars + F_A_STAGES;                 // [2] MMAs that read the chunk have completed
    uint64_t* b_full = a_empty + F_A_STAGES;               // [3] TMA bytes landed
    uint64_t* b_empty = b_full + F_B_STAGES;               // [3]
    uint64_t* tmem_full = b_empty + F_B_STAGES;
    uint64_t* tmem_empty = tmem_full + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 1);
    SchedSmem* sched = reinterpret_cast<SchedSmem*>(smem + F_OFF_BARS + 256);
    float* scratch = reinterpret_cast<float*>(smem + F_OFF_SCRATCH);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    auto stageA = [&](int s) { return smem + s * F_A_STAGE; };
    auto stageB = [&](int s) { return smem + F_OFF_B + s * F_B_STAGE; };

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmBhi);
        if (SPLIT) prefetch_tmap(&tmBlo);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < F_A_STAGES; ++s) {
            mbar_init(&a_ready[s], F_GATHER_WARPS);
            mbar_init(&a_empty[s], 1);
        }
        for (int s = 0; s < F_B_STAGES; ++s) {
            mbar_init(&b_full[s], 1);
            mbar_init(&b_empty[s], 1);
        }
        mbar_init(tmem_full, 1);
        mbar_init(tmem_empty, 4);
        for (int i = 0; i < SCHED_SLOTS; ++i) {
            mbar_init(&sched->full[i], 1);
            mbar_init(&sched->empty[i], F_SCHED_CONSUMERS);
        }
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(tmem_slot, 512);
    for (int i = threadIdx.x; i < F_SCRATCH / 4; i += F_THREADS) scratch[i] = 0.f;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int nblk = p.N >> 5;                             // 32-column blocks of W, 1 KB each per K=8 slice

    if (warp == 0) {
        // ===================================================== scheduler + TMA producer for the weight slices
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            const uint32_t tx = (uint32_t)nblk * 1024u * (SPLIT ? 2u : 1u);
            SchedState ss;
            for (;;) {
                const int item = sched_produce(sched, ss, p.counter);
                if (item >= p.n_items) break;
                for (int k8 = 0; k8 < p.n_k8; ++k8) {
                    mbar_wait(&b_empty[stage], phase ^ 1);
                    mbar_expect_tx(&b_full[stage], tx);
                    tma_load_2d(stageB(stage), &tmBhi, &b_full[stage], 0, k8 * nblk * 8);
                    if (SPLIT) tma_load_2d(stageB(stage) + F_B_HALF, &tmBlo, &b_full[stage], 0, k8 * nblk * 8);
                    if (++stage == F_B_STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================================================== MMA issuer
        if (lane == 0) {
            const uint32_t idesc0 = instr_desc_tf32(F_BM, p.N0, 0, 1);
            const uint32_t idesc1 = instr_desc_tf32(F_BM, p.N1 > 0 ? p.N1 : 16, 0, 1);
            int as = 0, bs = 0;
            uint32_t aphase = 0, bphase = 0, tphase = 0;
            SchedState ss;
            for (;;) {
                const int item = sched_consume_thread(sched, ss);
                if (item >= p.n_items) break;
                mbar_wait(tmem_empty, tphase ^ 1);
                tc_fence_after();
                uint32_t accumulate = 0;
                for (int kc = 0; kc < p.k_chunks; ++kc) {
                    mbar_wait(&a_ready[as], aphase);
                    tc_fence_after();
                    const uint32_t a_hi = smem_u32(stageA(as)), a_lo = a_hi + F_A_HALF;
                    const int ks_n = min(4, p.n_k8 - kc * 4);
                    for (int ks = 0; ks < ks_n; ++ks) {
                        mbar_wait(&b_full[bs], bphase);
                        tc_fence_after();
                        const uint32_t b_hi = smem_u32(stageB(bs)), b_lo = b_hi + F_B_HALF;
                        const uint64_t dah = smem_desc(a_hi + ks * 32, 16, 1024, 2);
                        const uint64_t dal = smem_desc(a_lo + ks * 32, 16, 1024, 2);
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            if (h == 1 && p.N1 == 0) break;
                            const uint32_t boff = h ? (uint32_t)(p.N0 >> 5) * 1024u : 0u;
                            const uint32_t tacc = tmem_base + (h ? 256u : 0u);
                            const uint32_t idesc = h ? idesc1 : idesc0;
                            const uint64_t dbh = smem_desc(b_hi + boff, 1024, 512, 1);
                            umma_tf32(tacc, dah, dbh, idesc, accumulate);
                            if (SPLIT) {
                                const uint64_t dbl = smem_desc(b_lo + boff, 1024, 512, 1);
                                umma_tf32(tacc, dal, dbh, idesc, 1);
                                umma_tf32(tacc, dah, dbl, idesc, 1);
                            }
                        }
                        accumulate = 1;
                        umma_commit(&b_empty[bs]);
                        if (++bs == F_B_STAGES) { bs = 0; bphase ^= 1; }
                    }
                    umma_commit(&a_empty[as]);
                    if (++as == F_A_STAGES) { as = 0; aphase ^= 1; }
                }
                umma_commit(tmem_full);
                tphase ^= 1;
            }
        }
    } else if (warp >= 4 && warp < 8) {
        // ===================================================== epilogue (TMEM -> registers -> global)
        const int q = warp & 3;
        const int row = q * 32 + lane;
        uint32_t tphase = 0;
        const bool vec_ok = ((p.ldy & 3) == 0) && ((p.strideY & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.Y) & 15u) == 0);
        const uint64_t stream_pol = l2_policy_evict_first();
        SchedState ss;
        for (;;) {
            const int item = sched_consume_warp(sched, ss, lane);
            if (item >= p.n_items) break;
            const int b = item / p.n_tiles, t = item - b * p.n_tiles;
            const int out_row = __ldg(p.tile_rows + t * F_BM + row);
            mbar_wait(tmem_full, tphase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
            float* dst = p.Y + (int64_t)b * p.strideY + (int64_t)(out_row < 0 ? 0 : out_row) * p.ldy;
            for (int c0 = 0; c0 < p.N; c0 += 32) {
                uint32_t r[32];
                tmem_ld16(taddr + c0, *reinterpret_cast<uint32_t(*)[16]>(&r[0]));
                tmem_ld16(taddr + c0 + 16, *reinterpret_cast<uint32_t(*)[16]>(&r[16]));
                tmem_ld_wait();
                if (out_row >= 0) {
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        const int n = c0 + j;
                        float v[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            float x = __uint_as_float(r[j + e]);
                            if (p.bias != nullptr) x += __ldg(p.bias + n + e);
                            v[e] = apply_act(x, p.act, p.slope);
                        }
                        if (vec_ok) {
                            stg4_l2(dst + n, make_float4(v[0], v[1], v[2], v[3]), stream_pol);
                        } else {
#pragma unroll
                            for (int e = 0; e < 4; ++e) dst[n + e] = v[e];
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tmem_empty);
            tphase ^= 1;
        }
    } else if (warp >= F_FIRST_GATHER_WARP) {
        // ===================================================== gather: Z chunk = sum_e val[e] * X[b, col[e], chunk]
        const int gw = warp - F_FIRST_GATHER_WARP;
        const int c = lane & 7, g = lane >> 3;
        const uint64_t keep_pol = l2_policy_evict_last();
        int as = 0;
        uint32_t aphase = 0;
        SchedState ss;
        for (;;) {
            const int item = sched_consume_warp(sched, ss, lane);
            if (item >= p.n_items) break;
            const int b = item / p.n_tiles, t = item - b * p.n_tiles;
            const char* Xb = reinterpret_cast<const char*>(p.X + (int64_t)b * p.strideX);
            const int seg0 = __ldg(p.tile_seg_ptr + t), seg1 = __ldg(p.tile_seg_ptr + t + 1);
            const int mul0 = __ldg(p.tile_multi_ptr + t), mul1 = __ldg(p.tile_multi_ptr + t + 1);
            for (int kc = 0; kc < p.k_chunks; ++kc) {
                mbar_wait(&a_empty[as], aphase ^ 1);
                uint8_t* hi = stageA(as);
                uint8_t* lo = hi + F_A_HALF;
                const bool kvalid = (kc * 32 + c * 4) < p.K;
                const char* Xk = Xb + kc * 128 + c * 16;
                for (int s = seg0 + gw; s < seg1; s += F_GATHER_WARPS) {
                    const int4 sd = __ldg(p.segs + s);
                    const int2* ep = p.edges + sd.x;
                    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
                    for (int base = 0; base < sd.y; base += 32) {
                        int2 ent = make_int2(0, 0);
                        if (base + lane < sd.y) ent = __ldg(ep + base + lane);
                        float4 x[8];
                        float v[8];
#pragma unroll
                        for (int u = 0; u < 8; ++u) {
                            const int idx = 4 * u + g;
                            const int off = __shfl_sync(0xffffffffu, ent.x, idx);
                            v[u] = __int_as_float(__shfl_sync(0xffffffffu, ent.y, idx));
                            x[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                            if (kvalid && base + idx < sd.y)
                                x[u] = ldg4_l2(reinterpret_cast<const float*>(Xk + (uint32_t)off), keep_pol);
                        }
#pragma unroll
                        for (int u = 0; u < 8; ++u) {
                            acc.x = fmaf(v[u], x[u].x, acc.x);
                            acc.y = fmaf(v[u], x[u].y, acc.y);
                            acc.z = fmaf(v[u], x[u].z, acc.z);
                            acc.w = fmaf(v[u], x[u].w, acc.w);
                        }
                    }
                    // the four 8-lane groups hold partial sums over every fourth edge
#pragma unroll
                    for (int o = 8; o <= 16; o <<= 1) {
                        acc.x += __shfl_xor_sync(0xffffffffu, acc.x, o);
                        acc.y += __shfl_xor_sync(0xffffffffu, acc.y, o);
                        acc.z += __shfl_xor_sync(0xffffffffu, acc.z, o);
                        acc.w += __shfl_xor_sync(0xffffffffu, acc.w, o);
                    }
                    if (g == 0) {
                        if (sd.w) {
                            split_store(hi, lo, sd.z, c, acc, SPLIT);
                        } else {
                            float* sp = scratch + sd.z * 32 + c * 4;
                            atomicAdd(sp + 0, acc.x);
                            atomicAdd(sp + 1, acc.y);
                            atomicAdd(sp + 2, acc.z);
                            atomicAdd(sp + 3, acc.w);
                        }
                    }
                }
                if (mul1 > mul0) {
                    gather_bar();                      // every partial sum of this chunk is in the scratch tile
                    for (int m = mul0 + gw * 4 + g; m < mul1; m += F_GATHER_WARPS * 4) {
                        const int r = __ldg(p.multi_rows + m);
                        float4* sp = reinterpret_cast<float4*>(scratch + r * 32 + c * 4);
                        const float4 v = *sp;
                        *sp = make_float4(0.f, 0.f, 0.f, 0.f);
                        split_store(hi, lo, r, c, v, SPLIT);
                    }
                    gather_bar();                      // scratch rows are zero again before the next chunk adds to them
                }
                fence_proxy_async();                   // generic-proxy writes -> visible to the tensor core's async proxy
                __syncwarp();
                if (lane == 0) mbar_arrive(&a_ready[as]);
                if (++as == F_A_STAGES) { as = 0; aphase ^= 1; }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 2) tmem_dealloc(tmem_base, 512);
}

// W[K,N] -> hi / lo in the blocked order the weight slices are fetched in: [k8][n/32][8][32], rows k >= K zero.
__global__ void split_block_weight_kernel(const float* __restrict__ W, int64_t ldw, int K, int N, int n_k8,
                                          float* __restrict__ hi, float* __restrict__ lo, int split) {
    const int64_t total = (int64_t)n_k8 * N * 8;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < total; i += stride) {
        const int k8 = (int)(i / ((int64_t)N * 8));
        const int rem = (int)(i - (int64_t)k8 * N * 8);
        const int nb = rem >> 8, ki = (rem >> 5) & 7, ni = rem & 31;
        const int k = k8 * 8 + ki, n = nb * 32 + ni;
        const float x = (k < K) ? W[(int64_t)k * ldw + n] : 0.f;
        if (split) {
            const float h = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
            hi[i] = h;
            lo[i] = x - h;
        } else {
            hi[i] = x;
        }
    }
}

template <bool SPLIT>
static int launch_fused(const CUtensorMap& bhi, const CUtensorMap& blo, const FusedParams& p, cudaStream_t st) {
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(gcn_fused_kernel<SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, F_SMEM_BYTES);
        MG_REQUIRE(e == cudaSuccess, "gcn_fused: cannot reserve %d bytes of shared memory: %s", F_SMEM_BYTES, cudaGetErrorString(e));
        configured = true;
    }
    int grid = tc_grid_limit();
    if (grid > p.n_items) grid = p.n_items;
    FusedParams q = p;
    q.counter = next_tile_counter(st);
    MG_REQUIRE(q.counter != nullptr, "gcn_fused: cannot set up the tile counter");
    gcn_fused_kernel<SPLIT><<<grid, F_THREADS, F_SMEM_BYTES, st>>>(bhi, blo, q);
    MG_LAUNCH_CHECK("gcn_fused");
    return 0;
}

}  // namespace tc
}  // namespace mgnns

using namespace mgnns;
using namespace mgnns::tc;

// floats of 16-byte aligned workspace mgnns_gcn_fused_tc needs for the blocked (and split) weight
extern "C" int64_t mgnns_gcn_fused_workspace(int N, int K, int precision) {
    const int64_t n_k8 = (K + 7) / 8;
    return n_k8 * N * 8 * (precision ? 2 : 1);
}

extern "C" int mgnns_gcn_fused_tc(const float* X, int64_t strideX, int batch,
                                  const int32_t* tile_seg_ptr, const int32_t* segs, const int32_t* edges,
                                  const int32_t* tile_rows, const int32_t* tile_multi_ptr, const int32_t* multi_rows,
                                  int n_tiles, const float* W, int64_t ldw, const float* bias, int act, float slope,
                                  int K, int N, int precision, float* workspace, int64_t workspace_floats,
                                  float* Y, int64_t ldy, int64_t strideY, void* stream) {
    MG_REQUIRE(batch >= 0 && n_tiles >= 0 && K >= 1 && N >= 1, "gcn_fused: bad dimensions");
    if (batch == 0 || n_tiles == 0) return 0;
    MG_REQUIRE(X && W && Y && tile_seg_ptr && segs && edges && tile_rows && tile_multi_ptr && multi_rows, "gcn_fused: null pointer");
    MG_REQUIRE(K % 4 == 0, "gcn_fused: in_features=%d must be a multiple of 4 (128-bit gathers)", K);
    MG_REQUIRE(N % 32 == 0 && N <= F_N_MAX, "gcn_fused: out_features=%d must be a multiple of 32 and at most %d", N, F_N_MAX);
    MG_REQUIRE(aligned16(X) && (strideX % 4) == 0, "gcn_fused: X must be 16-byte aligned with a sample stride that is a multiple of 4");
    MG_REQUIRE(ldw >= N && ldy >= N, "gcn_fused: leading dimension too small");
    MG_REQUIRE(act == MGNNS_ACT_NONE || act == MGNNS_ACT_RELU || act == MGNNS_ACT_LEAKY, "gcn_fused: bad activation %d", act);
    MG_REQUIRE((int64_t)batch * n_tiles < (1LL << 31), "gcn_fused: too many tiles");
    const int64_t need = mgnns_gcn_fused_workspace(N, K, precision);
    MG_REQUIRE(workspace && workspace_floats >= need && aligned16(workspace), "gcn_fused: needs %lld floats of aligned workspace", (long long)need);
    cudaStream_t st = as_stream(stream);

    FusedParams p{};
    p.X = X; p.strideX = strideX;
    p.Y = Y; p.ldy = ldy; p.strideY = strideY;
    p.bias = bias; p.act = act; p.slope = slope;
    p.K = K; p.N = N;
    p.N0 = N > 256 ? 256 : N;
    p.N1 = N - p.N0;
    p.k_chunks = (K + 31) / 32;
    p.n_k8 = (K + 7) / 8;
    p.n_tiles = n_tiles;
    p.n_items = batch * n_tiles;
    p.tile_seg_ptr = tile_seg_ptr;
    p.segs = reinterpret_cast<const int4*>(segs);
    p.edges = reinterpret_cast<const int2*>(edges);
    p.tile_rows = tile_rows;
    p.tile_multi_ptr = tile_multi_ptr;
    p.multi_rows = multi_rows;

    float* w_hi = workspace;
    float* w_lo = precision ? workspace + need / 2 : workspace;
    {
        const int64_t total = (int64_t)p.n_k8 * N * 8;
        int blocks = (int)((total + 255) / 256);
        if (blocks > 148 * 8) blocks = 148 * 8;
        split_block_weight_kernel<<<blocks, 256, 0, st>>>(W, ldw, K, N, p.n_k8, w_hi, w_lo, precision ? 1 : 0);
        MG_LAUNCH_CHECK("split_block_weight");
    }
    CUtensorMap mbh, mbl;
    {
        const int nblk = N / 32;
        uint64_t dims[2] = {32, (uint64_t)p.n_k8 * nblk * 8};
        uint64_t str[1] = {128};
        uint32_t box[2] = {32, (uint32_t)nblk * 8};
        if (int rc = make_map(&mbh, w_hi, 2, dims, str, box, true)) return rc;
        if (int rc = make_map(&mbl, w_lo, 2, dims, str, box, true)) return rc;
    }
    return precision ? launch_fused<true>(mbh, mbl, p, st) : launch_fused<false>(mbh, mbl, p, st);
}
