// Length-aware bidirectional LSTM recurrence over COMPACTED tokens (the reference
// runs a packed cuDNN bi-LSTM, models/Multi_GCN_Multihead_att.py:366-398).
//
// The input projection x_t W_ih^T + b for every valid token is one dense GEMM done by the caller
// (mgnns_gemm_f32 over the compact [N, in] matrix, N = sum of lengths), so padding (84 % of a
// TumEmo batch) is never touched.  This file holds the sequential part: for a tile of TS sequences
// of similar length (the host sorts by length) and one direction, a CTA steps through time keeping
// h in shared memory and c in registers; every step is a [TS x H] x [H x 4H] product against W_hh
// streamed from L2 (transposed copy, every element fetched once per CTA and step).  PyTorch gate order i,f,g,o.
//
// Forward saves the post-activation gates, the cell state and h_{t-1} per token; backward walks the
// same tiles in reverse time and emits the pre-activation gradients dG, from which the caller gets
// dW_ih / dx / db (GEMMs over the compact matrices) and dW_hh = dG^T . Hprev.
#include <stdlib.h>
#include "common.cuh"

namespace mgnns {

constexpr int TS = 8;        // sequences per tile
constexpr int HP = 160;      // hidden size padded to a multiple of 32 (H <= 160)
constexpr int LSTM_THREADS = HP * (TS / 2);   // thread = (unit, pair of sequences)

struct LstmPlan {
    const int32_t* offsets;   // [B+1] first compact row of each sequence
    const int32_t* lens;      // [B]
    const int32_t* tiles;     // [T*TS] sequence ids per tile, -1 = empty slot, sorted by length (desc)
    int H;
};

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }
// tanh through one accurate expf: 1 - 2/(1+e^{2x}); |error| ~1e-7, saturates correctly at +-inf
__device__ __forceinline__ float tanhf_(float x) { return 1.f - 2.f / (1.f + expf(2.f * x)); }

// fast forms for the cluster kernels' gate phase (ex2.approx + approximate reciprocal, ~2 ulp each: 1e-7 relative,
// against the 1e-4 / 1e-5 bound of the LSTM parity test): the accurate expf and IEEE division cost ~25 instructions per
// gate, and the gate phase was 22 % of a recurrence step
__device__ __forceinline__ float sigmoid_fast(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }
__device__ __forceinline__ float tanh_fast(float x) { return 1.f - __fdividef(2.f, 1.f + __expf(2.f * x)); }

constexpr int KCH = 16;      // W_hh elements fetched per thread before they are consumed (memory-level parallelism)

// WT[k][j] = Whh[j][k]   (j = gate*H + unit)
__global__ void lstm_prep_whh_kernel(const float* __restrict__ whh, float* __restrict__ wt, int H) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= H * H * 4) return;
    int j = idx % (4 * H), k = idx / (4 * H);
    wt[idx] = whh[(int64_t)j * H + k];
}

// grid (tiles, 2 directions).  Per step: (1) thread j < 4H computes the recurrent pre-activation
// z[s][j] = sum_k WT[k][j] h[s][k] for all TS sequences (every W element is loaded exactly once per
// CTA and step, KCH loads in flight per thread); (2) thread (unit, sequence pair) applies the gates.
__global__ void __launch_bounds__(LSTM_THREADS) lstm_rec_fwd_kernel(
    LstmPlan plan, const float* __restrict__ G /* [N, 2*4H] */, const float* __restrict__ wt_f,
    const float* __restrict__ wt_r, float* __restrict__ Y /* [N, 2H] */, float* __restrict__ gates /* [N,2,4,H] */,
    float* __restrict__ csave /* [N,2,H] */, float* __restrict__ hprev /* [N,2,H] */) {
    __shared__ __align__(16) float hs[2][HP][TS];        // [buffer][k][sequence]
    __shared__ float zs[TS][4 * HP];
    __shared__ int s_off[TS], s_len[TS];
    const int H = plan.H;
    const int dir = blockIdx.y;
    const float* __restrict__ W = dir ? wt_r : wt_f;
    const int j = threadIdx.x;
    const int u = threadIdx.x % HP, sp = threadIdx.x / HP;
    if (threadIdx.x < TS) {
        int s = plan.tiles[blockIdx.x * TS + threadIdx.x];
        s_off[threadIdx.x] = s >= 0 ? plan.offsets[s] : 0;
        s_len[threadIdx.x] = s >= 0 ? plan.lens[s] : 0;
    }
    for (int i = threadIdx.x; i < 2 * HP * TS; i += LSTM_THREADS) (&hs[0][0][0])[i] = 0.f;
    __syncthreads();
    int tile_len = 0;
#pragma unroll
    for (int s = 0; s < TS; ++s) tile_len = max(tile_len, s_len[s]);
    const int sA = 2 * sp, sB = 2 * sp + 1;
    const int lenA = s_len[sA], lenB = s_len[sB], offA = s_off[sA], offB = s_off[sB];
    float cA = 0.f, cB = 0.f;
    const bool live = u < H;
    const bool jlive = j < 4 * H;
    int cur = 0;
    for (int t = 0; t < tile_len; ++t) {
        if (jlive) {
            float acc[TS];
#pragma unroll
            for (int s = 0; s < TS; ++s) acc[s] = 0.f;
            for (int k0 = 0; k0 < H; k0 += KCH) {
                float w[KCH];
#pragma unroll
                for (int i = 0; i < KCH; ++i) w[i] = (k0 + i < H) ? __ldg(W + (int64_t)(k0 + i) * (4 * H) + j) : 0.f;
#pragma unroll
                for (int i = 0; i < KCH; ++i) {
                    if (k0 + i < H) {
                        const float4 h0 = *reinterpret_cast<const float4*>(&hs[cur][k0 + i][0]);
                        const float4 h1 = *reinterpret_cast<const float4*>(&hs[cur][k0 + i][4]);
                        acc[0] = fmaf(w[i], h0.x, acc[0]); acc[1] = fmaf(w[i], h0.y, acc[1]);
                        acc[2] = fmaf(w[i], h0.z, acc[2]); acc[3] = fmaf(w[i], h0.w, acc[3]);
                        acc[4] = fmaf(w[i], h1.x, acc[4]); acc[5] = fmaf(w[i], h1.y, acc[5]);
                        acc[6] = fmaf(w[i], h1.z, acc[6]); acc[7] = fmaf(w[i], h1.w, acc[7]);
                    }
                }
            }
#pragma unroll
            for (int s = 0; s < TS; ++s) zs[s][j] = acc[s];
        }
        __syncthreads();
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const int len = q ? lenB : lenA, off = q ? offB : offA, s = q ? sB : sA;
            if (live) {
                const float hold = hs[cur][u][s];
                float hnew = hold;
                if (t < len) {
                    const int tt = dir ? (len - 1 - t) : t;
                    const int64_t row = off + tt;
                    const float* g_in = G + row * (8 * H) + dir * 4 * H;
                    const float ig = sigmoidf_(zs[s][u] + g_in[u]);
                    const float fg = sigmoidf_(zs[s][H + u] + g_in[H + u]);
                    const float gg = tanhf_(zs[s][2 * H + u] + g_in[2 * H + u]);
                    const float og = sigmoidf_(zs[s][3 * H + u] + g_in[3 * H + u]);
                    float& c = q ? cB : cA;
                    c = fg * c + ig * gg;
                    hnew = og * tanhf_(c);
                    Y[row * (2 * H) + dir * H + u] = hnew;
                    float* gs = gates + (row * 2 + dir) * (4 * H);
                    gs[u] = ig; gs[H + u] = fg; gs[2 * H + u] = gg; gs[3 * H + u] = og;
                    csave[(row * 2 + dir) * H + u] = c;
                    hprev[(row * 2 + dir) * H + u] = hold;
                }
                hs[cur ^ 1][u][s] = hnew;
            }
        }
        __syncthreads();
        cur ^= 1;
    }
}

__global__ void __launch_bounds__(LSTM_THREADS) lstm_rec_bwd_kernel(
    LstmPlan plan, const float* __restrict__ dY /* [N, 2H] */, const float* __restrict__ gates,
    const float* __restrict__ csave, const float* __restrict__ whh_f /* [4H, H] */,
    const float* __restrict__ whh_r, float* __restrict__ dG /* [N, 2*4H] */) {
    __shared__ __align__(16) float dzs[4 * HP][TS];      // [gate row j][sequence]
    __shared__ float part[4][TS][HP];
    __shared__ float dh_rec[TS][HP];
    __shared__ int s_off[TS], s_len[TS];
    const int H = plan.H;
    const int dir = blockIdx.y;
    const float* __restrict__ W = dir ? whh_r : whh_f;
    const int u = threadIdx.x % HP, sp = threadIdx.x / HP;
    if (threadIdx.x < TS) {
        int s = plan.tiles[blockIdx.x * TS + threadIdx.x];
        s_off[threadIdx.x] = s >= 0 ? plan.offsets[s] : 0;
        s_len[threadIdx.x] = s >= 0 ? plan.lens[s] : 0;
    }
    for (int i = threadIdx.x; i < TS * HP; i += LSTM_THREADS) (&dh_rec[0][0])[i] = 0.f;
    __syncthreads();
    int tile_len = 0;
#pragma unroll
    for (int s = 0; s < TS; ++s) tile_len = max(tile_len, s_len[s]);
    const int sA = 2 * sp, sB = 2 * sp + 1;
    const int lenA = s_len[sA], lenB = s_len[sB], offA = s_off[sA], offB = s_off[sB];
    float dcA = 0.f, dcB = 0.f;
    const bool live = u < H;
    for (int t = tile_len - 1; t >= 0; --t) {
        // ---- phase 1: gate gradients of this step ------------------------------------------------
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const int len = q ? lenB : lenA, off = q ? offB : offA, s = q ? sB : sA;
            float dz[4] = {0.f, 0.f, 0.f, 0.f};
            if (live && t < len) {
                const int tt = dir ? (len - 1 - t) : t;
                const int64_t row = off + tt;
                const float* gs = gates + (row * 2 + dir) * (4 * H);
                const float ig = gs[u], fg = gs[H + u], gg = gs[2 * H + u], og = gs[3 * H + u];
                const float c = csave[(row * 2 + dir) * H + u];
                float cprev = 0.f;
                if (t > 0) {
                    const int64_t rp = dir ? row + 1 : row - 1;
                    cprev = csave[(rp * 2 + dir) * H + u];
                }
                const float dh = dY[row * (2 * H) + dir * H + u] + dh_rec[s][u];
                const float tc = tanhf_(c);
                float& dc = q ? dcB : dcA;
                const float dct = dc + dh * og * (1.f - tc * tc);
                dz[0] = dct * gg * ig * (1.f - ig);
                dz[1] = dct * cprev * fg * (1.f - fg);
                dz[2] = dct * ig * (1.f - gg * gg);
                dz[3] = dh * tc * og * (1.f - og);
                dc = dct * fg;
                float* out = dG + row * (8 * H) + dir * 4 * H;
                out[u] = dz[0]; out[H + u] = dz[1]; out[2 * H + u] = dz[2]; out[3 * H + u] = dz[3];
            }
            if (live) {
                dzs[u][s] = dz[0]; dzs[H + u][s] = dz[1]; dzs[2 * H + u][s] = dz[2]; dzs[3 * H + u][s] = dz[3];
            }
        }
        __syncthreads();
        // ---- phase 2: dh_{t-1} = W_hh^T dz ; thread (k = u, gate block sp) over all TS sequences ------
        if (t > 0) {
            float acc[TS];
#pragma unroll
            for (int s = 0; s < TS; ++s) acc[s] = 0.f;
            if (live) {
                const int j0 = sp * H;
                for (int jb = 0; jb < H; jb += KCH) {
                    float w[KCH];
#pragma unroll
                    for (int i = 0; i < KCH; ++i) w[i] = (jb + i < H) ? __ldg(W + (int64_t)(j0 + jb + i) * H + u) : 0.f;
#pragma unroll
                    for (int i = 0; i < KCH; ++i) {
                        if (jb + i < H) {
                            const float4 d0 = *reinterpret_cast<const float4*>(&dzs[j0 + jb + i][0]);
                            const float4 d1 = *reinterpret_cast<const float4*>(&dzs[j0 + jb + i][4]);
                            acc[0] = fmaf(w[i], d0.x, acc[0]); acc[1] = fmaf(w[i], d0.y, acc[1]);
                            acc[2] = fmaf(w[i], d0.z, acc[2]); acc[3] = fmaf(w[i], d0.w, acc[3]);
                            acc[4] = fmaf(w[i], d1.x, acc[4]); acc[5] = fmaf(w[i], d1.y, acc[5]);
                            acc[6] = fmaf(w[i], d1.z, acc[6]); acc[7] = fmaf(w[i], d1.w, acc[7]);
                        }
                    }
                }
            }
#pragma unroll
            for (int s = 0; s < TS; ++s) part[sp][s][u] = acc[s];
            __syncthreads();
            dh_rec[sA][u] = part[0][sA][u] + part[1][sA][u] + part[2][sA][u] + part[3][sA][u];
            dh_rec[sB][u] = part[0][sB][u] + part[1][sB][u] + part[2][sB][u] + part[3][sB][u];
            __syncthreads();
        }
    }
}


// =====================================================================================================
// Cluster variant: a pair of CTAs (thread-block cluster of 2, distributed shared memory) owns one
// (tile, direction).  Each CTA keeps HALF of W_hh resident in shared memory for the whole kernel
// (forward: the gate columns of its H/2 hidden units, 180 KB at H=150; backward: the matching H/2
// columns of W_hh), so the per-step weight traffic comes from shared memory instead of L2, and the two
// CTAs exchange only the new hidden state (forward) / gate gradients (backward) through DSMEM stores
// followed by one cluster barrier.
// =====================================================================================================
constexpr int CL_THREADS = 640;

__device__ __forceinline__ uint32_t cl_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cl_map(const void* p, uint32_t rank) {
    uint32_t a = (uint32_t)__cvta_generic_to_shared(p), r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(rank));
    return r;
}
__device__ __forceinline__ void cl_store(uint32_t addr, float v) {
    asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ void cl_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// Split form.  The release of an arrive covers every earlier write of the thread, global ones included, so the
// per-step result stores (Y / gates / c / h_prev, dG) are issued BETWEEN arrive and wait: the barrier then only
// waits for the shared-memory exchange, and the global stores drain while the next step's product runs.
__device__ __forceinline__ void cl_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cl_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }

// Per-step exchange without a cluster barrier.  barrier.cluster.arrive.release makes every thread drain ALL its
// outstanding stores first (ncu: 6 % of the forward kernel's and 18 % of the backward kernel's stall samples sat on
// that fence, the per-step global result stores included) and a full cluster barrier costs ~380 cycles.  Instead the
// remote stores are st.async ones that carry their own completion: each lands in the peer's shared memory and
// decrements the transaction count of an mbarrier there, the consumer arms that mbarrier with the byte count it
// expects per step and waits for its phase — one one-way trip (~215 cycles), no fence, and the global stores of the
// step are never waited for.
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mb_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mb_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mb_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// Default semantics (acquire at CTA scope), as for any transaction barrier fed by another CTA of the cluster (TMA
// multicast, st.async): the data lands in THIS CTA's shared memory and complete_tx orders it before the phase flip.
// The .acquire.cluster form makes ptxas emit CCTL.IVALL — an L1 invalidate that first waits for every outstanding
// global load, i.e. for the next step's prefetch: ncu showed 5 % of all stall samples on that one instruction.
__device__ __forceinline__ void mb_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "MB_WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@!p bra MB_WAIT_%=;\n\t}"
        ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// value -> the peer CTA's shared memory at `remote_addr`, 4 bytes counted on the peer's mbarrier `remote_bar`
__device__ __forceinline__ void st_async_f32(uint32_t remote_addr, float v, uint32_t remote_bar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];"
                 ::"r"(remote_addr), "r"(__float_as_uint(v)), "r"(remote_bar) : "memory");
}
// 16 bytes at once: the exchanged blocks are contiguous in shared memory, so a warp's stores form one 512-byte run
// (32 scattered 4-byte DSMEM stores per warp — the layout's natural (unit, sequence) -> address map has a 32-byte
// stride — cost a remote transaction each)
__device__ __forceinline__ void st_async_f32x4(uint32_t remote_addr, const float4& v, uint32_t remote_bar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];"
                 ::"r"(remote_addr), "r"(__float_as_uint(v.x)), "r"(__float_as_uint(v.y)), "r"(__float_as_uint(v.z)),
                   "r"(__float_as_uint(v.w)), "r"(remote_bar) : "memory");
}
__device__ __forceinline__ void mb_arrive_remote(uint32_t remote_bar) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote_bar) : "memory");
}

// HT > 0: hidden size known at compile time (every shared-memory stride becomes an immediate offset: the runtime-H
// version spent ~40 % of its issue slots on address arithmetic for the W / h loads); HT = 0: plan.H
//
// 640 threads (20 warps, five per scheduler).  The per-step product is bound by instruction issue (16 FMA + 4 LDS per
// thread and k), and with 320 threads — 2.5 warps per scheduler — the FMA pipe sat at 31 % (ncu): the loads of one
// warp were not covered by the FMAs of another.  So the K range is cut into FOUR groups of 160 threads (two columns x
// eight sequences per thread, as before), the four partial sums meet in shared memory, and the gate phase has one
// (unit, sequence) pair per thread instead of two in sequence.
constexpr int CL_KGROUPS = 4;
__device__ long long* g_lstm_dbg = nullptr;   // profiling aid: per-phase cycle sums of CTA (0,0), thread 0 and thread 608

template <int HT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(CL_THREADS, 1) lstm_rec_fwd_cl_kernel(
    LstmPlan plan, const float* __restrict__ G, const float* __restrict__ wt_f, const float* __restrict__ wt_r,
    float* __restrict__ Y, float* __restrict__ gates, float* __restrict__ csave, float* __restrict__ hprev,
    int n_tiles, int* __restrict__ tile_ctr_f, int* __restrict__ tile_ctr_r) {
    extern __shared__ __align__(16) float sm[];
    const int H = HT ? HT : plan.H, Hh = H / 2, NC = 4 * Hh;          // NC = gate columns owned by this CTA
    float* Wsm = sm;                                         // [H][NC]
    float* hs = Wsm + H * NC;                                // [2][H][TS]
    float* zp = hs + 2 * H * TS;                             // [CL_KGROUPS][TS][NC]
    __shared__ int s_off[TS], s_len[TS];
    __shared__ int s_tile;
    __shared__ __align__(8) uint64_t full[2];                // full[b]: the peer's half of hs[b] has landed
    const int dir = blockIdx.y;
    const uint32_t rank = cl_rank(), peer = rank ^ 1u;
    const float* __restrict__ WT = dir ? wt_r : wt_f;        // [H][4H]
    int* tile_ctr = dir ? tile_ctr_r : tile_ctr_f;
    const int tid = threadIdx.x;
    if (tid == 0) {
        mb_init(&full[0], 1);
        mb_init(&full[1], 1);
        mb_fence_init();
    }
    // W_hh is loaded ONCE per cluster: the clusters are persistent and draw (length-sorted) tiles from a queue
    for (int i = tid; i < H * NC; i += CL_THREADS) {
        const int k = i / NC, lc = i - k * NC, g = lc / Hh, uu = lc - g * Hh;
        Wsm[i] = WT[(int64_t)k * (4 * H) + g * H + rank * Hh + uu];
    }
    cl_sync();          // both CTAs are running and their mbarriers are initialised before any remote access

    // z-phase mapping: four groups of 160 threads split the k range; each thread owns two columns
    const int grp = tid / 160, jl = tid % 160;
    const bool zlive = jl < NC / 2;
    // gate mapping: one (unit, sequence) pair per thread, fixed for the whole kernel (Hh * TS <= 640 pairs)
    const int npairs = Hh * TS;
    const bool p_ok = tid < npairs;
    const int p_s = p_ok ? tid / Hh : 0, p_u = p_ok ? tid % Hh : 0;
    const uint32_t step_bytes = (uint32_t)(npairs * sizeof(float));     // what the peer sends per step
    const uint32_t peer_full0 = cl_map(&full[0], peer), peer_full1 = cl_map(&full[1], peer);
    const uint32_t peer_tile = cl_map(&s_tile, peer);
    const int u_glob = rank * Hh + p_u;
    long long* dbg = (blockIdx.x == 0 && blockIdx.y == 0 && (tid == 0 || tid == 608)) ? g_lstm_dbg : nullptr;
    long long ph[6] = {0, 0, 0, 0, 0, 0};
    int gstep = 0;                                           // steps done by this cluster over all its tiles
  for (;;) {
    // ---- next tile of the queue (longest first); both CTAs of the cluster must take the same one -----------------
    if (rank == 0 && tid == 0) {
        const int tl = atomicAdd(tile_ctr, 1);
        s_tile = tl;
        asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(peer_tile), "r"(tl) : "memory");
    }
    cl_sync();          // the tile number has landed on both sides; every exchange of the previous tile is complete
    const int tile = s_tile;
    if (tile >= n_tiles) break;
    if (tid < TS) {
        int s = plan.tiles[tile * TS + tid];
        s_off[tid] = s >= 0 ? plan.offsets[s] : 0;
        s_len[tid] = s >= 0 ? plan.lens[s] : 0;
    }
    {   // h_0 = 0: the buffer the first step reads
        float* h0 = hs + (gstep & 1) * H * TS;
        for (int i = tid; i < H * TS; i += CL_THREADS) h0[i] = 0.f;
    }
    __syncthreads();
    int tile_len = 0;
#pragma unroll
    for (int s = 0; s < TS; ++s) tile_len = max(tile_len, s_len[s]);
    const int p_len = s_len[p_s], p_off = s_off[p_s];
    float cst = 0.f;
    // input projections of one step for the thread's pair; the NEXT step's are fetched one step ahead, so their
    // L2 latency is off the per-step critical path
    struct Gin { float g[4]; int64_t row; bool on; };
    auto fetch = [&](int t, Gin& gi) {
        gi.on = false;
        gi.row = 0;
        if (!p_ok || t >= p_len) return;
        const int tt = dir ? (p_len - 1 - t) : t;
        gi.row = p_off + tt;
        const float* g_in = G + gi.row * (8 * H) + dir * 4 * H + rank * Hh + p_u;
#pragma unroll
        for (int g = 0; g < 4; ++g) gi.g[g] = __ldg(g_in + g * H);
        gi.on = true;
    };
    Gin gin, gnx;
    fetch(0, gin);
    for (int t = 0; t < tile_len; ++t, ++gstep) {
        const int cur = gstep & 1;
        long long c0 = 0, c1 = 0, c2 = 0, c3 = 0, c4 = 0, c5 = 0;
        if (dbg) c0 = clock64();
        if (tid == 0) mb_expect_tx(&full[cur ^ 1], step_bytes);
        if (zlive) {
            // two columns x eight sequences per thread; sequences in pairs, one FFMA2 per (column, pair) and k
            unsigned long long a0[TS / 2], a1[TS / 2];
#pragma unroll
            for (int s = 0; s < TS / 2; ++s) { a0[s] = 0ull; a1[s] = 0ull; }
            const float* hcur = hs + cur * H * TS;
            auto body = [&](int k) {
                const float w0 = Wsm[k * NC + jl], w1 = Wsm[k * NC + NC / 2 + jl];
                const unsigned long long w00 = pack_f32x2(w0, w0), w11 = pack_f32x2(w1, w1);
                const ulonglong2 h0 = *reinterpret_cast<const ulonglong2*>(hcur + k * TS);
                const ulonglong2 h1 = *reinterpret_cast<const ulonglong2*>(hcur + k * TS + 4);
                a0[0] = fma_f32x2(w00, h0.x, a0[0]); a0[1] = fma_f32x2(w00, h0.y, a0[1]);
                a0[2] = fma_f32x2(w00, h1.x, a0[2]); a0[3] = fma_f32x2(w00, h1.y, a0[3]);
                a1[0] = fma_f32x2(w11, h0.x, a1[0]); a1[1] = fma_f32x2(w11, h0.y, a1[1]);
                a1[2] = fma_f32x2(w11, h1.x, a1[2]); a1[3] = fma_f32x2(w11, h1.y, a1[3]);
            };
            // group g takes k = g, g + 4, g + 8, ...: H / 4 iterations for every group (a compile-time count when HT
            // is given: no peeled alignment iterations) plus one more for the first H % 4 groups
            int k = grp;
#pragma unroll 4
            for (int i = 0; i < (H / CL_KGROUPS) / 4 * 4; ++i, k += CL_KGROUPS) body(k);
#pragma unroll
            for (int i = (H / CL_KGROUPS) / 4 * 4; i < H / CL_KGROUPS; ++i, k += CL_KGROUPS) body(k);
            if (grp < H % CL_KGROUPS) body(k);
            float* z = zp + grp * TS * NC;
#pragma unroll
            for (int s = 0; s < TS / 2; ++s) {
                float lo, hi;
                unpack_f32x2(a0[s], lo, hi);
                z[(2 * s) * NC + jl] = lo; z[(2 * s + 1) * NC + jl] = hi;
                unpack_f32x2(a1[s], lo, hi);
                z[(2 * s) * NC + NC / 2 + jl] = lo; z[(2 * s + 1) * NC + NC / 2 + jl] = hi;
            }
        }
        // next step's input projections: issued AFTER the product (in flight during the gate phase and the exchange,
        // consumed a whole step later) — issued before it, the loads shared scoreboard slots with the product's
        // shared-memory loads and stalled it
        if (dbg) c1 = clock64();
        fetch(t + 1, gnx);
        __syncthreads();
        if (dbg) c2 = clock64();
        float* hnext = hs + (cur ^ 1) * H * TS;
        const uint32_t peer_full = cur ? peer_full0 : peer_full1;          // the peer's full[cur ^ 1]
        float sv[7];                                         // i, f, g, o, c, h_prev, h of the thread's pair
        if (p_ok) {
            const float hold = hs[cur * H * TS + u_glob * TS + p_s];
            float hnew = hold;
            if (gin.on) {
                float zg[4];
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    float acc = gin.g[g];
#pragma unroll
                    for (int kg = 0; kg < CL_KGROUPS; ++kg) acc += zp[(kg * TS + p_s) * NC + g * Hh + p_u];
                    zg[g] = acc;
                }
                const float ig = sigmoid_fast(zg[0]), fg = sigmoid_fast(zg[1]), gg = tanh_fast(zg[2]), og = sigmoid_fast(zg[3]);
                cst = fg * cst + ig * gg;
                hnew = og * tanh_fast(cst);
                sv[0] = ig; sv[1] = fg; sv[2] = gg; sv[3] = og; sv[4] = cst; sv[5] = hold; sv[6] = hnew;
            }
            hnext[u_glob * TS + p_s] = hnew;
            if (gin.on) {
                const int64_t row = gin.row;
                Y[row * (2 * H) + dir * H + u_glob] = sv[6];
                float* gs = gates + (row * 2 + dir) * (4 * H);
                gs[u_glob] = sv[0]; gs[H + u_glob] = sv[1]; gs[2 * H + u_glob] = sv[2]; gs[3 * H + u_glob] = sv[3];
                csave[(row * 2 + dir) * H + u_glob] = sv[4];
                hprev[(row * 2 + dir) * H + u_glob] = sv[5];
            }
        }
        // hs[cur ^ 1] complete: own half (barrier below; it is one contiguous block of Hh * TS floats, which the first
        // Hh * TS / 4 threads then copy to the peer 16 bytes each), the peer's half (its bytes counted on
        // full[cur ^ 1]).  The peer overwrites hs[cur] only in ITS step t + 1, i.e. after it has received every byte of
        // this step from here — which is sent after every thread's last read of hs[cur].
        if (dbg) c3 = clock64();
        __syncthreads();
        if (dbg) c4 = clock64();
        if (tid < npairs / 4) {
            const float* src = hnext + rank * Hh * TS + 4 * tid;
            st_async_f32x4(cl_map(src, peer), *reinterpret_cast<const float4*>(src), peer_full);
        }
        mb_wait(&full[cur ^ 1], (uint32_t)(gstep >> 1) & 1u);
        if (dbg) {
            c5 = clock64();
            ph[0] += c1 - c0; ph[1] += c2 - c1; ph[2] += c3 - c2; ph[3] += c4 - c3; ph[4] += c5 - c4; ph[5] += 1;
        }
        gin = gnx;
    }
  }
    if (dbg) {
        long long* o = dbg + (tid == 0 ? 0 : 8);
#pragma unroll
        for (int i = 0; i < 6; ++i) o[i] = ph[i];
    }
    cl_sync();          // neither CTA leaves (and frees its shared memory) while the other may still write into it
}

constexpr int CL_JGROUPS = 8;        // backward phase 2: the 4H gate rows are cut into eight groups of H/2

template <int HT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(CL_THREADS, 1) lstm_rec_bwd_cl_kernel(
    LstmPlan plan, const float* __restrict__ dY, const float* __restrict__ gates, const float* __restrict__ csave,
    const float* __restrict__ whh_f, const float* __restrict__ whh_r, float* __restrict__ dG,
    int n_tiles, int* __restrict__ tile_ctr_f, int* __restrict__ tile_ctr_r) {
    extern __shared__ __align__(16) float sm[];
    const int H = HT ? HT : plan.H, Hh = H / 2;
    float* Wb = sm;                                          // [4H][Hh]: W_hh columns of this CTA's hidden units
    float* dzs = Wb + 4 * H * Hh;                            // [4H][TS]
    float* part = dzs + 4 * H * TS;                          // [CL_JGROUPS][TS][Hh]
    float* dh = part + CL_JGROUPS * TS * Hh;                 // [TS][Hh]
    __shared__ int s_off[TS], s_len[TS];
    __shared__ int s_tile;
    __shared__ __align__(8) uint64_t full, freeb;            // full: the peer's gate gradients of this step have landed;
                                                             // freeb: the peer has finished reading what was sent to it
    const int dir = blockIdx.y;
    const uint32_t rank = cl_rank(), peer = rank ^ 1u;
    const float* __restrict__ W = dir ? whh_r : whh_f;       // [4H][H]
    int* tile_ctr = dir ? tile_ctr_r : tile_ctr_f;
    const int tid = threadIdx.x;
    if (tid == 0) {
        mb_init(&full, 1);
        mb_init(&freeb, 1);
        mb_fence_init();
    }
    // W_hh is loaded ONCE per cluster: the clusters are persistent and draw (length-sorted) tiles from a queue
    for (int i = tid; i < 4 * H * Hh; i += CL_THREADS) {
        const int j = i / Hh, kk = i - j * Hh;
        Wb[i] = W[(int64_t)j * H + rank * Hh + kk];
    }
    cl_sync();          // both CTAs are running and their mbarriers are initialised before any remote access

    // phase-1 mapping: one (unit, sequence) pair per thread
    const int npairs = Hh * TS;
    const bool p_ok = tid < npairs;
    const int p_s = p_ok ? tid / Hh : 0, p_u = p_ok ? tid % Hh : 0;
    const int u_glob = rank * Hh + p_u;
    // phase-2 mapping: thread (kk, one of eight groups of gate rows)
    const int kk2 = tid % 80, jq = tid / 80;
    const bool live2 = kk2 < Hh;
    const uint32_t step_bytes = (uint32_t)(4 * npairs * sizeof(float));  // what the peer sends per step
    const uint32_t peer_full = cl_map(&full, peer), peer_free = cl_map(&freeb, peer);
    const uint32_t peer_tile = cl_map(&s_tile, peer);
    int gstep = 0;                                           // steps done by this cluster over all its tiles
  for (;;) {
    // ---- next tile of the queue (longest first); both CTAs of the cluster must take the same one -----------------
    if (rank == 0 && tid == 0) {
        const int tl = atomicAdd(tile_ctr, 1);
        s_tile = tl;
        asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(peer_tile), "r"(tl) : "memory");
    }
    cl_sync();          // the tile number has landed on both sides; every exchange of the previous tile is complete
    const int tile = s_tile;
    if (tile >= n_tiles) break;
    if (tid < TS) {
        int s = plan.tiles[tile * TS + tid];
        s_off[tid] = s >= 0 ? plan.offsets[s] : 0;
        s_len[tid] = s >= 0 ? plan.lens[s] : 0;
    }
    for (int i = tid; i < TS * Hh; i += CL_THREADS) dh[i] = 0.f;
    __syncthreads();
    int tile_len = 0;
#pragma unroll
    for (int s = 0; s < TS; ++s) tile_len = max(tile_len, s_len[s]);
    const int p_len = s_len[p_s], p_off = s_off[p_s];
    float dcs = 0.f;
    // saved forward tensors of one time step for the thread's pair; the NEXT step's are fetched while this step's
    // exchange and product run, so their L2 latency is off the per-step critical path
    struct Saved { float ig, fg, gg, og, c, cprev, dy; int64_t row; bool on; };
    auto fetch = [&](int t, Saved& sv) {
        sv.on = false;
        if (!p_ok || t < 0 || t >= p_len) return;
        const int tt = dir ? (p_len - 1 - t) : t;
        const int64_t row = p_off + tt;
        const float* gs = gates + (row * 2 + dir) * (4 * H);
        sv.ig = __ldg(gs + u_glob); sv.fg = __ldg(gs + H + u_glob); sv.gg = __ldg(gs + 2 * H + u_glob); sv.og = __ldg(gs + 3 * H + u_glob);
        sv.c = __ldg(csave + (row * 2 + dir) * H + u_glob);
        sv.cprev = 0.f;
        if (t > 0) {
            const int64_t rp = dir ? row + 1 : row - 1;
            sv.cprev = __ldg(csave + (rp * 2 + dir) * H + u_glob);
        }
        sv.dy = __ldg(dY + row * (2 * H) + dir * H + u_glob);
        sv.row = row;
        sv.on = true;
    };
    Saved cur_sv, nxt_sv;
    fetch(tile_len - 1, cur_sv);
    for (int t = tile_len - 1; t >= 0; --t, ++gstep) {
        if (tid == 0) mb_expect_tx(&full, step_bytes);
        // ---- phase 1: gate gradients for this CTA's hidden units -------------------------------------
        float dz[4] = {0.f, 0.f, 0.f, 0.f};
        if (p_ok) {
            if (cur_sv.on) {
                const float ig = cur_sv.ig, fg = cur_sv.fg, gg = cur_sv.gg, og = cur_sv.og;
                const float c = cur_sv.c, cprev = cur_sv.cprev;
                const float dhv = cur_sv.dy + dh[p_s * Hh + p_u];
                const float tc = tanh_fast(c);
                const float dct = dcs + dhv * og * (1.f - tc * tc);
                dz[0] = dct * gg * ig * (1.f - ig);
                dz[1] = dct * cprev * fg * (1.f - fg);
                dz[2] = dct * ig * (1.f - gg * gg);
                dz[3] = dhv * tc * og * (1.f - og);
                dcs = dct * fg;
            }
#pragma unroll
            for (int g = 0; g < 4; ++g) dzs[(g * H + u_glob) * TS + p_s] = dz[g];
            if (cur_sv.on) {
                float* out = dG + cur_sv.row * (8 * H) + dir * 4 * H;
                out[u_glob] = dz[0]; out[H + u_glob] = dz[1]; out[2 * H + u_glob] = dz[2]; out[3 * H + u_glob] = dz[3];
            }
        }
        fetch(t - 1, nxt_sv);
        __syncthreads();                                     // own half of dzs: four contiguous blocks of Hh * TS floats
        if (p_ok) {
            // the peer's dzs may be overwritten once it has finished reading the previous step's (it says so every step)
            if (gstep > 0) mb_wait(&freeb, (uint32_t)(gstep - 1) & 1u);
            const int g = tid / (npairs / 4), r = tid - g * (npairs / 4);
            const float* src = dzs + (g * H + rank * Hh) * TS + 4 * r;
            st_async_f32x4(cl_map(src, peer), *reinterpret_cast<const float4*>(src), peer_full);
        }
        mb_wait(&full, (uint32_t)gstep & 1u);                // the peer's half
        // ---- phase 2: dh_{t-1}[s][k] = sum_j W_hh[j][k] dz[s][j] for this CTA's k half ------------------
        if (t > 0) {
            if (live2) {
                unsigned long long acc[TS / 2];
#pragma unroll
                for (int s = 0; s < TS / 2; ++s) acc[s] = 0ull;
                const int j0 = jq * Hh;                      // 4H / 8 gate rows per group
#pragma unroll 5
                for (int j = 0; j < Hh; ++j) {
                    const float w = Wb[(j0 + j) * Hh + kk2];
                    const unsigned long long ww = pack_f32x2(w, w);
                    const ulonglong2 d0 = *reinterpret_cast<const ulonglong2*>(dzs + (j0 + j) * TS);
                    const ulonglong2 d1 = *reinterpret_cast<const ulonglong2*>(dzs + (j0 + j) * TS + 4);
                    acc[0] = fma_f32x2(ww, d0.x, acc[0]); acc[1] = fma_f32x2(ww, d0.y, acc[1]);
                    acc[2] = fma_f32x2(ww, d1.x, acc[2]); acc[3] = fma_f32x2(ww, d1.y, acc[3]);
                }
#pragma unroll
                for (int s = 0; s < TS / 2; ++s) {
                    float lo, hi;
                    unpack_f32x2(acc[s], lo, hi);
                    part[(jq * TS + 2 * s) * Hh + kk2] = lo;
                    part[(jq * TS + 2 * s + 1) * Hh + kk2] = hi;
                }
            }
            __syncthreads();                                 // every thread has finished reading dzs
            if (tid == 0) mb_arrive_remote(peer_free);       // ... so the peer may send its next step's gradients
            if (tid < TS * Hh) {
                float v = 0.f;
#pragma unroll
                for (int g = 0; g < CL_JGROUPS; ++g) v += part[g * TS * Hh + tid];
                dh[tid] = v;
            }
            __syncthreads();                                 // dh complete before phase 1 of the next step reads it
        } else {
            // last step of the tile: nobody reads dzs, but the peer's next send (first step of its next tile) still
            // waits for this step's "free" — every step signals exactly once
            __syncthreads();
            if (tid == 0) mb_arrive_remote(peer_free);
        }
        cur_sv = nxt_sv;
    }
  }
    cl_sync();          // neither CTA leaves (and frees its shared memory) while the other may still write into it
}

// Persistent clusters: `lstm_clusters` per direction (2 CTAs = 2 SMs each) draw the length-sorted tiles from a queue.
// One cluster per tile (the earlier scheme) put 256 whole-SM CTAs on the GPU at every launch: for ~100 us nothing else
// ran (timeline: every other stream of the training step stalled), and each of those CTAs first loaded its 180 KB half
// of W_hh for what was mostly a handful of time steps.  The longest tile (100 steps) bounds the kernel either way; 20
// clusters per direction finish the other 63 tiles of a B = 512 batch (~1,100 steps in total) well inside that time and
// leave 68 SMs to the rest of the step.  MGNNS_LSTM_CLUSTERS overrides.
static int lstm_clusters(int n_tiles) {
    static int n = 0;
    if (!n) {
        const char* v = getenv("MGNNS_LSTM_CLUSTERS");
        n = v ? atoi(v) : 20;
        if (n < 1) n = 20;
    }
    return n_tiles < n ? n_tiles : n;
}

}  // namespace mgnns
namespace mgnns { namespace tc { int* next_tile_counter(cudaStream_t st); } }
namespace mgnns {

// two queue heads (forward / reverse direction), zeroed on the launch stream (capture-aware slots, tc_gemm.cu)
static int lstm_tile_counters(cudaStream_t st, int** cf, int** cr) {
    *cf = tc::next_tile_counter(st);
    *cr = tc::next_tile_counter(st);
    return (*cf && *cr) ? 0 : 1;
}

static size_t lstm_cl_fwd_smem(int H) { return sizeof(float) * ((size_t)H * 2 * H + 2 * H * TS + (size_t)CL_KGROUPS * TS * 2 * H); }
static size_t lstm_cl_bwd_smem(int H) { return sizeof(float) * ((size_t)4 * H * (H / 2) + 4 * H * TS + (size_t)CL_JGROUPS * TS * (H / 2) + TS * (H / 2)); }

}  // namespace mgnns

using namespace mgnns;

// wt [H,4H] <- whh [4H,H] (transpose)
extern "C" int mgnns_lstm_prep_whh(const float* whh, float* wt4, int H, void* stream) {
    MG_REQUIRE(H >= 1 && H <= HP, "lstm_prep_whh: hidden size %d unsupported (max %d)", H, HP);
    MG_REQUIRE(whh && wt4, "lstm_prep_whh: null pointer");
    int n = H * H * 4;
    lstm_prep_whh_kernel<<<(n + 255) / 256, 256, 0, as_stream(stream)>>>(whh, wt4, H);
    MG_LAUNCH_CHECK("lstm_prep_whh");
    return 0;
}

// profiling aid (not part of the public header): cycle sums per phase of the forward kernel's CTA (0,0) go to buf[16]
extern "C" int mgnns_lstm_debug_buffer(long long* buf) {
    return cudaMemcpyToSymbol(g_lstm_dbg, &buf, sizeof(buf)) == cudaSuccess ? 0 : 1;
}

extern "C" int mgnns_lstm_rec_fwd(const int32_t* offsets, const int32_t* lens, const int32_t* tiles, int n_tiles,
                                  int H, const float* G, const float* wt4_f, const float* wt4_r, float* Y,
                                  float* gates, float* csave, float* hprev, void* stream) {
    MG_REQUIRE(H >= 1 && H <= HP && H % 2 == 0, "lstm_rec_fwd: hidden size %d unsupported (even, <= %d)", H, HP);
    if (n_tiles == 0) return 0;
    MG_REQUIRE(offsets && lens && tiles && G && wt4_f && wt4_r && Y && gates && csave && hprev, "lstm_rec_fwd: null pointer");
    LstmPlan plan{offsets, lens, tiles, H};
    const size_t smem = lstm_cl_fwd_smem(H);
    if (smem <= 225 * 1024 && H <= 2 * 80) {
        // cluster-of-2 variant with W_hh resident in (distributed) shared memory
        static bool configured = false;
        if (!configured) {
            cudaError_t e = cudaFuncSetAttribute(lstm_rec_fwd_cl_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024);
            if (e == cudaSuccess) e = cudaFuncSetAttribute(lstm_rec_fwd_cl_kernel<150>, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024);
            MG_REQUIRE(e == cudaSuccess, "lstm_rec_fwd: cannot reserve shared memory: %s", cudaGetErrorString(e));
            configured = true;
        }
        int *cf = nullptr, *cr = nullptr;
        const int ncl = lstm_clusters(n_tiles);
        MG_REQUIRE(lstm_tile_counters(as_stream(stream), &cf, &cr) == 0, "lstm_rec_fwd: cannot set up the tile counters");
        if (H == 150)      // the reference's hidden size (entry: --hidden_size 150)
            lstm_rec_fwd_cl_kernel<150><<<dim3(2 * ncl, 2), CL_THREADS, smem, as_stream(stream)>>>(plan, G, wt4_f, wt4_r, Y, gates, csave, hprev, n_tiles, cf, cr);
        else
            lstm_rec_fwd_cl_kernel<0><<<dim3(2 * ncl, 2), CL_THREADS, smem, as_stream(stream)>>>(plan, G, wt4_f, wt4_r, Y, gates, csave, hprev, n_tiles, cf, cr);
        MG_LAUNCH_CHECK("lstm_rec_fwd_cl");
        return 0;
    }
    lstm_rec_fwd_kernel<<<dim3(n_tiles, 2), LSTM_THREADS, 0, as_stream(stream)>>>(plan, G, wt4_f, wt4_r, Y, gates, csave, hprev);
    MG_LAUNCH_CHECK("lstm_rec_fwd");
    return 0;
}

extern "C" int mgnns_lstm_rec_bwd(const int32_t* offsets, const int32_t* lens, const int32_t* tiles, int n_tiles,
                                  int H, const float* dY, const float* gates, const float* csave,
                                  const float* whh_f, const float* whh_r, float* dG, void* stream) {
    MG_REQUIRE(H >= 1 && H <= HP && H % 2 == 0, "lstm_rec_bwd: hidden size %d unsupported (even, <= %d)", H, HP);
    if (n_tiles == 0) return 0;
    MG_REQUIRE(offsets && lens && tiles && dY && gates && csave && whh_f && whh_r && dG, "lstm_rec_bwd: null pointer");
    LstmPlan plan{offsets, lens, tiles, H};
    const size_t smem = lstm_cl_bwd_smem(H);
    if (smem <= 225 * 1024 && H <= 2 * 80) {
        static bool configured = false;
        if (!configured) {
            cudaError_t e = cudaFuncSetAttribute(lstm_rec_bwd_cl_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024);
            if (e == cudaSuccess) e = cudaFuncSetAttribute(lstm_rec_bwd_cl_kernel<150>, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024);
            MG_REQUIRE(e == cudaSuccess, "lstm_rec_bwd: cannot reserve shared memory: %s", cudaGetErrorString(e));
            configured = true;
        }
        int *cf = nullptr, *cr = nullptr;
        const int ncl = lstm_clusters(n_tiles);
        MG_REQUIRE(lstm_tile_counters(as_stream(stream), &cf, &cr) == 0, "lstm_rec_bwd: cannot set up the tile counters");
        if (H == 150)
            lstm_rec_bwd_cl_kernel<150><<<dim3(2 * ncl, 2), CL_THREADS, smem, as_stream(stream)>>>(plan, dY, gates, csave, whh_f, whh_r, dG, n_tiles, cf, cr);
        else
            lstm_rec_bwd_cl_kernel<0><<<dim3(2 * ncl, 2), CL_THREADS, smem, as_stream(stream)>>>(plan, dY, gates, csave, whh_f, whh_r, dG, n_tiles, cf, cr);
        MG_LAUNCH_CHECK("lstm_rec_bwd_cl");
        return 0;
    }
    lstm_rec_bwd_kernel<<<dim3(n_tiles, 2), LSTM_THREADS, 0, as_stream(stream)>>>(plan, dY, gates, csave, whh_f, whh_r, dG);
    MG_LAUNCH_CHECK("lstm_rec_bwd");
    return 0;
}
