// Gradient all-reduce over NVLink peer memory, as ONE kernel that a CUDA graph can hold
// (ref: the training step, engine/Multi_GCN_Multihead_Att_engine.py:847-851 — loss.backward(); clip_grad_norm_;
//  optimizer.step().  SURVEY §8e: the path's only collective is the gradient all-reduce between backward and clip.)
//
// Every rank keeps its flat gradient buffer in a cudaMalloc allocation that the other ranks of the box have mapped
// through CUDA IPC (mgnns_p2p_export / _import), so a kernel can address all `world` copies directly; NVSwitch gives
// every pair full bandwidth.  The all-reduce is a single pass ("reduce + broadcast"):
//
//   barrier A   CTA c of rank r tells CTA c of every peer that r's gradients are packed, and waits for theirs
//   pass        rank r owns slice r of the buffer: for each 16-byte element of it, load the `world` copies (world-1 of
//               them over NVLink, volatile: no stale L1 lines), add them in rank order, scale by 1/world, and store the
//               result into all `world` buffers (world-1 NVLink stores).  Loads and stores use both directions of the
//               links at once; each rank moves (world-1)/world of the payload each way.
//   barrier B   all of r's stores are visible system-wide before the peers read their buffers (clip + Adam), and no
//               rank starts overwriting its buffer (next step's gradients) while a peer is still reading it
//
// Barriers are per CTA index (CTA c only ever touches sub-chunk c of every slice, on every rank), built from
// st.release.sys / ld.acquire.sys on flag words in peer memory with monotonically increasing epochs kept on the
// device, so replaying a captured graph needs no host-side state.  A rank reduces each element of its slice exactly
// once and in rank order, so all ranks end up with bit-identical sums.  A spin that lasts longer than 20 s sets an
// error word instead of hanging the GPU.
#include <string.h>
#include "common.cuh"

namespace mgnns {

constexpr int P2P_MAX_WORLD = 8;
constexpr int P2P_MAX_CTAS = 128;
constexpr int P2P_THREADS = 512;
// layout of a rank's flag allocation (uint32 words)
constexpr int P2P_FLAG_WORDS = P2P_MAX_CTAS * P2P_MAX_WORLD;     // flag[c][peer]
constexpr int P2P_EPOCH_OFF = P2P_FLAG_WORDS;                    // epoch[c]
constexpr int P2P_ERR_OFF = P2P_EPOCH_OFF + P2P_MAX_CTAS;        // err[1]
constexpr int P2P_WORDS = P2P_ERR_OFF + 4;

struct P2PParams {
    float* buf[P2P_MAX_WORLD];
    uint32_t* flag[P2P_MAX_WORLD];
    int rank, world;
    int64_t n4;           // float4 elements
    float scale;
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float4 ld_volatile4(const float* p) {
    float4 v;
    asm volatile("ld.volatile.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

__device__ __forceinline__ void p2p_barrier(const P2PParams& p, int c, uint32_t val) {
    __threadfence_system();
    __syncthreads();
    const int t = threadIdx.x;
    if (t < p.world && t != p.rank) {
        st_release_sys(p.flag[t] + c * P2P_MAX_WORLD + p.rank, val);
        const uint32_t* mine = p.flag[p.rank] + c * P2P_MAX_WORLD + t;
        const unsigned long long t0 = global_ns();
        // epochs only grow; the signed difference also survives the 32-bit wrap
        while ((int32_t)(ld_acquire_sys(mine) - val) < 0) {
            if (global_ns() - t0 > 20000000000ull) {
                p.flag[p.rank][P2P_ERR_OFF] = 1u;
                break;
            }
        }
    }
    __syncthreads();
}

// U elements per thread and iteration: (W-1)*U remote 128-bit loads in flight per thread.  A peer load takes 1-3 us
// under load, so NVLink bandwidth is bought with bytes in flight: ~100 CTAs x 512 threads x ~10 loads x 16 B = 8 MB.
template <int W, int U>
__global__ void __launch_bounds__(P2P_THREADS) allreduce_p2p_kernel(P2PParams p) {
    const int c = blockIdx.x;
    uint32_t* epoch = p.flag[p.rank] + P2P_EPOCH_OFF + c;
    const uint32_t e = *epoch;                       // only this CTA index ever touches epoch[c]
    p2p_barrier(p, c, e + 1);

    const int64_t slice = (p.n4 + W - 1) / W;
    const int64_t lo = slice * p.rank;
    const int64_t hi = min(lo + slice, p.n4);
    const int64_t stride = (int64_t)gridDim.x * P2P_THREADS;
    for (int64_t j = lo + (int64_t)c * P2P_THREADS + threadIdx.x; j < hi; j += U * stride) {
        float4 v[U][W];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t ju = j + u * stride;
            if (ju < hi) {
#pragma unroll
                for (int q = 0; q < W; ++q) v[u][q] = ld_volatile4(p.buf[q] + 4 * ju);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t ju = j + u * stride;
            if (ju < hi) {
                float4 a = v[u][0];
#pragma unroll
                for (int q = 1; q < W; ++q) { a.x += v[u][q].x; a.y += v[u][q].y; a.z += v[u][q].z; a.w += v[u][q].w; }
                a.x *= p.scale; a.y *= p.scale; a.z *= p.scale; a.w *= p.scale;
#pragma unroll
                for (int q = 0; q < W; ++q) *reinterpret_cast<float4*>(p.buf[q] + 4 * ju) = a;
            }
        }
    }
    p2p_barrier(p, c, e + 2);
    if (threadIdx.x == 0) *epoch = e + 2;
}

}  // namespace mgnns

using namespace mgnns;

extern "C" int mgnns_p2p_flag_bytes(void) { return P2P_WORDS * 4; }

extern "C" int mgnns_p2p_alloc(int64_t bytes, void** out) {
    MG_REQUIRE(bytes > 0 && out, "p2p_alloc: bad argument");
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, (size_t)bytes);
    MG_REQUIRE(e == cudaSuccess, "p2p_alloc: cudaMalloc(%lld) failed: %s", (long long)bytes, cudaGetErrorString(e));
    e = cudaMemset(p, 0, (size_t)bytes);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    MG_REQUIRE(e == cudaSuccess, "p2p_alloc: clearing the buffer failed: %s", cudaGetErrorString(e));
    *out = p;
    return 0;
}

extern "C" int mgnns_p2p_free(void* p) {
    if (p) cudaFree(p);
    return 0;
}

extern "C" int mgnns_p2p_export(void* p, void* handle64) {
    MG_REQUIRE(p && handle64, "p2p_export: null pointer");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaError_t e = cudaIpcGetMemHandle(reinterpret_cast<cudaIpcMemHandle_t*>(handle64), p);
    MG_REQUIRE(e == cudaSuccess, "p2p_export: cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e));
    return 0;
}

extern "C" int mgnns_p2p_import(const void* handle64, void** out) {
    MG_REQUIRE(handle64 && out, "p2p_import: null pointer");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, sizeof(h));
    void* p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    MG_REQUIRE(e == cudaSuccess, "p2p_import: cudaIpcOpenMemHandle failed: %s", cudaGetErrorString(e));
    *out = p;
    return 0;
}

extern "C" int mgnns_p2p_close(void* p) {
    if (p) cudaIpcCloseMemHandle(p);
    return 0;
}

// bufs / flags: HOST arrays of `world` device pointers (entry `rank` is this rank's own allocation, the others are
// the imported mappings).  n must be a multiple of 4 floats, every buffer 16-byte aligned.
extern "C" int mgnns_allreduce_p2p_f32(const uint64_t* bufs, const uint64_t* flags, int rank, int world, int64_t n,
                                       float scale, int ctas, void* stream) {
    MG_REQUIRE(world >= 1 && world <= P2P_MAX_WORLD && rank >= 0 && rank < world, "allreduce_p2p: rank %d / world %d unsupported (<= %d)",
               rank, world, P2P_MAX_WORLD);
    MG_REQUIRE(n >= 0 && (n & 3) == 0, "allreduce_p2p: n must be a multiple of 4 floats");
    MG_REQUIRE(bufs && flags, "allreduce_p2p: null pointer");
    if (n == 0) return 0;
    if (ctas <= 0) ctas = 96;
    if (ctas > P2P_MAX_CTAS) ctas = P2P_MAX_CTAS;
    P2PParams p{};
    for (int q = 0; q < world; ++q) {
        p.buf[q] = reinterpret_cast<float*>(bufs[q]);
        p.flag[q] = reinterpret_cast<uint32_t*>(flags[q]);
        MG_REQUIRE(p.buf[q] && p.flag[q] && aligned16(p.buf[q]), "allreduce_p2p: buffer %d is null or not 16-byte aligned", q);
    }
    p.rank = rank; p.world = world; p.n4 = n >> 2; p.scale = scale;
    cudaStream_t st = as_stream(stream);
    switch (world) {
        case 1: allreduce_p2p_kernel<1, 4><<<ctas, P2P_THREADS, 0, st>>>(p); break;
        case 2: allreduce_p2p_kernel<2, 8><<<ctas, P2P_THREADS, 0, st>>>(p); break;
        case 3: allreduce_p2p_kernel<3, 5><<<ctas, P2P_THREADS, 0, st>>>(p); break;
        case 4: allreduce_p2p_kernel<4, 4><<<ctas, P2P_THREADS, 0, st>>>(p); break;
        case 5: allreduce_p2p_kernel<5, 3><<<ctas, P2P_THREADS, 0, st>>>(p); break;
        case 6: allreduce_p2p_kernel<6, 3><<<ctas, P2P_THREADS, 0, st>>>(p); break;
        case 7: allreduce_p2p_kernel<7, 2><<<ctas, P2P_THREADS, 0, st>>>(p); break;
        default: allreduce_p2p_kernel<8, 2><<<ctas, P2P_THREADS, 0, st>>>(p); break;
    }
    MG_LAUNCH_CHECK("allreduce_p2p");
    return 0;
}

// 1 if a barrier of this rank ever timed out (device word read back; synchronises the stream)
extern "C" int mgnns_p2p_error(const void* own_flags, void* stream) {
    if (!own_flags) return 0;
    uint32_t v = 0;
    cudaStream_t st = as_stream(stream);
    if (cudaMemcpyAsync(&v, reinterpret_cast<const uint32_t*>(own_flags) + P2P_ERR_OFF, 4, cudaMemcpyDeviceToHost, st) != cudaSuccess) return -1;
    if (cudaStreamSynchronize(st) != cudaSuccess) return -1;
    return (int)v;
}
