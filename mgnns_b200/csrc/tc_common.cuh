// Shared tcgen05 / TMEM / TMA plumbing for the tensor-core translation units (sm_100a).
#pragma once
#include <cuda.h>
#include <stdlib.h>
#include "common.cuh"

namespace mgnns {
namespace tc {

// ---------------------------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, P1;\n\t"
        "}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {}
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---------------------------------------------------------------------------------- descriptors
// Shared-memory matrix descriptor (SM100): start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46),
// version=1 [46,48), layout type [61,64): 2 = SWIZZLE_128B (16-byte atoms; K-major operands),
// 1 = SWIZZLE_128B_BASE32B (32-byte atoms) — the only layout the tensor core accepts for MN-major tf32
// operands; it pairs with the TMA swizzle mode CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B.
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout = 2) {
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)layout << 61;
    return d;
}
// Instruction descriptor for kind::tf32, fp32 accumulate.
__host__ __device__ constexpr uint32_t instr_desc_tf32(int M, int N, int a_mn_major, int b_mn_major) {
    return (1u << 4)                     // c_format = F32
         | (2u << 7) | (2u << 10)        // a_format = b_format = TF32
         | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16)
         | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ---------------------------------------------------------------------------------- dynamic tile scheduler
// Persistent CTAs draw work items from a global counter instead of a static stride, so a CTA that gets its SM
// late (another stream's kernel — the LSTM recurrence — still holds it) simply takes fewer items and the kernel
// ends when the work does, not when the unluckiest CTA has finished a fixed share.  One thread (the TMA
// producer) fetches the next item with atomicAdd and publishes it through a small shared-memory ring guarded by
// mbarriers; the MMA issuer, the splitter warps and the epilogue warps consume it in order.
constexpr int SCHED_SLOTS = 4;
constexpr int SCHED_CONSUMERS = 9;           // MMA thread + 4 splitter warps + 4 epilogue warps
struct SchedSmem {
    uint64_t full[SCHED_SLOTS];
    uint64_t empty[SCHED_SLOTS];
    int item[SCHED_SLOTS];
};
struct SchedState {
    int slot = 0;
    uint32_t phase = 0;
    __device__ __forceinline__ void advance() { if (++slot == SCHED_SLOTS) { slot = 0; phase ^= 1; } }
};
__device__ __forceinline__ void sched_init(SchedSmem* sm) {
    for (int i = 0; i < SCHED_SLOTS; ++i) {
        mbar_init(&sm->full[i], 1);
        mbar_init(&sm->empty[i], SCHED_CONSUMERS);
    }
}
// producer thread: next item (values >= n_items tell every role to stop)
__device__ __forceinline__ int sched_produce(SchedSmem* sm, SchedState& st, int* counter) {
    mbar_wait(&sm->empty[st.slot], st.phase ^ 1);
    const int item = atomicAdd(counter, 1);
    sm->item[st.slot] = item;
    mbar_arrive(&sm->full[st.slot]);         // release: the store above is visible to the waiters
    st.advance();
    return item;
}
// a single consumer thread (the MMA issuer)
__device__ __forceinline__ int sched_consume_thread(SchedSmem* sm, SchedState& st) {
    mbar_wait(&sm->full[st.slot], st.phase);
    const int item = sm->item[st.slot];
    mbar_arrive(&sm->empty[st.slot]);
    st.advance();
    return item;
}
// a whole consumer warp: every lane reads the item, one lane releases the slot
__device__ __forceinline__ int sched_consume_warp(SchedSmem* sm, SchedState& st, int lane) {
    mbar_wait(&sm->full[st.slot], st.phase);
    const int item = sm->item[st.slot];
    __syncwarp();
    if (lane == 0) mbar_arrive(&sm->empty[st.slot]);
    st.advance();
    return item;
}

// ---------------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult st;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &st) == cudaSuccess &&
            st == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(ptr);
    }
    return fn;
}

// fp32 tensor map, 128-byte swizzle, zero fill out of bounds. dims/strides innermost first.
static int make_map(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                    const uint32_t* box, bool mn_major = false) {
    EncodeTiledFn enc = get_encode();
    MG_REQUIRE(enc != nullptr, "tc_gemm: cuTensorMapEncodeTiled is unavailable (driver too old?)");
    cuuint64_t gdim[5], gstr[5];
    cuuint32_t bx[5], es[5];
    for (int i = 0; i < rank; ++i) { gdim[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
    for (int i = 0; i + 1 < rank; ++i) gstr[i] = strides_bytes[i];
    CUresult rc = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, rank, const_cast<void*>(base), gdim, gstr, bx, es,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                      CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    MG_REQUIRE(rc == CUDA_SUCCESS, "tc_gemm: cuTensorMapEncodeTiled failed with code %d", (int)rc);
    return 0;
}

// one work counter per in-flight launch: a ring of device ints, zeroed on the launch stream before the kernel
int* next_tile_counter(cudaStream_t st);

// CTAs a persistent tensor-core kernel launches: all SMs unless MGNNS_TC_CTAS caps it (leaving SMs to the small
// latency-bound kernels of concurrent streams; the dynamic scheduler makes the grid size a free parameter)
static int sm_count();
static int tc_grid_limit() {
    static int n = 0;
    if (!n) {
        const char* v = getenv("MGNNS_TC_CTAS");
        n = v ? atoi(v) : 0;
        if (n <= 0 || n > sm_count()) n = sm_count();
    }
    return n;
}

static int sm_count() {
    static int n = 0;
    if (!n) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    }
    return n;
}

}  // namespace tc
}  // namespace mgnns
