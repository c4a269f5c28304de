// Stand-alone probe for the tcgen05/TMA building blocks used by mgnns_b200/csrc/tc_gemm.cu.
// nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o tc_probe tools/tc_probe.cu ; ./tc_probe
#include <cstdio>
#include <vector>
#include <cmath>
#include <cstdarg>
#define MGNNS_TC_PROBE 1
#include "../mgnns_b200/csrc/tc_gemm.cu"

using namespace mgnns::tc;

namespace mgnns { void set_error(const char* fmt, ...) { va_list ap; va_start(ap, fmt); vfprintf(stderr, fmt, ap); va_end(ap); fprintf(stderr, "\n"); }
                  void count_launch(int) {} }

// One CTA: TMA-load a K-major A tile [128 x 32] and a K-major B tile [160 x 32] from plain row-major
// matrices, dump shared memory, run 4 UMMAs (K = 32), dump the accumulator.
__global__ void __launch_bounds__(128, 1) probe_kernel(const __grid_constant__ CUtensorMap tmA,
                                                       const __grid_constant__ CUtensorMap tmB,
                                                       float* dump_a, float* dump_b, float* dump_d, uint32_t* info) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sa = smem;                  // 16 KB
    uint8_t* sb = smem + 16384;          // 20 KB
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 16384 + 20480);
    uint64_t* done = bar + 1;
    uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        mbar_init(done, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(slot, 256);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *slot;
    if (threadIdx.x == 0) {
        info[0] = tmem_base;
        mbar_expect_tx(bar, 16384 + 20480);
        tma_load_2d(sa, &tmA, bar, 0, 0);
        tma_load_2d(sb, &tmB, bar, 0, 0);
    }
    mbar_wait(bar, 0);
    for (int i = threadIdx.x; i < 4096; i += 128) dump_a[i] = reinterpret_cast<float*>(sa)[i];
    for (int i = threadIdx.x; i < 5120; i += 128) dump_b[i] = reinterpret_cast<float*>(sb)[i];
    __syncthreads();
    if (threadIdx.x == 0) {
        tc_fence_after();
        constexpr uint32_t idesc = instr_desc_tf32(128, 160, 0, 0);
        info[1] = idesc;
        for (int ks = 0; ks < 4; ++ks) {
            uint64_t da = smem_desc(smem_u32(sa) + ks * 32, 16, 1024);
            uint64_t db = smem_desc(smem_u32(sb) + ks * 32, 16, 1024);
            if (ks == 0) { info[2] = (uint32_t)da; info[3] = (uint32_t)(da >> 32); }
            umma_tf32(tmem_base, da, db, idesc, ks > 0);
        }
        umma_commit(done);
    }
    mbar_wait(done, 0);
    tc_fence_after();
    const int row = warp * 32 + lane;
    for (int c0 = 0; c0 < 160; c0 += 16) {
        uint32_t r[16];
        tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + c0, r);
        tmem_ld_wait();
        for (int j = 0; j < 16; ++j) dump_d[row * 160 + c0 + j] = __uint_as_float(r[j]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 256);
}

int main() {
    const int M = 128, N = 160, K = 32;
    std::vector<float> A(M * K), B(N * K);
    for (int i = 0; i < M; ++i) for (int k = 0; k < K; ++k) A[i * K + k] = (float)((i * 7 + k * 3) % 11) - 5.f;
    for (int j = 0; j < N; ++j) for (int k = 0; k < K; ++k) B[j * K + k] = (float)((j * 5 + k) % 7) - 3.f;
    float *dA, *dB, *da, *db, *dd; uint32_t* dinfo;
    cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4);
    cudaMalloc(&da, 4096 * 4); cudaMalloc(&db, 5120 * 4); cudaMalloc(&dd, M * N * 4); cudaMalloc(&dinfo, 64);
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
    cudaMemset(dd, 0xff, M * N * 4);
    CUtensorMap ma, mb;
    { uint64_t dims[2] = {(uint64_t)K, (uint64_t)M}; uint64_t str[1] = {(uint64_t)K * 4}; uint32_t box[2] = {32, 128};
      if (make_map(&ma, dA, 2, dims, str, box)) return 1; }
    { uint64_t dims[2] = {(uint64_t)K, (uint64_t)N}; uint64_t str[1] = {(uint64_t)K * 4}; uint32_t box[2] = {32, 160};
      if (make_map(&mb, dB, 2, dims, str, box)) return 1; }
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * 1024);
    probe_kernel<<<1, 128, 40 * 1024>>>(ma, mb, da, db, dd, dinfo);
    cudaError_t e = cudaDeviceSynchronize();
    printf("kernel: %s\n", cudaGetErrorString(e));
    std::vector<float> ha(4096), hb(5120), hd(M * N); uint32_t info[16];
    cudaMemcpy(ha.data(), da, 4096 * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(hb.data(), db, 5120 * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(hd.data(), dd, M * N * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(info, dinfo, 64, cudaMemcpyDeviceToHost);
    printf("tmem_base=0x%08x idesc=0x%08x adesc=0x%08x_%08x\n", info[0], info[1], info[3], info[2]);
    // check swizzled smem image of A: element (row r, k) lives at r*128 + ((k/4) ^ (r%8))*16 + (k%4)*4
    int bad = 0;
    for (int r = 0; r < M; ++r) for (int k = 0; k < K; ++k) {
        int off = r * 32 + (((k / 4) ^ (r % 8)) * 4) + (k % 4);
        if (ha[off] != A[r * K + k]) ++bad;
    }
    printf("smem A image mismatches vs expected 128B swizzle: %d of %d (first floats: %g %g %g %g)\n", bad, M * K, ha[0], ha[1], ha[2], ha[3]);
    double maxerr = 0; int nz = 0;
    for (int i = 0; i < M; ++i) for (int j = 0; j < N; ++j) {
        double ref = 0; for (int k = 0; k < K; ++k) ref += (double)A[i * K + k] * B[j * K + k];
        maxerr = fmax(maxerr, fabs(ref - hd[i * N + j])); if (hd[i * N + j] != 0) ++nz;
    }
    printf("D: max err %g, nonzeros %d of %d, D[0][0..3] = %g %g %g %g\n", maxerr, nz, M * N, hd[0], hd[1], hd[2], hd[3]);
    return 0;
}
