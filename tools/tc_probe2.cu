// Probe 2: the exact FWD stage (A = F[c][p] MN-major via a 3D map, B = W[o][c] K-major, N = 160 + 144).
#include <cstdio>
#include <vector>
#include <cmath>
#include <cstdarg>
#include "../mgnns_b200/csrc/tc_gemm.cu"
using namespace mgnns::tc;
namespace mgnns { void set_error(const char* fmt, ...) { va_list ap; va_start(ap, fmt); vfprintf(stderr, fmt, ap); va_end(ap); fprintf(stderr, "\n"); }
                  void count_launch(int) {} }

__global__ void __launch_bounds__(128, 1) probe_kernel(const __grid_constant__ CUtensorMap tmA,
                                                       const __grid_constant__ CUtensorMap tmB,
                                                       float* dump_a, float* dump_d, int mt) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sa = smem;                  // 16 KB
    uint8_t* sb = smem + 16384;          // 38 KB
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 16384 + 38912);
    uint64_t* done = bar + 1;
    uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_init(done, 1); fence_barrier_init(); }
    if (warp == 1) tmem_alloc(slot, 512);
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tmem_base = *slot;
    if (threadIdx.x == 0) {
        mbar_expect_tx(bar, 16384 + 38912);
        for (int j = 0; j < 4; ++j) tma_load_3d(sa + j * 4096, &tmA, bar, mt * 128 + j * 32, 0, 0);
        tma_load_2d(sb, &tmB, bar, 0, 0);
        tma_load_2d(sb + 152 * 128, &tmB, bar, 0, 152);
    }
    mbar_wait(bar, 0);
    for (int i = threadIdx.x; i < 4096; i += 128) dump_a[i] = reinterpret_cast<float*>(sa)[i];
    __syncthreads();
    if (threadIdx.x == 0) {
        tc_fence_after();
        constexpr uint32_t id0 = instr_desc_tf32(128, 160, 1, 0), id1 = instr_desc_tf32(128, 144, 1, 0);
        for (int ks = 0; ks < 4; ++ks) {
            uint64_t da = smem_desc(smem_u32(sa) + ks * 1024, 4096, 512, 1);
            uint64_t db0 = smem_desc(smem_u32(sb) + ks * 32, 16, 1024);
            uint64_t db1 = smem_desc(smem_u32(sb) + 160 * 128 + ks * 32, 16, 1024);
            umma_tf32(tmem_base, da, db0, id0, ks > 0);
            umma_tf32(tmem_base + 160, da, db1, id1, ks > 0);
        }
        umma_commit(done);
    }
    mbar_wait(done, 0);
    tc_fence_after();
    const int row = warp * 32 + lane;
    for (int c0 = 0; c0 < 304; c0 += 16) {
        uint32_t r[16];
        tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + c0, r);
        tmem_ld_wait();
        for (int j = 0; j < 16; ++j) dump_d[row * 304 + c0 + j] = __uint_as_float(r[j]);
    }
    tc_fence_before(); __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 512);
}

int main() {
    const int C = 32, P = 196, O = 300;
    std::vector<float> F(C * P), W(O * C);
    for (int c = 0; c < C; ++c) for (int p = 0; p < P; ++p) F[c * P + p] = (float)((p * 7 + c * 3) % 11) - 5.f;
    for (int o = 0; o < O; ++o) for (int c = 0; c < C; ++c) W[o * C + c] = (float)((o * 5 + c) % 7) - 3.f;
    float *dF, *dW, *da, *dd;
    cudaMalloc(&dF, F.size() * 4); cudaMalloc(&dW, W.size() * 4); cudaMalloc(&da, 4096 * 4); cudaMalloc(&dd, 128 * 304 * 4);
    cudaMemcpy(dF, F.data(), F.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dW, W.data(), W.size() * 4, cudaMemcpyHostToDevice);
    CUtensorMap ma, mb;
    { uint64_t dims[3] = {(uint64_t)P, (uint64_t)C, 1}; uint64_t str[2] = {(uint64_t)P * 4, (uint64_t)C * P * 4}; uint32_t box[3] = {32, 32, 1};
      if (make_map(&ma, dF, 3, dims, str, box, true)) return 1; }
    { uint64_t dims[2] = {(uint64_t)C, (uint64_t)O}; uint64_t str[1] = {(uint64_t)C * 4}; uint32_t box[2] = {32, 152};
      if (make_map(&mb, dW, 2, dims, str, box)) return 1; }
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    for (int mt = 0; mt < 2; ++mt) {
        cudaMemset(dd, 0xff, 128 * 304 * 4);
        probe_kernel<<<1, 128, 60 * 1024>>>(ma, mb, da, dd, mt);
        cudaError_t e = cudaDeviceSynchronize();
        printf("mt=%d kernel: %s\n", mt, cudaGetErrorString(e));
        std::vector<float> ha(4096), hd(128 * 304);
        cudaMemcpy(ha.data(), da, 4096 * 4, cudaMemcpyDeviceToHost);
        cudaMemcpy(hd.data(), dd, 128 * 304 * 4, cudaMemcpyDeviceToHost);
        // A image: atom j (32 p) at j*4096 B; k-row c at c*128 B; inside the row, 32B chunk index (pp/8) ^ (c%4)
        int bad = 0, zeros = 0;
        for (int j = 0; j < 4; ++j) for (int c = 0; c < 32; ++c) for (int pp = 0; pp < 32; ++pp) {
            int p = mt * 128 + j * 32 + pp;
            float want = p < P ? F[c * P + p] : 0.f;
            float got = ha[j * 1024 + c * 32 + (((pp / 8) ^ (c % 4)) * 8) + (pp % 8)];
            if (got != want) ++bad;
            if (got == 0.f) ++zeros;
        }
        printf("  smem A mismatches %d of 4096 (zeros %d)\n", bad, zeros);
        double maxerr = 0; int nz = 0;
        for (int i = 0; i < 128; ++i) for (int o = 0; o < 304; ++o) {
            int p = mt * 128 + i;
            double ref = 0;
            if (p < P && o < O) for (int c = 0; c < C; ++c) ref += (double)F[c * P + p] * W[o * C + c];
            maxerr = fmax(maxerr, fabs(ref - hd[i * 304 + o])); if (hd[i * 304 + o] != 0) ++nz;
        }
        printf("  D max err %g, nonzeros %d; D[0][0..3] = %g %g %g %g\n", maxerr, nz, hd[0], hd[1], hd[2], hd[3]);
    }
    return 0;
}
