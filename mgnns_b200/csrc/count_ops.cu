// Integer counting kernels behind the training loop's bookkeeping:
//   * per-batch confusion matrix from the class scores, on the device, so the engine does not copy predictions to
//     the host and call sklearn every batch (ref: engine/Multi_GCN_Multihead_Att_engine.py:831-838);
//   * label co-occurrence counts for the label-graph adjacency (ref: utils/util.py:336-357, generate_nums /
//     generate_Adj: nums[j] = images containing label j, Adj[a][b] = images containing both, a != b).
// Counts are exact integers (atomics), independent of arrival order.
#include "common.cuh"

namespace mgnns {

// one warp per sample: arg-max over C scores (first maximum wins, like torch.argmax on distinct values),
// conf[target*C + pred] += 1, pred_out[b] = pred
__global__ void __launch_bounds__(256) confusion_count_kernel(const float* __restrict__ scores, int64_t ld,
                                                              const int64_t* __restrict__ target, int B, int C,
                                                              int32_t* __restrict__ conf, int64_t* __restrict__ pred_out) {
    const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (b >= B) return;
    const int lane = threadIdx.x & 31;
    const float* row = scores + (int64_t)b * ld;
    float best = -INFINITY;
    int arg = 0x7fffffff;
    for (int c = lane; c < C; c += 32) {
        const float v = row[c];
        if (v > best || (v == best && c < arg)) { best = v; arg = c; }       // NaN compares false: never wins
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
        if (ob > best || (ob == best && oa < arg)) { best = ob; arg = oa; }
    }
    if (lane == 0) {
        if (arg == 0x7fffffff) arg = 0;                    // all-NaN row: class 0, as torch.argmax's first index
        const int64_t t = target[b];
        if (t >= 0 && t < C) atomicAdd(conf + t * C + arg, 1);
        if (pred_out != nullptr) pred_out[b] = arg;
    }
}

// one warp per image: labels[i, 0..len_i) (row stride max_len; entries outside [0,C) are skipped)
__global__ void __launch_bounds__(256) label_cooccurrence_kernel(const int32_t* __restrict__ labels,
                                                                 const int32_t* __restrict__ lens, int64_t n_images,
                                                                 int max_len, int C,
                                                                 unsigned long long* __restrict__ nums,
                                                                 unsigned long long* __restrict__ adj) {
    const int64_t img = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (img >= n_images) return;
    const int lane = threadIdx.x & 31;
    const int32_t* row = labels + img * max_len;
    const int n = min(lens[img], max_len);
    for (int j = lane; j < n; j += 32) {
        const int a = row[j];
        if (a >= 0 && a < C) atomicAdd(nums + a, 1ull);
    }
    const int pairs = n * n;
    for (int p = lane; p < pairs; p += 32) {
        const int a = row[p / n], b = row[p % n];
        if (a != b && a >= 0 && a < C && b >= 0 && b < C) atomicAdd(adj + (int64_t)a * C + b, 1ull);
    }
}

}  // namespace mgnns

using namespace mgnns;

extern "C" int mgnns_confusion_count(const float* scores, int64_t ld, const int64_t* target, int B, int C,
                                     int32_t* conf, int64_t* pred_out, void* stream) {
    MG_REQUIRE(B >= 0 && C >= 1 && ld >= C, "confusion_count: bad dimensions");
    if (B == 0) return 0;
    MG_REQUIRE(scores && target && conf, "confusion_count: null pointer");
    confusion_count_kernel<<<(B + 7) / 8, 256, 0, as_stream(stream)>>>(scores, ld, target, B, C, conf, pred_out);
    MG_LAUNCH_CHECK("confusion_count");
    return 0;
}

extern "C" int mgnns_label_cooccurrence(const int32_t* labels, const int32_t* lens, int64_t n_images, int max_len,
                                        int C, int64_t* nums, int64_t* adj, void* stream) {
    MG_REQUIRE(n_images >= 0 && max_len >= 0 && C >= 1, "label_cooccurrence: bad dimensions");
    if (n_images == 0 || max_len == 0) return 0;
    MG_REQUIRE(labels && lens && nums && adj, "label_cooccurrence: null pointer");
    const int64_t blocks = (n_images + 7) / 8;
    MG_REQUIRE(blocks < (1LL << 31), "label_cooccurrence: too many images for one launch");
    label_cooccurrence_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(
        labels, lens, n_images, max_len, C, reinterpret_cast<unsigned long long*>(nums),
        reinterpret_cast<unsigned long long*>(adj));
    MG_LAUNCH_CHECK("label_cooccurrence");
    return 0;
}
