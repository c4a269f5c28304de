// TextLevelGCN channel (ref: models/Text_GCN.py:142-275), one CTA per document.
//
// The reference builds a DGL graph per document on the host: nodes = unique
// tokens (PAD included), edges (src=tok[p], dst=tok[q]) for |p-q| <= ngram over
// the PAD-stripped text plus an explicit self loop, edge weight =
// seq_edge_w[edges_matrix[src,dst]], then h'[v] = max over in-edges of w*h[u]
// and the document vector is the sum over nodes (nodes without in-edges -> 0).
//
// Here no graph is materialised.  Because the window is symmetric, the in-edges
// of the node for word v are exactly {(tok[q] -> v) : tok[p]==v, |p-q|<=ngram};
// the CTA compacts the document, looks the edge ids up in the CSR form of the
// PMI map (binary search), and every thread owns a slice of the feature
// dimension, looping over unique words and their windows.  The backward kernel
// recomputes the arg-max instead of storing it.
#include "common.cuh"

namespace mgnns {

constexpr int TG_THREADS = 320;

struct TextArgs {
    const int64_t* doc_ids; int B, L, max_length, ngram;
    const float* node_hidden; int V, F;
    const float* edge_w; int64_t n_edge_w;
    const int32_t* rowptr; const int32_t* col; const int32_t* eid;
    int apply_relu;
};

__device__ __forceinline__ int lookup_eid(const TextArgs& a, int src, int dst) {
    int lo = a.rowptr[src], hi = a.rowptr[src + 1];
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        int c = a.col[mid];
        if (c < dst) lo = mid + 1; else hi = mid;
    }
    if (lo < a.rowptr[src + 1] && a.col[lo] == dst) return a.eid ? a.eid[lo] : lo + 1;
    return 0;
}

// Shared layout (dynamic): seq[Lc] int, next_same[Lc] int, first[Lc] int,
// w[Lc*(2n+1)] float, e[Lc*(2n+1)] int, (bwd) gw[Lc*(2n+1)] float
__device__ __forceinline__ int prepare_document(const TextArgs& a, int b, int* seq, int* next_same, int* first,
                                                float* wgt, int* eids, int* s_n) {
    const int Lc = min(a.L, a.max_length);
    const int W = 2 * a.ngram + 1;
    if (threadIdx.x < 32) {
        // stable compaction of non-PAD tokens by warp 0
        int n = 0;
        for (int base = 0; base < Lc; base += 32) {
            int i = base + threadIdx.x;
            long long t = (i < Lc) ? a.doc_ids[(int64_t)b * a.L + i] : 0;
            bool k = (i < Lc) && (t != 0);
            unsigned m = __ballot_sync(0xffffffffu, k);
            if (k) {
                // ids outside the vocabulary would fault the embedding gather; clamp instead
                long long tc = t < 0 ? 0 : (t >= a.V ? (long long)a.V - 1 : t);
                seq[n + __popc(m & ((1u << threadIdx.x) - 1u))] = (int)tc;
            }
            n += __popc(m);
        }
        if (threadIdx.x == 0) *s_n = n;
    }
    __syncthreads();
    const int n = *s_n;
    for (int p = threadIdx.x; p < n; p += blockDim.x) {
        int v = seq[p];
        int nx = -1;
        for (int q = p + 1; q < n; ++q)
            if (seq[q] == v) { nx = q; break; }
        next_same[p] = nx;
        int f = 1;
        for (int q = 0; q < p; ++q)
            if (seq[q] == v) { f = 0; break; }
        first[p] = f;
    }
    // edge (src = seq[q]) -> (dst = seq[p]) for q = p - ngram + d
    for (int i = threadIdx.x; i < n * W; i += blockDim.x) {
        int p = i / W, d = i - p * W;
        int q = p - a.ngram + d;
        int id = 0;
        float w = 0.f;
        if (q >= 0 && q < n) {
            int src = seq[q], dst = seq[p];
            if (src >= 0 && src < a.V && dst >= 0 && dst < a.V) id = lookup_eid(a, src, dst);
            w = (id >= 0 && id < a.n_edge_w) ? a.edge_w[id] : 0.f;
        }
        eids[i] = id;
        wgt[i] = w;
    }
    __syncthreads();
    return n;
}

__global__ void __launch_bounds__(TG_THREADS) text_maxagg_fwd_kernel(TextArgs a, float* __restrict__ out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int Lc = min(a.L, a.max_length);
    const int W = 2 * a.ngram + 1;
    int* seq = reinterpret_cast<int*>(smem_raw);
    int* next_same = seq + Lc;
    int* first = next_same + Lc;
    float* wgt = reinterpret_cast<float*>(first + Lc);
    int* eids = reinterpret_cast<int*>(wgt + Lc * W);
    __shared__ int s_n;
    const int b = blockIdx.x;
    const int n = prepare_document(a, b, seq, next_same, first, wgt, eids, &s_n);

    for (int f = threadIdx.x; f < a.F; f += blockDim.x) {
        float total = 0.f;
        for (int p0 = 0; p0 < n; ++p0) {
            if (!first[p0]) continue;
            float best = -INFINITY;
            for (int p = p0; p >= 0; p = next_same[p]) {
                const int qlo = max(0, p - a.ngram), qhi = min(n - 1, p + a.ngram);
                for (int q = qlo; q <= qhi; ++q) {
                    float h = __ldg(a.node_hidden + (int64_t)seq[q] * a.F + f);
                    float m = wgt[p * W + (q - p + a.ngram)] * h;
                    best = fmaxf(best, m);
                }
            }
            total += best;
        }
        if (a.apply_relu) total = fmaxf(total, 0.f);
        out[(int64_t)b * a.F + f] = total;
    }
}

__global__ void __launch_bounds__(TG_THREADS) text_maxagg_bwd_kernel(
    TextArgs a, const float* __restrict__ out, const float* __restrict__ grad_out,
    float* __restrict__ grad_h, float* __restrict__ grad_w) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int Lc = min(a.L, a.max_length);
    const int W = 2 * a.ngram + 1;
    int* seq = reinterpret_cast<int*>(smem_raw);
    int* next_same = seq + Lc;
    int* first = next_same + Lc;
    float* wgt = reinterpret_cast<float*>(first + Lc);
    int* eids = reinterpret_cast<int*>(wgt + Lc * W);
    float* gw = reinterpret_cast<float*>(eids + Lc * W);
    __shared__ int s_n;
    const int b = blockIdx.x;
    for (int i = threadIdx.x; i < Lc * W; i += blockDim.x) gw[i] = 0.f;
    const int n = prepare_document(a, b, seq, next_same, first, wgt, eids, &s_n);

    for (int f = threadIdx.x; f < a.F; f += blockDim.x) {
        float g = grad_out[(int64_t)b * a.F + f];
        if (a.apply_relu && !(out[(int64_t)b * a.F + f] > 0.f)) g = 0.f;
        if (g == 0.f) continue;
        for (int p0 = 0; p0 < n; ++p0) {
            if (!first[p0]) continue;
            float best = -INFINITY, best_h = 0.f;
            int best_slot = -1, best_src = 0;
            for (int p = p0; p >= 0; p = next_same[p]) {
                const int qlo = max(0, p - a.ngram), qhi = min(n - 1, p + a.ngram);
                for (int q = qlo; q <= qhi; ++q) {
                    float h = __ldg(a.node_hidden + (int64_t)seq[q] * a.F + f);
                    int slot = p * W + (q - p + a.ngram);
                    float m = wgt[slot] * h;
                    if (m > best || best_slot < 0) { best = m; best_slot = slot; best_src = seq[q]; best_h = h; }
                }
            }
            if (best_slot >= 0) {
                atomicAdd(grad_h + (int64_t)best_src * a.F + f, wgt[best_slot] * g);
                atomicAdd(gw + best_slot, best_h * g);
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n * W; i += blockDim.x) {
        float v = gw[i];
        int id = eids[i];
        if (v != 0.f && id >= 0 && id < a.n_edge_w) atomicAdd(grad_w + id, v);
    }
}

static int check_text_args(const TextArgs& a) {
    MG_REQUIRE(a.B >= 0 && a.L >= 1 && a.F >= 1 && a.V >= 1, "text_maxagg: bad dimensions");
    MG_REQUIRE(a.ngram >= 0 && a.ngram <= 64, "text_maxagg: ngram %d out of range [0,64]", a.ngram);
    MG_REQUIRE(a.max_length >= 1, "text_maxagg: max_length must be >= 1");
    MG_REQUIRE(a.doc_ids && a.node_hidden && a.edge_w && a.rowptr, "text_maxagg: null pointer");
    MG_REQUIRE(a.n_edge_w >= 1, "text_maxagg: edge weight table is empty");
    return 0;
}

}  // namespace mgnns

using namespace mgnns;

extern "C" int mgnns_text_maxagg_fwd(const int64_t* doc_ids, int B, int L, int max_length, int ngram,
                                     const float* node_hidden, int V, int F,
                                     const float* edge_w, int64_t n_edge_w,
                                     const int32_t* pmi_rowptr, const int32_t* pmi_col, const int32_t* pmi_eid,
                                     int apply_relu, float* out, void* stream) {
    TextArgs a{doc_ids, B, L, max_length, ngram, node_hidden, V, F, edge_w, n_edge_w,
               pmi_rowptr, pmi_col, pmi_eid, apply_relu};
    if (int rc = check_text_args(a)) return rc;
    if (B == 0) return 0;
    MG_REQUIRE(out, "text_maxagg_fwd: null output");
    const int Lc = L < max_length ? L : max_length;
    const int W = 2 * ngram + 1;
    size_t smem = (size_t)Lc * 3 * sizeof(int) + (size_t)Lc * W * (sizeof(float) + sizeof(int));
    MG_REQUIRE(smem <= 200 * 1024, "text_maxagg_fwd: document window table too large (%zu B)", smem);
    if (smem > 48 * 1024)
        cudaFuncSetAttribute(text_maxagg_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    text_maxagg_fwd_kernel<<<B, TG_THREADS, smem, as_stream(stream)>>>(a, out);
    MG_LAUNCH_CHECK("text_maxagg_fwd");
    return 0;
}

extern "C" int mgnns_text_maxagg_bwd(const int64_t* doc_ids, int B, int L, int max_length, int ngram,
                                     const float* node_hidden, int V, int F,
                                     const float* edge_w, int64_t n_edge_w,
                                     const int32_t* pmi_rowptr, const int32_t* pmi_col, const int32_t* pmi_eid,
                                     int apply_relu, const float* out, const float* grad_out,
                                     float* grad_node_hidden, float* grad_edge_w, void* stream) {
    TextArgs a{doc_ids, B, L, max_length, ngram, node_hidden, V, F, edge_w, n_edge_w,
               pmi_rowptr, pmi_col, pmi_eid, apply_relu};
    if (int rc = check_text_args(a)) return rc;
    if (B == 0) return 0;
    MG_REQUIRE(grad_out && grad_node_hidden && grad_edge_w, "text_maxagg_bwd: null pointer");
    MG_REQUIRE(!apply_relu || out, "text_maxagg_bwd: forward output needed for the ReLU mask");
    const int Lc = L < max_length ? L : max_length;
    const int W = 2 * ngram + 1;
    size_t smem = (size_t)Lc * 3 * sizeof(int) + (size_t)Lc * W * (2 * sizeof(float) + sizeof(int));
    MG_REQUIRE(smem <= 200 * 1024, "text_maxagg_bwd: document window table too large (%zu B)", smem);
    if (smem > 48 * 1024)
        cudaFuncSetAttribute(text_maxagg_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    text_maxagg_bwd_kernel<<<B, TG_THREADS, smem, as_stream(stream)>>>(a, out, grad_out, grad_node_hidden, grad_edge_w);
    MG_LAUNCH_CHECK("text_maxagg_bwd");
    return 0;
}
