// aggregate.cu -- K6: aggregate device power over the horizon, the one quantity agents exchange.
//   P_agg[k] = sum_b P_nom[b] * u[b, k]
// replaces GridAgentMpc.get_grid_device_powers_N_tilde + the grid model's D4 = ones(1, n_dev)
// (reference: examples/.../micro_grid_agents.py:625-646, 410-420; micro_grid_models.py:143).
// Deterministic two-pass tree (no atomics) so that a run is bit-reproducible: pass 1 reduces chunks of 16
// agents (thread <-> horizon step, coalesced along k), pass 2 sums the chunk partials in chunk order.
// Across GPUs the [Nt] result is all-reduced with NCCL on the same stream by the host layer.
#include "common.cuh"

namespace hmpc {
constexpr int kAggChunk = 16;   // agents per CTA of pass 1: short dependent chains, many CTAs in flight

__global__ void __launch_bounds__(128) aggregate_partial_kernel(int B, int Nt, const double* __restrict__ u,
                                                                int64_t sb, int sk, const double* __restrict__ P_nom,
                                                                double* __restrict__ partial) {
    const int chunk = blockIdx.x;
    const int b0 = chunk * kAggChunk, b1 = min(B, b0 + kAggChunk);
    for (int k = threadIdx.x; k < Nt; k += blockDim.x) {
        double val[kAggChunk];
#pragma unroll
        for (int i = 0; i < kAggChunk; ++i) {           // all loads in flight before the (ordered) summation
            const int b = b0 + i;
            val[i] = b < b1 ? (P_nom ? P_nom[b] : 1.0) * u[(int64_t)b * sb + (int64_t)k * sk] : 0.0;
        }
        double acc = 0.0;
#pragma unroll
        for (int i = 0; i < kAggChunk; ++i) acc += val[i];
        partial[(int64_t)chunk * Nt + k] = acc;
    }
}

__global__ void __launch_bounds__(128) aggregate_final_kernel(int chunks, int Nt, const double* __restrict__ partial,
                                                              double* __restrict__ out) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= Nt) return;
    double acc = 0.0;
    int c = 0;
    for (; c + 8 <= chunks; c += 8) {
        double v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = partial[(int64_t)(c + i) * Nt + k];
#pragma unroll
        for (int i = 0; i < 8; ++i) acc += v[i];
    }
    for (; c < chunks; ++c) acc += partial[(int64_t)c * Nt + k];
    out[k] = acc;
}

// FP64 FMA peak probe: 8 independent dependency chains per thread, 4096 FMAs each.
__global__ void __launch_bounds__(256) fp64_peak_kernel(double* sink, int iters) {
    double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6,
           a7 = a0 + 7;
    const double m = 1.0000001, c = 1e-7;
    for (int i = 0; i < iters; ++i) {
        a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
        a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
    }
    if (a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 == 12345.678) sink[0] = a0;
}
}  // namespace hmpc

extern "C" int hmpc_aggregate_power_f64(int32_t B, int32_t Nt, const double* u, int64_t u_stride_b,
                                        int32_t u_stride_k, const double* P_nom, double* partial, double* P_agg,
                                        void* stream) {
    using namespace hmpc;
    if (B < 0 || Nt < 0 || !u || !partial || !P_agg) return HMPC_ERR_ARG;
    if (Nt == 0) return HMPC_OK;
    const int chunks = B > 0 ? ceil_div(B, kAggChunk) : 0;
    if (chunks) {
        aggregate_partial_kernel<<<chunks, 128, 0, (cudaStream_t)stream>>>(B, Nt, u, u_stride_b, u_stride_k, P_nom,
                                                                           partial);
        HMPC_LAUNCH_CHECK("aggregate_partial_kernel");
    }
    aggregate_final_kernel<<<ceil_div(Nt, 128), 128, 0, (cudaStream_t)stream>>>(chunks, Nt, partial, P_agg);
    HMPC_LAUNCH_CHECK("aggregate_final_kernel");
    return HMPC_OK;
}

extern "C" int hmpc_fp64_peak_probe(double* tflops, void* stream) {
    using namespace hmpc;
    if (!tflops) return HMPC_ERR_ARG;
    double* sink = nullptr;
    HMPC_CUDA_TRY(cudaMalloc(&sink, sizeof(double)));
    cudaEvent_t e0, e1;
    HMPC_CUDA_TRY(cudaEventCreate(&e0));
    HMPC_CUDA_TRY(cudaEventCreate(&e1));
    const int iters = 4096, blocks = kNumSM * 8;
    cudaStream_t s = (cudaStream_t)stream;
    fp64_peak_kernel<<<blocks, 256, 0, s>>>(sink, iters);  // warm-up
    double best = 0.0;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0, s);
        fp64_peak_kernel<<<blocks, 256, 0, s>>>(sink, iters);
        cudaEventRecord(e1, s);
        HMPC_CUDA_TRY(cudaEventSynchronize(e1));
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        const double fl = 2.0 * 8.0 * iters * 256.0 * blocks;
        if (ms > 0) best = fmax(best, fl / (ms * 1e-3) / 1e12);
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(sink);
    *tflops = best;
    return HMPC_OK;
}
