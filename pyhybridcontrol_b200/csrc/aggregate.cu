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

// ---- K6 across the GPUs of one box without a collective library.  Every rank owns an exchange WINDOW in memory its peers
// can address (torch symmetric memory: NVLink peer mappings); the last pass of the local reduction writes this rank's
// [Nt] sums straight into the window of EVERY rank (plain stores to peer memory, 392 B per peer at N_p = 48), fences,
// and raises that step's flag in each of them.  The gather kernel of a rank waits for the world's flags of its own
// latest step (bounded spin on local memory) and adds the contributions in rank order (bit-reproducible).  Both are
// ordinary kernels on the step's stream: they sit inside the step's CUDA graph, no host call per step.
// Window (doubles unless noted): [0] step counter of the owner (u64), [1] error word (u64), [2 .. 2 + RING*W) flags
// (u64, [slot][rank]), then data [slot][rank][Nt].
constexpr int kXRing = 8;
__host__ __device__ inline int64_t xwin_flags(int slot, int world, int r) { return 2 + (int64_t)slot * world + r; }
__host__ __device__ inline int64_t xwin_data(int slot, int world, int r, int Nt) {
    return 2 + (int64_t)kXRing * world + ((int64_t)slot * world + r) * Nt;
}

__global__ void __launch_bounds__(128) aggregate_publish_kernel(int chunks, int Nt, const double* __restrict__ partial,
                                                                int world, int rank, double* const* __restrict__ windows,
                                                                double* __restrict__ out_prev, long long spin_limit, int lag) {
    unsigned long long* mine = reinterpret_cast<unsigned long long*>(windows[rank]);
    const unsigned long long seq = mine[0] + 1ull;
    const int slot = (int)(seq % kXRing);
    for (int k = threadIdx.x; k < Nt; k += blockDim.x) {
        double acc = 0.0;
        for (int c = 0; c < chunks; ++c) acc += partial[(int64_t)c * Nt + k];          // chunk order: deterministic
        for (int r = 0; r < world; ++r) windows[r][xwin_data(slot, world, rank, Nt) + k] = acc;
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x < world) {
        unsigned long long* flag = reinterpret_cast<unsigned long long*>(windows[threadIdx.x]) + xwin_flags(slot, world, rank);
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(flag), "l"(seq) : "memory");
    }
    if (threadIdx.x == 0) mine[0] = seq;
    // the same launch gathers an EARLIER step (pipelined loops: the peers' contributions of step seq - lag arrived long
    // ago, nobody waits; the ranks may drift `lag` steps apart) -- one launch less per control step
    if (out_prev) {
        if (seq <= (unsigned long long)lag) { for (int k = threadIdx.x; k < Nt; k += blockDim.x) out_prev[k] = 0.0; return; }
        const unsigned long long want = seq - (unsigned long long)lag;
        const int pslot = (int)(want % kXRing);
        __shared__ int s_bad;
        if (threadIdx.x == 0) s_bad = 0;
        __syncthreads();
        if (threadIdx.x < world) {
            const unsigned long long* flag = mine + xwin_flags(pslot, world, threadIdx.x);
            unsigned long long seen = 0;
            long long n = 0;
            do {
                asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(flag) : "memory");
            } while (seen < want && ++n < spin_limit);
            if (seen < want) { s_bad = 1; mine[1] = want; }
        }
        __syncthreads();
        const double* win = windows[rank];
        for (int k = threadIdx.x; k < Nt; k += blockDim.x) {
            double acc = 0.0;
            for (int r = 0; r < world; ++r) acc += win[xwin_data(pslot, world, r, Nt) + k];
            out_prev[k] = s_bad ? nan("") : acc;
        }
    }
}

__global__ void __launch_bounds__(128) aggregate_gather_kernel(int Nt, int world, int rank, double* __restrict__ window,
                                                               double* __restrict__ out, long long spin_limit, int lag) {
    unsigned long long* mine = reinterpret_cast<unsigned long long*>(window);
    if (mine[0] <= (unsigned long long)lag) {                 // nothing that old has been published yet
        for (int k = threadIdx.x; k < Nt; k += blockDim.x) out[k] = 0.0;
        return;
    }
    const unsigned long long seq = mine[0] - (unsigned long long)lag;
    const int slot = (int)(seq % kXRing);
    __shared__ int s_bad;
    if (threadIdx.x == 0) s_bad = 0;
    __syncthreads();
    if (threadIdx.x < world) {
        const unsigned long long* flag = mine + xwin_flags(slot, world, threadIdx.x);
        unsigned long long seen = 0;
        long long n = 0;
        do {
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(flag) : "memory");
        } while (seen < seq && ++n < spin_limit);
        if (seen < seq) { s_bad = 1; mine[1] = seq; }         // a peer never arrived: report instead of hanging
    }
    __syncthreads();
    for (int k = threadIdx.x; k < Nt; k += blockDim.x) {
        double acc = 0.0;
        for (int r = 0; r < world; ++r) acc += window[xwin_data(slot, world, r, Nt) + k];
        out[k] = s_bad ? nan("") : acc;
    }
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

extern "C" int64_t hmpc_aggregate_window_doubles(int32_t Nt, int32_t world) {
    return hmpc::xwin_data(hmpc::kXRing, world, 0, Nt);
}

extern "C" int hmpc_aggregate_publish_f64(int32_t B, int32_t Nt, const double* u, int64_t u_stride_b,
                                          int32_t u_stride_k, const double* P_nom, double* partial, int32_t world,
                                          int32_t rank, double* const* windows, double* P_total_prev, int32_t lag,
                                          void* stream) {
    using namespace hmpc;
    if (B < 0 || Nt < 1 || !u || !partial || !windows || world < 1 || world > 128 || rank < 0 || rank >= world) return HMPC_ERR_ARG;
    if (P_total_prev && (lag < 1 || lag > kXRing - 2)) return HMPC_ERR_ARG;
    const int chunks = B > 0 ? ceil_div(B, kAggChunk) : 0;
    if (chunks) {
        aggregate_partial_kernel<<<chunks, 128, 0, (cudaStream_t)stream>>>(B, Nt, u, u_stride_b, u_stride_k, P_nom,
                                                                           partial);
        HMPC_LAUNCH_CHECK("aggregate_partial_kernel");
    }
    aggregate_publish_kernel<<<1, 128, 0, (cudaStream_t)stream>>>(chunks, Nt, partial, world, rank, windows, P_total_prev,
                                                                  400000000ll, lag);
    HMPC_LAUNCH_CHECK("aggregate_publish_kernel");
    return HMPC_OK;
}

extern "C" int hmpc_aggregate_gather_f64(int32_t Nt, int32_t world, int32_t rank, double* window, double* P_total,
                                         int64_t spin_limit, int32_t lag, void* stream) {
    using namespace hmpc;
    if (Nt < 1 || !window || !P_total || world < 1 || world > 128 || rank < 0 || rank >= world || lag < 0 || lag > kXRing - 2) return HMPC_ERR_ARG;
    aggregate_gather_kernel<<<1, 128, 0, (cudaStream_t)stream>>>(Nt, world, rank, window, P_total,
                                                                 spin_limit > 0 ? spin_limit : 400000000ll, lag);
    HMPC_LAUNCH_CHECK("aggregate_gather_kernel");
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
