// Small dense layers and head post-processing (models/humaniflow_model.py:232-258, 116-150;
// utils/rigid_transform_utils.py:86-100).  M is the image batch (tens of rows): these are latency-bound
// GEMVs, written for coalesced weight streaming rather than tensor cores.
#include "common.cuh"
#include <algorithm>

namespace {

constexpr int LW = 8;        // warps per CTA, one output neuron each
constexpr int KC = 512;      // K chunk of the activations staged in shared memory (32 rows x KC fp32 = 64 KB)

__device__ __forceinline__ float act_fn(float a, int act) {
    if (act == 1) return a > 0.f ? a : expm1f(a);
    if (act == 2) return fmaxf(a, 0.f);
    return a;
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// y[m][o] (+)= act(sum_k x[m][k] W[o][k] + b[o]) for a tile of 32 batch rows.  K % 4 == 0, 16-byte aligned rows.
// A warp owns LNB = 4 output neurons; lanes split K (coalesced 128-bit weight loads, each weight read once per row tile); the
// 32 x KC activation chunk is staged in shared memory with cp.async, double buffered, and shared by the 8 warps of the CTA.  One
// 128-bit shared load of x feeds 16 FMAs (4 neurons x 4 k), which keeps the shared-memory pipe (4 wavefronts per load) below the
// FMA pipe; per-row partial sums are reduced across lanes with a 31-shuffle transpose-reduction per neuron.
// blockIdx.z = K-slice: with kslice < K every CTA handles k in [z * kslice, +kslice) and writes raw partial sums to
// partial[z][M][O]; linear_reduce_kernel adds the slices in a fixed order and applies bias / activation.
constexpr int LNB = 4;
__global__ void __launch_bounds__(LW * 32)
linear_vec_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ W, int ldw,
                  const float* __restrict__ b, float* __restrict__ y, int ldy, int M, int K, int O, int act,
                  int accumulate, int kslice, float* __restrict__ partial) {
    extern __shared__ __align__(16) float xs[];           // [2][32][KC]
    HF_PDL_SYNC();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int o0 = (blockIdx.x * LW + warp) * LNB;
    const int m0 = blockIdx.y * 32;
    const int kz = blockIdx.z * kslice;
    x += kz; W += kz;
    const int Kfull = K;
    K = min(kslice, K - kz);
    const float4* wrow[LNB];
#pragma unroll
    for (int n = 0; n < LNB; ++n) wrow[n] = reinterpret_cast<const float4*>(W + (size_t)min(o0 + n, O - 1) * ldw);
    auto stage = [&](int buf, int kc) {
        const int kn4 = min(KC, K - kc) >> 2;              // float4 per row in this chunk
        for (int idx = threadIdx.x; idx < 32 * kn4; idx += LW * 32) {
            const int r = idx / kn4, c4 = idx - r * kn4;
            const int m = min(m0 + r, M - 1);               // rows past M duplicate the last row (never stored)
            cp_async16(xs + ((size_t)buf * 32 + r) * KC + c4 * 4, x + (size_t)m * ldx + kc + c4 * 4);
        }
        cp_async_commit();
    };
    float acc[LNB][32];
#pragma unroll
    for (int n = 0; n < LNB; ++n)
#pragma unroll
        for (int m = 0; m < 32; ++m) acc[n][m] = 0.f;
    const int nchunks = (K + KC - 1) / KC;
    stage(0, 0);
    for (int c = 0; c < nchunks; ++c) {
        const int kc = c * KC, kn4 = min(KC, K - kc) >> 2;
        if (c + 1 < nchunks) { stage((c + 1) & 1, kc + KC); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
        __syncthreads();
        const float* xb = xs + (size_t)(c & 1) * 32 * KC;
        for (int k4 = lane; k4 < kn4; k4 += 32) {
            float4 w[LNB];
#pragma unroll
            for (int n = 0; n < LNB; ++n) w[n] = __ldg(wrow[n] + (kc >> 2) + k4);
#pragma unroll
            for (int m = 0; m < 32; ++m) {
                const float4 xv = *reinterpret_cast<const float4*>(xb + m * KC + k4 * 4);
#pragma unroll
                for (int n = 0; n < LNB; ++n)
                    acc[n][m] = fmaf(w[n].x, xv.x, fmaf(w[n].y, xv.y, fmaf(w[n].z, xv.z, fmaf(w[n].w, xv.w, acc[n][m]))));
            }
        }
        __syncthreads();
    }
    const int r = m0 + lane;
#pragma unroll
    for (int n = 0; n < LNB; ++n) {
        // transpose-reduction: after the step with stride s a lane keeps the half of the rows selected by its bit s; lane l ends
        // with the sum over all lanes of row l in acc[n][0]
#pragma unroll
        for (int sft = 16; sft > 0; sft >>= 1) {
#pragma unroll
            for (int i = 0; i < sft; ++i) {
                const bool up = (lane & sft) != 0;
                const float keep = up ? acc[n][i + sft] : acc[n][i];
                const float send = up ? acc[n][i] : acc[n][i + sft];
                acc[n][i] = keep + __shfl_xor_sync(0xffffffffu, send, sft);
            }
        }
        const float mine = acc[n][0];
        const int o = o0 + n;
        if (o < O && r < M) {
            if (kslice < Kfull) { partial[((size_t)blockIdx.z * M + r) * O + o] = mine; continue; }
            float* dst = y + (size_t)r * ldy + o;
            float a = mine + (b ? __ldg(b + o) : 0.f);
            if (accumulate) a += *dst;
            *dst = act_fn(a, act);
        }
    }
}

// y[m][o] (+)= act(sum_z partial[z][m][o] + b[o]): slices added in index order (deterministic)
__global__ void linear_reduce_kernel(const float* __restrict__ partial, int nz, const float* __restrict__ b, float* __restrict__ y, int ldy,
                                     int M, int O, int act, int accumulate) {
    HF_PDL_SYNC();
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= M * O) return;
    const int r = e / O, o = e - r * O;
    float a = 0.f;
    for (int z = 0; z < nz; ++z) a += partial[(size_t)z * M * O + e];
    a += b ? __ldg(b + o) : 0.f;
    float* dst = y + (size_t)r * ldy + o;
    if (accumulate) a += *dst;
    *dst = act_fn(a, act);
}

// Generic small-K fallback (K not a multiple of 4 or unaligned rows): lane = batch row, broadcast weight loads.
__global__ void __launch_bounds__(128)
linear_small_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ W, int ldw,
                    const float* __restrict__ b, float* __restrict__ y, int ldy, int M, int K, int O, int act,
                    int accumulate) {
    HF_PDL_SYNC();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int o = blockIdx.x * 4 + warp;
    const int r = blockIdx.y * 32 + lane;
    if (o >= O || r >= M) return;
    float a = b ? __ldg(b + o) : 0.f;
    for (int k = 0; k < K; ++k) a = fmaf(__ldg(W + (size_t)o * ldw + k), __ldg(x + (size_t)r * ldx + k), a);
    float* dst = y + (size_t)r * ldy + o;
    if (accumulate) a += *dst;
    *dst = act_fn(a, act);
}

// x.view(-1,3,2): a1 = (x0,x2,x4), a2 = (x1,x3,x5); F.normalize eps 1e-12; columns b1,b2,b3
__device__ __forceinline__ void rot6d_one(const float* x, float* o) {
    float a1[3] = {x[0], x[2], x[4]}, a2[3] = {x[1], x[3], x[5]};
    float n1 = fmaxf(sqrtf(a1[0] * a1[0] + a1[1] * a1[1] + a1[2] * a1[2]), 1e-12f);
    float b1[3] = {a1[0] / n1, a1[1] / n1, a1[2] / n1};
    float d = b1[0] * a2[0] + b1[1] * a2[1] + b1[2] * a2[2];
    float u[3] = {a2[0] - d * b1[0], a2[1] - d * b1[1], a2[2] - d * b1[2]};
    float n2 = fmaxf(sqrtf(u[0] * u[0] + u[1] * u[1] + u[2] * u[2]), 1e-12f);
    float b2[3] = {u[0] / n2, u[1] / n2, u[2] / n2};
    float b3[3] = {b1[1] * b2[2] - b1[2] * b2[1], b1[2] * b2[0] - b1[0] * b2[2], b1[0] * b2[1] - b1[1] * b2[0]};
    for (int r = 0; r < 3; ++r) { o[r * 3 + 0] = b1[r]; o[r * 3 + 1] = b2[r]; o[r * 3 + 2] = b3[r]; }
}

__global__ void rot6d_kernel(const float* __restrict__ x6, float* __restrict__ R, int n) {
    HF_PDL_SYNC();
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    rot6d_one(x6 + (size_t)i * 6, R + (size_t)i * 9);
}

// heads (B, 2*nb + 6 + 3) = [shape_mode | shape_log_std | glob6 | cam]  ->  cam + init, rot6d(glob6 + init),
// and the per-row shape table for the flow: rows [0, B*N) = mode + exp(log_std) * eps (or mode when eps == NULL),
// rows [B*N, B*N + B) = mode (point-estimate rows).  models/humaniflow_model.py:237-258.
__global__ void heads_finish_kernel(const float* __restrict__ heads, const float* __restrict__ init_glob,
                                    const float* __restrict__ init_cam, const float* __restrict__ shape_eps,
                                    int B, int N, int nb, float* __restrict__ cam, float* __restrict__ glob6,
                                    float* __restrict__ shape_rows, float* __restrict__ glob_R, float* __restrict__ shape_std) {
    HF_PDL_SYNC();
    const int ld = 2 * nb + 9;
    const int total = B * N * nb + B * nb;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
        int b, l;
        float v;
        if (e < B * N * nb) {
            const int row = e / nb;
            l = e - row * nb; b = row / N;
            const float mode = heads[b * ld + l];
            v = shape_eps ? mode + expf(heads[b * ld + nb + l]) * shape_eps[e] : mode;
        } else {
            const int q = e - B * N * nb;
            b = q / nb; l = q - b * nb;
            v = heads[b * ld + l];
        }
        shape_rows[e] = v;
    }
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < B * 6) { const int b = t / 6, c = t - b * 6; glob6[t] = heads[b * ld + 2 * nb + c] + init_glob[c]; }
    if (t < B * 3) { const int b = t / 3, c = t - b * 3; cam[t] = heads[b * ld + 2 * nb + 6 + c] + init_cam[c]; }
    if (glob_R && t < B) {            // rot6d -> matrix of the global rotation (from the heads directly: glob6 is written by other threads)
        float x6[6];
        for (int c = 0; c < 6; ++c) x6[c] = heads[t * ld + 2 * nb + c] + init_glob[c];
        rot6d_one(x6, glob_R + (size_t)t * 9);
    }
    if (shape_std && t < B * nb) { const int b = t / nb, l = t - b * nb; shape_std[t] = expf(heads[b * ld + nb + l]); }
}

}  // namespace

extern "C" int hf_heads_finish(const float* heads, const float* init_glob, const float* init_cam,
                               const float* shape_eps, int B, int N, int nb, float* cam, float* glob6,
                               float* shape_rows, float* glob_R, float* shape_std, void* stream) {
    if (!heads || !init_glob || !init_cam || !cam || !glob6 || !shape_rows) return hf::fail(HF_ERR_INVALID, "hf_heads_finish: null argument");
    if (B <= 0) return HF_OK;
    const int total = B * N * nb + B * nb;
    int blocks = hf::div_up(std::max(total, B * std::max(6, nb)), 256);
    if (blocks > 1024) blocks = 1024;
    if (blocks * 256 < B * 6 || blocks * 256 < B * nb) return hf::fail(HF_ERR_UNSUPPORTED, "hf_heads_finish: batch too large");
    HF_CUDA(hf::launch_pdl(heads_finish_kernel, dim3(blocks), dim3(256), 0, (cudaStream_t)stream, heads, init_glob, init_cam, shape_eps, B, N,
                           nb, cam, glob6, shape_rows, glob_R, shape_std));
    HF_LAUNCH_CHECK();
    return HF_OK;
}

namespace {
// K-slices of a layer: up to one CTA per SM (255 registers x 256 threads: one CTA per SM, stay within one wave), slices a multiple
// of 128 floats
int linear_slices(int M, int K, int O) {
    const bool vec4 = (K % 4 == 0) && (K >= 64);
    if (!vec4) return 1;
    const int ctas = hf::div_up(O, LW * LNB) * hf::div_up(M, 32);
    int nz = 1;
    while (nz < 8 && ctas * nz * 2 <= 148 && (K % (nz * 2 * 128)) == 0) nz *= 2;
    return nz;
}
}  // namespace

extern "C" size_t hf_linear_workspace_bytes(int M, int K, int O) {
    if (M <= 0 || K <= 0 || O <= 0) return 0;
    const int nz = linear_slices(M, K, O);
    return nz > 1 ? (size_t)nz * M * O * sizeof(float) : 0;
}

extern "C" int hf_linear(const float* x, int ldx, const float* W, int ldw, const float* b, float* y, int ldy,
                         int M, int K, int O, int act, int accumulate, void* stream) {
    return hf_linear_ws(x, ldx, W, ldw, b, y, ldy, M, K, O, act, accumulate, nullptr, 0, stream);
}

extern "C" int hf_linear_ws(const float* x, int ldx, const float* W, int ldw, const float* b, float* y, int ldy,
                            int M, int K, int O, int act, int accumulate, void* workspace, size_t workspace_bytes, void* stream) {
    if (!x || !W || !y) return hf::fail(HF_ERR_INVALID, "hf_linear: null argument");
    if (M <= 0 || O <= 0) return HF_OK;
    if (K < 0 || ldx < K || ldw < K || ldy < O) return hf::fail(HF_ERR_INVALID, "hf_linear: bad strides");
    const bool vec4 = (K % 4 == 0) && (K >= 64) && (ldw % 4 == 0) && (ldx % 4 == 0) && (((uintptr_t)W & 15) == 0) && (((uintptr_t)x & 15) == 0);
    if (vec4) {
        const int smem = 2 * 32 * KC * (int)sizeof(float);
        // per device / context attribute: set on every call (cheap), not cached in a process-wide static
        HF_CUDA(cudaFuncSetAttribute(linear_vec_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        // K-sliced across CTAs when the caller provides room for the partial sums (hf_linear_workspace_bytes); otherwise one CTA
        // per output tile walks the whole K
        int nz = workspace ? linear_slices(M, K, O) : 1;
        float* part = nullptr;
        if (nz > 1) {
            if (workspace_bytes < (size_t)nz * M * O * sizeof(float) || ((uintptr_t)workspace & 15))
                return hf::fail(HF_ERR_INVALID, "hf_linear_ws: workspace too small or unaligned (%zu < %zu)", workspace_bytes, (size_t)nz * M * O * sizeof(float));
            part = static_cast<float*>(workspace);
        }
        dim3 grid(hf::div_up(O, LW * LNB), hf::div_up(M, 32), nz);
        HF_CUDA(hf::launch_pdl(linear_vec_kernel, grid, dim3(LW * 32), (size_t)smem, (cudaStream_t)stream, x, ldx, W, ldw, b, y, ldy, M, K, O, act, accumulate,
                               nz > 1 ? K / nz : K, part));
        if (nz > 1) {
            HF_LAUNCH_CHECK();
            HF_CUDA(hf::launch_pdl(linear_reduce_kernel, dim3(hf::div_up(M * O, 256)), dim3(256), 0, (cudaStream_t)stream, (const float*)part, nz, b, y, ldy, M, O,
                                   act, accumulate));
        }
    } else {
        dim3 grid(hf::div_up(O, 4), hf::div_up(M, 32));
        HF_CUDA(hf::launch_pdl(linear_small_kernel, grid, dim3(128), 0, (cudaStream_t)stream, x, ldx, W, ldw, b, y, ldy, M, K, O, act, accumulate));
    }
    HF_LAUNCH_CHECK();
    return HF_OK;
}

extern "C" int hf_rot6d_to_rotmat(const float* rot6d, float* rotmats, int n, void* stream) {
    if (n <= 0) return HF_OK;
    HF_CUDA(hf::launch_pdl(rot6d_kernel, dim3(hf::div_up(n, 128)), dim3(128), 0, (cudaStream_t)stream, rot6d, rotmats, n));
    HF_LAUNCH_CHECK();
    return HF_OK;
}
