// Small dense layers and head post-processing (models/humaniflow_model.py:232-258, 116-150;
// utils/rigid_transform_utils.py:86-100).  M is the image batch (tens of rows): these are latency-bound
// GEMVs, written for coalesced weight streaming rather than tensor cores.
#include "common.cuh"

namespace {

constexpr int LW = 4;      // warps per CTA
constexpr int LO = 2;      // outputs per warp
constexpr int KC = 256;    // K chunk staged in shared memory

__device__ __forceinline__ float act_fn(float a, int act) {
    if (act == 1) return a > 0.f ? a : expm1f(a);
    if (act == 2) return fmaxf(a, 0.f);
    return a;
}

// y[m][o] (+)= act(sum_k x[m][k] W[o][k] + b[o]).  Lane = batch row (32 rows per CTA tile), a warp owns LO
// output neurons; the x tile is staged transposed in shared memory (conflict-free column reads), weights
// are streamed with warp-broadcast (vector) loads, each element read once per row tile.  No reductions.
template <bool VEC4>
__global__ void __launch_bounds__(LW * 32)
linear_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ W, int ldw,
              const float* __restrict__ b, float* __restrict__ y, int ldy, int M, int K, int O, int act,
              int accumulate) {
    __shared__ float xs[KC][33];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int o0 = (blockIdx.x * LW + warp) * LO;
    const int m0 = blockIdx.y * 32;
    float acc[LO];
#pragma unroll
    for (int i = 0; i < LO; ++i) acc[i] = 0.f;
    const float* wrow[LO];
#pragma unroll
    for (int i = 0; i < LO; ++i) wrow[i] = W + (size_t)min(o0 + i, O - 1) * ldw;
    for (int kc = 0; kc < K; kc += KC) {
        const int kn = min(KC, K - kc);
        __syncthreads();
        for (int idx = threadIdx.x; idx < 32 * kn; idx += LW * 32) {
            const int r = idx / kn, k = idx - r * kn;
            xs[k][r] = (m0 + r < M) ? __ldg(x + (size_t)(m0 + r) * ldx + kc + k) : 0.f;
        }
        __syncthreads();
        if (VEC4) {
#pragma unroll 4
            for (int k = 0; k < kn; k += 4) {
                const float x0 = xs[k][lane], x1 = xs[k + 1][lane], x2 = xs[k + 2][lane], x3 = xs[k + 3][lane];
#pragma unroll
                for (int i = 0; i < LO; ++i) {
                    const float4 w = __ldg(reinterpret_cast<const float4*>(wrow[i] + kc + k));
                    acc[i] = fmaf(w.x, x0, acc[i]); acc[i] = fmaf(w.y, x1, acc[i]);
                    acc[i] = fmaf(w.z, x2, acc[i]); acc[i] = fmaf(w.w, x3, acc[i]);
                }
            }
        } else {
#pragma unroll 4
            for (int k = 0; k < kn; ++k) {
                const float xv = xs[k][lane];
#pragma unroll
                for (int i = 0; i < LO; ++i) acc[i] = fmaf(__ldg(wrow[i] + kc + k), xv, acc[i]);
            }
        }
    }
    const int r = m0 + lane;
    if (r < M) {
#pragma unroll
        for (int i = 0; i < LO; ++i) {
            const int o = o0 + i;
            if (o < O) {
                float* dst = y + (size_t)r * ldy + o;
                float a = acc[i] + (b ? __ldg(b + o) : 0.f);
                if (accumulate) a += *dst;
                *dst = act_fn(a, act);
            }
        }
    }
}

__global__ void rot6d_kernel(const float* __restrict__ x6, float* __restrict__ R, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    // x.view(-1,3,2): a1 = (x0,x2,x4), a2 = (x1,x3,x5); F.normalize eps 1e-12; columns b1,b2,b3
    const float* x = x6 + (size_t)i * 6;
    float a1[3] = {x[0], x[2], x[4]}, a2[3] = {x[1], x[3], x[5]};
    float n1 = fmaxf(sqrtf(a1[0] * a1[0] + a1[1] * a1[1] + a1[2] * a1[2]), 1e-12f);
    float b1[3] = {a1[0] / n1, a1[1] / n1, a1[2] / n1};
    float d = b1[0] * a2[0] + b1[1] * a2[1] + b1[2] * a2[2];
    float u[3] = {a2[0] - d * b1[0], a2[1] - d * b1[1], a2[2] - d * b1[2]};
    float n2 = fmaxf(sqrtf(u[0] * u[0] + u[1] * u[1] + u[2] * u[2]), 1e-12f);
    float b2[3] = {u[0] / n2, u[1] / n2, u[2] / n2};
    float b3[3] = {b1[1] * b2[2] - b1[2] * b2[1], b1[2] * b2[0] - b1[0] * b2[2], b1[0] * b2[1] - b1[1] * b2[0]};
    float* o = R + (size_t)i * 9;
    for (int r = 0; r < 3; ++r) { o[r * 3 + 0] = b1[r]; o[r * 3 + 1] = b2[r]; o[r * 3 + 2] = b3[r]; }
}

// heads (B, 2*nb + 6 + 3) = [shape_mode | shape_log_std | glob6 | cam]  ->  cam + init, rot6d(glob6 + init),
// and the per-row shape table for the flow: rows [0, B*N) = mode + exp(log_std) * eps (or mode when eps == NULL),
// rows [B*N, B*N + B) = mode (point-estimate rows).  models/humaniflow_model.py:237-258.
__global__ void heads_finish_kernel(const float* __restrict__ heads, const float* __restrict__ init_glob,
                                    const float* __restrict__ init_cam, const float* __restrict__ shape_eps,
                                    int B, int N, int nb, float* __restrict__ cam, float* __restrict__ glob6,
                                    float* __restrict__ shape_rows) {
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
}

}  // namespace

extern "C" int hf_heads_finish(const float* heads, const float* init_glob, const float* init_cam,
                               const float* shape_eps, int B, int N, int nb, float* cam, float* glob6,
                               float* shape_rows, void* stream) {
    if (!heads || !init_glob || !init_cam || !cam || !glob6 || !shape_rows) return hf::fail(HF_ERR_INVALID, "hf_heads_finish: null argument");
    if (B <= 0) return HF_OK;
    const int total = B * N * nb + B * nb;
    int blocks = hf::div_up(total > B * 6 ? total : B * 6, 256);
    if (blocks > 1024) blocks = 1024;
    if (blocks * 256 < B * 6) return hf::fail(HF_ERR_UNSUPPORTED, "hf_heads_finish: batch too large");
    heads_finish_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(heads, init_glob, init_cam, shape_eps, B, N, nb, cam,
                                                                 glob6, shape_rows);
    HF_LAUNCH_CHECK();
    return HF_OK;
}

extern "C" int hf_linear(const float* x, int ldx, const float* W, int ldw, const float* b, float* y, int ldy,
                         int M, int K, int O, int act, int accumulate, void* stream) {
    if (!x || !W || !y) return hf::fail(HF_ERR_INVALID, "hf_linear: null argument");
    if (M <= 0 || O <= 0) return HF_OK;
    if (K < 0 || ldx < K || ldw < K || ldy < O) return hf::fail(HF_ERR_INVALID, "hf_linear: bad strides");
    dim3 grid(hf::div_up(O, LW * LO), hf::div_up(M, 32));
    const bool vec4 = (K % 4 == 0) && (ldw % 4 == 0) && (((uintptr_t)W & 15) == 0);
    if (vec4)
        linear_kernel<true><<<grid, LW * 32, 0, (cudaStream_t)stream>>>(x, ldx, W, ldw, b, y, ldy, M, K, O, act, accumulate);
    else
        linear_kernel<false><<<grid, LW * 32, 0, (cudaStream_t)stream>>>(x, ldx, W, ldw, b, y, ldy, M, K, O, act, accumulate);
    HF_LAUNCH_CHECK();
    return HF_OK;
}

extern "C" int hf_rot6d_to_rotmat(const float* rot6d, float* rotmats, int n, void* stream) {
    if (n <= 0) return HF_OK;
    rot6d_kernel<<<hf::div_up(n, 128), 128, 0, (cudaStream_t)stream>>>(rot6d, rotmats, n);
    HF_LAUNCH_CHECK();
    return HF_OK;
}
