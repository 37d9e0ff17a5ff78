// Small dense layers and head post-processing (models/humaniflow_model.py:232-258, 116-150;
// utils/rigid_transform_utils.py:86-100).  M is the image batch (tens of rows): these are latency-bound
// GEMVs, written for coalesced weight streaming rather than tensor cores.
#include "common.cuh"

namespace {

constexpr int OPW = 4;    // outputs per warp
constexpr int MT = 16;    // rows per CTA tile
constexpr int WARPS = 8;

__device__ __forceinline__ float act_fn(float a, int act) {
    if (act == 1) return a > 0.f ? a : expm1f(a);
    if (act == 2) return fmaxf(a, 0.f);
    return a;
}

// y[m][o] (+)= act(sum_k x[m][k] W[o][k] + b[o]); a warp owns OPW outputs x MT rows, lanes stride K.
__global__ void __launch_bounds__(WARPS * 32)
linear_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ W, int ldw,
              const float* __restrict__ b, float* __restrict__ y, int ldy, int M, int K, int O, int act,
              int accumulate) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int o0 = (blockIdx.x * WARPS + warp) * OPW;
    const int m0 = blockIdx.y * MT;
    if (o0 >= O) return;
    float acc[OPW][MT];
#pragma unroll
    for (int i = 0; i < OPW; ++i)
#pragma unroll
        for (int m = 0; m < MT; ++m) acc[i][m] = 0.f;
    for (int k = lane; k < K; k += 32) {
        float w[OPW];
#pragma unroll
        for (int i = 0; i < OPW; ++i) w[i] = (o0 + i < O) ? __ldg(W + (size_t)(o0 + i) * ldw + k) : 0.f;
#pragma unroll
        for (int m = 0; m < MT; ++m) {
            const float xv = (m0 + m < M) ? __ldg(x + (size_t)(m0 + m) * ldx + k) : 0.f;
#pragma unroll
            for (int i = 0; i < OPW; ++i) acc[i][m] = fmaf(w[i], xv, acc[i][m]);
        }
    }
#pragma unroll
    for (int i = 0; i < OPW; ++i)
#pragma unroll
        for (int m = 0; m < MT; ++m) {
            float v = acc[i][m];
#pragma unroll
            for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
            acc[i][m] = v;
        }
    // lane l writes element (i = l / MT ... ) : spread the OPW*MT results over the lanes
    for (int e = lane; e < OPW * MT; e += 32) {
        const int i = e / MT, m = e - i * MT;
        float v = 0.f;
#pragma unroll
        for (int ii = 0; ii < OPW; ++ii)
#pragma unroll
            for (int mm = 0; mm < MT; ++mm)
                if (ii == i && mm == m) v = acc[ii][mm];
        const int o = o0 + i, r = m0 + m;
        if (o < O && r < M) {
            float* dst = y + (size_t)r * ldy + o;
            float a = v + (b ? __ldg(b + o) : 0.f);
            if (accumulate) a += *dst;
            *dst = act_fn(a, act);
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
    dim3 grid(hf::div_up(O, WARPS * OPW), hf::div_up(M, MT));
    linear_kernel<<<grid, WARPS * 32, 0, (cudaStream_t)stream>>>(x, ldx, W, ldw, b, y, ldy, M, K, O, act, accumulate);
    HF_LAUNCH_CHECK();
    return HF_OK;
}

extern "C" int hf_rot6d_to_rotmat(const float* rot6d, float* rotmats, int n, void* stream) {
    if (n <= 0) return HF_OK;
    rot6d_kernel<<<hf::div_up(n, 128), 128, 0, (cudaStream_t)stream>>>(rot6d, rotmats, n);
    HF_LAUNCH_CHECK();
    return HF_OK;
}
