// Input proxy representation on the device, one fused kernel:
//   channel 0      Canny-style edge map of the RGB crop        models/canny_edge_detector.py:104-166
//                  (5-tap separable Gaussian blur -> Sobel gradients averaged over the channels [computed on the channel sum: the
//                  stencils are linear] -> magnitude,
//                  orientation binned to 45 degrees -> non-maximum suppression along the gradient -> threshold)
//   channels 1..J  Gaussian heatmaps of the 2-D joints          utils/label_conversions.py:106-125, times the visibility flags
//                  (predict_humaniflow.py:103-110)
// The reference runs ~15 single-channel conv / elementwise launches with a Python loop over channels; here one CTA
// produces a 32 x 16 pixel tile of all 1+J channels from an RGB halo tile held in shared memory.  Every stage is defined on
// the image domain and zero outside it, exactly like the zero padding of the reference's chained nn.Conv2d modules.
#include "common.cuh"
#include <cuda_bf16.h>

namespace {

constexpr int PT_W = 32, PT_H = 16, PR_THREADS = 256;
constexpr int RW = PT_W + 8, RH = PT_H + 8;        // RGB tile with halo 4
constexpr int HW_ = PT_W + 4, HH = PT_H + 8;       // horizontally blurred: cols halo 2, rows halo 4
constexpr int BW = PT_W + 4, BH = PT_H + 4;        // blurred: halo 2
constexpr int MW = PT_W + 2, MH = PT_H + 2;        // gradient magnitude: halo 1

struct ProxyParams {
    float g[5];          // normalised Gaussian taps
    float threshold;
    int nms;
    float heat_std;
};

__global__ void __launch_bounds__(PR_THREADS)
proxy_rep_kernel(const float* __restrict__ rgb, const float* __restrict__ joints2D, const float* __restrict__ vis, int C, int H,
                 int W, int J, ProxyParams prm, float* __restrict__ out, float* __restrict__ dbg_mag, float* __restrict__ dbg_ori,
                 __nv_bfloat16* __restrict__ staged, int Hp, int Wp, int Cp, int top, int left) {
    HF_PDL_SYNC();
    __shared__ float s_rgb[RH][RW];
    __shared__ float s_hb[HH][HW_];
    __shared__ float s_bl[BH][BW];
    __shared__ float s_gx[MH][MW], s_gy[MH][MW];
    __shared__ float s_mag[MH][MW];
    __shared__ float s_edge[PT_H][PT_W];
    // heatmaps are separable: exp(-(a*a)/2 - (c*c)/2) = exp(-(a*a)/2) exp(-(c*c)/2) with a = (row - v)/std, c = (col - u)/std.  The
    // two factors are computed once per (joint, tile row) and (joint, tile column); a pixel then costs ONE multiplication per joint
    // instead of two divisions and an expf (the product differs from the single exponential by <= 2 ulp: tolerance 2e-6).
    __shared__ float s_hrow[32][PT_H], s_hcol[32][PT_W], s_vis[32];
    const int b = blockIdx.z, x0 = blockIdx.x * PT_W, y0 = blockIdx.y * PT_H;
    const int tid = threadIdx.x;
    for (int i = tid; i < J * (PT_H + PT_W); i += PR_THREADS) {
        const int j = i / (PT_H + PT_W), k = i - j * (PT_H + PT_W);
        if (k < PT_H) {
            const float a = ((float)(y0 + k) - __ldg(joints2D + ((size_t)b * J + j) * 2 + 1)) / prm.heat_std;
            s_hrow[j][k] = expf(-(a * a) / 2.f);
        } else {
            const float c2 = ((float)(x0 + k - PT_H) - __ldg(joints2D + ((size_t)b * J + j) * 2)) / prm.heat_std;
            s_hcol[j][k - PT_H] = expf(-(c2 * c2) / 2.f);
        }
    }
    if (tid < J) s_vis[tid] = vis ? __ldg(vis + (size_t)b * J + tid) : 1.f;
    // Blur and Sobel are linear, so the gradients summed over the C channels (what the reference averages) are the gradients of the
    // channel SUM: one pass through the stencils instead of C (same map in real arithmetic; in fp32 the summation order differs,
    // which moves a gradient by ~1e-7 and can flip the threshold / non-maximum decision of an isolated pixel -- the tests bound the
    // flipped fraction).
    {
        __syncthreads();
        for (int i = tid; i < RH * RW; i += PR_THREADS) {
            const int r = i / RW, q = i - r * RW, y = y0 - 4 + r, x = x0 - 4 + q;
            float a = 0.f;
            if (y >= 0 && y < H && x >= 0 && x < W)
                for (int c = 0; c < C; ++c) a += __ldg(rgb + ((size_t)b * C + c) * H * W + (size_t)y * W + x);
            s_rgb[r][q] = a;
        }
        __syncthreads();
        for (int i = tid; i < HH * HW_; i += PR_THREADS) {          // horizontal 1x5, image col x0 - 2 + q
            const int r = i / HW_, q = i - r * HW_, y = y0 - 4 + r, x = x0 - 2 + q;
            float a = 0.f;
            if (y >= 0 && y < H && x >= 0 && x < W) {
#pragma unroll
                for (int k = 0; k < 5; ++k) a = fmaf(prm.g[k], s_rgb[r][q + k], a);
            }
            s_hb[r][q] = a;
        }
        __syncthreads();
        for (int i = tid; i < BH * BW; i += PR_THREADS) {           // vertical 5x1, image row y0 - 2 + r
            const int r = i / BW, q = i - r * BW, y = y0 - 2 + r, x = x0 - 2 + q;
            float a = 0.f;
            if (y >= 0 && y < H && x >= 0 && x < W) {
#pragma unroll
                for (int k = 0; k < 5; ++k) a = fmaf(prm.g[k], s_hb[r + k][q], a);
            }
            s_bl[r][q] = a;
        }
        __syncthreads();
        for (int i = tid; i < MH * MW; i += PR_THREADS) {           // Sobel (cross-correlation), image pixel (y0 - 1 + r, x0 - 1 + q)
            const int r = i / MW, q = i - r * MW;
            const float a00 = s_bl[r][q], a01 = s_bl[r][q + 1], a02 = s_bl[r][q + 2];
            const float a10 = s_bl[r + 1][q], a12 = s_bl[r + 1][q + 2];
            const float a20 = s_bl[r + 2][q], a21 = s_bl[r + 2][q + 1], a22 = s_bl[r + 2][q + 2];
            s_gx[r][q] = (a00 - a02) + 2.f * (a10 - a12) + (a20 - a22);
            s_gy[r][q] = (a00 - a20) + 2.f * (a01 - a21) + (a02 - a22);
        }
    }
    __syncthreads();
    const float nch = (float)C;
    for (int i = tid; i < MH * MW; i += PR_THREADS) {
        const int r = i / MW, q = i - r * MW, y = y0 - 1 + r, x = x0 - 1 + q;
        float m = 0.f;
        if (y >= 0 && y < H && x >= 0 && x < W) {
            const float gx = s_gx[r][q] / nch, gy = s_gy[r][q] / nch;
            m = sqrtf(gx * gx + gy * gy);
            s_gx[r][q] = gx; s_gy[r][q] = gy;
        }
        s_mag[r][q] = m;
    }
    __syncthreads();
    const int HWp = H * W;
    for (int i = tid; i < PT_H * PT_W; i += PR_THREADS) {
        const int r = i / PT_W, q = i - r * PT_W, y = y0 + r, x = x0 + q;
        if (y >= H || x >= W) continue;
        const float m = s_mag[r + 1][q + 1];
        // orientation in degrees (0, 360], binned to multiples of 45 (torch.round: half to even)
        const float ori = rintf((atan2f(s_gy[r + 1][q + 1], s_gx[r + 1][q + 1]) * 57.29577951308232f + 180.0f) / 45.0f) * 45.0f;
        float e = m;
        if (prm.nms) {
            const int bin = ((int)(ori / 45.0f)) % 8;
            const int p = bin & 3;                       // pair of opposite directions (p, p + 4)
            // directional differences centre - neighbour: 0: (y, x+1)  1: (y+1, x+1)  2: (y+1, x)  3: (y+1, x-1), and the opposite ones
            const int dy = (p == 0) ? 0 : 1, dx = (p == 0) ? 1 : ((p == 1) ? 1 : ((p == 2) ? 0 : -1));
            const float dpos = m - s_mag[r + 1 + dy][q + 1 + dx], dneg = m - s_mag[r + 1 - dy][q + 1 - dx];
            if (!(fminf(dpos, dneg) > 0.f)) e = 0.f;
        }
        if (e < prm.threshold) e = 0.f;
        const size_t pix = (size_t)y * W + x;
        if (staged) { s_edge[r][q] = e; continue; }
        out[(size_t)b * (1 + J) * HWp + pix] = e;
        if (dbg_mag) dbg_mag[(size_t)b * HWp + pix] = m;
        if (dbg_ori) dbg_ori[(size_t)b * HWp + pix] = ori;
        // heatmaps: exp(-((row - v) / std)^2 / 2 - ((col - u) / std)^2 / 2), joints2D = (u, v) = (column, row)
        for (int j = 0; j < J; ++j) {
            float h = s_hrow[j][r] * s_hcol[j][q];
            if (vis) h *= s_vis[j];
            out[((size_t)b * (1 + J) + 1 + j) * HWp + pix] = h;
        }
    }
    if (staged) {
        // the encoder's stem input layout directly: bf16 NHWC, Cp channels per pixel (edge map, J heatmaps, zero padding) inside the
        // zero border of the (Hp, Wp) padded image.  One 16-byte store per (pixel, group of 8 channels); consecutive threads write
        // consecutive 16-byte pieces, so a warp writes 512 contiguous bytes.
        __syncthreads();
        // a thread owns a pixel and writes its Cp channels as Cp/8 16-byte pieces (the lanes of a warp cover 32 consecutive pixels
        // = one contiguous 32*Cp*2-byte run across the pieces)
        const int G = Cp >> 3;
        for (int i = tid; i < PT_H * PT_W; i += PR_THREADS) {
            const int r = i / PT_W, q = i - r * PT_W, y = y0 + r, x = x0 + q;
            if (y >= H || x >= W) continue;
            __nv_bfloat16* dst = staged + (((size_t)b * Hp + y + top) * Wp + x + left) * Cp;
            for (int grp = 0; grp < G; ++grp) {
                float ch[8];
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const int cc = grp * 8 + c;            // warp-uniform
                    float val = 0.f;
                    if (cc == 0) val = s_edge[r][q];
                    else if (cc <= J) {
                        val = s_hrow[cc - 1][r] * s_hcol[cc - 1][q];
                        if (vis) val *= s_vis[cc - 1];
                    }
                    ch[c] = val;
                }
                uint32_t pk[4];
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    __nv_bfloat162 h2 = __floats2bfloat162_rn(ch[2 * c], ch[2 * c + 1]);
                    pk[c] = *reinterpret_cast<uint32_t*>(&h2);
                }
                *reinterpret_cast<uint4*>(dst + grp * 8) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            }
        }
    }
}

}  // namespace

extern "C" int hf_proxy_rep(const float* rgb, const float* joints2D, const float* joints_vis, int B, int C, int H, int W, int J,
                            const float* gauss5, float threshold, int nms, float heat_std, float* out, float* dbg_mag,
                            float* dbg_ori, void* stream) {
    if (!rgb || !out || !gauss5 || (J > 0 && !joints2D)) return hf::fail(HF_ERR_INVALID, "hf_proxy_rep: null argument");
    if (B <= 0 || H <= 0 || W <= 0) return HF_OK;
    if (C < 1 || heat_std <= 0.f) return hf::fail(HF_ERR_INVALID, "hf_proxy_rep: bad channel count / heatmap std");
    if (J > 32) return hf::fail(HF_ERR_UNSUPPORTED, "hf_proxy_rep: at most 32 joints (got %d)", J);
    ProxyParams prm;
    for (int i = 0; i < 5; ++i) prm.g[i] = gauss5[i];
    prm.threshold = threshold; prm.nms = nms; prm.heat_std = heat_std;
    dim3 grid(hf::div_up(W, PT_W), hf::div_up(H, PT_H), B);
    HF_CUDA(hf::launch_pdl(proxy_rep_kernel, grid, dim3(PR_THREADS), 0, (cudaStream_t)stream, rgb, joints2D, joints_vis, C, H, W, J, prm,
                           out, dbg_mag, dbg_ori, (__nv_bfloat16*)nullptr, 0, 0, 0, 0, 0));
    HF_LAUNCH_CHECK();
    return HF_OK;
}

extern "C" int hf_proxy_rep_staged(const float* rgb, const float* joints2D, const float* joints_vis, int B, int C, int H, int W, int J,
                                   const float* gauss5, float threshold, int nms, float heat_std, uint16_t* staged, int Hp, int Wp,
                                   int Cp, int top, int left, void* stream) {
    if (!rgb || !staged || !gauss5 || (J > 0 && !joints2D)) return hf::fail(HF_ERR_INVALID, "hf_proxy_rep_staged: null argument");
    if (B <= 0 || H <= 0 || W <= 0) return HF_OK;
    if (C < 1 || heat_std <= 0.f) return hf::fail(HF_ERR_INVALID, "hf_proxy_rep_staged: bad channel count / heatmap std");
    if (1 + J > Cp || Cp > 32 || Cp % 8 || H + top > Hp || W + left > Wp)
        return hf::fail(HF_ERR_INVALID, "hf_proxy_rep_staged: %d channels / %dx%d image do not fit the staged layout (%d, %d, %d)", 1 + J, H, W, Hp, Wp, Cp);
    ProxyParams prm;
    for (int i = 0; i < 5; ++i) prm.g[i] = gauss5[i];
    prm.threshold = threshold; prm.nms = nms; prm.heat_std = heat_std;
    dim3 grid(hf::div_up(W, PT_W), hf::div_up(H, PT_H), B);
    HF_CUDA(hf::launch_pdl(proxy_rep_kernel, grid, dim3(PR_THREADS), 0, (cudaStream_t)stream, rgb, joints2D, joints_vis, C, H, W, J, prm,
                           (float*)nullptr, (float*)nullptr, (float*)nullptr, (__nv_bfloat16*)staged, Hp, Wp, Cp, top, left));
    HF_LAUNCH_CHECK();
    return HF_OK;
}
