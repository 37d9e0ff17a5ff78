// Per-sample 3-D error metrics of the evaluation path on the device.
//
// For every predicted point set (a sampled mesh: P = 6890 vertices, or a joint set: P = 14 / 17) and its image's target:
//   plain  mean_i || p_i - t_i ||                                   metrics/eval_metrics_tracker.py:119-125, 169-174, 201-206
//   SC     after scale_and_translation_transform_batch(p, t)        utils/eval_utils.py:105-125  (tracker :128-136, 209-217)
//   PA     after procrustes_analysis_batch(p, t)                    utils/eval_utils.py:62-102   (tracker :139-147, 220-229)
// These per-sample means are what every PVE / PVE-SC / PVE-PA / PVE-T(-SC) / MPJPE(-SC/-PA) variant and their
// "samples_min" forms reduce (sum = mean * P, minimum over the samples of an image).  The reference copies all meshes to
// the host (2 x 265 MB per batch) and runs numpy incl. one 3x3 SVD per mesh; here a block per sample reads the mesh twice
// from HBM and nothing but 3 floats per sample leaves the GPU.
#include "common.cuh"
#include <algorithm>

namespace {

constexpr int ME_THREADS = 256;     // block per point set; eight such blocks per SM overlap the single-thread 3x3 solve of one set with the passes of the others
constexpr int SS_THREADS = 1024;    // sample_stats

// Jacobi eigen-decomposition of a symmetric 3x3 matrix (fp64): A = V diag(w) V^T, columns of V orthonormal.
__device__ void eig_sym3(double A[3][3], double V[3][3], double w[3]) {
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) V[i][j] = i == j ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 12; ++sweep) {
        const double off = fabs(A[0][1]) + fabs(A[0][2]) + fabs(A[1][2]);
        if (off <= 1e-18 * (fabs(A[0][0]) + fabs(A[1][1]) + fabs(A[2][2])) || off < 1e-300) break;     // converged (Jacobi: quadratic, ~5 sweeps)
        for (int p = 0; p < 2; ++p)
            for (int q = p + 1; q < 3; ++q) {
                if (fabs(A[p][q]) < 1e-300) continue;
                const double theta = (A[q][q] - A[p][p]) / (2.0 * A[p][q]);
                const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < 3; ++k) {          // A <- A J
                    const double akp = A[k][p], akq = A[k][q];
                    A[k][p] = c * akp - s * akq;
                    A[k][q] = s * akp + c * akq;
                }
                for (int k = 0; k < 3; ++k) {          // A <- J^T A
                    const double apk = A[p][k], aqk = A[q][k];
                    A[p][k] = c * apk - s * aqk;
                    A[q][k] = s * apk + c * aqk;
                }
                for (int k = 0; k < 3; ++k) {          // V <- V J
                    const double vkp = V[k][p], vkq = V[k][q];
                    V[k][p] = c * vkp - s * vkq;
                    V[k][q] = s * vkp + c * vkq;
                }
            }
    }
    for (int i = 0; i < 3; ++i) w[i] = A[i][i];
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
    return v;
}

// block-wide sums of NV doubles per thread -> every thread gets the totals (through shared memory)
template <int NV, int NTHR = ME_THREADS>
__device__ void block_sum(double* v, double* scratch /* [NTHR/32][NV] + [NV] */) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const double s = warp_sum(v[i]);
        if (lane == 0) scratch[warp * NV + i] = s;
    }
    __syncthreads();
    if (threadIdx.x < NV) {
        double s = 0.0;
        for (int w = 0; w < NTHR / 32; ++w) s += scratch[w * NV + threadIdx.x];
        scratch[(NTHR / 32) * NV + threadIdx.x] = s;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = scratch[(NTHR / 32) * NV + i];
    __syncthreads();
}

// pred (B*N, P, 3), target (B, P, 3) -> out (B*N, 3) = [plain, SC, PA] mean point errors
// STAGE: the predicted point set (P x 3 floats, 83 KB for a mesh) is copied into shared memory once (8-byte cp.async, all in
// flight at once) and both passes read it from there: HBM is read once per set, the second pass never misses L2
template <bool STAGE>
__global__ void __launch_bounds__(ME_THREADS, STAGE ? 2 : 4)
pointset_errors_kernel(const float* __restrict__ pred, const float* __restrict__ target, int N, int P, float* __restrict__ out) {
    HF_PDL_SYNC();
    extern __shared__ __align__(16) float pstage[];
    __shared__ double scratch[(ME_THREADS / 32 + 1) * 18];
    __shared__ float xf[16];                                   // SC: s, mu1, mu2;  PA: scale*R (9), t (3)
    const int m = blockIdx.x, b = m / N;
    const float* p = pred + (size_t)m * P * 3;
    const float* t = target + (size_t)b * P * 3;
    if (STAGE) {
        const int n2 = (P * 3) >> 1;                           // 8-byte pieces (the set starts 8-byte aligned: checked on the host)
        for (int i = threadIdx.x; i < n2; i += ME_THREADS)
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(pstage + 2 * i)), "l"(p + 2 * i) : "memory");
        if ((P * 3) & 1) { if (threadIdx.x == 0) pstage[P * 3 - 1] = __ldg(p + P * 3 - 1); }
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
        p = pstage;
    }
#define LDP(ptr) (STAGE ? *(ptr) : __ldg(ptr))
    // pass 1: raw moments + the plain error
    const float p0x = LDP(p), p0y = LDP(p + 1), p0z = LDP(p + 2), t0x = __ldg(t), t0y = __ldg(t + 1), t0z = __ldg(t + 2);
    double v[18];
#pragma unroll
    for (int i = 0; i < 18; ++i) v[i] = 0.0;
    {
        float sp[3] = {0.f, 0.f, 0.f}, st[3] = {0.f, 0.f, 0.f}, spp = 0.f, stt = 0.f, k[9], e = 0.f;
#pragma unroll
        for (int i = 0; i < 9; ++i) k[i] = 0.f;
        // moments are taken about the first point of each set (shift invariance of the centred quantities): no
        // cancellation when the sets sit far from the origin (camera-space meshes)
        auto acc1 = [&](float qx, float qy, float qz, float ux, float uy, float uz) {
            const float dx = qx - ux, dy = qy - uy, dz = qz - uz;
            const float px = qx - p0x, py = qy - p0y, pz = qz - p0z;
            const float tx = ux - t0x, ty = uy - t0y, tz = uz - t0z;
            sp[0] += px; sp[1] += py; sp[2] += pz;
            st[0] += tx; st[1] += ty; st[2] += tz;
            spp += px * px + py * py + pz * pz;
            stt += tx * tx + ty * ty + tz * tz;
            k[0] += px * tx; k[1] += px * ty; k[2] += px * tz;
            k[3] += py * tx; k[4] += py * ty; k[5] += py * tz;
            k[6] += pz * tx; k[7] += pz * ty; k[8] += pz * tz;
            e += sqrtf(dx * dx + dy * dy + dz * dz);
        };
        int i0 = 0;
        if (STAGE) {
            // four points (48 contiguous bytes) per thread and iteration: three 16-byte shared-memory loads of the staged set,
            // six 8-byte loads of the target; with the loop unrolled every target load of a thread is in flight at once
            const int ng = P >> 2;
#pragma unroll 2
            for (int gq = threadIdx.x; gq < ng; gq += ME_THREADS) {
                const float4* ps = reinterpret_cast<const float4*>(p + gq * 12);
                const float2* tg = reinterpret_cast<const float2*>(t + gq * 12);
                const float4 a = ps[0], b4 = ps[1], c = ps[2];
                const float2 u0 = __ldg(tg), u1 = __ldg(tg + 1), u2 = __ldg(tg + 2), u3 = __ldg(tg + 3), u4 = __ldg(tg + 4), u5 = __ldg(tg + 5);
                acc1(a.x, a.y, a.z, u0.x, u0.y, u1.x);
                acc1(a.w, b4.x, b4.y, u1.y, u2.x, u2.y);
                acc1(b4.z, b4.w, c.x, u3.x, u3.y, u4.x);
                acc1(c.y, c.z, c.w, u4.y, u5.x, u5.y);
            }
            i0 = ng << 2;
        }
#pragma unroll 4
        for (int i = i0 + threadIdx.x; i < P; i += ME_THREADS)
            acc1(LDP(p + i * 3), LDP(p + i * 3 + 1), LDP(p + i * 3 + 2), __ldg(t + i * 3), __ldg(t + i * 3 + 1), __ldg(t + i * 3 + 2));
        for (int i = 0; i < 3; ++i) { v[i] = sp[i]; v[3 + i] = st[i]; }
        v[6] = spp; v[7] = stt;
        for (int i = 0; i < 9; ++i) v[8 + i] = k[i];
        v[17] = e;
    }
    block_sum<18>(v, scratch);
    if (threadIdx.x == 0) {
        const double n = (double)P;
        double mu1[3], mu2[3];
        for (int i = 0; i < 3; ++i) { mu1[i] = v[i] / n; mu2[i] = v[3 + i] / n; }        // means of the SHIFTED sets
        const double var1 = v[6] - n * (mu1[0] * mu1[0] + mu1[1] * mu1[1] + mu1[2] * mu1[2]);      // sum |p - mu1|^2
        const double var2 = v[7] - n * (mu2[0] * mu2[0] + mu2[1] * mu2[1] + mu2[2] * mu2[2]);
        const double o1[3] = {(double)p0x, (double)p0y, (double)p0z}, o2[3] = {(double)t0x, (double)t0y, (double)t0z};
        // SC (eval_utils.py:105-125): (p - mu1) / sqrt(var1 / n) * sqrt(var2 / n) + mu2   (absolute means = shifted + origin)
        xf[0] = (float)sqrt(var2 / var1);
        for (int i = 0; i < 3; ++i) { xf[1 + i] = (float)(mu1[i] + o1[i]); xf[4 + i] = (float)(mu2[i] + o2[i]); }
        // PA (eval_utils.py:62-102): K = X1 X2^T; R = V Z U^T for K = U S V^T = V diag(1/s1, 1/s2, d/s3) V^T K^T, d = sign det K
        double K[3][3], KtK[3][3], V[3][3], w[3];
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) K[i][j] = v[8 + i * 3 + j] - n * mu1[i] * mu2[j];
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) KtK[i][j] = K[0][i] * K[0][j] + K[1][i] * K[1][j] + K[2][i] * K[2][j];
#ifndef PSE_NOSOLVE
        eig_sym3(KtK, V, w);
#else
        for (int i = 0; i < 3; ++i) { w[i] = KtK[i][i]; for (int j = 0; j < 3; ++j) V[i][j] = i == j; }
#endif
        int order[3] = {0, 1, 2};                                  // singular values in descending order, as numpy returns them
        for (int a = 0; a < 2; ++a)
            for (int c = a + 1; c < 3; ++c)
                if (w[order[c]] > w[order[a]]) { const int tmp = order[a]; order[a] = order[c]; order[c] = tmp; }
        const double detK = K[0][0] * (K[1][1] * K[2][2] - K[1][2] * K[2][1]) - K[0][1] * (K[1][0] * K[2][2] - K[1][2] * K[2][0]) +
                            K[0][2] * (K[1][0] * K[2][1] - K[1][1] * K[2][0]);
        const double d = detK < 0.0 ? -1.0 : 1.0;
        double sv[3], g[3];
        for (int a = 0; a < 3; ++a) sv[a] = sqrt(fmax(w[order[a]], 0.0));
        const double tiny = 1e-30;
        g[0] = 1.0 / fmax(sv[0], tiny); g[1] = 1.0 / fmax(sv[1], tiny); g[2] = d / fmax(sv[2], tiny);
        double M[3][3];                                            // V diag(g) V^T
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) {
                double s = 0.0;
                for (int a = 0; a < 3; ++a) s += V[i][order[a]] * g[a] * V[j][order[a]];
                M[i][j] = s;
            }
        double R[3][3];                                            // M K^T
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) R[i][j] = M[i][0] * K[j][0] + M[i][1] * K[j][1] + M[i][2] * K[j][2];
        const double scale = (sv[0] + sv[1] + d * sv[2]) / var1;
        for (int i = 0; i < 3; ++i) {
            for (int j = 0; j < 3; ++j) xf[7 + i * 3 + j] = (float)(scale * R[i][j]);
        }
        // translation folded with the means in fp64: t = mu2 - scale R mu1
        double tt[3];
        for (int i = 0; i < 3; ++i)
            tt[i] = (mu2[i] + o2[i]) - scale * (R[i][0] * (mu1[0] + o1[0]) + R[i][1] * (mu1[1] + o1[1]) + R[i][2] * (mu1[2] + o1[2]));
        scratch[0] = tt[0]; scratch[1] = tt[1]; scratch[2] = tt[2];
        scratch[3] = v[17] / n;
    }
    __syncthreads();
    const float s_sc = xf[0];
    const float m1x = xf[1], m1y = xf[2], m1z = xf[3], m2x = xf[4], m2y = xf[5], m2z = xf[6];
    const float r0 = xf[7], r1 = xf[8], r2 = xf[9], r3 = xf[10], r4 = xf[11], r5 = xf[12], r6 = xf[13], r7 = xf[14], r8 = xf[15];
    const float t0 = (float)scratch[0], t1 = (float)scratch[1], t2 = (float)scratch[2];
    const double plain = scratch[3];
    __syncthreads();
    // pass 2: errors after the two alignments
    double e2[2];
    {
        float esc = 0.f, epa = 0.f;
        auto acc2 = [&](float px, float py, float pz, float tx, float ty, float tz) {
            float dx = (px - m1x) * s_sc + m2x - tx, dy = (py - m1y) * s_sc + m2y - ty, dz = (pz - m1z) * s_sc + m2z - tz;
            esc += sqrtf(dx * dx + dy * dy + dz * dz);
            dx = r0 * px + r1 * py + r2 * pz + t0 - tx;
            dy = r3 * px + r4 * py + r5 * pz + t1 - ty;
            dz = r6 * px + r7 * py + r8 * pz + t2 - tz;
            epa += sqrtf(dx * dx + dy * dy + dz * dz);
        };
        int i0 = 0;
        if (STAGE) {
            const int ng = P >> 2;
#pragma unroll 2
            for (int gq = threadIdx.x; gq < ng; gq += ME_THREADS) {
                const float4* ps = reinterpret_cast<const float4*>(p + gq * 12);
                const float2* tg = reinterpret_cast<const float2*>(t + gq * 12);
                const float4 a = ps[0], b4 = ps[1], c = ps[2];
                const float2 u0 = __ldg(tg), u1 = __ldg(tg + 1), u2 = __ldg(tg + 2), u3 = __ldg(tg + 3), u4 = __ldg(tg + 4), u5 = __ldg(tg + 5);
                acc2(a.x, a.y, a.z, u0.x, u0.y, u1.x);
                acc2(a.w, b4.x, b4.y, u1.y, u2.x, u2.y);
                acc2(b4.z, b4.w, c.x, u3.x, u3.y, u4.x);
                acc2(c.y, c.z, c.w, u4.y, u5.x, u5.y);
            }
            i0 = ng << 2;
        }
#pragma unroll 4
        for (int i = i0 + threadIdx.x; i < P; i += ME_THREADS)
            acc2(LDP(p + i * 3), LDP(p + i * 3 + 1), LDP(p + i * 3 + 2), __ldg(t + i * 3), __ldg(t + i * 3 + 1), __ldg(t + i * 3 + 2));
        e2[0] = esc; e2[1] = epa;
    }
    block_sum<2>(e2, scratch);
    if (threadIdx.x == 0) {
        out[(size_t)m * 3 + 0] = (float)plain;
        out[(size_t)m * 3 + 1] = (float)(e2[0] / (double)P);
        out[(size_t)m * 3 + 2] = (float)(e2[1] / (double)P);
    }
#undef LDP
}

// ---- the same computation as three launches, for many point sets --------------------------------------------------------
// In the fused kernel 255 threads of every block wait while thread 0 solves the 3x3 problem (ncu: ~40 % of the warp samples sit at
// that barrier, and shared-memory staging limits an SM to two sets in flight).  With thousands of sets the passes run as their own
// memory-bound kernels (four blocks per SM, nothing staged) around a solve kernel with one THREAD per set:
//   pse_pass_kernel<1>  moments about the first point of each set -> mom[m][18] (fp64)
//   pse_solve_kernel    SC / PA transforms -> xfg[m][20] = [s, mu1 (3), mu2 (3), scale*R (9), t (3), plain error]
//   pse_pass_kernel<2>  the two aligned errors; pass 1 walks the sets downwards, pass 2 upwards, so that each starts on what its
//                       predecessor left in L2
struct Pts4 { float x[12]; };
__device__ __forceinline__ Pts4 load4(const float* q) {      // 4 points = 48 contiguous bytes, 8-byte aligned
    Pts4 r;
    if ((reinterpret_cast<uintptr_t>(q) & 15) == 0) {         // (uniform per block: a set's base is either 16- or only 8-byte aligned)
        const float4* g = reinterpret_cast<const float4*>(q);
#pragma unroll
        for (int i = 0; i < 3; ++i) { const float4 u = __ldg(g + i); r.x[4 * i] = u.x; r.x[4 * i + 1] = u.y; r.x[4 * i + 2] = u.z; r.x[4 * i + 3] = u.w; }
    } else {
        const float2* g = reinterpret_cast<const float2*>(q);
#pragma unroll
        for (int i = 0; i < 6; ++i) { const float2 u = __ldg(g + i); r.x[2 * i] = u.x; r.x[2 * i + 1] = u.y; }
    }
    return r;
}

__device__ void pse_solve(const double* v, int P, const float* p0, const float* t0, float* xf /* [20] */) {
    const double n = (double)P;
    double mu1[3], mu2[3];
    for (int i = 0; i < 3; ++i) { mu1[i] = v[i] / n; mu2[i] = v[3 + i] / n; }
    const double var1 = v[6] - n * (mu1[0] * mu1[0] + mu1[1] * mu1[1] + mu1[2] * mu1[2]);
    const double var2 = v[7] - n * (mu2[0] * mu2[0] + mu2[1] * mu2[1] + mu2[2] * mu2[2]);
    const double o1[3] = {(double)p0[0], (double)p0[1], (double)p0[2]}, o2[3] = {(double)t0[0], (double)t0[1], (double)t0[2]};
    xf[0] = (float)sqrt(var2 / var1);
    for (int i = 0; i < 3; ++i) { xf[1 + i] = (float)(mu1[i] + o1[i]); xf[4 + i] = (float)(mu2[i] + o2[i]); }
    double K[3][3], KtK[3][3], V[3][3], w[3];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) K[i][j] = v[8 + i * 3 + j] - n * mu1[i] * mu2[j];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) KtK[i][j] = K[0][i] * K[0][j] + K[1][i] * K[1][j] + K[2][i] * K[2][j];
    eig_sym3(KtK, V, w);
    int order[3] = {0, 1, 2};
    for (int a = 0; a < 2; ++a)
        for (int c = a + 1; c < 3; ++c)
            if (w[order[c]] > w[order[a]]) { const int tmp = order[a]; order[a] = order[c]; order[c] = tmp; }
    const double detK = K[0][0] * (K[1][1] * K[2][2] - K[1][2] * K[2][1]) - K[0][1] * (K[1][0] * K[2][2] - K[1][2] * K[2][0]) +
                        K[0][2] * (K[1][0] * K[2][1] - K[1][1] * K[2][0]);
    const double d = detK < 0.0 ? -1.0 : 1.0;
    double sv[3], g[3];
    for (int a = 0; a < 3; ++a) sv[a] = sqrt(fmax(w[order[a]], 0.0));
    const double tiny = 1e-30;
    g[0] = 1.0 / fmax(sv[0], tiny); g[1] = 1.0 / fmax(sv[1], tiny); g[2] = d / fmax(sv[2], tiny);
    double Mm[3][3];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            double sacc = 0.0;
            for (int a = 0; a < 3; ++a) sacc += V[i][order[a]] * g[a] * V[j][order[a]];
            Mm[i][j] = sacc;
        }
    double R[3][3];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) R[i][j] = Mm[i][0] * K[j][0] + Mm[i][1] * K[j][1] + Mm[i][2] * K[j][2];
    const double scale = (sv[0] + sv[1] + d * sv[2]) / var1;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) xf[7 + i * 3 + j] = (float)(scale * R[i][j]);
    for (int i = 0; i < 3; ++i)
        xf[16 + i] = (float)((mu2[i] + o2[i]) - scale * (R[i][0] * (mu1[0] + o1[0]) + R[i][1] * (mu1[1] + o1[1]) + R[i][2] * (mu1[2] + o1[2])));
    xf[19] = (float)(v[17] / n);
}

__global__ void __launch_bounds__(128)
pse_solve_kernel(const float* __restrict__ pred, const float* __restrict__ target, const double* __restrict__ mom, int N, int P, int M,
                 float* __restrict__ xfg) {
    HF_PDL_SYNC();
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M) return;
    double v[18];
#pragma unroll
    for (int i = 0; i < 18; ++i) v[i] = mom[(size_t)m * 18 + i];
    const float* p = pred + (size_t)m * P * 3;
    const float* t = target + (size_t)(m / N) * P * 3;
    const float p0[3] = {__ldg(p), __ldg(p + 1), __ldg(p + 2)}, t0[3] = {__ldg(t), __ldg(t + 1), __ldg(t + 2)};
    float xf[20];
    pse_solve(v, P, p0, t0, xf);
#pragma unroll
    for (int i = 0; i < 20; ++i) xfg[(size_t)m * 20 + i] = xf[i];
}

template <int PASS>
__global__ void __launch_bounds__(ME_THREADS, 4)
pse_pass_kernel(const float* __restrict__ pred, const float* __restrict__ target, int N, int P, double* __restrict__ mom,
                const float* __restrict__ xfg, float* __restrict__ out) {
    HF_PDL_SYNC();
    __shared__ double scratch[(ME_THREADS / 32 + 1) * 18];
    // pass 1 walks the sets from the LAST one down (the producer -- the skinning kernel -- wrote them in ascending order, so the tail is
    // what L2 still holds), pass 2 from the first one up (what pass 1 read last)
    const int m = PASS == 1 ? gridDim.x - 1 - blockIdx.x : blockIdx.x, b = m / N;
    const float* p = pred + (size_t)m * P * 3;
    const float* t = target + (size_t)b * P * 3;
    const int ng = P >> 2;
    if (PASS == 1) {
        const float p0x = __ldg(p), p0y = __ldg(p + 1), p0z = __ldg(p + 2), t0x = __ldg(t), t0y = __ldg(t + 1), t0z = __ldg(t + 2);
        float sp[3] = {0.f, 0.f, 0.f}, st[3] = {0.f, 0.f, 0.f}, spp = 0.f, stt = 0.f, k[9], e = 0.f;
#pragma unroll
        for (int i = 0; i < 9; ++i) k[i] = 0.f;
        auto acc1 = [&](float qx, float qy, float qz, float ux, float uy, float uz) {
            const float dx = qx - ux, dy = qy - uy, dz = qz - uz;
            const float px = qx - p0x, py = qy - p0y, pz = qz - p0z;
            const float tx = ux - t0x, ty = uy - t0y, tz = uz - t0z;
            sp[0] += px; sp[1] += py; sp[2] += pz;
            st[0] += tx; st[1] += ty; st[2] += tz;
            spp += px * px + py * py + pz * pz;
            stt += tx * tx + ty * ty + tz * tz;
            k[0] += px * tx; k[1] += px * ty; k[2] += px * tz;
            k[3] += py * tx; k[4] += py * ty; k[5] += py * tz;
            k[6] += pz * tx; k[7] += pz * ty; k[8] += pz * tz;
            e += sqrtf(dx * dx + dy * dy + dz * dz);
        };
#pragma unroll 2
        for (int gq = threadIdx.x; gq < ng; gq += ME_THREADS) {
            const Pts4 a = load4(p + gq * 12), u = load4(t + gq * 12);
#pragma unroll
            for (int q = 0; q < 4; ++q) acc1(a.x[3 * q], a.x[3 * q + 1], a.x[3 * q + 2], u.x[3 * q], u.x[3 * q + 1], u.x[3 * q + 2]);
        }
        for (int i = (ng << 2) + threadIdx.x; i < P; i += ME_THREADS)
            acc1(__ldg(p + i * 3), __ldg(p + i * 3 + 1), __ldg(p + i * 3 + 2), __ldg(t + i * 3), __ldg(t + i * 3 + 1), __ldg(t + i * 3 + 2));
        double v[18];
        for (int i = 0; i < 3; ++i) { v[i] = sp[i]; v[3 + i] = st[i]; }
        v[6] = spp; v[7] = stt;
        for (int i = 0; i < 9; ++i) v[8 + i] = k[i];
        v[17] = e;
        block_sum<18>(v, scratch);
        if (threadIdx.x < 18) {
            double mine = 0.0;
#pragma unroll
            for (int i = 0; i < 18; ++i) mine = (threadIdx.x == i) ? v[i] : mine;
            mom[(size_t)m * 18 + threadIdx.x] = mine;
        }
    } else {
        const float* xf = xfg + (size_t)m * 20;
        const float s_sc = __ldg(xf), m1x = __ldg(xf + 1), m1y = __ldg(xf + 2), m1z = __ldg(xf + 3), m2x = __ldg(xf + 4), m2y = __ldg(xf + 5), m2z = __ldg(xf + 6);
        const float r0 = __ldg(xf + 7), r1 = __ldg(xf + 8), r2 = __ldg(xf + 9), r3 = __ldg(xf + 10), r4 = __ldg(xf + 11), r5 = __ldg(xf + 12),
                    r6 = __ldg(xf + 13), r7 = __ldg(xf + 14), r8 = __ldg(xf + 15), t0 = __ldg(xf + 16), t1 = __ldg(xf + 17), t2 = __ldg(xf + 18);
        float esc = 0.f, epa = 0.f;
        auto acc2 = [&](float px, float py, float pz, float tx, float ty, float tz) {
            float dx = (px - m1x) * s_sc + m2x - tx, dy = (py - m1y) * s_sc + m2y - ty, dz = (pz - m1z) * s_sc + m2z - tz;
            esc += sqrtf(dx * dx + dy * dy + dz * dz);
            dx = r0 * px + r1 * py + r2 * pz + t0 - tx;
            dy = r3 * px + r4 * py + r5 * pz + t1 - ty;
            dz = r6 * px + r7 * py + r8 * pz + t2 - tz;
            epa += sqrtf(dx * dx + dy * dy + dz * dz);
        };
#pragma unroll 2
        for (int gq = threadIdx.x; gq < ng; gq += ME_THREADS) {
            const Pts4 a = load4(p + gq * 12), u = load4(t + gq * 12);
#pragma unroll
            for (int q = 0; q < 4; ++q) acc2(a.x[3 * q], a.x[3 * q + 1], a.x[3 * q + 2], u.x[3 * q], u.x[3 * q + 1], u.x[3 * q + 2]);
        }
        for (int i = (ng << 2) + threadIdx.x; i < P; i += ME_THREADS)
            acc2(__ldg(p + i * 3), __ldg(p + i * 3 + 1), __ldg(p + i * 3 + 2), __ldg(t + i * 3), __ldg(t + i * 3 + 1), __ldg(t + i * 3 + 2));
        double e2[2] = {esc, epa};
        block_sum<2>(e2, scratch);
        if (threadIdx.x == 0) {
            out[(size_t)m * 3 + 0] = __ldg(xf + 19);
            out[(size_t)m * 3 + 1] = (float)(e2[0] / (double)P);
            out[(size_t)m * 3 + 2] = (float)(e2[1] / (double)P);
        }
    }
}

// Per-image sample statistics of point sets (B, N, P, D), D in {2, 3} (metrics/eval_metrics_tracker.py:330-433):
//   out[b][0] = mean over (samples, points) of w_p ||x_{n,p} - mean_n x_{.,p}||            sample diversity (:397-433)
//   out[b][1] = sum over (samples, points) of w_p ||x_{n,p} - t_p|| / (N * sum_p w_p)       samples-L2E (:339-374); 0 without target
// w = per-image point weights (visibility flags) or 1.  One block per image, a thread per point (strided), fp64 block sums
// in a fixed order.
template <int D>
__global__ void __launch_bounds__(SS_THREADS)
sample_stats_kernel(const float* __restrict__ pts, const float* __restrict__ target, const float* __restrict__ weights, int N, int P,
                    int parts, float* __restrict__ out) {
    // `parts` threads share a point: thread (p, part) takes the samples n = part, part + parts, ...  (joint sets have fewer points
    // than the block has threads; one thread per point walked 2 x N dependent loads and took 20 us for 100 x 90 joints).  The
    // partial sums of the mean meet in shared memory and are added in part order.
    HF_PDL_SYNC();
    extern __shared__ float mu_part[];                       // [parts][P][D]
    __shared__ double scratch[(SS_THREADS / 32) * 3 + 3];
    const int b = blockIdx.x;
    const float* xb = pts + (size_t)b * N * P * D;
    double v[3] = {0.0, 0.0, 0.0};       // diversity sum, L2E sum, weight sum
    const int npp = P * parts;
    for (int base = 0; base < npp; base += SS_THREADS) {
        const int i = base + threadIdx.x;
        const bool act = i < npp;
        const int p = act ? i / parts : 0, part = act ? i - p * parts : 0;
        float mu[D];
#pragma unroll
        for (int d = 0; d < D; ++d) mu[d] = 0.f;
        if (act) {
            for (int n = part; n < N; n += parts) {
                const float* x = xb + ((size_t)n * P + p) * D;
#pragma unroll
                for (int d = 0; d < D; ++d) mu[d] += __ldg(x + d);
            }
            if (parts > 1) {
#pragma unroll
                for (int d = 0; d < D; ++d) mu_part[((size_t)part * P + p) * D + d] = mu[d];
            }
        }
        if (parts > 1) {
            __syncthreads();
            if (act) {
#pragma unroll
                for (int d = 0; d < D; ++d) {
                    float a = 0.f;
                    for (int q = 0; q < parts; ++q) a += mu_part[((size_t)q * P + p) * D + d];
                    mu[d] = a;
                }
            }
            __syncthreads();
        }
        if (act) {
#pragma unroll
            for (int d = 0; d < D; ++d) mu[d] /= (float)N;
            const float w = weights ? __ldg(weights + (size_t)b * P + p) : 1.f;
            float t[D];
#pragma unroll
            for (int d = 0; d < D; ++d) t[d] = target ? __ldg(target + ((size_t)b * P + p) * D + d) : 0.f;
            float div = 0.f, l2e = 0.f;
            for (int n = part; n < N; n += parts) {
                const float* x = xb + ((size_t)n * P + p) * D;
                float a = 0.f, c = 0.f;
#pragma unroll
                for (int d = 0; d < D; ++d) { const float xv = __ldg(x + d); a += (xv - mu[d]) * (xv - mu[d]); c += (xv - t[d]) * (xv - t[d]); }
                div += sqrtf(a); l2e += sqrtf(c);
            }
            v[0] += (double)(w * div); v[1] += (double)(w * l2e);
            if (part == 0) v[2] += (double)w;
        }
    }
    block_sum<3, SS_THREADS>(v, scratch);
    if (threadIdx.x == 0) {
        out[(size_t)b * 2] = (float)(v[0] / ((double)N * (double)P));
        out[(size_t)b * 2 + 1] = target ? (float)(v[1] / ((double)N * v[2])) : 0.f;
    }
}

// out[b][k] = min over the samples, out[b][K + k] = mean over the samples of err[b][n][k]; one CTA (128 threads) per image,
// fixed reduction order (thread-strided partials, then a shuffle / shared-memory tree).
__global__ void __launch_bounds__(128)
samples_reduce_kernel(const float* __restrict__ err, int N, int K, float* __restrict__ out) {
    HF_PDL_SYNC();
    __shared__ float s_min[4][8], s_sum[4][8];
    const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int k0 = 0; k0 < K; k0 += 8) {
        float mn[8], sm[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) { mn[k] = INFINITY; sm[k] = 0.f; }
        for (int n = tid; n < N; n += 128) {
#pragma unroll
            for (int k = 0; k < 8; ++k)
                if (k0 + k < K) {
                    const float v = err[((size_t)b * N + n) * K + k0 + k];
                    mn[k] = fminf(mn[k], v);
                    sm[k] += v;
                }
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
#pragma unroll
            for (int sft = 16; sft > 0; sft >>= 1) {
                mn[k] = fminf(mn[k], __shfl_xor_sync(0xffffffffu, mn[k], sft));
                sm[k] += __shfl_xor_sync(0xffffffffu, sm[k], sft);
            }
            if (lane == 0) { s_min[warp][k] = mn[k]; s_sum[warp][k] = sm[k]; }
        }
        __syncthreads();
        if (tid < 8 && k0 + tid < K) {
            out[(size_t)b * 2 * K + k0 + tid] = fminf(fminf(s_min[0][tid], s_min[1][tid]), fminf(s_min[2][tid], s_min[3][tid]));
            out[(size_t)b * 2 * K + K + k0 + tid] = ((s_sum[0][tid] + s_sum[1][tid]) + (s_sum[2][tid] + s_sum[3][tid])) / (float)N;
        }
        __syncthreads();
    }
}

}  // namespace

extern "C" int hf_samples_reduce(const float* err, int B, int N, int K, float* out, void* stream) {
    if (!err || !out) return hf::fail(HF_ERR_INVALID, "hf_samples_reduce: null argument");
    if (B <= 0 || N <= 0 || K <= 0) return HF_OK;
    HF_CUDA(hf::launch_pdl(samples_reduce_kernel, dim3(B), dim3(128), 0, (cudaStream_t)stream, err, N, K, out));
    HF_LAUNCH_CHECK();
    return HF_OK;
}

extern "C" int hf_sample_stats(const float* points, const float* target, const float* weights, int B, int N, int P, int D, float* out,
                               void* stream) {
    if (!points || !out) return hf::fail(HF_ERR_INVALID, "hf_sample_stats: null argument");
    if (B <= 0 || N <= 0 || P <= 0) return HF_OK;
    int parts = std::max(1, std::min(SS_THREADS / P, std::min(N, 16)));
    while (parts > 1 && (size_t)parts * P * D * sizeof(float) > 40 * 1024) --parts;
    const size_t smem = parts > 1 ? (size_t)parts * P * D * sizeof(float) : 0;
    if (D == 3) HF_CUDA(hf::launch_pdl(sample_stats_kernel<3>, dim3(B), dim3(SS_THREADS), smem, (cudaStream_t)stream, points, target, weights, N, P, parts, out));
    else if (D == 2) HF_CUDA(hf::launch_pdl(sample_stats_kernel<2>, dim3(B), dim3(SS_THREADS), smem, (cudaStream_t)stream, points, target, weights, N, P, parts, out));
    else return hf::fail(HF_ERR_INVALID, "hf_sample_stats: D must be 2 or 3");
    HF_LAUNCH_CHECK();
    return HF_OK;
}

extern "C" size_t hf_pointset_errors_workspace_bytes(int B, int N) {
    if (B <= 0 || N <= 0) return 0;
    return (size_t)B * N * (18 * sizeof(double) + 20 * sizeof(float));
}

extern "C" int hf_pointset_errors(const float* pred, const float* target, int B, int N, int P, float* out, void* stream) {
    return hf_pointset_errors_ws(pred, target, B, N, P, out, nullptr, 0, stream);
}

extern "C" int hf_pointset_errors_ws(const float* pred, const float* target, int B, int N, int P, float* out, void* workspace,
                                     size_t workspace_bytes, void* stream) {
    if (!pred || !target || !out) return hf::fail(HF_ERR_INVALID, "hf_pointset_errors: null argument");
    if (B <= 0 || N <= 0) return HF_OK;
    if (P < 3) return hf::fail(HF_ERR_INVALID, "hf_pointset_errors: at least 3 points per set are needed (got %d)", P);
    const size_t stage_bytes = (size_t)P * 3 * sizeof(float);
    const int M = B * N;
    if (workspace && M >= 512 && ((uintptr_t)pred & 7) == 0 && ((uintptr_t)target & 7) == 0 && (P * 3) % 2 == 0) {
        // many sets: moments / solve / errors as three launches (see pse_pass_kernel)
        if (workspace_bytes < hf_pointset_errors_workspace_bytes(B, N) || ((uintptr_t)workspace & 7))
            return hf::fail(HF_ERR_INVALID, "hf_pointset_errors_ws: workspace too small or unaligned (%zu < %zu)", workspace_bytes, hf_pointset_errors_workspace_bytes(B, N));
        double* mom = static_cast<double*>(workspace);
        float* xfg = reinterpret_cast<float*>(mom + (size_t)M * 18);
        HF_CUDA(hf::launch_pdl(pse_pass_kernel<1>, dim3(M), dim3(ME_THREADS), 0, (cudaStream_t)stream, pred, target, N, P, mom, (const float*)nullptr, (float*)nullptr));
        HF_LAUNCH_CHECK();
        HF_CUDA(hf::launch_pdl(pse_solve_kernel, dim3(hf::div_up(M, 128)), dim3(128), 0, (cudaStream_t)stream, pred, target, (const double*)mom, N, P, M, xfg));
        HF_LAUNCH_CHECK();
        HF_CUDA(hf::launch_pdl(pse_pass_kernel<2>, dim3(M), dim3(ME_THREADS), 0, (cudaStream_t)stream, pred, target, N, P, (double*)nullptr, (const float*)xfg, out));
        HF_LAUNCH_CHECK();
        return HF_OK;
    }
    if (stage_bytes <= 100 * 1024 && ((uintptr_t)pred & 7) == 0 && ((uintptr_t)target & 7) == 0 && (P * 3 * sizeof(float)) % 8 == 0) {    // two sets per SM
        HF_CUDA(cudaFuncSetAttribute(pointset_errors_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        HF_CUDA(hf::launch_pdl(pointset_errors_kernel<true>, dim3(B * N), dim3(ME_THREADS), stage_bytes, (cudaStream_t)stream, pred, target, N, P, out));
    } else
        HF_CUDA(hf::launch_pdl(pointset_errors_kernel<false>, dim3(B * N), dim3(ME_THREADS), 0, (cudaStream_t)stream, pred, target, N, P, out));
    HF_LAUNCH_CHECK();
    return HF_OK;
}
