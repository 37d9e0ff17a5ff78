// SMPL linear blend skinning for sm_100a: shape+pose blend, joint regression, kinematic chain,
// skinning and the reference's 90-joint assembly.
//
// Replaces models/smpl.py:27-41 and, beneath it, smplx 0.1.26 lbs()/batch_rigid_transform()/
// vertices2joints()/VertexJointSelector.
//
// HBM layout (built once by hf_smpl_create, immutable, ~19 MB -> L2 resident on B200):
//   blend [KB][3][Vp]  KB = num_betas + 9*(J-1); row l<nb = shapedirs[:,c,l], row nb+k = posedirs[k][3v+c]
//   vtemp [3][Vp]      v_template, structure-of-arrays so that a warp's 32 vertices are 128 contiguous bytes
//   J0 [J][3], Jd [J][3][nb]   J_regressor folded through v_template / shapedirs (fp64 on the host)
//   sj/sw [nslots][Vp] sparse skinning weights (SMPL: <=4 influences per vertex)
// Per call (workspace): F [M][KP] blend coefficients (betas | vec(R_i - I)), A [M][J][3][4] relative transforms.
#include "common.cuh"
#include "tma.cuh"
#include <cuda_fp16.h>
#include <vector>
#include <cmath>
#include <cstring>
#include <algorithm>

#define HF_MAXJ 24
#define HF_MAXB 16
#define LBS_KH 256          // K of one bf16 half (num_betas + 9*(J-1) <= 223, padded)
// fp16 single-pass layout (product path): K = 256 = [pose 0..207 | beta_hi 208..223 | beta_lo 224..239 | beta_hi 240..255]
// against basis columns              [posedirs        | shape_hi        | shape_hi        | shape_lo        ]
// i.e. the 207 pose terms take one fp16 x fp16 product each (fp32 accumulation; measured <= 5e-5 m at pose std 0.8)
// and the 10 shape terms, which carry ten times the magnitude, the three-product split (error ~1e-7 m).
#define LBS_K2 256
#define LBS_K2_POSE 208

struct hf_smpl {
    int V, Vp, nb, J, KB, KP, nslots, nvj, nextra, nnz;
    float *blend, *vtemp, *J0, *Jd, *sw;
    int *sj, *vj, *csr_ptr, *csr_col;
    float* csr_val;
    int parents[HF_MAXJ];
    // tensor-core path: blend basis as split bf16 [3][Vp][KT] = [hi(256) | lo(256)] per (coordinate, vertex) row
    __nv_bfloat16* Pbf;
    CUtensorMap mapA;
    int impl;                 // 0 = persistent fp16 tcgen05 blend (product path), 1 = FP32 CUDA-core blend (debug
                              // cross-check), 2 = split-bf16 three-pass tcgen05 blend (high-precision cross-check),
                              // 3 = lanes-as-samples fp16 tcgen05 blend (round-2 experiment, slower: DESIGN.md 4.1)
    // product path: fp16 basis [3][Vp][LBS_K2] scaled by 2^k = pose block | shape hi | shape hi | shape lo
    __half* Pf16;
    float inv_scale;
    // vertices read by the joint picks / extra regressors (~4 %) are flagged: the product skinning kernel stores them with an
    // L2 evict_last policy (everything else streams evict_first), so the extra-joint kernel's gathers hit L2 instead of DRAM
    int* vflag;
    CUtensorMap mapA2;
    const void* mapB2_ptr; int mapB2_M; CUtensorMap mapB2;
    // round-2 experiment (impl 3, lbs_skin_tc3_kernel): lanes = samples.  Basis rows re-ordered so that inside every
    // 32-vertex tile the vertices are sorted by their skinning-joint signature (runs of equal joints -> the 3x4 transforms
    // stay in registers); per sorted position one 64-byte record {v_template, original index, weights, joints, flagged slot}
    __half* Pf16s; float4* vconst; CUtensorMap mapP3;
    const void* mapF3_ptr; int mapF3_M; CUtensorMap mapF3;
    int NF;                   // vertices read by the joint picks / extra regressors ("flagged"), each with a slot in xvt
    int *pick_f, *csr_f;      // flagged slot of every pick / CSR entry
    // cached tensor map of the per-call coefficient matrix
    const void* mapB_ptr; int mapB_M; CUtensorMap mapB;
    // backward: transpose of (joint picks + extra regressors) per vertex, rows numbered from the first
    // non-chain output joint; blend basis as split tf32 [256][2 * 3Vp] = [hi | lo] (built on the first hf_lbs_backward call)
    int *csc_ptr, *csc_row; float* csc_val;
    float* blend_split; CUtensorMap mapBs;
    const void* mapG_ptr; int mapG_M; CUtensorMap mapG;
    // joints-only backward: everything restricted to the NF vertices that a pick / regressor row reads (~4 %), padded to NFp
    int NFp; int* fvert;                         // compact slot -> vertex
    float* blend_split_c; CUtensorMap mapBsc;    // [256][hi 3*NFp | lo 3*NFp]
    float* blend_c;                              // [KB][3][NFp] fp32: the compact columns of `blend` (coalesced reads in the vertex pass)
    const void* mapGc_ptr; int mapGc_M; CUtensorMap mapGc;
};

int hf_lbs_extra_joints(const hf_smpl* h, const float* vertices, float* joints, int M, cudaStream_t stream);

namespace {

struct Parents { int p[HF_MAXJ]; };

// Kinematic chain, joint regression and blend coefficients.  FOUR LANES per sample (8 samples per warp): lane r < 3 owns
// row r of every 3x4 global transform, G_i[r][:] = G_parent[r][0..2] . [R_i | t_i] (+ G_parent[r][3]), which depends on
// row r of the parent only, so the 24-step chain needs no cross-lane traffic and every lane keeps its own rows in
// shared memory (dynamic parent index).  The samples' rotation matrices are staged into shared memory with coalesced
// loads first, so the chain runs on shared-memory operands only.  Outputs: relative transforms A_i (one 16-byte store
// per lane and joint), posed joints, and the blend coefficients (fp16 layout LBS_K2 for the product path -> Fh, or
// fp32 -> F for the CUDA-core blend).
// One block = PS samples, 4 warps: all warps stage the inputs and regress the joints, then warp 0 walks the chain while
// warps 1..3 write the blend coefficients.
constexpr int PS = 8, PTHREADS = 128, PR = HF_MAXJ * 9 + 1;
__global__ void __launch_bounds__(PTHREADS)
lbs_pose_kernel(const float* __restrict__ betas, const float* __restrict__ rotmats, const float* __restrict__ glob, int rep,
                                const float* __restrict__ transl, const float* __restrict__ J0,
                                const float* __restrict__ Jd, Parents par, int M, int J, int nb, int KP,
                                int J_out, float* __restrict__ F, __half* __restrict__ Fh, float* __restrict__ A,
                                float* __restrict__ At, float* __restrict__ joints) {
    __shared__ __align__(16) float Gs[HF_MAXJ][PS][12];
    __shared__ float Js[HF_MAXJ][PS][3];
    __shared__ float Rs[PS][PR];
    __shared__ float Bs[PS][HF_MAXB];
    __shared__ float Jds[HF_MAXJ * 3 * HF_MAXB], J0s[HF_MAXJ * 3];
    __shared__ int pars[HF_MAXJ];          // kinematic tree: a dynamically indexed by-value kernel parameter compiles into one
                                           // predicated constant load per joint per step (ncu: a third of this kernel's stall samples)
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid < HF_MAXJ) pars[tid] = par.p[tid];
    const int mb = blockIdx.x * PS;                              // first sample of this block
    const int ns = min(PS, M - mb);
    const int J9 = J * 9, J3 = J * 3;
    for (int k = tid; k < J3 * nb; k += PTHREADS) Jds[k] = Jd[k];       // model constants: no dependency on the predecessor
    for (int k = tid; k < J3; k += PTHREADS) J0s[k] = J0[k];
    HF_PDL_SYNC();
    {
        if (glob) {   // split form: rotmats = body rotations (M, J-1, 3, 3), glob (M / rep, 3, 3) shared by `rep` consecutive samples
            const int B9 = J9 - 9;
            const float* Rm = rotmats + (size_t)mb * B9;
            for (int e = tid; e < ns * B9; e += PTHREADS) {
                const int s2 = e / B9;
                Rs[s2][9 + e - s2 * B9] = __ldg(Rm + e);
            }
            for (int e = tid; e < ns * 9; e += PTHREADS) {
                const int s2 = e / 9;
                Rs[s2][e - s2 * 9] = __ldg(glob + (size_t)((mb + s2) / rep) * 9 + (e - s2 * 9));
            }
        } else {
        const float* Rm = rotmats + (size_t)mb * J9;
        for (int e = tid; e < ns * J9; e += PTHREADS) {
            const int s2 = e / J9;
            Rs[s2][e - s2 * J9] = __ldg(Rm + e);
        }
        }
        for (int e = tid; e < ns * nb; e += PTHREADS) {
            const int s2 = e / nb;
            Bs[s2][e - s2 * nb] = __ldg(betas + (size_t)mb * nb + e);
        }
    }
    __syncthreads();
    // joints of the shaped template
    for (int o = tid; o < ns * J3; o += PTHREADS) {
        const int s2 = o / J3, jr = o - s2 * J3;
        float acc = J0s[jr];
        const float* jd = Jds + jr * nb;
        for (int l = 0; l < nb; ++l) acc = fmaf(jd[l], Bs[s2][l], acc);
        Js[jr / 3][s2][jr % 3] = acc;
    }
    __syncthreads();
    if (warp > 0) {
        const int t = tid - 32, nt = PTHREADS - 32;
        if (F) {   // fp32 blend coefficients (beta | vec(R_i - I)), only for the CUDA-core blend
            for (int e = t; e < ns * KP; e += nt) {
                const int s2 = e / KP, k = e - s2 * KP;
                float f = 0.f;
                if (k < nb) f = Bs[s2][k];
                else if (k < nb + 9 * (J - 1)) {
                    const int q = k - nb;
                    f = Rs[s2][9 + q] - (((q % 9) % 4 == 0) ? 1.f : 0.f);
                }
                F[(size_t)mb * KP + e] = f;
            }
        }
        if (Fh) {  // fp16 blend coefficients, layout LBS_K2: [vec(R_i - I) (208) | beta_hi (16) | beta_lo (16) | beta_hi (16)]
            __half2* Fm = reinterpret_cast<__half2*>(Fh + (size_t)mb * LBS_K2);
            for (int e = t; e < ns * (LBS_K2 / 2); e += nt) {
                const int s2 = e / (LBS_K2 / 2), k2 = e - s2 * (LBS_K2 / 2);
                float f[2];
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const int k = 2 * k2 + u;
                    float v = 0.f;
                    if (k < LBS_K2_POSE) {
                        if (k < 9 * (J - 1)) v = Rs[s2][9 + k] - (((k % 9) % 4 == 0) ? 1.f : 0.f);
                    } else {
                        const int blk = (k - LBS_K2_POSE) >> 4, l = (k - LBS_K2_POSE) & 15;   // blk 0: hi, 1: lo, 2: hi
                        if (l < nb) {
                            const float b = Bs[s2][l];
                            const float hi = __half2float(__float2half_rn(b));
                            v = (blk == 1) ? b - hi : hi;
                        }
                    }
                    f[u] = v;
                }
                Fm[e] = __floats2half2_rn(f[0], f[1]);
            }
        }
        return;
    }
    // warp 0: FOUR LANES per sample; lane r < 3 owns row r of every 3x4 global transform,
    // G_i[r][:] = G_parent[r][0..2] . [R_i | t_i] (+ G_parent[r][3]), which depends on row r of the parent only: the
    // 24-step chain needs no cross-lane traffic, every lane keeps its own rows in shared memory (dynamic parent index)
    const int sl = lane >> 2, r = lane & 3;
    const int m = mb + sl;
    if (sl >= ns || r == 3) return;
    float tr = 0.f;
    if (transl) tr = __ldg(transl + m * 3 + r);
    const float* Rm = Rs[sl];
    float* Am = A + (size_t)m * J * 12 + r * 4;
    float* jm = joints + (size_t)m * J_out * 3 + r;
    for (int i = 0; i < J; ++i) {
        const int p = pars[i];
        const float* Ri = Rm + i * 9;
        const float jx = Js[i][sl][0], jy = Js[i][sl][1], jz = Js[i][sl][2];
        float4 g;
        if (i == 0) {
            g = make_float4(Ri[r * 3], Ri[r * 3 + 1], Ri[r * 3 + 2], r == 0 ? jx : (r == 1 ? jy : jz));
        } else {
            const float4 gp = *reinterpret_cast<const float4*>(&Gs[p][sl][r * 4]);
            const float tx = jx - Js[p][sl][0], ty = jy - Js[p][sl][1], tz = jz - Js[p][sl][2];
            g.x = gp.x * Ri[0] + gp.y * Ri[3] + gp.z * Ri[6];
            g.y = gp.x * Ri[1] + gp.y * Ri[4] + gp.z * Ri[7];
            g.z = gp.x * Ri[2] + gp.y * Ri[5] + gp.z * Ri[8];
            g.w = gp.x * tx + gp.y * ty + gp.z * tz + gp.w;
        }
        *reinterpret_cast<float4*>(&Gs[i][sl][r * 4]) = g;
        const float aw = g.w - (g.x * jx + g.y * jy + g.z * jz);
        if (At)     // layout of the round-2 skinning kernel: At[sample tile][joint][128 samples][3 x 4]
            *reinterpret_cast<float4*>(At + (((size_t)(m >> 7) * J + i) * 128 + (m & 127)) * 12 + r * 4) = make_float4(g.x, g.y, g.z, aw);
        else
        *reinterpret_cast<float4*>(Am + i * 12) = make_float4(g.x, g.y, g.z, aw);
        jm[i * 3] = g.w + tr;
    }
}

// Split-bf16 blend coefficients for the tensor-core path: Fb[m] = [hi(KH) | lo(KH)] of (beta | vec(R_i - I)),
// zero padded; one thread per element, coalesced bf16 writes.
__global__ void lbs_coef_kernel(const float* __restrict__ betas, const float* __restrict__ rotmats, int split, int M, int J, int nb,
                                __nv_bfloat16* __restrict__ Fb) {
    HF_PDL_SYNC();
    const size_t total = (size_t)M * LBS_KH;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
        const int m = (int)(e / LBS_KH), k = (int)(e - (size_t)m * LBS_KH);
        float f = 0.f;
        if (k < nb) f = __ldg(betas + (size_t)m * nb + k);
        else if (k < nb + 9 * (J - 1)) {
            const int q = k - nb, i = q / 9 + 1, el = q - (i - 1) * 9;
            f = __ldg(rotmats + (split ? ((size_t)m * (J - 1) + i - 1) : ((size_t)m * J + i)) * 9 + el) - ((el % 4 == 0) ? 1.f : 0.f);
        }
        const __nv_bfloat16 hi = __float2bfloat16_rn(f);
        Fb[(size_t)m * 2 * LBS_KH + k] = hi;
        Fb[(size_t)m * 2 * LBS_KH + LBS_KH + k] = __float2bfloat16_rn(f - __bfloat162float(hi));
    }
}

// Tile = 128 vertices x TS samples.  Thread (vl, sg) owns one vertex and SPT samples: the blend
// contraction runs as SPT*3 independent FMA chains per thread with the basis read coalesced from L2
// (the SG sample groups of a CTA share the lines through L1) and coefficients broadcast from smem.
template <int SPT, int SG>
__global__ void __launch_bounds__(128 * SG, 1)
lbs_skin_kernel(const float* __restrict__ blend, const float* __restrict__ vtemp,
                const int* __restrict__ sj, const float* __restrict__ sw, const float* __restrict__ F,
                const float* __restrict__ A, const float* __restrict__ transl, int M, int V, int Vp, int KB,
                int KP, int J, int nslots, float* __restrict__ vertices) {
    constexpr int TS = SPT * SG;
    extern __shared__ __align__(16) float smem[];
    float* Fs = smem;                 // [KP][TS]
    float* As = smem + KP * TS;       // [TS][J*12]
    const int J12 = J * 12;
    const int m0 = blockIdx.y * TS;
    const int tid = threadIdx.x;
    for (int idx = tid; idx < TS * KP; idx += 128 * SG) {
        int s = idx / KP, k = idx - s * KP;
        int m = m0 + s;
        Fs[k * TS + s] = (m < M) ? F[(size_t)m * KP + k] : 0.f;
    }
    for (int idx = tid; idx < TS * J12; idx += 128 * SG) {
        int s = idx / J12;
        int m = m0 + s;
        As[idx] = (m < M) ? A[(size_t)m * J12 + (idx - s * J12)] : 0.f;
    }
    __syncthreads();

    const int vl = tid & 127, sg = tid >> 7;
    const int v = blockIdx.x * 128 + vl;          // < Vp always (arrays are padded)
    float acc[SPT][3];
#pragma unroll
    for (int s = 0; s < SPT; ++s) acc[s][0] = acc[s][1] = acc[s][2] = 0.f;
    const float* bp = blend + v;
    const float* fs = Fs + sg * SPT;
#pragma unroll 4
    for (int k = 0; k < KB; ++k) {
        float p0 = __ldg(bp + (size_t)(k * 3 + 0) * Vp);
        float p1 = __ldg(bp + (size_t)(k * 3 + 1) * Vp);
        float p2 = __ldg(bp + (size_t)(k * 3 + 2) * Vp);
        const float4* f4 = reinterpret_cast<const float4*>(fs + k * TS);
#pragma unroll
        for (int q = 0; q < SPT / 4; ++q) {
            float4 f = f4[q];
            acc[q * 4 + 0][0] = fmaf(p0, f.x, acc[q * 4 + 0][0]); acc[q * 4 + 0][1] = fmaf(p1, f.x, acc[q * 4 + 0][1]); acc[q * 4 + 0][2] = fmaf(p2, f.x, acc[q * 4 + 0][2]);
            acc[q * 4 + 1][0] = fmaf(p0, f.y, acc[q * 4 + 1][0]); acc[q * 4 + 1][1] = fmaf(p1, f.y, acc[q * 4 + 1][1]); acc[q * 4 + 1][2] = fmaf(p2, f.y, acc[q * 4 + 1][2]);
            acc[q * 4 + 2][0] = fmaf(p0, f.z, acc[q * 4 + 2][0]); acc[q * 4 + 2][1] = fmaf(p1, f.z, acc[q * 4 + 2][1]); acc[q * 4 + 2][2] = fmaf(p2, f.z, acc[q * 4 + 2][2]);
            acc[q * 4 + 3][0] = fmaf(p0, f.w, acc[q * 4 + 3][0]); acc[q * 4 + 3][1] = fmaf(p1, f.w, acc[q * 4 + 3][1]); acc[q * 4 + 3][2] = fmaf(p2, f.w, acc[q * 4 + 3][2]);
        }
    }
    const float t0 = vtemp[v], t1 = vtemp[Vp + v], t2 = vtemp[2 * Vp + v];
#pragma unroll
    for (int s = 0; s < SPT; ++s) { acc[s][0] += t0; acc[s][1] += t1; acc[s][2] += t2; }

    // skinning: v = sum_k w_k (R_k p + t_k), A rows gathered from smem (vertices of a warp mostly
    // share joints, so the 128-bit loads are broadcasts).
    constexpr int HS = (SPT >= 8) ? 8 : SPT;
#pragma unroll
    for (int h0 = 0; h0 < SPT; h0 += HS) {
        float o[HS][3];
#pragma unroll
        for (int s = 0; s < HS; ++s) o[s][0] = o[s][1] = o[s][2] = 0.f;
        for (int slot = 0; slot < nslots; ++slot) {
            const int j = sj[slot * Vp + v];
            const float w = sw[slot * Vp + v];
            const float* a0 = As + (sg * SPT + h0) * J12 + j * 12;
#pragma unroll
            for (int s = 0; s < HS; ++s) {
                const float4* a = reinterpret_cast<const float4*>(a0 + s * J12);
                float4 r0 = a[0], r1 = a[1], r2 = a[2];
                float px = acc[h0 + s][0], py = acc[h0 + s][1], pz = acc[h0 + s][2];
                o[s][0] = fmaf(w, fmaf(r0.x, px, fmaf(r0.y, py, fmaf(r0.z, pz, r0.w))), o[s][0]);
                o[s][1] = fmaf(w, fmaf(r1.x, px, fmaf(r1.y, py, fmaf(r1.z, pz, r1.w))), o[s][1]);
                o[s][2] = fmaf(w, fmaf(r2.x, px, fmaf(r2.y, py, fmaf(r2.z, pz, r2.w))), o[s][2]);
            }
        }
        if (v < V) {
#pragma unroll
            for (int s = 0; s < HS; ++s) {
                int m = m0 + sg * SPT + h0 + s;
                if (m < M) {
                    float tx = 0.f, ty = 0.f, tz = 0.f;
                    if (transl) { tx = transl[m * 3]; ty = transl[m * 3 + 1]; tz = transl[m * 3 + 2]; }
                    float* out = vertices + ((size_t)m * V + v) * 3;
                    __stcs(out + 0, o[s][0] + tx);
                    __stcs(out + 1, o[s][1] + ty);
                    __stcs(out + 2, o[s][2] + tz);
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Tensor-core blend + CUDA-core skinning.  Tile = 128 vertices x 64 samples.
//   D_c[v][s] = sum_k P_c[v][k] f[s][k]   for c in {x,y,z}, as split-bf16 (hi*hi + hi*lo + lo*hi, fp32 accumulate in
//   TMEM): A = basis rows (vertex-major, K-major) via TMA from the L2-resident [3][Vp][512] table, B = the per-sample
//   coefficients [M][512].  Per k-block stage: A_hi, A_lo (16 KB each), B_hi, B_lo (8 KB each); 3 products x 4 UMMA_K.
//   Epilogue (8 warps; warp w owns TMEM lanes 32*(w%4).., samples 32*(w/4)..): v_posed = D + v_template, then
//   v = sum_k w_k (R_k v_posed + t_k) with the 3x4 transforms of the tile's 64 samples in shared memory.
constexpr int TC_NS = 64, TC_STAGES = 3, TC_THREADS = 320;
constexpr int TC_STAGE_BYTES = 2 * 128 * 128 + 2 * TC_NS * 128;

__global__ void __launch_bounds__(TC_THREADS, 1)
lbs_skin_tc_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                   const float* __restrict__ vtemp, const int* __restrict__ sj, const float* __restrict__ sw,
                   const float* __restrict__ A, const float* __restrict__ transl, int M, int V, int Vp, int J,
                   int nslots, float* __restrict__ vertices) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bars[2 * TC_STAGES + 2];
    __shared__ uint32_t tmem_base_s;
    const uint32_t tile_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    float* As = reinterpret_cast<float*>(smem_raw + (tile_base - smem_u32(smem_raw)) + TC_STAGES * TC_STAGE_BYTES);   // [TC_NS][J*12]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int v0 = blockIdx.x * 128, m0 = blockIdx.y * TC_NS;
    const int J12 = J * 12;
    const uint32_t full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[TC_STAGES]), tfull = smem_u32(&bars[2 * TC_STAGES]);
    const uint32_t abar = smem_u32(&bars[2 * TC_STAGES + 1]);
    if (threadIdx.x == 0) {
        for (int s = 0; s < TC_STAGES; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
        mbar_init(tfull, 1);
        mbar_init(abar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(256u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const uint32_t tmem_base = tmem_base_s;

    constexpr int NIT = 3 * (LBS_KH / 64);          // (coordinate, k-block) iterations
    if (warp == 0) {
        if (lane == 0) {
            {   // the tile's rigid transforms: one contiguous block of the workspace -> one bulk copy
                const int rows = min(TC_NS, M - m0);
                const uint32_t bytes = (uint32_t)(rows * J12 * 4);
                mbar_expect_tx(abar, bytes);
                bulk_load_1d(tile_base + TC_STAGES * TC_STAGE_BYTES, A + (size_t)m0 * J12, bytes, abar);
            }
            for (int it = 0; it < NIT; ++it) {
                const int st = it % TC_STAGES;
                const uint32_t ph = (uint32_t)(it / TC_STAGES) & 1u;
                mbar_wait(empty0 + 8 * st, ph ^ 1u);
                const uint32_t sa = tile_base + st * TC_STAGE_BYTES;
                const uint32_t fb = full0 + 8 * st;
                mbar_expect_tx(fb, TC_STAGE_BYTES);
                const int c = it / (LBS_KH / 64), i = it - c * (LBS_KH / 64);
                tma_load_2d(sa, &mapA, fb, i * 64, c * Vp + v0);
                tma_load_2d(sa + 16384, &mapA, fb, LBS_KH + i * 64, c * Vp + v0);
                tma_load_2d(sa + 32768, &mapB, fb, i * 64, m0);
                tma_load_2d(sa + 32768 + TC_NS * 128, &mapB, fb, LBS_KH + i * 64, m0);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_bf16(128, TC_NS);
            for (int it = 0; it < NIT; ++it) {
                const int st = it % TC_STAGES;
                const uint32_t ph = (uint32_t)(it / TC_STAGES) & 1u;
                mbar_wait(full0 + 8 * st, ph);
                tcgen05_fence_after();
                const uint32_t sa = tile_base + st * TC_STAGE_BYTES;
                const uint64_t a_hi = umma_desc_sw128(sa), a_lo = umma_desc_sw128(sa + 16384);
                const uint64_t b_hi = umma_desc_sw128(sa + 32768), b_lo = umma_desc_sw128(sa + 32768 + TC_NS * 128);
                const int c = it / (LBS_KH / 64), i = it - c * (LBS_KH / 64);
                const uint32_t d = tmem_base + (uint32_t)(c * TC_NS);
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_bf16(d, a_hi + (uint64_t)(2 * k), b_hi + (uint64_t)(2 * k), idesc, (uint32_t)((i | k) != 0));
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_bf16(d, a_hi + (uint64_t)(2 * k), b_lo + (uint64_t)(2 * k), idesc, 1u);
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_bf16(d, a_lo + (uint64_t)(2 * k), b_hi + (uint64_t)(2 * k), idesc, 1u);
                umma_commit(empty0 + 8 * st);
            }
            umma_commit(tfull);
        }
    } else {
        const int q = warp & 3, half = (warp - 2) >> 2;
        const int v = v0 + q * 32 + lane;                     // < Vp (tables are padded)
        const float t0 = vtemp[v], t1 = vtemp[Vp + v], t2 = vtemp[2 * Vp + v];
        mbar_wait(abar, 0);
        mbar_wait(tfull, 0);
        tcgen05_fence_after();
#pragma unroll 1
        for (int chunk = 0; chunk < 2; ++chunk) {
            const int s0 = half * 32 + chunk * 16;             // first sample (tile-local) of this chunk
            uint32_t px[16], py[16], pz[16];
            const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)s0;
            tmem_ld16(ta, px);
            tmem_ld16(ta + TC_NS, py);
            tmem_ld16(ta + 2 * TC_NS, pz);
            tmem_ld_wait();
            float o[16][3];
#pragma unroll
            for (int s = 0; s < 16; ++s) o[s][0] = o[s][1] = o[s][2] = 0.f;
            for (int slot = 0; slot < nslots; ++slot) {
                const int j = sj[slot * Vp + v];
                const float w = sw[slot * Vp + v];
                const float* a0 = As + s0 * J12 + j * 12;
#pragma unroll
                for (int s = 0; s < 16; ++s) {
                    const float4* a = reinterpret_cast<const float4*>(a0 + s * J12);
                    const float4 r0 = a[0], r1 = a[1], r2 = a[2];
                    const float x = __uint_as_float(px[s]) + t0, y = __uint_as_float(py[s]) + t1, z = __uint_as_float(pz[s]) + t2;
                    o[s][0] = fmaf(w, fmaf(r0.x, x, fmaf(r0.y, y, fmaf(r0.z, z, r0.w))), o[s][0]);
                    o[s][1] = fmaf(w, fmaf(r1.x, x, fmaf(r1.y, y, fmaf(r1.z, z, r1.w))), o[s][1]);
                    o[s][2] = fmaf(w, fmaf(r2.x, x, fmaf(r2.y, y, fmaf(r2.z, z, r2.w))), o[s][2]);
                }
            }
            if (v < V) {
#pragma unroll
                for (int s = 0; s < 16; ++s) {
                    const int m = m0 + s0 + s;
                    if (m < M) {
                        float tx = 0.f, ty = 0.f, tz = 0.f;
                        if (transl) { tx = transl[m * 3]; ty = transl[m * 3 + 1]; tz = transl[m * 3 + 2]; }
                        float* out = vertices + ((size_t)m * V + v) * 3;
                        __stcs(out + 0, o[s][0] + tx);
                        __stcs(out + 1, o[s][1] + ty);
                        __stcs(out + 2, o[s][2] + tz);
                    }
                }
            }
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256u) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------
// Product path: persistent fp16 tensor-core blend + CUDA-core skinning.
//   unit = 128 vertices x 64 samples; units are numbered sample-tile major and every CTA (one per SM) walks a
//   contiguous range of them, so the 64 samples' coefficients (B operand, 32 KB) and 3x4 transforms (72 KB) stay
//   resident in shared memory while the basis tiles (A operand, 3 coordinates x 4 k-blocks of 16 KB) stream through a
//   TMA ring.  D_c[v][s] = sum_k P_c[v][k] f[s][k] accumulates in TMEM, double-buffered (2 x 3 x 64 columns), so the
//   skinning epilogue of unit u (16 warps) overlaps the loads and MMAs of unit u+1.
//   warp 0: TMA producer, warp 1: MMA issuer + TMEM owner, warps 2..17: epilogue; warp w reads TMEM lanes 32*(w%4)..
//   (= vertices) and 16 of the 64 samples.
constexpr int T2_NS = 64, T2_STAGES = 6, T2_EPI_WARPS = 16, T2_THREADS = (2 + T2_EPI_WARPS) * 32;
constexpr int T2_KBLK = LBS_K2 / 64, T2_NIT = 3 * T2_KBLK;
constexpr int T2_BRES_BYTES = T2_KBLK * T2_NS * 128;

__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    umma_bf16(tmem_d, desc_a, desc_b, idesc, accumulate);   // same instruction (kind::f16); the operand formats live in idesc
}
// instruction descriptor: fp16 x fp16 -> fp32, both operands K-major
__device__ __forceinline__ uint32_t umma_idesc_f16(int M, int N) {
    return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void st_hint(float* p, float v, uint64_t policy) {
    asm volatile("st.global.L2::cache_hint.f32 [%0], %1, %2;" ::"l"(p), "f"(v), "l"(policy) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t* v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr));
}

// NSLOT: skinning slots kept in registers (1..4; 0 = any number, read per chunk); TRANSL: add the per-sample translation
template <int NSLOT, bool TRANSL>
__global__ void __launch_bounds__(T2_THREADS, 1)
lbs_skin_tc2_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                    const float* __restrict__ vtemp, const int* __restrict__ sj, const float* __restrict__ sw,
                    const float* __restrict__ A, const float* __restrict__ transl, int M, int V, int Vp, int J,
                    int nslots, float inv_scale, int nvt, int num_units, float* __restrict__ vertices,
                    const int* __restrict__ vflag, float* __restrict__ xv, int NF) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bars[2 * T2_STAGES + 6];
    __shared__ uint32_t tmem_base_s;
    const uint32_t tile_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bres = tile_base + T2_STAGES * 16384;                    // resident B operand: 4 k-blocks [64 rows][128 B]
    const uint32_t as_addr = bres + T2_BRES_BYTES;                          // transforms [64][J*12] fp32
    const float* As = reinterpret_cast<const float*>(smem_raw + (as_addr - smem_u32(smem_raw)));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int J12 = J * 12;
    const uint32_t full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[T2_STAGES]);
    const uint32_t tfull0 = smem_u32(&bars[2 * T2_STAGES]), tempty0 = smem_u32(&bars[2 * T2_STAGES + 2]);
    const uint32_t bfull = smem_u32(&bars[2 * T2_STAGES + 4]), sdone = smem_u32(&bars[2 * T2_STAGES + 5]);
    if (threadIdx.x == 0) {
        for (int s = 0; s < T2_STAGES; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(tfull0 + 8 * b, 1); mbar_init(tempty0 + 8 * b, T2_EPI_WARPS); }
        mbar_init(bfull, 1);
        mbar_init(sdone, T2_EPI_WARPS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const uint32_t tmem_base = tmem_base_s;
    const int u0 = (int)((long long)blockIdx.x * num_units / gridDim.x);
    const int u1 = (int)((long long)(blockIdx.x + 1) * num_units / gridDim.x);

    if (warp == 0) {
        if (lane == 0) {
            uint32_t kbc = 0, sc = 0;
            int cur_st = -1;
            for (int u = u0; u < u1; ++u) {
                const int st = u / nvt, vt = u - st * nvt;
                const bool new_tile = st != cur_st;
                for (int it = 0; it < T2_NIT; ++it, ++kbc) {
                    const uint32_t sg = kbc % T2_STAGES, ph = (kbc / T2_STAGES) & 1u;
                    mbar_wait_long(empty0 + 8 * sg, ph ^ 1u);
                    const uint32_t fb = full0 + 8 * sg;
                    mbar_expect_tx(fb, 16384);
                    const int c = it / T2_KBLK, kb = it - c * T2_KBLK;
                    tma_load_2d(tile_base + sg * 16384, &mapA, fb, kb * 64, c * Vp + vt * 128);
                    // switch the resident sample tile once the ring is primed with the new unit's first blocks: the old
                    // tile's coefficients / transforms are free when every epilogue warp has finished its last unit
                    if (new_tile && it == (T2_STAGES < T2_NIT ? T2_STAGES : T2_NIT) - 1) {
                        if (cur_st >= 0) mbar_wait_long(sdone, (sc - 1) & 1u);
                        const int m0 = st * T2_NS;
                        const int rows = min(T2_NS, M - m0);
                        const uint32_t abytes = (uint32_t)(rows * J12 * 4);
                        mbar_expect_tx(bfull, (uint32_t)T2_BRES_BYTES + abytes);
#pragma unroll
                        for (int k = 0; k < T2_KBLK; ++k) tma_load_2d(bres + k * (T2_NS * 128), &mapB, bfull, k * 64, m0);
                        bulk_load_1d(as_addr, A + (size_t)m0 * J12, abytes, bfull);
                        cur_st = st; ++sc;
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_f16(128, T2_NS);
            uint32_t kbc = 0, lt = 0, sc = 0;
            int cur_st = -1;
            for (int u = u0; u < u1; ++u, ++lt) {
                const int st = u / nvt;
                if (st != cur_st) { mbar_wait_long(bfull, sc & 1u); ++sc; cur_st = st; }
                const uint32_t buf = lt & 1u;
                mbar_wait_long(tempty0 + 8 * buf, ((lt >> 1) & 1u) ^ 1u);
                tcgen05_fence_after();
                for (int it = 0; it < T2_NIT; ++it, ++kbc) {
                    const uint32_t sg = kbc % T2_STAGES, ph = (kbc / T2_STAGES) & 1u;
                    mbar_wait_long(full0 + 8 * sg, ph);
                    tcgen05_fence_after();
                    const int c = it / T2_KBLK, kb = it - c * T2_KBLK;
                    const uint64_t da = umma_desc_sw128(tile_base + sg * 16384), db = umma_desc_sw128(bres + kb * (T2_NS * 128));
                    const uint32_t d = tmem_base + buf * (3 * T2_NS) + (uint32_t)(c * T2_NS);
#pragma unroll
                    for (int k = 0; k < 4; ++k) umma_f16(d, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (uint32_t)((kb | k) != 0));
                    umma_commit(empty0 + 8 * sg);
                }
                umma_commit(tfull0 + 8 * buf);
            }
        }
    } else {
        const int q = warp & 3, sgrp = (warp - 2) >> 2;     // TMEM lane quarter (= vertex group), 16-sample group
        const size_t vstride = (size_t)V * 3;
        uint64_t pol_first, pol_last;
        asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_first));
        asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_last));
        uint32_t lt = 0, sc = 0;
        int cur_st = -1;
        for (int u = u0; u < u1; ++u, ++lt) {
            const int st = u / nvt, vt = u - st * nvt;
            const uint32_t buf = lt & 1u;
            const int v = vt * 128 + q * 32 + lane;                 // < Vp (tables are padded)
            const int m0 = st * T2_NS;
            // per-vertex constants first (global loads in flight while the accumulator is still being produced)
            const float t0 = __ldg(vtemp + v), t1 = __ldg(vtemp + Vp + v), t2 = __ldg(vtemp + 2 * Vp + v);
            const int fidx = __ldg(vflag + v);                       // >= 0: vertex feeds a joint pick / regressor (slot in xv)
            const uint64_t pol = fidx >= 0 ? pol_last : pol_first;
            float wgt[NSLOT > 0 ? NSLOT : 1];
            int jo[NSLOT > 0 ? NSLOT : 1];
            uint32_t live = 0;                                       // slots with a non-zero weight somewhere in this warp
            if (NSLOT > 0) {
#pragma unroll
                for (int k = 0; k < NSLOT; ++k) {
                    wgt[k] = __ldg(sw + k * Vp + v);
                    jo[k] = __ldg(sj + k * Vp + v) * 12;
                }
#pragma unroll
                for (int k = 0; k < NSLOT; ++k) live |= (__any_sync(0xffffffffu, wgt[k] != 0.f) ? 1u : 0u) << k;
            }
            if (st != cur_st) { mbar_wait(bfull, sc & 1u); ++sc; cur_st = st; }
            mbar_wait(tfull0 + 8 * buf, (lt >> 1) & 1u);
            tcgen05_fence_after();
            const int rows = min(T2_NS, M - m0);
            const bool store_v = v < V;
#pragma unroll 1
            for (int chunk = 0; chunk < 2; ++chunk) {
                const int s0 = sgrp * 16 + chunk * 8;               // first sample (tile-local) of this chunk
                uint32_t px[8], py[8], pz[8];
                const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + buf * (3 * T2_NS) + (uint32_t)s0;
                tmem_ld8(ta, px);
                tmem_ld8(ta + T2_NS, py);
                tmem_ld8(ta + 2 * T2_NS, pz);
                tmem_ld_wait();
                float x[8], y[8], z[8], o[8][3];
#pragma unroll
                for (int s = 0; s < 8; ++s) {
                    x[s] = fmaf(__uint_as_float(px[s]), inv_scale, t0);
                    y[s] = fmaf(__uint_as_float(py[s]), inv_scale, t1);
                    z[s] = fmaf(__uint_as_float(pz[s]), inv_scale, t2);
                    o[s][0] = o[s][1] = o[s][2] = 0.f;
                }
                const float* ab = As + s0 * J12;
                if (NSLOT > 0) {
#pragma unroll
                    for (int k = 0; k < NSLOT; ++k) {
                        if (!((live >> k) & 1u)) continue;          // warp-uniform
                        const float w = wgt[k];
                        const float* a0 = ab + jo[k];
#pragma unroll
                        for (int s = 0; s < 8; ++s) {
                            const float4* a = reinterpret_cast<const float4*>(a0 + s * J12);
                            const float4 r0 = a[0], r1 = a[1], r2 = a[2];
                            o[s][0] = fmaf(w, fmaf(r0.x, x[s], fmaf(r0.y, y[s], fmaf(r0.z, z[s], r0.w))), o[s][0]);
                            o[s][1] = fmaf(w, fmaf(r1.x, x[s], fmaf(r1.y, y[s], fmaf(r1.z, z[s], r1.w))), o[s][1]);
                            o[s][2] = fmaf(w, fmaf(r2.x, x[s], fmaf(r2.y, y[s], fmaf(r2.z, z[s], r2.w))), o[s][2]);
                        }
                    }
                } else {
                    for (int slot = 0; slot < nslots; ++slot) {
                        const float w = __ldg(sw + slot * Vp + v);
                        const float* a0 = ab + __ldg(sj + slot * Vp + v) * 12;
#pragma unroll
                        for (int s = 0; s < 8; ++s) {
                            const float4* a = reinterpret_cast<const float4*>(a0 + s * J12);
                            const float4 r0 = a[0], r1 = a[1], r2 = a[2];
                            o[s][0] = fmaf(w, fmaf(r0.x, x[s], fmaf(r0.y, y[s], fmaf(r0.z, z[s], r0.w))), o[s][0]);
                            o[s][1] = fmaf(w, fmaf(r1.x, x[s], fmaf(r1.y, y[s], fmaf(r1.z, z[s], r1.w))), o[s][1]);
                            o[s][2] = fmaf(w, fmaf(r2.x, x[s], fmaf(r2.y, y[s], fmaf(r2.z, z[s], r2.w))), o[s][2]);
                        }
                    }
                }
                if (fidx >= 0) {      // ~4 % of the vertices: compact per-sample copy xv[m][slot][3] for the extra-joint kernel
#pragma unroll
                    for (int s = 0; s < 8; ++s)
                        if (s0 + s < rows) {
                            float* xo = xv + ((size_t)(m0 + s0 + s) * NF + fidx) * 3;
                            float tx = 0.f, ty = 0.f, tz = 0.f;
                            if (TRANSL) { const int mm = m0 + s0 + s; tx = __ldg(transl + mm * 3); ty = __ldg(transl + mm * 3 + 1); tz = __ldg(transl + mm * 3 + 2); }
                            xo[0] = o[s][0] + tx; xo[1] = o[s][1] + ty; xo[2] = o[s][2] + tz;
                        }
                }
                if (store_v) {
                    float* out = vertices + ((size_t)(m0 + s0) * V + v) * 3;
                    if (!TRANSL && s0 + 8 <= rows) {                // full chunk, no translation: straight-line stores
#pragma unroll
                        for (int s = 0; s < 8; ++s) {
                            st_hint(out + 0, o[s][0], pol); st_hint(out + 1, o[s][1], pol); st_hint(out + 2, o[s][2], pol);
                            out += vstride;
                        }
                    } else {
#pragma unroll
                        for (int s = 0; s < 8; ++s) {
                            const int m = m0 + s0 + s;
                            if (s0 + s < rows) {
                                if (TRANSL) { o[s][0] += __ldg(transl + m * 3); o[s][1] += __ldg(transl + m * 3 + 1); o[s][2] += __ldg(transl + m * 3 + 2); }
                                st_hint(out + 0, o[s][0], pol); st_hint(out + 1, o[s][1], pol); st_hint(out + 2, o[s][2], pol);
                            }
                            out += vstride;
                        }
                    }
                }
            }
            // accumulator buffer may be refilled; on the last unit of a sample tile the resident tile may be replaced
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(tempty0 + 8 * buf);
                if (u + 1 < u1 && (u + 1) / nvt != st) mbar_arrive(sdone);
            }
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (warp == 1) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}


// ------------------------------------------------------------------------------------------------
// Round-2 experiment (impl 3; measured slower than the product kernel above, kept as a cross-check):
// fp16 tensor-core blend + CUDA-core skinning with LANES = SAMPLES.
//   unit = 128 samples x 32 vertices.  D[s][(c, v)] = sum_k f[s][k] P_c[v][k]: the per-sample coefficient tile is the
//   RESIDENT A operand (4 k-blocks x [128 rows][128 B] = 64 KB per sample tile), the basis rows of the unit's 32 vertices
//   (3 coordinates -> N = 96) stream through a TMA ring as the B operand; accumulators double-buffered in TMEM (2 x 96
//   columns).  Epilogue warp (q, g): TMEM lane quarter q = 32 samples, 8 of the unit's 32 vertices.  Every vertex is
//   warp-uniform, so its weights / joints come from one 64-byte record (uniform loads) and the 3x4 transforms of its <= 4
//   joints live in REGISTERS, re-loaded (coalesced over the samples, At[tile][joint][12][128]) only when the joint of a
//   slot changes; vertices are pre-sorted inside each 32-tile by joint signature, so changes are rare.  Results are
//   transposed through a shared-memory tile (pitch 97, conflict-free) and leave as 128-byte row segments of
//   vertices[m][v][3]; vertices that feed a joint pick / extra regressor are also written to xvt[tile][slot][c][sample].
#ifndef T3_BRANCHY
#define T3_BRANCHY 1
#endif
constexpr int T3_MS = 128, T3_NV = 32, T3_N = 3 * T3_NV, T3_STAGES = 6, T3_STAGE_BYTES = T3_N * 128;
constexpr int T3_EPI_WARPS = 16, T3_THREADS = (2 + T3_EPI_WARPS) * 32;
constexpr int T3_KBLK = LBS_K2 / 64, T3_ARES_BYTES = T3_KBLK * T3_MS * 128;
constexpr int T3_PITCH = T3_N + 1, T3_STG_BYTES = 4 * 32 * T3_PITCH * 4 + T3_EPI_WARPS * 512 + 16;

__device__ __forceinline__ void tmem_ld4(uint32_t taddr, uint32_t* v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]) : "r"(taddr));
}

__device__ __forceinline__ void tmem_ld1(uint32_t taddr, uint32_t& v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v) : "r"(taddr));
}

template <bool TRANSL>
__global__ void __launch_bounds__(T3_THREADS, 1)
lbs_skin_tc3_kernel(const __grid_constant__ CUtensorMap mapP, const __grid_constant__ CUtensorMap mapF,
                    const float4* __restrict__ vconst, const float* __restrict__ At, const float* __restrict__ transl,
                    int M, int V, int Vp, int J, float inv_scale, int nvt, int num_units, int NF,
                    float* __restrict__ vertices, float* __restrict__ xvt, int dbg) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bars[2 * T3_STAGES + 6];
    __shared__ uint32_t tmem_base_s;
    const uint32_t tile_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t ares = tile_base + T3_STAGES * T3_STAGE_BYTES;            // resident A operand: 4 k-blocks [128 rows][128 B]
    float* stg = reinterpret_cast<float*>(smem_raw + (ares + T3_ARES_BYTES - smem_u32(smem_raw)));   // [4 quarters][32][T3_PITCH]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[T3_STAGES]);
    const uint32_t tfull0 = smem_u32(&bars[2 * T3_STAGES]), tempty0 = smem_u32(&bars[2 * T3_STAGES + 2]);
    const uint32_t afull = smem_u32(&bars[2 * T3_STAGES + 4]), sdone = smem_u32(&bars[2 * T3_STAGES + 5]);
    if (threadIdx.x == 0) {
        for (int s = 0; s < T3_STAGES; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(tfull0 + 8 * b, 1); mbar_init(tempty0 + 8 * b, T3_EPI_WARPS); }
        mbar_init(afull, 1);
        mbar_init(sdone, T3_EPI_WARPS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(256u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const uint32_t tmem_base = tmem_base_s;
    const int u0 = (int)((long long)blockIdx.x * num_units / gridDim.x);
    const int u1 = (int)((long long)(blockIdx.x + 1) * num_units / gridDim.x);

    if (warp == 0) {
        if (lane == 0) {
            uint32_t kbc = 0, sc = 0;
            int cur_st = -1;
            for (int u = u0; u < u1; ++u) {
                const int st = u / nvt, vt = u - st * nvt;
                const bool new_tile = st != cur_st;
                for (int kb = 0; kb < T3_KBLK; ++kb, ++kbc) {
                    const uint32_t sg = kbc % T3_STAGES, ph = (kbc / T3_STAGES) & 1u;
                    mbar_wait_long(empty0 + 8 * sg, ph ^ 1u);
                    const uint32_t fb = full0 + 8 * sg, dst = tile_base + sg * T3_STAGE_BYTES;
                    mbar_expect_tx(fb, T3_STAGE_BYTES);
#pragma unroll
                    for (int c = 0; c < 3; ++c) tma_load_2d(dst + c * (T3_NV * 128), &mapP, fb, kb * 64, c * Vp + vt * T3_NV);
                    // switch the resident sample tile once this unit's basis blocks are on their way: the old tile's
                    // coefficients are free when every epilogue warp has finished its last unit (its MMAs completed before)
                    if (new_tile && kb == T3_KBLK - 1) {
                        if (cur_st >= 0) mbar_wait_long(sdone, (sc - 1) & 1u);
                        mbar_expect_tx(afull, (uint32_t)T3_ARES_BYTES);
#pragma unroll
                        for (int k = 0; k < T3_KBLK; ++k) tma_load_2d(ares + k * (T3_MS * 128), &mapF, afull, k * 64, st * T3_MS);
                        cur_st = st; ++sc;
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_f16(T3_MS, T3_N);
            uint32_t kbc = 0, lt = 0, sc = 0;
            int cur_st = -1;
            for (int u = u0; u < u1; ++u, ++lt) {
                const int st = u / nvt;
                if (st != cur_st) { mbar_wait_long(afull, sc & 1u); ++sc; cur_st = st; }
                const uint32_t buf = lt & 1u;
                mbar_wait_long(tempty0 + 8 * buf, ((lt >> 1) & 1u) ^ 1u);
                tcgen05_fence_after();
                const uint32_t d = tmem_base + buf * T3_N;
                for (int kb = 0; kb < T3_KBLK; ++kb, ++kbc) {
                    const uint32_t sg = kbc % T3_STAGES, ph = (kbc / T3_STAGES) & 1u;
                    mbar_wait_long(full0 + 8 * sg, ph);
                    tcgen05_fence_after();
                    const uint64_t da = umma_desc_sw128(ares + kb * (T3_MS * 128)), db = umma_desc_sw128(tile_base + sg * T3_STAGE_BYTES);
#pragma unroll
                    for (int k = 0; k < 4; ++k) umma_f16(d, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (uint32_t)((kb | k) != 0));
                    umma_commit(empty0 + 8 * sg);
                }
                umma_commit(tfull0 + 8 * buf);
            }
        }
    } else {
        const int q = warp & 3, g = (warp - 2) >> 2;        // TMEM lane quarter (32 samples), group of 8 vertices
        const int sl = q * 32 + lane;                        // sample within the tile
        float* sq = stg + q * (32 * T3_PITCH);               // this quarter's staging rows [32][T3_PITCH]
        float4* rec = reinterpret_cast<float4*>(stg + 4 * 32 * T3_PITCH) + (warp - 2) * 32;   // this warp's 8 vertex records (512 B)
        const uint32_t bar_id = 1 + q;
        float T[4][12];
        float trx = 0.f, try_ = 0.f, trz = 0.f;
        uint32_t lt = 0, curpack = 0xffffffffu;               // joints (one byte per slot) of the transforms held in T
        int cur_st = -1;
        const float* Ab = At;
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
            for (int e = 0; e < 12; ++e) T[k][e] = 0.f;
        // records of the warp's 8 vertices of a unit = 512 contiguous bytes = one 16-byte load per lane, fetched one unit ahead
        float4 rnext = make_float4(0.f, 0.f, 0.f, 0.f);
        if (u0 < u1) rnext = __ldg(vconst + ((size_t)(u0 % nvt) * T3_NV + g * 8) * 4 + lane);
        for (int u = u0; u < u1; ++u, ++lt) {
            const int st = u / nvt, vt = u - st * nvt;
            const uint32_t buf = lt & 1u;
            const int m = st * T3_MS + sl;
            if (st != cur_st) {
                cur_st = st;
                Ab = At + ((size_t)st * J * 128 + sl) * 12;
                curpack = 0xffffffffu;                               // register-resident transforms belong to the old tile
                if (TRANSL && m < M) { trx = __ldg(transl + (size_t)m * 3); try_ = __ldg(transl + (size_t)m * 3 + 1); trz = __ldg(transl + (size_t)m * 3 + 2); }
            }
            rec[lane] = rnext;
            if (u + 1 < u1) rnext = __ldg(vconst + ((size_t)((u + 1) % nvt) * T3_NV + g * 8) * 4 + lane);
            __syncwarp();
            mbar_wait_long(tfull0 + 8 * buf, (lt >> 1) & 1u);
            tcgen05_fence_after();
            if (dbg & 1) {     // timing experiment (HF_LBS_DBG=1): skip the epilogue work, keep the protocol
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) { mbar_arrive(tempty0 + 8 * buf); if (u + 1 < u1 && (u + 1) / nvt != st) mbar_arrive(sdone); }
                continue;
            }
            const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + buf * T3_N + (uint32_t)(g * 8);
            // one vertex per iteration (rolled loop: the body is large and its control flow data dependent); the accumulator
            // columns of vertex i+1 are requested before vertex i is processed
            uint32_t nx, ny, nz;
            tmem_ld1(ta, nx); tmem_ld1(ta + T3_NV, ny); tmem_ld1(ta + 2 * T3_NV, nz);
            // software pipeline: record (shared memory, broadcast) + control words (shuffles) of vertex i+1 are fetched while
            // vertex i is processed, so the per-vertex dependency chain is only the FMA part
            float4 n0 = rec[0], n1 = rec[1], n3 = rec[3];
            uint32_t nctrl = __shfl_sync(0xffffffffu, __float_as_uint(n3.x), 0);
            uint32_t njpack = __shfl_sync(0xffffffffu, __float_as_uint(n3.z), 0);
            uint32_t nlmask = __shfl_sync(0xffffffffu, __float_as_uint(n3.w), 0);
#pragma unroll 1
            for (int i = 0; i < ((dbg & 8) ? 0 : 8); ++i) {
                tmem_ld_wait();
                const uint32_t px = nx, py = ny, pz = nz;
                const float4 c0 = n0, c1 = n1, c3 = n3;
                const uint32_t ctrl = nctrl, jpack = njpack, lmask = nlmask;
                const float4* rp = rec + i * 4;
                if (i < 7) {
                    tmem_ld1(ta + i + 1, nx); tmem_ld1(ta + T3_NV + i + 1, ny); tmem_ld1(ta + 2 * T3_NV + i + 1, nz);
                    n0 = rp[4]; n1 = rp[5]; n3 = rp[7];
                }
                if (dbg & 16) { if (px == 0x12345678u && py == pz) sq[lane] = 1.f; continue; }   // timing experiment: TMEM loads only
                const float x = fmaf(__uint_as_float(px), inv_scale, c0.x);
                const float y = fmaf(__uint_as_float(py), inv_scale, c0.y);
                const float z = fmaf(__uint_as_float(pz), inv_scale, c0.z);
                if ((jpack ^ curpack) & lmask) {                             // rare with sorted tiles: a live slot's joint run ended
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint32_t jb = (jpack >> (8 * k)) & 0xffu;
                        if (jb != 0xffu && jb != ((curpack >> (8 * k)) & 0xffu)) {
                            const float4* ap = reinterpret_cast<const float4*>(Ab + (size_t)jb * (128 * 12));
                            const float4 r0 = __ldg(ap), r1 = __ldg(ap + 1), r2 = __ldg(ap + 2);
                            T[k][0] = r0.x; T[k][1] = r0.y; T[k][2] = r0.z; T[k][3] = r0.w;
                            T[k][4] = r1.x; T[k][5] = r1.y; T[k][6] = r1.z; T[k][7] = r1.w;
                            T[k][8] = r2.x; T[k][9] = r2.y; T[k][10] = r2.z; T[k][11] = r2.w;
                            curpack = (curpack & ~(0xffu << (8 * k))) | (jb << (8 * k));
                        }
                    }
                }
#if T3_BRANCHY
                // skip dead slots (slots are compacted live-first, slot 0 is always live): ns = live slots - 1, warp-uniform
                const uint32_t ns = ctrl & 3u;
                float ox, oy, oz;
                ox = c1.x * fmaf(T[0][0], x, fmaf(T[0][1], y, fmaf(T[0][2], z, T[0][3])));
                oy = c1.x * fmaf(T[0][4], x, fmaf(T[0][5], y, fmaf(T[0][6], z, T[0][7])));
                oz = c1.x * fmaf(T[0][8], x, fmaf(T[0][9], y, fmaf(T[0][10], z, T[0][11])));
                if (ns >= 1u) {
                    ox = fmaf(c1.y, fmaf(T[1][0], x, fmaf(T[1][1], y, fmaf(T[1][2], z, T[1][3]))), ox);
                    oy = fmaf(c1.y, fmaf(T[1][4], x, fmaf(T[1][5], y, fmaf(T[1][6], z, T[1][7]))), oy);
                    oz = fmaf(c1.y, fmaf(T[1][8], x, fmaf(T[1][9], y, fmaf(T[1][10], z, T[1][11]))), oz);
                    if (ns >= 2u) {
                        ox = fmaf(c1.z, fmaf(T[2][0], x, fmaf(T[2][1], y, fmaf(T[2][2], z, T[2][3]))), ox);
                        oy = fmaf(c1.z, fmaf(T[2][4], x, fmaf(T[2][5], y, fmaf(T[2][6], z, T[2][7]))), oy);
                        oz = fmaf(c1.z, fmaf(T[2][8], x, fmaf(T[2][9], y, fmaf(T[2][10], z, T[2][11]))), oz);
                        if (ns >= 3u) {
                            ox = fmaf(c1.w, fmaf(T[3][0], x, fmaf(T[3][1], y, fmaf(T[3][2], z, T[3][3]))), ox);
                            oy = fmaf(c1.w, fmaf(T[3][4], x, fmaf(T[3][5], y, fmaf(T[3][6], z, T[3][7]))), oy);
                            oz = fmaf(c1.w, fmaf(T[3][8], x, fmaf(T[3][9], y, fmaf(T[3][10], z, T[3][11]))), oz);
                        }
                    }
                }
#else
                // all four slots, branch-free: dead slots carry weight 0 and whatever (finite) transform the slot held last, so
                // the 12 inner products are independent FMA chains the scheduler can interleave (the branchy form, which skips
                // dead slots, measured latency-bound at 4 warps per scheduler)
                const float wk[4] = {c1.x, c1.y, c1.z, c1.w};
                float ox = 0.f, oy = 0.f, oz = 0.f;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    ox = fmaf(wk[k], fmaf(T[k][0], x, fmaf(T[k][1], y, fmaf(T[k][2], z, T[k][3]))), ox);
                    oy = fmaf(wk[k], fmaf(T[k][4], x, fmaf(T[k][5], y, fmaf(T[k][6], z, T[k][7]))), oy);
                    oz = fmaf(wk[k], fmaf(T[k][8], x, fmaf(T[k][9], y, fmaf(T[k][10], z, T[k][11]))), oz);
                }
#endif
                if (TRANSL) { ox += trx; oy += try_; oz += trz; }
                float* so = sq + lane * T3_PITCH + __float_as_int(c0.w);     // c0.w = 3 * (original index inside the 32-tile)
                so[0] = ox; so[1] = oy; so[2] = oz;
                if ((ctrl & 0x100u) && m < M) {                    // vertex feeds a joint pick / regressor: sample-contiguous copy
                    float* xo = xvt + (((size_t)st * NF + __float_as_int(c3.y)) * 3) * 128 + sl;
                    xo[0] = ox; xo[128] = oy; xo[256] = oz;
                }
                nctrl = __shfl_sync(0xffffffffu, __float_as_uint(n3.x), 0);
                njpack = __shfl_sync(0xffffffffu, __float_as_uint(n3.z), 0);
                nlmask = __shfl_sync(0xffffffffu, __float_as_uint(n3.w), 0);
            }
            // accumulator fully read (the last vertex's columns were waited for inside the loop): hand the TMEM buffer back
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty0 + 8 * buf);
            // the quarter's 32 x 96 tile is complete when its four warps have written their 8 vertices each
            asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
            {
                const int ebase = vt * T3_N;                         // first float of this unit inside a sample's vertex row
                const int emax = V * 3;
                const bool full = ebase + T3_N <= emax;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int r = g * 8 + i;                         // sample row of this quarter written by this warp
                    const int mr = st * T3_MS + q * 32 + r;
                    if (mr < M && !(dbg & 4)) {
                        float* out = vertices + (size_t)mr * emax + ebase + lane;
                        const float* sr = sq + r * T3_PITCH + lane;
                        if (dbg & 2) { if (sr[0] + sr[32] + sr[64] == 123.456f) out[0] = 0.f; }
                        else if (full) { __stcs(out, sr[0]); __stcs(out + 32, sr[32]); __stcs(out + 64, sr[64]); }
                        else {
#pragma unroll
                            for (int c3i = 0; c3i < 3; ++c3i)
                                if (ebase + c3i * 32 + lane < emax) __stcs(out + c3i * 32, sr[c3i * 32]);
                        }
                    }
                }
            }
            asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");    // staging rows may be overwritten by the next unit
            if (lane == 0 && u + 1 < u1 && (u + 1) / nvt != st) mbar_arrive(sdone);
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (warp == 1) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256u) : "memory");
    }
}

// joints[J ..) from the sample-contiguous copies of the flagged vertices written by lbs_skin_tc3_kernel:
// xvt[tile][slot][c][128 samples].  Block = (sample tile, chunk of XT_ROWS output rows), thread = sample; every row is summed in
// CSR order (fixed order -> deterministic).
constexpr int XT_ROWS = 6;
__global__ void __launch_bounds__(128)
lbs_extra_joints_t_kernel(const float* __restrict__ xvt, const int* __restrict__ pick_f, const int* __restrict__ csr_ptr,
                          const int* __restrict__ csr_f, const float* __restrict__ csr_val, int M, int J, int nvj, int nextra,
                          int J_out, int NF, float* __restrict__ joints) {
    HF_PDL_SYNC();
    const int st = blockIdx.x, s = threadIdx.x, m = st * 128 + s;
    if (m >= M) return;
    const float* xb = xvt + (size_t)st * NF * 3 * 128 + s;
    const int r0 = blockIdx.y * XT_ROWS, r1 = min(r0 + XT_ROWS, nvj + nextra);
    for (int r = r0; r < r1; ++r) {
        float x = 0.f, y = 0.f, z = 0.f;
        if (r < nvj) {
            const float* p = xb + (size_t)pick_f[r] * 3 * 128;
            x = p[0]; y = p[128]; z = p[256];
        } else {
            const int row = r - nvj;
            for (int e = csr_ptr[row]; e < csr_ptr[row + 1]; ++e) {
                const float w = csr_val[e];
                const float* p = xb + (size_t)csr_f[e] * 3 * 128;
                x = fmaf(w, p[0], x); y = fmaf(w, p[128], y); z = fmaf(w, p[256], z);
            }
        }
        float* o = joints + ((size_t)m * J_out + J + r) * 3;
        o[0] = x; o[1] = y; o[2] = z;
    }
}

// joints[J ..) from the compact copies xv[m][slot][3] of the flagged vertices (written by the skinning epilogue): a warp per
// sample, a lane per output row (picks, then the CSR regressor rows summed in CSR order); a sample's 3.3 KB are contiguous,
// so there are no dependent gathers into the 265 MB vertex array.
__global__ void __launch_bounds__(128)
lbs_extra_joints_c_kernel(const float* __restrict__ xv, const int* __restrict__ pick_f, const int* __restrict__ csr_ptr,
                          const int* __restrict__ csr_f, const float* __restrict__ csr_val, int M, int J, int nvj, int nextra,
                          int J_out, int NF, float* __restrict__ joints) {
    HF_PDL_SYNC();
    const int m = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (m >= M) return;
    const float* xb = xv + (size_t)m * NF * 3;
    for (int r = lane; r < nvj + nextra; r += 32) {
        float x = 0.f, y = 0.f, z = 0.f;
        if (r < nvj) {
            const float* p = xb + (size_t)__ldg(pick_f + r) * 3;
            x = p[0]; y = p[1]; z = p[2];
        } else {
            const int row = r - nvj;
            for (int e = __ldg(csr_ptr + row); e < __ldg(csr_ptr + row + 1); ++e) {
                const float w = __ldg(csr_val + e);
                const float* p = xb + (size_t)__ldg(csr_f + e) * 3;
                x = fmaf(w, p[0], x); y = fmaf(w, p[1], y); z = fmaf(w, p[2], z);
            }
        }
        float* o = joints + ((size_t)m * J_out + J + r) * 3;
        o[0] = x; o[1] = y; o[2] = z;
    }
}

// joints[J .. J+nvj) = picked vertices; joints[J+nvj ..) = sparse regressors applied to the final vertices.
// One block per sample: every (row, vertex) entry of the picks + CSR regressors is fetched by its own thread (all
// gathers of a sample in flight at once), then one thread per output row sums its entries in CSR order.
constexpr int XJ_THREADS = 256, XJ_MAXE = 1024;
__global__ void __launch_bounds__(XJ_THREADS)
lbs_extra_joints_kernel(const float* __restrict__ vertices, const int* __restrict__ vj,
                                        const int* __restrict__ csr_ptr, const int* __restrict__ csr_col,
                                        const float* __restrict__ csr_val, int M, int V, int J, int nvj,
                                        int nextra, int J_out, float* __restrict__ joints) {
    __shared__ float ps[XJ_MAXE][3];
    const int m = blockIdx.x;
    const int nnz = csr_ptr[nextra], E = nvj + nnz;
    HF_PDL_SYNC();
    const float* vm = vertices + (size_t)m * V * 3;
    for (int e = threadIdx.x; e < E; e += XJ_THREADS) {
        const int col = e < nvj ? vj[e] : csr_col[e - nvj];
        const float* p = vm + (size_t)col * 3;
        ps[e][0] = __ldg(p); ps[e][1] = __ldg(p + 1); ps[e][2] = __ldg(p + 2);
    }
    __syncthreads();
    for (int r = threadIdx.x; r < nvj + nextra; r += XJ_THREADS) {
        float x = 0.f, y = 0.f, z = 0.f;
        if (r < nvj) { x = ps[r][0]; y = ps[r][1]; z = ps[r][2]; }
        else {
            const int row = r - nvj;
            for (int e = csr_ptr[row]; e < csr_ptr[row + 1]; ++e) {
                const float w = csr_val[e];
                x = fmaf(w, ps[nvj + e][0], x); y = fmaf(w, ps[nvj + e][1], y); z = fmaf(w, ps[nvj + e][2], z);
            }
        }
        float* o = joints + ((size_t)m * J_out + J + r) * 3;
        o[0] = x; o[1] = y; o[2] = z;
    }
}

// Fallback for regressors with more than XJ_MAXE - nvj non-zeros: one thread per output row.
__global__ void lbs_extra_joints_rows_kernel(const float* __restrict__ vertices, const int* __restrict__ vj,
                                        const int* __restrict__ csr_ptr, const int* __restrict__ csr_col,
                                        const float* __restrict__ csr_val, int M, int V, int J, int nvj,
                                        int nextra, int J_out, float* __restrict__ joints) {
    HF_PDL_SYNC();
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    int per = nvj + nextra;
    if (idx >= M * per) return;
    int m = idx / per, r = idx - m * per;
    const float* vm = vertices + (size_t)m * V * 3;
    float x = 0.f, y = 0.f, z = 0.f;
    if (r < nvj) {
        const float* p = vm + (size_t)vj[r] * 3;
        x = p[0]; y = p[1]; z = p[2];
    } else {
        int row = r - nvj;
        for (int e = csr_ptr[row]; e < csr_ptr[row + 1]; ++e) {
            const float* p = vm + (size_t)csr_col[e] * 3;
            float w = csr_val[e];
            x = fmaf(w, p[0], x); y = fmaf(w, p[1], y); z = fmaf(w, p[2], z);
        }
    }
    float* o = joints + ((size_t)m * J_out + J + r) * 3;
    o[0] = x; o[1] = y; o[2] = z;
}

__global__ void rodrigues_kernel(const float* __restrict__ aa, float* __restrict__ R, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    // smplx batch_rodrigues: angle = ||v + 1e-8||, axis = v / angle, R = I + sin K + (1-cos) K^2 (fp32)
    float x = aa[i * 3], y = aa[i * 3 + 1], z = aa[i * 3 + 2];
    float ex = x + 1e-8f, ey = y + 1e-8f, ez = z + 1e-8f;
    float angle = sqrtf(ex * ex + ey * ey + ez * ez);
    float rx = x / angle, ry = y / angle, rz = z / angle;
    float s = sinf(angle), c1 = 1.f - cosf(angle);
    float* o = R + (size_t)i * 9;
    // K = [[0,-rz,ry],[rz,0,-rx],[-ry,rx,0]];  K^2 = r r^T - |r|^2 I
    float xx = rx * rx, yy = ry * ry, zz = rz * rz, xy = rx * ry, xz = rx * rz, yz = ry * rz;
    o[0] = 1.f + c1 * (-(yy + zz)); o[1] = -s * rz + c1 * xy;        o[2] = s * ry + c1 * xz;
    o[3] = s * rz + c1 * xy;        o[4] = 1.f + c1 * (-(xx + zz)); o[5] = -s * rx + c1 * yz;
    o[6] = -s * ry + c1 * xz;       o[7] = s * rx + c1 * yz;        o[8] = 1.f + c1 * (-(xx + yy));
}

constexpr int kSPT = 16, kSG = 4;

inline uint16_t f2bf(float f) {   // round-to-nearest-even float -> bf16 bits
    uint32_t u; memcpy(&u, &f, 4);
    if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40);
    u += 0x7fffu + ((u >> 16) & 1u);
    return (uint16_t)(u >> 16);
}
inline float bf2f(uint16_t b) { uint32_t u = (uint32_t)b << 16; float f; memcpy(&f, &u, 4); return f; }

}  // namespace

extern "C" int hf_smpl_create(hf_smpl_t** out, int V, int nb, int J, const float* v_template,
                              const float* shapedirs, const float* posedirs, const float* J_regressor,
                              const float* lbs_weights, const int* parents, const int* vertex_joint_ids,
                              int nvj, const float* extra_regressors, int nextra) {
    if (!out || V <= 0 || nb <= 0 || nb > HF_MAXB || J <= 1 || J > HF_MAXJ)
        return hf::fail(HF_ERR_INVALID, "hf_smpl_create: bad sizes V=%d nb=%d J=%d", V, nb, J);
    if (parents[0] != -1) return hf::fail(HF_ERR_INVALID, "hf_smpl_create: parents[0] must be -1");
    for (int i = 1; i < J; ++i)
        if (parents[i] < 0 || parents[i] >= i)
            return hf::fail(HF_ERR_INVALID, "hf_smpl_create: parents[%d]=%d is not an earlier joint", i, parents[i]);
    hf_smpl* h = new hf_smpl();
    h->V = V; h->nb = nb; h->J = J; h->nvj = nvj; h->nextra = nextra;
    h->Vp = hf::div_up(V, 128) * 128;
    h->KB = nb + 9 * (J - 1);
    h->KP = hf::div_up(h->KB, 4) * 4;
    for (int i = 0; i < J; ++i) h->parents[i] = parents[i];
    const int Vp = h->Vp, KB = h->KB;
    std::vector<float> blend((size_t)KB * 3 * Vp, 0.f), vt((size_t)3 * Vp, 0.f);
    for (int v = 0; v < V; ++v)
        for (int c = 0; c < 3; ++c) {
            vt[(size_t)c * Vp + v] = v_template[v * 3 + c];
            for (int l = 0; l < nb; ++l)
                blend[((size_t)l * 3 + c) * Vp + v] = shapedirs[((size_t)v * 3 + c) * nb + l];
            for (int k = 0; k < 9 * (J - 1); ++k)
                blend[((size_t)(nb + k) * 3 + c) * Vp + v] = posedirs[(size_t)k * V * 3 + v * 3 + c];
        }
    // joint regressor folded through the shape space (fp64 accumulation)
    std::vector<float> J0((size_t)J * 3), Jd((size_t)J * 3 * nb);
    for (int i = 0; i < J; ++i)
        for (int c = 0; c < 3; ++c) {
            double a = 0.0;
            std::vector<double> d(nb, 0.0);
            for (int v = 0; v < V; ++v) {
                double w = J_regressor[(size_t)i * V + v];
                if (w == 0.0) continue;
                a += w * v_template[v * 3 + c];
                for (int l = 0; l < nb; ++l) d[l] += w * shapedirs[((size_t)v * 3 + c) * nb + l];
            }
            J0[i * 3 + c] = (float)a;
            for (int l = 0; l < nb; ++l) Jd[(i * 3 + c) * nb + l] = (float)d[l];
        }
    // sparse skinning weights
    int nslots = 1;
    for (int v = 0; v < V; ++v) {
        int n = 0;
        for (int j = 0; j < J; ++j) n += lbs_weights[(size_t)v * J + j] != 0.f;
        if (n > nslots) nslots = n;
    }
    h->nslots = nslots;
    std::vector<int> sj((size_t)nslots * Vp, 0);
    std::vector<float> sw((size_t)nslots * Vp, 0.f);
    for (int v = 0; v < V; ++v) {
        int n = 0;
        for (int j = 0; j < J; ++j) {
            float w = lbs_weights[(size_t)v * J + j];
            if (w != 0.f) { sj[(size_t)n * Vp + v] = j; sw[(size_t)n * Vp + v] = w; ++n; }
        }
    }
    // extra joint regressors -> CSR
    std::vector<int> ptr(nextra + 1, 0), col;
    std::vector<float> val;
    for (int r = 0; r < nextra; ++r) {
        for (int v = 0; v < V; ++v) {
            float w = extra_regressors[(size_t)r * V + v];
            if (w != 0.f) { col.push_back(v); val.push_back(w); }
        }
        ptr[r + 1] = (int)col.size();
    }
    h->nnz = (int)col.size();
    if (col.empty()) { col.push_back(0); val.push_back(0.f); }
    for (int r = 0; r < nvj; ++r)
        if (vertex_joint_ids[r] < 0 || vertex_joint_ids[r] >= V) {
            delete h;
            return hf::fail(HF_ERR_INVALID, "hf_smpl_create: vertex_joint_ids[%d] out of range", r);
        }
    int rc;
    if (KB > LBS_KH) { delete h; return hf::fail(HF_ERR_UNSUPPORTED, "hf_smpl_create: %d blend coefficients exceed %d", KB, LBS_KH); }
    {   // split-bf16 basis for the tensor-core blend: row (c, v) = [hi(256) | lo(256)]
        std::vector<uint16_t> pb((size_t)3 * Vp * 2 * LBS_KH, 0);
        for (int c = 0; c < 3; ++c)
            for (int v = 0; v < V; ++v) {
                uint16_t* row = &pb[((size_t)c * Vp + v) * 2 * LBS_KH];
                for (int k = 0; k < KB; ++k) {
                    const float f = blend[((size_t)k * 3 + c) * Vp + v];
                    const uint16_t hi = f2bf(f);
                    row[k] = hi;
                    row[LBS_KH + k] = f2bf(f - bf2f(hi));
                }
            }
        if ((rc = hf::upload((uint16_t**)&h->Pbf, pb.data(), pb.size()))) return rc;
        const uint64_t dims[2] = {(uint64_t)2 * LBS_KH, (uint64_t)3 * Vp};
        const uint64_t st[1] = {(uint64_t)2 * LBS_KH * 2};
        const uint32_t box[2] = {64, 128};
        if ((rc = encode_map(&h->mapA, h->Pbf, 2, dims, st, box))) return rc;
        h->impl = 0; h->mapB_ptr = nullptr; h->mapB_M = 0;
    }
    std::vector<__half> pf;
    {   // fp16 basis for the product path: row (c, v) = [posedirs (208) | shape hi (16) | shape hi (16) | shape lo (16)] * 2^k
        if (9 * (J - 1) > LBS_K2_POSE || nb > 16) { delete h; return hf::fail(HF_ERR_UNSUPPORTED, "hf_smpl_create: J=%d nb=%d exceed the fp16 blend layout", J, nb); }
        float mx = 0.f;
        for (size_t i = 0; i < blend.size(); ++i) mx = std::max(mx, std::fabs(blend[i]));
        int ex = 0;
        if (mx > 0.f) { std::frexp(mx, &ex); }            // mx = m * 2^ex, 0.5 <= m < 1
        const float scale = std::ldexp(1.f, 10 - ex);       // max |entry| * scale in [512, 1024): far from fp16 overflow and subnormals
        h->inv_scale = 1.f / scale;
        pf.assign((size_t)3 * Vp * LBS_K2, __float2half(0.f));
        for (int c = 0; c < 3; ++c)
            for (int v = 0; v < V; ++v) {
                __half* row = &pf[((size_t)c * Vp + v) * LBS_K2];
                for (int k = 0; k < 9 * (J - 1); ++k) row[k] = __float2half_rn(blend[((size_t)(nb + k) * 3 + c) * Vp + v] * scale);
                for (int l = 0; l < nb; ++l) {
                    const float f = blend[((size_t)l * 3 + c) * Vp + v] * scale;
                    const __half hi = __float2half_rn(f);
                    row[LBS_K2_POSE + l] = hi;
                    row[LBS_K2_POSE + 16 + l] = hi;
                    row[LBS_K2_POSE + 32 + l] = __float2half_rn(f - __half2float(hi));
                }
            }
        if ((rc = hf::upload(&h->Pf16, pf.data(), pf.size()))) return rc;
        const uint64_t dims[2] = {(uint64_t)LBS_K2, (uint64_t)3 * Vp};
        const uint64_t st[1] = {(uint64_t)LBS_K2 * 2};
        const uint32_t box[2] = {64, 128};
        if ((rc = encode_map(&h->mapA2, h->Pf16, 2, dims, st, box))) return rc;
        h->mapB2_ptr = nullptr; h->mapB2_M = 0;
    }
    if ((rc = hf::upload(&h->blend, blend.data(), blend.size()))) return rc;
    if ((rc = hf::upload(&h->vtemp, vt.data(), vt.size()))) return rc;
    if ((rc = hf::upload(&h->J0, J0.data(), J0.size()))) return rc;
    if ((rc = hf::upload(&h->Jd, Jd.data(), Jd.size()))) return rc;
    if ((rc = hf::upload(&h->sj, sj.data(), sj.size()))) return rc;
    if ((rc = hf::upload(&h->sw, sw.data(), sw.size()))) return rc;
    std::vector<int> vj(vertex_joint_ids, vertex_joint_ids + nvj);
    if (vj.empty()) vj.push_back(0);
    if ((rc = hf::upload(&h->vj, vj.data(), vj.size()))) return rc;
    {
        // vflag[v] = slot of vertex v in the compact "flagged vertices" copy (vertices read by a joint pick / regressor), else -1
        std::vector<int> vflag((size_t)Vp, 0);
        for (int r = 0; r < nvj; ++r) vflag[vj[r]] = 1;
        for (int e = 0; e < h->nnz; ++e) vflag[col[e]] = 1;
        int nf = 0;
        for (int v = 0; v < Vp; ++v) vflag[v] = (v < V && vflag[v]) ? nf++ : -1;
        if ((rc = hf::upload(&h->vflag, vflag.data(), vflag.size()))) return rc;
    }
    {   // round-2 experiment (impl 3): per-tile sorted vertex order, basis rows in that order, 64-byte per-vertex records, flagged slots
        const int nvt3 = Vp / T3_NV;
        auto key = [&](int v, int k) -> int { return (v < V && k < nslots && sw[(size_t)k * Vp + v] != 0.f) ? sj[(size_t)k * Vp + v] : 999; };
        std::vector<int> order((size_t)Vp);
        for (int t = 0; t < nvt3; ++t) {
            int idx[T3_NV];
            for (int i = 0; i < T3_NV; ++i) idx[i] = t * T3_NV + i;
            std::stable_sort(idx, idx + T3_NV, [&](int a, int b) {
                if ((a < V) != (b < V)) return a < V;                 // padding vertices last
                for (int k = 0; k < 4; ++k) { const int ka = key(a, k), kb = key(b, k); if (ka != kb) return ka < kb; }
                return false;
            });
            for (int i = 0; i < T3_NV; ++i) order[(size_t)t * T3_NV + i] = idx[i];
        }
        std::vector<int> fmap((size_t)Vp, -1), vflag_h((size_t)Vp, 0);
        for (int r = 0; r < nvj; ++r) vflag_h[vj[r]] = 1;
        for (int e = 0; e < h->nnz; ++e) vflag_h[col[e]] = 1;
        int NF = 0;
        for (int v = 0; v < V; ++v) if (vflag_h[v]) fmap[v] = NF++;
        h->NF = std::max(NF, 1);
        std::vector<int> pick_f(std::max(nvj, 1), 0), csr_f(col.size(), 0);
        for (int r = 0; r < nvj; ++r) pick_f[r] = fmap[vj[r]];
        for (int e = 0; e < h->nnz; ++e) csr_f[e] = fmap[col[e]];
        if ((rc = hf::upload(&h->pick_f, pick_f.data(), pick_f.size()))) return rc;
        if ((rc = hf::upload(&h->csr_f, csr_f.data(), csr_f.size()))) return rc;
        std::vector<float> vc((size_t)Vp * 16, 0.f);
        std::vector<__half> ps((size_t)3 * Vp * LBS_K2, __float2half(0.f));
        for (int p2 = 0; p2 < Vp; ++p2) {
            const int v = order[p2];
            float* rec = &vc[(size_t)p2 * 16];
            int li = v % T3_NV, jj[4] = {0, 0, 0, 0}, f = -1;
            float ww[4] = {0.f, 0.f, 0.f, 0.f};
            if (v < V) {
                for (int c = 0; c < 3; ++c) rec[c] = vt[(size_t)c * Vp + v];
                for (int k = 0; k < std::min(nslots, 4); ++k) { ww[k] = sw[(size_t)k * Vp + v]; jj[k] = sj[(size_t)k * Vp + v]; }
                f = fmap[v];
                for (int c = 0; c < 3; ++c)
                    memcpy(&ps[((size_t)c * Vp + p2) * LBS_K2], &pf[((size_t)c * Vp + v) * LBS_K2], LBS_K2 * sizeof(__half));
            }
            const int li3 = li * 3;
            memcpy(&rec[3], &li3, 4);
            for (int k = 0; k < 4; ++k) { rec[4 + k] = ww[k]; memcpy(&rec[8 + k], &jj[k], 4); }
            // control word: live slots - 1 | flagged << 8; jpack: joint of every live slot, one byte each (0xff = dead slot)
            int live = 0;
            for (int k = 0; k < 4; ++k) live += ww[k] != 0.f;
            unsigned ctrl = (unsigned)std::max(live - 1, 0), jpack = 0;
            for (int k = 0; k < 4; ++k) jpack |= (unsigned)((v < V && k < live) ? jj[k] : 0xff) << (8 * k);
            if (f >= 0) ctrl |= 0x100u;
            memcpy(&rec[12], &ctrl, 4);
            const int fi = std::max(f, 0);
            memcpy(&rec[13], &fi, 4);
            memcpy(&rec[14], &jpack, 4);
            unsigned lmask = 0;
            for (int k = 0; k < 4; ++k) if (v < V && k < live) lmask |= 0xffu << (8 * k);
            memcpy(&rec[15], &lmask, 4);
        }
        if ((rc = hf::upload((float**)&h->vconst, vc.data(), vc.size()))) return rc;
        if ((rc = hf::upload(&h->Pf16s, ps.data(), ps.size()))) return rc;
        const uint64_t dims[2] = {(uint64_t)LBS_K2, (uint64_t)3 * Vp};
        const uint64_t st[1] = {(uint64_t)LBS_K2 * 2};
        const uint32_t box[2] = {64, (uint32_t)T3_NV};
        if ((rc = encode_map(&h->mapP3, h->Pf16s, 2, dims, st, box))) return rc;
        h->mapF3_ptr = nullptr; h->mapF3_M = 0;
    }
    {   // per-vertex transpose of the output-joint rows that read vertices: picks (weight 1) then the regressor rows
        std::vector<std::vector<std::pair<int, float>>> byv((size_t)V);
        for (int r = 0; r < nvj; ++r) byv[vertex_joint_ids[r]].push_back({r, 1.f});
        for (int r = 0; r < nextra; ++r)
            for (int e = ptr[r]; e < ptr[r + 1]; ++e) byv[col[e]].push_back({nvj + r, val[e]});
        std::vector<int> cptr((size_t)Vp + 1, 0), crow;
        std::vector<float> cval;
        for (int v = 0; v < Vp; ++v) {
            if (v < V) for (auto& pr : byv[v]) { crow.push_back(pr.first); cval.push_back(pr.second); }
            cptr[v + 1] = (int)crow.size();
        }
        if (crow.empty()) { crow.push_back(0); cval.push_back(0.f); }
        if ((rc = hf::upload(&h->csc_ptr, cptr.data(), cptr.size()))) return rc;
        if ((rc = hf::upload(&h->csc_row, crow.data(), crow.size()))) return rc;
        if ((rc = hf::upload(&h->csc_val, cval.data(), cval.size()))) return rc;
        h->blend_split = nullptr; h->mapG_ptr = nullptr; h->mapG_M = 0;
        std::vector<int> fvert;
        for (int v = 0; v < V; ++v) if (!byv[v].empty()) fvert.push_back(v);
        // ordered by skinning joints so that the lanes of a warp of the backward vertex pass share their joints
        std::stable_sort(fvert.begin(), fvert.end(), [&](int a, int b) {
            for (int k = 0; k < nslots; ++k) {
                const int ja = sw[(size_t)k * Vp + a] != 0.f ? sj[(size_t)k * Vp + a] : 999, jb = sw[(size_t)k * Vp + b] != 0.f ? sj[(size_t)k * Vp + b] : 999;
                if (ja != jb) return ja < jb;
            }
            return false;
        });
        h->NFp = std::max(128, hf::div_up((int)fvert.size(), 128) * 128);
        fvert.resize((size_t)h->NFp, -1);
        if ((rc = hf::upload(&h->fvert, fvert.data(), fvert.size()))) return rc;
        h->blend_split_c = nullptr; h->blend_c = nullptr; h->mapGc_ptr = nullptr; h->mapGc_M = 0;
    }
    if ((rc = hf::upload(&h->csr_ptr, ptr.data(), ptr.size()))) return rc;
    if ((rc = hf::upload(&h->csr_col, col.data(), col.size()))) return rc;
    if ((rc = hf::upload(&h->csr_val, val.data(), val.size()))) return rc;
    *out = h;
    return HF_OK;
}

extern "C" void hf_smpl_destroy(hf_smpl_t* h) {
    if (!h) return;
    cudaFree(h->blend); cudaFree(h->vtemp); cudaFree(h->J0); cudaFree(h->Jd); cudaFree(h->sj);
    cudaFree(h->sw); cudaFree(h->Pbf); cudaFree(h->Pf16); cudaFree(h->Pf16s); cudaFree(h->vconst); cudaFree(h->pick_f); cudaFree(h->csr_f); cudaFree(h->vflag); cudaFree(h->vj); cudaFree(h->csr_ptr); cudaFree(h->csr_col); cudaFree(h->csr_val);
    cudaFree(h->csc_ptr); cudaFree(h->csc_row); cudaFree(h->csc_val); cudaFree(h->blend_split); cudaFree(h->fvert); cudaFree(h->blend_split_c); cudaFree(h->blend_c);
    delete h;
}

extern "C" int hf_smpl_num_joints_out(const hf_smpl_t* h) { return h->J + h->nvj + h->nextra; }

// workspace layout: F [M][KP] f32 | A [M][J][12] f32 | coefficient rows (fp16 [M][256] or split bf16 [M][512]), 256-aligned |
// At [tiles][J][12][128] f32 | xvt [tiles][NF][3][128] f32   (tiles = ceil(M / 128); the last two feed lbs_skin_tc3_kernel)
static size_t lbs_ws_off_at(const hf_smpl* h, int M) {
    return (((size_t)M * (h->KP + h->J * 12) * sizeof(float) + 255) & ~(size_t)255) + (((size_t)M * 2 * LBS_KH * 2 + 255) & ~(size_t)255);
}
extern "C" size_t hf_lbs_workspace_bytes(const hf_smpl_t* h, int M) {
    const size_t tiles = (size_t)hf::div_up(M, T3_MS);
    return lbs_ws_off_at(h, M) + tiles * h->J * 12 * 128 * sizeof(float) + tiles * h->NF * 3 * 128 * sizeof(float) + 512;
}

extern "C" int hf_lbs_set_impl(hf_smpl_t* h, int impl) {
    if (!h || impl < 0 || impl > 3) return hf::fail(HF_ERR_INVALID, "hf_lbs_set_impl: bad argument");
    if (impl == 3 && h->nslots > 4) return hf::fail(HF_ERR_UNSUPPORTED, "hf_lbs_set_impl: the lanes-as-samples kernel holds 4 skinning slots in registers, model has %d", h->nslots);
    h->impl = impl;
    return HF_OK;
}

// joints[J ..) from finished vertices: 21 vertex picks + the sparse extra regressors (shared with hf_lbs_tpose)
int hf_lbs_extra_joints(const hf_smpl* h, const float* vertices, float* joints, int M, cudaStream_t stream) {
    const int per = h->nvj + h->nextra, J_out = hf_smpl_num_joints_out(h);
    if (per <= 0) return HF_OK;
    if (h->nvj + h->nnz <= XJ_MAXE)
        HF_CUDA(hf::launch_pdl(lbs_extra_joints_kernel, dim3(M), dim3(XJ_THREADS), 0, stream, vertices, h->vj, h->csr_ptr,
                               h->csr_col, h->csr_val, M, h->V, h->J, h->nvj, h->nextra, J_out, joints));
    else
        HF_CUDA(hf::launch_pdl(lbs_extra_joints_rows_kernel, dim3(hf::div_up(M * per, 256)), dim3(256), 0, stream, vertices, h->vj, h->csr_ptr,
                               h->csr_col, h->csr_val, M, h->V, h->J, h->nvj, h->nextra, J_out, joints));
    HF_LAUNCH_CHECK();
    return HF_OK;
}

extern "C" int hf_smpl_dims(const hf_smpl* h, int* V, int* Vp, int* nb, int* J, int* J_out) {
    if (!h) return hf::fail(HF_ERR_INVALID, "hf_smpl_dims: null handle");
    if (V) *V = h->V; if (Vp) *Vp = h->Vp; if (nb) *nb = h->nb; if (J) *J = h->J; if (J_out) *J_out = hf_smpl_num_joints_out(h);
    return HF_OK;
}

int hf_smpl_tpose_tables(const hf_smpl* h, const float** blend, const float** vtemp, const float** J0, const float** Jd) {
    *blend = h->blend; *vtemp = h->vtemp; *J0 = h->J0; *Jd = h->Jd;
    return HF_OK;
}

static int lbs_forward_any(const hf_smpl_t* h, const float* betas, const float* rotmats, const float* glob, int rep,
                           const float* transl, float* vertices, float* joints, void* workspace,
                           size_t workspace_bytes, int M, void* stream_) {
    if (!h || !betas || !rotmats || !vertices || !joints) return hf::fail(HF_ERR_INVALID, "hf_lbs_forward: null argument");
    if (glob && (rep < 1 || M % rep)) return hf::fail(HF_ERR_INVALID, "hf_lbs_forward_split: M = %d is not a multiple of rep = %d", M, rep);
    if (M <= 0) return HF_OK;
    if (!workspace || workspace_bytes < hf_lbs_workspace_bytes(h, M))
        return hf::fail(HF_ERR_INVALID, "hf_lbs_forward: workspace too small (%zu < %zu)", workspace_bytes,
                        hf_lbs_workspace_bytes(h, M));
    cudaStream_t stream = (cudaStream_t)stream_;
    uint8_t* wsb = (uint8_t*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    float* F = (float*)wsb;
    float* A = F + (size_t)M * h->KP;
    __nv_bfloat16* Fb = (__nv_bfloat16*)(wsb + (((size_t)M * (h->KP + h->J * 12) * sizeof(float) + 255) & ~(size_t)255));
    const int tiles3 = hf::div_up(M, T3_MS);
    float* At = (float*)(wsb + lbs_ws_off_at(h, M));
    float* xvt = At + (size_t)tiles3 * h->J * 12 * 128;
    const int J_out = hf_smpl_num_joints_out(h);
    Parents par;
    for (int i = 0; i < HF_MAXJ; ++i) par.p[i] = i < h->J ? h->parents[i] : 0;
    static const int stage_mask = getenv("HF_LBS_STAGES") ? atoi(getenv("HF_LBS_STAGES")) : 7;   // profiling aid: bit 0 pose, 1 skin, 2 extra joints
    if (stage_mask & 1)
    HF_CUDA(hf::launch_pdl(lbs_pose_kernel, dim3(hf::div_up(M, PS)), dim3(PTHREADS), 0, stream, betas, rotmats, glob, rep, transl, h->J0, h->Jd, par, M,
                           h->J, h->nb, h->KP, J_out, h->impl != 1 ? (float*)nullptr : F, (h->impl == 0 || h->impl == 3) ? (__half*)Fb : (__half*)nullptr, A,
                           h->impl == 3 ? At : (float*)nullptr, joints));
    HF_LAUNCH_CHECK();
    if (!(stage_mask & 2)) {
    } else if (h->impl == 3) {
        __half* Fh = (__half*)Fb;
        hf_smpl* hm = const_cast<hf_smpl*>(h);
        if (hm->mapF3_ptr != (const void*)Fh || hm->mapF3_M != M) {
            const uint64_t dims[2] = {(uint64_t)LBS_K2, (uint64_t)M};
            const uint64_t st[1] = {(uint64_t)LBS_K2 * 2};
            const uint32_t box[2] = {64, (uint32_t)T3_MS};
            int rc = encode_map(&hm->mapF3, Fh, 2, dims, st, box);
            if (rc) return rc;
            hm->mapF3_ptr = Fh; hm->mapF3_M = M;
        }
        const size_t t3_smem = (size_t)T3_STAGES * T3_STAGE_BYTES + T3_ARES_BYTES + T3_STG_BYTES + 1024;
        const int nvt = h->Vp / T3_NV, num_units = nvt * tiles3;
        int sms = 148, dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        auto launch = [&](auto kern) -> int {
            HF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)t3_smem));
            HF_CUDA(hf::launch_pdl(kern, dim3(std::min(num_units, sms)), dim3(T3_THREADS), t3_smem, stream, hm->mapP3, hm->mapF3, (const float4*)h->vconst,
                                   (const float*)At, transl, M, h->V, h->Vp, h->J, h->inv_scale, nvt, num_units, h->NF, vertices, xvt, getenv("HF_LBS_DBG") ? atoi(getenv("HF_LBS_DBG")) : 0));
            return HF_OK;
        };
        int rc = transl ? launch(lbs_skin_tc3_kernel<true>) : launch(lbs_skin_tc3_kernel<false>);
        if (rc) return rc;
        HF_LAUNCH_CHECK();
        if ((stage_mask & 4) && h->nvj + h->nextra > 0) {
            HF_CUDA(hf::launch_pdl(lbs_extra_joints_t_kernel, dim3(tiles3, hf::div_up(h->nvj + h->nextra, XT_ROWS)), dim3(128), 0, stream, (const float*)xvt,
                                   (const int*)h->pick_f, (const int*)h->csr_ptr, (const int*)h->csr_f, (const float*)h->csr_val, M, h->J, h->nvj, h->nextra,
                                   J_out, h->NF, joints));
            HF_LAUNCH_CHECK();
        }
        return HF_OK;
    } else if (h->impl == 0) {
        __half* Fh = (__half*)Fb;
        hf_smpl* hm = const_cast<hf_smpl*>(h);
        if (hm->mapB2_ptr != (const void*)Fh || hm->mapB2_M != M) {
            const uint64_t dims[2] = {(uint64_t)LBS_K2, (uint64_t)M};
            const uint64_t st[1] = {(uint64_t)LBS_K2 * 2};
            const uint32_t box[2] = {64, (uint32_t)T2_NS};
            int rc = encode_map(&hm->mapB2, Fh, 2, dims, st, box);
            if (rc) return rc;
            hm->mapB2_ptr = Fh; hm->mapB2_M = M;
        }
        const size_t t2_smem = (size_t)T2_STAGES * 16384 + T2_BRES_BYTES + (size_t)T2_NS * h->J * 12 * sizeof(float) + 1024;
        if (t2_smem > 226 * 1024) return hf::fail(HF_ERR_UNSUPPORTED, "hf_lbs_forward: tile needs %zu B of shared memory", t2_smem);
        const int nvt = h->Vp / 128, num_units = nvt * hf::div_up(M, T2_NS);
        int sms = 148, dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        auto launch = [&](auto kern) -> int {
            HF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
            HF_CUDA(hf::launch_pdl(kern, dim3(std::min(num_units, sms)), dim3(T2_THREADS), t2_smem, stream, hm->mapA2, hm->mapB2,
                                   h->vtemp, h->sj, h->sw, A, transl, M, h->V, h->Vp, h->J, h->nslots, h->inv_scale, nvt, num_units, vertices, h->vflag,
                                   xvt, h->NF));
            return HF_OK;
        };
        int rc;
        const bool tr = transl != nullptr;
        switch (h->nslots <= 4 ? h->nslots : 0) {
            case 1: rc = tr ? launch(lbs_skin_tc2_kernel<1, true>) : launch(lbs_skin_tc2_kernel<1, false>); break;
            case 2: rc = tr ? launch(lbs_skin_tc2_kernel<2, true>) : launch(lbs_skin_tc2_kernel<2, false>); break;
            case 3: rc = tr ? launch(lbs_skin_tc2_kernel<3, true>) : launch(lbs_skin_tc2_kernel<3, false>); break;
            case 4: rc = tr ? launch(lbs_skin_tc2_kernel<4, true>) : launch(lbs_skin_tc2_kernel<4, false>); break;
            default: rc = tr ? launch(lbs_skin_tc2_kernel<0, true>) : launch(lbs_skin_tc2_kernel<0, false>); break;
        }
        if (rc) return rc;
        HF_LAUNCH_CHECK();
        if ((stage_mask & 4) && h->nvj + h->nextra > 0) {
            HF_CUDA(hf::launch_pdl(lbs_extra_joints_c_kernel, dim3(hf::div_up(M, 4)), dim3(128), 0, stream, (const float*)xvt, (const int*)h->pick_f,
                                   (const int*)h->csr_ptr, (const int*)h->csr_f, (const float*)h->csr_val, M, h->J, h->nvj, h->nextra, J_out, h->NF, joints));
            HF_LAUNCH_CHECK();
        }
        return HF_OK;
    } else if (h->impl == 2) {
        HF_CUDA(hf::launch_pdl(lbs_coef_kernel, dim3(std::min(hf::div_up(M * LBS_KH, 256), 148 * 16)), dim3(256), 0, stream, betas, rotmats,
                               glob ? 1 : 0, M, h->J, h->nb, Fb));
        HF_LAUNCH_CHECK();
        hf_smpl* hm = const_cast<hf_smpl*>(h);
        if (hm->mapB_ptr != (const void*)Fb || hm->mapB_M != M) {
            const uint64_t dims[2] = {(uint64_t)2 * LBS_KH, (uint64_t)M};
            const uint64_t st[1] = {(uint64_t)2 * LBS_KH * 2};
            const uint32_t box[2] = {64, (uint32_t)TC_NS};
            int rc = encode_map(&hm->mapB, Fb, 2, dims, st, box);
            if (rc) return rc;
            hm->mapB_ptr = Fb; hm->mapB_M = M;
        }
        const size_t tc_smem = (size_t)TC_STAGES * TC_STAGE_BYTES + (size_t)TC_NS * h->J * 12 * sizeof(float) + 1024;
        HF_CUDA(cudaFuncSetAttribute(lbs_skin_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));   // per device: every call
        if (tc_smem > 226 * 1024) return hf::fail(HF_ERR_UNSUPPORTED, "hf_lbs_forward: tile needs %zu B of shared memory", tc_smem);
        dim3 tgrid(h->Vp / 128, hf::div_up(M, TC_NS));
        HF_CUDA(hf::launch_pdl(lbs_skin_tc_kernel, tgrid, dim3(TC_THREADS), tc_smem, stream, hm->mapA, hm->mapB, h->vtemp, h->sj, h->sw,
                               A, transl, M, h->V, h->Vp, h->J, h->nslots, vertices));
        HF_LAUNCH_CHECK();
    } else {
    constexpr int TS = kSPT * kSG;
    size_t smem = ((size_t)h->KP * TS + (size_t)TS * h->J * 12) * sizeof(float);
    HF_CUDA(cudaFuncSetAttribute(lbs_skin_kernel<kSPT, kSG>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));   // per device: every call
    if (smem > 200 * 1024) return hf::fail(HF_ERR_UNSUPPORTED, "hf_lbs_forward: tile needs %zu B of shared memory", smem);
    dim3 grid(h->Vp / 128, hf::div_up(M, TS));
    lbs_skin_kernel<kSPT, kSG><<<grid, 128 * kSG, smem, stream>>>(h->blend, h->vtemp, h->sj, h->sw, F, A, transl, M,
                                                                 h->V, h->Vp, h->KB, h->KP, h->J, h->nslots, vertices);
    HF_LAUNCH_CHECK();
    }
    if (stage_mask & 4) {
        int rc = hf_lbs_extra_joints(h, vertices, joints, M, stream);
        if (rc) return rc;
    }
    return HF_OK;
}

extern "C" int hf_lbs_forward(const hf_smpl_t* h, const float* betas, const float* rotmats,
                              const float* transl, float* vertices, float* joints, void* workspace,
                              size_t workspace_bytes, int M, void* stream) {
    return lbs_forward_any(h, betas, rotmats, nullptr, 1, transl, vertices, joints, workspace, workspace_bytes, M, stream);
}

extern "C" int hf_lbs_forward_split(const hf_smpl_t* h, const float* betas, const float* body_rotmats, const float* glob_rotmats,
                                    int rep, const float* transl, float* vertices, float* joints, void* workspace,
                                    size_t workspace_bytes, int M, void* stream) {
    if (!glob_rotmats) return hf::fail(HF_ERR_INVALID, "hf_lbs_forward_split: null argument");
    return lbs_forward_any(h, betas, body_rotmats, glob_rotmats, rep, transl, vertices, joints, workspace, workspace_bytes, M, stream);
}

extern "C" int hf_rodrigues(const float* aa, float* R, int n, void* stream) {
    if (n <= 0) return HF_OK;
    rodrigues_kernel<<<hf::div_up(n, 256), 256, 0, (cudaStream_t)stream>>>(aa, R, n);
    HF_LAUNCH_CHECK();
    return HF_OK;
}


// =====================================================================================================================
// Backward of the LBS forward (the gradient a fitting / training loop needs through models/smpl.py:27-41
// and [upstream] smplx lbs): d loss / d (betas, rotation matrices) from d loss / d (vertices, 90 joints).  fp32 throughout.
//
//   forward      p = v_template + [S | P] f            f = (beta | vec(R_i - I))         (blend)
//                v = sum_k w_vk (A_k.R p + A_k.t)      A_k from the kinematic chain        (skinning)
//                joints = (chain translations | picked vertices | regressor rows . v)
//   backward  1. lbs_bwd_vertex_kernel  per (vertex, sample): g = dL/dv (+ the rows of dL/djoints that read v), recompute p,
//                                       dL/dp = T.R^T g -> G2 (split tf32 [M][hi 3Vp | lo 3Vp]);  dL/dA_k += w_vk g (x) [p;1]
//                                       (shared-memory atomics per CTA, per-vertex-tile partials to global)
//             2. gemm_tf32x3_kernel     dL/df = G2 . [S | P]^T as split-tf32 tcgen05 GEMM, K = 3Vp split across CTAs
//             3. lbs_bwd_chain_kernel   sums the partials, walks the chain backwards (thread per sample), adds the pose-feature and
//                                       joint-regressor terms: dL/dR (M,J,3,3), dL/dbeta (M,nb)
// The shared-memory atomics make the summation order of dL/dA run-dependent (differences at the fp32 rounding level).
namespace {

constexpr int BW_SPT = 4, BW_SG = 4, BW_TS = BW_SPT * BW_SG;

__global__ void __launch_bounds__(128 * BW_SG, 2)
lbs_bwd_vertex_kernel(const float* __restrict__ blend, const float* __restrict__ vtemp, const int* __restrict__ sj,
                      const float* __restrict__ sw, const float* __restrict__ F, const float* __restrict__ A,
                      const int* __restrict__ csc_ptr, const int* __restrict__ csc_row, const float* __restrict__ csc_val,
                      const float* __restrict__ gV, const float* __restrict__ gJ, int M, int V, int Vp, int KB, int KP, int J, int J_out,
                      int nslots, const int* __restrict__ vmap, int ncols, float* __restrict__ G2, float* __restrict__ part) {
    // vmap != NULL: column c of this launch is vertex vmap[c] (or -1 = padding), ncols = padded number of columns (joints-only
    // backward over the vertices a pick / regressor reads); vmap == NULL: column = vertex, ncols = Vp.
    extern __shared__ __align__(16) float smem[];
    const int J12 = J * 12, NX = (J_out - J) * 3;
    float* Fs = smem;                       // [KP][TS]
    float* As = Fs + KP * BW_TS;            // [TS][J12]
    float* gAs = As + BW_TS * J12;          // [TS][J12]
    float* gXs = gAs + BW_TS * J12;         // [TS][NX]   dL/d(non-chain output joints)
    const int m0 = blockIdx.y * BW_TS, tid = threadIdx.x;
    for (int idx = tid; idx < BW_TS * KP; idx += 128 * BW_SG) {
        const int s = idx / KP, k = idx - s * KP, m = m0 + s;
        Fs[k * BW_TS + s] = (m < M) ? F[(size_t)m * KP + k] : 0.f;
    }
    for (int idx = tid; idx < BW_TS * J12; idx += 128 * BW_SG) {
        const int s = idx / J12, m = m0 + s;
        As[idx] = (m < M) ? A[(size_t)m * J12 + (idx - s * J12)] : 0.f;
        gAs[idx] = 0.f;
    }
    for (int idx = tid; idx < BW_TS * NX; idx += 128 * BW_SG) {
        const int s = idx / NX, m = m0 + s;
        gXs[idx] = (gJ && m < M) ? gJ[((size_t)m * J_out + J) * 3 + (idx - s * NX)] : 0.f;
    }
    __syncthreads();
    const int vl = tid & 127, sg = tid >> 7, lane = tid & 31;
    const int col = blockIdx.x * 128 + vl;        // < ncols
    const int vraw = vmap ? vmap[col] : col;      // vertex (< Vp: tables are padded), -1 for a padding column
    const bool live = vraw >= 0 && vraw < V;
    const int v = vraw >= 0 ? vraw : 0;
    float acc[BW_SPT][3];
#pragma unroll
    for (int s = 0; s < BW_SPT; ++s) acc[s][0] = acc[s][1] = acc[s][2] = 0.f;
    // `blend` is [KB][3][bstride]: the full table (column = vertex) or its compact copy (column = compact slot)
    const float* bp = blend + (vmap ? col : v);
    const int bstride = vmap ? ncols : Vp;
    const float* fs = Fs + sg * BW_SPT;
#pragma unroll 8
    for (int k = 0; k < KB; ++k) {
        const float p0 = __ldg(bp + (size_t)(k * 3 + 0) * bstride), p1 = __ldg(bp + (size_t)(k * 3 + 1) * bstride), p2 = __ldg(bp + (size_t)(k * 3 + 2) * bstride);
        const float4 f = *reinterpret_cast<const float4*>(fs + k * BW_TS);
        acc[0][0] = fmaf(p0, f.x, acc[0][0]); acc[0][1] = fmaf(p1, f.x, acc[0][1]); acc[0][2] = fmaf(p2, f.x, acc[0][2]);
        acc[1][0] = fmaf(p0, f.y, acc[1][0]); acc[1][1] = fmaf(p1, f.y, acc[1][1]); acc[1][2] = fmaf(p2, f.y, acc[1][2]);
        acc[2][0] = fmaf(p0, f.z, acc[2][0]); acc[2][1] = fmaf(p1, f.z, acc[2][1]); acc[2][2] = fmaf(p2, f.z, acc[2][2]);
        acc[3][0] = fmaf(p0, f.w, acc[3][0]); acc[3][1] = fmaf(p1, f.w, acc[3][1]); acc[3][2] = fmaf(p2, f.w, acc[3][2]);
    }
    const float t0 = vtemp[v], t1 = vtemp[Vp + v], t2 = vtemp[2 * Vp + v];
    const int e0 = csc_ptr[v], e1 = csc_ptr[v + 1];
#pragma unroll
    for (int s = 0; s < BW_SPT; ++s) {
        const int sl = sg * BW_SPT + s, m = m0 + sl;
        const float p[3] = {acc[s][0] + t0, acc[s][1] + t1, acc[s][2] + t2};
        float g[3] = {0.f, 0.f, 0.f};
        if (m < M && live) {
            if (gV) { const float* q = gV + ((size_t)m * V + v) * 3; g[0] = q[0]; g[1] = q[1]; g[2] = q[2]; }
            for (int e = e0; e < e1; ++e) {
                const float w = csc_val[e];
                const float* q = gXs + sl * NX + csc_row[e] * 3;
                g[0] = fmaf(w, q[0], g[0]); g[1] = fmaf(w, q[1], g[1]); g[2] = fmaf(w, q[2], g[2]);
            }
        }
        float gp[3] = {0.f, 0.f, 0.f};
        const bool any_g = g[0] != 0.f || g[1] != 0.f || g[2] != 0.f;
        // dL/dA_j += w_vj g (x) [p;1]: the lanes of a warp mostly share their joints, so the warp first sums the contributions of
        // all lanes with the same joint (butterfly over the 12 entries) and then issues ONE 12-lane shared-memory atomic per
        // distinct joint, instead of 12 per lane on a handful of addresses
        for (int slot = 0; slot < nslots; ++slot) {
            const float w = any_g ? sw[slot * Vp + v] : 0.f;
            const int j = (w != 0.f) ? sj[slot * Vp + v] : -1;
            if (j >= 0) {
                const float* a = As + sl * J12 + j * 12;
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                    const float wg = w * g[r];
                    gp[0] = fmaf(a[r * 4 + 0], wg, gp[0]); gp[1] = fmaf(a[r * 4 + 1], wg, gp[1]); gp[2] = fmaf(a[r * 4 + 2], wg, gp[2]);
                }
            }
            unsigned todo = __ballot_sync(0xffffffffu, j >= 0);
            while (todo) {
                const int jj = __shfl_sync(0xffffffffu, j, __ffs(todo) - 1);
                const bool mine = j == jj;
                float c[16];
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                    const float wg = mine ? w * g[r] : 0.f;
                    c[r * 4 + 0] = wg * p[0]; c[r * 4 + 1] = wg * p[1]; c[r * 4 + 2] = wg * p[2]; c[r * 4 + 3] = wg;
                }
                c[12] = c[13] = c[14] = c[15] = 0.f;
                // transpose-reduction: 8 + 4 + 2 + 1 + 1 shuffles instead of 12 x 5; afterwards lane l holds the warp sum of entry
                // 8*b4 + 4*b3 + 2*b2 + b1 of its lane index
#pragma unroll
                for (int sft = 16, n = 8; sft >= 2; sft >>= 1, n >>= 1) {
                    const bool up = (lane & sft) != 0;
#pragma unroll
                    for (int i = 0; i < n; ++i) {
                        const float keep = up ? c[i + n] : c[i], send = up ? c[i] : c[i + n];
                        c[i] = keep + __shfl_xor_sync(0xffffffffu, send, sft);
                    }
                }
                c[0] += __shfl_xor_sync(0xffffffffu, c[0], 1);
                const int ent = lane >> 1;       // = 8*b4 + 4*b3 + 2*b2 + b1
                if ((lane & 1) == 0 && ent < 12) atomicAdd(gAs + sl * J12 + jj * 12 + ent, c[0]);
                todo &= ~__ballot_sync(0xffffffffu, mine);
            }
        }
        if (m < M) {
            float* row = G2 + (size_t)m * 6 * ncols;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                uint32_t hb;
                asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"(gp[c]));
                const float hi = __uint_as_float(hb);
                row[c * ncols + col] = hi;
                row[3 * ncols + c * ncols + col] = gp[c] - hi;
            }
        }
    }
    __syncthreads();
    for (int idx = tid; idx < BW_TS * J12; idx += 128 * BW_SG) {
        const int s = idx / J12, m = m0 + s;
        if (m < M) part[((size_t)blockIdx.x * M + m) * J12 + (idx - s * J12)] = gAs[idx];
    }
}

// blend [KB][3][Vp] fp32 -> [256][hi 3Vp | lo 3Vp] (rows >= KB zero)
// (vmap != NULL: compact columns, out is [256][hi 3*ncols | lo 3*ncols] with column (c, slot) = blend[k][c][vmap[slot]])
__global__ void lbs_blend_split_kernel(const float* __restrict__ blend, int KB, int Vp, const int* __restrict__ vmap, int ncols,
                                       float* __restrict__ out) {
    const int Vp3 = 3 * ncols;
    const size_t total = (size_t)256 * Vp3;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
        const int k = (int)(e / Vp3), c = (int)(e - (size_t)k * Vp3);
        const int cc = c / ncols, slot = c - cc * ncols;
        const int v = vmap ? vmap[slot] : slot;
        const float x = (k < KB && v >= 0) ? blend[((size_t)k * 3 + cc) * Vp + v] : 0.f;
        uint32_t hb;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"(x));
        const float hi = __uint_as_float(hb);
        out[(size_t)k * 2 * Vp3 + c] = hi;
        out[(size_t)k * 2 * Vp3 + Vp3 + c] = x - hi;
    }
}

__global__ void lbs_blend_compact_kernel(const float* __restrict__ blend, int KB, int Vp, const int* __restrict__ vmap, int ncols,
                                         float* __restrict__ out) {
    const size_t total = (size_t)KB * 3 * ncols;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
        const int kc = (int)(e / ncols), slot = (int)(e - (size_t)kc * ncols);
        const int v = vmap[slot];
        out[e] = v >= 0 ? blend[(size_t)kc * Vp + v] : 0.f;
    }
}

// D[z][m][n] = sum over this CTA's K range of (Ahi + Alo)[m][k] (Bhi + Blo)[n][k] (dropping lo x lo): split-tf32 on tcgen05,
// both operands K-major [rows][hi K | lo K].  Tile 128 x 128, 32-float k-steps, 3-stage TMA ring (the structure of
// flow_ctx_gemm_kernel in flow.cu); grid (m tiles, n tiles, K splits).
constexpr int GB_STAGES = 3, GB_THREADS = 320, GB_BN = 128, GB_STAGE_BYTES = 2 * 128 * 128 + 2 * GB_BN * 128;
__global__ void __launch_bounds__(GB_THREADS, 1)
gemm_tf32x3_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, int lo_off, int ksteps_total,
                   int ksteps_per_split, int M, int N, float* __restrict__ D) {
    extern __shared__ uint8_t gb_smem[];
    __shared__ __align__(8) uint64_t bars[2 * GB_STAGES + 1];
    __shared__ uint32_t tmem_base_s;
    const uint32_t tile_base = (smem_u32(gb_smem) + 1023u) & ~1023u;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.x * 128, n0 = blockIdx.y * GB_BN;
    const int k_begin = blockIdx.z * ksteps_per_split, nit = min(ksteps_per_split, ksteps_total - k_begin);
    const uint32_t full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[GB_STAGES]), tfull = smem_u32(&bars[2 * GB_STAGES]);
    if (threadIdx.x == 0) {
        for (int s = 0; s < GB_STAGES; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
        mbar_init(tfull, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"((uint32_t)GB_BN) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    HF_PDL_SYNC();
    const uint32_t tmem_base = tmem_base_s;
    if (warp == 0) {
        if (lane == 0) {
            for (int it = 0; it < nit; ++it) {
                const int st = it % GB_STAGES;
                const uint32_t ph = (uint32_t)(it / GB_STAGES) & 1u;
                mbar_wait(empty0 + 8 * st, ph ^ 1u);
                const uint32_t sa = tile_base + st * GB_STAGE_BYTES, fb = full0 + 8 * st;
                const int kc = (k_begin + it) * 32;
                mbar_expect_tx(fb, GB_STAGE_BYTES);
                tma_load_2d(sa, &mapA, fb, kc, m0);
                tma_load_2d(sa + 16384, &mapA, fb, lo_off + kc, m0);
                tma_load_2d(sa + 32768, &mapB, fb, kc, n0);
                tma_load_2d(sa + 32768 + GB_BN * 128, &mapB, fb, lo_off + kc, n0);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_tf32(128, GB_BN);
            for (int it = 0; it < nit; ++it) {
                const int st = it % GB_STAGES;
                const uint32_t ph = (uint32_t)(it / GB_STAGES) & 1u;
                mbar_wait(full0 + 8 * st, ph);
                tcgen05_fence_after();
                const uint32_t sa = tile_base + st * GB_STAGE_BYTES;
                const uint64_t a_hi = umma_desc_sw128(sa), a_lo = umma_desc_sw128(sa + 16384);
                const uint64_t b_hi = umma_desc_sw128(sa + 32768), b_lo = umma_desc_sw128(sa + 32768 + GB_BN * 128);
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_tf32(tmem_base, a_lo + (uint64_t)(2 * k), b_hi + (uint64_t)(2 * k), idesc, (uint32_t)((it | k) != 0));
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_tf32(tmem_base, a_hi + (uint64_t)(2 * k), b_lo + (uint64_t)(2 * k), idesc, 1u);
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_tf32(tmem_base, a_hi + (uint64_t)(2 * k), b_hi + (uint64_t)(2 * k), idesc, 1u);
                umma_commit(empty0 + 8 * st);
            }
            umma_commit(tfull);
        }
    } else {
        const int q = warp & 3, half = (warp - 2) >> 2;     // TMEM lane quarter (= warp % 4), 64-column half
        const int r = m0 + q * 32 + lane;
        if (nit > 0) {
            mbar_wait_warp(tfull, 0);
            tcgen05_fence_after();
        }
        uint32_t v[64];
        const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(half * 64);
        if (nit > 0) {
            tmem_ld32(ta, v);
            tmem_ld32(ta + 32, v + 32);
            tmem_ld_wait();
        } else {
#pragma unroll
            for (int o = 0; o < 64; ++o) v[o] = 0u;
        }
        if (r < M) {
            float* dst = D + ((size_t)blockIdx.z * M + r) * N + n0 + half * 64;
#pragma unroll
            for (int o = 0; o < 64; o += 4)
                if (n0 + half * 64 + o < N) *reinterpret_cast<float4*>(dst + o) = make_float4(__uint_as_float(v[o]), __uint_as_float(v[o + 1]), __uint_as_float(v[o + 2]), __uint_as_float(v[o + 3]));
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)GB_BN) : "memory");
    }
}

// Chain backward, BC samples per block.  All per-sample arrays live in shared memory as [element][sample] (conflict-free for the
// one-thread-per-sample chain walk).  dL/dG of joint i is kept in the 12 floats of gA[i] (rotation part r*4+c, translation r*4+3).
constexpr int BC = 16, BC_THREADS = 512;
__global__ void __launch_bounds__(BC_THREADS)
lbs_bwd_chain_kernel(const float* __restrict__ betas, const float* __restrict__ rotmats, const float* __restrict__ J0, const float* __restrict__ Jd,
                     Parents par, const float* __restrict__ part, int nvt, const float* __restrict__ cpart, int nz, int ldc,
                     const float* __restrict__ gJ, int M, int J, int nb, int KB, int J_out, float* __restrict__ g_betas,
                     float* __restrict__ g_rot) {
    extern __shared__ __align__(16) float sm[];
    const int J12 = J * 12, J9 = J * 9, J3 = J * 3;
    float* gA = sm;                    // [J12][BC]
    float* Gs = gA + J12 * BC;         // [J12][BC]  global transforms (rotation r*4+c, translation r*4+3)
    float* Rs = Gs + J12 * BC;         // [J9][BC]   rotations in, dL/dR out
    float* Js = Rs + J9 * BC;          // [J3][BC]   rest joints
    float* gJs = Js + J3 * BC;         // [J3][BC]   dL/d(rest joints)
    float* dc = gJs + J3 * BC;         // [KB][BC]   dL/d(blend coefficients)
    float* Bs = dc + KB * BC;          // [nb][BC]
    __shared__ int pars[HF_MAXJ];
    const int tid = threadIdx.x, mb = blockIdx.x * BC, ns = min(BC, M - mb);
    if (tid < HF_MAXJ) pars[tid] = par.p[tid];
    HF_PDL_SYNC();
    // sums of the per-vertex-tile / per-K-split partials in a fixed order; four independent elements per thread so that the
    // global loads of one step overlap (a single running sum serialises on the ~600-cycle load latency)
    for (int idx0 = tid; idx0 < ns * J12; idx0 += 4 * BC_THREADS) {
        float a[4] = {0.f, 0.f, 0.f, 0.f};
        size_t off[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int idx = min(idx0 + u * BC_THREADS, ns * J12 - 1), s = idx / J12;
            off[u] = (size_t)(mb + s) * J12 + (idx - s * J12);
        }
        for (int t = 0; t < nvt; ++t)
#pragma unroll
            for (int u = 0; u < 4; ++u) a[u] += part[(size_t)t * M * J12 + off[u]];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int idx = idx0 + u * BC_THREADS;
            if (idx < ns * J12) { const int s = idx / J12; gA[(idx - s * J12) * BC + s] = a[u]; }
        }
    }
    for (int idx0 = tid; idx0 < ns * KB; idx0 += 4 * BC_THREADS) {
        float a[4] = {0.f, 0.f, 0.f, 0.f};
        size_t off[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int idx = min(idx0 + u * BC_THREADS, ns * KB - 1), s = idx / KB;
            off[u] = (size_t)(mb + s) * ldc + (idx - s * KB);
        }
        for (int z = 0; z < nz; ++z)
#pragma unroll
            for (int u = 0; u < 4; ++u) a[u] += cpart[(size_t)z * M * ldc + off[u]];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int idx = idx0 + u * BC_THREADS;
            if (idx < ns * KB) { const int s = idx / KB; dc[(idx - s * KB) * BC + s] = a[u]; }
        }
    }
    for (int idx = tid; idx < ns * J9; idx += BC_THREADS) {
        const int s = idx / J9, e = idx - s * J9;
        Rs[e * BC + s] = rotmats[(size_t)(mb + s) * J9 + e];
    }
    for (int idx = tid; idx < ns * nb; idx += BC_THREADS) {
        const int s = idx / nb, l = idx - s * nb;
        Bs[l * BC + s] = betas[(size_t)(mb + s) * nb + l];
    }
    __syncthreads();
    for (int idx = tid; idx < ns * J3; idx += BC_THREADS) {
        const int s = idx / J3, jr = idx - s * J3;
        float a = J0[jr];
        for (int l = 0; l < nb; ++l) a = fmaf(Jd[jr * nb + l], Bs[l * BC + s], a);
        Js[jr * BC + s] = a;
        gJs[jr * BC + s] = 0.f;
    }
    __syncthreads();
    if (tid < ns) {
        const int s = tid, m = mb + s;
        // per-joint base pointers: the element offsets below are compile-time immediates of the shared-memory loads / stores
        auto GAp = [&](int i) { return gA + i * 12 * BC + s; };
        auto GGp = [&](int i) { return Gs + i * 12 * BC + s; };
        auto RRp = [&](int i) { return Rs + i * 9 * BC + s; };
        auto JJp = [&](int i) { return Js + i * 3 * BC + s; };
        auto GJp = [&](int i) { return gJs + i * 3 * BC + s; };
        // forward chain
        {
            float* g0 = GGp(0); const float* r0 = RRp(0); const float* j0 = JJp(0);
#pragma unroll
            for (int r = 0; r < 3; ++r) {
#pragma unroll
                for (int c = 0; c < 3; ++c) g0[(r * 4 + c) * BC] = r0[(r * 3 + c) * BC];
                g0[(r * 4 + 3) * BC] = j0[r * BC];
            }
        }
        for (int i = 1; i < J; ++i) {
            const int p = pars[i];
            float* gi = GGp(i); const float* gpp = GGp(p); const float* ri = RRp(i); const float* ji = JJp(i); const float* jp = JJp(p);
            float R[9], Gp[12];
#pragma unroll
            for (int e = 0; e < 9; ++e) R[e] = ri[e * BC];
#pragma unroll
            for (int e = 0; e < 12; ++e) Gp[e] = gpp[e * BC];
            const float d0 = ji[0] - jp[0], d1 = ji[BC] - jp[BC], d2 = ji[2 * BC] - jp[2 * BC];
#pragma unroll
            for (int r = 0; r < 3; ++r) {
#pragma unroll
                for (int c = 0; c < 3; ++c) gi[(r * 4 + c) * BC] = Gp[r * 4] * R[c] + Gp[r * 4 + 1] * R[3 + c] + Gp[r * 4 + 2] * R[6 + c];
                gi[(r * 4 + 3) * BC] = Gp[r * 4] * d0 + Gp[r * 4 + 1] * d1 + Gp[r * 4 + 2] * d2 + Gp[r * 4 + 3];
            }
        }
        // dL/dA -> dL/dG (in place): A.R = G.R, A.t = G.t - G.R J;  posed joint i = G_i.t
        for (int i = 0; i < J; ++i) {
            float* ga = GAp(i); const float* gg = GGp(i); const float* ji = JJp(i); float* gj = GJp(i);
            float gt[3], Jv[3];
#pragma unroll
            for (int r = 0; r < 3; ++r) { gt[r] = ga[(r * 4 + 3) * BC]; Jv[r] = ji[r * BC]; }
#pragma unroll
            for (int c = 0; c < 3; ++c) gj[c * BC] -= gg[c * BC] * gt[0] + gg[(4 + c) * BC] * gt[1] + gg[(8 + c) * BC] * gt[2];
#pragma unroll
            for (int r = 0; r < 3; ++r) {
#pragma unroll
                for (int c = 0; c < 3; ++c) ga[(r * 4 + c) * BC] -= gt[r] * Jv[c];
                if (gJ) ga[(r * 4 + 3) * BC] += gJ[((size_t)m * J_out + i) * 3 + r];
            }
        }
        // children before parents
        for (int i = J - 1; i >= 1; --i) {
            const int p = pars[i];
            float* ga = GAp(i); float* gap = GAp(p); const float* gpp = GGp(p); float* ri = RRp(i);
            const float* ji = JJp(i); const float* jp = JJp(p); float* gji = GJp(i); float* gjp = GJp(p);
            float gr[9], gt[3], R[9], Gp[9], gd[3] = {0.f, 0.f, 0.f}, gR[9], acc[12];
#pragma unroll
            for (int r = 0; r < 3; ++r) {
#pragma unroll
                for (int c = 0; c < 3; ++c) { gr[r * 3 + c] = ga[(r * 4 + c) * BC]; Gp[r * 3 + c] = gpp[(r * 4 + c) * BC]; }
                gt[r] = ga[(r * 4 + 3) * BC];
            }
#pragma unroll
            for (int e = 0; e < 9; ++e) { R[e] = ri[e * BC]; gR[e] = 0.f; }
#pragma unroll
            for (int e = 0; e < 12; ++e) acc[e] = gap[e * BC];
            const float d[3] = {ji[0] - jp[0], ji[BC] - jp[BC], ji[2 * BC] - jp[2 * BC]};
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                // dL/dR_i = G_p.R^T dL/dG_i.R ; dL/d(J_i - J_p) = G_p.R^T dL/dG_i.t
#pragma unroll
                for (int c = 0; c < 3; ++c) { gR[c] += Gp[r * 3] * gr[r * 3 + c]; gR[3 + c] += Gp[r * 3 + 1] * gr[r * 3 + c]; gR[6 + c] += Gp[r * 3 + 2] * gr[r * 3 + c]; }
                gd[0] += Gp[r * 3] * gt[r]; gd[1] += Gp[r * 3 + 1] * gt[r]; gd[2] += Gp[r * 3 + 2] * gt[r];
                // dL/dG_p.R += dL/dG_i.R R_i^T + dL/dG_i.t (x) d ; dL/dG_p.t += dL/dG_i.t
#pragma unroll
                for (int c = 0; c < 3; ++c)
                    acc[r * 4 + c] += gr[r * 3] * R[c * 3] + gr[r * 3 + 1] * R[c * 3 + 1] + gr[r * 3 + 2] * R[c * 3 + 2] + gt[r] * d[c];
                acc[r * 4 + 3] += gt[r];
            }
#pragma unroll
            for (int e = 0; e < 12; ++e) gap[e * BC] = acc[e];
#pragma unroll
            for (int c = 0; c < 3; ++c) { gji[c * BC] += gd[c]; gjp[c * BC] -= gd[c]; }
#pragma unroll
            for (int e = 0; e < 9; ++e) ri[e * BC] = gR[e];
        }
        {
            const float* ga = GAp(0); float* r0 = RRp(0); float* gj = GJp(0);
#pragma unroll
            for (int r = 0; r < 3; ++r) {
#pragma unroll
                for (int c = 0; c < 3; ++c) r0[(r * 3 + c) * BC] = ga[(r * 4 + c) * BC];
                gj[r * BC] += ga[(r * 4 + 3) * BC];
            }
        }
    }
    __syncthreads();
    for (int idx = tid; idx < ns * J9; idx += BC_THREADS) {          // + pose-feature term (f = R_i - I for i >= 1)
        const int s = idx / J9, e = idx - s * J9;
        float a = Rs[e * BC + s];
        if (e >= 9) a += dc[(nb + e - 9) * BC + s];
        g_rot[(size_t)(mb + s) * J9 + e] = a;
    }
    for (int idx = tid; idx < ns * nb; idx += BC_THREADS) {          // + joint-regressor term
        const int s = idx / nb, l = idx - s * nb;
        float a = dc[l * BC + s];
        for (int jr = 0; jr < J3; ++jr) a = fmaf(Jd[jr * nb + l], gJs[jr * BC + s], a);
        g_betas[(size_t)(mb + s) * nb + l] = a;
    }
}

struct BwdLayout { size_t F, A, joints, G2, part, cpart, total; int nvt, nz, ksteps_per, ncols; };
// dense = a gradient w.r.t. all vertices comes in; otherwise only the NFp compact columns are processed
BwdLayout bwd_layout(const hf_smpl* h, int M, bool dense) {
    BwdLayout L;
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    const int J_out = h->J + h->nvj + h->nextra;
    L.ncols = dense ? h->Vp : h->NFp;
    L.nvt = L.ncols / 128;
    const int ksteps = 3 * L.ncols / 32, mt = hf::div_up(M, 128);
    int nz = std::max(1, std::min(8, 148 / (mt * 2)));
    L.ksteps_per = hf::div_up(ksteps, nz);
    L.nz = hf::div_up(ksteps, L.ksteps_per);
    size_t o = 0;
    L.F = o; o = al(o + (size_t)M * h->KP * 4);
    L.A = o; o = al(o + (size_t)M * h->J * 12 * 4);
    L.joints = o; o = al(o + (size_t)M * J_out * 3 * 4);
    L.G2 = o; o = al(o + (size_t)M * 6 * L.ncols * 4);
    L.part = o; o = al(o + (size_t)L.nvt * M * h->J * 12 * 4);
    L.cpart = o; o = al(o + (size_t)L.nz * M * 256 * 4);
    L.total = o + 256;
    return L;
}

}  // namespace

extern "C" size_t hf_lbs_backward_workspace_bytes(const hf_smpl_t* h, int M) {
    if (!h || M <= 0) return 0;
    return bwd_layout(h, M, true).total;        // the dense layout is the larger one
}

extern "C" int hf_lbs_backward(hf_smpl_t* h, const float* betas, const float* rotmats, const float* grad_vertices, const float* grad_joints,
                               float* grad_betas, float* grad_rotmats, void* workspace, size_t workspace_bytes, int M, void* stream_) {
    if (!h || !betas || !rotmats || !grad_betas || !grad_rotmats) return hf::fail(HF_ERR_INVALID, "hf_lbs_backward: null argument");
    if (M <= 0) return HF_OK;
    const bool dense = grad_vertices != nullptr;
    if (!dense && !grad_joints) return hf::fail(HF_ERR_INVALID, "hf_lbs_backward: both incoming gradients are NULL");
    const BwdLayout L = bwd_layout(h, M, dense);
    if (!workspace || workspace_bytes < L.total)
        return hf::fail(HF_ERR_INVALID, "hf_lbs_backward: workspace too small (%zu < %zu)", workspace_bytes, L.total);
    if (3 * L.ncols % 32) return hf::fail(HF_ERR_UNSUPPORTED, "hf_lbs_backward: padded column count %d", L.ncols);
    cudaStream_t stream = (cudaStream_t)stream_;
    uint8_t* wsb = (uint8_t*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    float *F = (float*)(wsb + L.F), *A = (float*)(wsb + L.A), *jtmp = (float*)(wsb + L.joints), *G2 = (float*)(wsb + L.G2);
    float *part = (float*)(wsb + L.part), *cpart = (float*)(wsb + L.cpart);
    const int J_out = hf_smpl_num_joints_out(h), C3 = 3 * L.ncols;
    int rc;
    float*& bsplit = dense ? h->blend_split : h->blend_split_c;
    CUtensorMap& mapBs = dense ? h->mapBs : h->mapBsc;
    if (!bsplit) {
        HF_CUDA(cudaMalloc(&bsplit, (size_t)256 * 2 * C3 * sizeof(float)));
        lbs_blend_split_kernel<<<1024, 256, 0, stream>>>(h->blend, h->KB, h->Vp, dense ? (const int*)nullptr : (const int*)h->fvert, L.ncols, bsplit);
        HF_LAUNCH_CHECK();
        const uint64_t dims[2] = {(uint64_t)(2 * C3), 256}, st[1] = {(uint64_t)(2 * C3) * 4};
        const uint32_t box[2] = {32, (uint32_t)GB_BN};
        if ((rc = encode_map(&mapBs, bsplit, 2, dims, st, box, CU_TENSOR_MAP_DATA_TYPE_FLOAT32))) return rc;
        if (!dense) {
            HF_CUDA(cudaMalloc(&h->blend_c, (size_t)h->KB * 3 * L.ncols * sizeof(float)));
            lbs_blend_compact_kernel<<<256, 256, 0, stream>>>(h->blend, h->KB, h->Vp, h->fvert, L.ncols, h->blend_c);
            HF_LAUNCH_CHECK();
        }
    }
    const void*& mapG_ptr = dense ? h->mapG_ptr : h->mapGc_ptr;
    int& mapG_M = dense ? h->mapG_M : h->mapGc_M;
    CUtensorMap& mapG = dense ? h->mapG : h->mapGc;
    if (mapG_ptr != (const void*)G2 || mapG_M != M) {
        const uint64_t dims[2] = {(uint64_t)(2 * C3), (uint64_t)M}, st[1] = {(uint64_t)(2 * C3) * 4};
        const uint32_t box[2] = {32, 128};
        if ((rc = encode_map(&mapG, G2, 2, dims, st, box, CU_TENSOR_MAP_DATA_TYPE_FLOAT32))) return rc;
        mapG_ptr = G2; mapG_M = M;
    }
    Parents par;
    for (int i = 0; i < HF_MAXJ; ++i) par.p[i] = i < h->J ? h->parents[i] : 0;
    // forward quantities the backward needs: fp32 blend coefficients F and relative transforms A
    HF_CUDA(hf::launch_pdl(lbs_pose_kernel, dim3(hf::div_up(M, PS)), dim3(PTHREADS), 0, stream, betas, rotmats, (const float*)nullptr, 1, (const float*)nullptr,
                           h->J0, h->Jd, par, M, h->J, h->nb, h->KP, J_out, F, (__half*)nullptr, A, (float*)nullptr, jtmp));
    HF_LAUNCH_CHECK();
    {
        const size_t smem = ((size_t)h->KP * BW_TS + 2 * (size_t)BW_TS * h->J * 12 + (size_t)BW_TS * (J_out - h->J) * 3) * sizeof(float);
        HF_CUDA(cudaFuncSetAttribute(lbs_bwd_vertex_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        HF_CUDA(hf::launch_pdl(lbs_bwd_vertex_kernel, dim3(L.nvt, hf::div_up(M, BW_TS)), dim3(128 * BW_SG), smem, stream, dense ? (const float*)h->blend : (const float*)h->blend_c, h->vtemp, h->sj, h->sw,
                               (const float*)F, (const float*)A, h->csc_ptr, h->csc_row, h->csc_val, grad_vertices, grad_joints, M, h->V, h->Vp, h->KB, h->KP,
                               h->J, J_out, h->nslots, dense ? (const int*)nullptr : (const int*)h->fvert, L.ncols, G2, part));
        HF_LAUNCH_CHECK();
    }
    {
        const size_t smem = (size_t)GB_STAGES * GB_STAGE_BYTES + 1024;
        HF_CUDA(cudaFuncSetAttribute(gemm_tf32x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        HF_CUDA(hf::launch_pdl(gemm_tf32x3_kernel, dim3(hf::div_up(M, 128), 256 / GB_BN, L.nz), dim3(GB_THREADS), smem, stream, mapG, mapBs, C3, C3 / 32,
                               L.ksteps_per, M, 256, cpart));
        HF_LAUNCH_CHECK();
    }
    {
        const size_t smem = ((size_t)(2 * h->J * 12 + h->J * 9 + 2 * h->J * 3 + h->KB + h->nb) * BC) * sizeof(float);
        HF_CUDA(cudaFuncSetAttribute(lbs_bwd_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        HF_CUDA(hf::launch_pdl(lbs_bwd_chain_kernel, dim3(hf::div_up(M, BC)), dim3(BC_THREADS), smem, stream, betas, rotmats, (const float*)h->J0,
                               (const float*)h->Jd, par, (const float*)part, L.nvt, (const float*)cpart, L.nz, 256, grad_joints, M, h->J, h->nb, h->KB, J_out,
                               grad_betas, grad_rotmats));
        HF_LAUNCH_CHECK();
    }
    return HF_OK;
}
