// Per-image reductions and projections applied to the sampled meshes right after the LBS:
//   hf_vertex_variance    utils/sampling_utils.py:22-33  compute_vertex_variance_from_samples, batched over images
//   hf_project_joints2d   utils/sampling_utils.py:50-58 / evaluate_humaniflow.py:186-206: joint selection, 180-degree flip
//                         about x (utils/rigid_transform_utils.py:67-83), utils/cam_utils.py:9-16 weak-perspective
//                         projection, utils/joints2d_utils.py:5-10 pixel un-normalisation
//   hf_lbs_tpose          models/smpl.py:27-41 with the default (zero) pose: predict_humaniflow.py:147,
//                         evaluate_humaniflow.py:131-133 -- shape blend only, no pose blend / skinning
#include "common.cuh"
#include <algorithm>

struct hf_smpl;
// accessors implemented in lbs.cu (the handle layout is private to that file)
extern "C" int hf_smpl_dims(const hf_smpl* h, int* V, int* Vp, int* nb, int* J, int* J_out);
int hf_smpl_tpose_tables(const hf_smpl* h, const float** blend, const float** vtemp, const float** J0, const float** Jd);
int hf_lbs_extra_joints(const hf_smpl* h, const float* vertices, float* joints, int M, cudaStream_t stream);

namespace {

// One block per (image, chunk of VC vertices): the chunk's N x (VC*3) samples are staged in shared memory once
// (coalesced rows of VC*12 bytes), so global memory is read exactly once; then the mean over samples, the RMS deviation
// per coordinate and the mean Euclidean distance from the mean come from shared memory.
constexpr int VC = 32, VV_THREADS = 128;
__global__ void __launch_bounds__(VV_THREADS)
vertex_variance_kernel(const float* __restrict__ vertices, int B, int N, int V, float* __restrict__ avg_dist,
                       float* __restrict__ dir_std) {
    extern __shared__ float sm[];              // [N][VC*3] samples, then [VC*3] mean
    HF_PDL_SYNC();
    const int b = blockIdx.y, v0 = blockIdx.x * VC;
    const int nv = min(VC, V - v0), ne = nv * 3;
    float* mean = sm + (size_t)N * VC * 3;
    const float* src = vertices + ((size_t)b * N * V + v0) * 3;
    for (int idx = threadIdx.x; idx < N * ne; idx += VV_THREADS) {
        const int s = idx / ne, e = idx - s * ne;
        sm[s * (VC * 3) + e] = __ldcs(src + (size_t)s * V * 3 + e);
    }
    __syncthreads();
    const float inv_n = 1.f / (float)N;
    for (int e = threadIdx.x; e < ne; e += VV_THREADS) {
        float acc = 0.f;
        for (int s = 0; s < N; ++s) acc += sm[s * (VC * 3) + e];
        mean[e] = acc * inv_n;
    }
    __syncthreads();
    // directional RMS deviation: sqrt(mean((x - mean)^2)) per coordinate
    for (int e = threadIdx.x; e < ne; e += VV_THREADS) {
        const float mu = mean[e];
        float acc = 0.f;
        for (int s = 0; s < N; ++s) { const float d = sm[s * (VC * 3) + e] - mu; acc = fmaf(d, d, acc); }
        dir_std[((size_t)b * V + v0) * 3 + e] = sqrtf(acc * inv_n);
    }
    // mean Euclidean distance from the mean: 4 threads per vertex, each a quarter of the samples
    {
        const int vl = threadIdx.x >> 2, part = threadIdx.x & 3;
        float acc = 0.f;
        if (vl < nv) {
            const float mx = mean[vl * 3], my = mean[vl * 3 + 1], mz = mean[vl * 3 + 2];
            for (int s = part; s < N; s += 4) {
                const float* p = sm + s * (VC * 3) + vl * 3;
                const float dx = p[0] - mx, dy = p[1] - my, dz = p[2] - mz;
                acc += sqrtf(dx * dx + dy * dy + dz * dz);
            }
        }
        acc += __shfl_xor_sync(0xffffffffu, acc, 1);
        acc += __shfl_xor_sync(0xffffffffu, acc, 2);
        if (vl < nv && part == 0) avg_dist[(size_t)b * V + v0 + vl] = acc * inv_n;
    }
}

// out[m][i] = img_scale( cam.s * (flip(joints[m][ids[i]]).xy + cam.t) ), cam = cam_wp[m / per_cam]
__global__ void project_joints2d_kernel(const float* __restrict__ joints, const float* __restrict__ cam_wp,
                                        const int* __restrict__ ids, int M, int per_cam, int J_in, int n_ids, int flip_x,
                                        float img_wh, float* __restrict__ out) {
    HF_PDL_SYNC();
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= M * n_ids) return;
    const int m = idx / n_ids, i = idx - m * n_ids;
    const float* p = joints + ((size_t)m * J_in + (ids ? ids[i] : i)) * 3;
    const float* c = cam_wp + (size_t)(m / per_cam) * 3;
    const float x = p[0], y = flip_x ? -p[1] : p[1];       // rotation by pi about x: (x, y, z) -> (x, -y, -z)
    float u = c[0] * (x + c[1]), v = c[0] * (y + c[2]);
    if (img_wh > 0.f) { u = (u + 1.f) * (img_wh * 0.5f); v = (v + 1.f) * (img_wh * 0.5f); }
    out[(size_t)idx * 2] = u;
    out[(size_t)idx * 2 + 1] = v;
}

// T-pose: vertices = v_template + shapedirs . beta (+ transl); the first J joints are the regressed rest joints
__global__ void lbs_tpose_kernel(const float* __restrict__ blend, const float* __restrict__ vtemp,
                                 const float* __restrict__ J0, const float* __restrict__ Jd, const float* __restrict__ betas,
                                 const float* __restrict__ transl, int M, int V, int Vp, int nb, int J, int J_out,
                                 float* __restrict__ vertices, float* __restrict__ joints) {
    HF_PDL_SYNC();
    __shared__ float bs[16];
    __shared__ float tr[3];
    const int m = blockIdx.y;
    if (threadIdx.x < nb) bs[threadIdx.x] = betas[(size_t)m * nb + threadIdx.x];
    if (threadIdx.x < 3) tr[threadIdx.x] = transl ? transl[m * 3 + threadIdx.x] : 0.f;
    __syncthreads();
    // thread = one output float (coalesced stores); basis rows are [l][c][Vp]
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < V * 3; e += gridDim.x * blockDim.x) {
        const int v = e / 3, c = e - v * 3;
        float acc = vtemp[c * Vp + v];
        for (int l = 0; l < nb; ++l) acc = fmaf(__ldg(blend + ((size_t)l * 3 + c) * Vp + v), bs[l], acc);
        __stcs(vertices + (size_t)m * V * 3 + e, acc + tr[c]);
    }
    if (blockIdx.x == 0) {
        for (int e = threadIdx.x; e < J * 3; e += blockDim.x) {
            float acc = J0[e];
            for (int l = 0; l < nb; ++l) acc = fmaf(Jd[e * nb + l], bs[l], acc);
            joints[((size_t)m * J_out) * 3 + e] = acc + tr[e % 3];
        }
    }
}

}  // namespace

extern "C" int hf_vertex_variance(const float* vertices, int B, int N, int V, float* avg_dist, float* dir_std, void* stream) {
    if (!vertices || !avg_dist || !dir_std) return hf::fail(HF_ERR_INVALID, "hf_vertex_variance: null argument");
    if (B <= 0 || V <= 0) return HF_OK;
    if (N <= 0) return hf::fail(HF_ERR_INVALID, "hf_vertex_variance: N must be positive");
    const size_t smem = ((size_t)N * VC * 3 + VC * 3) * sizeof(float);
    if (smem > 200 * 1024) return hf::fail(HF_ERR_UNSUPPORTED, "hf_vertex_variance: %d samples per image exceed the shared-memory tile (max 532)", N);
    HF_CUDA(cudaFuncSetAttribute(vertex_variance_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));   // per device: every call
    HF_CUDA(hf::launch_pdl(vertex_variance_kernel, dim3(hf::div_up(V, VC), B), dim3(VV_THREADS), smem, (cudaStream_t)stream, vertices, B, N, V,
                           avg_dist, dir_std));
    HF_LAUNCH_CHECK();
    return HF_OK;
}

extern "C" int hf_project_joints2d(const float* joints, const float* cam_wp, const int* joint_ids, int M, int per_cam, int J_in,
                                   int n_ids, int flip_x, float img_wh, float* out, void* stream) {
    if (!joints || !cam_wp || !out) return hf::fail(HF_ERR_INVALID, "hf_project_joints2d: null argument");
    if (M <= 0 || n_ids <= 0) return HF_OK;
    if (per_cam <= 0) return hf::fail(HF_ERR_INVALID, "hf_project_joints2d: per_cam must be positive");
    HF_CUDA(hf::launch_pdl(project_joints2d_kernel, dim3(hf::div_up(M * n_ids, 256)), dim3(256), 0, (cudaStream_t)stream, joints, cam_wp,
                           joint_ids, M, per_cam, J_in, n_ids, flip_x, img_wh, out));
    HF_LAUNCH_CHECK();
    return HF_OK;
}

extern "C" int hf_lbs_tpose(const hf_smpl_t* h, const float* betas, const float* transl, float* vertices, float* joints, int M,
                            void* stream) {
    if (!h || !betas || !vertices || !joints) return hf::fail(HF_ERR_INVALID, "hf_lbs_tpose: null argument");
    if (M <= 0) return HF_OK;
    int V, Vp, nb, J, J_out;
    hf_smpl_dims(h, &V, &Vp, &nb, &J, &J_out);
    const float *blend, *vtemp, *J0, *Jd;
    hf_smpl_tpose_tables(h, &blend, &vtemp, &J0, &Jd);
    HF_CUDA(hf::launch_pdl(lbs_tpose_kernel, dim3(std::min(hf::div_up(V * 3, 256), 27), M), dim3(256), 0, (cudaStream_t)stream, blend, vtemp,
                           J0, Jd, betas, transl, M, V, Vp, nb, J, J_out, vertices, joints));
    HF_LAUNCH_CHECK();
    return hf_lbs_extra_joints(h, vertices, joints, M, (cudaStream_t)stream);
}
