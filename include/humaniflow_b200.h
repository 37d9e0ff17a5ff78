/*
 * humaniflow_b200 — C ABI of the B200-native HuManiFlow sampling hot path.
 *
 * The reference (akashsengupta1997/HuManiFlow) is pure Python/PyTorch and has no FFI layer; its
 * boundary for this path is two Python classes.  This header is the C-ABI that the
 * Python mirrors of those classes (humaniflow_b200/humaniflow_model.py, humaniflow_b200/smpl.py)
 * bind with ctypes.  Every entry point cites the reference interface it replaces.
 *
 * Conventions
 *   - plain pointers and sizes only; no torch / C++ types in any signature;
 *   - `*_create` functions take HOST pointers, copy + repack onto the current CUDA device and return
 *     an opaque handle (immutable afterwards); `*_destroy` frees it;
 *   - all other pointer arguments are DEVICE pointers borrowed for the duration of the call;
 *     no allocation happens inside a forward call — scratch comes from the caller (`*_workspace_bytes`);
 *   - `stream` is a cudaStream_t passed as void*; calls are asynchronous on that stream;
 *   - return value 0 = ok, non-zero = error (hf_last_error() gives the message, thread-local);
 *   - tensors are dense row-major fp32 unless stated otherwise.
 */
#ifndef HUMANIFLOW_B200_H
#define HUMANIFLOW_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HF_OK 0
#define HF_ERR_INVALID 1
#define HF_ERR_CUDA 2
#define HF_ERR_UNSUPPORTED 3

/* Library / device bookkeeping. */
int hf_version(void);
const char* hf_last_error(void);
/* Number of kernels this library has launched in this process (bench.py's `gpu_launches`). */
long long hf_launch_count(void);

/* ------------------------------------------------------------------------------------------------
 * SMPL linear blend skinning.
 * Replaces: models/smpl.py:13-41 `SMPL.__init__/forward` and, beneath it, smplx 0.1.26
 * `SMPL.forward` -> `lbs` / `batch_rigid_transform` / `vertices2joints` / `VertexJointSelector`.
 * ---------------------------------------------------------------------------------------------- */
typedef struct hf_smpl hf_smpl_t;

/* Host arrays, smplx buffer layouts: v_template (V,3); shapedirs (V,3,num_betas); posedirs
 * ((J-1)*9, V*3); J_regressor (J,V); lbs_weights (V,J); parents (J) with parents[0] = -1;
 * vertex_joint_ids (num_vertex_joints) = vertices appended as joints (smplx VertexJointSelector);
 * extra_regressors (num_extra, V) = the reference's J_regressor_extra / cocoplus / h36m stacked
 * (models/smpl.py:16-25).  Output joint count = J + num_vertex_joints + num_extra (90 for SMPL). */
int hf_smpl_create(hf_smpl_t** out, int num_verts, int num_betas, int num_joints,
                   const float* v_template, const float* shapedirs, const float* posedirs,
                   const float* J_regressor, const float* lbs_weights, const int* parents,
                   const int* vertex_joint_ids, int num_vertex_joints,
                   const float* extra_regressors, int num_extra);
void hf_smpl_destroy(hf_smpl_t* h);
int hf_smpl_num_joints_out(const hf_smpl_t* h);
size_t hf_lbs_workspace_bytes(const hf_smpl_t* h, int M);

/* impl: 0 = persistent fp16 tcgen05 blend (product path: pose terms one fp16 product each, shape terms a three-product
 * split; vertex error <= 5e-5 m), 1 = FP32 CUDA-core blend, 2 = split-bf16 three-pass tcgen05 blend (cross-checks),
 * 3 = the same fp16 blend with TMEM lanes = samples and register-resident joint transforms (experiment, <= 4 influences). */
int hf_lbs_set_impl(hf_smpl_t* h, int impl);

/* Threading contract of every handle in this header: calls on ONE handle must not overlap (one stream / one host thread at a
 * time); the per-call tensor maps of the workspace are cached inside the handle.  Use one handle per stream.
 * betas (M,num_betas); rotmats (M,J,3,3) = [global_orient | body_pose] (the `pose2rot=False` form);
 * transl (M,3) or NULL.  vertices (M,V,3); joints (M,J_out,3).  models/smpl.py:27-41. */
int hf_lbs_forward(const hf_smpl_t* h, const float* betas, const float* rotmats, const float* transl,
                   float* vertices, float* joints, void* workspace, size_t workspace_bytes,
                   int M, void* stream);

/* Same with the rotations as the model produces them (models/humaniflow_model.py:272,311 + predict_humaniflow.py:138-141):
 * body_rotmats (M,J-1,3,3) and ONE global rotation per image, glob_rotmats (M/rep,3,3), shared by the `rep` consecutive
 * samples of that image: no concatenated (M,J,3,3) copy has to be built first. */
int hf_lbs_forward_split(const hf_smpl_t* h, const float* betas, const float* body_rotmats, const float* glob_rotmats,
                         int rep, const float* transl, float* vertices, float* joints, void* workspace,
                         size_t workspace_bytes, int M, void* stream);

/* Backward of hf_lbs_forward (the reference gets it from torch.autograd through models/smpl.py:27-41 and
 * [upstream] smplx lbs.lbs, e.g. in a fitting loop that optimises pose / shape against 2-D joints):
 *   grad_vertices (M,V,3) or NULL, grad_joints (M,J_out,3) or NULL  ->  grad_betas (M,num_betas), grad_rotmats (M,J,3,3).
 * A translation only shifts the outputs: d loss / d transl = sum over vertices and joints of the incoming gradients (caller).
 * fp32; the per-joint sums over the vertices use shared-memory atomics (run-to-run differences at the rounding level).
 * The first call on a handle builds the split-tf32 copies of the blend basis (cudaMalloc + one kernel: not capturable in a CUDA
 * graph; later calls only launch kernels).  A joints-only call (grad_vertices == NULL) touches only the ~4 % of the vertices that a
 * joint pick / regressor row reads. */
size_t hf_lbs_backward_workspace_bytes(const hf_smpl_t* h, int M);
int hf_lbs_backward(hf_smpl_t* h, const float* betas, const float* rotmats, const float* grad_vertices, const float* grad_joints,
                    float* grad_betas, float* grad_rotmats, void* workspace, size_t workspace_bytes, int M, void* stream);

/* T-pose forward (zero pose): vertices = v_template + shapedirs.betas (+ transl), joints as hf_lbs_forward would give for
 * identity rotations.  Replaces models/smpl.py:27-41 called with the default pose (predict_humaniflow.py:147,
 * evaluate_humaniflow.py:131-133); skips pose blend and skinning. */
int hf_lbs_tpose(const hf_smpl_t* h, const float* betas, const float* transl, float* vertices, float* joints, int M,
                 void* stream);
int hf_smpl_dims(const hf_smpl_t* h, int* V, int* Vp, int* num_betas, int* J, int* J_out);

/* ------------------------------------------------------------------------------------------------
 * Per-image reductions / projections of the sampled meshes.
 * ---------------------------------------------------------------------------------------------- */
/* vertices (B,N,V,3) -> avg_dist (B,V) = mean over the N samples of ||x - mean||, dir_std (B,V,3) =
 * sqrt(mean((x - mean)^2)).  Replaces utils/sampling_utils.py:22-33 compute_vertex_variance_from_samples, batched over
 * images (predict_humaniflow.py:168).  N <= 532. */
int hf_vertex_variance(const float* vertices, int B, int N, int V, float* avg_dist, float* dir_std, void* stream);

/* joints (M,J_in,3) -> out (M,n_ids,2): select joint_ids (NULL = first n_ids), optional rotation by pi about the x axis
 * (utils/rigid_transform_utils.py:67-83 with axes=x, angles=pi), weak-perspective projection with cam_wp[m / per_cam]
 * = (s, tx, ty) (utils/cam_utils.py:9-16), and, if img_wh > 0, un-normalisation to pixels (utils/joints2d_utils.py:5-10).
 * Replaces utils/sampling_utils.py:50-58 and evaluate_humaniflow.py:186-206. */
int hf_project_joints2d(const float* joints, const float* cam_wp, const int* joint_ids, int M, int per_cam, int J_in,
                        int n_ids, int flip_x, float img_wh, float* out, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Per-sample 3-D error metrics of the evaluation path.
 * pred (B*N, P, 3) point sets (sampled meshes or joint sets), target (B, P, 3), out (B*N, 3) =
 * [mean_i ||p_i - t_i||, the same after scale_and_translation_transform_batch, the same after procrustes_analysis_batch].
 * Replaces utils/eval_utils.py:62-125 + the norm / mean reductions of metrics/eval_metrics_tracker.py:119-280
 * (PVE, PVE-SC, PVE-PA, PVE-T(-SC), MPJPE(-SC/-PA) and their samples_min forms are sums / minima of these values).
 * ---------------------------------------------------------------------------------------------- */
int hf_pointset_errors(const float* pred, const float* target, int B, int N, int P, float* out, void* stream);
/* Same with a scratch buffer (hf_pointset_errors_workspace_bytes(B, N) bytes, 8-byte aligned): with >= 512 point sets the two
 * passes over the points run as their own memory-bound launches around a one-thread-per-set solve, instead of one block per
 * set in which 255 threads wait for the 3x3 solve.  Same values (same per-thread arithmetic; block sums in the same order). */
size_t hf_pointset_errors_workspace_bytes(int B, int N);
int hf_pointset_errors_ws(const float* pred, const float* target, int B, int N, int P, float* out, void* workspace,
                          size_t workspace_bytes, void* stream);

/* Per-image reduction over the samples of per-sample error values: err (B,N,K) -> out (B,2K) = [min over n (K) | mean over n (K)].
 * The "samples_min" metrics of metrics/eval_metrics_tracker.py:201-280 are the minimum over the samples of the per-sample mean
 * error (argmin + gather there), the sample-mean forms are the mean; one launch instead of one reduction per metric. */
int hf_samples_reduce(const float* err, int B, int N, int K, float* out, void* stream);

/* Per-image sample statistics of sampled point sets: points (B,N,P,D), D = 3 (vertices / 3-D joints) or 2 (projected joints);
 * target (B,P,D) or NULL; weights (B,P) (visibility flags) or NULL = 1.  out (B,2):
 *   [0] sample diversity = mean over (samples, points) of w ||x - mean over the samples||
 *       (metrics/eval_metrics_tracker.py:397-433: verts3D / joints3D / joints3D_(in)vis _sample_diversity, per frame)
 *   [1] samples-L2E = sum w ||x - target|| / (N sum w)   (:339-374: joints2Dsamples-L2E, input_joints2Dsamples-L2E). */
int hf_sample_stats(const float* points, const float* target, const float* weights, int B, int N, int P, int D, float* out,
                    void* stream);

/* ------------------------------------------------------------------------------------------------
 * Input proxy representation: rgb (B,C,H,W) fp32 + joints2D (B,J,2) = (column, row) pixel
 * coordinates [+ joints_vis (B,J) or NULL] -> out (B,1+J,H,W): channel 0 = edge map, channels 1..J = Gaussian heatmaps.
 * Replaces models/canny_edge_detector.py:104-166 (gauss5 = the five normalised Gaussian taps, threshold, nms = EDGE_NMS)
 * and utils/label_conversions.py:106-125 times the visibility flags (predict_humaniflow.py:96-110).
 * dbg_mag / dbg_ori: optional (B,H,W) outputs of the gradient magnitude / binned orientation (tests), else NULL.
 * ---------------------------------------------------------------------------------------------- */
int hf_proxy_rep(const float* rgb, const float* joints2D, const float* joints_vis, int B, int C, int H, int W, int J,
                 const float* gauss5, float threshold, int nms, float heat_std, float* out, float* dbg_mag, float* dbg_ori,
                 void* stream);

/* Same computation, written straight into the encoder's stem input (SURVEY.md 8f N4 as worded: "fused into the encoder's
 * first-layer staging"): staged = the buffer hf_encoder_stem_input returns, bf16 NHWC (B, Hp, Wp, Cp) with the image at
 * (top, left) inside a zero border and channels 1+J..Cp-1 zero.  Follow with hf_encoder_forward_staged. */
int hf_proxy_rep_staged(const float* rgb, const float* joints2D, const float* joints_vis, int B, int C, int H, int W, int J,
                        const float* gauss5, float threshold, int nms, float heat_std, uint16_t* staged, int Hp, int Wp, int Cp,
                        int top, int left, void* stream);

/* fp32 axis-angle -> rotation matrices, n rows.  Replaces smplx `lbs.batch_rodrigues`
 * (used by SMPL.forward when pose2rot=True and at models/humaniflow_model.py:299). */
int hf_rodrigues(const float* axis_angle, float* rotmats, int n, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Ancestor-conditioned SO(3) normalising flow over the kinematic tree.
 * Replaces: models/humaniflow_model.py:116-186 (context features), :286-320 (joint loop);
 * models/norm_flows/pyro_conditional_norm_flow.py:21-129; local_diffeo_transformed_distribution.py:
 * 72-142; transforms/{conditional_spline_coupling,scaled_radial_tanh,so3_exp,to}_transform.py;
 * utils/rigid_transform_utils.py:142-314; pyro 1.7.0 SplineCoupling / ConditionalDenseNN / Permute.
 * ---------------------------------------------------------------------------------------------- */
typedef struct hf_flow hf_flow_t;

typedef struct hf_flow_config {
    int num_joints;        /* body parts (23) */
    int feats_dim;         /* INPUT_SHAPE_GLOB_CAM_FEATS_DIM (256) */
    int context_dim;       /* NORM_FLOW.CONTEXT_DIM (64) */
    int num_transforms;    /* NORM_FLOW.NUM_TRANSFORMS (2) */
    int hidden[3];         /* NORM_FLOW.TRANSFORM_NN_HIDDEN_DIMS (64,32,32) */
    int num_bins;          /* NORM_FLOW.NUM_SPLINE_SEGMENTS (8) */
    int num_betas;         /* NUM_SMPL_BETAS (10) */
    float radius;          /* NORM_FLOW.COMPACT_SUPPORT_RADIUS (1.5*pi) */
    float base_std;        /* NORM_FLOW.BASE_DIST_STD (0.6) */
} hf_flow_config;

/* Host arrays.  ancestors: for joint j the list ancestors[anc_offsets[j] .. anc_offsets[j+1]) (nearest
 * first, root excluded; models/humaniflow_model.py:16-30).  beta_weight (feats_dim, num_betas) = the beta
 * columns of fc_input_shape_glob_cam_feats.weight.  ctx_weight[j] (context_dim, feats_dim + 9*a_j) and
 * ctx_bias[j] = fc_flow_context[j].  coupling layer l of transform t of joint j:
 * nn_weight[(j*num_transforms+t)*4 + l], nn_bias[...] with torch Linear layout (out,in)
 * (= pose_so3flow_transform_modules.{j*T+t}.nn.layers.{l}). */
int hf_flow_create(hf_flow_t** out, const hf_flow_config* cfg, const int* ancestors,
                   const int* anc_offsets, const float* beta_weight,
                   const float* const* ctx_weight, const float* const* ctx_bias,
                   const float* const* nn_weight, const float* const* nn_bias);
void hf_flow_destroy(hf_flow_t* h);

/* Hierarchical sampling down the kinematic tree, all joints in one launch.
 * img_base (B,feats_dim): pre-activation image-level features WITHOUT the beta term, i.e.
 *   W[:, feats|glob|cam] . [input_feats, vec(glob_R), cam] + bias   (models/humaniflow_model.py:133-148);
 * betas (R,num_betas) per-row shape; img_index (R) int32 row -> image;
 * base_noise (Rn,J,3): base-distribution draws ~ N(0, base_std^2) for rows [0,Rn) ("sample rows":
 *   fp32 flow -> fp64 exp map -> fp32 store, humaniflow_model.py:304-311);
 * rows [Rn,R) are "point-estimate rows": zero base sample, smplx fp32 Rodrigues (humaniflow_model.py:290-301),
 *   axis-angle written to axisangle_pe (R-Rn,J,3).
 * rotmats (R,J,3,3) fp32.  workspace: hf_flow_workspace_bytes(h, R) bytes of 16-byte aligned device scratch. */
size_t hf_flow_workspace_bytes(const hf_flow_t* h, int R);
int hf_flow_sample(const hf_flow_t* h, const float* img_base, const float* betas, const int* img_index,
                   const float* base_noise, int R, int Rn, float* rotmats, float* axisangle_pe,
                   void* workspace, size_t workspace_bytes, void* stream);

/* Contexts for teacher-forced log-likelihood (humaniflow_model.py:314-320): ancestors taken from
 * given rotations anc_rotmats (R,J,3,3) fp32.  ctx_out (R,J,context_dim). */
int hf_flow_context(const hf_flow_t* h, const float* img_base, const float* betas, const int* img_index,
                    const float* anc_rotmats, int R, float* ctx_out, void* stream);

/* log_prob on SO(3) (local_diffeo_transformed_distribution.py:84-142): for joints
 * [joint_first, joint_first+joint_count): ctx (R, ctx_row_stride floats per row; joint j's context at
 * offset j*context_dim), rot_f64 (R,joint_count,3,3) DOUBLE, out (R,joint_count) fp32. */
int hf_flow_log_prob(const hf_flow_t* h, const float* ctx, int ctx_row_stride, int joint_first,
                     int joint_count, const double* rot_f64, int R, float* out, void* stream);

/* log-density on the algebra so(3) (the `conditioned_pose_so3flow_dists_for_loglik` objects):
 * v (R,joint_count,3) fp32 -> out (R,joint_count) fp32. */
int hf_flow_algebra_log_prob(const hf_flow_t* h, const float* ctx, int ctx_row_stride, int joint_first,
                             int joint_count, const float* v, int R, float* out, void* stream);

/* Backward of the two log-densities w.r.t. their INPUTS, weights fixed (the reference: torch.autograd through
 * local_diffeo_transformed_distribution.py:84-142 and pyro's conditional spline coupling, as used by
 * optimise/optimise_humaniflow.py:96-114).  grad_out (R,joint_count) = d loss / d log_prob;
 * grad_ctx (R,joint_count,context_dim) = d loss / d (context row of that joint);
 * grad_rot (R,joint_count,3,3) fp64 = d loss / d target rotation (entries treated as free variables, like autograd does);
 * grad_v (R,joint_count,3) = d loss / d algebra vector. */
int hf_flow_log_prob_backward(const hf_flow_t* h, const float* ctx, int ctx_row_stride, int joint_first, int joint_count,
                              const double* rot_f64, const float* grad_out, int R, float* grad_ctx, double* grad_rot, void* stream);
int hf_flow_algebra_log_prob_backward(const hf_flow_t* h, const float* ctx, int ctx_row_stride, int joint_first, int joint_count,
                                      const float* v, const float* grad_out, int R, float* grad_ctx, float* grad_v, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Heads (models/humaniflow_model.py:232-258, 116-150) and small dense layers.
 * ---------------------------------------------------------------------------------------------- */
/* y (M,O) = act(x (M,K; row stride ldx) . W^T (O,K; row stride ldw) + b) ; act: 0 none, 1 ELU, 2 ReLU.
 * b may be NULL.  If accumulate != 0, y += instead of = (activation applied to the sum). */
int hf_linear(const float* x, int ldx, const float* W, int ldw, const float* b, float* y, int ldy,
              int M, int K, int O, int act, int accumulate, void* stream);
/* Same, with room for partial sums: a layer with few output tiles (M <= 32 rows x O outputs) is then split along K across CTAs
 * (slices summed in index order by a second small launch: deterministic).  workspace: hf_linear_workspace_bytes(M, K, O) bytes,
 * 16-byte aligned, private to the stream the call is issued on; NULL = hf_linear. */
size_t hf_linear_workspace_bytes(int M, int K, int O);
int hf_linear_ws(const float* x, int ldx, const float* W, int ldw, const float* b, float* y, int ldy,
                 int M, int K, int O, int act, int accumulate, void* workspace, size_t workspace_bytes, void* stream);

/* Head post-processing (models/humaniflow_model.py:237-258).  heads (B, 2*nb+9) = one Linear over the
 * concatenated [fc_shape | fc_glob | fc_cam] rows.  cam (B,3) = cam head + init_cam; glob6 (B,6) = glob head +
 * init_glob; shape_rows (B*N + B, nb): rows [0,B*N) = shape_mode + exp(shape_log_std) * shape_eps (shape_eps
 * (B,N,nb) ~ N(0,1); NULL = use the mode, `use_shape_mode_for_samples`), rows [B*N, B*N+B) = shape_mode.
 * Optional (NULL to skip): glob_R (B,3,3) = rot6d_to_rotmat(glob6) (utils/rigid_transform_utils.py:86-100) and shape_std (B,nb) =
 * exp(shape_log_std) (the scale of `shape_dist_for_loglik`), so that the step needs no separate launches for them. */
int hf_heads_finish(const float* heads, const float* init_glob, const float* init_cam, const float* shape_eps,
                    int B, int N, int nb, float* cam, float* glob6, float* shape_rows, float* glob_R, float* shape_std,
                    void* stream);

/* glob6 (B,6) -> rotation matrices (B,3,3): utils/rigid_transform_utils.py:86-100. */
int hf_rot6d_to_rotmat(const float* rot6d, float* rotmats, int n, void* stream);

/* ------------------------------------------------------------------------------------------------
 * ResNet encoder (models/resnet.py:202-217): a small op program over bf16 NHWC activation buffers.
 * ---------------------------------------------------------------------------------------------- */
typedef struct hf_encoder hf_encoder_t;

enum { HF_OP_CONV = 0, HF_OP_MAXPOOL3x3S2 = 1, HF_OP_GLOBAL_AVGPOOL = 2 };

typedef struct hf_enc_op {
    int kind;              /* HF_OP_* */
    int src, dst, res;     /* activation buffer ids; res = -1 for none (residual added before ReLU) */
    int cin, cout;         /* channels as stored (cin already padded for the stem) */
    int ksize, stride, pad;
    int relu;
    int weight_index;      /* into the weights/bias arrays given to hf_encoder_create */
    /* optional fused 1x1 branch (a residual block's downsample path, models/resnet.py:70-74,108-119): a second input buffer
     * src2 (-1 = none) with cin2 channels, sampled with stride2 at the output pixels; its weights are appended along K to
     * every weight row ([ksize*ksize*cin | cin2]) and its folded BatchNorm shift is added into the bias */
    int src2, cin2, stride2;
} hf_enc_op;

/* weights[i]: HOST bf16 (uint16 bit patterns) (cout, ksize, ksize, cin) with BatchNorm scale folded in;
 * bias[i]: HOST fp32 (cout) = folded BatchNorm shift.  in_channels = channels of the fp32 NCHW input,
 * stem_cin = padded channel count of the first conv. */
int hf_encoder_create(hf_encoder_t** out, const hf_enc_op* ops, int num_ops,
                      const uint16_t* const* weights, const float* const* bias, int num_weights,
                      int in_channels, int stem_cin, int feat_dim);
void hf_encoder_destroy(hf_encoder_t* h);
size_t hf_encoder_workspace_bytes(const hf_encoder_t* h, int B, int H, int W);
/* input (B,in_channels,H,W) fp32 NCHW -> feats (B,feat_dim) fp32.
 * Workspace contract: one stream at a time per handle; between calls with the same (B,H,W,workspace) the workspace
 * contents belong to the encoder (the zero border / channel padding of the staged stem input is written once, when the
 * launch plans are built).  A caller that recycled or overwrote the memory calls hf_encoder_invalidate first. */
int hf_encoder_forward(hf_encoder_t* h, const float* input, int B, int H, int W, float* feats,
                       void* workspace, size_t workspace_bytes, void* stream);
/* Same with a bf16 NCHW input (half the host->device bytes of the fp32 form; the encoder rounds to bf16 anyway). */
int hf_encoder_forward_bf16(hf_encoder_t* h, const uint16_t* input, int B, int H, int W, float* feats,
                            void* workspace, size_t workspace_bytes, void* stream);
/* The staged stem input inside the workspace: bf16 NHWC (B, Hp, Wp, Cp), image at (top, left) inside a zero border.
 * dims receives {Hp, Wp, Cp, top, left}.  A producer kernel (hf_proxy_rep_staged) fills the interior, then
 * hf_encoder_forward_staged runs the trunk without the NCHW->NHWC conversion pass. */
int hf_encoder_stem_input(hf_encoder_t* h, int B, int H, int W, void* workspace, size_t workspace_bytes, void* stream,
                          void** staged, int* dims);
int hf_encoder_forward_staged(hf_encoder_t* h, int B, int H, int W, float* feats, void* workspace,
                              size_t workspace_bytes, void* stream);
/* Forget the cached launch plans: the next call re-builds them and re-zeroes the staged input's padding. */
int hf_encoder_invalidate(hf_encoder_t* h);
/* Channels per pixel of the staged stem input chosen at the first forward: 24 (compact stem, sliding-window
 * tensor map) or 32 (pixel-pair layout, used when the driver refuses the overlapping map or in_channels > 24). */
int hf_encoder_stem_channels(const hf_encoder_t* h);
/* impl: 0 = tcgen05/TMA implicit GEMM (product path), 1 = SIMT direct convolution (debug cross-check). */
int hf_encoder_set_impl(hf_encoder_t* h, int impl);

/* Bring-up aid: run the program up to and including op `op_index`, copy that op's bf16 NHWC output to `out`
 * (device) and write its {H, W, C} to dims (host). */
int hf_encoder_debug_op_output(hf_encoder_t* h, int op_index, const float* input, int B, int H, int W,
                               void* workspace, size_t workspace_bytes, uint16_t* out, size_t out_bytes,
                               int* dims, float* feats, void* stream);

/* Timing aid: with HF_CONV_DBG bit 3 set, CTA 0 of every conv launch records %globaltimer at ten milestones; this
 * synchronises the device and returns the stamps of the most recent launch (16 values, ns). */
int hf_debug_conv_stamps(unsigned long long* out16);

/* Single convolution on bf16 NHWC tensors (unit-test entry point for the conv kernels).
 * x (B,H,W,cin) bf16; w (cout,k,k,cin) bf16; bias (cout) fp32; res (B,Ho,Wo,cout) bf16 or NULL;
 * y (B,Ho,Wo,cout) bf16. */
int hf_conv2d_nhwc(const uint16_t* x, const uint16_t* w, const float* bias, const uint16_t* res,
                   uint16_t* y, int B, int H, int W, int cin, int cout, int ksize, int stride, int pad,
                   int relu, int impl, void* stream);

/* Same, plus the fused 1x1 branch of hf_enc_op.src2 (a bottleneck block's downsample path, models/resnet.py:108-118):
 * x2 (B,H2,W2,cin2) bf16 sampled with stride2 at the output pixels; w rows are [k*k*cin | cin2]; bias is the sum of both
 * folded shifts.  Both convolutions accumulate in fp32 in one accumulator; the branch output is never rounded to bf16. */
int hf_conv2d_nhwc_branch(const uint16_t* x, const uint16_t* w, const float* bias, const uint16_t* res,
                          uint16_t* y, int B, int H, int W, int cin, int cout, int ksize, int stride, int pad,
                          int relu, const uint16_t* x2, int H2, int W2, int cin2, int stride2, int impl, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* HUMANIFLOW_B200_H */
