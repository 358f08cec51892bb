/* libchd -- C ABI of the B200-native batched physics-based trajectory optimiser and foot-contact
 * classifier (drop-in for the hot path of davrempe/contact-human-dynamics).
 *
 * Every entry point is plain C: pointers + sizes, int return (0 = ok, negative = error), no exceptions
 * cross the boundary, no ownership transfer unless stated.  Host pointers are marked [host], device
 * pointers [device].  All floating point is IEEE fp64 unless the name says otherwise.
 *
 * Reference interfaces replaced (paths relative to the reference repo):
 *   chd_phys_batch_create   <- towr_phys_optim/phys_optim.cpp:380-552  (Read*Info + NlpFormulation::
 *                              GetVariableSets / GetConstraints, src/nlp_formulation.cpp:79-360)
 *   chd_phys_solve          <- phys_optim.cpp:554-749  (the five/six staged ifopt::IpoptSolver::Solve calls)
 *   chd_phys_eval           <- ifopt ConstraintSet::GetValues / FillJacobianBlock and CostTerm::GetCost /
 *                              FillJacobianBlock as implemented in towr_phys_optim/src/{constraints,costs,models}
 *   chd_phys_sample         <- SaveSolution, phys_optim.cpp:63-143
 *   chd_contact_*           <- src/contact_learning/test.py:51-152 (val_full_video) +
 *                              src/contact_learning/models/openpose_only.py:29-78
 */
#ifndef CHD_H_
#define CHD_H_
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* One sequence's inputs, i.e. what phys_optim reads from skel_info.txt, motion_info.txt,
 * terrain_info.txt and contact_info.txt (phys_optim.cpp:155-267).  End-effector order is the solver's:
 * L toe, R toe, L heel, R heel (phys_optim.cpp:491-513).  All pointers [host]. */
typedef struct chd_phys_problem {
  int32_t n_frames;            /* --nframes */
  int32_t n_ee;                /* 2 (toes only) or 4 (reference: always 4, phys_optim.cpp:432) */
  double dt;
  const double* hip_left;      /* n_frames x 3 */
  const double* hip_right;     /* n_frames x 3 */
  double max_leg_length, max_heel_length, heel_dist, body_mass;
  const double* inertia;       /* n_frames x 6: Ixx Iyy Izz Ixy Ixz Iyz */
  const double* base_lin;      /* n_frames x 3 */
  const double* base_ang;      /* n_frames x 3 */
  const double* ee_pos;        /* n_ee x n_frames x 3 */
  double floor_normal[3], floor_point[3];
  const int32_t* ee_start_contact; /* n_ee */
  const int32_t* ee_n_phases;      /* n_ee */
  const double* ee_durations;      /* concatenated, sum(ee_n_phases) */
} chd_phys_problem;

/* gflags of phys_optim.cpp:27-31 */
typedef struct chd_phys_weights {
  double w_com_lin, w_com_ang, w_ee, w_smooth, w_dur;
} chd_phys_weights;

typedef struct chd_phys_batch chd_phys_batch;

/* Stage ids (phys_optim.cpp:554-749): 0 = 1.1, 1 = 1.2, 2 = 2.1, 3 = 2.2, 4 = 3, 5 = 4. */
enum { CHD_ST_11 = 0, CHD_ST_12 = 1, CHD_ST_21 = 2, CHD_ST_22 = 3, CHD_ST_3 = 4, CHD_ST_4 = 5 };

/* Sizes of the padded batch (strides of every per-sequence array below). */
typedef struct chd_phys_dims {
  int32_t batch, n_max, m_max, slots_max, n_splines, p_max, sets_max, na_max, nb_max, w_max, frames_out_max;
} chd_phys_dims;

/* Builds the NLP layouts of `batch` sequences on the host and uploads them to the current CUDA device.
 * device < 0 keeps the current device. */
int chd_phys_batch_create(const chd_phys_problem* problems, int32_t batch, const chd_phys_weights* weights,
                          int32_t device, chd_phys_batch** out);
void chd_phys_batch_destroy(chd_phys_batch* b);
int chd_phys_get_dims(const chd_phys_batch* b, chd_phys_dims* dims);
/* Per-sequence sizes: n (variables: node values, then the P-1 free phase durations of every foot), m (master rows),
 * nslots, Na, nb (border unknowns incl. the durations), w -- six int32 per sequence [host]. */
int chd_phys_get_sizes(const chd_phys_batch* b, int32_t* sizes6);
/* What the fixed-duration stages (all but stage 3) work with: border unknowns without the switch times, half bandwidth
 * of the static pattern, and the number of phase-duration variables (0: stage 3 not available) -- three int32 per
 * sequence [host]. */
int chd_phys_get_sizes_fixed(const chd_phys_batch* b, int32_t* sizes3);

/* Current iterate x: batch x n_max doubles.  [host] copies (synchronous). */
int chd_phys_get_x(const chd_phys_batch* b, double* x_host);
int chd_phys_set_x(chd_phys_batch* b, const double* x_host);

/* Function-level evaluation at the current x for `stage` (selects active rows and cost weights):
 *   cost[batch], grad[batch x n_max], g[batch x m_max] (master row order, inactive rows = 0),
 *   jac_vals[batch x slots_max] (block-row slots; columns via chd_phys_get_layout).  Any output may be NULL.
 * All outputs [host]. */
int chd_phys_eval(chd_phys_batch* b, int32_t stage, double* cost, double* grad, double* g, double* jac_vals);

/* Static layout tables, [host] outputs, any may be NULL:
 *   ent_ptr[batch x (m_max+1)], ent_col[batch x slots_max] (variable index or -1),
 *   row_lo / row_hi [batch x m_max], row_set[batch x m_max] (constraint-set type of each row),
 *   var_kkt[batch x n_max], row_kkt[batch x m_max] (KKT ordering; -1 = not an unknown). */
int chd_phys_get_layout(const chd_phys_batch* b, int32_t* ent_ptr, int32_t* ent_col, double* row_lo, double* row_hi,
                        int32_t* row_set, int32_t* var_kkt, int32_t* row_kkt);
/* Column-oriented view of the same Jacobian slots (what ifopt keeps as a column-compressed Eigen matrix,
 * ifopt::Composite::GetJacobian): ent_row[batch x slots_max] (row of every slot), col_ptr[batch x (n_max+1)],
 * col_ent[batch x slots_max] (slots grouped by variable).  [host] outputs, any may be NULL. */
int chd_phys_get_slot_index(const chd_phys_batch* b, int32_t* ent_row, int32_t* col_ptr, int32_t* col_ent);

/* Jacobian slot columns as they stand on the device: equal to chd_phys_get_layout's ent_col until stage 3 moves a
 * phase duration; from then on the polynomial active at a sample time is located at run time and the node / switch-time
 * columns of the time-located rows are rewritten by every evaluation.  Column index of a duration variable d_k = the
 * switch time tau_k = d_0 + ... + d_k the solver works with (derivatives are with respect to tau). [host] output. */
int chd_phys_get_ent_col(const chd_phys_batch* b, int32_t* ent_col);
/* Interior-point state of the last solved stage, master row order, batch x m_max each [host], any may be NULL:
 * constraint multipliers y, bound multipliers zL / zU of the slacks, slacks s (all of the scaled problem
 * min obj_scale * f  s.t.  row_scale * g(x) - s = 0), the row scaling and obj_scale[batch]
 * (what IpoptCalculatedQuantities reports as the scaled multipliers; unscaled y = y * row_scale / obj_scale). */
int chd_phys_get_duals(const chd_phys_batch* b, double* y, double* zL, double* zU, double* s, double* row_scale,
                       double* obj_scale);
/* Residuals of every stage of the last chd_phys_solve / chd_phys_solve_stage: stats [host] 6 x batch x 4 =
 * objective, scaled NLP error E0 (IPOPT's overall error, tol 1e-3), unscaled max constraint violation
 * (constr_viol_tol 1e-4), unscaled dual infeasibility. */
int chd_phys_stage_stats(const chd_phys_batch* b, double* stats);

/* Runs the interior-point solve of one stage for every sequence (warm start from the current x).
 * status[batch] [host]: 0 = converged (IPOPT "Solve_Succeeded" test), -1 = iteration cap, -2 = numerical failure.
 * iters[batch] [host], stats[batch x 8] [host]: f, E0, unscaled constraint violation, unscaled dual inf,
 * unscaled complementarity, mu, delta_w, #line-search failures.  Outputs may be NULL. */
int chd_phys_solve_stage(chd_phys_batch* b, int32_t stage, int32_t max_iter, int32_t* status, int32_t* iters,
                         double* stats);

/* Whole staged schedule of phys_optim.cpp:554-749 for the batch.  `samples` [host]:
 * 3 x batch x frames_out_max x (6 + 7*n_ee_max) doubles = the three SaveSolution snapshots
 * (no_dynamics, dynamics, durations), frames_out[batch] [host] = frames per sequence,
 * success[batch x 2] [host] = (dynamics_succeed, durations_succeed) of success_log.txt.
 * stage_status [host, 6 x batch] / stage_iters [host, 6 x batch] may be NULL; status -9 = stage not run (stage 4 after a
 * successful stage 3, phys_optim.cpp:713), -3 = stage 3 not attempted (more than 96 phase durations). */
int chd_phys_solve(chd_phys_batch* b, double* samples, int32_t* frames_out, int32_t* success,
                   int32_t* stage_status, int32_t* stage_iters);

/* SaveSolution sampling of the current x: out [host] batch x frames_out_max x (6+7*n_ee_max). */
int chd_phys_sample(chd_phys_batch* b, double* out, int32_t* frames_out);

/* Device-resident variant of the sampler for the multi-GPU gather: writes into a caller-provided
 * [device] buffer (e.g. an NCCL send buffer) on the given stream (cudaStream_t passed as void*). */
int chd_phys_sample_device(chd_phys_batch* b, double* out_device, void* stream);

/* Number of kernels launched by this batch so far. */
int64_t chd_phys_launch_count(const chd_phys_batch* b);
/* Bytes copied host -> device by chd_phys_batch_create (problem data + layout tables). */
int64_t chd_phys_h2d_bytes(const chd_phys_batch* b);
/* Restores the initial point of nlp_formulation.cpp:106-203 on the device (no host traffic). */
int chd_phys_reset(chd_phys_batch* b);

/* Per-kernel accumulated CUDA-event time (ms) and launch counts since the last reset:
 * names: 0 eval, 1 kkt (assemble+factor+solve), 2 linesearch, 3 init, 4 sample.  [host] arrays of 8. */
int chd_phys_kernel_times(chd_phys_batch* b, double* ms8, int64_t* launches8, int reset);
int chd_phys_set_timing(chd_phys_batch* b, int enable);

/* ---------------------------------------------------------------------------------------------------------
 * Foot-contact classifier (src/contact_learning/test.py:51-152 val_full_video + models/openpose_only.py:29-78).
 * weights: the five Linear weights in torch layout [out][in], concatenated in layer order
 *          (1024x351, 512x1024, 128x512, 32x128, 20x32); biases concatenated likewise;
 * bn: for each of the four BatchNorm1d layers gamma, beta, running_mean, running_var (4 x width floats), concatenated.
 * All [host].  Keys of the reference state_dict: model.{0,3,6,10,13}.{weight,bias}, model.{1,4,7,11}.*. */
typedef struct chd_contact_net chd_contact_net;
int chd_contact_create(const float* weights, const float* biases, const float* bn, float bn_eps, int32_t device,
                       chd_contact_net** out);
void chd_contact_destroy(chd_contact_net* net);
/* frames [host]: V x Fmax x 25 x 3 doubles = keypoints after the dataset's preprocessing (pad to the longest video,
 * scale to 1280 wide, low-confidence interpolation, division by 200.416..., real_video_dataset.py:132-163);
 * seq_lens [host] V; labels [host] V x Fmax x 4 int64 (columns L heel, L toe, R heel, R toe; rows >= seq_len are 0);
 * logits [host, optional] V x (Fmax-8) x 20; min_abs_logit [host, optional]: smallest |logit| that entered a vote. */
int chd_contact_forward(chd_contact_net* net, const double* frames, int32_t V, int32_t Fmax, const int32_t* seq_lens,
                        int64_t* labels, float* logits, float* min_abs_logit);
/* Same with every buffer already on the device (stream = cudaStream_t as void*, NULL = the net's own stream); all
 * device buffers, logits_dev and min_abs_dev included, are required (they are workspace of the vote kernel). */
int chd_contact_forward_device(chd_contact_net* net, const double* frames_dev, int32_t V, int32_t Fmax,
                               const int32_t* seq_lens_dev, int64_t* labels_dev, float* logits_dev, float* min_abs_dev,
                               void* stream);
/* Dataset preprocessing on the device (RealVideoDataset.__init__, real_video_dataset.py:132-163, and
 * process_openpose_data, openpose_dataset.py:49-121): raw [host] = concatenated OpenPose keypoints (sum F) x 25 x 3
 * doubles [x, y, confidence] as load_keypoint_dir returns them, seq_offsets [host] V+1 frame offsets, dim_w = width of
 * the source video (reference default 1920).  frames_out [host] V x Fmax x 25 x 3 (Fmax = longest video),
 * seq_lens_out [host, optional] V.  Bit identical to the reference's numpy result. */
int chd_contact_preprocess(chd_contact_net* net, const double* raw, const int32_t* seq_offsets, int32_t V, int32_t dim_w,
                           double* frames_out, int32_t* seq_lens_out);
/* test.py --full-video --save-contacts --real-data in one call: raw keypoints in, foot_contacts rows out.
 * labels_out [host] (sum F) x 4 int64 (columns L heel, L toe, R heel, R toe; the rows every video's foot_contacts.npy
 * holds, concatenated).  raw may be page-locked: the upload is asynchronous on the net's stream. */
int chd_contact_detect(chd_contact_net* net, const double* raw, const int32_t* seq_offsets, int32_t V, int32_t dim_w,
                       int64_t* labels_out, float* min_abs_logit);
int64_t chd_contact_launch_count(const chd_contact_net* net);

/* OpenPose keypoint ingestion on the host threads (replaces the serial per-file json.load of
 * src/utils/openpose_utils.py:48-76 load_keypoint_file / load_keypoint_dir): n_files `*_keypoints.json` files ->
 * out [host] n_files x num_joints x 3 fp64 (x, y, confidence of the FIRST person's "pose_keypoints_2d"; zeros for a frame
 * without people), bit-identical to the reference's arrays (strtod).  n_threads <= 0: all hardware threads.
 * Returns 0, -1 bad argument, -2 unreadable file, -3 malformed file / keypoint count != 3 * num_joints. */
int chd_openpose_load(const char* const* paths, int32_t n_files, int32_t num_joints, double* out, int32_t n_threads);

const char* chd_version(void);

/* Measurement helper (no reference counterpart): sustained fp64 throughput of the current device in GFLOP/s, for the
 * scalar FMA pipe (DFMA) and for the fp64 tensor-core instruction the KKT kernel uses (mma.sync m8n8k4, DMMA);
 * MEASURED_PEAKS.json carries no fp64 figure.  Either output may be NULL. */
int chd_measure_fp64_peak(double* dfma_gflops, double* dmma_gflops);

#ifdef __cplusplus
}
#endif
#endif /* CHD_H_ */
