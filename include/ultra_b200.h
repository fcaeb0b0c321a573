/*
 * ultra_b200.h - C ABI of the B200-native (sm_100a) training hot path for ULTRA (unbiased learning to rank).
 *
 * This header is the drop-in boundary.  Every entry point replaces a chain of ATen ops + autograd in the
 * reference (ULTR-Community/ULTRA_pytorch; paths below are relative to the reference root).  The reference has
 * no FFI of its own (it is pure Python on torch ops), so the "binding a maintainer would add" is a ctypes stub;
 * see INTEGRATION.md and ultra_pytorch_b200/_capi.py.
 *
 * Conventions
 *   - all pointers are DEVICE pointers unless the name ends in _host; no torch types anywhere;
 *   - every launcher takes the cudaStream_t (as void*) to launch on, never synchronises, never allocates:
 *     scratch comes from the caller's `workspace` (size from the matching *_workspace_bytes call; the caller
 *     must zero it ONCE after allocation - the kernels keep their ticket counters at zero between calls);
 *   - return value 0 = ok; otherwise an error code, message via ub200_last_error();
 *   - all floating point is IEEE fp32 ("f32"); reductions are deterministic (fixed order, no float atomics);
 *   - scores / labels / dscores are row-major [B, L] (one ranked list per row); the DNN processes M = L*B
 *     rows in the reference's position-major order (row = l*B + b, ultra/ranking_model/DNN.py:72,
 *     ultra/learning_algorithm/base_algorithm.py:132).
 *
 * Flat parameter layout (`params`, `grads`, optimizer state): for j = 0..n_hidden (n_hidden+1 linear layers,
 * the last one maps to 1 output), in the reference's state_dict order (DNN.py:43-55):
 *     layer_norm{j}.weight [K_j] | layer_norm{j}.bias [K_j] | linear{j}.weight [N_j, K_j] | linear{j}.bias [N_j]
 * with K_0 = F, K_j = hidden[j-1], N_j = hidden[j] (N_last = 1).
 */
#ifndef ULTRA_B200_H_
#define ULTRA_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define UB200_MAX_LAYERS 8

#if defined(__GNUC__)
#define UB200_API __attribute__((visibility("default")))
#else
#define UB200_API
#endif

/* ---- misc ------------------------------------------------------------------------------------------- */
UB200_API const char* ub200_last_error(void);
UB200_API int ub200_abi_version(void);
/* number of kernel launches issued through this library since it was loaded (bench.py's gpu_launches) */
UB200_API unsigned long long ub200_launch_count(void);
/* K1 GEMM engine, a bit mask: 1 = forward, 2 = data-gradient, 4 = weight-gradient GEMMs of the hidden layers whose
 * shapes qualify (K % 4 == 0, N % 64 == 0) run on the tcgen05 tensor cores with 3xTF32 error compensation; cleared
* bits use the CUDA-core fp32 kernels; 8 = run the whole forward pass as ONE fused kernel when the net fits the
 * tensor-memory plan (N_0 <= 256, N_j <= 128).  Default 15 (env UB200_TC overrides).  Both engines meet the same parity bound;
 * the switch exists for A/B tests and profiling.  Returns the previous mask. */
UB200_API int ub200_set_tc_mode(int mode);
/* number of parameters of the DNN ranker for (F, hidden[]) in the flat layout above */
UB200_API size_t ub200_mlp_param_count(int F, const int* hidden, int n_hidden);

/* ---- host side of the boundary: pack one input_feed into a (pinned) staging buffer ---------------------------
 * Replaces the numpy work of create_input_feed / get_ranking_scores (base_algorithm.py:148-152, 176-186) and the
 * f64 -> f32 cast (DNN.py:72-73).  HOST pointers.  Layout written to dst (and expected on the device after ONE
 * H2D copy of ub200_feed_bytes bytes):
 *     docid int32 [L, B] (position-major) | labels f32 [B, L] | pad to 256 B | feats f32 [n_docs + 1, F] (PAD row = 0)
 * feats_host: f64 [n_docs, F] row-major; docid_cols_host / label_cols_host: L pointers to f32 [B] (the feed's
 * "docid_input{l}" / "label{l}" arrays).  Multi-threaded (n_threads, persistent worker pool). */
UB200_API size_t ub200_feed_bytes(int n_docs, int F, int L, int B);
UB200_API int ub200_pack_feed_host(const double* feats_host, int n_docs, int F, const float* const* docid_cols_host,
                         const float* const* label_cols_host, int L, int B, void* dst_host, size_t dst_bytes,
                         int n_threads);

/* the same packing in two pieces, for a pipelined pack (convert a chunk of rows, start its H2D copy, convert the
 * next chunk): ids/labels part of the layout above, and a plain multi-threaded f64 -> f32 conversion.
 * Every packer validates the document ids: an id outside [0, max_id] (max_id = n_docs, the PAD row) is error 5 - the
 * kernels gather feats[id] unchecked, where the reference's np.take raises IndexError (base_algorithm.py:150). */
UB200_API int ub200_pack_ids_host(const float* const* docid_cols_host, const float* const* label_cols_host, int L, int B,
                        int max_id, void* dst_host, size_t dst_bytes);
UB200_API int ub200_convert_f64_f32_host(const double* src_host, float* dst_host, size_t n, int n_threads);

/* ub200_pack_feed_host into the PINNED buffer `pinned_host` + the H2D copy to `device_dst` on `stream`, pipelined: the
 * feature rows are converted on a persistent pool of n_threads host threads (the caller included) and the copy of each of
 * the n_groups groups of rows is enqueued as soon as the group is converted.  Returns when the last copy has been
 * enqueued; kernels launched on `stream` afterwards see the complete feed (the one H2D transfer of a train() call). */
UB200_API int ub200_stage_feed(const double* feats_host, int n_docs, int F, const float* const* docid_cols_host,
                     const float* const* label_cols_host, int L, int B, void* pinned_host, size_t pinned_bytes,
                     void* device_dst, int n_threads, int n_groups, void* stream);

/* Double-buffered staging (the host side of train(), base_algorithm.py:169-186 + DNN.py:72-75, overlapped with the
 * device work of the previous step): same packing as ub200_stage_feed, but the copies run on `copy_stream`.  The caller
 * alternates between two (pinned, device) buffer pairs; `slot_free` (cudaEvent_t, may be NULL) was recorded on
 * `compute_stream` behind the last kernel that reads this device buffer and is waited for before the first copy;
 * `ready` (cudaEvent_t) is recorded behind the last copy and `compute_stream` is made to wait for it.
 * ub200_stage_ids_pipelined is the variant for a data set resident in HBM (ids / labels block only).
 * ub200_event_record = cudaEventRecord(event, stream). */
UB200_API int ub200_stage_feed_pipelined(const double* feats_host, int n_docs, int F, const float* const* docid_cols_host,
                     const float* const* label_cols_host, int L, int B, void* pinned_host, size_t pinned_bytes,
                     void* device_dst, int n_threads, int n_groups, void* copy_stream, void* compute_stream,
                     void* slot_free, void* ready);
UB200_API int ub200_stage_ids_pipelined(const float* const* docid_cols_host, const float* const* label_cols_host, int L,
                     int B, int max_id, void* pinned_host, size_t pinned_bytes, void* device_dst, void* copy_stream,
                     void* compute_stream, void* slot_free, void* ready);
UB200_API int ub200_event_record(void* event, void* stream);
/* diagnostics: wall-clock stamps (ns since entry) of the last ub200_stage_feed* call - [0] ids packed, [2 + k] copy of
 * group k issued, [31] return */
UB200_API int ub200_stage_timeline(long long* out32);

/* ---- K1: DNN ranker forward / backward ----------------------------------------------------------------
 * Replaces: host gather base_algorithm.py:148-152, cat + f64->f32 cast DNN.py:72-73, the nn.Sequential of
 * [LayerNorm -> Linear -> ELU] x n_hidden + LayerNorm -> Linear(1) DNN.py:43-55,77, split/cat DNN.py:87-88 +
 * base_algorithm.py:132, and loss.backward() through all of it base_algorithm.py:222.
 *
 *   feats   [n_feat_rows, F] f32, row n_docs (the last one the caller uses) is the all-zero PAD row
 *   docid   [L*B] int32, position-major (docid[l*B+b]) row index into feats; NULL = identity (row r reads feats[r])
 *   scores  [B, L] f32 out (scores[b*L+l])
 *   training != 0 keeps the activations + LayerNorm statistics in `workspace` for ub200_mlp_backward
 */
UB200_API size_t ub200_mlp_workspace_bytes(int L, int B, int F, const int* hidden, int n_hidden, int training);
UB200_API int ub200_mlp_forward(const float* feats, const int32_t* docid, int L, int B, int F,
                      const int* hidden, int n_hidden, const float* params,
                      float* scores, void* workspace, size_t workspace_bytes, int training, void* stream);
/* dscores [B, L]; grads (flat layout) is overwritten.  Must follow ub200_mlp_forward(training=1) on the same
 * workspace, inputs and parameters. */
UB200_API int ub200_mlp_backward(const float* feats, const int32_t* docid, int L, int B, int F,
                       const int* hidden, int n_hidden, const float* params, const float* dscores,
                       float* grads, void* workspace, size_t workspace_bytes, void* stream);

/* The same two entry points for the other hidden-layer activations of the reference ranker (hparam activation_func,
 * base_ranking_model.py:63-69): activation = 0 elu (identical to the calls above), 1 relu, 2 selu, 3 tanh, 4 sigmoid.
 * ELU runs on the tensor cores; the others through the fp32 CUDA-core kernels. */
UB200_API int ub200_mlp_forward_act(const float* feats, const int32_t* docid, int L, int B, int F, const int* hidden,
                          int n_hidden, int activation, const float* params, float* scores, void* workspace,
                          size_t workspace_bytes, int training, void* stream);
UB200_API int ub200_mlp_backward_act(const float* feats, const int32_t* docid, int L, int B, int F, const int* hidden,
                           int n_hidden, int activation, const float* params, const float* dscores, float* grads,
                           void* workspace, size_t workspace_bytes, void* stream);

/* ---- K2: listwise softmax cross-entropy, forward + gradient ----------------------------------------------
 * Replaces BaseAlgorithm.softmax_loss base_algorithm.py:309-330 (+ :18-30) and its autograd, and the pure-Python
 * IPW weight loop ipw_rank.py:116-128 -> propensity_estimator.py:22-42.
 *   weight_mode 0: pw = 1 (NavieAlgorithm, navie_algorithm.py:105-106)
 *   weight_mode 1: pw_bl = table[min(l, table_len-1)] if labels_bl > 0 else 0 (IPWrank)
 * Outputs are UN-NORMALISED so that data-parallel ranks can sum them before dividing:
 *   dscores[b,l] = softmax(s_b)_l * W_b * sum_l(d_bl) - w_bl        (true gradient = dscores / sums[1])
 *   sums[0] = sum_b l_b,  sums[1] = sum_bl w_bl                      (loss = sums[0] / sums[1])
 */
UB200_API size_t ub200_loss_workspace_bytes(int B, int L);
UB200_API int ub200_softmax_ce(const float* scores, const float* labels, int B, int L, int weight_mode,
                     const float* table, int table_len, float* dscores, float* sums,
                     void* workspace, size_t workspace_bytes, void* stream);

/* DLA: both softmax losses of dla.py:196-224 in one pass, plus the DenoisingNet (dla.py:24-48) forward/backward
 * and get_normalized_weights (dla.py:287-301).
 *   prop_w [L] = DenoisingNet.linear_layer.weight[0,:], prop_b [1] = its bias
 *   dscores  [B,L]  un-normalised d(rank_loss)/d(scores)                 (divide by sums[1])
 *   dprop    [L+1]  un-normalised d(exam_loss)/d(prop_w[0..L-1], prop_b) (divide by sums[3])
 *   sums[4] = { rank num, rank den, exam num, exam den }; loss = sums[2]/sums[3] + ranker_loss_weight*sums[0]/sums[1]
 */
UB200_API int ub200_dla_loss(const float* scores, const float* clicks, int B, int L, const float* prop_w, const float* prop_b,
                   float* dscores, float* dprop, float* sums, void* workspace, size_t workspace_bytes, void* stream);

/* ---- K3: pairwise losses -----------------------------------------------------------------------------------
 * LambdaRank (lambda_rank.py:116-135, dcg :247-266, compute_delta_ndcg :268-291): i, j = predicted-rank positions
 * (stable descending sort), BCE-with-logits applied to p_ij, batch-global natural-log IDCG.
 *   out[0..L)   = T+_i = sum_j pair_ij / t-_j          out[L..2L) = T-_j = sum_i pair_ij / t+_i
 *   out[2L]     = loss,   out[2L+1] = idcg             ALL computed with idcg = 1:
 *   true loss = out[2L]/idcg, true gradient = dscores/idcg (T+/T- only enter the EM update as ratios).
 * PairDebias (pairwise_debias.py:142-157, base_algorithm.py:228-248): i, j = display positions; outputs are
 * WITHOUT the reference's x batch_size factor (the host multiplies by the global batch size): out has 2L+1 floats.
 */
UB200_API size_t ub200_pair_workspace_bytes(int B, int L);
UB200_API int ub200_lambdarank(const float* scores, const float* labels, int B, int L, float sigma,
                     const float* t_plus, const float* t_minus, float* dscores, float* out,
                     void* workspace, size_t workspace_bytes, void* stream);
UB200_API int ub200_pairdebias(const float* scores, const float* clicks, int B, int L,
                     const float* t_plus, const float* t_minus, float* dscores, float* out,
                     void* workspace, size_t workspace_bytes, void* stream);
/* t <- (1-em_step)*t + em_step * (T_i / T_0)^(1/(reg_p+1)) for t+ (T = out[0..L)) and t- (T = out[L..2L));
 * safe_div != 0 uses the reference's _safe_div (LambdaRank, lambda_rank.py:136-140), 0 a plain division
 * (PairDebias, pairwise_debias.py:159-163). */
/* PRSrank (prs_rank.py:94-151, N4): the same pair kernel on predicted ranks; pair (r ranked above s) weighted by
 * delta-NDCG * ipw_r / ipw_s with ipw = ipw_table[min(display position, table_len - 1)]; loss = weighted binary
 * cross-entropy of sigmoid(sigma (s_r - s_s)) against (1 + clamp(y_r - y_s)) / 2 (torch's -100 log clamp).  out as for
 * ub200_lambdarank: [2L] = un-normalised loss, [2L+1] = batch IDCG partial (T+/T- slots are zero). */
UB200_API int ub200_prsrank(const float* scores, const float* labels, int B, int L, float sigma, const float* ipw_table,
                  int table_len, float* dscores_unnorm, float* out, void* workspace, size_t workspace_bytes,
                  void* stream);
UB200_API int ub200_em_update(float* t_plus, float* t_minus, const float* out, int L, float em_step, float reg_p,
                    int safe_div, void* stream);

/* ---- optimizer: clip_grad_norm_ + Adagrad / SGD on a flat buffer ---------------------------------------------
 * Replaces BaseAlgorithm.opt_step base_algorithm.py:208-226 (clip_grad_norm_ + torch.optim.Adagrad/SGD.step) and
 * DLA.separate_gradient_update dla.py:141-166.
 *   g = grads * scale_const / (*den)            (den may be NULL = 1; this is where the loss normaliser is applied)
 *   norm = ||g||_2 ; coef = min(max_norm / (norm + 1e-6), 1) if max_norm > 0 else 1 ; g *= coef
 *   mode 0: Adagrad with persistent `state_sum`      (lr_decay 0, eps 1e-10, initial accumulator 0)
 *   mode 1: Adagrad re-created every step (DLA quirk, dla.py:153-154): the accumulator starts from 0 each call
 *   mode 2: SGD (grad_strategy == 'sgd')
 * grads is overwritten with the clipped gradient (what p.grad holds after the reference's opt_step);
 * norm_out[0] receives the pre-clip norm.
 */
/* Early read-back of a step's scalars (the reference's loss.item(), e.g. ipw_rank.py:181-182): copies n <= 32 floats
 * from src (device) to host_dst[(c & 1) * 32 ...] and then stores the launch count c (dev_counter, device,
 * zero-initialised, incremented by every launch) into host_seq; host_dst (2 x 32 floats) / host_seq are MAPPED PINNED
 * host memory (UVA: the host pointer is valid on the device).  The
 * host polls host_seq instead of synchronising the stream, so train() returns while the backward pass and the optimizer
 * step of the batch are still running; everything later on the stream stays ordered behind them. */
UB200_API int ub200_publish(const float* src, int n, float* host_dst, unsigned int* host_seq, unsigned int* dev_counter,
                  void* stream);
/* L2 regularisation (hparam l2_loss of NavieAlgorithm / IPWrank / DLA / PairDebias / RegressionEM, e.g.
 * ipw_rank.py:153-157 + base_algorithm.py:332-333): grads[i] += l2 * params[i] * f with f = den[0] (device, the
 * normaliser ub200_clip_update divides by) or `factor` when den is NULL; half_sumsq[0] = sum(params^2) / 2. */
UB200_API int ub200_l2_term(const float* params, float* grads, size_t n, float l2, const float* den, float factor,
                  float* half_sumsq, void* stream);
UB200_API size_t ub200_opt_workspace_bytes(size_t n);
UB200_API int ub200_clip_update(float* params, float* grads, float* state_sum, size_t n, const float* den, float scale_const,
                      float max_norm, float lr, int mode, float* norm_out,
                      void* workspace, size_t workspace_bytes, void* stream);

/* ---- N3: Plackett-Luce re-ranking for the stochastic online simulation ------------------------------------------
 * Replaces the per-query host loop of StochasticOnlineSimulationFeed.simulate_clicks_online
 * (stochastic_online_simulation_feed.py:100-177: np.random.choice(list_len, replace=False, p=softmax(tau * scores))).
 * scores [B, L] f32; docid [L, B] i32 position-major with PAD id == n_docs, or NULL (every position valid);
 * perm [B, L] i32: perm[b][r] = original position of the document ranked r-th; positions at or behind the list's
 * length (1 + last real document) stay in place.  Gumbel-top-k with Philox4x32-10 noise keyed by (seed, offset, b, l):
 * the same arguments give the same permutations.  Distribution-level parity with the reference (tested). */
UB200_API int ub200_pl_sample(const float* scores, const int32_t* docid, int n_docs, int B, int L, float tau,
                    unsigned long long seed, unsigned long long offset, int32_t* perm, void* stream);

/* ---- N2: validation metrics on the device -----------------------------------------------------------------
 * Replaces remove_padding_for_metric_eval (base_algorithm.py:88-116) and the per-list part of the reference's metric
 * code (ultra/utils/metrics.py:191-336, 456-495): scores of PAD documents (docid == n_docs; docid may be NULL) are
 * masked to -100000, every list is ranked (stable, descending) and  out[b] = { ndcg@topn[0..n) | err@topn[0..n) | mrr }
 * is written ([B, 2 n_topn + 1] floats; n_topn <= 8).  discount[r] = 1 / log2(r + 2) is supplied by the caller.
 * *flag is set to 1 when a label is not an integer in [0, 30] (gains are formed as exact powers of two).  The batch
 * means are left to the caller (B x n values instead of B x L scores cross the bus). */
UB200_API int ub200_rank_metrics(const float* scores, const float* labels, const int32_t* docid, int n_docs, int B, int L,
                       const float* discount, const int* topn, int n_topn, float max_label, float* out, int* flag,
                       void* stream);

/* ---- RegressionEM (N4) ----------------------------------------------------------------------------------------
 * Replaces RegressionEM.train's E-step / sampling / loss / M-step statistics (ultra/learning_algorithm/
 * regression_EM.py:122-183) for the scores [B, L] the ranker produced (the constant sigmoid_prob_b of the reference is 0):
 * posteriors from the current propensities prop[L], pseudo-labels ceil(p_r1 - u) with u from Philox (seed, offset) or
 * from `uniforms` [B, L] (parity tests), dscores = sigmoid(s) - label (UN-normalised: the mean over B*L is applied by the
 * optimizer scale), out = [sum of the BCE-with-logits terms, B*L, S_0..S_{L-1}] with S_l the M-step sums.  Workspace:
 * ub200_pair_workspace_bytes(B, L).  ub200_regem_update applies prop_l <- (1 - em_step) prop_l + em_step S_l / B. */
UB200_API int ub200_regression_em(const float* scores, const float* clicks, int B, int L, const float* prop,
                        const float* uniforms, unsigned long long seed, unsigned long long offset, float* dscores,
                        float* out, void* workspace, size_t workspace_bytes, void* stream);
UB200_API int ub200_regem_update(float* prop, const float* out, int L, float em_step, void* stream);

/* ---- N1: click simulation + batch assembly on the device ----------------------------------------------------------
 * Replaces ClickSimulationFeed.get_batch (click_simulation_feed.py:101-174) + PositionBiasedModel.sampleClicksForOneList
 * (click_models.py:80-110) for a data set resident in HBM.  init_list [nq, L] i32 (row ids, < 0 = PAD), rel [nq, L] f32
 * (true labels, 0 at PADs).  Per batch slot b: query ~ U{0..nq-1}; click_l ~ Bernoulli(exam_prob[min(l, n_exam-1)] *
 * click_prob[min(label, n_cp-1)]) (oracle_mode: click = label); with check_validation the draw repeats (at most
 * max_rounds times) until the list has a click.  Writes docid [L, B] i32 (PAD -> pad_id), labels [B, L] f32 and
 * query_idx [B] (may be NULL).  Philox4x32-10 keyed by (seed, offset, slot, round, position). */
UB200_API int ub200_click_batch(const int32_t* init_list, const float* rel, int nq, int L, const float* exam_prob,
                      int n_exam, const float* click_prob, int n_cp, int oracle_mode, int check_validation,
                      int max_rounds, int B, int pad_id, unsigned long long seed, unsigned long long offset,
                      int32_t* docid, float* labels, int32_t* query_idx, void* stream);
/* The same for the other click models of click_models.py:112-236: click_model = 0 position biased (as above), 1 cascade
 * (the first click ends the session), 2 user browsing (exam_prob is the [n_exam x n_exam] table exam_prob[rank][rank -
 * last_click_rank - 1] of getExamProb); lists of up to 256 positions for 1 and 2. */
UB200_API int ub200_click_batch_model(const int32_t* init_list, const float* rel, int nq, int L, const float* exam_prob,
                      int n_exam, const float* click_prob, int n_cp, int click_model, int oracle_mode,
                      int check_validation, int max_rounds, int B, int pad_id, unsigned long long seed,
                      unsigned long long offset, int32_t* docid, float* labels, int32_t* query_idx, void* stream);

/* ---- C1: data-parallel exchange of the flat gradient buffer over NVLink peer memory ----------------------------------
 * Nothing in the reference corresponds to this (it is single-process); it is the ONE collective of a data-parallel
 * step: the in-place SUM over ranks of [DNN grads | loss normalisers | EM / DenoisingNet partials].
 * peer_bufs[r] / peer_flags[r]: rank r's buffer of n floats / flag array (size: ub200_peer_flag_bytes), both in
 * symmetric memory (mapped into this process; HOST arrays of device pointers, world entries); flags zero before the
 * first call.  scratch: n floats, ctl: zeroed bytes (size: ub200_peer_ctl_bytes) (both local).  Every rank must issue the
 * same sequence of calls.  Ranks add in the order 0..world-1, so all replicas receive bitwise identical sums.
 * Graph-capturable, no host synchronisation. */
UB200_API size_t ub200_peer_ctl_bytes(void);
UB200_API size_t ub200_peer_flag_bytes(int world);
UB200_API int ub200_peer_allreduce(const void* const* peer_bufs, const void* const* peer_flags, int rank, int world,
                         float* scratch, size_t n, void* ctl, void* stream);

/* The same exchange FUSED with the optimizer step: one kernel = SUM over ranks of the flat buffer `own` (n floats: the
 * gradients of the n_params parameters followed by the trailing normalisers / partials) + ub200_clip_update on the
 * summed gradient (same arithmetic, same modes; the normaliser is own[den_index] AFTER the sum, den_index < 0 = none).
 * Push model over NVLink: every rank stores its buffer into every peer's inbox (peer_inbox[r]: ub200_dp_inbox_bytes
 * bytes of symmetric memory on rank r, double buffered by step parity) and signals per-block flags (peer_flags[r]:
 * ub200_dp_flag_bytes zeroed bytes of symmetric memory); one handshake per step, sums in rank order (bitwise equal
 * replicas).  ctl: ub200_dp_ctl_bytes zeroed local bytes.  On return (stream order) own[0, n_params) holds the clipped
 * summed gradient and own[n_params, n) the summed trailing floats.  world == 1 degenerates to ub200_clip_update. */
UB200_API size_t ub200_dp_inbox_bytes(int world, size_t n);
UB200_API size_t ub200_dp_flag_bytes(int world);
UB200_API size_t ub200_dp_ctl_bytes(void);
UB200_API int ub200_dp_reduce_update(float* own, size_t n, const void* const* peer_inbox, const void* const* peer_flags,
                           int rank, int world, float* params, float* state_sum, size_t n_params, long long den_index,
                           float scale_const, float max_norm, float lr, int mode, float* norm_out, void* ctl,
                           void* stream);

/* ub200_dp_reduce_update + ub200_publish of own[pub_index, pub_index + pub_n) (the step's summed loss scalars) from inside
 * the same kernel: a data-parallel train() returns the previous step's loss, read from host_dst / host_seq. */
UB200_API int ub200_dp_reduce_update_publish(float* own, size_t n, const void* const* peer_inbox,
                           const void* const* peer_flags, int rank, int world, float* params, float* state_sum,
                           size_t n_params, long long den_index, float scale_const, float max_norm, float lr, int mode,
                           float* norm_out, void* ctl, void* stream, long long pub_index, int pub_n, float* host_dst,
                           unsigned int* host_seq, unsigned int* dev_counter);

#ifdef __cplusplus
}
#endif
#endif /* ULTRA_B200_H_ */
