"""Shared host logic of the B200-native learning algorithms.

Mirrors `ultra.learning_algorithm.BaseAlgorithm` (reference: ultra/learning_algorithm/base_algorithm.py:32-333):
same constructor `(data_set, exp_settings)`, same `train(input_feed)` / `validation(input_feed,
is_online_simulation=False)` contract, same public attributes (`model`, `global_step`, `learning_rate`,
`rank_list_size`, `max_candidate_num`, `feature_size`, `letor_features_name`, `docid_inputs_name`, `labels_name`,
`hparams`, `is_cuda_avail`, `exp_settings`), so the reference's unmodified main.py and input feeds drive it.

What changed underneath: `create_input_feed` + `get_ranking_scores` (host gather, base_algorithm.py:134-186) become
one pinned-memory pack + one H2D copy, and ranking_model -> loss -> backward -> clip -> optimizer all run as
hand-written sm_100a kernels through the C ABI on the current CUDA stream.  One D2H read of the loss scalars per
step remains, as in the reference (`loss.item()`).
"""
import os
import sys

import numpy as np
import torch

from .. import metrics as b200_metrics
from ..engine import RankerEngine  # noqa: F401  (re-exported for subclasses)
from ..hparams import HParams  # noqa: F401


def _reference_base():
    """Subclass the reference's BaseAlgorithm when it is importable so `list_available()` sees the plugin
    (ultra/learning_algorithm/__init__.py:17-20); stand alone otherwise."""
    mod = sys.modules.get("ultra.learning_algorithm.base_algorithm")
    if mod is not None and hasattr(mod, "BaseAlgorithm"):
        return mod.BaseAlgorithm
    return object


def find_class(class_str):
    """ultra/utils/sys_tools.py:7-21."""
    mod_str, _, cls_str = class_str.rpartition('.')
    __import__(mod_str)
    return getattr(sys.modules[mod_str], cls_str)


_RANKER_ALIASES = {
    # the reference class path is accepted and mapped to the B200 implementation, so a settings JSON only has to
    # switch the learning algorithm to move the whole hot path onto the GPU kernels
    "ultra.ranking_model.DNN": "ultra_pytorch_b200.ranking_model.DNN",
    "ultra.ranking_model.Linear": "ultra_pytorch_b200.ranking_model.Linear",
}


_DEVICE_FEED_CLS = []


def _device_feed_cls():
    if not _DEVICE_FEED_CLS:
        from ..input_layer.resident import DeviceFeed
        _DEVICE_FEED_CLS.append(DeviceFeed)
    return _DEVICE_FEED_CLS[0]


class B200Algorithm(_reference_base()):
    PADDING_SCORE = -100000                      # base_algorithm.py:37
    VERBOSE = os.environ.get("UB200_QUIET", "0") != "1"
    # fixed-shape steps are captured once in a CUDA graph and replayed (one launch per step instead of ~25)
    USE_GRAPH = os.environ.get("UB200_GRAPH", "1") != "0"
    # data-parallel steps are captured as TWO graphs (compute | update) with the NCCL all-reduce launched eagerly
    # between them (capturing the collective itself dead-locked when ranks capture at different moments)
    USE_GRAPH_DP = os.environ.get("UB200_GRAPH_DP", "1") != "0"

    # ---- construction helpers ---------------------------------------------------------------------
    @staticmethod
    def _bootstrap_data_parallel():
        """Data parallel through the reference's UNMODIFIED main.py: under `torchrun --nproc-per-node N
        main.py ...` every process is one replica on one GPU.  The first B200 algorithm constructed in such a process
        (WORLD_SIZE > 1, no process group yet) binds the process to cuda:LOCAL_RANK and joins the NCCL group; every
        rank samples its own batches (global batch = N x batch_size) and the replicas stay bitwise equal, so the
        checkpoint main.py writes without any rank guard (main.py:199-214) is rank 0's: on the other ranks
        `torch.save` of a parameter dictionary is skipped (nothing else in main.py is saved with torch.save)."""
        import torch.distributed as dist
        world = int(os.environ.get("WORLD_SIZE", "1"))
        if world <= 1 or not dist.is_available() or dist.is_initialized():
            return
        if os.environ.get("UB200_DP_BOOTSTRAP", "1") == "0":
            return
        local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        if dist.get_rank() != 0:
            plain_save = torch.save

            def save_on_rank0_only(obj, f, *a, **kw):
                if isinstance(obj, dict) and obj and all(isinstance(v, torch.Tensor) for v in obj.values()):
                    return None
                return plain_save(obj, f, *a, **kw)
            torch.save = save_on_rank0_only

    def _register_param_dump(self, engine):
        """UB200_DP_DUMP=path_with_%d: every rank writes its flat parameter vector there at interpreter exit (used by
        the 2-rank main.py test to show that the replicas stayed bitwise equal)."""
        pattern = os.environ.get("UB200_DP_DUMP")
        if not pattern or getattr(B200Algorithm, "_dump_registered", False):
            return
        B200Algorithm._dump_registered = True
        import atexit
        import torch.distributed as dist

        def dump():
            rank = dist.get_rank() if dist.is_available() and dist.is_initialized() else 0
            torch.cuda.synchronize()
            # the unpatched torch.save: on ranks > 0 the bootstrap skips parameter dictionaries only, a tensor passes
            torch.save(engine.params.detach().cpu(), pattern % rank)
        atexit.register(dump)

    def _init_common(self, data_set, exp_settings, extra_floats):
        self._bootstrap_data_parallel()
        self.is_cuda_avail = True                # online feeds call .cpu() on the scores when this is set
        self.cuda = torch.device('cuda')
        self.train_summary = {}
        self.eval_summary = {}
        self.is_training = "is_train"
        self.exp_settings = exp_settings
        if 'selection_bias_cutoff' in self.exp_settings.keys():
            self.rank_list_size = self.exp_settings['selection_bias_cutoff']
        self.max_candidate_num = exp_settings['max_candidate_num']
        self.feature_size = data_set.feature_size
        self.letor_features_name = "letor_features"
        self.letor_features = None
        self.docid_inputs_name = []
        self.labels_name = []
        self.docid_inputs = []
        self.labels = []
        for i in range(self.max_candidate_num):
            self.docid_inputs_name.append("docid_input{0}".format(i))
            self.labels_name.append("label{0}".format(i))
        self.global_step = 0
        self._extra_floats = extra_floats
        self.last_h2d_bytes = 0
        self.last_d2h_bytes = 0
        self._graphs = {}
        self._feed_getters = {}
        self._phase = None       # None: whole step | 'pre': up to the all-reduce | 'post': after it

    def run_step(self, st):
        """device_step(st), replayed from CUDA graphs once the (B, L, buffer) combination has been seen twice."""
        out = self._run_step(st)
        if getattr(self, "_early", False) or getattr(self, "_late", False):
            self.engine._pub_launched += 1          # one execution of the publish kernel per launched step
        return out

    # Data parallel: the loss of a step is only final after the exchange at its end, so an early read-back does not
    # exist.  Instead the final scalars are published at the END of the step and train() of step i returns the loss of
    # step i - 1 (the first call returns its own): the host never waits for the step it has just launched and packs /
    # copies the next batch beside it.  The reference has no multi-process mode, so there is no contract to keep; the
    # value is the exact global loss, one step late.  UB200_DP_LAG_LOSS=0 restores the blocking read of the own step.
    LAG_LOSS_DP = os.environ.get("UB200_DP_LAG_LOSS", "1") != "0"
    L2_EXHAUSTS_CLIP_PARAMS = True      # see _exchange_and_update (DLA overrides)

    def _device_step_published(self, st):
        self._late_fused = False
        out = self.device_step(st)
        self._late = False
        if self._late_fused:
            self._late = True                    # the exchange kernel has published the scalars
        elif out is not None and self.LAG_LOSS_DP and self.world_size() > 1 and out.numel() <= 32:
            self.engine.publish(out)
            self.engine.join_publish()
            self._late = True
        return out

    def _run_step(self, st):
        # data parallel over peer memory: the exchange is one of OUR kernels, so the whole step is one graph again
        dp = self.world_size() > 1 and self.engine.peer is None
        if not self.USE_GRAPH or (dp and not self.USE_GRAPH_DP):
            return self._device_step_published(st)
        key = (st.B, st.L, st.feats.data_ptr(), st.docid.data_ptr())
        ent = self._graphs.get(key)
        if ent is not None and ent[4] != self.engine.generation:
            # a buffer this graph has baked in (workspace, scores, loss scratch) was released since the capture
            del self._graphs[key]
            ent = None
        if ent is None:
            if len(self._graphs) > 512:
                self._graphs.clear()
            ent = self._graphs[key] = [0, None, None, None, self.engine.generation]
        if ent[1] is not None:
            ent[1].replay()
            if ent[3] is not None:
                self._allreduce_gradbuf()
                ent[3].replay()
            return ent[2]
        ent[0] += 1
        if ent[0] <= 2:                      # warm-up: workspaces get allocated outside the capture
            out = self._device_step_published(st)
            ent[4] = self.engine.generation   # allocations of the warm-up itself are not releases of captured buffers
            return out
        torch.cuda.synchronize()
        if not dp:
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                out = self._device_step_published(st)
            ent[1], ent[2] = graph, out
            graph.replay()
            return out
        g_pre, g_post = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
        self._phase = "pre"
        with torch.cuda.graph(g_pre):
            self._device_step_published(st)
        self._phase = "post"
        with torch.cuda.graph(g_post):
            out = self._device_step_published(st)
        self._phase = None
        ent[1], ent[2], ent[3] = g_pre, out, g_post
        g_pre.replay()
        self._allreduce_gradbuf()
        g_post.replay()
        return out

    def create_model(self, feature_size):
        """base_algorithm.py:156-167 (class resolved from exp_settings['ranking_model'])."""
        cls_path = self.exp_settings['ranking_model']
        cls_path = _RANKER_ALIASES.get(cls_path, cls_path)
        cls = find_class(cls_path)
        model = cls(self.exp_settings['ranking_model_hparams'], feature_size, extra_floats=self._extra_floats)
        if not hasattr(model, "engine"):
            raise TypeError("%s is not a B200 ranking model (no .engine): the B200 learning algorithms drive the "
                            "fused kernels directly and have no fallback path" % cls_path)
        self._register_param_dump(model.engine)
        self.broadcast_initial_state(model.engine.params)
        return model

    def broadcast_initial_state(self, *tensors):
        """Data parallel: every replica starts from rank 0's initial values (the reference's main.py seeds nothing, so
        each process draws its own initial weights; torch's DDP broadcasts at construction for the same reason)."""
        import torch.distributed as dist
        if self.world_size() > 1:
            for t in tensors:
                dist.broadcast(t, src=0)

    @property
    def engine(self):
        return self.model.engine

    # ---- data-parallel plumbing ---------------------------------------------------------------------
    @staticmethod
    def world_size():
        import torch.distributed as dist
        return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1

    def _allreduce_gradbuf(self):
        """ONE all-reduce per step over [DNN grads | loss normalisers | EM / DenoisingNet partials]."""
        if self.world_size() > 1:
            self.engine.allreduce_gradbuf()

    def _exchange_and_update(self, state_sum, den, scale_const, lr, mode, norm_out):
        """The end of every training step: exchange of the flat gradient buffer between the data-parallel ranks (if
        any) + clip_grad_norm_ + optimizer step of the ranker (base_algorithm.py:208-226).  With peer memory both
        happen in one kernel; otherwise ONE all-reduce (NCCL) and the single-GPU optimizer kernel."""
        eng = self.engine
        mg = self.hparams.max_gradient_norm
        l2 = float(getattr(self.hparams, "l2_loss", 0.0) or 0.0)
        if l2 > 0:
            # loss += l2 * sum(p^2) / 2 over the ranker's parameters (e.g. ipw_rank.py:153-157): its gradient l2 * p joins
            # the un-normalised buffer scaled by what the update divides by.  Before the exchange every rank adds its
            # share (the normalisers sum over the ranks; a constant scale is split evenly); in the 'post' phase of the
            # two-graph NCCL path the buffer already holds the global sums.
            share = self.world_size() if self._phase is None else 1
            self._l2_half_sumsq = eng.l2_term(l2, den, 1.0 / (float(scale_const) * share))
            # Reference behaviour, kept: NA / IPW / PairDebias / RegressionEM loop over `params = self.model.parameters()`
            # - a generator - to add the L2 terms and then hand the EXHAUSTED generator to clip_grad_norm_ (e.g.
            # ipw_rank.py:153-159 -> base_algorithm.py:222-225), which therefore clips nothing when l2_loss > 0.
            # DLA re-creates the iterator (dla.py:161-163) and does clip.
            if self.L2_EXHAUSTS_CLIP_PARAMS:
                mg = 0.0
        fused = os.environ.get("UB200_DP_FUSED", "1") != "0"
        if self.world_size() > 1 and eng.peer is not None and fused:
            # lagged loss (LAG_LOSS_DP): the step's summed scalars leave through the exchange kernel itself when they live
            # in the flat buffer (they do for every algorithm: _publish_early records them)
            pub = getattr(self, "_dp_scalars", None) if self.LAG_LOSS_DP else None
            if pub is not None and not (eng.gradbuf.data_ptr() <= pub.data_ptr() and
                                        pub.data_ptr() + 4 * pub.numel() <= eng.gradbuf.data_ptr() + 4 * eng.gradbuf.numel()
                                        and pub.numel() <= 32):
                pub = None
            eng.dp_reduce_update(state_sum, den, scale_const, mg, lr, mode, norm_out, publish=pub)
            self._late_fused = pub is not None
            return
        if self._phase is None:
            self._allreduce_gradbuf()
        eng.clip_update(eng.params, eng.grads, state_sum, den, scale_const, mg, lr, mode, norm_out)

    # ---- input staging --------------------------------------------------------------------------------
    def _stage(self, input_feed, list_size):
        if isinstance(input_feed, _device_feed_cls()) and input_feed.L == list_size:
            # the batch was assembled on the device (N1): nothing to pack or copy
            self.letor_features = dict.__getitem__(input_feed, self.letor_features_name)
            st = self.engine.stage_device_feed(input_feed)
            self.last_h2d_bytes = 0
            return st
        getters = self._feed_getters.get(list_size)
        if getters is None:
            import operator
            one = list_size == 1                 # itemgetter with a single key returns the item, not a tuple
            getters = self._feed_getters[list_size] = (
                (lambda f, k=self.docid_inputs_name[0]: (f[k],)) if one else
                operator.itemgetter(*self.docid_inputs_name[:list_size]),
                (lambda f, k=self.labels_name[0]: (f[k],)) if one else
                operator.itemgetter(*self.labels_name[:list_size]))
        docids = getters[0](input_feed)
        labels = getters[1](input_feed)
        self.letor_features = input_feed[self.letor_features_name]
        st = self.engine.stage(self.letor_features, docids, labels)
        self.last_h2d_bytes = st.h2d_bytes
        return st

    # train() returns the loss as soon as the loss kernel has produced it (single GPU): the scalars are published into
    # mapped pinned host memory right after the loss kernel, the backward pass and the optimizer step of the batch keep
    # running while the host already packs the next batch; every later use of the parameters is ordered behind them on
    # the stream.  UB200_EARLY_LOSS=0 restores the blocking read after the whole step.
    EARLY_LOSS = os.environ.get("UB200_EARLY_LOSS", "1") != "0"

    def _publish_early(self, scalars):
        """Called by device_step right after the loss kernel; False when the scalars are only final at the end of the
        step (data parallel: they are summed by the exchange)."""
        if not self.EARLY_LOSS or self.world_size() > 1:
            self._early = False
            self._dp_scalars = scalars           # data parallel: published at the end of the step (see LAG_LOSS_DP)
            return False
        self.engine.publish(scalars)
        self._early = True
        return True

    def _read_scalars(self, t):
        """The one D2H read of a step (the reference's loss.item())."""
        if getattr(self, "_early", False):
            host = self.engine.read_published()
            self.last_d2h_bytes = host.size * 4
            return host
        if getattr(self, "_late", False):
            host = self.engine.read_published(lag=1 if self.engine._pub_launched > 1 else 0)
            self.last_d2h_bytes = host.size * 4
            return host
        host = t.detach().to("cpu", non_blocking=False)
        self.last_d2h_bytes = host.numel() * 4
        return host.numpy()

    # ---- validation (identical text in all five reference algorithms, e.g. ipw_rank.py:184-211) ----------
    def validation(self, input_feed, is_online_simulation=False):
        self.model.eval()
        L = self.max_candidate_num
        st = self._stage(input_feed, L)
        self._last_validation_stage = st          # the online feeds re-rank on the device from these buffers
        eng = self.engine
        with torch.no_grad():
            scores = eng.forward(st.feats, st.docid.view(-1), L, st.B, training=False)
            self.output = scores.clone()
        if not is_online_simulation:
            wanted = list(self.exp_settings['metrics'])
            topn = [int(n) for n in self.exp_settings['metrics_topn']]
            on_device = [m for m in wanted if m in b200_metrics.DEVICE_METRICS]
            host_vals = {}
            if on_device:
                # N2: PAD masking (base_algorithm.py:88-116), ranking and the per-list DCG / ERR / MRR chains run on the
                # device; B x (2 n + 1) floats come back instead of the B x L scores
                per_list, flag = b200_metrics.per_list_metrics(self.output, st.labels, st.docid, st.n_docs,
                                                               [min(n, L) for n in topn])
                host = torch.cat([per_list.view(-1), flag.to(torch.float32)]).cpu()
                self.last_d2h_bytes = host.numel() * 4
                if host[-1].item() != 0:
                    on_device = []                 # labels that are not small integers: the reference's module takes over
                else:
                    host_vals = b200_metrics.batch_means(host[:-1].view(st.B, -1), L, topn)
            rest = [m for m in wanted if m not in on_device]
            if rest:
                out_host = self.output.cpu()
                self.last_d2h_bytes = out_host.numel() * 4
                docid_bl = np.stack([np.asarray(input_feed[self.docid_inputs_name[i]]) for i in range(L)], axis=1)
                # same memory layout as the reference (a TRANSPOSED view of the [L, B] stack, base_algorithm.py:181-182)
                self.labels = torch.from_numpy(np.transpose(
                    np.asarray([np.asarray(input_feed[self.labels_name[i]], dtype=np.float32) for i in range(L)])))
                pad_removed = self.remove_padding_for_metric_eval(torch.from_numpy(docid_bl), out_host)
                for metric in rest:
                    host_vals[metric] = b200_metrics.reference_metric_fn(metric, topn)(self.labels, pad_removed, None)
            for metric in wanted:
                for n, metric_value in zip(topn, host_vals[metric]):
                    self.create_summary('%s_%d' % (metric, n), '%s_%d' % (metric, n), metric_value.item(), False)
        return None, self.output, self.eval_summary

    def remove_padding_for_metric_eval(self, docid_bl, model_output):
        """base_algorithm.py:88-116: documents whose id == n_docs (the PAD row) score PADDING_SCORE."""
        n_docs = self.letor_features.shape[0]
        valid = docid_bl.to(torch.int64) != n_docs
        return torch.where(valid, model_output, torch.ones_like(model_output) * self.PADDING_SCORE)

    def create_summary(self, scalar_name, summarize_name, value, is_training):
        if is_training:
            self.train_summary[summarize_name] = value
        else:
            self.eval_summary[summarize_name] = value

    def _say(self, loss):
        if self.VERBOSE:
            print(" Loss %f at Global Step %d: " % (loss, self.global_step))

    # ---- optimizer selection (ipw_rank.py:96-99) ---------------------------------------------------------
    def _opt_mode(self, fresh=False):
        if self.hparams.grad_strategy == 'sgd':
            return 2
        return 1 if fresh else 0

    def _check_l2(self):
        """l2_loss > 0 is applied in _exchange_and_update (gradient) and _l2_loss_value (reported loss)."""
        self._l2_half_sumsq = None

    def _l2_loss_value(self):
        """l2_loss * sum(p^2) / 2 of the parameters the step started from (one extra scalar read; 0.0 when off)."""
        l2 = float(getattr(self.hparams, "l2_loss", 0.0) or 0.0)
        if l2 <= 0 or getattr(self, "_l2_half_sumsq", None) is None:
            return 0.0
        return l2 * float(self._l2_half_sumsq.item())
