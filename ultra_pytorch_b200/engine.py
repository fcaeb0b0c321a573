"""Device-side state and launch plumbing for the hot path (PyTorch is used for device memory and streams only).

`RankerEngine` owns, for one DNN ranker on one GPU:
  - the flat fp32 parameter / gradient / Adagrad-accumulator buffers (layout of include/ultra_b200.h),
  - the kernel workspaces (zeroed once; the kernels keep their ticket counters at zero),
  - a pinned host staging buffer + its device mirror for one step's input_feed.
Every compute call goes through the C ABI (ultra_pytorch_b200._capi); there is no eager fallback.
"""
import ctypes
import os

import numpy as np
import torch

from . import _capi
from ._capi import check, int_array, lib


def _ptr(t):
    return 0 if t is None else t.data_ptr()


_RAW_STREAM = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def _stream():
    """cudaStream_t of torch's current stream on the current device.  torch.cuda.current_stream() builds a Stream object
    (~3 us per call, several calls per train()); the raw accessor - the one torch's own extensions use - costs ~0.3 us."""
    if _RAW_STREAM is not None:
        return _RAW_STREAM(torch.cuda.current_device())
    return torch.cuda.current_stream().cuda_stream


_RESIDENT_CLS = []


def _resident_features_cls():
    if not _RESIDENT_CLS:
        from .input_layer.resident import ResidentFeatures
        _RESIDENT_CLS.append(ResidentFeatures)
    return _RESIDENT_CLS[0]


def column_ptrs(arrays, B):
    """ctypes array of the data pointers of L per-position feed arrays (f32 [B] each) + an object to keep alive.

    Asking 80 numpy arrays for `.ctypes.data` costs ~2 us each; stacking them into one [L, B] block (a 40 KB copy) and
    deriving the row addresses arithmetically is ~6x cheaper, which matters on the host path of a 0.36 ms step."""
    L = len(arrays)
    try:
        block = np.array(arrays, dtype=np.float32)
    except ValueError:
        block = None
    if block is None or block.shape != (L, B):
        keep = [np.ascontiguousarray(x, dtype=np.float32) for x in arrays]
        if any(x.shape != (B,) for x in keep):
            raise ValueError("every docid_input / label array of a feed must have shape (%d,)" % B)
        return (ctypes.c_void_p * L)(*[x.ctypes.data for x in keep]), keep
    key = (L, B)
    offs = _ROW_OFFSETS.get(key)
    if offs is None:
        if len(_ROW_OFFSETS) > 256:
            _ROW_OFFSETS.clear()
        offs = _ROW_OFFSETS[key] = np.arange(L, dtype=np.uint64) * np.uint64(4 * B)
    addr = offs + np.uint64(block.__array_interface__['data'][0])
    return (ctypes.c_void_p * L).from_buffer(addr), (block, addr)


_ROW_OFFSETS = {}


def column_ptrs2(docid_arrays, label_arrays, B):
    """column_ptrs for the docid and the label arrays of a feed with ONE stacking copy ([2 L, B] block)."""
    L = len(docid_arrays)
    if len(label_arrays) != L:
        raise ValueError("a feed needs as many label arrays as docid_input arrays")
    ptrs, keep = column_ptrs(tuple(docid_arrays) + tuple(label_arrays), B)
    lab = (ctypes.c_void_p * L).from_address(ctypes.addressof(ptrs) + L * ctypes.sizeof(ctypes.c_void_p))
    return ptrs, lab, (ptrs, keep)


class Staged(object):
    """Device views of one staged input_feed."""
    __slots__ = ("feats", "docid", "labels", "B", "L", "n_docs", "h2d_bytes")


def make_staged(dev, L, B, n_docs, F):
    """Views of a staging buffer laid out as docid i32 [L, B] | labels f32 [B, L] | pad to 256 B | feats f32 [n_docs+1, F]."""
    off_l = 4 * L * B
    off_f = (8 * L * B + 255) // 256 * 256
    total = off_f + 4 * (n_docs + 1) * F
    st = Staged()
    st.docid = dev[:off_l].view(torch.int32).view(L, B)
    st.labels = dev[off_l:2 * off_l].view(torch.float32).view(B, L)
    st.feats = dev[off_f:total].view(torch.float32).view(n_docs + 1, F)
    st.B, st.L, st.n_docs, st.h2d_bytes = B, L, n_docs, total
    return st


class StagedCache(object):
    """Building the three tensor views costs ~10-20 us of Python per step; the same (buffer, L, B, n_docs) combination
    recurs every step of a fixed-shape run.  Keyed on the buffer's address: a cached entry keeps its buffer alive, so
    the address cannot be handed to another tensor while the entry exists."""

    def __init__(self, limit=64):
        self.limit = limit
        self.entries = {}

    def get(self, dev, L, B, n_docs, F):
        key = (dev.data_ptr(), dev.numel(), L, B, n_docs, F)
        st = self.entries.get(key)
        if st is None:
            if len(self.entries) >= self.limit:
                self.entries.clear()
            st = self.entries[key] = make_staged(dev, L, B, n_docs, F)
        return st


class StagingSlot(object):
    """One (pinned, device) buffer pair of the double-buffered input staging + its two events."""
    __slots__ = ("pin", "dev", "pin_np", "free_ev", "ready_ev", "free_valid", "used", "resident_views", "pin_ptr",
                 "pin_bytes", "dev_ptr", "free_h", "ready_h")

    def __init__(self):
        self.pin = self.dev = self.pin_np = None
        # free_ev: recorded on the compute stream behind the last kernel that reads `dev`;
        # ready_ev: recorded on the copy stream behind the last H2D copy into `dev`
        self.free_ev = torch.cuda.Event()
        self.ready_ev = torch.cuda.Event()
        cur = torch.cuda.current_stream()
        self.free_ev.record(cur)                 # torch creates the CUDA event lazily: materialise both handles now
        self.ready_ev.record(cur)
        self.free_h, self.ready_h = self.free_ev.cuda_event, self.ready_ev.cuda_event     # raw handles for the C calls
        self.pin_ptr = self.pin_bytes = self.dev_ptr = 0
        self.free_valid = False
        self.used = False
        self.resident_views = {}

    def ensure(self, nbytes, device):
        if self.pin is not None and self.pin.numel() >= nbytes:
            return False
        cap = int(nbytes * 1.5) + 1024
        self.pin = torch.empty(cap, dtype=torch.uint8, pin_memory=True)
        self.dev = torch.empty(cap, dtype=torch.uint8, device=device)
        self.pin_np = self.pin.numpy()
        self.pin_ptr, self.pin_bytes, self.dev_ptr = self.pin.data_ptr(), self.pin.numel(), self.dev.data_ptr()
        self.resident_views.clear()
        return True


class RankerEngine(object):
    def __init__(self, feature_size, hidden, device=None, extra_floats=0, activation=0):
        if not torch.cuda.is_available():
            raise _capi.UltraB200Error("ultra_pytorch_b200 needs a CUDA device (B200, sm_100a); no CPU fallback")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.F = int(feature_size)
        self.hidden = [int(h) for h in hidden]
        self._hidden_c = int_array(self.hidden)
        self.n_hidden = len(self.hidden)
        self.activation = int(activation)        # UB200_ACT_*: 0 elu (tensor cores), 1 relu, 2 selu, 3 tanh, 4 sigmoid
        self.P = int(lib.ub200_mlp_param_count(self.F, self._hidden_c, self.n_hidden))
        if self.P == 0:
            raise _capi.UltraB200Error("bad ranker spec F=%d hidden=%s" % (self.F, self.hidden))
        self.E = int(extra_floats)
        f32 = dict(dtype=torch.float32, device=self.device)
        self.params = torch.zeros(self.P, **f32)
        # gradient buffer with E trailing floats for the loss normalisers / EM partials, so that data-parallel
        # ranks need ONE all-reduce per step
        self.peer = None
        self.gradbuf = self._alloc_gradbuf(f32)
        self.grads = self.gradbuf[:self.P]
        self.extra = self.gradbuf[self.P:]
        self.state_sum = torch.zeros(self.P, **f32)
        self.norm = torch.zeros(1, **f32)
        self._ws = {}
        self._staged_cache = StagedCache()
        self._opt_ws = torch.zeros(int(lib.ub200_opt_workspace_bytes(self.P)), dtype=torch.uint8, device=self.device)
        self._loss_ws = None
        # input staging: two (pinned, device) buffer pairs used alternately, copies on their own stream, so that the
        # H2D transfer of step i overlaps the backward pass / optimizer step of step i - 1 (train() returns as soon as
        # the loss is known).  UB200_STAGE_SLOTS=1: one pair, copies on the compute stream (the round-1 behaviour).
        self._n_slots = max(1, int(os.environ.get("UB200_STAGE_SLOTS", "2")))
        self._slots = None
        self._slot_i = 0
        self._copy_stream = None
        # host packer threads: the ranks of a node share its cores (torchrun exports LOCAL_WORLD_SIZE); spinning workers
        # of 8 ranks x 16 threads on 16 cores would starve each other
        local_world = max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1")))
        self._pack_threads = int(os.environ.get("UB200_PACK_THREADS",
                                                str(max(1, min(16, (os.cpu_count() or 1) // local_world)))))
        self._pack_chunks = int(os.environ.get("UB200_PACK_CHUNKS", "4"))
        self._scores = {}
        self._dscores = {}
        # CUDA graphs captured by the learning algorithms have the raw pointers of the workspaces / score buffers baked
        # in.  Whenever one of those buffers is released the generation moves on, and every graph captured under an
        # older generation is discarded instead of replayed (base_algorithm.py: _run_step).
        self.generation = 0

    MAX_SHAPES = 64      # (L, B) shapes whose workspaces stay cached (DirectLabelFeed batches vary in size)

    @staticmethod
    def _evict_lru(cache, limit):
        """dicts keep insertion order and hits re-insert their key: the first key is the least recently used one"""
        evicted = False
        while len(cache) >= limit:
            cache.pop(next(iter(cache)))
            evicted = True
        return evicted

    # ---- data-parallel exchange buffer ----------------------------------------------------------------
    def _alloc_gradbuf(self, f32):
        """Single GPU: a plain device buffer.  Data parallel (NCCL process group, one rank per GPU of one node): the
        buffer and a flag array are allocated in SYMMETRIC memory so that every rank can read every peer's gradients
        over NVLink (csrc/peer.cu); if symmetric memory is unavailable the step falls back to ONE NCCL all-reduce."""
        import torch.distributed as dist
        from .learning_algorithm.base_algorithm import B200Algorithm
        n = self.P + self.E
        # NOTE: in data-parallel mode constructing an engine is a COLLECTIVE operation (symmetric-memory rendezvous):
        # every rank must build the same rankers in the same order, like wrapping a module in DDP
        if B200Algorithm.world_size() <= 1 or os.environ.get("UB200_DP_PEER", "1") == "0" or \
                dist.get_backend() != "nccl":
            return torch.zeros(n, **f32)
        try:
            import torch.distributed._symmetric_memory as symm
            world, rank = dist.get_world_size(), dist.get_rank()
            group = dist.group.WORLD
            buf = symm.empty(n, dtype=torch.float32, device=self.device)
            flags = symm.empty(max(64, 2 * world), dtype=torch.int32, device=self.device)
            buf.zero_()
            flags.zero_()
            # fused exchange + optimizer (push model): per-rank inboxes [2][world][n] and per-block flags
            inbox = symm.empty(int(lib.ub200_dp_inbox_bytes(world, n)) // 4, dtype=torch.float32, device=self.device)
            pflags = symm.empty(int(lib.ub200_dp_flag_bytes(world)) // 4, dtype=torch.int32, device=self.device)
            inbox.zero_()
            pflags.zero_()
            hb = symm.rendezvous(buf, group)
            hf = symm.rendezvous(flags, group)
            hi = symm.rendezvous(inbox, group)
            hp = symm.rendezvous(pflags, group)
            torch.cuda.synchronize()
            dist.barrier()                        # every rank's flags are zero before anybody signals
            PtrArr = ctypes.c_void_p * world
            self.peer = {
                "bufs": PtrArr(*[int(p) for p in hb.buffer_ptrs]), "flags": PtrArr(*[int(p) for p in hf.buffer_ptrs]),
                "inbox": PtrArr(*[int(p) for p in hi.buffer_ptrs]), "pflags": PtrArr(*[int(p) for p in hp.buffer_ptrs]),
                "rank": rank, "world": world, "handles": (hb, hf, hi, hp), "tensors": (flags, inbox, pflags),
                "scratch": torch.zeros(n, **f32),
                "ctl": torch.zeros(int(lib.ub200_peer_ctl_bytes()), dtype=torch.uint8, device=self.device),
                "dp_ctl": torch.zeros(int(lib.ub200_dp_ctl_bytes()), dtype=torch.uint8, device=self.device),
            }
            assert int(hb.buffer_ptrs[rank]) == buf.data_ptr()
            return buf
        except Exception as e:  # noqa: BLE001  (no symmetric memory on this system: NCCL all-reduce instead)
            if os.environ.get("UB200_DP_PEER", "1") == "2":
                raise
            print("ultra_pytorch_b200: symmetric memory unavailable (%s); data-parallel steps use ncclAllReduce" % (e,))
            self.peer = None
            return torch.zeros(n, **f32)

    def allreduce_gradbuf(self):
        """ONE exchange per step over [DNN grads | loss normalisers | EM / DenoisingNet partials]."""
        import torch.distributed as dist
        if self.peer is not None:
            p = self.peer
            check(lib.ub200_peer_allreduce(p["bufs"], p["flags"], p["rank"], p["world"], _ptr(p["scratch"]),
                                           self.gradbuf.numel(), _ptr(p["ctl"]), _stream()), "ub200_peer_allreduce")
        else:
            dist.all_reduce(self.gradbuf, op=dist.ReduceOp.SUM)

    def _ensure_publish_buffers(self):
        if getattr(self, "_pub_host", None) is None:
            self._pub_host = torch.zeros(128, dtype=torch.float32, pin_memory=True)   # 2 x 32 values, [64] seq
            self._pub_np = self._pub_host.numpy()
            self._pub_seq_np = self._pub_np[64:65].view(np.uint32)
            self._pub_counter = torch.zeros(1, dtype=torch.int32, device=self.device)
            self._pub_stream = torch.cuda.Stream(device=self.device)
            self._pub_launched = 0

    def dp_reduce_update(self, state_sum, den, scale_const, max_norm, lr, mode, norm_out=None, publish=None):
        """Fused exchange + optimizer of the ranker's parameters (csrc/peer.cu: dp_reduce_update_kernel): SUM of the
        whole flat buffer over the ranks through NVLink peer memory, then clip_grad_norm_ + Adagrad / SGD on the summed
        gradient, in ONE kernel.  `den` is a view INTO self.gradbuf (the normaliser is read after the sum)."""
        p = self.peer
        den_index = -1 if den is None else (den.data_ptr() - self.gradbuf.data_ptr()) // 4
        assert den is None or 0 <= den_index < self.gradbuf.numel()
        if publish is not None:
            # `publish`: a view INTO self.gradbuf (the step's loss scalars): written to mapped host memory by the kernel
            self._ensure_publish_buffers()
            pub_index = (publish.data_ptr() - self.gradbuf.data_ptr()) // 4
            assert 0 <= pub_index and pub_index + publish.numel() <= self.gradbuf.numel()
            check(lib.ub200_dp_reduce_update_publish(
                _ptr(self.gradbuf), self.gradbuf.numel(), p["inbox"], p["pflags"], p["rank"], p["world"],
                _ptr(self.params), _ptr(state_sum), self.P, den_index, float(scale_const), float(max_norm), float(lr),
                int(mode), _ptr(norm_out), _ptr(p["dp_ctl"]), _stream(), pub_index, publish.numel(),
                self._pub_host.data_ptr(), self._pub_host.data_ptr() + 256, _ptr(self._pub_counter)),
                "ub200_dp_reduce_update_publish")
            self._pub_n = publish.numel()
            return
        check(lib.ub200_dp_reduce_update(_ptr(self.gradbuf), self.gradbuf.numel(), p["inbox"], p["pflags"], p["rank"],
                                         p["world"], _ptr(self.params), _ptr(state_sum), self.P, den_index,
                                         float(scale_const), float(max_norm), float(lr), int(mode), _ptr(norm_out),
                                         _ptr(p["dp_ctl"]), _stream()), "ub200_dp_reduce_update")

    # ---- parameter views -------------------------------------------------------------------------
    def layer_slices(self):
        """[(name, offset, shape)] in the flat layout == the reference's state_dict order (DNN.py:43-55)."""
        out, off, k = [], 0, self.F
        for j, n in enumerate(self.hidden + [1]):
            for name, shape in (("layer_norm%d.weight" % j, (k,)), ("layer_norm%d.bias" % j, (k,)),
                                ("linear%d.weight" % j, (n, k)), ("linear%d.bias" % j, (n,))):
                cnt = int(np.prod(shape))
                out.append((name, off, shape))
                off += cnt
            k = n
        assert off == self.P
        return out

    # ---- workspaces ------------------------------------------------------------------------------
    def _mlp_ws(self, L, B, training):
        key = (L, B, int(bool(training)))
        ws = self._ws.pop(key, None)
        if ws is None:
            n = int(lib.ub200_mlp_workspace_bytes(L, B, self.F, self._hidden_c, self.n_hidden, key[2]))
            if self._evict_lru(self._ws, self.MAX_SHAPES):
                self.generation += 1
            ws = torch.zeros(max(n, 256), dtype=torch.uint8, device=self.device)
        self._ws[key] = ws                        # (re-)inserted last = most recently used
        return ws

    def loss_ws(self, B, L):
        n = max(int(lib.ub200_loss_workspace_bytes(B, L)), int(lib.ub200_pair_workspace_bytes(B, L)))
        if self._loss_ws is None or self._loss_ws.numel() < n:
            if self._loss_ws is not None:
                self.generation += 1
            self._loss_ws = torch.zeros(n, dtype=torch.uint8, device=self.device)
        return self._loss_ws

    def scores_buf(self, B, L):
        key = (B, L)
        t = self._scores.pop(key, None)
        if t is None:
            if self._evict_lru(self._scores, self.MAX_SHAPES):
                self.generation += 1
                for k in [k for k in self._dscores if k not in self._scores]:
                    del self._dscores[k]
            t = torch.empty(B, L, dtype=torch.float32, device=self.device)
            self._dscores[key] = torch.empty(B, L, dtype=torch.float32, device=self.device)
        self._scores[key] = t
        return t

    def dscores_buf(self, B, L):
        self.scores_buf(B, L)
        return self._dscores[(B, L)]

    # ---- staging: host input_feed -> device ---------------------------------------------------------
    def stage(self, letor_features, docid_arrays, label_arrays):
        """One pinned-memory pack + ONE async H2D copy of a step's inputs.

        letor_features: np [n_docs, F] (f64 in the reference feeds, cast like DNN.py:73);
        docid_arrays / label_arrays: L arrays of [B] (f32 in the feeds, base_algorithm.py:176-186).
        Device layout: docid i32 [L, B] (position-major) | labels f32 [B, L] |
                       feats f32 [n_docs+1, F] (last row = zero PAD, base_algorithm.py:148-149)."""
        if isinstance(letor_features, _resident_features_cls()):
            return self._stage_resident(letor_features, docid_arrays, label_arrays)
        feats = np.asarray(letor_features)
        n_docs = feats.shape[0] if feats.ndim == 2 else 0
        L = len(docid_arrays)
        B = len(docid_arrays[0])
        nf = (n_docs + 1) * self.F
        # fixed offsets for a given (L, B): docid | labels | feats (so captured CUDA graphs keep valid pointers)
        off_l = 4 * L * B
        off_f = (8 * L * B + 255) // 256 * 256
        total = off_f + 4 * nf
        cs = _stream()                    # (torch.cuda.current_stream() costs ~5 us: asked once per call)
        slot = self._acquire_slot(total, cs)
        def _f32_vec(x):
            return isinstance(x, np.ndarray) and x.dtype == np.float32 and x.flags.c_contiguous and x.shape == (B,)
        # feeds emit homogeneous per-position arrays (click_simulation_feed.py:141-150): checking the ends is enough
        fast = (feats.dtype == np.float64 and feats.flags.c_contiguous and n_docs > 0 and feats.shape[1] == self.F and
                _f32_vec(docid_arrays[0]) and _f32_vec(docid_arrays[-1]) and _f32_vec(label_arrays[0]) and
                _f32_vec(label_arrays[-1]))
        if fast:
            # ONE C call: ids/labels + the f64 -> f32 conversion of the feature rows on the persistent host thread pool,
            # straight into pinned memory, with the H2D copy of every finished group of rows enqueued on the current
            # stream while the rest is still being converted (csrc/hostpack.cpp).
            dptr, lptr, keep = column_ptrs2(docid_arrays, label_arrays, B)
            if self._n_slots > 1:
                check(lib.ub200_stage_feed_pipelined(
                    feats.__array_interface__['data'][0], n_docs, self.F, dptr, lptr, L, B, slot.pin_ptr, slot.pin_bytes, slot.dev_ptr,
                    self._pack_threads, self._pack_chunks, self._copy_stream_h, cs,
                    slot.free_h if slot.free_valid else None, slot.ready_h), "ub200_stage_feed_pipelined")
            else:
                check(lib.ub200_stage_feed(feats.__array_interface__['data'][0], n_docs, self.F, dptr, lptr, L, B, slot.pin_ptr,
                                           slot.pin_bytes, slot.dev_ptr, self._pack_threads, self._pack_chunks, cs),
                      "ub200_stage_feed")
            slot.used = True
            return self.staged_views(slot.dev, L, B, n_docs)
        else:
            buf = slot.pin_np
            hd = buf[:off_l].view(np.int32).reshape(L, B)
            hl = buf[off_l:2 * off_l].view(np.float32).reshape(B, L)
            hf = buf[off_f:total].view(np.float32).reshape(n_docs + 1, self.F)
            if n_docs:
                np.copyto(hf[:n_docs], feats, casting="same_kind")
            hf[n_docs] = 0.0
            for l in range(L):
                np.copyto(hd[l], docid_arrays[l], casting="unsafe")
                hl[:, l] = label_arrays[l]
        self._copy_slot(slot, total)
        return self.staged_views(slot.dev, L, B, n_docs)

    def _acquire_slot(self, nbytes, cs=None):
        """The staging buffer pair of this call.  Marks, on the compute stream, the end of everything launched so far:
        the pair used by the previous call may be overwritten (by the call after this one) once that point has passed.
        The pinned half is reused two calls later; by then its copy has long finished (every train() / validation()
        call reads a result of its own step back), which is checked rather than assumed."""
        if self._slots is None:
            self._slots = [StagingSlot() for _ in range(self._n_slots)]
            self._copy_stream = torch.cuda.Stream(device=self.device) if self._n_slots > 1 else None
            self._copy_stream_h = self._copy_stream.cuda_stream if self._n_slots > 1 else None
        if self._n_slots > 1:
            prev = self._slots[self._slot_i]
            if prev.used:
                check(lib.ub200_event_record(prev.free_h, _stream() if cs is None else cs), "ub200_event_record")
                prev.free_valid = True
            self._slot_i = (self._slot_i + 1) % self._n_slots
        slot = self._slots[self._slot_i]
        if slot.used and not slot.ready_ev.query():
            slot.ready_ev.synchronize()
        if slot.ensure(nbytes, self.device):
            self._staged_cache.entries.clear()       # views of the previous staging buffer
            slot.free_valid = False
        return slot

    @property
    def _dev(self):
        """device half of the staging pair used by the last stage() call"""
        return None if self._slots is None else self._slots[self._slot_i].dev

    def _copy_slot(self, slot, nbytes):
        """pinned -> device for the paths that packed with numpy (same ordering as ub200_stage_feed_pipelined)"""
        slot.used = True
        if self._n_slots == 1:
            slot.dev[:nbytes].copy_(slot.pin[:nbytes], non_blocking=True)
            return
        cur = torch.cuda.current_stream()
        with torch.cuda.stream(self._copy_stream):
            if slot.free_valid:
                self._copy_stream.wait_event(slot.free_ev)
            slot.dev[:nbytes].copy_(slot.pin[:nbytes], non_blocking=True)
            slot.ready_ev.record(self._copy_stream)
        cur.wait_event(slot.ready_ev)

    def _stage_resident(self, feats, docid_arrays, label_arrays):
        """The feed's `letor_features` is the data set's whole feature matrix (input_layer/resident.py) and the doc ids
        are global row ids: keep the matrix in HBM (uploaded once, fp32, zero PAD row appended) and move only ids and
        labels per step."""
        n_rows = feats.shape[0]
        self.ensure_resident(feats)
        L = len(docid_arrays)
        B = len(docid_arrays[0])
        nbytes = 8 * L * B
        cs = _stream()
        slot = self._acquire_slot(nbytes, cs)
        dptr, keep_d = column_ptrs(docid_arrays, B)
        lptr, keep_l = column_ptrs(label_arrays, B)
        if self._n_slots > 1:
            check(lib.ub200_stage_ids_pipelined(dptr, lptr, L, B, n_rows, slot.pin_ptr, slot.pin_bytes, slot.dev_ptr,
                                                self._copy_stream_h, cs, slot.free_h if slot.free_valid else None,
                                                slot.ready_h), "ub200_stage_ids_pipelined")
            slot.used = True
        else:
            check(lib.ub200_pack_ids_host(dptr, lptr, L, B, n_rows, slot.pin_ptr, slot.pin_bytes), "ub200_pack_ids_host")
            self._copy_slot(slot, nbytes)
        key = (L, B, n_rows, self._resident.data_ptr())
        st = slot.resident_views.get(key)
        if st is None:
            if len(slot.resident_views) >= 64:
                slot.resident_views.clear()
            st = slot.resident_views[key] = Staged()
            st.docid = slot.dev[:4 * L * B].view(torch.int32).view(L, B)
            st.labels = slot.dev[4 * L * B:nbytes].view(torch.float32).view(B, L)
            st.feats = self._resident
            st.B, st.L, st.n_docs, st.h2d_bytes = B, L, n_rows, nbytes
        return st

    def stage_device_feed(self, feed):
        """A batch assembled on the device (input_layer/resident.py: DeviceFeed): nothing to pack, nothing to copy."""
        self.ensure_resident(dict.__getitem__(feed, feed.model.letor_features_name))
        st = Staged()
        st.docid, st.labels, st.feats = feed.docid, feed.labels, self._resident
        st.B, st.L, st.n_docs, st.h2d_bytes = feed.B, feed.L, feed.n_rows, 0
        return st

    def click_batch(self, init_list, rel, exam_prob, click_prob, oracle_mode, check_validation, max_rounds, pad_id, seed,
                    offset, docid, labels, query_idx, click_model=0):
        """click_model: 0 position biased (exam_prob [n]), 1 cascade (exam_prob [n]), 2 user browsing (exam_prob [n, n])"""
        nq, L = init_list.shape
        B = docid.shape[1]
        n_exam = 0 if exam_prob is None else (exam_prob.shape[0] if click_model == 2 else exam_prob.numel())
        check(lib.ub200_click_batch_model(_ptr(init_list), _ptr(rel), nq, L, _ptr(exam_prob), n_exam, _ptr(click_prob),
                                          0 if click_prob is None else click_prob.numel(), int(click_model),
                                          int(oracle_mode), int(check_validation), int(max_rounds), B, int(pad_id),
                                          int(seed), int(offset), _ptr(docid), _ptr(labels), _ptr(query_idx), _stream()),
              "ub200_click_batch_model")

    # bytes of resident feature matrices kept in HBM at once (train / valid / test sets of one run stay resident side by
    # side; the least recently used one goes when the cap would be exceeded)
    RESIDENT_CAP_BYTES = int(float(os.environ.get("UB200_RESIDENT_CAP_GB", "96")) * (1 << 30))

    def ensure_resident(self, feats):
        """Makes the data set's whole feature matrix (fp32 + zero PAD row) resident and current.  Matrices are cached by
        `resident_key()`: main.py alternates train / validation / test feeds on one model, and re-converting GBs of
        f64 at every switch (and orphaning the CUDA graphs keyed on the old pointer) would defeat "uploaded once"."""
        n_rows, F = feats.shape
        if F != self.F:
            raise _capi.UltraB200Error("resident feature matrix has %d columns, the ranker expects %d" % (F, self.F))
        if n_rows >= (1 << 24):
            raise _capi.UltraB200Error("doc ids travel as float32 in the feed format: at most 2^24 rows per data set")
        key = feats.resident_key()
        cache = self.__dict__.setdefault("_resident_cache", {})
        ent = cache.pop(key, None)
        if ent is None:
            need = 4 * (n_rows + 1) * F
            while cache and sum(e[0].numel() * 4 for e in cache.values()) + need > self.RESIDENT_CAP_BYTES:
                cache.pop(next(iter(cache)))                         # least recently used
                self.generation += 1                                 # graphs keyed on its pointer must not be replayed
            dev = torch.empty((n_rows + 1) * F, dtype=torch.float32, device=self.device)
            chunk_rows = max(1, (32 << 20) // (4 * F))
            pin = torch.empty(chunk_rows * F, dtype=torch.float32, pin_memory=True)
            src = feats.ctypes.data
            for r0 in range(0, n_rows, chunk_rows):
                r1 = min(n_rows, r0 + chunk_rows)
                n = (r1 - r0) * F
                check(lib.ub200_convert_f64_f32_host(src + 8 * r0 * F, pin.data_ptr(), n, self._pack_threads),
                      "ub200_convert_f64_f32_host")
                dev[r0 * F:r1 * F].copy_(pin[:n], non_blocking=True)
                torch.cuda.current_stream().synchronize()           # the pinned chunk is reused
            dev[n_rows * F:].zero_()                                 # PAD row (base_algorithm.py:148-149)
            ent = (dev.view(n_rows + 1, F), feats)                   # the host array keeps its address alive while it is the key
        cache[key] = ent                                             # most recently used last
        self._resident = ent[0]
        self._resident_key = key

    def staged_views(self, dev, L, B, n_docs):
        return self._staged_cache.get(dev, L, B, n_docs, self.F)

    # ---- K1 ---------------------------------------------------------------------------------------
    def forward(self, feats, docid, L, B, training, scores=None):
        """feats f32 [rows, F] cuda, docid i32 [L*B] cuda or None (identity) -> scores [B, L]."""
        if scores is None:
            scores = self.scores_buf(B, L)
        ws = self._mlp_ws(L, B, training)
        if self.activation:
            check(lib.ub200_mlp_forward_act(_ptr(feats), _ptr(docid), L, B, self.F, self._hidden_c, self.n_hidden,
                                            self.activation, _ptr(self.params), _ptr(scores), _ptr(ws), ws.numel(),
                                            int(bool(training)), _stream()), "ub200_mlp_forward_act")
            return scores
        check(lib.ub200_mlp_forward(_ptr(feats), _ptr(docid), L, B, self.F, self._hidden_c, self.n_hidden,
                                    _ptr(self.params), _ptr(scores), _ptr(ws), ws.numel(), int(bool(training)),
                                    _stream()), "ub200_mlp_forward")
        return scores

    def backward(self, feats, docid, L, B, dscores):
        ws = self._mlp_ws(L, B, True)
        if self.activation:
            check(lib.ub200_mlp_backward_act(_ptr(feats), _ptr(docid), L, B, self.F, self._hidden_c, self.n_hidden,
                                             self.activation, _ptr(self.params), _ptr(dscores), _ptr(self.grads),
                                             _ptr(ws), ws.numel(), _stream()), "ub200_mlp_backward_act")
            return self.grads
        check(lib.ub200_mlp_backward(_ptr(feats), _ptr(docid), L, B, self.F, self._hidden_c, self.n_hidden,
                                     _ptr(self.params), _ptr(dscores), _ptr(self.grads), _ptr(ws), ws.numel(),
                                     _stream()), "ub200_mlp_backward")
        return self.grads

    # ---- K2 / K3 -----------------------------------------------------------------------------------
    def softmax_ce(self, scores, labels, weight_mode, table, dscores, sums):
        B, L = scores.shape
        ws = self.loss_ws(B, L)
        check(lib.ub200_softmax_ce(_ptr(scores), _ptr(labels), B, L, weight_mode, _ptr(table),
                                   0 if table is None else table.numel(), _ptr(dscores), _ptr(sums), _ptr(ws),
                                   ws.numel(), _stream()), "ub200_softmax_ce")

    def dla_loss(self, scores, clicks, prop_w, prop_b, dscores, dprop, sums):
        B, L = scores.shape
        ws = self.loss_ws(B, L)
        check(lib.ub200_dla_loss(_ptr(scores), _ptr(clicks), B, L, _ptr(prop_w), _ptr(prop_b), _ptr(dscores),
                                 _ptr(dprop), _ptr(sums), _ptr(ws), ws.numel(), _stream()), "ub200_dla_loss")

    def lambdarank(self, scores, labels, sigma, t_plus, t_minus, dscores, out):
        B, L = scores.shape
        ws = self.loss_ws(B, L)
        check(lib.ub200_lambdarank(_ptr(scores), _ptr(labels), B, L, float(sigma), _ptr(t_plus), _ptr(t_minus),
                                   _ptr(dscores), _ptr(out), _ptr(ws), ws.numel(), _stream()), "ub200_lambdarank")

    def prsrank(self, scores, labels, sigma, ipw_table, dscores, out):
        B, L = scores.shape
        ws = self.loss_ws(B, L)
        check(lib.ub200_prsrank(_ptr(scores), _ptr(labels), B, L, float(sigma), _ptr(ipw_table), ipw_table.numel(),
                                _ptr(dscores), _ptr(out), _ptr(ws), ws.numel(), _stream()), "ub200_prsrank")

    def pairdebias(self, scores, clicks, t_plus, t_minus, dscores, out):
        B, L = scores.shape
        ws = self.loss_ws(B, L)
        check(lib.ub200_pairdebias(_ptr(scores), _ptr(clicks), B, L, _ptr(t_plus), _ptr(t_minus), _ptr(dscores),
                                   _ptr(out), _ptr(ws), ws.numel(), _stream()), "ub200_pairdebias")

    def regression_em(self, scores, clicks, prop, uniforms, seed, offset, dscores, out):
        B, L = scores.shape
        ws = self.loss_ws(B, L)
        check(lib.ub200_regression_em(_ptr(scores), _ptr(clicks), B, L, _ptr(prop), _ptr(uniforms), int(seed), int(offset),
                                      _ptr(dscores), _ptr(out), _ptr(ws), ws.numel(), _stream()), "ub200_regression_em")

    def regem_update(self, prop, out, em_step):
        check(lib.ub200_regem_update(_ptr(prop), _ptr(out), prop.numel(), float(em_step), _stream()),
              "ub200_regem_update")

    def em_update(self, t_plus, t_minus, out, em_step, reg_p, safe_div):
        check(lib.ub200_em_update(_ptr(t_plus), _ptr(t_minus), _ptr(out), t_plus.numel(), float(em_step),
                                  float(reg_p), int(bool(safe_div)), _stream()), "ub200_em_update")

    # ---- N3: Plackett-Luce re-ranking (online simulation feeds) -------------------------------------------
    def pl_sample(self, scores, docid, n_docs, tau, seed, offset=0):
        """scores [B, L] cuda f32, docid [L, B] cuda i32 (PAD id == n_docs) or None -> perm [B, L] cuda i32."""
        B, L = scores.shape
        perm = torch.empty(B, L, dtype=torch.int32, device=self.device)
        check(lib.ub200_pl_sample(_ptr(scores), _ptr(docid), int(n_docs), B, L, float(tau), int(seed), int(offset),
                                  _ptr(perm), _stream()), "ub200_pl_sample")
        return perm

    # ---- early read-back of the loss scalars -----------------------------------------------------------------
    def publish(self, scalars):
        """Enqueues the copy of `scalars` (<= 32 floats, device) into mapped pinned host memory + a sequence number
        (csrc/optim.cu: publish_kernel) on a side stream forked from the current one; returns nothing - read with
        `read_published()` after the step has been launched."""
        self._ensure_publish_buffers()
        cur = torch.cuda.current_stream()
        self._pub_stream.wait_stream(cur)
        with torch.cuda.stream(self._pub_stream):
            check(lib.ub200_publish(_ptr(scalars), scalars.numel(), self._pub_host.data_ptr(),
                                    self._pub_host.data_ptr() + 256, _ptr(self._pub_counter),
                                    self._pub_stream.cuda_stream), "ub200_publish")
        self._pub_n = scalars.numel()
        self._pub_forked = True

    def join_publish(self):
        """The side stream must re-join before the step ends (CUDA-graph capture needs a single sink)."""
        if getattr(self, "_pub_forked", False):       # only behind a publish of this step (a join inside a CUDA-graph
            self._pub_forked = False                  # capture must not depend on the side stream's earlier, uncaptured work)
            torch.cuda.current_stream().wait_stream(self._pub_stream)

    def read_published(self, timeout_s=20.0, lag=0):
        """Waits (spinning on the host-visible sequence number) for the publish of the step launched last (lag = 0) or
        of the one before it (lag = 1) and returns its scalars as a numpy array."""
        import time
        want = (self._pub_launched - lag) & 0xFFFFFFFF  # B200Algorithm.run_step counts the launched steps
        seq = self._pub_seq_np
        spins = 0
        t0 = None
        # sequence numbers only grow: with lag = 1 the step in flight may already have published as well
        while ((int(seq[0]) - want) & 0xFFFFFFFF) > lag:
            spins += 1
            if (spins & 0xFFF) == 0:
                if t0 is None:
                    t0 = time.perf_counter()
                elif time.perf_counter() - t0 > timeout_s:
                    torch.cuda.synchronize()          # surfaces a CUDA error if there is one
                    raise _capi.UltraB200Error("early loss read-back timed out (sequence %d, expected %d)"
                                               % (int(seq[0]), want))
        half = 32 * (want & 1)
        return self._pub_np[half:half + self._pub_n].copy()

    # ---- optimizer ------------------------------------------------------------------------------------
    def l2_term(self, l2, den, factor):
        """grads += l2 * params * (den[0] or factor); returns the device scalar sum(params^2) / 2 (hparam l2_loss)."""
        if getattr(self, "_l2_half_sumsq", None) is None:
            self._l2_half_sumsq = torch.zeros(1, dtype=torch.float32, device=self.device)
        check(lib.ub200_l2_term(_ptr(self.params), _ptr(self.grads), self.P, float(l2), _ptr(den), float(factor),
                                _ptr(self._l2_half_sumsq), _stream()), "ub200_l2_term")
        return self._l2_half_sumsq

    def clip_update(self, params, grads, state_sum, den, scale_const, max_norm, lr, mode, norm_out=None):
        n = params.numel()
        check(lib.ub200_clip_update(_ptr(params), _ptr(grads), _ptr(state_sum), n, _ptr(den), float(scale_const),
                                    float(max_norm), float(lr), int(mode), _ptr(norm_out), _ptr(self._opt_ws),
                                    self._opt_ws.numel(), _stream()), "ub200_clip_update")
