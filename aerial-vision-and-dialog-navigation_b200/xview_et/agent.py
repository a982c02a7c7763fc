"""``NavCMTAgent`` -- the training / inference agent of the HAA-Transformer
(mirror of src/xview_et/agent.py, hot-path subset).

What is here
------------
* ``train_step``: ONE fused, device-resident training iteration of BASELINE
  configs[1]: render the B*T views of the batch (stage 1) -> Darknet trunk in
  train mode (stage 2) -> ET forward, loss, backward (stage 3) -> Darknet
  backward -> (data-parallel gradient all-reduce) -> clip + AdamW.  No autograd
  graph, no host round trip besides the final loss read.
* ``forward_loss``: the same forward + loss without backward (validation).
* ``NSS`` (agent.py:256-270), ``postprocess_waypoints`` (agent.py:637-653,745-752),
  ``move_view_corners`` / ``get_direction`` (agent.py:83-101,285-384: host float64,
  restated because the reference evaluates them on the host per sample),
  ``save`` / ``load`` (agent.py:899-940; ``lang_model`` is written when one is attached, optimiser
  moments in the flat-arena format).

* ``rollout_greedy``: the student-feedback (inference / validation) loop of ``rollout``
  (agent.py:580-760) with the growing episode history: every step renders the B current views,
  runs the trunk in eval mode, appends the features and the heading to the history, runs the ET
  over the t+1 steps seen so far, discretises the waypoint and moves the view -- all on the device;
  the host reads the B stop flags once per step (the reference's ``lenths`` bookkeeping).

* ``rollout(train_ml, not_in_train, nss_w)`` / ``train(loader, n_epochs, feedback, nss_w_weighting)`` /
  ``test(loader, ...)``: the reference's agent loop with its signatures over ``self.env`` (an ``ANDHNavBatch``):
  tokenizer -> ``CustomBERTModel`` -> poses of the rollout (teacher: ground-truth actions, ``avdn_teacher_action`` +
  ``avdn_waypoint_step`` on the device; student: a no-grad greedy rollout) -> ``train_rollout_step`` (views, trunk,
  T encoder passes with loss and backward) -> trajectory dicts, ``self.loss``, ``self.logs['IL_loss']``.
"""
from __future__ import annotations

import os

import numpy as np
import torch

from .. import _lib
from ..env import ViewRenderer
from ..models.ET_haa import ET
from ..models import dark_net as DN
from ..models.dark_net import Darknet
from ..optim import FusedAdamW
from .. import parallel

PI_REF = 3.14159


def get_direction(start, end):
    """src/xview_et/agent.py:83-101 (host float64)."""
    vec = np.array(end) - np.array(start)
    if vec[1] > 0:
        ang = np.arctan(vec[0] / vec[1]) / 1.57 * 90
    elif vec[1] < 0:
        ang = np.arctan(vec[0] / vec[1]) / 1.57 * 90 + 180
    else:
        ang = 90 if np.sign(vec[0]) == 1 else 270
    return (360 - ang + 90) % 360


class NavCMTAgent:
    def __init__(self, args, rank=0, world_size=1, device=None, process_group=None):
        if not torch.cuda.is_available():
            raise RuntimeError("NavCMTAgent needs a CUDA device (sm_100); there is no CPU fallback")
        _lib.lib()
        self.args = args
        self.rank, self.world = rank, world_size
        self.pg = process_group
        self.device = torch.device(device if device is not None else "cuda")
        self.results = {}
        self.losses = []
        self.logs = {"IL_loss": []}
        self.env = []
        self.vision_model = Darknet(args.darknet_model_file, 224).to(self.device)
        wf = getattr(args, "darknet_weight_file", None)
        if wf and os.path.exists(wf):                           # agent.py:136-141
            new_state = torch.load(wf, map_location=self.device)
            state = self.vision_model.state_dict()
            state.update({k: v for k, v in new_state["model"].items() if k in state})
            self.vision_model.load_state_dict(state)
        self.vln_model = ET(args).to(self.device)
        self.vln_model.drop_rank = rank              # data-parallel replicas draw different dropout masks
        self.tokenizer = getattr(args, "tokenizer", None)     # agent.py:124 (BertTokenizerFast; no vocabulary offline)
        self.feedback = "student"
        self.env_name = ""
        self.loss = 0
        self.exposed_events = None                   # bench: [(event, event)] around the wait for the last gradient bucket
        lr = getattr(args, "lr", 1e-5)
        if getattr(args, "optim", "adamW") not in ("adamW", "adam"):
            raise ValueError("optim must be 'adam' or 'adamW' (agent.py:151)")
        if getattr(args, "optim", "adamW") == "adam":
            raise NotImplementedError("optim='adam' (no decoupled weight decay) is not implemented: use 'adamW'")
        # agent.py:153-156: one AdamW per model, default weight decay; only ET is clipped (agent.py:247)
        self.et_optimizer = FusedAdamW(self.vln_model.used_parameters(), lr=lr, max_norm=40.0)
        self.vision_model_optimizer = FusedAdamW(dict(self.vision_model.named_parameters()), lr=lr)
        self.vln_model._grad_arena = self.et_optimizer.grads
        self.vision_model._grad_arena = self.vision_model_optimizer.grads
        self.optimizers = (self.et_optimizer, self.vision_model_optimizer)
        # language encoder (agent.py:125-126,155: trained with its own AdamW): built on demand, see attach_lang_model
        self.lang_model, self.lang_optimizer = None, None
        self.renderer = ViewRenderer(self.device)
        self.use_graphs = os.environ.get("AVDN_CUDA_GRAPHS", "1") != "0"     # rollouts replay the frozen trunk pass
        self.loss_total = torch.zeros(1, dtype=torch.float64, device=self.device)
        self._bufs = {}
        self.launches = 0
        self._comm_stream = torch.cuda.Stream(self.device) if world_size > 1 else None
        self._ar_bf16 = str(getattr(args, "allreduce_dtype", os.environ.get("AVDN_ALLREDUCE_DTYPE", "fp32"))) == "bf16"
        self._ar_bufs = {}
        self._buckets = None
        if world_size > 1:
            self.broadcast_parameters()

    # ------------------------------------------------------------ distributed
    def attach_lang_model(self, lang_model=None, lr=None):
        """Put ``CustomBERTModel`` into the training loop (src/xview_et/agent.py:125-126,155,249): ``train_step``
        then takes ``input_ids`` / ``attention_mask`` (and optionally ``cls_input_ids`` / ``cls_attention_mask``: the
        reference's second pass over pre_dialogs + instructions that produces ``linear_cls``, agent.py:530-538)
        instead of ``lang`` / ``lang_cls``, and the gradients of both
        flow back into BERT and its head (one encoder pass per step = the reference's ``train_val_on_full`` mode,
        agent.py:530-538)."""
        from ..models.bert import CustomBERTModel
        self.lang_model = (lang_model if lang_model is not None else CustomBERTModel()).to(self.device)
        self.lang_optimizer = FusedAdamW(self.lang_model.used_parameters(),
                                         lr=lr if lr is not None else getattr(self.args, "lr", 1e-5))
        self.lang_model._grad_arena = self.lang_optimizer.grads
        self.lang_model._engines.clear()
        self.lang_model.drop_rank = self.rank
        self.optimizers = (self.et_optimizer, self.vision_model_optimizer, self.lang_optimizer)
        if self.world > 1:
            # every rank built (and randomly initialised) its own encoder: replicas start from rank 0's, like the
            # other two models in __init__
            parallel.broadcast_([self.lang_optimizer.p], 0, self.pg)
        return self.lang_model

    def broadcast_parameters(self):
        """Replicas start identical (DDP semantics): rank 0's arenas and BN buffers."""
        parallel.broadcast_([opt.p for opt in self.optimizers], 0, self.pg)
        parallel.broadcast_(list(self.vision_model.buffers()), 0, self.pg)

    def _trunk_buckets(self, eng):
        """Slices of the trunk's gradient arena in backward-completion order.  The deep
        7x7 / 14x14 blocks hold most parameters and finish first; the shallow blocks
        hold most of the time: 3 buckets let NCCL run under the rest of the backward."""
        if self._buckets is None:
            offs = self.vision_model_optimizer.offsets
            first = [offs[f"module_list.{L.idx}.conv_{L.idx}.weight"][0] for L in eng.layers]
            self._buckets = parallel.bucket_edges(first, self.vision_model_optimizer.n)
        return self._buckets

    def _allreduce_async(self, flat, lo, hi):
        """Sum ``flat[lo:hi]`` (a slice of a gradient arena) over the ranks on the communication stream, behind
        everything enqueued so far.  ``args.allreduce_dtype == 'bf16'`` (or AVDN_ALLREDUCE_DTYPE=bf16) ships the
        bucket as bf16 -- half the NVLink bytes; every rank receives the same rounded sum, so the replicas stay
        bit-identical -- the default is fp32."""
        cs = self._comm_stream
        cs.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(cs):
            if self._ar_bf16:
                buf = self._ar_bufs.get(flat.data_ptr())
                if buf is None:
                    buf = torch.empty(flat.numel(), dtype=torch.bfloat16, device=flat.device)
                    self._ar_bufs[flat.data_ptr()] = buf
                buf[lo:hi].copy_(flat[lo:hi])
                parallel.allreduce_sum_(buf, lo, hi, self.pg)
                flat[lo:hi].copy_(buf[lo:hi])
            else:
                parallel.allreduce_sum_(flat, lo, hi, self.pg)

    # ----------------------------------------------------------------- buffers
    def _get_bufs(self, B, T):
        key = (B, T)
        b = self._bufs.get(key)
        if b is None:
            dev = self.device
            b = dict(
                x=torch.empty((B * T, 224, 224, 4), dtype=torch.bfloat16, device=dev),
                att=torch.empty((B, 224, 224), dtype=torch.uint8, device=dev),
                frames=torch.empty((B * T, 512, 7, 7), dtype=torch.float32, device=dev),
                d_frames=torch.empty((B * T, 512, 49), dtype=torch.float32, device=dev),
                d_output=torch.empty((B, 4), dtype=torch.float32, device=dev),
                d_h_sali=torch.empty((B, 64), dtype=torch.float32, device=dev),
                loss_i=torch.empty(B, dtype=torch.float64, device=dev),
            )
            self._bufs[key] = b
        return b

    # ------------------------------------------------------------------- steps
    def _forward(self, batch, train):
        """Stages 1-3 forward + loss.  ``batch`` (device tensors unless noted):
        ``corners_px`` i32 [B,T,4,2] pose corners (FL,FR,BR,BL; pixel x,y) or ``images``
        bf16 [B*T,224,224,4] (pre-rendered, normalised NHWC); ``tile_idx`` i32 [B,T] or None;
        ``lang`` f32 [B,L,768]; ``lang_cls`` f32 [B,49]; ``directions`` f32 [B,T,2];
        ``lenths`` host list[int]; ``gt_xy`` [B,2], ``gt_alt`` [B], ``gt_prog`` [B] f32;
        optional ``att`` u8 [B,224,224], ``jitter`` f32 [B]."""
        ptr = _lib.ptr
        self._lang_eng = None
        if "input_ids" in batch:
            if self.lang_model is None:
                raise RuntimeError("batch carries input_ids but no language model is attached (attach_lang_model)")
            lm = self.lang_model
            lm.train(bool(train) and not getattr(self.args, "no_dropout", False))
            ids, am = batch["input_ids"], batch["attention_mask"]
            leng = lm.engine(ids.shape[0], ids.shape[1], self.device)
            leng.set_dropout(*lm.dropout_config())
            l0l = leng.launches
            seq, lin, _ = leng.forward(ids.long(), am)
            self.launches += leng.launches - l0l
            self._lang_eng = leng
            self._lang_eng_cls = None
            if "cls_input_ids" in batch:
                # agent.py:530-538: ``linear_cls`` comes from a second pass over pre_dialogs + instructions
                ids2, am2 = batch["cls_input_ids"], batch["cls_attention_mask"]
                leng2 = lm.engine(ids2.shape[0], ids2.shape[1], self.device, slot=1)
                leng2.set_dropout(*lm.dropout_config())
                l0l = leng2.launches
                _, lin, _ = leng2.forward(ids2.long(), am2)
                self.launches += leng2.launches - l0l
                self._lang_eng_cls = leng2
            batch = dict(batch, lang=seq, lang_cls=lin)
        lang, lang_cls, dirs = batch["lang"], batch["lang_cls"], batch["directions"]
        B, T = dirs.shape[0], dirs.shape[1]
        L = lang.shape[1]
        bufs = self._get_bufs(B, T)
        n = 0
        att = batch.get("att")
        if "images" in batch:
            x = batch["images"]
        else:
            r = self.renderer
            c = batch["corners_px"].view(B * T, 4, 2)
            ti = batch.get("tile_idx")
            minv = r.homography(c)
            r.render(None, None if ti is None else ti.view(-1), views=False, norm_nhwc=True, minv=minv,
                     out={"norm_nhwc": bufs["x"]})
            n += 2
            x = bufs["x"]
            if att is None and self.nss_w != 0:
                # human-attention target of the current (last) step of every episode (env.py:292-293)
                last = minv.view(B, T, 3, 3)[:, -1].contiguous()
                r.render(None, None if ti is None else ti.view(B, T)[:, -1].contiguous(), views=False, att=True,
                         minv=last, out={"att": bufs["att"]})
                att = bufs["att"]
                n += 2
        vm = self.vision_model
        teng = vm.engine(B * T, 224, 224, self.device)
        l0 = teng.launches
        DN._trunk_forward(vm, teng, x, train, out=bufs["frames"])
        et = self.vln_model
        eng = et.engine(B, L, T, self.device)
        e0 = eng.launches
        et.train(bool(train) and not getattr(self.args, "no_dropout", False))
        eng.set_dropout(*et.dropout_config())
        output, h_sali = eng.forward(bufs["frames"].view(B * T, 512, 49), lang, lang_cls, dirs, batch["lenths"],
                                     et.encoder_vl.enc_pos.pe[0])
        self.loss_total.zero_()
        scale = float(self.train_ml) / B                               # agent.py:883-885
        _lib.call("avdn_loss", ptr(output), ptr(h_sali), ptr(batch["gt_xy"]), ptr(batch["gt_alt"]),
                  ptr(batch["gt_prog"]), ptr(att), ptr(batch.get("jitter")), B, float(self.nss_w),
                  int(getattr(self.args, "nss_r", 0)), scale, ptr(self.loss_total), ptr(bufs["loss_i"]),
                  ptr(bufs["d_output"]), ptr(bufs["d_h_sali"]))
        n += 2
        self._ctx = (teng, eng, bufs, l0, e0)
        self.launches += n
        return output, h_sali

    @property
    def nss_w(self):
        ov = getattr(self, "_nss_w_override", None)
        return float(ov if ov is not None else getattr(self.args, "nss_w", 1.0))        # parser.py:38 default 1

    @property
    def train_ml(self):
        ov = getattr(self, "_train_ml_override", None)
        return float(ov if ov is not None else getattr(self.args, "ml_weight", 0.2))     # parser.py:54

    class _Weights:
        """``with agent._weights(nss_w, train_ml):`` -- the per-rollout loss weights the reference passes to
        ``rollout(train_ml=..., nss_w=...)`` (agent.py:229-235); ``None`` keeps the value from ``args``."""

        def __init__(self, agent, nss_w, train_ml):
            self.a, self.v = agent, (nss_w, train_ml)

        def __enter__(self):
            self.old = (getattr(self.a, "_nss_w_override", None), getattr(self.a, "_train_ml_override", None))
            self.a._nss_w_override, self.a._train_ml_override = self.v

        def __exit__(self, *exc):
            self.a._nss_w_override, self.a._train_ml_override = self.old

    def _weights(self, nss_w=None, train_ml=None):
        return NavCMTAgent._Weights(self, nss_w, train_ml)

    def forward_loss(self, batch):
        """Validation: forward + loss (trunk in eval mode).  Returns (loss, output, h_sali)."""
        self.vision_model.eval()
        output, h_sali = self._forward(batch, False)
        return self.loss_total.clone(), output, h_sali

    def train_step(self, batch, sync_loss=False):
        """One training iteration (see the module docstring).  Returns the device
        tensor holding the step loss (float64 [1]); ``sync_loss=True`` returns a host float."""
        self.vision_model.train()
        for opt in self.optimizers:
            opt.zero_grad()
        self.launches += 2
        self._forward(batch, True)
        teng, eng, bufs, l0, e0 = self._ctx
        leng = self._lang_eng
        if leng is None:
            eng.backward(bufs["d_output"], bufs["d_h_sali"], d_frames=bufs["d_frames"])
        else:
            d_cls = torch.zeros((eng.B, 49), dtype=torch.float32, device=self.device)
            _, d_lang = eng.backward(bufs["d_output"], bufs["d_h_sali"], d_frames=bufs["d_frames"], need_lang_grad=True,
                                     d_lang_cls=d_cls)
            leng2 = self._lang_eng_cls
            l0l = leng.launches
            if leng2 is None:
                leng.backward(d_lang, d_cls, None)
            else:                          # both passes accumulate into the one gradient arena
                leng.backward(d_lang, None, None)
                l1 = leng2.launches
                leng2.backward(None, d_cls, None)
                self.launches += leng2.launches - l1
            self.launches += leng.launches - l0l
        dp = self.world > 1
        if dp:
            if leng is not None:
                self._allreduce_async(self.lang_optimizer.g, 0, self.lang_optimizer.n)
            self._allreduce_async(self.et_optimizer.g, 0, self.et_optimizer.n)
            buckets = {c: (lo, hi) for c, lo, hi in self._trunk_buckets(teng)}
            hook = lambda li: (self._allreduce_async(self.vision_model_optimizer.g, *buckets[li])
                               if li in buckets else None)
        else:
            hook = None
        DN._trunk_backward(self.vision_model, teng, bufs["d_frames"].view(-1, 512, 7, 7), after_layer=hook,
                           flush_layers=set(buckets) if dp else None)
        if dp:
            self._wait_comm()
        gs = 1.0 / self.world
        for opt in self.optimizers:
            self.launches += opt.step(grad_scale=gs)
        self.launches += (teng.launches - l0) + (eng.launches - e0)
        self._log_loss()
        if sync_loss:
            return float(self.loss_total.item())
        return self.loss_total

    def _wait_comm(self):
        """The optimiser waits for the last gradient bucket; with ``exposed_events`` set (bench) the wait is bracketed
        by CUDA events: the all-reduce time the backward pass did not hide."""
        cur = torch.cuda.current_stream()
        if self.exposed_events is not None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(cur)
            cur.wait_stream(self._comm_stream)
            e1.record(cur)
            self.exposed_events.append((e0, e1))
        else:
            cur.wait_stream(self._comm_stream)

    LOG_KEEP = 4096

    def _log_loss(self):
        """``self.logs['IL_loss']`` (agent.py:885): one entry per rollout.  Device scalars (a copy: ``loss_total`` is
        rewritten every step), converted lazily by ``il_losses()``; bounded."""
        log = self.logs["IL_loss"]
        log.append(self.loss_total.clone())
        if len(log) > self.LOG_KEEP:
            del log[: len(log) - self.LOG_KEEP]

    def il_losses(self):
        """The logged losses as host floats (one synchronisation)."""
        log = self.logs["IL_loss"]
        return torch.stack([x.reshape(()) for x in log]).cpu().tolist() if log else []

    def train_iteration(self, teacher_batch, student_batch=None, sync_loss=False, feedback="student",
                        nss_w_weighting=1):
        """One iteration of ``train`` (agent.py:225-251).  ``feedback='student'`` (the reference's recipe): the
        teacher-feedback rollout with ``train_ml = args.ml_weight`` and **nss_w = 0** (agent.py:232), then the
        student-feedback rollout with ``nss_w = args.nss_w * nss_w_weighting`` (agent.py:235); the two add their
        losses, ONE backward's worth of gradients (accumulated over both), clip, one step of every optimiser.
        ``feedback='teacher'``: the teacher rollout alone with ``train_ml = args.teacher_weight`` and the NSS term
        (agent.py:229).  ``student_batch`` comes from ``student_batch()``.  Returns the summed loss."""
        nss = float(getattr(self.args, "nss_w", 1.0)) * nss_w_weighting
        if feedback == "teacher":
            tot = self.train_rollout_step(teacher_batch, nss_w=nss,
                                          train_ml=float(getattr(self.args, "teacher_weight", 1.0))).clone()
        elif feedback == "student":
            l1 = self.train_rollout_step(teacher_batch, step=False, nss_w=0.0).clone()
            tot = l1 + self.train_rollout_step(student_batch, zero=False, nss_w=nss)
        else:
            raise ValueError("Invalid feedback option")
        return float(tot.item()) if sync_loss else tot

    def train_rollout_step(self, batch, sync_loss=False, zero=True, step=True, nss_w=None, train_ml=None,
                           collect=None, bn_per_step=None):
        with self._weights(nss_w, train_ml):
            return self._train_rollout_step(batch, sync_loss, zero, step, collect, bn_per_step)

    def _train_rollout_step(self, batch, sync_loss=False, zero=True, step=True, collect=None, bn_per_step=None):
        """One teacher-forced training ROLLOUT (agent.py:580-760, 883-885, 245-251) as one batched pass: the loss is
        taken at EVERY step t of every episode on the history ``[:t+1]``, summed over steps and samples and scaled
        by ``ml_weight / B``, as the reference accumulates ``ml_loss`` over its step loop.

        The poses of all steps are known before the pass (teacher feedback: the ground-truth path; student
        feedback: a no-grad ``rollout_greedy`` -- see ``student_batch``), so the B*T views are rendered and pushed
        through the trunk ONCE (forward and backward); the transformer then makes the reference's T encoder calls
        over the growing history, each followed at once by its loss and its backward pass; the gradient of a
        frame is the sum over the steps whose history contains it.  (The reference's train-mode trunk normalises
        each step's B views with their own batch statistics; here the statistics are over all B*T views,
        DESIGN.md §8.)

        ``batch``: ``corners_px`` i32 [B,T,4,2] (or ``images``), ``tile_idx`` i32 [B,T] | None, ``lang`` [B,L,768],
        ``lang_cls`` [B,49] (or, with an attached language model, ``input_ids`` / ``attention_mask`` and optionally
        ``cls_input_ids`` / ``cls_attention_mask``: the encoder runs once per rollout and receives the gradients of
        every step), ``directions`` f32 [B,T,2], ``gt_xy`` [B,T,2], ``gt_alt`` / ``gt_prog`` [B,T], optional
        ``lenths`` host int [B][T] (history length seen at step t: stops growing once a sample has ended,
        agent.py:603-620; default t+1), ``att`` u8 [B,T,224,224] (default: rendered from the poses), ``jitter``
        f32 [B,T].  ``zero=False`` keeps the gradients already in the arenas, ``step=False`` leaves them there
        without an optimiser step (``train_iteration`` chains the two rollouts of an iteration that way).
        ``nss_w`` / ``train_ml`` (``train_rollout_step``): this rollout's loss weights (default: ``args.nss_w``,
        ``args.ml_weight``); with ``nss_w == 0`` no human-attention maps are rendered.  ``collect``: a list that
        receives the per-step ``output`` [B,4] tensors (the trajectory log of ``rollout``).
        ``bn_per_step`` (default ``args.bn_per_step``, false): the reference's BatchNorm semantics -- one train-mode
        trunk pass per time step over that step's B views (its own batch statistics, one running-statistics update
        per step, agent.py:593), each pass kept for its own backward -- instead of one pass over all B*T views."""
        ptr = _lib.ptr
        self.vision_model.train()
        n = 0
        if zero:
            for opt in self.optimizers:
                opt.zero_grad()
            n += 2
        leng = leng2 = None
        if "input_ids" in batch:
            # the language encoder runs once per rollout (agent.py:519-538) and collects the gradients of all steps
            if self.lang_model is None:
                raise RuntimeError("batch carries input_ids but no language model is attached (attach_lang_model)")
            lm = self.lang_model
            lm.train(not getattr(self.args, "no_dropout", False))
            ids, am = batch["input_ids"], batch["attention_mask"]
            leng = lm.engine(ids.shape[0], ids.shape[1], self.device)
            leng.set_dropout(*lm.dropout_config())
            l0l = leng.launches
            seq, lin, _ = leng.forward(ids.long(), am)
            n += leng.launches - l0l
            if "cls_input_ids" in batch:
                ids2, am2 = batch["cls_input_ids"], batch["cls_attention_mask"]
                leng2 = lm.engine(ids2.shape[0], ids2.shape[1], self.device, slot=1)
                leng2.set_dropout(*lm.dropout_config())
                l0l = leng2.launches
                _, lin, _ = leng2.forward(ids2.long(), am2)
                n += leng2.launches - l0l
            batch = dict(batch, lang=seq, lang_cls=lin)
        dirs = batch["directions"].contiguous().float()
        B, T = dirs.shape[0], dirs.shape[1]
        BT = B * T
        lang, lang_cls = batch["lang"].contiguous().float(), batch["lang_cls"].contiguous().float()
        L = lang.shape[1]
        dev = self.device
        bufs = self._get_bufs(B, T)
        key = ("rollout_train", B, T)
        xb = self._bufs.get(key)
        if xb is None:
            xb = dict(att=torch.empty((BT, 224, 224), dtype=torch.uint8, device=dev),
                      d_frames=torch.empty((BT, 512, 49), dtype=torch.float32, device=dev))
            self._bufs[key] = xb
        att = batch.get("att")
        if "images" in batch:
            x = batch["images"]
        else:
            r = self.renderer
            ti = batch.get("tile_idx")
            ti = None if ti is None else ti.reshape(-1)
            minv = r.homography(batch["corners_px"].view(BT, 4, 2))
            r.render(None, ti, views=False, norm_nhwc=True, minv=minv, out={"norm_nhwc": bufs["x"]})
            n += 2
            x = bufs["x"]
            if att is None and self.nss_w != 0:
                r.render(None, ti, views=False, att=True, minv=minv, out={"att": xb["att"]})
                att = xb["att"]
                n += 1
        vm, et = self.vision_model, self.vln_model
        if bn_per_step is None:
            bn_per_step = bool(getattr(self.args, "bn_per_step", False))
        if bn_per_step:
            # the reference's loop: Darknet sees the B views of one time step per call (agent.py:593)
            key = ("rollout_steps", B, T)
            sb = self._bufs.get(key)
            if sb is None:
                sb = dict(x=[torch.empty((B, 224, 224, 4), dtype=torch.bfloat16, device=dev) for _ in range(T)],
                          f=torch.empty((B, 512, 7, 7), dtype=torch.float32, device=dev),
                          df=torch.empty((B, 512, 7, 7), dtype=torch.float32, device=dev))
                self._bufs[key] = sb
            tengs = [vm.engine(B, 224, 224, dev, slot=1 + t) for t in range(T)]
            l0 = sum(e.launches for e in tengs)
            x5 = x.view(B, T, 224, 224, 4)
            f5 = bufs["frames"].view(B, T, 512, 49)
            for t in range(T):
                sb["x"][t].copy_(x5[:, t])
                DN._trunk_forward(vm, tengs[t], sb["x"][t], True, out=sb["f"])
                f5[:, t].copy_(sb["f"].view(B, 512, 49))
            n += 2 * T
            teng = tengs[0]
        else:
            tengs = None
            teng = vm.engine(BT, 224, 224, dev)
            l0 = teng.launches
            DN._trunk_forward(vm, teng, x, True, out=bufs["frames"])
        # ---- step t: the encoder over the history [:t+1], its loss, and straight back (the loss is a sum over steps,
        #      so only one step's activations are alive at a time; every step adds into the one gradient arena) ----
        frames = bufs["frames"].view(B, T, 512, 49)
        d_frames = bufs["d_frames"].view(B, T, 512, 49)
        d_frames.zero_()
        lens = batch.get("lenths")
        gt_xy = batch["gt_xy"].float().transpose(0, 1).contiguous()     # [T,B,2]
        gt_alt = batch["gt_alt"].float().transpose(0, 1).contiguous()
        gt_prog = batch["gt_prog"].float().transpose(0, 1).contiguous()
        jit = batch.get("jitter")
        jit = None if jit is None else jit.float().transpose(0, 1).contiguous()
        att_t = None if att is None else att.view(B, T, 224, 224).transpose(0, 1).contiguous()
        n += 6
        et.train(not getattr(self.args, "no_dropout", False))
        self.loss_total.zero_()
        scale = float(self.train_ml) / B                               # agent.py:883-885
        pe = et.encoder_vl.enc_pos.pe[0]
        et_launches = 0
        if leng is not None:
            d_lang_sum = torch.zeros((B, L, 768), dtype=torch.float32, device=dev)
            d_cls_sum = torch.zeros((B, 49), dtype=torch.float32, device=dev)      # every step adds into it
        for t in range(T):
            Tc = t + 1
            eng = et.engine(B, L, Tc, dev)
            e0 = eng.launches
            eng.set_dropout(*et.dropout_config())
            f_t = frames[:, :Tc].contiguous().view(B * Tc, 512, 49)
            output, h_sali = eng.forward(f_t, lang, lang_cls, dirs[:, :Tc].contiguous(),
                                         [int(lens[i][t]) for i in range(B)] if lens is not None else [Tc] * B, pe)
            _lib.call("avdn_loss", ptr(output), ptr(h_sali), ptr(gt_xy[t]), ptr(gt_alt[t]), ptr(gt_prog[t]),
                      ptr(None if att_t is None else att_t[t]), ptr(None if jit is None else jit[t]), B,
                      float(self.nss_w), int(getattr(self.args, "nss_r", 0)), scale, ptr(self.loss_total),
                      ptr(bufs["loss_i"]), ptr(bufs["d_output"]), ptr(bufs["d_h_sali"]))
            if collect is not None:
                collect.append(output.detach().clone())
            if leng is None:
                df, _ = eng.backward(bufs["d_output"], bufs["d_h_sali"], d_frames=xb["d_frames"][:B * Tc])
            else:
                df, d_lang = eng.backward(bufs["d_output"], bufs["d_h_sali"], d_frames=xb["d_frames"][:B * Tc],
                                          need_lang_grad=True, d_lang_cls=d_cls_sum)
                d_lang_sum += d_lang.view(B, L, 768)
                n += 1
            d_frames[:, :Tc] += df.view(B, Tc, 512, 49)
            n += 5
            et_launches += eng.launches - e0
        e0 = eng.launches - et_launches
        if leng is not None:
            l0l = leng.launches
            if leng2 is None:
                leng.backward(d_lang_sum, d_cls_sum, None)
            else:
                leng.backward(d_lang_sum, None, None)
                l1 = leng2.launches
                leng2.backward(None, d_cls_sum, None)
                n += leng2.launches - l1
            n += leng.launches - l0l
        self._ctx = (teng, eng, bufs, l0, e0)
        dp = self.world > 1 and step               # gradients are reduced once, after the iteration's last rollout
        if dp:
            if self.lang_optimizer is not None:
                self._allreduce_async(self.lang_optimizer.g, 0, self.lang_optimizer.n)
            self._allreduce_async(self.et_optimizer.g, 0, self.et_optimizer.n)
            buckets = {c: (lo, hi) for c, lo, hi in self._trunk_buckets(teng)}
            hook = lambda li: (self._allreduce_async(self.vision_model_optimizer.g, *buckets[li])
                               if li in buckets else None)
        else:
            hook = None
        if tengs is None:
            DN._trunk_backward(vm, teng, bufs["d_frames"].view(-1, 512, 7, 7), after_layer=hook,
                               flush_layers=set(buckets) if dp else None)
            trunk_launches = teng.launches - l0
        else:
            # one backward pass per time step, each through the activations of its own forward pass; the parameter
            # gradients accumulate, the last pass releases the data-parallel buckets
            sb = self._bufs[("rollout_steps", B, T)]
            d5 = bufs["d_frames"].view(B, T, 512, 49)
            for t in reversed(range(T)):
                sb["df"].view(B, 512, 49).copy_(d5[:, t])
                last = (t == 0)
                DN._trunk_backward(vm, tengs[t], sb["df"], after_layer=hook if last else None,
                                   flush_layers=set(buckets) if (dp and last) else None)
            n += T
            trunk_launches = sum(e.launches for e in tengs) - l0
        if dp:
            self._wait_comm()
        if step:
            gs = 1.0 / self.world
            for opt in self.optimizers:
                n += opt.step(grad_scale=gs)
        self.launches += n + trunk_launches + (eng.launches - e0)
        self._log_loss()
        return float(self.loss_total.item()) if sync_loss else self.loss_total

    @torch.no_grad()
    def student_batch(self, batch, gt_path_corners, max_action_len=None):
        """The student-feedback half of a training iteration (agent.py:245-251 runs ``rollout`` twice): a no-grad
        greedy rollout fixes the poses, ``teacher_action`` supplies the target of every step (one launch), and the
        result is a ``train_rollout_step`` batch.  ``batch`` as for ``rollout_greedy``; returns the training batch
        (device tensors; ``lenths`` from the steps each sample was alive)."""
        res = self.rollout_greedy(batch, max_action_len=max_action_len)
        T = int(res["steps"])
        corners = res["corners"][:T]                                   # [T,B,4,2] gps
        B = corners.shape[1]
        ended_before = torch.zeros((T, B), dtype=torch.uint8, device=self.device)
        ended_before[1:] = res["ended"][:T - 1]
        tgt_xy, tgt_alt, prog = self.teacher_action(corners.reshape(T * B, 4, 2),
                                                    [gt_path_corners[k % B] for k in range(T * B)],
                                                    ended_before.reshape(-1), feedback="student")
        geo = batch["geo"].contiguous()
        px = torch.empty((T * B, 4, 2), dtype=torch.int32, device=self.device)
        _lib.call("avdn_gps_to_pixels", _lib.ptr(corners.reshape(T * B, 4, 2).contiguous()),
                  _lib.ptr(geo.repeat(T, 1).contiguous()), T * B, _lib.ptr(px))
        rad = res["directions"][:T].float() / 180 * PI_REF
        alive = (ended_before == 0).cpu().numpy()
        lens = np.cumsum(alive, axis=0).T                              # [B,T]: history length seen at step t
        tb = lambda a: a.view(T, B, *a.shape[1:]).transpose(0, 1).contiguous()
        ti = batch.get("tile_idx")
        return dict(corners_px=tb(px), tile_idx=None if ti is None else ti.view(B, 1).expand(B, T).contiguous(),
                    lang=batch["lang"], lang_cls=batch["lang_cls"],
                    directions=torch.stack([torch.sin(rad), torch.cos(rad)], -1).transpose(0, 1).contiguous(),
                    gt_xy=tb(tgt_xy.float()), gt_alt=tb(tgt_alt.float()), gt_prog=tb(prog.float()),
                    lenths=np.maximum(lens, 1).tolist())

    # ----------------------------------------------------------------- inference
    STOP_THRESHOLD = 0.5               # agent.py:738-745 (the LSTM agent uses 0.25)

    @torch.no_grad()
    def rollout_greedy(self, batch, max_action_len=None, stop_threshold=None, incremental=True):
        """Greedy (student-feedback) rollout of ``B`` episodes, the inference path of the HAA-Transformer
        (agent.py:580-760 with ``feedback == 'student'``, no teacher / loss).

        ``batch`` (device tensors): ``corners_gps`` f64 [B,4,2] (lat,lng; FL,FR,BR,BL = gt_path_corners[0]),
        ``directions`` [B] (starting_angle, degrees), ``geo`` f64 [B,5] = (bl_lat, bl_lng, tr_lat, tr_lng,
        lat_ratio), ``tile_idx`` i32 [B] or None, ``lang`` f32 [B,L,768], ``lang_cls`` f32 [B,49].

        Step t: observation at the current corners (env._get_obs, agent.py:769) -> Darknet, eval mode ->
        ``frames`` / ``directions`` grow by one step for EVERY sample, ``lenths`` only for the samples that
        have not ended (agent.py:603-620) -> ``ET`` over the t+1 steps -> post-processing, stop test
        (progress > 0.5 or last step) and ``move_view_corners`` (agent.py:637-653,736-760).  The loop ends
        early once every sample has ended (agent.py:773-774).

        ``incremental=True`` (default): step 0 runs the full encoder once, later steps compute only their
        two new rows per sample against cached keys / values (``ETDecodeState``: earlier rows cannot change
        under the causal mask); ``False`` re-runs the encoder over the whole history every step, exactly as
        the reference does.  The two agree to bf16 rounding for every sample that has not ended.

        Returns device tensors: ``corners`` [T+1,B,4,2], ``directions`` [T+1,B], ``ended`` [T,B],
        ``output`` [T,B,4], ``angle`` / ``altitude`` [T,B], ``dist`` [T,B], and ``steps`` (the number of
        steps actually run).  Rows after ``steps`` repeat the final state."""
        ptr, call = _lib.ptr, _lib.call
        T = int(max_action_len if max_action_len is not None else getattr(self.args, "max_action_len", 20))
        thr = float(self.STOP_THRESHOLD if stop_threshold is None else stop_threshold)
        lang, lang_cls = batch["lang"].contiguous().float(), batch["lang_cls"].contiguous().float()
        B, L = lang.shape[0], lang.shape[1]
        dev = self.device
        key = ("rollout", B, T)
        bf = self._bufs.get(key)
        if bf is None:
            f64, i32 = torch.float64, torch.int32
            bf = dict(x=torch.empty((B, 224, 224, 4), dtype=torch.bfloat16, device=dev),
                      frames=torch.empty((B, 512, 7, 7), dtype=torch.float32, device=dev),
                      frames_hist=torch.zeros((T, B, 512, 49), dtype=torch.float32, device=dev),
                      dirs_hist=torch.zeros((T, B, 2), dtype=torch.float32, device=dev),
                      corners=torch.empty((B, 4, 2), dtype=f64, device=dev),
                      cur_dir=torch.empty(B, dtype=f64, device=dev),
                      ended=torch.empty(B, dtype=torch.uint8, device=dev),
                      px=torch.empty((B, 4, 2), dtype=i32, device=dev),
                      minv=torch.empty((B, 3, 3), dtype=f64, device=dev),
                      corners_hist=torch.empty((T + 1, B, 4, 2), dtype=f64, device=dev),
                      dir_hist=torch.empty((T + 1, B), dtype=f64, device=dev),
                      ended_hist=torch.empty((T, B), dtype=torch.uint8, device=dev),
                      output_hist=torch.zeros((T, B, 4), dtype=torch.float32, device=dev),
                      angle=torch.zeros((T, B), dtype=i32, device=dev),
                      altitude=torch.zeros((T, B), dtype=i32, device=dev),
                      dist=torch.zeros((T, B), dtype=f64, device=dev))
            self._bufs[key] = bf
        geo = batch["geo"].contiguous()
        bounds = geo[:, :4].contiguous()
        ti = batch.get("tile_idx")
        bf["corners"].copy_(batch["corners_gps"])
        bf["cur_dir"].copy_(batch["directions"])
        bf["ended"].zero_()
        vm, et, r = self.vision_model, self.vln_model, self.renderer
        vm.eval()
        et.eval()
        teng = vm.engine(B, 224, 224, dev)
        l0 = teng.launches
        lens = [0] * B
        ended_host = [False] * B
        n, steps = 0, 0
        for t in range(T):
            steps = t + 1
            bf["corners_hist"][t].copy_(bf["corners"])
            bf["dir_hist"][t].copy_(bf["cur_dir"])
            # ---- observation -> trunk features of the step ----
            call("avdn_gps_to_pixels", ptr(bf["corners"]), ptr(geo), B, ptr(bf["px"]))
            call("avdn_homography_from_corners", ptr(bf["px"]), B, ptr(bf["minv"]))
            r.render(None, ti, views=False, norm_nhwc=True, minv=bf["minv"], out={"norm_nhwc": bf["x"]})
            if t == 0 or not self.use_graphs:
                DN._trunk_forward(vm, teng, bf["x"], False, out=bf["frames"], frozen=(t > 0))
            else:
                DN._trunk_forward_graphed(vm, teng, bf["x"], bf["frames"])
            bf["frames_hist"][t].copy_(bf["frames"].view(B, 512, 49))
            rad = bf["cur_dir"].float() / 180 * PI_REF                  # agent.py:605-606 (float32, pi = 3.14159)
            bf["dirs_hist"][t, :, 0] = torch.sin(rad)
            bf["dirs_hist"][t, :, 1] = torch.cos(rad)
            for i in range(B):
                if not ended_host[i]:
                    lens[i] += 1
            # ---- ET over the history [B, t+1, ...] ----
            Tc = t + 1
            pe = et.encoder_vl.enc_pos.pe[0]
            if incremental and t > 0:
                dec = et.decoder(B, L, T, dev)
                e0 = dec.launches
                output, _ = dec.step(t, bf["frames_hist"][t], bf["dirs_hist"][t], lang_cls, pe)
                n += dec.launches - e0
            else:
                frames = bf["frames_hist"][:Tc].permute(1, 0, 2, 3).contiguous().view(B * Tc, 512, 49)
                dirs = bf["dirs_hist"][:Tc].permute(1, 0, 2).contiguous()
                eng = et.engine(B, L, Tc, dev)
                e0 = eng.launches
                eng.set_dropout(0.0, 0.0, 0)
                output, _ = eng.forward(frames, lang, lang_cls, dirs, lens if not incremental else [1] * B, pe)
                n += eng.launches - e0
                if incremental:
                    et.decoder(B, L, T, dev).start()
            bf["output_hist"][t].copy_(output)
            # ---- simulator: post-processing, stop test, move_view_corners ----
            call("avdn_waypoint_step", ptr(output), ptr(bf["corners"]), ptr(bounds), ptr(bf["cur_dir"]),
                 ptr(bf["ended"]), B, thr, int(t == T - 1), ptr(bf["angle"][t]), ptr(bf["dist"][t]),
                 ptr(bf["altitude"][t]))
            bf["ended_hist"][t].copy_(bf["ended"])
            n += 12
            ended_host = [bool(v) for v in bf["ended"].cpu().tolist()]  # the step's only host read
            if all(ended_host):
                break
        for t2 in range(steps, T + 1):
            bf["corners_hist"][t2].copy_(bf["corners"])
            bf["dir_hist"][t2].copy_(bf["cur_dir"])
        for t2 in range(steps, T):
            bf["ended_hist"][t2].copy_(bf["ended"])
        self.launches += n + (teng.launches - l0)
        return dict(corners=bf["corners_hist"], directions=bf["dir_hist"], ended=bf["ended_hist"],
                    output=bf["output_hist"], angle=bf["angle"], altitude=bf["altitude"], dist=bf["dist"],
                    steps=steps)

    @staticmethod
    def trajectories(res):
        """Host view of a rollout result: per sample the list of (corners [4,2], direction) the reference
        appends to ``traj['path_corners']`` (agent.py:563,762-767)."""
        corners = res["corners"].cpu().numpy()
        dirs = res["directions"].cpu().numpy()
        ended = res["ended"].cpu().numpy().astype(bool)
        T, B = ended.shape
        out = []
        for i in range(B):
            path = [(corners[0, i], dirs[0, i])]
            for t in range(min(T, int(res.get("steps", T)))):
                if not ended[t, i]:
                    path.append((corners[t + 1, i], dirs[t + 1, i]))
            out.append(path)
        return out

    def test(self, batches, env_name="no_name_provided", feedback="student", not_in_train=False, **kwargs):
        """``agent.test`` (agent.py:191-206): greedy rollouts in eval mode, one trajectory dict per
        episode in ``self.results[instr_id]`` -- the structure ``ANDHNavBatch.eval_metrics`` scores.

        With ``self.env`` set to an ``ANDHNavBatch`` this is the reference's loop: ``batches`` is the data loader (it
        fills ``self.env`` as it is iterated) and every item triggers ``self.rollout(not_in_train=True)``.  Otherwise:

        Every batch is the ``rollout_greedy`` input dict plus host metadata: ``instr_id`` list[str],
        ``gt_path_corners`` list of ``[n_i,4,2]`` arrays (optional: without it ``gt_progress`` is omitted, as for
        the unseen test split, agent.py:655) and ``num_dia`` list[int] (optional).  ``gt_progress`` is the
        reference's per-step IoU of the current view with the goal (``teacher_action``'s progress, agent.py:660,
        721), evaluated for all steps of the batch in one ``avdn_teacher_action`` launch."""
        self.feedback, self.env_name = feedback, env_name
        self.results, self.losses = {}, []
        if hasattr(self.env, "_get_obs"):
            self.vln_model.eval(); self.vision_model.eval()
            if self.lang_model is not None:
                self.lang_model.eval()
            self.loss = 0
            for _ in batches:
                for traj in self.rollout(not_in_train=True, **kwargs):
                    self.loss = 0
                    self.results[traj["instr_id"]] = traj
            return self.results
        for batch in batches:
            meta = {k: batch[k] for k in ("instr_id", "gt_path_corners", "num_dia") if k in batch}
            res = self.rollout_greedy({k: v for k, v in batch.items() if k not in meta}, **kwargs)
            steps = int(res["steps"])
            corners = res["corners"]                                       # [T+1,B,4,2]
            B = corners.shape[1]
            ended = res["ended"].cpu().numpy().astype(bool)
            prog = None
            gts = meta.get("gt_path_corners")
            if gts is not None:
                flat = corners[:steps].reshape(steps * B, 4, 2)
                _, _, pr = self.teacher_action(flat, [gts[i % B] for i in range(steps * B)],
                                               np.zeros(steps * B, dtype=np.uint8), feedback="student")
                prog = pr.view(steps, B).cpu().numpy()
            c_host = corners.cpu().numpy()
            d_host = res["directions"].cpu().numpy()
            out_host = res["output"].cpu().numpy()
            for i in range(B):
                traj = {"instr_id": meta["instr_id"][i] if "instr_id" in meta else f"{env_name}_{len(self.results)}",
                        "path_corners": [(c_host[0, i], d_host[0, i])], "actions": [], "progress": []}
                if "num_dia" in meta:
                    traj["num_dia"] = int(meta["num_dia"][i])
                if gts is not None:
                    traj["gt_path_corners"] = gts[i]
                    traj["gt_progress"] = []
                alive = True
                for t in range(steps):
                    if alive:                                              # logged while the episode has not ended
                        traj["actions"].append(out_host[t, i, :3].copy())
                        traj["progress"].append(float(out_host[t, i, 3]))
                        if prog is not None:
                            traj["gt_progress"].append(float(prog[t, i]))
                    alive = not ended[t, i]
                    if alive:
                        traj["path_corners"].append((c_host[t + 1, i], d_host[t + 1, i]))
                self.results[traj["instr_id"]] = traj
        return self.results

    def teacher_action(self, corners, gt_path_corners, ended, feedback=None):
        """``teacher_action`` (agent.py:386-507) for the whole batch on the device; ``feedback`` 'student' /
        'teacher' (default ``self.feedback`` or 'student').
        ``corners`` [B,4,2] f64 (lat,lng); ``gt_path_corners``: list (len B) of ``[n_i,4,2]`` arrays or a padded
        ``[B,Pmax,4,2]`` tensor with ``gt_len``; ``ended`` [B].  Returns device tensors
        ``(next_pos_ratio [B,2] f32, altitude [B] f32, progress [B] f32)`` -- the targets of ``output[:,0:2]``,
        ``output[:,2]`` and ``output[:,3]`` (agent.py:655-671)."""
        dev = self.device
        c = torch.as_tensor(np.asarray(corners) if not torch.is_tensor(corners) else corners).to(dev, torch.float64).contiguous()
        B = c.shape[0]
        packed = isinstance(gt_path_corners, tuple) and len(gt_path_corners) == 2 and torch.is_tensor(gt_path_corners[0])
        if isinstance(gt_path_corners, (list, tuple)) and not packed:
            lens = [len(g) for g in gt_path_corners]
            pmax = max(lens)
            gt = np.zeros((B, pmax, 4, 2), dtype=np.float64)
            for i, g in enumerate(gt_path_corners):
                gt[i, :lens[i]] = np.asarray(g, dtype=np.float64)
            gt_t = torch.from_numpy(gt).to(dev)
            len_t = torch.tensor(lens, dtype=torch.int32, device=dev)
        else:
            gt_t, len_t = gt_path_corners
            gt_t = gt_t.to(dev, torch.float64).contiguous()
            len_t = len_t.to(dev, torch.int32).contiguous()
            pmax = gt_t.shape[1]
        e = torch.as_tensor(np.asarray(ended, dtype=np.uint8) if not torch.is_tensor(ended) else ended).to(dev, torch.uint8).contiguous()
        ratio = torch.empty((B, 2), dtype=torch.float32, device=dev)
        alt = torch.empty(B, dtype=torch.float32, device=dev)
        prog = torch.empty(B, dtype=torch.float32, device=dev)
        fb = feedback if feedback is not None else getattr(self, "feedback", "student")
        if fb not in ("student", "teacher"):
            raise ValueError("Invalid feedback option")
        _lib.call("avdn_teacher_action", _lib.ptr(c), _lib.ptr(gt_t), int(pmax), _lib.ptr(len_t), _lib.ptr(e), B,
                  int(fb == "teacher"), _lib.ptr(ratio), _lib.ptr(alt), _lib.ptr(prog))
        self.launches += 1
        return ratio, alt, prog

    # ------------------------------------------------------------ API helpers
    def NSS(self, sal, fix):
        """src/xview_et/agent.py:256-270 on device tensors (torch elementwise; off the hot
        path -- the training step uses the fused ``avdn_loss`` kernel instead)."""
        nss_r = int(getattr(self.args, "nss_r", 0))
        m = torch.mean(sal.view(-1, 224 * 224), 1).view(-1, 1, 1)
        std = torch.std(sal.view(-1, 224 * 224), 1).view(-1, 1, 1)
        n_sal = (sal - m) / std
        if nss_r == 1:
            n_sal = n_sal / 2 + 1
        elif nss_r == -1:
            n_sal = n_sal / 2 - 1
        s_fix = torch.sum(fix.view(-1, 224 * 224), 1) + 0.001
        s_ns = torch.sum((n_sal * fix).view(-1, 224 * 224), 1)
        return -torch.mean(s_ns / s_fix)

    def postprocess_waypoints(self, output, edge_len, stop_threshold=0.5):
        """agent.py:637-653,738,745-752 for the whole batch in one kernel.
        ``output`` [B,4] f32, ``edge_len`` [B] f64 -> dict of device tensors."""
        B = output.shape[0]
        dev = output.device
        res = dict(angle=torch.empty(B, dtype=torch.int32, device=dev),
                   dist=torch.empty(B, dtype=torch.float64, device=dev),
                   altitude=torch.empty(B, dtype=torch.int32, device=dev),
                   stop=torch.empty(B, dtype=torch.uint8, device=dev),
                   xy=torch.empty((B, 2), dtype=torch.float32, device=dev))
        el = torch.as_tensor(edge_len, dtype=torch.float64, device=dev).contiguous()
        out_c = output.contiguous()                  # named: a temporary inside ptr() would dangle
        _lib.call("avdn_postprocess_waypoints", _lib.ptr(out_c), _lib.ptr(el), B, float(stop_threshold),
                  _lib.ptr(res["angle"]), _lib.ptr(res["dist"]), _lib.ptr(res["altitude"]), _lib.ptr(res["stop"]),
                  _lib.ptr(res["xy"]))
        return res

    def gps_to_img_coords(self, gps, ob):
        """agent.py:508-509 (twin of env.py:189-196)."""
        return (int(round((gps[1] - ob["gps_botm_left"][1]) / ob["lat_ratio"])),
                int(round((ob["gps_top_right"][0] - gps[0]) / ob["lat_ratio"])))

    def move_view_corners(self, corners, angle, distance, altitude, gps_botm_left, gps_top_right,
                          input_current_direction=None):
        """agent.py:285-384: zoom to ``altitude``, rotate by ``-angle`` about the centre, move
        forward by ``distance``; each stage is rejected if a corner leaves the map.  Host
        float64, evaluated in the reference's operation order."""
        corners = np.asarray(corners, dtype=np.float64)
        norm = np.linalg.norm

        def inside(p):
            return gps_botm_left[0] < p[0] < gps_top_right[0] and gps_botm_left[1] < p[1] < gps_top_right[1]

        def rot(theta, p):
            t = theta / 180 * PI_REF
            M = np.array([[np.cos(t), np.sin(t)], [-np.sin(t), np.cos(t)]])
            return np.matmul(M, np.array([p[0], p[1]]))

        def zoom(cs, ch):
            o = np.zeros((4, 2))
            o[0] = cs[0] + (cs[0] - cs[1]) / norm(cs[1] - cs[0]) * ch
            o[0] += (cs[0] - cs[3]) / norm(cs[3] - cs[0]) * ch
            o[1] = cs[1] + (cs[1] - cs[0]) / norm(cs[1] - cs[0]) * ch
            o[1] += (cs[1] - cs[2]) / norm(cs[2] - cs[1]) * ch
            o[2] = cs[2] + (cs[2] - cs[3]) / norm(cs[2] - cs[3]) * ch
            o[2] += (cs[2] - cs[1]) / norm(cs[2] - cs[1]) * ch
            o[3] = cs[3] + (cs[3] - cs[2]) / norm(cs[2] - cs[3]) * ch
            o[3] += (cs[3] - cs[0]) / norm(cs[3] - cs[0]) * ch
            return o

        def forward(cs, ch):
            o = np.zeros((4, 2))
            o[0] = cs[0] + (cs[0] - cs[3]) / norm(cs[3] - cs[0]) * ch
            o[1] = cs[1] + (cs[1] - cs[2]) / norm(cs[2] - cs[1]) * ch
            o[2] = cs[2] + (cs[1] - cs[2]) / norm(cs[2] - cs[1]) * ch
            o[3] = cs[3] + (cs[0] - cs[3]) / norm(cs[3] - cs[0]) * ch
            return o

        cur = round(get_direction(np.mean(corners, axis=0), (corners[0] + corners[1]) / 2)) % 360
        if input_current_direction is not None and abs(input_current_direction - cur) > 2:
            angle += input_current_direction
        edge = norm(corners[1] - corners[0]) * 11.13 * 1e4
        zoomed = zoom(corners, 0.5 * (altitude - edge) / 11.13 / 1e4)
        if not all(inside(p) for p in zoomed):
            return np.array(corners), cur
        corners = zoomed
        centre = np.mean(corners, axis=0)
        rotated = [centre + rot(-angle, corners[i] - centre) for i in range(4)]
        if not all(inside(p) for p in rotated):
            return np.array(corners), cur
        moved = forward(np.array(rotated), distance)
        if not all(inside(p) for p in moved):
            return np.array(rotated), (cur + angle) % 360
        return np.array(moved), (cur + angle) % 360

    # ------------------------------------------------- the reference's agent loop (drop-in signatures)
    def _language_batch(self, items):
        """agent.py:519-538: the dialog of every episode through the tokenizer -- ``instructions`` for the
        transformer's language rows, ``pre_dialogs + instructions`` for ``linear_cls`` unless
        ``args.train_val_on_full``.  Needs ``self.tokenizer`` (the reference's ``BertTokenizerFast``: any callable
        ``tok(list[str], padding=True, return_tensors='pt')`` -> ``input_ids`` / ``attention_mask``) and an attached
        language model."""
        if self.lang_model is None or self.tokenizer is None:
            raise RuntimeError("rollout() tokenises the dialogs and runs CustomBERTModel (agent.py:124-126,519-538): "
                               "set agent.tokenizer and call attach_lang_model() first")
        vis_only = bool(getattr(self.args, "vision_only", False))
        texts = ["" if vis_only else it["instructions"] for it in items]
        enc = self.tokenizer(texts, padding=True, return_tensors="pt")
        out = {"input_ids": enc["input_ids"].to(self.device), "attention_mask": enc["attention_mask"].to(self.device)}
        lang_inputs = texts
        if not bool(getattr(self.args, "train_val_on_full", False)):
            lang_inputs = [it["pre_dialogs"] + it["instructions"] for it in items]
            enc2 = self.tokenizer(lang_inputs, padding=True, return_tensors="pt")
            out["cls_input_ids"] = enc2["input_ids"].to(self.device)
            out["cls_attention_mask"] = enc2["attention_mask"].to(self.device)
        return out, lang_inputs

    @torch.no_grad()
    def _teacher_rollout_poses(self, corners0, dirs0, geo, gt_paths, T):
        """The pose sequence of a teacher-feedback rollout (agent.py:580-770 with ``feedback == 'teacher'``): the
        action of every step is the ground-truth action, so the sequence does not depend on the model.  Per step:
        ``teacher_action`` (target waypoint / altitude / progress at the current view) and the simulator update with
        that target -- stop once the ground-truth progress exceeds 0.5 or at the last step.  All on the device; the
        host reads the stop flags once at the end.  Returns a dict of ``[T, B, ...]`` tensors and the step count."""
        ptr, call = _lib.ptr, _lib.call
        dev = self.device
        B = corners0.shape[0]
        f64, i32 = torch.float64, torch.int32
        corners = corners0.to(dev, f64).contiguous().clone()
        cur = dirs0.to(dev, f64).contiguous().clone()
        ended = torch.zeros(B, dtype=torch.uint8, device=dev)
        bounds = geo[:, :4].contiguous()
        H = dict(corners=torch.empty((T + 1, B, 4, 2), dtype=f64, device=dev), dirs=torch.empty((T + 1, B), dtype=f64, device=dev),
                 ended=torch.empty((T, B), dtype=torch.uint8, device=dev), xy=torch.empty((T, B, 2), device=dev),
                 alt=torch.empty((T, B), device=dev), prog=torch.empty((T, B), device=dev))
        ang = torch.empty(B, dtype=i32, device=dev); dist = torch.empty(B, dtype=f64, device=dev); alt_i = torch.empty(B, dtype=i32, device=dev)
        gts = [gt_paths[i] for i in range(B)]
        lens = [len(g) for g in gts]
        gt = np.zeros((B, max(lens), 4, 2), dtype=np.float64)
        for i, g in enumerate(gts):
            gt[i, :lens[i]] = np.asarray(g, dtype=np.float64)
        gt_pack = (torch.from_numpy(gt).to(dev), torch.tensor(lens, dtype=i32, device=dev))
        for t in range(T):
            H["corners"][t].copy_(corners); H["dirs"][t].copy_(cur)
            xy, alt, prog = self.teacher_action(corners, gt_pack, ended, feedback="teacher")
            H["xy"][t].copy_(xy); H["alt"][t].copy_(alt); H["prog"][t].copy_(prog)
            out = torch.cat((xy, alt[:, None], prog[:, None]), 1).contiguous()
            call("avdn_waypoint_step", ptr(out), ptr(corners), ptr(bounds), ptr(cur), ptr(ended), B,
                 float(self.STOP_THRESHOLD), int(t == T - 1), ptr(ang), ptr(dist), ptr(alt_i))
            H["ended"][t].copy_(ended)
            self.launches += 8
        H["corners"][T].copy_(corners); H["dirs"][T].copy_(cur)
        e = H["ended"].cpu().numpy().astype(bool)
        steps = T
        for t in range(T):
            if e[t].all():                                   # agent.py:773-774: early exit once every sample has ended
                steps = t + 1
                break
        return H, steps

    def rollout(self, train_ml=None, not_in_train=False, nss_w=0, **kwargs):
        """Drop-in for ``NavCMTAgent.rollout`` (src/xview_et/agent.py:512-894) over ``self.env`` (an
        ``ANDHNavBatch`` whose ``batch`` / maps the data loader has filled): one rollout of the current batch under
        ``self.feedback``; returns the list of trajectory dicts the reference returns.

        * training (``not_in_train`` false): the rollout's loss ``ml_loss * train_ml / B`` is added to ``self.loss`` and
          its gradients to the optimiser arenas -- the reference builds an autograd graph here and calls
          ``self.loss.backward()`` in ``train``; this implementation has no graph, so the backward of a rollout runs
          inside it and ``train`` only clips and steps.  Teacher feedback: the poses come from the ground-truth
          actions; student feedback: from a no-grad greedy rollout (``student_batch``).  Language: tokenizer +
          ``CustomBERTModel``, trained in the loop.
        * ``not_in_train``: no gradients; student feedback is ``rollout_greedy``; teacher feedback additionally logs
          ``human_att_performance`` / ``nss`` per step (agent.py:683-693)."""
        env = self.env
        items = list(env.batch)
        B = len(items)
        dev = self.device
        env._sync_maps()
        T = int(getattr(self.args, "max_action_len", 15))
        lb, lang_inputs = self._language_batch(items)
        gps0 = np.stack([np.asarray(it["gt_path_corners"][0], dtype=np.float64) for it in items])
        _, geo_h, tidx_h = env._gather_poses(None, 0)
        geo = torch.from_numpy(geo_h).to(dev)
        tidx = torch.from_numpy(tidx_h).to(dev)
        dirs0 = torch.tensor([float(it["angle"]) for it in items], dtype=torch.float64, device=dev)
        gt_paths = [np.asarray(it["gt_path_corners"], dtype=np.float64) for it in items]
        has_gt = "test" not in self.env_name
        traj = []
        for i, it in enumerate(items):
            rounds = lang_inputs[i].split("[QUE]")
            remove = sum(1 for r in rounds if "Yes" in r[0:5])
            traj.append({"instr_id": it["map_name"] + "__" + it["route_index"], "num_dia": len(rounds) - remove,
                         "path_corners": [(np.array(it["gt_path_corners"][0]), it["angle"])],
                         "gt_path_corners": it["gt_path_corners"], "actions": [], "gt_actions": [], "gt_progress": [],
                         "progress": []})
        prev_r, self.renderer = self.renderer, env.renderer
        try:
            if not_in_train and self.feedback == "student":
                lm = self.lang_model
                lm.eval()
                with torch.no_grad():
                    leng = lm.engine(lb["input_ids"].shape[0], lb["input_ids"].shape[1], dev)
                    leng.set_dropout(*lm.dropout_config())
                    seq, lin, _ = leng.forward(lb["input_ids"].long(), lb["attention_mask"])
                    if "cls_input_ids" in lb:
                        leng2 = lm.engine(lb["cls_input_ids"].shape[0], lb["cls_input_ids"].shape[1], dev, slot=1)
                        leng2.set_dropout(*lm.dropout_config())
                        _, lin, _ = leng2.forward(lb["cls_input_ids"].long(), lb["cls_attention_mask"])
                res = self.rollout_greedy(dict(corners_gps=torch.from_numpy(gps0).to(dev), directions=dirs0, geo=geo,
                                               tile_idx=tidx, lang=seq.float(), lang_cls=lin.float()), max_action_len=T)
                steps = int(res["steps"])
                ended = res["ended"][:steps].cpu().numpy().astype(bool)
                c_h, d_h = res["corners"].cpu().numpy(), res["directions"].cpu().numpy()
                outs = res["output"][:steps].cpu().numpy()
                tgt = None
                if has_gt:
                    flat = res["corners"][:steps].reshape(steps * B, 4, 2)
                    eb = np.zeros((steps, B), dtype=np.uint8)
                    eb[1:] = ended[:steps - 1]
                    xy, alt, pr = self.teacher_action(flat, [gt_paths[k % B] for k in range(steps * B)], eb.reshape(-1),
                                                      feedback="student")
                    tgt = (xy.view(steps, B, 2).cpu().numpy(), alt.view(steps, B).cpu().numpy(),
                           pr.view(steps, B).cpu().numpy())
                self._log_traj(traj, outs, tgt, ended, c_h, d_h, steps)
                return traj
            # ---- poses of the rollout ----
            if self.feedback == "teacher":
                Hh, steps = self._teacher_rollout_poses(torch.from_numpy(gps0), dirs0, geo, gt_paths, T)
                corners = Hh["corners"][:steps]
                ended = Hh["ended"][:steps]
                tgt_xy, tgt_alt, tgt_prog = Hh["xy"][:steps], Hh["alt"][:steps], Hh["prog"][:steps]
                heads = Hh["dirs"][:steps]
                c_all, d_all = Hh["corners"], Hh["dirs"]
            elif self.feedback == "student":
                # the student's own (no-grad, eval-mode) greedy rollout fixes the poses; the teacher supplies the
                # target of every visited pose (agent.py:655-661)
                with torch.no_grad():
                    lm = self.lang_model
                    lm.eval()
                    leng = lm.engine(lb["input_ids"].shape[0], lb["input_ids"].shape[1], dev)
                    leng.set_dropout(*lm.dropout_config())
                    seq, lin, _ = leng.forward(lb["input_ids"].long(), lb["attention_mask"])
                    if "cls_input_ids" in lb:
                        leng2 = lm.engine(lb["cls_input_ids"].shape[0], lb["cls_input_ids"].shape[1], dev, slot=1)
                        leng2.set_dropout(*lm.dropout_config())
                        _, lin, _ = leng2.forward(lb["cls_input_ids"].long(), lb["cls_attention_mask"])
                    res = self.rollout_greedy(dict(corners_gps=torch.from_numpy(gps0).to(dev), directions=dirs0, geo=geo,
                                                   tile_idx=tidx, lang=seq.float().clone(), lang_cls=lin.float().clone()),
                                              max_action_len=T)
                steps = int(res["steps"])
                corners, ended, heads = res["corners"][:steps], res["ended"][:steps], res["directions"][:steps]
                c_all, d_all = res["corners"], res["directions"]
                eb = torch.zeros((steps, B), dtype=torch.uint8, device=dev)
                eb[1:] = ended[:steps - 1]
                xy, alt, pr = self.teacher_action(corners.reshape(steps * B, 4, 2),
                                                  [gt_paths[k % B] for k in range(steps * B)], eb.reshape(-1),
                                                  feedback="student")
                tgt_xy, tgt_alt, tgt_prog = xy.view(steps, B, 2), alt.view(steps, B), pr.view(steps, B)
            else:
                raise SystemExit("Invalid feedback option")
            ended_h = ended.cpu().numpy().astype(bool)
            alive = np.ones((steps, B), dtype=bool)
            alive[1:] = ~ended_h[:steps - 1]
            lens = np.maximum(np.cumsum(alive, axis=0).T, 1)             # [B, steps]: history length seen at step t
            px = torch.empty((steps * B, 4, 2), dtype=torch.int32, device=dev)
            _lib.call("avdn_gps_to_pixels", _lib.ptr(corners.reshape(steps * B, 4, 2).contiguous()),
                      _lib.ptr(geo.repeat(steps, 1).contiguous()), steps * B, _lib.ptr(px))
            rad = heads.float() / 180 * PI_REF                            # agent.py:605-606
            tb = lambda a: a.reshape(steps, B, *a.shape[2:]).transpose(0, 1).contiguous()
            batch = dict(lb, corners_px=tb(px.view(steps, B, 4, 2)), tile_idx=tidx.view(B, 1).expand(B, steps).contiguous(),
                         directions=tb(torch.stack([torch.sin(rad), torch.cos(rad)], -1)),
                         gt_xy=tb(tgt_xy.float()), gt_alt=tb(tgt_alt.float()), gt_prog=tb(tgt_prog.float()),
                         lenths=lens.tolist())
            if getattr(self.args, "no_direction", False):
                batch["directions"] = torch.zeros_like(batch["directions"])
            outs = []
            if not_in_train:
                loss = self._eval_rollout(batch, traj, outs)
            else:
                loss = self.train_rollout_step(batch, zero=False, step=False, nss_w=float(nss_w),
                                               train_ml=train_ml if train_ml is not None else 0.0, collect=outs,
                                               bn_per_step=bool(getattr(self.args, "bn_per_step", True)))
            tgt = (tgt_xy.cpu().numpy(), tgt_alt.cpu().numpy(), tgt_prog.cpu().numpy()) if has_gt else None
            # what the rollout was made of (inspection / tests): poses, stop flags, targets, the training batch
            self._last_rollout = dict(corners=corners, ended=ended, steps=steps, tgt_xy=tgt_xy, tgt_alt=tgt_alt,
                                      tgt_prog=tgt_prog, batch=batch, lenths=lens)
            self._log_traj(traj, torch.stack(outs).cpu().numpy(), tgt, ended_h, c_all.cpu().numpy(), d_all.cpu().numpy(), steps)
            if train_ml is not None:
                l = loss.clone()
                self.loss = l if isinstance(self.loss, (int, float)) else self.loss + l
                self.losses.append(l)
            else:
                self.losses.append(0.0)
        finally:
            self.renderer = prev_r
        return traj

    @staticmethod
    def _log_traj(traj, outs, tgt, ended, corners, dirs, steps):
        """agent.py:637-653,725-770: per alive step the clipped prediction, the ground-truth action / progress and the
        pose the simulator moved to."""
        B = len(traj)
        for i in range(B):
            alive = True
            for t in range(steps):
                if alive:
                    o = outs[t, i]
                    m = max(abs(float(o[0])), abs(float(o[1])), 1.0)
                    traj[i]["actions"].append([np.array([o[0] / m, o[1] / m], dtype=np.float32),
                                               np.float32(min(1.0, max(0.0, float(o[2]))))])
                    traj[i]["progress"].append(float(o[3]))
                    if tgt is not None:
                        traj[i]["gt_actions"].append([tgt[0][t, i].copy(), float(tgt[1][t, i])])
                        traj[i]["gt_progress"].append(float(tgt[2][t, i]))
                alive = not ended[t, i]
                if alive:
                    traj[i]["path_corners"].append((corners[t + 1, i].copy(), float(dirs[t + 1, i])))

    @torch.no_grad()
    def _eval_rollout(self, batch, traj, outs):
        """``not_in_train`` with teacher feedback (validation of the human-attention head, agent.py:673-693): the ET in
        eval mode over the growing history of the given poses; per step the loss terms and, for samples with a
        fixation map, precision / recall of the clipped saliency map and the NSS value.  Returns the loss."""
        ptr = _lib.ptr
        dev = self.device
        self.vision_model.eval()
        et = self.vln_model
        et.eval()
        lm = self.lang_model
        lm.eval()
        ids, am = batch["input_ids"], batch["attention_mask"]
        leng = lm.engine(ids.shape[0], ids.shape[1], dev)
        leng.set_dropout(*lm.dropout_config())
        seq, lin, _ = leng.forward(ids.long(), am)
        if "cls_input_ids" in batch:
            leng2 = lm.engine(batch["cls_input_ids"].shape[0], batch["cls_input_ids"].shape[1], dev, slot=1)
            leng2.set_dropout(*lm.dropout_config())
            _, lin, _ = leng2.forward(batch["cls_input_ids"].long(), batch["cls_attention_mask"])
        lang, lang_cls = seq.float().contiguous(), lin.float().contiguous()
        dirs = batch["directions"].contiguous().float()
        B, T = dirs.shape[:2]
        L = lang.shape[1]
        r = self.renderer
        ti = batch["tile_idx"].reshape(-1)
        minv = r.homography(batch["corners_px"].view(B * T, 4, 2))
        x = torch.empty((B * T, 224, 224, 4), dtype=torch.bfloat16, device=dev)
        att = torch.empty((B * T, 224, 224), dtype=torch.uint8, device=dev)
        r.render(None, ti, views=False, norm_nhwc=True, minv=minv, out={"norm_nhwc": x})
        r.render(None, ti, views=False, att=True, minv=minv, out={"att": att})
        att = att.view(B, T, 224, 224)
        xs = x.view(B, T, 224, 224, 4)
        fr = torch.empty((B, T, 512, 49), dtype=torch.float32, device=dev)
        for t in range(T):                                   # the reference's eval-mode trunk sees B views per call
            teng = self.vision_model.engine(B, 224, 224, dev)
            o = torch.empty((B, 512, 7, 7), dtype=torch.float32, device=dev)
            DN._trunk_forward(self.vision_model, teng, xs[:, t].contiguous(), False, out=o)
            fr[:, t] = o.view(B, 512, 49)
        lens = batch["lenths"]
        pe = et.encoder_vl.enc_pos.pe[0]
        total = torch.zeros(1, dtype=torch.float64, device=dev)
        loss_i = torch.empty(B, dtype=torch.float64, device=dev)
        d_o = torch.empty((B, 4), device=dev); d_h = torch.empty((B, 64), device=dev)
        scale = float(self.train_ml) / B
        for t in range(T):
            Tc = t + 1
            eng = et.engine(B, L, Tc, dev)
            eng.set_dropout(0.0, 0.0, 0)
            output, h_sali = eng.forward(fr[:, :Tc].contiguous().view(B * Tc, 512, 49), lang, lang_cls,
                                         dirs[:, :Tc].contiguous(), [int(lens[i][t]) for i in range(B)], pe)
            outs.append(output.detach().clone())
            att_t = att[:, t].contiguous()
            _lib.call("avdn_loss", ptr(output), ptr(h_sali), ptr(batch["gt_xy"][:, t].contiguous()),
                      ptr(batch["gt_alt"][:, t].contiguous()), ptr(batch["gt_prog"][:, t].contiguous()), ptr(att_t), None, B,
                      float(self.nss_w), int(getattr(self.args, "nss_r", 0)), scale, ptr(total), ptr(loss_i), ptr(d_o), ptr(d_h))
            if self.feedback == "teacher":
                sal = torch.empty((B, 224, 224), dtype=torch.float32, device=dev)
                _lib.call("avdn_upsample_saliency", ptr(h_sali.contiguous()), B, ptr(sal))
                gt = att_t.double() / 255
                for i in range(B):
                    if float(gt[i].sum()) > 0:
                        nss = self.NSS(sal[i].double(), gt[i])
                        p = sal[i].clip(0, 1).double()
                        tp = float((p * gt[i]).sum())
                        sp = float(p.sum())
                        traj[i].setdefault("human_att_performance", []).append([tp / sp if sp != 0 else 0.0,
                                                                                tp / float(gt[i].sum())])
                        traj[i].setdefault("nss", []).append(float(nss))
        return total

    def train(self, loader, n_epochs, feedback="student", nss_w_weighting=1, **kwargs):
        """Drop-in for ``NavCMTAgent.train`` (src/xview_et/agent.py:208-254): for every batch the loader yields (the
        loader fills ``self.env`` as a side effect, as the reference's ``num_workers=0`` DataLoader does): zero the
        gradients, one teacher rollout (``feedback='teacher'``) or a teacher rollout without the NSS term followed by
        a student rollout (``'student'``), clip the ET gradients to 40, one step of every optimiser."""
        self.vision_model.train(); self.vln_model.train()
        if self.lang_model is not None:
            self.lang_model.train()
        self.losses = []
        for epoch in range(1, n_epochs + 1):
            for _ in loader:
                for opt in self.optimizers:
                    opt.zero_grad()
                self.loss = 0
                if feedback == "teacher":
                    self.feedback = "teacher"
                    self.rollout(train_ml=float(getattr(self.args, "teacher_weight", 1.0)), train_rl=False,
                                 nss_w=float(getattr(self.args, "nss_w", 1.0)) * nss_w_weighting, **kwargs)
                elif feedback == "student":
                    self.feedback = "teacher"
                    self.rollout(train_ml=float(getattr(self.args, "ml_weight", 0.2)), train_rl=False, nss_w=0, **kwargs)
                    self.feedback = "student"
                    self.rollout(train_ml=float(getattr(self.args, "ml_weight", 0.2)), train_rl=False,
                                 nss_w=float(getattr(self.args, "nss_w", 1.0)) * nss_w_weighting, **kwargs)
                else:
                    assert False
                # (the reference calls self.loss.backward() here; the rollouts above already left their gradients in
                #  the arenas)  gradient all-reduce, clip (ET only, agent.py:247), step
                self._reduce_and_step()

    def _reduce_and_step(self):
        if self.world > 1:
            for opt in self.optimizers:
                self._allreduce_async(opt.g, 0, opt.n)
            self._wait_comm()
        gs = 1.0 / self.world
        for opt in self.optimizers:
            self.launches += opt.step(grad_scale=gs)

    # ------------------------------------------------------------- checkpoints
    def _checkpoint_items(self):
        items = [("vision_model", self.vision_model, self.vision_model_optimizer),
                 ("vln_model", self.vln_model, self.et_optimizer)]
        if self.lang_model is not None:
            items.insert(0, ("lang_model", self.lang_model, self.lang_optimizer))
        return items

    def save(self, epoch, path):
        """agent.py:899-916: vision_model, vln_model and (when attached) lang_model."""
        d = os.path.dirname(path)
        if d:
            os.makedirs(d, exist_ok=True)
        states = {}
        for name, model, opt in self._checkpoint_items():
            states[name] = {"epoch": epoch + 1, "state_dict": model.state_dict(),
                            "optimizer": {"m": opt.m.clone(), "v": opt.v.clone(), "step": opt.step_count}}
        torch.save(states, path)

    def load(self, path):
        """agent.py:918-940: parameters (and optimiser moments when ``args.resume_optimizer``)."""
        states = torch.load(path, map_location=self.device)
        for name, model, opt in self._checkpoint_items():
            if name not in states:
                continue
            state = model.state_dict()                       # tensors alias the arena: copy IN PLACE
            with torch.no_grad():
                for k, v in states[name]["state_dict"].items():
                    if k in state:
                        state[k].copy_(v)
            if getattr(self.args, "resume_optimizer", False):
                o = states[name].get("optimizer", {})
                if "m" not in o:
                    raise ValueError(f"checkpoint '{name}': optimiser state is not in the flat-arena format this agent "
                                     "writes ({'m', 'v', 'step'}); torch per-parameter optimiser states (the reference's "
                                     "format) cannot be resumed -- load with resume_optimizer=False")
                opt.m.copy_(o["m"]); opt.v.copy_(o["v"]); opt.step_count = int(o["step"])
        return states.get("vln_model", {}).get("epoch", 0) - 1
