"""``NavCMTAgent`` of the LSTM baseline -- greedy waypoint-rollout inference
(mirror of src/xview_lstm/agent.py, hot-path subset = BASELINE config 5).

``rollout_greedy`` is the student-feedback loop of ``rollout`` (src/xview_lstm/agent.py:518-750)
with every per-step stage on the device and no host round trip inside the loop:

    corners (GPS) -> pixel corners (env.py:189-196) -> homography + rendered views (env.py:287-293,
    agent.py:575-582) -> Darknet trunk, eval mode (vln_model.py:216) -> ViT_LSTM step
    (vln_model.py:219-248) -> post-processing + stop test + move_view_corners
    (agent.py:607-626,700-730; xview_et/agent.py:285-384)

The language side (BERT -> ``lang_feature`` [B,L,768], ``cls_hidden`` = linear_cls [B,49],
agent.py:527-543) is an input of the path; the teacher / loss branch of the
reference loop needs ground truth and shapely and is outside config 5.
"""
from __future__ import annotations

import os

import torch

from .. import _lib
from ..env import ViewRenderer
from ..models import dark_net as DN
from ..models.dark_net import Darknet
from ..models.vln_model import ViT_LSTM, NCH, NSP

STOP_THRESHOLD_STUDENT = 0.25          # src/xview_lstm/agent.py:705 (the ET agent uses 0.5)


class NavCMTAgent:
    def __init__(self, args, rank=0, device=None):
        if not torch.cuda.is_available():
            raise RuntimeError("NavCMTAgent needs a CUDA device (sm_100); there is no CPU fallback")
        _lib.lib()
        self.args, self.rank = args, rank
        self.device = torch.device(device if device is not None else "cuda")
        self.results, self.losses, self.env = {}, [], []
        self.vision_model = Darknet(args.darknet_model_file, 224)
        wf = getattr(args, "darknet_weight_file", None)
        if wf and os.path.exists(wf):                           # agent.py:130-135
            new_state = torch.load(wf, map_location="cpu")
            state = self.vision_model.state_dict()
            state.update({k: v for k, v in new_state["model"].items() if k in state})
            self.vision_model.load_state_dict(state)
        self.vln_model = ViT_LSTM(args, self.vision_model).to(self.device)      # trunk is a sub-module (agent.py:140-142)
        self.renderer = ViewRenderer(self.device)
        self.use_graphs = os.environ.get("AVDN_CUDA_GRAPHS", "1") != "0"     # rollouts replay the frozen trunk pass
        self.launches = 0
        self._bufs = {}

    def _get_bufs(self, B, T):
        b = self._bufs.get((B, T))
        if b is None:
            dev = self.device
            b = dict(x=torch.empty((B, 224, 224, 4), dtype=torch.bfloat16, device=dev),
                     frames=torch.empty((B, NCH, 7, 7), dtype=torch.float32, device=dev),
                     corners=torch.empty((B, 4, 2), dtype=torch.float64, device=dev),
                     cur_dir=torch.empty(B, dtype=torch.float64, device=dev),
                     ended=torch.empty(B, dtype=torch.uint8, device=dev),
                     px=torch.empty((B, 4, 2), dtype=torch.int32, device=dev),
                     minv=torch.empty((B, 3, 3), dtype=torch.float64, device=dev),
                     corners_hist=torch.empty((T + 1, B, 4, 2), dtype=torch.float64, device=dev),
                     dir_hist=torch.empty((T + 1, B), dtype=torch.float64, device=dev),
                     ended_hist=torch.empty((T, B), dtype=torch.uint8, device=dev),
                     output_hist=torch.empty((T, B, 4), dtype=torch.float32, device=dev),
                     angle=torch.empty((T, B), dtype=torch.int32, device=dev),
                     altitude=torch.empty((T, B), dtype=torch.int32, device=dev),
                     dist=torch.empty((T, B), dtype=torch.float64, device=dev))
            self._bufs[(B, T)] = b
        return b

    @torch.no_grad()
    def rollout_greedy(self, batch, max_action_len=None, stop_threshold=STOP_THRESHOLD_STUDENT):
        """Greedy (student) rollout of ``B`` episodes for ``max_action_len`` steps.

        ``batch`` (device tensors): ``corners_gps`` f64 [B,4,2] (lat,lng; FL,FR,BR,BL = gt_path_corners[0]),
        ``directions`` [B] (starting_angle, degrees), ``geo`` f64 [B,5] = (bl_lat, bl_lng, tr_lat, tr_lng,
        lat_ratio), ``tile_idx`` i32 [B] or None, ``lang_feature`` f32 [B,L,768], ``cls_hidden`` f32 [B,49].

        Returns device tensors: ``corners`` [T+1,B,4,2] (pose before every step and after the last),
        ``directions`` [T+1,B], ``ended`` [T,B] (sticky stop flags after each step), ``output`` [T,B,4]
        (raw network outputs), ``angle`` / ``altitude`` [T,B] (discretised action), ``dist`` [T,B].
        A trajectory is the poses of the steps at which the sample had not ended yet
        (``traj['path_corners']``, agent.py:735-739)."""
        self.vln_model.eval()
        T = int(max_action_len if max_action_len is not None else getattr(self.args, "max_action_len", 20))
        lang, cls = batch["lang_feature"].contiguous(), batch["cls_hidden"].contiguous()
        B, L = lang.shape[0], lang.shape[1]
        ptr, call = _lib.ptr, _lib.call
        bf = self._get_bufs(B, T)
        geo = batch["geo"].contiguous()
        bounds = geo[:, :4].contiguous()
        ti = batch.get("tile_idx")
        bf["corners"].copy_(batch["corners_gps"])
        bf["cur_dir"].copy_(batch["directions"])
        bf["ended"].zero_()
        vm, r = self.vision_model, self.renderer
        eng = vm.engine(B, 224, 224, self.device)
        l0, v0 = eng.launches, self.vln_model.launches
        self.vln_model.reset_state(B, L, self.device)
        n = 0
        for t in range(T):
            bf["corners_hist"][t].copy_(bf["corners"])
            bf["dir_hist"][t].copy_(bf["cur_dir"])
            # ---- observation: env._get_obs(corners=current_view_corners) (agent.py:742) ----
            call("avdn_gps_to_pixels", ptr(bf["corners"]), ptr(geo), B, ptr(bf["px"]))
            call("avdn_homography_from_corners", ptr(bf["px"]), B, ptr(bf["minv"]))
            r.render(None, ti, views=False, norm_nhwc=True, minv=bf["minv"], out={"norm_nhwc": bf["x"]})
            # ---- policy (agent.py:592-602) ----
            if t == 0 or not self.use_graphs:
                DN._trunk_forward(vm, eng, bf["x"], False, out=bf["frames"], frozen=(t > 0))
            else:
                DN._trunk_forward_graphed(vm, eng, bf["x"], bf["frames"])
            output, _ = self.vln_model.step(bf["frames"].view(B, NCH, NSP), bf["cur_dir"], cls, lang,
                                            want_saliency=False)
            bf["output_hist"][t].copy_(output)
            # ---- simulator (agent.py:607-626,700-730) ----
            call("avdn_waypoint_step", ptr(output), ptr(bf["corners"]), ptr(bounds), ptr(bf["cur_dir"]),
                 ptr(bf["ended"]), B, float(stop_threshold), int(t == T - 1), ptr(bf["angle"][t]),
                 ptr(bf["dist"][t]), ptr(bf["altitude"][t]))
            bf["ended_hist"][t].copy_(bf["ended"])
            n += 4
        bf["corners_hist"][T].copy_(bf["corners"])
        bf["dir_hist"][T].copy_(bf["cur_dir"])
        self.launches += n + (eng.launches - l0) + (self.vln_model.launches - v0)
        return dict(corners=bf["corners_hist"], directions=bf["dir_hist"], ended=bf["ended_hist"],
                    output=bf["output_hist"], angle=bf["angle"], altitude=bf["altitude"], dist=bf["dist"])

    @staticmethod
    def trajectories(res):
        """Host view of a rollout result: per sample the list of (corners [4,2], direction) the
        reference appends to ``traj['path_corners']`` (agent.py:556-558,735-739)."""
        corners = res["corners"].cpu().numpy()
        dirs = res["directions"].cpu().numpy()
        ended = res["ended"].cpu().numpy().astype(bool)
        T, B = ended.shape
        out = []
        for i in range(B):
            path = [(corners[0, i], dirs[0, i])]
            for t in range(T):
                if not ended[t, i]:
                    path.append((corners[t + 1, i], dirs[t + 1, i]))
            out.append(path)
        return out
