"""Batched, GPU-resident replacement for the observation interface of the
reference environment (``/root/reference/src/env.py``).

``ViewRenderer``    — the device-side renderer: maps are packed once into HBM
                      (``avdn_pack_tile``), poses are rendered in one launch per
                      batch (``avdn_homography_from_corners`` + ``avdn_render_views``).
``ANDHNavBatch``    — host-side mirror of the reference class of the same name,
                      restricted to the hot-path method ``_get_obs`` (src/env.py:254-332)
                      and ``gps_to_img_coords`` (src/env.py:189-196).  It reads the same
                      attributes (``batch``, ``batch_size``, ``map_batch``,
                      ``attention_map_batch``) and returns the same list of dicts.

Dataset loading (``__init__`` JSON parsing, ``next_batch`` tif decoding) and the
evaluation metrics of the reference class are outside the hot path
 and are not provided here.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib

VIEW = 224

# src/xview_et/agent.py:115-116
RGB_MEAN = np.array([60.134, 49.697, 40.746], dtype=np.float32)
RGB_STD = np.array([29.99, 24.498, 22.046], dtype=np.float32)


def normalisation_lut() -> np.ndarray:
    """``lut[c][v] = (float32(v) - mean[c]) / std[c]`` for RGB channel ``c``; two
    separate float32 operations exactly as src/xview_et/agent.py:590-592."""
    v = np.arange(256, dtype=np.float32)[None, :].repeat(3, 0)
    v -= RGB_MEAN[:, None]
    v /= RGB_STD[:, None]
    return np.ascontiguousarray(v)


class ViewRenderer:
    """Holds packed maps in HBM and renders batches of drone views from them."""

    def __init__(self, device="cuda"):
        if not torch.cuda.is_available():
            raise RuntimeError("ViewRenderer needs a CUDA device (sm_100); there is no CPU fallback")
        _lib.lib()
        self.device = torch.device(device)
        self._maps = {}            # name -> dict(tile8=tensor, H, W, index)
        self._order = []           # index -> name
        self._table = None         # device array of avdn_tile_desc
        self._lut = torch.from_numpy(normalisation_lut()).to(self.device)

    # ------------------------------------------------------------------ maps
    def add_map(self, name, map_bgr, attention_map=None):
        """Upload and pack one satellite tile (``map_batch[name]``, BGR u8 HWC) and
        optionally its attention map (``attention_map_batch[name]``, u8 HWC or HW).

        The reference's attention maps are gray (R=G=B, src/env.py:224-231) and its
        ``BGR2GRAY`` is then the identity, so channel 0 is packed.  A non-gray
        attention map is rejected rather than silently mis-rendered.
        """
        m = torch.as_tensor(np.ascontiguousarray(map_bgr)) if not torch.is_tensor(map_bgr) else map_bgr
        if m.dtype != torch.uint8 or m.dim() != 3 or m.shape[2] != 3:
            raise ValueError("map must be uint8 [H,W,3] (BGR)")
        H, W = int(m.shape[0]), int(m.shape[1])
        m = m.to(self.device).contiguous()
        a, a_ch = None, 0
        if attention_map is not None:
            a = torch.as_tensor(np.ascontiguousarray(attention_map)) if not torch.is_tensor(attention_map) else attention_map
            if a.dim() == 2:
                a = a[:, :, None]
            if a.dtype != torch.uint8 or a.shape[0] != H or a.shape[1] != W:
                raise ValueError("attention map must be uint8 with the map's height and width")
            a = a.to(self.device).contiguous()
            if a.shape[2] == 3 and not (torch.equal(a[..., 0], a[..., 1]) and torch.equal(a[..., 0], a[..., 2])):
                raise ValueError("attention map must be gray (R=G=B), as built by src/env.py:224-231")
            a_ch = int(a.shape[2])
        tile8 = torch.empty((H + 1) * (W + 2), dtype=torch.int64, device=self.device)
        _lib.call("avdn_pack_tile", _lib.ptr(m), _lib.ptr(a), a_ch, H, W, _lib.ptr(tile8))
        if name in self._maps:
            idx = self._maps[name]["index"]
        else:
            idx = len(self._order)
            self._order.append(name)
        self._maps[name] = dict(tile8=tile8, H=H, W=W, index=idx, has_att=a is not None)
        self._rebuild_table()
        return idx

    @staticmethod
    def area_table(ssize, dsize):
        """OpenCV's ``computeResizeAreaTab`` for one axis (resize.cpp), float64 on the host: returns
        (ofs int32 [dsize+1], sidx int32 [n], alpha float32 [n]); entries of destination column dx are
        ``ofs[dx]:ofs[dx+1]`` in the order ``ResizeArea_`` accumulates them."""
        import math
        scale = ssize / dsize
        ofs, sidx, alpha = [0], [], []
        for dx in range(dsize):
            fsx1 = dx * scale
            fsx2 = fsx1 + scale
            cell = min(scale, ssize - fsx1)
            sx1, sx2 = math.ceil(fsx1), math.floor(fsx2)
            sx2 = min(sx2, ssize - 1)
            sx1 = min(sx1, sx2)
            if sx1 - fsx1 > 1e-3:
                sidx.append(sx1 - 1); alpha.append((sx1 - fsx1) / cell)
            for sx in range(sx1, sx2):
                sidx.append(sx); alpha.append(1.0 / cell)
            if fsx2 - sx2 > 1e-3:
                sidx.append(sx2); alpha.append(min(min(fsx2 - sx2, 1.0), cell) / cell)
            ofs.append(len(sidx))
        return (np.asarray(ofs, np.int32), np.asarray(sidx, np.int32), np.asarray(alpha, np.float64).astype(np.float32))

    def prepare_map(self, name, im_bgr, lng_ratio, lat_ratio, attention_spots_px, keep_host_copies=False):
        """Map preparation of ``next_batch`` (src/env.py:217-231) on the device, bit-exact with OpenCV:
        ``cv2.resize(im, (int(W*lng_ratio/lat_ratio), H), INTER_AREA)``, the attention raster (zeros + one filled
        white circle per ``(center (x, y) px, radius px)``) and the renderer's packed layout -- the decoded tile
        is uploaded once and never comes back.  Returns ``(index, (H, new_w, 3))``; with ``keep_host_copies`` also
        the two uint8 arrays the reference keeps in ``map_batch`` / ``attention_map_batch``."""
        m = torch.as_tensor(np.ascontiguousarray(im_bgr)) if not torch.is_tensor(im_bgr) else im_bgr
        if m.dtype != torch.uint8 or m.dim() != 3 or m.shape[2] != 3:
            raise ValueError("map must be uint8 [H,W,3] (BGR)")
        H, W = int(m.shape[0]), int(m.shape[1])
        new_w = int(W * lng_ratio / lat_ratio)                         # src/env.py:221
        if not 0 < new_w <= W:
            raise NotImplementedError("only the horizontal shrink of src/env.py:221 (lng_ratio <= lat_ratio) is implemented")
        dev = self.device
        m = m.to(dev).contiguous()
        if new_w == W:
            resized = m
        else:
            key = (W, new_w)
            tabs = getattr(self, "_area_tabs", None)
            if tabs is None:
                tabs = self._area_tabs = {}
            if key not in tabs:                                         # the table depends on the two widths only
                tabs[key] = tuple(torch.from_numpy(a).to(dev) for a in self.area_table(W, new_w))
            t_ofs, t_sidx, t_alpha = tabs[key]
            resized = torch.empty((H, new_w, 3), dtype=torch.uint8, device=dev)
            _lib.call("avdn_resize_area_width", _lib.ptr(m), H, W, new_w, _lib.ptr(t_ofs), _lib.ptr(t_sidx),
                      _lib.ptr(t_alpha), _lib.ptr(resized))
        spots = np.asarray([[int(c[0]), int(c[1]), int(r)] for (c, r) in attention_spots_px], dtype=np.int32).reshape(-1, 3)
        att = torch.empty((H, new_w, 1), dtype=torch.uint8, device=dev)
        rmax = int(spots[:, 2].max()) if len(spots) else 0
        t_spots = torch.from_numpy(spots).to(dev) if len(spots) else None
        scratch = torch.empty((max(len(spots), 1), rmax + 1), dtype=torch.int32, device=dev)
        _lib.call("avdn_raster_attention", _lib.ptr(t_spots), len(spots), rmax, _lib.ptr(scratch), H, new_w, 1,
                  _lib.ptr(att))
        idx = self.add_map(name, resized, att)
        if keep_host_copies:
            return idx, (H, new_w, 3), resized.cpu().numpy(), np.repeat(att.cpu().numpy(), 3, axis=2)
        return idx, (H, new_w, 3)

    def remove_map(self, name):
        """Mirror of the reference's eviction of unused maps (src/env.py:234-240)."""
        if name in self._maps:
            del self._maps[name]
            self._order = [n for n in self._order if n != name]
            for i, n in enumerate(self._order):
                self._maps[n]["index"] = i
            self._rebuild_table()

    def map_index(self, name):
        return self._maps[name]["index"]

    def has_map(self, name):
        return name in self._maps

    def map_shape(self, name):
        e = self._maps[name]
        return (e["H"], e["W"], 3)

    def _rebuild_table(self):
        n = len(self._order)
        if n == 0:
            self._table = None
            return
        arr = (_lib.TileDesc * n)()
        for i, name in enumerate(self._order):
            e = self._maps[name]
            arr[i].tile8 = e["tile8"].data_ptr()
            arr[i].H, arr[i].W = e["H"], e["W"]
        raw = np.frombuffer(bytes(arr), dtype=np.uint8).copy()
        self._table = torch.from_numpy(raw).to(self.device)

    # --------------------------------------------------------------- geometry
    def gps_to_pixels(self, corners_gps, geo):
        """``gps_to_img_coords`` (src/env.py:189-196) for ``[P,4,2]`` (lat,lng) float64
        corners; ``geo`` is ``[P,5]`` = (bl_lat, bl_lng, tr_lat, tr_lng, lat_ratio).
        Returns int32 ``[P,4,2]`` (x,y) on the device."""
        cg = torch.as_tensor(corners_gps, dtype=torch.float64).to(self.device).contiguous()
        g = torch.as_tensor(geo, dtype=torch.float64).to(self.device).contiguous()
        P = int(cg.shape[0])
        out = torch.empty((P, 4, 2), dtype=torch.int32, device=self.device)
        _lib.call("avdn_gps_to_pixels", _lib.ptr(cg), _lib.ptr(g), P, _lib.ptr(out))
        return out

    def homography(self, corners_px):
        """int32 ``[P,4,2]`` pixel corners -> ``[P,3,3]`` float64 inverse homographies."""
        c = torch.as_tensor(corners_px).to(device=self.device, dtype=torch.int32).contiguous()
        P = int(c.shape[0])
        minv = torch.empty((P, 3, 3), dtype=torch.float64, device=self.device)
        _lib.call("avdn_homography_from_corners", _lib.ptr(c), P, _lib.ptr(minv))
        return minv

    # ----------------------------------------------------------------- render
    def render(self, corners_px, tile_idx=None, *, views=True, att=False, norm_nchw=False,
               norm_nhwc=False, minv=None, out=None):
        """Render ``P`` poses.  Returns a dict with the requested outputs:
        ``views`` u8 ``[P,224,224,3]`` BGR, ``att`` u8 ``[P,224,224]``,
        ``norm_nchw`` f32 ``[P,3,224,224]`` RGB, ``norm_nhwc`` bf16 ``[P,224,224,4]``.
        ``out`` may hold preallocated tensors under the same keys."""
        if self._table is None:
            raise RuntimeError("ViewRenderer.render: no map has been added")
        if minv is None:
            minv = self.homography(corners_px)
        P = int(minv.shape[0])
        ti = None
        if tile_idx is not None:
            ti = torch.as_tensor(tile_idx).to(device=self.device, dtype=torch.int32).contiguous()
        out = dict(out or {})
        dev = self.device

        def get(key, want, shape, dtype):
            if not want:
                return None
            t = out.get(key)
            if t is None:
                t = torch.empty(shape, dtype=dtype, device=dev)
                out[key] = t
            return t

        v = get("views", views, (P, VIEW, VIEW, 3), torch.uint8)
        a = get("att", att, (P, VIEW, VIEW), torch.uint8)
        n1 = get("norm_nchw", norm_nchw, (P, 3, VIEW, VIEW), torch.float32)
        n2 = get("norm_nhwc", norm_nhwc, (P, VIEW, VIEW, 4), torch.bfloat16)
        _lib.call("avdn_render_views", _lib.ptr(self._table), len(self._order), _lib.ptr(ti),
                  _lib.ptr(minv), P, _lib.ptr(v), _lib.ptr(a), _lib.ptr(n1), _lib.ptr(n2),
                  _lib.ptr(self._lut))
        return out


class ANDHNavBatch:
    """Hot-path subset of the reference ``ANDHNavBatch`` (src/env.py:83-332).

    Construct it empty and fill ``batch`` / ``map_batch`` / ``attention_map_batch``
    the way the reference's ``next_batch`` does (src/env.py:203-249); ``_get_obs``
    then behaves like the reference's, with all poses of the batch rendered by
    one kernel launch instead of a per-sample cv2 loop.
    """

    def __init__(self, batch_size=4, device="cuda"):
        self.batch_size = batch_size
        self.batch = []
        self.map_batch = {}
        self.attention_map_batch = {}
        self.renderer = ViewRenderer(device)
        self._uploaded = {}        # map name -> (id(map array), id(att array))
        self._device_maps = {}     # map name -> shape of maps prepared on the device (load_map)

    def load_map(self, name, im_bgr, item):
        """The map-preparation branch of ``next_batch`` (src/env.py:217-231) for one decoded tile (``cv2.imread``
        output, BGR u8) on the device: INTER_AREA width rescale by ``lng_ratio / lat_ratio``, the attention raster
        from ``item['attention_list']`` = [((lat, lng), radius_px), ...] and the renderer's packed layout.  The
        map then serves ``_get_obs`` like an entry of ``map_batch`` (whose host copy is not needed any more)."""
        spots = [(self.gps_to_img_coords(a[0], item), a[1]) for a in item.get("attention_list", [])]
        _, shape = self.renderer.prepare_map(name, im_bgr, item["lng_ratio"], item["lat_ratio"], spots)
        self._device_maps[name] = shape
        return shape

    def drop_unused_maps(self, used_names):
        """src/env.py:234-240: forget the maps the next batch does not use."""
        for name in [n for n in self._device_maps if n not in used_names]:
            self.renderer.remove_map(name)
            del self._device_maps[name]

    def gps_to_img_coords(self, gps, ob):
        """src/env.py:189-196 (host scalar version, used by callers outside the path)."""
        lat_ratio = ob["lat_ratio"]
        return (int(round((gps[1] - ob["gps_botm_left"][1]) / lat_ratio)),
                int(round((ob["gps_top_right"][0] - gps[0]) / lat_ratio)))

    # ------------------------------------------------------------- evaluation
    @staticmethod
    def _contains(quad, point):
        """shapely ``Polygon(quad).contains(Point(point))`` for the simulator's convex quads: strictly inside."""
        q = np.asarray(quad, dtype=np.float64)
        x, y = float(point[0]), float(point[1])
        sign = 0
        for i in range(len(q)):
            a, b = q[i], q[(i + 1) % len(q)]
            cr = (b[0] - a[0]) * (y - a[1]) - (b[1] - a[1]) * (x - a[0])
            if cr == 0:
                return False
            s_ = 1 if cr > 0 else -1
            if sign == 0:
                sign = s_
            elif s_ != sign:
                return False
        return True

    def _eval_item(self, gt_path, gt_corners, path, corners, progress):
        """src/env.py:335-374: trajectory length, goal progress, (oracle) success, SPL of one trajectory."""
        scores = {}
        scores["trajectory_lengths"] = np.sum([np.linalg.norm(a - b) for a, b in zip(path[:-1], path[1:])])
        scores["trajectory_lengths"] = scores["trajectory_lengths"] * 11.13 * 1e4
        gt_whole_lengths = np.sum([np.linalg.norm(a - b) for a, b in zip(gt_path[:-1], gt_path[1:])]) * 11.13 * 1e4
        gt_net_lengths = np.linalg.norm(gt_path[0] - gt_path[-1]) * 11.13 * 1e4
        scores["iou"] = progress[-1]
        scores["gp"] = gt_net_lengths - np.linalg.norm(path[-1] - gt_path[-1]) * 11.13 * 1e4
        scores["oracle_gp"] = gt_net_lengths - np.min([np.linalg.norm(path[x] - gt_path[-1])
                                                       for x in range(len(path))]) * 11.13 * 1e4
        scores["success"] = float(progress[-1] >= 0.4)
        if not self._contains(corners[-1], np.mean(gt_corners[-1], axis=0)):
            scores["success"] = float(0)
        if not self._contains(gt_corners[-1], np.mean(corners[-1], axis=0)):
            scores["success"] = float(0)
        scores["oracle_success"] = float(any(np.array(progress) > 0.4))
        scores["gt_length"] = gt_whole_lengths
        scores["spl"] = scores["success"] * gt_net_lengths / max(scores["trajectory_lengths"], gt_net_lengths, 0.01)
        return scores

    def eval_metrics(self, preds, human_att_eval=False):
        """src/env.py:376-475: averages over the trajectories of ``preds`` (``instr_id`` -> trajectory dict as
        the rollouts produce them: ``path_corners``, ``gt_path_corners``, ``gt_progress``, ``num_dia`` ...)."""
        from collections import defaultdict
        metrics = defaultdict(list)
        if human_att_eval:
            for k in preds.keys():
                if "human_att_performance" in preds[k].keys():
                    metrics["human_att_performance"] += preds[k]["human_att_performance"]
                    nss = np.mean(preds[k]["nss"])
                    if nss == nss:
                        metrics["nss"].append(nss)
            metrics["human_att_performance"] = np.mean(metrics["human_att_performance"], axis=0)
            metrics["nss"] = np.mean(metrics["nss"])
            if metrics["nss"] == metrics["nss"]:
                # (the reference reports precision under both keys, src/env.py:396-398)
                avg = {"HA_precision": metrics["human_att_performance"][0],
                       "HA_recall": metrics["human_att_performance"][0], "nss": metrics["nss"]}
            else:
                avg = {"HA_precision": 0, "HA_recall": 0, "nss": 0}
            return avg, metrics
        for k in preds.keys():
            item = preds[k]
            dia_number = item.get("num_dia", 0)
            traj = [np.mean(x[0], axis=0) for x in item["path_corners"]]
            corners = [np.array(x[0]) for x in item["path_corners"]]
            progress = [x for x in item["gt_progress"]]
            gt_corners = [np.array(x) for x in item["gt_path_corners"]]
            gt_trajs = [np.mean(x, axis=0) for x in item["gt_path_corners"]]
            sc = self._eval_item(gt_trajs, gt_corners, traj, corners, progress)
            for kk, v in sc.items():
                metrics[kk].append(v)
            tag = {1: "1", 2: "2"}.get(dia_number, "else")
            metrics["success_" + tag].append(sc["success"])
            metrics["spl_" + tag].append(sc["spl"])
            metrics["gp_" + tag].append(sc["gp"])
            tag = "long" if sc["trajectory_lengths"] > 150 else "short"
            metrics["success_" + tag].append(sc["success"])
            metrics["spl_" + tag].append(sc["spl"])
            metrics["gp_" + tag].append(sc["gp"])
            metrics["instr_id"].append(item["instr_id"])
        avg = {"lengths": np.mean(metrics["trajectory_lengths"]), "sr": np.mean(metrics["success"]) * 100,
               "oracle_sr": np.mean(metrics["oracle_success"]) * 100, "spl": np.mean(metrics["spl"]) * 100,
               "gp": np.mean(metrics["gp"]), "oracle_gp": np.mean(metrics["oracle_gp"]),
               "gt_length": np.mean(metrics["gt_length"]), "iou": np.mean(metrics["iou"])}
        for tag in ("1", "2", "else"):
            if len(metrics["success_" + tag]) != 0:
                avg["num_" + tag] = len(metrics["success_" + tag])
                avg["spl_" + tag] = np.mean(metrics["spl_" + tag]) * 100
                avg["sr_" + tag] = np.mean(metrics["success_" + tag]) * 100
                avg["gp_" + tag] = np.mean(metrics["gp_" + tag])
        return avg, metrics

    def _sync_maps(self):
        for name in list(self._uploaded):
            if name not in self.map_batch and name not in self._device_maps:
                self.renderer.remove_map(name)
                del self._uploaded[name]
        for name, m in self.map_batch.items():
            a = self.attention_map_batch.get(name)
            key = (id(m), id(a))
            if self._uploaded.get(name) != key:
                self.renderer.add_map(name, m, a)
                self._uploaded[name] = key

    def _gather_poses(self, corners, t):
        gps = np.empty((self.batch_size, 4, 2), dtype=np.float64)
        geo = np.empty((self.batch_size, 5), dtype=np.float64)
        tidx = np.empty((self.batch_size,), dtype=np.int32)
        for i in range(self.batch_size):
            item = self.batch[i]
            if corners is None:
                n = len(item["gt_path_corners"])
                t_input = 0 if t is None else (t if t < n else n - 1)      # src/env.py:260-266
                gps[i] = np.array(item["gt_path_corners"][t_input], dtype=np.float64)
            else:
                gps[i] = np.array(corners[i], dtype=np.float64)
            geo[i, 0:2] = item["gps_botm_left"]
            geo[i, 2:4] = item["gps_top_right"]
            geo[i, 4] = item["lat_ratio"]
            tidx[i] = self.renderer.map_index(item["map_name"])
        return gps, geo, tidx

    def get_obs_device(self, corners=None, t=None, *, norm_nchw=False, norm_nhwc=False):
        """Device-resident observation: dict of tensors (no host copies).  Keys:
        ``views`` u8, ``att`` u8, ``corners_px`` i32 and the optional fused
        normalised trunk inputs."""
        self._sync_maps()
        gps, geo, tidx = self._gather_poses(corners, t)
        r = self.renderer
        px = r.gps_to_pixels(gps, geo)
        out = r.render(px, tidx, views=True, att=True, norm_nchw=norm_nchw, norm_nhwc=norm_nhwc)
        out["corners_px"] = px
        return out

    def _get_obs(self, corners=None, directions=None, t=None, shortest_teacher=False):
        """Drop-in for src/env.py:254-332: list of per-sample dicts with numpy values."""
        dev = self.get_obs_device(corners, t)
        views = dev["views"].cpu().numpy()
        att = dev["att"].cpu().numpy()
        px = dev["corners_px"].cpu().numpy()
        obs = []
        for i in range(self.batch_size):
            item = self.batch[i]
            obs.append({
                "map_name": item["map_name"],
                "map_size": (self._device_maps[item["map_name"]] if item["map_name"] in self._device_maps
                             else self.map_batch[item["map_name"]].shape),
                "route_index": item["route_index"],
                "gps_botm_left": item["gps_botm_left"],
                "gps_top_right": item["gps_top_right"],
                "lng_ratio": item["lng_ratio"],
                "lat_ratio": item["lat_ratio"],
                "starting_angle": item["angle"],
                "current_view": views[i],
                "gt_saliency": att[i].astype(np.float64) / 255,       # src/env.py:293
                "gt_path_corners": item["gt_path_corners"],
                # the reference returns the corners in PIXEL coordinates (aliasing
                # at src/env.py:280-283), as float64
                "view_area_corners": px[i].astype(np.float64),
                "instructions": item["instructions"],
                "pre_dialogs": item["pre_dialogs"],
            })
        return obs
