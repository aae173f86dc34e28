"""TransformerDetector — the text-spotting entry point called once per denoising step
(testr/adet/modeling/transformer_detector.py:40-152).  ``forward(extracted_feats, targets, MODE)`` keeps the
reference contract and returns ``(loss_dict | None, list[Instances])``; only MODE == 'VAL' is on the inference path
(the training criterion / matcher are out of scope)."""
from __future__ import annotations

from types import SimpleNamespace
from typing import List, Optional, Sequence

import numpy as np
import torch
from torch import nn

from .. import ops
from .models import TESTR
from .structures import Instances


def default_cfg(device: str = "cuda") -> SimpleNamespace:
    """The cfg keys TESTR reads: testr/adet/config/defaults.py:340-369 + testr/configs/TESTR/TESTR_R_50_Polygon.yaml."""
    tr = SimpleNamespace(ENABLED=True, INFERENCE_TH_TEST=0.5, VOC_SIZE=96, NUM_CHARS=25, AUX_LOSS=True, ENC_LAYERS=6,
                         DEC_LAYERS=6, DIM_FEEDFORWARD=1024, HIDDEN_DIM=256, DROPOUT=0.1, NHEADS=8, NUM_QUERIES=100,
                         ENC_N_POINTS=4, DEC_N_POINTS=4, POSITION_EMBEDDING_SCALE=6.283185307179586,
                         NUM_FEATURE_LEVELS=4, USE_POLYGON=True, NUM_CTRL_POINTS=16)
    return SimpleNamespace(MODEL=SimpleNamespace(DEVICE=device, TRANSFORMER=tr))


class TransformerDetector(nn.Module):
    def __init__(self, cfg=None):
        super().__init__()
        cfg = cfg or default_cfg()
        self.device = torch.device(cfg.MODEL.DEVICE)
        self.test_score_threshold = cfg.MODEL.TRANSFORMER.INFERENCE_TH_TEST
        self.use_polygon = cfg.MODEL.TRANSFORMER.USE_POLYGON
        self.testr = TESTR(cfg)

    @torch.no_grad()
    def forward(self, extracted_feats: Sequence[torch.Tensor], targets=None, MODE: str = "VAL"):
        if MODE != "VAL":
            raise NotImplementedError("tair_b200 covers the inference path (MODE='VAL'); training losses are out of scope")
        output = self.testr(extracted_feats)
        bs = output["pred_logits"].shape[0]
        image_sizes = [(512, 512) for _ in range(bs)]
        results = self.inference(output["pred_logits"], output["pred_ctrl_points"], output["pred_texts"], image_sizes)
        return None, results

    def detect_host(self, dense, image_size=(512, 512)):
        """The sampler's view of ``inference`` (spaced_sampler.py:294-306 only reads ``polygons`` and ``recs``): one
        post-processing kernel over all tiles x queries, ONE device->host copy, thresholding and string decode on the host.
        -> (texts [B][k] str, polys [B][k] int32 (16,2)); same detections, order and values as ``inference`` + ``decode``."""
        from ..prompt import texts_and_polys
        scores, polys, recs, pack = ops.testr_postprocess(dense["pred_logits"].contiguous(), dense["pred_ctrl_points"].contiguous(),
                                                          dense["pred_texts"].contiguous(), float(image_size[1]), float(image_size[0]))
        B, Q = scores.shape
        n_s, n_p = scores.numel() * 4, polys.numel() * 4
        host = pack.cpu().numpy()
        return texts_and_polys(host[:n_s].view(np.float32).reshape(B, Q),
                               host[n_s:n_s + n_p].view(np.float32).reshape(B, Q, -1),
                               host[n_s + n_p:].reshape(B, Q, -1), self.test_score_threshold)

    def inference(self, ctrl_point_cls, ctrl_point_coord, text_pred, image_sizes) -> List[Instances]:
        """transformer_detector.py:123-152: softmax over the vocabulary, score = sigmoid(mean point logit),
        threshold, scale control points to pixels, arg-max characters."""
        assert len(ctrl_point_cls) == len(image_sizes)
        # Batched: ONE device->host copy (the B x Q keep mask) instead of a boolean-mask sync per image and field; every
        # field is gathered once for the whole batch and split into per-image views on the host-known counts.
        B, Q = ctrl_point_cls.shape[:2]
        dev = ctrl_point_cls.device
        text_pred = torch.softmax(text_pred, dim=-1)
        prob = ctrl_point_cls.mean(-2).sigmoid()
        scores, labels = prob.max(-1)
        recs = text_pred.topk(1)[1].squeeze(-1)
        scale = torch.tensor([[float(sz[1]), float(sz[0])] for sz in image_sizes], dtype=ctrl_point_coord.dtype).to(dev, non_blocking=True)
        pts_px = ctrl_point_coord * scale[:, None, None, :]
        keep = (scores >= self.test_score_threshold).cpu().numpy()
        counts = keep.sum(1).tolist()
        flat = torch.from_numpy(keep.reshape(-1).nonzero()[0]).to(dev, non_blocking=True)
        g_scores = scores.reshape(B * Q).index_select(0, flat).split(counts)
        g_labels = labels.reshape(B * Q).index_select(0, flat).split(counts)
        g_pts = pts_px.reshape(B * Q, -1).index_select(0, flat).split(counts)
        g_txt = text_pred.reshape(B * Q, *text_pred.shape[2:]).index_select(0, flat).split(counts)
        g_recs = recs.reshape(B * Q, -1).index_select(0, flat).split(counts)
        results = []
        for i, size in enumerate(image_sizes):
            r = Instances(size)
            r.scores = g_scores[i]
            r.pred_classes = g_labels[i]
            r.rec_scores = g_txt[i]
            if self.use_polygon:
                r.polygons = g_pts[i]
            else:
                r.beziers = g_pts[i]
            r.recs = g_recs[i]
            results.append(r)
        return results
