"""TEST INFRASTRUCTURE — fp32 restatement of the TESTR text-spotting head as run inside every denoising step.

Evaluated functionally from the reference ``TransformerDetector.state_dict()`` (keys ``testr.*``):
  TESTR.forward                      testr/adet/modeling/testr/models.py:117-171 (diff_feat_proj :76-88)
  PositionalEncoding2D / 1D          testr/adet/layers/pos_encoding.py:46-82, :5-43
  DeformableTransformer.forward      testr/adet/layers/deformable_transformer.py:123-181
  encoder layer / reference points   :214-245 ; proposals :81-112 ; proposal pos embed :66-79
  composite decoder layer            :428-519 ; decoder loop :532-566
  TransformerDetector.inference      testr/adet/modeling/transformer_detector.py:123-152
Masks are all-False on this path (models.py:127), so valid ratios are 1 and padding fills are no-ops.
Parameter aliasing of the reference (ctrl_point_class[0..5] is one module, bbox_class == transformer.bbox_class_embed,
bbox_coord == transformer.bbox_embed; models.py:99-106) is resolved by ``canonical`` exactly as ``load_state_dict``
resolves it: the last key in state_dict order wins.
"""
from __future__ import annotations

import math
from typing import Dict, List, Sequence

import torch
import torch.nn.functional as F

from .msda import msda_core, msda_module

SD = Dict[str, torch.Tensor]
N_HEADS, N_LEVELS, N_POINTS, N_PROPOSALS, N_CTRL, N_CHARS, D_MODEL = 8, 4, 4, 100, 16, 25, 256


def canonical(sd: SD) -> SD:
    """Resolve shared parameters the way nn.Module.load_state_dict does (later keys overwrite earlier aliases)."""
    sd = dict(sd)
    for suffix in ("weight", "bias"):
        win = sd[f"testr.ctrl_point_class.5.{suffix}"]
        for i in range(6):
            sd[f"testr.ctrl_point_class.{i}.{suffix}"] = win
        sd[f"testr.transformer.bbox_class_embed.{suffix}"] = sd[f"testr.bbox_class.{suffix}"]
        for l in range(3):
            win = sd[f"testr.ctrl_point_coord.5.layers.{l}.{suffix}"]
            for i in range(6):
                sd[f"testr.ctrl_point_coord.{i}.layers.{l}.{suffix}"] = win
            sd[f"testr.transformer.bbox_embed.layers.{l}.{suffix}"] = sd[f"testr.bbox_coord.layers.{l}.{suffix}"]
    return sd


def _lin(sd, p, x):
    return F.linear(x, sd[p + ".weight"], sd[p + ".bias"])


def _ln(sd, p, x):
    return F.layer_norm(x, (x.shape[-1],), sd[p + ".weight"], sd[p + ".bias"], 1e-5)


def _mlp(sd, p, x, n=3):
    for i in range(n):
        x = _lin(sd, f"{p}.layers.{i}", x)
        if i < n - 1:
            x = F.relu(x)
    return x


def pos2d(B, H, W, device, num_pos_feats=128, temperature=10000.0):
    """pos_encoding.py:62-82 with normalize=True, scale 2*pi, all-valid mask -> (B, 256, H, W)."""
    scale, eps = 2 * math.pi, 1e-6
    y = torch.arange(1, H + 1, dtype=torch.float32, device=device).view(1, H, 1).expand(B, H, W)
    x = torch.arange(1, W + 1, dtype=torch.float32, device=device).view(1, 1, W).expand(B, H, W)
    y = (y - 0.5) / (H + eps) * scale
    x = (x - 0.5) / (W + eps) * scale
    dim_t = torch.arange(num_pos_feats, dtype=torch.float32, device=device)
    dim_t = temperature ** (2 * torch.div(dim_t, 2, rounding_mode="trunc") / num_pos_feats)
    px, py = x[..., None] / dim_t, y[..., None] / dim_t
    px = torch.stack((px[..., 0::2].sin(), px[..., 1::2].cos()), dim=4).flatten(3)
    py = torch.stack((py[..., 0::2].sin(), py[..., 1::2].cos()), dim=4).flatten(3)
    return torch.cat((py, px), dim=3).permute(0, 3, 1, 2)


def pos1d(n, channels, device, inv_freq, scale):
    """pos_encoding.py:23-43 with normalize=True: positions 1..n scaled by n + eps."""
    pos = torch.arange(1, n + 1, device=device).float()
    pos = pos / (pos[-1:] + 1e-6) * scale
    s = torch.einsum("i,j->ij", pos, inv_freq)
    return torch.cat((s.sin(), s.cos()), dim=-1)[:, :channels]


def _mha(sd, p, q, k, v):
    """nn.MultiheadAttention forward (batch-first here): [N, L, E] each; no masks, dropout off."""
    E = q.shape[-1]
    w, b = sd[p + ".in_proj_weight"], sd[p + ".in_proj_bias"]
    q = F.linear(q, w[:E], b[:E])
    k = F.linear(k, w[E:2 * E], b[E:2 * E])
    v = F.linear(v, w[2 * E:], b[2 * E:])
    N, L, _ = q.shape
    hd = E // N_HEADS

    def heads(t):
        return t.view(N, -1, N_HEADS, hd).transpose(1, 2)
    a = torch.softmax(heads(q) @ heads(k).transpose(-1, -2) / math.sqrt(hd), dim=-1) @ heads(v)
    return F.linear(a.transpose(1, 2).reshape(N, L, E), sd[p + ".out_proj.weight"], sd[p + ".out_proj.bias"])


def _branch(sd, p, sfx, tgt, qpos, boxes, memory, shapes, core):
    """One branch (location: sfx '' / text: sfx '_text') of the composite decoder layer, :454-483 / :485-513."""
    B, n_obj, n_pt, E = tgt.shape
    qk = tgt + qpos
    t2 = _mha(sd, f"{p}.attn_intra{sfx}", qk.flatten(0, 1), qk.flatten(0, 1), tgt.flatten(0, 1)).reshape(tgt.shape)
    tgt = _ln(sd, f"{p}.norm_intra{sfx}", tgt + t2)
    ti = tgt.transpose(1, 2)                                        # [B, n_pt, n_obj, E]
    t2 = _mha(sd, f"{p}.attn_inter{sfx}", ti.flatten(0, 1), ti.flatten(0, 1), ti.flatten(0, 1)).reshape(ti.shape)
    ti = _ln(sd, f"{p}.norm_inter{sfx}", ti + t2).transpose(1, 2)    # back to [B, n_obj, n_pt, E]
    ref = boxes[:, :, None, :, :].expand(B, n_obj, n_pt, boxes.shape[2], boxes.shape[3])
    t2 = msda_module(sd, f"{p}.attn_cross{sfx}", (ti + qpos).flatten(1, 2), ref.flatten(1, 2), memory, shapes,
                     N_HEADS, N_POINTS, core).reshape(ti.shape)
    tgt = _ln(sd, f"{p}.norm_cross{sfx}", ti + t2)
    t2 = _lin(sd, f"{p}.linear2{sfx}", F.relu(_lin(sd, f"{p}.linear1{sfx}", tgt)))
    return _ln(sd, f"{p}.norm3{sfx}", tgt + t2)


def testr_forward(sd: SD, feats: Sequence[torch.Tensor], core=msda_core, proposal_indices=None) -> Dict[str, torch.Tensor]:
    """feats: 4 decoder feature maps (B,C,H,W) fp32 -> pred_logits (B,100,16,1), pred_ctrl_points (B,100,16,2),
    pred_texts (B,100,25,97) of the LAST decoder layer (what inference consumes), plus encoder proposals."""
    sd = canonical(sd)
    T = "testr.transformer"
    B, dev = feats[0].shape[0], feats[0].device
    srcs, poss, shapes = [], [], []
    for l, f in enumerate(feats):
        p = f"testr.diff_feat_proj.{l}"
        h = F.conv2d(f, sd[p + ".0.weight"], sd[p + ".0.bias"])
        h = F.gelu(F.group_norm(h, 32, sd[p + ".1.weight"], sd[p + ".1.bias"], 1e-5))
        h = F.conv2d(h, sd[p + ".3.weight"], sd[p + ".3.bias"], padding=1)
        h = F.gelu(F.group_norm(h, 32, sd[p + ".4.weight"], sd[p + ".4.bias"], 1e-5))
        H, W = h.shape[-2:]
        shapes.append((H, W))
        srcs.append(h.flatten(2).transpose(1, 2))
        poss.append(pos2d(B, H, W, dev).flatten(2).transpose(1, 2) + sd[T + ".level_embed"][l].view(1, 1, -1))
    src, pos = torch.cat(srcs, 1), torch.cat(poss, 1)

    # encoder reference points: pixel centres, normalised per level, same for every level slot (:233-245)
    refs = []
    for (H, W) in shapes:
        ry, rx = torch.meshgrid(torch.linspace(0.5, H - 0.5, H, device=dev), torch.linspace(0.5, W - 0.5, W, device=dev),
                                indexing="ij")
        refs.append(torch.stack((rx.reshape(-1) / W, ry.reshape(-1) / H), -1))
    enc_ref = torch.cat(refs, 0)[None, :, None, :].expand(B, -1, N_LEVELS, 2)
    mem = src
    for i in range(6):
        p = f"{T}.encoder.layers.{i}"
        mem = _ln(sd, p + ".norm1", mem + msda_module(sd, p + ".self_attn", mem + pos, enc_ref, mem, shapes, N_HEADS,
                                                      N_POINTS, core))
        mem = _ln(sd, p + ".norm2", mem + _lin(sd, p + ".linear2", F.relu(_lin(sd, p + ".linear1", mem))))

    # two-stage proposals (:81-112, :154-167)
    props = []
    for lvl, (H, W) in enumerate(shapes):
        gy, gx = torch.meshgrid(torch.linspace(0, H - 1, H, device=dev), torch.linspace(0, W - 1, W, device=dev),
                                indexing="ij")
        grid = (torch.stack((gx, gy), -1) + 0.5) / torch.tensor([W, H], dtype=torch.float32, device=dev)
        wh = torch.ones_like(grid) * 0.05 * (2.0 ** lvl)
        props.append(torch.cat((grid, wh), -1).view(-1, 4))
    props = torch.cat(props, 0)[None].expand(B, -1, -1)
    valid = ((props > 0.01) & (props < 0.99)).all(-1, keepdim=True)
    props_logit = torch.log(props / (1 - props)).masked_fill(~valid, float("inf"))
    out_mem = _ln(sd, T + ".enc_output_norm", _lin(sd, T + ".enc_output", mem.masked_fill(~valid, 0.0)))
    enc_class = _lin(sd, T + ".bbox_class_embed", out_mem)
    enc_coord = _mlp(sd, T + ".bbox_embed", out_mem) + props_logit
    top = torch.topk(enc_class[..., 0], N_PROPOSALS, dim=1)[1] if proposal_indices is None else proposal_indices
    top_coord = torch.gather(enc_coord, 1, top[..., None].expand(-1, -1, 4))
    boxes = top_coord.sigmoid()                                     # (B,100,4) reference boxes
    # proposal positional embedding (:66-79) -> query_pos
    dim_t = torch.arange(64, dtype=torch.float32, device=dev)
    dim_t = 10000 ** (2 * torch.div(dim_t, 2, rounding_mode="trunc") / 64)
    pe = (top_coord.sigmoid() * (2 * math.pi))[..., None] / dim_t
    pe = torch.stack((pe[..., 0::2].sin(), pe[..., 1::2].cos()), dim=4).flatten(2)
    qpos = _ln(sd, T + ".pos_trans_norm", _lin(sd, T + ".pos_trans", pe))       # (B,100,256)

    tgt = sd["testr.ctrl_point_embed.weight"][None, None].expand(B, N_PROPOSALS, N_CTRL, D_MODEL)
    tgt_text = sd["testr.text_embed.weight"][None, None].expand(B, N_PROPOSALS, N_CHARS, D_MODEL)
    text_pos = pos1d(N_CHARS, D_MODEL, dev, sd["testr.text_pos_embed.inv_freq"], 2 * math.pi)
    qpos_loc = qpos[:, :, None, :].expand(B, N_PROPOSALS, N_CTRL, D_MODEL)
    qpos_text = text_pos[None, None].expand(B, N_PROPOSALS, N_CHARS, D_MODEL)
    ref_in = boxes[:, :, None, :].expand(B, N_PROPOSALS, N_LEVELS, 4)          # valid ratios are 1 (:541-546)
    for i in range(6):
        p = f"{T}.decoder.layers.{i}"
        tgt = _branch(sd, p, "", tgt, qpos_loc, ref_in, mem, shapes, core)
        tgt_text = _branch(sd, p, "_text", tgt_text, qpos_text, ref_in, mem, shapes, core)

    # heads on the last layer (models.py:139-163); USE_POLYGON -> plain sigmoid / inverse_sigmoid (utils/misc.py:128-145)
    b = boxes.clamp(0, 1)
    ref_logit = torch.log(b.clamp(min=1e-5) / (1 - b).clamp(min=1e-5))
    logits = _lin(sd, "testr.ctrl_point_class.5", tgt)
    coords = (_mlp(sd, "testr.ctrl_point_coord.5", tgt) + ref_logit[:, :, None, :2]).sigmoid()
    texts = _lin(sd, "testr.text_class", tgt_text)
    return dict(pred_logits=logits, pred_ctrl_points=coords, pred_texts=texts, enc_logits=enc_class,
                enc_boxes=enc_coord.sigmoid(), boxes=boxes, topk_indices=top)


def inference(out: Dict[str, torch.Tensor], threshold: float = 0.5, image_size=(512, 512)) -> List[Dict[str, torch.Tensor]]:
    """transformer_detector.py:123-152 -> per image dict(scores, pred_classes, rec_scores, polygons, recs)."""
    text = torch.softmax(out["pred_texts"], dim=-1)
    prob = out["pred_logits"].mean(-2).sigmoid()
    scores, labels = prob.max(-1)
    res = []
    for s, lab, pts, tx in zip(scores, labels, out["pred_ctrl_points"], text):
        keep = s >= threshold
        pts = pts[keep].clone()
        pts[..., 0] *= image_size[1]
        pts[..., 1] *= image_size[0]
        res.append(dict(scores=s[keep], pred_classes=lab[keep], rec_scores=tx[keep], polygons=pts.flatten(1),
                        recs=tx[keep].topk(1)[1].squeeze(-1)))
    return res
