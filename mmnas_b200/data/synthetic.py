"""Synthetic MMnas batches with the shapes and distributions fixed in SURVEY.md §8(d) (there are no datasets
and no network here), plus the host-side geometry producer the loader runs per sample.

VQA input tuple (full_vqa.py:86): (frcn_feat [B,100,2048], bbox_feat [B,100,5], rel_img [B,100,100,4],
ques_ix [B,14] int64, rel_ques [B,14,14,3]) and target ans [B,3129].
"""
import numpy as np
import torch


class SynthSpec:
    def __init__(self, task='vqa', batch=64, n_regions=100, n_tokens=14, feat=2048, vocab=20000, n_ans=3129,
                 word_embed=300, ragged=True):
        self.task, self.batch, self.n_regions, self.n_tokens = task, batch, n_regions, n_tokens
        self.feat, self.vocab, self.n_ans, self.word_embed, self.ragged = feat, vocab, n_ans, word_embed, ragged


def box_geometry(boxes):
    """4-d log-geometry between every pair of boxes [n,4] (x1,y1,x2,y2) -> [n,n,4]; the per-sample CPU step of the
    reference loader (relation_embedding, load_data_vqa.py:7-33): log|dcx/w| and log|dcy/h| clamped at 1e-3,
    log(w_i/w_j), log(h_i/h_j), with w = x2-x1+1, h = y2-y1+1."""
    x1, y1, x2, y2 = boxes.unbind(1)
    w, h = (x2 - x1) + 1., (y2 - y1) + 1.
    cx, cy = (x1 + x2) * 0.5, (y1 + y2) * 0.5
    dx = torch.log(torch.clamp(torch.abs((cx[:, None] - cx[None, :]) / w[:, None]), min=1e-3))
    dy = torch.log(torch.clamp(torch.abs((cy[:, None] - cy[None, :]) / h[:, None]), min=1e-3))
    dw = torch.log(w[:, None] / w[None, :])
    dh = torch.log(h[:, None] / h[None, :])
    return torch.stack((dx, dy, dw, dh), dim=-1)


def random_boxes(n, gen, img_w=640., img_h=480.):
    w = 8 + 192 * torch.rand(n, generator=gen)
    h = 8 + 152 * torch.rand(n, generator=gen)
    x1 = (img_w - w) * torch.rand(n, generator=gen)
    y1 = (img_h - h) * torch.rand(n, generator=gen)
    return torch.stack((x1, y1, x1 + w, y1 + h), dim=1)


def _box_iou(boxes, gt):
    """IoU of boxes [n,4] with one box gt [4] (the +1 pixel convention of the reference's bbox_overlaps, bbox.pyx)."""
    iw = (torch.minimum(boxes[:, 2], gt[2]) - torch.maximum(boxes[:, 0], gt[0]) + 1).clamp(min=0)
    ih = (torch.minimum(boxes[:, 3], gt[3]) - torch.maximum(boxes[:, 1], gt[1]) + 1).clamp(min=0)
    area = (boxes[:, 2] - boxes[:, 0] + 1) * (boxes[:, 3] - boxes[:, 1] + 1)
    ga = (gt[2] - gt[0] + 1) * (gt[3] - gt[1] + 1)
    inter = iw * ih
    return inter / (area + ga - inter)


def _box_deltas(boxes, gt):
    """Fast-RCNN regression targets of gt w.r.t. every box (bbox_transform.py), normalised like BBOX_NORM does."""
    w, h = boxes[:, 2] - boxes[:, 0] + 1, boxes[:, 3] - boxes[:, 1] + 1
    cx, cy = boxes[:, 0] + 0.5 * w, boxes[:, 1] + 0.5 * h
    gw, gh = gt[2] - gt[0] + 1, gt[3] - gt[1] + 1
    gx, gy = gt[0] + 0.5 * gw, gt[1] + 0.5 * gh
    d = torch.stack(((gx - cx) / w, (gy - cy) / h, torch.log(gw / w), torch.log(gh / h)), dim=1)
    return d / torch.tensor([0.1, 0.1, 0.2, 0.2])


def compact(batch):
    """The compact loader format of SURVEY §8f row 3 for a batch from make_batch: region
    features as bf16 (the bf16 arm's own first step: bit-identical results) and raw boxes [B,N,4] in place of the
    [B,N,N,4] log-geometry, which the net then builds on the device (mmnas_box_geometry).  Cuts the host->device bytes
    of a VQA batch from 63.8 MB to 26.7 MB."""
    (frcn, bbox, rel_img, ques, rel_ques), target = batch
    boxes = getattr(rel_img, '_boxes', None)
    if boxes is None:
        raise ValueError('compact() needs a batch produced by make_batch (raw boxes attached to the geometry tensor)')
    return (frcn.to(torch.bfloat16), bbox, boxes, ques, rel_ques), target


def _regions(B, N, feat, ragged, g):
    frcn = torch.relu(torch.randn(B, N, feat, generator=g))
    bbox = torch.zeros(B, N, 5)
    rel_img = torch.zeros(B, N, N, 4)
    boxes_all = []
    for b in range(B):
        n_obj = int(torch.randint(10, N + 1, (1,), generator=g)) if (ragged and N >= 10) else N
        frcn[b, n_obj:] = 0
        boxes = random_boxes(n_obj, g)
        rel_img[b, :n_obj, :n_obj] = box_geometry(boxes)
        bbox[b, :n_obj, :4] = boxes / torch.tensor([640., 480., 640., 480.])
        bbox[b, :n_obj, 4] = ((boxes[:, 2] - boxes[:, 0]) * (boxes[:, 3] - boxes[:, 1])) / (640. * 480.)
        boxes_all.append(boxes)
    raw = torch.zeros(B, N, 4)
    for b, boxes in enumerate(boxes_all):
        raw[b, :boxes.shape[0]] = boxes
    rel_img._boxes = raw                      # raw boxes ride along for compact()
    return frcn, bbox, rel_img, boxes_all


def _tokens(B, T, vocab, lo, hi, g):
    ques = torch.zeros(B, T, dtype=torch.int64)
    for b in range(B):
        n_tok = int(torch.randint(lo, hi + 1, (1,), generator=g))
        ques[b, :n_tok] = torch.randint(3, vocab, (n_tok,), generator=g)
    return ques


def make_batch(spec, seed=888):
    """One host batch (CPU tensors).  Returns (input_tuple, target).

    vqa: target = ans [B, n_ans] soft scores                                         (load_data_vqa.py:221-275)
    vgd: every region valid, queries of n_tokens (= 15: max_token + 1, the last slot always 0); target =
         (scores [B,N] IoU-proportional over regions with IoU >= 0.5, scores_mask [B,1], transformed_bbox [B,N,4],
         bbox_mask [B,N,1])                                                          (load_data_vgd.py:175-186,241-283)
    itm: the three forwards of a step stacked along the batch: rows [0,B) positive pairs, [B,2B) the same images with
         negative captions, [2B,3B) negative images with the positive captions (train_itm.py:380-389); target = None."""
    g = torch.Generator().manual_seed(seed)
    B, N, T = spec.batch, spec.n_regions, spec.n_tokens
    if spec.task == 'vgd':
        frcn, bbox, rel_img, boxes_all = _regions(B, N, spec.feat, False, g)
        ques = _tokens(B, T, spec.vocab, min(3, T - 1), T - 1, g)
        scores, smask = torch.zeros(B, N), torch.zeros(B, 1)
        tbox, bmask = torch.zeros(B, N, 4), torch.zeros(B, N, 1)
        for b, boxes in enumerate(boxes_all):
            k = int(torch.randint(0, boxes.shape[0], (1,), generator=g))
            gt = boxes[k] + 6 * torch.randn(4, generator=g)          # the referred object: a jittered region box
            gt = torch.stack((torch.minimum(gt[0], gt[2] - 4), torch.minimum(gt[1], gt[3] - 4), gt[2], gt[3]))
            iou = _box_iou(boxes, gt)
            hit = iou >= 0.5
            if bool(hit.any()):
                smask[b] = 1
                sc = torch.where(hit, iou, torch.zeros_like(iou))
                scores[b, :boxes.shape[0]] = sc / (sc.sum() + 1e-8)
                bmask[b, :boxes.shape[0], 0] = hit.float()
            tbox[b, :boxes.shape[0]] = _box_deltas(boxes, gt)
        return (frcn, bbox, rel_img, ques, torch.zeros(B, T, T, 3)), (scores, smask, tbox, bmask)
    if spec.task == 'itm':
        frcn, bbox, rel_img, _ = _regions(2 * B, N, spec.feat, spec.ragged, g)      # positive and negative images
        caps = _tokens(2 * B, T, spec.vocab, min(5, T), min(30, T), g)              # positive and negative captions
        pos, neg = slice(0, B), slice(B, 2 * B)
        cat = lambda t, order: torch.cat([t[o] for o in order], 0)                  # noqa: E731
        img_order, cap_order = (pos, pos, neg), (pos, neg, pos)
        rel = cat(rel_img, img_order)
        rel._boxes = cat(rel_img._boxes, img_order)
        return ((cat(frcn, img_order), cat(bbox, img_order), rel, cat(caps, cap_order),
                 torch.zeros(3 * B, T, T, 3)), torch.zeros(1))
    frcn, bbox, rel_img, _ = _regions(B, N, spec.feat, spec.ragged, g)
    ques = _tokens(B, T, spec.vocab, min(3, T), T, g) if spec.ragged else torch.randint(3, spec.vocab, (B, T), generator=g)
    rel_ques = torch.zeros(B, T, T, 3)
    levels = torch.tensor([0., .3, .6, .9, 1.])
    ans = torch.zeros(B, spec.n_ans)
    for b in range(B):
        k = int(torch.randint(1, 5, (1,), generator=g))
        idx = torch.randint(0, spec.n_ans, (k,), generator=g)
        ans[b, idx] = levels[torch.randint(1, 5, (k,), generator=g)]
    return (frcn, bbox, rel_img, ques, rel_ques), ans


def spec_for(task, batch=64, **over):
    """The BASELINE shapes of each task (T/S 100 regions x 14 tokens; G 100 x 15; I 36 x 50)."""
    if task == 'vgd':
        return SynthSpec(task='vgd', batch=batch, n_regions=100, n_tokens=15, ragged=False, **over)
    if task == 'itm':
        return SynthSpec(task='itm', batch=batch, n_regions=36, n_tokens=50, **over)
    return SynthSpec(task='vqa', batch=batch, **over)


def init_dict(spec, seed=888):
    rng = np.random.RandomState(seed)
    return {'token_size': spec.vocab, 'ans_size': spec.n_ans,
            'pretrained_emb': (0.1 * rng.randn(spec.vocab, spec.word_embed)).astype(np.float32)}


class Cfg:
    """Attribute bag with the hyper-parameters the model package reads (train_vqa.py:130-160 /
    search_vqa.py:90-161 defaults)."""

    def __init__(self, mode='train', genotype=None, **over):
        search = mode == 'search'
        self.LAYERS = 1
        self.HSIZE = 256 if search else 512
        self.DROPOUT_R = 0.1
        self.OPS_RESIDUAL = True
        self.OPS_NORM = True
        self.REL_SIZE = 64
        self.BBOX_FEATURE = False
        self.FRCNFEAT_LEN = 100
        self.FRCNFEAT_SIZE = 2048
        self.BBOXFEAT_EMB_SIZE = 2048
        self.WORD_EMBED_SIZE = 300
        self.ATTFLAT_GLIMPSES = 1
        self.ATTFLAT_MLP_SIZE = 512
        self.NODES = {'enc': 12, 'dec': 18}
        self.ALPHA_INIT_TYPE = 'normal'
        self.NET_LR_BASE = 0.0004 if search else 0.00012
        self.NET_GRAD_CLIP = 1.
        self.OPT_BETAS = (0.9, 0.98)
        self.OPT_EPS = 1e-9
        self.ALPHA_LR_BASE = 0.1
        self.ALPHA_OPT_BETAS = (0., 0.999)
        self.ALPHA_EVERY = 5
        self.ALPHA_BINARY_MODE = 'full'
        self.GENOTYPE = genotype
        self.__dict__.update(over)
        self.ATTFLAT_OUT_SIZE = over.get('ATTFLAT_OUT_SIZE', self.HSIZE * 2)
