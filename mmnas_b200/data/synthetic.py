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


def make_batch(spec, seed=888):
    """One host batch (CPU tensors).  Returns (input_tuple, target)."""
    g = torch.Generator().manual_seed(seed)
    B, N, T = spec.batch, spec.n_regions, spec.n_tokens
    frcn = torch.relu(torch.randn(B, N, spec.feat, generator=g))
    bbox = torch.zeros(B, N, 5)
    rel_img = torch.zeros(B, N, N, 4)
    ques = torch.zeros(B, T, dtype=torch.int64)
    for b in range(B):
        n_obj = int(torch.randint(10, N + 1, (1,), generator=g)) if (spec.ragged and N >= 10) else N
        frcn[b, n_obj:] = 0
        boxes = random_boxes(n_obj, g)
        rel_img[b, :n_obj, :n_obj] = box_geometry(boxes)
        bbox[b, :n_obj, :4] = boxes / torch.tensor([640., 480., 640., 480.])
        bbox[b, :n_obj, 4] = ((boxes[:, 2] - boxes[:, 0]) * (boxes[:, 3] - boxes[:, 1])) / (640. * 480.)
        lo = min(3, T)
        n_tok = int(torch.randint(lo, T + 1, (1,), generator=g)) if spec.ragged else T
        ques[b, :n_tok] = torch.randint(3, spec.vocab, (n_tok,), generator=g)
    rel_ques = torch.zeros(B, T, T, 3)
    levels = torch.tensor([0., .3, .6, .9, 1.])
    ans = torch.zeros(B, spec.n_ans)
    for b in range(B):
        k = int(torch.randint(1, 5, (1,), generator=g))
        idx = torch.randint(0, spec.n_ans, (k,), generator=g)
        ans[b, idx] = levels[torch.randint(1, 5, (k,), generator=g)]
    return (frcn, bbox, rel_img, ques, rel_ques), ans


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
