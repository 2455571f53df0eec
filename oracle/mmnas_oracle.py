"""CPU oracle for the MMnas operator hot path.  TEST INFRASTRUCTURE ONLY.

This file is the checker, never the product: only ``tests/``, ``__graft_entry__.smoke()``
and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.
Nothing under ``mmnas_b200/`` imports it, and the product path raises when the CUDA
library is missing instead of falling back to this code.

It restates, as stateless functions over a ``{state_dict key: tensor}`` mapping, the
arithmetic of the reference's operator path (all citations relative to
``/root/reference``):

  mmnas/model/modules.py   LayerNorm :44-56, FC/MLP :13-41, AttFlat :59-85,
                           MHAtt :158-199, RelMHAtt :202-245, SelfAtt :248-271,
                           RelSelfAtt :274-298, GuidedAtt :301-325, FeedForward :328-362
  mmnas/model/mixed.py     MixedOp.forward :59-106, set_arch_param_grad :171-198
  mmnas/model/full_vqa.py  Cell_Full :24-28, Backbone_Full :46-53, Net_Full.forward :85-114
  mmnas/model/hygr_vqa.py  Cell_Search :23-27, Net_Search.forward :112-141
  mmnas/loader/load_data_vqa.py  relation_embedding :7-33

The arithmetic of that path lives in PyTorch (un-vendored third-party dependency;
installed wheel torch 2.11.0+cu128, the reference pins only "PyTorch >= 0.4.1").  The
reference ships no tests, golden vectors or fixtures for this path, so parity is pinned
the second way the task allows: ``tests/golden/make_golden.py`` imports the *unmodified*
reference modules from /root/reference in the authoring container, runs them on seeded
inputs and commits the inputs/outputs/gradients under ``tests/golden/``;
``tests/test_oracle_golden.py`` checks every function here against those vectors (and
against the live reference when /root/reference is mounted).

Gradients come from torch autograd over these functions; run them in float64 to get a
reference that is tighter than the fp32 reference itself.
"""
import math

import torch
import torch.nn.functional as F

HEAD_DIM = 64  # the '*_64' operator names fix the per-head width (ops_adapter.py:33,40,46)


# --------------------------------------------------------------------------------------
# elementary pieces
# --------------------------------------------------------------------------------------
def layer_norm(z, a_2, b_2, eps=1e-6):
    """modules.py:52-56 — unbiased std, eps added to sigma (not to the variance)."""
    mu = z.mean(-1, keepdim=True)
    sigma = z.std(-1, keepdim=True)
    return a_2 * (z - mu) / (sigma + eps) + b_2


def drop(t, p, training):
    return F.dropout(t, p, training) if (training and p > 0) else t


def split_heads(t, n_batch, heads, base):
    return t.view(n_batch, -1, heads, base).transpose(1, 2)


def mh_att(P, pre, v_in, k_in, q_in, mask, rel_embed=None, base=HEAD_DIM, p=0.0, training=False):
    """MHAtt.forward modules.py:178-199 / RelMHAtt.forward :224-245 (when rel_embed given).

    ``pre`` is the state-dict prefix of the ``mhatt`` sub-module (ending in '.').
    ``mask``: bool [B,1,1,Nk], True = padded key.
    """
    Wv, Wk, Wq, Wm = (P[pre + 'linear_%s.weight' % n] for n in ('v', 'k', 'q', 'merge'))
    inside = Wq.shape[0]
    heads = inside // base
    nb = q_in.size(0)
    v = split_heads(F.linear(v_in, Wv), nb, heads, base)
    k = split_heads(F.linear(k_in, Wk), nb, heads, base)
    q = split_heads(F.linear(q_in, Wq), nb, heads, base)
    scores = torch.matmul(q, k.transpose(-2, -1)) / math.sqrt(base)
    if rel_embed is not None:
        r = F.relu(F.linear(rel_embed, P[pre + 'linear_r.weight'], P[pre + 'linear_r.bias']))
        scores = torch.log(torch.clamp(r.permute(0, 3, 1, 2), min=1e-6)) + scores
    if mask is not None:
        scores = scores.masked_fill(mask, -1e9)
    att = drop(F.softmax(scores, dim=-1), p, training)
    out = torch.matmul(att, v).transpose(1, 2).contiguous().view(nb, -1, inside)
    return F.linear(out, Wm)


def mlp(P, pre, x, p=0.0, training=False):
    """MLP.forward modules.py:40-41 over FC.forward :24-31."""
    h = F.relu(F.linear(x, P[pre + 'fc.linear.weight'], P[pre + 'fc.linear.bias']))
    h = drop(h, p, training)
    return F.linear(h, P[pre + 'linear.weight'], P[pre + 'linear.bias'])


def block_tail(P, pre, x, branch, norm, residual, p=0.0, training=False):
    """modules.py:261-271 — residual add of the (dropped-out) branch, then LayerNorm."""
    branch = drop(branch, p, training)
    z = x + branch if residual else branch
    if norm:
        z = layer_norm(z, P[pre + 'ln.a_2'], P[pre + 'ln.b_2'])
    return z


def op_forward(name, P, pre, x, y=None, x_mask=None, y_mask=None, rel_embed=None,
               norm=True, residual=True, p=0.0, training=False):
    """One candidate block, addressed by its registry name (ops_adapter.py:24-74)."""
    if name == 'self_att_64':           # SelfAtt.forward modules.py:260
        br = mh_att(P, pre + 'mhatt.', x, x, x, x_mask, None, HEAD_DIM, p, training)
    elif name == 'rel_self_att_64':     # RelSelfAtt.forward modules.py:286
        assert rel_embed is not None
        br = mh_att(P, pre + 'mhatt.', x, x, x, x_mask, rel_embed, HEAD_DIM, p, training)
    elif name == 'guided_att_64':       # GuidedAtt.forward modules.py:313
        assert y is not None
        br = mh_att(P, pre + 'mhatt.', y, y, x, y_mask, None, HEAD_DIM, p, training)
    elif name == 'feed_forward':        # FeedForward.forward modules.py:351
        br = mlp(P, pre + 'mlp.', x, p, training)
    else:
        raise KeyError(name)
    return block_tail(P, pre, x, br, norm, residual, p, training)


# --------------------------------------------------------------------------------------
# supernet mixed-op
# --------------------------------------------------------------------------------------
def mixed_forward(used_ops, P, pre, s, pre_s, s_mask, pre_mask, rel_embed, mode, active, inactive,
                  norm=True, residual=True, p=0.0, training=False):
    """MixedOp.forward mixed.py:59-106.  ``pre`` = prefix of the MixedOp module."""
    def cand(i):
        return op_forward(used_ops[i], P, '%scandidate_ops.%d.' % (pre, i), s, pre_s, s_mask, pre_mask,
                          rel_embed, norm, residual, p, training)
    if mode in ('full', 'two'):
        gate = P[pre + 'alpha_gate']
        out = 0
        for i in active:
            out = out + gate[i] * cand(i)
        for i in inactive:
            out = out + gate[i] * cand(i).detach()
        return out
    return cand(active[0])


def arch_param_grad(alpha_prob, gate_grad):
    """MixedOp.set_arch_param_grad mixed.py:193-197 ('full' mode), double loop kept literal:
    grad_i += sum_j gate_grad_j * p_j * (delta_ij - p_i),  p = softmax(alpha_prob)."""
    probs = F.softmax(alpha_prob.detach(), dim=0)
    n = probs.numel()
    out = torch.zeros_like(probs)
    for i in range(n):
        for j in range(n):
            out[i] += gate_grad[j] * probs[j] * ((1 if i == j else 0) - probs[i])
    return out


# --------------------------------------------------------------------------------------
# geometry producer
# --------------------------------------------------------------------------------------
def relation_embedding(f_g):
    """load_data_vqa.py:7-33.  boxes [n,4] (x1,y1,x2,y2) -> [n,n,4] log-geometry."""
    x_min, y_min, x_max, y_max = torch.chunk(f_g, 4, dim=1)
    cx = (x_min + x_max) * 0.5
    cy = (y_min + y_max) * 0.5
    w = (x_max - x_min) + 1.
    h = (y_max - y_min) + 1.
    dx = torch.log(torch.clamp(torch.abs((cx - cx.view(1, -1)) / w), min=1e-3))
    dy = torch.log(torch.clamp(torch.abs((cy - cy.view(1, -1)) / h), min=1e-3))
    dw = torch.log(w / w.view(1, -1))
    dh = torch.log(h / h.view(1, -1))
    return torch.stack((dx, dy, dw, dh), -1)


# --------------------------------------------------------------------------------------
# callers: cells, backbone, VQA nets
# --------------------------------------------------------------------------------------
def make_mask(feature):
    """full_vqa.py:113-114."""
    return (feature.abs().sum(-1) == 0).unsqueeze(1).unsqueeze(2)


def lstm(P, pre, x):
    """nn.LSTM(num_layers=1, batch_first=True) zero initial state (full_vqa.py:63-68,95),
    written out as the recurrence torch documents: gates i,f,g,o."""
    W_ih, W_hh = P[pre + 'weight_ih_l0'], P[pre + 'weight_hh_l0']
    b_ih, b_hh = P[pre + 'bias_ih_l0'], P[pre + 'bias_hh_l0']
    nb, T, _ = x.shape
    hid = W_hh.shape[1]
    h = x.new_zeros(nb, hid)
    c = x.new_zeros(nb, hid)
    outs = []
    for t in range(T):
        g = F.linear(x[:, t], W_ih, b_ih) + F.linear(h, W_hh, b_hh)
        i, f, gg, o = g.chunk(4, dim=1)
        c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(gg)
        h = torch.sigmoid(o) * torch.tanh(c)
        outs.append(h)
    return torch.stack(outs, 1)


def att_flat(P, pre, x, x_mask, p=0.0, training=False):
    """AttFlat.forward modules.py:74-85 (ATTFLAT_GLIMPSES glimpses)."""
    att = mlp(P, pre + 'mlp.', x, p, training)
    if x_mask is not None:
        att = att.masked_fill(x_mask.squeeze(1).squeeze(1).unsqueeze(2), -1e9)
    att = F.softmax(att, dim=1)
    g = att.shape[-1]
    flat = torch.cat([torch.sum(att[:, :, i:i + 1] * x, dim=1) for i in range(g)], dim=1)
    return F.linear(flat, P[pre + 'linear_merge.weight'], P[pre + 'linear_merge.bias'])


def stem_vqa(P, inputs, search=False):
    """Net_Full.forward full_vqa.py:86-103 / Net_Search.forward hygr_vqa.py:113-135 up to the backbone."""
    frcn_feat, bbox_feat, y_rel, ques_ix, x_rel = inputs
    x_mask = make_mask(ques_ix.unsqueeze(2))
    y_mask = make_mask(frcn_feat)
    x_in = lstm(P, 'lstm.', F.embedding(ques_ix, P['embedding.weight']))
    y_in = F.linear(frcn_feat, P['imgfeat_linear.weight'], P['imgfeat_linear.bias'])
    if search:
        x_rel = F.relu(F.linear(x_rel, P['linear_x_rel.weight'], P['linear_x_rel.bias']))
    y_rel = F.relu(F.linear(y_rel, P['linear_y_rel.weight'], P['linear_y_rel.bias']))
    return x_in, y_in, x_mask, y_mask, x_rel, y_rel


def head_vqa(P, x, y, x_mask, y_mask, p=0.0, training=False):
    """full_vqa.py:105-109."""
    xy = att_flat(P, 'attflat_x.', x, x_mask, p, training) + att_flat(P, 'attflat_y.', y, y_mask, p, training)
    xy = layer_norm(xy, P['proj_norm.a_2'], P['proj_norm.b_2'])
    return F.linear(xy, P['proj.weight'], P['proj.bias'])


def head_vgd(P, x, y, x_mask, scores_loss='kld', p=0.0, training=False):
    """full_vgd.py:105-112: per-region scores (log-softmax over regions for the KLD loss) and box regression."""
    xy = att_flat(P, 'attflat_x.', x, x_mask, p, training).unsqueeze(1) + F.linear(y, P['attfc_y.weight'], P['attfc_y.bias'])
    xy = layer_norm(xy, P['proj_norm.a_2'], P['proj_norm.b_2'])
    scores = F.linear(xy, P['proj_scores.weight'], P['proj_scores.bias']).squeeze(-1)
    if scores_loss == 'kld':
        scores = F.log_softmax(scores, dim=-1)
    return scores, F.linear(xy, P['proj_reg.weight'], P['proj_reg.bias'])


def head_itm(P, x, y, x_mask, y_mask, p=0.0, training=False):
    """full_itm.py:105-110: one matching logit per (image, caption) pair, through a sigmoid."""
    return torch.sigmoid(head_vqa(P, x, y, x_mask, y_mask, p, training).squeeze(-1))


def net_full(P, inputs, genotype, task='vqa', scores_loss='kld', p=0.0, training=False):
    """Net_Full of full_vqa.py / full_vgd.py / full_itm.py: same stem and backbone, task-specific head."""
    x, y = net_full_vqa(P, inputs, genotype, p, training, return_backbone=True)
    x_mask, y_mask = make_mask(inputs[3].unsqueeze(2)), make_mask(inputs[0])
    if task == 'vgd':
        return head_vgd(P, x, y, x_mask, scores_loss, p, training)
    if task == 'itm':
        return head_itm(P, x, y, x_mask, y_mask, p, training)
    return head_vqa(P, x, y, x_mask, y_mask, p, training)


def net_full_vqa(P, inputs, genotype, p=0.0, training=False, return_backbone=False):
    """Net_Full (full_vqa.py:56-114) on the genotype {'enc': [[op]...], 'dec': [[op]...]}."""
    x, y, x_mask, y_mask, x_rel, y_rel = stem_vqa(P, inputs)
    for i, node in enumerate(genotype['enc']):      # Cell_Full.forward: s = sum(op(...) for op in ops)
        x = sum(op_forward(n, P, 'backnone.cells_enc.0.dag.%d.%d.' % (i, j), x, None, x_mask, None, x_rel,
                           True, True, p, training) for j, n in enumerate(node))
    for i, node in enumerate(genotype['dec']):
        y = sum(op_forward(n, P, 'backnone.cells_dec.0.dag.%d.%d.' % (i, j), y, x, y_mask, x_mask, y_rel,
                           True, True, p, training) for j, n in enumerate(node))
    if return_backbone:
        return x, y
    return head_vqa(P, x, y, x_mask, y_mask, p, training)


ENC_SAFE = ['self_att_64', 'feed_forward']                                       # ops_adapter.py:8-12
DEC_SAFE = ['self_att_64', 'rel_self_att_64', 'guided_att_64', 'feed_forward']   # ops_adapter.py:13-19


def net_search_vqa(P, inputs, mode, choices, n_enc=12, n_dec=18, p=0.0, training=False):
    """Net_Search (hygr_vqa.py:55-141).  ``choices`` = {'enc': [active_idx]*n_enc, 'dec': [...]}:
    the indices MixedOp.binarize (mixed.py:151-156) would have sampled."""
    x, y, x_mask, y_mask, x_rel, y_rel = stem_vqa(P, inputs, search=True)

    def node(kind, used, i, s, pre_s, s_mask, pre_mask, rel):
        a = choices[kind][i]
        inactive = [k for k in range(len(used)) if k != a]
        return 0 + mixed_forward(used, P, 'backnone.cells_%s.0.dag.%d.0.' % (kind, i), s, pre_s, s_mask, pre_mask,
                                 rel, mode, [a], inactive, True, True, p, training)
    for i in range(n_enc):
        x = node('enc', ENC_SAFE, i, x, None, x_mask, None, x_rel)
    for i in range(n_dec):
        y = node('dec', DEC_SAFE, i, y, x, y_mask, x_mask, y_rel)
    return head_vqa(P, x, y, x_mask, y_mask, p, training)


# --------------------------------------------------------------------------------------
# step bodies (what bench.py's reference arm times, and what the step tests compare)
# --------------------------------------------------------------------------------------
def leaf_params(state, dtype=torch.float32, requires_grad=True):
    """Detached copies of a state dict as autograd leaves."""
    out = {}
    for k, v in state.items():
        t = v.detach().clone()
        if t.is_floating_point():
            t = t.to(dtype).requires_grad_(requires_grad)
        out[k] = t
    return out


def train_step_vqa(P, inputs, target, genotype, p=0.0, training=False):
    """Loss + grads of the train step body train_vqa.py:294-299 (BCE-with-logits, reduction='sum')."""
    pred = net_full_vqa(P, inputs, genotype, p, training)
    loss = F.binary_cross_entropy_with_logits(pred, target, reduction='sum')
    loss.backward()
    return loss.detach(), pred.detach()


def vgd_loss(pred_scores, pred_reg, scores, scores_mask, transformed_bbox, bbox_mask, loss_lambda=0.5):
    """train_vgd.py:320-334 with the shipped configuration (SCORES_LOSS 'kld' :159, LOSS_AVG True :160, LOSS_LAMBDA
    0.5 :161, REDUCTION 'sum' :178): KLDivLoss(sum) over masked log-probabilities :323, SmoothL1Loss(sum) over masked
    box deltas :324, each divided by its mask count :327-333."""
    loss_scores = F.kl_div(pred_scores * scores_mask, scores * scores_mask, reduction='sum')
    loss_reg = F.smooth_l1_loss(pred_reg * bbox_mask, transformed_bbox * bbox_mask, reduction='sum')
    return loss_scores / scores_mask.sum() + loss_lambda * (loss_reg / bbox_mask.sum())


def itm_bce_loss(scores_pos, scores_negc, scores_negi):
    """mmnas/utils/itm_loss.py:13-24 (BCE_Loss, REDUCTION 'sum' train_itm.py:171): loss_pos + loss_negc + loss_pos +
    loss_negi — the positive term counted twice (:22), kept literally."""
    loss_pos = F.binary_cross_entropy(scores_pos, torch.ones_like(scores_pos), reduction='sum')
    loss_negc = F.binary_cross_entropy(scores_negc, torch.zeros_like(scores_negc), reduction='sum')
    loss_negi = F.binary_cross_entropy(scores_negi, torch.zeros_like(scores_negi), reduction='sum')
    return loss_pos + loss_negc + loss_pos + loss_negi


def train_step_vgd(P, inputs, target, genotype, p=0.0, training=False):
    """Loss + grads of the VGD step body train_vgd.py:317-335."""
    pred_scores, pred_reg = net_full(P, inputs, genotype, task='vgd', scores_loss='kld', p=p, training=training)
    loss = vgd_loss(pred_scores, pred_reg, *target)
    loss.backward()
    return loss.detach(), (pred_scores.detach(), pred_reg.detach())


def train_step_itm(P, input_pos, input_negc, input_negi, genotype, p=0.0, training=False):
    """Loss + grads of the ITM step body train_itm.py:384-392: THREE forwards of the same net, then BCE_Loss."""
    s_pos = net_full(P, input_pos, genotype, task='itm', p=p, training=training)
    s_negc = net_full(P, input_negc, genotype, task='itm', p=p, training=training)
    s_negi = net_full(P, input_negi, genotype, task='itm', p=p, training=training)
    loss = itm_bce_loss(s_pos, s_negc, s_negi)
    loss.backward()
    return loss.detach(), (s_pos.detach(), s_negc.detach(), s_negi.detach())


def binarize_two(alpha_prob, generator=None):
    """MixedOp.binarize in MODE 'two' (mixed.py:136-148): draw two candidates without replacement from
    softmax(alpha_prob), then pick the active one of the pair from the softmax over just those two alphas."""
    probs = F.softmax(alpha_prob.detach(), dim=0)
    pair = torch.multinomial(probs, 2, replacement=False, generator=generator)
    sub = F.softmax(torch.stack([alpha_prob.detach()[i] for i in pair]), dim=0)
    c = torch.multinomial(sub, 1, generator=generator)[0]
    return pair[c].item(), pair[1 - c].item()


def arch_param_grad_two(alpha_prob, gate_grad, active, inactive):
    """MixedOp.set_arch_param_grad in MODE 'two' (mixed.py:179-186): the 2 x 2 rule over the involved pair only."""
    involved = [active, inactive]
    probs = F.softmax(torch.stack([alpha_prob.detach()[i] for i in involved]), dim=0)
    out = torch.zeros_like(alpha_prob.detach())
    for i in range(2):
        for j in range(2):
            out[involved[i]] += gate_grad[involved[j]] * probs[j] * ((1 if i == j else 0) - probs[i])
    return out


def rescale_two(alpha_new, alpha_old, active, inactive):
    """MixedOp.rescale_updated_arch_param (mixed.py:200-208): shift the two touched alphas so that their logsumexp is
    what it was before the optimizer step."""
    involved = [active, inactive]
    offset = math.log(sum(math.exp(float(alpha_new[i])) for i in involved) /
                      sum(math.exp(float(alpha_old[i])) for i in involved))
    out = alpha_new.detach().clone()
    for i in involved:
        out[i] -= offset
    return out


def clip_and_adam(params, state, step, lr, max_norm=1.0, betas=(0.9, 0.98), eps=1e-9):
    """clip_grad_norm_ (train_vqa.py:310) then Adam (train_vqa.py:311 via optimizer.py:14-20), in place."""
    grads = [p.grad for p in params if p.grad is not None]
    total = torch.sqrt(sum((g.double() ** 2).sum() for g in grads)).float()
    coef = torch.clamp(max_norm / (total + 1e-6), max=1.0)
    b1, b2 = betas
    with torch.no_grad():
        for p in params:
            if p.grad is None:
                continue
            g = p.grad * coef
            m, v = state.setdefault(id(p), (torch.zeros_like(p), torch.zeros_like(p)))
            m.mul_(b1).add_(g, alpha=1 - b1)
            v.mul_(b2).addcmul_(g, g, value=1 - b2)
            denom = (v.sqrt() / math.sqrt(1 - b2 ** step)).add_(eps)
            p.addcdiv_(m, denom, value=-lr / (1 - b1 ** step))
    return total
